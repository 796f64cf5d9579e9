"""Measured tcgen05.mma issue rate per operand kind on this GPU (clocks per M128 x N128 x 32-byte-K instruction)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hnsw_clj_b200 import _lib
_lib.check(_lib.lib().hb_init(0))
for kind, kbytes in (("i8", 32), ("bf16", 32), ("e4m3", 32), ("tf32", 32)):
    c = _lib.get_stat("mma_clocks_" + kind)
    elems = {"i8": 32, "bf16": 16, "e4m3": 32, "tf32": 8}[kind]
    print(f"{kind}: {c:.1f} clocks per MMA (M128 N128 K{elems}) -> {128*128*elems/c:.0f} MAC/clk/SM")
