"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md):
    python tools/sass_summary.py hnsw_clj_b200/build/hb_tc.o > profiles/rNN_sass_tc_pass_summary.txt"""
import collections
import re
import subprocess
import sys

obj = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
fn, cnt = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        cnt[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn and re.match(r"UTC|LDTM|STTM|UBLKCP|UTMA|SYNCS|HMMA|IMMA|DFMA|DADD|DMUL", m.group(1)):
        cnt[fn][m.group(1)] += 1
print(f"# cuobjdump -sass {obj}: per kernel, tcgen05 (UTC*MMA), TMEM (LDTM / STTM), bulk copies (UBLKCP / UTMA*), mbarrier (SYNCS), fp64 (D*)")
for fn, c in cnt.items():
    dem = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print(dem)
    for op, n in sorted(c.items()):
        print(f"    {n:5d}  {op}")
