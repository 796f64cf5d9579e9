"""CPU model of warp_sort_regs (hnsw_clj_b200/csrc/hb_fastprep.cu): the bitonic network over 32 lanes x EPL registers
that cand_select_warp_kernel runs.  Position i = lane * EPL + e; strides below EPL swap a lane's own registers, the
others exchange with lane ^ (stride / EPL).  The model applies exactly those rules and must sort every input — the GPU
parity tests then only have to show that the kernel implements the model."""
import numpy as np
import pytest


def warp_sort_model(keys: np.ndarray, epl: int) -> np.ndarray:
    k = keys.reshape(32, epl).copy()  # k[lane][e]
    n = 32 * epl
    size = 2
    while size <= n:
        stride = size >> 1
        while stride > 0:
            if stride < epl:
                for lane in range(32):
                    for e in range(epl):
                        if e & stride:
                            continue
                        asc = ((lane * epl + e) & size) == 0
                        a, b = k[lane][e], k[lane][e | stride]
                        if (a > b) == asc:
                            k[lane][e], k[lane][e | stride] = b, a
            else:
                ld = stride // epl
                new = k.copy()
                for lane in range(32):
                    for e in range(epl):
                        o = k[lane ^ ld][e]  # __shfl_xor_sync(k[e], ld)
                        i = lane * epl + e
                        keep_min = ((i & size) == 0) == ((i & stride) == 0)
                        new[lane][e] = min(o, k[lane][e]) if keep_min else max(o, k[lane][e])
                k = new
            stride >>= 1
        size <<= 1
    return k.reshape(-1)


@pytest.mark.parametrize("epl", [2, 4, 8, 16])
def test_register_bitonic_network_sorts(epl):
    rng = np.random.default_rng(epl)
    n = 32 * epl
    for trial in range(6):
        if trial == 0:
            keys = np.arange(n, dtype=np.uint64)[::-1].copy()
        elif trial == 1:
            keys = np.arange(n, dtype=np.uint64)
        else:
            # unique 64-bit keys as the kernel builds them: (score key << 32) | slot, many equal scores
            score = rng.integers(0, 8 if trial == 2 else 1 << 32, size=n, dtype=np.uint64)
            keys = (score << np.uint64(32)) | np.arange(n, dtype=np.uint64)
            keys = keys[rng.permutation(n)]
        out = warp_sort_model(keys, epl)
        assert np.array_equal(out, np.sort(keys))


def test_padding_keys_sort_last():
    # unused positions hold ~0 (all ones) and must end up behind every real key
    epl, n_real = 8, 100
    keys = np.full(32 * epl, np.iinfo(np.uint64).max, dtype=np.uint64)
    rng = np.random.default_rng(1)
    keys[:n_real] = (rng.integers(0, 1 << 31, size=n_real, dtype=np.uint64) << np.uint64(32)) | np.arange(n_real, dtype=np.uint64)
    out = warp_sort_model(keys[rng.permutation(keys.size)], epl)
    assert np.array_equal(out[:n_real], np.sort(keys[:n_real]))
    assert (out[n_real:] == np.iinfo(np.uint64).max).all()
