"""Host model of the chunked ordered sum of hb_kpp.cu: the reference's k-means++ adds d_i^2 left to right in fp64
(src/hnsw/ann/partition/ivf_flat.clj:51-58); the device replaces the adds of a chunk by one integer add whenever all the
chunk's partial sums provably stay in one binade and no element falls on a rounding tie.  This file states that algorithm in
plain Python and checks it against the sequential loop on adversarial inputs (the CUDA kernels are held to the same inputs
in tests/test_gpu_kpp.py)."""
import math
import random

import numpy as np

KC = 512
DELTA = 2.0 ** -24


def sequential(xs):
    s, out = 0.0, []
    for x in xs:
        s = s + x
        out.append(s)
    return out


def chunked(xs):
    """(total, number of chunks that needed real adds) by the device's rules."""
    n = len(xs)
    nchunks = -(-n // KC)
    approx = [float(np.sum(np.asarray(xs[k * KC:(k + 1) * KC], dtype=np.float64))) for k in range(nchunks)]  # any order
    prefix = [0.0] * nchunks
    run = 0.0
    for k in range(nchunks):
        prefix[k] = run
        run += approx[k]
    cum, walked, starts = 0.0, 0, []
    for k in range(nchunks):
        starts.append(cum)
        chunk = xs[k * KC:(k + 1) * KC]
        lo, hi = prefix[k] * (1 - DELTA), (prefix[k] + approx[k]) * (1 + DELTA)
        e = None
        if lo > 0 and hi < 1.7e308 and lo >= 2.2250738585072014e-308 * 2.0 ** 52:
            elo, ehi = math.frexp(lo)[1] - 1, math.frexp(hi)[1] - 1
            if elo == ehi:
                e = elo
        q, tie = 0, False
        if e is not None:
            for x in chunk:
                if not (0.0 <= x < 1.7e308):
                    tie = True
                    continue
                sc = math.ldexp(x, 52 - e)
                fl = math.floor(sc)
                fr = sc - fl
                if sc >= 2.0 ** 53:
                    tie = True
                q += int(fl) + (1 if fr > 0.5 else 0)
                if fr == 0.5:
                    tie = True
        done = False
        if e is not None and not tie and cum > 0 and math.frexp(cum)[1] - 1 == e:
            mm = int(math.ldexp(cum, 52 - e)) + q
            if mm < 2 ** 53:
                cum = math.ldexp(float(mm), e - 52)
                done = True
        if not done:
            for x in chunk:
                cum = cum + x
            walked += 1
    return cum, walked, starts


def cases():
    r = random.Random(7)
    g = np.random.default_rng(7)
    yield "uniform squares", [(r.random() * 2) ** 2 for _ in range(20000)]
    yield "cosine-distance squares", list((g.random(30000) * 0.3) ** 2)
    yield "many ties", [2.0 ** -30 * r.randrange(1, 8) for _ in range(5000)] + [0.5] + [2.0 ** -54 * (2 * r.randrange(1, 2 ** 20) + 1) for _ in range(20000)]
    yield "half-ulp exactly", [1.0] + [2.0 ** -53] * 3000 + [2.0 ** -52] * 3000 + [3 * 2.0 ** -54] * 3000
    yield "power-of-two crossings", [2.0 ** -20] * (1 << 15) + [1.0, 1.0, 2.0, 4.0] + [2.0 ** -40 * r.randrange(1, 999) for _ in range(9000)]
    yield "huge range", [10.0 ** r.uniform(-200, 100) for _ in range(8000)]
    yield "zeros and tiny", [0.0] * 1500 + [5e-324] * 100 + [1e-310] * 700 + [2.2250738585072014e-308] * 900 + [1e-300 * r.random() for _ in range(3000)]
    yield "decreasing", sorted(((g.random(12000)) ** 4).tolist(), reverse=True)
    yield "one element", [0.3]
    yield "all equal", [0.1] * 10000


def test_chunked_sum_equals_sequential_bits():
    for name, xs in cases():
        want = sequential(xs)
        got, walked, starts = chunked(xs)
        assert got.hex() == want[-1].hex(), name
        for k, s in enumerate(starts):  # the running sum at every chunk start is exact too (the pick walks from there)
            assert s.hex() == (want[k * KC - 1] if k else 0.0).hex(), (name, k)
        assert walked <= max(3, len(xs) // KC // 2 + 70), (name, walked)  # the integer path carries most chunks


def test_large_random_mostly_integer_path():
    g = np.random.default_rng(3)
    xs = ((g.random(400000) * 0.2) ** 2).tolist()
    got, walked, _ = chunked(xs)
    assert got.hex() == sequential(xs)[-1].hex()
    assert walked <= 40  # ~log2(n / 512) binade crossings + the first chunks + the rare tie


def test_adding_ulps_is_an_integer_add_on_the_bit_pattern():
    """kpp_compose_pick_kernel composes a chunk as bits(cum) + Q: for a normal cum in the binade 2^e and 0 <= Q with
    significand(cum) + Q < 2^53 that IS cum + Q * ulp(cum), exactly."""
    import struct

    rng = random.Random(5)
    for _ in range(20000):
        e = rng.randint(-900, 900)
        sig = rng.randint(1 << 52, (1 << 53) - 1)
        cum = math.ldexp(float(sig), e - 52)
        assert math.frexp(cum)[1] - 1 == e
        q = rng.randint(0, (1 << 53) - 1 - sig)
        bits = struct.unpack("<q", struct.pack("<d", cum))[0]
        assert ((bits >> 52) & 0x7FF) - 1023 == e and (bits & ((1 << 52) - 1)) | (1 << 52) == sig
        got = struct.unpack("<d", struct.pack("<q", bits + q))[0]
        want = math.ldexp(float(sig + q), e - 52)  # sig + q < 2^53: exact
        assert got == want
