"""The arithmetic behind the packed-rank scan (csrc/acsfit_kernels.cuh, RankLayout), restated in numpy:
req_d <= thr_d for every d  <=>  (((T | G) - R) & G) == G  word by word, with R the pods' 1-based ranks among
the sorted distinct request values, T the count of table values <= thr, one guard bit per field.  The CUDA
implementation itself is checked bit for bit against the oracle in the -m gpu tests (both scan forms)."""
import numpy as np


def layout(U):
    word, used, fields, guard = 0, 0, [], [0, 0, 0, 0]
    for u in U:
        bits = int(u).bit_length()
        if used + bits + 1 > 32:
            word, used = word + 1, 0
        if word >= 4:
            return None
        fields.append((word, used))
        guard[word] |= 1 << (used + bits)
        used += bits + 1
    return fields, guard, word + 1


def test_packed_predicate_equals_float_compares():
    rng = np.random.Generator(np.random.PCG64(5))
    for trial in range(200):
        D = int(rng.integers(1, 17))
        tables = [np.unique(rng.choice([0.0, 0.1, 0.25, 1.0, 2.5, 64.0 * 2 ** 20, 1e9, 3.0, 7.5, np.inf],
                                       size=int(rng.integers(1, 10)))) if rng.random() < 0.7
                  else np.unique(rng.random(int(rng.integers(1, 3000)))) for _ in range(D)]
        lay = layout([len(t) for t in tables])
        if lay is None:
            continue
        fields, guard, nw = lay
        P, N = 64, 64
        req = np.stack([t[rng.integers(0, len(t), size=P)] for t in tables], axis=1)
        # thresholds: -1 (nothing fits), +inf, table values, and values between table values
        thr = np.stack([rng.choice(np.concatenate([[-1.0, np.inf], t, t * 0.999, t * 1.001 + 1e-9]), size=N)
                        for t in tables], axis=1)
        R = np.zeros((P, 4), dtype=np.uint64)
        T = np.zeros((N, 4), dtype=np.uint64)
        for d, (w, sh) in enumerate(fields):
            R[:, w] |= (np.searchsorted(tables[d], req[:, d], side="left").astype(np.uint64) + 1) << np.uint64(sh)
            T[:, w] |= np.searchsorted(tables[d], thr[:, d], side="right").astype(np.uint64) << np.uint64(sh)
        G = np.array(guard, dtype=np.uint64)
        assert np.all(R < 2 ** 32) and np.all(T < 2 ** 32) and np.all((T & G) == 0) and np.all((R & G) == 0)
        X = ((T | G)[None, :, :] - R[:, None, :]) & np.uint64(0xFFFFFFFF)   # 32-bit wraparound subtract
        packed = np.all((X & G) == G, axis=2)
        direct = np.all(req[:, None, :] <= thr[None, :, :], axis=2)
        assert np.array_equal(packed, direct), trial


def test_layout_limits():
    assert layout([8192] * 8)[2] == 4           # 8 x (14 + 1) bits: two fields per word
    assert layout([8192] * 9) is None            # more than 128 bits: the float64 scan is used
    assert layout([7, 6, 2, 1, 5, 5, 5, 5])[2] == 1   # BASELINE config 3's request table fits one word
