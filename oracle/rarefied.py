"""numpy restatement of the two RE-DRAWN outputs of the hot path (test infrastructure only).

    clonTR                      calculate_rarefied_clonality     inStrain/profile/snv_utilities.py:233-247 (set at :98-102)
    r2_normalized, d_prime_normalized        _calc_ld_single      inStrain/profile/linkage.py:200-228

The reference draws both with an UNSEEDED np.random.choice, so nothing of theirs can be pinned bit for bit (its own tests
delete these columns before comparing: test/tests/test_profile.py:896-900).  The CUDA path draws from a counter-based
generator instead (instrain_b200/csrc/isb_k2_site.cuh: splitmix64's finaliser over (seed, stream tag, site key, word index),
the n draws sampled bit-sliced as a chain of binomials: isb_redraw4).  This module restates exactly that
generator and the arithmetic around it, so the CUDA outputs ARE bit-exact against the oracle for a given seed; that the
construction has the reference's distribution is checked separately against the reference's own functions
(tests/test_reference_mirrors.py).
"""
import numpy as np

TAG_CLONR = np.uint64(0x636c6f6e54520001)
TAG_LD = np.uint64(0x6c646e6f726d0002)
_K1, _K2, _K3 = np.uint64(0x9e3779b97f4a7c15), np.uint64(0xd1b54a32d192ed03), np.uint64(0x8cb92ba72f3d8dd7)
_M1, _M2 = np.uint64(0xbf58476d1ce4e5b9), np.uint64(0x94d049bb133111eb)


def mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


_K4 = np.uint64(0xa0761d6478bd642f)


def rng_base(seed, tag, a, b):
    with np.errstate(over="ignore"):
        return mix64(np.uint64(seed) + tag) ^ (np.asarray(a, np.uint64) * _K1) ^ (np.asarray(b, np.uint64) * _K2)


def redraw4(c, T, n, rb):
    """isb_redraw4 (instrain_b200/csrc/isb_k2_site.cuh), row-wise: n draws with replacement from 4 categories with counts
    c[rows, 4] (totals T), bit-sliced -- the trials are the bit lanes of a 64-bit word, categories are split off one after
    the other (a binomial with p = c_i / remaining), a random word decides one bit of the comparison "uniform < p" for all
    trials at once.  Returns int64 r[rows, 4] (row sums = n)."""
    c = np.asarray(c, dtype=np.int64)
    rows = len(c)
    r = np.zeros((rows, 4), dtype=np.int64)
    q = 0
    with np.errstate(over="ignore", divide="ignore", invalid="ignore"):
        while q * 64 < n:
            nt = min(64, n - q * 64)
            left = np.full(rows, np.uint64(0xFFFFFFFFFFFFFFFF) if nt == 64 else np.uint64((1 << nt) - 1), dtype=np.uint64)
            base = rb ^ (np.uint64(q) * _K4)
            kk = np.zeros(rows, dtype=np.uint64)
            rem = np.asarray(T, dtype=np.int64).copy()
            for i in range(4):
                ci = c[:, i]
                act = (ci != 0) & (left != 0)
                last = act & (ci == rem)
                r[last, i] += _popc64(left[last])
                left[last] = 0
                go = act & ~last
                P = np.zeros(rows, dtype=np.uint64)
                P[go] = np.floor(ci[go].astype(np.float64) / rem[go].astype(np.float64) * 4294967296.0).astype(np.uint64)
                und = np.where(go, left, np.uint64(0))
                hit = np.zeros(rows, dtype=np.uint64)
                for b in range(31, -1, -1):
                    run = und != 0
                    if not run.any():
                        break
                    W = mix64(base ^ kk)
                    kk = np.where(run, kk + _K3, kk)
                    one = ((P >> np.uint64(b)) & np.uint64(1)) != 0
                    hit = np.where(run & one, hit | (~W & und), hit)
                    und = np.where(run, np.where(one, und & W, und & ~W), und)
                r[:, i] += np.where(go, _popc64(hit), 0)
                left = np.where(go, left & ~hit, left)
                rem = np.where(go, rem - ci, rem)
            q += 1
    return r


def _popc64(x):
    x = np.asarray(x, dtype=np.uint64)
    out = np.zeros(len(x), dtype=np.int64)
    for sh in range(0, 64, 16):
        out += _POP16[((x >> np.uint64(sh)) & np.uint64(0xFFFF)).astype(np.int64)]
    return out


_POP16 = np.array([bin(i).count("1") for i in range(1 << 16)], dtype=np.int64)


def clonality_of(c, s):
    """calculate_clonality (snv_utilities.py:225-231): double, A,C,T,G order."""
    f = c.astype(np.float64) / np.float64(s)
    return ((f[:, 0] * f[:, 0] + f[:, 1] * f[:, 1]) + f[:, 2] * f[:, 2]) + f[:, 3] * f[:, 3]


def clonTR(counts, nmask, rarefied_coverage=50, seed=0, start=0):
    """Dense float32[L, M]: at every (position, level) whose level is a key of the position's MMcounts and whose cumulative
    coverage reaches rarefied_coverage, the clonality of rarefied_coverage re-drawn bases; NaN elsewhere."""
    L, M, _ = counts.shape
    out = np.full((L, M), np.nan, dtype=np.float32)
    if rarefied_coverage <= 0:
        return out
    present = (counts.sum(2) > 0) | (((nmask[:, None] >> np.arange(M, dtype=np.uint64)[None, :]) & np.uint64(1)) != 0)
    cum = np.cumsum(np.where(present[:, :, None], counts, 0).astype(np.int64), axis=1)       # counts of the levels present
    T = cum.sum(2)
    sel = present & (T >= rarefied_coverage)
    p, m = np.nonzero(sel)
    if len(p) == 0:
        return out
    C = cum[p, m]
    Ts = T[p, m]
    single = C.max(1) == Ts
    val = np.ones(len(p), dtype=np.float64)
    q = np.nonzero(~single)[0]
    if len(q):
        rb = rng_base(seed, TAG_CLONR, (p[q] + start).astype(np.uint64), m[q].astype(np.uint64))
        val[q] = clonality_of(redraw4(C[q], Ts[q], rarefied_coverage, rb), rarefied_coverage)
    out[p, m] = val.astype(np.float32)
    return out


def normalized_ld(rows, min_snp=20, seed=0):
    """(r2_normalized, d_prime_normalized) float64 arrays for LD rows (batch coordinates in pos_a / pos_b)."""
    n = len(rows)
    r2n, dpn = np.full(n, np.nan), np.full(n, np.nan)
    if n == 0 or min_snp < 1:
        return r2n, dpn
    c = np.stack([rows["c_AB"], rows["c_Ab"], rows["c_aB"], rows["c_ab"]], 1).astype(np.int64)
    total = c.sum(1)
    key = (rows["pos_a"].astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)) << np.uint64(32)
    key |= rows["pos_b"].astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    g = redraw4(c, total, min_snp, rng_base(seed, TAG_LD, key, rows["mm"].astype(np.uint64))) / np.float64(min_snp)
    gAB, gAb, gaB, gab = g[:, 0], g[:, 1], g[:, 2], g[:, 3]
    gA, ga, gB, gb = gAB + gAb, gab + gaB, gAB + gaB, gab + gAb
    ldn = gab - ga * gb
    ok = ~((ga == 0) | (gA == 0) | (gB == 0) | (gb == 0))
    with np.errstate(divide="ignore", invalid="ignore"):
        r2n[ok] = (ldn[ok] * ldn[ok]) / (((gA[ok] * ga[ok]) * gB[ok]) * gb[ok])
        neg, pos = ldn < 0, ldn > 0
        dpn[neg] = ldn[neg] / np.maximum(-gA[neg] * gB[neg], -ga[neg] * gb[neg])
        dpn[pos] = ldn[pos] / np.minimum(gA[pos] * gb[pos], ga[pos] * gB[pos])
    return r2n, dpn
