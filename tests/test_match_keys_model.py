"""CPU model of the matcher's key epilogue (csrc/surf_match_tc.cu, k_tc_candidates3 / k_tc_candidates3w).

The kernels turn every approximate score x = |t|^2 - 2 q.t into one sortable 32-bit key, push the keys through sorted
three-entry lists (quad tournament: only the smallest of four is listed) and hand k_tc_rescore four candidates per query
plus a lower bound `worst` of every score that is NOT among them.  The margin test of k_tc_rescore is only as good as
that bound, so the obligations are restated here with numpy float32 arithmetic and checked on random problems:

  1. the key order is the order of the scores up to the quantisation grid, padded rows sort last;
  2. decode(key) - slack <= score for every key (the bound never overshoots);
  3. for every query, every train row outside the four candidates has score >= worst;
  4. the best-scoring row is always a candidate; the second best is a candidate or worst <= its score
     (the query is then unproven and redone exactly).

This is a model of the arithmetic, not of the GPU schedule: it mirrors the column ownership (a tile of 128 train rows,
four column quarters of 32, quads of four consecutive columns) because obligations 3 and 4 depend on it."""
import numpy as np
import pytest

F = np.float32


def _unit_rows(rng, n, dim):
    t = np.abs(rng.standard_normal((n, dim))).astype(F)
    return t / np.linalg.norm(t, axis=1, keepdims=True)


def _fma32(a, b, c):
    """fl32(a * b + c) with one rounding: exact in float64 for float32 inputs of this magnitude."""
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(F)


class KeyModel:
    def __init__(self, qn_max, tn_max, rows_per_split):
        self.s = 9
        while (1 << self.s) < rows_per_split:
            self.s += 1
        assert self.s <= 14
        self.vbits = 32 - self.s
        self.pad_field = (1 << self.vbits) - 1
        nsum = F(qn_max) + F(tn_max)
        self.nsum = nsum
        self.off = F(F(qn_max) + F(0.01) * nsum)
        self.inv_sigma = F(F(2.04) * nsum + F(1e-30))
        self.sigma = F(F(1.0) / self.inv_sigma)
        self.m2sigma = F(F(-2.0) * self.sigma)
        self.base2 = F(2.0 ** (self.s - 9))
        self.grid = F(2.0 ** -self.vbits)

    def biased_norms(self, tn):
        out = _fma32(tn + self.off, self.sigma, np.full_like(tn, self.base2))
        pad = np.nextafter(F(2.0 ** (self.s - 8)), F(0))
        return np.where(np.isfinite(tn), out, pad).astype(F)

    def keys(self, acc, tnb, idx):
        """acc: (rows, cols) approximate q.t; tnb: (cols,); idx: (cols,) row of the split."""
        t = _fma32(acc, self.m2sigma, np.broadcast_to(tnb, acc.shape).astype(F))
        bits = t.view(np.uint32).astype(np.uint64)
        return ((bits << np.uint64(self.s)) & np.uint64(0xFFFFFFFF)) | idx.astype(np.uint64)

    def field(self, key):
        return key >> np.uint64(self.s)

    def decode_bound(self, field):
        """What the kernel writes into TcCand::worst for a bound key with this value field."""
        return F((F(field) - F(2.5)) * self.grid * self.inv_sigma - self.off - F(2e-6) * self.nsum)


def _epilogue(model, acc, tn, nt):
    """Per query row: the four candidates and `worst`, exactly as the kernels form them (one split)."""
    nq, ntp = acc.shape
    tnb = model.biased_norms(tn)
    keys = model.keys(acc, tnb, np.arange(ntp))
    empty = np.uint64(model.pad_field << model.s)
    cands, worst = [], []
    for r in range(nq):
        lists = []
        for quarter in range(4):
            m = [empty, empty, empty]
            bnd = np.uint64(0xFFFFFFFF)
            for tile0 in range(0, ntp, 128):
                c0 = tile0 + quarter * 32
                if c0 >= nt:  # a chunk of padding rows only
                    continue
                for c in range(c0, c0 + 32, 4):
                    q4 = sorted(int(k) for k in keys[r, c:c + 4])
                    bnd = min(bnd, np.uint64(q4[1]))          # the quad's second smallest bounds its three losers
                    m = sorted(m + [np.uint64(q4[0])])[:3]    # only the smallest is listed
            lists.append((m, min(m[2], bnd)))
        all12 = sorted(int(k) for m, _ in lists for k in m)
        third = min(int(b) for _, b in lists)
        best = all12[:5]
        cand = [k & ((1 << model.s) - 1) for k in best[:4] if (k >> model.s) != model.pad_field]
        bfield = min(third, best[4]) >> model.s
        cands.append(cand)
        worst.append(np.inf if bfield == model.pad_field else float(model.decode_bound(bfield)))
    return cands, np.array(worst), keys


@pytest.mark.parametrize("nq,nt,seed", [(48, 1000, 1), (32, 4230, 2), (40, 130, 3), (16, 5, 4)])
def test_candidates_and_bound_of_the_key_epilogue(nq, nt, seed):
    rng = np.random.default_rng(seed)
    D = 64
    t = _unit_rows(rng, nt, D)
    q = t[rng.integers(0, nt, nq)] + F(0.05) * rng.standard_normal((nq, D)).astype(F)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    ntp = (nt + 255) // 256 * 256                     # the pool pads a slot to a multiple of 256 rows
    tn = np.full(ntp, np.inf, F)
    tn[:nt] = (t.astype(np.float64) ** 2).sum(1).astype(F)
    qn = (q.astype(np.float64) ** 2).sum(1).astype(F)
    acc = np.zeros((nq, ntp), F)
    acc[:, :nt] = (q.astype(np.float64) @ t.astype(np.float64).T).astype(F)   # stands for the tensor-core accumulators
    model = KeyModel(qn.max(), tn[:nt].max(), ntp)
    cands, worst, keys = _epilogue(model, acc, tn, nt)
    score = (tn[None, :nt].astype(np.float64) - 2.0 * acc[:, :nt].astype(np.float64))   # x = |t|^2 - 2 q.t

    # 1. key order = score order up to the grid; padded rows last
    fields = model.field(keys)
    assert (fields[:, nt:] == model.pad_field).all() and (fields[:, :nt] < model.pad_field).all()
    grid_x = float(model.grid) * float(model.inv_sigma)
    for r in range(0, nq, 7):
        o = np.argsort(keys[r, :nt], kind="stable")
        assert (np.diff(score[r, o]) > -2.5 * grid_x).all()
    # 2. the decoded bound of ANY key never overshoots the score it stands for
    dec = (fields[:, :nt].astype(np.float64) - 2.5) * float(model.grid) * float(model.inv_sigma) - float(model.off) \
        - 2e-6 * float(model.nsum)
    assert (dec <= score + 1e-9).all()
    for r in range(nq):
        c = set(cands[r])
        assert 1 <= len(c) <= 4 and all(0 <= i < nt for i in c)   # (a quad lists one of its four rows: tiny train sets give fewer)
        rest = np.array([i for i in range(nt) if i not in c], int)
        # 3. everything outside the candidates is bounded
        if len(rest):
            assert worst[r] <= score[r, rest].min() + 1e-9
        else:
            assert worst[r] == np.inf
        # 4. best always a candidate; second best a candidate or flagged by the bound
        o = np.argsort(score[r], kind="stable")
        assert int(o[0]) in c
        if nt > 1 and int(o[1]) not in c:
            assert worst[r] <= score[r, o[1]] + 1e-9


def test_key_fields_use_every_bit_of_the_value_range():
    """Scores span [-max|q|^2, 2 max|t|^2 + max|q|^2]: both ends must land inside the value field, below the pad value."""
    for rows in (512, 4352, 16384):
        m = KeyModel(1.0, 1.0, rows)
        tnb = m.biased_norms(np.array([1.0], F))
        for qt in (-1.0, 0.0, 1.0):   # q.t of unit vectors
            k = m.keys(np.array([[qt]], F), tnb, np.array([0]))
            f = int(m.field(k)[0, 0])
            assert 0 < f < m.pad_field
        lo = int(m.field(m.keys(np.array([[1.0]], F), tnb, np.array([0])))[0, 0])
        hi = int(m.field(m.keys(np.array([[-1.0]], F), tnb, np.array([0])))[0, 0])
        assert hi - lo > 0.9 * m.pad_field   # the range of the scores fills > 90 % of the field: the grid is not wasted
