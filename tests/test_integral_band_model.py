"""CPU model of the single-pass integral kernel (csrc/surf_kernels.cu, k_integral_band): the decomposition it relies on,
restated with numpy integer arithmetic and checked against the plain double cumulative sum.

  band of R rows, 4 row groups of R/4 rows:  V[r][x]  = sum of the pixels of column x from the top of r's row group to r   (u16)
                                             G[g][x]  = column totals of row group g                                      (u16)
                                             T[x]     = column totals of the band (published for the bands below)          (u32)
  carry of row group g of band k:            C[g][x]  = sum_{j<k} T_j[x] + sum_{g'<g} G[g'][x]                             (u32)
  output row y0 + r + 1:                     S[.][x+1] = inclusive prefix over x of (C[g(r)][x] + V[r][x]), S[.][0] = 0

The 16-bit fields must hold their worst case (a row group of eight saturated rows), the 32-bit ones the whole frame."""
import numpy as np
import pytest


def _model(img, R):
    H, W = img.shape
    gR = R // 4
    out = np.zeros((H + 1, W + 1), np.uint32)
    totals = []
    for y0 in range(0, H, R):
        rows = min(R, H - y0)
        band = img[y0:y0 + rows].astype(np.uint32)
        V = np.zeros((rows, W), np.uint32)
        G = np.zeros((4, W), np.uint32)
        for g in range(4):
            r0, r1 = g * gR, min(rows, (g + 1) * gR)
            if r0 < r1:
                V[r0:r1] = np.cumsum(band[r0:r1], axis=0)
                G[g] = V[r1 - 1]
        assert V.max(initial=0) <= 0xFFFF and G.max(initial=0) <= 0xFFFF      # the kernel keeps them in 16 bits
        carry = np.sum(totals, axis=0, dtype=np.uint64).astype(np.uint32) if totals else np.zeros(W, np.uint32)
        C = np.zeros((4, W), np.uint32)
        C[0] = carry
        for g in range(1, 4):
            C[g] = C[g - 1] + G[g - 1]
        for r in range(rows):
            vals = C[min(r // gR, 3)] + V[r]
            # the kernel's lanes own output columns 4l..4l+3 = the prefix of the values at image columns 4l-1..4l+2
            out[y0 + r + 1, 1:] = np.cumsum(vals, dtype=np.uint64).astype(np.uint32)
        totals.append(G.sum(axis=0, dtype=np.uint64).astype(np.uint32))
    return out


@pytest.mark.parametrize("shape,R", [((480, 640), 32), ((70, 1400), 28), ((9, 4100), 4), ((33, 129), 32), ((1, 1), 32)])
def test_band_decomposition_equals_the_cumulative_sum(shape, R):
    rng = np.random.default_rng(shape[0] + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    want = np.zeros((shape[0] + 1, shape[1] + 1), np.uint64)
    want[1:, 1:] = np.cumsum(np.cumsum(img.astype(np.uint64), axis=0), axis=1)
    assert np.array_equal(_model(img, R).astype(np.uint64), want)


def test_sixteen_bit_fields_hold_a_saturated_row_group():
    img = np.full((64, 40), 255, np.uint8)
    got = _model(img, 32)          # asserts V, G <= 0xFFFF inside (8 rows x 255 = 2 040)
    assert got[-1, -1] == 255 * 64 * 40
