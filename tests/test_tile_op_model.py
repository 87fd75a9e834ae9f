"""Lane-level model of `tile_op_body_v2` (alfi_b200/csrc/condense.cu) in numpy.

The CUDA tile op keeps a ring of UB column loads per lane, switches the gathered source every 32
columns and sums G column groups with xor-shuffles.  This restates exactly that index arithmetic —
32 lanes, the same slot / column / shuffle-lane formulas — and checks it against `M @ x` for every
row count 1..64 and a spread of column counts, so that an off-by-one in the ring refill, the chunk
switch or the group reduction is caught without a GPU.  (The GPU tests then check the real kernel.)
"""
import numpy as np
import pytest

UB = 8


def tile_op_v2_model(M, x):
    nrows, n = M.shape
    half = (nrows + 1) // 2
    G = 4 if half <= 8 else 2 if half <= 16 else 1
    LPG, CPB = 32 // G, UB * G
    tile = np.zeros((n, 2 * half))                  # column-major, roundup2(nrows) rows per column
    tile[:, :nrows] = M.T
    lanes = np.arange(32)
    grp, l = lanes // LPG, lanes % LPG
    active = l < half

    def column(c):                                  # per-lane column index -> the lane's double2 (or zeros)
        out = np.zeros((32, 2))
        for ln in range(32):
            if active[ln] and c[ln] < n:
                out[ln] = tile[c[ln], 2 * l[ln]:2 * l[ln] + 2]
        return out

    def value_at(pos):
        return np.array([x[p] if p < n else 0.0 for p in pos])

    a = [column(u * G + grp) for u in range(UB)]
    xv, xnext = value_at(lanes), value_at(32 + lanes)
    acc, bcc = np.zeros((32, 2)), np.zeros((32, 2))
    for cb in range(0, n, CPB):
        assert cb // 32 == (cb + CPB - 1) // 32     # a round never straddles a chunk
        if cb > 0 and cb % 32 == 0:
            xv, xnext = xnext, value_at(cb + 32 + lanes)
        for u in range(UB):
            c = cb + u * G + grp
            xc = xv[c & 31]                         # __shfl_sync(full mask, xv, c & 31)
            if u & 1:
                bcc += a[u] * xc[:, None]
            else:
                acc += a[u] * xc[:, None]
            a[u] = column(c + CPB)
    acc = acc + bcc
    off = 16
    while off >= LPG:
        acc = acc + acc[lanes ^ off]
        off >>= 1
    y = np.zeros(nrows)
    for ln in range(32):
        if grp[ln] == 0 and active[ln]:
            r = 2 * l[ln]
            y[r] = acc[ln, 0]
            if r + 1 < nrows:
                y[r + 1] = acc[ln, 1]
    return y


@pytest.mark.parametrize("nrows", list(range(1, 65)))
def test_model_equals_matvec(nrows):
    rng = np.random.default_rng(nrows)
    for n in sorted({1, 2, 7, 8, 9, 31, 32, 33, 45, 63, 64, 65, 75, 96, 105, 128, int(rng.integers(1, 129))}):
        M = rng.standard_normal((nrows, n))
        x = rng.standard_normal(n)
        y = tile_op_v2_model(M, x)
        assert np.allclose(y, M @ x, rtol=1e-12, atol=1e-12), (nrows, n)


def test_x_ss_tile_width():
    """X_SS ops have up to 64 rows and as many columns as the separator has dofs (195 in 3-D, more in edge cases)."""
    rng = np.random.default_rng(0)
    for nrows, n in [(64, 195), (3, 195), (64, 257), (45, 300)]:
        M, x = rng.standard_normal((nrows, n)), rng.standard_normal(n)
        assert np.allclose(tile_op_v2_model(M, x), M @ x, rtol=1e-12, atol=1e-12)
