"""CPU: known answers for the hash-grid level table (SURVEY 8a R6 / reference core/nerf/gridencoder/grid.py:120-133,
gridencoder.cu:137-150) and consistency of the oracle's forward / backward, plus the host-side table the CUDA path uses."""
import numpy as np

from oracle import grid as ogrid


def test_level_table_known_answers():
    offsets, per_level_scale, _, scale, res = ogrid.level_table()
    sizes = np.diff(np.asarray(offsets, dtype=np.int64))
    # level sizes of the avatar grid (tiled, L=16, base 16, desired 4096, log2_hashmap_size 19)
    assert sizes[:5].tolist() == [4920, 15632, 42880, 125000, 373248]
    assert np.all(sizes[5:] == 524288) and len(sizes) == 16
    assert int(offsets[-1]) == 6328848
    assert abs(per_level_scale - 2.0 ** (8.0 / 15.0)) < 1e-6


def test_host_table_matches_oracle_table():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'dreamwaltz-g_b200'))
    from dwg import ops
    h_offsets, h_pls, h_S = ops.grid_level_table()
    offsets, pls, S, scale, res = ogrid.level_table()
    assert np.array_equal(h_offsets, offsets) and h_pls == pls and np.float32(h_S) == np.float32(S)
    # the per-level constants are the GPU-evaluated ones of the fixture (within 1 ulp of the correctly rounded value)
    lv = np.arange(16, dtype=np.float32)
    host = (np.exp2(lv * S).astype(np.float32) * np.float32(16) - np.float32(1.0)).astype(np.float32)
    assert np.all(np.abs(scale.view(np.int32).astype(np.int64) - host.view(np.int32).astype(np.int64)) <= 2)
    assert np.array_equal(res, np.ceil(scale).astype(np.uint32) + 1)


def test_constant_table_and_out_of_bounds():
    offsets, _, _, scale, res = ogrid.level_table()
    rng = np.random.default_rng(0)
    x = rng.uniform(-1.9, 1.9, size=(64, 3)).astype(np.float32)
    table = np.full((int(offsets[-1]), 2), 0.25, dtype=np.float32)
    out, dy, _ = ogrid.forward(x, table, offsets, scale, res, bound=2.0)
    np.testing.assert_allclose(out, 0.25, rtol=0, atol=1e-6)       # interpolation weights sum to 1 at every level
    np.testing.assert_allclose(dy, 0.0, atol=1e-5)                 # and a constant field has no spatial gradient
    xo = np.array([[2.5, 0.0, 0.0], [0.0, -3.0, 0.0]], dtype=np.float32)
    out_o, _, _ = ogrid.forward(xo, table, offsets, scale, res, bound=2.0)
    assert np.all(out_o == 0.0)                                    # gridencoder.cu:110-135: outside [0,1]^3 -> zeros


def test_backward_is_the_adjoint_of_forward():
    offsets, _, _, scale, res = ogrid.level_table()
    rng = np.random.default_rng(1)
    x = rng.uniform(-1.5, 1.5, size=(32, 3)).astype(np.float32)
    table = rng.normal(size=(int(offsets[-1]), 2)).astype(np.float32) * 0.1
    out, dy, _ = ogrid.forward(x, table, offsets, scale, res, bound=2.0)
    g = rng.normal(size=out.shape).astype(np.float32)
    gt, gx = ogrid.backward(g, x, table.shape, offsets, scale, res, dy_dx=dy, bound=2.0)
    # the encoding is linear in the table: <g, enc(table)> == <grad_table, table>
    np.testing.assert_allclose(float((g.astype(np.float64) * out).sum()), float((gt.astype(np.float64) * table).sum()), rtol=1e-4)
    # input gradient against central differences of the (smoothstep-interpolated, C1) encoding; only the three
    # coarsest levels carry features here (cells >= 0.1 wide), so a 1e-3 step stays inside one cell
    table = table.copy()
    table[int(offsets[3]):] = 0.0
    out, dy, _ = ogrid.forward(x, table, offsets, scale, res, bound=2.0)
    gt, gx = ogrid.backward(g, x, table.shape, offsets, scale, res, dy_dx=dy, bound=2.0)
    eps = 1e-3
    for d in range(3):
        xp, xm = x.copy(), x.copy()
        xp[:, d] += eps; xm[:, d] -= eps
        op, _, _ = ogrid.forward(xp, table, offsets, scale, res, bound=2.0)
        om, _, _ = ogrid.forward(xm, table, offsets, scale, res, bound=2.0)
        fd = ((op.astype(np.float64) - om) * g).sum(1) / (2 * eps)
        np.testing.assert_allclose(gx[:, d], fd, rtol=5e-2, atol=5e-2 * np.abs(fd).max())
