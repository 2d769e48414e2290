"""§8(f)-2, host side: pcg32 pinned by the PCG reference known-answer vector, the Python m_rng mirror, and the oracle's
restatement of Testbed::update_density_grid_nerf_operator (testbed_nerf.cu:3533-3639) checked through properties."""
import numpy as np

import edit_fixtures as fx
from nerfshop_b200 import abi
from nerfshop_b200.rng import Pcg32
from oracle import oracle as orc


def test_pcg32_known_answer_and_advance():
    # pcg32-demo of the PCG reference implementation: seed (42, 54) -> these six 32-bit outputs
    kat = [0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E]
    st = orc.pcg32_seed(42, 54)
    assert orc.pcg32_next(st, 6).tolist() == kat
    r = Pcg32(42, 54)
    assert [r.next_uint() for _ in range(6)] == kat
    # skip-ahead == stepping, in the oracle and in the host mirror (tcnn default seed/stream)
    st = orc.pcg32_seed(1337, 0xDA3E39CB94B95BDB)
    st2 = st.copy()
    seq = orc.pcg32_next(st, 5000)
    for k in (0, 1, 4, 997, 4096):
        s = st2.copy()
        orc.pcg32_advance(s, k)
        assert orc.pcg32_next(s, 3).tolist() == seq[k:k + 3].tolist()
        r = Pcg32(1337)
        assert (r.state, r.inc) == (int(st2[0]), int(st2[1]))
        r.advance(k)
        assert [r.next_uint() for _ in range(3)] == seq[k:k + 3].tolist()
    r = Pcg32(1337)
    r.advance()  # 2^32, the tcnn default
    s = st2.copy()
    orc.pcg32_advance(s, 1 << 32)
    assert r.state == int(s[0])


def grid_params(n_uniform, n_nonuniform, rng, ema_step=0, reset=True, n_cascades=3, apply_ops=True):
    u = abi.NsbGridUpdate()
    u.n_uniform_samples, u.n_nonuniform_samples, u.reset_grid, u.n_cascades = n_uniform, n_nonuniform, int(reset), n_cascades
    u.decay, u.ema_step, u.rng_state, u.rng_inc = 0.95, ema_step, rng.state, rng.inc
    u.train_aabb_min[:] = (-1.5, -1.5, -1.5)
    u.train_aabb_max[:] = (2.5, 2.5, 2.5)
    u.density_activation, u.apply_operators = abi.NSB_ACT_EXPONENTIAL, int(apply_ops)
    return u


def test_oracle_grid_update_properties(scene, oracle):
    model, occ = scene
    G = 128 ** 3
    rng = Pcg32(1337)
    u = grid_params(300_000, 0, rng)
    grid0 = np.full(abi.NSB_GRID_CELLS, 0.5, np.float32)  # reset_grid must wipe this
    grid, bits, mean, samples = oracle.update_density_grid(u, grid0, want_samples=True)
    assert np.isfinite(grid).all() and (grid >= 0).all()
    assert (grid[3 * G:] == 0).all(), "cascades beyond n_cascades are never sampled"
    touched = grid > 0
    assert 150_000 < touched.sum() <= 300_000
    # every sample lies in the unit cube of its cascade, and the density it splats is exp(raw) * MIN_CONE_STEPSIZE >= its cell's value
    assert (samples[:, :3] > 0.5 - 4 / 4 / 2 - 1e-6).all() and (samples[:, :3] < 0.5 + 4 / 4 / 2 + 1e-6).all()
    assert np.isclose(grid.max(), samples[:, 3].max() * np.float32(np.sqrt(3) / 1024), rtol=1e-6)
    # bitfield == grid_to_bitfield + max-pool of the same grid
    bits2, mean2 = orc.density_grid_to_bitfield(grid)
    assert np.array_equal(bits, bits2) and mean == mean2
    # determinism, and a second step (ema_step 1, no reset) keeps max(prev*decay, new) and never lowers an untouched cell below decay
    g2, b2, m2 = oracle.update_density_grid(u, grid0)
    assert np.array_equal(g2, grid) and np.array_equal(b2, bits)
    rng.advance(); rng.advance()
    u2 = grid_params(100_000, 50_000, rng, ema_step=1, reset=False)
    g3, _, _ = oracle.update_density_grid(u2, grid)
    assert (g3 >= grid * np.float32(0.95)).all()
    assert ((g3 > grid * np.float32(0.95)) & touched).sum() > 1000
    # cells marked untrained (< 0) stay as they are and are avoided by the sampler
    gneg = grid.copy()
    gneg[:G:2] = -1.0
    g4, _, _ = oracle.update_density_grid(u2, gneg)
    assert (g4[:G:2] == -1.0).all()


def test_oracle_grid_update_moves_occupancy_with_the_cage(scene):
    """A cage that pulls the fox head outwards must switch on cells in the vacated->deformed region that are empty without it."""
    model, occ = scene
    rng = Pcg32(7)
    u = grid_params(128 ** 3, 0, rng, n_cascades=1)
    base = orc.Oracle(model.desc, model.params, occ)
    g0, b0, _ = base.update_density_grid(u, np.zeros(abi.NSB_GRID_CELLS, np.float32))
    ops = [o.to_op() for o in fx.e1(model)]
    edited = orc.Oracle(model.desc, model.params, occ, ops)
    g1, b1, _ = edited.update_density_grid(u, np.zeros(abi.NSB_GRID_CELLS, np.float32))
    diff = g0 != g1
    assert 100 < diff.sum() < 200_000
    # differences are confined to cascade-0 cells inside the union of the deformed and canonical boxes
    idx = np.nonzero(diff[:128 ** 3])[0]
    def inv(x):
        x = x & 0x49249249; x = (x | (x >> 2)) & 0xC30C30C3; x = (x | (x >> 4)) & 0x0F00F00F; x = (x | (x >> 8)) & 0xFF0000FF; x = (x | (x >> 16)) & 0xFFFF
        return x
    cells = np.stack([inv(idx), inv(idx >> 1), inv(idx >> 2)], 1).astype(np.float32)
    lo, hi = (cells / 128), (cells + 1) / 128
    op = fx.e1(model)[0]
    bmin = np.minimum(op.vertices.min(0), op.original_vertices.min(0)) - 1e-4
    bmax = np.maximum(op.vertices.max(0), op.original_vertices.max(0)) + 1e-4
    assert (hi >= bmin).all() and (lo <= bmax).all()


def test_python_sample_draw_equals_the_oracle(scene, oracle):
    """generate_grid_samples_nerf_nonuniform (common_nerf.cu:179-208) restated in Python on top of the host pcg32 mirror: the warped
    position of every checked sample equals the oracle's, bit for bit — for both launches (the second one runs on m_rng advanced by 2^32)."""
    import struct

    G = 128 ** 3
    rng0 = Pcg32(4242)
    grid0 = np.zeros(abi.NSB_GRID_CELLS, np.float32)
    grid0[::3] = 0.02      # some cells above the non-uniform threshold 0.01
    grid0[5::7] = -1.0     # some untrained cells, which both passes must step over
    n_uni, n_non, step = 5000, 3000, 7
    u = grid_params(n_uni, n_non, rng0, ema_step=step, reset=False, n_cascades=3, apply_ops=False)
    _, _, _, samples = oracle.update_density_grid(u, grid0, want_samples=True)

    def next_float(r):
        return np.float32(struct.unpack("<f", struct.pack("<I", (r.next_uint() >> 9) | 0x3F800000))[0]) - np.float32(1.0)

    def inv_morton(m):
        return [sum(((m >> (3 * b + k)) & 1) << b for b in range(7)) for k in range(3)]

    f = np.float32
    for i in list(range(0, n_uni, 97)) + list(range(n_uni, n_uni + n_non, 61)):
        second = i >= n_uni
        li, n_el, thresh = (i - n_uni, n_non, 0.01) if second else (i, n_uni, -0.01)
        r = rng0.copy()
        if second:
            r.advance(1 << 32)
        r.advance((li * 4) & 0xFFFFFFFF)
        level = int(next_float(r) * f(3)) % 3
        idx = 0
        for j in range(10):
            idx = (((li + step * n_el) * 56924617 + j * 19349663 + 96925573) & 0xFFFFFFFF) % G + level * G
            if grid0[idx] > thresh:
                break
        x, y, z = inv_morton(idx % G)
        rnd = [next_float(r) for _ in range(3)]
        sc = f(2.0 ** level)
        pos = [((f(c) + q) / f(128) - f(0.5)) * sc + f(0.5) for c, q in zip((x, y, z), rnd)]
        pw = np.array([(pos[k] - f(-1.5)) / (f(2.5) - f(-1.5)) for k in range(3)], np.float32)
        assert np.array_equal(pw.view(np.uint32), samples[i, :3].view(np.uint32)), (i, pw, samples[i, :3])
