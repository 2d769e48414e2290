"""§8(f)-3/-2: reference-format snapshot round trip and the density-grid -> occupancy-bitfield step of load_snapshot."""
import numpy as np
import pytest

from nerfshop_b200 import abi, snapshot, synthetic as syn
from oracle import oracle as orc


def _grid_from_occupancy(occ):
    return np.unpackbits(occ, bitorder="little").astype(np.float32)  # 1.0 where occupied: mean of cascade 0 >> 0.01 -> threshold 0.01


def test_grid_to_bitfield_reproduces_the_occupancy(scene):
    model, occ = scene
    bits, mean = orc.density_grid_to_bitfield(_grid_from_occupancy(occ))
    assert mean > 0.01
    n = 3 * 128 ** 3 // 8  # cascades 0..max_cascade(2): thresholding is exact and the max-pool of an already pooled grid is idempotent
    assert np.array_equal(bits[:n], occ[:n])
    b = np.unpackbits(bits, bitorder="little").reshape(5, -1)
    assert b[3].sum() > 0 and b[4].sum() > 0  # the reference max-pools into EVERY coarser cascade (unused above max_cascade)
    assert b[3].sum() * 8 >= b[2].sum() >= b[3].sum()


def test_grid_threshold_is_min_of_mean_and_optical_thickness():
    g = np.zeros(abi.NSB_GRID_CELLS, np.float32)
    g[:1000] = 0.004          # mean of cascade 0 = 1000*0.004/128^3 ~ 1.9e-6 < 0.01 -> threshold = mean
    g[5] = -1.0               # negative cells count as 0 in the mean
    bits, mean = orc.density_grid_to_bitfield(g)
    assert 0 < mean < 0.01
    b = np.unpackbits(bits, bitorder="little")
    assert b[:1000].sum() == 999 and b[5] == 0 and b[1000:128 ** 3].sum() == 0
    # cascade 1 = max-pool of cascade 0 (cells 0..999 are Morton-contiguous: the first 2x2x2 blocks of the inner cube)
    assert b[128 ** 3:2 * 128 ** 3].sum() == int(np.ceil(1000 / 8))


def test_snapshot_roundtrip(tmp_path, scene):
    model, occ = scene
    path = str(tmp_path / "synthetic.msgpack")
    grid = _grid_from_occupancy(occ)
    snapshot.save_snapshot(path, model.desc, model.params, grid, model.aabb_scale, training_step=123)
    desc, params, grid2, aabb_scale = snapshot.load_snapshot(path)
    assert aabb_scale == 4 and np.array_equal(params, model.params) and np.array_equal(grid2, grid)
    for f, _ in abi.NsbModelDesc._fields_:
        assert getattr(desc, f) == getattr(model.desc, f)


def test_ingp_is_the_same_stream_behind_zlib(tmp_path, scene):
    """.ingp = the msgpack stream through zstr (gzip-wrapped deflate, testbed.cu:168-171,3173-3176); uncompressed .ingp also loads."""
    import gzip
    import os

    model, occ = scene
    grid = _grid_from_occupancy(occ)
    plain, packed, stored = (str(tmp_path / n) for n in ("s.msgpack", "s.ingp", "stored.ingp"))
    snapshot.save_snapshot(plain, model.desc, model.params, grid, model.aabb_scale)
    snapshot.save_snapshot(packed, model.desc, model.params, grid, model.aabb_scale)
    snapshot.save_snapshot(stored, model.desc, model.params, grid, model.aabb_scale, compress=False)
    assert gzip.open(packed, "rb").read() == open(plain, "rb").read() == gzip.open(stored, "rb").read()
    assert os.path.getsize(packed) < os.path.getsize(plain) <= os.path.getsize(stored)
    for path in (packed, stored):
        desc, params, grid2, aabb_scale = snapshot.load_snapshot(path)
        assert aabb_scale == 4 and np.array_equal(params, model.params) and np.array_equal(grid2, grid)
    # a file named .ingp that holds the plain stream (zstr reads it through unchanged)
    raw_ingp = str(tmp_path / "raw.ingp")
    open(raw_ingp, "wb").write(open(plain, "rb").read())
    assert np.array_equal(snapshot.load_snapshot(raw_ingp)[1], model.params)


@pytest.mark.gpu
def test_snapshot_to_render(tmp_path, scene, renderer):
    """load_snapshot path end to end: msgpack -> upload_model + upload_density_grid -> the same frame as the direct uploads."""
    from nerfshop_b200.renderer import NerfRenderer

    model, occ = scene
    path = str(tmp_path / "synthetic.msgpack")
    snapshot.save_snapshot(path, model.desc, model.params, _grid_from_occupancy(occ), model.aabb_scale)
    desc, params, grid, _ = snapshot.load_snapshot(path)
    r = NerfRenderer(0)
    try:
        r.upload_model(desc, params)
        bits = r.upload_density_grid(grid)
        ref_bits, _ = orc.density_grid_to_bitfield(grid)
        n = 3 * 128 ** 3 // 8
        assert np.array_equal(bits, ref_bits) and np.array_equal(bits[:n], occ[:n])
        f = syn.make_frame(model, syn.orbit_cameras(120)[17], 96, 54)
        a, _ = r.render(f)
        b, _ = renderer.render(f)
        assert (a == b).all()
        g = np.random.default_rng(1).random(abi.NSB_GRID_CELLS, dtype=np.float32) * 0.02  # generic grid: mean of cascade 0 = 0.01 boundary region
        g[3 * 128 ** 3:] = 0
        assert np.array_equal(r.upload_density_grid(g), orc.density_grid_to_bitfield(g)[0])
    finally:
        r.close()
