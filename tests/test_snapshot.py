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


def test_reads_the_export_snapshot_schema(tmp_path, scene):
    """Testbed::export_snapshot (testbed.cu:3118-3183): fp16 density grid of (max_cascade+1) cascades, aabb_scale at snapshot.nerf.aabb_scale, no dataset,
    .ingp = the same stream behind zlib/gzip. Written here key by key from that function, then loaded."""
    import zlib

    import msgpack

    model, occ = scene
    grid = _grid_from_occupancy(occ)
    n_casc = 3  # fox: max_cascade = 2
    cfg = {"encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16},
           "network": {"otype": "FullyFusedMLP", "n_neurons": 64, "n_hidden_layers": 1}, "rgb_network": {"otype": "FullyFusedMLP", "n_neurons": 64, "n_hidden_layers": 2},
           "snapshot": {"version": 1, "n_params": int(model.params.size), "params_type": "__half", "params_binary": model.params.tobytes(),
                        "density_grid_size": 128, "density_grid_binary": grid[: n_casc * 128 ** 3].astype(np.float16).tobytes(),
                        "training_step": 35000, "loss": 0.001, "aabb": {"min": [-1.5] * 3, "max": [2.5] * 3},
                        "nerf": {"aabb_scale": 4, "rgb": {"rays_per_batch": 4096, "measured_batch_size": 0, "measured_batch_size_before_compaction": 0}}}}
    blob = msgpack.packb(cfg, use_bin_type=True)
    co = zlib.compressobj(zlib.Z_DEFAULT_COMPRESSION, zlib.DEFLATED, 15 + 16)
    path = tmp_path / "export.ingp"
    path.write_bytes(co.compress(blob) + co.flush())
    desc, params, g, aabb_scale = snapshot.load_snapshot(str(path))
    assert aabb_scale == 4 and abs(desc.per_level_scale - model.desc.per_level_scale) < 1e-6  # derived like Testbed::reset_network (the file has none)
    assert np.array_equal(params, model.params)
    assert np.array_equal(g[: n_casc * 128 ** 3], grid[: n_casc * 128 ** 3]) and not g[n_casc * 128 ** 3:].any()
    bits, _ = orc.density_grid_to_bitfield(g)
    assert np.array_equal(bits[: n_casc * 128 ** 3 // 8], occ[: n_casc * 128 ** 3 // 8])


def test_written_snapshot_has_every_key_the_reference_reads(tmp_path, scene):
    """save_snapshot's file must survive Testbed::load_snapshot: NerfDataset::from_json (json_binding.h:164-194) reads these keys with .at()."""
    import msgpack

    model, occ = scene
    path = str(tmp_path / "full.msgpack")
    snapshot.save_snapshot(path, model.desc, model.params, _grid_from_occupancy(occ), model.aabb_scale)
    cfg = msgpack.unpackb(open(path, "rb").read(), raw=False)
    ds = cfg["snapshot"]["nerf"]["dataset"]
    for key in ("n_images", "xforms", "render_aabb", "up", "offset", "image_resolution", "envmap_resolution", "scale", "aabb_scale", "from_mitsuba"):
        assert key in ds, key
    assert ds["n_images"] == 0 and ds["xforms"] == [] and ds["render_aabb"]["min"] == [-1.5] * 3 and ds["render_aabb"]["max"] == [2.5] * 3
    assert cfg["snapshot"]["nerf"]["aabb_scale"] == 4 and cfg["snapshot"]["density_grid_size"] == 128
    for key in ("rays_per_batch", "measured_batch_size", "measured_batch_size_before_compaction"):  # testbed.cu:3064-3066
        assert key in cfg["snapshot"]["nerf"]["rgb"]


def test_missing_aabb_scale_is_an_error(tmp_path, scene):
    import msgpack

    model, occ = scene
    cfg = {"encoding": {}, "network": {}, "rgb_network": {}, "snapshot": {"n_params": 0, "params_type": "__half", "params_binary": b"", "density_grid_size": 128,
                                                                          "density_grid_binary": b"", "nerf": {"rgb": {}}}}
    p = tmp_path / "bad.msgpack"
    p.write_bytes(msgpack.packb(cfg, use_bin_type=True))
    with pytest.raises(abi.NsbError, match="aabb_scale"):
        snapshot.load_snapshot(str(p))
