"""Reference-format snapshots: what Testbed::save_snapshot / load_snapshot exchange
(src/testbed.cu:3054-3113): a msgpack map = the network config JSON + a "snapshot" map with
  n_params, params_type ("__half"), params_binary   <- tcnn::Trainer::serialize (external; schema as described in SURVEY
                                                       Appendix B — not verifiable in this tree)
  density_grid_size (128), density_grid_binary (float32[5*128^3]), training_step, loss, nerf.{rgb, dataset}
`.ingp` files are the same msgpack stream behind zstr (zlib with a gzip wrapper, `testbed.cu:168-171,3173-3176`; zstr reads plain data
through unchanged). The render path needs three things from it: the parameter block (nsb_upload_model), the density grid
(nsb_upload_density_grid -> occupancy bitfield) and aabb_scale (render/train AABB, cone angle, per_level_scale).
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import abi

_BASE_CONFIG = {
    "encoding": {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19, "base_resolution": 16},
    "network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1},
    "dir_encoding": {"otype": "Composite", "nested": [{"n_dims_to_encode": 3, "otype": "SphericalHarmonics", "degree": 4},
                                                       {"otype": "Identity", "n_bins": 4, "degree": 4}]},
    "rgb_network": {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 2},
}


def save_snapshot(path: str, desc: abi.NsbModelDesc, params_u16: np.ndarray, density_grid: np.ndarray, aabb_scale: int, training_step: int = 0,
                  loss: float = 0.0, compress: bool = True) -> None:
    import msgpack

    cfg = json.loads(json.dumps(_BASE_CONFIG))
    cfg["encoding"].update({"n_levels": desc.n_levels, "n_features_per_level": desc.n_features_per_level, "log2_hashmap_size": desc.log2_hashmap_size,
                            "base_resolution": desc.base_resolution, "per_level_scale": float(desc.per_level_scale), "n_pos_dims": 3})
    p = np.ascontiguousarray(params_u16, np.uint16)
    g = np.ascontiguousarray(density_grid, np.float32).reshape(-1)
    assert g.size == abi.NSB_GRID_CELLS
    cfg["snapshot"] = {
        "n_params": int(p.size), "params_type": "__half", "params_binary": p.tobytes(),
        "density_grid_size": abi.NSB_NERF_GRIDSIZE, "density_grid_binary": g.tobytes(),
        "training_step": int(training_step), "loss": float(loss),
        "nerf": {"rgb": {"rays_per_batch": 1 << 12, "measured_batch_size": 0, "measured_batch_size_before_compaction": 0},
                 "dataset": {"aabb_scale": int(aabb_scale), "scale": 0.33, "offset": [0.5, 0.5, 0.5]}},
    }
    blob = msgpack.packb(cfg, use_bin_type=True)
    if path.lower().endswith(".ingp"):  # zstr::ostream: deflate with a gzip header (windowBits 15 + 16)
        import zlib

        co = zlib.compressobj(zlib.Z_DEFAULT_COMPRESSION if compress else zlib.Z_NO_COMPRESSION, zlib.DEFLATED, 15 + 16)
        blob = co.compress(blob) + co.flush()
    with open(path, "wb") as fh:
        fh.write(blob)


def load_snapshot(path: str):
    """-> (NsbModelDesc, params uint16[n], density_grid float32[5*128^3], aabb_scale)."""
    import msgpack

    with open(path, "rb") as fh:
        blob = fh.read()
    if blob[:2] == b"\x1f\x8b" or (len(blob) > 1 and blob[0] == 0x78 and (blob[0] * 256 + blob[1]) % 31 == 0):  # zstr::istream: gzip or zlib, else plain
        import zlib

        blob = zlib.decompress(blob, 15 + 32)
    cfg = msgpack.unpackb(blob, raw=False)
    enc, net, rgb = cfg["encoding"], cfg["network"], cfg["rgb_network"]
    snap = cfg["snapshot"]
    if snap.get("params_type", "__half") != "__half":
        raise abi.NsbError(f"unsupported params_type {snap.get('params_type')!r} (the render path is fp16)")
    if snap.get("density_grid_size", 128) != abi.NSB_NERF_GRIDSIZE:
        raise abi.NsbError("density_grid_size must be 128")
    aabb_scale = int(snap.get("nerf", {}).get("dataset", {}).get("aabb_scale", 1))
    pls = enc.get("per_level_scale", 0.0)
    if not pls:  # Testbed::reset_network derives it (testbed.cu:2280-2292)
        from .synthetic import per_level_scale

        pls = per_level_scale(aabb_scale, enc.get("n_levels", 16), enc.get("base_resolution", 16))
    desc = abi.NsbModelDesc(enc.get("n_levels", 16), enc.get("n_features_per_level", 2), enc.get("log2_hashmap_size", 19), enc.get("base_resolution", 16),
                            float(pls), net.get("n_neurons", 64), net.get("n_hidden_layers", 1), rgb.get("n_hidden_layers", 2), 4)
    params = np.frombuffer(snap["params_binary"], np.uint16).copy()
    if params.size != snap["n_params"]:
        raise abi.NsbError("params_binary does not hold n_params halves")
    grid = np.frombuffer(snap["density_grid_binary"], np.float32).copy()
    if grid.size != abi.NSB_GRID_CELLS:
        raise abi.NsbError(f"density grid has {grid.size} floats, expected {abi.NSB_GRID_CELLS}")
    return desc, params, grid, aabb_scale
