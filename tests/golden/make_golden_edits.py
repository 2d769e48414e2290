"""Mints tests/golden/synthetic_edits.npz from the CPU oracle and the host geometry: the §8(f) rows and BASELINE configs 2-3
(cage edits, membrane, occupancy update, boundary values). Same status as make_golden.py: self-minted pins ("parity unpinned").
Run:  python tests/golden/make_golden_edits.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import edit_fixtures as fx  # noqa: E402
from conftest import random_coords  # noqa: E402
from nerfshop_b200 import abi, synthetic as syn  # noqa: E402
from nerfshop_b200.rng import Pcg32  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from test_grid_update_cpu import grid_params  # noqa: E402
from test_membrane_cpu import sampling  # noqa: E402

W, H = 96, 54


def coords_around(model, center, half, n, seed):
    rng = np.random.default_rng(seed)
    p = np.asarray(center, np.float32) + (rng.random((n, 3)).astype(np.float32) * 2 - 1) * np.asarray(half, np.float32) * 1.6
    c = random_coords(n, seed)
    c[:, :3] = (p - model.aabb_min) / (model.aabb_max - model.aabb_min)
    return c


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8).copy()


def build(model, occ):
    out = {}
    e3 = fx.e3(model)
    # host geometry of the three cages: deformed vertices and rotations stored, CSR tables hashed (42 MB each)
    for i, op in enumerate(e3):
        out[f"e3_{i}_vertices"] = op.vertices
        out[f"e3_{i}_rotations"] = op.rotations
        out[f"e3_{i}_csr_sha256"] = sha(op.lut_offsets, op.lut_idx)
        out[f"e3_{i}_n_idx"] = np.array([op.lut_idx.size], np.uint64)
    ops = [o.to_op() for o in e3]
    o = orc.Oracle(model.desc, model.params, occ, ops)
    c = np.concatenate([coords_around(model, (0.5, 0.62, 0.78), (0.2, 0.2, 0.2), 1500, 1), coords_around(model, (0.5, 0.5, 0.3), (0.3, 0.3, 0.4), 1500, 2)])
    mapped, mask = o.map_rays(c)
    sh, od, rd = o.poisson_residuals(c[:1500])
    out.update(map_in=c, map_out=mapped, map_mask=mask, res_sh=sh, res_od=od, res_rd=rd)
    for name, fixture, target in (("e1", fx.e1(model), False), ("e3", e3, False), ("e3_target", e3, True)):
        oo = orc.Oracle(model.desc, model.params, occ, [x.to_op() for x in fixture])
        f = syn.make_frame(model, syn.fox_camera0(), W, H, apply_operators=True, poisson_target=target)
        fb, depth, st, margin = oo.render(f, want_margin=True)
        out[f"{name}_rgba"], out[f"{name}_depth"], out[f"{name}_margin"] = fb, depth, margin
        out[f"{name}_stats"] = np.array([st.n_rays, st.n_rays_alive, st.n_hit, st.n_samples, st.n_old_samples], np.uint64)
        print(name, "samples", st.n_samples, "old", st.n_old_samples)
    # occupancy update through E3 (uniform pass over cascade 0 + a non-uniform pass), boundary values, blend
    rng = Pcg32(99)
    grid, bits, mean = o.update_density_grid(grid_params(128 ** 3, 0, rng, n_cascades=1), np.zeros(abi.NSB_GRID_CELLS, np.float32))
    rng.advance(); rng.advance()
    grid, bits, mean = o.update_density_grid(grid_params(100_000, 60_000, rng, ema_step=1, reset=False, n_cascades=3), grid)
    nz = np.nonzero(grid)[0]
    out.update(grid_touched=np.packbits(grid != 0), grid_sub_idx=nz[::97].astype(np.uint32), grid_sub_val=grid[nz[::97]], grid_bits=bits, grid_mean=np.array([mean], np.float32))
    cage = fx.e1(model)[0]
    d_in, s_in = o.poisson_boundary(cage.cage_original, sampling(seed=5, w=10, inside=True))
    d_out, s_out = o.poisson_boundary(cage.cage_vertices, sampling(seed=6, w=10, inside=False))
    b_shs, b_od, b_rd = orc.membrane_blend(cage.mvc, d_in, d_out, s_in, s_out)
    out.update(bnd_d_in=d_in, bnd_s_in=s_in, bnd_d_out=d_out, bnd_s_out=s_out, blend_shs=b_shs, blend_od=b_od, blend_rd=b_rd)
    return out


def main():
    model = syn.make_model(seed=1337)
    occ = syn.make_occupancy(model)
    out = build(model, occ)
    path = os.path.join(HERE, "synthetic_edits.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
