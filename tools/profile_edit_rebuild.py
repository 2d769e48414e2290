"""Latency of one cage drag: the reference's flow — rebuild on the host (MVC, rotations, tet grid) and re-upload
the operator (42 MB of CSR offsets) — against nsb_cage_deform, which rebuilds in place on the device.
  python tools/profile_edit_rebuild.py [lattice_n]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import edit_fixtures as fx  # noqa: E402
from nerfshop_b200 import synthetic as syn  # noqa: E402
from nerfshop_b200.renderer import NerfRenderer  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = syn.make_model(seed=1337)
r = NerfRenderer(0)
r.upload_model(model.desc, model.params)
r.upload_occupancy(syn.make_occupancy(model))
op = fx.make_cage(model, (0.5, 0.62, 0.78), (0.17, 0.17, 0.17), n_lattice=N)
print(f"cage: {op.cage_original.shape[0]} cage vertices, {op.vertices.shape[0]} tet vertices, {op.tets.shape[0]} tets, {op.lut_idx.size} CSR entries", flush=True)
r.set_edit_operators([op.to_op()])
r.cage_attach_mvc(0, op.mvc)
def dragged(k):
    moved = op.cage_original.copy()
    moved[moved[:, 0] > op.cage_original[:, 0].mean()] += np.float32(0.01 * (k + 1))
    return moved


host_ms, dev_ms = [], []
for k in range(5):  # the reference's flow: rebuild on the host, upload the operator again
    op.cage_vertices = dragged(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    op.update_tet_mesh()
    t1 = time.perf_counter()
    r.set_edit_operators([op.to_op()])
    torch.cuda.synchronize()
    host_ms.append(((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
r.cage_attach_mvc(0, op.mvc)
for k in range(6):  # the device flow: the operator stays resident
    moved = dragged(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r.cage_deform(0, moved)
    torch.cuda.synchronize()
    dev_ms.append((time.perf_counter() - t0) * 1e3)
op.cage_vertices = moved
op.update_tet_mesh()
got = r.cage_download(0, op.vertices.shape[0], op.tets.shape[0])
assert np.array_equal(got["lut_offsets"], op.lut_offsets) and np.array_equal(got["lut_idx"], op.lut_idx)
h = np.median(np.array(host_ms)[1:], axis=0)
print(f"host rebuild (nsb_host_geometry.cpp, 1 thread) {h[0]:.1f} ms + operator re-upload {h[1]:.1f} ms = {h.sum():.1f} ms per drag")
print(f"nsb_cage_deform (device, wall clock incl. its two stream syncs) {np.median(dev_ms[1:]):.2f} ms per drag; identical CSR")
r.close()
