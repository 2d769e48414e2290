"""Times nsb_update_density_grid for the reference's post-edit setting — update_density_grid_nerf_render
(testbed_nerf.cu:3514-3520): 128^3 * (max_cascade+1) uniform samples per iteration — with and without edit operators, and the
oracle on the host for the same call (bounded: 1/8 of the samples).
  python tools/profile_grid_update.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import edit_fixtures as fx  # noqa: E402
from nerfshop_b200 import abi, synthetic as syn  # noqa: E402
from nerfshop_b200.renderer import NerfRenderer  # noqa: E402
from nerfshop_b200.rng import Pcg32  # noqa: E402

model = syn.make_model(seed=1337)
occ = syn.make_occupancy(model)
r = NerfRenderer(0)
r.upload_model(model.desc, model.params)
r.upload_occupancy(occ)
N = 128 ** 3 * 3
for name, ops in (("no operators", []), ("3 cages + membrane (E3)", [o.to_op() for o in fx.e3(model)])):
    r.set_edit_operators(ops)
    rng = Pcg32(1337)
    ms = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.update_density_grid(rng, it, N, 0, reset_grid=(it == 0), n_cascades=3)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms[2:]))
    print(f"nsb_update_density_grid [{name}]: {N} samples, {t:.3f} ms = {N / t / 1e6:.2f} Gsamples/s "
          f"(hash lookup {N * 512 / t / 1e6:.0f} GB/s algorithmic); all iterations {['%.2f' % m for m in ms]}", flush=True)
if "--cpu" in sys.argv:
    from oracle import oracle as orc
    from test_grid_update_cpu import grid_params

    o = orc.Oracle(model.desc, model.params, occ, [o_.to_op() for o_ in fx.e3(model)])
    u = grid_params(N // 8, 0, Pcg32(1337))
    t0 = time.perf_counter()
    o.update_density_grid(u, np.zeros(abi.NSB_GRID_CELLS, np.float32))
    dt = time.perf_counter() - t0
    print(f"oracle (host, {orc.set_threads(0)} threads): {N // 8} samples in {dt:.2f} s = {N // 8 / dt / 1e6:.3f} Msamples/s")
r.close()
