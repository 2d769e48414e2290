"""Mints tests/golden/*.npz from the CPU oracle on the seeded synthetic scene.

The reference has no golden vectors, fixtures or tests for this path and cannot be built here
(SURVEY.md §0, §4, §8c), so these pins come from the restatement itself ("parity unpinned").
They freeze today's oracle so that any later change to it — or to the synthetic generator — is caught,
and they travel to the GPU box where the CUDA path is compared against them.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import random_coords  # noqa: E402
from nerfshop_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as orc  # noqa: E402

GOLDEN_W, GOLDEN_H = 96, 54
TRACE_PIXELS = np.array([y * GOLDEN_W + x for y in range(3, GOLDEN_H, 10) for x in range(5, GOLDEN_W, 12)], np.uint32)
TRACE_MAX = 96


def main():
    model = syn.make_model(seed=1337)
    occ = syn.make_occupancy(model)
    o = orc.Oracle(model.desc, model.params, occ)
    coords = random_coords(384, seed=7)
    enc = o.encode(coords)
    inf = o.inference(coords)
    den = o.inference(coords, density_only=True)
    out = dict(coords=coords, encode=enc, inference=inf, density=den)
    for name, cam in (("fox0", syn.fox_camera0()), ("orbit17", syn.orbit_cameras(120)[17])):
        f = syn.make_frame(model, cam, GOLDEN_W, GOLDEN_H)
        fb, depth, st, margin = o.render(f, want_margin=True)
        rec, idx, cnt = o.march_trace(f, TRACE_PIXELS, TRACE_MAX)
        out[f"{name}_rgba"] = fb
        out[f"{name}_depth"] = depth
        out[f"{name}_margin"] = margin
        out[f"{name}_stats"] = np.array([st.n_rays, st.n_rays_alive, st.n_hit, st.n_samples], np.uint64)
        out[f"{name}_trace_rec"] = rec
        out[f"{name}_trace_idx"] = idx
        out[f"{name}_trace_cnt"] = cnt
        print(name, "samples", st.n_samples, "hit", st.n_hit)
    out["params_checksum"] = np.array([int(model.params.astype(np.uint64).sum()), int(occ.astype(np.uint64).sum())], np.uint64)
    np.savez_compressed(os.path.join(HERE, "synthetic_fox.npz"), **out)
    print("wrote", os.path.join(HERE, "synthetic_fox.npz"))


if __name__ == "__main__":
    main()
