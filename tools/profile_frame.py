"""One 1080p frame (plus warm-up) of the bench workload, for ncu: python tools/profile_frame.py [n_frames] [width height]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nerfshop_b200 import synthetic as syn  # noqa: E402
from nerfshop_b200.renderer import NerfRenderer  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
W = int(sys.argv[2]) if len(sys.argv) > 3 else 1920
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
model = syn.make_model(seed=1337)
occ = syn.make_occupancy(model)
r = NerfRenderer(0)
r.upload_model(model.desc, model.params)
r.upload_occupancy(occ)
cams = syn.orbit_cameras(120)
fb = torch.zeros((H, W, 4), device="cuda")
dp = torch.zeros((H, W), device="cuda")
for i in range(n):
    fb.zero_()
    r.render(syn.make_frame(model, cams[(17 + 7 * i) % 120], W, H), fb, dp)
    st = r.stats()
    print(f"frame {i}: {st.gpu_ms:.3f} ms, {st.n_samples} samples, {W * H / st.gpu_ms / 1e3:.1f} Mrays/s")
r.close()
