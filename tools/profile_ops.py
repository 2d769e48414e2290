"""The hash lookup and the fused MLPs ALONE, on the sample positions of a real frame (not synthetic coordinates):
  python tools/profile_ops.py [width height max_samples]
Prints achieved algorithmic GB/s of nsb_encode (512 B gathered per sample) and TFLOP/s of the MLPs inside nsb_inference,
for two sample orders: ray-major (one ray's consecutive samples adjacent) and step-major (neighbouring pixels at the same
step adjacent = what a warp of the fused kernel sees). Run under ncu with -k regex:k_encode / -k regex:k_inference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nerfshop_b200 import abi, synthetic as syn  # noqa: E402
from nerfshop_b200.renderer import NerfRenderer  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 2 else 960
H = int(sys.argv[2]) if len(sys.argv) > 2 else 540
MAXS = int(sys.argv[3]) if len(sys.argv) > 3 else 32

model = syn.make_model(seed=1337)
occ = syn.make_occupancy(model)
r = NerfRenderer(0)
r.upload_model(model.desc, model.params)
r.upload_occupancy(occ)
f = syn.make_frame(model, syn.orbit_cameras(120)[17], W, H)
# pixels in the renderer's tile order (16x8 tiles, 8x4 blocks)
tx, ty = (W + 15) // 16, (H + 7) // 8
order = []
for t in range(tx * ty):
    for s in range(128):
        b, l = s >> 5, s & 31
        px, py = (t % tx) * 16 + (b & 1) * 8 + (l & 7), (t // tx) * 8 + (b >> 1) * 4 + (l >> 3)
        if px < W and py < H:
            order.append(py * W + px)
pix = np.array(order, np.uint32)
rec, idx, cnt = r.march_trace(f, pix, MAXS)
n_rays = pix.size
cnt = np.minimum(cnt, MAXS)
pos = rec[:, :, 2:5]
warped = (pos - model.aabb_min) / (model.aabb_max - model.aabb_min)
valid = np.arange(MAXS)[None, :] < cnt[:, None]
ray_major = warped[valid]
step_major = np.concatenate([warped[:, k][valid[:, k]] for k in range(MAXS)])
print(f"{n_rays} rays, {ray_major.shape[0]} samples ({ray_major.shape[0] / n_rays:.1f} per ray, first {MAXS} of each)", flush=True)

peaks = abi  # noqa
for name, p in (("ray-major", ray_major), ("step-major", step_major)):
    n = p.shape[0]
    c = np.zeros((n, 7), np.float32)
    c[:, :3] = p
    c[:, 4:] = 0.5
    ct = torch.from_numpy(c).cuda()
    n_pad = ((n + 127) // 128) * 128
    enc = torch.zeros((32, n_pad), dtype=torch.int16, device="cuda")
    out = torch.zeros((16, n_pad), dtype=torch.int16, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for fn, nm, dst in ((r.lib.nsb_encode, "encode", enc), (r.lib.nsb_inference, "inference", out)):
        ms = []
        for rep in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            abi.check(r.lib, fn(r.ctx, ct.data_ptr(), n, dst.data_ptr(), n_pad, torch.cuda.current_stream().cuda_stream), nm)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = float(np.median(ms[1:]))
        print(f"{nm:9s} [{name:10s}] {n} samples: {t:.3f} ms = {n / t / 1e6:.2f} Gsamples/s; hash lookup {n * 512 / t / 1e6:.0f} GB/s algorithmic "
              f"({n * 512 / t / 1e6 / 6576.7 * 100:.0f} % of measured HBM peak); MLP {n * 20480 / t / 1e9:.1f} TFLOP/s", flush=True)
r.close()
