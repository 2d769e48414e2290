"""GPU exploration: this repository's CUDA path vs the reference's own CUDA path (oracle/_ref/libnerfshop_ref_cuda.so) vs the CPU oracle.

Prints, for the occupancy march: how many samples differ (t stream, warped position, warped direction) between
  native (nsb_march_trace) / ref_cuda (the reference's kernels, nvcc) / oracle (CPU restatement)
and for a frame: L-inf RGBA between the three, and the frame time of ref_cuda vs native at 1080p."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from nerfshop_b200 import synthetic as syn
from nerfshop_b200.renderer import NerfRenderer
from oracle import oracle as orc
from oracle import ref

model = syn.make_model(seed=1337)
occ = syn.make_occupancy(model)
r = NerfRenderer(0)
r.upload_model(model.desc, model.params)
r.upload_occupancy(occ)
o = orc.Oracle(model.desc, model.params, occ)
rc = ref.RefCuda(occ)
cams = syn.orbit_cameras(120)

# ---- march ------------------------------------------------------------------------------------------------------------
W, H, MS = 192, 108, 64
for ci in (17, 63):
    f = syn.make_frame(model, cams[ci], W, H)
    rng = np.random.default_rng(ci)
    pix = rng.choice(W * H, 4000, replace=False).astype(np.uint32)
    rec_n, idx_n, cnt_n = r.march_trace(f, pix, MS)
    rec_o, idx_o, cnt_o = o.march_trace(f, pix, MS)
    rec_r, ray_r, cnt_r = rc.march_trace(f, pix, MS)
    amin = np.array(list(f.train_aabb_min), np.float32)
    amax = np.array(list(f.train_aabb_max), np.float32)
    for name, (rec, cnt) in {"native": (rec_n, cnt_n), "oracle": (rec_o, cnt_o)}.items():
        tot = bad_t = bad_p = bad_c = 0
        maxulp = 0
        for k in range(len(pix)):
            c = int(min(cnt[k], MS))
            if c != int(cnt_r[k]):
                bad_c += 1
                c = min(c, int(cnt_r[k]))
            if c == 0:
                continue
            t, dt, pos = rec[k, :c, 0], rec[k, :c, 1], rec[k, :c, 2:5]
            ta = (t + dt).astype(np.float32)
            wp = ((pos - amin) / (amax - amin)).astype(np.float32)
            tot += c
            bad_t += int((ta != rec_r[k, :c, 7]).sum())
            d = np.abs(wp.view(np.int32).astype(np.int64) - rec_r[k, :c, 0:3].view(np.int32).astype(np.int64))
            bad_p += int((d > 0).any(axis=1).sum())
            maxulp = max(maxulp, int(d.max()))
        print(f"cam {ci} march {name} vs ref_cuda: {tot} samples, count mismatches {bad_c}, t mismatches {bad_t}, pos mismatches {bad_p} (max {maxulp} ulp)")
    print("  native vs oracle bit-equal:", np.array_equal(rec_n, rec_o), np.array_equal(cnt_n, cnt_o))

# ---- frames -------------------------------------------------------------------------------------------------------------
f = syn.make_frame(model, cams[17], W, H)
fb_n, d_n = r.render(f)
fb_n = fb_n.cpu().numpy()
st = r.stats()
fb_r, d_r, info = rc.render(f, r)
torch.cuda.synchronize()
fb_r = fb_r.cpu().numpy()
fb_o, d_o, st_o, margin = o.render(f, want_margin=True)
print(f"frame {W}x{H}: native samples {st.n_samples}, oracle {st_o.n_samples}, ref_cuda inferred {info['n_inferred']} in {info['n_calls']} calls")
print("  L-inf native-ref_cuda %.3e  oracle-ref_cuda %.3e  native-oracle %.3e" % (np.abs(fb_n - fb_r).max(), np.abs(fb_o - fb_r).max(), np.abs(fb_n - fb_o).max()))
e = np.abs(fb_n - fb_r).max(-1)
print("  native-ref_cuda: pixels > 1e-4: %d, > 1e-3: %d of %d" % ((e > 1e-4).sum(), (e > 1e-3).sum(), e.size))

# ---- 1080p timing -----------------------------------------------------------------------------------------------------------
for (w, h) in ((1920, 1080),):
    fr = [syn.make_frame(model, cams[i], w, h) for i in (0, 30, 60, 90)]
    fb = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    dp = torch.zeros((h, w), dtype=torch.float32, device="cuda")
    for arm in ("native", "ref_cuda"):
        ms = []
        for it in range(3 + len(fr)):
            f = fr[it % len(fr)]
            fb.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if arm == "native":
                r.render(f, fb, dp)
            else:
                rc.render(f, r, fb, dp)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ms.append(e0.elapsed_time(e1))
        print(f"{arm} {w}x{h}: {np.mean(ms):.2f} ms/frame  ({w * h / np.mean(ms) / 1e3:.1f} Mrays/s)  per-frame {['%.1f' % m for m in ms]}")
    fb_nat = torch.zeros_like(fb)
    r.render(fr[0], fb_nat, dp)
    fb_ref = torch.zeros_like(fb)
    rc.render(fr[0], r, fb_ref, dp)
    torch.cuda.synchronize()
    e = (fb_nat - fb_ref).abs().amax(-1)
    print("1080p native vs ref_cuda L-inf %.3e, pixels > 1e-3: %d" % (e.max().item(), (e > 1e-3).sum().item()))
