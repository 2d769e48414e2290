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

# ---- edits: map_rays / poisson through E3 + affine ------------------------------------------------------------------
from edit_fixtures import e3
from nerfshop_b200 import editing
a_ = 0.4
ROT = np.array([[np.cos(a_), -np.sin(a_), 0], [np.sin(a_), np.cos(a_), 0], [0, 0, 1]], np.float32)
ops = [c.to_op() for c in e3(model)]
ops.append(editing.AffineDuplication((0.5, 0.5, 0.5), (0.12, 0.12, 0.12), (0.03, 0.0, -0.1), rotation=ROT, hide_original=True, correct_dir=True).to_op())
r.set_edit_operators(ops)
rce = ref.RefCuda(occ, ops)
oe = orc.Oracle(model.desc, model.params, occ, ops)
rng = np.random.default_rng(3)
n = 400_000
c = np.zeros((n, 7), np.float32)
c[:, :3] = rng.uniform(0.38, 0.64, (n, 3))
d = rng.standard_normal((n, 3)).astype(np.float32)
c[:, 4:] = (d / np.linalg.norm(d, axis=1, keepdims=True) + 1) * 0.5
cn, mn_ = r.map_rays(c)
cr_, mr_ = rce.map_rays(c)
co_, mo_ = oe.map_rays(c)
print("map_rays native vs oracle bit-equal:", np.array_equal(cn, co_), np.array_equal(mn_, mo_))
moved = (cr_[:, :3] != c[:, :3]).any(axis=1)
print(f"map_rays native vs ref_cuda: moved {moved.sum()}, masked {mr_.sum()}; mask mismatches {(mn_ != mr_).sum()}; rows differing pos {(cn[:, :3] != cr_[:, :3]).any(axis=1).sum()} dir {(cn[:, 4:] != cr_[:, 4:]).any(axis=1).sum()}; max abs pos {np.abs(cn[:, :3] - cr_[:, :3]).max():.3e} dir {np.abs(cn[:, 4:] - cr_[:, 4:]).max():.3e}")
shn, odn, rdn = r.poisson_residuals(c)
shr, odr, rdr = rce.poisson_residuals(c)
print(f"poisson native vs ref_cuda: inside {(odr != 0).sum()}; membership mismatches {((odn != 0) != (odr != 0)).sum()}; od differ {(odn != odr).sum()} max {np.abs(odn - odr).max():.3e}; sh differ {(shn != shr).any(axis=1).sum()} max {np.abs(shn - shr).max():.3e}")
fe = syn.make_frame(model, cams[17], W, H)
fe.apply_operators, fe.poisson_target = 1, 1
fbn, _ = r.render(fe)
fbr, _, _ = rce.render(fe, r)
torch.cuda.synchronize()
e = (fbn - fbr).abs().amax(-1)
print("edited frame native vs ref_cuda L-inf %.3e, > 1e-4: %d, > 1e-3: %d" % (e.max().item(), (e > 1e-4).sum().item(), (e > 1e-3).sum().item()))
r.set_edit_operators([])
rce.close()

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
