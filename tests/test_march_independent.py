"""CPU: an independent pure-Python/numpy restatement of ray generation + the occupancy march, written from SURVEY.md Appendix A
(formulas of init_rays_with_payload_kernel_nerf :2512-2616, advance_pos_nerf :557-606, common_nerf.cu:89-177, random_val.cuh:159-322),
against the C++ oracle's march_trace: every t, dt, position, cascade and Morton cell of the sampled rays must agree BIT FOR BIT.
float32 arithmetic is numpy float32 scalars; the FMAs the contract places (o + d*t, dot products, cell-scale) are emulated as
round32(float64 product + float64 addend) — exact products, one extra rounding that can differ from a true FMA only on 2^-29-rare ties."""
import math

import numpy as np

from nerfshop_b200 import synthetic as syn

F = np.float32
MIN_STEP = F(1.73205080757) / F(1024.0)
MAX_STEP = MIN_STEP * F(16.0) * F(1024.0) / F(128.0)
M32 = 0xFFFFFFFF


def fma(a, b, c):
    return F(np.float64(a) * np.float64(b) + np.float64(c))


def reverse_bits(x):
    return int(f"{x & M32:032b}"[::-1], 2)


def laine_karras(x, seed):
    x = (x + seed) & M32
    for k in (0x6C50B47C, 0xB82F1E52, 0xC7AFE638, 0x8D22F6E6):
        x ^= (x * k) & M32
    return x


def scramble(x, seed):
    return reverse_bits(laine_karras(reverse_bits(x), seed))


def ld_random_val(index, seed):  # Owen-scrambled Sobol, dimension 0 (= bit reversal)
    index = scramble(index, seed)
    sob = reverse_bits(index)
    hc = (seed ^ ((0 + ((seed << 6) & M32) + (seed >> 2)) & M32)) & M32
    return F(scramble(sob, hc)) * F(1.0 / 4294967296.0)


def calc_dt(t, cone):
    return min(max(t * cone, MIN_STEP), MAX_STEP)


def frexp_exp(v):
    return math.frexp(float(v))[1]


def mip_from_pos(p):
    m = max(abs(p[0] - F(0.5)), abs(p[1] - F(0.5)), abs(p[2] - F(0.5)))
    return min(4, max(0, frexp_exp(m) + 1))


def mip_from_dt(dt, p):
    mip = mip_from_pos(p)
    d = dt * F(256.0)
    return mip if d < F(1.0) else min(4, max(frexp_exp(d), mip))


def morton(x, y, z):
    r = 0
    for b in range(7):
        r |= ((x >> b) & 1) << (3 * b) | ((y >> b) & 1) << (3 * b + 1) | ((z >> b) & 1) << (3 * b + 2)
    return r


def cell_index(p, mip):
    s = F(math.ldexp(1.0, -mip))
    out = []
    for k in range(3):
        q = fma(p[k] - F(0.5), s, F(0.5))
        out.append(min(max(int(q * F(128.0)), 0), 127))  # int() truncates toward zero like the C cast
    return morton(*out)


def box_tmin(mn, mx, o, d):
    big = F(3.402823466e38)
    tmin, tmax = (mn[0] - o[0]) / d[0], (mx[0] - o[0]) / d[0]
    if tmin > tmax: tmin, tmax = tmax, tmin
    tymin, tymax = (mn[1] - o[1]) / d[1], (mx[1] - o[1]) / d[1]
    if tymin > tymax: tymin, tymax = tymax, tymin
    if tmin > tymax or tymin > tmax: return big
    if tymin > tmin: tmin = tymin
    if tymax < tmax: tmax = tymax
    tzmin, tzmax = (mn[2] - o[2]) / d[2], (mx[2] - o[2]) / d[2]
    if tzmin > tzmax: tzmin, tzmax = tzmax, tzmin
    if tmin > tzmax or tzmin > tmax: return big
    return max(tmin, tzmin)


def march(frame, bits, px, py, max_samples, want_ray=False):
    W, H = frame.width, frame.height
    cam = np.array(frame.camera1[:], np.float32)  # rolling shutter 0 -> camera1 exactly
    mn, mx = np.array(frame.render_aabb_min[:], np.float32), np.array(frame.render_aabb_max[:], np.float32)
    cone = F(frame.cone_angle_constant)
    uvx, uvy = (F(px) + F(0.5)) / F(W), (F(py) + F(0.5)) / F(H)   # spp 0: pixel centre
    dl = [(uvx - F(frame.screen_center[0])) * F(W) / F(frame.focal_length[0]), (uvy - F(frame.screen_center[1])) * F(H) / F(frame.focal_length[1]), F(1.0)]
    d = [fma(cam[r], dl[0], fma(cam[3 + r], dl[1], cam[6 + r] * dl[2])) for r in range(3)]  # Eigen tree m0*d0 + (m1*d1 + m2*d2), nvcc's contraction
    n = np.sqrt(fma(d[0], d[0], fma(d[1], d[1], d[2] * d[2])))
    d = [v / n for v in d]
    o = [cam[9], cam[10], cam[11]]
    t = max(box_tmin(mn, mx, o, d), F(0.05)) + F(1e-6)
    pos = [fma(d[k], t, o[k]) for k in range(3)]
    if not all(mn[k] <= pos[k] <= mx[k] for k in range(3)):
        return ([], o, d) if want_ray else []
    idir = [F(1.0) / v for v in d]
    pix = px + W * py
    t = fma(ld_random_val(frame.spp_index, (pix * 786433) & M32), calc_dt(t, cone), t)   # first-step jitter
    out = []
    for _ in range(100000):
        pos = [fma(d[k], t, o[k]) for k in range(3)]
        if not all(mn[k] <= pos[k] <= mx[k] for k in range(3)):
            break
        dt = calc_dt(t, cone)
        mip = max(frame.min_mip, mip_from_dt(dt, pos))
        cell = cell_index(pos, mip)
        if bits[mip * 128 ** 3 + cell]:
            out.append((t, dt, pos[0], pos[1], pos[2], mip, cell))
            if len(out) >= max_samples:
                break
            t = t + dt
            continue
        res = F(128 >> mip)
        tt = []
        for k in range(3):
            p = res * pos[k]
            sgn = F(math.copysign(1.0, float(d[k])))
            tt.append((np.floor(fma(F(0.5), sgn, p + F(0.5))) - p) * idir[k])
        t_target = t + max(min(tt) / res, F(0.0))
        while True:
            t = t + calc_dt(t, cone)
            if not (t < t_target):
                break
    return (out, o, d) if want_ray else out


def test_python_march_equals_the_oracle_bit_for_bit(scene, oracle):
    model, occ = scene
    bits = np.unpackbits(occ, bitorder="little")
    total = 0
    for cam, (W, H) in ((syn.fox_camera0(), (40, 22)), (syn.orbit_cameras(120)[77], (36, 20))):
        frame = syn.make_frame(model, cam, W, H)
        pix = np.arange(3, W * H, 41, dtype=np.uint32)
        rec, idx, cnt = oracle.march_trace(frame, pix, 48)
        for i, p in enumerate(pix):
            mine = march(frame, bits, int(p) % W, int(p) // W, 48)
            n = min(int(cnt[i]), 48)
            assert len(mine) == n, (int(p), len(mine), n)
            for k in range(n):
                got = np.array(mine[k][:5], np.float32)
                assert np.array_equal(got.view(np.uint32), rec[i, k].view(np.uint32)), (int(p), k, got, rec[i, k])
                assert (mine[k][5], mine[k][6]) == (int(idx[i, k, 0]), int(idx[i, k, 1]))
            total += n
    assert total > 500


def test_python_composite_equals_the_oracle_frame(scene, oracle):
    """The rest of the frame from the formulas (composite_kernel_nerf :750-955, compact :2503, shade :2464-2482): Python march above ->
    network outputs from the oracle's inference (cross-checked against numpy separately) -> float32 composite in Python, against
    oracle.render pixel by pixel. exp/pow come from numpy instead of libm: agreement to a few float32 ulps, not bits."""
    from nerfshop_b200 import abi

    model, occ = scene
    bits = np.unpackbits(occ, bitorder="little")
    W, H = 14, 8
    frame = syn.make_frame(model, syn.fox_camera0(), W, H)
    fb_o, depth_o, st_o, margin = oracle.render(frame, want_margin=True)
    tmin, tmax = np.array(frame.train_aabb_min[:], np.float32), np.array(frame.train_aabb_max[:], np.float32)
    diag = tmax - tmin
    cam = np.array(frame.camera1[:], np.float32)
    fwd, org = cam[6:9], cam[9:12]
    dt_range = MIN_STEP * F(16.0) - MIN_STEP
    sat = F(1.0) - F(frame.min_transmittance)
    n_samples = n_checked = 0
    for py in range(H):
        for px in range(W):
            samples, o, d = march(frame, bits, px, py, 4000, want_ray=True)
            c = np.zeros(4, np.float32)
            depth, max_w, done_sat = F(0.0), F(0.0), False
            if samples:
                coords = np.zeros((len(samples), 7), np.float32)
                for i, smp in enumerate(samples):
                    coords[i, :3] = [(smp[2 + k] - tmin[k]) / diag[k] for k in range(3)]
                    coords[i, 3] = (smp[1] - MIN_STEP) / dt_range
                    coords[i, 4:] = [(d[k] + F(1.0)) * F(0.5) for k in range(3)]
                raw = oracle.inference(coords).view(np.float16).astype(np.float32)  # [16, n]: rows 0-2 rgb, row 3 density
                for i in range(len(samples)):
                    n_samples += 1
                    cpos = [fma(coords[i, k], diag[k], tmin[k]) for k in range(3)]
                    T = F(1.0) - c[3]
                    dtu = fma(coords[i, 3], dt_range, MIN_STEP)
                    sigma = np.exp(raw[3, i])
                    alpha = F(1.0) - np.exp(-sigma * dtu)
                    w = alpha * T
                    for k in range(3):
                        c[k] = fma(F(1.0) / (F(1.0) + np.exp(-raw[k, i])), w, c[k])
                    c[3] = c[3] + w
                    if w > max_w:
                        max_w = w
                        q = [cpos[k] - org[k] for k in range(3)]
                        depth = fma(fwd[0], q[0], fma(fwd[1], q[1], fwd[2] * q[2]))
                    if c[3] > sat:
                        c = c / c[3]
                        done_sat = True
                        break
            pix_ambiguous = margin[py, px] <= 2e-5
            expect = np.zeros(4, np.float32)
            expect_depth = F(1e10)
            if c[3] > F(0.001):
                lin = [v / F(12.92) if v <= F(0.04045) else np.power((v + F(0.055)) / F(1.055), F(2.4)) for v in c[:3]]
                expect = np.array(lin + [c[3]], np.float32)
                if c[3] > F(0.2):
                    expect_depth = depth
            if not pix_ambiguous:
                assert np.allclose(expect, fb_o[py, px], atol=3e-6, rtol=1e-5), (px, py, expect, fb_o[py, px])
                assert np.isclose(expect_depth, depth_o[py, px], rtol=1e-5), (px, py, expect_depth, depth_o[py, px])
                n_checked += 1
    assert n_checked >= W * H - 2 and n_samples > 1000
    assert n_samples == st_o.n_samples  # same termination decisions on every ray
