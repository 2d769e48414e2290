"""§8(f)-4 frame post-process: accumulate (running mean over spp) and tonemap (background blend, exposure, curve, output space)."""
import numpy as np
import pytest

from nerfshop_b200 import abi
from oracle import oracle as orc


def _params(cs=0, out=1, curve=0, clamp=0, exposure=0.0, bg=(0.2, 0.4, 0.6, 1.0)):
    p = abi.NsbTonemap()
    p.color_space, p.output_color_space, p.tonemap_curve, p.clamp_output_color, p.exposure = cs, out, curve, clamp, exposure
    for i in range(4):
        p.background_color[i] = bg[i]
    return p


def _frames(n, seed=0):
    rng = np.random.default_rng(seed)
    f = rng.random((n, 24, 40, 4)).astype(np.float32)
    f[..., :3] *= f[..., 3:4]  # premultiplied
    return f


def test_accumulate_is_a_running_mean():
    fr = _frames(5)
    acc = None
    for k in range(5):
        acc = orc.accumulate(fr[k], acc, k)
    assert np.allclose(acc, fr.mean(0), atol=1e-6)
    srgb = orc.accumulate(fr[0], None, 0, abi.NSB_COLOR_SRGB)
    lin = fr[0][..., :3]
    ref = np.where(lin < 0.0031308, 12.92 * lin, 1.055 * np.power(lin, 0.41666) - 0.055)
    assert np.allclose(srgb[..., :3], ref, atol=1e-6) and np.allclose(srgb[..., 3], fr[0][..., 3])


def test_tonemap_identity_and_background():
    fr = _frames(1)[0]
    out = orc.tonemap(fr, _params(out=0, bg=(0, 0, 0, 0)))
    assert np.allclose(out, fr)                                   # linear in, linear out, no background: identity
    out = orc.tonemap(fr, _params(out=0, bg=(1.0, 1.0, 1.0, 1.0)))
    assert np.allclose(out[..., 3], 1.0, atol=1e-6)               # opaque background completes alpha
    assert np.allclose(out[..., :3], fr[..., :3] + (1 - fr[..., 3:4]), atol=1e-6)
    brighter = orc.tonemap(fr, _params(out=0, exposure=1.0, bg=(0, 0, 0, 0)))
    assert np.allclose(brighter[..., :3], 2 * fr[..., :3], rtol=1e-6)
    for curve in (abi.NSB_TONEMAP_ACES, abi.NSB_TONEMAP_HABLE, abi.NSB_TONEMAP_REINHARD):
        o = orc.tonemap(fr, _params(curve=curve, out=1, clamp=1))
        assert np.isfinite(o).all() and o.min() >= 0 and o.max() <= 1


@pytest.mark.gpu
def test_postprocess_gpu_matches_oracle(renderer):
    import torch

    fr = _frames(4, seed=3)
    acc_o, acc_g = None, torch.zeros((24, 40, 4), device="cuda")
    for cs in (abi.NSB_COLOR_LINEAR, abi.NSB_COLOR_SRGB, abi.NSB_COLOR_VISPOSNEG):
        for k in range(4):
            acc_o = orc.accumulate(fr[k], acc_o, k, cs)
            renderer.accumulate(torch.from_numpy(fr[k]).cuda(), acc_g, k, cs)
        assert np.allclose(acc_g.cpu().numpy(), acc_o, rtol=2e-6, atol=1e-7)
    for cs in (0, 1):
        for out in (0, 1):
            for curve in range(4):
                p = _params(cs, out, curve, clamp=curve % 2, exposure=0.5)
                ref = orc.tonemap(acc_o, p)
                got = renderer.tonemap(torch.from_numpy(acc_o).cuda(), p).cpu().numpy()
                assert np.allclose(got, ref, rtol=1e-5, atol=1e-6), (cs, out, curve)
