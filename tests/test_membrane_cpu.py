"""§8(f)-4, host side: properties of the oracle's compute_poisson_boundary / interpolate_poisson_boundary restatements
(growing_selection.cu:2220-2398)."""
import numpy as np

import edit_fixtures as fx
from nerfshop_b200 import abi
from oracle import oracle as orc


def sampling(seed=11, w=6, inside=False):
    p = abi.NsbBoundarySampling()
    p.sampling_width, p.hemisphere_width, p.seed = w, w, seed
    p.train_aabb_min[:] = (-1.5, -1.5, -1.5)
    p.train_aabb_max[:] = (2.5, 2.5, 2.5)
    p.rgb_activation, p.density_activation, p.is_inside = abi.NSB_ACT_LOGISTIC, abi.NSB_ACT_EXPONENTIAL, int(inside)
    return p


def test_boundary_sampling_properties(scene, oracle):
    model, occ = scene
    op = fx.e1(model)[0]
    pts = op.cage_vertices
    p = sampling()
    dens, shs, coords = oracle.poisson_boundary(pts, p, want_coords=True)
    w = p.sampling_width
    d = coords[:, 4:] * 2 - 1
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
    # stratified: u = (1 + cos(phi)) / 2 = (1 + z) / 2 lies in cell i, theta / 2pi in cell j
    c = coords.reshape(pts.shape[0], w, w, 7)
    u = (1 + (c[..., 6] * 2 - 1)) / 2
    assert (np.floor(u * w + 1e-4).astype(int) == np.arange(w)[None, :, None]).mean() > 0.97
    assert np.allclose(c[..., :3], ((pts - model.aabb_min) / (model.aabb_max - model.aabb_min))[:, None, None, :])
    # DC term = 4pi/n * sum(rgb) * Y00 with rgb in (0,1)
    assert (shs[:, [0, 9, 18]] > 0).all() and (shs[:, [0, 9, 18]] < 4 * np.pi * 0.282095 + 1e-4).all()
    assert (dens > 0).all() and np.isfinite(shs).all()
    # determinism and seed dependence
    d2, s2 = oracle.poisson_boundary(pts, p)
    assert np.array_equal(d2, dens) and np.array_equal(s2, shs)
    d3, s3 = oracle.poisson_boundary(pts, sampling(seed=12))
    assert not np.array_equal(s3, shs) and np.allclose(s3[:, [0, 9, 18]], shs[:, [0, 9, 18]], atol=0.3)
    # is_inside: density is zeroed exactly at points whose occupancy cell is empty
    far = np.array([[2.3, 2.3, 2.3], [0.5, 0.62, 0.78]], np.float32)
    di, _ = oracle.poisson_boundary(far, sampling(inside=True))
    do, _ = oracle.poisson_boundary(far, sampling(inside=False))
    assert di[0] == 0.0 and do[0] > 0.0 and di[1] == do[1]


def test_membrane_blend_against_float64():
    rng = np.random.default_rng(3)
    nv, ncv = 57, 13
    gamma = rng.random((nv, ncv)).astype(np.float32)
    gamma /= gamma.sum(1, keepdims=True)
    d_in, d_out = rng.uniform(0, 40, ncv).astype(np.float32), rng.uniform(1, 60, ncv).astype(np.float32)
    s_in, s_out = rng.normal(0, 0.5, (ncv, 27)).astype(np.float32), rng.normal(0, 0.5, (ncv, 27)).astype(np.float32)
    b_shs, b_od, b_rd = orc.membrane_blend(gamma, d_in, d_out, s_in, s_out)
    ms = np.sqrt(3) / 1024
    a_out, a_in = 1 - np.exp(-d_out.astype(np.float64) * ms), 1 - np.exp(-d_in.astype(np.float64) * ms)
    w_in = np.minimum(a_in / a_out, 1.0)
    diff = s_out - w_in[:, None] * s_in
    ga = gamma * a_out[None, :]
    ref = (ga @ diff) / (ga.sum(1, keepdims=True) + 1e-6)
    assert np.allclose(b_shs, ref, atol=2e-5)
    assert np.allclose(b_od, gamma @ d_out, rtol=1e-5)
    assert np.allclose(b_rd, np.maximum(gamma @ (d_out - d_in), 0), atol=1e-4) and (b_rd >= 0).all()
    # equal inside/outside values: no colour residual, no density residual
    z_shs, _, z_rd = orc.membrane_blend(gamma, d_out, d_out, s_out, s_out)
    assert np.abs(z_shs).max() < 1e-6 and (z_rd == 0).all()
