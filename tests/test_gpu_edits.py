"""GPU parity of the edit path (BASELINE configs 2 and 3): map_rays, membrane residuals and full frames with
cage operators, CUDA vs oracle. map_rays/residuals use only pinned fp32 ops -> bit-exact."""
import numpy as np
import pytest

import edit_fixtures as fx
from conftest import random_coords
from nerfshop_b200 import editing, synthetic as syn
from oracle import oracle as orc
from test_gpu_parity import _compare_frames

pytestmark = pytest.mark.gpu


def _coords_around(model, center, half, n, seed):
    rng = np.random.default_rng(seed)
    p = np.asarray(center, np.float32) + (rng.random((n, 3)).astype(np.float32) * 2 - 1) * np.asarray(half, np.float32) * 1.6
    c = random_coords(n, seed)
    c[:, :3] = (p - model.aabb_min) / (model.aabb_max - model.aabb_min)
    return c


@pytest.fixture()
def edited(scene, renderer):
    model, occ = scene
    yield model, occ
    renderer.reset_edit_operators()


def test_map_rays_bit_exact(edited, renderer):
    model, occ = edited
    ops = [o.to_op() for o in fx.e3(model)] + [editing.AffineDuplication((0.6, 0.55, 0.5), (0.05, 0.05, 0.05), (0.04, 0.02, 0.0), correct_dir=True,
                                                                           rotation=[[0.8, -0.6, 0], [0.6, 0.8, 0], [0, 0, 1]]).to_op()]
    o = orc.Oracle(model.desc, model.params, occ, ops)
    renderer.set_edit_operators(ops)
    c = np.concatenate([_coords_around(model, (0.5, 0.62, 0.78), (0.2, 0.2, 0.2), 20000, 1), _coords_around(model, (0.5, 0.5, 0.3), (0.3, 0.3, 0.4), 20001, 2)])
    ref, mref = o.map_rays(c)
    got, mgot = renderer.map_rays(c)
    moved = (ref[:, :3] != c[:, :3]).any(1)
    assert moved.sum() > 2000 and mref.sum() > 100
    assert np.array_equal(mref, mgot)
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32)), f"{(ref != got).any(1).sum()} samples differ"


def test_poisson_residuals_bit_exact(edited, renderer):
    model, occ = edited
    ops = [o.to_op() for o in fx.e3(model)]
    o = orc.Oracle(model.desc, model.params, occ, ops)
    renderer.set_edit_operators(ops)
    c = _coords_around(model, (0.55, 0.62, 0.78), (0.2, 0.2, 0.2), 10007, 3)
    sh_o, od_o, rd_o = o.poisson_residuals(c)
    sh_g, od_g, rd_g = renderer.poisson_residuals(c)
    assert (od_o > 1e-9).sum() > 1000
    assert np.array_equal(od_o.view(np.uint32), od_g.view(np.uint32)) and np.array_equal(rd_o.view(np.uint32), rd_g.view(np.uint32))
    assert np.array_equal(sh_o.view(np.uint32), sh_g.view(np.uint32))


@pytest.mark.parametrize("fixture,poisson_target", [("e1", False), ("e3", False), ("e3", True)])
def test_render_with_cage_edits(edited, renderer, fixture, poisson_target):
    """configs[2] (one cage) and configs[3] (three concurrent cages + membrane interpolation)."""
    model, occ = edited
    ops = [o.to_op() for o in getattr(fx, fixture)(model)]
    o = orc.Oracle(model.desc, model.params, occ, ops)
    renderer.set_edit_operators(ops)
    for cam in (syn.orbit_cameras(120)[17], syn.fox_camera0()):
        f = syn.make_frame(model, cam, 144, 81, apply_operators=True, poisson_target=poisson_target)
        fb_o, depth_o, st_o, margin = o.render(f, want_margin=True)
        fb, depth = renderer.render(f)
        st = renderer.stats()
        _compare_frames(fb.cpu().numpy(), depth.cpu().numpy(), fb_o, depth_o, margin)
        assert abs(int(st.n_samples) - int(st_o.n_samples)) <= max(20, st_o.n_samples // 2000)
        if poisson_target:
            assert st.n_old_samples == st_o.n_old_samples and st.n_old_samples > 0
        else:
            assert st.n_old_samples == 0
    # operators uploaded but apply_operators = 0 -> the unedited frame
    f0 = syn.make_frame(model, syn.orbit_cameras(120)[17], 144, 81)
    a, _ = renderer.render(f0)
    renderer.reset_edit_operators()
    b, _ = renderer.render(f0)
    assert (a == b).all()


def test_edited_frames_match_golden(edited, renderer):
    """The committed pins of configs 2-3 (tests/golden/synthetic_edits.npz): one cage, three cages + membrane, and the
    poisson-target blend, 96x54."""
    import os

    model, occ = edited
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synthetic_edits.npz"))
    for name, fixture, target in (("e1", fx.e1(model), False), ("e3", fx.e3(model), False), ("e3_target", fx.e3(model), True)):
        renderer.set_edit_operators([o.to_op() for o in fixture])
        f = syn.make_frame(model, syn.fox_camera0(), 96, 54, apply_operators=True, poisson_target=target)
        fb, depth = renderer.render(f)
        st = renderer.stats()
        _compare_frames(fb.cpu().numpy(), depth.cpu().numpy(), g[f"{name}_rgba"], g[f"{name}_depth"], g[f"{name}_margin"])
        assert [st.n_rays, st.n_rays_alive, st.n_samples] == [int(g[f"{name}_stats"][k]) for k in (0, 1, 3)]
    ops = [o.to_op() for o in fx.e3(model)]
    renderer.set_edit_operators(ops)
    got, mask = renderer.map_rays(g["map_in"])
    assert np.array_equal(got.view(np.uint32), g["map_out"].view(np.uint32)) and np.array_equal(mask, g["map_mask"])
