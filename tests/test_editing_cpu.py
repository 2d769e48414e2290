"""CPU tests of the edit path: host geometry (MVC, tet lookup table, rotations) and the oracle's map_rays /
membrane residuals through invariants (identity cage, translated cage, sum of barycentrics)."""
import numpy as np

import edit_fixtures as fx
from nerfshop_b200 import abi, editing, synthetic as syn
from oracle import oracle as orc


def _coords_in_box(model, lo, hi, n, seed):
    rng = np.random.default_rng(seed)
    p = lo + rng.random((n, 3)).astype(np.float32) * (np.asarray(hi, np.float32) - np.asarray(lo, np.float32))
    c = np.zeros((n, 7), np.float32)
    c[:, :3] = (p - model.aabb_min) / (model.aabb_max - model.aabb_min)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    c[:, 4:] = (d / np.linalg.norm(d, axis=1, keepdims=True) + 1) * 0.5
    return c, p


def test_mvc_partition_of_unity_and_linear_precision(built_lib):
    lib = abi.load_library()
    cv, ct = editing.box_cage((0.2, 0.3, 0.1), (0.7, 0.9, 0.8))
    rng = np.random.default_rng(0)
    pts = (np.array([0.2, 0.3, 0.1]) + rng.random((200, 3)) * np.array([0.5, 0.6, 0.7])).astype(np.float32)
    w = np.zeros((200, 8), np.float32)
    assert lib.nsb_compute_mvc(cv.ctypes.data, 8, ct.ctypes.data, 12, pts.ctypes.data, 200, 0.0, w.ctypes.data) == 0
    assert np.allclose(w.sum(1), 1.0, atol=1e-5)
    assert np.allclose(w @ cv, pts, atol=1e-5)          # mean value coordinates reproduce linear functions
    assert (w > -1e-6).all()                               # convex cage, interior points: non-negative
    one = np.zeros((1, 8), np.float32)
    v3 = np.ascontiguousarray(cv[3:4])
    assert lib.nsb_compute_mvc(cv.ctypes.data, 8, ct.ctypes.data, 12, v3.ctypes.data, 1, 0.0, one.ctypes.data) == 0
    assert np.allclose(one[0], np.eye(8)[3])               # on a cage vertex -> one-hot
    out = np.zeros_like(pts)
    assert lib.nsb_interpolate_with_mvc(w.ctypes.data, 200, 8, cv.ctypes.data, out.ctypes.data) == 0
    assert np.allclose(out, pts, atol=1e-5)


def test_local_rotations_recover_a_rigid_rotation(built_lib):
    lib = abi.load_library()
    tv, tets = editing.lattice_tets((0, 0, 0), (1, 1, 1), 2)
    a = 0.4
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
    rotated = np.ascontiguousarray((tv @ R.T + np.array([0.3, -0.1, 0.2], np.float32)).astype(np.float32))
    rots = np.zeros((tets.shape[0], 9), np.float32)
    tets = np.ascontiguousarray(tets.reshape(-1))
    assert lib.nsb_local_rotations(rotated.ctypes.data, tv.ctypes.data, tets.ctypes.data, rots.shape[0], rots.ctypes.data) == 0
    # R_tet maps deformed-space directions to canonical space: dir' = R_tet dir = R^T dir  (column-major storage)
    for r in rots:
        assert np.allclose(r.reshape(3, 3).T, R.T, atol=1e-5)


def test_tet_grid_lists_the_containing_tet(scene, built_lib):
    model, _ = scene
    op = fx.e1(model)[0]
    rng = np.random.default_rng(3)
    # random points inside random tets of the DEFORMED mesh
    t = rng.integers(0, op.tets.shape[0], 400)
    b = rng.dirichlet(np.ones(4), 400).astype(np.float32)
    pts = np.einsum("nk,nkd->nd", b, op.vertices[op.tets[t]])
    o = orc.lib()
    for p, ti in zip(pts, t):
        level = o.orc_mip_from_pos(float(p[0]), float(p[1]), float(p[2]))
        cell = level * 128 ** 3 + o.orc_cascaded_grid_idx_at(float(p[0]), float(p[1]), float(p[2]), level)
        cand = op.lut_idx[op.lut_offsets[cell]:op.lut_offsets[cell + 1]]
        assert ti in cand
        assert (np.diff(cand.astype(np.int64)) > 0).all()  # ascending tet ids, like the reference's merged lists
    assert op.lut_offsets[-1] == op.lut_idx.size
    assert op.original_bitfield.any()


def test_identity_cage_is_identity(scene):
    model, _ = scene
    cv, ct = editing.box_cage((0.3, 0.4, 0.5), (0.7, 0.8, 0.95))
    tv, tets = editing.lattice_tets((0.31, 0.41, 0.51), (0.69, 0.79, 0.94), 3)
    op = editing.CageDeformation(model.aabb_min, model.aabb_max, cv, ct, tv, tets)
    o = orc.Oracle(model.desc, model.params, None, [op.to_op()])
    c, _ = _coords_in_box(model, (0.25, 0.35, 0.45), (0.75, 0.85, 1.0), 3000, 1)
    out, mask = o.map_rays(c)
    assert np.abs(out[:, :3] - c[:, :3]).max() < 2e-6   # barycentric round trip through identical vertices
    assert np.abs(out[:, 4:] - c[:, 4:]).max() < 2e-6   # per-tet rotation of an undeformed mesh is the identity
    assert mask.sum() == 0                              # nothing is vacated


def test_translated_cage_moves_and_vacates(scene):
    model, _ = scene
    lo, hi = np.array([0.3, 0.4, 0.5], np.float32), np.array([0.6, 0.7, 0.8], np.float32)
    delta = np.array([0.5, 0.0, 0.0], np.float32)  # disjoint from the original box
    cv, ct = editing.box_cage(lo, hi)
    tv, tets = editing.lattice_tets(lo + 0.005, hi - 0.005, 3)
    op = editing.CageDeformation(model.aabb_min, model.aabb_max, cv, ct, tv, tets)
    op.cage_vertices = op.cage_original + delta
    op.update_tet_mesh()
    o = orc.Oracle(model.desc, model.params, None, [op.to_op()])
    diag = model.aabb_max - model.aabb_min
    c_in, p_in = _coords_in_box(model, lo + delta + 0.02, hi + delta - 0.02, 2000, 2)   # inside the moved cage
    out, mask = o.map_rays(c_in)
    assert np.abs((out[:, :3] * diag + model.aabb_min) - (p_in - delta)).max() < 5e-6   # backward map: deformed -> canonical
    assert mask.sum() == 0
    c_old, _ = _coords_in_box(model, lo + 0.02, hi - 0.02, 2000, 3)                     # the vacated canonical region
    out2, mask2 = o.map_rays(c_old)
    assert np.array_equal(out2, c_old) and mask2.all()                                   # masked empty, not moved
    op.copy = True                                                                       # copy mode keeps the original
    o.set_ops([op.to_op()])
    _, mask3 = o.map_rays(c_old)
    assert mask3.sum() == 0


def test_operators_apply_in_reverse_order(scene):
    model, _ = scene
    a = editing.AffineDuplication((0.5, 0.5, 0.5), (0.1, 0.1, 0.1), (0.2, 0.0, 0.0), hide_original=False)
    b = editing.AffineDuplication((0.7, 0.5, 0.5), (0.1, 0.1, 0.1), (0.0, 0.2, 0.0), hide_original=False)
    c = np.zeros((1, 7), np.float32)
    c[0, :3] = (0.7, 0.7, 0.5)   # in b's destination -> back to (0.7,0.5,0.5) = a's destination -> back to (0.5,0.5,0.5)
    c[0, 4:] = 0.5
    o = orc.Oracle(model.desc, model.params, None, [a.to_op(), b.to_op()])   # list order a, b: b is applied first
    out, _ = o.map_rays(c)
    assert np.allclose(out[0, :3], (0.5, 0.5, 0.5), atol=1e-6)
    o.set_ops([b.to_op(), a.to_op()])                                          # a first: (0.7,0.7) is not in a's box; then b
    out, _ = o.map_rays(c)
    assert np.allclose(out[0, :3], (0.7, 0.5, 0.5), atol=1e-6)


def test_membrane_residuals_are_barycentric(scene):
    model, _ = scene
    op = fx.make_cage(model, (0.5, 0.62, 0.78), (0.17, 0.17, 0.17), membrane_seed=5)
    o = orc.Oracle(model.desc, model.params, None, [op.to_op()])
    rng = np.random.default_rng(4)
    t = rng.integers(0, op.tets.shape[0], 300)
    b = rng.dirichlet(np.ones(4), 300).astype(np.float32)
    pts = np.einsum("nk,nkd->nd", b, op.vertices[op.tets[t]]).astype(np.float32)
    c = np.zeros((300, 7), np.float32)
    c[:, :3] = (pts - model.aabb_min) / (model.aabb_max - model.aabb_min)
    sh, od, rd = o.poisson_residuals(c)
    exp_od = np.einsum("nk,nk->n", b, op.boundary_outside_density[op.tets[t]])
    exp_sh = np.einsum("nk,nkc->nc", b, op.boundary_shs[op.tets[t]])
    ok = np.abs(od - exp_od) < 1e-3 * (1 + np.abs(exp_od))   # a point on a shared face may resolve to the neighbouring tet: same value
    assert ok.mean() > 0.98
    assert np.abs(sh - exp_sh)[ok].max() < 1e-3
    outside = np.zeros((10, 7), np.float32)
    outside[:, :3] = 0.9
    sh0, od0, rd0 = o.poisson_residuals(outside)
    assert (sh0 == 0).all() and (od0 == 0).all() and (rd0 == 0).all()


def test_render_with_identity_cage_matches_unedited(scene, oracle):
    model, occ = scene
    cv, ct = editing.box_cage((0.3, 0.4, 0.5), (0.7, 0.85, 0.95))
    tv, tets = editing.lattice_tets((0.31, 0.41, 0.51), (0.69, 0.84, 0.94), 3)
    op = editing.CageDeformation(model.aabb_min, model.aabb_max, cv, ct, tv, tets)
    f0 = syn.make_frame(model, syn.orbit_cameras(120)[17], 64, 36)
    f1 = syn.make_frame(model, syn.orbit_cameras(120)[17], 64, 36, apply_operators=True)
    base, _, st0, _ = oracle.render(f0)
    o = orc.Oracle(model.desc, model.params, occ, [op.to_op()])
    edited, _, st1, _ = o.render(f1)
    d = np.abs(base - edited).max(-1)
    assert np.quantile(d, 0.99) < 5e-3 and st0.n_hit == st1.n_hit  # positions move by ~1e-6: same image up to fp16 feature noise


def test_edit_changes_the_image_only_near_the_cage(scene, oracle):
    model, occ = scene
    ops = [o.to_op() for o in fx.e1(model)]
    f0 = syn.make_frame(model, syn.orbit_cameras(120)[17], 64, 36)
    f1 = syn.make_frame(model, syn.orbit_cameras(120)[17], 64, 36, apply_operators=True)
    base, _, _, _ = oracle.render(f0)
    o = orc.Oracle(model.desc, model.params, occ, ops)
    edited, _, st, _ = o.render(f1)
    changed = np.abs(base - edited).max(-1) > 1e-2
    assert 0.005 < changed.mean() < 0.6
    off, _, _, _ = o.render(f0)  # apply_operators = false: operators uploaded but disabled (m_enable_edits)
    assert np.array_equal(off, base)
