"""§8(f)-3: the edits file (Testbed::save_edits / load_edits, testbed.cu:3190-3236) — schema round trip and that a restored
operator list produces the same device-facing PODs (NsbEditOp) as the operators that were saved."""
import ctypes as C
import json

import numpy as np

import edit_fixtures as fx
from nerfshop_b200 import abi, edits_io
from oracle import oracle as orc


def _pod_bytes(op):
    pod, keep = op.to_op()
    scalars = bytes(pod)[: abi.NsbEditOp.tet_lut_offsets.offset] + bytes(pod)[abi.NsbEditOp.selection_box.offset:]
    return scalars, keep


def test_edits_round_trip(tmp_path, scene):
    model, _ = scene
    cage = fx.e1(model)[0]
    rng = np.random.default_rng(0)
    nc = cage.cage_original.shape[0]
    cage.cage_shs = {"inside_shs": rng.normal(0, .3, (nc, 27)).astype(np.float32), "outside_shs": rng.normal(0, .3, (nc, 27)).astype(np.float32),
                     "inside_density": rng.uniform(0, 20, nc).astype(np.float32), "outside_density": rng.uniform(1, 30, nc).astype(np.float32)}
    cage.set_membrane(*orc.membrane_blend(cage.mvc, cage.cage_shs["inside_density"], cage.cage_shs["outside_density"], cage.cage_shs["inside_shs"],
                                          cage.cage_shs["outside_shs"]), amplitude=0.7, apply=True)
    R = np.array([[0.8, -0.6, 0], [0.6, 0.8, 0], [0, 0, 1]], np.float32)
    aff = edits_io.AffineDuplicationWorld(edits_io.AffineBox((0.9, 0.7, 0.5), (0.2, 0.3, 0.2)), (0.15, 0.05, -0.1), model.aabb_min, model.aabb_max,
                                          scale=(1.2, 1.0, 0.8), rotation=R, hide_original=True, correct_dir=True)
    path = str(tmp_path / "edits.json")
    edits_io.save_edits(path, [cage, aff])
    j = json.load(open(path))
    assert [o["type"] for o in j["edit_operators"]] == ["cage_deformation", "affine_duplication"]
    cj = j["edit_operators"][0]
    # the reference's keys (growing_selection.cu:2459-2482, cage.h:100-121, tet_mesh.h:137-154)
    assert {"proxy_cage", "interpolation_mesh", "projected_pixels", "selection_points", "m_selection_grid_bitfield", "growing_level", "region_growing",
            "selection_mesh"} <= set(cj)
    assert {"vertices", "indices", "normals", "original_vertices", "inside_shs", "outside_shs", "inside_density", "outside_density"} <= set(cj["proxy_cage"])
    assert {"bbox", "original_bbox", "warped_bbox", "original_warped_bbox", "vertices", "original_vertices", "mvc_coordinates", "gamma_coordinates", "tets"} <= set(cj["interpolation_mesh"])
    assert np.asarray(cj["proxy_cage"]["inside_shs"]).shape == (nc, 9, 3)
    assert np.asarray(j["edit_operators"][1]["rotation_matrix"]).shape == (3, 3) and len(j["edit_operators"][1]["selection_box"]["u"]) == 3

    ops = edits_io.load_edits(path, model.aabb_min, model.aabb_max, blend=orc.membrane_blend)
    assert len(ops) == 2
    a, ka = _pod_bytes(cage)
    b, kb = _pod_bytes(ops[0])
    assert a == b, "scalar members of the cage POD differ after the round trip"
    for k in ("off", "idx", "tets", "v", "ov", "rot", "bits", "shs", "od", "rd"):
        assert np.array_equal(ka[k], kb[k]), k
    assert bytes(aff.to_op()[0]) == bytes(ops[1].to_op()[0])
    # second generation is identical text (a fixed point)
    path2 = str(tmp_path / "edits2.json")
    edits_io.save_edits(path2, ops)
    assert json.load(open(path2)) == j


def test_affine_world_matches_update_destination(scene):
    """Hand-computed AffineDuplication::update_destination for an axis-aligned box (affine_duplication.h:72-84)."""
    model, _ = scene
    aff = edits_io.AffineDuplicationWorld(edits_io.AffineBox((0.5, 0.5, 0.5), (0.4, 0.2, 0.2)), (1.0, 0.0, 0.0), model.aabb_min, model.aabb_max, hide_original=True)
    op, _ = aff.to_op()
    diag = model.aabb_max - model.aabb_min
    assert np.allclose(list(op.translation), np.array([1.0, 0, 0]) / diag)
    assert np.allclose(list(op.selection_box.center), (np.array([0.5, 0.5, 0.5]) - model.aabb_min) / diag)
    assert np.allclose(list(op.destination_box.center), (np.array([1.5, 0.5, 0.5]) - model.aabb_min) / diag)
    assert np.allclose(list(op.selection_box.u), [0.4 / diag[0], 0, 0]) and np.allclose(list(op.destination_box.w), [0, 0, 0.2 / diag[2]])
    assert np.allclose(list(op.selection_box.min), np.array(list(op.selection_box.center)) - 0.5 * np.array([0.4, 0.2, 0.2]) / diag)
    # the oracle maps a point of the destination box back into the selection box
    o = orc.Oracle(model.desc, model.params, None, [aff.to_op()])
    c = np.zeros((2, 7), np.float32)
    c[0, :3] = (np.array([1.45, 0.52, 0.48]) - model.aabb_min) / diag   # inside the destination
    c[1, :3] = (np.array([0.45, 0.52, 0.48]) - model.aabb_min) / diag   # inside the (hidden) original
    out, mask = o.map_rays(c)
    assert np.allclose(out[0, :3], (np.array([0.45, 0.52, 0.48]) - model.aabb_min) / diag, atol=1e-6) and mask.tolist() == [0, 1]
