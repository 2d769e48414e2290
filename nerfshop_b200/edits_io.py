"""Reader/writer of the reference's edits file — Testbed::save_edits / load_edits (testbed.cu:3190-3236).

Schema (nlohmann::json, json_binding.h:27-60: Eigen vectors as [x,y,z], matrices as nested rows):
  {"edit_operators": [
     {"type": "affine_duplication", "selection_box": AffineBoundingBox, "translation", "scale", "rotation_matrix", "hide_original",
      "correct_dir"}                                                          (affine_duplication.cu:356-369, affine_bounding_box.cuh:116-142)
     {"type": "cage_deformation", ...GrowingSelection::to_json (growing_selection.cu:2459-2482):
        "proxy_cage": Cage (cage.h:100-121), "interpolation_mesh": TetMesh (tet_mesh.h:137-154), selection bookkeeping arrays}
  ]}
Only what the render path consumes is interpreted (proxy cage, tet mesh, membrane values); the selection bookkeeping the GUI keeps
(projected_*, selection_*, region_growing, selection_mesh) is carried through untouched on a load->save round trip.
World-space quantities are converted exactly like the reference's constructors do (AffineDuplication::update_destination,
affine_duplication.h:72-84; AffineBoundingBox::warp_box, affine_bounding_box.cuh:95-103).
"""
from __future__ import annotations

import json

import numpy as np

from . import abi, editing

_PASSTHROUGH = ("projected_pixels", "projected_labels", "projected_cell_idx", "selection_points", "selection_labels", "selection_cell_idx",
                "m_selection_grid_bitfield", "growing_level", "region_growing", "selection_mesh")


def _f32(a):
    return np.asarray(a, np.float32)


# ---- AffineBoundingBox (affine_bounding_box.cuh) ------------------------------------------------------------------------------
class AffineBox:
    def __init__(self, center, scale, rot=None):
        self.center, self.scale = _f32(center), _f32(scale)
        self.rot = np.eye(3, dtype=np.float32) if rot is None else _f32(rot)
        self._refresh()

    def _refresh(self):
        R, s = self.rot, self.scale
        self.u, self.v, self.w = (R @ (s[0] * _f32([1, 0, 0]))).astype(np.float32), (R @ (s[1] * _f32([0, 1, 0]))).astype(np.float32), (R @ (s[2] * _f32([0, 0, 1]))).astype(np.float32)
        self.min = (np.float32(-0.5) * (R @ s) + self.center).astype(np.float32)
        self.max = (np.float32(0.5) * (R @ s) + self.center).astype(np.float32)

    def copy(self):
        b = AffineBox(self.center.copy(), self.scale.copy(), self.rot.copy())
        b.u, b.v, b.w, b.min, b.max = self.u.copy(), self.v.copy(), self.w.copy(), self.min.copy(), self.max.copy()
        return b

    def translate(self, t):          # :37-39 (note: min/max/u/v/w are NOT refreshed by the reference here)
        self.center = (self.center + _f32(t)).astype(np.float32)

    def scale_with_vector(self, s):  # :45-52
        self.scale = (self.scale * _f32(s)).astype(np.float32)
        self._refresh()

    def rotate(self, R):             # :54-61: u,v,w are rotated in place, rot_matrix pre-multiplied
        R = _f32(R)
        self.rot = (R @ self.rot).astype(np.float32)
        self.u, self.v, self.w = (R @ self.u).astype(np.float32), (R @ self.v).astype(np.float32), (R @ self.w).astype(np.float32)
        self.min = (np.float32(-0.5) * (self.rot @ self.scale) + self.center).astype(np.float32)
        self.max = (np.float32(0.5) * (self.rot @ self.scale) + self.center).astype(np.float32)

    def warp_box(self, aabb_min, aabb_max):  # :95-103
        diag = _f32(aabb_max) - _f32(aabb_min)
        self.center = ((self.center - _f32(aabb_min)) / diag).astype(np.float32)
        self.scale = (self.scale / diag).astype(np.float32)
        self._refresh()

    def to_json(self):
        return {"min": self.min.tolist(), "max": self.max.tolist(), "rot_matrix": self.rot.tolist(), "u": self.u.tolist(), "v": self.v.tolist(),
                "w": self.w.tolist(), "center": self.center.tolist(), "scale": self.scale.tolist()}

    @classmethod
    def from_json(cls, j):
        b = cls(j["center"], j["scale"], j["rot_matrix"])
        b.u, b.v, b.w, b.min, b.max = _f32(j["u"]), _f32(j["v"]), _f32(j["w"]), _f32(j["min"]), _f32(j["max"])
        return b

    def to_abi(self) -> abi.NsbAffineBox:
        b = abi.NsbAffineBox()
        for i in range(3):
            b.min[i], b.u[i], b.v[i], b.w[i], b.center[i] = float(self.min[i]), float(self.u[i]), float(self.v[i]), float(self.w[i]), float(self.center[i])
        return b


class AffineDuplicationWorld:
    """AffineDuplication with the reference's WORLD-space members (affine_duplication.h:86-97)."""

    def __init__(self, selection_box: AffineBox, translation, scene_aabb_min, scene_aabb_max, scale=(1, 1, 1), rotation=None, hide_original=False,
                 correct_dir=True):
        self.selection_box, self.translation, self.scale = selection_box, _f32(translation), _f32(scale)
        self.rotation = np.eye(3, dtype=np.float32) if rotation is None else _f32(rotation)
        self.hide_original, self.correct_dir = bool(hide_original), bool(correct_dir)
        self.aabb_min, self.aabb_max = _f32(scene_aabb_min), _f32(scene_aabb_max)

    def to_op(self):
        # update_destination (affine_duplication.h:72-84)
        dst = self.selection_box.copy()
        dst.translate(self.translation)
        dst.scale_with_vector(self.scale)
        dst.rotate(self.rotation)
        wdst, wsel = dst.copy(), self.selection_box.copy()
        wdst.warp_box(self.aabb_min, self.aabb_max)
        wsel.warp_box(self.aabb_min, self.aabb_max)
        wt = (self.translation / (self.aabb_max - self.aabb_min)).astype(np.float32)
        op = abi.NsbEditOp()
        op.type = abi.NSB_OP_AFFINE
        op.selection_box, op.destination_box = wsel.to_abi(), wdst.to_abi()
        for i in range(3):
            op.translation[i], op.scale[i] = float(wt[i]), float(self.scale[i])
        flat = self.rotation.T.reshape(-1)  # column-major
        for i in range(9):
            op.rotation[i] = float(flat[i])
        op.hide_original, op.correct_dir = int(self.hide_original), int(self.correct_dir)
        return op, {}

    def to_json(self):
        return {"type": "affine_duplication", "selection_box": self.selection_box.to_json(), "translation": self.translation.tolist(),
                "scale": self.scale.tolist(), "rotation_matrix": self.rotation.tolist(), "hide_original": self.hide_original, "correct_dir": self.correct_dir}


def _cage_to_json(op: editing.CageDeformation, extra: dict) -> dict:
    nc = op.cage_original.shape[0]
    zeros3 = np.zeros((nc, 3), np.float32).tolist()
    shs = getattr(op, "cage_shs", None) or {}
    z_sh = np.zeros((nc, 9, 3), np.float32)
    cage = {
        "vertices": op.cage_vertices.tolist(), "indices": op.cage_triangles.reshape(-1).tolist(), "normals": zeros3, "initial_normals": zeros3,
        "labels": [0] * nc, "original_vertices": op.cage_original.tolist(), "colors": zeros3, "outside_colors": zeros3, "initial_colors": zeros3,
        "new_shs": z_sh.tolist(), "initial_shs": z_sh.tolist(),
        "inside_shs": _sh_rows(shs.get("inside_shs"), nc), "outside_shs": _sh_rows(shs.get("outside_shs"), nc),
        "inside_density": _f32(shs.get("inside_density", np.zeros(0))).tolist(), "outside_density": _f32(shs.get("outside_density", np.zeros(0))).tolist(),
    }
    host, _ = op.to_op()
    box = lambda mn, mx: {"min": list(mn), "max": list(mx)}  # noqa: E731
    obmin, obmax = op.original_vertices.min(0), op.original_vertices.max(0)
    mesh = {
        "bbox": box(host.bbox_min, host.bbox_max), "original_bbox": box(obmin.tolist(), obmax.tolist()),
        "warped_bbox": box(host.warped_bbox_min, host.warped_bbox_max), "original_warped_bbox": box(host.original_warped_bbox_min, host.original_warped_bbox_max),
        "vertices": op.vertices.tolist(), "indices": [], "original_vertices": op.original_vertices.tolist(), "mvc_coordinates": op.mvc.tolist(),
        "gamma_coordinates": getattr(op, "gamma_coordinates", op.mvc).tolist(), "tets": op.tets.reshape(-1).tolist(), "labels": [0] * op.vertices.shape[0],
        "colors": np.zeros_like(op.vertices).tolist(), "all_indices": [],
    }
    j = {"type": "cage_deformation", "proxy_cage": cage, "interpolation_mesh": mesh}
    for k in _PASSTHROUGH:
        j[k] = extra.get(k, 0 if k == "growing_level" else ({} if k in ("region_growing", "selection_mesh") else []))
    # members the reference keeps outside the JSON (GUI state) but a headless host needs to reproduce the frame:
    j["nsb_state"] = {"copy": op.copy, "apply_poisson": op.apply_poisson, "residual_amplitude": op.residual_amplitude, "local_rotations": op.use_local_rotations}
    return j


def _sh_rows(a, nc):
    """SH9RGB = Eigen 9x3 -> 9 rows of 3 (json_binding.h:27-41); ours is [nc, 27] column-major (col*9 + k)."""
    if a is None or len(a) == 0:
        return []
    a = _f32(a).reshape(nc, 3, 9)
    return np.transpose(a, (0, 2, 1)).tolist()


def _sh_flat(rows):
    a = _f32(rows)
    if a.size == 0:
        return None
    return np.ascontiguousarray(np.transpose(a.reshape(-1, 9, 3), (0, 2, 1))).reshape(-1, 27)


def save_edits(path: str, operators) -> None:
    """Testbed::save_edits: operators in m_edit_operators order."""
    out = []
    for op in operators:
        if isinstance(op, editing.CageDeformation):
            out.append(_cage_to_json(op, getattr(op, "json_extra", {})))
        elif isinstance(op, AffineDuplicationWorld):
            out.append(op.to_json())
        else:
            raise TypeError(f"cannot serialise {type(op).__name__}: use AffineDuplicationWorld / CageDeformation")
    with open(path, "w") as f:
        json.dump({"edit_operators": out}, f)
        f.write("\n")


def load_edits(path: str, scene_aabb_min, scene_aabb_max, blend=None):
    """Testbed::load_edits. `blend(gamma, d_in, d_out, shs_in, shs_out) -> (boundary_shs, outside_density, residual_density)` is the
    interpolate_poisson_boundary implementation to use for cages that carry membrane values (NerfRenderer.cage_set_membrane does it on
    the device once the operator is uploaded; pass a callable here to have the arrays at upload time)."""
    with open(path) as f:
        j = json.load(f)
    ops = []
    for oj in j["edit_operators"]:
        t = oj["type"]
        if t == "affine_duplication":
            ops.append(AffineDuplicationWorld(AffineBox.from_json(oj["selection_box"]), oj["translation"], scene_aabb_min, scene_aabb_max, oj["scale"],
                                              oj["rotation_matrix"], oj["hide_original"], oj["correct_dir"]))
        elif t == "cage_deformation":
            cage, mesh, st = oj["proxy_cage"], oj.get("interpolation_mesh"), oj.get("nsb_state", {})
            if mesh is None:
                raise ValueError("cage_deformation without interpolation_mesh: nothing to render (the reference keeps such operators inert)")
            op = editing.CageDeformation.from_arrays(scene_aabb_min, scene_aabb_max, cage["original_vertices"], cage["vertices"], cage["indices"],
                                                     mesh["original_vertices"], mesh["tets"], mesh["mvc_coordinates"], mesh.get("gamma_coordinates") or None,
                                                     copy=st.get("copy", False), local_rotations=st.get("local_rotations", True))
            op.json_extra = {k: oj[k] for k in _PASSTHROUGH if k in oj}
            shs_in, shs_out = _sh_flat(cage.get("inside_shs", [])), _sh_flat(cage.get("outside_shs", []))
            if shs_in is not None and shs_out is not None:
                op.cage_shs = {"inside_shs": shs_in, "outside_shs": shs_out, "inside_density": _f32(cage["inside_density"]), "outside_density": _f32(cage["outside_density"])}
                if blend is not None:
                    b = blend(op.gamma_coordinates, op.cage_shs["inside_density"], op.cage_shs["outside_density"], shs_in, shs_out)
                    op.set_membrane(*b, amplitude=st.get("residual_amplitude", 1.0), apply=st.get("apply_poisson", True))
            ops.append(op)
        elif t == "twist":
            continue  # commented out in the reference (testbed.cu:3218-3219)
        else:
            raise RuntimeError("Invalid edit operator!")  # testbed.cu:3232
    return ops
