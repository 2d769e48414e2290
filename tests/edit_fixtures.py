"""Edit-operator fixtures: E1 = one box cage around the fox head, lattice tets, MVC, one face
pulled out; E3 = three disjoint cages, the first with membrane (Poisson) arrays from seeded values."""
import numpy as np

from nerfshop_b200 import editing


def make_cage(model, center, half, pull=(0.10, 0.03, 0.0), n_lattice=3, copy=False, membrane_seed=None):
    center, half = np.asarray(center, np.float32), np.asarray(half, np.float32)
    cv, ct = editing.box_cage(center - half, center + half)
    tv, tets = editing.lattice_tets(center - 0.97 * half, center + 0.97 * half, n_lattice)
    op = editing.CageDeformation(model.aabb_min, model.aabb_max, cv, ct, tv, tets, copy=copy)
    moved = op.cage_original.copy()
    moved[moved[:, 0] > center[0]] += np.asarray(pull, np.float32)  # pull the +x face
    op.cage_vertices = moved
    op.update_tet_mesh()
    if membrane_seed is not None:
        rng = np.random.default_rng(membrane_seed)
        nv = op.vertices.shape[0]
        shs = rng.uniform(-0.3, 0.3, (nv, 27)).astype(np.float32)
        shs[:, [0, 9, 18]] += 1.2  # DC terms: a visible base colour
        op.set_membrane(shs, rng.uniform(0.0, 30.0, nv).astype(np.float32), rng.uniform(-2.0, 6.0, nv).astype(np.float32), amplitude=1.0, apply=True)
    return op


def e1(model):
    return [make_cage(model, (0.5, 0.62, 0.78), (0.17, 0.17, 0.17))]


def e3(model):
    return [
        make_cage(model, (0.5, 0.62, 0.78), (0.17, 0.17, 0.17), membrane_seed=5),
        make_cage(model, (0.5, 0.55, 0.12), (0.12, 0.12, 0.2), pull=(0.0, 0.08, 0.0)),
        make_cage(model, (0.41, 0.30, 0.33), (0.08, 0.14, 0.08), pull=(0.05, 0.0, 0.05), copy=True),
    ]
