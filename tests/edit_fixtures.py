"""Edit-operator fixtures: E1 = one box cage around the fox head, lattice tets, MVC, one face
pulled out; E3 = three disjoint cages, the first with membrane (Poisson) arrays from seeded values."""
import numpy as np

from nerfshop_b200 import editing


def make_cage(model, center, half, pull=(0.10, 0.03, 0.0), n_lattice=3, copy=False, membrane_seed=None):
    return editing.lattice_cage(model, center, half, pull=pull, n_lattice=n_lattice, copy=copy, membrane_seed=membrane_seed)


def e1(model):
    return [make_cage(model, (0.5, 0.62, 0.78), (0.17, 0.17, 0.17))]


def e3(model):
    return [
        make_cage(model, (0.5, 0.62, 0.78), (0.17, 0.17, 0.17), membrane_seed=5),
        make_cage(model, (0.5, 0.55, 0.12), (0.12, 0.12, 0.2), pull=(0.0, 0.08, 0.0)),
        make_cage(model, (0.41, 0.30, 0.33), (0.08, 0.14, 0.08), pull=(0.05, 0.0, 0.05), copy=True),
    ]
