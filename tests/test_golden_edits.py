"""CPU: the oracle and the host geometry reproduce tests/golden/synthetic_edits.npz (minted by make_golden_edits.py) — the pins of
BASELINE configs 2-3 and the SURVEY §8(f) rows. Integer/pinned-fp32 results bit for bit; anything through libm (expf, cosf, ...)
to the last bits."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "synthetic_edits.npz")


def _minted(scene):
    spec = importlib.util.spec_from_file_location("make_golden_edits", os.path.join(HERE, "golden", "make_golden_edits.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    model, occ = scene
    return mg.build(model, occ)


def test_oracle_and_host_geometry_match_golden_edits(scene):
    g = np.load(GOLDEN)
    now = _minted(scene)
    assert set(now) == set(g.files)
    exact = [k for k in g.files if k.endswith(("_vertices", "_csr_sha256", "_n_idx", "_mask", "_stats", "_touched", "_sub_idx", "_bits")) or k in ("map_in", "map_out", "res_sh", "res_od", "res_rd")]
    for k in exact:
        a, b = now[k], g[k]
        assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8)), k
    for k in set(g.files) - set(exact):
        tol = 2e-6 if k.endswith(("_rgba", "_rotations")) else 1e-5
        assert np.allclose(now[k], g[k], rtol=tol, atol=tol), (k, float(np.abs(now[k] - g[k]).max()))
