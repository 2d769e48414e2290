"""CPU: the oracle reproduces the committed golden vectors bit for bit (tests/golden/make_golden.py minted them;
the reference itself has none — see the oracle's header). Guards the oracle and the synthetic generator against drift."""
import os

import numpy as np

from nerfshop_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synthetic_fox.npz")


def test_synthetic_inputs_are_reproducible(scene):
    model, occ = scene
    g = np.load(GOLDEN)
    assert int(model.params.astype(np.uint64).sum()) == int(g["params_checksum"][0])
    assert int(occ.astype(np.uint64).sum()) == int(g["params_checksum"][1])


def test_oracle_ops_match_golden(oracle):
    g = np.load(GOLDEN)
    assert np.array_equal(oracle.encode(g["coords"]), g["encode"])
    assert np.array_equal(oracle.inference(g["coords"]), g["inference"])
    assert np.array_equal(oracle.inference(g["coords"], density_only=True), g["density"])


def test_oracle_render_matches_golden(scene, oracle):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLDEN), "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    model, _ = scene
    g = np.load(GOLDEN)
    for name, cam in (("fox0", syn.fox_camera0()), ("orbit17", syn.orbit_cameras(120)[17])):
        f = syn.make_frame(model, cam, mg.GOLDEN_W, mg.GOLDEN_H)
        fb, depth, st, margin = oracle.render(f, want_margin=True)
        # expf/powf come from libm: allow the last bits to differ between hosts, nothing more
        assert np.allclose(fb, g[f"{name}_rgba"], atol=2e-6, rtol=0)
        assert np.allclose(depth, g[f"{name}_depth"], rtol=1e-6)
        assert [st.n_rays, st.n_rays_alive, st.n_hit, st.n_samples] == [int(v) for v in g[f"{name}_stats"]]
        rec, idx, cnt = oracle.march_trace(f, mg.TRACE_PIXELS, mg.TRACE_MAX)
        assert np.array_equal(cnt, g[f"{name}_trace_cnt"]) and np.array_equal(idx, g[f"{name}_trace_idx"])
        assert np.array_equal(rec.view(np.uint32), g[f"{name}_trace_rec"].view(np.uint32))  # march is bit-exact by contract
