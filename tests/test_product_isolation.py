"""The oracle is test infrastructure: nothing in the product may import, call, link or execute it, and there is no CPU
fallback path in the product (a missing CUDA library is an error)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _product_files():
    for base in ("nerfshop_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            if "__pycache__" in d or d.endswith(os.sep + "lib"):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    yield os.path.join(d, f)


def test_product_never_touches_the_oracle():
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b|oracle[/\\.]|libnerfshop_oracle|orc_[a-z_]+\()", re.M)
    offenders = [p for p in _product_files() if pat.search(open(p).read())]
    assert not offenders, offenders


def test_bench_uses_the_oracle_only_in_the_baseline_legs():
    """bench.py touches oracle/ in exactly two functions, both baselines and never the thing measured: the CPU arm (cpu_baseline / --impl reference:
    the oracle port or oracle/_ref's CPU build of the reference) and the gpu_baseline arm (oracle/_ref's nvcc build of the reference's own CUDA path)."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"from oracle import", src)]
    assert uses
    owners = set()
    for u in uses:
        fn_start = src.rfind("\ndef ", 0, u) + 1
        owners.add(re.match(r"def (\w+)", src[fn_start:]).group(1))
    assert owners <= {"cpu_reference_run", "measure_reference_cuda_and_edit_configs"}, owners


def test_missing_library_is_an_error(tmp_path):
    from nerfshop_b200 import abi

    try:
        abi.load_library(str(tmp_path / "nope.so"))
        raise AssertionError("expected NsbError")
    except abi.NsbError as e:
        assert "no CPU fallback" in str(e)
