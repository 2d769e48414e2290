"""The C++ shim that keeps the reference's class names compiles against the C ABI and links with the library (CPU check)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_shim_compiles_and_links(built_lib, tmp_path):
    src = tmp_path / "shim.cpp"
    src.write_text(
        '#include "nerfshop_b200/host/nerfshop_host.hpp"\n'
        "#include <cstdio>\n"
        "int main() {\n"
        "  NsbModelDesc d{16, 2, 19, 16, 1.5157166f, 64, 1, 2, 4};\n"
        "  uint64_t n = 0;\n"
        "  ngp_b200::check(nsb_model_n_params(&d, &n), \"n_params\");\n"
        "  std::printf(\"%llu\\n\", (unsigned long long)n);\n"
        "  try { auto ctx = std::make_shared<ngp_b200::Context>(0); ngp_b200::NerfTracer t(ctx); ngp_b200::NerfNetwork net(ctx, d);\n"
        "        t.add_edit_operator(std::make_shared<ngp_b200::AffineDuplication>()); t.reset_edit_operator(); std::printf(\"gpu\\n\"); }\n"
        "  catch (const std::runtime_error& e) { std::printf(\"no-gpu: %s\\n\", e.what()); }\n"
        "  return 0;\n}\n"
    )
    exe = tmp_path / "shim"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.dirname(built_lib)
    subprocess.run([cxx, "-std=c++17", "-I", ROOT, str(src), "-o", str(exe), "-L", libdir, "-lnerfshop_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0] == str(10240 + 13074912)
    assert out[1].startswith("gpu") or out[1].startswith("no-gpu")
