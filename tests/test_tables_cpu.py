"""Host-side exhaustive checks of the constant tables the CUDA kernels use (no GPU needed):
the cell LUT reproduces the double-precision curve for every possible matrix output, and the
encode path's three-instruction constant division equals IEEE division for every 8-bit RGB."""
from metalbt709decoder_b200 import _lib


def test_tables_selftest():
    assert _lib.lib().bt709_selftest_host() == 0


def test_biased_log_luts_on_every_float_of_the_unit_interval(tmp_path):
    """tools/probes/blog_exhaustive.cpp: the three biased log-cell LUTs the kernels use (sRGB encode of pass 2 and the
    Linear-mode decode; BT709_from_linear for the Apple and sRGB curves of the 4:2:0 encode) equal their step functions
    on all 1 065 353 217 float32 values of [0, 1].  Host code only (bt709_tables.cpp compiled with g++)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "blog_exhaustive")
    csrc = os.path.join(root, "metalbt709decoder_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + csrc, "-o", exe, os.path.join(root, "tools", "probes", "blog_exhaustive.cpp"),
                    os.path.join(csrc, "bt709_tables.cpp"), "-lpthread"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(" 0 mismatches") == 3, r.stdout
