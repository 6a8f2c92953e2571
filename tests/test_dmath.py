"""log10_fast (crisp_b200/csrc/dmath.cuh): the EM's own logarithm, checked on the CPU with the device's table and constants."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_log10_fast_accuracy(tmp_path):
    src = open(os.path.join(ROOT, "crisp_b200", "csrc", "dmath.cuh")).read()
    tab = re.search(r"kLog10Tab\[256\] = \{(.*?)\};", src, re.S).group(1)
    k = re.search(r"kLog10K\[11\] = \{(.*?)\};", src, re.S).group(1)
    assert len([v for v in tab.replace("\n", " ").split(",") if v.strip()]) == 256
    (tmp_path / "log10_tab.h").write_text(f"static const double kLog10Tab[256] = {{{tab}}};\nstatic const double kLog10K[11] = {{{k}}};\n")
    exe = str(tmp_path / "check")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", f"-I{tmp_path}", "-o", exe, os.path.join(ROOT, "tests", "csrc", "check_log10_fast.c"), "-lm"], check=True)
    out = subprocess.run([exe, "4000000"], capture_output=True, text=True, check=True).stdout.split()
    max_ulp, max_abs_near_one = float(out[0]), float(out[1])
    assert max_ulp < 2.5, max_ulp                  # the statistics' parity bar is 1e-9 relative
    assert max_abs_near_one < 4e-18, max_abs_near_one   # |log10 x| < 0.01: absolute error (what the EM sums see)
