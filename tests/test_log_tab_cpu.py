"""CPU: accuracy of the table logarithm the EM task kernel sums log-likelihoods with (log_tab, basevar_b200/csrc/bv_em_kernels.cuh).
tools/log_tab_check.c restates it operation by operation on the host (same table construction as bv_api.cu, fma where the device
code says fma) and compares with 80-bit logl over marginals from 1e-13 to 1 and arguments next to 1."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_log_tab_against_long_double(tmp_path):
    exe = str(tmp_path / "log_tab_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tools", "log_tab_check.c"), "-lm"])
    out = subprocess.run([exe, "4000000"], capture_output=True, text=True, check=True).stdout
    val = {k: float(v) for k, v in re.findall(r"^(max_\S+) (\S+)$", out, re.M)}
    # absolute error relative to max(|log x|, ln 2): what a sum of same-signed terms c * log m inherits
    assert val["max_err_over_max_abs_log_ln2"] < 3.0e-16, out
    # next to x = 1 (|log x| < 0.05) the error is absolute
    assert val["max_abs_err_where_abs_log_lt_0.05"] < 1.0e-16, out


def test_device_code_and_check_program_share_their_constants():
    dev = open(os.path.join(ROOT, "basevar_b200", "csrc", "bv_em_kernels.cuh")).read()
    chk = open(os.path.join(ROOT, "tools", "log_tab_check.c")).read()
    for const in ("0.693147180559945309417", "-1.0 / 6", "1.0 / 3"):
        assert const in dev and const.replace("-1.0 / 6", "-1.0/6").replace("1.0 / 3", "1.0/3") in chk, const
    api = open(os.path.join(ROOT, "basevar_b200", "csrc", "bv_api.cu")).read()
    assert "((double)i + 0.5) / (double)bv::kLogTabEntries" in api and "(i+0.5)/128" in chk
