"""The C++ host layer (basevar_b200/host): compiled test programs, run from pytest.
CPU: packer / encodings / error messages / BaseType getters / strand_bias / region sharding / no-fallback.
GPU: BaseType + strand_bias mirror vs the oracle and the compiled reference, drop-in constructor, sharded region run."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "bin")


@pytest.fixture(scope="module")
def host_built(built_lib, oracle_lib):
    from basevar_b200 import build
    build.build_host()
    return BIN


def _run(exe, *args, env=None):
    p = subprocess.run([exe, *args], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "ALL OK" in p.stdout, p.stdout[-4000:] + p.stderr[-2000:]
    return p.stdout


def test_host_layer_cpu(host_built):
    import torch
    env = dict(os.environ)
    if torch.cuda.is_available():
        env["BV_EXPECT_GPU"] = "1"
    _run(os.path.join(host_built, "test_host_cpu"), env=env)


def test_fast_fisher_matches_the_reference_algorithm(host_built):
    """csrc/bv_fisher_fast.h (O(log range) two-sided Fisher of the product, compiled for the host) vs the oracle's
    restatement of kt_fisher_exact on 1.6e5 tables."""
    out = _run(os.path.join(host_built, "test_fisher_fast"), ROOT)
    assert "through the fast path" in out


@pytest.mark.gpu
def test_host_layer_gpu(host_built):
    out = _run(os.path.join(host_built, "test_host_gpu"), ROOT)
    assert "region sharding" in out and "batch path" in out


def test_caller_text_side_cpu(host_built):
    """basevar_b200/host/bv_caller: number formatting, CVG / VCF rows from hand-made records, headers, no-fallback."""
    import torch
    env = dict(os.environ)
    if torch.cuda.is_available():
        env["BV_EXPECT_GPU"] = "1"
    _run(os.path.join(host_built, "test_caller_cpu"), env=env)


@pytest.mark.gpu
def test_caller_c1_end_to_end_gpu(host_built):
    """BASELINE.json configs[0]: the reference's own batchfile rows -> our driver -> VCF / CVG text, byte-identical with
    what the unmodified reference CLI wrote (tests/golden/c1, with and without --pop-group)."""
    out = _run(os.path.join(host_built, "test_caller_gpu"), ROOT)
    assert out.count("identical, CVG") >= 4 and "DIFFERENT" not in out and "error behaviour: ok" in out
