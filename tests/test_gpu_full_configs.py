"""-m gpu: every BASELINE.json configuration at its FULL size (C4: the shard one GPU of eight takes), streamed in tiles
generated on the device (tools/full_configs.py): the size-independent properties -- depth conservation, fwd + rev ==
depth, a checksum of checksums that does not depend on the tile size -- and random sites of the region against the CPU
oracle through the host twin of the generator."""
import pytest

from tools import full_configs

pytestmark = pytest.mark.gpu


def _assert_ok(r):
    print(r)
    assert r["depth_conservation_violations"] == 0
    assert r["strand_sum_violations"] == 0
    assert r["tiling_invariant"], (r["checksum"], r["checksum_b"])
    # flips: calls that differ at sites the ORACLE marks as sitting on the LRT threshold or as a tie of two subsets (tests/util.py)
    assert r["spot_exact_mismatches"] == 0 and r["spot_float_mismatches"] == 0 and r["spot_flips"] <= max(2, r["spot_sites"] // 1000)
    assert r["variant_sites"] > 0 and r["covered_sites"] > 0.9 * r["sites"]


def test_c2_full_size(built_lib, oracle_lib):
    r = full_configs.run_config("C2", spot=20000)
    assert r["sites"] == 1_000_000 and r["n_samples"] == 1_000
    _assert_ok(r)


def test_c3_full_size(built_lib, oracle_lib):
    """10,000 samples x 10,000,000 sites = 1e11 sample-sites, about 300 GB of planes streamed through one GPU."""
    r = full_configs.run_config("C3", spot=10000)
    assert r["sites"] == 10_000_000 and r["n_samples"] == 10_000
    _assert_ok(r)


def test_c4_one_shard_of_eight(built_lib, oracle_lib):
    """100,000 samples x 64,000,000 sites over 8 GPUs: GPU 3's contiguous shard (8,000,000 sites = 8e11 sample-sites)."""
    r = full_configs.run_config("C4", shard=(3, 8), spot=2000)
    assert r["sites"] == 8_000_000 and r["site_range"] == [24_000_000, 32_000_000] and r["n_samples"] == 100_000
    _assert_ok(r)


@pytest.mark.parametrize("abs_mode", [0, 1])
def test_c5_full_size_both_abs_modes(built_lib, oracle_lib, abs_mode):
    r = full_configs.run_config("C5", abs_mode=abs_mode, spot=10000)
    assert r["sites"] == 1_000_000 and r["n_samples"] == 2_000
    _assert_ok(r)
    assert r["variant_sites"] > 200_000


def test_kernel_modes_leave_no_trace(built_lib, oracle_lib):
    """Work that is shaped by the size of a tile, on the same 160,000 deep multi-allelic sites as one tile and in tiles of 8,000:
    bv_fisher_kernel gives a test a pair of lanes (one tail each) while that is one trip of its grid, else one thread (~200,000
    listed tests in the large tile, ~10,000 per small one); bv_em_task_kernel runs its high-occupancy build with a thread per EM
    task on the large tile (~148,000 tasks) and the other build with four lanes per task on the small ones (~7,400).  The records
    are the same, bit for bit: every sum is formed in the same order in either mode."""
    r = full_configs.run_config("C5", max_sites=160_000, tile_sites=160_000, tile_sites_b=8_000, spot=1000)
    assert r["sites"] == 160_000
    _assert_ok(r)
