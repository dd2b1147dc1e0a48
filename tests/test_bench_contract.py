"""CPU-side checks of bench.py's bookkeeping (no GPU, no timing): the roofline classification the
round-1 review asked for (`bound` l2 vs hbm from measured DRAM traffic, no HBM fraction above 1
passed off as an HBM number), the committed traffic keys the default run looks up, the argument
defaults of the contract, and that both arms describe the SAME workload."""
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


@pytest.fixture(scope="module")
def bench():
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        import bench as b
    finally:
        sys.argv = argv
    return b


def test_defaults_follow_the_contract(bench):
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        a = bench.parse()
    finally:
        sys.argv = argv
    assert a.gpus == 1 and a.warmup >= 3 and a.steps >= 1
    assert a.batch == 65536 and a.dim == 128 and a.shape == "ml-20m" and a.opt == "sgd"  # BASELINE configs[1]
    assert a.impl == "ours"


def test_roofline_bound_follows_measured_traffic(bench):
    alg = 24.0 * 128 * 65536
    # the committed capture of the headline kernel: 50 MB of DRAM traffic for 201 MB algorithmic -> L2-bound
    r = bench.roofline("bpr_phase_a", "bpr_phase_a:ml-20m:D128:B65536:sgd:N1", alg, 0.0282, 1e9)
    assert r["traffic"] is not None and r["traffic"] < 0.5 * alg
    assert r["bound"] == "l2"
    assert r["frac_hbm_dram"] < 1.0 < r["frac_algorithmic"] * 1.2  # the >1 number is labelled l2, the DRAM one is honest
    assert r["frac"] == r["frac_algorithmic"]
    # MSD / Adam: DRAM traffic above the 24*D formula -> hbm
    r = bench.roofline("bpr_phase_a", "bpr_phase_a:msd:D256:B65536:adam:N1", 24.0 * 256 * 65536, 0.146, 1e9)
    assert r["bound"] == "hbm" and r["frac"] < 1.0 and r["frac_hbm_dram"] < 1.0
    # no capture for the key: the working set decides
    r = bench.roofline("k", "no:such:key", alg, 0.03, 10e6)
    assert r["traffic"] is None and r["bound"] == "l2" and r["frac_hbm_dram"] is None
    r = bench.roofline("k", "no:such:key", alg, 0.03, 1e9)
    assert r["bound"] == "hbm"
    # an untimed kernel yields no fractions instead of a division by zero
    r = bench.roofline("k", "no:such:key", alg, None, 1e9)
    assert r["achieved"] is None and r["frac"] is None


def test_traffic_file_has_the_keys_the_default_run_quotes(bench):
    z = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    for key in ("bpr_phase_a:ml-20m:D128:B65536:sgd:N1", "bpr_phase_a:ml-20m:D128:B262144:sgd:N1",
                "bpr_phase_a:msd:D256:B65536:adam:N1", "score:ml-20m:D128:U10000:N1"):
        assert isinstance(z[key], int) and z[key] > 0
        assert key in z["source"], f"{key}: say which capture the number comes from"
        assert bench.ncu_traffic(key) == z[key]


def test_both_arms_describe_the_same_workload(bench):
    dims = (136677, 20108, 9676553)
    ours = bench.workload("ml-20m", dims, 128, "sgd", "uniform", 65536, 1)
    again = bench.workload("ml-20m", dims, 128, "sgd", "uniform", 65536, 1)
    assert ours == again and "workload" in ours and "model" not in ours
    assert "65536" in ours["workload"].replace(" ", "").replace(",", "") or ours.get("batch") == 65536
    eight = bench.workload("ml-20m", dims, 128, "sgd", "uniform", 65536, 8)
    assert eight != ours  # the sharding is part of the description


def test_peaks_come_from_the_driver_file_when_present(bench):
    hbm, bf16, src = bench.peaks()
    assert 3000 < hbm < 9000 and 500 < bf16 < 2500
    if (ROOT / "MEASURED_PEAKS.json").exists():
        assert "measured" in src
