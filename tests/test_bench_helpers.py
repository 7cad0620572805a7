"""bench.py's host-side helpers (no GPU): the parity block, the L2 policy, the workload description shared by both arms."""

import gc
import importlib.util
import os

import numpy as np

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_parity_stats_counts_outliers_and_step_differences():
    b = _bench()
    rng = np.random.default_rng(0)
    ref = rng.uniform(1, 2, (1000, 1, 3))
    y = ref * (1 + 1e-8)            # 0.01 of the tolerance
    nacc_ref = np.full(1000, 400)
    nacc = nacc_ref.copy()
    y[7] = ref[7] * (1 + 5e-6)      # 5 x the tolerance
    y[11, 0, 1] = np.nan            # a non-finite result is an outlier, not a pass
    nacc[3] += 1                    # within the bar
    nacc[5] += 2                    # outside
    st = b.parity_stats(y, nacc, ref, nacc_ref)
    assert st["trajectories"] == 1000 and st["n_outliers"] == 3
    assert abs(st["frac_within_tol"] - 0.998) < 1e-12 and abs(st["frac_steps_within_1"] - 0.999) < 1e-12
    assert abs(st["frac_steps_identical"] - 0.998) < 1e-12
    assert {o["trajectory"] for o in st["outliers"]} == {5, 7, 11}
    assert st["outliers"][0]["trajectory"] == 11  # non-finite first


def test_workload_description_is_the_same_for_both_arms_and_l2_policy():
    b = _bench()
    c1, c8 = b.workload_config(1), b.workload_config(8)
    assert c1["trajectories_total"] == c8["trajectories_total"] == 1_000_000
    assert c1["trajectories_per_gpu"] == 1_000_000 and c8["trajectories_per_gpu"] == 125_000
    assert set(c1) == {"workload", "trajectories_total", "trajectories_per_gpu", "parallelism", "l2"}
    ring1, note1 = b.l2_policy(1_000_000)
    ring8, note8 = b.l2_policy(125_000)
    assert ring1 == 1 and "exceeds" in note1
    assert ring8 * 125_000 * 416 / 1e6 > 252 and "rotate" in note8  # the rotation is larger than twice the L2


def test_no_gc_restores_the_collector():
    b = _bench()
    assert gc.isenabled()
    with b.no_gc():
        assert not gc.isenabled()
    assert gc.isenabled()
