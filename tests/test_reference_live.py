"""The LIVE reference (nbkode, numba mode, installed unmodified under the git-ignored baseline/_ref) as it is used
by bench.py's reference arm / `cpu_baseline` / `parity` legs (baseline/nbkode_ref.py).

CPU-only.  Skipped when baseline/_ref is not populated (`python baseline/nbkode_ref.py --install` in the build
container; __graft_entry__.build() does it whenever /root/reference is present).

What is pinned here:
 * the numba-mode reference driven through the compile-once parameter buffer reproduces the
   committed golden rows of the 1M Lorenz ensemble, which tests/golden/make_golden.py generated with the SAME
   reference in pure-Python mode through the ordinary ``params=`` path: bit-identical end states, identical
   accepted-step counts -- so the buffer trick and the JIT change nothing;
 * the bit-exact NumPy restatement and the C port agree with the live numba run on fresh trajectories that are
   in no fixture.
"""

import numpy as np
import pytest

from baseline import nbkode_ref

pytestmark = pytest.mark.skipif(not nbkode_ref.installed(), reason="baseline/_ref (the live reference) is not installed")


@pytest.fixture(scope="module")
def pool():
    p = nbkode_ref.Pool(2, "lorenz", "RungeKutta45", rtol=1e-6, atol=1e-9)
    yield p
    p.close()


def test_loaded_from_baseline_ref():
    nbkode = nbkode_ref.load()
    assert "baseline/_ref" in nbkode.__file__.replace("\\", "/")
    assert len(nbkode.get_solvers()) == 26


def test_numba_path_reproduces_golden_rows(pool, golden):
    g = golden("lorenz_ensemble_rk45.json")
    rows = g["rows"][:12]
    y0 = np.array([r["y0"] for r in rows])
    p = np.array([r["p"] for r in rows])
    out = pool.run(y0, p, [g["T"]], count=True)
    for i, r in enumerate(rows):
        assert out["n_accepted"][i] == r["n_accepted"]
        assert out["t_end"][i] == r["t_end"]
        # pure-Python mode (fixtures) and numba mode may differ by BLAS / LLVM contraction in the last bits only
        np.testing.assert_allclose(out["y_end"][i], r["y_end"], rtol=1e-9, atol=0)
        np.testing.assert_allclose(out["ys"][i], np.array(r["ys"]), rtol=1e-9, atol=0)


def test_oracles_agree_with_live_numba_on_fresh_trajectories(pool):
    from oracle import nbk_oracle_c as oc
    from oracle import nbk_oracle_np as onp

    y0, p = onp.lorenz_ensemble(48, seed=1234)
    live = pool.run(y0, p, [10.0], count=True)
    c = oc.run_ensemble("RungeKutta45", "lorenz", y0, p, [10.0], rtol=1e-6, atol=1e-9, nthreads=2)
    assert np.array_equal(c["n_accepted"], live["n_accepted"])
    tol = 1e-6 * np.abs(live["ys"]) + 1e-9
    assert np.all(np.abs(c["ys"] - live["ys"]) <= 1e-3 * tol)  # three orders inside the north-star bar
    for i in range(4):
        s = onp.AdaptiveRK("RungeKutta45", onp.rhs_lorenz, 0.0, y0[i], p[i], rtol=1e-6, atol=1e-9)
        _, ys = onp.run(s, [10.0])
        np.testing.assert_allclose(ys[0], live["ys"][i, 0], rtol=1e-9, atol=0)
        assert s.t == live["t_end"][i]
