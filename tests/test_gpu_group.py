"""GPU parity tests of the sub-warp regime (csrc/nbk_group.cuh): G lanes of a warp per trajectory for small and mid-size
component-wise systems under RungeKutta23 / RungeKutta45 (the north star's register-resident regime up to n = 32 and the
band up to n = 64).  Checkers: the C port and the bit-exact NumPy restatement of the reference (test infrastructure).
Bars: the north star's -- final states / dense output within rtol |y| + atol, accepted steps within +-1."""

import os

import numpy as np
import pytest

import nbkode_b200 as nbk
from conftest import parity_bar
from oracle import nbk_oracle_c as oc
from oracle import nbk_oracle_np as onp

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-6, 1e-9


def decay_ensemble(B, n, seed, scalar=False):
    rng = np.random.default_rng(seed)
    return rng.uniform(0.5, 2, (B, n)), rng.uniform(0.1, 1.0, (B, 1 if scalar else n))


@pytest.mark.parametrize("cls", ["RungeKutta45", "RungeKutta23"])
@pytest.mark.parametrize("n,lanes", [(13, 4), (16, 4), (24, 8), (32, 8), (33, 16), (50, 16), (64, 16)])
def test_decay_midsize_vs_oracle(cls, n, lanes):
    """The built-in decay from n = 13 on takes the component-wise form: 4, 8 or 16 lanes per trajectory, three or four
    components per lane, ragged last lanes at n = 13, 33, 50.  4097 trajectories: 129 warps' worth at G = 4, every group
    refills from the queue several times."""
    B = 4097
    y0, k = decay_ensemble(B, n, seed=n)
    ts = [1.0, 4.0, 10.0]
    ref = oc.run_ensemble(cls, "decay", y0, k, ts, rtol=RTOL, atol=ATOL, nthreads=8)
    s = getattr(nbk, cls)("decay", 0.0, y0, k, rtol=RTOL, atol=ATOL)
    assert s._regime == 1
    t, y = s.run(ts)
    assert s.last_launch()["lanes_per_trajectory"] == lanes and s.last_launch()["block"] == 128
    assert (s.status == 0).all() and np.abs(s.n_accepted - ref["n_accepted"]).max() <= 1
    parity_bar(f"group/decay/{cls}/n{n}/gpu_vs_c_port", y, ref["ys"], RTOL, ATOL, 1.0)
    assert np.abs(s.t - ref["t_end"]).max() < 1e-7 and np.abs(s.h - ref["h_end"]).max() / ref["h_end"].max() < 1e-6
    # the per-thread kernels on the same ensemble (regime="thread"): two regimes, one answer
    if n <= 24:
        th = getattr(nbk, cls)("decay", 0.0, y0[:512], k[:512], rtol=RTOL, atol=ATOL, regime="thread")
        assert th._regime == 0
        _, yt = th.run(ts)
        parity_bar(f"group/decay/{cls}/n{n}/group_vs_thread_regime", y[:512], yt, RTOL, ATOL, 1.0)
        assert np.abs(th.n_accepted - s.n_accepted[:512]).max() <= 1


def test_scalar_rate_and_unbatched():
    y0, k = decay_ensemble(300, 20, seed=1, scalar=True)
    ref = oc.run_ensemble("RungeKutta45", "decay", y0, k, [2.0, 5.0], rtol=RTOL, atol=ATOL, nthreads=4)
    s = nbk.RungeKutta45("decay", 0.0, y0, k, rtol=RTOL, atol=ATOL)
    _, y = s.run([2.0, 5.0])
    assert s.last_launch()["lanes_per_trajectory"] == 8
    parity_bar("group/decay_scalar/n20", y, ref["ys"], RTOL, ATOL, 1.0)
    one = nbk.RungeKutta45("decay", 0.0, y0[7], k[7], rtol=RTOL, atol=ATOL)  # a single trajectory: one group of one warp works
    t1, y1 = one.run([2.0, 5.0])
    assert y1.shape == (2, 20) and np.array_equal(y1, y[7])


@pytest.mark.parametrize("cls,n,B", [("RungeKutta45", 8, 600), ("RungeKutta45", 24, 600), ("RungeKutta45", 32, 333), ("RungeKutta23", 40, 200),
                                     ("RungeKutta45", 64, 130)])
def test_rd1d_small_grids_vs_restatement(cls, n, B):
    """The stencil right-hand side (neighbours through the group's stage buffer, periodic wrap-around inside the group)
    against the bit-exact NumPy restatement on a prefix and the C port on all trajectories; dense output on a grid."""
    y0, p = onp.rd1d_ensemble(B, n=n, seed=3)
    ts = [0.25, 0.5, 1.0]
    ref = oc.run_ensemble(cls, "rd1d", y0, p, ts, rtol=RTOL, atol=ATOL, nthreads=8)
    s = getattr(nbk, cls)("rd1d", 0.0, y0, p, rtol=RTOL, atol=ATOL)
    _, y = s.run(ts)
    assert s._regime == 1 and s.last_launch()["lanes_per_trajectory"] in (2, 8, 16)
    assert np.abs(s.n_accepted - ref["n_accepted"]).max() <= 1
    parity_bar(f"group/rd1d/{cls}/n{n}/gpu_vs_c_port", y, ref["ys"], RTOL, ATOL, 1.0)
    ynp = []
    for i in range(6):
        o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[i], p[i], rtol=RTOL, atol=ATOL)
        ynp.append(onp.run(o, ts)[1])
        assert abs(int(s.n_accepted[i]) - o.n_accepted) <= 1
    parity_bar(f"group/rd1d/{cls}/n{n}/gpu_vs_numpy_restatement", y[:6], np.array(ynp), RTOL, ATOL, 1.0)


USER_RING = """
NBK_RHS_I nbk_rhs_i(double t, nbk_vec y, const double* p, int i, int n) {
    const double yl = y[i == 0 ? n - 1 : i - 1], yr = y[i + 1 == n ? 0 : i + 1];
    return p[0] * (yr - 2 * y[i] + yl) - p[1] * y[i] * y[i] * y[i] + sin(t);
}
"""


def ring(t, y, p):
    return p[0] * (np.roll(y, 1) - 2 * y + np.roll(y, -1)) - p[1] * y ** 3 + np.sin(t)


def test_user_cuda_and_traced_python_rhs_and_the_solver_api():
    """NVRTC user source (nbk_rhs_i) and a traced NumPy callable run in the sub-warp kernels; step / skip / resumed run /
    interpolate / t_bound behave as in the other regimes; the CTA kernels (NBK_NO_GROUP=1) give the same answer."""
    n, B = 20, 500
    rng = np.random.default_rng(8)
    y0, p = rng.uniform(-1, 1, (B, n)), rng.uniform(0.2, 1.0, (B, 2))
    ts = [0.5, 2.0, 6.0]
    a = nbk.RungeKutta45(nbk.CudaRHS(USER_RING), 0.0, y0, p, rtol=RTOL, atol=ATOL)
    b = nbk.RungeKutta45(ring, 0.0, y0, p, rtol=RTOL, atol=ATOL)
    _, ya = a.run(ts)
    _, yb = b.run(ts)
    for s in (a, b):
        assert s.is_jit and s._regime == 1 and s.last_launch()["lanes_per_trajectory"] == 8
    ynp, nacc = [], []
    for i in range(8):
        o = onp.AdaptiveRK("RungeKutta45", ring, 0.0, y0[i], p[i], rtol=RTOL, atol=ATOL)
        ynp.append(onp.run(o, ts)[1])
        nacc.append(o.n_accepted)
    assert np.abs(a.n_accepted[:8] - np.array(nacc)).max() <= 1 and np.abs(b.n_accepted[:8] - np.array(nacc)).max() <= 1
    parity_bar("group/user_cuda_rhs/n20/gpu_vs_numpy_restatement", ya[:8], np.array(ynp), RTOL, ATOL, 1.0)
    parity_bar("group/traced_python_rhs/n20/gpu_vs_numpy_restatement", yb[:8], np.array(ynp), RTOL, ATOL, 1.0)
    os.environ["NBK_NO_GROUP"] = "1"
    try:
        c = nbk.RungeKutta45(nbk.CudaRHS(USER_RING), 0.0, y0, p, rtol=RTOL, atol=ATOL)
    finally:
        del os.environ["NBK_NO_GROUP"]
    _, yc = c.run(ts)
    assert "lanes_per_trajectory" not in c.last_launch()
    parity_bar("group/user_cuda_rhs/n20/group_vs_cta_regime", ya, yc, RTOL, ATOL, 1.0)
    assert np.abs(a.n_accepted - c.n_accepted).max() <= 1
    # the solver API on top of the sub-warp kernels
    d = nbk.RungeKutta45(nbk.CudaRHS(USER_RING), 0.0, y0, p, rtol=RTOL, atol=ATOL)
    t1, y1 = d.step(n=5)
    assert t1.shape == (B, 5) and y1.shape == (B, 5, n)
    o = onp.AdaptiveRK("RungeKutta45", ring, 0.0, y0[0], p[0], rtol=RTOL, atol=ATOL)
    tr, yr = onp.steps(o, n=5)
    assert np.abs(t1[0] - tr).max() < 1e-9 and np.abs(y1[0] - yr).max() < 1e-9
    d.skip(n=3)
    _, yd = d.run([6.0])  # resumed from eight steps
    assert np.array_equal(yd[:, 0], ya[:, 2])  # same kernels, same arithmetic: the split run is bit-identical
    lo, hi = d.cache.ts
    te = float(max(lo.max(), hi.min() - 1e-6))
    if lo.max() <= te <= hi.min():
        assert d.interpolate(te).shape == (B, n)
    e = nbk.RungeKutta45(nbk.CudaRHS(USER_RING), 0.0, y0[0], p[0], rtol=RTOL, atol=ATOL, t_bound=0.3)
    with pytest.raises(ValueError):
        e.run([0.5])
    te_, ye_ = e.step(upto_t=0.3)
    assert len(te_) > 0 and te_[-1] <= 0.3 and e.status == 1


def test_position_invariance_and_failures():
    """A trajectory's result does not depend on its position in the ensemble nor on its neighbours in the warp; a
    non-finite trajectory fails alone."""
    n, B = 24, 1000
    y0, k = decay_ensemble(B, n, seed=5)
    ts = [3.0, 7.0]
    s = nbk.RungeKutta45("decay", 0.0, y0, k, rtol=RTOL, atol=ATOL)
    _, y = s.run(ts)
    perm = np.random.default_rng(0).permutation(B)
    s2 = nbk.RungeKutta45("decay", 0.0, y0[perm], k[perm], rtol=RTOL, atol=ATOL)
    _, y2 = s2.run(ts)
    assert np.array_equal(y2, y[perm]) and np.array_equal(s2.n_accepted, s.n_accepted[perm])
    bad = y0.copy()
    bad[17, 5] = np.nan
    s3 = nbk.RungeKutta45("decay", 0.0, bad, k, rtol=RTOL, atol=ATOL)
    with pytest.raises(RuntimeError):
        s3.run(ts)
    st = s3.status
    assert st[17] == 2 and (np.delete(st, 17) == 0).all()


def test_run_events_in_the_sub_warp_regime():
    """run_events for a mid-size system: traced event callables (non-terminal, directional, terminal) through the
    sub-warp kernels against the NumPy restatement of the reference's run_events, and against the per-thread kernels
    (regime="thread") on the same ensemble: same events in the same order, times to the bisection's resolution."""
    n, B = 12, 300
    y0, k = decay_ensemble(B, n, seed=9)

    def half(t, y):
        return y[0] - 0.25

    def cross(t, y):
        return y[3] - y[7]

    cross.direction = 0

    def low(t, y):
        return y[1] + y[2] - 0.2

    low.terminal = True
    low.direction = -1
    evs = [half, cross, low]
    ts = [1.0, 3.0, 8.0]
    g = nbk.RungeKutta45("decay", 0.0, y0, k, rtol=RTOL, atol=ATOL)
    th = nbk.RungeKutta45("decay", 0.0, y0, k, rtol=RTOL, atol=ATOL, regime="thread")
    assert g._regime == 1 and g._group == 4 and th._regime == 0
    tg, yg, teg, yeg = g.run_events(ts, evs)
    tt, yt, tet, yet = th.run_events(ts, evs)
    assert np.array_equal(g.event_counts, th.event_counts) and g.event_counts.sum() > B
    assert np.array_equal(g.status, th.status) and (g.status == -1).sum() > 0
    assert np.array_equal(np.isnan(yg), np.isnan(yt)) and np.nanmax(np.abs(yg - yt)) < 1e-6
    for j in range(3):
        assert np.array_equal(np.isnan(teg[j]), np.isnan(tet[j]))
        if teg[j].size:
            assert np.nanmax(np.abs(teg[j] - tet[j])) < 1e-6 and np.nanmax(np.abs(yeg[j] - yet[j])) < 1e-5
    for b in range(6):
        o = onp.AdaptiveRK("RungeKutta45", onp.rhs_decay, 0.0, y0[b], k[b], rtol=RTOL, atol=ATOL)
        t_ref, y_ref, te_ref, ye_ref = onp.run_events(o, ts, evs)
        for j in range(3):
            assert g.event_counts[b, j] == len(te_ref[j])
            if len(te_ref[j]):
                assert np.abs(teg[j][b, :len(te_ref[j])] - np.array(te_ref[j])).max() < 1e-6
    # the stencil right-hand side (stage buffers in use next to the event buffer), user CUDA C events
    y0r, pr = onp.rd1d_ensemble(200, n=24, seed=3)
    ev = nbk.CudaEvents("NBK_EVENT nbk_event(int k, double t, const double* y, const double* p) { return y[5] - y[17]; }", 1)
    r = nbk.RungeKutta45("rd1d", 0.0, y0r, pr, rtol=RTOL, atol=ATOL)
    assert r._group == 8
    tr_, yr_, ter, yer = r.run_events([0.25, 0.5], ev)
    _, yplain = nbk.RungeKutta45("rd1d", 0.0, y0r, pr, rtol=RTOL, atol=ATOL).run([0.25, 0.5])
    assert np.array_equal(yr_, yplain)  # a non-terminal event does not change the trajectory
    cnt = r.event_counts
    assert cnt.shape == (200, 1)
    for b in np.nonzero(cnt[:, 0] > 0)[0][:20]:
        for i in range(cnt[b, 0]):
            assert abs(yer[0][b, i, 5] - yer[0][b, i, 17]) < 1e-7  # the recorded state is a root of the event function


@pytest.mark.parametrize("cls,n,kw", [("AdamsBashforth4", 24, dict(h=0.002)), ("AdamsBashforth2", 100, dict(h=0.002)), ("ForwardEuler", 50, dict(h=0.001)),
                                      ("RungeKutta4", 20, dict(h=0.005)), ("Heun3", 64, dict(h=0.005)),
                                      ("DOP853", 32, dict(rtol=1e-7, atol=1e-10)), ("DOP853", 13, dict(rtol=1e-6, atol=1e-9))])
def test_other_classes_as_sub_warp_teams(cls, n, kw):
    """The classes without a dedicated sub-warp kernel (DOP853, Adams-Bashforth, ForwardEuler, fixed-step Runge-Kutta) run
    the CTA regime's source with G lanes per trajectory (NBK_TEAM_G): against the C port, and against the same classes
    with one CTA per trajectory (NBK_NO_TEAM=1) -- fixed step 1e-12 (against both), DOP853 within the tolerance with step
    counts within +-1."""
    B = 1500
    y0, p = onp.rd1d_ensemble(B, n=n, seed=3)
    p[:, 0] *= 0.05 * (100.0 / n) ** 2
    ts = [0.05, 0.2, 0.5]
    adaptive = cls == "DOP853"
    ref = oc.run_ensemble(cls, "rd1d", y0, p, ts, nthreads=8, **kw)
    s = getattr(nbk, cls)("rd1d", 0.0, y0, p, **kw)
    t, y = s.run(ts)
    assert s._regime == 1 and s._group in (4, 8, 16, 32) and s.last_launch()["block"] == 128
    os.environ["NBK_NO_TEAM"] = "1"
    try:
        c = getattr(nbk, cls)("rd1d", 0.0, y0, p, **kw)
    finally:
        del os.environ["NBK_NO_TEAM"]
    _, yc = c.run(ts)
    assert c._group == 0
    if adaptive:
        parity_bar(f"group/team/{cls}/n{n}/gpu_vs_c_port", y, ref["ys"], kw["rtol"], kw["atol"], 1.0)
        parity_bar(f"group/team/{cls}/n{n}/team_vs_cta", y, yc, kw["rtol"], kw["atol"], 1.0)
        assert np.abs(s.n_accepted - ref["n_accepted"]).max() <= 1 and np.abs(s.n_accepted - c.n_accepted).max() <= 1
    else:
        scale = max(1.0, np.abs(ref["ys"]).max())
        assert np.abs(y - ref["ys"]).max() <= 1e-12 * scale
        assert np.abs(y - yc).max() <= 1e-12 * scale and np.abs(s.y - c.y).max() <= 1e-12 * scale and np.abs(s.f - c.f).max() <= 1e-11 * scale
    # step / resumed run on top
    t1, y1 = s.step(n=3)
    tc, yc1 = c.step(n=3)
    assert t1.shape == (B, 3) and np.abs(y1 - yc1).max() < (1e-6 if adaptive else 1e-12 * max(1.0, np.abs(yc1).max()))
