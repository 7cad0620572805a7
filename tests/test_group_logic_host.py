"""Kernel-logic tests of the sub-warp regime without a GPU: csrc/nbk_group.cuh compiled for the host, the run / step kernels
emulated as one warp of 32 POSIX threads (tests/emul/nbk_emul_group.cpp: group barrier = pthread barrier of the group's
lanes, group shuffles = exchange between two such barriers, real atomics on the trajectory queue), against the NumPy
restatement of the reference.  Covers what neither the single-lane emulation of the per-thread kernels nor the block
emulation of the CTA kernels can: several trajectories in flight in one warp, each group taking its next trajectory from
the queue on its own, group-local stage buffers, ragged groups (n < G E), the warp's exit vote.  GPU-only aspects (FMA
contraction, MUFU seeds, real warps) are covered by the -m gpu parity tests (tests/test_gpu_group.py)."""

import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
from emul import EmulGroupEnsemble, group_shape  # noqa: E402

from oracle import nbk_oracle_np as onp  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b))) / max(1.0, float(np.max(np.abs(b)))) if a.size else 0.0


def ensemble(B, n, seed, scale=0.05):
    y0, p = onp.rd1d_ensemble(B, n=n, seed=seed)
    p[:, 0] *= scale * (100.0 / n) ** 2  # a non-stiff grid: rounding differences stay at rounding level
    return y0, p


def test_shape_rule_matches_the_launcher():
    """emul.group_shape restates group_shape() of csrc/nbk_api.cu: the smallest power-of-two group with at most four
    components per lane; the construction kernel's block is the CTA rule for that E."""
    assert group_shape(3) == (2, 2, 32)
    assert group_shape(8) == (2, 4, 32)
    assert group_shape(9) == (4, 3, 32)
    assert group_shape(16) == (4, 4, 32)
    assert group_shape(24) == (8, 3, 32)
    assert group_shape(32) == (8, 4, 32)
    assert group_shape(33) == (16, 3, 32)
    assert group_shape(64) == (16, 4, 32)
    assert group_shape(100) == (32, 4, 32)


@pytest.mark.parametrize("cls,n,B", [("RungeKutta45", 24, 9), ("RungeKutta45", 32, 6), ("RungeKutta45", 13, 19), ("RungeKutta23", 50, 5),
                                     ("RungeKutta23", 7, 21), ("RungeKutta45", 64, 3), ("RungeKutta45", 100, 3)])
def test_group_kernels_vs_numpy_restatement(cls, n, B):
    """run() with dense output, a resumed run, step() and the t_bound guard; B exceeds the 32 / G groups of the warp in most
    cases (queue refill per group), ragged last lanes (n = 13, 50, 7, 100), full groups (n = 24, 32, 64)."""
    y0, p = ensemble(B, n, seed=21)
    kw = dict(rtol=1e-6, atol=1e-9)
    g = EmulGroupEnsemble(cls, "rd1d", y0, p, **kw)
    ts = [0.4, 1.0, 2.5]
    out = g.run(ts[:2])
    out2 = g.run(ts[2:])
    assert (g.status == 0).all()
    for b in range(B):
        o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], **kw)
        _, yo = onp.run(o, ts)
        got = np.concatenate([out[b], out2[b]])
        assert np.all(np.abs(got - yo) <= 1e-3 * (1e-6 * np.abs(yo) + 1e-9)), (cls, n, b)
        assert g.nacc[b] == o.n_accepted and g.nrej[b] == o.n_rejected
        assert rel(g.t[b], o.t) < 1e-9 and rel(g.y[b], o.y) < 1e-9
    s = EmulGroupEnsemble(cls, "rd1d", y0, p, **kw)
    t_out, y_out, counts = s.step(4)
    for b in range(B):
        o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], **kw)
        to, yo = onp.steps(o, n=4)
        assert counts[b] == 4 and rel(t_out[b], to) < 1e-9 and rel(y_out[b], yo) < 1e-9
    q = EmulGroupEnsemble(cls, "rd1d", y0, p, **kw)
    outb = q.run([0.2, 5.0], t_bound=0.6)
    assert (q.status == 1).all() and np.isnan(outb[:, 1]).all() and (q.t <= 0.6).all()


@pytest.mark.parametrize("rhs,n", [("decay", 16), ("decay1", 20)])
def test_elementwise_rhs_skips_the_stage_buffers(rhs, n):
    """The component-wise decay needs no neighbour (LOCAL): stages are not staged in shared memory at all."""
    B = 12
    rng = np.random.default_rng(4)
    y0 = rng.uniform(0.5, 2, (B, n))
    k = rng.uniform(0.1, 1.0, (B, n if rhs == "decay" else 1))
    g = EmulGroupEnsemble("RungeKutta45", rhs, y0, k, rtol=1e-7, atol=1e-10)
    out = g.run([1.0, 3.0])
    for b in range(B):
        o = onp.AdaptiveRK("RungeKutta45", onp.rhs_decay, 0.0, y0[b], k[b], rtol=1e-7, atol=1e-10)
        _, yo = onp.run(o, [1.0, 3.0])
        assert np.all(np.abs(out[b] - yo) <= 1e-3 * (1e-7 * np.abs(yo) + 1e-10))
        assert g.nacc[b] == o.n_accepted and g.nrej[b] == o.n_rejected


@pytest.mark.parametrize("cls,n,lanes", [("RungeKutta45", 24, None), ("RungeKutta23", 13, None), ("RungeKutta45", 40, None),
                                         ("RungeKutta45", 24, 4), ("RungeKutta45", 12, 32)])
def test_group_kernels_are_race_free_under_real_thread_scheduling(cls, n, lanes):
    """The 32 threads of the emulated warp are scheduled by the OS, far from lock-step: a missing group barrier (stage
    buffers reused across stages and attempts, the queue broadcast) shows up as a run-to-run difference.  Repetitions of
    run (several grid points per step, resumed) + step must be bit-identical; forced shapes: six components per lane
    (lanes = 4) and a whole warp for a small system (lanes = 32, most lanes own nothing)."""
    B = 11
    y0, p = ensemble(B, n, seed=5)
    kw = dict(rtol=1e-6, atol=1e-9)
    ref = None
    for rep in range(4):
        g = EmulGroupEnsemble(cls, "rd1d", y0, p, lanes=lanes, **kw)
        got = [g.run([0.1, 0.11, 0.5]), g.run([0.9]), *g.step(3)[:2], g.t, g.y, g.f, g.nacc.copy(), g.nrej.copy()]
        if ref is None:
            ref = got
        else:
            for a, b in zip(got, ref):
                assert np.array_equal(a, b)
    o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[3], p[3], **kw)
    _, yo = onp.run(o, [0.1, 0.11, 0.5, 0.9])
    assert np.all(np.abs(np.concatenate([ref[0][3], ref[1][3]]) - yo) <= 1e-3 * (1e-6 * np.abs(yo) + 1e-9))


def test_traced_python_rhs_in_the_group_kernels():
    """A NumPy-style callable traced to the component-wise form (rhs.trace_cta), spliced in like the NVRTC path."""
    from nbkode_b200 import rhs as _rhs

    n, B = 20, 7

    def f(t, y, p):
        return p[0] * (np.roll(y, 1) - 2 * y + np.roll(y, -1)) - p[1] * y ** 3 + np.sin(t)

    src = _rhs.trace_cta(f, n, (2,), True).source
    rng = np.random.default_rng(8)
    y0, p = rng.uniform(-1, 1, (B, n)), rng.uniform(0.2, 1.0, (B, 2))
    g = EmulGroupEnsemble("RungeKutta45", src, y0, p, rtol=1e-6, atol=1e-9)
    out = g.run([0.5, 2.0])
    assert (g.status == 0).all()
    for b in range(B):
        o = onp.AdaptiveRK("RungeKutta45", f, 0.0, y0[b], p[b], rtol=1e-6, atol=1e-9)
        _, yo = onp.run(o, [0.5, 2.0])
        assert np.all(np.abs(out[b] - yo) <= 1e-3 * (1e-6 * np.abs(yo) + 1e-9))
        assert g.nacc[b] == o.n_accepted and g.nrej[b] == o.n_rejected


def test_failure_paths_of_one_group_leave_its_neighbours_alone():
    """A trajectory that goes non-finite retires with NBK_TRAJ_NONFINITE and NaN outputs while the other groups of the
    warp finish; the status is sticky in a second call; an attempt budget ends a call with NBK_TRAJ_MAXATTEMPTS."""
    n, B = 16, 9
    y0, p = ensemble(B, n, seed=2)
    y0[4, 3] = np.nan
    g = EmulGroupEnsemble("RungeKutta45", "rd1d", y0, p, rtol=1e-6, atol=1e-9)
    out = g.run([0.3, 0.6])
    assert g.status[4] == 2 and (np.delete(g.status, 4) == 0).all()
    assert np.isnan(out[4]).all() and np.isfinite(np.delete(out, 4, axis=0)).all()
    assert g.counter[1] == 1
    out2 = g.run([0.9])
    assert g.status[4] == 2 and np.isnan(out2[4]).all() and np.isfinite(np.delete(out2, 4, axis=0)).all()
    ok = EmulGroupEnsemble("RungeKutta45", "rd1d", np.delete(y0, 4, axis=0), np.delete(p, 4, axis=0), rtol=1e-6, atol=1e-9)
    assert np.array_equal(np.concatenate([ok.run([0.3, 0.6]), ok.run([0.9])], axis=1),
                          np.delete(np.concatenate([out, out2], axis=1), 4, axis=0))


def _events(n, terminal_at=None):
    def crossing(t, y):
        return y[0] - 0.5

    def gap(t, y):
        return y[n // 2] - y[n // 2 + 2]

    gap.direction = -1

    def level(t, y):
        return y[1] + y[2] - (1.0 if terminal_at is None else terminal_at)

    level.terminal = terminal_at is not None
    return [crossing, gap, level]


@pytest.mark.parametrize("cls,n,terminal_at", [("RungeKutta45", 12, None), ("RungeKutta45", 12, 0.9), ("RungeKutta23", 20, None),
                                               ("RungeKutta45", 7, 1.1)])
def test_run_events_in_the_group_kernels(cls, n, terminal_at):
    """run_events in the sub-warp regime (the group's lanes publish the whole vector for the event functions, every lane
    evaluates them): same events in the same order as the NumPy restatement of the reference's run_events, event times to
    the bisection's resolution, terminal events truncate the output; bit-identical across OS-scheduled repetitions."""
    from nbkode_b200 import rhs as nrhs

    B = 9
    y0, p = ensemble(B, n, seed=17, scale=0.4)
    kw = dict(rtol=1e-6, atol=1e-9)
    evs = _events(n, terminal_at)
    ev = nrhs.trace_events(evs, n, (2,), True)
    ts = [0.5, 1.5, 3.0]
    first = None
    for rep in range(3):
        g = EmulGroupEnsemble(cls, "rd1d", y0, p, **kw)
        got = g.run_events(ts, ev)
        if first is None:
            first = (g, got)
        else:
            for a, b in zip(got, first[1]):
                assert np.array_equal(a, b, equal_nan=True)
    g, (out, ev_t, ev_y, ev_c, t_term) = first
    n_events = 0
    for b in range(B):
        o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], **kw)
        t_ref, y_ref, te_ref, ye_ref = onp.run_events(o, ts, evs)
        k = int(g.nout[b])
        assert k == len(t_ref)
        if g.status[b] == -1:
            assert rel(t_term[b], t_ref[-1]) < 1e-7 and rel(out[b, k - 1], y_ref[-1]) < 1e-6
            if k > 1:
                assert rel(out[b, :k - 1], y_ref[:k - 1]) < 1e-7
        else:
            assert g.status[b] == 0 and rel(out[b, :k], y_ref) < 1e-7
        assert np.isnan(out[b, k:]).all()
        for j in range(3):
            assert ev_c[b, j] == len(te_ref[j])
            n_events += len(te_ref[j])
            if len(te_ref[j]):
                assert rel(ev_t[b, j, :len(te_ref[j])], te_ref[j]) < 1e-6
                assert rel(ev_y[b, j, :len(te_ref[j])], np.array(ye_ref[j])) < 1e-5
    assert n_events > 0


# ---------------------------------------------------------------------------------------------------------------------
# The classes without a dedicated sub-warp kernel run the CTA regime's source with the team = G lanes of a warp
# (NBK_TEAM_G, csrc/nbk_cta.cuh): construction, run and step kernels all in the emulated warp.
# ---------------------------------------------------------------------------------------------------------------------
def _reference(cls, y0, p, kw):
    if cls in ("RungeKutta23", "RungeKutta45", "DOP853"):
        return onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0, p, **kw)
    if cls in ("RungeKutta4", "Heun3", "Runge2", "Runge3", "RungeKutta3_8"):
        return onp.FixedRK(cls, onp.rhs_rd1d, 0.0, y0, p, h=kw["h"])
    order = 1 if cls == "ForwardEuler" else int(cls[-1])
    return onp.AdamsBashforth(order, onp.rhs_rd1d, 0.0, y0, p, h=kw["h"],
                              first_stepper={45: "RungeKutta45", 23: "RungeKutta23", 0: None}[kw.get("first_stepper", 45)])


@pytest.mark.parametrize("cls,n,kw", [("AdamsBashforth4", 24, dict(h=0.01, first_stepper=45)), ("AdamsBashforth5", 13, dict(h=0.01, first_stepper=23)),
                                      ("AdamsBashforth2", 50, dict(h=0.01, first_stepper=45)), ("ForwardEuler", 32, dict(h=0.01, first_stepper=0)),
                                      ("RungeKutta4", 20, dict(h=0.01)), ("Heun3", 7, dict(h=0.01)),
                                      ("DOP853", 20, dict(rtol=1e-6, atol=1e-9)), ("DOP853", 64, dict(rtol=1e-7, atol=1e-10))])
def test_team_mode_classes_vs_numpy_restatement(cls, n, kw):
    """Adams-Bashforth (ring-buffered window, RungeKutta45 / RungeKutta23 / no start-up), ForwardEuler, fixed-step Runge-
    Kutta and DOP853 (two-norm error estimate, dense output with three extra stages behind team barriers) as sub-warp
    teams: dense output on a grid that straddles steps, a resumed run, step(); more trajectories than teams in the warp;
    fixed step to 1e-12, DOP853 to 1e-3 of the tolerance with identical step counts."""
    B = 11
    y0, p = ensemble(B, n, seed=33)
    adaptive = cls == "DOP853"
    g = EmulGroupEnsemble(cls, "rd1d", y0, p, **kw)
    assert g._shape[0] * g._shape[1] >= n and g._shape[1] <= (2 if adaptive else 4)
    ts = np.array([0.005, 0.0312, 0.1, 0.1049, 0.25])
    out = np.concatenate([g.run(ts[:3]), g.run(ts[3:])], axis=1)
    tq, yq, counts = g.step(5)
    assert (g.status == 0).all() and (counts == 5).all()
    for b in range(B):
        o = _reference(cls, y0[b], p[b], kw)
        _, yo = onp.run(o, ts)
        to, ys = onp.steps(o, n=5)
        if adaptive:
            assert np.all(np.abs(out[b] - yo) <= 1e-3 * (kw["rtol"] * np.abs(yo) + kw["atol"]))
            assert g.nacc[b] == o.n_accepted and rel(tq[b], to) < 1e-6 and rel(yq[b], ys) < 1e-6
        else:
            assert rel(out[b], yo) < 1e-12, (cls, n, b, rel(out[b], yo))
            assert rel(tq[b], to) < 1e-12 and rel(yq[b], ys) < 1e-12
            assert rel(g.t[b], o.t) < 1e-12 and rel(g.y[b], o.y) < 1e-12 and rel(g.f[b], o.f) < 1e-11


@pytest.mark.parametrize("cls,n,kw", [("AdamsBashforth4", 24, dict(h=0.01, first_stepper=45)), ("RungeKutta4", 13, dict(h=0.01)),
                                      ("DOP853", 20, dict(rtol=1e-6, atol=1e-9))])
def test_team_mode_is_race_free_and_dop853_events(cls, n, kw):
    """Repetitions under OS scheduling are bit-identical (every barrier names exactly the team's lanes); DOP853 also runs
    run_events through the team routine (its dense output prepares three extra stages inside the bisection)."""
    B = 9
    y0, p = ensemble(B, n, seed=5, scale=0.4 if cls == "DOP853" else 0.05)
    ref = None
    for rep in range(3):
        g = EmulGroupEnsemble(cls, "rd1d", y0, p, **kw)
        got = [g.run([0.1, 0.11, 0.5]), g.run([0.9]), *g.step(3)[:2], g.t, g.y, g.f]
        if ref is None:
            ref = got
        else:
            for a, b in zip(got, ref):
                assert np.array_equal(a, b)
    if cls == "DOP853":
        from nbkode_b200 import rhs as nrhs

        evs = _events(n, 0.95)
        ev = nrhs.trace_events(evs, n, (2,), True)
        g = EmulGroupEnsemble(cls, "rd1d", y0, p, **kw)
        out, ev_t, ev_y, ev_c, t_term = g.run_events([0.5, 1.5, 3.0], ev)
        total = 0
        for b in range(B):
            o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], **kw)
            t_ref, y_ref, te_ref, ye_ref = onp.run_events(o, [0.5, 1.5, 3.0], evs)
            assert int(g.nout[b]) == len(t_ref) and (g.status[b] == -1) == (len(te_ref[2]) > 0)
            for j in range(3):
                assert ev_c[b, j] == len(te_ref[j])
                total += len(te_ref[j])
                if len(te_ref[j]):
                    assert rel(ev_t[b, j, :len(te_ref[j])], te_ref[j]) < 1e-6
        assert total > 0


@pytest.mark.parametrize("cls,n,terminal_at", [("RungeKutta4", 13, None), ("AdamsBashforth2", 24, 0.9), ("ForwardEuler", 40, None)])
def test_run_events_of_the_fixed_step_classes_as_sub_warp_teams(cls, n, terminal_at):
    """run_events of the fixed-step classes with G lanes per trajectory (cta_advance_fixed<..., EV> behind NBK_TEAM_G): several
    teams of one warp detect, bisect and record their own events at different steps; against the restatement of the
    reference's run_events, bit-identical across OS-scheduled repetitions."""
    from nbkode_b200 import rhs as nrhs

    evs = _events(n, terminal_at)
    ev = nrhs.trace_events(evs, n, (2,), True)
    B, h, ts = 9, 0.01, [0.5, 1.5, 3.0]
    y0, p = ensemble(B, n, seed=5, scale=0.4)
    kw = dict(h=h) if cls == "RungeKutta4" else dict(h=h, first_stepper=0)
    first = None
    for rep in range(2):
        g = EmulGroupEnsemble(cls, "rd1d", y0, p, **kw)
        got = g.run_events(ts, ev)
        if first is None:
            first = (g, got)
        else:
            for a, b in zip(got, first[1]):
                assert np.array_equal(a, b, equal_nan=True)
    g, (out, ev_t, ev_y, ev_c, t_term) = first
    total = 0
    for b in range(B):
        if cls == "RungeKutta4":
            o = onp.FixedRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], h=h)
        else:
            o = onp.AdamsBashforth(1 if cls == "ForwardEuler" else int(cls[-1]), onp.rhs_rd1d, 0.0, y0[b], p[b], h=h, first_stepper=None)
        t_ref, y_ref, te_ref, ye_ref = onp.run_events(o, ts, evs)
        k = int(g.nout[b])
        assert k == len(t_ref) and (g.status[b] == -1) == (terminal_at is not None and len(te_ref[2]) > 0)
        keep = k - 1 if g.status[b] == -1 else k
        if keep:
            assert rel(out[b, :keep], y_ref[:keep]) < 1e-12
        for j in range(3):
            assert ev_c[b, j] == len(te_ref[j])
            total += len(te_ref[j])
            if len(te_ref[j]):
                assert rel(ev_t[b, j, :len(te_ref[j])], te_ref[j]) < 1e-7
    assert total > 0
