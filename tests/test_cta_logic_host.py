"""Kernel-logic tests of the CTA-per-trajectory regime without a GPU: csrc/nbk_cta.cuh + nbk_program_cta.cu compiled
for the host, a thread block emulated by T POSIX threads (tests/emul/nbk_emul_cta.cpp: __syncthreads = barrier, warp
shuffles = exchange through a scratch array), against the NumPy restatement of the reference.  Covers what the
single-lane emulation of the per-thread kernels cannot: stage arguments through "shared memory", halo reads, block
reductions, ragged and "full" shapes, the ring-buffered multistep window, DOP853's two-norm error estimate and its
dense output evaluated by the whole block.  GPU-only aspects (FMA contraction, MUFU seeds, real warps) are covered by
the -m gpu parity tests."""

import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
from emul import EmulCtaEnsemble, EmulMmaEnsemble, cta_shape  # noqa: E402

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
    """emul.cta_shape restates cta_shape() / cta_block() of csrc/nbk_api.cu (components per thread, block, full)."""
    assert cta_shape(45, "rd1d", 48) == (1, 64, 0)
    assert cta_shape(45, "rd1d", 64) == (1, 64, 1)
    assert cta_shape(45, "rd1d", 100) == (2, 64, 0)
    assert cta_shape(45, "rd1d", 128) == (4, 32, 1)
    assert cta_shape(45, "rd1d", 1000) == (4, 256, 0)
    assert cta_shape(45, "rd1d", 2048) == (4, 512, 1)
    assert cta_shape(4, "rd1d", 2048) == (4, 512, 0)      # fixed step: the any-n program
    assert cta_shape(853, "rd1d", 512) == (1, 512, 1) and cta_shape(853, "rd1d", 700) == (2, 352, 0)
    assert cta_shape(853, "rd1d", 2048) == (4, 512, 1) and cta_shape(853, "rd1d", 1100) == (4, 288, 0)
    assert cta_shape(45, "linear", 128) == (1, 128, 1)
    assert cta_shape(45, "rd1d", 3000) == (12, 256, 0)


@pytest.mark.parametrize("cls,n", [("RungeKutta45", 48), ("RungeKutta45", 100), ("RungeKutta45", 128), ("RungeKutta23", 128),
                                   ("RungeKutta23", 70), ("DOP853", 100), ("DOP853", 96)])
def test_adaptive_cta_vs_numpy_restatement(cls, n):
    """run() with dense output, a resumed run, step() and the t_bound guard for the adaptive classes; ragged shapes
    (n = 48, 70, 100), the "full" programs (n = 96, 128)."""
    B = 3
    y0, p = ensemble(B, n, seed=21)
    kw = dict(rtol=1e-6, atol=1e-9)
    # step SIZES carry the rounding of the error estimate (std::fma in the stencil against NumPy's two roundings);
    # DOP853's E5.K / E3.K cancel more digits than the 5(4) and 3(2) pairs' E.K
    tol_t = 1e-6 if cls == "DOP853" else 1e-9
    g = EmulCtaEnsemble(cls, "rd1d", y0, p, **kw)
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
        assert rel(g.t[b], o.t) < tol_t and rel(g.y[b], o.y) < tol_t
    # step(): every accepted step of each trajectory
    s = EmulCtaEnsemble(cls, "rd1d", y0, p, **kw)
    t_out, y_out, counts = s.step(4)
    for b in range(B):
        o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], **kw)
        to, yo = onp.steps(o, n=4)
        assert counts[b] == 4 and rel(t_out[b], to) < tol_t and rel(y_out[b], yo) < tol_t
    # t_bound: trajectories stop before the first step that would pass it, later grid points stay NaN
    q = EmulCtaEnsemble(cls, "rd1d", y0, p, **kw)
    outb = q.run([0.2, 5.0], t_bound=0.6)
    assert (q.status == 1).all() and np.isnan(outb[:, 1]).all() and (q.t <= 0.6).all()


@pytest.mark.parametrize("cls,n,fs", [("AdamsBashforth4", 100, 45), ("AdamsBashforth4", 128, 0), ("AdamsBashforth3", 48, 45),
                                      ("AdamsBashforth5", 70, 23), ("AdamsBashforth2", 100, 45), ("ForwardEuler", 128, 0),
                                      ("RungeKutta4", 100, 0), ("Heun3", 128, 0)])
def test_fixed_step_cta_vs_numpy_restatement(cls, n, fs):
    """Fixed-step classes in the CTA regime: AdamsBashforth3..5 through the ring-buffered loop (start-up RungeKutta45 /
    RungeKutta23 / none), AdamsBashforth2 / ForwardEuler through the shifting loop, the fixed-step Runge-Kutta classes;
    dense output on a grid that straddles steps, a resumed run and step(), all to 1e-12."""
    B, h = 2, 0.01
    y0, p = ensemble(B, n, seed=33)
    g = EmulCtaEnsemble(cls, "rd1d", y0, p, h=h, first_stepper=fs)
    ts = np.array([0.005, 0.0312, 0.1, 0.1049, 0.25])
    out = np.concatenate([g.run(ts[:3]), g.run(ts[3:])], axis=1)
    tq, yq, counts = g.step(5)
    assert (g.status == 0).all() and (counts == 5).all()
    for b in range(B):
        if cls in ("RungeKutta4", "Heun3"):
            o = onp.FixedRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], h=h)
        else:
            order = 1 if cls == "ForwardEuler" else int(cls[-1])
            o = onp.AdamsBashforth(order, onp.rhs_rd1d, 0.0, y0[b], p[b], h=h,
                                   first_stepper={45: "RungeKutta45", 23: "RungeKutta23", 0: None}[fs])
        _, yo = onp.run(o, ts)
        assert rel(out[b], yo) < 1e-12, (cls, n, fs, b, rel(out[b], yo))
        to, ys = onp.steps(o, n=5)
        assert rel(tq[b], to) < 1e-12 and rel(yq[b], ys) < 1e-12
        assert rel(g.t[b], o.t) < 1e-12 and rel(g.y[b], o.y) < 1e-12 and rel(g.f[b], o.f) < 1e-11


@pytest.mark.parametrize("cls,n,kw", [("RungeKutta45", 100, dict(rtol=1e-6, atol=1e-9)), ("RungeKutta23", 128, dict(rtol=1e-5, atol=1e-8)),
                                      ("DOP853", 100, dict(rtol=1e-6, atol=1e-9)), ("DOP853", 96, dict(rtol=1e-7, atol=1e-10)),
                                      ("AdamsBashforth4", 100, dict(h=0.01)), ("AdamsBashforth2", 70, dict(h=0.01)),
                                      ("RungeKutta4", 128, dict(h=0.01))])
def test_cta_kernels_are_race_free_under_real_thread_scheduling(cls, n, kw):
    """The emulated block's threads are scheduled by the OS, far from lock-step: any missing barrier shows up as a
    run-to-run difference.  Five repetitions of run (several grid points per step and per call, resumed) + step must be
    bit-identical, and must equal a run at one output point per call."""
    B = 3
    y0, p = ensemble(B, n, seed=5)
    ts = np.array([0.05, 0.06, 0.3, 0.31, 0.32, 0.9])
    results = []
    for _ in range(5):
        g = EmulCtaEnsemble(cls, "rd1d", y0, p, **kw)
        out = np.concatenate([g.run(ts[:4]), g.run(ts[4:])], axis=1)
        tq, yq, _ = g.step(3)
        results.append((out, tq, yq, g.y, g.h.copy(), g.K.copy()))
    for r in results[1:]:
        for a, b in zip(results[0], r):
            assert np.array_equal(a, b, equal_nan=True)
    one = EmulCtaEnsemble(cls, "rd1d", y0, p, **kw)
    single = np.concatenate([one.run([t]) for t in ts], axis=1)
    assert np.array_equal(single, results[0][0])


def test_linear_rhs_cta_vs_numpy_restatement():
    """y' = A y with one shared dense A through the generic CTA kernel (one component per thread, the whole vector read
    from "shared memory" by every thread)."""
    B, n = 3, 40
    y0, A = onp.linear_ensemble(B, n=n, seed=2)
    kw = dict(rtol=1e-6, atol=1e-9)
    g = EmulCtaEnsemble("RungeKutta45", "linear", y0, A.T.copy(), params_shared=True, **kw)  # the kernel takes A transposed
    ts = [0.5, 1.0, 4.0]
    out = g.run(ts)
    for b in range(B):
        o = onp.AdaptiveRK("RungeKutta45", onp.rhs_linear, 0.0, y0[b], A, **kw)
        _, yo = onp.run(o, ts)
        assert np.all(np.abs(out[b] - yo) <= 1e-2 * (1e-6 * np.abs(yo) + 1e-9)) and g.nacc[b] == o.n_accepted


def test_traced_python_rhs_through_the_emulated_block():
    """A NumPy-style Python right-hand side -> vector tracer -> nbk_rhs_i -> the user path of the CTA kernels (the
    parameter block held in registers, NBK_CTA_NP), spliced in as csrc/nbk_jit.cpp does for NVRTC: the same results
    as the built-in right-hand side, for an adaptive and a fixed-step class, plus operations only the tracer has."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "numbakit-ode_b200"))
    from nbkode_b200 import rhs

    def rd(t, y, p):
        return p[0] * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + p[1] * y * (1 - y)

    B, n = 2, 100
    y0, p = ensemble(B, n, seed=8)
    src = rhs.trace_cta(rd, n, (2,), True).source
    ts = [0.3, 1.1]
    for cls, kw, tol in (("RungeKutta45", dict(rtol=1e-6, atol=1e-9), 1e-9), ("AdamsBashforth4", dict(h=0.01), 1e-12)):
        a = EmulCtaEnsemble(cls, src, y0, p, **kw)
        b = EmulCtaEnsemble(cls, "rd1d", y0, p, **kw)
        assert (a.E, a.T) == (b.E, b.T)
        assert rel(a.run(ts), b.run(ts)) < tol and np.array_equal(a.nacc, b.nacc)

    def limited(t, y, p):  # comparisons / np.where / np.diff / a reduction (a loop in the source: one component per thread)
        d = np.diff(np.append(y, y[0]))
        return p[0] * np.where(d > 0, d, 0.5 * d) - p[1] * (y - y.mean())

    src2 = rhs.trace_cta(limited, n, (2,), True).source
    c = EmulCtaEnsemble("RungeKutta23", src2, y0, p, rtol=1e-6, atol=1e-9)
    assert c.E == 1 and "for (" in src2
    out = c.run(ts)
    for k in range(B):
        o = onp.AdaptiveRK("RungeKutta23", limited, 0.0, y0[k], p[k], rtol=1e-6, atol=1e-9)
        _, yo = onp.run(o, ts)
        assert np.all(np.abs(out[k] - yo) <= 1e-2 * (1e-6 * np.abs(yo) + 1e-9)) and c.nacc[k] == o.n_accepted


def test_dmma_tile_kernel_k4_on_the_host():
    """K4 (csrc/nbk_mma.cuh) with the mma.sync fragment exchange emulated per warp: a partial tile (5 of 16 columns),
    n = 40 zero-padded to 128, per-column controllers and emission from the accumulator layout, against the NumPy
    restatement; two OS-scheduled repetitions must be bit-identical (race detector); a resumed run continues."""
    B, n = 5, 40
    y0, A = onp.linear_ensemble(B, n=n, seed=2)
    kw = dict(rtol=1e-6, atol=1e-9)
    ts = [0.2, 0.6]
    runs = []
    for _ in range(2):
        g = EmulMmaEnsemble("RungeKutta45", y0, A, **kw)
        runs.append((g.run(ts[:1]), g.run(ts[1:]), g.y, g.h.copy(), g.nacc.copy(), g.K.copy()))
    for a, b in zip(*runs):
        assert np.array_equal(a, b)
    out = np.concatenate(runs[0][:2], axis=1)
    for b in range(B):
        o = onp.AdaptiveRK("RungeKutta45", onp.rhs_linear, 0.0, y0[b], A, **kw)
        _, yo = onp.run(o, ts)
        assert np.all(np.abs(out[b] - yo) <= 1e-3 * (1e-6 * np.abs(yo) + 1e-9))
        assert runs[0][4][b] == o.n_accepted and rel(runs[0][2][b], o.y) < 1e-7 and rel(runs[0][3][b], o.h) < 1e-6  # summation order of A.y


@pytest.mark.parametrize("cls,n", [("RungeKutta45", 130), ("RungeKutta45", 300), ("RungeKutta23", 2100), ("AdamsBashforth3", 300),
                                   ("DOP853", 1100), ("DOP853", 2100)])
def test_more_cta_shapes_on_the_host(cls, n):
    """Further shapes: four components per thread with a ragged last thread (n = 130, 300), the large-n shape
    (n = 2100: nine components per thread, 256 threads), DOP853 beyond n = 1024 (four components per thread, and the
    large-n shape), against the NumPy restatement and repeated for determinism."""
    B = 2
    y0, p = ensemble(B, n, seed=13)
    kw = dict(h=0.01) if cls.startswith("Adams") else dict(rtol=1e-6, atol=1e-9)
    ts = [0.15, 0.4]
    outs = []
    for _ in range(2):
        g = EmulCtaEnsemble(cls, "rd1d", y0, p, **kw)
        outs.append((g.run(ts), g.y, g.nacc.copy()))
    assert all(np.array_equal(a, b) for a, b in zip(*outs))
    for b in range(B):
        if cls.startswith("Adams"):
            o = onp.AdamsBashforth(int(cls[-1]), onp.rhs_rd1d, 0.0, y0[b], p[b], h=0.01)
            _, yo = onp.run(o, ts)
            assert rel(outs[0][0][b], yo) < 1e-12
        else:
            o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], **kw)
            _, yo = onp.run(o, ts)
            assert np.all(np.abs(outs[0][0][b] - yo) <= 1e-3 * (1e-6 * np.abs(yo) + 1e-9)) and outs[0][2][b] == o.n_accepted


@pytest.mark.parametrize("cls", ["RungeKutta45", "DOP853", "AdamsBashforth4"])
def test_failure_paths_of_the_cta_kernels(cls):
    """A trajectory whose right-hand side turns non-finite (NaN parameter) fails alone: status NONFINITE for the
    adaptive classes (the reference would spin forever), NaN rows in the output, the neighbours untouched; the status is
    sticky across calls; the attempt budget stops a run with MAXATTEMPTS."""
    B, n = 3, 100
    y0, p = ensemble(B, n, seed=17)
    p[1, 1] = np.nan
    kw = dict(h=0.01) if cls.startswith("Adams") else dict(rtol=1e-6, atol=1e-9)
    g = EmulCtaEnsemble(cls, "rd1d", y0, p, **kw)
    out = g.run([0.2, 0.5])
    good = [0, 2]
    assert np.isfinite(out[good]).all() and (g.status[good] == 0).all()
    if cls.startswith("Adams"):
        # a fixed-step solver has no error estimate to notice: the start-up (adaptive RungeKutta45) fails the trajectory
        assert g.status[1] == 2 and np.isnan(out[1]).all()
    else:
        assert g.status[1] == 2 and np.isnan(out[1]).all() and g.counter[1] >= 1
    out2 = g.run([0.6])
    assert g.status[1] == 2 and np.isnan(out2[1]).all() and np.isfinite(out2[good]).all()
    for b in good:
        q = dict(kw)
        if cls.startswith("Adams"):
            o = onp.AdamsBashforth(4, onp.rhs_rd1d, 0.0, y0[b], p[b], h=0.01)
            _, yo = onp.run(o, [0.2, 0.5, 0.6])
            assert rel(np.concatenate([out[b], out2[b]]), yo) < 1e-12
        else:
            o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], **q)
            _, yo = onp.run(o, [0.2, 0.5, 0.6])
            assert np.all(np.abs(np.concatenate([out[b], out2[b]]) - yo) <= 1e-3 * (1e-6 * np.abs(yo) + 1e-9))
    if not cls.startswith("Adams"):
        y1, p1 = ensemble(2, n, seed=18)
        m = EmulCtaEnsemble(cls, "rd1d", y1, p1, **kw)
        m.base["max_attempts"] = 2
        m.run([5.0])
        assert (m.status == 3).all() and (m.nacc + m.nrej == 2).all()


@pytest.mark.parametrize("cls,n,terminal_at", [("RungeKutta45", 140, None), ("RungeKutta45", 128, 0.9), ("DOP853", 100, None),
                                               ("RungeKutta23", 70, 1.05), ("DOP853", 96, 0.95)])
def test_run_events_with_one_cta_per_trajectory(cls, n, terminal_at):
    """run_events in the CTA regime (round 2): the block's threads publish the whole vector in natural order for the event
    functions (the same nbk_event source as per-thread), every thread evaluates them, detection / bisection / recording
    are block-uniform; for DOP853 the bisection evaluates the 7th-order dense output whose preparation runs three more
    stages with block barriers.  Against the NumPy restatement of the reference's run_events; bit-identical across
    OS-scheduled repetitions (any missing barrier around the event buffer shows as a difference)."""
    from nbkode_b200 import rhs as nrhs

    def crossing(t, y):
        return y[0] - 0.5

    def gap(t, y):
        return y[n // 2] - y[n // 2 + 2]

    gap.direction = -1

    def level(t, y):
        return y[1] + y[2] - (1.0 if terminal_at is None else terminal_at)

    level.terminal = terminal_at is not None
    evs = [crossing, gap, level]
    B = 3
    y0, p = ensemble(B, n, seed=17, scale=0.4)
    kw = dict(rtol=1e-6, atol=1e-9)
    ev = nrhs.trace_events(evs, n, (2,), True)
    ts = [0.5, 1.5, 3.0]
    first = None
    for rep in range(2):
        g = EmulCtaEnsemble(cls, "rd1d", y0, p, **kw)
        got = g.run_events(ts, ev)
        if first is None:
            first = (g, got)
        else:
            for a, b in zip(got, first[1]):
                assert np.array_equal(a, b, equal_nan=True)
    g, (out, ev_t, ev_y, ev_c, t_term) = first
    n_events = 0
    tol = 1e-6 if cls == "DOP853" else 1e-7
    for b in range(B):
        o = onp.AdaptiveRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], **kw)
        t_ref, y_ref, te_ref, ye_ref = onp.run_events(o, ts, evs)
        k = int(g.nout[b])
        assert k == len(t_ref)
        if g.status[b] == -1:
            assert rel(t_term[b], t_ref[-1]) < 1e-6 and rel(out[b, k - 1], y_ref[-1]) < 1e-5
            if k > 1:
                assert rel(out[b, :k - 1], y_ref[:k - 1]) < tol
        else:
            assert g.status[b] == 0 and rel(out[b, :k], y_ref) < tol
        assert np.isnan(out[b, k:]).all()
        for j in range(3):
            assert ev_c[b, j] == len(te_ref[j])
            n_events += len(te_ref[j])
            if len(te_ref[j]):
                assert rel(ev_t[b, j, :len(te_ref[j])], te_ref[j]) < 1e-6
                assert rel(ev_y[b, j, :len(te_ref[j])], np.array(ye_ref[j])) < 1e-5
    assert n_events > 0


@pytest.mark.parametrize("cls,n,terminal_at", [("RungeKutta4", 100, None), ("AdamsBashforth2", 70, 0.9), ("ForwardEuler", 64, 1.0),
                                               ("Heun3", 128, None)])
def test_run_events_of_the_fixed_step_classes_with_one_cta_per_trajectory(cls, n, terminal_at):
    """run_events of the fixed-step classes in the CTA regime: the team's event routine after every step, bracketing the
    previous point and the new one on the window Hermite (cta_advance_fixed<..., EV>), against the NumPy restatement of the
    reference's run_events; bit-identical across OS-scheduled repetitions."""
    from nbkode_b200 import rhs as nrhs

    def crossing(t, y):
        return y[0] - 0.5

    def gap(t, y):
        return y[n // 2] - y[n // 2 + 2]

    gap.direction = -1

    def level(t, y):
        return y[1] + y[2] - (1.0 if terminal_at is None else terminal_at)

    level.terminal = terminal_at is not None
    evs = [crossing, gap, level]
    B, h = 3, 0.01
    y0, p = ensemble(B, n, seed=17, scale=0.4)
    ev = nrhs.trace_events(evs, n, (2,), True)
    ts = [0.5, 1.5, 3.0]
    first = None
    for rep in range(2):
        g = EmulCtaEnsemble(cls, "rd1d", y0, p, h=h, first_stepper=0)
        got = g.run_events(ts, ev)
        if first is None:
            first = (g, got)
        else:
            for a, b in zip(got, first[1]):
                assert np.array_equal(a, b, equal_nan=True)
    g, (out, ev_t, ev_y, ev_c, t_term) = first
    n_events = 0
    for b in range(B):
        if cls in ("RungeKutta4", "Heun3"):
            o = onp.FixedRK(cls, onp.rhs_rd1d, 0.0, y0[b], p[b], h=h)
        else:
            o = onp.AdamsBashforth(1 if cls == "ForwardEuler" else int(cls[-1]), onp.rhs_rd1d, 0.0, y0[b], p[b], h=h, first_stepper=None)
        t_ref, y_ref, te_ref, ye_ref = onp.run_events(o, ts, evs)
        k = int(g.nout[b])
        assert k == len(t_ref)
        if g.status[b] == -1:
            assert rel(t_term[b], t_ref[-1]) < 1e-7 and rel(out[b, k - 1], y_ref[-1]) < 1e-7
            if k > 1:
                assert rel(out[b, :k - 1], y_ref[:k - 1]) < 1e-12
        else:
            assert g.status[b] == 0 and rel(out[b, :k], y_ref) < 1e-12
        assert np.isnan(out[b, k:]).all()
        for j in range(3):
            assert ev_c[b, j] == len(te_ref[j])
            n_events += len(te_ref[j])
            if len(te_ref[j]):
                assert rel(ev_t[b, j, :len(te_ref[j])], te_ref[j]) < 1e-7
                assert rel(ev_y[b, j, :len(te_ref[j])], np.array(ye_ref[j])) < 1e-7
    assert n_events > 0
