"""Seeded differential test: random call sequences (run / step / skip, with and without upto_t) on random
(solver class, built-in right-hand side) pairs, every trajectory of a small batched ensemble compared with the NumPy
restatement of the reference (oracle/nbk_oracle_np.py, bit-exact against the live-reference golden vectors) driven
through the same sequence.  Bars as everywhere: fixed-step 1e-12 relative, adaptive rtol |y| + atol on values and
identical numbers of steps (the sequences are short and non-chaotic enough for that)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

import nbkode_b200 as nbk  # noqa: E402
from oracle import nbk_oracle_np as onp  # noqa: E402

RHS = {
    "decay": (onp.rhs_decay, 1, lambda r, B: (r.uniform(0.5, 2.0, (B, 1)), r.uniform(0.05, 0.6, (B, 1)))),
    "vdp": (onp.rhs_vdp, 2, lambda r, B: (r.uniform(-2, 2, (B, 2)), r.uniform(0.2, 2.0, (B, 1)))),
    "lorenz": (onp.rhs_lorenz, 3, lambda r, B: (np.stack([r.uniform(-5, 5, B), r.uniform(-5, 5, B), r.uniform(15, 30, B)], 1),
                                                 np.stack([r.uniform(8, 12, B), r.uniform(20, 30, B), r.uniform(2, 3, B)], 1))),
}
ADAPTIVE = ("RungeKutta23", "RungeKutta45", "DOP853")
FIXED_RK = ("Runge2", "Runge3", "Heun3", "RungeKutta4", "RungeKutta3_8")
AB = ("ForwardEuler", "AdamsBashforth1", "AdamsBashforth2", "AdamsBashforth3", "AdamsBashforth4", "AdamsBashforth5")


def make_pair(cls, rname, y0, p, rng):
    fn = RHS[rname][0] if rname in RHS else onp.rhs_rd1d
    if cls in ADAPTIVE:
        rtol, atol = 10.0 ** rng.uniform(-8, -4), 10.0 ** rng.uniform(-11, -7)
        g = getattr(nbk, cls)(rname, 0.0, y0, p, rtol=rtol, atol=atol)
        o = [onp.AdaptiveRK(cls, fn, 0.0, y0[b], p[b], rtol=rtol, atol=atol) for b in range(len(y0))]
        return g, o, ("adaptive", rtol, atol)
    h = float(rng.choice([0.004, 0.01, 0.02]))
    if rname == "rd1d" and cls in AB and cls not in ("ForwardEuler", "AdamsBashforth1"):
        # CTA regime: the start-up is RungeKutta45 (default) or none
        fs = [None, "RungeKutta45"][int(rng.integers(0, 2))]
        order = int(cls[-1])
        g = getattr(nbk, cls)(rname, 0.0, y0, p, h=h, first_stepper_cls=fs)
        o = [onp.AdamsBashforth(order, fn, 0.0, y0[b], p[b], h=h, first_stepper=fs) for b in range(len(y0))]
        return g, o, ("fixed", h, 0.0)
    if cls in FIXED_RK:
        g = getattr(nbk, cls)(rname, 0.0, y0, p, h=h)
        o = [onp.FixedRK(cls, fn, 0.0, y0[b], p[b], h=h) for b in range(len(y0))]
    else:
        order = 1 if cls == "ForwardEuler" else int(cls[-1])
        g = getattr(nbk, cls)(rname, 0.0, y0, p, h=h)
        o = [onp.AdamsBashforth(order, fn, 0.0, y0[b], p[b], h=h) for b in range(len(y0))]
    return g, o, ("fixed", h, 0.0)


def osteps(s, n=None, upto_t=None):
    """onp.steps that also accepts "no step fits" (the oracle's reshape needs at least one row)."""
    bound = s.t_bound if upto_t is None else upto_t
    tt, yy = [], []
    while n is None or len(tt) < n:
        if not s.step(bound):
            break
        tt.append(s.t)
        yy.append(np.array(s.y, dtype=float).copy())
    return np.array(tt), np.array(yy).reshape(len(tt), np.size(s.y))


def close(kind, a, b, what):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if kind[0] == "fixed":
        err = np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))) if a.size else 0.0
        assert err <= 1e-12, (what, err)
    else:
        _, rtol, atol = kind
        bad = np.abs(a - b) > rtol * np.abs(b) + atol
        assert not bad.any(), (what, float(np.max(np.abs(a - b) / (rtol * np.abs(b) + atol))))


@pytest.mark.parametrize("seed", range(96))
def test_random_call_sequences(seed):
    rng = np.random.default_rng(1000 + seed)
    cls = str(rng.choice(ADAPTIVE + FIXED_RK + AB))
    rname = str(rng.choice(list(RHS)))
    B = int(rng.integers(2, 6))
    y0, p = RHS[rname][2](rng, B)
    _drive(rng, cls, rname, y0, p)


@pytest.mark.parametrize("seed", range(48))
def test_random_call_sequences_cta_regime(seed):
    """The same through the CTA-per-trajectory kernels: reaction-diffusion grids of assorted sizes (ragged and "full"
    shapes, one / two / four components per thread), diffusion scaled down so that the grid is not stiff."""
    rng = np.random.default_rng(5000 + seed)
    cls = str(rng.choice(ADAPTIVE + FIXED_RK + AB))
    n = int(rng.choice([40, 64, 100, 128, 300, 512]))
    B = int(rng.integers(2, 5))
    y0, p = onp.rd1d_ensemble(B, n=n, seed=int(rng.integers(0, 1 << 30)))
    p[:, 0] *= 0.05 * (100.0 / n) ** 2
    g = _drive(rng, cls, "rd1d", y0, p)
    assert g._regime == 1


def _drive(rng, cls, rname, y0, p, step_slack=1.0):
    g, o, kind = make_pair(cls, rname, y0, p, rng)
    # step times of the adaptive pairs: the controller feeds the rounding of the error estimate back into h.  E.K cancels
    # down to the local error (8+ digits at rtol 1e-8; more for DOP853's E5.K / E3.K), so an FMA-contracted and a plain
    # evaluation differ by up to ~1e-8 (RungeKutta23/45) and ~1e-7 (DOP853) in h -- not part of the parity bar, which
    # is values within rtol |y| + atol and step counts within 1
    tkind = kind if kind[0] == "fixed" else ("adaptive", 1e-6 if cls == "DOP853" else 1e-7, 0.0)
    # states AFTER a step are compared at the two implementations' own step times, which differ by the figure above: the
    # caller may widen the bar of those comparisons (not of run(), which interpolates at given times)
    skind = kind if kind[0] == "fixed" else ("adaptive", kind[1] * step_slack, kind[2] * step_slack)
    for _ in range(int(rng.integers(3, 7))):
        op = str(rng.choice(["run", "step", "skip", "step_upto", "skip_upto"]))
        tmax = max(float(s.t) for s in o)
        if tmax > 1.2:  # keep Lorenz short of its chaotic divergence
            break
        if op == "run":
            ts = np.sort(tmax + rng.uniform(0.0, 0.4, int(rng.integers(1, 6))))
            tg, yg = g.run(ts)
            for b, s in enumerate(o):
                _, yo = onp.run(s, ts)
                close(kind, yg[b], yo, (cls, rname, op, b))
        elif op in ("step", "skip"):
            k = int(rng.integers(1, 5))
            if op == "step":
                tg, yg = g.step(n=k)
                for b, s in enumerate(o):
                    to, yo = osteps(s, n=k)
                    close(tkind, tg[b], to, (cls, rname, "step t", b))
                    close(skind, yg[b], yo, (cls, rname, "step y", b))
            else:
                g.skip(n=k)
                for s in o:
                    osteps(s, n=k)
        else:
            upto = tmax + float(rng.uniform(0.01, 0.2))
            if op == "step_upto":
                tg, yg = g.step(upto_t=upto)
                for b, s in enumerate(o):
                    to, yo = osteps(s, upto_t=upto)
                    k = len(to)  # batched step(upto_t=) pads with NaN
                    assert np.isnan(tg[b, k:]).all(), (cls, rname, op, b)
                    close(tkind, tg[b, :k], to, (cls, rname, "upto t", b))
                    close(skind, yg[b, :k], yo, (cls, rname, "upto y", b))
            else:
                g.skip(upto_t=upto)
                for s in o:
                    osteps(s, upto_t=upto)
        # the state after every operation
        close(tkind, g.t, [s.t for s in o], (cls, rname, op, "t"))
        if kind[0] == "fixed":
            close(kind, g.y, np.array([s.y for s in o]), (cls, rname, op, "y"))
        else:
            # each implementation's state sits at its OWN step end: allow |f| |t_gpu - t_oracle| on top of the tolerance
            for b, s in enumerate(o):
                dt = abs(float(np.atleast_1d(g.t)[b]) - float(s.t))
                lim = kind[1] * np.abs(s.y) + kind[2] + 2 * np.abs(s.f) * dt
                assert (np.abs(np.atleast_2d(g.y)[b] - s.y) <= lim).all(), (cls, rname, op, "y", b)
    assert (np.asarray(g.status) <= 1).all()
    return g


@pytest.mark.parametrize("seed", range(36))
def test_random_call_sequences_sub_warp_regime(seed):
    """The same for small and mid-size component-wise systems: sub-warp groups (RungeKutta23 / RungeKutta45: dedicated
    kernels) and sub-warp teams (DOP853, Adams-Bashforth, fixed-step Runge-Kutta: the CTA regime's source with G lanes per
    trajectory), 2 .. 32 lanes per trajectory, ragged and full groups, more trajectories than groups in a warp."""
    rng = np.random.default_rng(9000 + seed)
    cls = str(rng.choice(ADAPTIVE + FIXED_RK + AB))
    n = int(rng.choice([5, 8, 13, 16, 24, 32, 50, 64, 100]))
    B = int(rng.integers(2, 40))
    y0, p = onp.rd1d_ensemble(B, n=n, seed=int(rng.integers(0, 1 << 30)))
    p[:, 0] *= 0.05 * (100.0 / n) ** 2
    # (DOP853's error estimate is a sum over the team: another summation order than the oracle's moves its step times by
    # ~1e-7 relative -- measured 1.75 x the state bar after a step on one of these sequences)
    g = _drive(rng, cls, "rd1d", y0, p, step_slack=3.0 if cls == "DOP853" else 1.0)
    assert g._regime == 1
    if not (cls == "DOP853" and n > 64):
        assert g._group in (2, 4, 8, 16, 32)
