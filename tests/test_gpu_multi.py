"""Single-process multi-GPU solver (`devices=[...]`, nbkode_b200/sharded.py): the ensemble is
split over the devices, every call is launched on all of them at once, results are gathered once -- and are
BIT-IDENTICAL to the single-GPU solver, because a trajectory's arithmetic does not depend on where it sits.
Needs at least two visible GPUs for the multi-device cases (skipped otherwise); the one-device spelling
`devices=[0]` is always covered."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

import nbkode_b200 as nbk  # noqa: E402
from nbkode_b200 import ensembles  # noqa: E402

NDEV = torch.cuda.device_count()
multi = pytest.mark.skipif(NDEV < 2, reason="needs two GPUs")


def test_devices_list_with_one_entry_is_the_plain_solver():
    y0, p = ensembles.lorenz_ensemble(257, seed=5)
    a = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-6, atol=1e-9, devices=[0])
    b = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-6, atol=1e-9)
    assert type(a) is type(b) and not isinstance(a, nbk.ShardedSolver)
    assert np.array_equal(a.run([1.0, 2.0])[1], b.run([1.0, 2.0])[1])
    with pytest.raises(ValueError):
        nbk.RungeKutta45("lorenz", 0.0, y0, p, devices=[0], device=0)


@multi
@pytest.mark.parametrize("cls,kw", [("RungeKutta45", dict(rtol=1e-6, atol=1e-9)), ("RungeKutta23", dict(rtol=1e-5, atol=1e-8)),
                                    ("AdamsBashforth4", dict(h=0.002)), ("RungeKutta4", dict(h=0.01)), ("DOP853", dict(rtol=1e-7, atol=1e-9))])
def test_sharded_equals_single_gpu_bitwise(cls, kw):
    B = 10_007  # odd: shards differ in size
    y0, p = ensembles.lorenz_ensemble(B, seed=11)
    devs = list(range(min(NDEV, 8)))
    ts = [0.5, 1.0, 3.0]
    one = getattr(nbk, cls)("lorenz", 0.0, y0, p, **kw)
    many = getattr(nbk, cls)("lorenz", 0.0, y0, p, devices=devs, **kw)
    assert isinstance(many, nbk.ShardedSolver) and isinstance(many, nbk.Solver) and len(many.shards) == len(devs)
    assert sorted({s._dev_index for s in many.shards}) == devs
    t1, y1 = one.run(ts)
    tm, ym = many.run(ts)
    assert np.array_equal(t1, tm) and ym.shape == (B, 3, 3)
    # (equal_nan: a fixed-step method may overflow on a few of the random Lorenz systems -- on both solvers alike)
    assert np.array_equal(y1, ym, equal_nan=True) and np.isfinite(y1).mean() > 0.99
    assert np.array_equal(one.t, many.t) and np.array_equal(one.y, many.y, equal_nan=True) and np.array_equal(one.f, many.f, equal_nan=True)
    assert np.array_equal(one.n_accepted, many.n_accepted) and (many.status == 0).all()
    assert np.array_equal(one.cache.ys, many.cache.ys, equal_nan=True) and np.array_equal(one.cache.ts, many.cache.ts)
    # step / skip / interpolate go through the same per-device fan-out
    a1, b1 = one.step(n=3)
    am, bm = many.step(n=3)
    assert np.array_equal(a1, am) and np.array_equal(b1, bm, equal_nan=True)
    one.skip(n=2)
    many.skip(n=2)
    assert np.array_equal(one.y, many.y, equal_nan=True)
    if cls != "AdamsBashforth4":  # (its window Hermite spans the whole history: interpolate inside the last step only)
        ti = float(np.max(one.cache.ts[0]) * 0.5 + np.min(one.cache.ts[-1]) * 0.5)
        if np.max(one.cache.ts[0]) <= ti <= np.min(one.cache.ts[-1]):
            assert np.array_equal(one.interpolate(ti), many.interpolate(ti), equal_nan=True)


@multi
def test_sharded_torch_in_torch_out_and_shared_params():
    devs = list(range(min(NDEV, 8)))
    # CUDA tensors in -> one CUDA tensor out, on the device the inputs live on
    y0, p = ensembles.lorenz_ensemble(4096, seed=3)
    y0d, pd = torch.from_numpy(y0).cuda(0), torch.from_numpy(p).cuda(0)
    s = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9, devices=devs)
    t, y = s.run([2.0])
    s.raise_for_status()
    assert isinstance(y, torch.Tensor) and y.device == y0d.device and y.shape == (4096, 1, 3)
    ref = nbk.RungeKutta45("lorenz", 0.0, y0d, pd, rtol=1e-6, atol=1e-9)
    assert torch.equal(ref.run([2.0])[1], y)
    # a shared (replicated) parameter block and a traced Python right-hand side (traced once, reused by every shard)
    rng = np.random.default_rng(0)
    y0v = rng.uniform(-2, 2, (1001, 2))

    def vdp(t, y, p):
        return np.array([y[1], p[0] * (1 - y[0] ** 2) * y[1] - y[0]])

    a = nbk.RungeKutta23(vdp, 0.0, y0v, np.array([1.5]), rtol=1e-6, atol=1e-9)
    b = nbk.RungeKutta23(vdp, 0.0, y0v, np.array([1.5]), rtol=1e-6, atol=1e-9, devices=devs)
    assert all(sh.is_jit for sh in b.shards)
    assert np.array_equal(a.run(np.linspace(0, 5, 11))[1], b.run(np.linspace(0, 5, 11))[1])
    # fewer trajectories than devices: only as many shards as trajectories
    c = nbk.RungeKutta45("lorenz", 0.0, y0[:1], p[:1], rtol=1e-6, atol=1e-9, devices=devs)
    assert len(c.shards) == 1 and c.run([1.0])[1].shape == (1, 1, 3)


@multi
def test_sharded_events_and_errors():
    devs = list(range(min(NDEV, 8)))
    y0, p = ensembles.lorenz_ensemble(64, seed=9)

    def z25(t, y):
        return y[2] - 25.0

    z25.direction = 1
    a = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-8, atol=1e-10)
    b = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-8, atol=1e-10, devices=devs)
    ta, ya, tea, yea = a.run_events([1.0, 2.0], [z25])
    tb, yb, teb, yeb = b.run_events([1.0, 2.0], [z25])
    assert np.array_equal(ya, yb) and np.array_equal(a.event_counts, b.event_counts)
    assert np.array_equal(tea[0], teb[0], equal_nan=True) and np.array_equal(yea[0], yeb[0], equal_nan=True)
    with pytest.raises(ValueError):
        nbk.RungeKutta45("lorenz", 0.0, y0[0], p[0], devices=devs)       # unbatched
    with pytest.raises(ValueError):
        nbk.RungeKutta45("lorenz", 0.0, y0, p, devices=[0, 0])            # duplicate device
    with pytest.raises(ValueError):
        nbk.RungeKutta45("lorenz", 0.0, y0, p, devices=[0, NDEV + 3])     # out of range


@multi
@pytest.mark.parametrize("cls,n,kw", [("RungeKutta45", 24, dict(rtol=1e-6, atol=1e-9)), ("DOP853", 20, dict(rtol=1e-7, atol=1e-10)),
                                      ("AdamsBashforth4", 32, dict(h=0.002)), ("RungeKutta45", 300, dict(rtol=1e-6, atol=1e-9))])
def test_sharded_component_wise_regimes(cls, n, kw):
    """devices=[...] over the sub-warp groups / teams and the CTA regime (trajectory-major state): bit-identical to one GPU,
    incl. run_events for the adaptive classes."""
    devs = list(range(min(NDEV, 8)))
    B = 1003
    y0, p = ensembles.rd1d_ensemble(B, n=n, seed=3)
    p[:, 0] *= 0.3 * (100.0 / n) ** 2
    one = getattr(nbk, cls)("rd1d", 0.0, y0, p, **kw)
    many = getattr(nbk, cls)("rd1d", 0.0, y0, p, devices=devs, **kw)
    ts = [0.1, 0.3]
    assert np.array_equal(one.run(ts)[1], many.run(ts)[1])
    assert np.array_equal(one.y, many.y) and np.array_equal(one.n_accepted, many.n_accepted)
    a1, b1 = one.step(n=2)
    am, bm = many.step(n=2)
    assert np.array_equal(a1, am) and np.array_equal(b1, bm)
    if cls != "AdamsBashforth4":
        def crossing(t, y):
            return y[0] - 0.5

        e1 = getattr(nbk, cls)("rd1d", 0.0, y0, p, **kw)
        em = getattr(nbk, cls)("rd1d", 0.0, y0, p, devices=devs, **kw)
        r1, rm = e1.run_events([0.5, 1.5], [crossing]), em.run_events([0.5, 1.5], [crossing])
        assert np.array_equal(r1[1], rm[1], equal_nan=True) and np.array_equal(e1.event_counts, em.event_counts)
        assert np.array_equal(r1[2][0], rm[2][0], equal_nan=True)
