"""BASELINE.json configs[2..4] at their FULL sizes on one B200, checked through size-independent properties:
rows of the full ensemble that the live reference integrated (golden vectors) or that the C oracle integrates here
sit at their places in the full result; every trajectory ends cleanly; trajectory arithmetic does not depend on the
position in the ensemble.  (configs[1], the 1 M Lorenz ensemble, is in test_gpu_parity.py.)"""

import math

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():  # pragma: no cover
    pytest.skip("needs a CUDA device", allow_module_level=True)

import nbkode_b200 as nbk  # noqa: E402
from oracle import nbk_oracle_c as oc  # noqa: E402
from oracle import nbk_oracle_np as onp  # noqa: E402


def within(y, ref, rtol, atol, slack=1.0):
    return np.abs(np.asarray(y) - np.asarray(ref)) <= slack * (rtol * np.abs(ref) + atol)


def test_c3_vdp_4m_dense_output_full_size():
    """configs[2]: 4 M Van der Pol trajectories, 1000-point grid: 64 GB of dense output per solver, kept on the device."""
    if torch.cuda.get_device_properties(0).total_memory < 150e9:
        pytest.skip("needs a 180 GB device")
    g = load_golden("vdp_ensemble.json")
    B = g["B_total"]
    y0, p = onp.vdp_ensemble(B, seed=g["seed"])
    ts = np.linspace(0, 20, g["n_ts"])
    y0d, pd = torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda()
    keep = torch.tensor(g["keep"], device="cuda")
    for cls, kw, key in (("RungeKutta23", dict(rtol=1e-6, atol=1e-9), "rk23"), ("AdamsBashforth4", dict(h=0.01), "ab4")):
        s = getattr(nbk, cls)("vdp", 0.0, y0d, pd, **kw)
        t, y = s.run(ts)
        assert isinstance(y, torch.Tensor) and y.shape == (B, g["n_ts"], 2)
        s.raise_for_status()
        assert bool((s.status == 0).all())
        rows = torch.tensor([r["row"] for r in g[key]], device="cuda")
        got = y[rows][:, keep].cpu().numpy()
        ref = np.array([r["ys"] for r in g[key]])
        if key == "rk23":
            assert within(got, ref, 1e-6, 1e-9).all()
            nacc = s.n_accepted[rows].cpu().numpy()
            assert np.abs(nacc - np.array([r["n_accepted"] for r in g[key]])).max() <= 1
        else:
            assert float(np.max(np.abs(got - ref))) / max(1.0, float(np.max(np.abs(ref)))) <= 1e-12
            assert np.array_equal(s.t[rows].cpu().numpy(), [r["t_end"] for r in g[key]])
        # the whole output is finite and starts at y0 (first grid point is t0)
        assert bool(torch.isfinite(y[:, -1]).all()) and bool((y[:, 0] == y0d).all())
        del y, s
        torch.cuda.empty_cache()


def test_c4_linear_262k_full_size():
    """configs[3]: y' = A y, n = 128, shared dense A, 262 144 trajectories, RK45 through the DMMA kernel (K4)."""
    B, n = 262_144, 128
    y0, A = onp.linear_ensemble(B, n=n, seed=2)
    s = nbk.RungeKutta45("linear", 0.0, torch.from_numpy(y0).cuda(), torch.from_numpy(A).cuda(), rtol=1e-6, atol=1e-9)
    t, y = s.run([10.0])
    s.raise_for_status()
    assert bool((s.status == 0).all()) and bool((s.t >= 10.0).all())
    idx = np.r_[0:8, B // 2:B // 2 + 8, B - 8:B]
    ref = oc.run_ensemble("RungeKutta45", "linear", y0[idx], A, [10.0], rtol=1e-6, atol=1e-9, params_shared=True, nthreads=8)
    got = y[torch.from_numpy(idx).cuda()].cpu().numpy()
    assert within(got, ref["ys"], 1e-6, 1e-9).all()
    assert np.abs(s.n_accepted[torch.from_numpy(idx).cuda()].cpu().numpy() - ref["n_accepted"]).max() <= 1
    # position independence: the same 4096 trajectories at other places of another ensemble give the same bits
    perm = np.random.default_rng(1).permutation(B)[:4096]
    s2 = nbk.RungeKutta45("linear", 0.0, torch.from_numpy(y0[perm]).cuda(), torch.from_numpy(A).cuda(), rtol=1e-6, atol=1e-9)
    t2, y2 = s2.run([10.0])
    assert bool((y2 == y[torch.from_numpy(perm).cuda()]).all())


def test_c5_reaction_diffusion_65k_full_size():
    """configs[4]: 1-D reaction-diffusion, n = 2048, 65 536 trajectories, RK45, one CTA per trajectory (K3)."""
    B, n = 65_536, 2048
    y0, p = onp.rd1d_ensemble(B, n=n, seed=3)
    s = nbk.RungeKutta45("rd1d", 0.0, torch.from_numpy(y0).cuda(), torch.from_numpy(p).cuda(), rtol=1e-6, atol=1e-9)
    t, y = s.run([1.0])
    s.raise_for_status()
    assert bool((s.status == 0).all()) and bool((s.t >= 1.0).all()) and bool(torch.isfinite(y).all())
    idx = np.r_[0:4, B // 2:B // 2 + 4, B - 4:B]
    ref = oc.run_ensemble("RungeKutta45", "rd1d", y0[idx], p[idx], [1.0], rtol=1e-6, atol=1e-9, nthreads=8)
    sel = torch.from_numpy(idx).cuda()
    got = y[sel].cpu().numpy()
    # BASELINE configs[4] itself: the plain north-star bar (measured 0.045x against the C port and 0.042x against the
    # bit-exact NumPy restatement on 12 trajectories, profiles/r02_parity.json), identical step counts +-1
    from conftest import parity_bar

    parity_bar("c5/full_size_65536/gpu_vs_c_port", got, ref["ys"], 1e-6, 1e-9, 1.0)
    assert np.abs(s.n_accepted[sel].cpu().numpy() - ref["n_accepted"]).max() <= 1
    # conservation-type sanity of the logistic reaction term: the solution stays inside (0, 1]
    assert float(y.min()) > 0.0 and float(y.max()) <= 1.0 + 1e-9


def test_state_arrays_beyond_2_to_31_elements():
    """120 M Lorenz trajectories: the stage-row array K [7][3][B] has 2.5e9 elements, so every index expression has
    to be 64-bit.  Inputs are generated on the device; trajectories at both ends and around every 2^31-element
    boundary of K are checked against the C oracle."""
    if torch.cuda.get_device_properties(0).total_memory < 150e9:
        pytest.skip("needs a 180 GB device")
    B = 120_000_000
    gen = torch.Generator(device="cuda").manual_seed(7)
    y0 = torch.empty((B, 3), dtype=torch.float64, device="cuda").uniform_(-10, 10, generator=gen)
    y0[:, 2] += 25.0
    p = torch.tensor([10.0, 28.0, 8 / 3], dtype=torch.float64, device="cuda")  # shared parameters
    s = nbk.RungeKutta45("lorenz", 0.0, y0, p, rtol=1e-6, atol=1e-9, params_batched=False)
    t, y = s.run([0.25])
    s.raise_for_status()
    assert y.shape == (B, 1, 3) and bool((s.status == 0).all())
    cross = 2**31 - 17 * B  # (row 17 of K) * B + cross == 2^31: the first element a 32-bit offset cannot reach
    assert 0 < cross < B
    probe = [0, 1, cross - 1, cross, cross + 1, B // 2, B - 2, B - 1]
    idx = torch.tensor(probe, device="cuda")
    ref = oc.run_ensemble("RungeKutta45", "lorenz", y0[idx].cpu().numpy(), np.array([10.0, 28.0, 8 / 3]), [0.25],
                          rtol=1e-6, atol=1e-9, params_shared=True)
    assert within(y[idx].cpu().numpy(), ref["ys"], 1e-6, 1e-9).all()
    assert np.array_equal(s.n_accepted[idx].cpu().numpy(), ref["n_accepted"])
    k_last = s._K[6, 2, idx].cpu().numpy()  # last stage row, last component: the far end of the 2.5e9-element array
    assert np.isfinite(k_last).all() and np.allclose(s.f[idx].cpu().numpy()[:, 2], k_last, rtol=1e-12)
    del y, s, y0
    torch.cuda.empty_cache()
