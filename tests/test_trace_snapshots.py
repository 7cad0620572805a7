"""Codegen snapshots: the CUDA C the tracers emit for the right-hand sides and events the GPU parity tests use is pinned by
hash (tests/golden/traced_sources.json).  The emitted text is what NVRTC compiles and what the B200 runs were verified
with; a tracer change that alters it for these callables shows up here, on the CPU, as a reminder to re-run the GPU
tier (and then to refresh the snapshot: ``python tests/test_trace_snapshots.py --update``)."""

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SNAP = os.path.join(HERE, "golden", "traced_sources.json")


def _sources():
    from nbkode_b200 import rhs
    from oracle import nbk_oracle_np as onp

    A = np.arange(16.0).reshape(4, 4) / 7
    grid = np.linspace(0, 1, 64)
    M = np.arange(64.0 * 64).reshape(64, 64) / 4096

    def rd(t, y, p):
        return p[0] * (np.roll(y, -1) - 2 * y + np.roll(y, 1)) + p[1] * y * (1 - y)

    def heat(t, y, p):
        out = np.zeros_like(y)
        out[1:-1] = p[0] * (y[2:] - 2 * y[1:-1] + y[:-2])
        return out

    def ring(t, y, p):
        return p[0] * (np.roll(y, 1) - y) + np.where(y > 0.5, -y, p[1] * y) + grid * y.sum() * 0.01

    def mix(t, y, p):
        return (np.concatenate([y[32:], y[:32]]) * np.clip(y, 0, 1) + np.diff(np.append(y, 0.0)) - np.flip(y) * np.sign(y)
                + np.maximum(y, p[0]) + y[::2].mean())

    out = {}
    scalar = {
        "decay": (onp.rhs_decay, 3, (3,), True), "lorenz": (onp.rhs_lorenz, 3, (3,), True), "vdp": (onp.rhs_vdp, 2, (1,), True),
        "pwl": (onp.rhs_pwl, 2, (3,), True), "lin": (lambda t, y: A @ y, 4, (), False),
        "readme": (lambda t, y, k: -k * y, 1, (1,), True), "cannon": (lambda t, y: np.array([y[1], -0.5]), 2, (), False),
        "trig": (lambda t, y, p: np.array([np.sin(y[0]) * p[0] + np.exp(-t), y[0] ** 2.5, abs(y[1])]), 3, (1,), True),
    }
    for name, (f, n, ps, hp) in scalar.items():
        out["thread/" + name] = rhs.trace(f, n, ps, hp).source
    for name, f in {"rd": rd, "heat": heat, "ring": ring, "mix": mix, "rd1d": onp.rhs_rd1d}.items():
        out["vector/" + name] = rhs.trace_cta(f, 64, (2,), True).source
    out["vector/matvec"] = rhs.trace_cta(lambda t, y: M @ y, 64, (), False).source
    evs = [onp.ev_y0, onp.ev_z25_up, onp.ev_x_m8_terminal, onp.ev_time, onp.ev_switch]
    out["events/catalogue"] = rhs.trace_events(evs, 3, (3,), True).source
    out["resolve/rd_n12"] = rhs.resolve(rd, 12, (2,), True, "auto", True, True, None)[1]
    out["resolve/rd_n8"] = rhs.resolve(rd, 8, (2,), True)[1]
    return {k: hashlib.sha1(v.encode()).hexdigest() for k, v in out.items()}


def test_emitted_sources_match_the_snapshot():
    want = json.load(open(SNAP))
    got = _sources()
    assert set(got) == set(want)
    changed = sorted(k for k in got if got[k] != want[k])
    assert not changed, f"the tracer now emits different CUDA C for {changed}: re-run the GPU tier, then refresh the snapshot"


if __name__ == "__main__":
    sys.path[:0] = [os.path.join(HERE, ".."), os.path.join(HERE, "..", "numbakit-ode_b200")]
    if "--update" in sys.argv:
        json.dump(_sources(), open(SNAP, "w"), indent=1, sort_keys=True)
        print("wrote", SNAP)
