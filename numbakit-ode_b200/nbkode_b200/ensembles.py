"""Synthetic ensembles of the BASELINE.json configurations.

Input generators only -- no solver arithmetic.  They live in the package so that ``bench.py``'s
GPU arm and ``tools/`` draw the C2..C5 inputs from the product tree alone; the CPU checker of the
test-suite re-exports them, so that tests and the CPU baseline see the same arrays.
"""

from __future__ import annotations

import math

import numpy as np


def lorenz_ensemble(B, seed=0):
    """C2 generator of SURVEY.md section 8(d): draw order x0,y0,z0,sigma,rho,beta."""
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(-15, 15, B)
    y0 = rng.uniform(-20, 20, B)
    z0 = rng.uniform(5, 45, B)
    sg = rng.uniform(8, 12, B)
    rh = rng.uniform(20, 36, B)
    be = rng.uniform(2, 3.5, B)
    return np.stack([x0, y0, z0], 1), np.stack([sg, rh, be], 1)


def vdp_ensemble(B, seed=1):
    """C3 generator of SURVEY.md section 8(d)."""
    rng = np.random.default_rng(seed)
    mu = np.linspace(0.1, 5, B)
    y0 = np.stack([rng.uniform(-2.5, 2.5, B), rng.uniform(-2.5, 2.5, B)], 1)
    return y0, mu[:, None].copy()


def linear_ensemble(B, n=128, seed=2):
    """C4 generator: A = G/sqrt(n) - I (shared), y0 ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) / math.sqrt(n) - np.eye(n)
    y0 = rng.standard_normal((B, n))
    return y0, A


def rd1d_ensemble(B, n=2048, seed=3):
    """C5 generator: kappa~U(0.5e-4,2e-4), r~U(0.5,2), y0=0.5+0.4 sin(2 pi x + phi)."""
    rng = np.random.default_rng(seed)
    kappa = rng.uniform(0.5e-4, 2e-4, B)
    r = rng.uniform(0.5, 2, B)
    phi = rng.uniform(0, 2 * np.pi, B)
    x = np.arange(n) / n
    y0 = 0.5 + 0.4 * np.sin(2 * np.pi * x[None, :] + phi[:, None])
    d = kappa * n * n  # kappa / dx^2
    return y0, np.stack([d, r], 1)
