"""pytest configuration: registers the ``gpu`` marker and puts the repo root and the
package directory (``numbakit-ode_b200/``, which holds the importable ``nbkode_b200``)
on sys.path."""

import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "numbakit-ode_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden():
    return load_golden


# ---------------------------------------------------------------------------------------------------------------
# Parity bars that are not the plain north-star bar are stated as "measured + margin": `parity_bar` computes
# max |got - ref| / (rtol |ref| + atol), records it (gpurun_out/parity_measured.json, copied into profiles/ per
# round) and asserts it against the limit given at the call site.
# ---------------------------------------------------------------------------------------------------------------
_MEASURED = {}


def parity_bar(name, got, ref, rtol, atol, limit):
    import numpy as np

    got, ref = np.asarray(got, dtype=float), np.asarray(ref, dtype=float)
    ratio = float(np.max(np.abs(got - ref) / (rtol * np.abs(ref) + atol)))
    _MEASURED[name] = {"max_err_over_tol": ratio, "limit": limit, "n": int(ref.size)}
    assert ratio <= limit, f"{name}: max |d| / (rtol |y| + atol) = {ratio:.3f} > {limit}"
    return ratio


def pytest_sessionfinish(session, exitstatus):
    if not _MEASURED:
        return
    try:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_measured.json"), "w") as fh:
            json.dump(_MEASURED, fh, indent=1, sort_keys=True)
    except OSError:
        pass
