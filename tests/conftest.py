import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


warnings.simplefilter("ignore")


def inputs(name):
    return os.path.join(GOLD, "inputs", name)


def outputs(name):
    return os.path.join(GOLD, "outputs", name)


@pytest.fixture(scope="session")
def lw():
    import lwfbrook90.jl_b200 as m
    return m


@pytest.fixture(scope="session")
def dav_full(lw):
    return lw.setup(lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True))


@pytest.fixture(scope="session")
def bea(lw):
    return lw.setup(lw.loadSPAC(inputs("BEA2016-reset-FALSE-input"), "BEA2016-reset-FALSE", simulate_isotopes=False))


@pytest.fixture(scope="session")
def isobea(lw):
    return lw.setup(lw.loadSPAC(inputs("isoBEA2016-reset-FALSE-input"), "isoBEA2016-reset-FALSE", simulate_isotopes=True))


@pytest.fixture(scope="session")
def infexp1(lw):
    m = lw.loadSPAC(inputs("infiltrationSaturationINFEXP1"), "infiltrationSaturationINFEXP1", simulate_isotopes=False,
                    Δz_thickness_m=[0.05] * 42,
                    IC_soil=dict(PSIM_init_kPa=-6.3, delta18O_init_permil=-13.0, delta2H_init_permil=-95.0),
                    root_distribution=dict(beta=0.90, z_rootMax_m=None))
    return lw.setup(m, requested_tspan=(0, 300))


def read_table(path):
    import csv
    rows = list(csv.reader(open(path, encoding="utf-8")))
    hdr = rows[0]

    def num(x):
        try:
            return float(x)
        except ValueError:
            return np.nan
    return hdr, np.array([[num(x) for x in r] for r in rows[1:]])


def aligned_grid(N):
    """N layer thicknesses (m) for DAV2020-full whose boundaries fall on the horizon interfaces (0.03 m, 0.40 m), so that
    the discretisation keeps exactly N layers (a grid that cuts through a horizon gets extra layers at the interfaces)."""
    a = max(1, round((N - 1) * 0.35))
    b = N - 1 - a
    return [0.03] + [0.37 / a] * a + [0.7 / b] * b
