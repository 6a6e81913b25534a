"""Shared test plumbing.

* `-m "not gpu"`: oracle vs the reference's golden vectors, host logic, C-ABI symbol check.
* `-m gpu`     : parity tests proper — the CUDA path (through the C ABI) against the oracle.

The oracle (oracle/) is only ever the checker here; the product package never sees it.
"""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_product():
    """import the package directory `alphabeta-rs_b200` (not a valid identifier) as `alphabeta_rs_b200`"""
    name = "alphabeta_rs_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "alphabeta-rs_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def ab():
    return load_product()


@pytest.fixture(scope="session")
def oracle():
    from oracle import abref_py

    abref_py.build()
    return abref_py


@pytest.fixture(scope="session")
def ctx(ab):
    c = ab.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


def resolve_golden(path: str) -> str:
    """node files in the fixtures' nodelists are relative to the reference's repo root / data dir"""
    p = path
    if p.startswith("./data/"):
        p = p[len("./data/"):]
    elif p.startswith("./"):
        p = os.path.join("desired_output", p[2:])
    return os.path.join(GOLDEN, p)
