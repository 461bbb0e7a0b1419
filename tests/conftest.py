import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oraclelib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_avx():
    from oraclelib import Ref, cpu_has_avx512
    if not cpu_has_avx512():
        pytest.skip("host CPU has no AVX-512")
    try:
        return Ref("avx")
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="session")
def ref_scalar():
    from oraclelib import Ref
    try:
        return Ref("scalar")
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="session")
def gd():
    import gdiet_b200
    return gdiet_b200


@pytest.fixture(scope="session")
def ctx(gd):
    c = gd.Context(0)
    yield c
    c.close()
