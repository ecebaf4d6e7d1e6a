import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def table_arrays():
    from lineprofiles_b200 import synthetic as syn
    return syn.make_table()


@pytest.fixture(scope="session")
def oracle_table(table_arrays):
    from oracle import oracle_py as orc
    orc.build()
    return orc.OracleTable(table_arrays)


@pytest.fixture(scope="session")
def gpu_table(table_arrays):
    import lineprofiles_b200 as klp
    t = klp.Table.from_arrays(table_arrays, device=0)
    yield t
    t.close()
