import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Build the native pieces once per session if they are missing."""
    import drone_collision_rp_b200 as pk

    if not (os.path.exists(pk.LIBDCRP) and os.path.exists(pk.LIBDCRP_HOST) and os.path.exists(pk.MONTE_CARLO)
            and os.path.exists(pk.LIBDCRP_BENCH)):
        pk.build_native()
    import oracle_api as oa

    if not os.path.exists(oa.LIBORACLE):
        oa.build_oracle()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    import oracle_api as oa

    return oa.Oracle()


@pytest.fixture(scope="session")
def ref(built):
    import oracle_api as oa

    if not oa.RefReplay.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oa.RefReplay()


@pytest.fixture(scope="session")
def hostsim(built):
    import oracle_api as oa

    return oa.HostSim()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "ref_replay_vectors.npz"))


@pytest.fixture(scope="session")
def day(tmp_path_factory, built):
    """The 6-flight synthetic day the golden vectors were made on: (tracks, csv path)."""
    from drone_collision_rp_b200 import scenario

    d = tmp_path_factory.mktemp("day")
    return scenario.synthetic_tracks(str(d), 6)


@pytest.fixture(scope="session")
def tracks(day):
    return day[0]


@pytest.fixture(scope="session")
def ring1(built):
    import drone_collision_rp_b200 as pk
    from drone_collision_rp_b200 import scenario

    return scenario.load_ring(pk.RING_1KM)


@pytest.fixture(scope="session")
def ring5(built):
    import drone_collision_rp_b200 as pk
    from drone_collision_rp_b200 import scenario

    return scenario.load_ring(pk.RING_5KM)


@pytest.fixture(scope="session")
def engine(built):
    from drone_collision_rp_b200 import capi

    e = capi.Engine(0)
    yield e
    e.close()
