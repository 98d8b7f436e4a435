"""drone-collision-rp_b200 -- B200-native Monte-Carlo collision-sampling engine.

The product is native: ``libdcrp.so`` (hand-written sm_100a CUDA kernels behind
the C ABI of ``include/dcrp.h``) and ``libdcrp_host.so`` / ``Monte_Carlo`` (the
C++ host keeping the reference's Aircraft / Drone / Monte_Carlo API).  This
Python package is only the ctypes binding the tests and ``bench.py`` drive them
through; it contains no arithmetic of the trial path and no CPU fallback.
"""
import os as _os
import subprocess as _sp

PKG_DIR = _os.path.dirname(_os.path.abspath(__file__))
REPO_ROOT = _os.path.dirname(PKG_DIR)
DATA_DIR = _os.path.join(REPO_ROOT, "data")
LIBDCRP = _os.environ.get("DCRP_LIBDCRP") or _os.path.join(PKG_DIR, "libdcrp.so")    # (override: A/B runs of two builds, tools/)
LIBDCRP_HOST = _os.path.join(PKG_DIR, "libdcrp_host.so")
LIBDCRP_BENCH = _os.path.join(PKG_DIR, "libdcrp_bench.so")    # measurement only (bench.py, tools/)
MONTE_CARLO = _os.path.join(PKG_DIR, "Monte_Carlo")
RING_1KM = _os.path.join(DATA_DIR, "drone_inital_positions_1km.csv")
RING_5KM = _os.path.join(DATA_DIR, "drone_inital_positions_5km.csv")


def build_native(verbose=False):
    """Compile libdcrp.so (nvcc, sm_100a), libdcrp_host.so and Monte_Carlo in-tree."""
    res = _sp.run(["make", "-C", PKG_DIR, "all"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building drone-collision-rp_b200 failed (make exit %d)" % res.returncode)


from . import capi  # noqa: E402
from . import scenario  # noqa: E402
from .capi import DcrpError, Engine, Scenario  # noqa: E402,F401
