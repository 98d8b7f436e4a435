"""GPU tier: the engine-owned communicator (include/dcrp.h "multi-process").  On the one-GPU test box only
the single-rank communicator can be built (NCCL refuses two ranks on one device); the two-rank exchange is
exercised by tools/comm_two_ranks.py under `gpurun --gpus 2` and by bench.py at N > 1."""
import numpy as np
import pytest

from drone_collision_rp_b200 import capi

pytestmark = pytest.mark.gpu


def test_single_rank_communicator(tracks, ring1):
    eng = capi.Engine(0)
    with pytest.raises(capi.DcrpError) as e:
        eng.allreduce_counts()
    assert e.value.status == capi.ERR_STATE                      # no communicator yet
    uid = capi.Engine.comm_unique_id()
    assert len(uid) == 128 and any(uid)
    eng.attach_comm(uid, 0, 1)
    info = eng.comm_info()
    assert info["attached"] and info["nranks"] == 1 and info["nccl_version"] >= 22000
    with pytest.raises(capi.DcrpError):
        eng.attach_comm(uid, 0, 1)                               # already attached
    eng.upload(tracks[:1], ring1)
    run = eng.run_async(seed=11, trial_count=20000, precision=capi.MIXED)
    eng.allreduce_counts()                                       # one rank: the identity
    got = eng.fetch(run)
    ref = eng.simulate(tracks[:1], ring1, seed=11, trial_count=20000, precision=capi.FP64)
    assert np.array_equal(got.counts, ref.counts) and got.hits > 0
    eng.detach_comm()
    assert not eng.comm_info()["attached"]
    with pytest.raises(capi.DcrpError):
        eng.attach_comm(uid, 3, 2)                               # rank out of range
    eng.close()
