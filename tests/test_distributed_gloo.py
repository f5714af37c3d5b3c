"""N > 1 host logic on CPU: world_size-2 gloo run of the sharded sweep.  The engine is replaced by
the oracle-backed test double (no GPU here); sharding, padding and the single all_gather are the
code under test.  Compared with the unsharded result and with the golden reference eigenvalues."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from tests.conftest import ROOT, load_golden


def _worker(rank, world, port, name, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pycqed_b200.distributed import sweep_sharded
    from tests.test_dropin_vs_reference import OracleEngine
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, plan = load_golden(name)
    grid = g["params"][:13]                       # odd count: ranks get 7 and 6 points
    res = sweep_sharded(lambda: OracleEngine(plan), grid, k, want_vectors=True, gather_vectors=True)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), E=res["E"], V=res["V"], info=res["info"],
             lo=res["shard"][0], hi=res["shard"][1], **{"s_" + n: v for n, v in res["scalars"].items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    from pycqed_b200.distributed import shard_bounds
    for n, w in [(13, 2), (10000, 8), (3, 8), (64, 8), (1, 4)]:
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert sum(hi - lo for lo, hi in spans) == n


@pytest.mark.timeout(300)
def test_two_rank_gloo_sweep(tmp_path):
    name, k = "c4_averin", 6
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, name, k, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert (int(r0["lo"]), int(r0["hi"]), int(r1["lo"]), int(r1["hi"])) == (0, 7, 7, 13)
    # every rank holds the full gathered result, identical
    assert np.array_equal(r0["E"], r1["E"]) and np.array_equal(r0["V"], r1["V"])
    g, plan = load_golden(name)
    assert np.max(np.abs(r0["E"] - g["E"][:13, :k])) < 1e-9 * np.abs(g["E"]).max()
    assert r0["V"].shape == (13, k, plan.hilbert_dim)
    assert np.max(np.abs(r0["s_ChargingEnergy"] - g["eval_getChargingEnergies"][:13])) < 1e-12
    assert not r0["info"].any()
