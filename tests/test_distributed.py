"""CPU (gloo, world_size 2): the host logic of the multi-GPU path -- sharding, gather, min-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from compas_tno_b200.distributed import shard_bounds


def test_shard_bounds_tile_the_batch():
    for N in (0, 1, 7, 256, 65536, 65537):
        for G in (1, 2, 3, 4, 8):
            cuts = [shard_bounds(N, G, r) for r in range(G)]
            assert cuts[0][0] == 0 and cuts[-1][1] == N
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(G - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, N, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from compas_tno_b200.distributed import ShardedSweep, argmin_feasible, count_feasible, gather_objectives
    rng = np.random.default_rng(123)                    # same "global sweep" on every rank
    f = rng.normal(size=N)
    gmin = rng.normal(scale=0.5, size=N)
    status = (rng.uniform(size=N) < 0.1).astype(np.int32)
    sweep = ShardedSweep(plan=None, n_candidates=N)
    lo, hi = sweep.lo, sweep.hi
    res = {"f": torch.from_numpy(f[lo:hi].copy()), "gmin": torch.from_numpy(gmin[lo:hi].copy()),
           "status": torch.from_numpy(status[lo:hi].copy())}
    f_all, s_all = sweep.gather(res)
    assert np.array_equal(f_all.numpy(), f) and np.array_equal(s_all.numpy(), status)
    f_all2, _ = gather_objectives(res["f"], res["status"])       # sizes discovered by a collective
    assert np.array_equal(f_all2.numpy(), f)
    best, idx = sweep.best(res, tol=0.0)
    ok = (gmin >= 0) & (status == 0)
    exp = np.where(ok, f, np.inf)
    assert idx == int(np.argmin(exp)) and best == float(exp.min())
    assert count_feasible(res["gmin"], res["status"], tol=0.0) == int(ok.sum())
    # the packed single-collective form gives the same answer on every rank
    from compas_tno_b200.distributed import best_of_packed
    packed = sweep.gather_packed(res)
    assert tuple(packed.shape) == (N, 3)
    assert np.array_equal(packed[:, 0].numpy(), f) and np.array_equal(packed[:, 1].numpy(), gmin)
    assert np.array_equal(packed[:, 2].numpy().astype(np.int32), status)
    assert best_of_packed(packed, tol=0.0) == (float(exp.min()), int(np.argmin(exp)), int(ok.sum()))
    assert best_of_packed(torch.stack([packed[:, 0], packed[:, 1] - 100.0, packed[:, 2]], dim=1), tol=0.0) == (float("inf"), -1, 0)
    # nothing feasible
    none = argmin_feasible(res["f"], res["gmin"] - 100.0, res["status"], lo, tol=0.0)
    assert none == (float("inf"), -1)
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [1001, 64])
def test_gather_and_argmin_world2_gloo(tmp_path, N):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, N, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
