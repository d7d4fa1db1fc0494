"""N > 1 host logic on CPU: world_size-2 gloo processes (no GPU).  Covers the partition rule, the timing / counter
aggregation bench.py uses, and -- with the CPU oracle standing in for the kernels -- that sharding the environments
over ranks by global id reproduces the single-process results (the counter-based RNG is keyed by global env id)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, steps, seed, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    # the package's env module needs no GPU to be imported (only to construct an environment)
    d = importlib.import_module("drl_uav_cellularnet_b200.dist")
    from oracle import mobi_oracle as orc
    d.init("gloo")
    assert d.world() == (rank, world, rank)
    lo, hi = d.shard_range(n_total, rank, world)
    envs = [orc.OracleEnv(orc.default_cfg(), seed=seed, env_id=e) for e in range(lo, hi)]
    for o in envs:
        o.reset()
    acts = np.random.RandomState(99).randint(0, 625, size=(steps, n_total))
    chk = np.zeros((hi - lo, 3))
    for t in range(steps):
        for i, o in enumerate(envs):
            s, r, dn, info = o.step(int(acts[t, lo + i]), want_state=False)
            chk[i] += (r, info["n_out"], info["n_ho"])
    d.barrier()
    value, ms = d.whole_job_throughput(hi - lo, steps, ms_this_rank=10.0 * (rank + 1))
    tot = d.sum_over_ranks([float(hi - lo), chk[:, 1].sum()])
    mx = d.max_over_ranks([float(rank)])
    q.put((rank, lo, hi, chk, value, ms, tot, mx))
    import torch.distributed as dist
    dist.destroy_process_group()


def test_shard_range_partitions():
    from drl_uav_cellularnet_b200 import shard_range
    for n, w in ((65536, 8), (4096, 1), (10, 3), (7, 8)):
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 8, 8)


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharding_matches_single_process():
    from oracle import mobi_oracle as orc
    orc.lib()
    n_total, steps, seed, world = 5, 12, 321, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, steps, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=240) for _ in procs], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference run over all global env ids
    envs = [orc.OracleEnv(orc.default_cfg(), seed=seed, env_id=e) for e in range(n_total)]
    for o in envs:
        o.reset()
    acts = np.random.RandomState(99).randint(0, 625, size=(steps, n_total))
    want = np.zeros((n_total, 3))
    for t in range(steps):
        for e, o in enumerate(envs):
            s, r, dn, info = o.step(int(acts[t, e]), want_state=False)
            want[e] += (r, info["n_out"], info["n_ho"])
    got = np.concatenate([o[3] for o in outs])
    assert [(o[1], o[2]) for o in outs] == [(0, 3), (3, 5)]
    assert np.array_equal(got, want)
    for o in outs:
        assert o[5] == 20.0                                     # max over ranks of (10, 20) ms
        assert abs(o[4] - n_total * steps / 20e-3) < 1e-6       # all envs / slowest rank
        assert o[6] == [float(n_total), float(want[:, 1].sum())]
        assert o[7] == [1.0]
