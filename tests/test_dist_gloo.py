"""CPU, world_size 2, gloo: the N>1 host path -- contiguous sharding + one integer all-reduce of
the confusion matrix, and the prototype mean reduction.  The per-tile compute is stood in for by
the oracle (the CUDA kernels cannot run here); what is under test is the sharding/reduction."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_ops
from segland_b200 import sweep


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_tiles, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        K = 12
        rng = np.random.default_rng(7)
        gts = rng.integers(0, K, size=(n_tiles, 32, 32)).astype(np.uint8)
        gts[rng.random(gts.shape) < 0.05] = 255
        prs = rng.integers(0, K, size=(n_tiles, 32, 32)).astype(np.uint8)
        cm = torch.zeros(K, K, dtype=torch.int64)
        for t in sweep.shard_range(n_tiles, rank, world):
            cm += torch.from_numpy(ref_ops.ref_confusion(gts[t], prs[t], K)).long()
        sweep.all_reduce_sum_(cm)
        full = sum(ref_ops.ref_confusion(gts[t], prs[t], K) for t in range(n_tiles))
        ok_cm = bool(np.array_equal(cm.numpy().astype(np.float64), full))
        # prototype mean: 5 support images split 3 / 2
        protos = torch.from_numpy(rng.standard_normal((5, 16)).astype(np.float32))
        mine = sweep.shard_range(5, rank, world)
        s = protos[mine.start:mine.stop].sum(0)
        c = torch.tensor([float(len(mine))])
        mean = sweep.prototype_mean_all_reduce_(s, c)
        ok_proto = bool(torch.allclose(mean, protos.mean(0), atol=1e-6))
        out.put((rank, ok_cm, ok_proto, int(cm.sum())))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_confusion_and_prototypes():
    world, n_tiles = 2, 7
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_tiles, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] and r[2] for r in results)
    assert results[0][3] == results[1][3] > 0
