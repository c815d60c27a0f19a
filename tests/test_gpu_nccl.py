"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the sharded sweep's confusion matrix equals the single-rank
matrix bit for bit (one int64 all-reduce, sweep.TileEvaluator.finalize; the reference never reduces its eval matrix,
eval_base.py:132-133,201), and the sharded masked-average-pooling prototypes (sweep.novel_prototypes_from_support,
networks/pspnet.py:7-15 per image) equal the single-rank ones within 1e-6."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import numpy as np
    import torch.distributed as dist
    from segland_b200 import ops, sweep, synth
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        dev = torch.device('cuda', rank)
        # ---- sharded eval sweep
        n_tiles, C = 7, 64                                      # odd count: shards of 4 and 3
        st = synth.make_trained_like_state(C, 7, 4, seed=11, n_bg_units=16)
        labels = synth.make_labels(n_tiles, 256, 256, st.n_classes, seed=11, coarse=8)
        feats = synth.make_features(labels, st, 8, seed=11)
        head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n, device=dev)
        mine = sweep.shard_range(n_tiles, rank, world)
        ev = sweep.TileEvaluator(head, (256, 256))
        for t in mine:
            ev.step(feats[t:t + 1].to(dev), labels[t:t + 1].to(dev))
        cm, (base, novel, total, _) = ev.finalize(base_classes=7)
        single = sweep.TileEvaluator(head, (256, 256))
        single.step(feats.to(dev), labels.to(dev))
        assert torch.equal(cm, single.cm), f'rank {rank}: sharded confusion matrix differs from the single-rank one'
        assert int(cm.sum()) == int((labels != 255).sum())
        # ---- sharded 5-shot prototypes: 4 novel classes x 5 shots = 20 support tiles
        n_sup, Kn = 20, 4
        sf = synth.make_random_features(n_sup, C, 32, 32, seed=5)
        sm = synth.make_support_masks(n_sup, 256, 256, seed=5)
        cls_of = torch.tensor([i // 5 for i in range(n_sup)])
        idx = list(sweep.shard_range(n_sup, rank, world))
        sharded = sweep.novel_prototypes_from_support(sf[idx].to(dev), sm[idx].to(dev), cls_of[idx], Kn)
        ref = torch.stack([ops.masked_average_pooling(sf[5 * k:5 * k + 5].to(dev), sm[5 * k:5 * k + 5].to(dev)).view(-1)
                           for k in range(Kn)])
        err = (sharded - ref).abs().max().item() / ref.abs().max().item()
        assert err <= 1e-6, f'rank {rank}: sharded prototypes off by {err}'
        # ---- validate()'s all-reduce pattern (ft_pop.py:276-277): inter / union areas
        pred = torch.randint(0, 12, (2, 64, 64), device=dev, generator=torch.Generator(dev).manual_seed(rank))
        tgt = torch.randint(0, 12, (2, 64, 64), device=dev, generator=torch.Generator(dev).manual_seed(10 + rank))
        inter, union, _ = ops.intersectionAndUnionGPU(pred.clone(), tgt, 12)
        sweep.all_reduce_sum_(inter); sweep.all_reduce_sum_(union)
        if rank == 0:
            np.save(os.path.join(out_dir, 'ok.npy'), np.array([float(total), err, float(inter.sum().item())]))
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_sweep_and_prototypes(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    import numpy as np
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    res = np.load(os.path.join(str(tmp_path), 'ok.npy'))
    assert res[0] > 0.3 and res[1] <= 1e-6
