"""GPU: sweep.PipelinedTileEvaluator (head of batch t on one stream, post-processing of batch t-1 underneath it on a
second stream) must return exactly what TileEvaluator returns -- per-batch predictions, the confusion matrix and the
mIoU split of eval_base.py:193-199 -- for device and pinned-host inputs, base and ft heads, with and without labels."""
import numpy as np
import pytest
import torch

from oracle import ref_ops
from segland_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mods():
    from segland_b200 import ops, sweep
    ops.check_device()
    return ops, sweep


def make_batches(st, stride, n_batches, B, seed, size=1024):
    out = []
    for i in range(n_batches):
        labels = synth.make_labels(B, size, size, st.n_classes, seed=seed + i)
        out.append((synth.make_features(labels, st, stride, seed=seed + i), labels))
    return out


@pytest.mark.parametrize('C,Kn,stride,B', [(512, 0, 8, 8), (512, 0, 8, 1), (192, 4, 4, 2), (96, 4, 4, 1)])
def test_pipelined_sweep_equals_sequential(mods, C, Kn, stride, B):
    ops, sweep = mods
    st = synth.make_trained_like_state(C, 7, Kn, seed=11)
    head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n)
    batches = make_batches(st, stride, 5, B, seed=40)
    seq = sweep.TileEvaluator(head, (1024, 1024))
    ref_preds = [seq.step(f.cuda(), l.cuda())['pred'].clone() for f, l in batches]
    cm_ref, miou_ref = seq.finalize(base_classes=7)
    for host in (False, True):
        pipe = sweep.PipelinedTileEvaluator(head, (1024, 1024))
        got = []
        for f, l in batches:
            if host:
                f, l = f.pin_memory(), l.pin_memory()
            else:
                f, l = f.cuda(), l.cuda()
            r = pipe.step(f, l)
            if r is not None:
                got.append(r['pred'].clone())
        assert len(got) == len(batches) - 1                      # one batch in flight
        got.append(pipe.flush()['pred'].clone())
        assert pipe.flush() is None
        cm, miou = pipe.finalize(base_classes=7)
        for a, b in zip(got, ref_preds):                         # the same kernels in a different schedule: identical
            assert torch.equal(a, b)
        assert torch.equal(cm, cm_ref)
        assert np.array_equal(np.asarray(miou[:3], dtype=np.float64), np.asarray(miou_ref[:3], dtype=np.float64), equal_nan=True)
    # and against the oracle's confusion matrix on the returned maps
    want = sum(ref_ops.ref_confusion(l[i].numpy(), p[i].cpu().numpy(), st.n_classes)
               for (f, l), p in zip(batches, ref_preds) for i in range(B))
    assert np.array_equal(cm_ref.cpu().numpy().astype(np.float64), want)


def test_pipelined_reset_and_unlabelled_batches(mods):
    ops, sweep = mods
    st = synth.make_trained_like_state(512, 7, 0, seed=12)
    head = ops.PopHead(st.base_emb, st.cls, None, None)
    batches = make_batches(st, 8, 3, 2, seed=50)
    pipe = sweep.PipelinedTileEvaluator(head, (1024, 1024))
    seq = sweep.TileEvaluator(head, (1024, 1024))
    assert pipe.step(batches[0][0].cuda(), batches[0][1].cuda()) is None
    pipe.reset()                                                 # flushes the pending batch, then clears the counts
    assert int(pipe.cm.sum()) == 0
    r0 = pipe.step(batches[1][0].cuda())                         # no labels: prediction only, nothing counted
    assert r0 is None
    r1 = pipe.step(batches[2][0].cuda(), batches[2][1].cuda(), want_probs=True)
    assert torch.equal(r1['pred'], seq.step(batches[1][0].cuda())['pred'])
    r2 = pipe.flush()
    s2 = seq.step(batches[2][0].cuda(), batches[2][1].cuda(), want_probs=True)
    assert torch.equal(r2['pred'], s2['pred']) and torch.equal(r2['probs'], s2['probs'])
    assert torch.equal(pipe.cm, seq.cm)



def test_pipelined_ragged_last_batch(mods):
    """80 tiles in batches of 32 end with a batch of 16: buffers are re-allocated mid-sweep while both streams are busy."""
    ops, sweep = mods
    st = synth.make_trained_like_state(512, 7, 0, seed=13)
    head = ops.PopHead(st.base_emb, st.cls, None, None)
    labels = synth.make_labels(10, 1024, 1024, st.n_classes, seed=60)
    feats = synth.make_features(labels, st, 8, seed=60).cuda()
    labels = labels.cuda()
    seq = sweep.TileEvaluator(head, (1024, 1024))
    pipe = sweep.PipelinedTileEvaluator(head, (1024, 1024))
    ref, got = [], []
    for rep in range(3):                                         # 4 + 4 + 2, three times over
        for lo, hi in ((0, 4), (4, 8), (8, 10)):
            ref.append(seq.step(feats[lo:hi], labels[lo:hi])['pred'].clone())
            r = pipe.step(feats[lo:hi], labels[lo:hi])
            if r is not None:
                got.append(r['pred'].clone())
            junk = torch.full((4, 8, 128, 128), float(rep), device='cuda')      # churn the allocator on the caller's stream
            del junk
    got.append(pipe.flush()['pred'].clone())
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    assert torch.equal(pipe.finalize(7)[0], seq.finalize(7)[0])


def test_evaluators_with_view_and_window_reduction(mods):
    """reduce= hook: flip-TTA averaging and sliding-window stitching run between the head and the up-sampling (on the
    post-processing stream of the pipelined evaluator); probability maps included.  Both evaluators == the manual
    composition of the same operators."""
    ops, sweep = mods
    st = synth.make_head_state(192, 7, 4, seed=3)
    head = ops.PopHead(st.base_emb, st.cls, st.novel_emb, st.cls_n)
    K, T, hw = head.n_classes, 2, 64
    g = torch.Generator().manual_seed(9)
    labels = torch.randint(0, K, (T, 256, 256), generator=g).to(torch.uint8).cuda()
    # (a) two views per tile, view-major
    feats = [torch.randn(2 * T, 192, hw, hw, generator=g).to(torch.bfloat16).cuda() for _ in range(3)]
    red = lambda lg: ops.aggregate_views(lg.view(2, T, K, hw, hw), [0, 1])
    manual = [ops.upsample_argmax(red(head(f)), (256, 256), want_probs=True) for f in feats]
    for cls in (sweep.TileEvaluator, sweep.PipelinedTileEvaluator):
        ev = cls(head, (256, 256), reduce=red)
        outs = []
        for f in feats:
            r = ev.step(f, labels, want_probs=True)
            if r is not None:
                outs.append({k: v.clone() for k, v in r.items()})
        if cls is sweep.PipelinedTileEvaluator:
            outs.append({k: v.clone() for k, v in ev.flush().items()})
        assert len(outs) == 3
        for o, m in zip(outs, manual):
            assert torch.equal(o['pred'], m['pred']) and torch.equal(o['probs'], m['probs'])
        cm, _ = ev.finalize(7)
        assert int(cm.sum()) == 3 * T * 256 * 256
    # (b) sliding windows + flips
    plan = ops.WindowPlan((256, 256), (128, 128), (96, 96), 4)
    flips = (0, 1)
    E = plan.n_windows * len(flips)
    hc, wc = plan.crop_lr_hw
    crops = [torch.randn(T * E, 192, hc, wc, generator=g).to(torch.bfloat16).cuda() for _ in range(2)]
    redw = lambda lg: ops.window_accumulate(lg.view(T, E, K, hc, wc), plan, flips)
    manual = [ops.upsample_argmax(redw(head(c)), (256, 256)) for c in crops]
    pev = sweep.PipelinedTileEvaluator(head, (256, 256), reduce=redw)
    assert pev.step(crops[0]) is None
    a = pev.step(crops[1])['pred'].clone()
    b = pev.flush()['pred'].clone()
    assert torch.equal(a, manual[0]['pred']) and torch.equal(b, manual[1]['pred'])
