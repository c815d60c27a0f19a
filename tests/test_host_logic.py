"""CPU: host-side logic that needs no device -- sharding, mIoU reduction, state-dict handling,
synthetic generators."""
import numpy as np
import torch

from oracle import ref_ops
from segland_b200 import ops, sweep, synth


def test_shard_range_partitions():
    for n in (0, 1, 7, 80, 81, 1000):
        for world in (1, 2, 3, 8):
            parts = [sweep.shard_range(n, r, world) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1


def test_miou_matches_oracle():
    rng = np.random.default_rng(0)
    cm = rng.integers(0, 1000, size=(12, 12)).astype(np.float64)
    cm[9, :] = 0
    cm[:, 9] = 0                      # an absent class -> NaN IoU, skipped by nanmean
    mine = ops.miou_from_confusion(cm, 7)
    ref = ref_ops.ref_miou(cm, 7)
    for a, b in zip(mine[:3], ref[:3]):
        assert a == b
    assert np.array_equal(mine[3], ref[3], equal_nan=True)
    assert np.isnan(mine[3][9])
    mine_t = ops.miou_from_confusion(torch.from_numpy(cm.astype(np.int64)), 7)
    assert mine_t[2] == ref[2]


def test_synth_is_seeded_and_shaped():
    st = synth.make_head_state(64, 7, 4, seed=3)
    st2 = synth.make_head_state(64, 7, 4, seed=3)
    assert torch.equal(st.base_emb, st2.base_emb) and torch.equal(st.cls_n[0], st2.cls_n[0])
    assert st.n_classes == 12
    gram = st.base_emb @ st.base_emb.t()
    assert torch.allclose(gram, torch.eye(7), atol=1e-5)          # orthogonal init
    lab = synth.make_labels(2, 64, 64, 12, seed=3, coarse=8)
    assert lab.dtype == torch.uint8 and lab.shape == (2, 64, 64)
    assert set(np.unique(lab.numpy()).tolist()) <= set(range(12)) | {255}
    f = synth.make_features(lab, st, 8, seed=3)
    assert f.dtype == torch.bfloat16 and f.shape == (2, 64, 8, 8)


def test_oracle_collapse_identity():
    """The algebra the CUDA path relies on, checked against the materialising oracle:
    fg logit = p>=0 ? p*alpha : -p*beta, bg logit = MLP(W1 (I - S^T S) q)."""
    st = synth.make_head_state(64, 7, 4, seed=5)
    lab = synth.make_labels(1, 64, 64, 12, seed=5, coarse=8)
    feats = synth.make_features(lab, st, 8, seed=5).float()
    ref = ref_ops.ref_head(feats, st.base_emb, st.novel_emb, st.cls, st.cls_n)
    protos = torch.cat([st.base_emb, st.novel_emb], 0)
    s = torch.nn.functional.normalize(protos, dim=-1)
    q = feats.flatten(2)[0]                                           # [C,N]
    p = s @ q
    mlp = lambda ws, x: ws[2] @ torch.relu(ws[1] @ torch.relu(ws[0] @ x))
    out = torch.empty(12, q.shape[1])
    for k in range(11):
        ws = st.cls if k < 7 else st.cls_n
        a, b = mlp(ws, s[k][:, None]).squeeze(), mlp(ws, -s[k][:, None]).squeeze()
        out[1 + k] = torch.where(p[k] >= 0, p[k] * a, -p[k] * b)
    W1p = st.cls_n[0] - (st.cls_n[0] @ s.t()) @ s
    out[0] = mlp((W1p, st.cls_n[1], st.cls_n[2]), q)
    assert torch.allclose(out.view(1, 12, 8, 8), ref, rtol=0, atol=2e-6)
