"""Pins oracle/ref_ops.py against outputs of the reference's own functions
(tests/golden/*.npz, written by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ref_ops
from helpers import bf16_from_bits, rel_err, state_from_npz

HEAD_CASES = ['head_base_c64', 'head_ft_c64', 'head_base_c512', 'head_ft_c192_s4', 'head_ft_c96_rand']


@pytest.mark.parametrize('name', HEAD_CASES)
def test_head_matches_reference(golden, name):
    z = golden(name)
    st = state_from_npz(z)
    feats = bf16_from_bits(z['feats_bf16_bits']).float()
    logits = ref_ops.ref_head(feats, st.base_emb, st.novel_emb, st.cls, st.cls_n)
    assert logits.shape == z['logits'].shape
    # same ops in the same order on the same CPU -> bit-exact
    assert np.array_equal(logits.numpy(), z['logits'])
    H, W = z['labels'].shape[-2:]
    pred = ref_ops.ref_upsample_argmax(logits, (H, W))
    assert np.array_equal(pred, z['pred'])
    cm = sum(ref_ops.ref_confusion(z['labels'][t], pred[t], st.n_classes) for t in range(pred.shape[0]))
    assert np.array_equal(cm, z['cm'])
    assert cm.dtype == np.float64 and cm.sum() == (z['labels'] != 255).sum()
    if 'upsampled' in z.files:
        assert np.array_equal(ref_ops.ref_upsample(logits, (H, W)).numpy(), z['upsampled'])


def test_upsample_cases(golden):
    z = golden('upsample_cases')
    i = 0
    while f'in{i}' in z.files:
        lg = torch.from_numpy(z[f'in{i}'])
        H, W = z[f'out{i}'].shape[-2:]
        assert np.array_equal(ref_ops.ref_upsample(lg, (H, W)).numpy(), z[f'out{i}'])
        assert np.array_equal(ref_ops.ref_upsample_argmax(lg, (H, W)), z[f'pred{i}'])
        i += 1
    assert i >= 6


def test_metrics(golden):
    z = golden('metrics')
    K = 12
    cm = ref_ops.ref_confusion(z['gt'], z['pred'], K)
    assert np.array_equal(cm, z['cm'])
    out = torch.from_numpy(z['pred'].astype(np.int64))
    tgt = torch.from_numpy(z['gt'].astype(np.int64))
    inter, union, target = ref_ops.ref_inter_union(out, tgt, K, 255)
    for mine, key in ((inter, 'inter'), (union, 'union'), (target, 'target')):
        assert np.array_equal(mine.numpy(), z[key + '_t'])
        assert np.array_equal(mine.numpy().astype(np.int64), z[key + '_np'])
    assert np.array_equal(out.numpy(), z['output_after'])          # in-place side effect
    # inter/union are the diagonal / row+col-diag of the confusion matrix
    assert np.array_equal(np.diag(cm), z['inter_np'])
    assert np.array_equal(cm.sum(0) + cm.sum(1) - np.diag(cm), z['union_np'])
    base, novel, total, arr = ref_ops.ref_miou(cm, 7)
    assert 0 < base < 1 and 0 < novel < 1 and abs(total - np.nanmean(arr)) < 1e-15


def test_map(golden):
    z = golden('map')
    p = ref_ops.ref_masked_average_pooling(bf16_from_bits(z['feats_bf16_bits']).float(), torch.from_numpy(z['masks']))
    assert np.array_equal(p.numpy(), z['proto'])
    p2 = ref_ops.ref_masked_average_pooling(bf16_from_bits(z['feats2_bf16_bits']).float(), torch.from_numpy(z['masks2']))
    assert np.array_equal(p2.numpy(), z['proto2'])


def test_orth_and_pseudo(golden):
    z = golden('orth_pseudo')
    stb, stf = state_from_npz(z, 'b_'), state_from_npz(z, 'f_')
    sim_b = ref_ops.ref_proto_sim_base(stb.base_emb)
    assert np.allclose(sim_b.numpy(), z['sim_b'], rtol=0, atol=1e-7)
    assert abs(ref_ops.ref_orth_loss(sim_b).item() - float(z['orth_b'])) < 1e-7
    sim_f = ref_ops.ref_proto_sim_ft(stf.novel_emb, stf.base_emb)
    assert np.array_equal(sim_f.numpy(), z['sim_f'])
    assert ref_ops.ref_orth_loss(sim_f).item() == float(z['orth_f'])
    assert sim_f.shape == (4, 11) and int(torch.triu(torch.ones_like(sim_f), 1).sum()) == 34
    # pseudo-labelling of the base images (second half of the batch in forward_novel)
    preds_all = torch.from_numpy(z['preds_all'])                   # eval-mode forward_all == train-mode preds
    B = preds_all.shape[0]
    preds2_base = torch.cat([preds_all[B // 2:, :1], preds_all[B // 2:, 1 + 7:]], dim=1)
    mask_b = torch.from_numpy(z['mask_b_before'].copy())
    new = ref_ops.ref_pseudo_label(preds2_base, mask_b, 7)
    assert np.array_equal(new.numpy(), z['mask_b_after'])
    assert np.array_equal(mask_b.numpy(), z['mask_b_after'])       # in place, like the reference


def test_fuse(golden):
    z = golden('fuse')
    for tile in ('tile_a', 'tile_b'):
        mats = [z[f'{tile}_m{m}'] for m in range(3)]
        pred, fused = ref_ops.ref_fuse(mats)
        assert fused.dtype == np.float32
        assert np.array_equal(pred, z[f'{tile}_pred'])


def test_seg_ce(golden):
    z = golden('ce')
    for i in range(3):
        preds = torch.from_numpy(z[f'preds{i}']).requires_grad_(True)
        target = torch.from_numpy(z[f'target{i}'])
        loss = ref_ops.ref_seg_ce(preds, target)
        assert loss.item() == float(z[f'loss{i}'])
        loss.backward()
        assert np.array_equal(preds.grad.numpy(), z[f'grad{i}'])


def _grads_of(loss, tensors):
    gs = torch.autograd.grad(loss, tensors, allow_unused=True)
    return [None if g is None else g.detach() for g in gs]


@pytest.mark.parametrize('name', ['ft_c64', 'ft_c64b', 'ft_c96'])
def test_forward_novel_gradients(golden, name):
    """Oracle forward_novel + OrthLoss + autograd == the reference's (train mode, identity backbone)."""
    z = golden('train_grads')
    st = state_from_npz(z, name + '_')
    req = lambda t: t.clone().requires_grad_(True)
    novel, cls, cls_n = req(st.novel_emb), tuple(req(t) for t in st.cls), tuple(req(t) for t in st.cls_n)
    img_n, img_b = req(bf16_from_bits(z[name + '_img_n_bits']).float()), req(bf16_from_bits(z[name + '_img_b_bits']).float())
    mask_b = torch.from_numpy(z[name + '_mask_b_before'].copy())
    loss, _ = ref_ops.ref_forward_novel(torch.cat([img_n, img_b], 0), torch.from_numpy(z[name + '_mask_n']), mask_b,
                                        st.base_emb, novel, cls, cls_n)
    assert np.array_equal(mask_b.numpy(), z[name + '_mask_b_after'])
    for key in ('total', 'seg', 'orth'):
        assert abs(loss[key + '_loss'].item() - float(z[f'{name}_{key}'])) < 2e-6, key
    got = _grads_of(loss['total_loss'], [novel, *cls, *cls_n, img_n, img_b])
    want = ['g_novel_emb', 'g_W1', 'g_W2', 'g_w3', 'g_W1n', 'g_W2n', 'g_w3n', 'g_img_n', 'g_img_b']
    for g, key in zip(got, want):
        ref = torch.from_numpy(z[f'{name}_{key}']).reshape(g.shape)
        assert (g - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-9, key


def test_forward_base_gradients(golden):
    z = golden('train_grads')
    st = state_from_npz(z, 'base_c64_')
    req = lambda t: t.clone().requires_grad_(True)
    base, cls = req(st.base_emb), tuple(req(t) for t in st.cls)
    img = req(bf16_from_bits(z['base_c64_img_bits']).float())
    loss, _ = ref_ops.ref_forward_base_loss(img, torch.from_numpy(z['base_c64_mask']).long(), base, cls)
    for key in ('total', 'seg', 'orth'):
        assert abs(loss[key + '_loss'].item() - float(z[f'base_c64_{key}'])) < 2e-6, key
    got = _grads_of(loss['total_loss'], [base, *cls, img])
    for g, key in zip(got, ['g_base_emb', 'g_W1', 'g_W2', 'g_w3', 'g_img']):
        ref = torch.from_numpy(z[f'base_c64_{key}']).reshape(g.shape)
        assert (g - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-9, key


# ------------------------------------------------------------------ decoder tails
def test_tails_match_reference_decoders(golden):
    """oracle tail functions vs what the reference's PSPModule / FPN_Seg_OCR_Decoder / UperNet_Decoder_Plus computed
    from the hooked tail inputs (oracle/gen_golden_tails.py)."""
    z = golden('tails')
    t = lambda k: torch.from_numpy(z[k].copy())
    for name in ('psp_c64', 'psp_c512', 'psp_c96'):
        out = ref_ops.ref_tail_bn_relu_conv(t(name + '_x'), t(name + '_bn_w'), t(name + '_bn_b'), t(name + '_bn_m'),
                                            t(name + '_bn_v'), float(z[name + '_bn_eps']), t(name + '_W'), t(name + '_bias'))
        # same ATen kernels; the convolution's blocking may differ with the thread count, hence 1e-6 not bit-exact
        assert rel_err(out, z[name + '_out']) <= 1e-6, name
    for name in ('ln_c192', 'ln_c96', 'ln_c480', 'ln_c40'):
        out = ref_ops.ref_tail_layernorm(t(name + '_x'), t(name + '_gamma'), t(name + '_beta'), float(z[name + '_eps']))
        assert np.array_equal(out.numpy(), z[name + '_out']), name
    maps = [t(f'sum_swin_map{i}') for i in range(4)]
    assert np.array_equal(ref_ops.ref_tail_sum(maps).numpy(), z['sum_swin_out'])
    out = ref_ops.ref_tail_bn_relu(t('aspp_x'), t('aspp_bn_w'), t('aspp_bn_b'), t('aspp_bn_m'), t('aspp_bn_v'),
                                   float(z['aspp_bn_eps']))
    assert np.array_equal(out.numpy(), z['aspp_out'])
    assert np.array_equal(ref_ops.ref_tail_concat([t(f'cat_hr_map{i}') for i in range(4)]).numpy(), z['cat_hr_out'])
    # the report helper: ideal rounding is <= 0.5 spacings everywhere and 100 % identical patterns
    ref = t('ln_c96_out')
    exact, d = ref_ops.bf16_ulp_report(ref.to(torch.bfloat16), ref)
    assert exact == 1.0 and d.max().item() <= 0.5


def test_full_size_head_matches_reference(golden):
    """The oracle at configs[1] size against the reference's own full-tile logits / prediction / confusion matrix
    (oracle/gen_golden_full.py); features come from the seeded generator and are checked by digest."""
    import hashlib
    from oracle import gen_golden_full
    z = golden('head_base_c512_full')
    kind, _, head_seed, data_seed = gen_golden_full.CASES[0]
    st, labels, feats = gen_golden_full.make_case(kind, head_seed, data_seed)
    assert hashlib.sha256(feats.view(torch.int16).numpy().tobytes()).hexdigest() == str(z[f'{kind}_feat_sha256'])
    pred, cm, logits = ref_ops.ref_eval_tile(feats.float(), labels.numpy(), st.base_emb, None, st.cls, None, (1024, 1024),
                                             st.n_classes)
    assert np.array_equal(logits.numpy(), z[f'{kind}_logits'])
    assert np.array_equal(pred, z[f'{kind}_pred']) and np.array_equal(cm, z[f'{kind}_cm'])
