"""Fused confusion inside the upsample kernel vs a separate sl_confusion pass over (label, pred)."""
import sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth
T, K = 32, 8
st = synth.make_trained_like_state(512, 7, 0, seed=1234)
labels_h = synth.make_labels(4, 1024, 1024, K, seed=1234)
feats_h = synth.make_features(labels_h, st, 8, seed=1234)
head = ops.PopHead(st.base_emb, st.cls)
feats = feats_h.cuda().repeat(8, 1, 1, 1)[:T].contiguous()
labels = labels_h.cuda().repeat(8, 1, 1)[:T].contiguous()
logits = head(feats)
cm = torch.zeros(K, K, dtype=torch.int64, device='cuda')
def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
def fused(): ops.upsample_argmax(logits, (1024, 1024), label=labels, cm=cm)
def split():
    p = ops.upsample_argmax(logits, (1024, 1024))['pred']
    ops.confusion_update(cm, labels, p)
def pred_only(): ops.upsample_argmax(logits, (1024, 1024))
print(f'fused {timeit(fused):.4f} ms  split {timeit(split):.4f} ms  pred only {timeit(pred_only):.4f} ms  (32 tiles)')
