"""One launch of the prediction + confusion path on the bench data (for ncu captures)."""
import sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth
T, K = 32, 8
st = synth.make_trained_like_state(512, 7, 0, seed=1234)
labels_h = synth.make_labels(8, 1024, 1024, K, seed=1234)
feats_h = synth.make_features(labels_h, st, 8, seed=1234)
head = ops.PopHead(st.base_emb, st.cls)
feats = feats_h.cuda().repeat(4, 1, 1, 1)[:T].contiguous()
labels = labels_h.cuda().repeat(4, 1, 1)[:T].contiguous()
logits = head(feats)
cm = torch.zeros(K, K, dtype=torch.int64, device='cuda')
for _ in range(3): ops.upsample_argmax(logits, (1024, 1024), label=labels, cm=cm)
torch.cuda.synchronize()
