import sys, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth, _cabi
dev = torch.device('cuda', 0)
st = synth.make_trained_like_state(512, 7, 0, seed=1234)
K = st.n_classes
labels_h = synth.make_labels(1, 1024, 1024, K, seed=1234)
feats = synth.make_features(labels_h, st, 8, seed=1234).to(dev)
labels = labels_h.to(dev)
head = ops.PopHead(st.base_emb, st.cls, None, None, device=dev)
lg = torch.empty(1, K, 128, 128, device=dev)
cm = torch.zeros(K, K, dtype=torch.int64, device=dev)
def t(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
print('fg  us', t(lambda: head(feats, out=lg, fg_only=True)))
print('bg  us', t(lambda: head.bg_tc(feats, lg)))
print('post us', t(lambda: ops.upsample_argmax(lg, (1024, 1024), label=labels, cm=cm)))
for B in (2, 4, 8):
    f = feats.repeat(B, 1, 1, 1).contiguous(); l = torch.empty(B, K, 128, 128, device=dev)
    print(B, 'fg us/tile', t(lambda: head(f, out=l, fg_only=True)) / B, 'bg us/tile', t(lambda: head.bg_tc(f, l)) / B)
