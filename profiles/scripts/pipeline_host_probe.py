import sys, time, torch
sys.path.insert(0, '.')
from segland_b200 import ops, synth, sweep
dev = torch.device('cuda', 0)
st = synth.make_trained_like_state(512, 7, 0, seed=1234)
K = st.n_classes
labels_h = synth.make_labels(8, 1024, 1024, K, seed=1234)
feats_h = synth.make_features(labels_h, st, 8, seed=1234)
feats = feats_h.to(dev).repeat(4, 1, 1, 1).contiguous()
labels = labels_h.to(dev).repeat(4, 1, 1).contiguous()
head = ops.PopHead(st.base_emb, st.cls, None, None, device=dev)
for name, ev in (('sequential', sweep.TileEvaluator(head, (1024, 1024))), ('pipelined', sweep.PipelinedTileEvaluator(head, (1024, 1024)))):
    for _ in range(10): ev.step(feats, labels)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n): ev.step(feats, labels)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'{name}: host {1e6*(t1-t0)/n:.0f} us/step, total {1e6*(t2-t0)/n:.0f} us/step', flush=True)
import cProfile, pstats
ev = sweep.PipelinedTileEvaluator(head, (1024, 1024))
for _ in range(10): ev.step(feats, labels)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(100): ev.step(feats, labels)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
