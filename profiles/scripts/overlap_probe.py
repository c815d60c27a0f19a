#!/usr/bin/env python
"""Does the up-sample/arg-max kernel hide under the background MLP?  configs[1] pass over 32 resident tiles:
   sequential:  fg(p) -> bg(p) -> post(p)                              (one stream, the round-1/2 bench pass)
   pipelined:   fg(p) -> bg(p) on the main stream, post(p-1) on a second stream launched right after bg(p), with
                double-buffered low-res logits: the post kernel's 128-thread CTAs (16 K registers, 256 B smem) fit on an
                SM next to the pair kernel's CTA, so its issue-bound work runs in the slots the tensor-bound kernel leaves idle.
Long timed regions (power-capped clocks), CUDA events, predictions and confusion matrices compared between the two."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from segland_b200 import _cabi, ops, synth, sweep  # noqa: E402
from bench import ClockSampler, physical_gpu_index  # noqa: E402

C, KB, TILE, STRIDE, HW = 512, 7, 1024, 8, 128


def main():
    dev = torch.device('cuda', 0)
    T = 32
    passes = int(os.environ.get('PROBE_PASSES', '1000'))
    st = synth.make_trained_like_state(C, KB, 0, seed=1234)
    K = st.n_classes
    labels_h = synth.make_labels(8, TILE, TILE, K, seed=1234)
    feats_h = synth.make_features(labels_h, st, STRIDE, seed=1234)
    feats = feats_h.to(dev).repeat(4, 1, 1, 1).contiguous()
    labels = labels_h.to(dev).repeat(4, 1, 1).contiguous()
    head = ops.PopHead(st.base_emb, st.cls, None, None, device=dev)
    lg = [torch.empty(T, K, HW, HW, dtype=torch.float32, device=dev) for _ in range(2)]
    pred = [torch.empty(T, TILE, TILE, dtype=torch.uint8, device=dev) for _ in range(2)]
    hi = torch.cuda.Stream(device=dev, priority=-1)             # fg / bg: placed first when both kernels have CTAs pending
    torch.cuda.set_stream(hi)
    main_s = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev, priority=0)

    no_cm = bool(os.environ.get('PROBE_NOCM'))                # prediction only: how much of the interference is the counting?

    def post(buf, cm, stream):
        _cabi.call('sl_upsample_argmax', _cabi.ptr(lg[buf]), T, K, HW, HW, TILE, TILE, None if no_cm else _cabi.ptr(labels), 255,
                   _cabi.ptr(pred[buf]), None, None, None, None if no_cm else _cabi.ptr(cm), stream.cuda_stream)

    def sequential(n, cm):
        for p in range(n):
            head(feats, out=lg[0], fg_only=True)
            head.bg_tc(feats, lg[0])
            post(0, cm, main_s)

    trace = {}

    def pipelined(n, cm):
        """post(p-1) becomes eligible together with bg(p): it waits (on the low-priority side stream) for fg(p)."""
        tr = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        done_fg = [torch.cuda.Event() for _ in range(2)]
        done_bg = [torch.cuda.Event() for _ in range(2)]
        done_post = [torch.cuda.Event() for _ in range(2)]
        for p in range(n):
            b = p & 1
            if p >= 2:
                main_s.wait_event(done_post[b])              # logits buffer b is free again
            rec = p == n // 2
            if rec: tr[0].record(main_s)
            head(feats, out=lg[b], fg_only=True)
            done_fg[b].record(main_s)
            if rec: tr[1].record(main_s)
            head.bg_tc(feats, lg[b])
            if rec: tr[2].record(main_s)
            done_bg[b].record(main_s)
            if p >= 1:
                side.wait_event(done_bg[b ^ 1])              # (implied by the next wait; kept for clarity)
                side.wait_event(done_fg[b])
                if rec: tr[3].record(side)
                post(b ^ 1, cm, side)
                if rec: tr[4].record(side)
                done_post[b ^ 1].record(side)
        side.wait_event(done_bg[(n - 1) & 1])
        post((n - 1) & 1, cm, side)
        main_s.wait_stream(side)
        trace['ev'] = tr

    results = {}
    for name, fn, env in (('sequential, 256-thread post CTAs', sequential, 1), ('sequential, 128-thread post CTAs', sequential, 3),
                          ('pipelined, 128-thread post CTAs', pipelined, 3),
                          ('sequential, row-cached post', sequential, 0), ('pipelined, row-cached post', pipelined, 0)):
        _cabi.set_env(SL_POST_REGS=env)
        cm = torch.zeros(K, K, dtype=torch.int64, device=dev)
        fn(8, cm)
        torch.cuda.synchronize()
        cm.zero_()
        sampler = ClockSampler(physical_gpu_index(0), period=0.01)
        sampler.start()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.begin()
        a.record(main_s)
        fn(passes, cm)
        b.record(main_s)
        torch.cuda.synchronize()
        print('    clocks', sampler.finish())
        ms = a.elapsed_time(b) / passes
        results[name] = (cm.clone(), pred[0].clone(), pred[1].clone())
        if fn is pipelined:
            tr = trace['ev']
            print(f'    mid-run pass: fg {tr[0].elapsed_time(tr[1]):.3f} ms, bg {tr[1].elapsed_time(tr[2]):.3f} ms, '
                  f'post (side stream) {tr[3].elapsed_time(tr[4]):.3f} ms, post start - bg start {tr[1].elapsed_time(tr[3]):+.3f} ms, '
                  f'post end - bg end {tr[2].elapsed_time(tr[4]):+.3f} ms')
        print(f'{name:36s}: {ms:7.4f} ms per pass = {T / ms * 1e3:9.0f} tiles/s  (region {a.elapsed_time(b) / 1e3:.2f} s)', flush=True)
    _cabi.set_env(SL_POST_REGS=None)
    ref = results['sequential, row-cached post']
    for name, r in results.items():
        ok = torch.equal(r[0], ref[0]) and torch.equal(r[1], ref[1])
        print(f'{name:36s}: confusion matrix and predictions identical to the sequential row-cached run: {ok}')


if __name__ == '__main__':
    main()
