#!/usr/bin/env python
"""Benchmark of the POP head + dense post-processing hot path (BASELINE.json metric:
1024^2 tiles/sec, head+postproc, device-timed; roofline fractions).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE configs[1]): PSPNet-POP base-class eval on synthetic OEM-shaped 1024x1024
tiles: features [T,512,128,128] bf16, 8 classes (bg + 7 base), labels [T,1024,1024] u8.
One "pass" = the whole path over T = 32 distinct resident tiles per GPU,
    sl_pop_fg_lowres -> sl_pop_bg_tc -> sl_upsample_argmax (+ confusion)
(sl_pop_prepare runs once per weight update, outside the sweep, as in PopHead); one "step" = --passes passes
(default 48: 1,536 tiles, ~55 ms), so K = 20 steps time > 1 s and the SM clocks are sampled > 50 times under load.
The sweep runs through sweep.PipelinedTileEvaluator: fg(p), bg(p) on one stream, the up-sampling / arg-max / confusion
kernel of pass p-1 on a second stream underneath bg(p) (the two kernels are sized to share an SM), flushed inside the
timed region; --no-pipeline runs the three kernels back to back, and a short second leg always does, so that the line
carries every kernel's duration alone as well (`sequential`, `roofline.frac_kernel_alone`).
T*16.8 MB of features is larger than the 126 MB L2, so every pass streams from HBM.  Weak scaling: every rank owns
its own T tiles; the only collective is one int64 all-reduce of the confusion matrix at the end of the sweep
(inside the timed region).

The same JSON line carries, in `configs`, device-timed numbers for the other BASELINE configs at the same N
(configs[2] 5-shot prototype update, configs[3] ConvNeXt / Swin flip + sliding-window inference with probability
maps, configs[4] fusion sweep), the single-tile latency (`latency_b1`), and the CPU / eager-GPU baselines.

The oracle (oracle/) is executed only for the cpu_baseline / gpu_eager_baseline legs and for --impl reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C, HW_LR, STRIDE, KB = 512, 128, 8, 7              # pspnet_pop / resnet50 geometry
TILE = HW_LR * STRIDE                              # 1024
N_PIX = HW_LR * HW_LR
WORKLOAD = 'PSPNet-POP base eval, 1024x1024 tiles, fused head+argmax+mIoU (BASELINE configs[1])'


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tflops_burst': d['bf16_tflops'],
                'tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'source': 'fallback'}


def load_traffic():
    for name in ('r2b_traffic.json', 'r2_traffic.json', 'r1_traffic.json'):
        p = os.path.join(ROOT, 'profiles', name)
        if os.path.exists(p):
            return json.load(open(p)), name
    return {}, None


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU (NVML).  It is created and started BEFORE the barrier in
    front of the timed region (NVML initialisation takes a rank-dependent few ms); `begin()` marks the first sample
    that belongs to the region."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self._first = 0
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:                                  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4),
                 'hw_power_brake': getattr(nv, 'nvmlClocksThrottleReasonHwPowerBrakeSlowdown', 0x80)}
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, [k for k, bit in names.items() if r & bit]))
            except Exception:                                   # noqa: BLE001
                pass
            self._stop_evt.wait(self.period)

    def begin(self):
        self._first = len(self.samples)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        region = self.samples[self._first:]
        if not region:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'note': 'no NVML samples'}
        reasons = sorted({k for _, ks in region for k in ks})
        return {'sm_mhz': statistics.median(m for m, _ in region), 'sm_max_mhz': self.max_mhz,
                'reasons': reasons, 'samples': len(region)}


def physical_gpu_index(local):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        try:
            return int(vis.split(',')[local])
        except (ValueError, IndexError):
            return local
    return local


def ev_pair():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def time_loop(fn, iters, warm=3, finish=None):
    """Average device time of fn() in seconds (CUDA events on the current stream, synchronised both sides); `finish`
    (e.g. the flush of a pipelined evaluator) runs once inside the timed region after the last fn()."""
    for _ in range(warm):
        fn()
    if finish is not None:
        finish()
    torch.cuda.synchronize()
    a, b = ev_pair()
    a.record()
    for _ in range(iters):
        fn()
    if finish is not None:
        finish()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def max_over_ranks(x, world, dev):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------ CPU reference
def cpu_reference_tiles_per_s(n_tiles, repeats=1, seed=1234):
    """The reference's CPU path for the same workload (oracle port of eval_base.py:166-178 after
    the decoder: head -> F.interpolate -> np.argmax -> get_confusion_matrix), all host threads."""
    import numpy as np
    from oracle import ref_ops
    from segland_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = synth.make_trained_like_state(C, KB, 0, seed=seed)
    labels = synth.make_labels(n_tiles, TILE, TILE, st.n_classes, seed=seed)
    feats = synth.make_features(labels, st, STRIDE, seed=seed).float()       # fp32 copy of the bf16 values
    K = st.n_classes
    times = []
    with torch.no_grad():
        ref_ops.ref_eval_tile(feats[:1], labels[:1].numpy(), st.base_emb, None, st.cls, None, (TILE, TILE), K)  # warm-up
        for _ in range(repeats):
            t0 = time.perf_counter()
            cm = np.zeros((K, K))
            for t in range(n_tiles):
                _, c, _ = ref_ops.ref_eval_tile(feats[t:t + 1], labels[t:t + 1].numpy(), st.base_emb, None,
                                                st.cls, None, (TILE, TILE), K)
                cm += c
            times.append(time.perf_counter() - t0)
    return n_tiles / statistics.median(times), cores


def gpu_eager_reference(n_tiles, dev, seed=1234):
    """The reference's LIVE path on the same GPU (eval_base.py:164-178 after the decoder): the materialising head in
    stock eager PyTorch on CUDA tensors, F.interpolate on the GPU, then -- as the scripts do -- a device-to-host copy
    of the full up-sampled logits, np.argmax and the numpy confusion matrix on the host.  fp32 and TF32 convolutions.
    Returns tiles/s and the per-tile latency (host-timed, one tile per forward: scripts/evaluate_oem.sh:16-17)."""
    from oracle import ref_ops
    from segland_b200 import synth
    st = synth.make_trained_like_state(C, KB, 0, seed=seed)
    labels = synth.make_labels(n_tiles, TILE, TILE, st.n_classes, seed=seed)
    feats = synth.make_features(labels, st, STRIDE, seed=seed).float().to(dev)
    base, cls = st.base_emb.to(dev), tuple(t.to(dev) for t in st.cls)
    K = st.n_classes
    out = {}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (('fp32', False), ('tf32', True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            lat = []
            with torch.no_grad():
                for t in range(-2, n_tiles):                     # two warm-up tiles
                    i = max(t, 0)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    ref_ops.ref_eval_tile(feats[i:i + 1], labels[i:i + 1].numpy(), base, None, cls, None, (TILE, TILE), K)
                    torch.cuda.synchronize()
                    if t >= 0:
                        lat.append(time.perf_counter() - t0)
            out[name] = {'tiles_per_s': n_tiles / sum(lat), 'latency_b1_us_p50': 1e6 * statistics.median(lat)}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    out['sample'] = (f'{n_tiles} tiles, one per forward: eager PyTorch head (orthogonal_decompose + 1x1-conv MLPs on the '
                     'materialised rank-1 tensors) + F.interpolate on the GPU, then D2H of the up-sampled logits, '
                     'np.argmax and numpy confusion on the host, as eval_base.py:164-178 does')
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    tiles_per_step = 1
    times = []
    import numpy as np
    from oracle import ref_ops
    from segland_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st = synth.make_trained_like_state(C, KB, 0, seed=1234)
    n_distinct = 4
    labels = synth.make_labels(n_distinct, TILE, TILE, st.n_classes, seed=1234)
    feats = synth.make_features(labels, st, STRIDE, seed=1234).float()
    K = st.n_classes
    cm = np.zeros((K, K))
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t = i % n_distinct
            t0 = time.perf_counter()
            _, c, _ = ref_ops.ref_eval_tile(feats[t:t + 1], labels[t:t + 1].numpy(), st.base_emb, None, st.cls,
                                            None, (TILE, TILE), K)
            cm += c
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    value = tiles_per_step * args.steps / total
    line = {
        'impl': 'reference', 'metric': 'tiles_per_sec', 'value': value, 'unit': '1024x1024 tiles/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'tiles_per_step': tiles_per_step, 'C': C, 'feature_hw': HW_LR,
                   'classes': K, 'mode': 'base'},
        'cpu_baseline': {'value': value, 'unit': '1024x1024 tiles/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{args.steps} timed single-tile steps of the oracle port (reference is pure '
                                   'Python/PyTorch; run on host cores, rank 0 only)'},
        'e2e': {'value': value, 'unit': '1024x1024 tiles/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- ours
class LaunchCounter:
    """Counts this library's kernel launches by wrapping the C-ABI call table (per entry point: how many kernels it
    launches at the benchmarked shapes, from the sources; checked against the ncu launch list under profiles/)."""
    PER_CALL = {'sl_pop_prepare': 3, 'sl_pop_fg_lowres': 1, 'sl_pop_bg_tc': 1, 'sl_pop_bg_simt': 1, 'sl_pop_head_tc': 2,
                'sl_upsample_argmax': 1, 'sl_confusion': 1, 'sl_views_reduce': 1, 'sl_window_accumulate': 1,
                'sl_map_proto': 4, 'sl_orth_loss': 1, 'sl_orth_from_sim': 1, 'sl_fuse_argmax': 1, 'sl_fuse_argmax_tiles': 1, 'sl_pseudo_label': 1,
                'sl_inter_union': 3}

    def __init__(self):
        from segland_b200 import _cabi
        self.n = 0
        self.by_name = {}
        self._cabi = _cabi
        self._orig = _cabi.call
        counter = self

        def counted(name, *a):
            rc = counter._orig(name, *a)
            k = counter.PER_CALL.get(name, 0)
            # sl_upsample_argmax: a second launch (confusion over label, pred) only when SL_POST_FUSED_CM=0 asks for
            # the counting outside the interpolation kernel (argument 13 = cm)
            if name == 'sl_upsample_argmax' and a[13] is not None and os.environ.get('SL_POST_FUSED_CM', '') == '0':
                k += 1
            counter.n += k
            counter.by_name[name] = counter.by_name.get(name, 0) + k
            return rc
        self._counted = counted

    def __enter__(self):
        from segland_b200 import ops
        self._cabi.call = self._counted
        ops.call = self._counted
        return self

    def __exit__(self, *exc):
        from segland_b200 import ops
        self._cabi.call = self._orig
        ops.call = self._orig


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from segland_b200 import ops, sweep, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    ops.check_device()
    peaks = load_peaks()
    T, P = args.tiles, args.passes
    sampler = ClockSampler(physical_gpu_index(local_rank))      # NVML init happens here, long before the timed region
    st = synth.make_trained_like_state(C, KB, 0, seed=1234)     # argmax follows the labels: mIoU is meaningful
    K = st.n_classes
    # distinct tiles per rank (weak scaling): different seeds per rank
    n_gen = min(T, 8)                                # generate 8 distinct tiles on the host, then decorrelate on device
    labels_h = synth.make_labels(n_gen, TILE, TILE, K, seed=1234 + rank)
    feats_h = synth.make_features(labels_h, st, STRIDE, seed=1234 + rank)
    reps = (T + n_gen - 1) // n_gen
    feats = feats_h.to(dev).repeat(reps, 1, 1, 1)[:T].contiguous()
    labels = labels_h.to(dev).repeat(reps, 1, 1)[:T].contiguous()
    if reps > 1:                                     # make the repeats distinct bytes (not that caches care)
        noise = torch.randn(T, 1, HW_LR, HW_LR, device=dev, generator=torch.Generator(dev).manual_seed(rank))
        feats = (feats.float() + 0.05 * noise).to(torch.bfloat16).contiguous()
    head = ops.PopHead(st.base_emb, st.cls, None, None, device=dev, bg_mode=args.bg_mode,
                       tc_precision=args.tc_precision)
    ev = sweep.TileEvaluator(head, (TILE, TILE))
    use_tc = head._use_tc(N_PIX)
    fused = use_tc and args.fuse and head.fused_ok(N_PIX)
    # default: the sweep is software-pipelined over two streams (post-processing of pass p-1 underneath the background
    # MLP of pass p, sweep.PipelinedTileEvaluator); --no-pipeline runs the three kernels back to back on one stream
    pipelined = not args.no_pipeline and not fused
    pev = sweep.PipelinedTileEvaluator(head, (TILE, TILE)) if pipelined else None

    stream = torch.cuda.current_stream()
    ev_k = {k: [] for k in ('fg', 'bg', 'post')}
    lg = torch.empty(T, K, HW_LR, HW_LR, dtype=torch.float32, device=dev)
    ev._logits = lg

    def one_pass(record, seq=False, sink=None):
        """One pass of the hot path over the T resident tiles; per-kernel events on the recorded passes."""
        sink = ev_k if sink is None else sink
        if pipelined and not seq:
            pev.trace = [] if record else None
            out = pev.step(feats, labels)
            if record:                       # fg(p), bg(p) and the post of pass p-1 that ran underneath bg(p)
                for name, a, b in pev.trace:
                    sink[name].append((a, b))
                pev.trace = None
            return out
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if record: marks[0].record(stream)
        if not fused:
            head(feats, out=lg, fg_only=True)
        if record: marks[1].record(stream)
        if fused:
            head.head_tc(feats, lg)          # fg logits + background MLP in one launch
        elif use_tc:
            head.bg_tc(feats, lg)
        else:
            head.bg_simt(feats, lg)
        if record: marks[2].record(stream)
        out = ops.upsample_argmax(lg, (TILE, TILE), label=labels, cm=ev.cm)
        if record:
            marks[3].record(stream)
            for name, a, b in (('fg', 0, 1), ('bg', 1, 2), ('post', 2, 3)):
                sink[name].append((marks[a], marks[b]))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # once per weight update, outside the sweep (PopHead caches the plan): timed on its own
    prepare_ms = 1e3 * time_loop(head.refresh, 10)
    for _ in range(max(args.warmup, 3)):
        for _ in range(min(P, 4)):
            one_pass(False)
    (pev or ev).finalize(base_classes=KB)                    # warm the (lazy) NCCL communicator as well
    (pev or ev).reset()
    sampler.start()
    barrier()
    counter = LaunchCounter()
    t_start, t_end = ev_pair()
    sampler.begin()
    with counter:
        t_start.record(stream)
        for _ in range(args.steps):
            for p in range(P):
                one_pass(p == P // 2)                        # per-kernel events on one pass per step
        cm, mious = (pev or ev).finalize(base_classes=KB)    # flushes the pipeline; the one all-reduce of the sweep
        t_end.record(stream)
    barrier()
    clocks = sampler.finish()
    elapsed_ms = max_over_ranks(t_start.elapsed_time(t_end), world, dev)
    kern_ms = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in ev_k.items()}
    value = world * T * P * args.steps / (elapsed_ms * 1e-3)
    gpu_launches = counter.n

    # ---- the same passes back to back on one stream (no overlap): per-kernel times of each kernel running alone
    seq_k = {k: [] for k in ('fg', 'bg', 'post')}
    ev.reset()
    n_seq = 4 * P
    for p in range(8):
        one_pass(False, seq=True)
    s0, s1 = ev_pair()
    s0.record(stream)
    for p in range(n_seq):
        one_pass(p % 8 == 4, seq=True, sink=seq_k)
    s1.record(stream)
    torch.cuda.synchronize()
    seq_pass_ms = max_over_ranks(s0.elapsed_time(s1), world, dev) / n_seq
    seq_ms = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in seq_k.items()}

    # ---- stage S (everything but the background MLP): fg + post, timed on its own
    ev.reset()

    def stage_s():
        head(feats, out=lg, fg_only=True)
        ops.upsample_argmax(lg, (TILE, TILE), label=labels, cm=ev.cm)
    stage_s_ms = 1e3 * time_loop(stage_s, max(args.steps, 20))

    # ---- the opt-in tensor modes on the same pass (informational; not parity-grade on near-tie data)
    alt_modes = {}
    if use_tc and not args.skip_configs:
        for mode in ('mid', 'balanced'):
            if mode == args.tc_precision:
                continue
            alt = ops.PopHead(st.base_emb, st.cls, None, None, device=dev, bg_mode=args.bg_mode, tc_precision=mode)

            def alt_pass():
                alt(feats, out=lg, fg_only=True)
                alt.bg_tc(feats, lg)
                ops.upsample_argmax(lg, (TILE, TILE), label=labels, cm=ev.cm)
            t_pass = max_over_ranks(time_loop(alt_pass, 4 * P, warm=P), world, dev)
            t_bg = time_loop(lambda: alt.bg_tc(feats, lg), 20)
            alt_modes[mode] = {'tiles_per_s': world * T / t_pass, 'ms_per_pass': 1e3 * t_pass, 'bg_ms': 1e3 * t_bg,
                               'mma_passes': {'mid': 4, 'balanced': 3}[mode]}
            del alt
        ev.reset()

    # ---- end to end through the public API with HOST buffers (double-buffered H2D, D2H of preds)
    Te = min(T, args.e2e_tiles)
    # --e2e-tiles 0 (profiler runs only) skips the host-buffer leg; the driver's default run always measures it
    e2e = run_e2e(ev, head, feats_h, labels_h, Te, args.steps * args.e2e_repeat, dev, world,
                  dist if world > 1 else None) if Te > 0 else None

    # ---- single-tile latency: the reference's operating point (scripts/evaluate_oem.sh:16-17)
    latency = run_latency_b1(ev, feats, labels, args.latency_tiles) if args.latency_tiles > 0 else None

    # ---- the other BASELINE configs, device-timed at this N
    del feats, labels, lg
    ev._logits = None
    torch.cuda.empty_cache()
    configs = {}
    if not args.skip_configs:
        configs['configs[2] 5-shot novel-class update'] = run_config2(rank, world, dev, peaks, args)
        configs['configs[3] ConvNeXt/Swin flip + sliding-window inference, probability maps'] = run_config3(rank, world, dev, peaks, args)
        configs['configs[4] fusemat fusion sweep + confusion mIoU'] = run_config4(rank, world, dev, peaks, args)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel
    traffic, traffic_file = load_traffic()
    bg_flops = (4.0 * C * C + 2.0 * C) * N_PIX * T           # algorithmic: two CxC layers + w3 dot, per launch
    passes = {'precise': 5, 'mid': 4, 'balanced': 3}[args.tc_precision] if use_tc else 2
    bg_exec_flops = (2.0 * C * C * passes + 2.0 * C) * N_PIX * T   # MMA FLOPs actually issued (split operands)
    fg_bytes = (C * N_PIX * 2 + KB * N_PIX * 4) * T          # features in + K fg logits out
    post_bytes = (K * N_PIX * 4 + 2 * TILE * TILE) * T       # low-res logits in + label in + pred out
    dominant = max(kern_ms, key=kern_ms.get)
    timed_s = elapsed_ms * 1e-3
    if dominant == 'bg':
        ach = bg_flops / (kern_ms['bg'] * 1e-3) / 1e12
        # a region of a second or more runs at the sustained (power-capped) clocks; a sub-second burst is judged against
        # the burst peak
        sustained = timed_s >= 1.0
        peak = peaks['tflops_sustained'] if sustained else peaks['tflops_burst']
        roofline = {'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                    'frac_of_burst_peak': ach / peaks['tflops_burst'], 'frac_of_sustained_peak': ach / peaks['tflops_sustained'],
                    'traffic': traffic.get('bg_pair_kernel', {}).get(args.tc_precision) if use_tc else None,
                    'traffic_note': f'dram read+write bytes per 32-tile launch from profiles/{traffic_file} (ncu --set full)',
                    'kernel': 'bg_pair_kernel (sl_pop_bg_tc, cta_group::2)' if use_tc else 'pop_bg_simt_kernel',
                    'mma_passes': passes,
                    'achieved_executed': bg_exec_flops / (kern_ms['bg'] * 1e-3) / 1e12,
                    'frac_executed_of_burst': bg_exec_flops / (kern_ms['bg'] * 1e-3) / 1e12 / peaks['tflops_burst'],
                    'peak_source': peaks['source'] + (' bf16 sustained (timed region %.2f s)' % timed_s if sustained
                                                      else ' bf16 burst (timed region %.2f s)' % timed_s),
                    'flops_per_launch': bg_flops,
                    'frac_kernel_alone': bg_flops / (seq_ms['bg'] * 1e-3) / 1e12 / peak,
                    'note': ('kernel duration measured inside the pipelined timed region, where the up-sampling kernel of the '
                             'previous pass shares the SMs (and the 1 kW power cap) with it; frac_kernel_alone is the same '
                             'kernel in the sequential leg') if pipelined else 'sequential passes'}
    else:
        byts = fg_bytes if dominant == 'fg' else post_bytes
        ach = byts / (kern_ms[dominant] * 1e-3) / 1e9
        roofline = {'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                    'frac': ach / peaks['hbm_gbs'], 'traffic': None, 'kernel': dominant,
                    'peak_source': peaks['source'], 'bytes_per_launch': byts}
    fg_gbs = fg_bytes / (seq_ms['fg'] * 1e-3) / 1e9
    s_bytes_tile = C * N_PIX * 2 + 2 * TILE * TILE
    stage_s = {'value': T / (stage_s_ms * 1e-3), 'unit': '1024x1024 tiles/s (fg logits + upsample/argmax/confusion; '
               'no background MLP)', 'ms_per_pass': stage_s_ms,
               'roofline_hbm_stage': {'achieved': s_bytes_tile * T / (stage_s_ms * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'],
                                      'unit': 'GB/s', 'frac': s_bytes_tile * T / (stage_s_ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
                                      'note': 'whole stage: algorithmic bytes per tile x tiles / stage time'},
               'roofline_hbm': {'kernel': 'pop_fg_mma_kernel (sl_pop_fg_lowres)', 'achieved': fg_gbs, 'peak': peaks['hbm_gbs'],
                                'unit': 'GB/s', 'frac': fg_gbs / peaks['hbm_gbs'], 'bytes_per_launch': fg_bytes,
                                'traffic': traffic.get('pop_fg_kernel', {}).get('base')},
               'post_roofline_hbm': {'kernel': 'sl_upsample_argmax (+ confusion)', 'achieved': post_bytes / (seq_ms['post'] * 1e-3) / 1e9,
                                     'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                                     'frac': post_bytes / (seq_ms['post'] * 1e-3) / 1e9 / peaks['hbm_gbs'],
                                     'ms_alone': seq_ms['post'],
                                     'bound_note': 'instruction-issue bound, not HBM (ncu: profiles/r2b_ncu_post_regs.txt)'},
               'algorithmic_bytes_per_tile': s_bytes_tile}
    if pipelined:
        # what the HBM-bound stage costs once it is pipelined: the pass minus the background MLP running alone
        marg_ms = elapsed_ms / (args.steps * P) - seq_ms['bg']
        stage_s['marginal_in_pipelined_pass'] = {
            'ms_per_pass': marg_ms, 'achieved': s_bytes_tile * T / (marg_ms * 1e-3) / 1e9, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
            'frac': s_bytes_tile * T / (marg_ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
            'note': 'pipelined pass time minus the background MLP alone (sequential leg): foreground kernel + what hiding the '
                    'post-processing underneath the MLP costs it'}
    cpu_tps, cores = cpu_reference_tiles_per_s(args.cpu_tiles) if world == 1 and args.cpu_tiles > 0 else (None, None)
    eager = gpu_eager_reference(args.eager_tiles, dev) if world == 1 and args.eager_tiles > 0 else None
    line = {
        'metric': 'tiles_per_sec', 'value': value, 'unit': '1024x1024 tiles/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': elapsed_ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16 features, f32 accumulate',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'tiles_per_step_per_gpu': T * P, 'passes_per_step': P, 'tiles_per_pass_per_gpu': T,
                   'C': C, 'feature_hw': HW_LR, 'classes': K, 'mode': 'base',
                   'bg_path': ('tcgen05 ' + {'precise': 'precise (split-bf16, 2+3 passes)', 'mid': 'mid (2+2 passes)',
                                             'balanced': 'balanced (split-bf16 L1, fp16 L2, 2+1 passes)'}[args.tc_precision])
                   if use_tc else 'fp32 CUDA cores',
                   'head_launches': 'one (fg logits fused into the tcgen05 kernel)' if fused else 'two (fg kernel + bg kernel)',
                   'schedule': 'pipelined (post of pass p-1 under bg of pass p)' if pipelined else 'sequential',
                   'l2_policy': f'inputs larger than L2 ({T * C * N_PIX * 2 / 1e6:.0f} MB features per pass)',
                   'parallelism': f'dp{world}'},
        'timed_region_s': timed_s,
        'kernel_ms_per_pass': kern_ms,
        'schedule': ('pipelined over two streams: fg(p), bg(p) on a high-priority stream, post(p-1) underneath bg(p) on a second '
                     'stream (sweep.PipelinedTileEvaluator); kernel_ms_per_pass are durations while co-running'
                     if pipelined else 'sequential: fg -> bg -> post on one stream'),
        'sequential': {'tiles_per_s': world * T / (seq_pass_ms * 1e-3), 'ms_per_pass': seq_pass_ms,
                       'kernel_ms_per_pass': seq_ms, 'passes': n_seq,
                       'note': 'the same passes back to back on one stream, each kernel alone on the GPU'},
        'prepare_ms_per_weight_update': prepare_ms,
        'opt_in_tensor_modes': alt_modes,
        'roofline': roofline,
        'stage_s': stage_s,
        'e2e': e2e,
        'latency_b1': latency,
        'gpu_launches': gpu_launches,
        'gpu_launches_by_entry_point': counter.by_name,
        'clocks': clocks,
        'miou_total': float(mious[2]),
        'configs': configs,
    }
    if cpu_tps is not None:
        line['cpu_baseline'] = {'value': cpu_tps, 'unit': '1024x1024 tiles/s', 'cores': cores, 'kind': 'port',
                                'sample': f'{args.cpu_tiles} tiles of the same workload through the oracle port '
                                          '(head + F.interpolate + np.argmax + get_confusion_matrix), median of 1'}
    if eager is not None:
        line['gpu_eager_baseline'] = eager
        if latency is not None:
            latency['eager_pytorch_gpu_us_p50'] = {k: eager[k]['latency_b1_us_p50'] for k in ('fp32', 'tf32')}
    print(json.dumps(line), flush=True)


def run_e2e(ev, head, feats_h, labels_h, Te, steps, dev, world, dist):
    """Same metric through the public API with HOST inputs: every step copies its batch of
    features+labels from pinned host memory and reads the predictions (and, at the end of the
    sweep, the confusion matrix) back.  Copies run on a second stream, double-buffered.  Also measures the plain
    pinned H2D copy rate of the same buffers in the same run (the ceiling this leg can reach)."""
    n_gen = feats_h.shape[0]
    reps = (Te + n_gen - 1) // n_gen
    # first-touch the pinned staging buffers on the GPU's NUMA node (8 ranks otherwise pull from one socket), then
    # give the process its CPUs back for the CPU-baseline leg
    from segland_b200 import sweep
    cpus_before = os.sched_getaffinity(0)
    numa_node = sweep.bind_to_gpu_numa_node(dev.index)
    f_pin = feats_h.repeat(reps, 1, 1, 1)[:Te].contiguous().pin_memory()
    l_pin = labels_h.repeat(reps, 1, 1)[:Te].contiguous().pin_memory()
    pred_pin = torch.empty(Te, TILE, TILE, dtype=torch.uint8).pin_memory()
    pred_pin.zero_()
    os.sched_setaffinity(0, cpus_before)
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    bufs = [(torch.empty_like(f_pin, device=dev), torch.empty_like(l_pin, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    ev.reset()
    ev._logits = None

    def upload(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            bufs[b][0].copy_(f_pin, non_blocking=True)
            bufs[b][1].copy_(l_pin, non_blocking=True)
            ready[b].record(copy_stream)

    def run(n):
        upload(0)
        for i in range(n):
            b = i % 2
            if i + 1 < n:
                upload(i + 1)
            main.wait_event(ready[b])
            out = ev.step(bufs[b][0], bufs[b][1])
            consumed[b].record(main)
            pred_pin.copy_(out['pred'], non_blocking=True)
        cm_host = ev.cm.cpu()                                    # D2H of the step results' reduction
        torch.cuda.synchronize()
        return cm_host

    for b in range(2):
        consumed[b].record(main)
    run(3)
    ev.reset()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = ev_pair()
    e0.record(main)
    run(steps)
    e1.record(main)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), wall * 1e3)                    # host-visible completion is what a caller sees
    ms = max_over_ranks(ms, world, dev)
    # the ceiling: the same pinned buffers copied host -> device back to back, nothing else running (all ranks at once)
    if dist is not None:
        dist.barrier()
    h2d_s = time_loop(lambda: (bufs[0][0].copy_(f_pin, non_blocking=True), bufs[0][1].copy_(l_pin, non_blocking=True)), 10)
    h2d_s = max_over_ranks(h2d_s, world, dev)
    h2d_bytes = int(f_pin.numel() * 2 + l_pin.numel())
    value = world * Te * steps / (ms * 1e-3)
    ceiling = world * Te / h2d_s
    return {'value': value, 'unit': '1024x1024 tiles/s',
            'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': int(pred_pin.numel()), 'tiles_per_step_per_gpu': Te,
            'steps': steps, 'timed_region_s': ms * 1e-3,
            'api': 'segland_b200.sweep.TileEvaluator.step(features_host->device, labels) + pred D2H',
            'h2d_ceiling': {'gb_per_s_per_gpu': h2d_bytes / h2d_s / 1e9, 'tiles_per_s': ceiling,
                            'note': 'plain pinned host->device copy of the same buffers, all ranks at once, same run'},
            'frac_of_h2d_ceiling': value / ceiling,
            'host_numa_node': numa_node}


def run_latency_b1(ev, feats, labels, n):
    """Per-tile latency at batch 1 (the reference evaluates one tile per forward): host-timed from the call to the
    synchronised result, features already on the device; eager C-ABI calls and the CUDA-graph replay of the same step,
    next to the device time of the step's kernels alone (graph replays back to back, CUDA events)."""
    from segland_b200 import sweep
    T = feats.shape[0]
    ev.reset()
    ev._logits = None

    def timed(fn):
        lat = []
        for i in range(-10, n):
            t = max(i, 0) % T
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(t)
            torch.cuda.synchronize()
            if i >= 0:
                lat.append(time.perf_counter() - t0)
        lat.sort()
        return {'p50': 1e6 * lat[len(lat) // 2], 'p90': 1e6 * lat[int(len(lat) * 0.9)], 'mean': 1e6 * sum(lat) / len(lat)}

    eager = timed(lambda t: ev.step(feats[t:t + 1], labels[t:t + 1]))
    graphed = sweep.GraphedTileStep(ev, (1,) + tuple(feats.shape[1:]))
    graph = timed(lambda t: graphed.run(feats[t:t + 1], labels[t:t + 1]))
    in_place = timed(lambda t: graphed.graph.replay())           # inputs already in the graph's buffers
    kernels_s = time_loop(graphed.graph.replay, 200, warm=10)
    ev.reset()
    ev._logits = None
    return {'unit': 'us per 1024x1024 tile, batch 1, host-timed (call -> synchronised result), features on the device',
            'tiles': n, 'eager_c_abi_us': eager, 'cuda_graph_us': graph, 'cuda_graph_inputs_in_place_us': in_place,
            'kernels_only_us': 1e6 * kernels_s,
            'ratio_graph_p50_to_kernels': graph['p50'] / (1e6 * kernels_s),
            'note': 'cuda_graph_us includes the device-to-device copy of the tile (16.8 MB) into the graph input buffer'}


# ------------------------------------------------------------------------ other BASELINE configs
def run_config2(rank, world, dev, peaks, args):
    """configs[2]: 5-shot novel-class update: masked-average-pooling prototypes of the 20 OEM support tiles (4 novel
    classes x 5 shots, dataset/list/oem/all_5shot_seed123.txt; networks/pspnet.py:7-15), sharded over the ranks, one
    all-reduce of the per-class sums and shot counts (ft_pop.py:276-277's collective pattern), then the orthogonal
    loss of the new prototypes against the base prototypes (loss/criterion.py:37-43, pspnet_pop.py:234-239)."""
    from segland_b200 import ops, sweep, synth
    n_support, Kn = 20, 4
    mine = sweep.shard_range(n_support, rank, world)
    n_local = len(mine)
    st = synth.make_head_state(C, KB, Kn, seed=7)
    base = st.base_emb.to(dev)
    # rotate through enough distinct copies of the local shard that every update streams from HBM (> L2)
    shard_bytes = max(1, n_local) * (C * N_PIX * 2 + TILE * TILE * 4)
    n_sets = max(2, min(8, -(-300 * 1024 * 1024 // shard_bytes)))
    g = torch.Generator(dev).manual_seed(100 + rank)
    sets = []
    for s in range(n_sets):
        f = torch.randn(n_local, C, HW_LR, HW_LR, device=dev, generator=g).to(torch.bfloat16)
        m = (torch.rand(n_local, 1, TILE // 32, TILE // 32, device=dev, generator=g) < 0.3).float()
        m = torch.nn.functional.interpolate(m, size=(TILE, TILE), mode='nearest').contiguous()
        sets.append((f, m))
    cls_of = torch.tensor([i // 5 for i in mine], dtype=torch.int64, device=dev)
    state = {'i': 0}

    def update():
        f, m = sets[state['i'] % n_sets]
        state['i'] += 1
        novel = sweep.novel_prototypes_from_support(f, m, cls_of, Kn)
        loss, sim = ops.orth_loss(novel, base)
        return loss

    def map_only():
        f, m = sets[state['i'] % n_sets]
        state['i'] += 1
        if n_local:
            ops.masked_average_pooling(f, m)

    iters = max(50, args.steps * 5)
    t_update = max_over_ranks(time_loop(update, iters, warm=5), world, dev)
    t_map = time_loop(map_only, iters, warm=3)
    map_bytes = n_local * (C * N_PIX * 2 + TILE * TILE * 4)
    gbs = map_bytes / t_map / 1e9 if n_local else 0.0
    del sets
    return {'workload': '20 support tiles [512,128,128] bf16 + fp32 masks [1,1024,1024], 4 novel classes x 5 shots, sharded '
                        f'by rank ({n_local} on rank 0); MAP -> per-class sums + counts all-reduce (NCCL) -> orthogonal loss',
            'updates_per_s': 1.0 / t_update, 'support_tiles_per_s': n_support / t_update, 'ms_per_update': 1e3 * t_update,
            'scaling': 'strong (20 tiles shared by all ranks; the update is latency-bound: 4 MAP launches, one all-reduce, '
                       'index_add, orth loss)',
            'roofline': {'bound': 'hbm', 'kernel': 'sl_map_proto (4 launches)', 'achieved': gbs, 'peak': peaks['hbm_gbs'],
                         'unit': 'GB/s', 'frac': gbs / peaks['hbm_gbs'], 'ms': 1e3 * t_map,
                         'bytes_per_launch': map_bytes, 'note': 'rank 0 shard, features + fp32 mask per tile'},
            'collective': 'one torch.distributed all_reduce (NCCL) of 4 x 513 fp32 per update' if world > 1 else 'none (1 rank)'}


def run_config3(rank, world, dev, peaks, args):
    """configs[3]: ConvNeXt-POP / Swin-POP ft-mode inference (12 classes) with test-time views and full fp32
    probability-map output: (a) whole 1024^2 tile, original + h-flip views averaged at feature resolution
    (sl_views_reduce); (b) sliding window -- 512-px crops at stride 384 (3 x 3 windows, the last pulled back to the
    border), original + h-flip per crop, stitched with sl_window_accumulate -- then the fused up-sample + softmax."""
    from segland_b200 import ops, sweep, synth
    out = {}
    Tt = args.c3_tiles
    n_it = max(10, args.steps)

    def pipelined_time(head, reduce, feats):
        """The same step through sweep.PipelinedTileEvaluator (up-sampling + soft-max of step t-1 on a second stream
        underneath the head of step t; the register-resident kernel shares an SM with the pair kernel)."""
        pev = sweep.PipelinedTileEvaluator(head, (TILE, TILE), reduce=reduce)
        t = time_loop(lambda: pev.step(feats, want_probs=True), n_it, finish=pev.flush)
        return max_over_ranks(t, world, dev)
    for name, Cm in (('ConvNeXt-T C=192', 192), ('Swin-T/S C=96', 96)):
        stc = synth.make_head_state(Cm, KB, 4, seed=2)
        head = ops.PopHead(stc.base_emb, stc.cls, stc.novel_emb, stc.cls_n, device=dev)
        K = head.n_classes
        hw = TILE // 4
        g = torch.Generator(dev).manual_seed(200 + rank)
        feats = torch.randn(2 * Tt, Cm, hw, hw, device=dev, generator=g).to(torch.bfloat16)     # [V*T]: view-major
        lg = torch.empty(2 * Tt, K, hw, hw, device=dev)
        marks = {}

        def whole(record=False):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
            if record: e[0].record()
            head(feats, out=lg)
            if record: e[1].record()
            mean = ops.aggregate_views(lg.view(2, Tt, K, hw, hw), [0, 1])
            if record: e[2].record()
            r = ops.upsample_argmax(mean, (TILE, TILE), want_probs=True)
            if record:
                e[3].record()
                marks['whole'] = e
            return r
        t_whole_seq = max_over_ranks(time_loop(whole, n_it), world, dev)
        # narrow heads (C <= 128) run on the weights-resident kernel, whose 14 warps x 128 registers fill the register
        # file: nothing can share its SMs, so those models keep the back-to-back schedule
        use_pipe = not args.no_pipeline and Cm > 128
        t_whole = t_whole_seq if not use_pipe else pipelined_time(
            head, lambda x: ops.aggregate_views(x.view(2, Tt, K, hw, hw), [0, 1]), feats)
        whole(True)
        torch.cuda.synchronize()
        e = marks['whole']
        k_whole = {'head_ms': e[0].elapsed_time(e[1]), 'views_reduce_ms': e[1].elapsed_time(e[2]),
                   'upsample_softmax_ms': e[2].elapsed_time(e[3])}
        prob_bytes = Tt * (K * hw * hw * 4 + K * TILE * TILE * 4 + TILE * TILE)
        prob_gbs = prob_bytes / (k_whole['upsample_softmax_ms'] * 1e-3) / 1e9
        del feats, lg
        # sliding window
        plan = ops.WindowPlan((TILE, TILE), (512, 512), (384, 384), 4)
        flips = (0, 1)
        E = plan.n_windows * len(flips)
        hc, wc = plan.crop_lr_hw
        cf = torch.randn(Tt * E, Cm, hc, wc, device=dev, generator=g).to(torch.bfloat16)        # [T,E] crops
        clg = torch.empty(Tt * E, K, hc, wc, device=dev)
        canvas = torch.empty(Tt, K, *plan.canvas_hw, device=dev)

        def sliding(record=False):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
            if record: e[0].record()
            head(cf, out=clg)
            if record: e[1].record()
            ops.window_accumulate(clg.view(Tt, E, K, hc, wc), plan, flips, out=canvas)
            if record: e[2].record()
            r = ops.upsample_argmax(canvas, (TILE, TILE), want_probs=True)
            if record:
                e[3].record()
                marks['sliding'] = e
            return r
        t_slide_seq = max_over_ranks(time_loop(sliding, n_it), world, dev)
        t_slide = t_slide_seq if not use_pipe else pipelined_time(
            head, lambda x: ops.window_accumulate(x.view(Tt, E, K, hc, wc), plan, flips), cf)
        sliding(True)
        torch.cuda.synchronize()
        e = marks['sliding']
        k_slide = {'head_ms': e[0].elapsed_time(e[1]), 'window_accumulate_ms': e[1].elapsed_time(e[2]),
                   'upsample_softmax_ms': e[2].elapsed_time(e[3])}
        win_bytes = (clg.numel() + canvas.numel()) * 4
        win_gbs = win_bytes / (k_slide['window_accumulate_ms'] * 1e-3) / 1e9
        bg_flops = (4.0 * Cm * Cm + 2.0 * Cm) * hw * hw * 2 * Tt
        out[name] = {
            'tiles_per_step_per_gpu': Tt,
            'schedule': ('pipelined (views / window reduction + up-sampling + soft-max of step t-1 underneath the head of step t)'
                         if use_pipe else 'sequential (the narrow-head kernel leaves no room on the SM)'),
            'whole_tile_2_views': {'tiles_per_s': world * Tt / t_whole, 'ms_per_step': 1e3 * t_whole,
                                   'sequential_tiles_per_s': world * Tt / t_whole_seq, 'kernels_sequential': k_whole, 'kernels': k_whole,
                                   'head_tflops_algorithmic': bg_flops / (k_whole['head_ms'] * 1e-3) / 1e12},
            'sliding_window_3x3_2_views': {'tiles_per_s': world * Tt / t_slide, 'ms_per_step': 1e3 * t_slide,
                                           'sequential_tiles_per_s': world * Tt / t_slide_seq, 'kernels': k_slide,
                                           'crops_per_tile': E,
                                           'window_accumulate_roofline': {'bound': 'hbm', 'achieved': win_gbs, 'peak': peaks['hbm_gbs'],
                                                                          'unit': 'GB/s', 'frac': win_gbs / peaks['hbm_gbs'],
                                                                          'bytes_per_launch': win_bytes}},
            'roofline': {'bound': 'hbm', 'kernel': 'sl_upsample_argmax with probability-map output', 'achieved': prob_gbs,
                         'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': prob_gbs / peaks['hbm_gbs'],
                         'bytes_per_launch': prob_bytes}}
        del cf, clg, canvas
        torch.cuda.empty_cache()
    out['scaling'] = 'weak'
    return out


def run_config4(rank, world, dev, peaks, args):
    """configs[4]: fusemat.py:37-48 over a full synthetic test set: the 80 OEM test tiles (dataset/list/oem/test.txt),
    M = 3 models' up-sampled logit stacks [12,1024,1024] fp32 per tile (12 GB in all, resident in HBM), summed in list
    order, divided by M, argmax, with the confusion matrix against the labels in the same launch; tiles are sharded
    by rank and the confusion matrix is all-reduced once per sweep."""
    from segland_b200 import ops, sweep
    n_tiles, M, K = args.c4_tiles, 3, 12
    mine = sweep.shard_range(n_tiles, rank, world)
    n_local = len(mine)
    g = torch.Generator(dev).manual_seed(300 + rank)
    mats = [torch.randn(n_local, K, TILE, TILE, device=dev, generator=g) for _ in range(M)]
    labels = torch.randint(0, K, (n_local, TILE, TILE), device=dev, dtype=torch.uint8, generator=g)
    cm = torch.zeros(K, K, dtype=torch.int64, device=dev)
    preds = []

    def sweep_once():
        if n_local:
            ops.fuse_logits_sweep(mats, labels=labels, cm=cm)
        sweep.all_reduce_sum_(cm)

    reps = max(3, args.steps // 4)
    t_sweep = max_over_ranks(time_loop(sweep_once, reps, warm=1), world, dev)
    byts = n_local * (M * K * 4 + 2) * TILE * TILE
    t_local = time_loop(lambda: ops.fuse_logits_sweep(mats, labels=labels, cm=cm), reps, warm=1) if n_local else 1.0
    gbs = byts / t_local / 1e9
    miou = ops.miou_from_confusion(cm, KB)[2]
    del mats, labels, preds
    torch.cuda.empty_cache()
    return {'workload': f'{n_tiles} tiles x M=3 x [12,1024,1024] fp32 logits + u8 labels, sharded by rank ({n_local} on rank 0), '
                        'one sl_fuse_argmax_tiles call per sweep (per tile: sum, /M, argmax, confusion), one int64 all-reduce per sweep',
            'tiles_per_s': n_tiles / t_sweep, 'ms_per_sweep': 1e3 * t_sweep, 'scaling': 'strong (80 tiles shared by all ranks)',
            'roofline': {'bound': 'hbm', 'kernel': 'fuse_argmax_kernel (sl_fuse_argmax)', 'achieved': gbs, 'peak': peaks['hbm_gbs'],
                         'unit': 'GB/s', 'frac': gbs / peaks['hbm_gbs'], 'bytes_per_sweep_rank0': byts},
            'miou_total_random_logits': float(miou)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--tiles', type=int, default=32, help='distinct resident 1024^2 tiles per pass per GPU')
    ap.add_argument('--no-pipeline', action='store_true', help='run fg -> bg -> post back to back on one stream')
    ap.add_argument('--passes', type=int, default=48, help='passes over the resident tiles per step (48 x 32 tiles: '
                                                           '~55 ms per step, so 20 steps time > 1 s)')
    ap.add_argument('--e2e-tiles', type=int, default=8)
    ap.add_argument('--e2e-repeat', type=int, default=10, help='host-buffer steps = steps x this (8 tiles each)')
    ap.add_argument('--latency-tiles', type=int, default=1000, help='single-tile latency sample (0 = skip)')
    ap.add_argument('--cpu-tiles', type=int, default=8, help='tiles in the bounded CPU-baseline sample (0 = skip)')
    ap.add_argument('--eager-tiles', type=int, default=16, help='tiles in the eager-PyTorch-on-GPU baseline (0 = skip)')
    ap.add_argument('--c3-tiles', type=int, default=16, help='tiles per step in configs[3]')
    ap.add_argument('--c4-tiles', type=int, default=80, help='tiles in the configs[4] fusion sweep (all ranks together)')
    ap.add_argument('--skip-configs', action='store_true', help='only the headline workload (profiler runs)')
    ap.add_argument('--bg-mode', default='auto', choices=['auto', 'tc', 'simt'])
    ap.add_argument('--fuse', action='store_true', help='single-launch head (sl_pop_head_tc); slower on B200, see DESIGN.md')
    ap.add_argument('--tc-precision', default='precise', choices=['precise', 'mid', 'balanced'],
                    help="tensor-core background MLP mode; 'precise' is the parity-grade default")
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
