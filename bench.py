#!/usr/bin/env python
"""Benchmark of the POP head + dense post-processing hot path (BASELINE.json metric:
1024^2 tiles/sec, head+postproc, device-timed; roofline fractions).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE configs[1]): PSPNet-POP base-class eval on synthetic OEM-shaped 1024x1024
tiles: features [T,512,128,128] bf16, 8 classes (bg + 7 base), labels [T,1024,1024] u8.
One "step" = one pass of the whole path over a batch of T distinct tiles per GPU:
    sl_pop_prepare -> sl_pop_fg_lowres -> sl_pop_bg_(tc|simt) -> sl_upsample_argmax(+confusion)
T*16.8 MB of features is larger than the 126 MB L2, so every step streams from HBM.
Weak scaling: every rank owns its own T tiles; the only collective is one int64 all-reduce of
the confusion matrix at the end of the sweep (inside the timed region).

The oracle (oracle/) is executed only for the cpu_baseline leg and for --impl reference.
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
    p = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    return json.load(open(p)) if os.path.exists(p) else {}


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
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
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:                                   # noqa: BLE001
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'note': 'no NVML samples'}
        return {'sm_mhz': statistics.median(self.samples), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def physical_gpu_index(local):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        try:
            return int(vis.split(',')[local])
        except (ValueError, IndexError):
            return local
    return local


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
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from segland_b200 import ops, sweep, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    ops.check_device()
    peaks = load_peaks()
    T = args.tiles
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

    stream = torch.cuda.current_stream()
    ev_k = {k: [] for k in ('prepare', 'fg', 'bg', 'post')}

    def step(record):
        """One pass of the hot path over the T resident tiles, with per-kernel events."""
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if record else None
        if record: marks[0].record(stream)
        head.refresh()
        if record: marks[1].record(stream)
        lg = ev._logits
        if lg is None or lg.shape[0] != T:
            ev._logits = lg = torch.empty(T, K, HW_LR, HW_LR, dtype=torch.float32, device=dev)
        if not fused:
            head(feats, out=lg, fg_only=True)
        if record: marks[2].record(stream)
        if fused:
            head.head_tc(feats, lg)          # fg logits + background MLP in one launch
        elif use_tc:
            head.bg_tc(feats, lg)
        else:
            head.bg_simt(feats, lg)
        if record: marks[3].record(stream)
        out = ops.upsample_argmax(lg, (TILE, TILE), label=labels, cm=ev.cm)
        if record:
            marks[4].record(stream)
            for name, a, b in (('prepare', 0, 1), ('fg', 1, 2), ('bg', 2, 3), ('post', 3, 4)):
                ev_k[name].append((marks[a], marks[b]))
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(False)
    ev.finalize(base_classes=KB)                             # warm the (lazy) NCCL communicator as well
    ev.reset()
    barrier()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for _ in range(args.steps):
        step(True)
    cm, mious = ev.finalize(base_classes=KB)                 # the one all-reduce of the sweep
    t_end.record(stream)
    barrier()
    clocks = sampler.finish()
    elapsed_ms = t_start.elapsed_time(t_end)
    if world > 1:
        tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    kern_ms = {k: statistics.mean(a.elapsed_time(b) for a, b in v) for k, v in ev_k.items()}
    value = world * T * args.steps / (elapsed_ms * 1e-3)

    # ---- stage S (everything but the background MLP): fg + post, timed on its own
    ev.reset()
    for _ in range(3):
        head(feats, out=ev._logits, fg_only=True)
        ops.upsample_argmax(ev._logits, (TILE, TILE), label=labels, cm=ev.cm)
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    for _ in range(args.steps):
        head(feats, out=ev._logits, fg_only=True)
        ops.upsample_argmax(ev._logits, (TILE, TILE), label=labels, cm=ev.cm)
    s1.record(stream)
    torch.cuda.synchronize()
    stage_s_ms = s0.elapsed_time(s1) / args.steps

    # ---- end to end through the public API with HOST buffers (double-buffered H2D, D2H of preds)
    Te = min(T, args.e2e_tiles)
    # --e2e-tiles 0 (profiler runs only) skips the host-buffer leg; the driver's default run always measures it
    e2e = run_e2e(ev, head, feats_h, labels_h, Te, args.steps, dev, world, dist if world > 1 else None) if Te > 0 else None

    if rank != 0:
        return
    # ---- roofline of the dominant kernel
    traffic = load_traffic()
    bg_flops = (4.0 * C * C + 2.0 * C) * N_PIX * T           # algorithmic: two CxC layers + w3 dot, per launch
    passes = (5 if args.tc_precision == 'precise' else 3) if use_tc else 2
    bg_exec_flops = (2.0 * C * C * passes + 2.0 * C) * N_PIX * T   # MMA FLOPs actually issued (split operands)
    fg_bytes = (C * N_PIX * 2 + KB * N_PIX * 4) * T          # features in + K fg logits out
    post_bytes = (K * N_PIX * 4 + 2 * TILE * TILE) * T       # low-res logits in + label in + pred out
    dominant = max(kern_ms, key=kern_ms.get)
    if dominant == 'bg':
        ach = bg_flops / (kern_ms['bg'] * 1e-3) / 1e12
        roofline = {'bound': 'tensor', 'achieved': ach, 'peak': peaks['tflops_sustained'], 'unit': 'TFLOP/s',
                    'frac': ach / peaks['tflops_sustained'],
                    'traffic': traffic.get('bg_pair_kernel', {}).get(args.tc_precision) if use_tc else None,
                    'traffic_note': 'dram read+write bytes per 32-tile launch from profiles/r1_traffic.json (ncu --set full)',
                    'kernel': 'bg_pair_kernel (sl_pop_bg_tc, cta_group::2)' if use_tc else 'pop_bg_simt_kernel',
                    'mma_passes': passes,
                    'achieved_executed': bg_exec_flops / (kern_ms['bg'] * 1e-3) / 1e12,
                    'frac_executed_of_burst': bg_exec_flops / (kern_ms['bg'] * 1e-3) / 1e12 / peaks['tflops_burst'],
                    'peak_source': peaks['source'] + ' bf16 sustained (kernel timed inside a long step)',
                    'flops_per_launch': bg_flops}
    else:
        byts = fg_bytes if dominant == 'fg' else post_bytes
        ach = byts / (kern_ms[dominant] * 1e-3) / 1e9
        roofline = {'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                    'frac': ach / peaks['hbm_gbs'], 'traffic': None, 'kernel': dominant,
                    'peak_source': peaks['source'], 'bytes_per_launch': byts}
    fg_gbs = fg_bytes / (kern_ms['fg'] * 1e-3) / 1e9
    stage_s = {'value': T / (stage_s_ms * 1e-3), 'unit': '1024x1024 tiles/s (fg logits + upsample/argmax/confusion; '
               'no background MLP)', 'ms_per_step': stage_s_ms,
               'roofline_hbm': {'kernel': 'pop_fg_kernel (sl_pop_fg_lowres)', 'achieved': fg_gbs, 'peak': peaks['hbm_gbs'],
                                'unit': 'GB/s', 'frac': fg_gbs / peaks['hbm_gbs'], 'bytes_per_launch': fg_bytes,
                                'traffic': traffic.get('pop_fg_kernel', {}).get('base')},
               'algorithmic_bytes_per_tile': C * N_PIX * 2 + 2 * TILE * TILE}
    cpu_tps, cores = cpu_reference_tiles_per_s(args.cpu_tiles) if world == 1 and args.cpu_tiles > 0 else (None, None)
    line = {
        'metric': 'tiles_per_sec', 'value': value, 'unit': '1024x1024 tiles/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': elapsed_ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16 features, f32 accumulate',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'tiles_per_step_per_gpu': T, 'C': C, 'feature_hw': HW_LR, 'classes': K,
                   'mode': 'base', 'bg_path': (f'tcgen05 {args.tc_precision} ' + ('(split-bf16, 2+3 passes)' if args.tc_precision == 'precise'
                                                                    else '(split-bf16 L1, fp16 L2, 2+1 passes)'))
                   if use_tc else 'fp32 CUDA cores',
                   'head_launches': 'one (fg logits fused into the tcgen05 kernel)' if fused else 'two (fg kernel + bg kernel)',
                   'l2_policy': f'inputs larger than L2 ({T * C * N_PIX * 2 / 1e6:.0f} MB features per step)',
                   'parallelism': f'dp{world}'},
        'kernel_ms_per_step': kern_ms,
        'roofline': roofline,
        'stage_s': stage_s,
        'e2e': e2e,
        # per step: 3 (prepare: {normalise + layer 1 + fold}, layer 2, layer 3; 5 with SL_PREP_SPLIT=1) + head (2 launches
        # fused incl. the prototype transpose, else fg + bg) + upsample/argmax + confusion over (label, pred)
        'gpu_launches': args.steps * ((5 if os.environ.get('SL_PREP_SPLIT', '')[:1] == '1' else 3) + 4),
        'clocks': clocks,
        'miou_total': float(mious[2]),
    }
    if cpu_tps is not None:
        line['cpu_baseline'] = {'value': cpu_tps, 'unit': '1024x1024 tiles/s', 'cores': cores, 'kind': 'port',
                                'sample': f'{args.cpu_tiles} tiles of the same workload through the oracle port '
                                          '(head + F.interpolate + np.argmax + get_confusion_matrix), median of 1'}
    print(json.dumps(line), flush=True)


def run_e2e(ev, head, feats_h, labels_h, Te, steps, dev, world, dist):
    """Same metric through the public API with HOST inputs: every step copies its batch of
    features+labels from pinned host memory and reads the predictions (and, at the end of the
    sweep, the confusion matrix) back.  Copies run on a second stream, double-buffered."""
    from segland_b200 import ops
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    run(steps)
    e1.record(main)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = max(e0.elapsed_time(e1), wall * 1e3)                    # host-visible completion is what a caller sees
    if dist is not None:
        tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    return {'value': world * Te * steps / (ms * 1e-3), 'unit': '1024x1024 tiles/s',
            'h2d_bytes_per_step': int(f_pin.numel() * 2 + l_pin.numel()),
            'd2h_bytes_per_step': int(pred_pin.numel()), 'tiles_per_step_per_gpu': Te,
            'api': 'segland_b200.sweep.TileEvaluator.step(features_host->device, labels) + pred D2H',
            'host_numa_node': numa_node}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--tiles', type=int, default=32, help='distinct 1024^2 tiles per step per GPU')
    ap.add_argument('--e2e-tiles', type=int, default=8)
    ap.add_argument('--cpu-tiles', type=int, default=8, help='tiles in the bounded CPU-baseline sample (0 = skip)')
    ap.add_argument('--bg-mode', default='auto', choices=['auto', 'tc', 'simt'])
    ap.add_argument('--fuse', action='store_true', help='single-launch head (sl_pop_head_tc); slower on B200, see DESIGN.md')
    ap.add_argument('--tc-precision', default='precise', choices=['precise', 'balanced'],
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
