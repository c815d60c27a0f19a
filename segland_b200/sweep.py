"""Data-parallel tile sweeps: tiles are independent, so the path shards with
no data-path collective.  Each rank owns a contiguous range of tile indices (no
DistributedSampler padding, which would double-count tiles), accumulates its confusion matrix
on the device for the whole sweep, and takes part in ONE integer all-reduce at the end
(1,152 bytes for OEM's 12 classes).  The reference never reduces its eval confusion matrix
across ranks (eval_base.py:132-133,201), so distributed eval is new functionality here.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


def shard_range(n_items, rank, world_size):
    """Contiguous, non-overlapping, exhaustive partition of range(n_items); sizes differ by <= 1."""
    if not (0 <= rank < world_size):
        raise ValueError('rank out of range')
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def bind_to_gpu_numa_node(device_index=None):
    """Pin the calling process (and so its first-touch pinned host buffers) to the CPUs of the NUMA node the GPU hangs
    off, so that eight ranks streaming features host->device do not all pull from one socket's memory.  Best effort:
    returns the node index, or None when sysfs does not tell (containers, single-node hosts)."""
    import os
    try:
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        props = torch.cuda.get_device_properties(idx)
        bus = f'{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        node = int(open(f'/sys/bus/pci/devices/{bus}/numa_node').read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f'/sys/devices/system/node/node{node}/cpulist').read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:                                                  # noqa: BLE001  (no sysfs / no permission)
        return None


def all_reduce_sum_(t, group=None):
    """In-place SUM all-reduce when a process group is up; identity otherwise.  Integer tensors
    stay integer, so confusion matrices reduce bit-exactly in any rank order."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class TileEvaluator:
    """eval_base.py:162-199 / eval_ft.py:162-202 after the decoder, for a stream of tile batches:
    head -> align-corners upsample -> argmax -> confusion accumulation, all on the device."""

    def __init__(self, head: ops.PopHead, out_size, ignore_label=ops.IGNORE_LABEL, reduce=None):
        """reduce: optional callable applied to the head's low-res logits before the up-sampling, on the stream the
        post-processing runs on -- test-time view averaging (`lambda lg: ops.aggregate_views(lg.view(V, -1, *lg.shape[1:]),
        flips)`) or sliding-window stitching (`lambda lg: ops.window_accumulate(lg.view(B, E, ...), plan, flips)`); the
        features of a step then hold all views / crops of its tiles and labels stay per tile."""
        self.head = head
        self.out_size = (int(out_size[0]), int(out_size[1]))
        self.ignore_label = ignore_label
        self.reduce = reduce
        K = head.n_classes
        self.cm = torch.zeros(K, K, dtype=torch.int64, device=head.device)
        self._logits = None

    def reset(self):
        self.cm.zero_()

    def step(self, features, labels=None, want_pred=True, **kw):
        """features [B,C,h,w] bf16 (device, or pinned host -> async H2D), labels [B,H,W] uint8.
        Returns the upsample_argmax dict (device tensors); no host synchronisation."""
        feats = ops._cuda(features, torch.bfloat16)
        B, _, h, w = feats.shape
        if self._logits is None or self._logits.shape[0] != B or self._logits.shape[-2:] != (h, w):
            self._logits = torch.empty(B, self.head.n_classes, h, w, dtype=torch.float32, device=feats.device)
        logits = self.head(feats, out=self._logits)
        if self.reduce is not None:
            logits = self.reduce(logits)
        return ops.upsample_argmax(logits, self.out_size, label=labels, cm=self.cm if labels is not None else None,
                                   ignore_label=self.ignore_label, want_pred=want_pred, **kw)

    def finalize(self, base_classes, group=None):
        """One all-reduce for the whole sweep, then the mIoU split of eval_ft.py:196-202.
        Returns (cm int64 [K,K] on the device, (base, novel, total, per-class) float64)."""
        all_reduce_sum_(self.cm, group)
        return self.cm, ops.miou_from_confusion(self.cm, base_classes)


class PipelinedTileEvaluator(TileEvaluator):
    """TileEvaluator with the sweep software-pipelined over two streams: the head of batch t (foreground kernel, then
    the tensor-bound background MLP) runs on a high-priority stream while the up-sample / arg-max / confusion kernel of
    batch t-1 runs on a second stream UNDER the background MLP.  The two kernels are complementary -- the MLP keeps the
    tensor pipe busy and leaves most issue slots idle, the post-processing kernel is bound by instruction issue and
    touches neither tensor cores nor shared memory -- and they are sized to share an SM (pair kernel: 12 warp slots x
    128 registers, 225 KB; post-processing: 128-thread CTAs of 128 registers, 256 B), so a pass costs
    fg + max(bg, post) instead of fg + bg + post.  Results are identical to TileEvaluator (same kernels, integer
    accumulation); low-res logits are double-buffered.  (The foreground kernel stays in front of the MLP: it is
    HBM-bound and cheap, and moving it underneath as well was measured to gain nothing under the power cap,
    profiles/r2b_overlap_probe.txt.)

    `step` returns the result dict of the PREVIOUS batch (None on the first call); `flush()` returns the last one.
    Returned tensors are safe to use on the caller's current stream.  The map-sized outputs (pred, conf, probs) live in
    one of two buffer sets owned by the evaluator -- a fresh map per batch on the second stream would have to be handed
    back to the caller's stream with record_stream, and with the host running ahead of the GPU the caching allocator then
    cannot recycle it and falls back to one cudaMalloc per step (2 ms) -- so they stay valid until the next-but-one
    `step`.  `finalize` flushes."""

    def __init__(self, head: ops.PopHead, out_size, ignore_label=ops.IGNORE_LABEL, reduce=None):
        super().__init__(head, out_size, ignore_label, reduce)
        dev = head.device
        self._hi = torch.cuda.Stream(device=dev, priority=-1)    # head kernels: placed first when both have CTAs pending
        self._lo = torch.cuda.Stream(device=dev, priority=0)
        self._lg = [None, None]
        self._pred = [None, None]
        self._ev_fg = [torch.cuda.Event() for _ in range(2)]
        self._ev_bg = [torch.cuda.Event() for _ in range(2)]
        self._ev_post = [torch.cuda.Event() for _ in range(2)]
        self._n = 0
        self._pending = None                                     # (buffer, labels, want_pred, kw) of the batch awaiting its post
        self.last = None
        self.trace = None                                        # set to a list to collect (name, start, end) CUDA events

    def _retire(self, t):
        """A buffer replaced because the batch shape changed (the ragged last batch of a sweep) may still be read or
        written by kernels queued on the two streams: tell the caching allocator before the reference is dropped."""
        if t is not None:
            t.record_stream(self._hi)
            t.record_stream(self._lo)
            t.record_stream(torch.cuda.current_stream(t.device))  # the caller may still be reading a returned map

    def _mark(self, stream):
        if self.trace is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        return e

    def _post(self, gate, cur):
        if self._pending is None:
            return None
        b, labels, want_pred, kw = self._pending
        self._pending = None
        lo = self._lo
        lo.wait_event(self._ev_bg[b])
        if gate is not None:
            lo.wait_event(gate)                                  # start together with the NEXT batch's background MLP
        with torch.cuda.stream(lo):
            t0 = self._mark(lo)
            lg = self._lg[b] if self.reduce is None else self.reduce(self._lg[b])
            # every map-sized output lives in one of two evaluator-owned buffers (see the class docstring)
            Bt, Kt = lg.shape[0], lg.shape[1]
            wants = {'pred': (want_pred, (Bt, *self.out_size), torch.uint8),
                     'conf': (kw.get('want_conf', False), (Bt, *self.out_size), torch.float32),
                     'probs': (kw.get('want_probs', False), (Bt, Kt, *self.out_size), torch.float32)}
            bufs = self._pred[b] if self._pred[b] is not None else {}
            self._pred[b] = bufs
            for name, (want, shape, dtype) in wants.items():
                if want and (name not in bufs or tuple(bufs[name].shape) != shape):
                    self._retire(bufs.get(name))
                    bufs[name] = torch.empty(shape, dtype=dtype, device=lg.device)
            out = ops.upsample_argmax(lg, self.out_size, label=labels, cm=self.cm if labels is not None else None,
                                      ignore_label=self.ignore_label, want_pred=want_pred,
                                      out_bufs={k: v for k, v in bufs.items() if wants[k][0]}, **kw)
            t1 = self._mark(lo)
            self._ev_post[b].record(lo)
        if t0 is not None:
            self.trace.append(('post', t0, t1))
        cur.wait_event(self._ev_post[b])
        for k, v in out.items():
            if k not in ('pred', 'conf', 'probs'):               # the up-sampled logits are allocated per call
                v.record_stream(cur)
        if labels is not None:
            labels.record_stream(lo)
        self.last = out
        return out

    def step(self, features, labels=None, want_pred=True, **kw):
        dev = self.head.device
        cur = torch.cuda.current_stream(dev)
        feats = ops._cuda(features, torch.bfloat16)
        if labels is not None:
            labels = ops._cuda(labels, torch.uint8)
        B, _, h, w = feats.shape
        b = self._n & 1
        lg = self._lg[b]
        if lg is None or lg.shape[0] != B or lg.shape[-2:] != (h, w):
            self._retire(lg)
            lg = self._lg[b] = torch.empty(B, self.head.n_classes, h, w, dtype=torch.float32, device=feats.device)
        hi = self._hi
        split = (h * w) % 8 == 0 and not self.head.fuse           # padded / single-launch heads have no split point
        hi.wait_stream(cur)                                      # inputs (and any H2D copy) are ready
        if self._n >= 2:
            hi.wait_event(self._ev_post[b])                      # logits buffer b has been consumed
        with torch.cuda.stream(hi):
            t0 = self._mark(hi)
            if split:
                self.head(feats, out=lg, fg_only=True)
            t1 = self._mark(hi)
            self._ev_fg[b].record(hi)
            if split:
                self.head.bg(feats, lg)
            else:
                self.head(feats, out=lg)
            t2 = self._mark(hi)
            self._ev_bg[b].record(hi)
        if t0 is not None:
            self.trace += [('fg', t0, t1), ('bg', t1, t2)]
        feats.record_stream(hi)
        # second stream: the post-processing of the previous batch, gated on this batch's foreground kernel so that it
        # starts together with this batch's background MLP
        prev = self._post(self._ev_fg[b], cur)
        self._pending = (b, labels, want_pred, kw)
        self._n += 1
        return prev

    def flush(self):
        """Run the post-processing of the last batch; returns its result dict (None when nothing is pending)."""
        cur = torch.cuda.current_stream(self.head.device)
        out = self._post(None, cur)
        cur.wait_stream(self._hi)
        return out

    def reset(self):
        self.flush()
        super().reset()

    def finalize(self, base_classes, group=None):
        self.flush()
        return super().finalize(base_classes, group)


def prototype_mean_all_reduce_(per_image_sum, count, group=None):
    """Multi-GPU masked-average-pool prototypes: MAP is a mean of PER-IMAGE ratios
    (networks/pspnet.py:14-15), so every support image lives wholly on one rank; ranks sum
    their per-image prototypes [C] and image counts, all-reduce both, then divide."""
    all_reduce_sum_(per_image_sum, group)
    all_reduce_sum_(count, group)
    return per_image_sum / count.to(per_image_sum.dtype)


def novel_prototypes_from_support(features, masks, class_of_tile, n_novel, group=None):
    """Novel-class prototypes from a (sharded) support set -- north_star's "masked-average-pooling prototype
    generation and novel-prototype update from the 5-shot support set" (spec: this repo; the reference defines
    masked_average_pooling, networks/pspnet.py:7-15, but learns novel_emb by SGD).
    features [n,C,h,w] bf16 and masks [n,1,H,W] float are THIS rank's support tiles, class_of_tile [n] (int64, values
    in [0,n_novel)) their novel-class indices.  Every image lives wholly on one rank (MAP is a mean of per-image
    ratios); ranks exchange per-class sums of per-image prototypes and shot counts in ONE all-reduce.
    Returns novel_emb [n_novel,C] fp32 (rows of classes without a shot anywhere are zero)."""
    dev = features.device
    C = features.shape[1]
    acc = torch.zeros(n_novel, C + 1, dtype=torch.float32, device=dev)      # [per-class prototype sums | shot counts]
    if features.shape[0] > 0:
        _, per_image = ops.masked_average_pooling(features, masks, return_per_image=True)
        idx = class_of_tile.to(dev, torch.int64)
        acc.index_add_(0, idx, torch.cat([per_image, torch.ones_like(per_image[:, :1])], dim=1))
    all_reduce_sum_(acc, group)                                             # one collective: 4 x 513 fp32 = 8 KB
    return acc[:, :C] / acc[:, C:].clamp_min(1.0)


class GraphedTileStep:
    """TileEvaluator.step captured once as a CUDA graph for a fixed batch shape (the reference's operating point is
    one tile per forward, scripts/evaluate_oem.sh:16-17, eval_ft.py:162-167): a replay costs one graph launch instead
    of four kernel launches plus their host-side set-up, so the per-tile latency is the kernels' own time.
    `run(features, labels)` copies the batch into the graph's static input buffers (device-to-device, or H2D from
    pinned memory) and replays; outputs are the graph's static tensors (valid until the next run)."""

    def __init__(self, evaluator: TileEvaluator, feat_shape, want_pred=True, with_labels=True, **kw):
        self.ev = evaluator
        dev = evaluator.head.device
        B = int(feat_shape[0])
        self.feats = torch.zeros(tuple(feat_shape), dtype=torch.bfloat16, device=dev)
        self.labels = torch.full((B, *evaluator.out_size), ops.IGNORE_LABEL, dtype=torch.uint8, device=dev) \
            if with_labels else None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            cm_before = evaluator.cm.clone()
            for _ in range(2):                                   # warm-up: allocations, lazy module loads, plan caches
                evaluator.step(self.feats, self.labels, want_pred=want_pred, **kw)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.out = evaluator.step(self.feats, self.labels, want_pred=want_pred, **kw)
            evaluator.cm.copy_(cm_before)                        # the warm-up calls counted ignore-only labels: nothing, but be exact
        torch.cuda.current_stream(dev).wait_stream(side)

    def run(self, features, labels=None):
        self.feats.copy_(features, non_blocking=True)
        if labels is not None and self.labels is not None:
            self.labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        return self.out


class LogitBank:
    """Device-resident replacement for the .mat round trip between evaluation and fusion:
    eval_base.py:168,190-191 up-samples every tile's logits to full resolution and dumps them with
    scipy.io.savemat (50 MB/tile/model); fusemat.py:37-48 loads them back, sums them in directory order,
    divides by len(fusion_list) and takes the argmax.  Here each model's sweep deposits its LOW-RES logits
    (<= 3 MB/tile, 80 OEM test tiles x 12 x 256^2 fp32 = 252 MB per model -- nothing next to 180 GB of HBM) and
    `fuse` averages the models at low resolution (sl_views_reduce) and runs the fused up-sample + argmax
    (+ confusion) kernel once.  Bilinear interpolation is linear, so upsample(mean_m L_m) equals
    mean_m upsample(L_m) up to fp32 rounding (~1e-7 relative): argmax maps agree with fusemat.py except on
    exact near-ties (tests: >= 99.99 %, every disagreement within 1e-4 of the top-2 gap).  All models must share
    one low-res geometry; full-resolution stacks of mixed geometry go through ops.fuse_logits instead."""

    def __init__(self, n_tiles, n_classes, lowres_hw, device=None):
        self.n_tiles, self.K = int(n_tiles), int(n_classes)
        self.hw = (int(lowres_hw[0]), int(lowres_hw[1]))
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.models = []                       # one [n_tiles,K,h,w] fp32 tensor per deposited model, in order

    def new_model(self):
        """Start depositing the next model's sweep; returns its index."""
        self.models.append(torch.zeros(self.n_tiles, self.K, *self.hw, dtype=torch.float32, device=self.device))
        return len(self.models) - 1

    def deposit(self, model_index, tile_index, logits_lr):
        """logits_lr [B,K,h,w] fp32 (device) for tiles tile_index .. tile_index+B-1 of that model."""
        B = logits_lr.shape[0]
        if tuple(logits_lr.shape[1:]) != (self.K, *self.hw):
            raise ValueError(f'expected [B,{self.K},{self.hw[0]},{self.hw[1]}], got {tuple(logits_lr.shape)}')
        self.models[model_index][tile_index:tile_index + B].copy_(logits_lr, non_blocking=True)

    def fuse(self, out_size, n_lists=None, labels=None, cm=None, tiles_per_call=16, ignore_label=ops.IGNORE_LABEL):
        """fusemat.py:42-48 for every tile: uint8 [n_tiles,H,W] on the device; with labels [n_tiles,H,W] uint8 and an
        int64 [K,K] cm the confusion matrix is accumulated in the same kernel."""
        M = len(self.models)
        if M == 0:
            raise ValueError('no model deposited')
        n = M if n_lists is None else int(n_lists)
        H, W = int(out_size[0]), int(out_size[1])
        pred = torch.empty(self.n_tiles, H, W, dtype=torch.uint8, device=self.device)
        for t0 in range(0, self.n_tiles, tiles_per_call):
            t1 = min(self.n_tiles, t0 + tiles_per_call)
            stack = torch.stack([m[t0:t1] for m in self.models], 0)               # [M,B,K,h,w]
            mean = ops.aggregate_views(stack, [0] * M, scale=1.0 / n)
            out = ops.upsample_argmax(mean, (H, W), label=None if labels is None else labels[t0:t1], cm=cm,
                                      ignore_label=ignore_label)
            pred[t0:t1] = out['pred']
        return pred


# --------------------------------------------------------------------------- GeoTIFF label maps (no rasterio needed)
GEO_TAGS = (33550, 33922, 34264, 34735, 34736, 34737, 42113)   # ModelPixelScale, ModelTiepoint, ModelTransformation,
#                                                                 GeoKeyDirectory, GeoDoubleParams, GeoAsciiParams, GDAL_NODATA
_TIFF_TYPES = {1: ('B', 1), 2: ('c', 1), 3: ('H', 2), 4: ('I', 4), 5: ('II', 8), 12: ('d', 8), 16: ('Q', 8)}


def read_geotiff_tags(path):
    """The georeferencing tags of a (classic, non-Big) TIFF as {tag: (tiff_type, values)}: what eval_base.py:180-188
    takes from `rasterio.open(input).profile` (CRS + affine transform + nodata), read straight from the image file
    directory so a tile's georeferencing can be copied onto its label map without rasterio."""
    import struct
    with open(path, 'rb') as f:
        data = f.read()
    bo = {b'II': '<', b'MM': '>'}[data[:2]]
    if struct.unpack(bo + 'H', data[2:4])[0] != 42:
        raise ValueError('not a classic TIFF')
    off = struct.unpack(bo + 'I', data[4:8])[0]
    n = struct.unpack(bo + 'H', data[off:off + 2])[0]
    out = {}
    for i in range(n):
        tag, typ, cnt, val = struct.unpack(bo + 'HHI4s', data[off + 2 + 12 * i: off + 14 + 12 * i])
        if tag not in GEO_TAGS or typ not in _TIFF_TYPES:
            continue
        fmt, size = _TIFF_TYPES[typ]
        raw = val[:size * cnt] if size * cnt <= 4 else data[struct.unpack(bo + 'I', val)[0]:][:size * cnt]
        if typ == 2:
            out[tag] = (typ, raw.decode('ascii', 'replace'))
        else:
            out[tag] = (typ, list(struct.unpack(bo + fmt[0] * (cnt * len(fmt)), raw)))
    return out


def write_geotiff(path, pred_u8, palette=None, geo_tags=None, nodata=0):
    """A single-band uint8 GeoTIFF of a label map, as eval_base.py:180-188 / eval_ft.py:184-194 write with rasterio
    (`driver GTiff, dtype uint8, count 1, nodata 0`, `write_colormap(1, colormap)`): baseline little-endian TIFF, one
    uncompressed strip, palette colour when `palette` (flat [r,g,b]*256 list or {index: (r,g,b[,a])} as the scripts'
    `colormap`) is given, the georeferencing tags of `read_geotiff_tags` copied through, GDAL_NODATA = nodata."""
    import struct
    import numpy as np
    a = np.ascontiguousarray(pred_u8, dtype=np.uint8)
    if a.ndim != 2:
        raise ValueError('write_geotiff takes one [H,W] uint8 map')
    H, W = a.shape
    entries = []                                     # (tag, type, count, packed bytes)
    add = lambda tag, typ, vals: entries.append((tag, typ, len(vals), struct.pack('<' + _TIFF_TYPES[typ][0][0] * len(vals), *vals)))
    add(256, 4, [W]); add(257, 4, [H]); add(258, 3, [8]); add(259, 3, [1])
    add(262, 3, [3 if palette is not None else 1])
    add(277, 3, [1]); add(278, 4, [H]); add(279, 4, [H * W]); add(284, 3, [1]); add(339, 3, [1])
    if palette is not None:
        lut = np.zeros((3, 256), dtype=np.uint16)
        if isinstance(palette, dict):
            for k, rgb in palette.items():
                lut[:, int(k)] = [int(c) * 257 for c in rgb[:3]]
        else:
            flat = list(palette) + [0] * (768 - len(palette))
            lut[:] = (np.asarray(flat[:768], dtype=np.uint16).reshape(256, 3).T) * 257
        add(320, 3, lut.reshape(-1).tolist())
    tags = dict(geo_tags or {})
    if nodata is not None and 42113 not in tags:
        tags[42113] = (2, str(nodata) + '\0')
    for tag, (typ, vals) in tags.items():
        if typ == 2:
            b = vals.encode('ascii') if isinstance(vals, str) else bytes(vals)
            entries.append((tag, 2, len(b), b))
        else:
            add(tag, typ, vals)
    entries.sort(key=lambda e: e[0])
    n = len(entries) + 1                              # + StripOffsets
    ifd_off = 8
    extra_off = ifd_off + 2 + 12 * n + 4
    extra = b''
    packed = []
    for tag, typ, cnt, b in entries:
        if len(b) <= 4:
            packed.append((tag, typ, cnt, b.ljust(4, b'\0')))
        else:
            if len(extra) % 2:
                extra += b'\0'
            packed.append((tag, typ, cnt, struct.pack('<I', extra_off + len(extra))))
            extra += b
    if len(extra) % 2:
        extra += b'\0'
    strip_off = extra_off + len(extra)
    packed.append((273, 4, 1, struct.pack('<I', strip_off)))
    packed.sort(key=lambda e: e[0])
    with open(path, 'wb') as f:
        f.write(b'II' + struct.pack('<HI', 42, ifd_off) + struct.pack('<H', n))
        for tag, typ, cnt, v in packed:
            f.write(struct.pack('<HHI', tag, typ, cnt) + v)
        f.write(struct.pack('<I', 0) + extra)
        f.write(a.tobytes())


class AsyncMapWriter:
    """Asynchronous uint8 label-map writer (eval_base.py:180-188 writes one GeoTIFF per tile and
    fusemat.py:49-53 one palettised PNG per tile, both synchronously inside the eval loop).  `submit` copies the
    device map into a pinned staging slot on a side stream and returns at once; worker threads wait for the
    copy's event, encode a 'P'-mode PNG with the class palette (as fusemat.py does) or, with fmt='tif', the palettised
    single-band GeoTIFF of eval_base.py:180-188 (`write_geotiff`; pass the input tile's `read_geotiff_tags(...)` to
    `submit` to carry its georeferencing), and recycle the slot."""

    def __init__(self, out_dir, shape, palette=None, slots=8, workers=2, fmt='png'):
        import os
        import queue
        import threading
        os.makedirs(out_dir, exist_ok=True)
        if fmt not in ('png', 'tif'):
            raise ValueError("fmt must be 'png' or 'tif'")
        self.out_dir, self.palette, self.fmt = out_dir, palette, fmt
        self._free, self._work = queue.Queue(), queue.Queue()
        self._stream = torch.cuda.Stream()
        for _ in range(slots):
            self._free.put(torch.empty(tuple(shape), dtype=torch.uint8).pin_memory())
        self._errors = []
        self._threads = [threading.Thread(target=self._run, daemon=True) for _ in range(workers)]
        for t in self._threads:
            t.start()

    def _run(self):
        import os
        from PIL import Image
        while True:
            item = self._work.get()
            if item is None:
                return
            name, buf, event, geo = item
            try:
                event.synchronize()
                if self.fmt == 'tif':
                    write_geotiff(os.path.join(self.out_dir, name + '.tif'), buf.numpy(), self.palette, geo)
                else:
                    img = Image.fromarray(buf.numpy(), 'P')
                    if self.palette is not None:
                        img.putpalette(self.palette)
                    img.save(os.path.join(self.out_dir, name + '.png'))
            except Exception as e:                                  # noqa: BLE001  (reported by close())
                self._errors.append((name, e))
            finally:
                self._free.put(buf)

    def submit(self, name, pred_u8, geo_tags=None):
        """pred_u8: uint8 [H,W] device tensor.  Blocks only when every staging slot is in flight."""
        buf = self._free.get()
        self._stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._stream):
            buf.copy_(pred_u8, non_blocking=True)
            event = torch.cuda.Event()
            event.record(self._stream)
        pred_u8.record_stream(self._stream)
        self._work.put((name, buf, event, geo_tags))

    def close(self):
        for _ in self._threads:
            self._work.put(None)
        for t in self._threads:
            t.join()
        if self._errors:
            raise RuntimeError(f'AsyncMapWriter: {len(self._errors)} writes failed, first: {self._errors[0]}')
