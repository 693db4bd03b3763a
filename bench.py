#!/usr/bin/env python
"""bench.py — images/sec of RetinaNet decode + top-k + NMS (BASELINE.json metric), every BASELINE config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c1|c2|c3|c4|c5|c2s]
                  [--logits dense|sparse|clustered] [--scaling weak|strong] [--configs all|none]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
              bench.py --gpus N --steps K --warmup W

Headline workload (default, BASELINE.json configs[1] = "c2"): 640x640 COCO shapes (N = 76 725 anchors, C = 80), batch
64 per GPU, PerClassHardNMS (iou 0.5, score 0.05), pre_nms_top_k 5000 per class over the fused anchor axis (= the
reference's FilterTopKDetections), max_detections 100; dense N(0,1) logits and N(0, 0.5^2) box deltas
generated on the device.  One "step" = one pass of the fused path (rpp_detect) over the batch; the timed region
replays CUDA graphs of that step, alternating between two distinct input sets.  Images are sharded by rank with no
collective (weak scaling).

One JSON line on stdout (rank 0):
  value / ms_per_step   device-resident inputs (headline workload)
  e2e                   pinned HOST buffers in, host detections out, through rpp_detect_host; with the box's measured
                        concurrent H2D ceiling next to it
  roofline              the streaming collect kernel against measured HBM bandwidth
  configs               (N = 1) every other BASELINE config on this GPU: images/s, ms_per_step, path_frac (SURVEY §8d
                        bytes / step time / peak), per-stage times + dominant stage, and bit_exact_images = images
                        whose detections equal the CPU oracle's over ALL images of a 16-image sub-batch, for dense,
                        sparse and clustered logits
  strong_scaling        (N > 1) configs[2] (GlobalSoftNMS, B = 64 TOTAL) and configs[3] (1024^2 CombinedNMS, B = 32
                        TOTAL) sharded over the N GPUs, with the same GPU's full-batch time beside it
  worst_case            the headline step with every sampled list bypassed (rpp_debug_force_exact_scan) and on
                        clustered logits
  cpu_baseline / --impl reference   the CPU oracle (oracle/, the restated TF path) on this box's host cores.
"""
import argparse
import copy
import ctypes
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'retinanet-tensorflow2.x_b200'))

M = 100
BASE_CONFIG = {
    'input': {'input_shape': [640, 640], 'channels': 3},
    'architecture': {'feature_fusion': {'min_level': 3, 'max_level': 7},
                     'head': {'num_classes': 80, 'num_anchors': 9}},
    'anchor_params': {'areas': [1024.0, 4096.0, 16384.0, 65536.0, 262144.0],
                      'aspect_ratios': [0.5, 1.0, 2.0],
                      'scales': [1, 1.2599210498948732, 1.5874010519681994]},
    'encoder_params': {'box_variance': [0.1, 0.1, 0.2, 0.2], 'scale_box_targets': False},
    'inference': {'batch_size': 64, 'mode': 'PerClassHardNMS', 'iou_threshold': 0.5, 'score_threshold': 0.05,
                  'soft_nms_sigma': 0.5, 'pre_nms_top_k': 5000, 'filter_per_class': True, 'max_detections': M},
}

# BASELINE.json `configs`, in order.  `batch` is per GPU for weak workloads and TOTAL for strong ones.
WORKLOADS = {
    'c1': dict(cfg='configs[0]', H=640, C=80, batch=1, scaling='weak', mode='CombinedNMS', per_class=True,
               text='640x640 80-class, batch 1, CombinedNMS, per-class filter k=5000 (latency row)'),
    'c2': dict(cfg='configs[1]', H=640, C=80, batch=64, scaling='weak', mode='PerClassHardNMS', per_class=True,
               text='640x640 80-class synthetic logits, per-class hard NMS (iou 0.5, score 0.05), pre_nms_top_k '
                    '5000/class over the fused anchor axis, 100 dets'),
    'c3': dict(cfg='configs[2]', H=640, C=80, batch=64, scaling='strong', mode='GlobalSoftNMS', per_class=False,
               text='640x640 80-class, GlobalSoftNMS (Gaussian sigma 0.5), global filter k=5000, batch 64 TOTAL'),
    'c4': dict(cfg='configs[3]', H=1024, C=80, batch=32, scaling='strong', mode='CombinedNMS', per_class=True,
               text='1024x1024 (196 416 anchors) 80-class, CombinedNMS, per-class filter k=5000, batch 32 TOTAL'),
    'c5': dict(cfg='configs[4]', H=320, C=5, batch=512, scaling='weak', mode='GlobalHardNMS', per_class=False,
               text='320x320 5-class, GlobalHardNMS (reference-exact: no suppression), global filter k=5000, '
                    'batch 512'),
    'c2s': dict(cfg='configs[1] variant', H=640, C=80, batch=64, scaling='weak', mode='PerClassSoftNMS',
                per_class=True, text='configs[1] with PerClassSoftNMS (sigma 0.5)'),
}


def workload_params(wl):
    from retinanet.cfg.config import AttrDict
    cfg = copy.deepcopy(BASE_CONFIG)
    cfg['input']['input_shape'] = [wl['H'], wl['H']]
    cfg['architecture']['head']['num_classes'] = wl['C']
    cfg['inference'].update(mode=wl['mode'], filter_per_class=wl['per_class'], batch_size=wl['batch'])
    return AttrDict(cfg)


def num_anchors(H):
    return sum(int(math.ceil(H / 2 ** l)) ** 2 * 9 for l in range(3, 8))


def path_bytes_per_image(wl):
    n = num_anchors(wl['H'])
    return 4 * n * wl['C'] + 16 * n + 24 * M + 4


def metric_name(wl):
    return 'images/sec decode+NMS @{0}x{0} {1}-cls ({2}, top-k 5000, 100 dets)'.format(wl['H'], wl['C'], wl['mode'])


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
            'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
            'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
            'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'samples': len(s),
                'reasons': sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU oracle legs (the checker and the reported baseline; never the thing measured as `value`)
# ---------------------------------------------------------------------------------------------------------------
def oracle_detect(params, logits, deltas, threads):
    from oracle import ref
    inf = params.inference
    H, W = params.input.input_shape
    ap = params.anchor_params
    anchors, _ = ref.anchors(H, W, 3, 7, ap['areas'], ap['aspect_ratios'], ap['scales'])
    return ref.detect(logits, deltas, anchors, H, W, inf['mode'], iou_threshold=inf['iou_threshold'],
                      score_threshold=inf['score_threshold'], soft_nms_sigma=inf['soft_nms_sigma'],
                      pre_nms_top_k=inf['pre_nms_top_k'], filter_per_class=inf['filter_per_class'],
                      max_detections=inf['max_detections'], threads=threads)


def cpu_oracle_rate(wl, batch, dist, threads, min_seconds, max_passes):
    """images/s of the CPU oracle on `batch` images per pass; one warm-up pass, then passes until min_seconds."""
    import numpy as np
    from oracle import ref
    ref.build()
    params = workload_params(wl)
    n = num_anchors(wl['H'])
    rng = np.random.default_rng(42)
    logits = rng.standard_normal((batch, n, wl['C'])).astype(np.float32)
    if dist == 'sparse':
        logits = logits * 1.5 - 4.595
    deltas = np.clip(rng.standard_normal((batch, n, 4)) * 0.5, -4, 4).astype(np.float32)
    oracle_detect(params, logits, deltas, threads)          # warm-up (page faults, thread pool)
    passes, t0 = 0, time.perf_counter()
    while passes < max_passes:
        oracle_detect(params, logits, deltas, threads)
        passes += 1
        if time.perf_counter() - t0 >= min_seconds:
            break
    dt = time.perf_counter() - t0
    return batch * passes / dt, dt / passes * 1e3, passes


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import ref
    ref.build()
    wl = WORKLOADS[args.workload]
    threads = ref.hardware_threads()
    sample = 8 if wl['H'] <= 640 else 4
    dist = args.logits if args.logits != 'clustered' else 'dense'
    # the driver's K / W bound the run: each step is one pass over `sample` images
    import numpy as np
    params = workload_params(wl)
    n = num_anchors(wl['H'])
    rng = np.random.default_rng(42)
    logits = rng.standard_normal((sample, n, wl['C'])).astype(np.float32)
    if dist == 'sparse':
        logits = logits * 1.5 - 4.595
    deltas = np.clip(rng.standard_normal((sample, n, 4)) * 0.5, -4, 4).astype(np.float32)
    for _ in range(args.warmup):
        oracle_detect(params, logits, deltas, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_detect(params, logits, deltas, threads)
    dt = time.perf_counter() - t0
    rate, ms = sample * args.steps / dt, dt / args.steps * 1e3
    line = {
        'impl': 'reference', 'metric': metric_name(wl), 'value': rate, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': wl['scaling'], 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(wl, args.batch or wl['batch'], dist, args.gpus, wl['scaling']),
        'cpu_baseline': {'value': rate, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                         'sample': '{} images per step of the same workload (oracle/retinapost_ref.cpp: the '
                                   'reference path with its TF kernels restated; TensorFlow is not installable '
                                   'here), {} host threads'.format(sample, threads)},
        'e2e': {'value': rate, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(wl, batch_per_gpu, dist, n_gpus, scaling):
    n = num_anchors(wl['H'])
    in_bytes = batch_per_gpu * (4 * n * wl['C'] + 16 * n)
    return {
        'workload': '{}: {}'.format(wl['cfg'], wl['text']),
        'batch_per_gpu': batch_per_gpu, 'global_batch': batch_per_gpu * n_gpus, 'anchors': n, 'classes': wl['C'],
        'logits': dist, 'sharding': 'by image, no collective', 'scaling': scaling,
        'l2': ('inputs ({:.2f} GB per step) exceed the 126 MB L2; no flush needed'.format(in_bytes / 1e9)
               if in_bytes > 190e6 else 'inputs fit in L2: a 256 MB buffer is rewritten between timed steps'),
    }


# ---------------------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------------------
def note(rank, msg):
    """Progress marker on stderr (a stuck run then shows where it stopped)."""
    sys.stderr.write('[bench rank {} +{:.1f}s] {}\n'.format(rank, time.perf_counter() - _T0, msg))
    sys.stderr.flush()


_T0 = time.perf_counter()


class Bench:
    def __init__(self, args, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        from retinanet import _native
        self.torch, self.dist, self.native = torch, dist, _native
        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        local_rank = local_rank % torch.cuda.device_count()   # (debug runs: several ranks on one GPU over gloo)
        torch.cuda.set_device(local_rank)
        self.dev = torch.device('cuda', local_rank)
        self.L = _native.lib()
        self.peak, self.peak_src = peaks()
        self._flush = None
        self._anchors = {}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def anchors(self, params):
        H = params.input.input_shape[0]
        if H not in self._anchors:
            from retinanet.dataloader.anchor_generator import AnchorBoxGenerator
            self._anchors[H] = AnchorBoxGenerator(H, H, 3, 7, params.anchor_params).boxes
        return self._anchors[H]

    def inputs(self, wl, params, B, dist, seed_offset=0):
        from tools import synth_inputs
        lg, dl = synth_inputs.make_inputs(dist, B, self.anchors(params), wl['C'], wl['H'], wl['H'], self.dev,
                                          seed_logits=42 + seed_offset, seed_deltas=1234 + seed_offset)
        return {'class_logits': lg, 'encoded_boxes': dl}

    def flush_l2(self):
        if self._flush is None:
            self._flush = self.torch.empty(256 << 20, dtype=self.torch.uint8, device=self.dev)
        self._flush.zero_()

    def time_steps(self, fns, steps, flush):
        """ms per step of `steps` calls cycling through fns, CUDA events on the launch stream.  flush: rewrite a
        256 MB buffer between steps (inputs that fit in L2) — then every step is timed on its own."""
        torch = self.torch
        if not flush:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for i in range(steps):
                fns[i % len(fns)]()
            e.record()
            torch.cuda.synchronize()
            return s.elapsed_time(e) / steps
        tot = 0.0
        for i in range(steps):
            self.flush_l2()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fns[i % len(fns)]()
            e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        return tot / steps

    def stages(self, layer, h, x, steps):
        self.native.check(self.L.rpp_debug_stage_timing(h.ptr, 1))
        for _ in range(steps):
            layer(x)
        self.torch.cuda.synchronize()
        rep, calls = self.native.stage_report(h.ptr)
        self.native.check(self.L.rpp_debug_stage_timing(h.ptr, 0))
        return rep, calls

    def parity(self, wl, params, layer, dist, n_images):
        """Images (of n_images, run in calls of min(batch, n_images)) whose detections equal the oracle's."""
        import numpy as np
        from oracle import ref
        bc = min(wl['batch'], n_images)
        ok = tot = 0
        threads = ref.hardware_threads()
        for i in range(0, n_images, bc):
            x = self.inputs(wl, params, bc, dist, seed_offset=1000 + i)
            out = layer(x)
            got = {k: v.cpu().numpy() for k, v in out.items()}
            exp = oracle_detect(params, x['class_logits'].cpu().numpy(), x['encoded_boxes'].cpu().numpy(), threads)
            for b in range(bc):
                same = (got['valid_detections'][b] == exp['valid_detections'][b]
                        and got['classes'].dtype == exp['classes'].dtype
                        and np.array_equal(got['classes'][b], exp['classes'][b])
                        and np.array_equal(got['scores'][b], exp['scores'][b])
                        and np.allclose(got['boxes'][b], exp['boxes'][b], rtol=1e-5, atol=1e-6))
                ok += int(same)
                tot += 1
            del x, out
        return '{}/{}'.format(ok, tot)

    def config_row(self, key, B, steps, check, dists=('dense', 'sparse', 'clustered'), timed_dists=('dense',)):
        """One BASELINE config on this GPU: throughput (graph replay), stage split, parity counts."""
        from retinanet.model.layers import FusedPostProcessing
        torch = self.torch
        wl = WORKLOADS[key]
        params = workload_params(wl)
        layer = FusedPostProcessing(params)
        h = layer.handle(wl['C'])
        row = {'workload': '{}: {}'.format(wl['cfg'], wl['text']), 'batch': B, 'anchors': int(h.num_anchors)}
        bpi = path_bytes_per_image(wl)
        flush = B * bpi < 190e6
        for dist in timed_dists:
            x = self.inputs(wl, params, B, dist, seed_offset=self.rank)
            # inputs that are not flushed out of L2 between steps alternate between two distinct sets (as the headline
            # does): a 200-400 MB input replayed alone would find a third of itself in the 126 MB L2
            xs = [x] if flush else [x, self.inputs(wl, params, B, dist, seed_offset=self.rank + 100)]
            replays = []
            for xi in xs:
                for _ in range(3):
                    layer(xi)
                replays.append(layer.capture(xi)[0])
            for rp in replays:
                rp()
            torch.cuda.synchronize()
            ms = self.time_steps(replays, steps, flush)
            rep, _ = self.stages(layer, h, x, min(steps, 10))
            dom = max(rep.items(), key=lambda kv: kv[1]) if rep else (None, None)
            r = {'ms_per_step': ms, 'images_per_s': B / ms * 1e3,
                 'path_frac': (B * bpi / (ms * 1e-3) / 1e9) / self.peak,
                 'stage_ms': rep, 'dominant_stage': dom[0], 'dominant_stage_ms': dom[1],
                 'gpu_launches_per_step': int(self.L.rpp_last_launch_count())}
            if dist == 'dense':
                row.update(r)
            else:
                row[dist] = r
            del x, xs, replays
            torch.cuda.empty_cache()
        row['l2_flush_between_steps'] = bool(flush)
        row['path_bytes_per_image'] = bpi
        if check > 0:
            row['bit_exact_images'] = {d: self.parity(wl, params, layer, d, check) for d in dists}
        del layer
        torch.cuda.empty_cache()
        return row


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: libretinapost has no CPU path')
    bn = Bench(args, rank, local_rank, world)
    dev, L, _native = bn.dev, bn.L, bn.native
    from retinanet.model.layers import FusedPostProcessing
    # bind this rank to the CPUs / memory node next to its GPU: the end-to-end leg streams 1.65 GB per step out of
    # pinned host memory, and with 8 ranks remote-socket buffers would share one inter-socket link
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    local_rank = local_rank % torch.cuda.device_count()
    if world > 1:
        backend = os.environ.get('BENCH_DIST_BACKEND', 'nccl')   # gloo: debug runs with several ranks on one GPU
        if backend == 'nccl':
            dist.init_process_group('nccl', device_id=dev)
        else:
            dist.init_process_group(backend)

    key = args.workload
    wl = WORKLOADS[key]
    scaling = args.scaling or wl['scaling']
    total = args.batch or wl['batch']
    if scaling == 'strong':
        from retinanet.distributed import shard_range
        lo, hi = shard_range(total, rank, world)
        B = max(1, hi - lo)
        global_batch = total if total >= world else world
    else:
        B = total
        global_batch = B * world
    C, H = wl['C'], wl['H']
    params = workload_params(wl)
    layer = FusedPostProcessing(params)
    h = layer.handle(C)
    N = int(h.num_anchors)
    bpi = path_bytes_per_image(wl)
    logit_bytes = 4 * N * C
    flush = B * bpi < 190e6
    steps, warmup = args.steps, max(args.warmup, 3)

    note(rank, 'headline {}: inputs'.format(key))
    # two distinct input sets, replayed alternately (so that the number is not the property of one tensor)
    xs = [bn.inputs(wl, params, B, args.logits, seed_offset=rank + 100 * i) for i in range(1 if flush else 2)]
    for _ in range(warmup):
        for x in xs:
            out = layer(x)
    launches_per_step = int(L.rpp_last_launch_count())
    bn.barrier()
    note(rank, 'eager timing')
    ms_eager = bn.time_steps([lambda x=x: layer(x) for x in xs], steps, flush)
    note(rank, 'graph capture + timed region')
    replays = []
    for x in xs:
        rp, og = layer.capture(x)
        replays.append(rp)
    for rp in replays:
        rp()
    bn.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_step_local = bn.time_steps(replays, steps, flush)
    bn.barrier()
    clocks = sampler.stop() if sampler else None
    out = layer(xs[-1])
    assert bool((og['scores'] == out['scores']).all().item())
    valid_mean = float(out['valid_detections'].float().mean().item())
    ms_step = bn.max_over_ranks(ms_step_local)
    value = global_batch / (ms_step * 1e-3)

    note(rank, 'stage timing')
    # per-stage times: a second K-step pass with CUDA events at the stage boundaries on the launch stream
    rep, ncalls = bn.stages(layer, h, xs[0], steps)
    collect_ms = sum(v for k, v in rep.items() if k.startswith('collect') or k.startswith('emit:collect'))
    dom = max(rep.items(), key=lambda kv: kv[1]) if rep else (None, 0.0)
    logits, deltas = xs[0]['class_logits'], xs[0]['encoded_boxes']

    extras = {}
    if key == 'c2' and not args.quick:
        # the same step fed with the model-side input of the path, the per-level NHWC head outputs (views of the
        # layout a detector emits): consumed in place (rpp_detect_levels) vs the reference's FuseDetections concat
        note(rank, 'head levels / half precision / worst case')
        from retinanet.model.builder import ModelBuilder
        bounds = [0, 57600, 72000, 75600, 76500, 76725]
        heads = {'class-predictions': {}, 'box-predictions': {}}
        for li, level in enumerate(range(3, 8)):
            f = -(-H // 2 ** level)
            heads['class-predictions'][str(level)] = logits[:, bounds[li]:bounds[li + 1]].contiguous().view(B, f, f, 9 * C)
            heads['box-predictions'][str(level)] = deltas[:, bounds[li]:bounds[li + 1]].contiguous().view(B, f, f, 36)
        levels_ms = {}
        for name, fused in (('in_place', True), ('concat_then_detect', False)):
            model = ModelBuilder(params).add_post_processing_stage(None)
            model.layers[0].lazy = fused
            for _ in range(3):
                o2 = model(heads)
            torch.cuda.synchronize()
            levels_ms[name] = bn.time_steps([lambda: model(heads)], steps, False)
            o2 = model(heads)
            assert bool((o2['scores'] == layer(xs[0])['scores']).all().item())
            del model
        del heads
        torch.cuda.empty_cache()
        extras['from_head_levels'] = {
            'ms_per_step': levels_ms, 'images_per_s': {k: B / v * 1e3 for k, v in levels_ms.items()},
            'note': 'per-level NHWC head outputs as input (rank 0): rpp_detect_levels in place vs FuseDetections '
                    'concat + rpp_detect'}
        # 16-bit head outputs (what a mixed-precision detector can emit; the reference casts them to fp32 first):
        # the dominant stream halves.  Informational: the headline `value` stays on fp32 inputs.
        half_ms = {}
        for name, tdt in (('bf16', torch.bfloat16), ('f16', torch.float16)):
            xh = {'class_logits': logits.to(tdt), 'encoded_boxes': deltas.to(tdt)}
            for _ in range(3):
                layer(xh)
            torch.cuda.synchronize()
            half_ms[name] = bn.time_steps([lambda: layer(xh)], steps, False)
            del xh
        torch.cuda.empty_cache()
        extras['half_precision_inputs'] = {
            'ms_per_step': half_ms, 'images_per_s': {k: B / v * 1e3 for k, v in half_ms.items()},
            'note': 'same workload with f16 / bf16 logits and deltas read in place (rpp_detect_typed), eager calls'}
        # the perf cliff, bounded: every sampled list bypassed (each problem scans its column exactly), and a
        # trained-detector-like clustered input
        _native.check(L.rpp_debug_force_exact_scan(h.ptr, 1))
        layer(xs[0])
        torch.cuda.synchronize()
        ms_scan = bn.time_steps([lambda: layer(xs[0])], 3, False)
        _native.check(L.rpp_debug_force_exact_scan(h.ptr, 0))
        xc = bn.inputs(wl, params, B, 'clustered', seed_offset=rank)
        for _ in range(3):
            layer(xc)
        torch.cuda.synchronize()
        _native.exact_scans(h.ptr, reset=True)
        layer(xs[0])
        scans_dense = _native.exact_scans(h.ptr, reset=True)
        layer(xc)
        scans_clu = _native.exact_scans(h.ptr, reset=True)
        ms_clu = bn.time_steps([lambda: layer(xc)], steps, False)
        del xc
        torch.cuda.empty_cache()
        extras['worst_case'] = {
            'forced_exact_scan_ms_per_step': ms_scan, 'clustered_ms_per_step': ms_clu,
            'clustered_images_per_s': B / ms_clu * 1e3,
            'exact_scans_per_step': {'dense': scans_dense, 'clustered': scans_clu, 'problems': B * C,
                                     'note': 'rpp_debug_exact_scans: problems whose sampled list ran dry'},
            'note': 'same step, eager calls: rpp_debug_force_exact_scan(1) = no sampled list is used, every '
                    '(image, class) problem selects from its whole column; clustered = tools/synth_inputs.py'}

        # optional per-level pre-selection (north_star "per-level top-k"; the reference filters over the fused axis):
        # FilterTopKDetectionsPerLevel (rpp_topk_levels), top-1000 per class and level on 16 images of the same
        # geometry, scores resident; parity of 2 images against the oracle's composition of the reference filter
        try:
            from retinanet.model.layers import FilterTopKDetectionsPerLevel
            from oracle import ref
            import numpy as np
            Bl = min(B, 16)
            bounds = [0, 57600, 72000, 75600, 76500, 76725]
            sc = torch.sigmoid(logits[:Bl])
            bx = deltas[:Bl].contiguous()
            lv_s = [sc[:, a:b_].contiguous() for a, b_ in zip(bounds[:-1], bounds[1:])]
            lv_b = [bx[:, a:b_].contiguous() for a, b_ in zip(bounds[:-1], bounds[1:])]
            del sc
            flt = FilterTopKDetectionsPerLevel(1000, True)
            xin = {'scores': lv_s, 'boxes': lv_b}
            for _ in range(2):
                o3 = flt(xin)
            torch.cuda.synchronize()
            ms_lv = bn.time_steps([lambda: flt(xin)], max(3, steps // 4), False)
            es, eb, _ = ref.filter_per_level(np.concatenate([t[:2].cpu().numpy() for t in lv_s], axis=1),
                                             np.concatenate([t[:2].cpu().numpy() for t in lv_b], axis=1), 1000, bounds,
                                             per_class=True, threads=ref.hardware_threads())
            same = bool(np.array_equal(o3['scores'][:2].cpu().numpy(), es)
                        and np.array_equal(o3['boxes'][:2].cpu().numpy(), eb))
            extras['per_level_top_k'] = {
                'batch': Bl, 'k_per_level': 1000, 'rows_out': int(o3['scores'].shape[1]), 'ms_per_step': ms_lv,
                'images_per_s': Bl / ms_lv * 1e3, 'bit_exact_vs_oracle': same,
                'note': 'stage entry rpp_topk_levels on per-level score tensors [B,n_l,80] (per-class filter, '
                        'outputs [B,4125,80] + boxes [B,4125,80,4] materialised); reported separately: the '
                        'reference has no per-level mode'}
            del lv_s, lv_b, o3, flt, xin
            torch.cuda.empty_cache()
        except Exception as exc:   # informational row: never fails the bench
            extras['per_level_top_k'] = {'error': repr(exc)[:200]}

    # ---- end to end: pinned host buffers -> rpp_detect_host -> host outputs ------------------------------------
    e2e = None
    if not args.quick:
        note(rank, 'end to end')
        e2e_steps = max(1, min(steps, 10))
        h_logits = torch.empty(logits.shape, dtype=torch.float32, pin_memory=True).copy_(logits)
        h_deltas = torch.empty(deltas.shape, dtype=torch.float32, pin_memory=True).copy_(deltas)
        ho = {'boxes': torch.empty((B, M, 4), dtype=torch.float32, pin_memory=True),
              'scores': torch.empty((B, M), dtype=torch.float32, pin_memory=True),
              'classes': torch.empty((B, M), dtype=h.class_dtype, pin_memory=True),
              'valid': torch.empty((B,), dtype=torch.int32, pin_memory=True)}

        def host_step():
            _native.check(L.rpp_detect_host(h.ptr, local_rank, h_deltas.data_ptr(), h_logits.data_ptr(), B,
                                            ho['boxes'].data_ptr(), ho['scores'].data_ptr(),
                                            ho['classes'].data_ptr(), ho['valid'].data_ptr()))
        host_step()
        host_step()
        ref_out = layer(xs[0])
        same = bool((ho['valid'].to(dev) == ref_out['valid_detections']).all().item()) and \
            bool((ho['scores'].to(dev) == ref_out['scores']).all().item())
        bn.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()
        e2e_s = bn.max_over_ranks(time.perf_counter() - t0)
        e2e_value = global_batch * e2e_steps / e2e_s
        h2d = B * (N * C * 4 + N * 16)
        d2h = B * (M * 16 + M * 4 + M * h.class_dtype.itemsize + 4)
        # the box's ceiling for this copy pattern: the same pinned buffers through plain cudaMemcpyAsync, all ranks
        # at once (what rpp_detect_host can reach if compute hides completely under the copies)
        bn.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            logits.copy_(h_logits, non_blocking=True)
            deltas.copy_(h_deltas, non_blocking=True)
            torch.cuda.synchronize()
        copy_s = bn.max_over_ranks(time.perf_counter() - t0)
        ceil_gbs = h2d * e2e_steps / copy_s / 1e9
        ceil_img = global_batch * e2e_steps / copy_s
        e2e = {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
               'steps': e2e_steps, 'matches_device_path': same,
               'api': 'rpp_detect_host (pinned host buffers in, host detections out, chunked H2D overlapped)',
               'h2d_ceiling': {'gb_per_s_per_rank': ceil_gbs, 'images_per_s': ceil_img,
                               'how': 'same pinned buffers, plain cudaMemcpyAsync, all {} rank(s) concurrently, '
                                      'max over ranks'.format(world)},
               'frac_of_h2d_ceiling': e2e_value / ceil_img}
        if key == 'c2':
            # informational: the same host call with bf16 host buffers (half the PCIe bytes)
            hb_logits = torch.empty(logits.shape, dtype=torch.bfloat16, pin_memory=True).copy_(logits)
            hb_deltas = torch.empty(deltas.shape, dtype=torch.bfloat16, pin_memory=True).copy_(deltas)

            def host_step_bf16():
                _native.check(L.rpp_detect_host_typed(h.ptr, local_rank, hb_deltas.data_ptr(), hb_logits.data_ptr(),
                                                      2, B, ho['boxes'].data_ptr(), ho['scores'].data_ptr(),
                                                      ho['classes'].data_ptr(), ho['valid'].data_ptr()))
            host_step_bf16()
            bn.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                host_step_bf16()
            e2e['bf16_host_buffers_images_per_s_rank0'] = B * e2e_steps / (time.perf_counter() - t0)
            del hb_logits, hb_deltas
        del h_logits, h_deltas

    del xs, replays, logits, deltas, out, og
    torch.cuda.empty_cache()

    # ---- strong scaling of the two BASELINE configs that are quoted on a TOTAL batch (N > 1) ---------------------
    strong = None
    if world > 1 and key == 'c2' and not args.quick:
        from retinanet.distributed import shard_range
        strong = {}
        for k2 in ('c3', 'c4'):
            note(rank, 'strong scaling {}'.format(k2))
            w2 = WORKLOADS[k2]
            p2 = workload_params(w2)
            lay = FusedPostProcessing(p2)
            lo, hi = shard_range(w2['batch'], rank, world)
            res = {}
            for tag, b2 in (('sharded', hi - lo), ('one_gpu_full_batch', w2['batch'])):
                x = bn.inputs(w2, p2, b2, 'dense', seed_offset=rank)
                for _ in range(3):
                    lay(x)
                rp, _ = lay.capture(x)
                rp()
                bn.barrier()
                ms = bn.time_steps([rp], steps, b2 * path_bytes_per_image(w2) < 190e6)
                bn.barrier()
                res[tag] = bn.max_over_ranks(ms)
                del x, rp
                torch.cuda.empty_cache()
            strong[k2] = {
                'workload': '{}: {}'.format(w2['cfg'], w2['text']), 'scaling': 'strong',
                'global_batch': w2['batch'], 'batch_per_gpu': hi - lo,
                'ms_per_step': res['sharded'], 'images_per_s': w2['batch'] / res['sharded'] * 1e3,
                'one_gpu_full_batch_ms': res['one_gpu_full_batch'],
                'one_gpu_images_per_s': w2['batch'] / res['one_gpu_full_batch'] * 1e3,
                'speedup_vs_one_gpu': res['one_gpu_full_batch'] / res['sharded'],
                'efficiency': res['one_gpu_full_batch'] / res['sharded'] / world,
                'note': 'max over ranks, CUDA events; one_gpu_* = the full batch on each GPU alone in the same run'}
            del lay
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- every other BASELINE config on this GPU (N = 1) ---------------------------------------------------------
    configs = None
    if world == 1 and args.configs == 'all' and key == 'c2' and not args.quick:
        configs = {}
        csteps = max(5, min(steps, 20))
        for k2 in ('c1', 'c3', 'c4', 'c5', 'c2s'):
            note(rank, 'config {}'.format(k2))
            configs[k2] = bn.config_row(k2, WORKLOADS[k2]['batch'], csteps, args.check)
        note(rank, 'config c2 (parity, sparse, clustered)')
        # the headline config's own parity counts + its sparse / clustered throughput
        configs['c2'] = bn.config_row('c2', B, csteps, args.check, timed_dists=('sparse', 'clustered'))

    achieved = B * logit_bytes / (collect_ms * 1e-3) / 1e9 if collect_ms > 0 else None
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')) as f:
            tj = json.load(f)
        if key == 'c2' and B == 64:
            traffic = tj.get('collect_dram_bytes_per_launch')
            traffic_src = 'static: profiles/roofline_traffic.json (ncu --set full capture of this kernel, B = 64)'
    except Exception:
        pass
    line = {
        'metric': metric_name(wl), 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': steps,
        'warmup': warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': scaling,
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(wl, B, args.logits, world, scaling),
        'clocks': clocks,
        'e2e': e2e,
        'gpu_launches': launches_per_step * steps,
        'gpu_launches_per_step': launches_per_step,
        'roofline': {
            'bound': 'hbm', 'kernel': 'collect kernel (streams class_logits once)',
            'achieved': achieved, 'peak': bn.peak, 'unit': 'GB/s', 'frac': achieved / bn.peak if achieved else None,
            'peak_source': bn.peak_src, 'traffic': traffic, 'traffic_source': traffic_src,
            'algorithmic_bytes_per_launch': B * logit_bytes,
            'kernel_ms': collect_ms,
            'path_bytes_per_image': bpi,
            'path_frac': (B * bpi / (ms_step * 1e-3) / 1e9) / bn.peak,
        },
        'stage_ms': dict(rep, calls=ncalls, dominant=dom[0],
                         note='separate K-step pass, eager calls, CUDA events at the stage boundaries'),
        'mean_valid_detections': valid_mean,
        'ms_per_step_eager': ms_eager,
        'launch': 'timed region replays CUDA graphs of one step (FusedPostProcessing.capture), alternating between '
                  '{} distinct input set(s); ms_per_step_eager = plain calls'.format(1 if flush else 2),
    }
    line.update(extras)
    if strong is not None:
        line['strong_scaling'] = strong
    if configs is not None:
        line['configs'] = configs
    if world == 1 and not args.no_cpu_baseline and not args.quick:
        from oracle import ref
        ref.build()
        threads = ref.hardware_threads()
        sample = 16 if H <= 640 else 4
        dname = args.logits if args.logits != 'clustered' else 'dense'
        note(rank, 'cpu baseline')
        rate, ms_pass, passes = cpu_oracle_rate(wl, sample, dname, threads, 12.0, 40)
        line['cpu_baseline'] = {
            'value': rate, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
            'sample': '{} images of the same workload per pass; one warm-up pass, then {} timed passes ({:.1f} s), '
                      '{} host threads (oracle/retinapost_ref.cpp)'.format(sample, passes, passes * ms_pass / 1e3,
                                                                            threads)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--scaling', default=None, choices=['weak', 'strong'],
                    help='default: the workload\'s own (c3 / c4 are quoted on a TOTAL batch: strong)')
    ap.add_argument('--batch', type=int, default=0, help='images per GPU per step (weak) or in total (strong)')
    ap.add_argument('--logits', default='dense', choices=['dense', 'sparse', 'clustered'])
    ap.add_argument('--configs', default='all', choices=['all', 'none'],
                    help='N = 1, workload c2: also report every other BASELINE config')
    ap.add_argument('--check', type=int, default=16, help='images per distribution checked against the oracle')
    ap.add_argument('--quick', action='store_true', help='headline numbers only (profiling runs)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--watchdog', type=int, default=900,
                    help='seconds after which a stuck run dumps its Python stacks and exits (0: off)')
    args = ap.parse_args()
    if args.watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
