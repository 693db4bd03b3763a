#!/usr/bin/env python
"""bench.py — images/sec of RetinaNet decode + top-k + NMS at 640x640, 80 classes (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
              bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): 640x640 COCO shapes (N = 76 725 anchors, C = 80), batch 64 per GPU,
PerClassHardNMS (iou 0.5, score 0.05), pre_nms_top_k 5000 per class over the fused anchor axis (= the reference's
FilterTopKDetections), max_detections 100; dense N(0,1) logits (worst case for selection: 99.8 % of
scores exceed the threshold) and N(0, 0.5^2) box deltas, generated on the device.  One "step" = one pass of the
fused path (rpp_detect) over the batch.  Images are sharded by rank with no collective (weak scaling).

One JSON line on stdout (rank 0).  `value` = device-resident inputs; `e2e` = host (pinned) buffers in, host
buffers out, through rpp_detect_host; `roofline` = the streaming collect kernel against measured HBM bandwidth;
`cpu_baseline` / `--impl reference` = the CPU oracle (oracle/, the restated TF path) on this box's host cores.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'retinanet-tensorflow2.x_b200'))

METRIC = 'images/sec decode+NMS @640x640 80-cls (PerClassHardNMS, top-k 5000, 100 dets)'
H = W = 640
C = 80
N_ANCHORS = 76725
M = 100
BYTES_PER_IMAGE = 4 * N_ANCHORS * C + 16 * N_ANCHORS + 24 * M + 4   # 25 782 004
LOGIT_BYTES_PER_IMAGE = 4 * N_ANCHORS * C                           # what the collect kernel streams

CONFIG = {
    'input': {'input_shape': [H, W], 'channels': 3},
    'architecture': {'feature_fusion': {'min_level': 3, 'max_level': 7},
                     'head': {'num_classes': C, 'num_anchors': 9}},
    'anchor_params': {'areas': [1024.0, 4096.0, 16384.0, 65536.0, 262144.0],
                      'aspect_ratios': [0.5, 1.0, 2.0],
                      'scales': [1, 1.2599210498948732, 1.5874010519681994]},
    'encoder_params': {'box_variance': [0.1, 0.1, 0.2, 0.2], 'scale_box_targets': False},
    'inference': {'batch_size': 64, 'mode': 'PerClassHardNMS', 'iou_threshold': 0.5, 'score_threshold': 0.05,
                  'soft_nms_sigma': 0.5, 'pre_nms_top_k': 5000, 'filter_per_class': True, 'max_detections': M},
}


def workload_config(batch, dist, n_gpus):
    return {
        'workload': 'configs[1]: 640x640 80-class synthetic logits, per-class hard NMS (iou 0.5, score 0.05), '
                    'pre_nms_top_k 5000/class over the fused anchor axis, 100 dets',
        'batch_per_gpu': batch, 'global_batch': batch * n_gpus, 'anchors': N_ANCHORS, 'classes': C,
        'logits': dist, 'sharding': 'by image, no collective',
        'l2': 'inputs (1.65 GB per step) exceed the 126 MB L2; no flush needed',
    }


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


def cpu_oracle_rate(batch, dist, steps, warmup, threads):
    """images/s of the CPU oracle (restated TF path) on `batch` images per step."""
    import numpy as np
    from oracle import ref
    ref.build()
    rng = np.random.default_rng(42)
    logits = rng.standard_normal((batch, N_ANCHORS, C)).astype(np.float32)
    if dist == 'sparse':
        logits = logits * 1.5 - 4.595
    deltas = np.clip(rng.standard_normal((batch, N_ANCHORS, 4)) * 0.5, -4, 4).astype(np.float32)
    ap = CONFIG['anchor_params']
    inf = CONFIG['inference']
    anchors, _ = ref.anchors(H, W, 3, 7, ap['areas'], ap['aspect_ratios'], ap['scales'])

    def step():
        return ref.detect(logits, deltas, anchors, H, W, inf['mode'], iou_threshold=inf['iou_threshold'],
                          score_threshold=inf['score_threshold'], soft_nms_sigma=inf['soft_nms_sigma'],
                          pre_nms_top_k=inf['pre_nms_top_k'], filter_per_class=inf['filter_per_class'],
                          max_detections=inf['max_detections'], threads=threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import ref
    ref.build()
    threads = ref.hardware_threads()
    sample = 8
    rate, ms = cpu_oracle_rate(sample, args.logits, args.steps, args.warmup, threads)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.batch, args.logits, args.gpus),
        'cpu_baseline': {'value': rate, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                         'sample': '{} images per step of the same workload (oracle/retinapost_ref.cpp: the '
                                   'reference path with its TF kernels restated; TensorFlow is not installable '
                                   'here), {} host threads'.format(sample, threads)},
        'e2e': {'value': rate, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from retinanet import _native
    from retinanet.cfg.config import AttrDict
    from retinanet.model.layers import FusedPostProcessing

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: libretinapost has no CPU path')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    # bind this rank to the CPUs / memory node next to its GPU: the end-to-end leg streams 1.65 GB per step out of
    # pinned host memory, and with 8 ranks remote-socket buffers would share one inter-socket link
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    L = _native.lib()
    B = args.batch
    params = AttrDict(CONFIG)
    layer = FusedPostProcessing(params)
    h = layer.handle(C)
    assert h.num_anchors == N_ANCHORS

    g = torch.Generator(device=dev)
    g.manual_seed(42 + rank)
    logits = torch.randn((B, N_ANCHORS, C), generator=g, device=dev, dtype=torch.float32)
    if args.logits == 'sparse':
        logits.mul_(1.5).add_(-4.595)
    g.manual_seed(1234 + rank)
    deltas = (torch.randn((B, N_ANCHORS, 4), generator=g, device=dev, dtype=torch.float32) * 0.5).clamp_(-4, 4)
    inputs = {'class_logits': logits, 'encoded_boxes': deltas}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = layer(inputs)
    launches_per_step = int(L.rpp_last_launch_count())
    barrier()

    # eager calls (one ctypes call + 8 launches per step) ...
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        out = layer(inputs)
    end.record()
    barrier()
    ms_eager = start.elapsed_time(end) / args.steps
    # ... and the same step captured once in a CUDA graph (public API: FusedPostProcessing.capture) — the timed region
    replay, out_g = layer.capture(inputs)
    for _ in range(3):
        replay()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    start.record()
    for _ in range(args.steps):
        replay()
    end.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = start.elapsed_time(end)
    assert bool((out_g['scores'] == out['scores']).all().item())
    valid_mean = float(out['valid_detections'].float().mean().item())

    # second pass, same K steps, with CUDA events at the stage boundaries on the launch stream: per-kernel durations
    # for the roofline.  Stage timing serialises the image chunks (no collect/NMS overlap), i.e. each kernel is timed
    # running alone, back to back inside a step.
    _native.check(L.rpp_debug_stage_timing(h.ptr, 1))
    for _ in range(args.steps):
        layer(inputs)
    torch.cuda.synchronize()
    stage = (ctypes.c_float * 4)()
    ncalls = ctypes.c_int()
    _native.check(L.rpp_debug_stage_ms(h.ptr, stage, ctypes.byref(ncalls)))
    _native.check(L.rpp_debug_stage_timing(h.ptr, 0))

    # the same step fed with the model-side input of the path, the per-level NHWC head outputs (views of the same
    # memory layout a detector emits): consumed in place (rpp_detect_levels) vs the reference's FuseDetections concat
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
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for _ in range(args.steps):
            o2 = model(heads)
        e2.record()
        torch.cuda.synchronize()
        levels_ms[name] = s2.elapsed_time(e2) / args.steps
        assert bool((o2['scores'] == out['scores']).all().item())
        del model
    del heads
    torch.cuda.empty_cache()
    # 16-bit head outputs (what a mixed-precision detector can emit; the reference casts them to fp32 first): the
    # dominant stream halves.  Informational: the headline `value` stays on fp32 inputs.
    half_ms = {}
    for name, tdt in (('bf16', torch.bfloat16), ('f16', torch.float16)):
        xh = {'class_logits': logits.to(tdt), 'encoded_boxes': deltas.to(tdt)}
        for _ in range(3):
            layer(xh)
        torch.cuda.synchronize()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for _ in range(args.steps):
            layer(xh)
        e2.record()
        torch.cuda.synchronize()
        half_ms[name] = s2.elapsed_time(e2) / args.steps
        del xh
    torch.cuda.empty_cache()

    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total * 1e-3)

    # ---- end to end: pinned host buffers -> rpp_detect_host -> host outputs ------------------------------------
    e2e_steps = max(1, min(args.steps, 10))
    h_logits = torch.empty(logits.shape, dtype=torch.float32, pin_memory=True).copy_(logits)
    h_deltas = torch.empty(deltas.shape, dtype=torch.float32, pin_memory=True).copy_(deltas)
    ho = {'boxes': torch.empty((B, M, 4), dtype=torch.float32, pin_memory=True),
          'scores': torch.empty((B, M), dtype=torch.float32, pin_memory=True),
          'classes': torch.empty((B, M), dtype=torch.int32, pin_memory=True),
          'valid': torch.empty((B,), dtype=torch.int32, pin_memory=True)}

    def host_step():
        _native.check(L.rpp_detect_host(h.ptr, local_rank, h_deltas.data_ptr(), h_logits.data_ptr(), B,
                                        ho['boxes'].data_ptr(), ho['scores'].data_ptr(), ho['classes'].data_ptr(),
                                        ho['valid'].data_ptr()))
    host_step()
    host_step()
    same = bool((ho['valid'].to(dev) == out['valid_detections']).all().item()) and \
        bool((ho['scores'].to(dev) == out['scores']).all().item())
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = B * world * e2e_steps / float(t.item())
    # informational: the same host call with bf16 host buffers (half the PCIe bytes)
    hb_logits = torch.empty(logits.shape, dtype=torch.bfloat16, pin_memory=True).copy_(logits)
    hb_deltas = torch.empty(deltas.shape, dtype=torch.bfloat16, pin_memory=True).copy_(deltas)

    def host_step_bf16():
        _native.check(L.rpp_detect_host_typed(h.ptr, local_rank, hb_deltas.data_ptr(), hb_logits.data_ptr(), 2, B,
                                              ho['boxes'].data_ptr(), ho['scores'].data_ptr(),
                                              ho['classes'].data_ptr(), ho['valid'].data_ptr()))
    host_step_bf16()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step_bf16()
    e2e_bf16 = B * e2e_steps / (time.perf_counter() - t0)
    del hb_logits, hb_deltas
    h2d = B * (N_ANCHORS * C * 4 + N_ANCHORS * 16)
    d2h = B * (M * 16 + M * 4 + M * 4 + 4)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    collect_ms = float(stage[1])
    achieved = B * LOGIT_BYTES_PER_IMAGE / (collect_ms * 1e-3) / 1e9 if collect_ms > 0 else None
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')) as f:
            traffic = json.load(f).get('collect_dram_bytes_per_launch')
    except Exception:
        pass
    line = {
        'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(B, args.logits, world),
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'steps': e2e_steps, 'matches_device_path': same,
                'api': 'rpp_detect_host (pinned host buffers in, host detections out, chunked H2D overlapped)',
                'bf16_host_buffers_images_per_s_rank0': e2e_bf16},
        'gpu_launches': launches_per_step * args.steps,
        'gpu_launches_per_step': launches_per_step,
        'roofline': {
            'bound': 'hbm', 'kernel': 'collect_cols4_kernel (streams class_logits once)',
            'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak if achieved else None,
            'peak_source': peak_src, 'traffic': traffic,
            'algorithmic_bytes_per_launch': B * LOGIT_BYTES_PER_IMAGE,
            'kernel_ms': collect_ms,
            'path_bytes_per_image': BYTES_PER_IMAGE,
            'path_frac': (B * BYTES_PER_IMAGE / (ms_step * 1e-3) / 1e9) / peak,
        },
        'stage_ms': {'sample': float(stage[0]), 'collect': float(stage[1]), 'nms': float(stage[2]),
                     'merge': float(stage[3]), 'calls': int(ncalls.value),
                     'note': 'separate K-step pass, stages serialised (in the timed region collect of image chunk '
                             'i+1 overlaps NMS+merge of chunk i on a side stream)'},
        'mean_valid_detections': valid_mean,
        'ms_per_step_eager': ms_eager,
        'half_precision_inputs': {'ms_per_step': half_ms, 'images_per_s': {k: B / v * 1e3 for k, v in half_ms.items()},
                                  'note': 'same workload with f16 / bf16 logits and deltas read in place '
                                          '(rpp_detect_typed), eager calls'},
        'launch': 'timed region replays a CUDA graph of one step (FusedPostProcessing.capture); ms_per_step_eager = '
                  'plain calls',
        'from_head_levels': {'ms_per_step': levels_ms, 'images_per_s': {k: B / v * 1e3 for k, v in levels_ms.items()},
                             'note': 'per-level NHWC head outputs as input (rank 0): rpp_detect_levels in place vs '
                                     'FuseDetections concat + rpp_detect'},
    }
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ref
        ref.build()
        threads = ref.hardware_threads()
        sample = 16
        rate, _ = cpu_oracle_rate(sample, args.logits, 1, 0, threads)
        line['cpu_baseline'] = {
            'value': rate, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
            'sample': '{} images of the same workload, one pass, {} host threads (oracle/retinapost_ref.cpp)'
                      .format(sample, threads)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='images per GPU per step')
    ap.add_argument('--logits', default='dense', choices=['dense', 'sparse'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
