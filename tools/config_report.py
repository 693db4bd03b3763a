#!/usr/bin/env python
"""Throughput + parity of every BASELINE.json configuration on one GPU (markdown table on stdout).
Parity: the first `--check` images of the batch against the CPU oracle (bit-exact kept detections, classes, order,
scores; boxes within 1e-5)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'retinanet-tensorflow2.x_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from _util import image_mismatches, make_params, oracle_detect  # noqa: E402
from oracle import ref  # noqa: E402
from retinanet.model.layers import FusedPostProcessing  # noqa: E402

CONFIGS = [
    ('C1', 640, 80, 1, dict(mode='CombinedNMS', pre_nms_top_k=5000, filter_per_class=True)),
    ('C2', 640, 80, 64, dict(mode='PerClassHardNMS', pre_nms_top_k=5000, filter_per_class=True)),
    ('C3', 640, 80, 64, dict(mode='GlobalSoftNMS', pre_nms_top_k=5000, filter_per_class=False, soft_nms_sigma=0.5)),
    ('C4', 1024, 80, 32, dict(mode='CombinedNMS', pre_nms_top_k=5000, filter_per_class=True)),
    ('C5', 320, 5, 512, dict(mode='GlobalHardNMS', pre_nms_top_k=5000, filter_per_class=False)),
    ('C2s', 640, 80, 64, dict(mode='PerClassSoftNMS', pre_nms_top_k=5000, filter_per_class=True)),
    # the reference's TPUStrategy branches (opt-in tpu_semantics)
    ('C2t', 640, 80, 64, dict(mode='PerClassHardNMS', pre_nms_top_k=5000, filter_per_class=True, tpu_semantics=True)),
    ('C3t', 640, 80, 64, dict(mode='GlobalHardNMS', pre_nms_top_k=5000, filter_per_class=False, tpu_semantics=True)),
    ('C5t', 320, 5, 512, dict(mode='GlobalHardNMS', pre_nms_top_k=5000, filter_per_class=False, tpu_semantics=True)),
]


def effnms_report(steps, check):
    """the EfficientNMS_TRT-shaped entry at configs[1] geometry."""
    from retinanet.onnx_utils import EfficientNMSPlugin
    H, C, B = 640, 80, 64
    p = make_params(H, num_classes=C, max_detections=100)
    plugin = EfficientNMSPlugin(p)
    N = plugin.anchor_boxes.shape[1]
    for dist in ('dense', 'sparse'):
        g = torch.Generator(device='cuda')
        g.manual_seed(42)
        logits = torch.randn((B, N, C), generator=g, device='cuda')
        if dist == 'sparse':
            logits.mul_(1.5).add_(-4.595)
        g.manual_seed(1234)
        deltas = (torch.randn((B, N, 4), generator=g, device='cuda') * 0.5).clamp_(-4, 4)
        for _ in range(3):
            out = plugin(deltas, logits)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(steps):
            s.record()
            out = plugin(deltas, logits)
            e.record()
            torch.cuda.synchronize()
            tot += s.elapsed_time(e)
        ms = tot / steps
        t0 = time.perf_counter()
        ev, eb, es, ec = ref.efficient_nms(deltas[:check].cpu().numpy(), logits[:check].cpu().numpy(),
                                           plugin.anchor_boxes.cpu().numpy(), 100, p.inference.score_threshold,
                                           p.inference.iou_threshold, threads=ref.hardware_threads())
        cpu_rate = check / (time.perf_counter() - t0)
        ok = sum(int(np.array_equal(out[0][b].cpu().numpy(), ev[b]) and np.array_equal(out[3][b].cpu().numpy(), ec[b])
                     and np.array_equal(out[2][b].cpu().numpy(), es[b])
                     and np.allclose(out[1][b].cpu().numpy(), eb[b], rtol=1e-5, atol=1e-4)) for b in range(check))
        print('| Eff | {}x{}, C={} (N={}) | EfficientNMS_TRT entry / 4096 best pairs | {} | {} | {:.3f} | {:,.0f} | {:.1f} | {}/{} |'
              .format(H, H, C, N, B, dist, ms, B / ms * 1e3, cpu_rate, ok, check), flush=True)



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--check', type=int, default=8)
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    ref.build()
    print('| cfg | shape | mode / filter | B | logits | ms/step | images/s | oracle images/s ({} thr) | bit-exact images |'
          .format(ref.hardware_threads()))
    print('|---|---|---|---|---|---|---|---|---|')
    for name, H, C, B, inf in CONFIGS:
        if args.only and name not in args.only.split(','):
            continue
        p = make_params(H, num_classes=C, max_detections=100, **inf)
        layer = FusedPostProcessing(p)
        N = layer.handle(C).num_anchors
        for dist in ('dense', 'sparse'):
            g = torch.Generator(device='cuda')
            g.manual_seed(42)
            logits = torch.randn((B, N, C), generator=g, device='cuda')
            if dist == 'sparse':
                logits.mul_(1.5).add_(-4.595)
            g.manual_seed(1234)
            deltas = (torch.randn((B, N, 4), generator=g, device='cuda') * 0.5).clamp_(-4, 4)
            x = {'class_logits': logits, 'encoded_boxes': deltas}
            for _ in range(3):
                out = layer(x)
            torch.cuda.synchronize()
            if B * N * C * 4 < 200e6:   # fits in L2: flush between iterations
                flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
            else:
                flush = None
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tot = 0.0
            for _ in range(args.steps):
                if flush is not None:
                    flush.zero_()
                s.record()
                out = layer(x)
                e.record()
                torch.cuda.synchronize()
                tot += s.elapsed_time(e)
            ms = tot / args.steps
            nchk = min(args.check, B)
            t0 = time.perf_counter()
            exp = oracle_detect(ref, p, logits[:nchk].cpu().numpy(), deltas[:nchk].cpu().numpy(), threads=ref.hardware_threads())
            cpu_rate = nchk / (time.perf_counter() - t0)
            got = {k: v[:nchk].cpu().numpy() for k, v in out.items()}
            bad = image_mismatches(got, exp)
            print('| {} | {}x{}, C={} (N={}) | {} / {} | {} | {} | {:.3f} | {:,.0f} | {:.1f} | {}/{} |'.format(
                name, H, H, C, N, inf['mode'],
                ('per-class k=5000' if inf['filter_per_class'] else 'global k=5000') +
                (' (TPU branch)' if inf.get('tpu_semantics') else ''), B, dist, ms, B / ms * 1e3,
                cpu_rate, nchk - len(bad), nchk), flush=True)
        del layer
        torch.cuda.empty_cache()
    if not args.only or 'Eff' in args.only.split(','):
        effnms_report(args.steps, min(args.check, 8))


if __name__ == '__main__':
    main()
