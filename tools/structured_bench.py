#!/usr/bin/env python
"""Detector-like inputs at configs[1] shapes: low background logits plus a few dozen "objects" per image, each a blob
of neighbouring anchors (same class, near-identical boxes, high scores) — heavy suppression, few classes active.
Checks parity on a few images and times the fused call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'retinanet-tensorflow2.x_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, bench
from _util import image_mismatches, make_params, oracle_detect
from oracle import ref
from retinanet.model.layers import FusedPostProcessing

B, N, C = 64, bench.N_ANCHORS, bench.C
rng = np.random.default_rng(7)
for mode in ('PerClassHardNMS', 'CombinedNMS', 'PerClassSoftNMS'):
    p = make_params(640, num_classes=C, mode=mode)
    layer = FusedPostProcessing(p)
    logits = torch.randn((B, N, C), device='cuda') * 1.0 - 6.0          # background p ~ 0.0025
    deltas = (torch.randn((B, N, 4), device='cuda') * 0.05)
    anchors, _ = ref.anchors(640, 640, 3, 7, p.anchor_params.areas, p.anchor_params.aspect_ratios, p.anchor_params.scales)
    a = torch.from_numpy(anchors).cuda()
    for b in range(B):
        nobj = int(rng.integers(5, 40))
        for _ in range(nobj):
            cx, cy = rng.uniform(50, 590, 2); size = rng.uniform(30, 300); cls = int(rng.integers(0, C))
            # anchors whose centre is near the object and whose size matches: they all fire
            d = ((a[:, 0] - cx).abs() < size * 0.2) & ((a[:, 1] - cy).abs() < size * 0.2) & \
                (a[:, 2] > size * 0.6) & (a[:, 2] < size * 1.6) & (a[:, 3] > size * 0.6) & (a[:, 3] < size * 1.6)
            idx = d.nonzero()[:, 0]
            if len(idx) == 0: continue
            logits[b, idx, cls] = torch.randn(len(idx), device='cuda') * 1.0 + 3.0
            # regress every firing anchor onto (almost) the same box
            tx = (cx - a[idx, 0]) / a[idx, 2]; ty = (cy - a[idx, 1]) / a[idx, 3]
            tw = torch.log(size / a[idx, 2]); th = torch.log(size / a[idx, 3])
            deltas[b, idx] = torch.stack([tx, ty, tw, th], 1).float() + torch.randn((len(idx), 4), device='cuda') * 0.03
    x = {'class_logits': logits, 'encoded_boxes': deltas}
    for _ in range(3): out = layer(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): out = layer(x)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    nchk = 6
    exp = oracle_detect(ref, p, logits[:nchk].cpu().numpy(), deltas[:nchk].cpu().numpy(), threads=16)
    got = {k: v[:nchk].cpu().numpy() for k, v in out.items()}
    bad = image_mismatches(got, exp)
    print('%-16s %.3f ms/step  %.0f images/s  mean valid %.1f  bit-exact %d/%d' % (
        mode, ms, B / ms * 1e3, out['valid_detections'].float().mean().item(), nchk - len(bad), nchk))
