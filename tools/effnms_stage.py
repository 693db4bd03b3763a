#!/usr/bin/env python
"""Stage times (thresholds | collect | emission | NMS) of the EfficientNMS_TRT-shaped entry at configs[1] geometry."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'retinanet-tensorflow2.x_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from _util import make_params  # noqa: E402
from retinanet import _native  # noqa: E402
from retinanet.onnx_utils import EfficientNMSPlugin  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
p = make_params(640, num_classes=80, max_detections=100)
plugin = EfficientNMSPlugin(p)
N = plugin.anchor_boxes.shape[1]
for dist in ('dense', 'sparse'):
    g = torch.Generator(device='cuda')
    g.manual_seed(42)
    logits = torch.randn((B, N, 80), generator=g, device='cuda')
    if dist == 'sparse':
        logits.mul_(1.5).add_(-4.595)
    deltas = (torch.randn((B, N, 4), generator=g, device='cuda') * 0.5).clamp_(-4, 4)
    for _ in range(2):
        plugin(deltas, logits)
    h = plugin._handle(80)
    L = _native.lib()
    L.rpp_debug_stage_timing(h.ptr, 1)
    for _ in range(5):
        plugin(deltas, logits)
    torch.cuda.synchronize()
    ms = (ctypes.c_float * 4)()
    n = ctypes.c_int()
    L.rpp_debug_stage_ms(h.ptr, ms, ctypes.byref(n))
    L.rpp_debug_stage_timing(h.ptr, 0)
    print(dist, 'thresholds %.3f  collect %.3f  emission %.3f  nms %.3f ms (%d calls)' % (ms[0], ms[1], ms[2], ms[3], n.value))
