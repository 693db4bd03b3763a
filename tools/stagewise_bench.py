#!/usr/bin/env python
"""Times the reference-shaped layer-by-layer graph (fused=False: rpp_decode -> rpp_topk -> rpp_nms, intermediates in
HBM exactly like the reference) next to the fused call, at configs[1] shapes."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'retinanet-tensorflow2.x_b200'))
import torch
import bench
from retinanet.cfg.config import AttrDict
from retinanet.model.builder import ModelBuilder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
params = AttrDict(bench.CONFIG)
g = torch.Generator(device='cuda'); g.manual_seed(42)
logits = torch.randn((B, bench.N_ANCHORS, bench.C), generator=g, device='cuda')
deltas = (torch.randn((B, bench.N_ANCHORS, 4), generator=g, device='cuda') * 0.5).clamp_(-4, 4)
x = {'class_logits': logits, 'encoded_boxes': deltas}


def timeit(fn, K=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(K):
        out = fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / K, out

fused = ModelBuilder(params).add_post_processing_stage(None).layers[-1]
t_f, o_f = timeit(lambda: fused(x), 20)
stages = ModelBuilder(params).add_post_processing_stage(None, fused=False).layers[1:]
y = x
for st in stages:
    t, y2 = timeit(lambda st=st, y=y: st(y))
    print('%-28s %.3f ms' % (type(st).__name__, t))
    y = y2
same = all(bool((y[k] == o_f[k]).all()) for k in o_f)
print('fused rpp_detect             %.3f ms   (B=%d)  identical=%s' % (t_f, B, same))
