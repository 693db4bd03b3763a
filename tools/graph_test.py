import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/retinanet-tensorflow2.x_b200')
import torch
import bench
from retinanet.cfg.config import AttrDict
from retinanet.model.layers import FusedPostProcessing
for B in (64, 1, 8):
    params = AttrDict(bench.CONFIG)
    layer = FusedPostProcessing(params)
    g = torch.Generator(device='cuda'); g.manual_seed(42)
    logits = torch.randn((B, bench.N_ANCHORS, bench.C), generator=g, device='cuda')
    deltas = (torch.randn((B, bench.N_ANCHORS, 4), generator=g, device='cuda') * 0.5).clamp_(-4, 4)
    x = {'class_logits': logits, 'encoded_boxes': deltas}
    for _ in range(5): out = layer(x)
    torch.cuda.synchronize()
    def timeit(fn, K=200):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(K): fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / K
    t_plain = timeit(lambda: layer(x))
    graph = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3): layer(x)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        out_g = layer(x)
    torch.cuda.synchronize()
    t_graph = timeit(graph.replay)
    same = all(bool((out_g[k] == out[k]).all()) for k in out)
    print('B=%d plain %.4f ms  graph %.4f ms  same=%s' % (B, t_plain, t_graph, same))
