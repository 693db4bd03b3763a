import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'retinanet-tensorflow2.x_b200'))
import torch, bench
from retinanet.cfg.config import AttrDict
from retinanet.model.layers import FusedPostProcessing
B = 64
g = torch.Generator(device='cuda'); g.manual_seed(42)
xs = []
for i in range(2):
    logits = torch.randn((B, bench.N_ANCHORS, bench.C), generator=g, device='cuda')
    deltas = (torch.randn((B, bench.N_ANCHORS, 4), generator=g, device='cuda') * 0.5).clamp_(-4, 4)
    xs.append({'class_logits': logits, 'encoded_boxes': deltas})
layers = [FusedPostProcessing(AttrDict(bench.CONFIG)) for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]
for i in range(2):
    with torch.cuda.stream(streams[i]):
        for _ in range(3): layers[i](xs[i])
torch.cuda.synchronize()
K = 100
# serial on one stream
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for k in range(K): layers[0](xs[k % 2])
e.record(); torch.cuda.synchronize()
print('serial      %.4f ms/step' % (s.elapsed_time(e) / K))
# two streams, alternating
torch.cuda.synchronize()
s.record()
for k in range(K):
    with torch.cuda.stream(streams[k % 2]):
        if k < 2: streams[k % 2].wait_event(s)
        layers[k % 2](xs[k % 2])
for st in streams: torch.cuda.current_stream().wait_stream(st)
e.record(); torch.cuda.synchronize()
print('2 streams   %.4f ms/step' % (s.elapsed_time(e) / K))
