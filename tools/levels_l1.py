import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'retinanet-tensorflow2.x_b200'))
import torch, bench
from retinanet import _native
from retinanet.cfg.config import AttrDict
from retinanet.model.layers import FusedPostProcessing
B = 64
layer = FusedPostProcessing(AttrDict(bench.CONFIG))
h = layer.handle(bench.C)
g = torch.Generator(device='cuda'); g.manual_seed(42)
logits = torch.randn((B, bench.N_ANCHORS, bench.C), generator=g, device='cuda')
deltas = (torch.randn((B, bench.N_ANCHORS, 4), generator=g, device='cuda') * 0.5).clamp_(-4, 4)
L = _native.lib()
def run(fn, name):
    for _ in range(5): fn()
    L.rpp_debug_stage_timing(h.ptr, 1)
    for _ in range(30): fn()
    torch.cuda.synchronize()
    st = (ctypes.c_float * 4)(); n = ctypes.c_int()
    L.rpp_debug_stage_ms(h.ptr, st, ctypes.byref(n)); L.rpp_debug_stage_timing(h.ptr, 0)
    print(name, 'sample %.4f collect %.4f nms %.4f merge %.4f' % tuple(st))
run(lambda: layer({'class_logits': logits, 'encoded_boxes': deltas}), 'fused kernel      ')
run(lambda: layer._call_pieces([logits], [deltas]), 'levels kernel, L=1')
bounds = [0, 57600, 72000, 75600, 76500, 76725]
torch.cuda.empty_cache()
cl = [logits[:, bounds[i]:bounds[i+1]].contiguous() for i in range(5)]
bl = [deltas[:, bounds[i]:bounds[i+1]].contiguous() for i in range(5)]
run(lambda: layer._call_pieces(cl, bl), 'levels kernel, L=5')
# experiments: same 5 pieces but as VIEW slices of one big allocation per piece order
big = torch.empty((B * bench.N_ANCHORS * bench.C,), device='cuda')
off = 0; cl2 = []
for i in range(5):
    n = bounds[i+1] - bounds[i]
    t = big[off:off + B * n * bench.C].view(B, n, bench.C); t.copy_(cl[i]); cl2.append(t); off += B * n * bench.C
run(lambda: layer._call_pieces(cl2, bl), 'L=5, pieces packed in one allocation')
