import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'retinanet-tensorflow2.x_b200'))
import torch, bench
from retinanet import _native
from retinanet.cfg.config import AttrDict
from retinanet.model.builder import ModelBuilder
B, C, H = 64, 80, 640
model = ModelBuilder(AttrDict(bench.BASE_CONFIG)).add_post_processing_stage(None)
layer = model.layers[-1]
h = layer.handle(C)
g = torch.Generator(device='cuda'); g.manual_seed(42)
bounds = [0, 57600, 72000, 75600, 76500, 76725]
heads = {'class-predictions': {}, 'box-predictions': {}}
for li, level in enumerate(range(3, 8)):
    f = -(-H // 2 ** level); n = bounds[li + 1] - bounds[li]
    heads['class-predictions'][str(level)] = torch.randn((B, f, f, 9 * C), generator=g, device='cuda')
    heads['box-predictions'][str(level)] = (torch.randn((B, f, f, 36), generator=g, device='cuda') * 0.5).clamp_(-4, 4)
L = _native.lib()
for _ in range(5): model(heads)
L.rpp_debug_stage_timing(h.ptr, 1)
for _ in range(30): model(heads)
torch.cuda.synchronize()
st = (ctypes.c_float * 4)(); n = ctypes.c_int()
L.rpp_debug_stage_ms(h.ptr, st, ctypes.byref(n)); L.rpp_debug_stage_timing(h.ptr, 0)
print('levels: sample %.4f collect %.4f nms %.4f merge %.4f' % tuple(st))
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): model(heads)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('host launch time per call %.1f us, total %.1f us' % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
