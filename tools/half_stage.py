import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'retinanet-tensorflow2.x_b200'))
import torch, bench
from retinanet import _native
from retinanet.cfg.config import AttrDict
from retinanet.model.layers import FusedPostProcessing
B = 64
layer = FusedPostProcessing(AttrDict(bench.BASE_CONFIG))
h = layer.handle(80)
g = torch.Generator(device='cuda'); g.manual_seed(42)
logits = torch.randn((B, 76725, 80), generator=g, device='cuda')
deltas = (torch.randn((B, 76725, 4), generator=g, device='cuda') * 0.5).clamp_(-4, 4)
L = _native.lib()
for name, tdt in (('f32', torch.float32), ('bf16', torch.bfloat16), ('f16', torch.float16)):
    x = {'class_logits': logits.to(tdt), 'encoded_boxes': deltas.to(tdt)}
    for _ in range(5): layer(x)
    L.rpp_debug_stage_timing(h.ptr, 1)
    for _ in range(30): layer(x)
    torch.cuda.synchronize()
    st = (ctypes.c_float * 4)(); n = ctypes.c_int()
    L.rpp_debug_stage_ms(h.ptr, st, ctypes.byref(n)); L.rpp_debug_stage_timing(h.ptr, 0)
    print(name, 'sample %.4f collect %.4f nms %.4f merge %.4f' % tuple(st))
