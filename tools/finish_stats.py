#!/usr/bin/env python
"""Diagnostics of the per-class finish pass: worklist size, list lengths and kept counts of the listed problems, read
straight out of the workspace (layout of run_problem_set in csrc/rpp_api.cu).  usage: tools/finish_stats.py [dist]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'retinanet-tensorflow2.x_b200'))
import numpy as np, torch, bench
from retinanet.model.layers import FusedPostProcessing

dist = sys.argv[1] if len(sys.argv) > 1 else 'clustered'
wl = bench.WORKLOADS['c2']; params = bench.workload_params(wl)
B, C, M = 64, wl['C'], bench.M
class A: pass
bb = bench.Bench(A(), 0, 0, 1)
layer = FusedPostProcessing(params); h = layer.handle(C)
x = bb.inputs(wl, params, B, dist)
out = layer(x); torch.cuda.synchronize()
ws = h.workspace(B, 0, x["class_logits"].device)
P = B * C
al = lambda v: (v + 255) // 256 * 256
off = 0
T_off = off; off = al(off + P * 4)
cc_off = off; off = al(off + P * 4)
tc_off = off; off = al(off + 64 * 4)
G = 96
gm_off = off; off = al(off + B * G * C * 4)
CAP = int(os.environ.get('RPP_LIST_CAP', 8192))
cand_off = off; off = al(off + P * CAP * 8)
selcnt_off = off; off = al(off + P * 4)
selkey_off = off; off = al(off + P * M * 8)
selbox_off = off; off = al(off + P * M * 16)
bound_off = off; off = al(off + P * 4)
stopL_off = off; off = al(off + B * 4)
work_off = off; off = al(off + P * 4)
w = ws.cpu().numpy()
def view(o, n, dt): return w[o:o + n * np.dtype(dt).itemsize].view(dt)
cc = view(cc_off, P, np.uint32) & 0x7fffffff
ctl = view(tc_off, 64, np.uint32)
nwork = int(ctl[16])
items = view(work_off, P, np.uint32)[:nwork]
selcnt = view(selcnt_off, P, np.int32)
bound = view(bound_off, P, np.float32)
stopL = view(stopL_off, B, np.float32)
T = view(T_off, P, np.float32)
print('dist', dist, 'work items', nwork, 'of', P, ' popped', int(ctl[17]))
print('ctl', ctl[:40]); print('cc head', cc[:8], 'T head', T[:4])
if nwork == 0: sys.exit(0)
print('list length all: mean %.0f max %d ; work items: mean %.0f p50 %d p90 %d max %d' % (
    cc.mean(), cc.max(), cc[items].mean(), np.percentile(cc[items], 50), np.percentile(cc[items], 90), cc[items].max()))
print('kept (work items): mean %.1f max %d ; stop_L mean %.4f min %.4f' % (selcnt[items].mean(), selcnt[items].max(), stopL.mean(), stopL.min()))
per_img = np.bincount(items // C, minlength=B)
print('work items per image: mean %.1f max %d' % (per_img.mean(), per_img.max()))
print('T == T_min fraction %.3f' % float((T <= T.min() + 1e-6).mean()))
