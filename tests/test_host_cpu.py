"""CPU-side checks: the C-ABI library loads and exports every symbol include/retinapost.h declares (no compute
without a GPU), host-side config / builder logic, and image sharding over a world_size-2 gloo group."""
import copy
import ctypes
import json
import os

import numpy as np
import re
import subprocess
import sys

import pytest

from conftest import PKG, REFERENCE_CONFIG, ROOT


def test_abi_exports_every_declared_symbol():
    from retinanet import _native
    header = open(os.path.join(ROOT, 'include', 'retinapost.h')).read()
    declared = sorted(set(re.findall(r'\b(rpp_[a-z0-9_]+)\s*\(', header)))
    assert len(declared) >= 16
    lib = ctypes.CDLL(_native.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert missing == []
    assert sorted(_native.EXPORTS) == declared


def test_no_cpu_fallback_without_gpu():
    """On a box without CUDA the product must fail loudly instead of computing somewhere else."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from retinanet import _native
    from retinanet.model.layers import GenerateDetections
    cfg = _native.RppConfig()
    h = ctypes.c_void_p()
    assert _native.lib().rpp_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    layer = GenerateDetections(mode='PerClassHardNMS', num_classes=2)
    with pytest.raises(RuntimeError):
        layer({'scores': torch.zeros(1, 4, 2), 'boxes': torch.zeros(1, 4, 4)})


def test_tpu_flag_is_validated_like_the_reference_constructor():
    """postprocessing_ops.py:199-208: under a TPUStrategy only GlobalHardNMS / PerClassHardNMS are accepted.  The
    Python layers raise the reference's AssertionError at construction; rpp_create returns RPP_EMODE (validation runs
    before any CUDA call, so this is checkable without a GPU)."""
    from retinanet import _native
    from retinanet.cfg.config import AttrDict
    from retinanet.model.layers import FusedPostProcessing, GenerateDetections
    for mode in ['CombinedNMS', 'GlobalSoftNMS', 'PerClassSoftNMS']:
        with pytest.raises(AssertionError, match='not supported on Cloud TPUs'):
            GenerateDetections(mode=mode, soft_nms_sigma=0.5, tpu_semantics=True)
        params = AttrDict(copy.deepcopy(REFERENCE_CONFIG))
        params.inference.mode = mode
        params.inference.tpu_semantics = True
        with pytest.raises(AssertionError, match='not supported on Cloud TPUs'):
            FusedPostProcessing(params)
    for mode in ['GlobalHardNMS', 'PerClassHardNMS']:
        assert GenerateDetections(mode=mode, tpu_semantics=True)._running_on_tpu
    areas = (ctypes.c_double * 5)(1024.0, 4096.0, 16384.0, 65536.0, 262144.0)
    one = (ctypes.c_double * 1)(1.0)
    cfg = _native.RppConfig()
    cfg.H = cfg.W = 64
    cfg.min_level, cfg.max_level, cfg.num_classes = 3, 7, 4
    cfg.n_areas, cfg.areas = 5, areas
    cfg.n_ratios, cfg.aspect_ratios = 1, one
    cfg.n_scales, cfg.scales = 1, one
    cfg.iou_threshold, cfg.score_threshold, cfg.soft_nms_sigma, cfg.max_detections = 0.5, 0.05, 0.5, 10
    cfg.tpu_semantics = 1
    h = ctypes.c_void_p()
    for mode in (0, 1, 3):
        cfg.mode = mode
        assert _native.lib().rpp_create(ctypes.byref(cfg), ctypes.byref(h)) == _native.RPP_EMODE
        assert 'Cloud TPUs' in _native.last_error()
    cfg.mode, cfg.iou_threshold = 2, 0.0
    assert _native.lib().rpp_create(ctypes.byref(cfg), ctypes.byref(h)) == _native.RPP_EINVAL


def test_product_never_imports_oracle():
    bad = []
    for root, _, files in os.walk(PKG):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(root, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle\b|retinapost_ref|rpp_ref_', text, re.M):
                    bad.append(f)
    assert bad == []


def test_attrdict_and_reference_json(tmp_path):
    from retinanet.cfg.config import AttrDict, Config
    path = tmp_path / 'cfg.json'
    path.write_text(json.dumps(REFERENCE_CONFIG))
    p = Config(str(path)).params
    assert p.inference.mode == 'PerClassHardNMS' and p['inference']['max_detections'] == 100
    assert p.architecture.feature_fusion.min_level == 3
    p.inference.pre_nms_top_k = -1
    assert p['inference']['pre_nms_top_k'] == -1
    assert json.loads(json.dumps(p.inference))['iou_threshold'] == 0.5     # still a plain dict for json.dumps
    with pytest.raises(AttributeError):
        p.inference.nope
    assert isinstance(AttrDict({'a': [{'b': 1}]}).a[0], AttrDict)
    ref_json = '/root/reference/configs/v3-32/mscoco-retinanet-resnet50-640x640-30x-256.json'
    if os.path.exists(ref_json):   # only in the build container
        q = Config(ref_json).params
        assert q.inference == REFERENCE_CONFIG['inference']
        assert q.anchor_params == REFERENCE_CONFIG['anchor_params']


def test_builder_stage_gating():
    from retinanet.cfg.config import AttrDict
    from retinanet.model.builder import ModelBuilder
    from retinanet.model.layers import (FilterTopKDetections, FuseDetections, FusedPostProcessing,
                                        GenerateDetections, TransformBoxesAndScores)
    names = lambda m: [type(layer) for layer in m.layers]   # noqa: E731
    p = AttrDict(REFERENCE_CONFIG)
    b = ModelBuilder(p, run_mode='export')
    assert names(b.add_post_processing_stage(None)) == [FuseDetections, FusedPostProcessing]
    assert names(b.add_post_processing_stage(None, fused=False)) == [
        FuseDetections, TransformBoxesAndScores, FilterTopKDetections, GenerateDetections]
    assert names(b.add_post_processing_stage(None, skip_decoding=True, skip_nms=True)) == [FuseDetections]
    assert names(b.prepare_model_for_export(None, mode='onnx_tensorrt')) == [FuseDetections]
    # tf_tensorrt / onnx force pre_nms_top_k = -1 (builder.py:133-138)
    b2 = ModelBuilder(AttrDict(REFERENCE_CONFIG))
    m = b2.prepare_model_for_export(None, mode='onnx')
    assert b2.params.inference.pre_nms_top_k == -1 and m.fused
    assert names(b2.add_post_processing_stage(None, fused=False)) == [
        FuseDetections, TransformBoxesAndScores, GenerateDetections]
    with pytest.raises(ValueError):
        b.prepare_model_for_export(None, mode='bogus')
    with pytest.raises(AssertionError):
        GenerateDetections(mode='NotAMode')
    bad = AttrDict(REFERENCE_CONFIG)
    bad.inference.mode = 'NotAMode'
    with pytest.raises(AssertionError):
        ModelBuilder(bad).add_post_processing_stage(None)


def test_fuse_detections_layout():
    """FuseDetections on CPU tensors (pure reshapes): anchor order (level, y, x, a), class fastest."""
    import torch
    from retinanet.model.layers import FuseDetections
    B, A, C = 2, 9, 3
    cls, box = {}, {}
    for level, f in zip(range(3, 8), [8, 4, 2, 1, 1]):
        cls[str(level)] = torch.arange(B * f * f * A * C, dtype=torch.float32).reshape(B, f, f, A * C) + level * 1e5
        box[str(level)] = torch.arange(B * f * f * A * 4, dtype=torch.float32).reshape(B, f, f, A * 4) + level * 1e5
    out = FuseDetections(3, 7)({'class-predictions': cls, 'box-predictions': box})
    N = (64 + 16 + 4 + 1 + 1) * A
    assert out['class_logits'].shape == (B, N, C) and out['encoded_boxes'].shape == (B, N, 4)
    assert out['class_logits'][1, 0, 1] == cls['3'][1, 0, 0, 1]
    assert out['class_logits'][0, 64 * A + 10, 2] == cls['4'][0, 0, 1, 1 * C + 2]   # level 4, x=1, anchor 1
    assert out['encoded_boxes'][1, N - 1, 3] == box['7'][1, 0, 0, A * 4 - 1]


def test_shard_range():
    from retinanet.distributed import shard_range
    for n in (0, 1, 7, 8, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from retinanet.distributed import shard_batch, shard_range, gather_detections
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:' + os.environ['PORT'], rank=rank, world_size=world)
B, M = 5, 3                                   # ragged: ranks get 3 and 2 images
full = {'boxes': torch.arange(B * M * 4, dtype=torch.float32).reshape(B, M, 4),
        'scores': torch.arange(B * M, dtype=torch.float32).reshape(B, M),
        'classes': torch.arange(B * M, dtype=torch.int32).reshape(B, M),
        'valid_detections': torch.arange(B, dtype=torch.int32)}
mine = shard_batch(full, rank, world)
lo, hi = shard_range(B, rank, world)
assert mine['scores'].shape[0] == hi - lo
got = gather_detections(mine, num_images=B)
for k in full:
    assert got[k].dtype == full[k].dtype and torch.equal(got[k], full[k]), k
dist.barrier()
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_gather_detections_gloo_world2(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(_GLOO_WORKER)
    import socket
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE='2', PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), PKG], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out.decode()


def test_config_struct_layout_matches_the_header(tmp_path):
    """The ctypes mirror of rpp_config (retinanet/_native.py) has the size and field offsets the C compiler gives the
    struct of include/retinapost.h."""
    from retinanet import _native
    fields = [f[0] for f in _native.RppConfig._fields_]
    src = tmp_path / 'layout.c'
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "retinapost.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(rpp_config));\n' +
                   ''.join('  printf("%zu\\n", offsetof(rpp_config, {}));\n'.format(f) for f in fields) +
                   '  return 0;\n}\n')
    exe = str(tmp_path / 'layout')
    subprocess.check_call(['gcc', str(src), '-I' + os.path.join(ROOT, 'include'), '-o', exe])
    out = [int(v) for v in subprocess.check_output([exe], text=True).split()]
    assert out[0] == ctypes.sizeof(_native.RppConfig)
    assert out[1:] == [getattr(_native.RppConfig, f).offset for f in fields]


def _simulate_list_lengths(plan, n, trials, rng):
    """List length (= elements >= the sampled threshold) over `trials` synthetic N(0,1) columns, mirroring
    sample_max_kernel / sample_rank_*: group g holds rows (r * G + g) * stride, r < rows_per_group; the threshold is
    the rank-th smallest group maximum."""
    on, stride, G, rows, rank, cap, target, fine = plan
    idx = ((np.arange(rows)[:, None] * G + np.arange(G)[None, :]) * stride).ravel()
    assert idx.max() < n
    counts = []
    for _ in range(trials):
        x = rng.standard_normal(n).astype(np.float32)
        gm = x[idx].reshape(rows, G).max(0)
        counts.append(int((x >= np.sort(gm)[rank]).sum()))
    return np.array(counts)


def test_sampling_plan_margins():
    """The sampled pre-threshold only steers speed, but a top-k list that comes up short of k (or overflows) costs a
    fallback: the plans of the BASELINE geometries must keep both tails far away ("Top-k emission robustness")."""
    from retinanet import _native
    rng = np.random.default_rng(0)
    out = (ctypes.c_int * 8)()

    def plan(n, C, k, emit):
        _native.check(_native.lib().rpp_debug_sample_plan(n, C, k, emit, out))
        return list(out)

    # NMS problems of configs[1]: columns of 76 725 logits, lists of ~768 candidates, capacity 8192
    p = plan(76725, 80, 0, 0)
    assert p[0] == 1 and p[5] == 8192 and p[6] == 768
    c = _simulate_list_lengths(p, 76725, 400, rng)
    assert 600 < c.mean() < 950 and c.max() < 4096 and c.min() > 250
    # flat top-k of the global filter (C3) and of the EfficientNMS entry: 6.1 M elements per image, fine plan
    for k in (5000, 4096):
        p = plan(76725 * 80, 1, k, 1)
        assert p[0] == 1 and p[7] == 1 and p[2] >= 256
        c = _simulate_list_lengths(p, 76725 * 80, 24, rng)
        sigma = c.std()
        assert c.min() >= k and c.max() <= min(p[5], 16384)
        assert (c.mean() - k) / sigma > 4.0 and (min(p[5], 16384) - c.mean()) / sigma > 4.0, (k, c.mean(), sigma)
    # short columns keep the cheap plan (their fallback is a scan of < 100 K elements)
    assert plan(19206 * 5, 1, 5000, 1)[7] == 0


def test_kernel_resource_budgets():
    """Build-time guard of the occupancy the design relies on, read from the cubin with cuobjdump
    (no GPU needed): the collect kernels must fit 3 x 512 threads per SM (<= 42 registers, no spills) and stream with
    128-bit loads; the sample kernels 2 x 1024 threads (<= 32 registers); the warp probe <= 64 registers."""
    import shutil
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    lib = os.path.join(PKG, 'libretinapost.so')
    txt = subprocess.check_output([cuobjdump, '-res-usage', lib], text=True, stderr=subprocess.STDOUT)
    usage = {}
    name = None
    for line in txt.splitlines():
        m = re.search(r'Function (\S+):', line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r'REG:(\d+) STACK:(\d+)', line)
        if m and name:
            usage[name] = (int(m.group(1)), int(m.group(2)))
            name = None

    def find(prefix):
        hits = {k: v for k, v in usage.items() if prefix in k}
        assert hits, prefix
        return hits

    for k, (reg, stack) in find('collect_cols4_kernelILi4ELi3E').items():
        assert reg <= 42 and stack == 0, (k, reg, stack)
    for k, (reg, stack) in find('collect_cols4_levels_kernelILi4ELi3E').items():
        assert reg <= 42 and stack == 0, (k, reg, stack)
    for k, (reg, stack) in find('collect_cols8_half_kernelILi8E').items():
        assert reg <= 64 and stack == 0, (k, reg, stack)          # 2 CTAs of 512 threads per SM
    for k, (reg, stack) in find('sample_max').items():
        assert reg <= 32, (k, reg)                                 # 2 blocks of up to 1024 threads per SM
    for k, (reg, stack) in find('probe_warp_kernel').items():
        assert reg <= 64 and stack == 0, (k, reg, stack)
    fused = [k for k in usage if 'collect_cols4_kernelILi4ELi3E' in k][0]
    sass = subprocess.check_output([cuobjdump, '-sass', '-fun', fused, lib], text=True, stderr=subprocess.STDOUT)
    assert sass.count('LDG.E.128') >= 4, 'the collect kernel must keep four 128-bit streaming loads in flight'


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the oracle timed on the host cores) needs no GPU and prints the contract's JSON
    line: impl, metric / unit of the own arm, cpu_baseline describing the run, a zero-copy e2e."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'images/s' and line['higher_is_better'] is True
    assert line['value'] > 0 and line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
    assert line['e2e']['value'] == line['value'] and 'workload' in line['config']


def _build_abi_smoke(tmp_path):
    exe = str(tmp_path / 'abi_smoke')
    cmd = ['gcc', os.path.join(ROOT, 'tests', 'abi_smoke.c'), '-I' + os.path.join(ROOT, 'include'),
           '-I/usr/local/cuda/include', '-L' + PKG, '-lretinapost', '-L/usr/local/cuda/lib64', '-lcudart', '-lm',
           '-Wl,-rpath,' + PKG, '-Wl,-rpath,/usr/local/cuda/lib64', '-o', exe]
    subprocess.check_call(cmd)
    return exe


def test_abi_compiles_and_links_from_plain_c(tmp_path):
    """include/retinapost.h is a C header and the library links from C (no torch, no Python)."""
    assert os.path.exists(_build_abi_smoke(tmp_path))


@pytest.mark.gpu
def test_abi_smoke_runs_from_plain_c(tmp_path):
    out = subprocess.run([_build_abi_smoke(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'abi_smoke ok' in out.stdout


def test_serving_signature_rekeying():
    """export.py:233-253: positional re-keying of the sorted frozen outputs, including the
    mis-named keys of the skip_nms (onnx_tensorrt) case."""
    from retinanet.export import InferenceModule, frozen_outputs, make_inference_module
    det = {'scores': 's', 'boxes': 'b', 'classes': 'c', 'valid_detections': 'v'}
    assert frozen_outputs(det) == ['b', 'c', 's', 'v']
    m = InferenceModule(lambda **kw: det, skip_nms=False)
    assert m.run_inference({}) == {'boxes': 'b', 'scores': 's', 'classes': 'c', 'valid_detections': 'v'}
    fused = {'class_logits': 'L', 'encoded_boxes': 'E'}
    m = InferenceModule(lambda **kw: fused, skip_nms=True)
    assert m.run_inference({}) == {'boxes': 'L', 'scores': 'E'}          # B22: the reference's mis-named keys
    m = make_inference_module(lambda x: dict(det, seen=None) and det, mode='tf')
    assert m.run_inference({'predictions': 1})['valid_detections'] == 'v'
    assert make_inference_module(lambda x: fused, mode='onnx_tensorrt').skip_nms


def test_clustered_synthetic_inputs_are_seeded_and_peaked():
    """tools/synth_inputs.py (bench.py's third logit distribution): same seed -> same tensors; a background floor with
    a small fraction of anchors above the score threshold, clustered on the objects' classes."""
    import torch
    from oracle import ref as _ref
    sys.path.insert(0, ROOT)
    from tools import synth_inputs
    ap = REFERENCE_CONFIG['anchor_params']
    anchors, _ = _ref.anchors(128, 128, 3, 7, ap['areas'], ap['aspect_ratios'], ap['scales'])
    a = synth_inputs.clustered_inputs(2, torch.from_numpy(anchors), 8, 128, 128, 'cpu', seed=5)
    b = synth_inputs.clustered_inputs(2, torch.from_numpy(anchors), 8, 128, 128, 'cpu', seed=5)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    frac = float((torch.sigmoid(a[0]) > 0.05).float().mean())
    assert 0.0005 < frac < 0.2
    assert float(a[1].abs().max()) <= 4.0
    assert synth_inputs.make_inputs('sparse', 1, torch.from_numpy(anchors), 8, 128, 128, 'cpu')[0].mean() < -3


def test_bench_workloads_match_baseline_configs():
    """bench.py's workload table = BASELINE.json `configs`: shapes, modes, batch sizes, byte counts."""
    sys.path.insert(0, ROOT)
    import bench
    w = bench.WORKLOADS
    assert (w['c2']['H'], w['c2']['C'], w['c2']['batch'], w['c2']['mode']) == (640, 80, 64, 'PerClassHardNMS')
    assert (w['c3']['mode'], w['c3']['per_class'], w['c3']['scaling']) == ('GlobalSoftNMS', False, 'strong')
    assert (w['c4']['H'], w['c4']['batch'], w['c4']['mode']) == (1024, 32, 'CombinedNMS')
    assert (w['c5']['H'], w['c5']['C'], w['c5']['batch'], w['c5']['mode']) == (320, 5, 512, 'GlobalHardNMS')
    assert w['c1']['batch'] == 1 and w['c1']['mode'] == 'CombinedNMS'
    assert bench.num_anchors(640) == 76725 and bench.num_anchors(1024) == 196416 and bench.num_anchors(320) == 19206
    assert bench.path_bytes_per_image(w['c2']) == 25782004
    assert bench.path_bytes_per_image(w['c4']) == 65998180 and bench.path_bytes_per_image(w['c5']) == 693820
    p = bench.workload_params(w['c3'])
    assert p.inference.mode == 'GlobalSoftNMS' and p.inference.filter_per_class is False
    assert p.inference.pre_nms_top_k == 5000 and p.inference.soft_nms_sigma == 0.5


def test_collect_kernels_keep_their_loads_in_flight():
    """SASS guard (cuobjdump, no GPU): in the streaming loop of the fused, per-level and 16-bit column collects the 128-bit
    loads of a round must go to DISTINCT destination registers.  At the 40-register budget of 3 x 512 threads per SM any
    extra live value makes ptxas re-use a destination, i.e. wait for one load before issuing the next: that cost the
    per-level collect 10 % and several rejected variants of the fused one 70 %."""
    import shutil
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    lib = os.path.join(PKG, 'libretinapost.so')
    txt = subprocess.check_output([cuobjdump, '-sass', lib], text=True, stderr=subprocess.STDOUT)
    checked = 0
    for fn in re.split(r'\n\s*Function : ', txt):
        name = fn.split('\n', 1)[0]
        m = re.search(r'collect_cols4_kernelILi4ELi3E|collect_cols4_levels_kernelILi4ELi3E|collect_cols8_half_kernelILi(\d)E', name)
        if not m:
            continue
        unroll = int(m.group(1)) if m.group(1) else 4
        dests = re.findall(r'LDG\.E\.128\.CONSTANT (R\d+),', fn)
        # the last `unroll` 128-bit loads of the function are the round's (threshold loads come first)
        assert len(dests) >= unroll, (name, dests)
        assert len(set(dests[-unroll:])) == unroll, (name, dests)
        checked += 1
    assert checked >= 4
