"""Global* modes behind the global filter (rpp_global.cuh: rows resolved from the sorted keys, block-parallel soft NMS)
against the CPU oracle: full-size trained-detector-like inputs, duplicate-heavy rows, ties, the k > 8192 fallback to
the generic problem kernel, thresholds that empty the queue, and flat axes that are not a multiple of four."""
import os
import sys

import numpy as np
import pytest

from _util import image_mismatches, make_params, oracle_detect, to_numpy
from conftest import ROOT

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')
sys.path.insert(0, ROOT)


def _run(ref, p, logits, deltas):
    from retinanet.model.layers import FusedPostProcessing
    layer = FusedPostProcessing(p)
    got = to_numpy(layer({'class_logits': torch.from_numpy(logits).cuda(),
                          'encoded_boxes': torch.from_numpy(deltas).cuda()}))
    exp = oracle_detect(ref, p, logits, deltas, threads=8)
    return got, exp


def _anchors(ref, p):
    H, W = p.input.input_shape
    ap = p.anchor_params
    return ref.anchors(H, W, 3, 7, ap.areas, ap.aspect_ratios, ap.scales)[0]


@pytest.mark.parametrize('mode', ['GlobalSoftNMS', 'GlobalHardNMS'])
@pytest.mark.parametrize('dist', ['clustered', 'dense', 'sparse'])
def test_full_size_640_c80(ref, mode, dist):
    """BASELINE configs[2] geometry, every image of a 6-image batch."""
    from tools import synth_inputs
    p = make_params(640, num_classes=80, mode=mode, pre_nms_top_k=5000, filter_per_class=False)
    anchors = torch.from_numpy(_anchors(ref, p))
    lg, dl = synth_inputs.make_inputs(dist, 6, anchors, 80, 640, 640, 'cpu', seed_logits=7, seed_deltas=8)
    got, exp = _run(ref, p, lg.numpy(), dl.numpy())
    assert image_mismatches(got, exp) == []


@pytest.mark.parametrize('mode', ['GlobalSoftNMS', 'GlobalHardNMS'])
@pytest.mark.parametrize('H,C,k', [(320, 5, 5000), (320, 5, 333), (128, 3, 1000), (96, 1, 100), (192, 2, 8192),
                                   (192, 7, 9000)])
def test_odd_flat_axes_duplicates_and_big_k(ref, mode, H, C, k):
    """N*C not a multiple of 4 (320^2 x 5 classes: 96 030), few classes (many duplicate anchors among the k best
    pairs), k at and beyond the soft kernel's 8192-row limit (generic fallback)."""
    p = make_params(H, num_classes=C, mode=mode, pre_nms_top_k=k, filter_per_class=False, max_detections=100)
    N = _anchors(ref, p).shape[0]
    rng = np.random.default_rng(H * 131 + C * 7 + k)
    logits = rng.standard_normal((3, N, C)).astype(np.float32)
    logits[:, :, 0] += 0.5 * logits[:, :, -1]          # correlated classes -> several classes of one anchor in the top k
    deltas = np.clip(rng.standard_normal((3, N, 4)) * 0.3, -4, 4).astype(np.float32)
    got, exp = _run(ref, p, logits, deltas)
    assert image_mismatches(got, exp) == []


@pytest.mark.parametrize('inf', [
    dict(score_threshold=0.6, soft_nms_sigma=0.1),      # decayed scores fall to the threshold: the queue drains
    dict(score_threshold=0.0, soft_nms_sigma=1.5, max_detections=300),
    dict(score_threshold=0.05, soft_nms_sigma=0.5, max_detections=1),
    dict(score_threshold=0.3, soft_nms_sigma=0.5, max_detections=1000, pre_nms_top_k=3000),
    dict(score_threshold=0.999, soft_nms_sigma=0.5),    # (almost) nothing above the threshold
])
def test_soft_parameter_edges(ref, inf):
    p = make_params(320, num_classes=12, mode='GlobalSoftNMS', filter_per_class=False,
                    **dict(dict(pre_nms_top_k=5000), **inf))
    N = _anchors(ref, p).shape[0]
    rng = np.random.default_rng(99)
    logits = (rng.standard_normal((4, N, 12)) * 1.5 - 1.0).astype(np.float32)
    deltas = np.clip(rng.standard_normal((4, N, 4)) * 0.2, -4, 4).astype(np.float32)   # small deltas: heavy overlap
    got, exp = _run(ref, p, logits, deltas)
    assert image_mismatches(got, exp) == []


def test_soft_heavy_overlap_and_ties(ref):
    """Every candidate overlaps every other (tiny deltas on few anchors' worth of boxes) and scores come from a
    coarse grid: long rounds (hundreds of visits before the next selection), exact score ties broken by row index."""
    p = make_params(128, num_classes=4, mode='GlobalSoftNMS', filter_per_class=False, pre_nms_top_k=5000,
                    max_detections=100, soft_nms_sigma=0.5, score_threshold=0.05)
    N = _anchors(ref, p).shape[0]
    rng = np.random.default_rng(5)
    logits = (np.round(rng.standard_normal((3, N, 4)) * 4) / 4).astype(np.float32)
    deltas = (rng.standard_normal((3, N, 4)) * 0.02).astype(np.float32)
    got, exp = _run(ref, p, logits, deltas)
    assert image_mismatches(got, exp) == []


def test_old_tf_soft_form_global(ref):
    """soft_ignores_iou_threshold = False (TF <= 2.2 kernel form): overlaps above the IoU threshold drop the box."""
    from retinanet.model.layers import FusedPostProcessing
    p = make_params(192, num_classes=6, mode='GlobalSoftNMS', filter_per_class=False, pre_nms_top_k=2000)
    N = _anchors(ref, p).shape[0]
    rng = np.random.default_rng(17)
    logits = rng.standard_normal((2, N, 6)).astype(np.float32)
    deltas = (rng.standard_normal((2, N, 4)) * 0.1).astype(np.float32)
    layer = FusedPostProcessing(p)
    h = layer.handle(6)
    # the flag is a handle property: rebuild the handle with it off
    from retinanet.model.layers.postprocessing_ops import _Handle
    inf = p.inference
    h2 = _Handle(H=192, W=192, min_level=3, max_level=7, num_classes=6, anchor_params=p.anchor_params,
                 mode=inf.mode, iou_threshold=inf.iou_threshold, score_threshold=inf.score_threshold,
                 soft_nms_sigma=inf.soft_nms_sigma, pre_nms_top_k=inf.pre_nms_top_k, filter_per_class=False,
                 max_detections=inf.max_detections, soft_ignores_iou_threshold=False)
    layer._handles[(6, torch.cuda.current_device())] = h2
    h.close()
    got = to_numpy(layer({'class_logits': torch.from_numpy(logits).cuda(),
                          'encoded_boxes': torch.from_numpy(deltas).cuda()}))
    exp = oracle_detect(ref, p, logits, deltas, threads=8, soft_ignores_iou_threshold=False)
    assert image_mismatches(got, exp) == []


def test_unaligned_flat_tensor(ref):
    """A logit tensor whose base is only 4-byte aligned (a slice of a larger buffer): the flat collect and sample
    kernels realign their 128-bit loads."""
    from retinanet.model.layers import FusedPostProcessing
    p = make_params(320, num_classes=5, mode='GlobalHardNMS', filter_per_class=False, pre_nms_top_k=5000)
    N = _anchors(ref, p).shape[0]
    rng = np.random.default_rng(3)
    logits = rng.standard_normal((2, N, 5)).astype(np.float32)
    deltas = np.clip(rng.standard_normal((2, N, 4)) * 0.5, -4, 4).astype(np.float32)
    exp = oracle_detect(ref, p, logits, deltas, threads=8)
    layer = FusedPostProcessing(p)
    for lead in (1, 2, 3):
        buf = torch.zeros(logits.size + 8, dtype=torch.float32, device='cuda')
        view = buf[lead:lead + logits.size].view(logits.shape)
        view.copy_(torch.from_numpy(logits))
        assert view.data_ptr() % 16 == 4 * lead
        got = to_numpy(layer({'class_logits': view, 'encoded_boxes': torch.from_numpy(deltas).cuda()}))
        assert image_mismatches(got, exp) == [], lead


@pytest.mark.parametrize('mode', ['PerClassHardNMS', 'CombinedNMS', 'PerClassSoftNMS'])
@pytest.mark.parametrize('H,C,B', [(320, 5, 4), (320, 91, 2), (448, 6, 3), (320, 3, 5), (192, 7, 2)])
def test_any_class_count_uses_vector_loads(ref, mode, H, C, B):
    """num_classes % 4 != 0 (5-class custom heads, 91-class COCO heads): collect_colsv_kernel — flat 128-bit words
    with a class-phase-preserving stride, image boundaries that are not 16-byte aligned, tiles that straddle images."""
    p = make_params(H, num_classes=C, mode=mode, pre_nms_top_k=5000, filter_per_class=True)
    N = _anchors(ref, p).shape[0]
    rng = np.random.default_rng(H + 13 * C + B)
    logits = rng.standard_normal((B, N, C)).astype(np.float32)
    deltas = np.clip(rng.standard_normal((B, N, 4)) * 0.5, -4, 4).astype(np.float32)
    got, exp = _run(ref, p, logits, deltas)
    assert image_mismatches(got, exp) == []


def test_many_classes_fall_back_to_the_generic_collect(ref):
    """num_classes = 512: the per-class stage of the vectorised collect would need more shared memory than an SM has
    (ADVICE r01): the generic kernel takes over instead of a failed launch."""
    p = make_params(64, num_classes=512, mode='PerClassHardNMS', pre_nms_top_k=200, filter_per_class=True)
    N = _anchors(ref, p).shape[0]
    rng = np.random.default_rng(512)
    logits = (rng.standard_normal((2, N, 512)) - 2.0).astype(np.float32)
    deltas = np.clip(rng.standard_normal((2, N, 4)) * 0.5, -4, 4).astype(np.float32)
    got, exp = _run(ref, p, logits, deltas)
    assert image_mismatches(got, exp) == []


def _level_heads(rng, H, C, B, dtype):
    cls, box = {}, {}
    for level in range(3, 8):
        f = int(np.ceil(H / 2 ** level))
        cls[str(level)] = torch.from_numpy(rng.standard_normal((B, f, f, 9 * C)).astype(np.float32)).to(dtype)
        box[str(level)] = torch.from_numpy(
            np.clip(rng.standard_normal((B, f, f, 36)) * 0.5, -4, 4).astype(np.float32)).to(dtype)
    return cls, box


@pytest.mark.parametrize('dtype', ['float32', 'bfloat16', 'float16'])
@pytest.mark.parametrize('mode,k,fpc,C', [
    ('GlobalSoftNMS', 3000, False, 8), ('GlobalHardNMS', 3000, False, 8), ('GlobalSoftNMS', -1, False, 8),
    ('GlobalHardNMS', -1, False, 5), ('PerClassHardNMS', 3000, False, 8), ('CombinedNMS', 3000, False, 5),
    ('GlobalSoftNMS', 5000, False, 5), ('GlobalHardNMS', 5000, False, 3)])
def test_every_mode_reads_head_outputs_in_place(ref, monkeypatch, dtype, mode, k, fpc, C):
    """SURVEY §8f-1 for the Global* modes and the global filter: per-level pieces and f16 / bf16 elements go through
    rpp_detect_typed (no torch.cat, no .to(float32)); the result equals the oracle on the values the reference sees
    after its tf.cast (:111-112), and the fused fp32 route."""
    from retinanet.model.builder import ModelBuilder
    from retinanet.model.layers import postprocessing_ops as ops
    H, B = 320, 3
    tdt = getattr(torch, dtype)
    p = make_params(H, num_classes=C, mode=mode, pre_nms_top_k=k, filter_per_class=fpc, max_detections=60)
    cls, box = _level_heads(np.random.default_rng(77 + C), H, C, B, tdt)
    heads = {'class-predictions': {k_: v.cuda() for k_, v in cls.items()},
             'box-predictions': {k_: v.cuda() for k_, v in box.items()}}
    model = ModelBuilder(p).add_post_processing_stage(None)
    # the native route must be taken: the eager fallbacks (cast / concat) are made to fail
    monkeypatch.setattr(ops, '_as_f32', lambda t: (_ for _ in ()).throw(AssertionError('eager cast / concat used')))
    got = to_numpy(model(heads))
    monkeypatch.undo()
    logits = np.concatenate([cls[str(l)].float().numpy().reshape(B, -1, C) for l in range(3, 8)], 1)
    deltas = np.concatenate([box[str(l)].float().numpy().reshape(B, -1, 4) for l in range(3, 8)], 1)
    exp = oracle_detect(ref, p, logits, deltas)
    assert image_mismatches(got, exp) == []


def _coarse(rng, shape):
    """a handful of distinct values: huge tie groups, saturated scores (sigmoid(20) == sigmoid(40) == 1.0f)"""
    return rng.choice(np.array([-30.0, -3.0, -0.5, 0.0, 0.5, 3.0, 20.0, 40.0], np.float32), size=shape)


@pytest.mark.parametrize('H,C,B,k,M', [(320, 5, 6, 5000, 100), (192, 3, 3, 1000, 100), (320, 80, 2, 5000, 100),
                                       (320, 5, 3, 5000, 1000), (128, 2, 2, 50, 100)])
@pytest.mark.parametrize('dist', ['dense', 'quantized', 'coarse', 'constant', 'near_ties'])
def test_global_hard_direct_kernel_vs_sorted_path_and_oracle(ref, monkeypatch, H, C, B, k, M, dist):
    """GlobalHardNMS behind the global filter: global_top_direct_kernel (raw-key selection, tie group of the k-th score
    resolved with the sigmoid itself) against the sorted-emission path (RPP_TOP_DIRECT=0) and the oracle — dense logits,
    heavy exact ties, saturated scores and constant logits (the direct kernel gives those images up: the kernels behind
    it take over), and logits a few ulps apart that round to ONE score around the k-th position."""
    from retinanet.model.layers import FusedPostProcessing
    p = make_params(H, num_classes=C, mode='GlobalHardNMS', pre_nms_top_k=k, filter_per_class=False, max_detections=M,
                    score_threshold=0.05)
    N = _anchors(ref, p).shape[0]
    rng = np.random.default_rng(H * 7 + C * 13 + k + len(dist))
    if dist == 'dense':
        logits = rng.standard_normal((B, N, C)).astype(np.float32)
        logits[:, :, 0] += 0.5 * logits[:, :, -1]
    elif dist == 'quantized':
        logits = (np.round(rng.standard_normal((B, N, C)) * 8) / 8).astype(np.float32)
    elif dist == 'coarse':
        logits = _coarse(rng, (B, N, C))
    elif dist == 'constant':
        logits = np.full((B, N, C), 1.25, np.float32)
    else:   # values spaced by single ulps around 3.0: groups of ~8 consecutive floats share one fp32 score
        base = np.float32(3.0).view(np.int32)
        logits = (base + rng.integers(-600, 600, size=(B, N, C)).astype(np.int32)).view(np.float32)
    deltas = np.clip(rng.standard_normal((B, N, 4)) * 0.3, -4, 4).astype(np.float32)
    x = {'class_logits': torch.from_numpy(logits).cuda(), 'encoded_boxes': torch.from_numpy(deltas).cuda()}
    monkeypatch.setenv('RPP_TOP_DIRECT', '1')
    got = to_numpy(FusedPostProcessing(p)(x))
    monkeypatch.setenv('RPP_TOP_DIRECT', '0')
    srt = to_numpy(FusedPostProcessing(p)(x))
    exp = oracle_detect(ref, p, logits, deltas, threads=8)
    assert image_mismatches(got, exp) == []
    assert image_mismatches(srt, exp) == []
    for key in ('scores', 'classes', 'valid_detections', 'boxes'):
        assert np.array_equal(got[key], srt[key]), key
