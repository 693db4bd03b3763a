"""The opt-in `tpu_semantics` of GlobalHardNMS / PerClassHardNMS (the reference's TPUStrategy branches,
postprocessing_ops.py:288-432) on the GPU: against the golden fixtures tests/golden/tpu_*.npz (made
by the unmodified reference branches over the numpy TensorFlow stand-in), against the oracle's tile-by-tile
restatement of tf.image.non_max_suppression_padded on random configurations, and at BASELINE sizes."""
import glob
import os

import numpy as np
import pytest

from _util import image_mismatches, make_params, oracle_detect, synth_inputs, to_numpy

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TPU = sorted(glob.glob(os.path.join(GOLDEN, 'tpu_*.npz')))


def _gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _check(got, g, box_exact):
    assert got['classes'].dtype == g['out_classes'].dtype == np.int32
    assert np.array_equal(got['valid_detections'], g['out_valid'])
    assert np.array_equal(got['classes'], g['out_classes'])
    assert np.array_equal(got['scores'], g['out_scores'])
    if box_exact:
        assert np.array_equal(got['boxes'], g['out_boxes'])
    else:
        np.testing.assert_allclose(got['boxes'], g['out_boxes'], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('path', TPU, ids=os.path.basename)
def test_fused_and_stagewise_vs_reference_tpu_branches(ref, path):
    from retinanet.model.builder import ModelBuilder
    from retinanet.model.layers import FilterTopKDetections, GenerateDetections
    g = np.load(path)
    H, C, M, k = int(g['H']), int(g['C']), int(g['M']), int(g['k'])
    mode, fpc = str(g['mode']), bool(g['filter_per_class'])
    p = make_params(H, num_classes=C, mode=mode, pre_nms_top_k=k, filter_per_class=fpc, max_detections=M,
                    tpu_semantics=True)
    x = {'class_logits': _gpu(g['logits']), 'encoded_boxes': _gpu(g['deltas'])}
    fused = ModelBuilder(p).add_post_processing_stage(None).layers[-1]
    _check(to_numpy(fused(x)), g, box_exact=False)
    stages = ModelBuilder(p).add_post_processing_stage(None, fused=False).layers[1:]
    y = x
    for layer in stages:
        y = layer(y)
    _check(to_numpy(y), g, box_exact=False)
    # GenerateDetections alone on IDENTICAL inputs (the oracle's decoded / filtered tensors): bit-exact
    H_, W_ = p.input.input_shape
    ap = p.anchor_params
    anchors, _ = ref.anchors(H_, W_, 3, 7, ap.areas, ap.aspect_ratios, ap.scales)
    z = {'scores': _gpu(ref.sigmoid(g['logits'])), 'boxes': _gpu(ref.decode_boxes(g['deltas'], anchors, H_, W_))}
    if k > 0:
        z = FilterTopKDetections(k, fpc)(z)
    det = GenerateDetections(0.5, 0.05, M, 0.5, C, mode, tpu_semantics=True)(z)
    _check(to_numpy(det), g, box_exact=True)


def test_known_answers_through_the_layer(ref):
    from retinanet.model.layers import GenerateDetections
    boxes = np.array([[[0.0, 0.0, 0.5, 0.5], [0.0, 0.0, 0.45, 0.45], [0.5, 0.5, 1.0, 1.0]]], np.float32)
    scores = np.array([[[0.9], [0.8], [0.7]]], np.float32)
    # padded slots gather (box 0, score 0.9): emitted three times (tests/test_oracle_tpu.py)
    out = to_numpy(GenerateDetections(0.5, 0.05, 4, None, 1, 'PerClassHardNMS', tpu_semantics=True)(
        {'scores': _gpu(scores), 'boxes': _gpu(boxes)}))
    exp = ref.generate_detections_tpu('PerClassHardNMS', scores, boxes, 0.5, 0.05, 4)
    assert image_mismatches(out, exp) == []
    assert out['scores'].tolist() == [[np.float32(0.9)] * 3 + [np.float32(0.7)]]
    # global branch: a real hard NMS across classes, -1 padding in every field, int32 classes
    boxes = np.array([[[0.0, 0.0, 0.5, 0.5], [0.0, 0.0, 0.45, 0.45], [0.5, 0.5, 1.0, 1.0], [0.2, 0.6, 0.4, 0.9]]],
                     np.float32)
    scores = np.array([[[0.9, 0.1], [0.2, 0.8], [0.3, 0.7], [0.01, 0.04]]], np.float32)
    out = to_numpy(GenerateDetections(0.5, 0.05, 4, None, 2, 'GlobalHardNMS', tpu_semantics=True)(
        {'scores': _gpu(scores), 'boxes': _gpu(boxes)}))
    assert out['valid_detections'].tolist() == [2] and out['classes'].dtype == np.int32
    assert out['classes'].tolist() == [[0, 1, -1, -1]] and (out['boxes'][0, 2:] == -1).all()
    # iou >= threshold suppresses (the NonMaxSuppressionV5 kernel of the non-TPU branch needs >)
    eq = np.array([[[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 1.0, 0.5]]], np.float32)
    sc = np.array([[[0.9], [0.8]]], np.float32)
    out = to_numpy(GenerateDetections(0.5, 0.05, 2, None, 1, 'GlobalHardNMS', tpu_semantics=True)(
        {'scores': _gpu(sc), 'boxes': _gpu(eq)}))
    assert out['valid_detections'].tolist() == [1]


def test_tpu_flag_rejects_other_modes_and_a_non_positive_threshold():
    from retinanet.model.layers import GenerateDetections
    from retinanet.model.layers.postprocessing_ops import _Handle
    rng = np.random.default_rng(3)
    s = _gpu(rng.uniform(0, 1, (2, 50, 3)).astype(np.float32))
    c = rng.uniform(0.1, 0.9, (2, 50, 2)).astype(np.float32)
    b = _gpu(np.concatenate([c - 0.1, c + 0.1], -1).astype(np.float32))
    for mode in ['CombinedNMS', 'GlobalSoftNMS', 'PerClassSoftNMS']:
        with pytest.raises(AssertionError):        # the reference's constructor under a TPUStrategy (:202-206)
            GenerateDetections(0.5, 0.05, 10, 0.5, 3, mode, tpu_semantics=True)
        with pytest.raises(AssertionError):        # and the C ABI (RPP_EMODE)
            _Handle(num_classes=3, mode=mode, soft_nms_sigma=0.5, tpu_semantics=True)
    with pytest.raises(ValueError):
        GenerateDetections(0.0, 0.05, 10, None, 3, 'GlobalHardNMS', tpu_semantics=True)({'scores': s, 'boxes': b})


def _case(seed):
    rng = np.random.default_rng(seed)
    mode = ['GlobalHardNMS', 'PerClassHardNMS'][rng.integers(2)]
    H = int(rng.choice([64, 96, 128, 192, 320, 320, 448]))
    W = int(rng.choice([H, H, 64, 160])) if H < 320 else H
    C = int(rng.choice([1, 2, 3, 4, 5, 8, 12, 20, 40])) if H < 448 else int(rng.choice([1, 4, 8, 12]))
    B = int(rng.integers(1, 5))
    M = int(rng.choice([1, 5, 20, 100, 150]))
    k = int(rng.choice([-1, 1, 10, 100, 1000, 5000]))
    fpc = bool(rng.integers(2)) and not mode.startswith('Global')
    inf = dict(mode=mode, pre_nms_top_k=k, filter_per_class=fpc, max_detections=M, tpu_semantics=True,
               iou_threshold=float(rng.choice([0.1, 0.3, 0.5, 0.75, 1.0])),
               score_threshold=float(rng.choice([0.0, 0.05, 0.3, 0.6])))
    dist = str(rng.choice(['dense', 'sparse', 'quantized', 'coarse', 'clustered']))
    return H, W, C, B, inf, dist, rng


@pytest.mark.parametrize('seed', range(int(os.environ.get('RPP_FUZZ_CASES_TPU', '60'))))
def test_fuzz_tpu_branches_vs_oracle(ref, seed):
    from retinanet.model.builder import ModelBuilder
    from test_gpu_fuzz import _inputs
    H, W, C, B, inf, dist, rng = _case(5000 + seed)
    p = make_params(H, W, num_classes=C, **inf)
    layer = ModelBuilder(p).add_post_processing_stage(None).layers[-1]
    N = layer.handle(C).num_anchors
    logits, deltas = _inputs(rng, B, N, C, dist)
    got = to_numpy(layer({'class_logits': torch.from_numpy(logits).cuda(),
                          'encoded_boxes': torch.from_numpy(deltas).cuda()}))
    exp = oracle_detect(ref, p, logits, deltas, threads=8)
    assert image_mismatches(got, exp) == [], (H, W, C, B, inf, dist)


@pytest.mark.parametrize('mode,k,fpc,dist', [('PerClassHardNMS', 5000, True, 'dense'),
                                            ('PerClassHardNMS', 5000, True, 'sparse'),
                                            ('PerClassHardNMS', -1, True, 'sparse'),
                                            ('GlobalHardNMS', 5000, False, 'dense'),
                                            ('GlobalHardNMS', -1, False, 'sparse')])
def test_baseline_size_tpu_branches(ref, mode, k, fpc, dist):
    # 640 x 640, 80 classes (BASELINE configs[1] geometry), 3 images
    from retinanet.model.builder import ModelBuilder
    p = make_params(640, num_classes=80, mode=mode, pre_nms_top_k=k, filter_per_class=fpc, tpu_semantics=True)
    layer = ModelBuilder(p).add_post_processing_stage(None).layers[-1]
    N = layer.handle(80).num_anchors
    logits, deltas = synth_inputs(3, N, 80, seed=11, dist=dist)
    got = to_numpy(layer({'class_logits': torch.from_numpy(logits).cuda(),
                          'encoded_boxes': torch.from_numpy(deltas).cuda()}))
    exp = oracle_detect(ref, p, logits, deltas, threads=8)
    assert image_mismatches(got, exp) == []
