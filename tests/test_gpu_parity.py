"""GPU parity: libretinapost (through the reference-shaped Python layers and the C ABI) against the CPU oracle on
identical seeded inputs.  Run on the B200 box: python -m pytest tests -m gpu."""
import numpy as np
import pytest

from _util import image_mismatches, make_params, oracle_detect, synth_inputs, to_numpy

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')


def _gpu(a):
    return torch.from_numpy(a).cuda()


def _fused(params):
    from retinanet.model.builder import ModelBuilder
    return ModelBuilder(params, run_mode='export').add_post_processing_stage(None).layers[-1]


def test_anchors_bit_exact(ref):
    from retinanet.dataloader.anchor_generator import AnchorBoxGenerator
    for hw, levels in [(640, (3, 7)), (1024, (3, 7)), (320, (3, 7)), (448, (3, 6)), (100, (3, 7))]:
        p = make_params(hw)
        g = AnchorBoxGenerator(hw, hw, levels[0], levels[1], p.anchor_params)
        exp, bounds = ref.anchors(hw, hw, levels[0], levels[1], p.anchor_params.areas,
                                  p.anchor_params.aspect_ratios, p.anchor_params.scales)
        assert g.anchor_boundaries == bounds
        got = g.boxes.cpu().numpy()
        assert got.shape == exp.shape
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


def test_decode_stage(ref):
    from retinanet.model.layers import TransformBoxesAndScores
    p = make_params(320, num_classes=8)
    layer = TransformBoxesAndScores(p)
    anchors, _ = ref.anchors(320, 320, 3, 7, p.anchor_params.areas, p.anchor_params.aspect_ratios,
                             p.anchor_params.scales)
    logits, deltas = synth_inputs(2, len(anchors), 8, seed=1)
    logits[0, :8, 0] = [-120, -90, -30, 0, 17, 30, 90, 120]
    out = to_numpy(layer({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}))
    es = ref.sigmoid(logits)
    eb = ref.decode_boxes(deltas, anchors, 320, 320)
    np.testing.assert_allclose(out['scores'], es, rtol=1e-5, atol=1e-37)   # tolerance stated by north_star
    np.testing.assert_allclose(out['boxes'], eb, rtol=1e-5, atol=1e-6)
    # in practice both sides round the exact value once: report how close to bit-exact we are
    assert (out['scores'] == es).mean() > 0.9999
    assert (out['boxes'] == eb).mean() > 0.9999


CASES = [
    # H, C, B, mode, k, per_class, dist
    (64, 6, 3, 'PerClassHardNMS', 50, True, 'dense'),
    (64, 6, 3, 'CombinedNMS', 50, True, 'dense'),
    (64, 6, 3, 'PerClassHardNMS', -1, True, 'dense'),
    (64, 6, 3, 'CombinedNMS', -1, True, 'sparse'),
    (128, 4, 2, 'PerClassHardNMS', 5000, True, 'quantized'),
    (320, 8, 2, 'PerClassHardNMS', 300, True, 'dense'),      # N = 19206: sampled pre-threshold path
    (320, 8, 2, 'CombinedNMS', 5000, True, 'sparse'),
    (320, 12, 2, 'PerClassHardNMS', 5000, True, 'quantized'),
    (320, 5, 2, 'PerClassHardNMS', 1000, True, 'dense'),     # C % 4 != 0: scalar collect kernel
    # soft NMS (lazy re-scoring, bit-exact expf)
    (64, 6, 3, 'PerClassSoftNMS', 50, True, 'dense'),
    (64, 6, 3, 'PerClassSoftNMS', -1, True, 'sparse'),
    (320, 8, 2, 'PerClassSoftNMS', 5000, True, 'dense'),
    (320, 8, 2, 'PerClassSoftNMS', 300, True, 'quantized'),
    # global filter (top-k over anchors x classes) feeding the per-class modes
    (64, 6, 3, 'PerClassHardNMS', 80, False, 'dense'),
    (320, 8, 2, 'CombinedNMS', 5000, False, 'dense'),
    (320, 8, 2, 'PerClassSoftNMS', 2000, False, 'sparse'),
    # Global* modes: no filter, or the global filter
    (64, 6, 3, 'GlobalHardNMS', -1, True, 'dense'),
    (64, 6, 3, 'GlobalSoftNMS', -1, False, 'dense'),
    (64, 6, 3, 'GlobalSoftNMS', 80, False, 'dense'),
    (320, 8, 2, 'GlobalHardNMS', 5000, False, 'dense'),
    (320, 8, 2, 'GlobalSoftNMS', 5000, False, 'dense'),
    (320, 8, 2, 'GlobalSoftNMS', -1, False, 'sparse'),
    (320, 5, 2, 'GlobalHardNMS', 5000, False, 'sparse'),
    (320, 8, 2, 'GlobalSoftNMS', 5000, False, 'quantized'),
]


@pytest.mark.parametrize('H,C,B,mode,k,fpc,dist', CASES)
def test_fused_detect_vs_oracle(ref, H, C, B, mode, k, fpc, dist):
    p = make_params(H, num_classes=C, mode=mode, pre_nms_top_k=k, filter_per_class=fpc, max_detections=100)
    layer = _fused(p)
    N = layer.handle(C).num_anchors
    logits, deltas = synth_inputs(B, N, C, seed=H + C, dist=dist)
    got = to_numpy(layer({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}))
    exp = oracle_detect(ref, p, logits, deltas)
    assert image_mismatches(got, exp) == []


def test_global_mode_with_per_class_filter_is_rejected():
    # the reference fails with a rank error; we raise ValueError
    p = make_params(64, num_classes=4, mode='GlobalSoftNMS', pre_nms_top_k=50, filter_per_class=True)
    layer = _fused(p)
    N = layer.handle(4).num_anchors
    logits, deltas = synth_inputs(1, N, 4)
    with pytest.raises(ValueError):
        layer({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)})


def test_unsupported_mode_raises_assertion():
    from retinanet.model.layers import GenerateDetections
    with pytest.raises(AssertionError):
        GenerateDetections(mode='FancyNMS')


STAGE_CASES = [
    ('PerClassHardNMS', 300, True), ('CombinedNMS', 300, True), ('PerClassSoftNMS', 300, True),
    ('PerClassHardNMS', 400, False), ('GlobalSoftNMS', 400, False), ('GlobalHardNMS', 400, False),
    ('CombinedNMS', -1, True), ('GlobalSoftNMS', -1, False),
]


@pytest.mark.parametrize('mode,k,fpc', STAGE_CASES)
def test_stagewise_layers_vs_oracle(ref, mode, k, fpc):
    """The reference's layer-by-layer graph (fused=False): every stage output is checked against the oracle's
    stage; NMS runs on IDENTICAL inputs (the oracle's own decoded tensors) and must be bit-exact."""
    from retinanet.model.builder import ModelBuilder
    from retinanet.model.layers import FilterTopKDetections, GenerateDetections
    H, C, B = 320, 8, 2
    p = make_params(H, num_classes=C, mode=mode, pre_nms_top_k=k, filter_per_class=fpc, max_detections=60)
    model = ModelBuilder(p, run_mode='export').add_post_processing_stage(None, fused=False)
    anchors, _ = ref.anchors(H, H, 3, 7, p.anchor_params.areas, p.anchor_params.aspect_ratios,
                             p.anchor_params.scales)
    N = len(anchors)
    logits, deltas = synth_inputs(B, N, C, seed=11)
    # whole chain through the layers
    x = {'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}
    for layer in model.layers[1:]:
        x = layer(x)
    exp = oracle_detect(ref, p, logits, deltas)
    assert image_mismatches(to_numpy(x), exp) == []
    # stage 2 on identical inputs: oracle's scores/boxes in, bit-exact out
    es, eb = ref.sigmoid(logits), ref.decode_boxes(deltas, anchors, H, H)
    if k > 0:
        got = to_numpy(FilterTopKDetections(k, fpc)({'scores': _gpu(es), 'boxes': _gpu(eb)}))
        fs, fb, _ = (ref.filter_per_class if fpc else ref.filter_global)(es, eb, k)
        assert np.array_equal(got['scores'], fs) and np.array_equal(got['boxes'], fb)
    else:
        fs, fb = es, eb
    inf = p.inference
    gd = GenerateDetections(inf.iou_threshold, inf.score_threshold, inf.max_detections, inf.soft_nms_sigma, C, mode)
    got = to_numpy(gd({'scores': _gpu(fs), 'boxes': _gpu(fb)}))
    e2 = ref.generate_detections(mode, fs, fb, max_detections=inf.max_detections)
    for key in e2:
        assert got[key].dtype == e2[key].dtype
        assert np.array_equal(got[key], e2[key]), key


@pytest.mark.parametrize('mode', ['PerClassHardNMS', 'CombinedNMS'])
def test_exact_scan_path(ref, mode):
    """The slow path (exact radix select straight over the column) must give the same answer as the lists."""
    from retinanet import _native
    p = make_params(128, num_classes=4, mode=mode, pre_nms_top_k=500, max_detections=50)
    layer = _fused(p)
    h = layer.handle(4)
    _native.check(_native.lib().rpp_debug_force_exact_scan(h.ptr, 1))
    logits, deltas = synth_inputs(2, h.num_anchors, 4, seed=9)
    got = to_numpy(layer({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}))
    exp = oracle_detect(ref, p, logits, deltas)
    assert image_mismatches(got, exp) == []


@pytest.mark.parametrize('mode', ['PerClassHardNMS', 'CombinedNMS'])
@pytest.mark.parametrize('value', [0.0, -10.0, 25.0])
def test_constant_logits(ref, mode, value):
    """All-equal logits: every score ties, order is pure index order; -10 -> nothing above the threshold (the
    reference's padding conventions); 25 -> saturated scores of exactly 1.0."""
    p = make_params(320, num_classes=4, mode=mode, pre_nms_top_k=5000, max_detections=100)
    layer = _fused(p)
    N = layer.handle(4).num_anchors
    logits = np.full((2, N, 4), value, np.float32)
    _, deltas = synth_inputs(2, N, 4, seed=3)
    got = to_numpy(layer({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}))
    exp = oracle_detect(ref, p, logits, deltas)
    assert image_mismatches(got, exp) == []


def test_full_size_640_c80(ref):
    """BASELINE config 2 shapes (640x640, 80 classes, PerClassHardNMS, k=5000, M=100) on a small batch."""
    p = make_params(640, num_classes=80)
    layer = _fused(p)
    N = layer.handle(80).num_anchors
    assert N == 76725
    for dist in ('dense', 'sparse'):
        logits, deltas = synth_inputs(4, N, 80, seed=42, dist=dist)
        got = to_numpy(layer({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}))
        exp = oracle_detect(ref, p, logits, deltas)
        assert image_mismatches(got, exp) == [], dist
