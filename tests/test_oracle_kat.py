"""Known-answer tests of the CPU oracle.  CPU only."""
import numpy as np
import pytest

from conftest import REFERENCE_CONFIG

AP = REFERENCE_CONFIG['anchor_params']


def test_anchor_kat_640(ref):
    # D.1
    a, bounds = ref.anchors(640, 640, 3, 7, AP['areas'], AP['aspect_ratios'], AP['scales'])
    assert bounds == [0, 57600, 72000, 75600, 76500, 76725]
    assert a.shape == (76725, 4)
    assert a[0].view(np.uint32).tolist() == [0x40800000, 0x40800000, 0x41B504F3, 0x423504F3]
    wh = [(22.627417, 45.254833), (28.508759, 57.017517), (35.918785, 71.83757), (32, 32),
          (40.317474, 40.317474), (50.796833, 50.796833), (45.254833, 22.627417), (57.017517, 28.508759),
          (71.83757, 35.918785)]
    np.testing.assert_allclose(a[:9, 2:], np.array(wh, np.float32), rtol=2e-7)
    assert np.all(a[:9, :2] == 4.0)
    assert a[9].tolist()[:2] == [12.0, 4.0]          # x advances before y
    assert a[9, 2:].tolist() == a[0, 2:].tolist()
    np.testing.assert_allclose(a[-3:], np.array([[576, 576, 724.07733, 362.03867], [576, 576, 912.2803, 456.14014],
                                                 [576, 576, 1149.4011, 574.70056]], np.float32), rtol=2e-7)
    s = a.astype(np.float64).sum(0)
    assert s[0] == s[1] == 24552000.0
    np.testing.assert_allclose(s[2:], 4765061.933136, rtol=1e-9)


@pytest.mark.parametrize('hw,n', [(1024, 196416), (320, 19206), (512, 49104), (768, 110484), (896, 150381),
                                  (1280, 306900)])
def test_anchor_counts(ref, hw, n):
    a, bounds = ref.anchors(hw, hw, 3, 7, AP['areas'], AP['aspect_ratios'], AP['scales'])
    assert a.shape[0] == n == bounds[-1]


def test_anchor_levels_3_6(ref):
    a, bounds = ref.anchors(448, 448, 3, 6, AP['areas'], AP['aspect_ratios'], AP['scales'])
    assert a.shape[0] == 37485


def test_decode_kat(ref):
    # D.2
    anc = np.array([[4, 4, 32, 32]], np.float32)
    out = ref.decode_boxes(np.zeros((1, 1, 4), np.float32), anc, 640, 640)
    np.testing.assert_array_equal(out[0, 0], np.array([-0.01875, -0.01875, 0.03125, 0.03125], np.float32))
    d = np.array([[[0.5, -0.25, np.log(2.0), 0.0]]], np.float32)
    out = ref.decode_boxes(d, anc, 640, 640)
    np.testing.assert_allclose(out[0, 0], np.array([-12, -20, 52, 12], np.float32) / 640, rtol=1e-6)


def test_decode_box_variance(ref):
    anc = np.array([[100, 50, 32, 64]], np.float32)
    d = np.array([[[1.0, -2.0, 0.5, 1.0]]], np.float32)
    a = ref.decode_boxes(d, anc, 640, 320, scale_box_targets=True)
    b = ref.decode_boxes(d * np.array([0.1, 0.1, 0.2, 0.2], np.float32), anc, 640, 320)
    assert a.tolist() == b.tolist()
    # normalisation quirk (B13): [x1,y1,x2,y2] / [H,W,H,W]
    cx, cy = 100 + 0.1 * 32, 50 - 0.2 * 64
    w, h = 32 * np.exp(0.1), 64 * np.exp(0.2)
    exp = np.array([(cx - w / 2) / 640, (cy - h / 2) / 320, (cx + w / 2) / 640, (cy + h / 2) / 320])
    np.testing.assert_allclose(a[0, 0], exp, rtol=1e-6)


def test_decode_encode_round_trip(ref):
    # LabelEncoder._compute_box_target (dataloader/label_encoder.py:57-76) is the inverse transform:
    # t_xy = (gt_xy - a_xy) / a_wh ; t_wh = log(gt_wh / a_wh)
    rng = np.random.default_rng(0)
    a, _ = ref.anchors(640, 640, 3, 7, AP['areas'], AP['aspect_ratios'], AP['scales'])
    a = a[rng.choice(len(a), 512, replace=False)]
    gt = np.stack([rng.uniform(0, 640, 512), rng.uniform(0, 640, 512), rng.uniform(8, 400, 512),
                   rng.uniform(8, 400, 512)], -1)
    t = np.concatenate([(gt[:, :2] - a[:, :2]) / a[:, 2:], np.log(gt[:, 2:] / a[:, 2:])], -1).astype(np.float32)
    out = ref.decode_boxes(t[None], a, 640, 640)[0] * 640
    exp = np.concatenate([gt[:, :2] - gt[:, 2:] / 2, gt[:, :2] + gt[:, 2:] / 2], -1)
    np.testing.assert_allclose(out, exp, rtol=2e-5, atol=2e-3)


def test_sigmoid(ref):
    x = np.array([-80, -4.59512, 0, 1, 20, 100], np.float32)
    y = ref.sigmoid(x)
    np.testing.assert_allclose(y, 1 / (1 + np.exp(-x.astype(np.float64))), rtol=1e-7)
    assert y[2] == 0.5 and y[-1] == 1.0


BOXES5 = np.array([[0, 0, 1, 1], [0, 0, 1, .9], [0, 0, 1, .5], [2, 2, 3, 3], [0, 0, 1, 1]], np.float32)
SCORES5 = np.array([.9, .8, .7, .6, .9], np.float32)


def test_iou_kat(ref):
    assert ref.iou(BOXES5[0], BOXES5[1]) == np.float32(0.9)
    assert ref.iou(BOXES5[0], BOXES5[2]) == 0.5
    np.testing.assert_allclose(ref.iou(BOXES5[1], BOXES5[2]), 0.5555556, rtol=1e-6)
    assert ref.iou(BOXES5[0], BOXES5[3]) == 0.0
    # flipped corners are canonicalised; zero-area boxes never overlap
    assert ref.iou(np.array([1, 1, 0, 0], np.float32), BOXES5[0]) == 1.0
    assert ref.iou(np.array([0, 0, 0, 1], np.float32), BOXES5[0]) == 0.0


def test_nms_v5_hard(ref):
    # D.3: strict '>' on IoU, index tie-break
    idx, sc, valid = ref.nms_v5(BOXES5, SCORES5, 10, 0.5, 0.05)
    assert valid == 3 and idx[:3].tolist() == [0, 2, 3]
    np.testing.assert_array_equal(sc[:3], np.array([.9, .7, .6], np.float32))
    assert idx[3:].tolist() == [0] * 7 and sc[3:].tolist() == [0.0] * 7


def test_nms_v5_iou_one_quirk(ref):
    # D.3: GlobalHardNMS passes iou_threshold=1.0 -> nothing suppressed
    idx, sc, valid = ref.nms_v5(BOXES5, SCORES5, 10, 1.0, 0.05)
    assert valid == 5 and idx[:5].tolist() == [0, 4, 1, 2, 3]


def test_nms_v5_soft(ref):
    # D.3 soft, sigma_tf = 0.25
    idx, sc, valid = ref.nms_v5(BOXES5, SCORES5, 10, 0.5, 0.05, soft_nms_sigma=0.25)
    assert valid == 4 and idx[:4].tolist() == [0, 3, 2, 1]
    np.testing.assert_allclose(sc[:4], [.9, .6, .42457145, .08539844], rtol=2e-6)
    # the pre-2.3 kernel form also hard-drops IoU > threshold: the TF-version discriminator
    idx, sc, valid = ref.nms_v5(BOXES5, SCORES5, 10, 0.5, 0.05, soft_nms_sigma=0.25,
                                soft_ignores_iou_threshold=False)
    assert valid == 3 and idx[:3].tolist() == [0, 3, 2]


def test_nms_v5_max_output(ref):
    idx, sc, valid = ref.nms_v5(BOXES5, SCORES5, 2, 1.0, 0.05)
    assert valid == 2 and idx.tolist() == [0, 4]


def test_nms_v5_score_threshold_strict(ref):
    idx, sc, valid = ref.nms_v5(BOXES5, SCORES5, 10, 1.0, 0.7)   # 0.7 is not > 0.7
    assert valid == 3 and idx[:3].tolist() == [0, 4, 1]


def test_topk_tiebreak(ref):
    # D.5
    v = np.array([[.5, .7, .7, .1]], np.float32)
    assert ref.topk(v, 2).tolist() == [[1, 2]]
    assert ref.topk(v, 1).tolist() == [[1]]
    assert ref.topk(v, 9).tolist() == [[1, 2, 0, 3]]
    # sorted=False: same SET, heap layout, slot 0 is the worst of the top-k
    u = ref.topk(v, 2, sorted=False)
    assert sorted(u[0].tolist()) == [1, 2]
    rng = np.random.default_rng(1)
    w = rng.standard_normal((7, 300)).astype(np.float32)
    s = ref.topk(w, 50, sorted=True)
    u = ref.topk(w, 50, sorted=False)
    for r in range(7):
        assert sorted(s[r].tolist()) == sorted(u[r].tolist())
        assert u[r, 0] == s[r, -1]
        assert s[r].tolist() == np.argsort(-w[r], kind='stable')[:50].tolist()


def test_padding_and_dtypes(ref):
    # D.4: one image, C=2, 3 anchors all below threshold
    scores = np.full((1, 3, 2), 0.01, np.float32)
    boxes = np.array([[[-0.2, 0.1, 0.5, 1.4], [0, 0, .3, .3], [.5, .5, .9, .9]]], np.float32)
    o = ref.generate_detections('CombinedNMS', scores, boxes, max_detections=4)
    assert o['classes'].dtype == np.float32 and o['valid_detections'].tolist() == [0]
    assert not o['boxes'].any() and not o['scores'].any() and not o['classes'].any()
    o = ref.generate_detections('GlobalHardNMS', scores, boxes, max_detections=4)
    assert o['classes'].dtype == np.int64 and o['valid_detections'].tolist() == [0]
    assert o['scores'].tolist() == [[-1.0] * 4] and o['classes'].tolist() == [[-1] * 4]
    np.testing.assert_array_equal(o['boxes'][0], np.tile(np.array([0, .1, .5, 1], np.float32), (4, 1)))
    o = ref.generate_detections('PerClassHardNMS', scores, boxes, max_detections=4)
    assert o['classes'].dtype == np.int32 and o['valid_detections'].tolist() == [0]
    assert o['scores'].tolist() == [[-1.0] * 4] and o['classes'].tolist() == [[-1] * 4]
    np.testing.assert_array_equal(o['boxes'][0], np.tile(np.array([0, .1, .5, 1], np.float32), (4, 1)))
    with pytest.raises(ValueError):
        ref.generate_detections('GlobalSoftNMS', scores, np.zeros((1, 3, 2, 4), np.float32))


def _rand_problem(rng, B, n, C, q):
    ctr = rng.uniform(0.1, 0.9, (B, n, q, 2))
    wh = rng.uniform(0.02, 0.4, (B, n, q, 2))
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], -1).astype(np.float32)
    scores = rng.uniform(0, 1, (B, n, C)).astype(np.float32) ** 3
    return scores, (boxes[:, :, 0] if q == 1 else boxes)


def test_per_class_hard_vs_torchvision(ref):
    # independent cross-check of the hard-NMS restatement (same greedy strict-'>' rule)
    import torch
    from torchvision.ops import nms
    rng = np.random.default_rng(2)
    scores, boxes = _rand_problem(rng, 2, 400, 3, 1)
    o = ref.generate_detections('PerClassHardNMS', scores, boxes, iou_threshold=0.5, score_threshold=0.05,
                                max_detections=50)
    bc = np.clip(boxes, 0, 1)
    for b in range(2):
        exp = []
        for c in range(3):
            m = np.nonzero(scores[b, :, c] > 0.05)[0]
            keep = nms(torch.from_numpy(bc[b, m]), torch.from_numpy(scores[b, m, c]), 0.5).numpy()[:50]
            exp += [(float(scores[b, m[i], c]), c) + tuple(bc[b, m[i]].tolist()) for i in keep]
        exp.sort(key=lambda t: -t[0])
        exp = exp[:50]
        v = int(o['valid_detections'][b])
        assert v == len(exp)
        np.testing.assert_array_equal(o['scores'][b, :v], np.array([e[0] for e in exp], np.float32))
        np.testing.assert_array_equal(o['classes'][b, :v], np.array([e[1] for e in exp], np.int32))
        np.testing.assert_array_equal(o['boxes'][b, :v], np.array([e[2:] for e in exp], np.float32))


def test_combined_matches_per_class_hard_when_no_clipping(ref):
    # with boxes inside [0,1] Combined and PerClassHard keep the same detections (different dtype/padding)
    rng = np.random.default_rng(3)
    scores, boxes = _rand_problem(rng, 3, 300, 4, 4)
    a = ref.generate_detections('CombinedNMS', scores, boxes, max_detections=40)
    p = ref.generate_detections('PerClassHardNMS', scores, boxes, max_detections=40)
    np.testing.assert_array_equal(a['valid_detections'], p['valid_detections'])
    for b in range(3):
        v = a['valid_detections'][b]
        np.testing.assert_array_equal(a['scores'][b, :v], p['scores'][b, :v])
        np.testing.assert_array_equal(a['classes'][b, :v].astype(np.int32), p['classes'][b, :v])
        np.testing.assert_array_equal(a['boxes'][b, :v], p['boxes'][b, :v])


def test_global_hard_is_topm_by_max_score(ref):
    rng = np.random.default_rng(4)
    scores, boxes = _rand_problem(rng, 2, 500, 5, 1)
    o = ref.generate_detections('GlobalHardNMS', scores, boxes, max_detections=30)
    for b in range(2):
        s = scores[b].max(-1)
        order = np.argsort(-s, kind='stable')[:30]
        np.testing.assert_array_equal(o['scores'][b], s[order])
        np.testing.assert_array_equal(o['classes'][b], scores[b].argmax(-1)[order])
        np.testing.assert_array_equal(o['boxes'][b], np.clip(boxes[b][order], 0, 1))


def test_detect_composition(ref):
    # detect == decode -> filter -> generate, for each filter flavour
    rng = np.random.default_rng(5)
    H = W = 64
    a, _ = ref.anchors(H, W, 3, 7, AP['areas'], AP['aspect_ratios'], AP['scales'])
    N, C, B = len(a), 6, 2
    logits = rng.standard_normal((B, N, C)).astype(np.float32)
    deltas = np.clip(rng.standard_normal((B, N, 4)) * 0.5, -4, 4).astype(np.float32)
    sc = ref.sigmoid(logits)
    bx = ref.decode_boxes(deltas, a, H, W)
    for mode, fpc, k in [('PerClassHardNMS', True, 50), ('CombinedNMS', True, 50), ('PerClassSoftNMS', False, 80),
                         ('GlobalSoftNMS', False, 80), ('GlobalHardNMS', False, -1), ('CombinedNMS', False, -1)]:
        o = ref.detect(logits, deltas, a, H, W, mode, pre_nms_top_k=k, filter_per_class=fpc, max_detections=20)
        if k > 0 and fpc:
            fs, fb, _ = ref.filter_per_class(sc, bx, k)
        elif k > 0:
            fs, fb, _ = ref.filter_global(sc, bx, k)
        else:
            fs, fb = sc, bx
        e = ref.generate_detections(mode, fs, fb, max_detections=20)
        for key in e:
            np.testing.assert_array_equal(o[key], e[key])
    with pytest.raises(ValueError):
        ref.detect(logits, deltas, a, H, W, 'GlobalSoftNMS', pre_nms_top_k=50, filter_per_class=True)
