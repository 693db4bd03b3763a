"""The oracle's restatement of the TPUStrategy branches of GenerateDetections (postprocessing_ops.py:288-432, the
product's opt-in `tpu_semantics`): tf.image.non_max_suppression_padded restated tile by tile in C++ against (a) an
independent greedy scan in numpy, (b) hand-derived known answers, (c) the golden fixtures tests/golden/tpu_*.npz made
by running the unmodified reference branches over the numpy TensorFlow stand-in.  CPU only."""
import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
AP = dict(areas=[1024.0, 4096.0, 16384.0, 65536.0, 262144.0], aspect_ratios=[0.5, 1.0, 2.0],
          scales=[1, 1.2599210498948732, 1.5874010519681994])
TPU = sorted(glob.glob(os.path.join(GOLDEN, 'tpu_*.npz')))


def greedy(ref, boxes, scores, M, thr, sthr):
    cand = [i for i in range(len(scores)) if sthr is None or scores[i] > sthr]
    cand.sort(key=lambda i: (-scores[i], i))
    kept = []
    for i in cand:
        if len(kept) >= M:
            break
        if not (boxes[i] > 0).any():
            continue
        if all(ref.iou_padded(boxes[j], boxes[i]) < thr for j in kept):
            kept.append(i)
    return kept


def random_boxes(rng, n, spread, size):
    c = rng.uniform(0.5 - spread, 0.5 + spread, (n, 2)).astype(np.float32)
    wh = rng.uniform(size / 4, size, (n, 2)).astype(np.float32)
    return np.clip(np.concatenate([c - wh / 2, c + wh / 2], 1), 0, 1).astype(np.float32)


def test_fixture_inventory():
    assert len(TPU) == 24


@pytest.mark.parametrize('n,M,spread,size', [(30, 10, 0.5, 0.3), (700, 100, 0.5, 0.3), (1500, 100, 0.5, 0.2),
                                             (513, 600, 0.5, 0.1), (2000, 100, 0.15, 0.3), (3000, 200, 0.2, 0.25),
                                             (1200, 64, 0.05, 0.4)])
@pytest.mark.parametrize('sthr', [None, 0.3])
def test_tiled_iteration_equals_greedy(ref, n, M, spread, size, sthr):
    # spread/size: from scattered (little suppression, the first tile fills M) to one tight cluster (cross-tile
    # suppression over several 512-box tiles and fewer than M survivors)
    rng = np.random.default_rng(n * 7 + M)
    boxes = random_boxes(rng, n, spread, size)
    boxes[rng.integers(0, n, 5)] = 0                      # all-zero boxes are never selected
    scores = rng.uniform(0, 1, n).astype(np.float32)
    scores[rng.integers(0, n, 20)] = scores[0]            # ties -> lower index first
    idx, valid = ref.nms_padded(boxes, scores, M, 0.5, sthr)
    idx_tf, valid_tf = ref.nms_padded(boxes, scores, M, 0.5, sthr, exact_fixed_point=False)
    kept = greedy(ref, boxes, scores, M, 0.5, sthr)
    assert valid == len(kept) and idx[:valid].tolist() == kept
    assert (idx[valid:] == 0).all()
    assert valid_tf == valid and np.array_equal(idx_tf, idx)


def test_padded_iou_known_answers(ref):
    a = [0.0, 0.0, 1.0, 1.0]
    assert ref.iou_padded(a, a) == np.float32(1.0) / (np.float32(1.0) + np.float32(1e-8))
    # inter 0.5, union 1: 0.5 / (1 + 1e-8) rounds to 0.5 in fp32 -> suppressed at threshold 0.5 (>=), while the
    # NonMaxSuppressionV5 kernel (strict >) keeps it
    b = [0.0, 0.0, 1.0, 0.5]
    assert ref.iou_padded(a, b) == np.float32(0.5)
    idx, valid = ref.nms_padded(np.array([a, b], np.float32), np.array([0.9, 0.8], np.float32), 2, 0.5)
    assert valid == 1 and idx.tolist() == [0, 0]
    i5, _, v5 = ref.nms_v5(np.array([a, b], np.float32), np.array([0.9, 0.8], np.float32), 2, 0.5, 0.0)
    assert v5 == 2 and i5.tolist() == [0, 1]
    # no canonicalisation: a flipped box has a negative "area" term and never reaches the threshold
    assert ref.iou_padded(a, [1.0, 1.0, 0.0, 0.0]) <= 0.0
    # zero-area box: IoU 0 with everything, but it is selected when a coordinate is positive
    z = [0.5, 0.2, 0.5, 0.7]
    idx, valid = ref.nms_padded(np.array([a, z], np.float32), np.array([0.9, 0.8], np.float32), 2, 0.5)
    assert valid == 2 and idx.tolist() == [0, 1]


def test_global_branch_known_answer(ref):
    # 4 boxes, 2 classes: box 1 overlaps box 0 (IoU 0.81 >= 0.5) and is dropped although its class differs (global
    # NMS on the row maximum); box 3 scores below the threshold.  Beyond `valid` every field is -1; classes int32.
    boxes = np.array([[[0.0, 0.0, 0.5, 0.5], [0.0, 0.0, 0.45, 0.45], [0.5, 0.5, 1.0, 1.0], [0.2, 0.6, 0.4, 0.9]]],
                     np.float32)
    scores = np.array([[[0.9, 0.1], [0.2, 0.8], [0.3, 0.7], [0.01, 0.04]]], np.float32)
    out = ref.generate_detections_tpu('GlobalHardNMS', scores, boxes, 0.5, 0.05, 4)
    assert out['valid_detections'].tolist() == [2]
    assert out['classes'].dtype == np.int32 and out['classes'].tolist() == [[0, 1, -1, -1]]
    assert out['scores'].tolist() == [[np.float32(0.9), np.float32(0.7), -1.0, -1.0]]
    assert out['boxes'][0, :2].tolist() == boxes[0, [0, 2]].tolist() and (out['boxes'][0, 2:] == -1).all()
    # the non-TPU GlobalHardNMS passes IoU threshold 1.0: nothing is suppressed there
    plain = ref.generate_detections('GlobalHardNMS', scores, boxes, 0.5, 0.05, 4)
    assert plain['valid_detections'].tolist() == [3]


def test_per_class_branch_pads_gather_row0(ref):
    # one class, 3 boxes, M = 4: boxes 0 and 1 overlap (1 is dropped), box 2 survives -> 2 kept; the two padded slots
    # gather index 0 = (box 0, score 0.9), which is above the score threshold: the reference emits it THREE times
    boxes = np.array([[[0.0, 0.0, 0.5, 0.5], [0.0, 0.0, 0.45, 0.45], [0.5, 0.5, 1.0, 1.0]]], np.float32)
    scores = np.array([[[0.9], [0.8], [0.7]]], np.float32)
    out = ref.generate_detections_tpu('PerClassHardNMS', scores, boxes, 0.5, 0.05, 4)
    assert out['valid_detections'].tolist() == [4]
    assert out['scores'].tolist() == [[np.float32(0.9)] * 3 + [np.float32(0.7)]]
    assert out['classes'].tolist() == [[0, 0, 0, 0]]
    assert out['boxes'][0, :3].tolist() == [boxes[0, 0].tolist()] * 3
    # with row 0 at or below the threshold the padded slots are masked to -1 like every sub-threshold position
    scores2 = np.array([[[0.04], [0.8], [0.7]]], np.float32)
    out2 = ref.generate_detections_tpu('PerClassHardNMS', scores2, boxes, 0.5, 0.05, 4)
    assert out2['valid_detections'].tolist() == [2]
    assert out2['scores'].tolist() == [[np.float32(0.8), np.float32(0.7), -1.0, -1.0]]
    assert out2['classes'].tolist() == [[0, 0, -1, -1]] and (out2['boxes'][0, 2:] == -1).all()
    # sub-threshold boxes take part in the NMS (no score filter inside, :323-330): they fill slots, so nothing is
    # padded, and are masked afterwards
    boxes3 = np.array([[[0.0, 0.0, 0.3, 0.3], [0.5, 0.5, 0.8, 0.8], [0.0, 0.6, 0.3, 0.9], [0.6, 0.0, 0.9, 0.3]]],
                      np.float32)
    scores3 = np.array([[[0.9], [0.03], [0.02], [0.01]]], np.float32)
    out3 = ref.generate_detections_tpu('PerClassHardNMS', scores3, boxes3, 0.5, 0.05, 4)
    assert out3['valid_detections'].tolist() == [1]
    assert out3['scores'].tolist() == [[np.float32(0.9), -1.0, -1.0, -1.0]]


@pytest.mark.parametrize('path', TPU, ids=os.path.basename)
def test_tpu_chain_vs_reference(ref, path):
    g = np.load(path)
    H, W, C, M, k = int(g['H']), int(g['W']), int(g['C']), int(g['M']), int(g['k'])
    mode, fpc = str(g['mode']), bool(g['filter_per_class'])
    anchors, _ = ref.anchors(H, W, 3, 7, AP['areas'], AP['aspect_ratios'], AP['scales'])
    scores = ref.sigmoid(g['logits'])
    boxes = ref.decode_boxes(g['deltas'], anchors, H, W)
    fs, fb = scores, boxes
    if k > 0:
        fs, fb, _ = (ref.filter_per_class if fpc else ref.filter_global)(scores, boxes, k)
    out = ref.generate_detections_tpu(mode, fs, fb, max_detections=M)
    full = ref.detect_tpu(g['logits'], g['deltas'], anchors, H, W, mode, pre_nms_top_k=k, filter_per_class=fpc,
                          max_detections=M)
    for o in (out, full):
        assert o['classes'].dtype == g['out_classes'].dtype == np.int32
        assert np.array_equal(o['valid_detections'], g['out_valid'])
        assert np.array_equal(o['classes'], g['out_classes'])
        assert np.array_equal(o['scores'], g['out_scores'])
        assert np.array_equal(o['boxes'], g['out_boxes'])


def test_tpu_fixtures_cover_the_padded_slot_quirk():
    # at least one per-class fixture must contain a duplicated detection (padded slots gathering row 0)
    dup = 0
    for path in TPU:
        g = np.load(path)
        if str(g['mode']) != 'PerClassHardNMS':
            continue
        for b in range(g['out_boxes'].shape[0]):
            v = int(g['out_valid'][b])
            rows = {tuple(r) for r in np.concatenate([g['out_boxes'][b, :v], g['out_scores'][b, :v, None],
                                                      g['out_classes'][b, :v, None].astype(np.float32)], 1).tolist()}
            dup += v - len(rows)
    assert dup > 0
