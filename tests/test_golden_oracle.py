"""The CPU oracle against the golden fixtures produced by executing the UNMODIFIED reference modules over the numpy
TensorFlow stand-in (tests/golden/make_golden.py): pins the oracle's restatement of the reference's Python glue
(anchor op order, decode, filters, mode dispatch, clipping, padding, dtypes).  CPU only."""
import glob
import hashlib
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
AP = dict(areas=[1024.0, 4096.0, 16384.0, 65536.0, 262144.0], aspect_ratios=[0.5, 1.0, 2.0],
          scales=[1, 1.2599210498948732, 1.5874010519681994])
DETECT = sorted(glob.glob(os.path.join(GOLDEN, 'detect_*.npz')))
ANCHORS = sorted(glob.glob(os.path.join(GOLDEN, 'anchors_*.npz')))


def test_fixture_inventory():
    assert len(DETECT) == 39 and len(ANCHORS) == 4


@pytest.mark.parametrize('path', ANCHORS, ids=os.path.basename)
def test_anchors_vs_reference(ref, path):
    g = np.load(path)
    boxes, bounds = ref.anchors(int(g['H']), int(g['W']), int(g['min_level']), int(g['max_level']), AP['areas'],
                                AP['aspect_ratios'], AP['scales'])
    assert bounds == g['boundaries'].tolist() and len(boxes) == int(g['n'])
    assert hashlib.sha256(boxes.tobytes()).hexdigest() == str(g['sha256'])
    keep = boxes if len(boxes) <= 4096 else np.concatenate([boxes[:512], boxes[-512:]])
    assert np.array_equal(keep.view(np.uint32), g['rows'].view(np.uint32))


@pytest.mark.parametrize('path', DETECT, ids=os.path.basename)
def test_chain_vs_reference(ref, path):
    g = np.load(path)
    H, W, C, M, k = int(g['H']), int(g['W']), int(g['C']), int(g['M']), int(g['k'])
    mode, fpc, sbt = str(g['mode']), bool(g['filter_per_class']), bool(g['scale_box_targets'])
    anchors, _ = ref.anchors(H, W, 3, 7, AP['areas'], AP['aspect_ratios'], AP['scales'])
    scores = ref.sigmoid(g['logits'])
    boxes = ref.decode_boxes(g['deltas'], anchors, H, W, scale_box_targets=sbt)
    assert np.array_equal(scores, g['scores'])
    assert np.array_equal(boxes, g['boxes'])
    fs, fb = scores, boxes
    if k > 0:
        fs, fb, _ = (ref.filter_per_class if fpc else ref.filter_global)(scores, boxes, k)
        assert np.array_equal(fs, g['filtered_scores']) and np.array_equal(fb, g['filtered_boxes'])
    out = ref.generate_detections(mode, fs, fb, max_detections=M)
    full = ref.detect(g['logits'], g['deltas'], anchors, H, W, mode, pre_nms_top_k=k, filter_per_class=fpc,
                      max_detections=M, scale_box_targets=sbt)
    for o in (out, full):
        assert o['classes'].dtype == g['out_classes'].dtype
        assert np.array_equal(o['valid_detections'], g['out_valid'])
        assert np.array_equal(o['classes'], g['out_classes'])
        assert np.array_equal(o['scores'], g['out_scores'])
        assert np.array_equal(o['boxes'], g['out_boxes'])
