"""Known answers for the oracle's restatement of EfficientNMS_TRT as the reference configures it (onnx_utils.py:38-46;
SURVEY.md §8f-4).  Parity unpinned (TensorRT's plugin is absent): these pin the STATED algorithm.  CPU only."""
import numpy as np


def test_class_aware_suppression_and_output_layout(ref):
    # anchors 0 and 1 overlap (IoU 0.88), anchor 2 is far away; zero deltas -> boxes = anchors
    anchors = np.array([[10, 10, 8, 8], [10.5, 10, 8, 8], [30, 30, 8, 8]], np.float32)
    raw = np.zeros((1, 3, 4), np.float32)
    logits = np.array([[[2.0, -9.0], [1.0, 1.5], [0.5, -9.0]]], np.float32)
    valid, boxes, scores, classes = ref.efficient_nms(raw, logits, anchors, 4, 0.05, 0.5)
    # (a0,c0) 0.8808 kept; (a1,c1) 0.8176 kept (other class); (a1,c0) 0.7311 dropped by (a0,c0); (a2,c0) 0.6225 kept
    assert valid.tolist() == [[3]] and classes.tolist() == [[0, 1, 0, 0]]
    np.testing.assert_allclose(scores[0], [0.8807971, 0.8175745, 0.62245935, 0.0], rtol=1e-6)
    assert boxes[0].tolist() == [[10, 10, 8, 8], [10.5, 10, 8, 8], [30, 30, 8, 8], [0, 0, 0, 0]]   # centre-size, pixels
    # score filter is >= : a logit of exactly 0 passes a threshold of 0.5
    v2, _, s2, _ = ref.efficient_nms(raw, np.zeros((1, 3, 2), np.float32), anchors, 4, 0.5, 0.5)
    assert v2.tolist() == [[4]] and (s2 == 0.5).all()


def test_decode_and_truncation(ref):
    # centre-size decode: cx = dx * aw + ax, w = aw * exp(dw); no variance scaling, no normalisation
    anchors = np.array([[16, 16, 8, 4]], np.float32)
    raw = np.array([[[0.5, -0.25, np.log(2.0), 0.0]]], np.float32)
    _, boxes, _, _ = ref.efficient_nms(raw, np.array([[[3.0]]], np.float32), anchors, 1, 0.05, 0.5)
    np.testing.assert_allclose(boxes[0, 0], [20.0, 15.0, 16.0, 4.0], rtol=1e-6)
    # only the 4096 best (anchor, class) pairs of an image enter the NMS: 5000 disjoint boxes, M = 5000 -> 4096 kept
    n = 5000
    anchors = np.stack([np.arange(n) * 10.0 + 5, np.full(n, 5.0), np.full(n, 4.0), np.full(n, 4.0)], 1).astype(np.float32)
    logits = np.linspace(3.0, 1.0, n, dtype=np.float32).reshape(1, n, 1)
    valid, _, scores, _ = ref.efficient_nms(np.zeros((1, n, 4), np.float32), logits, anchors, n, 0.05, 0.5)
    assert valid.tolist() == [[4096]] and scores[0, 4095] > 0 and scores[0, 4096] == 0
