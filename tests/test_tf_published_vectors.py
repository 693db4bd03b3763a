"""TensorFlow's OWN published unit-test vectors for the kernels the reference calls, run against the oracle (CPU) and,
in test_gpu_tf_vectors.py, against the CUDA path.

The reference's arithmetic lives in TensorFlow (tf-nightly 2.8.0-dev20210925), which is neither vendored in
/root/reference nor installable here, and the reference itself has no tests.  The closest thing to golden vectors at
that boundary are the expectations TensorFlow's test-suite holds for these kernels; they are quoted here from the
TensorFlow repository (tensorflow/core/kernels/image/non_max_suppression_op_test.cc: NonMaxSuppressionOpTest /
V2 / V3 / V4 / V5 and CombinedNonMaxSuppressionOpTest; tensorflow/python/ops/image_ops_test.py:
NonMaxSuppressionWithScoresTest, NonMaxSuppressionPaddedTest).  They pin: the strict `>` IoU test, the coordinate
canonicalisation, the score-threshold filter, zero-padding of V4/V5 outputs, the soft-NMS decay and re-queueing
order (and with it the `is_soft ||` weight form), CombinedNMS's clipping and output order.
TensorFlow's sources are not available offline: the vectors are restated from the upstream test files, not copied
from a checkout; cases marked "derived" extend a published vector and are not claimed to exist upstream.  CPU only."""
import numpy as np

# the six boxes of "SelectFromThreeClusters" and their scores
BOXES = np.array([[0, 0, 1, 1], [0, 0.1, 1, 1.1], [0, -0.1, 1, 0.9], [0, 10, 1, 11], [0, 10.1, 1, 11.1],
                  [0, 100, 1, 101]], np.float32)
FLIPPED = np.array([[1, 1, 0, 0], [0, 0.1, 1, 1.1], [0, 0.9, 1, -0.1], [0, 10, 1, 11], [1, 10.1, 0, 11.1],
                    [1, 101, 0, 100]], np.float32)
SCORES = np.array([0.9, 0.75, 0.6, 0.95, 0.5, 0.3], np.float32)
NO_THRESHOLD = float('-inf')


def test_select_from_three_clusters(ref):
    idx, sc, valid = ref.nms_v5(BOXES, SCORES, 3, 0.5, NO_THRESHOLD)
    assert valid == 3 and idx.tolist() == [3, 0, 5]
    assert sc.tolist() == [np.float32(0.95), np.float32(0.9), np.float32(0.3)]      # V5: selected_scores


def test_select_from_three_clusters_flipped_coordinates(ref):
    idx, _, valid = ref.nms_v5(FLIPPED, SCORES, 3, 0.5, NO_THRESHOLD)
    assert valid == 3 and idx.tolist() == [3, 0, 5]


def test_select_at_most_two_and_at_most_thirty(ref):
    idx, _, valid = ref.nms_v5(BOXES, SCORES, 2, 0.5, NO_THRESHOLD)
    assert valid == 2 and idx.tolist() == [3, 0]
    idx, _, valid = ref.nms_v5(BOXES, SCORES, 30, 0.5, NO_THRESHOLD)
    assert valid == 3 and idx[:3].tolist() == [3, 0, 5] and (idx[3:] == 0).all()


def test_select_with_negative_scores(ref):
    idx, _, valid = ref.nms_v5(BOXES, SCORES - np.float32(10.0), 6, 0.5, NO_THRESHOLD)
    assert valid == 3 and idx[:3].tolist() == [3, 0, 5]


def test_select_single_box_ten_identical_boxes_and_empty_input(ref):
    idx, _, valid = ref.nms_v5(BOXES[:1], SCORES[:1], 3, 0.5, NO_THRESHOLD)
    assert valid == 1 and idx[0] == 0
    idx, _, valid = ref.nms_v5(np.tile(BOXES[:1], (10, 1)), np.full(10, 0.9, np.float32), 3, 0.5, NO_THRESHOLD)
    assert valid == 1 and idx[0] == 0
    _, _, valid = ref.nms_v5(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 30, 0.5, NO_THRESHOLD)
    assert valid == 0


def test_v3_score_threshold_and_v4_padding(ref):
    # NonMaxSuppressionV3OpTest.TestSelectFromThreeClustersWithScoreThreshold: score_threshold 0.4 drops box 5
    idx, _, valid = ref.nms_v5(BOXES, SCORES, 3, 0.5, 0.4)
    assert valid == 2 and idx[:2].tolist() == [3, 0]
    # NonMaxSuppressionV4OpTest.TestSelectFromThreeClustersPadFive / PadFiveScoreThr: zero padding + valid_outputs
    idx, _, valid = ref.nms_v5(BOXES, SCORES, 5, 0.5, 0.0)
    assert valid == 3 and idx.tolist() == [3, 0, 5, 0, 0]
    idx, _, valid = ref.nms_v5(BOXES, SCORES, 5, 0.5, 0.4)
    assert valid == 2 and idx.tolist() == [3, 0, 0, 0, 0]


def test_v5_soft_nms(ref):
    # NonMaxSuppressionV5OpTest.TestSelectFromThreeClustersWithSoftNMS / image_ops_test
    # testSelectFromThreeClustersWithSoftNMS: sigma 0.5, score_threshold 0, max_output_size 6.
    for iou_threshold in (1.0, 0.5):   # with soft_nms_sigma > 0 the IoU threshold does not hard-suppress (TF >= 2.3)
        idx, sc, valid = ref.nms_v5(BOXES, SCORES, 6, iou_threshold, 0.0, soft_nms_sigma=0.5)
        assert valid == 6 and idx.tolist() == [3, 0, 1, 5, 4, 2]
        np.testing.assert_allclose(sc, [0.95, 0.9, 0.384, 0.3, 0.256, 0.197], rtol=1e-2, atol=1e-2)
    # the pre-2.3 weight form (flag soft_ignores_iou_threshold=False) fails that vector at threshold 0.5:
    idx, _, valid = ref.nms_v5(BOXES, SCORES, 6, 0.5, 0.0, soft_nms_sigma=0.5, soft_ignores_iou_threshold=False)
    assert idx[:valid].tolist() != [3, 0, 1, 5, 4, 2]


def test_combined_nms_select_from_three_clusters(ref):
    # CombinedNonMaxSuppressionOpTest.TestSelectFromThreeClusters: boxes [1,6,1,4], scores [1,6,1],
    # max_output_size_per_class 3, max_total_size 3, iou 0.5, score_threshold 0, clip_boxes (default true)
    out = ref.generate_detections('CombinedNMS', SCORES.reshape(1, 6, 1), BOXES.reshape(1, 6, 4), 0.5, 0.0, 3)
    assert out['valid_detections'].tolist() == [3]
    assert out['scores'].tolist() == [[np.float32(0.95), np.float32(0.9), np.float32(0.3)]]
    assert out['classes'].tolist() == [[0.0, 0.0, 0.0]]
    assert out['boxes'].tolist() == [[[0, 1, 1, 1], [0, 0, 1, 1], [0, 1, 1, 1]]]       # clipped to [0, 1]
    # ...WithScoreThreshold: 0.4 -> two detections, third slot zero-padded
    out = ref.generate_detections('CombinedNMS', SCORES.reshape(1, 6, 1), BOXES.reshape(1, 6, 4), 0.5, 0.4, 3)
    assert out['valid_detections'].tolist() == [2]
    assert out['scores'].tolist() == [[np.float32(0.95), np.float32(0.9), 0.0]]
    assert out['boxes'].tolist() == [[[0, 1, 1, 1], [0, 0, 1, 1], [0, 0, 0, 0]]]


def test_combined_nms_two_classes_derived(ref):
    # derived: the same six boxes with two score columns — per-class NMS, merged by score, class ids as floats
    scores = np.array([[[0.1, 0.9], [0.75, 0.8], [0.6, 0.3], [0.95, 0.1], [0.5, 0.5], [0.3, 0.1]]], np.float32)
    out = ref.generate_detections('CombinedNMS', scores, BOXES.reshape(1, 6, 4), 0.5, 0.0, 3)
    assert out['valid_detections'].tolist() == [3]
    assert out['scores'].tolist() == [[np.float32(0.95), np.float32(0.9), np.float32(0.75)]]
    assert out['classes'].tolist() == [[0.0, 1.0, 0.0]]
    assert out['boxes'].tolist() == [[[0, 1, 1, 1], [0, 0, 1, 1], [0, float(np.float32(0.1)), 1, 1]]]


def test_padded_nms_select_from_three_clusters(ref):
    # image_ops_test NonMaxSuppressionPaddedTest.testSelectFromThreeClusters: padded to max_output_size 5
    idx, valid = ref.nms_padded(BOXES, SCORES, 5, 0.5)
    assert valid == 3 and idx.tolist() == [3, 0, 5, 0, 0]
    # derived: with a score threshold of 0.4 box 5 is filtered (scores and boxes zeroed before the sort)
    idx, valid = ref.nms_padded(BOXES, SCORES, 3, 0.5, 0.4)
    assert valid == 2 and idx.tolist() == [3, 0, 0]


def test_topk_v2_vectors(ref):
    # topk_op_test.py TopKTest.testTop3Vector / testTop2 (and the tf.math.top_k docstring example); ties go to the
    # lower index (testTensorStableSort / "stable" in the op's documentation: derived vector below)
    v = np.array([[3, 6, 15, 18, 6, 12, 1, 17, 3, 0, 4, 19, 1, 6]], np.float32)
    assert ref.topk(v, 3).tolist() == [[11, 3, 7]]
    v = np.array([[0.1, 0.3, 0.2, 0.4], [0.1, 0.3, 0.4, 0.2]], np.float32)
    assert ref.topk(v, 2).tolist() == [[3, 1], [2, 1]]
    v = np.array([[5, 7, 7, 5, 7, 1]], np.float32)           # derived: ties -> lower index first
    assert ref.topk(v, 4).tolist() == [[1, 2, 4, 0]]
    assert ref.topk(v, 6).tolist() == [[1, 2, 4, 0, 3, 5]]   # k == num_cols: fully sorted
