"""TensorFlow's published unit-test vectors (see tests/test_tf_published_vectors.py) through the CUDA path, via the
reference's layer API.  The layers clip boxes to [0, 1] before NonMaxSuppressionV5 (postprocessing_ops.py:275, :501),
so the vectors' boxes (coordinates from -0.1 to 101) are translated by +0.125 and scaled by 1/128 (IoU is invariant
under both; the soft-NMS scores are compared at TF's own 1e-2 tolerance) — except for CombinedNMS, which takes them
as they are and clips on output like TF's kernel."""
import numpy as np
import pytest

from test_tf_published_vectors import BOXES, FLIPPED, SCORES

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

S = np.float32(1.0 / 128.0)
SHIFT = np.array([0, 0.125, 0, 0.125], np.float32)


def _run(mode, boxes, scores, M, iou=0.5, thr=0.0, sigma=None, **kw):
    from retinanet.model.layers import GenerateDetections
    layer = GenerateDetections(iou, thr, M, sigma, scores.shape[-1], mode, **kw)
    out = layer({'scores': torch.from_numpy(scores).cuda(), 'boxes': torch.from_numpy(boxes).cuda()})
    return {k: v.cpu().numpy() for k, v in out.items()}


def _indices(out_boxes, boxes):
    return [int(np.flatnonzero((boxes == b).all(1))[0]) for b in out_boxes]


@pytest.mark.parametrize('boxes', [BOXES, FLIPPED], ids=['as_published', 'flipped_coordinates'])
def test_hard_nms_select_from_three_clusters(boxes):
    b = ((boxes + SHIFT) * S).reshape(1, 6, 4)
    clipped = np.clip(b[0], 0, 1)
    assert np.array_equal(clipped, b[0])
    for M, exp in [(3, [3, 0, 5]), (2, [3, 0]), (30, [3, 0, 5])]:
        out = _run('PerClassHardNMS', b, SCORES.reshape(1, 6, 1), M)
        v = int(out['valid_detections'][0])
        assert v == len(exp) and _indices(out['boxes'][0, :v], clipped) == exp
        assert out['scores'][0, :v].tolist() == SCORES[exp].tolist()
    out = _run('PerClassHardNMS', b, SCORES.reshape(1, 6, 1), 3, thr=0.4)        # V3: score threshold
    assert out['valid_detections'].tolist() == [2] and _indices(out['boxes'][0, :2], clipped) == [3, 0]


def test_soft_nms_vector():
    # V5 soft: config sigma 1.0 -> NonMaxSuppressionV5(soft_nms_sigma=0.5) (postprocessing_ops.py:255)
    b = (BOXES + SHIFT) * S
    out = _run('GlobalSoftNMS', b.reshape(1, 6, 4), SCORES.reshape(1, 6, 1), 6, iou=0.5, thr=0.0, sigma=1.0)
    assert out['valid_detections'].tolist() == [6]
    assert _indices(out['boxes'][0], b) == [3, 0, 1, 5, 4, 2]
    np.testing.assert_allclose(out['scores'][0], [0.95, 0.9, 0.384, 0.3, 0.256, 0.197], rtol=1e-2, atol=1e-2)


def test_combined_nms_vector():
    out = _run('CombinedNMS', BOXES.reshape(1, 6, 4), SCORES.reshape(1, 6, 1), 3)
    assert out['valid_detections'].tolist() == [3]
    assert out['scores'].tolist() == [[np.float32(0.95), np.float32(0.9), np.float32(0.3)]]
    assert out['classes'].tolist() == [[0.0, 0.0, 0.0]]
    assert out['boxes'].tolist() == [[[0, 1, 1, 1], [0, 0, 1, 1], [0, 1, 1, 1]]]
    out = _run('CombinedNMS', BOXES.reshape(1, 6, 4), SCORES.reshape(1, 6, 1), 3, thr=0.4)
    assert out['valid_detections'].tolist() == [2]
    assert out['scores'].tolist() == [[np.float32(0.95), np.float32(0.9), 0.0]]
    assert out['boxes'].tolist() == [[[0, 1, 1, 1], [0, 0, 1, 1], [0, 0, 0, 0]]]


def test_padded_nms_vector():
    # tf.image.non_max_suppression_padded through the TPU branch of GlobalHardNMS: [3, 0, 5], num_valid 3
    b = ((BOXES + SHIFT) * S).reshape(1, 6, 4)
    out = _run('GlobalHardNMS', b, SCORES.reshape(1, 6, 1), 5, tpu_semantics=True)
    assert out['valid_detections'].tolist() == [3]
    assert _indices(out['boxes'][0, :3], np.clip(b[0], 0, 1)) == [3, 0, 5]
    assert (out['scores'][0, 3:] == -1).all() and (out['classes'][0, 3:] == -1).all()
