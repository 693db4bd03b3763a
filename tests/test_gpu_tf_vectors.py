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


def test_topk_v2_vectors_through_the_filter_layer():
    """TopKV2's published vectors through FilterTopKDetections (per-class filter on a one-class tensor): values in
    tf.nn.top_k sorted order, ties to the lower index."""
    from retinanet.model.layers import FilterTopKDetections
    for vals, k, exp in [([3, 6, 15, 18, 6, 12, 1, 17, 3, 0, 4, 19, 1, 6], 3, [11, 3, 7]),
                         ([0.1, 0.3, 0.2, 0.4], 2, [3, 1]), ([0.1, 0.3, 0.4, 0.2], 2, [2, 1]),
                         ([5, 7, 7, 5, 7, 1], 4, [1, 2, 4, 0]), ([5, 7, 7, 5, 7, 1], 6, [1, 2, 4, 0, 3, 5])]:
        n = len(vals)
        scores = torch.tensor(vals, dtype=torch.float32).reshape(1, n, 1).cuda()
        boxes = torch.arange(n, dtype=torch.float32).reshape(1, n, 1).repeat(1, 1, 4).cuda()   # box = row index
        out = FilterTopKDetections(top_k=k, filter_per_class=True)({'scores': scores, 'boxes': boxes})
        assert out['boxes'][0, :, 0, 0].cpu().tolist() == [float(i) for i in exp]
        assert out['scores'][0, :, 0].cpu().tolist() == [float(np.float32(vals[i])) for i in exp]


def test_hard_modes_against_torchvision_batched_nms():
    """Independent implementation cross-check (not authoritative on exact ties): PerClassHardNMS at full 640x640 / 80
    classes against torchvision.ops.nms run class by class on torch-decoded boxes and torch.sigmoid scores."""
    tv = pytest.importorskip('torchvision')
    from _util import make_params
    from retinanet.dataloader.anchor_generator import AnchorBoxGenerator
    from retinanet.model.layers import FusedPostProcessing
    p = make_params(640, num_classes=80, mode='PerClassHardNMS', pre_nms_top_k=5000, filter_per_class=True)
    layer = FusedPostProcessing(p)
    N = layer.handle(80).num_anchors
    g = torch.Generator(device='cuda').manual_seed(11)
    logits = torch.randn((2, N, 80), generator=g, device='cuda') * 1.5 - 4.0
    deltas = (torch.randn((2, N, 4), generator=g, device='cuda') * 0.3).clamp_(-4, 4)
    out = layer({'class_logits': logits, 'encoded_boxes': deltas})
    a = AnchorBoxGenerator(640, 640, 3, 7, p.anchor_params).boxes
    xy = deltas[..., :2] * a[:, 2:] + a[:, :2]
    wh = torch.exp(deltas[..., 2:]) * a[:, 2:]
    boxes = (torch.cat([xy - wh / 2, xy + wh / 2], -1) / 640.0).clamp_(0, 1)
    scores = torch.sigmoid(logits)
    for b in range(2):
        cand_s, cand_b, cand_c = [], [], []
        for c in range(80):
            s, idx = scores[b, :, c].topk(5000)
            keep = s > 0.05
            s, idx = s[keep], idx[keep]
            k = tv.ops.nms(boxes[b, idx], s, 0.5)[:100]
            cand_s.append(s[k]); cand_b.append(boxes[b, idx[k]]); cand_c.append(torch.full_like(k, c))
        s, bx, cl = torch.cat(cand_s), torch.cat(cand_b), torch.cat(cand_c)
        top = s.argsort(descending=True, stable=True)[:100]
        assert int(out['valid_detections'][b]) == 100
        assert torch.equal(out['classes'][b].long(), cl[top])
        assert torch.allclose(out['scores'][b], s[top], rtol=1e-6, atol=0)
        assert torch.allclose(out['boxes'][b], bx[top], rtol=1e-5, atol=1e-6)
