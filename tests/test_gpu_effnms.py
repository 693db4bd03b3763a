"""The EfficientNMS_TRT-compatible entry (reference: retinanet/onnx_utils.py:13-85) against the
oracle's restatement of TensorRT's efficientNMSPlugin.  Parity is UNPINNED for this row: the plugin is third-party
code absent from the reference tree and from this image; both sides follow the published algorithm."""
import numpy as np
import pytest

from _util import make_params, synth_inputs

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def _run(ref, H, W, C, B, M, score_thr, iou_thr, dist, seed):
    from retinanet.onnx_utils import EfficientNMSPlugin
    p = make_params(H, W, num_classes=C, max_detections=M, score_threshold=score_thr, iou_threshold=iou_thr)
    plugin = EfficientNMSPlugin(p)
    anchors = plugin.anchor_boxes
    N = anchors.shape[1]
    logits, deltas = synth_inputs(B, N, C, seed=seed, dist=dist)
    if dist == 'quantized':
        deltas = (deltas * 0.1).astype(np.float32)   # near-identical boxes: heavy same-class suppression
    valid, boxes, scores, classes = plugin(torch.from_numpy(deltas).cuda(), torch.from_numpy(logits).cuda(), anchors)
    assert valid.dtype == torch.int32 and tuple(valid.shape) == (B, 1)
    assert classes.dtype == torch.int32 and tuple(boxes.shape) == (B, M, 4) and tuple(scores.shape) == (B, M)
    ev, eb, es, ec = ref.efficient_nms(deltas, logits, anchors.cpu().numpy(), M, score_thr, iou_thr, threads=8)
    assert np.array_equal(valid.cpu().numpy(), ev)
    assert np.array_equal(classes.cpu().numpy(), ec)
    assert np.array_equal(scores.cpu().numpy(), es)
    np.testing.assert_allclose(boxes.cpu().numpy(), eb, rtol=1e-5, atol=1e-4)
    return ev


@pytest.mark.parametrize('H,W,C,B,M,dist', [(64, 64, 5, 2, 20, 'dense'), (64, 96, 3, 3, 100, 'sparse'),
                                          (128, 128, 8, 2, 100, 'quantized'), (320, 320, 12, 2, 100, 'dense'),
                                          (320, 320, 1, 2, 50, 'sparse'), (64, 64, 7, 1, 300, 'quantized'),
                                          (448, 448, 4, 2, 100, 'sparse')])
@pytest.mark.parametrize('score_thr,iou_thr', [(0.05, 0.5), (0.3, 0.3), (0.9, 0.75)])
def test_efficient_nms_vs_oracle(ref, H, W, C, B, M, dist, score_thr, iou_thr):
    _run(ref, H, W, C, B, M, score_thr, iou_thr, dist, seed=H + C + M)


def test_efficient_nms_baseline_size(ref):
    # 640 x 640, 80 classes: 6.1 M (anchor, class) pairs per image, the 4096-best truncation is active
    v = _run(ref, 640, 640, 80, 2, 100, 0.05, 0.5, 'dense', seed=5)
    assert (v == 100).all()
    _run(ref, 640, 640, 80, 2, 100, 0.05, 0.5, 'sparse', seed=6)


def test_onnx_tensorrt_export_pipeline(ref):
    """prepare_model_for_export(mode='onnx_tensorrt') stops after FuseDetections (builder.py:140-142) and
    onnx_utils._add_nms_plugin appends the NMS node (onnx_utils.py:13-85): node name, attributes, output order."""
    from retinanet import onnx_utils
    from retinanet.model.builder import ModelBuilder
    p = make_params(64, num_classes=4, max_detections=10)
    inference_model = ModelBuilder(p).prepare_model_for_export(None, mode='onnx_tensorrt')
    model = onnx_utils._add_nms_plugin(inference_model, p)
    assert model.plugin.op == 'EfficientNMS_TRT' and model.plugin.name == 'non_maximum_suppression'
    assert model.plugin.attributes == {'plugin_version': '1', 'background_class': -1, 'max_output_boxes': 10,
                                       'score_threshold': p.inference.score_threshold,
                                       'iou_threshold': p.inference.iou_threshold, 'score_activation': True,
                                       'box_coding': 1}
    assert model.plugin.output_names == ['valid_detections', 'detection_boxes', 'detection_scores',
                                         'detection_classes']
    rng = np.random.default_rng(0)
    heads = {'class-predictions': {}, 'box-predictions': {}}
    for level in range(3, 8):
        s = -(-64 // 2 ** level)
        heads['class-predictions'][str(level)] = torch.from_numpy(
            rng.standard_normal((2, s, s, 9 * 4)).astype(np.float32)).cuda()
        heads['box-predictions'][str(level)] = torch.from_numpy(
            (rng.standard_normal((2, s, s, 9 * 4)) * 0.3).astype(np.float32)).cuda()
    valid, boxes, scores, classes = model(heads)
    fused = inference_model(heads)
    ev, eb, es, ec = ref.efficient_nms(fused['encoded_boxes'].cpu().numpy(), fused['class_logits'].cpu().numpy(),
                                       model.plugin.anchor_boxes.cpu().numpy(), 10, p.inference.score_threshold,
                                       p.inference.iou_threshold)
    assert np.array_equal(valid.cpu().numpy(), ev) and np.array_equal(classes.cpu().numpy(), ec)
    assert np.array_equal(scores.cpu().numpy(), es)
    np.testing.assert_allclose(boxes.cpu().numpy(), eb, rtol=1e-5, atol=1e-4)
