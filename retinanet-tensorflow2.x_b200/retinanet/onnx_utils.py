"""EfficientNMS stage of the `onnx_tensorrt` export (reference: retinanet/onnx_utils.py:13-85).

The reference exports the network up to FuseDetections (`skip_decoding=True, skip_nms=True`, model/builder.py:140-142),
converts it to ONNX and appends one `EfficientNMS_TRT` node that TensorRT executes (`_add_nms_plugin`, :13-85).
ONNX / TensorRT are out of scope here (and absent from the image); what this module mirrors is that node — inputs,
attributes and outputs exactly as `_add_nms_plugin` wires them — executed by libretinapost's `rpp_efficient_nms`, so
that an `onnx_tensorrt`-style pipeline keeps working without TensorRT.

Semantics are TensorRT's efficientNMSPlugin as published (include/retinapost.h has the statement): PARITY UNPINNED —
the plugin is third-party code that is neither in the reference tree nor installed here.
"""
import ctypes

import torch

from retinanet import _native
from retinanet.dataloader.anchor_generator import AnchorBoxGenerator
from retinanet.model.layers.postprocessing_ops import _Handle, _as_f32, _device_guard, _stream


def nms_plugin_attributes(params):
    """The attribute dict of onnx_utils.py:38-46."""
    inference_params = params.inference
    return {
        'plugin_version': '1',
        'background_class': -1,
        'max_output_boxes': inference_params['max_detections'],
        'score_threshold': inference_params['score_threshold'],
        'iou_threshold': inference_params['iou_threshold'],
        'score_activation': True,
        'box_coding': 1,
    }


class EfficientNMSPlugin:
    """`EfficientNMS_TRT(raw-boxes [B,N,4], class-logits [B,N,C], anchor-boxes [1,N,4]) ->
    (valid_detections [B,1] i32, detection_boxes [B,M,4], detection_scores [B,M], detection_classes [B,M] i32)`."""

    op = 'EfficientNMS_TRT'
    name = 'non_maximum_suppression'
    output_names = ['valid_detections', 'detection_boxes', 'detection_scores', 'detection_classes']

    def __init__(self, params):
        self._params = params
        self.attributes = nms_plugin_attributes(params)
        min_level = params.architecture.feature_fusion.min_level
        max_level = params.architecture.feature_fusion.max_level
        # the 'anchor-boxes' constant of onnx_utils.py:18-24
        self.anchor_boxes = AnchorBoxGenerator(*params.input.input_shape, min_level, max_level,
                                               params.anchor_params).boxes.unsqueeze(0)
        self._handles = {}

    def _handle(self, num_classes):
        key = (num_classes, torch.cuda.current_device())
        h = self._handles.get(key)
        if h is None:
            p = self._params
            h = _Handle(H=p.input.input_shape[0], W=p.input.input_shape[1],
                        min_level=p.architecture.feature_fusion.min_level,
                        max_level=p.architecture.feature_fusion.max_level,
                        num_classes=num_classes, anchor_params=p.anchor_params,
                        mode='PerClassHardNMS', iou_threshold=self.attributes['iou_threshold'],
                        score_threshold=self.attributes['score_threshold'],
                        max_detections=self.attributes['max_output_boxes'])
            self._handles[key] = h
        return h

    def __call__(self, raw_boxes, class_logits, anchor_boxes=None):
        with _device_guard(class_logits):
            return self._call(raw_boxes, class_logits, anchor_boxes)

    def _call(self, raw_boxes, class_logits, anchor_boxes=None):
        raw_boxes = _as_f32(raw_boxes)
        class_logits = _as_f32(class_logits)
        anchors = _as_f32(self.anchor_boxes if anchor_boxes is None else anchor_boxes).to(class_logits.device)
        B, N, C = class_logits.shape
        h = self._handle(C)
        if N != h.num_anchors or tuple(raw_boxes.shape) != (B, N, 4) or anchors.numel() != N * 4:
            raise ValueError('expected raw-boxes [B,{0},4], class-logits [B,{0},C] and anchor-boxes [1,{0},4]'
                             .format(h.num_anchors))
        M = h.max_detections
        dev = class_logits.device
        valid = torch.empty((B, 1), dtype=torch.int32, device=dev)
        boxes = torch.empty((B, M, 4), dtype=torch.float32, device=dev)
        scores = torch.empty((B, M), dtype=torch.float32, device=dev)
        classes = torch.empty((B, M), dtype=torch.int32, device=dev)
        ws = h.workspace(B, 0, dev)
        _native.check(_native.lib().rpp_efficient_nms(h.ptr, raw_boxes.data_ptr(), class_logits.data_ptr(),
                                                      anchors.data_ptr(), B, valid.data_ptr(), boxes.data_ptr(),
                                                      scores.data_ptr(), classes.data_ptr(), ws.data_ptr(),
                                                      ws.numel(), _stream()))
        return valid, boxes, scores, classes


def _add_nms_plugin(model, params):
    """onnx_utils.py:13-85: `model` maps images to {'class_logits', 'encoded_boxes'} (what
    ModelBuilder.prepare_model_for_export(model, 'onnx_tensorrt') returns); the result maps images to the plugin's
    four outputs, in the node's order."""
    plugin = EfficientNMSPlugin(params)

    def model_with_nms(x):
        y = model(x)
        return plugin(y['encoded_boxes'], y['class_logits'])

    model_with_nms.plugin = plugin
    return model_with_nms
