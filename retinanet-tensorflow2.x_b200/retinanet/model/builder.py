"""ModelBuilder.prepare_model_for_export / add_post_processing_stage (reference: retinanet/model/builder.py:121-190).

Only the post-processing composition is mirrored: backbones, necks, heads, losses and optimizers are out of scope.
`model` is any callable returning the head-output dict
    {'class-predictions': {level: [B,H,W,A*C]}, 'box-predictions': {level: [B,H,W,A*4]}}
(torch CUDA tensors), or None when the caller already holds that dict.
"""
import json
import logging

from retinanet.model.layers import (FilterTopKDetections, FuseDetections, FusedPostProcessing, GenerateDetections,
                                    TransformBoxesAndScores)


class InferenceModel:
    """What add_post_processing_stage returns: `inference_model(x)` runs model -> post-processing stages."""

    def __init__(self, model, stages, fused):
        self.model = model
        self.layers = stages
        self.fused = fused

    def __call__(self, x, training=False):
        if self.model is not None:
            x = self.model(x)
        for layer in self.layers:
            x = layer(x)
        return x

    call = __call__


class ModelBuilder:

    def __init__(self, params, run_mode='export'):
        self.params = params
        self.run_mode = run_mode

    def prepare_model_for_export(self, model, mode='tf'):
        skip_decoding = False
        skip_nms = False

        if mode == 'tf':
            pass

        elif mode == 'tf_tensorrt' or mode == 'onnx':
            if self.params.inference.pre_nms_top_k > 0:
                logging.warning('Inference is faster with top-k filtering disabled '
                                'when running on Tensorrt/ONNX. Forcefully '
                                'disabling top-k filtering !!!')
                self.params.inference.pre_nms_top_k = -1

        elif mode == 'onnx_tensorrt':
            skip_decoding = True
            skip_nms = True

        else:
            raise ValueError('Invalid export model requested!')

        return self.add_post_processing_stage(model=model, skip_decoding=skip_decoding, skip_nms=skip_nms)

    def add_post_processing_stage(self, model, skip_decoding=False, skip_nms=False, fused=True):
        """`fused=True` (default) collapses decode -> top-k -> NMS into one rpp_detect call when the whole chain is
        requested; `fused=False` keeps the reference's layer-by-layer graph (same results, intermediates in HBM)."""
        params = self.params
        logging.info('Postprocessing stage config:\n{}'.format(json.dumps(params.inference, indent=4)))

        stages = [FuseDetections(min_level=params.architecture.feature_fusion.min_level,
                                 max_level=params.architecture.feature_fusion.max_level)]

        if fused and not skip_decoding and not skip_nms:
            stages[0].lazy = True     # hand the per-level tensors over; the concat happens only if it is needed
            stages.append(FusedPostProcessing(params=params))
            return InferenceModel(model, stages, fused=True)

        if not skip_decoding:
            stages.append(TransformBoxesAndScores(params=params))
        else:
            logging.warning('Skipping decoding of predictions !!!')

        if params.inference.pre_nms_top_k > 0 and not skip_nms:
            stages.append(FilterTopKDetections(top_k=params.inference.pre_nms_top_k,
                                               filter_per_class=params.inference.filter_per_class))
        else:
            logging.warning('Skipping top-k anchors filtering !!!')

        if not skip_nms:
            stages.append(GenerateDetections(iou_threshold=params.inference.iou_threshold,
                                             score_threshold=params.inference.score_threshold,
                                             max_detections=params.inference.max_detections,
                                             soft_nms_sigma=params.inference.soft_nms_sigma,
                                             num_classes=params.architecture.head.num_classes,
                                             mode=params.inference.mode,
                                             # optional key, absent from the reference's JSONs (the reference detects
                                             # a TPUStrategy instead, postprocessing_ops.py:199-208)
                                             tpu_semantics=bool(params.inference.get('tpu_semantics', False))))
        else:
            logging.warning('Skipping NMS filtering !!!')

        return InferenceModel(model, stages, fused=False)
