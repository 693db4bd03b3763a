"""The serving call site of the path: `InferenceModule.run_inference` of the reference's export script
(retinanet/export.py:233-253), i.e. the re-keying of the post-processing outputs into the serving signature.

The reference freezes the model into a graph whose outputs come back as a LIST in sorted-key order and re-keys them
by POSITION:

    with NMS      sorted keys  boxes, classes, scores, valid_detections
                  -> {'boxes': [0], 'scores': [2], 'classes': [1], 'valid_detections': [3]}             (:244-248)
    skip_nms      (mode onnx_tensorrt: decode and NMS are left to the EfficientNMS_TRT plugin) the frozen outputs are
                  class_logits, encoded_boxes (sorted), and the reference labels them 'boxes' <- [0], 'scores' <- [1]
                  (:249-252) — i.e. 'boxes' holds the class logits and 'scores' the box deltas (B22);
                  onnx_utils._add_nms_plugin compensates by unpacking `class_logits, raw_boxes = outputs` (:28-32).

`InferenceModule` reproduces exactly that contract on top of a post-processing callable of this package (what
`ModelBuilder.add_post_processing_stage` / `prepare_model_for_export` return), so code written against the reference's
serving signature keeps working — including the mis-named keys of the skip_nms case.
"""


def frozen_outputs(outputs):
    """The output list of the frozen serving function: the dict's values in sorted-key order
    (convert_variables_to_constants_v2_as_graph flattens structured outputs that way, export.py:229-231)."""
    return [outputs[k] for k in sorted(outputs)]


class InferenceModule:
    """export.py:233-253.  `inference_function(**sample)` returns the dict of the post-processing stage (or of
    FuseDetections alone when `skip_nms`); `run_inference(sample)` returns the serving dict."""

    def __init__(self, inference_function, skip_nms):
        self.inference_function = inference_function
        self.skip_nms = skip_nms

    def run_inference(self, sample):
        raw_outputs = frozen_outputs(self.inference_function(**sample))
        outputs = {}
        if not self.skip_nms:
            outputs.update({
                'boxes': raw_outputs[0],
                'scores': raw_outputs[2],
                'classes': raw_outputs[1],
                'valid_detections': raw_outputs[3]})
        else:
            outputs.update({
                'boxes': raw_outputs[0],
                'scores': raw_outputs[1]})
        return outputs


def make_inference_module(post_processing_model, mode='tf'):
    """`post_processing_model`: the callable built by ModelBuilder.prepare_model_for_export(model, mode); its input is
    the dict of per-level head outputs, passed as sample = {'predictions': ...}.  NMS is skipped for onnx_tensorrt
    (export.py:255-259)."""
    return InferenceModule(lambda predictions: post_processing_model(predictions), skip_nms=(mode == 'onnx_tensorrt'))
