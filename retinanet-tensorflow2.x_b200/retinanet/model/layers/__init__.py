"""Public import surface of the reference's retinanet.model.layers for the post-processing path
(reference: retinanet/model/layers/__init__.py:3-14; the FPN helper layers are out of scope)."""
from retinanet.model.layers.postprocessing_ops import (
    FilterTopKDetections, FuseDetections, FusedPostProcessing, GenerateDetections,
    TransformBoxesAndScores)

__all__ = [
    'FilterTopKDetections',
    'FuseDetections',
    'FusedPostProcessing',
    'GenerateDetections',
    'TransformBoxesAndScores'
]
