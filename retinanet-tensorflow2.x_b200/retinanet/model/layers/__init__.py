"""Public import surface of the reference's retinanet.model.layers for the post-processing path
(reference: retinanet/model/layers/__init__.py:3-14; the FPN helper layers are out of scope).

    from retinanet.model.layers import FuseDetections, TransformBoxesAndScores, FilterTopKDetections, GenerateDetections

works as it does in the reference; FusedPostProcessing (the whole chain as one rpp_detect call) is this package's own."""
from retinanet.model.layers import postprocessing_ops as _ops

_REFERENCE_NAMES = ('FuseDetections', 'TransformBoxesAndScores', 'FilterTopKDetections', 'GenerateDetections')
_OWN_NAMES = ('FusedPostProcessing', 'FilterTopKDetectionsPerLevel')

__all__ = sorted(_REFERENCE_NAMES + _OWN_NAMES)
globals().update({name: getattr(_ops, name) for name in __all__})
