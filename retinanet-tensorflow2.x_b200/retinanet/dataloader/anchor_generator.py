"""AnchorBoxGenerator with the reference's constructor and properties (dataloader/anchor_generator.py:5-112).

The boxes are produced on the device by libretinapost's anchor kernel (one thread per anchor, the reference's fp32
operation order); `boxes` is a torch CUDA tensor [N, 4] = [cx, cy, w, h], `anchor_boundaries` a Python list.
"""
import ctypes

import torch

from retinanet import _native


def _get(params, key):
    return params[key] if isinstance(params, dict) else getattr(params, key)


class AnchorBoxGenerator:

    def __init__(self, img_h, img_w, min_level, max_level, params):
        self.image_height = img_h
        self.image_width = img_w

        self.areas = _get(params, 'areas')
        self.aspect_ratios = _get(params, 'aspect_ratios')
        self.scales = _get(params, 'scales')

        self._num_anchors = len(self.aspect_ratios) * len(self.scales)
        self._min_level = min_level
        self._max_level = max_level
        self._strides = [2**i for i in range(min_level, max_level + 1)]

        from retinanet.model.layers.postprocessing_ops import _Handle
        h = _Handle(H=img_h, W=img_w, min_level=min_level, max_level=max_level, num_classes=1,
                    anchor_params={'areas': self.areas, 'aspect_ratios': self.aspect_ratios, 'scales': self.scales})
        L = _native.lib()
        n = L.rpp_num_anchors(h.ptr)
        bounds = (ctypes.c_long * (L.rpp_num_levels(h.ptr) + 1))()
        _native.check(L.rpp_anchor_boundaries(h.ptr, bounds))
        self._anchor_boundaries = [int(b) for b in bounds]
        boxes = torch.empty((n, 4), dtype=torch.float32, device='cuda')
        _native.check(L.rpp_anchors(h.ptr, boxes.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.current_stream().synchronize()
        self._boxes = boxes
        h.close()

    @property
    def anchor_boundaries(self):
        return self._anchor_boundaries

    @property
    def boxes(self):
        return self._boxes
