"""The detection post-formatting half of the reference's COCOEvaluator (retinanet/eval/coco_evaluator.py:95-134):
`accumulate_results` turns a batch of padded detections into COCO result dicts.  The per-image numpy loop of the
reference (slice by valid_detections, rescale, int32 truncation, xyxy -> xywh, class-id remap) runs as one kernel
(rpp_coco_format) and one compact device->host copy.  `evaluate()` needs pycocotools, like the reference.
"""
import ctypes
import json

import numpy as np
import torch

from retinanet import _native


class COCOEvaluator:

    def __init__(self, input_shape, annotation_file_path=None, prediction_file_path='predictions.json',
                 remap_class_ids=False, class_id_map=None):
        """`class_id_map[i]` = original COCO category id of sorted class i.  The reference derives it from the
        annotation file (:36-57: categories sorted by name); pass either the file or the list."""
        self._input_shape = input_shape
        self.annotation_file_path = annotation_file_path
        self.prediction_file_path = prediction_file_path
        self._remap_class_ids = remap_class_ids
        self._processed_detections = []
        if class_id_map is None and annotation_file_path is not None and remap_class_ids:
            with open(annotation_file_path, 'r') as f:
                cats = json.load(f)['categories']
            by_name = {c['name']: c['id'] for c in cats}
            class_id_map = [by_name[name] for name in sorted(by_name)]
        self._class_id_map = list(class_id_map) if class_id_map is not None else None
        self._handles = {}

    def _handle(self, mode, num_classes, max_detections):
        from retinanet.model.layers.postprocessing_ops import _Handle
        key = (mode, num_classes, max_detections, torch.cuda.current_device())   # a handle belongs to its device
        if key not in self._handles:
            self._handles[key] = _Handle(H=self._input_shape[0], W=self._input_shape[1], num_classes=num_classes,
                                         mode=mode, max_detections=max_detections)
        return self._handles[key]

    @property
    def processed_detections(self):
        return self._processed_detections

    def accumulate_results(self, results, rescale_detections=True):
        image_ids = results['image_id']
        det = results['detections']
        # the reference takes numpy / TF host arrays here (:111-116); the kernel reads device memory, so everything is
        # brought to one CUDA device, made contiguous and given the dtypes the kernel expects before its raw pointers
        # are handed over (host tensors, e.g. the result of gather_detections(...).cpu(), are copied)
        def as_tensor(v):
            return v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
        boxes, scores, classes, valid = (as_tensor(det[k]) for k in ('boxes', 'scores', 'classes', 'valid_detections'))
        dev = next((t.device for t in (boxes, scores, classes, valid) if t.is_cuda),
                   torch.device('cuda', torch.cuda.current_device()))
        if classes.dtype not in (torch.float32, torch.int64, torch.int32):
            raise TypeError('classes must be float32 (CombinedNMS), int64 (Global*) or int32 (PerClass*), got {}'
                            .format(classes.dtype))
        boxes = boxes.to(device=dev, dtype=torch.float32).contiguous()
        scores = scores.to(device=dev, dtype=torch.float32).contiguous()
        classes = classes.to(device=dev).contiguous()
        valid = valid.to(device=dev, dtype=torch.int32).contiguous()
        if scores.dim() != 2 or tuple(boxes.shape) != tuple(scores.shape) + (4,) or classes.shape != scores.shape \
                or valid.shape != scores.shape[:1]:
            raise ValueError('detections must be boxes [B,M,4], scores [B,M], classes [B,M], valid_detections [B]; got '
                             '{}, {}, {}, {}'.format(tuple(boxes.shape), tuple(scores.shape), tuple(classes.shape),
                                                     tuple(valid.shape)))
        B, M = scores.shape
        if len(image_ids) != B:
            raise ValueError('{} image ids for {} images'.format(len(image_ids), B))
        mode = {torch.float32: 'CombinedNMS', torch.int64: 'GlobalHardNMS', torch.int32: 'PerClassHardNMS'}[classes.dtype]
        num_classes = len(self._class_id_map) if self._class_id_map else 1
        with torch.cuda.device(dev):
            return self._accumulate_on_device(results, image_ids, boxes, scores, classes, valid, B, M, mode, num_classes,
                                              dev, rescale_detections)

    def _accumulate_on_device(self, results, image_ids, boxes, scores, classes, valid, B, M, mode, num_classes, dev,
                              rescale_detections):
        h = self._handle(mode, num_classes, M)
        scale = None
        if rescale_detections:
            scale = torch.as_tensor(np.asarray(results['resize_scale']), dtype=torch.float32).to(dev)
            if scale.numel() != 2 * B:
                raise ValueError('resize_scale must hold [B,2] values')
            scale = scale.reshape(B, 2).contiguous()
        cmap = None
        if self._remap_class_ids and self._class_id_map:
            cmap = torch.tensor(self._class_id_map, dtype=torch.int32, device=dev)
        bbox = torch.empty((B * M, 4), dtype=torch.int32, device=dev)
        cat = torch.empty((B * M,), dtype=torch.int32, device=dev)
        sc = torch.empty((B * M,), dtype=torch.float32, device=dev)
        img = torch.empty((B * M,), dtype=torch.int32, device=dev)
        total = torch.zeros((1,), dtype=torch.int32, device=dev)
        _native.check(_native.lib().rpp_coco_format(
            h.ptr, boxes.data_ptr(), scores.data_ptr(), classes.data_ptr(), valid.data_ptr(), B,
            scale.data_ptr() if scale is not None else None,
            cmap.data_ptr() if cmap is not None else None, bbox.data_ptr(), cat.data_ptr(), sc.data_ptr(),
            img.data_ptr(), total.data_ptr(), torch.cuda.current_stream().cuda_stream))
        t = int(total.item())
        bbox, cat, sc, img = bbox[:t].cpu().tolist(), cat[:t].cpu().tolist(), sc[:t].cpu().tolist(), img[:t].cpu().tolist()
        ids = [int(i) for i in (image_ids.tolist() if hasattr(image_ids, 'tolist') else image_ids)]
        for r in range(t):
            self._processed_detections.append({'image_id': ids[img[r]], 'category_id': cat[r], 'bbox': bbox[r],
                                               'score': sc[r]})

    def evaluate(self):
        with open(self.prediction_file_path, 'w') as f:
            json.dump(self._processed_detections, f, indent=4)
        try:
            from pycocotools.coco import COCO
            from pycocotools.cocoeval import COCOeval
        except ImportError as e:
            raise ImportError('evaluate() needs pycocotools (as the reference does); the processed detections were '
                              'written to {}'.format(self.prediction_file_path)) from e
        gt = COCO(self.annotation_file_path)
        dt = gt.loadRes(self.prediction_file_path)
        ev = COCOeval(gt, dt, 'bbox')
        ev.evaluate()
        ev.accumulate()
        ev.summarize()
        return ev.stats
