"""The detection post-formatting half of the reference's COCOEvaluator (retinanet/eval/coco_evaluator.py:95-134):
`accumulate_results` turns a batch of padded detections into COCO result dicts.  The per-image numpy loop of the
reference (slice by valid_detections, rescale, int32 truncation, xyxy -> xywh, class-id remap) runs as one kernel
(rpp_coco_format) and one compact device->host copy.  `evaluate()` needs pycocotools, like the reference.
"""
import ctypes
import json

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
        key = (mode, num_classes, max_detections)
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
        boxes, scores, classes, valid = det['boxes'], det['scores'], det['classes'], det['valid_detections']
        B, M = scores.shape
        mode = {torch.float32: 'CombinedNMS', torch.int64: 'GlobalHardNMS', torch.int32: 'PerClassHardNMS'}[classes.dtype]
        num_classes = len(self._class_id_map) if self._class_id_map else 1
        h = self._handle(mode, num_classes, M)
        dev = boxes.device
        scale = None
        if rescale_detections:
            scale = torch.as_tensor(results['resize_scale'], dtype=torch.float32, device=dev).reshape(B, 2).contiguous()
        cmap = None
        if self._remap_class_ids and self._class_id_map:
            cmap = torch.tensor(self._class_id_map, dtype=torch.int32, device=dev)
        bbox = torch.empty((B * M, 4), dtype=torch.int32, device=dev)
        cat = torch.empty((B * M,), dtype=torch.int32, device=dev)
        sc = torch.empty((B * M,), dtype=torch.float32, device=dev)
        img = torch.empty((B * M,), dtype=torch.int32, device=dev)
        total = torch.zeros((1,), dtype=torch.int32, device=dev)
        _native.check(_native.lib().rpp_coco_format(
            h.ptr, boxes.contiguous().data_ptr(), scores.contiguous().data_ptr(), classes.contiguous().data_ptr(),
            valid.contiguous().data_ptr(), B, scale.data_ptr() if scale is not None else None,
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
