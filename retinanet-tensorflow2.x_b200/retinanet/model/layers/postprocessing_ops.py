"""The four post-processing layers of the reference (retinanet/model/layers/postprocessing_ops.py), same names,
constructor signatures, dict keys, output dtypes and padding, running on hand-written sm_100a kernels.

    FuseDetections           postprocessing_ops.py:7-56
    TransformBoxesAndScores  postprocessing_ops.py:59-117
    FilterTopKDetections     postprocessing_ops.py:120-173
    GenerateDetections       postprocessing_ops.py:176-561   (non-TPU branches; the TPU branches :288-432 are the
                                                             opt-in `tpu_semantics=True`)

Tensors are torch CUDA tensors.  Each layer alone calls its stage entry point of libretinapost.so (rpp_decode,
rpp_topk, rpp_nms); `FusedPostProcessing` — what ModelBuilder.add_post_processing_stage builds when the whole chain
is requested — calls rpp_detect, which never materialises the [B,N,C] score tensor.  There is no CPU fallback.
"""
import ctypes

import torch

from retinanet import _native

_CLASS_DTYPES = {
    'CombinedNMS': torch.float32,      # TF kernel output (B7)
    'GlobalSoftNMS': torch.int64,      # tf.argmax (:259)
    'GlobalHardNMS': torch.int64,
    'PerClassSoftNMS': torch.int32,    # tf.fill of python ints (:468)
    'PerClassHardNMS': torch.int32,
}


def _get(params, key):
    return params[key] if isinstance(params, dict) else getattr(params, key)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _device_guard(t):
    """Context manager that makes the tensor's CUDA device current (handles belong to the device they were created
    on); CPU tensors are rejected here: there is no CPU fallback."""
    if not isinstance(t, torch.Tensor):
        raise TypeError('expected a torch.Tensor, got {}'.format(type(t)))
    if not t.is_cuda:
        raise RuntimeError('retinapost layers run on CUDA tensors only (no CPU fallback); got a {} tensor'
                           .format(t.device))
    return torch.cuda.device(t.device)


def _as_f32(t):
    if not isinstance(t, torch.Tensor):
        raise TypeError('expected a torch.Tensor, got {}'.format(type(t)))
    if not t.is_cuda:
        raise RuntimeError('retinapost layers run on CUDA tensors only (no CPU fallback); got a {} tensor'
                           .format(t.device))
    return t.to(torch.float32).contiguous()   # tf.cast(..., tf.float32) (:111-112, :539-540)


class _Handle:
    """Owns one rpp handle (include/retinapost.h rpp_create/rpp_destroy) and its cached workspaces."""

    _DUMMY_ANCHORS = {'areas': [1.0], 'aspect_ratios': [1.0], 'scales': [1.0]}

    def __init__(self, H=8, W=8, min_level=3, max_level=3, num_classes=1, anchor_params=None,
                 box_variance=(0.1, 0.1, 0.2, 0.2), scale_box_targets=False, mode='CombinedNMS',
                 iou_threshold=0.5, score_threshold=0.05, soft_nms_sigma=0.0, pre_nms_top_k=-1,
                 filter_per_class=True, max_detections=100, soft_ignores_iou_threshold=True, tpu_semantics=False):
        if not torch.cuda.is_available():
            raise RuntimeError('retinapost needs a CUDA device: there is no CPU fallback')
        ap = anchor_params or self._DUMMY_ANCHORS
        areas = [float(a) for a in _get(ap, 'areas')]
        ratios = [float(a) for a in _get(ap, 'aspect_ratios')]
        scales = [float(a) for a in _get(ap, 'scales')]
        self._keep = (
            (ctypes.c_double * len(areas))(*areas),
            (ctypes.c_double * len(ratios))(*ratios),
            (ctypes.c_double * len(scales))(*scales),
        )
        cfg = _native.RppConfig()
        cfg.H, cfg.W = int(H), int(W)
        cfg.min_level, cfg.max_level = int(min_level), int(max_level)
        cfg.num_classes = int(num_classes)
        cfg.n_areas, cfg.areas = len(areas), self._keep[0]
        cfg.n_ratios, cfg.aspect_ratios = len(ratios), self._keep[1]
        cfg.n_scales, cfg.scales = len(scales), self._keep[2]
        cfg.box_variance = (ctypes.c_float * 4)(*[float(v) for v in box_variance])
        cfg.scale_box_targets = int(bool(scale_box_targets))
        cfg.mode = _native.MODES.index(mode) if mode in _native.MODES else -1
        cfg.iou_threshold = float(iou_threshold)
        cfg.score_threshold = float(score_threshold)
        cfg.soft_nms_sigma = float('nan') if soft_nms_sigma is None else float(soft_nms_sigma)
        cfg.pre_nms_top_k = int(pre_nms_top_k)
        cfg.filter_per_class = int(bool(filter_per_class))
        cfg.max_detections = int(max_detections)
        cfg.soft_ignores_iou_threshold = int(bool(soft_ignores_iou_threshold))
        cfg.tpu_semantics = int(bool(tpu_semantics))
        # both TPU branches cast the classes to int32 (:375, :425-426)
        self.class_dtype = (torch.int32 if tpu_semantics and mode in ('GlobalHardNMS', 'PerClassHardNMS')
                            else _CLASS_DTYPES.get(mode, torch.float32))
        self.mode = mode
        self.num_classes = int(num_classes)
        self.max_detections = int(max_detections)
        self.ptr = ctypes.c_void_p()
        _native.check(_native.lib().rpp_create(ctypes.byref(cfg), ctypes.byref(self.ptr)))
        self.num_anchors = int(_native.lib().rpp_num_anchors(self.ptr))
        self._ws = {}

    def workspace(self, B, n, device):
        """Workspace tensor for calls of this geometry on the CURRENT stream.  One tensor per (B, n, device, stream)
        is kept for the life of the handle: a CUDA graph captured from a call bakes the tensor's address in, and two
        streams must not share scratch memory, so an entry is never replaced or handed to another stream
        (`release_workspaces()` frees them all when no captured graph is alive)."""
        key = (int(B), int(n), str(device), int(torch.cuda.current_stream(device).cuda_stream))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = int(_native.lib().rpp_workspace_bytes(self.ptr, int(B), int(n)))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    def release_workspaces(self):
        self._ws = {}

    def outputs(self, B, device):
        M = self.max_detections
        return {
            'boxes': torch.empty((B, M, 4), dtype=torch.float32, device=device),
            'scores': torch.empty((B, M), dtype=torch.float32, device=device),
            'classes': torch.empty((B, M), dtype=self.class_dtype, device=device),
            'valid_detections': torch.empty((B,), dtype=torch.int32, device=device),
        }

    def close(self):
        if self.ptr:
            _native.lib().rpp_destroy(self.ptr)
            self.ptr = ctypes.c_void_p()
        self._ws = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _LazyFused(dict):
    """FuseDetections' output with the concat deferred: behaves like {'class_logits', 'encoded_boxes'} (the tensors
    are concatenated on first access) and carries the per-level views for rpp_detect_levels."""

    def __init__(self, class_levels, box_levels):
        super().__init__()
        self.class_levels = class_levels
        self.box_levels = box_levels

    def __missing__(self, key):
        if key == 'class_logits':
            self[key] = torch.cat(self.class_levels, dim=1)
        elif key == 'encoded_boxes':
            self[key] = torch.cat(self.box_levels, dim=1)
        else:
            raise KeyError(key)
        return dict.__getitem__(self, key)


class Layer:
    """Stand-in for tf.keras.layers.Layer: `layer(x)` dispatches to `layer.call(x)`; accepts `name=`."""

    def __init__(self, name=None, **kwargs):
        self.name = name or type(self).__name__.lower()

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


class FuseDetections(Layer):
    """postprocessing_ops.py:7-56 — reshape each level's NHWC head output to [B, H*W*A, C] / [B, H*W*A, 4] and
    concatenate the levels.  Pure data movement (torch views + one concat)."""

    def __init__(self, min_level, max_level, lazy=False, **kwargs):
        super(FuseDetections, self).__init__(**kwargs)

        self.min_level = min_level
        self.max_level = max_level
        # lazy (not in the reference; set by the builder in front of FusedPostProcessing): also hand over the
        # per-level [B, n_l, C] / [B, n_l, 4] views and defer the concat, which rpp_detect_levels makes unnecessary
        self.lazy = lazy

    def call(self, predictions):
        class_predictions = predictions['class-predictions']
        box_predictions = predictions['box-predictions']

        box_shape = list(box_predictions[str(self.min_level)].shape)
        class_shape = list(class_predictions[str(self.min_level)].shape)

        anchors_at_each_location = box_shape[-1] // 4
        num_classes = class_shape[-1] // anchors_at_each_location
        batch_size = box_shape[0] or 1

        class_logits = []
        encoded_boxes = []
        for level in range(self.min_level, self.max_level + 1):
            level = str(level)
            class_logits.append(class_predictions[level].reshape(batch_size, -1, num_classes))
            encoded_boxes.append(box_predictions[level].reshape(batch_size, -1, 4))

        if self.lazy:
            return _LazyFused(class_logits, encoded_boxes)
        return {
            'class_logits': torch.cat(class_logits, dim=1),
            'encoded_boxes': torch.cat(encoded_boxes, dim=1)
        }


class TransformBoxesAndScores(Layer):
    """postprocessing_ops.py:59-117 — scores = sigmoid(class_logits), boxes = decode(encoded_boxes, anchors)
    normalised by [H, W, H, W].  Stage entry point: rpp_decode."""

    def __init__(self, params, **kwargs):
        super(TransformBoxesAndScores, self).__init__(**kwargs)
        self._params = params
        self._handles = {}

    def _handle(self, num_classes):
        key = (num_classes, torch.cuda.current_device())   # a handle belongs to the device it was created on
        h = self._handles.get(key)
        if h is None:
            p = self._params
            h = _Handle(H=p.input.input_shape[0], W=p.input.input_shape[1],
                        min_level=p.architecture.feature_fusion.min_level,
                        max_level=p.architecture.feature_fusion.max_level,
                        num_classes=num_classes, anchor_params=p.anchor_params,
                        box_variance=p.encoder_params.box_variance,
                        scale_box_targets=p.encoder_params.scale_box_targets)
            self._handles[key] = h
        return h

    def call(self, predictions):
        with _device_guard(predictions['class_logits']):
            return self._call(predictions)

    def _call(self, predictions):
        class_logits = _as_f32(predictions['class_logits'])
        encoded_boxes = _as_f32(predictions['encoded_boxes'])
        B, N, C = class_logits.shape
        h = self._handle(C)
        if N != h.num_anchors or tuple(encoded_boxes.shape) != (B, N, 4):
            raise ValueError('expected class_logits [B,{0},C] and encoded_boxes [B,{0},4], got {1} and {2}'.format(
                h.num_anchors, tuple(class_logits.shape), tuple(encoded_boxes.shape)))
        scores = torch.empty_like(class_logits)
        boxes = torch.empty_like(encoded_boxes)
        _native.check(_native.lib().rpp_decode(h.ptr, class_logits.data_ptr(), encoded_boxes.data_ptr(), B,
                                               scores.data_ptr(), boxes.data_ptr(), _stream()))
        return {'scores': scores, 'boxes': boxes}


class FilterTopKDetections(Layer):
    """postprocessing_ops.py:120-173 — pre-NMS top-k over the fused anchor axis, per class (:128-147) or over
    anchors x classes (:149-161).  Rows come back in tf.nn.top_k sorted=True order.  Stage entry: rpp_topk."""

    def __init__(self, top_k=100, filter_per_class=True, **kwargs):
        super(FilterTopKDetections, self).__init__(**kwargs)

        self.top_k = top_k
        self.filter_per_class = filter_per_class
        self._handles = {}

    def _handle(self, num_classes):
        key = (num_classes, torch.cuda.current_device())
        h = self._handles.get(key)
        if h is None:
            h = _Handle(num_classes=num_classes, pre_nms_top_k=self.top_k, filter_per_class=self.filter_per_class)
            self._handles[key] = h
        return h

    def call(self, predictions):
        with _device_guard(predictions['scores']):
            return self._call(predictions)

    def _call(self, predictions):
        scores = _as_f32(predictions['scores'])
        boxes = _as_f32(predictions['boxes'])
        B, n, C = scores.shape
        if tuple(boxes.shape) != (B, n, 4):
            raise ValueError('expected boxes [B,n,4], got {}'.format(tuple(boxes.shape)))
        h = self._handle(C)
        if self.filter_per_class:
            k = min(self.top_k, n)
            out_scores = torch.empty((B, k, C), dtype=torch.float32, device=scores.device)
            out_boxes = torch.empty((B, k, C, 4), dtype=torch.float32, device=scores.device)
        else:
            k = min(self.top_k, n * C)
            out_scores = torch.empty((B, k, C), dtype=torch.float32, device=scores.device)
            out_boxes = torch.empty((B, k, 4), dtype=torch.float32, device=scores.device)
        ws = h.workspace(B, n, scores.device)
        _native.check(_native.lib().rpp_topk(h.ptr, scores.data_ptr(), boxes.data_ptr(), B, n,
                                             out_scores.data_ptr(), out_boxes.data_ptr(), None,
                                             ws.data_ptr(), ws.numel(), _stream()))
        return {'scores': out_scores, 'boxes': out_boxes}


class FilterTopKDetectionsPerLevel(FilterTopKDetections):
    """Optional extension (BASELINE.json north_star: "per-level top-k pre-selection"; the reference's own filter runs
    over the fused anchor axis, postprocessing_ops.py:128-161): FilterTopKDetections applied to every pyramid level's
    segment of the anchor axis — `anchor_boundaries` of dataloader/anchor_generator.py:42-49 — and concatenated in
    level order.  `top_k` is per level.  Stage entry: rpp_topk_levels.

    Input: {'scores': [B,N,C], 'boxes': [B,N,4]} plus `anchor_boundaries`, or per-level lists / dicts of tensors
    ([B,n_l,C] / [B,n_l,4], the natural layout of the head outputs: no copy)."""

    def __init__(self, top_k=100, filter_per_class=True, anchor_boundaries=None, **kwargs):
        super(FilterTopKDetectionsPerLevel, self).__init__(top_k=top_k, filter_per_class=filter_per_class, **kwargs)
        self.anchor_boundaries = None if anchor_boundaries is None else [int(v) for v in anchor_boundaries]

    @staticmethod
    def _as_list(x):
        if isinstance(x, dict):
            return [x[k] for k in sorted(x, key=lambda v: int(v))]
        return list(x)

    def _levels(self, predictions):
        scores, boxes = predictions['scores'], predictions['boxes']
        if torch.is_tensor(scores):
            bounds = self.anchor_boundaries
            if bounds is None or bounds[0] != 0 or bounds[-1] != scores.shape[1] or sorted(bounds) != bounds:
                raise ValueError('anchor_boundaries must partition the fused anchor axis [0, {}]'.format(
                    scores.shape[1]))
            pairs = [(scores[:, a:b], boxes[:, a:b]) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        else:
            pairs = list(zip(self._as_list(scores), self._as_list(boxes)))
        return [(_as_f32(s), _as_f32(b)) for s, b in pairs]

    def call(self, predictions):
        first = predictions['scores']
        first = first if torch.is_tensor(first) else self._as_list(first)[0]
        with _device_guard(first):
            return self._call(predictions)

    def _call(self, predictions):
        levels = self._levels(predictions)
        B, _, C = levels[0][0].shape
        n_rows = [int(s.shape[1]) for s, _ in levels]
        for (s, b), n in zip(levels, n_rows):
            if tuple(s.shape) != (B, n, C) or tuple(b.shape) != (B, n, 4):
                raise ValueError('expected per-level scores [B,n_l,C] and boxes [B,n_l,4]')
        h = self._handle(C)
        K = sum(min(self.top_k, n if self.filter_per_class else n * C) for n in n_rows)
        dev = levels[0][0].device
        out_scores = torch.empty((B, K, C), dtype=torch.float32, device=dev)
        out_boxes = torch.empty((B, K, C, 4) if self.filter_per_class else (B, K, 4), dtype=torch.float32, device=dev)
        L = len(levels)
        sp = (ctypes.c_void_p * L)(*[s.data_ptr() for s, _ in levels])
        bp = (ctypes.c_void_p * L)(*[b.data_ptr() for _, b in levels])
        nr = (ctypes.c_long * L)(*n_rows)
        ws = h.workspace(B, max(n_rows, key=lambda n: _native.lib().rpp_workspace_bytes(h.ptr, B, n)), dev)
        _native.check(_native.lib().rpp_topk_levels(h.ptr, L, sp, bp, nr, B, out_scores.data_ptr(),
                                                    out_boxes.data_ptr(), None, ws.data_ptr(), ws.numel(), _stream()))
        return {'scores': out_scores, 'boxes': out_boxes}


class GenerateDetections(Layer):
    """postprocessing_ops.py:176-561 — the five NMS modes, non-TPU branches, with the reference's per-mode output
    dtypes and padding.  Stage entry point: rpp_nms."""

    _SUPPORTED_NMS_MODES = [
        'CombinedNMS',
        'GlobalSoftNMS',
        'GlobalHardNMS',
        'PerClassSoftNMS',
        'PerClassHardNMS',
    ]

    def __init__(self,
                 iou_threshold=0.5,
                 score_threshold=0.05,
                 max_detections=100,
                 soft_nms_sigma=None,
                 num_classes=None,
                 mode='CombinedNMS',
                 soft_ignores_iou_threshold=True,
                 tpu_semantics=False,
                 **kwargs):
        # soft_ignores_iou_threshold (not in the reference): True = NonMaxSuppressionV5 of TF >= 2.3, where the IoU
        # threshold is ignored when soft_nms_sigma > 0; False = the older kernel form.
        # tpu_semantics (not in the reference, which detects a TPUStrategy instead, :199-208): True = GlobalHardNMS /
        # PerClassHardNMS run _tpu_global_hard_nms (:381-432) / _tpu_per_class_hard_nms (:288-379): a real global hard
        # NMS, tf.image.non_max_suppression_padded arithmetic, int32 classes, -1 in every padded field.  As in the
        # reference under a TPUStrategy, any other mode is rejected (:202-206).

        if mode not in GenerateDetections._SUPPORTED_NMS_MODES:
            raise AssertionError(
                'Requested unsupported mode: {}, available modes are: {}'
                .format(mode, GenerateDetections._SUPPORTED_NMS_MODES))

        self._running_on_tpu = bool(tpu_semantics)

        if self._running_on_tpu:
            if mode != 'GlobalHardNMS' and mode != 'PerClassHardNMS':
                raise AssertionError(
                    'Requested mode not supported on Cloud TPUs.'
                    ' Please use `GlobalHardNMS` or `PerClassHardNMS`')

        super(GenerateDetections, self).__init__(**kwargs)

        self.iou_threshold = iou_threshold
        self.score_threshold = score_threshold
        self.max_detections = max_detections
        self.soft_nms_sigma = soft_nms_sigma
        self.num_classes = num_classes
        self.mode = mode
        self.soft_ignores_iou_threshold = soft_ignores_iou_threshold
        self._handles = {}

    def _handle(self, num_classes):
        key = (num_classes, torch.cuda.current_device())
        h = self._handles.get(key)
        if h is None:
            if self.mode in ('GlobalSoftNMS', 'PerClassSoftNMS') and self.soft_nms_sigma is None:
                # the reference evaluates `None / 2` here (:255, :450)
                raise TypeError("unsupported operand type(s) for /: 'NoneType' and 'int' "
                                "(soft NMS modes need soft_nms_sigma)")
            h = _Handle(num_classes=num_classes, mode=self.mode, iou_threshold=self.iou_threshold,
                        score_threshold=self.score_threshold, soft_nms_sigma=self.soft_nms_sigma or 0.0,
                        max_detections=self.max_detections,
                        soft_ignores_iou_threshold=self.soft_ignores_iou_threshold,
                        tpu_semantics=self._running_on_tpu)
            self._handles[key] = h
        return h

    def call(self, predictions):
        with _device_guard(predictions['scores']):
            return self._call(predictions)

    def _call(self, predictions):
        scores = _as_f32(predictions['scores'])
        boxes = _as_f32(predictions['boxes'])
        B, n, C = scores.shape
        q = 1 if boxes.dim() == 3 else boxes.shape[2]
        if self.num_classes is not None and self.mode.startswith('PerClass') and self.num_classes != C:
            raise ValueError('num_classes={} but scores have {} classes'.format(self.num_classes, C))
        h = self._handle(C)
        out = h.outputs(B, scores.device)
        ws = h.workspace(B, n, scores.device)
        _native.check(_native.lib().rpp_nms(h.ptr, scores.data_ptr(), boxes.data_ptr(), B, n, q,
                                            out['boxes'].data_ptr(), out['scores'].data_ptr(),
                                            out['classes'].data_ptr(), out['valid_detections'].data_ptr(),
                                            ws.data_ptr(), ws.numel(), _stream()))
        return {
            'scores': out['scores'],
            'boxes': out['boxes'],
            'classes': out['classes'],
            'valid_detections': out['valid_detections'],
        }


class FusedPostProcessing(Layer):
    """TransformBoxesAndScores -> FilterTopKDetections -> GenerateDetections as ONE call (rpp_detect): the chain
    ModelBuilder.add_post_processing_stage wires (model/builder.py:162-181), without the [B,N,C] score tensor, the
    transposes or the [B,k,C,4] gather ever reaching HBM.  Input: {'class_logits', 'encoded_boxes'}."""

    def __init__(self, params, **kwargs):
        super(FusedPostProcessing, self).__init__(**kwargs)
        inf = params.inference
        if inf.mode not in GenerateDetections._SUPPORTED_NMS_MODES:
            raise AssertionError(
                'Requested unsupported mode: {}, available modes are: {}'
                .format(inf.mode, GenerateDetections._SUPPORTED_NMS_MODES))
        self._params = params
        self.mode = inf.mode
        # optional key `inference.tpu_semantics` (absent from the reference's JSONs -> False)
        self.tpu_semantics = bool(inf.get('tpu_semantics', False)) if hasattr(inf, 'get') else False
        if self.tpu_semantics and inf.mode not in ('GlobalHardNMS', 'PerClassHardNMS'):
            raise AssertionError('Requested mode not supported on Cloud TPUs.'
                                 ' Please use `GlobalHardNMS` or `PerClassHardNMS`')   # :202-206
        self._handles = {}

    def handle(self, num_classes):
        key = (num_classes, torch.cuda.current_device())
        h = self._handles.get(key)
        if h is None:
            p = self._params
            inf = p.inference
            if self.mode in ('GlobalSoftNMS', 'PerClassSoftNMS') and inf.soft_nms_sigma is None:
                raise TypeError("unsupported operand type(s) for /: 'NoneType' and 'int' "
                                "(soft NMS modes need soft_nms_sigma)")
            h = _Handle(H=p.input.input_shape[0], W=p.input.input_shape[1],
                        min_level=p.architecture.feature_fusion.min_level,
                        max_level=p.architecture.feature_fusion.max_level,
                        num_classes=num_classes, anchor_params=p.anchor_params,
                        box_variance=p.encoder_params.box_variance,
                        scale_box_targets=p.encoder_params.scale_box_targets,
                        mode=inf.mode, iou_threshold=inf.iou_threshold, score_threshold=inf.score_threshold,
                        soft_nms_sigma=inf.soft_nms_sigma or 0.0, pre_nms_top_k=inf.pre_nms_top_k,
                        filter_per_class=inf.filter_per_class, max_detections=inf.max_detections,
                        tpu_semantics=self.tpu_semantics)
            self._handles[key] = h
        return h

    def capture(self, predictions, warmup=3):
        """CUDA-graphs the call for FIXED input tensors (the whole step is stream-ordered: one memset and seven kernel
        launches, no host synchronisation, so it can be captured).  Returns (replay, outputs): `replay()` re-runs the
        step on the current contents of the same input tensors and refreshes `outputs` (the same dict) in place.
        Worth ~3 % at batch 64 and ~15 % at batch 1, where launch latency dominates."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.call(predictions)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        before = {id(h): set(h._ws) for h in self._handles.values()}
        with torch.cuda.graph(graph):
            # the workspace this call uses is allocated here, inside the capture, from the graph's private pool (its
            # cache key carries the capture stream): no later eager call can evict or reuse it
            outputs = self.call(predictions)
        # ... and it is owned by the replay closure from here on, so that it lives exactly as long as the graph
        pinned = [h._ws.pop(k) for h in self._handles.values() for k in set(h._ws) - before.get(id(h), set())]

        def replay(_graph=graph, _inputs=predictions, _pinned=pinned):
            _graph.replay()
        return replay, outputs

    _DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}

    def _native_pieces(self, cls, box):
        """Can rpp_detect_typed read these tensors where they lie (per-level pieces and / or 16-bit elements)?
        Every mode / filter combination can; the staged column collect of the per-class filter (or no filter) in the
        per-class modes needs num_classes % 4 == 0 (% 8 for 16-bit elements) and at most 392 classes."""
        inf = self._params.inference
        dts = {t.dtype for t in cls + box}
        if len(dts) != 1 or next(iter(dts)) not in self._DTYPES:
            return False
        half = next(iter(dts)) != torch.float32
        C = cls[0].shape[2]
        columns = not self.mode.startswith('Global') and (inf.pre_nms_top_k <= 0 or inf.filter_per_class)
        if columns and (C % (8 if half else 4) != 0 or C > 392):
            return False
        if self.mode.startswith('Global') and inf.pre_nms_top_k > 0 and inf.filter_per_class:
            return False    # (invalid combination: the fused call reports it)
        return all(t.is_cuda and t.is_contiguous() and t.data_ptr() % 16 == 0 for t in cls + box)

    def _call_pieces(self, cls, box):
        """Head outputs in place (rpp_detect_typed): per-level pieces without the FuseDetections concat, f16 / bf16
        elements without the fp32 cast of postprocessing_ops.py:111-112 (both are folded into the loads)."""
        B, C = cls[0].shape[0], cls[0].shape[2]
        h = self.handle(C)
        if sum(t.shape[1] for t in cls) != h.num_anchors or any(b.shape[1] != c.shape[1] for b, c in zip(box, cls)):
            raise ValueError('head outputs do not add up to the {} anchors of the configured input shape'
                             .format(h.num_anchors))
        out = h.outputs(B, cls[0].device)
        ws = h.workspace(B, 0, cls[0].device)
        n = len(cls)
        cls_p = (ctypes.c_void_p * n)(*[t.data_ptr() for t in cls])
        box_p = (ctypes.c_void_p * n)(*[t.data_ptr() for t in box])
        _native.check(_native.lib().rpp_detect_typed(h.ptr, n, box_p, cls_p, self._DTYPES[cls[0].dtype], B,
                                                     out['boxes'].data_ptr(), out['scores'].data_ptr(),
                                                     out['classes'].data_ptr(), out['valid_detections'].data_ptr(),
                                                     ws.data_ptr(), ws.numel(), _stream()))
        return {
            'scores': out['scores'],
            'boxes': out['boxes'],
            'classes': out['classes'],
            'valid_detections': out['valid_detections'],
        }

    def call(self, predictions):
        first = predictions.class_levels[0] if isinstance(predictions, _LazyFused) else predictions['class_logits']
        with _device_guard(first):
            return self._call(predictions)

    def _call(self, predictions):
        if isinstance(predictions, _LazyFused):
            if self._native_pieces(predictions.class_levels, predictions.box_levels):
                return self._call_pieces(predictions.class_levels, predictions.box_levels)
        elif predictions['class_logits'].dtype != torch.float32:
            cls, box = [predictions['class_logits']], [predictions['encoded_boxes']]
            if self._native_pieces(cls, box):
                return self._call_pieces(cls, box)
        class_logits = _as_f32(predictions['class_logits'])
        encoded_boxes = _as_f32(predictions['encoded_boxes'])
        B, N, C = class_logits.shape
        h = self.handle(C)
        if N != h.num_anchors or tuple(encoded_boxes.shape) != (B, N, 4):
            raise ValueError('expected class_logits [B,{0},C] and encoded_boxes [B,{0},4], got {1} and {2}'.format(
                h.num_anchors, tuple(class_logits.shape), tuple(encoded_boxes.shape)))
        out = h.outputs(B, class_logits.device)
        ws = h.workspace(B, 0, class_logits.device)   # n = 0: sized for rpp_detect
        _native.check(_native.lib().rpp_detect(h.ptr, encoded_boxes.data_ptr(), class_logits.data_ptr(), B,
                                               out['boxes'].data_ptr(), out['scores'].data_ptr(),
                                               out['classes'].data_ptr(), out['valid_detections'].data_ptr(),
                                               ws.data_ptr(), ws.numel(), _stream()))
        return {
            'scores': out['scores'],
            'boxes': out['boxes'],
            'classes': out['classes'],
            'valid_detections': out['valid_detections'],
        }
