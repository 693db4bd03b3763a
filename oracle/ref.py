"""ctypes/numpy front-end of the CPU oracle (oracle/retinapost_ref.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs.  The product package never imports this module.  See the header of retinapost_ref.cpp for what the
oracle restates and for the "parity unpinned (TF kernels)" caveat.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libretinapost_ref.so')

MODES = ['CombinedNMS', 'GlobalSoftNMS', 'GlobalHardNMS', 'PerClassSoftNMS', 'PerClassHardNMS']
_CLASS_DTYPE = {0: np.float32, 1: np.int64, 2: np.int64, 3: np.int32, 4: np.int32}

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_long)


def build(force=False):
    src = os.path.join(_HERE, 'retinapost_ref.cpp')
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libretinapost_ref.so'])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.rpp_ref_anchors.restype = ctypes.c_long
        _lib.rpp_ref_iou.restype = ctypes.c_float
    return _lib


def _p(a, t=_f32p):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def hardware_threads():
    return int(lib().rpp_ref_hardware_threads())


def anchors(H, W, min_level, max_level, areas, aspect_ratios, scales):
    """AnchorBoxGenerator(...).boxes and .anchor_boundaries (dataloader/anchor_generator.py)."""
    ar = np.asarray(areas, np.float64)
    ra = np.asarray(aspect_ratios, np.float64)
    sc = np.asarray(scales, np.float64)
    nl = max_level - min_level + 1
    bounds = np.zeros(nl + 1, np.int64)
    n = lib().rpp_ref_anchors(H, W, min_level, max_level, _p(ar, _f64p), len(ar), _p(ra, _f64p), len(ra),
                              _p(sc, _f64p), len(sc), None, _p(bounds, _i64p))
    out = np.empty((n, 4), np.float32)
    r = lib().rpp_ref_anchors(H, W, min_level, max_level, _p(ar, _f64p), len(ar), _p(ra, _f64p), len(ra),
                              _p(sc, _f64p), len(sc), _p(out), _p(bounds, _i64p))
    if r < 0:
        raise ValueError('anchor_params.areas shorter than the number of levels')
    return out, bounds.tolist()


def sigmoid(x, threads=1):
    x = _f32(x)
    y = np.empty_like(x)
    lib().rpp_ref_sigmoid(_p(x), _p(y), ctypes.c_long(x.size), threads)
    return y


def decode_boxes(deltas, anchor_boxes, H, W, box_variance=(0.1, 0.1, 0.2, 0.2), scale_box_targets=False, threads=1):
    """TransformBoxesAndScores._transform_box_predictions (postprocessing_ops.py:87-105)."""
    d = _f32(deltas)
    a = _f32(anchor_boxes)
    B, N = d.shape[0], d.shape[1]
    bv = np.asarray(box_variance, np.float32)
    out = np.empty((B, N, 4), np.float32)
    lib().rpp_ref_decode_boxes(_p(d), _p(a), ctypes.c_long(B), ctypes.c_long(N), H, W, _p(bv),
                               int(bool(scale_box_targets)), _p(out), threads)
    return out


def topk(values, k, sorted=True, threads=1):
    v = _f32(values)
    rows, cols = v.shape
    kk = min(k, cols)
    idx = np.empty((rows, kk), np.int32)
    lib().rpp_ref_topk(_p(v), ctypes.c_long(rows), cols, k, int(sorted), _p(idx, _i32p), threads)
    return idx


def filter_per_class(scores, boxes, k, sorted=True, threads=1):
    """FilterTopKDetections._filter_per_class (postprocessing_ops.py:128-147)."""
    s = _f32(scores)
    b = _f32(boxes)
    B, N, C = s.shape
    kk = min(k, N)
    so = np.empty((B, kk, C), np.float32)
    bo = np.empty((B, kk, C, 4), np.float32)
    idx = np.empty((B, C, kk), np.int32)
    lib().rpp_ref_filter_per_class(_p(s), _p(b), ctypes.c_long(B), ctypes.c_long(N), C, k, int(sorted), _p(so),
                                   _p(bo), _p(idx, _i32p), threads)
    return so, bo, idx


def filter_global(scores, boxes, k, sorted=True, threads=1):
    """FilterTopKDetections._filter_global (postprocessing_ops.py:149-161)."""
    s = _f32(scores)
    b = _f32(boxes)
    B, N, C = s.shape
    kk = min(k, N * C)
    so = np.empty((B, kk, C), np.float32)
    bo = np.empty((B, kk, 4), np.float32)
    idx = np.empty((B, kk), np.int32)
    lib().rpp_ref_filter_global(_p(s), _p(b), ctypes.c_long(B), ctypes.c_long(N), C, k, int(sorted), _p(so), _p(bo),
                                _p(idx, _i32p), threads)
    return so, bo, idx


def filter_per_level(scores, boxes, k, anchor_boundaries, per_class=True, sorted=True, threads=1):
    """FilterTopKDetections applied to every pyramid level's segment of the fused anchor axis (anchor_boundaries of
    dataloader/anchor_generator.py:42-49) and concatenated in level order — the optional per-level pre-selection of
    BASELINE.json's north_star.  The reference has no such mode: this is its own filter (:128-161), composed.
    Indices are positions on the fused axis."""
    s = _f32(scores)
    b = _f32(boxes)
    C = s.shape[2]
    so, bo, io = [], [], []
    for lo, hi in zip(anchor_boundaries[:-1], anchor_boundaries[1:]):
        if hi <= lo:
            continue
        f = filter_per_class if per_class else filter_global
        s_l, b_l, i_l = f(s[:, lo:hi], b[:, lo:hi], k, sorted=sorted, threads=threads)
        so.append(s_l)
        bo.append(b_l)
        io.append(i_l + (lo if per_class else lo * C))
    return np.concatenate(so, axis=1), np.concatenate(bo, axis=1), np.concatenate(io, axis=-1)


def iou(a, b):
    a = _f32(a)
    b = _f32(b)
    return float(lib().rpp_ref_iou(_p(a), _p(b)))


def nms_v5(boxes, scores, max_output_size, iou_threshold, score_threshold, soft_nms_sigma=0.0,
           soft_ignores_iou_threshold=True):
    """tf.raw_ops.NonMaxSuppressionV5(pad_to_max_output_size=True) -> (indices[M], scores[M], valid)."""
    b = _f32(boxes)
    s = _f32(scores)
    M = int(max_output_size)
    idx = np.zeros(M, np.int32)
    sc = np.zeros(M, np.float32)
    valid = lib().rpp_ref_nms_v5(_p(b), _p(s), len(s), M, ctypes.c_float(iou_threshold),
                                 ctypes.c_float(score_threshold), ctypes.c_float(soft_nms_sigma),
                                 int(bool(soft_ignores_iou_threshold)), _p(idx, _i32p), _p(sc))
    return idx, sc, int(valid)


def _mode_id(mode):
    return MODES.index(mode) if isinstance(mode, str) else int(mode)


def generate_detections(mode, scores, boxes, iou_threshold=0.5, score_threshold=0.05, max_detections=100,
                        soft_nms_sigma=0.5, topk_sorted=True, soft_ignores_iou_threshold=True, threads=1):
    """GenerateDetections(...).call({'scores','boxes'}) (postprocessing_ops.py:537-561), non-TPU branches."""
    m = _mode_id(mode)
    s = _f32(scores)
    b = _f32(boxes)
    B, n, C = s.shape
    q = 1 if b.ndim == 3 else b.shape[2]
    M = int(max_detections)
    bo = np.empty((B, M, 4), np.float32)
    so = np.empty((B, M), np.float32)
    co = np.empty((B, M), _CLASS_DTYPE[m])
    vo = np.empty((B,), np.int32)
    sigma = 0.0 if soft_nms_sigma is None else float(soft_nms_sigma)
    r = lib().rpp_ref_generate_detections(m, _p(s), _p(b), ctypes.c_long(B), ctypes.c_long(n), q, C,
                                          ctypes.c_float(iou_threshold), ctypes.c_float(score_threshold), M,
                                          ctypes.c_float(sigma), int(topk_sorted),
                                          int(bool(soft_ignores_iou_threshold)), _p(bo), _p(so),
                                          co.ctypes.data_as(ctypes.c_void_p), _p(vo, _i32p), threads)
    if r != 0:
        raise ValueError('invalid mode / box rank combination (Global* modes need 3-D boxes)')
    return {'boxes': bo, 'scores': so, 'classes': co, 'valid_detections': vo}


def detect(class_logits, encoded_boxes, anchor_boxes, H, W, mode, iou_threshold=0.5, score_threshold=0.05,
           soft_nms_sigma=0.5, pre_nms_top_k=5000, filter_per_class=True, max_detections=100,
           box_variance=(0.1, 0.1, 0.2, 0.2), scale_box_targets=False, topk_sorted=True,
           soft_ignores_iou_threshold=True, threads=1):
    """TransformBoxesAndScores -> FilterTopKDetections -> GenerateDetections (model/builder.py:162-181)."""
    m = _mode_id(mode)
    lg = _f32(class_logits)
    d = _f32(encoded_boxes)
    a = _f32(anchor_boxes)
    B, N, C = lg.shape
    M = int(max_detections)
    bv = np.asarray(box_variance, np.float32)
    bo = np.empty((B, M, 4), np.float32)
    so = np.empty((B, M), np.float32)
    co = np.empty((B, M), _CLASS_DTYPE[m])
    vo = np.empty((B,), np.int32)
    sigma = 0.0 if soft_nms_sigma is None else float(soft_nms_sigma)
    r = lib().rpp_ref_detect(_p(lg), _p(d), _p(a), ctypes.c_long(B), ctypes.c_long(N), C, H, W, _p(bv),
                             int(bool(scale_box_targets)), m, ctypes.c_float(iou_threshold),
                             ctypes.c_float(score_threshold), ctypes.c_float(sigma), int(pre_nms_top_k),
                             int(bool(filter_per_class)), M, int(topk_sorted), int(bool(soft_ignores_iou_threshold)),
                             _p(bo), _p(so), co.ctypes.data_as(ctypes.c_void_p), _p(vo, _i32p), threads)
    if r != 0:
        raise ValueError('invalid mode / filter combination (Global* modes need filter_per_class=False)')
    return {'boxes': bo, 'scores': so, 'classes': co, 'valid_detections': vo}


def iou_padded(a, b):
    """_bbox_overlap of tf.image.non_max_suppression_padded for one pair of boxes."""
    a = _f32(a)
    b = _f32(b)
    lib().rpp_ref_iou_padded.restype = ctypes.c_float
    return float(lib().rpp_ref_iou_padded(_p(a), _p(b)))


def nms_padded(boxes, scores, max_output_size, iou_threshold, score_threshold=None, exact_fixed_point=True):
    """tf.image.non_max_suppression_padded(..., pad_to_max_output_size=True, canonicalized_coordinates=True) for one
    image -> (indices[M], num_valid).  score_threshold=None: no score filter (the per-class TPU branch)."""
    b = _f32(boxes)
    s = _f32(scores)
    M = int(max_output_size)
    idx = np.zeros(M, np.int32)
    use = score_threshold is not None
    valid = lib().rpp_ref_nms_padded(_p(b), _p(s), len(s), M, ctypes.c_float(iou_threshold), int(use),
                                     ctypes.c_float(score_threshold if use else 0.0), int(bool(exact_fixed_point)),
                                     _p(idx, _i32p))
    return idx, int(valid)


def generate_detections_tpu(mode, scores, boxes, iou_threshold=0.5, score_threshold=0.05, max_detections=100,
                            exact_fixed_point=True, threads=1):
    """GenerateDetections under TPUStrategy (postprocessing_ops.py:288-432): mode GlobalHardNMS / PerClassHardNMS."""
    m = _mode_id(mode)
    s = _f32(scores)
    b = _f32(boxes)
    B, n, C = s.shape
    q = 1 if b.ndim == 3 else b.shape[2]
    M = int(max_detections)
    bo = np.empty((B, M, 4), np.float32)
    so = np.empty((B, M), np.float32)
    co = np.empty((B, M), np.int32)
    vo = np.empty((B,), np.int32)
    r = lib().rpp_ref_generate_detections_tpu(m, _p(s), _p(b), ctypes.c_long(B), ctypes.c_long(n), q, C,
                                              ctypes.c_float(iou_threshold), ctypes.c_float(score_threshold), M,
                                              int(bool(exact_fixed_point)), _p(bo), _p(so), _p(co, _i32p),
                                              _p(vo, _i32p), threads)
    if r != 0:
        raise ValueError('the TPU branches exist for GlobalHardNMS (3-D boxes) and PerClassHardNMS only')
    return {'boxes': bo, 'scores': so, 'classes': co, 'valid_detections': vo}


def detect_tpu(class_logits, encoded_boxes, anchor_boxes, H, W, mode, iou_threshold=0.5, score_threshold=0.05,
               pre_nms_top_k=5000, filter_per_class=True, max_detections=100, box_variance=(0.1, 0.1, 0.2, 0.2),
               scale_box_targets=False, threads=1):
    """TransformBoxesAndScores -> FilterTopKDetections -> GenerateDetections (TPU branches)."""
    m = _mode_id(mode)
    lg = _f32(class_logits)
    d = _f32(encoded_boxes)
    a = _f32(anchor_boxes)
    B, N, C = lg.shape
    M = int(max_detections)
    bv = np.asarray(box_variance, np.float32)
    bo = np.empty((B, M, 4), np.float32)
    so = np.empty((B, M), np.float32)
    co = np.empty((B, M), np.int32)
    vo = np.empty((B,), np.int32)
    r = lib().rpp_ref_detect_tpu(_p(lg), _p(d), _p(a), ctypes.c_long(B), ctypes.c_long(N), C, H, W, _p(bv),
                                 int(bool(scale_box_targets)), m, ctypes.c_float(iou_threshold),
                                 ctypes.c_float(score_threshold), int(pre_nms_top_k), int(bool(filter_per_class)), M,
                                 _p(bo), _p(so), _p(co, _i32p), _p(vo, _i32p), threads)
    if r != 0:
        raise ValueError('invalid mode / filter combination for the TPU branches')
    return {'boxes': bo, 'scores': so, 'classes': co, 'valid_detections': vo}


def efficient_nms(raw_boxes, class_logits, anchor_boxes, max_output_boxes=100, score_threshold=0.05, iou_threshold=0.5,
                  threads=1):
    """EfficientNMS_TRT as the reference configures it (onnx_utils.py:38-46) -> (valid_detections [B,1],
    detection_boxes [B,M,4] centre-size, detection_scores [B,M], detection_classes [B,M] i32).  Parity unpinned."""
    d = _f32(raw_boxes)
    lg = _f32(class_logits)
    a = _f32(anchor_boxes).reshape(-1, 4)
    B, N, C = lg.shape
    M = int(max_output_boxes)
    vo = np.empty((B, 1), np.int32)
    bo = np.empty((B, M, 4), np.float32)
    so = np.empty((B, M), np.float32)
    co = np.empty((B, M), np.int32)
    lib().rpp_ref_efficient_nms(_p(d), _p(lg), _p(a), ctypes.c_long(B), ctypes.c_long(N), C, M,
                                ctypes.c_float(score_threshold), ctypes.c_float(iou_threshold), _p(vo, _i32p), _p(bo),
                                _p(so), _p(co, _i32p), threads)
    return vo, bo, so, co


def coco_format(detections, image_ids, resize_scales, input_shape, rescale_detections=True, class_id_map=None):
    """COCOEvaluator.accumulate_results (eval/coco_evaluator.py:95-134) restated in numpy, line by line."""
    out = []
    f = np.float32
    for i in range(len(image_ids)):
        v = int(detections['valid_detections'][i])                      # :113
        boxes = np.array(detections['boxes'][i][:v], dtype=f)           # :114
        classes = detections['classes'][i][:v]                          # :115
        scores = detections['scores'][i][:v]                            # :116
        if rescale_detections:
            rs = np.asarray(resize_scales[i], f) / np.asarray(input_shape, f)   # :119
            boxes = boxes / np.tile(rs[None], (1, 2))                   # :120-123
        boxes = np.int32(boxes)                                         # :125
        boxes[:, 2:] = boxes[:, 2:] - boxes[:, :2]                      # :126
        for box, c, s in zip(boxes, classes, scores):
            c = int(c)
            if class_id_map is not None:                                # _maybe_remap_class_ids :89-93
                c = class_id_map[c]
            out.append({'image_id': int(image_ids[i]), 'category_id': c, 'bbox': box.tolist(), 'score': float(s)})
    return out
