"""A numpy stand-in for the handful of tf.* symbols the reference's post-processing modules touch, so that the
UNMODIFIED reference files
    /root/reference/retinanet/dataloader/anchor_generator.py
    /root/reference/retinanet/model/layers/postprocessing_ops.py
can be imported and executed in a container without TensorFlow (make_golden.py).  What this pins: the reference's
own Python glue — op order of the anchor generator and the decode, the transposes/gathers of the filters, mode
dispatch, clipping, thresholds passed to the NMS ops, padding and dtypes.  What it cannot pin: the TF C++ kernels
themselves (NonMaxSuppressionV5, CombinedNonMaxSuppression, TopKV2, Eigen sigmoid/exp), which are restated here in
plain Python/numpy from SURVEY.md Appendix A, independently of oracle/retinapost_ref.cpp.

Test infrastructure only; never imported by the product.
"""
import math
import sys
import types

import numpy as np


class Tensor(np.ndarray):
    def get_shape(self):
        return _Shape(self.shape)

    def numpy(self):
        return np.asarray(self)


class _Shape(tuple):
    def as_list(self):
        return list(self)


def T(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(Tensor)


float32 = np.float32
int32 = np.int32
int64 = np.int64


def _f32(x):
    """Python numbers become float32 tensors (TF's default conversion for floats)."""
    if isinstance(x, np.ndarray):
        return x
    return np.asarray(x, dtype=np.float32)


def constant(v, dtype=None):
    return T(v, dtype)


def convert_to_tensor(v, dtype=None):
    return T(v, dtype)


def cast(x, dtype=None):
    return T(np.asarray(x).astype(dtype))


def range(*args, dtype=None):  # noqa: A001
    args = [int(a) for a in args]
    return T(np.arange(*args), dtype or np.int32)


def meshgrid(x, y):
    gx, gy = np.meshgrid(np.asarray(x), np.asarray(y))
    return T(gx), T(gy)


def stack(values, axis=0):
    return T(np.stack([_f32(v) for v in values], axis=axis))


def concat(values, axis=0):
    return T(np.concatenate([np.asarray(v) for v in values], axis=axis))


def expand_dims(x, axis):
    return T(np.expand_dims(np.asarray(x), axis))


def tile(x, multiples):
    return T(np.tile(np.asarray(x), [int(m) for m in multiples]))


def reshape(x, shape):
    return T(np.reshape(np.asarray(x), [int(s) for s in shape]))


def transpose(x, perm):
    return T(np.transpose(np.asarray(x), perm))


def fill(dims, value):
    dtype = np.int32 if isinstance(value, (int, np.integer)) else np.float32   # TF: python int -> int32
    return T(np.full([int(d) for d in dims], value, dtype=dtype))


def where(cond, a, b):
    a = np.asarray(a)
    return T(np.where(np.asarray(cond), a, np.asarray(b, dtype=a.dtype)))


def less(a, b):
    return T(np.less(np.asarray(a), np.asarray(b)))


def greater(a, b):
    return T(np.greater(np.asarray(a), np.asarray(b)))


def reduce_max(x, axis=None):
    return T(np.max(np.asarray(x), axis=axis))


def reduce_sum(x, axis=None):
    return T(np.sum(np.asarray(x), axis=axis, dtype=np.asarray(x).dtype))


def argmax(x, axis=None):
    return T(np.argmax(np.asarray(x), axis=axis).astype(np.int64))   # first maximum, int64 like tf.argmax


def clip_by_value(x, lo, hi):
    return T(np.minimum(np.maximum(np.asarray(x), np.float32(lo)), np.float32(hi)))


def gather(params, indices, batch_dims=0):
    p, i = np.asarray(params), np.asarray(indices)
    if batch_dims == 0:
        return T(p[i])
    assert batch_dims == 1
    return T(np.stack([p[b][i[b]] for b in builtins_range(p.shape[0])]))


def gather_nd(params, indices, batch_dims=0):
    p, i = np.asarray(params), np.asarray(indices)
    assert batch_dims == 1 and i.shape[-1] == 1
    return T(np.stack([p[b][i[b, :, 0]] for b in builtins_range(p.shape[0])]))


def vectorized_map(fn, elems):
    n = np.asarray(elems[0]).shape[0]
    outs = [fn(tuple(T(np.asarray(e)[b]) for e in elems)) for b in builtins_range(n)]
    return tuple(T(np.stack([np.asarray(o[j]) for o in outs])) for j in builtins_range(len(outs[0])))


builtins_range = __builtins__['range'] if isinstance(__builtins__, dict) else __builtins__.range


# ---- restated TF kernels, plain Python --------------------------------------------------
def _sigmoid(x):
    x = np.asarray(x, np.float32)
    with np.errstate(over='ignore'):
        return T((1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(np.float32))


def _exp(x):
    x = np.asarray(x, np.float32)
    return T(np.exp(x.astype(np.float64)).astype(np.float32))


def _sqrt(x):
    return T(np.sqrt(np.asarray(x, np.float32)))


def _ceil(x):
    return T(np.ceil(np.asarray(x, np.float32)))


class _TopK(tuple):
    values = property(lambda s: s[0])
    indices = property(lambda s: s[1])


def top_k(input, k=1, sorted=True, name=None):  # noqa: A002
    """TopKV2: best = higher value, ties -> lower index.  sorted=False on TF-CPU returns gtl::TopN's heap layout;
    the fixtures use the canonical sorted order for both."""
    v = np.asarray(input)
    flat = v.reshape(-1, v.shape[-1])
    idx = np.stack([np.argsort(-row, kind='stable')[:k] for row in flat]).astype(np.int32)
    val = np.take_along_axis(flat, idx, -1)
    shp = v.shape[:-1] + (k,)
    return _TopK((T(val.reshape(shp)), T(idx.reshape(shp))))


def _iou(a, b):
    f = np.float32
    ymin_i, xmin_i, ymax_i, xmax_i = min(a[0], a[2]), min(a[1], a[3]), max(a[0], a[2]), max(a[1], a[3])
    ymin_j, xmin_j, ymax_j, xmax_j = min(b[0], b[2]), min(b[1], b[3]), max(b[0], b[2]), max(b[1], b[3])
    area_i = f(f(ymax_i - ymin_i) * f(xmax_i - xmin_i))
    area_j = f(f(ymax_j - ymin_j) * f(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return f(0)
    iymin, ixmin = max(ymin_i, ymin_j), max(xmin_i, xmin_j)
    iymax, ixmax = min(ymax_i, ymax_j), min(xmax_i, xmax_j)
    inter = f(max(f(iymax - iymin), f(0)) * max(f(ixmax - ixmin), f(0)))
    return f(inter / f(f(area_i + area_j) - inter))


def _expf(x):
    # libm expf via ctypes (what Eigen::numext::exp<float> calls)
    import ctypes
    import ctypes.util
    lib = _expf.lib = getattr(_expf, 'lib', None) or ctypes.CDLL(ctypes.util.find_library('m'))
    lib.expf.restype = ctypes.c_float
    lib.expf.argtypes = [ctypes.c_float]
    return np.float32(lib.expf(float(x)))


def _nms_v5(boxes, scores, max_output_size, iou_threshold, score_threshold, soft_nms_sigma):
    f = np.float32
    boxes, scores = np.asarray(boxes, f), np.asarray(scores, f)
    iou_threshold, score_threshold, soft_nms_sigma = f(iou_threshold), f(score_threshold), f(soft_nms_sigma)
    cands = [[f(scores[i]), i, 0] for i in builtins_range(len(scores)) if scores[i] > score_threshold]
    is_soft = soft_nms_sigma > 0
    scale = f(-0.5) / soft_nms_sigma if is_soft else f(0)
    selected, sel_scores = [], []
    while len(selected) < max_output_size and cands:
        best = max(builtins_range(len(cands)), key=lambda t: (cands[t][0], -cands[t][1]))
        score, idx, begin = cands.pop(best)
        original, hard = score, False
        for j in builtins_range(len(selected) - 1, begin - 1, -1):
            sim = _iou(boxes[idx], boxes[selected[j]])
            w = _expf(f(f(scale * sim) * sim))
            if not (is_soft or sim <= iou_threshold):
                w = f(0)
            score = f(score * w)
            if not is_soft and sim > iou_threshold:
                hard = True
                break
            if score <= score_threshold:
                break
        begin = len(selected)
        if not hard:
            if score == original:
                selected.append(idx)
                sel_scores.append(score)
                continue
            if score > score_threshold:
                cands.append([score, idx, begin])
    valid = len(selected)
    pad = max_output_size - valid
    return (T(np.array(selected + [0] * pad, np.int32)), T(np.array(sel_scores + [0.0] * pad, np.float32)),
            T(np.int32(valid)))


class _RawOps:
    @staticmethod
    def NonMaxSuppressionV5(boxes, scores, max_output_size, iou_threshold, score_threshold, soft_nms_sigma,
                            pad_to_max_output_size=False):
        assert pad_to_max_output_size
        assert np.asarray(boxes).ndim == 2, 'NonMaxSuppressionV5: boxes must be rank 2'
        return _nms_v5(boxes, scores, int(max_output_size), iou_threshold, score_threshold, soft_nms_sigma)


class _Combined(tuple):
    nmsed_boxes = property(lambda s: s[0])
    nmsed_scores = property(lambda s: s[1])
    nmsed_classes = property(lambda s: s[2])
    valid_detections = property(lambda s: s[3])


def _combined_nms(boxes, scores, max_output_size_per_class, max_total_size, iou_threshold=0.5,
                  score_threshold=float('-inf'), pad_per_class=False, clip_boxes=True, name=None):
    f = np.float32
    boxes, scores = np.asarray(boxes, f), np.asarray(scores, f)
    B, n, q, _ = boxes.shape
    C = scores.shape[2]
    M = int(max_total_size)
    per_class = min(int(max_output_size_per_class), n)
    ob, os_, oc, ov = np.zeros((B, M, 4), f), np.zeros((B, M), f), np.zeros((B, M), f), np.zeros((B,), np.int32)
    for b in builtins_range(B):
        res = []
        for c in builtins_range(C):
            order = [i for i in np.argsort(-scores[b, :, c], kind='stable') if scores[b, i, c] > f(score_threshold)]
            kept = []
            for i in order:
                if len(kept) >= per_class:
                    break
                bx = boxes[b, i, c if q > 1 else 0]
                if all(not (_iou(bx, boxes[b, j, c if q > 1 else 0]) > f(iou_threshold)) for j in reversed(kept)):
                    kept.append(i)
            res += [(scores[b, i, c], c, boxes[b, i, c if q > 1 else 0]) for i in kept]
        res.sort(key=lambda t: -t[0])   # stable: (score desc, class asc, selection order) — canonical tie order
        res = res[:M]
        ov[b] = len(res)
        for t, (s, c, bx) in enumerate(res):
            ob[b, t] = np.clip(bx, 0, 1) if clip_boxes else bx
            os_[b, t] = s
            oc[b, t] = c
    return _Combined((T(ob), T(os_), T(oc), T(ov)))


def ones_like(x):
    return T(np.ones_like(np.asarray(x)))


def _bbox_overlap(a, b):
    """image_ops_impl._bbox_overlap for one pair, fp32 step by step."""
    f = np.float32
    i_xmin, i_xmax = max(a[1], b[1]), min(a[3], b[3])
    i_ymin, i_ymax = max(a[0], b[0]), min(a[2], b[2])
    i_area = f(max(f(i_xmax - i_xmin), f(0))) * f(max(f(i_ymax - i_ymin), f(0)))
    a_area = f(f(a[2] - a[0]) * f(a[3] - a[1]))
    b_area = f(f(b[2] - b[0]) * f(b[3] - b[1]))
    u_area = f(f(f(a_area + b_area) - i_area) + f(1e-8))
    return f(i_area / u_area)


def _nms_padded(boxes, scores, max_output_size, iou_threshold=0.5, score_threshold=float('-inf'),
                pad_to_max_output_size=False, name=None, sorted_input=False, canonicalized_coordinates=False,
                tile_size=512):
    """tf.image.non_max_suppression_padded (-> non_max_suppression_padded_v2) for batched [B,n,4] / [B,n] inputs, as
    the reference's TPU branches call it.  Stated as the greedy scan that the op's tiled fixed-point iteration
    converges to: (score desc, index asc) order, a box is dropped when its _bbox_overlap with an earlier kept box is
    >= iou_threshold, boxes at or below score_threshold are zeroed first, only boxes with a coordinate > 0 count as
    selected, indices beyond num_valid are 0.  (The C++ oracle restates the tile iteration itself; the two are
    compared in tests/test_oracle_tpu.py.)"""
    assert pad_to_max_output_size and canonicalized_coordinates and not sorted_input
    boxes = np.asarray(boxes, np.float32)
    scores = np.asarray(scores, np.float32)
    B, n = scores.shape
    M = int(max_output_size)
    idx = np.zeros((B, M), np.int32)
    valid = np.zeros((B,), np.int32)
    for b in builtins_range(B):
        sc, bx = scores[b].copy(), boxes[b].copy()
        if score_threshold != float('-inf'):
            mask = (sc > np.float32(score_threshold)).astype(np.float32)
            sc = sc * mask
            bx = bx * mask[:, None]
        order = np.argsort(-sc, kind='stable')
        kept = []
        for i in order:
            if len(kept) >= M:
                break
            if not (bx[i] > 0).any():
                continue
            if all(_bbox_overlap(bx[j], bx[i]) < np.float32(iou_threshold) for j in kept):
                kept.append(i)
        valid[b] = len(kept)
        idx[b, :len(kept)] = kept
    return T(idx), T(valid)


class _Layer:
    def __init__(self, **kwargs):
        pass

    def __call__(self, *a, **k):
        return self.call(*a, **k)


class _TPUStrategy:
    pass


_strategy = [object()]


def set_tpu_strategy(on):
    """Makes tf.distribute.get_strategy() return a TPUStrategy (the reference's TPU detection, :199-208)."""
    _strategy[0] = _TPUStrategy() if on else object()


def install():
    """Registers this module tree as `tensorflow` and returns it."""
    tf = types.ModuleType('tensorflow')
    for name in ['float32', 'int32', 'int64', 'constant', 'convert_to_tensor', 'cast', 'range', 'meshgrid', 'stack',
                 'concat', 'expand_dims', 'tile', 'reshape', 'transpose', 'fill', 'where', 'less', 'greater',
                 'reduce_max', 'reduce_sum', 'argmax', 'clip_by_value', 'gather', 'gather_nd', 'vectorized_map',
                 'ones_like']:
        setattr(tf, name, globals()[name])
    tf.math = types.SimpleNamespace(sqrt=_sqrt, ceil=_ceil, exp=_exp, top_k=top_k)
    tf.nn = types.SimpleNamespace(sigmoid=_sigmoid, top_k=top_k)
    tf.image = types.SimpleNamespace(combined_non_max_suppression=_combined_nms,
                                     non_max_suppression_padded=_nms_padded)
    tf.raw_ops = _RawOps
    tf.keras = types.SimpleNamespace(layers=types.SimpleNamespace(Layer=_Layer))
    tf.nest = types.SimpleNamespace(map_structure=lambda fn, d: {k: fn(v) for k, v in d.items()})
    tf.distribute = types.SimpleNamespace(get_strategy=lambda: _strategy[0], TPUStrategy=_TPUStrategy)
    sys.modules['tensorflow'] = tf
    return tf
