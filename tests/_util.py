"""Shared helpers of the parity tests: seeded synthetic inputs and detection comparison."""
import copy

import numpy as np

from conftest import REFERENCE_CONFIG


def make_params(H=640, W=None, num_classes=80, **inference):
    from retinanet.cfg.config import AttrDict
    cfg = copy.deepcopy(REFERENCE_CONFIG)
    cfg['input']['input_shape'] = [H, W or H]
    cfg['architecture']['head']['num_classes'] = num_classes
    cfg['inference'].update(inference)
    return AttrDict(cfg)


def synth_inputs(B, N, C, seed=0, dist='dense'):
    """deltas ~ N(0, 0.5^2) clipped to +-4; logits dense N(0,1) or sparse N(-4.595, 1.5^2)."""
    rng = np.random.default_rng(seed)
    deltas = np.clip(rng.standard_normal((B, N, 4)) * 0.5, -4, 4).astype(np.float32)
    if dist == 'dense':
        logits = rng.standard_normal((B, N, C)).astype(np.float32)
    elif dist == 'sparse':
        logits = (rng.standard_normal((B, N, C)) * 1.5 - 4.595).astype(np.float32)
    elif dist == 'quantized':   # many exact ties (bf16-like grid)
        logits = (np.round(rng.standard_normal((B, N, C)) * 8) / 8).astype(np.float32)
    else:
        raise ValueError(dist)
    return logits, deltas


def oracle_detect(ref, params, logits, deltas, threads=8, **kw):
    inf = params.inference
    H, W = params.input.input_shape
    ff = params.architecture.feature_fusion
    ap = params.anchor_params
    anchors, _ = ref.anchors(H, W, ff.min_level, ff.max_level, ap.areas, ap.aspect_ratios, ap.scales)
    if inf.get('tpu_semantics', False) and inf.mode in ('GlobalHardNMS', 'PerClassHardNMS'):
        return ref.detect_tpu(logits, deltas, anchors, H, W, inf.mode, iou_threshold=inf.iou_threshold,
                              score_threshold=inf.score_threshold, pre_nms_top_k=inf.pre_nms_top_k,
                              filter_per_class=inf.filter_per_class, max_detections=inf.max_detections,
                              box_variance=params.encoder_params.box_variance,
                              scale_box_targets=params.encoder_params.scale_box_targets, threads=threads)
    return ref.detect(logits, deltas, anchors, H, W, inf.mode, iou_threshold=inf.iou_threshold,
                      score_threshold=inf.score_threshold, soft_nms_sigma=inf.soft_nms_sigma,
                      pre_nms_top_k=inf.pre_nms_top_k, filter_per_class=inf.filter_per_class,
                      max_detections=inf.max_detections, box_variance=params.encoder_params.box_variance,
                      scale_box_targets=params.encoder_params.scale_box_targets, threads=threads, **kw)


def to_numpy(out):
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


def image_mismatches(got, exp, box_rtol=1e-5, box_atol=1e-6, score_exact=True):
    """Per image: kept count, classes and order bit-exact; scores bit-exact (or 1e-6); boxes within 1e-5 relative.
    Returns the list of image indices that differ (pads included: they are part of the reference contract)."""
    bad = []
    B = len(exp['valid_detections'])
    for b in range(B):
        ok = got['valid_detections'][b] == exp['valid_detections'][b]
        ok = ok and got['classes'].dtype == exp['classes'].dtype
        ok = ok and np.array_equal(got['classes'][b], exp['classes'][b])
        if score_exact:
            ok = ok and np.array_equal(got['scores'][b], exp['scores'][b])
        else:
            ok = ok and np.allclose(got['scores'][b], exp['scores'][b], rtol=1e-6, atol=0)
        ok = ok and np.allclose(got['boxes'][b], exp['boxes'][b], rtol=box_rtol, atol=box_atol)
        if not ok:
            bad.append(b)
    return bad
