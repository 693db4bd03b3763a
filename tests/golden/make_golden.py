#!/usr/bin/env python
"""Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE modules
    /root/reference/retinanet/dataloader/anchor_generator.py
    /root/reference/retinanet/model/layers/postprocessing_ops.py
over the numpy stand-in for TensorFlow in tf_shim.py (TensorFlow itself is not installable in this image).
Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py

Each fixture holds seeded inputs and the reference layers' outputs:
  anchors_*.npz      AnchorBoxGenerator(...).boxes / anchor_boundaries
  stage_*.npz        TransformBoxesAndScores / FilterTopKDetections outputs
  detect_*.npz       the full chain FuseDetections-less: Transform -> [Filter] -> GenerateDetections, every mode
  tpu_*.npz          the same chain with GenerateDetections constructed under a TPUStrategy (the _tpu_* branches)
"""
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = '/root/reference'


def load_reference():
    import tf_shim
    tf_shim.install()
    # package shells so that the reference's package __init__ files (which pull in the training stack) do not run
    for name, path in [('retinanet', 'retinanet'), ('retinanet.dataloader', 'retinanet/dataloader'),
                       ('retinanet.model', 'retinanet/model'), ('retinanet.model.layers', 'retinanet/model/layers')]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, path)]
        sys.modules[name] = m
    import importlib
    ag = importlib.import_module('retinanet.dataloader.anchor_generator')
    po = importlib.import_module('retinanet.model.layers.postprocessing_ops')
    assert ag.__file__.startswith(REF) and po.__file__.startswith(REF)
    return ag, po


class AD(dict):
    __getattr__ = dict.__getitem__


def params_for(H, W, C, min_level=3, max_level=7, scale_box_targets=False):
    return AD(input=AD(input_shape=[H, W]),
              architecture=AD(feature_fusion=AD(min_level=min_level, max_level=max_level), head=AD(num_classes=C)),
              anchor_params=AD(areas=[1024.0, 4096.0, 16384.0, 65536.0, 262144.0], aspect_ratios=[0.5, 1.0, 2.0],
                               scales=[1, 1.2599210498948732, 1.5874010519681994]),
              encoder_params=AD(box_variance=[0.1, 0.1, 0.2, 0.2], scale_box_targets=scale_box_targets))


def synth(B, N, C, seed, dist):
    rng = np.random.default_rng(seed)
    deltas = np.clip(rng.standard_normal((B, N, 4)) * 0.5, -4, 4).astype(np.float32)
    logits = rng.standard_normal((B, N, C)).astype(np.float32)
    if dist == 'sparse':
        logits = (logits * 1.5 - 4.595).astype(np.float32)
    if dist == 'quantized':
        logits = (np.round(logits * 4) / 4).astype(np.float32)
    if dist == 'clustered':   # boxes stay close to their anchors: heavy suppression, classes run out of boxes
        deltas = (deltas * 0.1).astype(np.float32)
    return logits, deltas


def main():
    ag, po = load_reference()
    out_dir = HERE
    # ---- anchors -------------------------------------------------------------------------------------------------
    for (H, W, lo, hi) in [(640, 640, 3, 7), (320, 320, 3, 7), (448, 448, 3, 6), (96, 160, 3, 7)]:
        p = params_for(H, W, 1, lo, hi)
        g = ag.AnchorBoxGenerator(H, W, lo, hi, p.anchor_params)
        boxes = np.asarray(g.boxes, np.float32)
        keep = boxes if len(boxes) <= 4096 else np.concatenate([boxes[:512], boxes[-512:]])
        np.savez_compressed(os.path.join(out_dir, 'anchors_{}x{}_l{}{}.npz'.format(H, W, lo, hi)),
                            H=H, W=W, min_level=lo, max_level=hi, n=len(boxes),
                            boundaries=np.asarray(g.anchor_boundaries, np.int64), rows=keep,
                            sha256=hashlib.sha256(boxes.tobytes()).hexdigest())
    # ---- stages + full chain ---------------------------------------------------------------------------------------
    H = W = 64
    C, B, M = 5, 2, 20
    cases = []
    for mode in po.GenerateDetections._SUPPORTED_NMS_MODES:
        for (k, fpc) in [(60, True), (90, False), (-1, True)]:
            if mode.startswith('Global') and k > 0 and fpc:
                continue
            for dist in ['dense', 'sparse', 'quantized']:
                cases.append((mode, k, fpc, dist))
    for ci, (mode, k, fpc, dist) in enumerate(cases):
        sbt = ci % 4 == 3   # exercise encoder_params.scale_box_targets on a quarter of the cases
        p = params_for(H, W, C, scale_box_targets=sbt)
        N = ag.AnchorBoxGenerator(H, W, 3, 7, p.anchor_params).boxes.shape[0]
        logits, deltas = synth(B, N, C, 1000 + ci, dist)
        x = po.TransformBoxesAndScores(p)({'class_logits': logits, 'encoded_boxes': deltas})
        stage = {'scores': np.asarray(x['scores']), 'boxes': np.asarray(x['boxes'])}
        if k > 0:
            x = po.FilterTopKDetections(top_k=k, filter_per_class=fpc)(x)
            stage['filtered_scores'] = np.asarray(x['scores'])
            stage['filtered_boxes'] = np.asarray(x['boxes'])
        det = po.GenerateDetections(iou_threshold=0.5, score_threshold=0.05, max_detections=M, soft_nms_sigma=0.5,
                                    num_classes=C, mode=mode)(x)
        name = 'detect_{}_k{}_{}_{}.npz'.format(mode, k, 'pc' if fpc else 'gl', dist)
        np.savez_compressed(os.path.join(out_dir, name), H=H, W=W, C=C, M=M, mode=mode, k=k, filter_per_class=fpc,
                            scale_box_targets=sbt, logits=logits, deltas=deltas,
                            out_boxes=np.asarray(det['boxes']), out_scores=np.asarray(det['scores']),
                            out_classes=np.asarray(det['classes']),
                            out_valid=np.asarray(det['valid_detections']), **stage)
    # ---- the TPUStrategy branches (_tpu_global_hard_nms / _tpu_per_class_hard_nms, :288-432) -------------------------
    import tf_shim
    tpu_cases = []
    for mode in ['GlobalHardNMS', 'PerClassHardNMS']:
        for (k, fpc) in [(60, True), (25, True), (90, False), (-1, True)]:
            if mode.startswith('Global') and k > 0 and fpc:
                continue
            for dist in ['dense', 'sparse', 'quantized', 'clustered']:
                tpu_cases.append((mode, k, fpc, dist))
    for ci, (mode, k, fpc, dist) in enumerate(tpu_cases):
        p = params_for(H, W, C)
        N = ag.AnchorBoxGenerator(H, W, 3, 7, p.anchor_params).boxes.shape[0]
        logits, deltas = synth(B, N, C, 2000 + ci, dist)
        x = po.TransformBoxesAndScores(p)({'class_logits': logits, 'encoded_boxes': deltas})
        if k > 0:
            x = po.FilterTopKDetections(top_k=k, filter_per_class=fpc)(x)
        tf_shim.set_tpu_strategy(True)
        try:
            layer = po.GenerateDetections(iou_threshold=0.5, score_threshold=0.05, max_detections=M,
                                          soft_nms_sigma=0.5, num_classes=C, mode=mode)
        finally:
            tf_shim.set_tpu_strategy(False)
        assert layer._running_on_tpu
        det = layer(x)
        name = 'tpu_{}_k{}_{}_{}.npz'.format(mode, k, 'pc' if fpc else 'gl', dist)
        np.savez_compressed(os.path.join(out_dir, name), H=H, W=W, C=C, M=M, mode=mode, k=k, filter_per_class=fpc,
                            scale_box_targets=False, logits=logits, deltas=deltas,
                            out_boxes=np.asarray(det['boxes']), out_scores=np.asarray(det['scores']),
                            out_classes=np.asarray(det['classes']),
                            out_valid=np.asarray(det['valid_detections']))
    print('wrote {} tpu fixtures'.format(len(tpu_cases)))
    # ---- the rank error of Global* modes on per-class filtered boxes ------------------------------------
    p = params_for(H, W, C)
    N = ag.AnchorBoxGenerator(H, W, 3, 7, p.anchor_params).boxes.shape[0]
    logits, deltas = synth(1, N, C, 7, 'dense')
    x = po.FilterTopKDetections(30, True)(po.TransformBoxesAndScores(p)({'class_logits': logits,
                                                                         'encoded_boxes': deltas}))
    try:
        po.GenerateDetections(mode='GlobalHardNMS', soft_nms_sigma=0.5, num_classes=C)(x)
        raise SystemExit('expected the reference to fail on 4-D boxes in a Global* mode')
    except AssertionError:
        pass
    print('wrote {} detect fixtures'.format(len(cases)))


if __name__ == '__main__':
    main()
