#!/usr/bin/env python
"""Golden vectors for the OPTIONAL per-level pre-NMS top-k (rpp_topk_levels): the UNMODIFIED reference
FilterTopKDetections (/root/reference/retinanet/model/layers/postprocessing_ops.py:120-173, over tf_shim.py) applied
to every pyramid level's segment of the fused axis — segments = the unmodified AnchorBoxGenerator.anchor_boundaries
(dataloader/anchor_generator.py:42-49) — and concatenated in level order.  The reference has no per-level mode; this
pins the composition rpp_topk_levels is defined as against the reference's own filter.
Run in the build container only:  python tests/golden/make_golden_per_level.py   -> tests/golden/perlevel_*.npz"""
import os

import numpy as np

from make_golden import HERE, load_reference, params_for, synth


def main():
    ag, po = load_reference()
    H = W = 96
    C, B = 4, 2
    p = params_for(H, W, C)
    gen = ag.AnchorBoxGenerator(H, W, 3, 7, p.anchor_params)
    bounds = [int(v) for v in gen.anchor_boundaries]
    N = int(gen.boxes.shape[0])
    assert bounds[0] == 0 and bounds[-1] == N
    n = 0
    for ci, (k, fpc, dist) in enumerate([(k, fpc, dist) for k in (50, 400) for fpc in (True, False)
                                         for dist in ('dense', 'quantized')]):
        logits, deltas = synth(B, N, C, 3000 + ci, dist)
        x = po.TransformBoxesAndScores(p)({'class_logits': logits, 'encoded_boxes': deltas})
        scores, boxes = np.asarray(x['scores']), np.asarray(x['boxes'])
        fs, fb = [], []
        import tf_shim
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            y = po.FilterTopKDetections(top_k=k, filter_per_class=fpc)({'scores': tf_shim.T(scores[:, lo:hi]),
                                                                       'boxes': tf_shim.T(boxes[:, lo:hi])})
            fs.append(np.asarray(y['scores']))
            fb.append(np.asarray(y['boxes']))
        name = 'perlevel_k{}_{}_{}.npz'.format(k, 'pc' if fpc else 'gl', dist)
        np.savez_compressed(os.path.join(HERE, name), H=H, W=W, C=C, k=k, filter_per_class=fpc,
                            boundaries=np.asarray(bounds, np.int64), scores=scores, boxes=boxes,
                            filtered_scores=np.concatenate(fs, axis=1), filtered_boxes=np.concatenate(fb, axis=1))
        n += 1
    print('wrote {} per-level fixtures'.format(n))


if __name__ == '__main__':
    main()
