"""Differential fuzzing: random configurations (mode, filter, sizes, thresholds, logit distribution, ties) through
the fused GPU path against the CPU oracle.  Seeded, so failures reproduce."""
import numpy as np
import pytest

from _util import image_mismatches, make_params, oracle_detect, to_numpy

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

MODES = ['CombinedNMS', 'GlobalSoftNMS', 'GlobalHardNMS', 'PerClassSoftNMS', 'PerClassHardNMS']


def _case(seed):
    rng = np.random.default_rng(seed)
    mode = MODES[rng.integers(len(MODES))]
    H = int(rng.choice([64, 96, 128, 192, 320, 320, 448]))   # >= 320: the sampled pre-threshold path
    W = int(rng.choice([H, H, 64, 160])) if H < 320 else H
    C = int(rng.choice([1, 2, 3, 4, 5, 8, 12, 20, 40])) if H < 448 else int(rng.choice([1, 4, 8, 12]))
    B = int(rng.integers(1, 5))
    M = int(rng.choice([1, 5, 20, 100, 150]))
    k = int(rng.choice([-1, 1, 10, 100, 1000, 5000]))
    fpc = bool(rng.integers(2))
    if mode.startswith('Global'):
        fpc = False
    inf = dict(mode=mode, pre_nms_top_k=k, filter_per_class=fpc, max_detections=M,
               iou_threshold=float(rng.choice([0.0, 0.3, 0.5, 0.75, 1.0])),
               score_threshold=float(rng.choice([0.0, 0.05, 0.3, 0.6])),
               soft_nms_sigma=float(rng.choice([0.1, 0.5, 1.5])))
    dist = str(rng.choice(['dense', 'sparse', 'quantized', 'coarse', 'clustered']))
    return H, W, C, B, inf, dist, rng


def _inputs(rng, B, N, C, dist):
    deltas = np.clip(rng.standard_normal((B, N, 4)) * 0.5, -4, 4).astype(np.float32)
    x = rng.standard_normal((B, N, C))
    if dist == 'sparse':
        x = x * 1.5 - 4.595
    elif dist == 'quantized':
        x = np.round(x * 8) / 8
    elif dist == 'coarse':        # a handful of distinct values: huge tie groups, saturated scores
        x = rng.choice(np.array([-30.0, -3.0, -0.5, 0.0, 0.5, 3.0, 20.0, 40.0]), size=(B, N, C))
    elif dist == 'clustered':     # a few "objects": many anchors with near-identical boxes and high scores
        x = x * 0.5 - 5.0
        hot = rng.integers(0, N, size=(B, 40))
        for b in range(B):
            x[b, hot[b], rng.integers(0, C, 40)] += rng.uniform(6, 10, 40)
        deltas *= 0.05
    return x.astype(np.float32), deltas


@pytest.mark.parametrize('seed', range(int(__import__('os').environ.get('RPP_FUZZ_CASES', '120'))))
def test_fuzz_fused_vs_oracle(ref, seed):
    from retinanet.model.builder import ModelBuilder
    H, W, C, B, inf, dist, rng = _case(1000 + seed)
    p = make_params(H, W, num_classes=C, **inf)
    layer = ModelBuilder(p).add_post_processing_stage(None).layers[-1]
    N = layer.handle(C).num_anchors
    logits, deltas = _inputs(rng, B, N, C, dist)
    got = to_numpy(layer({'class_logits': torch.from_numpy(logits).cuda(),
                          'encoded_boxes': torch.from_numpy(deltas).cuda()}))
    exp = oracle_detect(ref, p, logits, deltas, threads=8)
    assert image_mismatches(got, exp) == [], (H, W, C, B, inf, dist)
