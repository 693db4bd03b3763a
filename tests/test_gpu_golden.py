"""GPU path against the committed golden fixtures (outputs of the unmodified reference modules executed over the
numpy TensorFlow stand-in, tests/golden/make_golden.py).  Both the fused entry (rpp_detect) and the reference's
layer-by-layer graph (rpp_decode -> rpp_topk -> rpp_nms) are checked.  Needs no /root/reference at run time."""
import glob
import hashlib
import os

import numpy as np
import pytest

from _util import make_params, to_numpy

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DETECT = sorted(glob.glob(os.path.join(GOLDEN, 'detect_*.npz')))
ANCHORS = sorted(glob.glob(os.path.join(GOLDEN, 'anchors_*.npz')))


def _gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize('path', ANCHORS, ids=os.path.basename)
def test_anchor_kernel_vs_reference(path):
    from retinanet.dataloader.anchor_generator import AnchorBoxGenerator
    g = np.load(path)
    p = make_params(64)
    gen = AnchorBoxGenerator(int(g['H']), int(g['W']), int(g['min_level']), int(g['max_level']), p.anchor_params)
    boxes = gen.boxes.cpu().numpy()
    assert gen.anchor_boundaries == g['boundaries'].tolist()
    assert hashlib.sha256(boxes.tobytes()).hexdigest() == str(g['sha256'])


def _check(got, g, box_exact):
    assert got['classes'].dtype == g['out_classes'].dtype
    assert np.array_equal(got['valid_detections'], g['out_valid'])
    assert np.array_equal(got['classes'], g['out_classes'])
    assert np.array_equal(got['scores'], g['out_scores'])
    if box_exact:
        assert np.array_equal(got['boxes'], g['out_boxes'])
    else:
        np.testing.assert_allclose(got['boxes'], g['out_boxes'], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('path', DETECT, ids=os.path.basename)
def test_fused_and_stagewise_vs_reference(path):
    from retinanet.model.builder import ModelBuilder
    from retinanet.model.layers import FilterTopKDetections, GenerateDetections
    g = np.load(path)
    H, C, M, k = int(g['H']), int(g['C']), int(g['M']), int(g['k'])
    mode, fpc = str(g['mode']), bool(g['filter_per_class'])
    p = make_params(H, num_classes=C, mode=mode, pre_nms_top_k=k, filter_per_class=fpc, max_detections=M)
    p.encoder_params.scale_box_targets = bool(g['scale_box_targets'])
    x = {'class_logits': _gpu(g['logits']), 'encoded_boxes': _gpu(g['deltas'])}
    # fused
    fused = ModelBuilder(p).add_post_processing_stage(None).layers[-1]
    _check(to_numpy(fused(x)), g, box_exact=False)
    # layer by layer
    stages = ModelBuilder(p).add_post_processing_stage(None, fused=False).layers[1:]
    y = stages[0](x)
    np.testing.assert_allclose(y['scores'].cpu().numpy(), g['scores'], rtol=1e-5, atol=1e-37)
    np.testing.assert_allclose(y['boxes'].cpu().numpy(), g['boxes'], rtol=1e-5, atol=1e-6)
    for layer in stages[1:]:
        y = layer(y)
    _check(to_numpy(y), g, box_exact=False)
    # stage 2 on IDENTICAL inputs (the reference's own decoded tensors): bit-exact
    z = {'scores': _gpu(g['scores']), 'boxes': _gpu(g['boxes'])}
    if k > 0:
        z = FilterTopKDetections(k, fpc)(z)
        assert np.array_equal(z['scores'].cpu().numpy(), g['filtered_scores'])
        assert np.array_equal(z['boxes'].cpu().numpy(), g['filtered_boxes'])
    det = GenerateDetections(0.5, 0.05, M, 0.5, C, mode)(z)
    _check(to_numpy(det), g, box_exact=True)
