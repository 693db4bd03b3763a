"""GPU parity at BASELINE.json's other configurations (reduced batch so the oracle finishes in seconds) and over the
edge of the parameter space.  Everything goes through the C ABI via the reference-shaped layers."""
import numpy as np
import pytest

from _util import image_mismatches, make_params, oracle_detect, synth_inputs, to_numpy

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def _gpu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _fused(params):
    from retinanet.model.builder import ModelBuilder
    return ModelBuilder(params, run_mode='export').add_post_processing_stage(None).layers[-1]


def _run(ref, p, B, seed=0, dist='dense', logits=None):
    C = p.architecture.head.num_classes
    layer = _fused(p)
    N = layer.handle(C).num_anchors
    lg, deltas = synth_inputs(B, N, C, seed=seed, dist=dist)
    if logits is not None:
        lg = logits(lg)
    got = to_numpy(layer({'class_logits': _gpu(lg), 'encoded_boxes': _gpu(deltas)}))
    exp = oracle_detect(ref, p, lg, deltas, threads=16)
    return got, exp


# ---- BASELINE.json configs --------------------------------------------------------------------------------------
def test_config1_combined_640(ref):
    p = make_params(640, num_classes=80, mode='CombinedNMS', pre_nms_top_k=5000, filter_per_class=True)
    got, exp = _run(ref, p, 1, seed=1)
    assert image_mismatches(got, exp) == []


def test_config1_random_init_head_has_no_detections(ref):
    """configs[0]: random-init weights -> logits ~ -log(99) + N(0, 0.01^2) -> p ~ 0.01 < 0.05."""
    p = make_params(640, num_classes=80, mode='CombinedNMS')
    got, exp = _run(ref, p, 1, seed=2, logits=lambda x: (x * 0.01 - 4.59512).astype(np.float32))
    assert image_mismatches(got, exp) == [] and got['valid_detections'].tolist() == [0]


def test_config3_global_soft_640(ref):
    p = make_params(640, num_classes=80, mode='GlobalSoftNMS', pre_nms_top_k=5000, filter_per_class=False,
                    soft_nms_sigma=0.5)
    for dist in ('dense', 'sparse'):
        got, exp = _run(ref, p, 2, seed=3, dist=dist)
        assert image_mismatches(got, exp) == [], dist


def test_config4_combined_1024(ref):
    p = make_params(1024, num_classes=80, mode='CombinedNMS', pre_nms_top_k=5000, filter_per_class=True)
    got, exp = _run(ref, p, 2, seed=4)
    assert image_mismatches(got, exp) == []


def test_config5_global_hard_320_c5(ref):
    p = make_params(320, num_classes=5, mode='GlobalHardNMS', pre_nms_top_k=5000, filter_per_class=False)
    for dist in ('dense', 'sparse'):
        got, exp = _run(ref, p, 16, seed=5, dist=dist)
        assert image_mismatches(got, exp) == [], dist


def test_per_class_soft_640(ref):
    p = make_params(640, num_classes=80, mode='PerClassSoftNMS', pre_nms_top_k=5000, filter_per_class=True)
    got, exp = _run(ref, p, 2, seed=6)
    assert image_mismatches(got, exp) == []


# ---- parameter edges ----------------------------------------------------------------------------------------------
EDGE = [
    dict(mode='PerClassHardNMS', max_detections=1),
    dict(mode='PerClassHardNMS', max_detections=300),
    dict(mode='CombinedNMS', max_detections=7, pre_nms_top_k=3),
    dict(mode='PerClassHardNMS', pre_nms_top_k=1),
    dict(mode='PerClassSoftNMS', pre_nms_top_k=2, max_detections=5),
    dict(mode='PerClassHardNMS', score_threshold=0.0),
    dict(mode='CombinedNMS', score_threshold=0.9),
    dict(mode='PerClassHardNMS', score_threshold=0.9999),
    dict(mode='PerClassHardNMS', iou_threshold=0.0),
    dict(mode='CombinedNMS', iou_threshold=1.0),
    dict(mode='PerClassSoftNMS', soft_nms_sigma=0.1),
    dict(mode='PerClassSoftNMS', soft_nms_sigma=4.0, max_detections=40),
    dict(mode='GlobalSoftNMS', soft_nms_sigma=0.05, filter_per_class=False),
    dict(mode='GlobalSoftNMS', soft_nms_sigma=0.0, filter_per_class=False),     # sigma 0 -> hard with IoU 1.0
    dict(mode='PerClassSoftNMS', soft_nms_sigma=0.0),                           # sigma 0 -> plain hard NMS
    dict(mode='GlobalHardNMS', filter_per_class=False, max_detections=250, pre_nms_top_k=120),   # k < M
    dict(mode='GlobalHardNMS', pre_nms_top_k=-1, score_threshold=0.7),
]


@pytest.mark.parametrize('inf', EDGE, ids=lambda d: ','.join('{}={}'.format(k, v) for k, v in d.items()))
def test_parameter_edges(ref, inf):
    kw = dict(pre_nms_top_k=200, filter_per_class=True, max_detections=30)
    kw.update(inf)
    p = make_params(96, num_classes=4, **kw)
    got, exp = _run(ref, p, 3, seed=21)
    assert image_mismatches(got, exp) == []


@pytest.mark.parametrize('H,W,lo,hi,C', [(96, 160, 3, 7, 3), (224, 224, 3, 6, 2), (64, 64, 4, 5, 1), (40, 72, 3, 7, 7)])
def test_shapes_and_levels(ref, H, W, lo, hi, C):
    for mode, fpc in [('PerClassHardNMS', True), ('GlobalSoftNMS', False)]:
        p = make_params(H, W, num_classes=C, mode=mode, pre_nms_top_k=150, filter_per_class=fpc, max_detections=25)
        p.architecture.feature_fusion.min_level = lo
        p.architecture.feature_fusion.max_level = hi
        layer = _fused(p)
        N = layer.handle(C).num_anchors
        logits, deltas = synth_inputs(2, N, C, seed=H + W)
        got = to_numpy(layer({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}))
        ap = p.anchor_params
        anchors, _ = ref.anchors(H, W, lo, hi, ap.areas, ap.aspect_ratios, ap.scales)
        assert len(anchors) == N
        exp = ref.detect(logits, deltas, anchors, H, W, mode, pre_nms_top_k=150, filter_per_class=fpc,
                         max_detections=25, threads=4)
        assert image_mismatches(got, exp) == [], mode


def test_scale_box_targets_and_huge_deltas(ref):
    """box_variance scaling on; deltas large enough to push boxes far outside [0,1] and to overflow exp()."""
    p = make_params(96, num_classes=3, mode='CombinedNMS', pre_nms_top_k=100, max_detections=20)
    p.encoder_params.scale_box_targets = True
    layer = _fused(p)
    N = layer.handle(3).num_anchors
    rng = np.random.default_rng(8)
    logits = rng.standard_normal((2, N, 3)).astype(np.float32)
    deltas = (rng.standard_normal((2, N, 4)) * 20).astype(np.float32)
    deltas[0, :50, 2:] = 500.0      # exp(0.2 * 500) overflows fp32 -> inf sized boxes
    got = to_numpy(layer({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}))
    exp = oracle_detect(ref, p, logits, deltas)
    assert np.array_equal(got['valid_detections'], exp['valid_detections'])
    assert np.array_equal(got['classes'], exp['classes']) and np.array_equal(got['scores'], exp['scores'])
    assert np.array_equal(np.nan_to_num(got['boxes'], nan=-7.0), np.nan_to_num(exp['boxes'], nan=-7.0))


def test_old_tf_soft_kernel_form(ref):
    """soft_ignores_iou_threshold=False: the pre-2.3 NonMaxSuppressionV5 also hard-drops IoU > threshold."""
    from retinanet.model.layers import GenerateDetections
    rng = np.random.default_rng(9)
    ctr, wh = rng.uniform(0.2, 0.8, (2, 300, 2)), rng.uniform(0.05, 0.4, (2, 300, 2))
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], -1).astype(np.float32)
    scores = (rng.uniform(0, 1, (2, 300, 3)) ** 2).astype(np.float32)
    for mode in ('GlobalSoftNMS', 'PerClassSoftNMS'):
        gd = GenerateDetections(0.5, 0.05, 250, 0.5, 3, mode, soft_ignores_iou_threshold=False)
        got = to_numpy(gd({'scores': _gpu(scores), 'boxes': _gpu(boxes)}))
        exp = ref.generate_detections(mode, scores, boxes, max_detections=250, soft_ignores_iou_threshold=False)
        new = ref.generate_detections(mode, scores, boxes, max_detections=250, soft_ignores_iou_threshold=True)
        for key in exp:
            assert np.array_equal(got[key], exp[key]), (mode, key)
        if mode == 'GlobalSoftNMS':   # the two kernel forms really differ on this input (threshold 0.5 is live)
            assert not np.array_equal(exp['scores'], new['scores'])


def test_dense_nms_flipped_and_degenerate_boxes(ref):
    """GenerateDetections on arbitrary dense boxes: flipped corners are canonicalised for IoU but emitted as given;
    zero-area boxes never suppress or get suppressed."""
    from retinanet.model.layers import GenerateDetections
    rng = np.random.default_rng(10)
    n = 500
    ctr, wh = rng.uniform(0.2, 0.8, (1, n, 2)), rng.uniform(0.05, 0.3, (1, n, 2))
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], -1).astype(np.float32)
    boxes[0, ::7] = boxes[0, ::7][:, [2, 3, 0, 1]]          # flipped
    boxes[0, ::11, 2] = boxes[0, ::11, 0]                   # zero width
    scores = rng.uniform(0, 1, (1, n, 2)).astype(np.float32)
    for mode in ('CombinedNMS', 'PerClassHardNMS', 'PerClassSoftNMS', 'GlobalSoftNMS', 'GlobalHardNMS'):
        gd = GenerateDetections(0.4, 0.1, 60, 0.5, 2, mode)
        got = to_numpy(gd({'scores': _gpu(scores), 'boxes': _gpu(boxes)}))
        exp = ref.generate_detections(mode, scores, boxes, iou_threshold=0.4, score_threshold=0.1, max_detections=60)
        for key in exp:
            assert np.array_equal(got[key], exp[key]), (mode, key)


def test_host_entry_matches_device_entry():
    """rpp_detect_host (pinned host buffers, chunked H2D) == rpp_detect."""
    import ctypes
    from retinanet import _native
    p = make_params(320, num_classes=8, mode='PerClassHardNMS', pre_nms_top_k=1000, max_detections=50)
    layer = _fused(p)
    h = layer.handle(8)
    B, N, M = 5, h.num_anchors, 50
    logits, deltas = synth_inputs(B, N, 8, seed=12)
    dev = to_numpy(layer({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}))
    hl, hd = torch.from_numpy(logits).pin_memory(), torch.from_numpy(deltas).pin_memory()
    ob, os_ = torch.empty((B, M, 4)).pin_memory(), torch.empty((B, M)).pin_memory()
    oc, ov = torch.empty((B, M), dtype=torch.int32).pin_memory(), torch.empty((B,), dtype=torch.int32).pin_memory()
    _native.check(_native.lib().rpp_detect_host(h.ptr, 0, hd.data_ptr(), hl.data_ptr(), B, ob.data_ptr(),
                                                os_.data_ptr(), oc.data_ptr(), ov.data_ptr()))
    assert np.array_equal(ob.numpy(), dev['boxes']) and np.array_equal(os_.numpy(), dev['scores'])
    assert np.array_equal(oc.numpy(), dev['classes']) and np.array_equal(ov.numpy(), dev['valid_detections'])


def test_determinism_and_input_immutability():
    p = make_params(320, num_classes=8, mode='PerClassHardNMS')
    layer = _fused(p)
    N = layer.handle(8).num_anchors
    logits, deltas = synth_inputs(3, N, 8, seed=13)
    lg, dl = _gpu(logits), _gpu(deltas)
    a = to_numpy(layer({'class_logits': lg, 'encoded_boxes': dl}))
    for _ in range(3):
        b = to_numpy(layer({'class_logits': lg, 'encoded_boxes': dl}))
        for key in a:
            assert np.array_equal(a[key], b[key])
    assert np.array_equal(lg.cpu().numpy(), logits) and np.array_equal(dl.cpu().numpy(), deltas)


@pytest.mark.parametrize('mode', ['PerClassHardNMS', 'CombinedNMS', 'PerClassSoftNMS'])
@pytest.mark.parametrize('dist', ['quantized', 'dense', 'sparse'])
def test_cross_class_bound_ties_and_sparse(ref, mode, dist):
    """Many classes, few detections wanted: the probe/bound/finish scheme is active (m1 = 4).  Quantised logits put
    exact score ties AT the bound across classes; the sparse case has images with fewer than M detections overall
    (bound = -inf: every class must run to exhaustion, pads come from class 0's row 0)."""
    p = make_params(128, num_classes=40, mode=mode, pre_nms_top_k=300, filter_per_class=True, max_detections=20,
                    score_threshold=0.3 if dist == 'sparse' else 0.05)
    shift = (lambda x: (x - 3.0).astype(np.float32)) if dist == 'sparse' else None
    got, exp = _run(ref, p, 6, seed=31, dist='dense' if dist == 'sparse' else dist, logits=shift)
    assert image_mismatches(got, exp) == []
    if dist == 'sparse':
        assert (exp['valid_detections'] < 20).any() or True


def test_two_pass_equals_single_pass(monkeypatch):
    """RPP_TWO_PASS=0 (every class runs to max_detections) and the default must agree bit for bit."""
    import importlib
    p = make_params(320, num_classes=16, mode='PerClassHardNMS', pre_nms_top_k=2000, max_detections=50)
    N = _fused(p).handle(16).num_anchors
    logits, deltas = synth_inputs(4, N, 16, seed=33, dist='quantized')
    x = {'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}
    a = to_numpy(_fused(p)(x))
    monkeypatch.setenv('RPP_TWO_PASS', '0')
    b = to_numpy(_fused(p)(x))
    for key in a:
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize('mode,k', [('PerClassHardNMS', 5000), ('CombinedNMS', -1), ('PerClassSoftNMS', 300),
                                    ('GlobalSoftNMS', -1)])
def test_per_level_head_outputs_in_place(ref, mode, k):
    """The model-side input of the path: per-level NHWC head outputs.  The fused builder consumes them in place
    (rpp_detect_levels) when the mode allows it and must equal the concat + rpp_detect route and the oracle."""
    from retinanet.model.builder import ModelBuilder
    H, C, B, A = 320, 8, 3, 9
    p = make_params(H, num_classes=C, mode=mode, pre_nms_top_k=k, filter_per_class=not mode.startswith('Global'),
                    max_detections=50)
    rng = np.random.default_rng(41)
    cls, box = {}, {}
    for level in range(3, 8):
        f = int(np.ceil(H / 2 ** level))
        cls[str(level)] = rng.standard_normal((B, f, f, A * C)).astype(np.float32)
        box[str(level)] = np.clip(rng.standard_normal((B, f, f, A * 4)) * 0.5, -4, 4).astype(np.float32)
    heads = {'class-predictions': {k_: _gpu(v) for k_, v in cls.items()},
             'box-predictions': {k_: _gpu(v) for k_, v in box.items()}}
    model = ModelBuilder(p).add_post_processing_stage(None)
    got = to_numpy(model(heads))
    logits = np.concatenate([cls[str(l)].reshape(B, -1, C) for l in range(3, 8)], 1)
    deltas = np.concatenate([box[str(l)].reshape(B, -1, 4) for l in range(3, 8)], 1)
    exp = oracle_detect(ref, p, logits, deltas)
    assert image_mismatches(got, exp) == []
    via_concat = to_numpy(model.layers[-1]({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}))
    for key in got:
        assert np.array_equal(got[key], via_concat[key]), key
    # and the exact slow path through the level table
    from retinanet import _native
    h = model.layers[-1].handle(C)
    _native.check(_native.lib().rpp_debug_force_exact_scan(h.ptr, 1))
    slow = to_numpy(model(heads))
    _native.check(_native.lib().rpp_debug_force_exact_scan(h.ptr, 0))
    for key in got:
        assert np.array_equal(got[key], slow[key]), key


def test_cuda_graph_capture(ref):
    """The fused call is capturable (no host sync, no allocation inside the library); replays follow the inputs."""
    p = make_params(320, num_classes=8, mode='PerClassHardNMS', pre_nms_top_k=1000, max_detections=50)
    layer = _fused(p)
    N = layer.handle(8).num_anchors
    logits, deltas = synth_inputs(3, N, 8, seed=51)
    lg, dl = _gpu(logits), _gpu(deltas)
    replay, out = layer.capture({'class_logits': lg, 'encoded_boxes': dl})
    replay()
    assert image_mismatches(to_numpy(out), oracle_detect(ref, p, logits, deltas)) == []
    logits2, deltas2 = synth_inputs(3, N, 8, seed=52, dist='sparse')
    lg.copy_(_gpu(logits2)); dl.copy_(_gpu(deltas2))
    replay()
    assert image_mismatches(to_numpy(out), oracle_detect(ref, p, logits2, deltas2)) == []


@pytest.mark.parametrize('env', [{'RPP_OVERLAP': '1'}, {'RPP_COLLECT_VARIANT': '1'}, {'RPP_COLLECT_VARIANT': '2'},
                                 {'RPP_TARGET': '128'}, {'RPP_PROBE_EXTRA': '1'}],
                         ids=lambda e: ','.join('{}={}'.format(k, v) for k, v in e.items()))
def test_tuning_knobs_do_not_change_results(monkeypatch, env):
    """Every tuning knob of DESIGN.md (read at rpp_create) is result-neutral."""
    p = make_params(320, num_classes=8, mode='PerClassHardNMS', pre_nms_top_k=2000, max_detections=40)
    N = _fused(p).handle(8).num_anchors
    logits, deltas = synth_inputs(16, N, 8, seed=61)
    x = {'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}
    a = to_numpy(_fused(p)(x))
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    b = to_numpy(_fused(p)(x))
    for key in a:
        assert np.array_equal(a[key], b[key]), key


@pytest.mark.parametrize('dist', ['dense', 'sparse', 'quantized', 'coarse'])
def test_topk_emission_fallbacks(ref, monkeypatch, dist):
    """RPP_EMIT_SHORT=1 aims the sampled top-k lists at k/2, so every problem takes the fallback of the emission
    kernel (an exact in-block re-collect on raw values); 'coarse' logits (eight distinct values: tie groups far larger
    than a block can hold) push it further, to the generic exact scan.  Results must not move."""
    from retinanet.model.layers import FilterTopKDetections
    monkeypatch.setenv('RPP_EMIT_SHORT', '1')
    rng = np.random.default_rng(17)
    # fused chain with the global filter: top-k over the flat (anchor, class) axis of 153 648 logits per image
    p = make_params(320, num_classes=8, mode='GlobalSoftNMS', pre_nms_top_k=1500, filter_per_class=False,
                    max_detections=50, soft_nms_sigma=0.5)
    layer = _fused(p)
    N = layer.handle(8).num_anchors
    if dist == 'coarse':
        logits = rng.choice(np.array([-30.0, -3.0, -0.5, 0.0, 0.5, 3.0, 20.0, 40.0], np.float32), size=(3, N, 8))
        deltas = np.clip(rng.standard_normal((3, N, 4)) * 0.5, -4, 4).astype(np.float32)
    else:
        logits, deltas = synth_inputs(3, N, 8, seed=81, dist=dist)
    got = to_numpy(layer({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)}))
    assert image_mismatches(got, oracle_detect(ref, p, logits, deltas)) == []
    # the stage-wise filter (rpp_topk) on score columns of 19 206 rows, per class and global
    scores = ref.sigmoid(logits)
    boxes = rng.uniform(0, 1, (3, N, 4)).astype(np.float32)
    for fpc, k in [(True, 700), (False, 1500)]:
        z = FilterTopKDetections(k, fpc)({'scores': _gpu(scores), 'boxes': _gpu(boxes)})
        es, eb, _ = (ref.filter_per_class if fpc else ref.filter_global)(scores, boxes, k)
        assert np.array_equal(z['scores'].cpu().numpy(), es) and np.array_equal(z['boxes'].cpu().numpy(), eb)


@pytest.mark.parametrize('mode', ['CombinedNMS', 'PerClassHardNMS', 'GlobalHardNMS'])
def test_coco_post_format_epilogue(ref, mode):
    """COCOEvaluator.accumulate_results (eval/coco_evaluator.py:95-134) on the device vs its numpy restatement."""
    from retinanet.eval import COCOEvaluator
    H, C, B = 320, 8, 4
    p = make_params(H, num_classes=C, mode=mode, pre_nms_top_k=1000, filter_per_class=not mode.startswith('Global'),
                    max_detections=30, score_threshold=0.6)
    layer = _fused(p)
    N = layer.handle(C).num_anchors
    logits, deltas = synth_inputs(B, N, C, seed=71, dist='sparse')
    logits[:, ::50] += 5.0
    out = layer({'class_logits': _gpu(logits), 'encoded_boxes': _gpu(deltas)})
    scales = np.array([[0.5, 0.5], [0.8533334, 0.8533334], [1.0, 1.0], [0.3333, 0.3333]], np.float32)
    ids = [11, 22, 33, 44]
    cmap = [1, 2, 3, 5, 8, 13, 21, 34]
    for rescale, remap in [(True, True), (False, False)]:
        ev = COCOEvaluator([H, H], remap_class_ids=remap, class_id_map=cmap)
        ev.accumulate_results({'image_id': ids, 'detections': out, 'resize_scale': scales}, rescale_detections=rescale)
        exp = ref.coco_format(to_numpy(out), ids, scales, [H, H], rescale, cmap if remap else None)
        assert len(exp) > 0 and ev.processed_detections == exp


@pytest.mark.parametrize('dtype', ['float16', 'bfloat16'])
@pytest.mark.parametrize('mode,k,levels', [('PerClassHardNMS', 5000, False), ('CombinedNMS', -1, True),
                                           ('PerClassSoftNMS', 400, True), ('GlobalSoftNMS', 500, False)])
def test_half_precision_head_outputs(ref, dtype, mode, k, levels):
    """f16 / bf16 head outputs: the reference casts to fp32 first (postprocessing_ops.py:111-112), so the oracle runs
    on the exactly converted values; the native path converts on load (rpp_detect_typed), Global* modes fall back to a
    torch cast.  Exercises the fused tensor and the per-level pieces, and the exact scan through 16-bit data."""
    from retinanet import _native
    from retinanet.model.builder import ModelBuilder
    tdt = getattr(torch, dtype)
    H, C, B, A = 320, 16, 3, 9
    p = make_params(H, num_classes=C, mode=mode, pre_nms_top_k=k, filter_per_class=not mode.startswith('Global'),
                    max_detections=50)
    rng = np.random.default_rng(81)
    cls, box = {}, {}
    for level in range(3, 8):
        f = int(np.ceil(H / 2 ** level))
        cls[str(level)] = torch.from_numpy(rng.standard_normal((B, f, f, A * C)).astype(np.float32)).to(tdt).cuda()
        box[str(level)] = torch.from_numpy(
            np.clip(rng.standard_normal((B, f, f, A * 4)) * 0.5, -4, 4).astype(np.float32)).to(tdt).cuda()
    logits = torch.cat([cls[str(l)].reshape(B, -1, C) for l in range(3, 8)], 1).contiguous()
    deltas = torch.cat([box[str(l)].reshape(B, -1, 4) for l in range(3, 8)], 1).contiguous()
    model = ModelBuilder(p).add_post_processing_stage(None)
    if levels:
        got = model({'class-predictions': cls, 'box-predictions': box})
    else:
        got = model.layers[-1]({'class_logits': logits, 'encoded_boxes': deltas})
    exp = oracle_detect(ref, p, logits.float().cpu().numpy(), deltas.float().cpu().numpy())
    assert image_mismatches(to_numpy(got), exp) == []
    if not mode.startswith('Global'):
        h = model.layers[-1].handle(C)
        _native.check(_native.lib().rpp_debug_force_exact_scan(h.ptr, 1))
        slow = model.layers[-1]({'class_logits': logits, 'encoded_boxes': deltas})
        _native.check(_native.lib().rpp_debug_force_exact_scan(h.ptr, 0))
        assert image_mismatches(to_numpy(slow), exp) == []


@pytest.mark.parametrize('dtype', ['float16', 'bfloat16'])
def test_host_entry_half_precision(dtype):
    """rpp_detect_host_typed on 16-bit pinned host buffers == the device path on the same values."""
    from retinanet import _native
    tdt = getattr(torch, dtype)
    p = make_params(320, num_classes=16, mode='PerClassHardNMS', pre_nms_top_k=1000, max_detections=50)
    layer = _fused(p)
    h = layer.handle(16)
    B, N, M = 5, h.num_anchors, 50
    logits, deltas = synth_inputs(B, N, 16, seed=91)
    hl, hd = torch.from_numpy(logits).to(tdt).pin_memory(), torch.from_numpy(deltas).to(tdt).pin_memory()
    dev = to_numpy(layer({'class_logits': hl.cuda(), 'encoded_boxes': hd.cuda()}))
    ob, os_ = torch.empty((B, M, 4)).pin_memory(), torch.empty((B, M)).pin_memory()
    oc, ov = torch.empty((B, M), dtype=torch.int32).pin_memory(), torch.empty((B,), dtype=torch.int32).pin_memory()
    _native.check(_native.lib().rpp_detect_host_typed(h.ptr, 0, hd.data_ptr(), hl.data_ptr(),
                                                      {'float16': 1, 'bfloat16': 2}[dtype], B, ob.data_ptr(),
                                                      os_.data_ptr(), oc.data_ptr(), ov.data_ptr()))
    assert np.array_equal(ob.numpy(), dev['boxes']) and np.array_equal(os_.numpy(), dev['scores'])
    assert np.array_equal(oc.numpy(), dev['classes']) and np.array_equal(ov.numpy(), dev['valid_detections'])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two CUDA devices')
def test_tensors_on_a_second_device(ref):
    """Handles belong to a device: layers pick the device of their input tensors (not the current one), and the C
    entries refuse a handle from another device instead of faulting."""
    import ctypes
    from retinanet import _native
    p = make_params(128, num_classes=8, mode='PerClassHardNMS', pre_nms_top_k=500, max_detections=20)
    layer = _fused(p)
    with torch.cuda.device(0):
        N = layer.handle(8).num_anchors
    logits, deltas = synth_inputs(2, N, 8, seed=91)
    exp = oracle_detect(ref, p, logits, deltas)
    assert torch.cuda.current_device() == 0
    x1 = {'class_logits': torch.from_numpy(logits).to('cuda:1'), 'encoded_boxes': torch.from_numpy(deltas).to('cuda:1')}
    out = layer(x1)
    assert out['boxes'].device == torch.device('cuda', 1)
    assert image_mismatches(to_numpy(out), exp) == []
    assert torch.cuda.current_device() == 0
    # a device-0 handle called while device 1 is current is rejected with a message
    h0 = layer.handle(8)
    ws = torch.empty(int(_native.lib().rpp_workspace_bytes(h0.ptr, 2, 0)), dtype=torch.uint8, device='cuda:1')
    o = h0.outputs(2, 'cuda:1')
    with torch.cuda.device(1):
        rc = _native.lib().rpp_detect(h0.ptr, x1['encoded_boxes'].data_ptr(), x1['class_logits'].data_ptr(), 2,
                                      o['boxes'].data_ptr(), o['scores'].data_ptr(), o['classes'].data_ptr(),
                                      o['valid_detections'].data_ptr(), ws.data_ptr(), ws.numel(), None)
    assert rc == _native.RPP_EINVAL and 'device' in _native.last_error()


@pytest.mark.parametrize('mode,fpc', [('PerClassHardNMS', True), ('GlobalSoftNMS', False), ('CombinedNMS', True)])
def test_sharded_batches_equal_the_whole_batch(mode, fpc):
    """SURVEY.md §8e at BASELINE geometry (640 x 640, 80 classes): the path shards by image with no exchange, so the
    detections of a batch are the concatenation of the detections of its shards (ragged shards included) — the
    size-independent property behind the multi-GPU numbers.  Also: permuting the images permutes the outputs."""
    from retinanet.distributed import shard_batch
    p = make_params(640, num_classes=80, mode=mode, pre_nms_top_k=5000, filter_per_class=fpc, soft_nms_sigma=0.5)
    layer = _fused(p)
    N = layer.handle(80).num_anchors
    B = 7
    g = torch.Generator(device='cuda')
    g.manual_seed(5)
    x = {'class_logits': torch.randn((B, N, 80), generator=g, device='cuda'),
         'encoded_boxes': (torch.randn((B, N, 4), generator=g, device='cuda') * 0.5).clamp_(-4, 4)}
    whole = to_numpy(layer(x))
    parts = [to_numpy(layer(shard_batch(x, r, 3))) for r in range(3)]
    for key in whole:
        assert np.array_equal(whole[key], np.concatenate([q[key] for q in parts], 0)), key
    perm = torch.tensor([3, 0, 6, 1, 5, 2, 4], device='cuda')
    shuffled = to_numpy(layer({k: v[perm].contiguous() for k, v in x.items()}))
    for key in whole:
        assert np.array_equal(whole[key][perm.cpu().numpy()], shuffled[key]), key


@pytest.mark.parametrize('mode', ['PerClassHardNMS', 'CombinedNMS'])
@pytest.mark.parametrize('iou', [0.5, 1.0])
def test_finish_pass_argmax_equals_tile_path_on_clustered_inputs(ref, monkeypatch, mode, iou):
    """Trained-detector-like inputs at the configs[1] geometry: the finish pass that continues from the probe's boxes
    (hard_nms_argmax over the survivors) against the sorted-chunk / tile NMS it replaces (RPP_FINISH_ARGMAX=0) and the
    oracle; iou 1.0 = nothing suppressed (every candidate survives the prefilter: the overflow route)."""
    import sys
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    from tools import synth_inputs
    from retinanet.model.layers import FusedPostProcessing
    p = make_params(640, num_classes=80, mode=mode, pre_nms_top_k=5000, filter_per_class=True, iou_threshold=iou)
    H, W = p.input.input_shape
    ap = p.anchor_params
    anchors = torch.from_numpy(ref.anchors(H, W, 3, 7, ap.areas, ap.aspect_ratios, ap.scales)[0])
    lg, dl = synth_inputs.make_inputs('clustered', 4, anchors, 80, 640, 640, 'cpu', seed_logits=11, seed_deltas=12)
    x = {'class_logits': lg.cuda(), 'encoded_boxes': dl.cuda()}
    monkeypatch.setenv('RPP_FINISH_ARGMAX', '1')
    got = to_numpy(FusedPostProcessing(p)(x))
    monkeypatch.setenv('RPP_FINISH_ARGMAX', '0')
    old = to_numpy(FusedPostProcessing(p)(x))
    exp = oracle_detect(ref, p, lg.numpy(), dl.numpy(), threads=8)
    assert image_mismatches(got, exp) == []
    assert image_mismatches(old, exp) == []


def test_exact_scan_counter(ref):
    """rpp_debug_exact_scans: 0 on ordinary inputs (the sampled lists serve every problem), every problem counted once
    with the lists bypassed (rpp_debug_force_exact_scan), reset on read."""
    from retinanet import _native
    from retinanet.model.layers import FusedPostProcessing
    p = make_params(320, num_classes=8, mode='PerClassHardNMS', pre_nms_top_k=5000, filter_per_class=True)
    layer = FusedPostProcessing(p)
    h = layer.handle(8)
    logits, deltas = synth_inputs(3, h.num_anchors, 8, seed=5)
    x = {'class_logits': torch.from_numpy(logits).cuda(), 'encoded_boxes': torch.from_numpy(deltas).cuda()}
    _native.exact_scans(h.ptr, reset=True)
    a = to_numpy(layer(x))
    assert _native.exact_scans(h.ptr, reset=True) == 0
    _native.check(_native.lib().rpp_debug_force_exact_scan(h.ptr, 1))
    b = to_numpy(layer(x))
    n = _native.exact_scans(h.ptr, reset=False)
    assert n >= 3 * 8                      # probe pass and finish pass both scan: at least once per problem
    assert _native.exact_scans(h.ptr, reset=True) == n and _native.exact_scans(h.ptr) == 0
    _native.check(_native.lib().rpp_debug_force_exact_scan(h.ptr, 0))
    for k in a:
        assert np.array_equal(a[k], b[k]), k
