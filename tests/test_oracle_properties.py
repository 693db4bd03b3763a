"""Property tests of the oracle (hypothesis): invariants of the reference's NMS semantics that hold at any size and
that the parity tests rely on implicitly.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st


def _case(seed, n, C, clustered):
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.2, 0.8, (n, 2)) if not clustered else 0.5 + 0.08 * rng.standard_normal((n, 2))
    wh = rng.uniform(0.05, 0.4, (n, 2))
    boxes = np.clip(np.concatenate([c - wh / 2, c + wh / 2], 1), 0, 1).astype(np.float32)
    scores = rng.uniform(0, 1, (n, C)).astype(np.float32)
    return boxes.reshape(1, n, 4), scores.reshape(1, n, C)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 120), C=st.integers(1, 4), M=st.integers(1, 30),
       clustered=st.booleans(), thr=st.sampled_from([0.3, 0.5, 0.75]))
def test_per_class_hard_invariants(ref, seed, n, C, M, clustered, thr):
    boxes, scores = _case(seed, n, C, clustered)
    out = ref.generate_detections('PerClassHardNMS', scores, boxes, thr, 0.05, M)
    v = int(out['valid_detections'][0])
    s, b, c = out['scores'][0], out['boxes'][0], out['classes'][0]
    assert 0 <= v <= M and (s[v:] == -1).all() and (c[v:] == -1).all()
    assert (np.diff(s[:v]) <= 0).all() and (s[:v] > 0.05).all()
    for i in range(v):                      # survivors of one class never overlap by more than the threshold
        for j in range(i):
            if c[i] == c[j]:
                assert ref.iou(b[i], b[j]) <= thr
    # every output row is an input (box, score, class) triple
    for i in range(v):
        hit = np.flatnonzero((boxes[0] == b[i]).all(1) & (scores[0, :, c[i]] == s[i]))
        assert hit.size > 0
    # a larger max_detections only appends: the first M detections do not change
    more = ref.generate_detections('PerClassHardNMS', scores, boxes, thr, 0.05, M + 7)
    vm = int(more['valid_detections'][0])
    assert vm >= v and np.array_equal(more['scores'][0][:v], s[:v]) and np.array_equal(more['boxes'][0][:v], b[:v])
    # IoU is invariant under a power-of-two scaling of the boxes: same detections, scaled
    half = ref.generate_detections('PerClassHardNMS', scores, boxes * np.float32(0.5), thr, 0.05, M)
    assert np.array_equal(half['scores'], out['scores']) and np.array_equal(half['classes'], out['classes'])
    assert np.array_equal(half['boxes'][0][:v], b[:v] * np.float32(0.5))


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 100), M=st.integers(1, 25), clustered=st.booleans())
def test_soft_nms_invariants(ref, seed, n, M, clustered):
    boxes, scores = _case(seed, n, 1, clustered)
    out = ref.generate_detections('GlobalSoftNMS', scores, boxes, 0.5, 0.05, M, soft_nms_sigma=0.5)
    v = int(out['valid_detections'][0])
    s = out['scores'][0]
    assert (np.diff(s[:v]) <= 0).all() and (s[:v] > 0.05).all() and (s[v:] == -1).all()
    # decayed scores never exceed the original score of the box they belong to
    for i in range(v):
        rows = np.flatnonzero((boxes[0] == out['boxes'][0][i]).all(1))
        assert rows.size > 0 and s[i] <= scores[0, rows, 0].max()
    # with a huge sigma the decay vanishes and soft NMS degenerates to "top-M above the threshold"
    flat = ref.generate_detections('GlobalSoftNMS', scores, boxes, 0.5, 0.05, M, soft_nms_sigma=1e9)
    top = np.sort(scores[0, :, 0][scores[0, :, 0] > 0.05])[::-1][:M]
    assert np.allclose(flat['scores'][0][:len(top)], top, rtol=1e-6)


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 150), M=st.integers(1, 40), clustered=st.booleans())
def test_padded_nms_is_the_greedy_scan(ref, seed, n, M, clustered):
    boxes, scores = _case(seed, n, 1, clustered)
    b, s = boxes[0], scores[0, :, 0]
    idx, valid = ref.nms_padded(b, s, M, 0.5, 0.05)
    order = sorted([i for i in range(n) if s[i] > 0.05], key=lambda i: (-s[i], i))
    kept = []
    for i in order:
        if len(kept) >= M:
            break
        if (b[i] > 0).any() and all(ref.iou_padded(b[j], b[i]) < 0.5 for j in kept):
            kept.append(i)
    assert valid == len(kept) and idx[:valid].tolist() == kept and (idx[valid:] == 0).all()


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10 ** 6), C=st.integers(1, 5), k=st.integers(1, 40), per_class=st.booleans(),
       cuts=st.lists(st.integers(1, 59), min_size=0, max_size=4, unique=True), ties=st.booleans())
def test_per_level_filter_invariants(ref, seed, C, k, per_class, cuts, ties):
    """rpp_topk_levels' definition (ref.filter_per_level): the reference filter per anchor_boundaries segment.  Every
    segment contributes min(k, its size) rows taken from inside it, in the filter's own order; one segment = the fused
    filter; splitting a segment never loses a row the fused filter keeps among that segment's first k."""
    rng = np.random.default_rng(seed)
    n = 60
    s = rng.random((2, n, C)).astype(np.float32)
    if ties:
        s = (np.round(s * 8) / 8).astype(np.float32)
    b = rng.random((2, n, 4)).astype(np.float32)
    bounds = [0] + sorted(cuts) + [n]
    fs, fb, fi = ref.filter_per_level(s, b, k, bounds, per_class=per_class)
    off = 0
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        kk = min(k, (hi - lo) if per_class else (hi - lo) * C)
        idx = fi[..., off:off + kk]
        anchor = idx if per_class else idx // C
        assert ((anchor >= lo) & (anchor < hi)).all()
        seg = fs[:, off:off + kk]
        if per_class:      # per class: scores of a segment are sorted descending and are the segment's own top-kk
            assert (np.diff(seg, axis=1) <= 0).all()
            top = -np.sort(-s[:, lo:hi], axis=1)[:, :kk]
            assert np.array_equal(seg, top)
        off += kk
    assert off == fs.shape[1] == fb.shape[1]
    one_s, one_b, one_i = ref.filter_per_level(s, b, k, [0, n], per_class=per_class)
    f = ref.filter_per_class if per_class else ref.filter_global
    es, eb, ei = f(s, b, k)
    assert np.array_equal(one_s, es) and np.array_equal(one_b, eb) and np.array_equal(one_i, ei)
