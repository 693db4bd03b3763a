"""Per-level (segmented) pre-NMS top-k — the optional extension named by BASELINE.json's north_star ("per-level top-k
pre-selection", configs[1] "top-1000/level").  The reference has no such mode (its FilterTopKDetections runs over the
fused anchor axis, postprocessing_ops.py:128-161), so the contract is its own filter applied to every segment
[anchor_boundaries[l], anchor_boundaries[l+1]) (dataloader/anchor_generator.py:42-49) and concatenated."""
import numpy as np
import pytest

from _util import to_numpy

BOUNDS = [0, 1296, 1620, 1701, 1728, 1737]   # 96x96 input, levels 3..7, 9 anchors per cell


def _inputs(B, N, C, seed, quantized=False):
    rng = np.random.default_rng(seed)
    s = rng.random((B, N, C)).astype(np.float32)
    if quantized:
        s = (np.round(s * 32) / 32).astype(np.float32)   # many exact ties: index order decides
    b = rng.random((B, N, 4)).astype(np.float32)
    return s, b


@pytest.mark.parametrize('per_class', [True, False])
def test_oracle_per_level_is_the_reference_filter_on_every_segment(ref, per_class):
    s, b = _inputs(2, BOUNDS[-1], 3, 1, quantized=True)
    k = 50   # levels 6 and 7 (27 and 9 rows) are shorter than k: per class they contribute all their rows
    so, bo, io = ref.filter_per_level(s, b, k, BOUNDS, per_class=per_class)
    f = ref.filter_per_class if per_class else ref.filter_global
    off = 0
    for lo, hi in zip(BOUNDS[:-1], BOUNDS[1:]):
        es, eb, ei = f(s[:, lo:hi], b[:, lo:hi], k)
        kk = es.shape[1]
        assert kk == min(k, (hi - lo) if per_class else (hi - lo) * 3)
        assert np.array_equal(so[:, off:off + kk], es) and np.array_equal(bo[:, off:off + kk], eb)
        assert np.array_equal(io[..., off:off + kk], ei + (lo if per_class else lo * 3))
        off += kk
    assert off == so.shape[1]
    # a single segment is the reference's fused filter itself
    fs, fb, fi = ref.filter_per_level(s, b, k, [0, BOUNDS[-1]], per_class=per_class)
    es, eb, ei = f(s, b, k)
    assert np.array_equal(fs, es) and np.array_equal(fb, eb) and np.array_equal(fi, ei)
    # every selected index lies inside its level
    if per_class:
        assert ((io[..., :k] >= 0) & (io[..., :k] < BOUNDS[1])).all()


@pytest.mark.gpu
@pytest.mark.parametrize('per_class', [True, False])
@pytest.mark.parametrize('quantized', [False, True])
def test_gpu_per_level_topk_vs_oracle(ref, per_class, quantized):
    torch = pytest.importorskip('torch')
    from retinanet.model.layers import FilterTopKDetectionsPerLevel
    B, C, k = 3, 4, 200
    s, b = _inputs(B, BOUNDS[-1], C, 7, quantized)
    es, eb, _ = ref.filter_per_level(s, b, k, BOUNDS, per_class=per_class)
    layer = FilterTopKDetectionsPerLevel(k, per_class, anchor_boundaries=BOUNDS)
    got = to_numpy(layer({'scores': torch.from_numpy(s).cuda(), 'boxes': torch.from_numpy(b).cuda()}))
    assert got['scores'].shape == es.shape and got['boxes'].shape == eb.shape
    assert np.array_equal(got['scores'], es) and np.array_equal(got['boxes'], eb)
    # the head outputs' own layout: one tensor per level, read in place (dict keyed by level, as FuseDetections takes)
    sl = {str(3 + i): torch.from_numpy(np.ascontiguousarray(s[:, lo:hi])).cuda()
          for i, (lo, hi) in enumerate(zip(BOUNDS[:-1], BOUNDS[1:]))}
    bl = {str(3 + i): torch.from_numpy(np.ascontiguousarray(b[:, lo:hi])).cuda()
          for i, (lo, hi) in enumerate(zip(BOUNDS[:-1], BOUNDS[1:]))}
    got2 = to_numpy(FilterTopKDetectionsPerLevel(k, per_class)({'scores': sl, 'boxes': bl}))
    assert np.array_equal(got2['scores'], es) and np.array_equal(got2['boxes'], eb)


@pytest.mark.gpu
def test_gpu_per_level_topk_indices_and_full_size(ref):
    """640x640 geometry, top-1000 per level (BASELINE configs[1] wording), per class: indices through the C ABI."""
    import ctypes
    torch = pytest.importorskip('torch')
    from retinanet import _native
    from retinanet.model.layers import FilterTopKDetectionsPerLevel
    bounds = [0, 57600, 72000, 75600, 76500, 76725]
    B, C, k = 2, 8, 1000
    rng = np.random.default_rng(3)
    s = (1.0 / (1.0 + np.exp(-rng.standard_normal((B, bounds[-1], C))))).astype(np.float32)
    b = rng.random((B, bounds[-1], 4)).astype(np.float32)
    es, eb, ei = ref.filter_per_level(s, b, k, bounds, per_class=True, threads=8)
    layer = FilterTopKDetectionsPerLevel(k, True, anchor_boundaries=bounds)
    h = layer._handle(C)
    lv = [(torch.from_numpy(np.ascontiguousarray(s[:, lo:hi])).cuda(),
           torch.from_numpy(np.ascontiguousarray(b[:, lo:hi])).cuda()) for lo, hi in zip(bounds[:-1], bounds[1:])]
    K = es.shape[1]
    assert K == 1000 * 3 + 900 + 225
    so = torch.empty((B, K, C), device='cuda'); bo = torch.empty((B, K, C, 4), device='cuda')
    io = torch.empty((B, C, K), dtype=torch.int32, device='cuda')
    L = len(lv)
    sp = (ctypes.c_void_p * L)(*[x.data_ptr() for x, _ in lv]); bp = (ctypes.c_void_p * L)(*[y.data_ptr() for _, y in lv])
    nr = (ctypes.c_long * L)(*[hi - lo for lo, hi in zip(bounds[:-1], bounds[1:])])
    n_ws = max(nr, key=lambda n: _native.lib().rpp_workspace_bytes(h.ptr, B, n))   # the level that needs the most
    ws = h.workspace(B, n_ws, so.device)
    _native.check(_native.lib().rpp_topk_levels(h.ptr, L, sp, bp, nr, B, so.data_ptr(), bo.data_ptr(), io.data_ptr(),
                                                ws.data_ptr(), ws.numel(), None))
    torch.cuda.synchronize()
    assert np.array_equal(so.cpu().numpy(), es) and np.array_equal(bo.cpu().numpy(), eb)
    assert np.array_equal(io.cpu().numpy(), ei)


# ---- golden vectors: the UNMODIFIED reference FilterTopKDetections run on every anchor_boundaries segment --------------
# (tests/golden/make_golden_per_level.py, executed in the build container over tests/golden/tf_shim.py)
import glob
import os

_GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'perlevel_*.npz')))


def test_per_level_fixture_inventory():
    assert len(_GOLDEN) == 8


@pytest.mark.parametrize('path', _GOLDEN, ids=os.path.basename)
def test_oracle_per_level_vs_reference_filter(ref, path):
    g = np.load(path)
    fs, fb, _ = ref.filter_per_level(g['scores'], g['boxes'], int(g['k']), g['boundaries'].tolist(),
                                     per_class=bool(g['filter_per_class']))
    assert fs.shape == g['filtered_scores'].shape and fb.shape == g['filtered_boxes'].shape
    assert np.array_equal(fs, g['filtered_scores']) and np.array_equal(fb, g['filtered_boxes'])


@pytest.mark.gpu
@pytest.mark.parametrize('path', _GOLDEN, ids=os.path.basename)
def test_gpu_per_level_vs_reference_filter(path):
    torch = pytest.importorskip('torch')
    from retinanet.model.layers import FilterTopKDetectionsPerLevel
    g = np.load(path)
    layer = FilterTopKDetectionsPerLevel(int(g['k']), bool(g['filter_per_class']),
                                         anchor_boundaries=g['boundaries'].tolist())
    got = to_numpy(layer({'scores': torch.from_numpy(g['scores']).cuda(), 'boxes': torch.from_numpy(g['boxes']).cuda()}))
    assert np.array_equal(got['scores'], g['filtered_scores']) and np.array_equal(got['boxes'], g['filtered_boxes'])
