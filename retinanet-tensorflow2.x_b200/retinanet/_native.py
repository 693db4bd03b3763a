"""ctypes binding of libretinapost.so (C ABI: include/retinapost.h).

There is no CPU path and no fallback: importing this module without the built library, or calling it without a CUDA
device, raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('RPP_LIB') or os.path.join(os.path.dirname(_HERE), 'libretinapost.so')   # RPP_LIB: tuning builds

MODES = ['CombinedNMS', 'GlobalSoftNMS', 'GlobalHardNMS', 'PerClassSoftNMS', 'PerClassHardNMS']

RPP_OK, RPP_EINVAL, RPP_EMODE, RPP_ECOMBO, RPP_EWORKSPACE, RPP_ECUDA = 0, -1, -2, -3, -4, -5

EXPORTS = [
    'rpp_create', 'rpp_destroy', 'rpp_last_error', 'rpp_num_anchors', 'rpp_num_levels', 'rpp_anchor_boundaries',
    'rpp_anchors', 'rpp_workspace_bytes', 'rpp_decode', 'rpp_topk', 'rpp_topk_levels', 'rpp_nms', 'rpp_detect', 'rpp_detect_levels', 'rpp_detect_typed',
    'rpp_detect_host', 'rpp_detect_host_typed', 'rpp_coco_format', 'rpp_efficient_nms',
    'rpp_last_launch_count', 'rpp_classes_itemsize', 'rpp_debug_force_exact_scan', 'rpp_debug_stage_timing',
    'rpp_debug_stage_ms', 'rpp_debug_stage_report', 'rpp_debug_sample_plan', 'rpp_debug_exact_scans',
]


class RppConfig(ctypes.Structure):
    _fields_ = [
        ('H', ctypes.c_int), ('W', ctypes.c_int),
        ('min_level', ctypes.c_int), ('max_level', ctypes.c_int),
        ('num_classes', ctypes.c_int),
        ('n_areas', ctypes.c_int), ('areas', ctypes.POINTER(ctypes.c_double)),
        ('n_ratios', ctypes.c_int), ('aspect_ratios', ctypes.POINTER(ctypes.c_double)),
        ('n_scales', ctypes.c_int), ('scales', ctypes.POINTER(ctypes.c_double)),
        ('box_variance', ctypes.c_float * 4),
        ('scale_box_targets', ctypes.c_int),
        ('mode', ctypes.c_int),
        ('iou_threshold', ctypes.c_float),
        ('score_threshold', ctypes.c_float),
        ('soft_nms_sigma', ctypes.c_float),
        ('pre_nms_top_k', ctypes.c_int),
        ('filter_per_class', ctypes.c_int),
        ('max_detections', ctypes.c_int),
        ('soft_ignores_iou_threshold', ctypes.c_int),
        ('tpu_semantics', ctypes.c_int),
        ('reserved', ctypes.c_int * 6),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'libretinapost.so is not built ({}). Run `python -c "import __graft_entry__ as g; g.build()"` or '
                '`make -C retinanet-tensorflow2.x_b200/csrc`. There is no CPU fallback.'.format(LIB_PATH))
        L = ctypes.CDLL(LIB_PATH)
        vp, ci, cl, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_float
        L.rpp_create.argtypes = [ctypes.POINTER(RppConfig), ctypes.POINTER(vp)]
        L.rpp_destroy.argtypes = [vp]
        L.rpp_last_error.restype = ctypes.c_char_p
        L.rpp_num_anchors.argtypes = [vp]
        L.rpp_num_anchors.restype = cl
        L.rpp_num_levels.argtypes = [vp]
        L.rpp_anchor_boundaries.argtypes = [vp, ctypes.POINTER(cl)]
        L.rpp_anchors.argtypes = [vp, vp, vp]
        L.rpp_workspace_bytes.argtypes = [vp, ci, cl]
        L.rpp_workspace_bytes.restype = ctypes.c_size_t
        L.rpp_decode.argtypes = [vp, vp, vp, ci, vp, vp, vp]
        L.rpp_topk.argtypes = [vp, vp, vp, ci, cl, vp, vp, vp, vp, ctypes.c_size_t, vp]
        if hasattr(L, 'rpp_topk_levels'):
            L.rpp_topk_levels.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(cl), ci, vp, vp,
                                          vp, vp, ctypes.c_size_t, vp]
        L.rpp_nms.argtypes = [vp, vp, vp, ci, cl, ci, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
        L.rpp_detect.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
        if hasattr(L, 'rpp_detect_levels'):   # absent only in old tuning builds loaded through RPP_LIB
            L.rpp_detect_levels.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ci, vp, vp, vp, vp, vp,
                                            ctypes.c_size_t, vp]
        if hasattr(L, 'rpp_detect_typed'):
            L.rpp_detect_typed.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(vp), ci, ci, vp, vp, vp, vp, vp,
                                           ctypes.c_size_t, vp]
        L.rpp_detect_host.argtypes = [vp, ci, vp, vp, ci, vp, vp, vp, vp]
        if hasattr(L, 'rpp_detect_host_typed'):
            L.rpp_detect_host_typed.argtypes = [vp, ci, vp, vp, ci, ci, vp, vp, vp, vp]
        L.rpp_coco_format.argtypes = [vp] * 5 + [ci] + [vp] * 8
        L.rpp_efficient_nms.argtypes = [vp, vp, vp, vp, ci, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
        L.rpp_classes_itemsize.argtypes = [vp]
        L.rpp_debug_force_exact_scan.argtypes = [vp, ci]
        if hasattr(L, 'rpp_debug_exact_scans'):
            L.rpp_debug_exact_scans.argtypes = [vp, ctypes.POINTER(ctypes.c_ulonglong), ci]
        L.rpp_debug_sample_plan.argtypes = [cl, ci, cl, ci, ctypes.POINTER(ci)]
        L.rpp_debug_stage_timing.argtypes = [vp, ci]
        L.rpp_debug_stage_ms.argtypes = [vp, ctypes.POINTER(cf), ctypes.POINTER(ci)]
        L.rpp_debug_stage_report.argtypes = [vp, ctypes.c_char_p, ci, ctypes.POINTER(ci)]
        _lib = L
    return _lib


def stage_report(handle_ptr):
    """{label: mean ms per call} of the segments recorded since rpp_debug_stage_timing(handle, 1), and the call count."""
    buf = ctypes.create_string_buffer(4096)
    n = ctypes.c_int()
    check(lib().rpp_debug_stage_report(handle_ptr, buf, len(buf), ctypes.byref(n)))
    out = {}
    for item in buf.value.decode().split(';'):
        if item:
            k, v = item.rsplit('=', 1)
            out[k] = float(v)
    return out, int(n.value)


def exact_scans(handle_ptr, reset=True):
    """Problems that left their candidate list for an exact scan of the whole column since the last reset (how often
    the sampled pre-threshold missed; include/retinapost.h rpp_debug_exact_scans).  Synchronises the device."""
    n = ctypes.c_ulonglong()
    check(lib().rpp_debug_exact_scans(handle_ptr, ctypes.byref(n), int(bool(reset))))
    return int(n.value)


def last_error():
    return lib().rpp_last_error().decode('utf-8', 'replace')


def check(rc):
    """Map return codes to the reference's exception types (postprocessing_ops.py:194-197 raises AssertionError)."""
    if rc == RPP_OK:
        return
    msg = last_error()
    if rc == RPP_EMODE:
        raise AssertionError(msg)
    if rc in (RPP_EINVAL, RPP_ECOMBO, RPP_EWORKSPACE):
        raise ValueError(msg)
    raise RuntimeError(msg)
