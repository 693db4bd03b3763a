import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'retinanet-tensorflow2.x_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE_CONFIG = {
    'input': {'input_shape': [640, 640], 'channels': 3},
    'architecture': {'feature_fusion': {'min_level': 3, 'max_level': 7},
                     'head': {'num_classes': 80, 'num_anchors': 9}},
    'anchor_params': {'areas': [1024.0, 4096.0, 16384.0, 65536.0, 262144.0],
                      'aspect_ratios': [0.5, 1.0, 2.0],
                      'scales': [1, 1.2599210498948732, 1.5874010519681994]},
    'encoder_params': {'box_variance': [0.1, 0.1, 0.2, 0.2], 'scale_box_targets': False},
    'inference': {'batch_size': 1, 'mode': 'PerClassHardNMS', 'iou_threshold': 0.5, 'score_threshold': 0.05,
                  'soft_nms_sigma': 0.5, 'pre_nms_top_k': 5000, 'filter_per_class': True, 'max_detections': 100},
}


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def ref():
    from oracle import ref as _ref
    _ref.build()
    return _ref
