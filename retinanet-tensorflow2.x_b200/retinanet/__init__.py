"""Host-side mirror of the reference's `retinanet` package for the detection post-processing path only.

Same module paths, class names, constructor signatures, config keys and dict keys as
srihari-humbarwadi/retinanet-tensorflow2.x (retinanet/model/layers, retinanet/dataloader/anchor_generator.py,
retinanet/model/builder.py, retinanet/cfg/config.py); tensors are torch CUDA tensors and all arithmetic runs in
the hand-written sm_100a kernels of libretinapost.so through the C ABI of include/retinapost.h.
"""
