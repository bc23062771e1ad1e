"""flowdenoising_b200 -- B200 (sm_100a) implementation of FlowDenoising's optical-flow-driven separable
Gaussian filter (the ``filter_along_Z/Y/X`` hot path of microscopy-processing/FlowDenoising).

    from flowdenoising_b200 import flowdenoising as fd      # drop-in for the reference module surface
    from flowdenoising_b200.engine import DeviceEngine       # device-resident API

The compute path is hand-written CUDA behind a C ABI (include/fdn_b200.h, libfdn_b200.so); there is no CPU
fallback.
"""
__version__ = "0.1.0"
