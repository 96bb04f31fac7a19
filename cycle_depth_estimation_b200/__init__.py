"""B200-native (sm_100a) dense-convolution hot path behind the API of the reference
JosephineRabbit/cycle_depth_estimation (models/networks.py, util/image_pool.py, new_multi/my_eval.py).

Host code is Python/PyTorch; all GPU arithmetic goes through the C ABI of ``lib/libcdb200.so``
(include/cdb200.h) into hand-written CUDA kernels.  There is no CPU path and no cuDNN/Triton
fallback: importing the package works anywhere, calling an op without the built library or
without a B200 raises.
"""
__version__ = "0.1.0"
