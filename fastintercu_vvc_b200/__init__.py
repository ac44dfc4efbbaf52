"""fastintercu_vvc_b200 -- B200-native (sm_100a) MLT-CNN inter CU-split predictor.

The product is the C-ABI shared library `libmltcnn.so` (include/mltcnn.h; sources in csrc/), a drop-in
for the libtorch/OpenCV inference block of the reference's EncCu::xCompressCU (EncCu.cpp:803-926).
This package holds that library, the offline weight packer and a thin ctypes binding used by the
tests and bench.py.  There is no CPU or PyTorch fallback: importing `capi` without the built library,
or creating a predictor without a Blackwell GPU, raises.
"""
from .capi import MltCuPredictor, MltError, MltPredictor, MltResult, lib_path, load_library  # noqa: F401
from .pack_weights import pack, write_blob, write_cu_blob  # noqa: F401

__all__ = ["MltPredictor", "MltCuPredictor", "write_cu_blob", "MltResult", "MltError", "load_library", "lib_path", "pack", "write_blob"]
