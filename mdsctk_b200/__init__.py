"""mdsctk_b200 -- B200-native all-pairs distance + kNN stage of MDSCTK (knn_rms / knn_data).

The product is the C-ABI CUDA library ``libmdsctk_knn.so`` (include/mdsctk_knn.h) and the
C++ command-line tools built on it (mdsctk_b200/host).  This Python module is a thin ctypes
binding over the same C ABI, used by the tests and bench.py; it contains no arithmetic and
has no CPU fallback: importing works anywhere, but creating a context without the built
library or without an sm_100 GPU raises.
"""
from .api import (KnnContext, KnnError, knn_data, knn_rms, library_path, load_library,  # noqa: F401
                  RMS_SIMT_FP32, RMS_TC_1XFP16, RMS_TC_1XTF32, RMS_TC_2XFP16, RMS_TC_3XBF16, RMS_TC_3XFP16, RMS_TC_3XTF32)

__all__ = ["KnnContext", "KnnError", "knn_rms", "knn_data", "load_library", "library_path",
           "RMS_SIMT_FP32", "RMS_TC_3XTF32", "RMS_TC_1XTF32", "RMS_TC_3XBF16", "RMS_TC_3XFP16", "RMS_TC_2XFP16", "RMS_TC_1XFP16"]
