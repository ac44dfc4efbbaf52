"""ctypes view of oracle/libmltcnn_oracle.so (CPU oracle; test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "libmltcnn_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        L = C.CDLL(so)
        L.mlto_load.restype = C.c_void_p
        L.mlto_load.argtypes = [C.c_char_p]
        L.mlto_free.argtypes = [C.c_void_p]
        L.mlto_stage.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.mlto_forward_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.mlto_predict.restype = C.c_int
        L.mlto_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.mlto_predict_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.mlto_cu_stage.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.mlto_cu_forward.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.mlto_cu_predict_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.mlto_picture_ctus.restype = C.c_int
        L.mlto_picture_ctus.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.mlto_picture_pred.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.mlto_picture_me.restype = C.c_uint32
        L.mlto_picture_me.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.mlto_picture_block_pred.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


def picture_ctus(w: int, h: int) -> np.ndarray:
    """Luma positions [n, 2] of the CTUs the gate of EncCu.cpp:755 lets through, raster order."""
    n = lib().mlto_picture_ctus(w, h, None, 0)
    xy = np.zeros((n, 2), np.int32)
    lib().mlto_picture_ctus(w, h, xy.ctypes.data, n)
    return xy


def picture_pred(ref: np.ndarray, x: int, y: int, mvx: int = 0, mvy: int = 0, size: int = 128) -> np.ndarray:
    """Integer-MV prediction block (size x size, top-left at (x, y)) from a border-replicated reference plane (frame-level pre-pass)."""
    assert ref.dtype == np.int16 and ref.ndim == 2 and ref.strides[1] == 2
    h, w = ref.shape
    out = np.zeros((size, size), np.int16)
    if size == 128:
        lib().mlto_picture_pred(ref.ctypes.data, ref.strides[0] // 2, w, h, int(x), int(y), int(mvx), int(mvy), out.ctypes.data)
    else:
        lib().mlto_picture_block_pred(ref.ctypes.data, ref.strides[0] // 2, w, h, size, int(x), int(y), int(mvx), int(mvy), out.ctypes.data)
    return out


class OracleModel:
    """Loads the raw fp32 parameters of a state_dict (numpy arrays) into the C oracle."""

    def __init__(self, sd: dict):
        from oracle.weights_io import write_raw_blob

        with tempfile.NamedTemporaryFile(suffix=".mltr", delete=False) as f:
            path = f.name
        try:
            write_raw_blob(sd, path)
            self.h = lib().mlto_load(path.encode())
        finally:
            os.unlink(path)
        if not self.h:
            raise RuntimeError("mlto_load failed")

    def close(self):
        if self.h:
            lib().mlto_free(self.h)
            self.h = None

    __del__ = close

    def stage(self, org: np.ndarray, pred: np.ndarray) -> np.ndarray:
        """org/pred: int16 2-D views (may be strided along rows)."""
        assert org.dtype == np.int16 and pred.dtype == np.int16
        assert org.strides[1] == 2 and pred.strides[1] == 2
        x = np.empty((2, 128, 128), np.float32)
        lib().mlto_stage(org.ctypes.data, org.strides[0] // 2, pred.ctypes.data, pred.strides[0] // 2, x.ctypes.data)
        return x

    def forward_ex(self, x: np.ndarray, poc: int, qp: int):
        x = np.ascontiguousarray(x, np.float32)
        lg = np.empty(9, np.float32)
        g1, g2, g3 = np.empty(64, np.float32), np.empty(128, np.float32), np.empty(256, np.float32)
        lib().mlto_forward_ex(self.h, x.ctypes.data, int(poc), int(qp), lg.ctypes.data, g1.ctypes.data, g2.ctypes.data, g3.ctypes.data)
        return lg, (g1, g2, g3)

    def predict_batch(self, orgpred: np.ndarray, pocqp: np.ndarray, nthreads: int | None = None):
        orgpred = np.ascontiguousarray(orgpred, np.int16)
        pocqp = np.ascontiguousarray(pocqp, np.int32)
        n = len(orgpred)
        lg = np.empty((n, 9), np.float32)
        sp = np.empty(n, np.int32)
        nt = nthreads or (os.cpu_count() or 1)
        lib().mlto_predict_batch(self.h, n, orgpred.ctypes.data, pocqp.ctypes.data, lg.ctypes.data, sp.ctypes.data, nt)
        return lg, sp


class OracleCuModel:
    """C oracle of the smaller-CU model (64 / 32 / 16 px): raw fp32 parameters of a CU state_dict."""

    def __init__(self, sd: dict, size: int):
        from oracle.weights_io import write_raw_cu_blob

        self.size = size
        with tempfile.NamedTemporaryFile(suffix=".mltr", delete=False) as f:
            path = f.name
        try:
            write_raw_cu_blob(sd, size, path)
            self.h = lib().mlto_load(path.encode())
        finally:
            os.unlink(path)
        if not self.h:
            raise RuntimeError("mlto_load failed")

    def close(self):
        if self.h:
            lib().mlto_free(self.h)
            self.h = None

    __del__ = close

    def stage(self, org: np.ndarray, pred: np.ndarray) -> np.ndarray:
        assert org.dtype == np.int16 and pred.dtype == np.int16 and org.strides[1] == 2 and pred.strides[1] == 2
        x = np.empty((2, self.size, self.size), np.float32)
        lib().mlto_cu_stage(self.size, org.ctypes.data, org.strides[0] // 2, pred.ctypes.data, pred.strides[0] // 2, x.ctypes.data)
        return x

    def predict_batch(self, orgpred: np.ndarray, pocqp: np.ndarray, nthreads: int | None = None) -> np.ndarray:
        orgpred = np.ascontiguousarray(orgpred, np.int16)
        pocqp = np.ascontiguousarray(pocqp, np.int32)
        n = len(orgpred)
        assert orgpred.shape[1:] == (2, self.size, self.size)
        lg = np.empty((n, 15), np.float32)
        lib().mlto_cu_predict_batch(self.h, self.size, n, orgpred.ctypes.data, pocqp.ctypes.data, lg.ctypes.data, nthreads or (os.cpu_count() or 1))
        return lg


def picture_me(org: np.ndarray, ref: np.ndarray, x: int, y: int, search_range: int):
    """Integer full-search block matching of the CTU at (x, y): ((mvx, mvy), cost) -- the library's mlt_estimate_picture_mv."""
    assert org.dtype == ref.dtype == np.int16 and org.shape == ref.shape and org.strides[1] == ref.strides[1] == 2
    h, w = org.shape
    mv = np.zeros(2, np.int16)
    cost = lib().mlto_picture_me(org.ctypes.data, org.strides[0] // 2, ref.ctypes.data, ref.strides[0] // 2, w, h, int(x), int(y),
                                 int(search_range), mv.ctypes.data)
    return (int(mv[0]), int(mv[1])), int(cost)
