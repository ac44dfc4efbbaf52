"""ctypes binding of libmltcnn.so (include/mltcnn.h) -- plumbing only, no compute and no fallback.

Mirrors the reference hook's call shape (EncCu.cpp:806-921): `predict_ctu(org, pred, poc, qp)` returns the
level-3 argmax the encoder hands to `setNewModeList`, plus logits / probabilities / per-level flags.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CTU = 128


class MltResult(C.Structure):
    _fields_ = [
        ("split_l3", C.c_int32),
        ("split_l2", C.c_int32),
        ("split_l1", C.c_int32),
        ("flags", C.c_uint32),
        ("logits", C.c_float * 9),
        ("probs", C.c_float * 9),
    ]


RESULT_DTYPE = np.dtype(
    [("split_l3", "<i4"), ("split_l2", "<i4"), ("split_l1", "<i4"), ("flags", "<u4"), ("logits", "<f4", 9), ("probs", "<f4", 9)]
)
assert RESULT_DTYPE.itemsize == C.sizeof(MltResult) == 88


class CtuDesc(C.Structure):
    _fields_ = [
        ("org", C.c_void_p),
        ("pred", C.c_void_p),
        ("org_stride", C.c_int32),
        ("pred_stride", C.c_int32),
        ("poc", C.c_int32),
        ("qp", C.c_int32),
    ]


class MltError(RuntimeError):
    def __init__(self, rc: int, what: str, detail: str = ""):
        super().__init__(f"{what} failed: rc={rc} ({detail})")
        self.rc = rc


EXPORTS = {
    "mlt_abi_version": (C.c_int, []),
    "mlt_strerror": (C.c_char_p, [C.c_int]),
    "mlt_last_error": (C.c_char_p, [C.c_void_p]),
    "mlt_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p, C.c_int]),
    "mlt_create_ex": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p, C.c_int, C.c_int]),
    "mlt_destroy": (None, [C.c_void_p]),
    "mlt_predict_ctu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mlt_predict_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mlt_predict_batch_dense": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mlt_predict_batch_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mlt_submit_batch_dense": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mlt_collect": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "mlt_pack10": (C.c_uint64, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "mlt_predict_batch_packed10": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mlt_submit_batch_packed10": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mlt_begin_picture": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "mlt_predict_ctu_in_picture": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mlt_pin_host_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "mlt_unpin_host_buffer": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mlt_picture_ctu_count": (C.c_int, [C.c_void_p]),
    "mlt_estimate_picture_mv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mlt_predict_picture": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "mlt_debug_picture_pred": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "mlt_set_engine": (C.c_int, [C.c_void_p, C.c_int]),
    "mlt_launch_count": (C.c_uint64, [C.c_void_p]),
    "mlt_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "mlt_get_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "mlt_debug_stage": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mlt_debug_activation": (C.c_int64, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
}

CU_RESULT_DTYPE = np.dtype([("split", "<i4", 4), ("logits", "<f4", 15), ("probs", "<f4", 15)])
assert CU_RESULT_DTYPE.itemsize == 136

# include/mltcnn_cu.h -- the smaller-CU models (64 / 32 / 16 px)
EXPORTS.update({
    "mlt_cu_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "mlt_cu_destroy": (None, [C.c_void_p]),
    "mlt_cu_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mlt_cu_predict_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mlt_cu_predict_batch_dense": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mlt_cu_submit_batch_dense": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mlt_cu_collect": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "mlt_cu_picture_cu_count": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "mlt_cu_predict_picture": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_int]),
    "mlt_cu_predict_batch_device": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mlt_cu_last_error": (C.c_char_p, [C.c_void_p]),
    "mlt_cu_size": (C.c_int, [C.c_void_p]),
    "mlt_cu_layer_info": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "mlt_cu_launch_count": (C.c_uint64, [C.c_void_p]),
    "mlt_cu_debug_activation": (C.c_int64, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
})

_LIB = None
PACKED10_BYTES = 40960  # MLT_CTU_PACKED10_BYTES


def pack10(orgpred: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """Host-side packer of the 10-bit transport (mlt_pack10): int16 [n,2,128,128] -> uint8 [n,40960].  Raises if a sample lies
    outside [0, 1023] (such blocks must use the int16 entry points)."""
    if orgpred.dtype != np.int16 or orgpred.shape[1:] != (2, CTU, CTU) or not orgpred.flags.c_contiguous:
        raise ValueError("orgpred must be C-contiguous int16 [n,2,128,128]")
    n = len(orgpred)
    if out is None:
        out = np.empty((n, PACKED10_BYTES), np.uint8)
    bad = load_library().mlt_pack10(orgpred.ctypes.data, orgpred.size, out.ctypes.data)
    if bad:
        raise ValueError(f"{bad} samples outside [0, 1023]: not packable, use the int16 entry points")
    return out


def lib_path() -> str:
    return os.environ.get("MLT_LIBRARY", os.path.join(HERE, "libmltcnn.so"))


def load_library():
    """dlopen libmltcnn.so and bind every symbol of include/mltcnn.h.  Raises if the library is missing."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it with `python -m fastintercu_vvc_b200.build` (there is no CPU fallback)"
            )
        L = C.CDLL(path)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _i16_2d(a: np.ndarray):
    if a.dtype != np.int16 or a.ndim != 2 or a.shape != (CTU, CTU) or a.strides[1] != 2 or a.strides[0] % 2:
        raise ValueError("expected an int16 [128,128] view with unit element stride (row stride may be larger)")
    return a.ctypes.data, a.strides[0] // 2


class MltPredictor:
    """One context per process per GPU (single-threaded, synchronous), like the reference hook."""

    def __init__(self, weights_path: str, device: int = 0, max_batch: int = 512):
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.mlt_create_ex(C.byref(self._h), os.fsencode(weights_path), int(device), int(max_batch))
        if rc != 0:
            self._h = C.c_void_p()
            raise MltError(rc, "mlt_create_ex", self._lib.mlt_strerror(rc).decode())
        self.max_batch = max_batch

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.mlt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc < 0:
            raise MltError(rc, what, self._lib.mlt_last_error(self._h).decode() or self._lib.mlt_strerror(rc).decode())
        return rc

    # -- the hook call (EncCu.cpp:806-921)
    def predict_ctu(self, org: np.ndarray, pred: np.ndarray, poc: int, qp: int) -> np.void:
        op, os_ = _i16_2d(org)
        pp, ps = _i16_2d(pred)
        out = np.zeros(1, RESULT_DTYPE)
        self._check(self._lib.mlt_predict_ctu(self._h, op, os_, pp, ps, int(poc), int(qp), out.ctypes.data), "mlt_predict_ctu")
        return out[0]

    def predict_batch(self, ctus) -> np.ndarray:
        """ctus: sequence of (org, pred, poc, qp) with int16 [128,128] views (strided rows allowed)."""
        n = len(ctus)
        descs = (CtuDesc * max(n, 1))()
        for i, (org, pred, poc, qp) in enumerate(ctus):
            descs[i].org, descs[i].org_stride = _i16_2d(org)
            descs[i].pred, descs[i].pred_stride = _i16_2d(pred)
            descs[i].poc, descs[i].qp = int(poc), int(qp)
        out = np.zeros(n, RESULT_DTYPE)
        self._check(self._lib.mlt_predict_batch(self._h, n, descs, out.ctypes.data), "mlt_predict_batch")
        return out

    def predict_batch_dense(self, orgpred: np.ndarray, pocqp: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        n = len(orgpred)
        if orgpred.dtype != np.int16 or orgpred.shape[1:] != (2, CTU, CTU) or not orgpred.flags.c_contiguous:
            raise ValueError("orgpred must be C-contiguous int16 [n,2,128,128]")
        pocqp = np.ascontiguousarray(pocqp, np.int32)
        if out is None:
            out = np.zeros(n, RESULT_DTYPE)
        self._check(
            self._lib.mlt_predict_batch_dense(self._h, n, orgpred.ctypes.data, pocqp.ctypes.data, out.ctypes.data),
            "mlt_predict_batch_dense",
        )
        return out

    def submit_batch_dense(self, orgpred: np.ndarray, pocqp: np.ndarray):
        """Pipelined form: enqueue and return; at most two batches in flight.  `orgpred` must stay alive and unchanged
        until the matching collect()."""
        n = len(orgpred)
        if orgpred.dtype != np.int16 or orgpred.shape[1:] != (2, CTU, CTU) or not orgpred.flags.c_contiguous:
            raise ValueError("orgpred must be C-contiguous int16 [n,2,128,128]")
        pocqp = np.ascontiguousarray(pocqp, np.int32)
        self._check(self._lib.mlt_submit_batch_dense(self._h, n, orgpred.ctypes.data, pocqp.ctypes.data), "mlt_submit_batch_dense")

    # -- 10-bit packed transport (40 KiB per CTU instead of 64 KiB; samples must be in [0, 1023])
    @staticmethod
    def _packed_view(packed: np.ndarray) -> int:
        if packed.dtype != np.uint8 or packed.ndim != 2 or packed.shape[1] != PACKED10_BYTES or not packed.flags.c_contiguous:
            raise ValueError(f"packed must be C-contiguous uint8 [n,{PACKED10_BYTES}] (pack10())")
        return len(packed)

    def predict_batch_packed10(self, packed: np.ndarray, pocqp: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        n = self._packed_view(packed)
        pocqp = np.ascontiguousarray(pocqp, np.int32)
        if out is None:
            out = np.zeros(n, RESULT_DTYPE)
        self._check(self._lib.mlt_predict_batch_packed10(self._h, n, packed.ctypes.data, pocqp.ctypes.data, out.ctypes.data), "mlt_predict_batch_packed10")
        return out

    def submit_batch_packed10(self, packed: np.ndarray, pocqp: np.ndarray):
        """Pipelined like submit_batch_dense; `packed` must stay alive and unchanged until the matching collect()."""
        n = self._packed_view(packed)
        pocqp = np.ascontiguousarray(pocqp, np.int32)
        self._check(self._lib.mlt_submit_batch_packed10(self._h, n, packed.ctypes.data, pocqp.ctypes.data), "mlt_submit_batch_packed10")

    def collect(self, out: np.ndarray | None = None) -> np.ndarray:
        """Blocks for the oldest submitted batch; returns its results (a view of `out` when given)."""
        if out is None:
            out = np.zeros(self.max_batch, RESULT_DTYPE)
        n = C.c_int(0)
        self._check(self._lib.mlt_collect(self._h, out.ctypes.data, C.byref(n)), "mlt_collect")
        return out[: n.value]

    def predict_batch_device(self, n: int, d_orgpred: int, d_pocqp: int, d_out: int, stream: int = 0):
        """Raw device pointers (e.g. torch.Tensor.data_ptr()) and a cudaStream_t handle; asynchronous."""
        self._check(
            self._lib.mlt_predict_batch_device(self._h, int(n), C.c_void_p(d_orgpred), C.c_void_p(d_pocqp), C.c_void_p(d_out), C.c_void_p(stream)),
            "mlt_predict_batch_device",
        )

    # -- per-picture staging
    def begin_picture(self, org_luma: np.ndarray, poc: int):
        if org_luma.dtype != np.int16 or org_luma.ndim != 2 or org_luma.strides[1] != 2:
            raise ValueError("org_luma must be an int16 2-D view")
        h, w = org_luma.shape
        self._check(
            self._lib.mlt_begin_picture(self._h, org_luma.ctypes.data, org_luma.strides[0] // 2, w, h, int(poc)), "mlt_begin_picture"
        )

    def predict_ctu_in_picture(self, x: int, y: int, pred: np.ndarray, qp: int) -> np.void:
        pp, ps = _i16_2d(pred)
        out = np.zeros(1, RESULT_DTYPE)
        self._check(
            self._lib.mlt_predict_ctu_in_picture(self._h, int(x), int(y), pp, ps, int(qp), out.ctypes.data), "mlt_predict_ctu_in_picture"
        )
        return out[0]

    def pin_host_buffer(self, a: np.ndarray):
        """Page-lock the memory behind `a` (its base allocation must outlive the pin); see mlt_pin_host_buffer."""
        self._check(self._lib.mlt_pin_host_buffer(self._h, C.c_void_p(a.ctypes.data), C.c_uint64(a.nbytes)), "mlt_pin_host_buffer")

    def unpin_host_buffer(self, a: np.ndarray):
        self._check(self._lib.mlt_unpin_host_buffer(self._h, C.c_void_p(a.ctypes.data)), "mlt_unpin_host_buffer")

    # -- frame-level pre-pass
    def picture_ctu_count(self) -> int:
        n = self._lib.mlt_picture_ctu_count(self._h)
        if n < 0:
            self._check(n, "mlt_picture_ctu_count")
        return n

    def estimate_picture_mv(self, ref_luma: np.ndarray | None, search_range: int):
        """Integer full-search block matching per eligible CTU on the device -> (mv [n, 2] int16 (x, y), cost [n] uint32).
        ref_luma = None reuses the reference plane already uploaded for this picture."""
        if ref_luma is not None and (ref_luma.dtype != np.int16 or ref_luma.ndim != 2 or ref_luma.strides[1] != 2):
            raise ValueError("ref_luma must be an int16 2-D view")
        n = self.picture_ctu_count()
        mv, cost = np.zeros((n, 2), np.int16), np.zeros(n, np.uint32)
        rc = self._lib.mlt_estimate_picture_mv(
            self._h, ref_luma.ctypes.data if ref_luma is not None else None, ref_luma.strides[0] // 2 if ref_luma is not None else 0,
            int(search_range), mv.ctypes.data, cost.ctypes.data,
        )
        if rc < 0:
            self._check(rc, "mlt_estimate_picture_mv")
        return mv[:rc], cost[:rc]

    def predict_picture(self, ref_luma: np.ndarray | None, slice_qp: int, mv: np.ndarray | None = None, ctu_qp: np.ndarray | None = None) -> np.ndarray:
        """All eligible CTUs of the picture begun, in raster order, from integer-MV prediction out of `ref_luma` (None: the
        reference plane already on the device, e.g. after estimate_picture_mv)."""
        if ref_luma is not None and (ref_luma.dtype != np.int16 or ref_luma.ndim != 2 or ref_luma.strides[1] != 2):
            raise ValueError("ref_luma must be an int16 2-D view")
        n = self.picture_ctu_count()
        if mv is not None:
            mv = np.ascontiguousarray(mv, np.int16)
            if mv.shape != (n, 2):
                raise ValueError(f"mv must be [{n}, 2]")
        if ctu_qp is not None:
            ctu_qp = np.ascontiguousarray(ctu_qp, np.int32)
            if ctu_qp.shape != (n,):
                raise ValueError(f"ctu_qp must be [{n}]")
        out = np.zeros(n, RESULT_DTYPE)
        rc = self._lib.mlt_predict_picture(
            self._h, ref_luma.ctypes.data if ref_luma is not None else None, ref_luma.strides[0] // 2 if ref_luma is not None else 0,
            mv.ctypes.data if mv is not None else None,
            ctu_qp.ctypes.data if ctu_qp is not None else None, int(slice_qp), out.ctypes.data, n,
        )
        if rc < 0:
            self._check(rc, "mlt_predict_picture")
        return out[:rc]

    def debug_picture_pred(self) -> np.ndarray:
        n = self.picture_ctu_count()
        out = np.zeros((n, CTU, CTU), np.int16)
        rc = self._lib.mlt_debug_picture_pred(self._h, out.ctypes.data, n)
        if rc < 0:
            self._check(rc, "mlt_debug_picture_pred")
        return out[:rc]

    # -- test hooks
    def set_engine(self, engine: int):
        self._check(self._lib.mlt_set_engine(self._h, int(engine)), "mlt_set_engine")

    @property
    def launch_count(self) -> int:
        return int(self._lib.mlt_launch_count(self._h))

    def set_profiling(self, on: bool):
        self._check(self._lib.mlt_set_profiling(self._h, int(bool(on))), "mlt_set_profiling")

    def get_profile(self) -> np.ndarray:
        """Device time (ms) of each kernel of the last batch: [stage+conv1, 16 convs, head]."""
        ms = np.zeros(18, np.float32)
        n = self._check(self._lib.mlt_get_profile(self._h, ms.ctypes.data, 18), "mlt_get_profile")
        return ms[:n]

    def debug_stage(self, ctus) -> np.ndarray:
        n = len(ctus)
        descs = (CtuDesc * n)()
        for i, (org, pred, poc, qp) in enumerate(ctus):
            descs[i].org, descs[i].org_stride = _i16_2d(org)
            descs[i].pred, descs[i].pred_stride = _i16_2d(pred)
            descs[i].poc, descs[i].qp = int(poc), int(qp)
        out = np.empty((n, 2, CTU, CTU), np.float32)
        self._check(self._lib.mlt_debug_stage(self._h, n, descs, out.ctypes.data), "mlt_debug_stage")
        return out

    def debug_activation(self, layer: int, n: int) -> np.ndarray:
        shapes = [(128, 32)] + [(64, 32)] * 4 + [(32, 64)] * 4 + [(16, 128)] * 4 + [(8, 256)] * 4
        h, c = shapes[layer]
        out = np.empty((n, h, h, c), np.float32)
        got = self._check(self._lib.mlt_debug_activation(self._h, layer, out.ctypes.data, out.size), "mlt_debug_activation")
        assert got == out.size, (got, out.size)
        return out


class MltCuPredictor:
    """Smaller-CU model (64 / 32 / 16 px): the hook's cuw != 128 branch (EncCu.cpp:754,899,916-919).  `split[:, 0]` of a
    result is the level-1 argmax the encoder would hand to setNewModeList."""

    def __init__(self, weights_path: str, size: int, device: int = 0, max_batch: int = 4096):
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.mlt_cu_create(C.byref(self._h), os.fsencode(weights_path), int(device), int(size), int(max_batch))
        if rc != 0:
            self._h = C.c_void_p()
            raise MltError(rc, "mlt_cu_create", self._lib.mlt_strerror(rc).decode())
        self.size, self.max_batch = size, max_batch

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.mlt_cu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc < 0:
            raise MltError(rc, what, self._lib.mlt_cu_last_error(self._h).decode() or self._lib.mlt_strerror(rc).decode())
        return rc

    def _view(self, a: np.ndarray):
        if a.dtype != np.int16 or a.shape != (self.size, self.size) or a.strides[1] != 2 or a.strides[0] % 2:
            raise ValueError(f"expected an int16 [{self.size},{self.size}] view with unit element stride")
        return a.ctypes.data, a.strides[0] // 2

    def predict(self, org: np.ndarray, pred: np.ndarray, poc: int, qp: int) -> np.void:
        op, os_ = self._view(org)
        pp, ps = self._view(pred)
        out = np.zeros(1, CU_RESULT_DTYPE)
        self._check(self._lib.mlt_cu_predict(self._h, op, os_, pp, ps, int(poc), int(qp), out.ctypes.data), "mlt_cu_predict")
        return out[0]

    def predict_batch(self, cus) -> np.ndarray:
        n = len(cus)
        descs = (CtuDesc * max(n, 1))()
        for i, (org, pred, poc, qp) in enumerate(cus):
            descs[i].org, descs[i].org_stride = self._view(org)
            descs[i].pred, descs[i].pred_stride = self._view(pred)
            descs[i].poc, descs[i].qp = int(poc), int(qp)
        out = np.zeros(n, CU_RESULT_DTYPE)
        self._check(self._lib.mlt_cu_predict_batch(self._h, n, descs, out.ctypes.data), "mlt_cu_predict_batch")
        return out

    def predict_batch_dense(self, orgpred: np.ndarray, pocqp: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        n = len(orgpred)
        if orgpred.dtype != np.int16 or orgpred.shape[1:] != (2, self.size, self.size) or not orgpred.flags.c_contiguous:
            raise ValueError("orgpred must be C-contiguous int16 [n,2,size,size]")
        pocqp = np.ascontiguousarray(pocqp, np.int32)
        if out is None:
            out = np.zeros(n, CU_RESULT_DTYPE)
        self._check(self._lib.mlt_cu_predict_batch_dense(self._h, n, orgpred.ctypes.data, pocqp.ctypes.data, out.ctypes.data),
                    "mlt_cu_predict_batch_dense")
        return out

    def submit_batch_dense(self, orgpred: np.ndarray, pocqp: np.ndarray):
        """Pipelined form: enqueue and return; at most two batches in flight.  `orgpred` must stay alive and unchanged
        until the matching collect()."""
        n = len(orgpred)
        if orgpred.dtype != np.int16 or orgpred.shape[1:] != (2, self.size, self.size) or not orgpred.flags.c_contiguous:
            raise ValueError("orgpred must be C-contiguous int16 [n,2,size,size]")
        pocqp = np.ascontiguousarray(pocqp, np.int32)
        self._check(self._lib.mlt_cu_submit_batch_dense(self._h, n, orgpred.ctypes.data, pocqp.ctypes.data), "mlt_cu_submit_batch_dense")

    def collect(self, out: np.ndarray | None = None) -> np.ndarray:
        """Blocks for the oldest submitted batch; returns its results (a view of `out` when given)."""
        if out is None:
            out = np.zeros(self.max_batch, CU_RESULT_DTYPE)
        n = C.c_int(0)
        self._check(self._lib.mlt_cu_collect(self._h, out.ctypes.data, C.byref(n)), "mlt_cu_collect")
        return out[: n.value]

    def predict_picture(self, org_luma: np.ndarray, ref_luma: np.ndarray, poc: int, qp: int, mv: np.ndarray | None = None) -> np.ndarray:
        """Every size x size block of the picture's CU raster (fully inside the picture), raster order, in one batch;
        pred = integer-MV prediction out of `ref_luma` built on the device."""
        for a in (org_luma, ref_luma):
            if a.dtype != np.int16 or a.ndim != 2 or a.strides[1] != 2 or a.shape != org_luma.shape:
                raise ValueError("org_luma / ref_luma must be int16 2-D views of the same shape")
        h, w = org_luma.shape
        n = self._lib.mlt_cu_picture_cu_count(self.size, w, h)
        if mv is not None:
            mv = np.ascontiguousarray(mv, np.int16)
            if mv.shape != (n, 2):
                raise ValueError(f"mv must be [{n}, 2]")
        out = np.zeros(max(n, 1), CU_RESULT_DTYPE)
        rc = self._lib.mlt_cu_predict_picture(
            self._h, org_luma.ctypes.data, org_luma.strides[0] // 2, ref_luma.ctypes.data, ref_luma.strides[0] // 2, w, h, int(poc),
            mv.ctypes.data if mv is not None else None, int(qp), out.ctypes.data, len(out),
        )
        if rc < 0:
            self._check(rc, "mlt_cu_predict_picture")
        return out[:rc]

    def predict_batch_device(self, n: int, d_orgpred: int, d_pocqp: int, d_out: int, stream: int = 0):
        self._check(self._lib.mlt_cu_predict_batch_device(self._h, int(n), C.c_void_p(d_orgpred), C.c_void_p(d_pocqp), C.c_void_p(d_out),
                                                          C.c_void_p(stream)), "mlt_cu_predict_batch_device")

    @property
    def launch_count(self) -> int:
        return int(self._lib.mlt_cu_launch_count(self._h))

    def debug_activation(self, layer: int, n: int) -> np.ndarray:
        """fp32 NHWC [n][H][H][C] of activation `layer` (0 = conv1 out, 1 + li = conv li out) of the last host batch."""
        planes = (32, 64, 96, 128, 256)
        if layer == 0:
            h, c = self.size, 32
        else:
            L = (layer - 1) // 4
            h, c = max(self.size >> (L + 1), 1), planes[L]
        out = np.empty((n, h, h, c), np.float32)
        got = self._check(self._lib.mlt_cu_debug_activation(self._h, layer, out.ctypes.data, out.size), "mlt_cu_debug_activation")
        assert got == out.size, (got, out.size)
        return out
