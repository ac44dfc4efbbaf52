"""CPU ORACLE helper (test infrastructure): write the raw, un-folded fp32 parameter blob ("MLTR")
that oracle/mltcnn_oracle.c loads.  Tensor order = forward order of the reference state_dict
(mlt_ctu_or_pq_arch.py:244-254; keys as saved under 'params', model2torchScript.py:23-32)."""
from __future__ import annotations

import numpy as np

from .ref_arch import bn_for, conv_keys, cu_conv_keys


def write_raw_blob(sd: dict, path: str) -> None:
    parts = [np.array([0x52544C4D, 1, 128, 0], np.uint32).tobytes()]  # "MLTR", version, arch, reserved
    for prefix, *_ in conv_keys():
        parts.append(np.ascontiguousarray(sd[f"{prefix}.weight"], np.float32).tobytes())
        if prefix != "conv1":
            bn = bn_for(prefix)
            for k in ("weight", "bias", "running_mean", "running_var"):
                parts.append(np.ascontiguousarray(sd[f"{bn}.{k}"], np.float32).tobytes())
    for i in (1, 2, 3):
        parts.append(np.ascontiguousarray(sd[f"branch{i}.weight"], np.float32).tobytes())
        parts.append(np.ascontiguousarray(sd[f"branch{i}.bias"], np.float32).tobytes())
    with open(path, "wb") as f:
        f.write(b"".join(parts))


def write_raw_cu_blob(sd: dict, size: int, path: str) -> None:
    """Same container for the smaller-CU model (mlt_cu_or_pq_arch.py:59-130); header arch field = CU size."""
    parts = [np.array([0x52544C4D, 1, size, 0], np.uint32).tobytes()]
    for prefix, *_ in cu_conv_keys():
        parts.append(np.ascontiguousarray(sd[f"{prefix}.weight"], np.float32).tobytes())
        if prefix != "conv1":
            bn = bn_for(prefix)
            for k in ("weight", "bias", "running_mean", "running_var"):
                parts.append(np.ascontiguousarray(sd[f"{bn}.{k}"], np.float32).tobytes())
    for i in (1, 2, 3, 4):
        parts.append(np.ascontiguousarray(sd[f"branch{i}.weight"], np.float32).tobytes())
        parts.append(np.ascontiguousarray(sd[f"branch{i}.bias"], np.float32).tobytes())
    with open(path, "wb") as f:
        f.write(b"".join(parts))
