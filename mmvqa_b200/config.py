"""Compute-dtype selection for the CUDA path.

``bf16`` (default): activations stored as bf16, GEMMs on tcgen05/TMEM tensor cores with fp32
accumulation, LayerNorm statistics / RealFormer scores / parameter gradients in fp32.
``fp32``: fp32 activations and SIMT fp32 GEMMs -- the validation path whose argmax must equal
the reference's.  This is a precision choice, not a fallback: both run the same CUDA library.
"""
from __future__ import annotations

import contextlib
import os

import torch

_NAMES = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp32": torch.float32, "float32": torch.float32}
_state = {"dtype": _NAMES[os.environ.get("MMVQA_DTYPE", "bf16").lower()]}


def set_compute_dtype(dtype) -> None:
    if isinstance(dtype, str):
        dtype = _NAMES[dtype.lower()]
    if dtype not in (torch.bfloat16, torch.float32):
        raise ValueError("compute dtype must be torch.bfloat16 or torch.float32")
    _state["dtype"] = dtype


def compute_dtype() -> torch.dtype:
    return _state["dtype"]


@contextlib.contextmanager
def compute_dtype_scope(dtype):
    old = _state["dtype"]
    set_compute_dtype(dtype)
    try:
        yield
    finally:
        _state["dtype"] = old
