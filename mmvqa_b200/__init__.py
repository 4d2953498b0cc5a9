"""mmvqa_b200 -- B200-native (sm_100a) implementation of the MMBERT fusion-encoder hot path of
DannielSilva/MM-VQA behind the reference's nn.Module surface.  See DESIGN.md / INTEGRATION.md."""
from .config import compute_dtype, compute_dtype_scope, set_compute_dtype  # noqa: F401
from ._lib import MMVQAError, launch_count, lib  # noqa: F401

__all__ = ["compute_dtype", "compute_dtype_scope", "set_compute_dtype", "MMVQAError", "launch_count", "lib"]
