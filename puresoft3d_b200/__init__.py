"""puresoft3d_b200 — the per-frame rasterisation hot path of Puresoft3D as hand-written sm_100a CUDA behind the
reference's pipeline API (include/ps3d.h is the C-ABI; pipeline.py / include/puresoft3d_b200.hpp mirror
`PuresoftPipeline`). Importing the package does not load the CUDA library; constructing a PuresoftPipeline does,
and fails loudly when it is missing (no CPU fallback)."""
from . import _capi  # noqa: F401
from .pipeline import *  # noqa: F401,F403
from .pipeline import PuresoftPipeline, PuresoftVBO, PuresoftProcessor  # noqa: F401
