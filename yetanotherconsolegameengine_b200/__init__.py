"""B200-native frame producer for YetAnotherConsoleGameEngine's per-frame ray tracing path.

Layout (the path only — see DESIGN.md):
  csrc/      hand-written sm_100a CUDA kernels + the C ABI of include/ycge.h      -> libycge.so
  host/      C++ mirror of the reference's C# host side (scenes, MeshLoader, ...)  -> libycge_host.so
  host_cs/   the C# wrapper a maintainer adds to the reference (cannot be compiled here)
  api.py     ctypes bindings mirroring the reference-facing names
"""
from .api import (CELL_DTYPE, BENCH_POSE, CudaRaytraceRenderer, HostScene, YcgeError, ansi_from_cells, default_params,  # noqa: F401
                  load_host, load_lib)
