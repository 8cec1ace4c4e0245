"""eleven-b200: B200-native implementation of Eleven Renderer's per-sample path-tracing hot path.

Layout (only what the path needs):
  csrc/      hand-written sm_100a CUDA kernels + the C-ABI (include/eleven_b200.h) -> libeleven_b200.so
  host/      C++ host: scene loading, `eleven <scene> <spp> <out.bmp>` CLI, BMP output
  _capi.py   ctypes mirror of include/eleven_b200.h (no torch types cross the boundary)
  renderer.py  host-side mirror of the reference's renderSetup/renderCuda/getBuffers/getSamples
  scenes.py  scene containers, procedural generators, reference-format writers
  dist.py    one-process-per-GPU sample split + NCCL film reduce (torch.distributed is plumbing only)
"""
from . import scenes  # noqa: F401

__all__ = ["scenes"]
