"""katome_b200 -- B200-native De Bruijn graph build stage (GIR producer) for katome.

Only the hot path lives here: `csrc/` (CUDA kernels + the C ABI of
include/katome_gpu.h) and the host-side mirror of the reference's GIR interface.
"""
from .gir import DeviceArray, GpuGIR, KatomeError, ReadTooShort, random_access_probe, synth_reads_device  # noqa: F401
from . import workloads  # noqa: F401

__all__ = ["DeviceArray", "GpuGIR", "KatomeError", "ReadTooShort", "random_access_probe", "synth_reads_device", "workloads"]
