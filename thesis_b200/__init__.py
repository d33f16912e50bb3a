"""thesis_b200 -- B200-native per-scan update of the amansanghvi/Thesis RBPF.

Only what the hot path needs: csrc/ (CUDA kernels + C ABI), _lib (ctypes
binding), particles (device-resident particle set behind the reference's
Robot / resample API), models / sensors / loaders (host value types and log
readers mirroring the reference's), harness (headless main.py loop).
"""
__all__ = ["particles", "models", "sensors", "loaders", "harness", "synth"]
