"""mmlspark_b200 — B200-native engine behind MMLSpark's LightGBM-on-Spark training path.

Layout (only what the hot path needs):
  csrc/       CUDA kernels (sm_100a) + engine + the LGBM_* C ABI   -> lib/libb200gbm.so
  capi.py     ctypes binding of the C ABI (stand-in for the SWIG lightgbmlib class)
  modeltext.py  LightGBM model-text v3 parser / comparator
  lightgbm/   host-side mirror of the reference's estimator / trainCore / rendezvous layer
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
