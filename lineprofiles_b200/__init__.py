"""lineprofiles_b200 -- B200-native (sm_100a) implementation of the klineprofiles hot path.

Python is only a thin ctypes binding over the C ABI of include/klineprofiles_b200.h
(lineprofiles_b200/libklineprofiles_b200.so, built by __graft_entry__.build()).  There is no
Python / CPU compute path: if the shared library is missing, or no CUDA device is present,
evaluation raises.
"""
from .api import (MODELS, NPARAMS, PRECISIONS, Group, KlpError, Table, nccl_unique_id, nccl_version, kconv, kconv5, kconv10, kline, kline5,  # noqa: F401
                  library_path, load_library, set_default_table, set_legacy_precision, legacy_reset)
from . import synthetic, tableio  # noqa: F401
