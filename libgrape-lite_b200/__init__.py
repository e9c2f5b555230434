"""libgrape-lite_b200 — B200-native PIE graph engine (host-side Python mirror).

The product is the C-ABI shared library built from csrc/ (include/grape_b200.h);
this package is a thin ctypes mirror of that ABI for tests and bench.py.
The directory name contains a hyphen, so import it with
``importlib.import_module("libgrape-lite_b200")``.

There is NO CPU fallback: every entry point raises when the CUDA library is
missing or no device is usable.
"""
from .capi import *  # noqa: F401,F403
