#!/usr/bin/env python
"""Launches a few probe kernels once each (for an ncu capture of their pipe utilisation)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "zk_stark_tutor_b200", "lib", "libzkb200_probe.so"))
r, ms, cs = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_uint64(0)
for cfg in (0, 4, 28, 20):
    lib.zkb_probe_blakex(0, cfg, ctypes.byref(r), ctypes.byref(ms), ctypes.byref(cs))
    print("blakex", cfg, r.value / 1e9)
