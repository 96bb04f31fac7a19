"""Prints the error of the TF32 weight gradient cases (was used to pin the MN-major TF32 descriptor fields:
LBO = chunk distance, SBO = 512 B, 128B swizzle with 32-byte atoms; log in profiles/r01_tf32_probe.txt)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import conv_cases as cc  # noqa: E402
from cycle_depth_estimation_b200 import _lib  # noqa: E402

for name in sorted(cc.TF32_WGRAD_CASES):
    res = cc.conv_wgrad_tf32_case(**cc.TF32_WGRAD_CASES[name])
    print(name, "err %.3e" % res["err"], "abort", _lib.lib().cdb_device_abort_flag(), flush=True)
