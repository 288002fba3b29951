#!/usr/bin/env python
"""Per-chunk timeline of the shift-GEMM wgrad kernel (NN_UMMA_DEBUG=1): operand wait vs MMA issue."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NN_UMMA_DEBUG"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

from noisynet_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
s_a = 1.0 / 15.0
x = torch.randint(0, 16, (512, 3, 32, 32), device=dev).float() * s_a
gy = torch.randn(512, 65, 28, 28, device=dev) * 0.01
for _ in range(3):
    ops.conv_wgrad(gy, x, (65, 3, 5, 5), 1, 0, precision="bf16", a_code_scale=s_a)
buf = np.zeros((296 * 16, 8), dtype=np.int64)
n = lib.nn_debug_cta_timeline(buf.ctypes.data_as(C.c_void_p), 296 * 16)
t = buf[:n].reshape(-1, 32, 4).astype(np.float64)
ok = t[:, :, 2] > 0
us = 1.0 / 1965.0
wait = (t[:, :, 1] - t[:, :, 0])[ok] * us
iss = (t[:, :, 2] - t[:, :, 1])[ok] * us
per = np.diff(t[:, :, 2], axis=1)[ok[:, 1:]] * us
print("%d CTAs x %.1f chunks" % (t.shape[0], ok.sum() / t.shape[0]))
for lab, a in (("wait for operands", wait), ("issue 8 x KH MMAs + commit", iss), ("chunk period", per)):
    print("   %-30s mean %6.2f us  p10 %6.2f  p90 %6.2f" % (lab, a.mean(), np.percentile(a, 10), np.percentile(a, 90)))
print("   first chunk operand wait: %.2f us" % ((t[:, 0, 1] - t[:, 0, 0]).mean() * us))
span = (t[:, :, 2].max(axis=1) - t[:, 0, 0]) * us
print("   CTA span mean %.1f us max %.1f us" % (span.mean(), span.max()))
