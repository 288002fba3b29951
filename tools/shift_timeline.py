#!/usr/bin/env python
"""Per-tile timeline of the persistent shift-GEMM kernel (NN_UMMA_DEBUG=1): MMA issue, accumulator latency, epilogue."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NN_UMMA_DEBUG"] = "1"
import numpy as np  # noqa: E402
import torch  # noqa: E402

from noisynet_b200 import _lib, ops  # noqa: E402
from noisynet_b200._lib import NOISE_MERGED, NOISE_NONE  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
s_a = 1.0 / 15.0
x = torch.randint(0, 16, (512, 3, 32, 32), device=dev).float() * s_a
w_raw = torch.randn(65, 3, 5, 5, device=dev) * 0.1
wq = ops.quantize_fwd(w_raw, 4, -1.0, 1.0, 0.0)
scale = ops.tensor_stats(w_raw)[1:2]
for name, kw in (("plain", dict(noise_mode=NOISE_NONE)), ("noisy", dict(noise_mode=NOISE_MERGED, current=1.0, scale_dev=scale, want_y=False))):
    for _ in range(2):
        ops.noisy_conv_fwd(x, wq, w_raw, None, 1, 0, precision="bf16", a_code_scale=s_a, w_code_scale=1.0 / 15.0, **kw)
    buf = np.zeros((148 * 16, 8), dtype=np.int64)
    n = lib.nn_debug_cta_timeline(buf.ctypes.data_as(C.c_void_p), 148 * 16)
    t = buf[:n].reshape(-1, 32, 4).astype(np.float64)
    ok = t[:, :, 3] > 0
    us = 1.0 / 1965.0
    iss = (t[:, :, 1] - t[:, :, 0])[ok] * us
    lat = (t[:, :, 2] - t[:, :, 1])[ok] * us
    epi = (t[:, :, 3] - t[:, :, 2])[ok] * us
    per = np.diff(t[:, :, 3], axis=1)[ok[:, 1:]] * us
    print("%s: %d CTAs x %.1f tiles" % (name, t.shape[0], ok.sum() / t.shape[0]))
    for lab, a in (("MMA issue (13 MMAs + commits)", iss), ("commit -> accumulator seen by epilogue", lat),
                   ("epilogue (warp 2)", epi), ("tile period (epilogue end to end)", per)):
        print("   %-42s mean %6.2f us  p10 %6.2f  p90 %6.2f" % (lab, a.mean(), np.percentile(a, 10), np.percentile(a, 90)))
    span = (t[:, :, 3].max(axis=1) - t[:, 0, 0]) * us
    print("   CTA span mean %.1f us max %.1f us" % (span.mean(), span.max()))
