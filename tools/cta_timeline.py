#!/usr/bin/env python
"""Per-CTA phase timeline of the tcgen05 forward kernel (run with NN_UMMA_DEBUG=1)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from noisynet_b200 import _lib, ops
from noisynet_b200._lib import NOISE_EXTERNAL, NOISE_MERGED

dev = torch.device("cuda:0")
lib = _lib.load()
B = 512
for name, (Cin, H, Cout, k, mode, a_cs) in {"conv1": (3, 32, 65, 5, NOISE_MERGED, 1 / 15.0), "conv2": (65, 14, 120, 5, NOISE_EXTERNAL, 1 / 3.0)}.items():
    x = torch.randint(0, 16, (B, Cin, H, H), device=dev).float() * a_cs
    w_raw = torch.randn(Cout, Cin, k, k, device=dev) * 0.1
    wq = ops.quantize_fwd(w_raw, 4, -1.0, 1.0, 0.0)
    scale = ops.tensor_stats(w_raw)[1:2] if mode == NOISE_MERGED else ops.tensor_stats(x)[0:1]
    for _ in range(3):
        ops.noisy_conv_fwd(x, wq, w_raw, None, 1, 0, noise_mode=mode, current=1.0, scale_dev=scale, precision="bf16",
                           a_code_scale=a_cs, w_code_scale=1 / 15.0, want_y=False)
    n = 4096
    buf = np.zeros((n, 8), dtype=np.int64)
    got = lib.nn_debug_cta_timeline(buf.ctypes.data_as(C.c_void_p), n)
    t = buf[:got].astype(np.float64)
    t0 = t[:, 0].min()
    d = lambda a, b: (t[:, a] - t[:, b])
    print("%s: %d CTAs" % (name, got))
    for lab, a, b in (("setup (start->tmem/barriers ready)", 1, 0), ("mainloop (setup->last MMA issued)", 2, 1),
                      ("producers done - setup", 6, 1), ("MMA drain (issued->acc ready)", 3, 2), ("epilogue", 4, 3),
                      ("teardown", 5, 4), ("total CTA", 5, 0)):
        v = d(a, b) / 1900.0
        print("   %-38s mean %6.2f us  p10 %6.2f  p90 %6.2f" % (lab, v.mean(), np.percentile(v, 10), np.percentile(v, 90)))
    # per-SM view (clock64 is per SM): busy span and number of CTAs
    sm = t[:, 7].astype(int)
    spans = []
    for sid in np.unique(sm):
        sel = t[sm == sid]
        spans.append(((sel[:, 5].max() - sel[:, 0].min()) / 1900.0, len(sel)))
    spans = np.array(spans)
    print("   per SM: span mean %.1f us max %.1f us, CTAs per SM mean %.2f max %d" % (spans[:, 0].mean(), spans[:, 0].max(), spans[:, 1].mean(), spans[:, 1].max()))
