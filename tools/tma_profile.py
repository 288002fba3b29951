#!/usr/bin/env python
"""Where the cycles of k_conv_tma go, per role (instrumented build: -DNN_KDEBUG, see below), at NoisyNet conv2 shapes.

    nvcc ... -DNN_KDEBUG -o noisynet_b200/lib/libnoisynet_b200_kdebug.so noisynet_b200/csrc/*.cu
    NN_LIB_PATH=noisynet_b200/lib/libnoisynet_b200_kdebug.so python tools/tma_profile.py
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from noisynet_b200 import _lib, ops  # noqa: E402
from noisynet_b200._lib import NOISE_EXTERNAL  # noqa: E402


def dump(lib, title):
    buf = (C.c_longlong * (512 * 16))()
    lib.nn_debug_tma_profile.restype = C.c_int
    lib.nn_debug_tma_profile.argtypes = [C.c_void_p, C.c_int]
    n = lib.nn_debug_tma_profile(buf, 512)
    a = np.frombuffer(buf, dtype=np.int64).reshape(512, 16)[:n]
    lead, peer = a[0::2], a[1::2]
    ghz = 1.965e3   # cycles per us
    f = lambda v: "%.1f" % (float(np.mean(v)) / ghz)
    print("== %s: %d CTAs (us, mean over CTAs)" % (title, n))
    print("  producer0 (leader): wait-empty %s  issue %s  total %s" % (f(lead[:, 0]), f(lead[:, 1]), f(lead[:, 2])))
    print("  producer0 (peer)  : wait-empty %s  issue %s  total %s" % (f(peer[:, 0]), f(peer[:, 1]), f(peer[:, 2])))
    print("  MMA thread        : wait acc-empty %s  wait full %s  wait peer %s  issue+commit %s  total %s"
          % (f(lead[:, 4]), f(lead[:, 5]), f(lead[:, 6]), f(lead[:, 7]), f(lead[:, 8])))
    print("  epilogue warp 0   : wait acc-full %s  total %s  items/CTA %.2f" % (f(lead[:, 10]), f(lead[:, 11]), float(np.mean(lead[:, 12]))))


def main():
    B = int(os.environ.get("B", "512"))
    dev = torch.device("cuda:0")
    lib = _lib.load()
    s_a = 5.0 / 15.0
    x = torch.randint(0, 16, (B, 65, 14, 14), device=dev).float() * s_a
    w_raw = torch.randn(120, 65, 5, 5, device=dev) * 0.1
    wq = ops.quantize_fwd(w_raw, 4, -1.0, 1.0, 0.0)
    scale = ops.tensor_stats(x)[0:1]
    for _ in range(3):
        r = ops.noisy_conv_fwd(x, wq, w_raw, None, 1, 0, noise_mode=NOISE_EXTERNAL, current=1.0, scale_dev=scale, precision="bf16",
                               a_code_scale=s_a, w_code_scale=1.0 / 15.0, want_y=False)
    dump(lib, "conv2 forward (noisy)")
    gy = torch.randn_like(r["y_noisy"])
    for _ in range(3):
        ops.conv_dgrad(gy, wq, x.shape, precision="bf16", w_code_scale=1.0 / 15.0)
    dump(lib, "conv2 dgrad")
    for _ in range(3):
        ops.noisy_conv_fwd(x, wq, None, None, 1, 0, precision="bf16", a_code_scale=s_a, w_code_scale=1.0 / 15.0)
    dump(lib, "conv2 forward (plain)")
    assert ops.error_flag() == 0


if __name__ == "__main__":
    main()
