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


def trace(lib, title, n=75):
    buf = (C.c_longlong * (1024 * 4))()
    lib.nn_debug_tma_trace.restype = C.c_int
    lib.nn_debug_tma_trace.argtypes = [C.c_void_p, C.c_int]
    k = lib.nn_debug_tma_trace(buf, 1024)
    a = np.frombuffer(buf, dtype=np.int64).reshape(1024, 4)[:n]
    t0 = a[0, 0]
    print("== %s: MMA thread of cluster 0, per group (cycles): wait for the stage | issue MMAs + commit | start-to-start" % title)
    prev = None
    line = []
    for g in range(n):
        w, i = a[g, 1] - a[g, 0], a[g, 2] - a[g, 1]
        line.append("%4d|%4d|%5d" % (w, i, (a[g, 0] - prev) if prev is not None else 0))
        prev = a[g, 0]
        if len(line) == 5:
            print("   " + "   ".join(line)); line = []
    if line:
        print("   " + "   ".join(line))


def main():
    B = int(os.environ.get("B", "512"))
    dev = torch.device("cuda:0")
    lib = _lib.load()
    s_a = 5.0 / 15.0
    x = torch.randint(0, 16, (B, 65, 14, 14), device=dev).float() * s_a
    w_raw = torch.randn(120, 65, 5, 5, device=dev) * 0.1
    wq = ops.quantize_fwd(w_raw, 4, -1.0, 1.0, 0.0)
    scale = ops.tensor_stats(x)[0:1]
    fwd = lambda: ops.noisy_conv_fwd(x, wq, w_raw, None, 1, 0, noise_mode=NOISE_EXTERNAL, current=1.0, scale_dev=scale, precision="bf16",
                                     a_code_scale=s_a, w_code_scale=1.0 / 15.0, want_y=False)
    for _ in range(3):
        r = fwd()
    dump(lib, "conv2 forward (noisy)")
    trace(lib, "conv2 forward (noisy)")
    gy = torch.randn_like(r["y_noisy"])
    dgrad = lambda: ops.conv_dgrad(gy, wq, x.shape, precision="bf16", w_code_scale=1.0 / 15.0)
    for _ in range(3):
        dgrad()
    dump(lib, "conv2 dgrad")
    # experiments (kdebug build): which role bounds the kernel?  mode bits: 1 no epilogue work, 2 stale operands (no copies
    # after the first ring revolution), 4 no per-group clock reads
    lib.nn_debug_tma_mode.restype = C.c_int
    lib.nn_debug_tma_mode.argtypes = [C.c_int]
    names = {4: "all roles", 5: "no epilogue work", 6: "no operand copies", 7: "MMA issue only", 13: "free-running MMA thread", 29: "MMA thread alone"}
    lib.nn_debug_main_kernel_ms(1)
    for title, call in (("conv2 forward (noisy)", fwd), ("conv2 dgrad", None)):
        for mode in (4, 7, 13, 29):
            lib.nn_debug_tma_mode(mode)
            ts = []
            for _ in range(5):
                if call is not None:
                    call()
                    ts.append(lib.nn_debug_main_kernel_ms(-1))
                else:
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); dgrad(); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
            print("  %-24s %-20s %.1f us%s" % (title, names[mode], 1e3 * min(ts), "" if call is not None else "  (incl. the two pack kernels, ~20 us)"))
    lib.nn_debug_main_kernel_ms(0)
    lib.nn_debug_tma_mode(0)
    assert ops.error_flag() == 0


if __name__ == "__main__":
    main()
