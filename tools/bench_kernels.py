#!/usr/bin/env python
"""Micro-benchmark of the hot-path kernels at NoisyNet batch-512 shapes (CUDA events, inputs cycled)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from noisynet_b200 import _lib, ops  # noqa: E402
from noisynet_b200._lib import NOISE_EXTERNAL, NOISE_MERGED  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i, (a, b) in enumerate(ev):
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / reps


def main():
    B = int(os.environ.get("B", "512"))
    dev = torch.device("cuda:0")
    lib = _lib.load()
    res = {}
    s_a = 5.0 / 15.0
    shapes = {"conv1": (3, 32, 65, 5, NOISE_MERGED, 1.0 / 15.0), "conv2": (65, 14, 120, 5, NOISE_EXTERNAL, s_a),
              "fc1": (3000, 1, 390, 1, NOISE_MERGED, s_a), "fc2": (390, 1, 10, 1, NOISE_EXTERNAL, s_a)}
    for name, (Cin, H, Cout, k, mode, a_cs) in shapes.items():
        xs = [(torch.randint(0, 16, (B, Cin, H, H), device=dev).float() * a_cs) for _ in range(6)]
        w_raw = torch.randn(Cout, Cin, k, k, device=dev) * 0.1
        wq = ops.quantize_fwd(w_raw, 4, -1.0, 1.0, 0.0)
        scale = ops.tensor_stats(w_raw)[1:2] if mode == NOISE_MERGED else ops.tensor_stats(xs[0])[0:1]

        def fwd(i):
            return ops.noisy_conv_fwd(xs[i % 6], wq, w_raw, None, 1, 0, noise_mode=mode, current=1.0, scale_dev=scale,
                                      precision="bf16", a_code_scale=a_cs, w_code_scale=1.0 / 15.0, want_y=False)
        total = timeit(fwd)
        lib.nn_debug_main_kernel_ms(1)
        mains = []
        for i in range(10):
            fwd(i)
            mains.append(lib.nn_debug_main_kernel_ms(-1))
        lib.nn_debug_main_kernel_ms(0)
        y = fwd(0)["y_noisy"]
        gys = [torch.randn_like(y) for _ in range(3)]
        res[name] = {"fwd_total_us": 1e3 * total, "fwd_main_us": 1e3 * sum(mains) / len(mains)}
        if name != "conv1":
            res[name]["dgrad_us"] = 1e3 * timeit(lambda i: ops.conv_dgrad(gys[i % 3], wq, xs[0].shape, precision="bf16",
                                                                          w_code_scale=1.0 / 15.0))
        res[name]["wgrad_us"] = 1e3 * timeit(lambda i: ops.conv_wgrad(gys[i % 3], xs[i % 6], wq.shape, precision="bf16",
                                                                      a_code_scale=a_cs))
        assert ops.error_flag() == 0
    res["env"] = {k: v for k, v in os.environ.items() if k.startswith("NN_UMMA")}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
