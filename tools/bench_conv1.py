#!/usr/bin/env python
"""conv1 forward (3 -> 65 channels, 5x5, batch 512) main-kernel time: shift-GEMM vs tiled kernel, and the share of
the noise generation (plain / Philox noise / injected z).  CUDA events around the main kernel only."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from noisynet_b200 import _lib, ops  # noqa: E402
from noisynet_b200._lib import NOISE_MERGED, NOISE_NONE  # noqa: E402


def main():
    B = int(os.environ.get("B", "512"))
    dev = torch.device("cuda:0")
    lib = _lib.load()
    s_a = 1.0 / 15.0
    xs = [(torch.randint(0, 16, (B, 3, 32, 32), device=dev).float() * s_a) for _ in range(6)]
    w_raw = torch.randn(65, 3, 5, 5, device=dev) * 0.1
    wq = ops.quantize_fwd(w_raw, 4, -1.0, 1.0, 0.0)
    scale = ops.tensor_stats(w_raw)[1:2]
    z = torch.randn(B, 65, 28, 28, device=dev)
    res = {}
    for shift in (1, 0):
        lib.nn_debug_shift_enable(shift)
        for name, kw in (("plain", dict(noise_mode=NOISE_NONE)),
                         ("noisy", dict(noise_mode=NOISE_MERGED, current=1.0, scale_dev=scale, want_y=False)),
                         ("inject", dict(noise_mode=NOISE_MERGED, current=1.0, scale_dev=scale, want_y=False, z=z))):
            def fwd(i):
                return ops.noisy_conv_fwd(xs[i % 6], wq, w_raw, None, 1, 0, precision="bf16", a_code_scale=s_a,
                                          w_code_scale=1.0 / 15.0, **kw)
            for i in range(3):
                fwd(i)
            lib.nn_debug_main_kernel_ms(1)
            t = []
            for i in range(20):
                fwd(i)
                t.append(lib.nn_debug_main_kernel_ms(-1))
            lib.nn_debug_main_kernel_ms(0)
            res["%s_%s" % ("shift" if shift else "tiled", name)] = round(1e3 * sorted(t)[len(t) // 2], 1)
    assert ops.error_flag() == 0
    lib.nn_debug_shift_enable(1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
