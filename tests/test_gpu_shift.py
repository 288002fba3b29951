"""The persistent shift-GEMM forward (k_conv_shift: narrow-input layers, weights resident in shared memory, the
im2col rows read through shifted SWIZZLE_NONE descriptors) against the tiled tcgen05 kernel, the fp32 CUDA-core
kernel and a float64 evaluation.

Stated tolerances (same as test_gpu_umma.py): integer-code main contraction exact up to the final scale multiply
(rtol 1e-6); sigma within rtol 3e-3 of fp32; the Philox draws are the same stream with the same (m, n) -> group
mapping as every other path, so with equal sigma the noisy outputs agree to the sigma tolerance.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    import __graft_entry__ as entry
    entry.build()
    return torch.device("cuda:0")


SHAPES = [  # B, Cin, H, W, Cout, k
    (4, 3, 32, 32, 65, 5),        # NoisyNet conv1 (25 taps: odd -> one padding tap)
    (3, 3, 32, 32, 9, 5),         # the narrow golden net
    (2, 1, 28, 28, 16, 4),        # MNIST-like, even tap count, 784 pixels/image: tiles straddle images
    (5, 8, 9, 11, 120, 3),        # Cin = 8 exactly, W not a divisor of 128, widest accumulator (2 x 120 -> 240 cols)
    (1, 3, 5, 5, 4, 5),           # single output pixel, fewer pixels than one tile
    (300, 2, 12, 12, 33, 3),      # more tiles than SMs: several tiles per CTA, both accumulator buffers reused
]


def _mk(shape, gen):
    B, Cin, H, W, Cout, k = shape
    s_a = 5.0 / 15.0
    ka = torch.randint(0, 16, (B, Cin, H, W), generator=gen).float()
    ka = ka * (torch.rand(ka.shape, generator=gen) > 0.3).float()
    cw = (torch.randint(0, 16, (Cout, Cin, k, k), generator=gen) * 2 - 15).float()
    wq = (cw * (2.0 / 15.0) / 2.0).float()
    w_raw = torch.randn(Cout, Cin, k, k, generator=gen) * 0.3
    return s_a, ka, (ka * s_a).float(), cw, wq, w_raw


@pytest.mark.parametrize("shape", SHAPES)
def test_shift_matches_tiled_and_exact(dev, shape):
    from noisynet_b200 import _lib, ops
    from noisynet_b200._lib import NOISE_EXTERNAL, NOISE_MERGED, NOISE_NONE, PACK_SHIFT, PREC_BF16, ConvGeom
    import ctypes as C
    lib = _lib.load()
    B, Cin, H, W, Cout, k = shape
    g = ConvGeom(B, Cin, H, W, Cout, k, k, 1, 0)
    assert lib.nn_conv_pack_layout(C.byref(g), NOISE_MERGED, PREC_BF16) == PACK_SHIFT
    gen = torch.Generator().manual_seed(sum(shape))
    s_a, ka, x, cw, wq, w_raw = _mk(shape, gen)
    xd, wqd, wrd = x.to(dev), wq.to(dev), w_raw.to(dev)
    exact = F.conv2d(ka.double(), cw.double()) * (float(np.float32(s_a)) * float(np.float32(1.0 / 15.0)))
    kw = dict(precision="bf16", a_code_scale=s_a, w_code_scale=1.0 / 15.0)
    try:
        # plain (no noise)
        r = ops.noisy_conv_fwd(xd, wqd, None, None, 1, 0, noise_mode=NOISE_NONE, **kw)
        assert ops.error_flag() == 0
        assert torch.allclose(r["y"].cpu().double(), exact, rtol=1e-6, atol=1e-9)
        for mode in (NOISE_MERGED, NOISE_EXTERNAL):
            scale = ops.tensor_stats(wrd)[1:2] if mode == NOISE_MERGED else ops.tensor_stats(xd)[0:1]
            common = dict(noise_mode=mode, current=1.0, scale_dev=scale)
            lib.nn_debug_shift_enable(1)
            a = ops.noisy_conv_fwd(xd, wqd, wrd, None, 1, 0, rng=ops._fixed_rng(11, 5), **common, **kw)
            assert ops.error_flag() == 0
            lib.nn_debug_shift_enable(0)
            b = ops.noisy_conv_fwd(xd, wqd, wrd, None, 1, 0, rng=ops._fixed_rng(11, 5), want_z=True, want_sigma=True,
                                   **common, **kw)
            assert ops.error_flag() == 0
            assert torch.allclose(a["y"].cpu().double(), exact, rtol=1e-6, atol=1e-9)
            assert torch.equal(a["y"], b["y"])
            # same z; sigma from the same bf16 operands, only the fp32 summation order may differ
            tol = 1e-5 * float(b["sigma"].abs().max()) * float(b["z"].abs().max()) + 1e-6
            assert (a["y_noisy"] - b["y_noisy"]).abs().max().item() <= tol
            # injected z (parity hook variant of the kernel)
            lib.nn_debug_shift_enable(1)
            c = ops.noisy_conv_fwd(xd, wqd, wrd, None, 1, 0, z=b["z"], **common, **kw)
            assert ops.error_flag() == 0
            assert (c["y_noisy"] - a["y_noisy"]).abs().max().item() <= tol
            # fp32 CUDA-core kernel: sigma tolerance
            f = ops.noisy_conv_fwd(xd, wqd, wrd, None, 1, 0, z=b["z"], want_sigma=True, precision="fp32", **common)
            assert torch.allclose(b["sigma"], f["sigma"], rtol=3e-3, atol=1e-6)
            lim = 3e-3 * (f["sigma"] * b["z"].abs()) + 2e-5 * f["y"].abs().max()
            assert ((a["y_noisy"] - f["y_noisy"]).abs() <= lim).all()
    finally:
        lib.nn_debug_shift_enable(1)


def test_shift_layout_rejected_when_not_served(dev):
    from noisynet_b200 import _lib
    from noisynet_b200._lib import NOISE_MERGED, PACK_SHIFT, PACK_TILED, PREC_BF16, ConvGeom
    import ctypes as C
    lib = _lib.load()
    for g in (ConvGeom(4, 65, 14, 14, 120, 5, 5, 1, 0), ConvGeom(4, 3, 32, 32, 65, 5, 5, 1, 2),
              ConvGeom(4, 3, 32, 32, 65, 5, 5, 2, 0), ConvGeom(4, 3, 32, 32, 200, 5, 5, 1, 0)):
        assert lib.nn_conv_pack_layout(C.byref(g), NOISE_MERGED, PREC_BF16) == PACK_TILED
    g = ConvGeom(4, 3, 32, 32, 200, 5, 5, 1, 0)
    assert lib.nn_conv_pack_layout(C.byref(g), 0, PREC_BF16) == PACK_SHIFT      # 200 plain columns fit, 400 do not
