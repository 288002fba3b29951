"""Pins the CPU oracle (oracle/noisynet_oracle.py) against golden vectors minted from the
UNMODIFIED reference (oracle/gen_golden.py).  CPU only."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import noisynet_oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_quant_kat(golden):
    g = golden("quant")
    for name in ("hm", "q"):
        x = T(g[f"kat_{name}_x"])
        y = O.uniform_quantize_fwd(x, 4, 0.0, 5.0)
        assert torch.equal(y, T(g[f"kat_{name}_y"]))
        gin = O.uniform_quantize_bwd(x, torch.arange(1.0, 12.0), 0.0, 5.0)
        assert torch.equal(gin, T(g[f"kat_{name}_g"]))
    # SURVEY §8c known answers (half-to-even after fp32 division)
    assert np.allclose(g["kat_hm_y"], [0, 0, 0, 0.6666667, 0, 2.3333335, 2.6666667, 2.6666667, 5, 5, 5])
    assert np.array_equal(g["kat_hm_g"], [0, 2, 3, 4, 5, 6, 7, 8, 9, 0, 0])
    yw = O.uniform_quantize_fwd(T(g["wq_x"]), 4, -1.0, 1.0)
    assert torch.equal(yw, T(g["wq_y"]))
    assert np.array_equal(np.round(g["wq_y"] * 15), [-15, -15, -15, -7, -1, -1, 1, 1, 5, 13, 15, 15])
    assert torch.equal(O.uniform_quantize_fwd(T(g["floor_x"]), 4, 0.0, 0.0), T(g["floor_y"]))


def test_quant_random_cases(golden):
    g = golden("quant")
    for ci, (bits, lo, hi, st) in enumerate(g["r_params"]):
        x, u = T(g[f"r{ci}_x"]), T(g[f"r{ci}_u"])
        y = O.uniform_quantize_fwd(x, int(bits), lo, hi, st, u)
        assert torch.equal(y, T(g[f"r{ci}_y"])), ci
        gin = O.uniform_quantize_bwd(x, T(g[f"r{ci}_gout"]), lo, hi)
        assert torch.equal(gin, T(g[f"r{ci}_gin"])), ci


def test_quant_measure_ranges(golden):
    g = golden("quant")
    x = T(g["qm_run_x"])
    # calculate_running (activations): percentile recorded, live max used now (hardware_model.py:240-253)
    assert torch.equal(O.kth_percentile(x, 99.0), T(g["qm_run_pctl"]))
    y = O.uniform_quantize_fwd(x, 4, 0.0, x.max().item(), 0.5, T(g["qm_run_u"]))
    assert torch.equal(y, T(g["qm_run_y"]))
    # running_max frozen, eval (stochastic off)
    y = O.uniform_quantize_fwd(x, 4, 0.0, float(g["qm_run_pctl"]), 0.0)
    assert torch.equal(y, T(g["qm_fixed_y"]))
    # signed-weight percentile branch :232-239
    w = T(g["qm_w_x"])
    lo, hi = O.weight_percentile_range(w, 99.0)
    assert torch.equal(lo.reshape(()), T(g["qm_w_min"]).reshape(())) and torch.equal(hi, T(g["qm_w_max"]))
    assert torch.equal(O.uniform_quantize_fwd(w, 4, lo.item(), hi.item(), 0.0), T(g["qm_w_y"]))
    # quant.py variant: pctl is a fraction (quant.py:109)
    assert torch.equal(O.kth_percentile(x, 0.99 * 100.0), T(g["qq_pctl"]))
    y = O.uniform_quantize_fwd(x, 4, 0.0, x.max().item(), 0.5, T(g["qq_u"]))
    assert torch.equal(y, T(g["qq_y"]))


def test_add_weight_noise(golden):
    g = golden("layers")
    y = O.add_weight_noise(T(g["an_w"]), 0.3, T(g["an_u"]))
    assert torch.equal(y, T(g["an_y"]))


def _conv_case(g, ci, cfg):
    bw, nz, train, stride, pad, bias = int(cfg[0]), cfg[1], bool(cfg[2]), int(cfg[3]), int(cfg[4]), bool(cfg[5])
    x = T(g[f"conv{ci}_x"]).requires_grad_(True)
    w = T(g[f"conv{ci}_w"]).requires_grad_(True)
    b = T(g[f"conv{ci}_b"]).requires_grad_(True) if bias else None
    u = T(g[f"conv{ci}_u"])
    w_eff = O.effective_weight(w, bw, nz, 0.0, train, 0.5, u)
    if bw > 0:
        mask = ((w >= -1.0) & (w <= 1.0)).float()
        w_ste = w * mask + (w_eff - w * mask).detach()
    else:
        w_ste = w + (w_eff - w).detach()
    b_eff = b
    if bias and nz > 0 and train and bw == 0:
        b_eff = b + (O.add_weight_noise(b, nz, T(g[f"conv{ci}_ub"])) - b).detach()
    y = O.noisy_conv2d_fwd(x, w_ste, b_eff, stride, pad)
    y.backward(T(g[f"conv{ci}_gy"]))
    return x, w, y


def test_noisy_conv_fwd_bwd(golden):
    g = golden("layers")
    for ci, cfg in enumerate(g["conv_cfgs"]):
        x, w, y = _conv_case(g, ci, cfg)
        assert torch.allclose(y, T(g[f"conv{ci}_y"]), rtol=0, atol=1e-6), ci
        assert torch.allclose(x.grad, T(g[f"conv{ci}_gx"]), rtol=0, atol=1e-6), ci
        assert torch.allclose(w.grad, T(g[f"conv{ci}_gw"]), rtol=0, atol=2e-5), ci


def test_noisy_conv_with_input_quant(golden):
    g = golden("layers")
    x = T(g["cq_x"])
    w = T(g["cq_w"])
    xq = O.uniform_quantize_fwd(x, 4, 0.0, 2.0, 0.5, T(g["cq_ua"]))
    wq = O.uniform_quantize_fwd(w, 4, -1.0, 1.0, 0.5, T(g["cq_uw"]))
    y = F.conv2d(xq, wq, None, 1, 1)
    assert torch.allclose(y, T(g["cq_y"]), atol=1e-6, rtol=0)
    gy = T(g["cq_gy"])
    gx = torch.nn.grad.conv2d_input(x.shape, wq, gy, padding=1)
    gw = torch.nn.grad.conv2d_weight(xq, w.shape, gy, padding=1)
    assert torch.allclose(O.uniform_quantize_bwd(x, gx, 0.0, 2.0), T(g["cq_gx"]), atol=1e-5, rtol=0)
    assert torch.allclose(O.uniform_quantize_bwd(w, gw, -1.0, 1.0), T(g["cq_gw"]), atol=2e-5, rtol=0)


def test_noisy_linear(golden):
    g = golden("layers")
    for ci, cfg in enumerate(g["lin_cfgs"]):
        bw, nz, train, bias = int(cfg[0]), cfg[1], bool(cfg[2]), bool(cfg[3])
        x, w, u = T(g[f"lin{ci}_x"]), T(g[f"lin{ci}_w"]), T(g[f"lin{ci}_u"])
        w_eff = O.effective_weight(w, bw, nz, 0.0, train, 0.5, u, linear=True)
        b = T(g[f"lin{ci}_b"]) if bias else None
        if bias and f"lin{ci}_ub" in g:
            b = O.add_weight_noise(b, nz, T(g[f"lin{ci}_ub"]))
        y = O.noisy_linear_fwd(x, w_eff, b)
        assert torch.allclose(y, T(g[f"lin{ci}_y"]), atol=1e-6, rtol=0), ci
        gy = T(g[f"lin{ci}_gy"])
        assert torch.allclose(gy @ w_eff, T(g[f"lin{ci}_gx"]), atol=1e-5, rtol=0), ci
        gw = gy.t() @ x
        if 0 < bw < 8:
            gw = O.uniform_quantize_bwd(w, gw, -1.0, 1.0)
        assert torch.allclose(gw, T(g[f"lin{ci}_gw"]), atol=1e-5, rtol=0), ci


def test_current_noise(golden):
    g = golden("noise")
    for ci, (kind, merged, cur) in enumerate(g["n_meta"]):
        kind = "conv" if kind == 0 else "linear"
        x, w, y, z = (T(g[f"n{ci}_{k}"]) for k in "xwyz")
        out, st = O.add_noise_calculate_power(x, w, y, cur, kind, bool(merged), z=z)
        assert torch.allclose(out, T(g[f"n{ci}_out"]), atol=1e-6, rtol=0), ci
        ref = g[f"n{ci}_stats"]
        assert np.allclose([st["power"], st["nsr"], st["sparsity"]], ref, rtol=1e-5), ci


def test_alt_noise_models(golden):
    g = golden("noise")
    for kind in ("uniform_ind", "uniform_dep", "normal_ind", "normal_dep"):
        y = T(g[f"alt_{kind}_y"])
        out = O.alt_noise(y, kind, 0.2, T(g[f"alt_{kind}_rnd"]))
        assert torch.allclose(out, T(g[f"alt_{kind}"]), atol=2e-6, rtol=1e-6), kind


def _net_step(golden, tag, q):
    g = golden("net_step")
    a = O.default_args(q_a=q, q_w=q, quant_max2=4.0, quant_max4=4.5, fm1=9, fm2=12, fc=24)
    m = O.OracleNet(a)
    sd = {k[len(tag) + 5:]: T(v) for k, v in g.items() if k.startswith(f"{tag}_sd0_")}
    sd = {k: v for k, v in sd.items() if k in m.state_dict()}
    m.load_state_dict(sd)
    opt = O.make_optimizer(m, a)
    m.train()
    torch.manual_seed(31337)      # same generator state as the reference run: same draw order
    loss, logits = O.train_step(m, opt, T(g[f"{tag}_x"]), T(g[f"{tag}_label"]), i=0)
    return g, m, loss, logits


def test_net_step_fp(golden):
    g, m, loss, logits = _net_step(golden, "fp", 0)
    assert torch.allclose(logits, T(g["fp_logits"]), atol=2e-5, rtol=1e-5)
    assert abs(loss.item() - float(g["fp_loss"])) < 1e-5
    for k, v in m.state_dict().items():
        assert torch.allclose(v, T(g[f"fp_sd1_{k}"]), atol=1e-5, rtol=1e-4), k
    assert np.allclose([p[0] for p in m.power], g["fp_power"], rtol=1e-5)
    assert np.allclose([p[0] for p in m.nsr], g["fp_nsr"], rtol=1e-4)
    assert np.allclose([p[0] for p in m.input_sparsity], g["fp_sparsity"], rtol=1e-6)


def test_net_step_q4(golden):
    g, m, loss, logits = _net_step(golden, "q4", 4)
    assert torch.allclose(logits, T(g["q4_logits"]), atol=2e-5, rtol=1e-5)
    assert abs(loss.item() - float(g["q4_loss"])) < 1e-5
    for k, v in m.state_dict().items():
        assert torch.allclose(v, T(g[f"q4_sd1_{k}"]), atol=1e-5, rtol=1e-4), k


def test_philox_known_answer():
    # Philox4x32-10 known-answer vectors (Random123 kat_vectors): counter/key all zero and all ones.
    r = O.philox4x32_10(np.array([0], dtype=np.uint64), 0, 0)[0]
    assert [hex(v) for v in r] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    ff = np.uint64(0xFFFFFFFFFFFFFFFF)
    r = O.philox4x32_10(np.array([ff], dtype=np.uint64), int(ff), int(ff))[0]
    assert [hex(v) for v in r] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_philox_normal_moments():
    z = O.philox_normal_flat(400000, seed=1234, offset=5)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1.0) < 5e-3
    assert abs((z ** 3).mean()) < 2e-2 and abs((z ** 4).mean() - 3.0) < 5e-2
    u = O.philox_uniform_sym(400000, 99, 0, 0.5)
    assert u.min() >= -0.5 and u.max() < 0.5 and abs(u.mean()) < 2e-3
    zz = O.philox_normal_mn(7, 65, 1, 2)
    assert zz.shape == (7, 65)
