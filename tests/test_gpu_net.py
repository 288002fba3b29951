"""Whole-net GPU parity: NoisyNet on the boundary modules (drop-in two-call flow AND the fused
single-launch flow) against the CPU oracle's restatement of noisynet.Net, which itself is pinned to
the reference's own Net by tests/test_oracle_golden.py::test_net_step_*.

All random draws (stochastic rounding u, current-noise z) are generated on the CPU, fed to the oracle
by name and injected into the CUDA ops in call order, so the comparison is elementwise.
Tolerance: fp32 CUDA-core path 2e-4 (logits/grads; BatchNorm amplifies summation-order noise).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import noisynet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    import __graft_entry__ as entry
    entry.build()
    return torch.device("cuda:0")


def _shapes(a, B):
    c1 = (B, a.fm1, 28, 28)
    c2 = (B, a.fm2, 10, 10)
    return dict(ua1=(B, 3, 32, 32), uw0=(a.fm1, 3, 5, 5), z0=c1, ua2=(B, a.fm1, 14, 14), uw1=(a.fm2, a.fm1, 5, 5),
                z1=c2, ua3=(B, a.fm2 * 25), uw2=(a.fc, a.fm2 * 25), z2=(B, a.fc), ua4=(B, a.fc), uw3=(10, a.fc),
                z3=(B, 10))


def _make_rnd(a, B, q, seed):
    g = torch.Generator().manual_seed(seed)
    rnd = {}
    for k, shp in _shapes(a, B).items():
        if k.startswith("z"):
            rnd[k] = torch.randn(shp, generator=g)
        elif q:
            rnd[k] = torch.rand(shp, generator=g) - 0.5
    return rnd


def _run_pair(dev, q, fused, widths, B, precision="fp32", steps=1):
    from noisynet_b200 import ops
    from noisynet_b200.net import NoisyNet, default_args, make_optimizer, train_step, with_quant
    oa = O.default_args(q_a=q, q_w=q, quant_max2=4.0, quant_max4=4.5, **widths)
    torch.manual_seed(1)
    om = O.OracleNet(oa).init_like_reference()
    na = default_args(**widths)
    if q:
        with_quant(na, q, q)
    nm = NoisyNet(na, fused=fused, precision=precision).to(dev)
    nm.load_state_dict({k: v for k, v in om.state_dict().items()}, strict=False)
    if q:
        nm.quantize2.running_max = torch.tensor(4.0, device=dev)
        nm.quantize4.running_max = torch.tensor(4.5, device=dev)
    oopt, nopt = O.make_optimizer(om, oa), make_optimizer(nm, na)
    om.train(), nm.train()
    out = []
    for s in range(steps):
        x, lab = O.synthetic_cifar(B, seed=10 + s)
        rnd = _make_rnd(oa, B, q, 100 + s)
        oloss, ologits = O.train_step(om, oopt, x, lab, i=s, rnd=rnd)
        order_u = [rnd[k] for k in ("ua1", "uw0", "ua2", "uw1", "ua3", "uw2", "ua4", "uw3")] if q else []
        order_z = [rnd[k] for k in ("z0", "z1", "z2", "z3")]
        with ops.inject_random([u.to(dev) for u in order_u], [z.to(dev) for z in order_z]) as inj:
            nloss, nlogits = train_step(nm, nopt, x.to(dev), lab.to(dev), i=s)
            assert not inj["u"] and not inj["z"], "not all injected draws were consumed"
        out.append((oloss, ologits, nloss.cpu(), nlogits.cpu()))
    return om, nm, out


def _assert_grads_and_weights(om, nm, gtol=3e-4):
    """Gradients: elementwise, relative to the tensor's max.  Updated weights: Adam's first update is
    +-lr*sign(g), so a gradient within rounding of zero may flip a weight by 2*lr -> fraction criterion."""
    ograds = dict(om.named_parameters())
    for k, p in nm.named_parameters():
        og = ograds[k].grad
        assert og is not None and p.grad is not None, k
        scale = og.abs().max().item() + 1e-12
        assert (p.grad.cpu() - og).abs().max().item() <= gtol * scale + 1e-7, (k, (p.grad.cpu() - og).abs().max().item(), scale)
    osd = om.state_dict()
    for k, v in nm.state_dict().items():
        if k in osd and v.dtype == torch.float32:
            bad = ((v.cpu() - osd[k]).abs() > 2e-4 + 2e-3 * osd[k].abs()).float().mean().item()
            assert bad <= 2e-3, (k, bad)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("q", [0, 4])
def test_train_step_parity_narrow(dev, q, fused):
    widths = dict(fm1=9, fm2=12, fc=24)
    om, nm, out = _run_pair(dev, q, fused, widths, B=8, steps=1)
    oloss, ologits, nloss, nlogits = out[0]
    assert torch.allclose(nlogits, ologits, rtol=2e-4, atol=2e-4), (nlogits - ologits).abs().max()
    assert abs(nloss.item() - oloss.item()) < 2e-4
    _assert_grads_and_weights(om, nm)
    # side statistics collected for i < 20 (power, nsr, sparsity) agree with the oracle's
    for name in ("power", "nsr", "input_sparsity"):
        a, b = np.asarray(getattr(nm, name)), np.asarray(getattr(om, name))
        assert a.shape == b.shape == (4, 1)
        assert np.allclose(a, b, rtol=2e-3), (name, a, b)


@pytest.mark.parametrize("q", [0, 4])
def test_train_step_parity_full_width(dev, q):
    """NoisyNet default widths (65/120/390), batch 16, fused flow."""
    om, nm, out = _run_pair(dev, q, True, {}, B=16, steps=1)
    oloss, ologits, nloss, nlogits = out[0]
    assert torch.allclose(nlogits, ologits, rtol=5e-4, atol=5e-4), (nlogits - ologits).abs().max()
    _assert_grads_and_weights(om, nm, gtol=1e-3)


def test_weight_clamp_projection(dev):
    from noisynet_b200.net import NoisyNet, default_args
    nm = NoisyNet(default_args(fm1=9, fm2=12, fc=24, w_max1=0.05, w_max2=0.01)).to(dev)
    nm.clamp_weights_()
    f32 = lambda v: torch.tensor(v, dtype=torch.float32).item()        # the clamp bound as the fp32 the kernel sees
    assert nm.conv1.weight.abs().max().item() <= f32(0.05) and nm.conv2.weight.abs().max().item() <= f32(0.01)
    assert nm.w_absmax[0][1].item() == pytest.approx(nm.conv1.weight.abs().max().item())


def test_training_reduces_loss(dev):
    """End-to-end sanity at I = 1 nA, 4-bit, in-kernel Philox: the loss of a small learnable problem falls."""
    from noisynet_b200.net import NoisyNet, default_args, init_like_reference, make_optimizer, train_step, with_quant
    torch.manual_seed(0)
    a = with_quant(default_args(), 4, 4)
    nm = init_like_reference(NoisyNet(a, fused=True)).to(dev)
    nm.quantize2.running_max = torch.tensor(5.0, device=dev)
    nm.quantize4.running_max = torch.tensor(5.0, device=dev)
    nm.collect_stats = False
    opt = make_optimizer(nm, a)
    g = torch.Generator().manual_seed(3)
    protos = torch.randint(0, 16, (10, 3, 32, 32), generator=g).float() / 15
    lab = torch.randint(0, 10, (256,), generator=g)
    x = (protos[lab] * 0.7 + 0.3 * torch.randint(0, 16, (256, 3, 32, 32), generator=g).float() / 15)
    x = (x * 15).round() / 15
    x, lab = x.to(dev), lab.to(dev)
    nm.train()
    losses = [train_step(nm, opt, x, lab, i=100 + s)[0].item() for s in range(30)]
    assert losses[-1] < 0.5 * losses[0], losses
    assert all(np.isfinite(losses))


def test_training_with_dropout_module_path(dev):
    """--dropout / --dropout_conv (noisynet.py:456, :512, :565: nn.Dropout between clamp and the next quantizer; quantize3's
    range is act_max / (1 - p), :346) on the module path: trains, drops ~p of the activations in train mode and nothing in eval."""
    from noisynet_b200.net import NoisyNet, default_args, init_like_reference, make_optimizer, train_step, with_quant
    torch.manual_seed(0)
    a = with_quant(default_args(dropout=0.2, dropout_conv=0.2), 4, 4)
    nm = init_like_reference(NoisyNet(a, fused=True)).to(dev)
    assert nm.quantize3.max_value == pytest.approx(5.0 / 0.8)
    nm.quantize2.running_max = torch.tensor(5.0, device=dev)
    nm.quantize4.running_max = torch.tensor(5.0, device=dev)
    nm.collect_stats = False
    opt = make_optimizer(nm, a)
    g = torch.Generator().manual_seed(3)
    protos = torch.randint(0, 16, (10, 3, 32, 32), generator=g).float() / 15
    lab = torch.randint(0, 10, (256,), generator=g)
    x = (protos[lab] * 0.7 + 0.3 * torch.randint(0, 16, (256, 3, 32, 32), generator=g).float() / 15)
    x = ((x * 15).round() / 15).to(dev)
    lab = lab.to(dev)
    seen = {}
    hook = nm.dropout.register_forward_hook(                # fraction of the nonzero activations that were dropped
        lambda m, i, o: seen.setdefault("z", []).append((((o == 0) & (i[0] != 0)).sum() / (i[0] != 0).sum().clamp(min=1)).item()))
    nm.train()
    losses = [train_step(nm, opt, x, lab, i=100 + s)[0].item() for s in range(30)]
    assert losses[-1] < 0.6 * losses[0] and all(np.isfinite(losses)), losses
    assert len(seen["z"]) == 90 and all(0.15 < f < 0.25 for f in seen["z"][-3:])            # three dropout sites per step
    seen["z"] = []
    nm.eval()
    with torch.no_grad():
        nm(x, 0, 100)
    hook.remove()
    assert all(f == 0.0 for f in seen["z"])


def test_fused_adamw_matches_torch(dev):
    """nn_adamw_step == torch.optim.AdamW + clamp_ over several steps (rtol 1e-5), incl. per-group lr / weight
    decay, the fused clamp and the max|W| side output."""
    from noisynet_b200.optim import FusedAdamW
    torch.manual_seed(0)
    shapes = [(65, 3, 5, 5), (120, 65, 5, 5), (390, 3000), (10,), (7,)]
    ps = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.2) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    groups = lambda xs: [{"params": [xs[0]], "lr": 0.005, "weight_decay": 0.0005, "clamp": 0.3},
                         {"params": [xs[1], xs[2]], "lr": 0.002, "weight_decay": 0.0, "clamp": 0.0},
                         {"params": [xs[3], xs[4]], "lr": 0.005, "weight_decay": 0.01, "clamp": 0.0}]
    ref = torch.optim.AdamW([{k: v for k, v in g.items() if k != "clamp"} for g in groups(qs)], lr=0.005)
    fus = FusedAdamW(groups(ps), lr=0.005, grad_scale=0.5)
    for step in range(5):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad = (g * 2.0).clone() if p.grad is None else p.grad.copy_(g * 2.0)     # grad_scale 0.5 undoes the 2x
            q.grad = g.clone()
        fus.step()
        ref.step()
        qs[0].data.clamp_(-0.3, 0.3)
        for p, q in zip(ps, qs):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), (step, (p - q).abs().max())
    assert fus.absmax_of(ps[0]).item() == pytest.approx(ps[0].abs().max().item())
    assert fus.absmax_of(ps[2]).item() == pytest.approx(ps[2].abs().max().item())
    assert fus.step_dev.item() == 5


def test_fused_adamw_partial_steps_equal_one_step(dev):
    """nn_adamw_step_part: a step split into an early subset (advance = 0) and the rest (advance = 1) gives bit-identical
    parameters, moments, max|W| and step count to the single launch; also across a step counter set from outside."""
    from noisynet_b200.optim import FusedAdamW
    torch.manual_seed(1)
    shapes = [(65, 3, 5, 5), (120, 65, 5, 5), (390, 3000), (10,), (7,)]
    ps = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.2) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    groups = lambda xs: [{"params": [xs[0]], "lr": 0.005, "weight_decay": 0.0005, "clamp": 0.3},
                         {"params": [xs[1], xs[2]], "lr": 0.002, "weight_decay": 0.0, "clamp": 0.25},
                         {"params": [xs[3], xs[4]], "lr": 0.005, "weight_decay": 0.01, "clamp": 0.0}]
    one, two = FusedAdamW(groups(qs), lr=0.005), FusedAdamW(groups(ps), lr=0.005)
    for step in range(6):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad = g.clone() if p.grad is None else p.grad.copy_(g)
            q.grad = g.clone() if q.grad is None else q.grad.copy_(g)
        if step == 3:                       # e.g. a checkpoint restored: the cached bias-correction constants are stale
            one.step_dev.fill_(40)
            two.step_dev.fill_(40)
        one.step()
        two.step_part([ps[2], ps[1]], advance=False)
        assert two.step_dev.item() == one.step_dev.item() - 1
        two.step_part([ps[0], ps[3], ps[4]], advance=True)
        for p, q in zip(ps, qs):
            assert torch.equal(p, q), (step, (p - q).abs().max())
            assert torch.equal(two.state[p]["exp_avg_sq"], one.state[q]["exp_avg_sq"])
        for p, q in zip(ps, qs):
            assert two.absmax_of(p).item() == one.absmax_of(q).item() == p.abs().max().item()
    assert two.step_dev.item() == one.step_dev.item() == 43
