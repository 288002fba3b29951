#!/usr/bin/env python
"""Debug: engine forward vs the CPU oracle, tensor by tensor (same injected random draws)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.nn.functional as F

from oracle import noisynet_oracle as O
from noisynet_b200.engine import NoisyNetEngine
from noisynet_b200.net import NoisyNet, default_args, make_fused_optimizer, with_quant
from test_gpu_net import _make_rnd

dev = torch.device("cuda:0")
CUR = float(os.environ.get("CUR", "1.0"))
widths, B, q = dict(fm1=9, fm2=12, fc=24), 8, 4
oa = O.default_args(q_a=q, q_w=q, quant_max2=4.0, quant_max4=4.5, current=CUR, **widths)
torch.manual_seed(1)
om = O.OracleNet(oa).init_like_reference()
na = with_quant(default_args(layer_currents=[CUR] * 4, **widths), q, q)
nm = NoisyNet(na, fused=True, precision="bf16").to(dev)
nm.load_state_dict(om.state_dict(), strict=False)
nm.quantize2.running_max = torch.tensor(4.0, device=dev)
nm.quantize4.running_max = torch.tensor(4.5, device=dev)
om.train(), nm.train()
eng = NoisyNetEngine(nm, B, opt=None)
x, lab = O.synthetic_cifar(B, seed=10)
rnd = _make_rnd(oa, B, q, 100)

# oracle forward, step by step
a = oa
with torch.no_grad():
    xq = O.uniform_quantize_fwd(x, 4, 0.0, 1.0, 0.5, rnd["ua1"])
    w0 = O.uniform_quantize_fwd(om.conv1.weight, 4, -1.0, 1.0, 0.5, rnd["uw0"])
    y1 = F.conv2d(xq, w0)
    y1n = y1 if CUR <= 0 else O.add_noise_calculate_power(xq, om.conv1.weight, y1, CUR, "conv", True, z=rnd["z0"], want_stats=False)[0]
    p1 = F.max_pool2d(y1n, 2, 2)
    bn1 = F.batch_norm(p1, None, None, om.bn1.weight, om.bn1.bias, True, 0.1, 1e-5)
    h1 = torch.clamp(F.relu(bn1), max=5.0)
    q2 = O.uniform_quantize_fwd(h1, 4, 0.0, 4.0, 0.5, rnd["ua2"])
    w1 = O.uniform_quantize_fwd(om.conv2.weight, 4, -1.0, 1.0, 0.5, rnd["uw1"])
    y2 = F.conv2d(q2, w1)
    y2n = y2 if CUR <= 0 else O.add_noise_calculate_power(q2, om.conv2.weight, y2, CUR, "conv", False, z=rnd["z1"], want_stats=False)[0]
    p2 = F.max_pool2d(y2n, 2, 2)
    bn2 = F.batch_norm(p2, None, None, om.bn2.weight, om.bn2.bias, True, 0.1, 1e-5)
    h2 = torch.clamp(F.relu(bn2), max=5.0).view(B, -1)
    q3 = O.uniform_quantize_fwd(h2, 4, 0.0, 5.0, 0.5, rnd["ua3"])
    w2 = O.uniform_quantize_fwd(om.linear1.weight, 4, -1.0, 1.0, 0.5, rnd["uw2"])
    l1 = F.linear(q3, w2)
    bn3 = F.batch_norm(l1, None, None, om.bn3.weight, om.bn3.bias, True, 0.1, 1e-5)
    h3 = torch.clamp(F.relu(bn3), max=5.0)
    q4 = O.uniform_quantize_fwd(h3, 4, 0.0, 4.5, 0.5, rnd["ua4"])
    w3 = O.uniform_quantize_fwd(om.linear2.weight, 4, -1.0, 1.0, 0.5, rnd["uw3"])
    l2 = F.linear(q4, w3)

eng.inject = dict(u=[rnd[k].to(dev) for k in ("ua1", "ua2", "ua3", "ua4")], uw=[rnd[k].to(dev) for k in ("uw0", "uw1", "uw2", "uw3")],
                  z=[rnd[k].to(dev) for k in ("z0", "z1", "z2", "z3")] if CUR > 0 else [])
loss = eng.train_step(x.to(dev), lab.to(dev))
torch.cuda.synchronize()


def cmp(name, a_, b_):
    d = (a_ - b_).abs()
    print("%-10s max|ref| %.4f  max err %.5f  mean err %.6f  frac>1e-3 %.4f" % (name, b_.abs().max().item(), d.max().item(), d.mean().item(), (d > 1e-3).float().mean().item()))


xp1 = eng.xp1.float().cpu()[..., :3].permute(0, 3, 1, 2) / 15.0
cmp("xq", xp1, xq)
cmp("wq0", eng.wq[0].cpu(), w0)
cmp("y1n", eng.y1n.cpu(), y1n)
cmp("pool1", eng.pool1.cpu(), p1)
xp2 = eng.xp2.float().cpu()[..., :9].permute(0, 3, 1, 2) * (4.0 / 15)
cmp("q2", xp2, q2)
cmp("wq1", eng.wq[1].cpu(), w1)
cmp("y2n", eng.y2n.cpu(), y2n)
print("xmax2", eng.xmax2.item(), q2.max().item())
if CUR <= 0:
    cmp("pool2", eng.pool2.cpu(), p2)
    xp3 = eng.xp3.float().cpu()[..., :12].permute(0, 3, 1, 2).reshape(B, -1) * (5.0 / 15)
    cmp("q3", xp3, q3)
    cmp("l1", eng.l1n.cpu(), l1)
    xp4 = eng.xp4.float().cpu()[..., :24] * (4.5 / 15)
    cmp("q4", xp4, q4)
    cmp("l2", eng.l2n.cpu(), l2)
oloss, _ = O.train_step(om, O.make_optimizer(om, oa), x, lab, i=100, rnd=rnd)
print("loss", loss.item(), oloss.item())
if CUR <= 0:
    og = dict(om.named_parameters())
    for k, p_ in nm.named_parameters():
        g_ref = og[k].grad
        d = (p_.grad.cpu() - g_ref).abs()
        print("grad %-16s max|ref| %.5f  max err %.5f  rel-L2 %.4f" % (k, g_ref.abs().max().item(), d.max().item(),
              ((p_.grad.cpu() - g_ref).norm() / (g_ref.norm() + 1e-12)).item()))
if CUR <= 0:
    from noisynet_b200 import ops
    gy2 = eng.gyp2.float()[..., :12].permute(0, 3, 1, 2).contiguous()          # NCHW fp32 view of the packed gradient
    x2 = q2.to(dev)
    ref = og["conv2.weight"].grad
    for prec in ("fp32", "bf16"):
        gw = ops.conv_wgrad(gy2, x2, (12, 9, 5, 5), precision=prec, w_raw=nm.conv2.weight.detach(), w_lo=-1.0, w_hi=1.0).cpu()
        print("conv2 wgrad from unpacked gyp2 (%s): rel-L2 vs oracle %.4f, vs engine %.4f" % (
            prec, ((gw - ref).norm() / ref.norm()).item(), ((gw - nm.conv2.weight.grad.cpu()).norm() / ref.norm()).item()))
    # oracle's own gradient w.r.t. y2n: recompute with autograd
    y2r = y2n.clone().requires_grad_(True)
    p2r = F.max_pool2d(y2r, 2, 2)
    bn2r = F.batch_norm(p2r, None, None, om.bn2.weight.detach(), om.bn2.bias.detach(), True, 0.1, 1e-5)
    h2r = torch.clamp(F.relu(bn2r), max=5.0)
    q3r = O.OracleNet._STEQuant.apply(h2r, 4, 0.0, 5.0, 0.5, rnd["ua3"].view(B, 12, 5, 5))
    q3r.backward(eng.gx3.cpu())
    d = (gy2.cpu() - y2r.grad)
    print("gyp2 vs autograd through stage 2 (given engine gx3): rel-L2 %.4f max %.5f" % ((d.norm() / y2r.grad.norm()).item(), d.abs().max().item()))
