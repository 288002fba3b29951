"""data path and formats.

  * nn_input_gather_quant_pack == the script's torch batch assembly (permute, slice, crop at a random offset of the
    zero-padded images, flip: noisynet.py:1232-1269, utils.py:165-167) followed by quantize1 + the NHWC code pack: bit exact;
  * the ``cifar_RGB_4bit.npz`` loader (utils.py:130-176);
  * checkpoint round trip: the state_dict of the drop-in model after engine training steps loads into the oracle's
    restatement of the reference Net (same keys, noisynet.py:979-1002 / :1636) and evaluates identically, and back;
  * the engine trains through the gather path (loss decreases on learnable synthetic data).
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import noisynet_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    import __graft_entry__ as entry
    entry.build()
    return torch.device("cuda:0")


def test_npz_loader_and_gather_kernel(dev, tmp_path):
    import ctypes as C
    from gen_script_trace import synthetic_learnable_cifar
    from noisynet_b200 import _lib, data
    x, y = synthetic_learnable_cifar(600, seed=5)
    xt, yt = synthetic_learnable_cifar(100, seed=6)
    path = str(tmp_path / "cifar_RGB_4bit.npz")
    np.savez(path, x, y, xt, yt)
    tr, trl, te, tel = data.load_cifar_npz(path, dev, augment=True)
    assert tr.shape == (600, 3, 40, 40) and te.shape == (100, 3, 32, 32) and trl.dtype == torch.int64
    assert torch.equal(tr[:, :, 4:36, 4:36].cpu(), torch.from_numpy(x.reshape(600, 3, 32, 32))) and tr[:, :, :4].abs().sum() == 0
    lib = _lib.load()
    B = 37
    np.random.seed(3)
    import random
    random.seed(3)
    sched = data.EpochBatches(600, B, trl, dev, augment=True)
    n = 0
    for idx, aug, lab in sched.epoch():
        ref = data.reference_batch(tr, idx, aug)                     # the script's torch ops
        assert torch.equal(lab, trl[idx])
        u = (torch.rand(B, 3, 32, 32, device=dev) - 0.5)
        xp = torch.zeros(B, 32, 32, 8, dtype=torch.bfloat16, device=dev)
        act = torch.empty(B, 3, 32, 32, device=dev)
        _lib.check(lib.nn_input_gather_quant_pack(tr.data_ptr(), idx.data_ptr(), B, 3, 40, 40, 32, 32, 0, 0, 0, aug.data_ptr(),
                                                  xp.data_ptr(), act.data_ptr(), 8, 4, 1.0, 0.5, u.data_ptr(), _lib.Rng(0, 0, None), 0,
                                                  torch.cuda.current_stream().cuda_stream), "nn_input_gather_quant_pack")
        want = O.uniform_quantize_fwd(ref.cpu(), 4, 0.0, 1.0, 0.5, u.cpu())
        assert torch.equal(act.cpu(), want)
        codes = torch.round(want * 15.0)
        assert torch.equal(xp[..., :3].float().cpu(), codes.permute(0, 2, 3, 1))
        n += 1
        if n == 3:
            break
    assert n == 3


def test_checkpoint_round_trip_and_training_through_gather(dev, tmp_path):
    from gen_script_trace import synthetic_learnable_cifar
    from noisynet_b200 import data, ops
    from noisynet_b200.engine import NoisyNetEngine
    from noisynet_b200.net import NoisyNet, default_args, init_like_reference, make_fused_optimizer, with_quant
    widths = dict(fm1=16, fm2=24, fc=48)
    B = 64
    x, y = synthetic_learnable_cifar(64 * 40, seed=7)
    tr = torch.nn.functional.pad(torch.from_numpy(x.reshape(-1, 3, 32, 32)), (4, 4, 4, 4)).to(dev)
    trl = torch.from_numpy(y).to(dev)
    torch.manual_seed(0)
    na = with_quant(default_args(**widths), 4, 4)
    nm = init_like_reference(NoisyNet(na, fused=True, precision="bf16")).to(dev)
    nm.quantize2.running_max = torch.tensor(5.0, device=dev)
    nm.quantize4.running_max = torch.tensor(5.0, device=dev)
    nm.train()
    eng = NoisyNetEngine(nm, B, opt=make_fused_optimizer(nm, na))
    np.random.seed(0)
    import random
    random.seed(0)
    losses = []
    sched = data.EpochBatches(tr.shape[0], B, trl, dev, augment=True)
    for ep in range(3):
        for idx, aug, lab in sched.epoch():
            losses.append(eng.train_step(tr, lab, gather=(idx, aug)).item())
    assert ops.error_flag() == 0
    assert np.isfinite(losses).all() and np.mean(losses[-10:]) < np.mean(losses[:10]) - 0.2, (losses[:10], losses[-10:])
    eng.sync_bn_counters()
    # ---- checkpoint round trip through the reference model's key set
    ck = str(tmp_path / "model.pth")
    torch.save({"model": nm.state_dict()}, ck)
    sd = torch.load(ck, map_location="cpu")["model"]
    oa = O.default_args(q_a=4, q_w=4, quant_max2=5.0, quant_max4=5.0, current=0.0, **widths)
    om = O.OracleNet(oa)
    ref_keys = set(om.state_dict().keys())
    missing = ref_keys - set(sd.keys())
    assert not missing, missing                                   # every key of the reference Net is in our checkpoint
    om.load_state_dict({k: v for k, v in sd.items() if k in ref_keys})
    assert int(om.bn1.num_batches_tracked) == len(losses)
    om.eval(), nm.eval()
    xe = torch.from_numpy(x[:B].reshape(B, 3, 32, 32))
    with torch.no_grad():
        want = om(xe, i=100)
    got = eng.eval_forward(xe.to(dev), currents=[1e12] * 4).cpu()                 # noise scaled to nothing: compare the clean path
    assert torch.allclose(got, want, rtol=2e-3, atol=2e-3), (got - want).abs().max()
    # and back: the oracle's state into a fresh drop-in model
    nm2 = NoisyNet(na, fused=True, precision="bf16").to(dev)
    res = nm2.load_state_dict(om.state_dict(), strict=False)
    assert not res.unexpected_keys
    assert torch.equal(nm2.conv2.weight.cpu(), om.conv2.weight) and torch.equal(nm2.bn3.running_var.cpu(), om.bn3.running_var)
