"""The reference script's training loop on the drop-in modules, on the GPU.

`tests/golden/script_trace.npz` holds the loss curve of the UNMODIFIED reference script (noisynet.py, q_a = q_w = 4,
I = 1 nA, batch 64, --seed 0) over its first 60 steps on a deterministic synthetic dataset, minted in the build
container by oracle/gen_script_trace.py.  The reference tree does not exist on the GPU box, so:

  * test_script_equivalent_loop_matches_reference_trace drives the drop-in modules (NoisyConv2d / NoisyLinear /
    QuantMeasure / add_noise_calculate_power in the script's two-call flow, the script's AdamW groups, per-step weight
    clamp, np.random permutation, `input.max()` live ranges of quantize2 / quantize4) through the SAME steps on the
    same data and compares the loss curve with the reference's own.  The generators differ (Philox streams of this
    library vs. ATen's), so the comparison is by windows: mean loss of steps 0-9, 25-34 and 50-59 within 0.12 of the
    reference's, and the same downward trend;
  * test_unmodified_script_on_gpu runs the real script under the drop-in modules when a reference tree IS present next
    to a GPU (a maintainer's machine): same windows.
"""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NOISYNET_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _windows(losses):
    l = np.asarray(losses, dtype=np.float64)
    return np.array([l[0:10].mean(), l[25:35].mean(), l[50:60].mean()])


def _check_against_trace(losses, trace):
    ref = trace["losses"]
    assert len(losses) == len(ref) == 60 and np.all(np.isfinite(losses))
    w, wr = _windows(losses), _windows(ref)
    assert np.all(np.abs(w - wr) <= 0.12), (w, wr)
    assert w[2] < w[0] - 0.15 and wr[2] < wr[0] - 0.15, (w, wr)          # both learn


def test_script_equivalent_loop_matches_reference_trace(golden):
    from gen_script_trace import synthetic_learnable_cifar
    from noisynet_b200 import ops
    from noisynet_b200.net import NoisyNet, default_args, init_like_reference, make_optimizer, with_quant
    import __graft_entry__ as entry
    entry.build()
    trace = golden("script_trace")
    B, steps = int(trace["batch"]), int(trace["steps"])
    dev = torch.device("cuda:0")
    x, y = synthetic_learnable_cifar(50000, seed=int(trace["data_seed"]))
    train_inputs = torch.from_numpy(x.reshape(50000, 3, 32, 32)).to(dev)          # utils.py:143-153: whole set on the device
    train_labels = torch.from_numpy(y).to(dev)
    torch.manual_seed(0)                                                             # --seed 0 (noisynet.py:314-317)
    np.random.seed(int(trace["np_seed"]))
    a = with_quant(default_args(), 4, 4)
    model = init_like_reference(NoisyNet(a, fused=False, precision="bf16")).to(dev)  # the script's flow: layer, then noise call
    opt = make_optimizer(model, a)                                                   # noisynet.py:1135-1169
    model.train()
    rnd_idx = np.random.permutation(len(train_inputs))                               # noisynet.py:1232-1234
    train_inputs, train_labels = train_inputs[rnd_idx], train_labels[rnd_idx]
    losses = []
    for i in range(steps):
        inp = train_inputs[i * B:(i + 1) * B]
        lab = train_labels[i * B:(i + 1) * B]
        out = model(inp, 0, i)                                                       # :1276 (i < 20: side statistics collected)
        loss = torch.nn.CrossEntropyLoss()(out, lab)                                 # :1278
        opt.zero_grad()                                                              # :1346
        loss.backward()                                                              # :1372
        opt.step()                                                                   # :1520
        model.clamp_weights_()                                                       # :1527-1542
        losses.append(loss.item())
    assert ops.error_flag() == 0
    assert len(model.power[0]) == 20 and len(model.nsr[3]) == 20 and all(np.isfinite(model.power[1]))
    _check_against_trace(losses, trace)


RUNNER = textwrap.dedent('''
    import sys, types, runpy, collections.abc, json
    import numpy as np, torch
    sys.path.insert(0, {ref!r})
    sys.path.insert(0, {dropin!r})
    six = types.ModuleType('torch._six'); six.container_abcs = collections.abc
    sys.modules['torch._six'] = six
    mpl = types.ModuleType('matplotlib'); mpl.use = lambda *a, **k: None
    plt = types.ModuleType('matplotlib.pyplot'); mpl.pyplot = plt
    sys.modules['matplotlib'] = mpl; sys.modules['matplotlib.pyplot'] = plt
    import hardware_model
    assert 'noisynet_b200' in hardware_model.NoisyConv2d.__module__
    np.random.seed(0)
    losses = []
    class Stop(Exception):
        pass
    _fwd = torch.nn.CrossEntropyLoss.forward
    def fwd(self, out, lab):
        l = _fwd(self, out, lab)
        if torch.is_grad_enabled():
            losses.append(float(l))
        return l
    torch.nn.CrossEntropyLoss.forward = fwd
    _step = torch.optim.AdamW.step
    count = [0]
    def step(self, *a, **k):
        r = _step(self, *a, **k)
        count[0] += 1
        if count[0] >= {steps}:
            raise Stop()
        return r
    torch.optim.AdamW.step = step
    sys.argv = ['noisynet.py'] + {argv!r}
    try:
        runpy.run_path({script!r}, run_name='__main__')
    except Stop:
        pass
    json.dump(losses, open('losses.json', 'w'))
''')


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "noisynet.py")), reason="reference tree not present on this machine")
def test_unmodified_script_on_gpu(golden, tmp_path):
    from gen_script_trace import synthetic_learnable_cifar
    trace = golden("script_trace")
    os.makedirs(tmp_path / "data")
    x, y = synthetic_learnable_cifar(50000, seed=int(trace["data_seed"]))
    xt, yt = synthetic_learnable_cifar(10000, seed=int(trace["data_seed"]) + 1)
    np.savez(tmp_path / "data" / "cifar_RGB_4bit.npz", x, y, xt, yt)
    code = RUNNER.format(ref=REF, dropin=os.path.join(ROOT, "dropin"), steps=int(trace["steps"]),
                         argv=[str(v) for v in trace["argv"]], script=os.path.join(REF, "noisynet.py"))
    r = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, capture_output=True, text=True, timeout=1800)
    assert os.path.isfile(tmp_path / "losses.json"), r.stdout[-1500:] + r.stderr[-3000:]
    _check_against_trace(json.load(open(tmp_path / "losses.json")), trace)
