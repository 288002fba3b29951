"""The UNMODIFIED reference script on top of the drop-in modules (only where the reference tree exists, i.e. the
build container; skipped on the GPU box).  Without a GPU the run must get through argument parsing, data
loading, `Net(args)` construction with OUR NoisyConv2d / NoisyLinear / QuantMeasure, init, optimizer setup, and
stop at the first kernel call with the loud no-CPU-fallback error -- which proves the import boundary
(noisynet.py:14), the constructor signatures (noisynet.py:344-359) and the module attributes the script touches."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NOISYNET_REFERENCE", "/root/reference")

RUNNER = textwrap.dedent('''
    import sys, types, runpy, collections.abc, torch
    sys.path.insert(0, {ref!r})
    sys.path.insert(0, {dropin!r})                     # OUR hardware_model / quant / plot_histograms first
    six = types.ModuleType('torch._six'); six.container_abcs = collections.abc
    sys.modules['torch._six'] = six                    # import rot of the reference under torch 2.x
    mpl = types.ModuleType('matplotlib'); mpl.use = lambda *a, **k: None
    plt = types.ModuleType('matplotlib.pyplot'); mpl.pyplot = plt
    sys.modules['matplotlib'] = mpl; sys.modules['matplotlib.pyplot'] = plt
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    import hardware_model, quant
    assert 'noisynet_b200' in hardware_model.NoisyConv2d.__module__, hardware_model.__file__
    assert 'noisynet_b200' in quant.QuantMeasure.__module__
    sys.argv = ['noisynet.py'] + {argv!r}
    runpy.run_path({script!r}, run_name='__main__')
''')


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "noisynet.py")), reason="reference tree not present")
def test_unmodified_noisynet_script_reaches_our_kernels(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-side boundary check")
    os.makedirs(tmp_path / "data")
    np.savez(tmp_path / "data" / "cifar_RGB_4bit.npz", np.zeros((50000, 3072), np.uint8), np.zeros(50000, np.int64),
             np.zeros((10000, 3072), np.uint8), np.zeros(10000, np.int64))
    argv = ["--current", "1", "--act_max", "5", "--w_max1", "0.3", "--LR", "0.005", "--L2_1", "0.0005", "--L2_2", "0.0002",
            "--q_a", "4", "--q_w", "4", "--batch_size", "16", "--nepochs", "1", "--no-augment"]
    code = RUNNER.format(ref=REF, dropin=os.path.join(ROOT, "dropin"), argv=argv, script=os.path.join(REF, "noisynet.py"))
    r = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    tail = (r.stdout[-1500:] + r.stderr[-3000:])
    assert r.returncode != 0, tail
    assert "NoisyNetLibraryError" in r.stderr and "there is no CPU fallback" in r.stderr, tail
    # it failed inside OUR module's forward, called from the script's training loop
    assert "noisynet_b200/hardware_model.py" in r.stderr and "noisynet.py" in r.stderr, tail
