"""Host side of the data path on CPU: the cifar_RGB_4bit.npz format of utils.load_cifar
(utils.py:130-176), the per-epoch schedule of noisynet.py:1232-1269 (numpy permutation, python `random` crop offsets and
flips -- the generators the script seeds) and the torch restatement of the batch assembly the gather kernel is tested against
on the GPU (tests/test_gpu_data_path.py)."""
import random

import numpy as np
import torch

from noisynet_b200 import data


def _write_npz(path, n_train=40, n_test=12, seed=0):
    rs = np.random.RandomState(seed)
    tr_x = rs.randint(0, 16, (n_train, 3072)).astype(np.float32) / 15.0
    te_x = rs.randint(0, 16, (n_test, 3072)).astype(np.float32) / 15.0
    tr_y, te_y = rs.randint(0, 10, n_train), rs.randint(0, 10, n_test)
    np.savez(path, tr_x, tr_y, te_x, te_y)                 # positional: arr_0 .. arr_3, as utils.py:150-153 reads them
    return tr_x, tr_y, te_x, te_y


def test_npz_loader_shapes_and_padding(tmp_path):
    p = str(tmp_path / "cifar_RGB_4bit.npz")
    tr_x, tr_y, te_x, te_y = _write_npz(p)
    tr, ty, te, tey = data.load_cifar_npz(p, torch.device("cpu"), augment=True)
    assert tr.shape == (40, 3, 40, 40) and te.shape == (12, 3, 32, 32) and ty.dtype == torch.int64
    assert torch.equal(tr[:, :, 4:36, 4:36], torch.from_numpy(tr_x.reshape(-1, 3, 32, 32)))
    border = tr.clone()
    border[:, :, 4:36, 4:36] = 0
    assert float(border.abs().sum()) == 0.0                # nn.ZeroPad2d(4), utils.py:165-167
    assert torch.equal(tey, torch.from_numpy(te_y.astype(np.int64)))
    tr2, *_ = data.load_cifar_npz(p, torch.device("cpu"), augment=False)
    assert tr2.shape == (40, 3, 32, 32)


def test_epoch_schedule_follows_the_script_generators(tmp_path):
    labels = torch.arange(40) % 10
    sched = data.EpochBatches(40, 8, labels, torch.device("cpu"), augment=True)
    np.random.seed(3); random.seed(3)
    got = [(i.clone(), a.clone(), y.clone()) for i, a, y in sched.epoch()]
    np.random.seed(3); random.seed(3)
    perm = torch.from_numpy(np.random.permutation(40))     # noisynet.py:1232
    assert len(got) == 5
    for b, (idx, aug, y) in enumerate(got):
        k, j = random.randint(0, 8), random.randint(0, 8)  # :1265-1266
        flip = 1 if random.random() < 0.5 else 0           # :1268
        assert torch.equal(idx, perm[b * 8:(b + 1) * 8]) and aug.tolist() == [k, j, flip]
        assert torch.equal(y, labels[idx])
    # every sample exactly once per epoch
    assert sorted(torch.cat([g[0] for g in got]).tolist()) == list(range(40))
    plain = data.EpochBatches(40, 8, labels, torch.device("cpu"), augment=False)
    assert all(a.tolist() == [0, 0, 0] for _, a, _ in plain.epoch())


def test_reference_batch_is_slice_crop_flip(tmp_path):
    g = torch.Generator().manual_seed(1)
    x = torch.rand(10, 3, 40, 40, generator=g)
    idx = torch.tensor([7, 2, 2, 9])
    out = data.reference_batch(x, idx, torch.tensor([3, 5, 0], dtype=torch.int32))
    assert torch.equal(out, x[idx][:, :, 3:35, 5:37])
    out = data.reference_batch(x, idx, torch.tensor([8, 0, 1], dtype=torch.int32))
    assert torch.equal(out, torch.flip(x[idx][:, :, 8:40, 0:32], [3])) and out.is_contiguous()
    assert torch.equal(data.reference_batch(x[:, :, :32, :32], idx, torch.tensor([0, 0, 0], dtype=torch.int32), augment=False),
                       x[idx][:, :, :32, :32])
