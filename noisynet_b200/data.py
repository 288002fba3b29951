"""Data path of the CIFAR script on the device: the ``cifar_RGB_4bit.npz`` format of
utils.load_cifar (utils.py:130-176), the per-epoch permutation and the crop / flip augmentation of noisynet.py:1232-1269.

The reference keeps the whole dataset on the GPU, permutes all of it every epoch and slices batches; here the padded
dataset stays put and every batch is gathered by index, cropped and flipped inside the input-quantizer kernel
(nn_input_gather_quant_pack) -- see NoisyNetEngine.train_step(gather=...).
"""
import random

import numpy as np
import torch


def load_cifar_npz(path, device, augment=True, pad=4):
    """utils.load_cifar: arr_0 [50000 x 3072] floats in {0..15}/15, arr_1 labels, arr_2 / arr_3 the test split.
    Returns (train_inputs [N,3,32+2 pad,32+2 pad] zero-padded when augmenting, train_labels, test_inputs, test_labels) on ``device``."""
    with np.load(path) as f:
        tr_x = f["arr_0"].reshape(-1, 3, 32, 32).astype(np.float32)
        tr_y = f["arr_1"].astype(np.int64)
        te_x = f["arr_2"].reshape(-1, 3, 32, 32).astype(np.float32)
        te_y = f["arr_3"].astype(np.int64)
    tr = torch.from_numpy(tr_x).to(device)
    if augment:
        tr = torch.nn.functional.pad(tr, (pad, pad, pad, pad))                      # nn.ZeroPad2d(4), utils.py:165-167
    return tr, torch.from_numpy(tr_y).to(device), torch.from_numpy(te_x).to(device), torch.from_numpy(te_y).to(device)


class EpochBatches:
    """Batch schedule of noisynet.py:1232-1269: one numpy permutation per epoch; per batch a crop offset in [0, 8]^2 and a
    coin flip from Python's ``random`` (the generators the script uses, so --seed keeps its meaning).  Yields
    (idx [B] int64 device tensor, aug int32[3] device tensor {off_y, off_x, flip}, labels [B])."""

    def __init__(self, n, batch, labels, device, augment=True, pad=4):
        self.n, self.batch, self.labels, self.device, self.augment, self.pad = n, batch, labels, device, augment, pad

    def epoch(self):
        perm = torch.from_numpy(np.random.permutation(self.n)).to(self.device)      # :1232
        for i in range(self.n // self.batch):
            idx = perm[i * self.batch:(i + 1) * self.batch]
            if self.augment:
                k, j = random.randint(0, 2 * self.pad), random.randint(0, 2 * self.pad)   # :1265-1266
                flip = 1 if random.random() < 0.5 else 0                                   # :1268
            else:
                k = j = flip = 0
            aug = torch.tensor([k, j, flip], dtype=torch.int32, device=self.device)
            yield idx, aug, self.labels[idx]


def reference_batch(train_inputs, idx, aug, augment=True):
    """The batch exactly as the script's torch ops would build it (slice, crop, flip) -- the torch restatement the gather kernel is tested against."""
    x = train_inputs[idx]
    k, j, flip = (int(v) for v in aug.tolist())
    if augment:
        x = x[:, :, k:k + 32, j:j + 32]
        if flip:
            x = torch.flip(x, [3])
    return x.contiguous()
