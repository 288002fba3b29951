"""Drop-in replacement for the reference's ``quant`` module (quant.py), used by
models/mobilenet.py:9 and models/resnet.py:9.  Differences from ``hardware_model.QuantMeasure``
are the reference's own: ``pctl`` is a FRACTION (quant.py:109), the percentile branch is gated on
``self.training`` (quant.py:102), there is no signed-weight branch, and a ``debug_quant`` argument.
"""
import torch
import torch.nn as nn

from .ops import UniformQuantize  # noqa: F401  (quant.py:10)

__all__ = ["UniformQuantize", "QuantMeasure"]


class QuantMeasure(nn.Module):
    """quant.py:64-137."""

    def __init__(self, num_bits=8, momentum=0.0, stochastic=0.5, min_value=0, max_value=0, scale=1,
                 calculate_running=False, pctl=.999, debug=False, debug_quant=False, inplace=False):
        super(QuantMeasure, self).__init__()
        self.register_buffer('running_min', torch.zeros(1))
        self.register_buffer('running_max', torch.zeros([]))
        self.momentum = momentum
        self.num_bits = num_bits
        self.stochastic = stochastic
        self.inplace = inplace
        self.debug = debug
        self.debug_quant = debug_quant
        self.max_value = max_value
        self.min_value = min_value
        self.scale = scale
        self.calculate_running = calculate_running
        self.running_list = []
        self.pctl = pctl

    def _host_running_max(self):
        t = self.running_max
        hit = self.__dict__.get('_rm_cache')
        # keyed on the tensor OBJECT (strong reference; the scripts replace the buffer and ids are reused) + its version
        if hit is None or hit[0] is not t or hit[1] != t._version:
            hit = (t, t._version, float(t.item()))
            self.__dict__['_rm_cache'] = hit
        return hit[2]

    def forward(self, input):
        from . import ops
        range_dev = None
        max_value = self.max_value
        with torch.no_grad():
            if self.calculate_running and self.training:                      # quant.py:102-115
                if 224 in list(input.shape):
                    pctl = torch.tensor(0.92) if self.num_bits == 4 else torch.tensor(1.0)
                    mx = ops.tensor_stats(input.detach())[0:1]
                    range_dev = torch.cat([torch.full_like(mx, float(self.min_value)), mx])
                else:       # kthvalue + input.max().item() on the device: exact radix select, no host sync
                    out4 = ops.range_select(input.detach(), mode=0, min_value=float(self.min_value),
                                            k_rank=int(input.numel() * self.pctl))
                    pctl, range_dev = out4[0].reshape([]), out4[2:4]
                self.running_list.append(pctl)
            else:                                                             # quant.py:117-124
                if self.max_value > 0:
                    max_value = self.max_value
                elif self._host_running_max() > 0:
                    max_value = self._host_running_max()
                else:
                    mx = ops.tensor_stats(input.detach())[0:1]
                    range_dev = torch.cat([torch.full_like(mx, float(self.min_value)), mx])
            stoch = self.stochastic if self.training else 0                   # quant.py:132-135
        if range_dev is not None:
            return UniformQuantize.apply(input, self.num_bits, None, None, stoch, self.inplace, self.debug_quant, None, range_dev)
        out = UniformQuantize.apply(input, self.num_bits, float(self.min_value), float(max_value), stoch,
                                      self.inplace, self.debug_quant)
        out._nn_quant = (int(self.num_bits), float(self.min_value), float(max_value))
        return out
