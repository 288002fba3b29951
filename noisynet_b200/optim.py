"""Fused AdamW + weight clamp (+ max|W|) for the NoisyNet training step.

One kernel launch (`nn_adamw_step`) replaces `torch.optim.AdamW.step()` (noisynet.py:1163, :1520) for ALL
parameter tensors and the per-layer `weight.data.clamp_(-w_max, w_max)` that follows it
(noisynet.py:1527-1542); it also leaves max|W| of every clamped tensor in a device scalar for the next
forward's merged-DAC noise (hardware_model.py:47).  Same update rule as torch.optim.AdamW (decoupled weight
decay, bias-corrected first/second moments, eps outside the sqrt, no amsgrad); the step count lives on the
device so a captured CUDA graph can be replayed.
"""
import ctypes as C

import torch

from . import _lib
from .ops import _dev, _stream


class FusedAdamW:
    def __init__(self, param_groups, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.param_groups = []
        for g in param_groups:
            g = dict(g)
            g["params"] = list(g["params"])
            for k, v in self.defaults.items():
                g.setdefault(k, v)
            g.setdefault("clamp", 0.0)          # w_max of noisynet.py:1527-1542 (0 = no clamp)
            self.param_groups.append(g)
        self.grad_scale = float(grad_scale)
        self.state = {}
        self._table = None
        self._table_key = None
        self.step_dev = None
        self.absmax = None                      # [num_tensors] device floats: max|p| after the step

    def _params(self):
        for g in self.param_groups:
            for p in g["params"]:
                if p.requires_grad:
                    yield g, p

    def _build(self):
        ps = list(self._params())
        if len(ps) > 24:
            raise ValueError("FusedAdamW: more than 24 parameter tensors")
        dev = ps[0][1].device
        if self.step_dev is None:
            self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
            self.absmax = torch.zeros(len(ps), dtype=torch.float32, device=dev)
        arr = (_lib.AdamWTensor * len(ps))()
        self.index = {}
        for i, (g, p) in enumerate(ps):
            if p.grad is None:
                raise RuntimeError("FusedAdamW.step(): a parameter has no gradient")
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                raise _lib.NoisyNetLibraryError("FusedAdamW: parameters must be contiguous CUDA float32 tensors")
            st = self.state.get(p)
            if st is None:
                st = self.state[p] = dict(exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))
            t = arr[i]
            t.p, t.g, t.m, t.v = p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            t.n = p.numel()
            t.lr, t.weight_decay, t.clamp = float(g["lr"]), float(g["weight_decay"]), float(g["clamp"])
            t.absmax_out = self.absmax[i:i + 1].data_ptr()
            self.index[p] = i
        self._table = arr
        self._n = len(ps)

    def _key(self):
        return tuple((id(p), p.data_ptr(), p.grad.data_ptr() if p.grad is not None else 0, g["lr"], g["weight_decay"],
                      g["clamp"]) for g, p in self._params())

    def zero_grad(self, set_to_none=False):
        for _, p in self._params():
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self):
        key = self._key()
        if self._table is None or key != self._table_key:
            self._build()
            self._table_key = key
            self._parts = {}
        b1, b2 = self.param_groups[0]["betas"]
        dev = _dev(self.step_dev)
        _lib.check(_lib.load().nn_adamw_step(self._table, self._n, float(b1), float(b2),
                                             float(self.param_groups[0]["eps"]), self.grad_scale,
                                             self.step_dev.data_ptr(), dev, _stream(dev)), "nn_adamw_step")

    @torch.no_grad()
    def step_part(self, params, advance):
        """The update of a subset of the step's tensors (``nn_adamw_step_part``): the engine updates the layers whose
        gradients are final (and exchanged) early, under the rest of the backward pass, and passes ``advance=True`` with
        the last subset only.  The subsets of one step must cover every parameter exactly once."""
        key = self._key()
        if self._table is None or key != self._table_key:
            self._build()
            self._table_key = key
            self._parts = {}
        ids = tuple(id(p) for p in params)
        part = self._parts.get(ids) if hasattr(self, "_parts") else None
        if part is None:
            if not hasattr(self, "_parts"):
                self._parts = {}
            arr = (_lib.AdamWTensor * len(params))()
            for j, p in enumerate(params):
                C.memmove(C.byref(arr[j]), C.byref(self._table[self.index[p]]), C.sizeof(_lib.AdamWTensor))
            part = self._parts[ids] = arr
        b1, b2 = self.param_groups[0]["betas"]
        dev = _dev(self.step_dev)
        _lib.check(_lib.load().nn_adamw_step_part(part, len(params), float(b1), float(b2), float(self.param_groups[0]["eps"]),
                                                  self.grad_scale, self.step_dev.data_ptr(), 1 if advance else 0, dev, _stream(dev)),
                   "nn_adamw_step_part")

    def absmax_of(self, p):
        i = self.index[p]
        return self.absmax[i:i + 1]
