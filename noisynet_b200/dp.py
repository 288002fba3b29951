"""Data parallelism for the hot path: one process per GPU, batch sharded across ranks, ONE exchange
per step = mean all-reduce of the parameter gradients (the reference's own
wiring is main.py:748-802 / train_efficientnet.py:225-320 with apex ``delay_allreduce=True``, i.e.
a single flat all-reduce after backward).  The model object stays a plain module (no DDP wrapper),
so script-style attribute access (model.conv1.weight.data.clamp_, noisynet.py:1532) keeps working.

All gradients live in ONE flat fp32 buffer (``p.grad`` are views into it): the exchange is a single
NCCL all-reduce of 5.5 MB for NoisyNet -- latency-bound over NVLink 5 / NVSwitch (NVLS), issued on the
compute stream so it is capturable in the step's CUDA graph.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


class FlatGradAllReduce:
    """Owns the flat gradient buffer of ``model`` and performs the per-step mean all-reduce."""

    def __init__(self, model, world=None, early=None):
        """``early``: parameters whose gradients are final early in the backward -- a list of parameters (one bucket)
        or a list of such lists (several buckets, in the order they become final; e.g. the fully connected layers of
        NoisyNet, 4.7 of the 5.5 MB, then conv2).  They are laid out first in the flat buffer, bucket by bucket, so
        that each bucket's all-reduce can be started (``start_early(k)``) while the remaining backward still runs."""
        params = [p for p in model.parameters() if p.requires_grad]
        buckets = early or []
        if buckets and not isinstance(buckets[0], (list, tuple)):
            buckets = [buckets]
        buckets = [[p for p in b if p.requires_grad] for b in buckets]
        flat_early = [p for b in buckets for p in b]
        ids = {id(p) for p in flat_early}
        self.params = flat_early + [p for p in params if id(p) not in ids]
        self.bounds = [0]
        for b in buckets:
            self.bounds.append(self.bounds[-1] + sum(p.numel() for p in b))
        self.n_early = self.bounds[-1]
        self._works = [None] * len(buckets)
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.nbytes = n * self.flat.element_size()

    @property
    def _work(self):            # the single-bucket interface of round 1 (tests)
        return next((w for w in self._works if w is not None), None)

    def zero_(self):
        self.flat.zero_()

    def broadcast_parameters(self, model, src=0):
        """Rank 0's parameters and buffers to everyone (as DDP does at construction, main.py:798-802)."""
        if self.world > 1:
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src)

    def start_early(self, k=0):
        """Asynchronous SUM all-reduce of early bucket ``k`` on the process group's stream (it waits for the CURRENT
        stream at the call); overlaps with the kernels enqueued afterwards (all_reduce_sum_ joins)."""
        if self.world > 1 and k < len(self._works) and self._works[k] is None and self.bounds[k + 1] > self.bounds[k]:
            self._works[k] = dist.all_reduce(self.flat[self.bounds[k]:self.bounds[k + 1]], op=dist.ReduceOp.SUM, async_op=True)

    def all_reduce_sum_(self):
        if self.world > 1:
            # buckets never started are folded into the final all-reduce together with the remaining parameters
            lo = 0
            for k, w in enumerate(self._works):
                if w is None:
                    break
                lo = self.bounds[k + 1]
            started = [w for w in self._works if w is not None]
            if any(w is None for w in self._works[:len(started)]) or \
               any(w is not None for w in self._works[len(started):]):
                raise RuntimeError("early buckets must be started in order")
            rest = self.flat[lo:]
            if rest.numel():
                dist.all_reduce(rest, op=dist.ReduceOp.SUM)
            for w in started:
                w.wait()
            self._works = [None] * len(self._works)
        return self.flat

    def all_reduce_mean_(self):
        if self.world > 1:
            self.all_reduce_sum_()
            self.flat.mul_(1.0 / self.world)
        return self.flat


def rank_seed(seed, rank):
    """Independent shuffles / crops / Philox noise per rank (train_efficientnet.py:240: seed + rank)."""
    return int(seed) + int(rank)


def shard_batch(global_batch, rank, world):
    """Contiguous equal shards of the global batch (DistributedSampler semantics without padding)."""
    if global_batch % world:
        raise ValueError("global batch %d not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per
