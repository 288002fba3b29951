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

    def wait_early(self, k=0):
        """The current stream waits for early bucket ``k``'s sums (so that its parameters can be updated before the
        step's last exchange); all_reduce_sum_ still joins everything."""
        if self.world > 1 and k < len(self._works) and self._works[k] is not None:
            self._works[k].wait()

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


class SymmGradAllReduce(FlatGradAllReduce):
    """FlatGradAllReduce whose flat gradient buffer lives in SYMMETRIC memory (every rank maps every rank's buffer and
    the NVSwitch multicast alias of all of them, torch.distributed._symmetric_memory) and whose exchange is this
    library's own kernel instead of an NCCL call: nn_allreduce_start / nn_allreduce_wait (csrc/nn_collective.cu) --
    rank r reduces slice r of a bucket with multimem.ld_reduce (in-switch reduction) and multicasts the sums back with
    multimem.st, in place; epoch flags in the same buffer order the ranks.  Same interface as the NCCL class:
    ``start_early(k)`` launches bucket k's exchange on the current stream (the engine calls it on its side stream, under
    the conv backward), ``all_reduce_sum_()`` launches the remaining range and waits for every bucket on the current
    stream.  Capturable in the step's CUDA graph; no host synchronisation."""

    def __init__(self, model, world=None, early=None, ctas=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self._C, self._lib_mod = C, _lib
        params = [p for p in model.parameters() if p.requires_grad]
        buckets = early or []
        if buckets and not isinstance(buckets[0], (list, tuple)):
            buckets = [buckets]
        buckets = [[p for p in b if p.requires_grad] for b in buckets]
        ids = {id(p) for b in buckets for p in b}
        buckets = buckets + [[p for p in params if id(p) not in ids]]           # the last range: everything else
        self.world = world if world is not None else dist.get_world_size()
        self.rank = dist.get_rank()
        ref = params[0]
        dev = ref.device
        self.dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        lib = _lib.load()
        self.ctl = int(lib.nn_allreduce_ctl_bytes())
        q = 4 * self.world                                                       # floats: ranges are multiples of 16 B * world
        self.ranges, off = [], 0
        for b in buckets:
            n = sum(p.numel() for p in b)
            n_pad = (n + q - 1) // q * q
            self.ranges.append((off, n_pad))
            off += n_pad
        total = off
        group_name = dist.group.WORLD.group_name
        try:
            symm.enable_symm_mem_for_group(group_name)
        except Exception:       # noqa: BLE001  (newer torch enables it implicitly)
            pass
        self.buf = symm.empty(self.ctl + total * 4, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        torch.cuda.synchronize(dev)
        self.hdl = symm.rendezvous(self.buf, group_name)
        dist.barrier()                                                           # every rank's control words are zero
        self.flat = self.buf[self.ctl:].view(torch.float32)
        self.params = [p for b in buckets for p in b]
        for (o, _), b in zip(self.ranges, buckets):
            for p in b:
                p.grad = self.flat[o:o + p.numel()].view_as(p)
                o += p.numel()
        self.nbytes = total * 4
        self.n_buckets = len(self.ranges)
        self._started = [0] * self.n_buckets      # 0 idle, 1 exchange enqueued, 2 sums already waited for on some stream
        ptrs = list(self.hdl.buffer_ptrs)
        self._peer = (C.c_void_p * self.world)(*ptrs)
        self._mc = 0 if os.environ.get("NN_DP_MULTICAST", "1") == "0" else int(self.hdl.multicast_ptr or 0)
        self._local = ptrs[self.rank]
        assert self._local == self.buf.data_ptr()
        self.ctas = int(ctas) if ctas else 32
        self.comm = None                 # the exchange's own stream (created on first use, joins CUDA-graph captures)
        self.n_early = sum(n for _, n in self.ranges[:-1])
        self.bounds = [0]
        for _, n in self.ranges[:-1]:
            self.bounds.append(self.bounds[-1] + n)
        self._works = [None] * (self.n_buckets - 1)

    def _st(self):
        return torch.cuda.current_stream(self.dev_index).cuda_stream

    def _start(self, k):
        """Bucket k's exchange on the reducer's OWN stream, ordered after everything enqueued so far on the current stream:
        the kernel waits in-stream for the slowest rank, which must not hold up whatever the caller enqueues next."""
        o, n = self.ranges[k]
        if n and not self._started[k]:
            lib = self._lib_mod.load()
            if self.comm is None:
                self.comm = torch.cuda.Stream(device=self.dev_index)
            self.comm.wait_stream(torch.cuda.current_stream(self.dev_index))
            self._lib_mod.check(lib.nn_allreduce_start(self._peer, self._mc or None, self.rank, self.world, k, self.ctl + o * 4, n,
                                                       self.ctas, self.dev_index, self.comm.cuda_stream), "nn_allreduce_start")
            self._started[k] = 1

    def start_early(self, k=0):
        if self.world > 1 and k < self.n_buckets - 1:
            self._start(k)

    def wait_early(self, k=0):
        if self.world > 1 and self._started[k] == 1:
            cur = torch.cuda.current_stream(self.dev_index)
            cur.wait_stream(self.comm)
            self._lib_mod.check(self._lib_mod.load().nn_allreduce_wait(self._local, self.world, k, self.dev_index, cur.cuda_stream),
                                "nn_allreduce_wait")
            self._started[k] = 2

    def all_reduce_sum_(self):
        if self.world > 1:
            lib = self._lib_mod.load()
            for k in range(self.n_buckets):
                self._start(k)                   # whatever has not been started yet, the tail range included
            torch.cuda.current_stream(self.dev_index).wait_stream(self.comm)
            for k in range(self.n_buckets):
                if self._started[k] == 1:
                    self._lib_mod.check(lib.nn_allreduce_wait(self._local, self.world, k, self.dev_index, self._st()), "nn_allreduce_wait")
                self._started[k] = 0
        return self.flat


def make_grad_reducer(model, world, early=None):
    """SymmGradAllReduce (this library's NVSwitch kernel) on NCCL process groups with CUDA tensors; the NCCL / gloo
    all-reduce class otherwise (CPU tests) or when NN_DP_SYMM=0."""
    if world > 1 and dist.is_initialized() and dist.get_backend() == "nccl" and os.environ.get("NN_DP_SYMM", "1") != "0":
        try:
            return SymmGradAllReduce(model, world, early=early)
        except Exception as e:      # noqa: BLE001  (no symmetric memory on this system: peer access / driver)
            import sys
            sys.stderr.write("[dp] symmetric-memory all-reduce unavailable (%s); using NCCL\n" % (e,))
    return FlatGradAllReduce(model, world, early=early)


def rank_seed(seed, rank):
    """Independent shuffles / crops / Philox noise per rank (train_efficientnet.py:240: seed + rank)."""
    return int(seed) + int(rank)


def shard_batch(global_batch, rank, world):
    """Contiguous equal shards of the global batch (DistributedSampler semantics without padding)."""
    if global_batch % world:
        raise ValueError("global batch %d not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per
