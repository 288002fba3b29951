#!/usr/bin/env python
"""Multi-GPU check of the engine's data-parallel step, run under torchrun on one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dp_engine_check.py

Trains a few steps (fixed seeds, fixed data) twice from the same initial state: with this library's symmetric-memory exchange
(buckets started early, fc2/fc1/conv2 updated early on the side stream) and with the NCCL reducer.  Checks: every rank ends
with bit-identical parameters (replicas stay in sync), and -- on 2 ranks, where a + b has one summation order -- the two
exchange implementations give bit-identical parameters and losses."""
import copy
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from noisynet_b200 import dp  # noqa: E402
from noisynet_b200.engine import NoisyNetEngine  # noqa: E402
from noisynet_b200.net import NoisyNet, default_args, init_like_reference, make_fused_optimizer, with_quant  # noqa: E402


def digest(model):
    h = hashlib.sha256()
    for p in model.parameters():
        h.update(p.detach().cpu().numpy().tobytes())
    return h.hexdigest()[:16]


def run(kind, state, rank, world, dev, steps=6, B=128):
    a = default_args()
    with_quant(a, 4, 4)
    torch.manual_seed(5)
    m = init_like_reference(NoisyNet(a, fused=True, precision="bf16")).to(dev)
    m.load_state_dict(state)
    m.quantize2.running_max = torch.tensor(5.0, device=dev)
    m.quantize4.running_max = torch.tensor(5.0, device=dev)
    m.collect_stats = False
    m.train()
    opt = make_fused_optimizer(m, a, grad_scale=1.0 / world)
    early = [[m.linear1.weight, m.linear2.weight], [m.conv2.weight]]
    red = dp.SymmGradAllReduce(m, world, early=early) if kind == "symm" else dp.FlatGradAllReduce(m, world, early=early)
    eng = NoisyNetEngine(m, B, opt=opt, reducer=red)
    torch.manual_seed(dp.rank_seed(77, rank))            # per-rank Philox streams, the same in both runs
    gen = torch.Generator().manual_seed(900 + rank)
    losses = []
    for s in range(steps):
        x = (torch.randint(0, 16, (B, 3, 32, 32), generator=gen).float() / 15).to(dev)
        y = torch.randint(0, 10, (B,), generator=gen).to(dev)
        losses.append(float(eng.train_step(x, y)[0]))
    torch.cuda.synchronize()
    return digest(m), losses


def main():
    rank, world, local = dp.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    a = default_args()
    with_quant(a, 4, 4)
    torch.manual_seed(5)
    m0 = init_like_reference(NoisyNet(a, fused=True, precision="bf16")).to(dev)
    state = copy.deepcopy(m0.state_dict())
    for t in state.values():
        dist.broadcast(t, 0)
    d_symm, l_symm = run("symm", state, rank, world, dev)
    d_nccl, l_nccl = run("nccl", state, rank, world, dev)
    all_symm = [None] * world
    dist.all_gather_object(all_symm, d_symm)
    ok = len(set(all_symm)) == 1
    if world == 2:
        ok = ok and d_symm == d_nccl and l_symm == l_nccl
    if rank == 0:
        print("replicas in sync:", len(set(all_symm)) == 1, all_symm[0], "| symm == nccl:", d_symm == d_nccl,
              "| losses", ["%.5f" % v for v in l_symm], ["%.5f" % v for v in l_nccl], flush=True)
        print("RESULT", "ok" if ok else "FAILED", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
