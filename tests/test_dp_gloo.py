"""N > 1 host logic on CPU: world_size-2 gloo processes exercise noisynet_b200.dp (flat-gradient mean
all-reduce, parameter broadcast, batch sharding, per-rank seeds).  No CUDA kernels are called."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, two_buckets=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from noisynet_b200 import dp
    r, w, _ = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(dp.rank_seed(0, rank))                       # different init per rank on purpose
    model = torch.nn.Sequential(torch.nn.Linear(12, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    if two_buckets:                                                 # buckets in the order they become final
        red = dp.FlatGradAllReduce(model, world, early=[[model[2].weight], [model[2].bias, model[0].weight]])
        assert red.params[:3] == [model[2].weight, model[2].bias, model[0].weight] and red.bounds == [0, 21, 21 + 3 + 84]
    else:
        red = dp.FlatGradAllReduce(model, world, early=[model[2].weight])      # early bucket: last layer's weight
        assert red.params[0] is model[2].weight and red.n_early == 21
    red.broadcast_parameters(model)                                 # now identical to rank 0
    g = torch.Generator().manual_seed(123)
    X, Y = torch.randn(16, 12, generator=g), torch.randint(0, 3, (16,), generator=g)
    lo, hi = dp.shard_batch(16, rank, world)
    red.zero_()
    loss = torch.nn.functional.cross_entropy(model(X[lo:hi]), Y[lo:hi])
    loss.backward()
    assert all(p.grad.data_ptr() >= red.flat.data_ptr() for p in model.parameters())   # grads are views
    red.start_early()                                            # async bucket, joined by all_reduce_mean_
    if two_buckets:
        red.start_early(1)
        # the engine updates the early buckets' parameters before the step's last exchange (optim.FusedAdamW.step_part):
        # their sums must be complete after wait_early
        red.wait_early(0)
        red.wait_early(1)
        early_now = red.flat[:red.n_early].clone()
    red.all_reduce_mean_()
    if two_buckets:
        assert torch.allclose(red.flat[:red.n_early], early_now / world)      # all_reduce_mean_ only rescaled them
    assert red._work is None
    torch.save({"flat": red.flat.clone(), "params": [p.detach().clone() for p in model.parameters()]},
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("two_buckets", [False, True])
def test_flat_grad_allreduce_world2(tmp_path, two_buckets):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), two_buckets), nprocs=world, join=True)
    a = torch.load(tmp_path / "rank0.pt")
    b = torch.load(tmp_path / "rank1.pt")
    assert torch.equal(a["flat"], b["flat"])
    for pa, pb in zip(a["params"], b["params"]):
        assert torch.equal(pa, pb)
    # mean of the two shard gradients == gradient of the full batch (equal shard sizes, mean loss)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(12, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    with torch.no_grad():
        for p, q in zip(model.parameters(), a["params"]):
            p.copy_(q)
    g = torch.Generator().manual_seed(123)
    X, Y = torch.randn(16, 12, generator=g), torch.randint(0, 3, (16,), generator=g)
    torch.nn.functional.cross_entropy(model(X), Y).backward()
    first = [model[2].weight, model[2].bias, model[0].weight] if two_buckets else [model[2].weight]
    order = first + [p for p in model.parameters() if all(p is not q for q in first)]          # early buckets first
    full = torch.cat([p.grad.flatten() for p in order])
    assert torch.allclose(a["flat"], full, atol=1e-6)


def test_shard_batch_and_seed():
    from noisynet_b200 import dp
    assert dp.shard_batch(512, 3, 8) == (192, 256)
    assert dp.rank_seed(5, 3) == 8
    import pytest
    with pytest.raises(ValueError):
        dp.shard_batch(10, 0, 4)
