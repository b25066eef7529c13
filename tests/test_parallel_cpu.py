"""CPU (gloo, world_size 2): the data-parallel host logic — batch sharding, SUM all-reduce of flat gradients, 1/R
averaging and the xR correction of the sum-type TV term — reproduces single-process big-batch gradients of the oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dsgan_oracle as O
from dsgan_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat(d):
    return torch.cat([v.flatten() for v in d.values()])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    PG, PD, PV = O.init_params_G(20), O.init_params_D(20), O.init_params_vgg(20)
    A, B = O.synthetic_pair(4, 32, 32, seed=1)
    batch = parallel.shard_batch({"A": A, "B": B, "A_paths": ["a"] * 4, "B_paths": ["b"] * 4, "note": "x"})
    assert batch["A"].shape[0] == 2 and len(batch["A_paths"]) == 2 and batch["note"] == "x"
    out = O.train_step(PG, PD, PV, batch["A"], batch["B"], update=False, tv_scale=parallel.tv_grad_scale())
    gD = parallel.allreduce_grads(_flat(out["grads_D"])) * parallel.adam_grad_scale()
    gG = parallel.allreduce_grads(_flat(out["grads_G"])) * parallel.adam_grad_scale()
    wrong = O.train_step(PG, PD, PV, batch["A"], batch["B"], update=False)          # without the TV correction
    gG_wrong = parallel.allreduce_grads(_flat(wrong["grads_G"])) * parallel.adam_grad_scale()
    p = parallel.broadcast_params(torch.full((3,), float(rank)))
    if rank == 0:
        q.put(tuple(t.numpy() for t in (gD, gG, gG_wrong, p)))  # by value: shared-memory handles die with the child
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gradients_equal_big_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gD, gG, gG_wrong, bc = (torch.from_numpy(a) for a in q.get(timeout=500))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    PG, PD, PV = O.init_params_G(20), O.init_params_D(20), O.init_params_vgg(20)
    A, B = O.synthetic_pair(4, 32, 32, seed=1)
    ref = O.train_step(PG, PD, PV, A, B, update=False)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    assert rel(gD, _flat(ref["grads_D"])) < 1e-4
    assert rel(gG, _flat(ref["grads_G"])) < 2e-3
    assert rel(gG_wrong, _flat(ref["grads_G"])) > 5 * rel(gG, _flat(ref["grads_G"]))  # the TV term really needs xR
    assert bc.tolist() == [0.0, 0.0, 0.0]


def test_shard_batch_rejects_ragged():
    A = torch.zeros(3, 3, 16, 16)
    with pytest.raises(ValueError):
        parallel.shard_batch({"A": A, "B": A}, 0, 2)
    assert parallel.shard_batch({"A": A, "B": A}, 0, 1)["A"].shape[0] == 3
    assert parallel.tv_grad_scale(8) == 8.0 and parallel.adam_grad_scale(8) == 0.125
