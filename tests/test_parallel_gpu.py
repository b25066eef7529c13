"""Data parallelism on the REAL kernels: two ranks (two processes sharing cuda:0, gloo all-reduce staged through the
host -- no rank ever spins on the GPU waiting for the other) run optimize_parameters() on the two halves of a batch; the
all-reduced flat gradients, averaged, must equal the single-process full-batch gradients (TV xR correction, all-reduce
placement between the CUDA-graph segments), eagerly and under graph replay."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

import dsgan_oracle as O  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(prec, graph):
    from dsgan_b200.models import create_model
    from dsgan_b200.options.train_options import TrainOptions
    opt = TrainOptions().parse("/tmp/none", "/tmp/dsgan_b200_dp", argv=["--precision", prec, "--cuda_graph", str(graph)],
                               quiet=True)
    m = create_model(opt)
    m.setup(opt)
    PG, PD, PV = O.init_params_G(20), O.init_params_D(20), O.init_params_vgg(20)
    m.netG.load_state_dict(PG)
    m.netD.load_state_dict(PD)
    m.vgg.load_state_dict(PV, strict=False)
    return m


def _run(m, batches, world=1, rank=0):
    from dsgan_b200 import parallel
    rows = []
    for A, B in batches:
        b = parallel.shard_batch({"A": A, "B": B, "A_paths": [""] * A.shape[0], "B_paths": [""] * A.shape[0]}, rank, world)
        m.set_input(b)
        m.optimize_parameters()
        torch.cuda.synchronize()
        rows.append([float(m._loss[i]) for i in range(7)])
    gD = m.netD.flat_buffers()[1].detach().cpu().clone()
    gG = m.netG.flat_buffers()[1].detach().cpu().clone()
    return rows, gD, gG


def _worker(rank, world, port, prec, graph, nsteps, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _make(prec, graph)
    assert m.world == world
    batches = [O.synthetic_pair(4, 64, 64, seed=50 + i) for i in range(nsteps)]
    rows, gD, gG = _run(m, batches, world, rank)
    if graph:
        assert m._gs is not None and m._gs["plan"] is not None and any(k == "eager" for k, _ in m._gs["plan"])
    # losses are per-rank means over the local half: average them for the comparison with the full batch (TV is a SUM)
    t = torch.tensor(rows, dtype=torch.float64)
    dist.all_reduce(t)
    if rank == 0:
        q.put((t.numpy(), gD.numpy(), gG.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _spawn(prec, graph, nsteps):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, prec, graph, nsteps, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows, gD, gG = q.get(timeout=500)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return torch.from_numpy(rows), torch.from_numpy(gD), torch.from_numpy(gG)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("prec,graph,nsteps", [("fp32", 0, 1), ("bf16", 0, 1), ("fp32", 1, 4)])
def test_two_ranks_equal_the_full_batch(prec, graph, nsteps):
    rows2, gD2, gG2 = _spawn(prec, graph, nsteps)
    m = _make(prec, graph)
    batches = [O.synthetic_pair(4, 64, 64, seed=50 + i) for i in range(nsteps)]
    rows1, gD1, gG1 = _run(m, batches)
    # flat gradient buffers hold the SUM over ranks; the fused Adam applies 1/R
    eD, eG = _rel(gD2 / 2, gD1), _rel(gG2 / 2, gG1)
    print("2-rank vs full batch (%s, graph=%d): D %.2e  G %.2e" % (prec, graph, eD, eG))
    if nsteps == 1:
        # fp32: only the fp32-atomic summation order differs (measured D 2.4e-3 -- D's gradient is the ill-conditioned
        # real/fake difference -- and G 3.1e-4); a missing TV xR correction or 1/R would show up at >= 1e-2 in G.
        # bf16: per-sample roundings differ when the batch is split differently.
        tolD, tolG = (5e-3, 1e-3) if prec == "fp32" else (8e-2, 4e-2)
        assert eD < tolD and eG < tolG
    # losses: batch means -> mean of the two ranks; TV (slot 3) is a batch SUM -> sum of the two ranks
    r2 = rows2.clone()
    r2 /= 2
    r2[:, 3] *= 2
    # slot 4 holds 1 - ssim with ssim a batch mean: mean of ranks as well
    r1 = torch.tensor(rows1, dtype=torch.float64)
    ltol = (2e-4 if prec == "fp32" else 5e-3) * (1 if nsteps == 1 else 20)   # later steps: trajectories drift apart slowly
    assert float((r2 - r1).abs().max()) < ltol * max(1.0, float(r1.abs().max())), (r2, r1)
