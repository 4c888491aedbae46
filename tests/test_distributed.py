"""World-size-2 gloo test (CPU) of the host-side logic of data-parallel map training: the packed
all-reduce that keeps the replicated map / decoder state identical on every rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pin_slam_b200.utils.mapper import allreduce_gradients, allreduce_map_statistics

    g = torch.Generator().manual_seed(100 + rank)
    m_rows, f, n_dec = 50, 8, 30
    red = torch.randn(m_rows * f + n_dec, generator=g) / world   # gradients pre-scaled by 1/world
    cert_before = torch.arange(m_rows, dtype=torch.float32)
    delta = torch.rand(m_rows, generator=g)
    cert = cert_before + delta
    ts = torch.randint(0, 100, (m_rows,), generator=g, dtype=torch.int32)
    grads_local = red[: m_rows * f + n_dec].clone()
    allreduce_gradients(red)
    allreduce_map_statistics(cert_before, cert, ts)
    # plain lists: torch tensors in an mp.Queue are passed through shared-memory handles that can
    # outlive the producer badly
    q.put((rank, grads_local.tolist(), delta.tolist(), red[: m_rows * f + n_dec].tolist(), cert.tolist(), ts.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_packed_allreduce_keeps_ranks_identical():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, g0, d0, r0, c0, t0), (_, g1, d1, r1, c1, t1) = [(r[0],) + tuple(torch.tensor(x) for x in r[1:]) for r in res]
    assert torch.equal(r0, r1) and torch.equal(c0, c1) and torch.equal(t0, t1)
    torch.testing.assert_close(r0, g0 + g1)
    torch.testing.assert_close(c0, torch.arange(50, dtype=torch.float32) + d0 + d1)
    assert int(t0.max()) <= 99
