"""Data-parallel gradient exchange on CPU (gloo, world_size 2): the flat-buffer all-reduce of
engine.GradAllReduce averages every parameter gradient across ranks and leaves parameters,
shapes and non-gradient state alone."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from titanet_b200 import losses, models
    from titanet_b200.engine import GradAllReduce
    torch.manual_seed(0)
    net = models.TitaNet.get_titanet(n_mega_blocks=1, model_size="s", loss_function=losses.CELoss(192, 11), dropout=0.0)
    params = list(net.parameters())
    g = torch.Generator().manual_seed(100 + rank)
    for p in params:                                    # stand-in for a per-shard backward
        p.grad = torch.randn(p.shape, generator=g)
    mine = [p.grad.clone() for p in params]
    before = [p.detach().clone() for p in params]
    GradAllReduce(params, world)()
    gathered = [None] * world
    dist.all_gather_object(gathered, [m.numpy() for m in mine])
    ok = True
    for i, p in enumerate(params):
        mean = sum(torch.from_numpy(gathered[r][i]) for r in range(world)) / world
        ok &= torch.allclose(p.grad, mean, atol=1e-6) and p.grad.shape == p.shape
        ok &= torch.equal(p.detach(), before[i])
    out.put((rank, bool(ok), sum(p.numel() for p in params)))
    dist.destroy_process_group()


def _worker_arena(rank, world, port, out):
    """Gradients that already live in one ZeroArena buffer are exchanged in place (no pack / unpack)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from titanet_b200 import _ops as ops
    from titanet_b200.engine import GradAllReduce
    g = torch.Generator().manual_seed(7 + rank)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((5, 3), (7,), (2, 4, 1), (1,))]
    arena = ops.ZeroArena("cpu")
    arena.measuring = False
    arena.buf = torch.zeros(4096, dtype=torch.uint8)       # what begin_step() allocates + clears on the device
    prev = ops.set_arena(arena)
    try:
        junk = ops.zeros((3,), params[0], torch.float64)    # a non-gradient accumulator in between
        for p in params:
            p.grad = ops.gempty(p.shape, p)
            p.grad.copy_(torch.randn(p.shape, generator=g))
        mine = [p.grad.clone() for p in params]
        ar = GradAllReduce(params, world)
        span = ar._arena_span([p.grad for p in params])
        ar()
    finally:
        ops.set_arena(prev)
    gathered = [None] * world
    dist.all_gather_object(gathered, [m.numpy() for m in mine])
    ok = span is not None and ar.flat is None and float(junk.abs().sum()) == 0.0
    for i, p in enumerate(params):
        mean = sum(torch.from_numpy(gathered[r][i]) for r in range(world)) / world
        ok &= torch.allclose(p.grad, mean, atol=1e-6)
        ok &= arena.buf.data_ptr() <= p.grad.data_ptr() < arena.buf.data_ptr() + arena.buf.numel()
    out.put((rank, bool(ok), 0))
    dist.destroy_process_group()


def test_arena_gradients_are_exchanged_in_place_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_arena, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2] > 1_000_000


def test_single_process_is_a_no_op():
    from titanet_b200.engine import GradAllReduce
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    GradAllReduce([p], 1)()
    assert torch.equal(p.grad, torch.full((3,), 2.0))
