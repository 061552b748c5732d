"""CPU, world_size 2, gloo: the data-parallel exchange (trainner_b200/parallel.py) averages the flat
gradient buffers across ranks and broadcasts parameters from rank 0."""
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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from trainner_b200.parallel import GradExchange, flat_buffers_of
    from trainner_b200.runtime import FlatGrads
    torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Conv2d(8, 4, 3), torch.nn.Linear(4, 2))

    class Eng:
        pass
    eng = Eng()
    eng.grads = FlatGrads(list(net[0].parameters()) + list(net[1].parameters()), "cpu")
    net._engine = [eng]
    ex = GradExchange()
    ex.broadcast_params([net])
    psum = float(sum(p.double().sum() for p in net.parameters()))
    eng.grads.attach()
    for p in eng.grads.params:
        p.grad.fill_(float(rank + 1))
    net[2].weight.grad = torch.full_like(net[2].weight, 10.0 * (rank + 1))
    net[2].bias.grad = torch.full_like(net[2].bias, 10.0 * (rank + 1))
    assert len(flat_buffers_of(net)) == 3
    ex.all_reduce_grads(net)
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in eng.grads.params)
    ok = ok and torch.allclose(net[2].weight.grad, torch.full_like(net[2].weight, 15.0))
    q.put((rank, psum, ok, ex.world))
    dist.destroy_process_group()


def test_grad_exchange_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1], "parameters differ after broadcast"
    assert all(r[2] for r in res) and all(r[3] == 2 for r in res)


def test_flatgrads_attach_semantics():
    from trainner_b200.runtime import FlatGrads
    net = torch.nn.Conv2d(3, 4, 3)
    fg = FlatGrads(list(net.parameters()), "cpu")
    fg.attach()
    assert net.weight.grad.data_ptr() == fg.view(net.weight).data_ptr()
    fg.flat.fill_(2.0)
    for p in net.parameters():
        p.grad = None                 # optimizer.zero_grad(set_to_none=True)
    fg.attach()
    assert float(fg.flat.abs().sum()) == 0.0 and net.bias.grad is not None
    net.bias.requires_grad = False    # frozen params are left alone
    net.bias.grad = None
    fg.attach()
    assert net.bias.grad is None
