"""CPU-only, world_size 2 over gloo: the N>1 plumbing (batch sharding, bucketed gradient all-reduce,
max-over-ranks timing) used by bench.py --gpus N and by a data-parallel training step."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from mvp_benchmark_b200 import dist as mdist
    r, w, _ = mdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    batch = torch.arange(7 * 3, dtype=torch.float32).view(7, 3)          # 7 clouds over 2 ranks -> 4 + 3
    mine = mdist.shard_batch(batch, r, w)
    model = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.Linear(5, 2))
    torch.manual_seed(0)
    for p in model.parameters():
        torch.nn.init.normal_(p)
    model(mine).sum().backward()
    calls = mdist.allreduce_gradients(list(model.parameters()), bucket_bytes=64, average=False)
    slowest = mdist.max_over_ranks(1.0 + rank, torch.device("cpu"))
    mdist.barrier()
    q.put((rank, mine.shape[0], calls, slowest, [p.grad.tolist() for p in model.parameters()]))
    torch.distributed.destroy_process_group()


def _worker_overlapped(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from mvp_benchmark_b200 import dist as mdist
    r, w, _ = mdist.init_from_env(backend="gloo")
    batch = torch.arange(8 * 3, dtype=torch.float32).view(8, 3) / 10
    mine = mdist.shard_batch(batch, r, w)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Linear(4, 2))
    unused = torch.nn.Linear(2, 2)                                   # never part of the graph: no gradient
    params = list(model.parameters()) + list(unused.parameters())
    reducer = mdist.OverlappedGradientAllReduce(params, bucket_bytes=64, average=True)
    out = []
    for step in range(2):                                             # state resets between steps
        for p in params:
            p.grad = None
        model(mine * (step + 1)).square().sum().backward()
        calls = reducer.finish()
        out.append((calls, [None if p.grad is None else p.grad.tolist() for p in params]))
    mdist.barrier()
    q.put((rank, out))
    torch.distributed.destroy_process_group()


def test_overlapped_gradient_all_reduce_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_overlapped, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = torch.arange(8 * 3, dtype=torch.float32).view(8, 3) / 10
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Linear(4, 2))
    for step in range(2):
        for p in model.parameters():
            p.grad = None
        model(batch * (step + 1)).square().sum().backward()
        (c0, g0), (c1, g1) = res[0][1][step], res[1][1][step]
        assert c0 == c1 and c0 >= 3                                  # same collectives on both ranks, several buckets
        assert g0[-2:] == [None, None] and g1[-2:] == [None, None]   # the unused layer stays without gradient
        for a, b, p in zip(g0, g1, model.parameters()):
            a, b = torch.tensor(a), torch.tensor(b)
            assert torch.allclose(a, b) and torch.allclose(a, p.grad / 2, rtol=1e-5, atol=1e-6)   # averaged over 2 ranks


def test_shard_bounds_cover_everything():
    from mvp_benchmark_b200.dist import shard_bounds
    for total in (1, 7, 32, 256):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharded_step_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [4, 3]                      # contiguous batch slices
    assert res[0][2] == res[1][2] and res[0][2] >= 2           # bucketed: several small collectives
    assert res[0][3] == res[1][3] == 2.0                       # max over ranks
    # summed gradients equal the single-process gradient on the whole batch
    batch = torch.arange(7 * 3, dtype=torch.float32).view(7, 3)
    model = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.Linear(5, 2))
    torch.manual_seed(0)
    for p in model.parameters():
        torch.nn.init.normal_(p)
    model(batch).sum().backward()
    for g0, g1, p in zip(res[0][4], res[1][4], model.parameters()):
        g0, g1 = torch.tensor(g0), torch.tensor(g1)
        assert torch.allclose(g0, g1) and torch.allclose(g0, p.grad, rtol=1e-5, atol=1e-5)


def _worker_asymmetric(rank, world, port, q):
    """Rank 0's graph uses the branch `extra` and rank 1's does not (a data-dependent branch): the ranks finish
    their buckets in different orders, and rank 1 has no gradient for `extra` at all."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from mvp_benchmark_b200 import dist as mdist
    r, w, _ = mdist.init_from_env(backend="gloo")
    torch.manual_seed(0)
    first, extra, last = torch.nn.Linear(3, 4), torch.nn.Linear(4, 4), torch.nn.Linear(4, 2)
    never = torch.nn.Linear(2, 2)
    params = [*first.parameters(), *extra.parameters(), *last.parameters(), *never.parameters()]
    x = torch.arange(6, dtype=torch.float32).view(2, 3) / 10 + rank
    out = {}
    for mode in ("overlapped", "after"):
        for p in params:
            p.grad = None
        reducer = mdist.OverlappedGradientAllReduce(params, bucket_bytes=16, average=True) if mode == "overlapped" else None
        h = first(x)
        if rank == 0:
            h = extra(h)
        last(h).square().sum().backward()
        if reducer is not None:
            reducer.finish()
            reducer.remove()
        else:
            mdist.allreduce_gradients(params, bucket_bytes=16, average=True)
        out[mode] = [None if p.grad is None else p.grad.tolist() for p in params]
    mdist.barrier()
    q.put((rank, out))
    torch.distributed.destroy_process_group()


def test_rank_asymmetric_unused_parameter():
    """collectives must be issued in the same order on every rank, and a parameter only ONE rank used
    must end up with the same averaged gradient on all ranks (else the replicas' weights diverge)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_asymmetric, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process expectation: mean of the two ranks' gradients, zeros standing in for rank 1's missing branch
    torch.manual_seed(0)
    first, extra, last = torch.nn.Linear(3, 4), torch.nn.Linear(4, 4), torch.nn.Linear(4, 2)
    mods = [first, extra, last]
    want = None
    for rank in range(2):
        for mod in mods:
            mod.zero_grad(set_to_none=True)
        x = torch.arange(6, dtype=torch.float32).view(2, 3) / 10 + rank
        h = first(x)
        if rank == 0:
            h = extra(h)
        last(h).square().sum().backward()
        g = [torch.zeros_like(p) if p.grad is None else p.grad.clone() for mod in mods for p in mod.parameters()]
        want = g if want is None else [a + b for a, b in zip(want, g)]
    want = [w / 2 for w in want]
    for mode in ("overlapped", "after"):
        g0, g1 = res[0][1][mode], res[1][1][mode]
        assert g0[-2:] == [None, None] and g1[-2:] == [None, None]   # used by no rank: stays None
        for a, b, w in zip(g0[:-2], g1[:-2], want):
            assert a is not None and b is not None
            a, b = torch.tensor(a), torch.tensor(b)
            assert torch.equal(a, b), mode                           # identical on both ranks
            assert torch.allclose(a, w, rtol=1e-5, atol=1e-6), mode
