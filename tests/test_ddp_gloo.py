"""Host-side logic of the data-parallel path on CPU: world_size 2, gloo backend (no GPU needed).
The compute path itself never runs on CPU; these tests cover bucketing, averaging and broadcast."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from midi_emotion_b200.ddp import DataParallel, _buckets, allreduce_mean_, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                      # different init on every rank
        model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
        ddp = DataParallel(model, bucket_mb=0.0001)        # tiny buckets: several allreduces
        w0 = [p.detach().clone() for p in model.parameters()]
        gathered = [torch.zeros_like(w0[0]) for _ in range(world)]
        dist.all_gather(gathered, w0[0])
        same_init = all(torch.equal(g, gathered[0]) for g in gathered)
        for i, p in enumerate(model.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        ddp.sync_gradients()
        want = [(sum(range(1, world + 1)) / world) * (i + 1) for i in range(4)]
        ok = all(torch.allclose(p.grad, torch.full_like(p, w)) for p, w in zip(model.parameters(), want))
        ts = [torch.full((3,), float(rank)), None, torch.full((2, 2), 2.0 * rank)]
        allreduce_mean_(ts)
        ok2 = torch.allclose(ts[0], torch.full((3,), (world - 1) / 2)) and torch.allclose(
            ts[2], torch.full((2, 2), float(world - 1)))
        q.put((rank, same_init, ok, ok2))
    finally:
        dist.destroy_process_group()


def _accum_worker(rank, world, port, q):
    """The overlapped hook under gradient accumulation: backward twice per optimiser step, with and
    without no_sync(); the result must be the average over ranks of the summed micro-step gradients."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = torch.nn.Linear(4, 3)
        ddp = DataParallel(model)
        model._grad_ready_hook = ddp._on_grads_ready        # what DataParallel installs on the CUDA model
        params = list(model.parameters())

        def fake_backward(scale):
            """What the model's hand-written backward does: gradients of a group in one flat buffer, hook, then
            autograd's AccumulateGrad (`p.grad = g` or `p.grad += g`).  The first parameter is produced last."""
            for p in reversed(params):
                flat = torch.full((p.numel(),), scale * (rank + 1.0))
                model._grad_ready_hook(flat)
                g = flat.view_as(p)
                if p.grad is None:
                    p.grad = g
                else:
                    p.grad += g

        ok = True
        mean_rank = sum(range(1, world + 1)) / world
        for use_no_sync in (True, False):
            for p in params:
                p.grad = None
            if use_no_sync:
                with ddp.no_sync():
                    fake_backward(1.0)
            else:
                fake_backward(1.0)                           # starts overlapped reductions
            fake_backward(10.0)                              # accumulates: must not race, must not double count
            ddp.sync_gradients()
            ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 11.0 * mean_rank)) for p in params)
            ok = ok and not ddp._pending
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gradient_accumulation_with_overlap_hook_world_size_2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_accum_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results


def test_gradient_allreduce_world_size_2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same_init, ok, ok2 in results:
        assert same_init, f"rank {rank}: weights differ after broadcast"
        assert ok, f"rank {rank}: averaged gradients wrong"
        assert ok2, f"rank {rank}: allreduce_mean_ wrong"


def test_bucketing_respects_capacity_and_order():
    ts = [torch.zeros(10), torch.zeros(30), torch.zeros(5), torch.zeros(100)]
    b = _buckets(ts, cap_bytes=160)          # 40 floats per bucket
    assert [len(x) for x in b] == [2, 1, 1]
    assert b[0][0] is ts[0] and b[2][0] is ts[3]
    assert sum(t.numel() for bb in b for t in bb) == 145


def test_shard_batch_is_even_and_disjoint():
    parts = [list(shard_batch(32, r, 8)) for r in range(8)]
    assert all(len(p) == 4 for p in parts)
    assert sorted(i for p in parts for i in p) == list(range(32))


def test_single_process_sync_is_a_noop():
    model = torch.nn.Linear(3, 2)
    for p in model.parameters():
        p.grad = torch.ones_like(p)
    DataParallel(model).sync_gradients()
    assert all(torch.equal(p.grad, torch.ones_like(p)) for p in model.parameters())
