"""world_size-2 gloo test (CPU) of the episode sharding + metric all-reduce: the reduced
statistics must be IDENTICAL to a single-process run (integer counts bit-exact)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _fake_video(e, v):
    g = torch.Generator().manual_seed(1991 + 17 * e + v)
    return torch.randn(20 + (e + v) % 7, 5, generator=g), (e + v) % 5


def _run(rank, world, port, n_episodes, out):
    import torch.distributed as dist
    from orbit_b200.evaluation import ShardedFrameAccuracy, shard_episodes
    if world > 1:
        os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
    ev = ShardedFrameAccuracy(torch.device('cpu'))
    for e in shard_episodes(n_episodes, rank, world):
        for v in range(3):
            logits, label = _fake_video(e, v)
            ev.append_video(logits, label)
    stats = ev.reduce()
    if rank == 0:
        out.put(stats)
    if world > 1:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_two_rank_metric_reduce_equals_single_process():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    _run(0, 1, 0, 11, q)
    single = q.get()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, 2, port, 11, q)) for r in range(2)]
    for p in procs:
        p.start()
    double = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert double['correct_frames'] == single['correct_frames'] and double['frames'] == single['frames']
    assert double['videos'] == single['videos'] == 33
    assert abs(double['frame_acc_mean_over_videos'] - single['frame_acc_mean_over_videos']) < 1e-12
    assert abs(double['frame_acc_ci95'] - single['frame_acc_ci95']) < 1e-12


def test_shard_episodes_partitions_everything_once():
    from orbit_b200.evaluation import shard_episodes
    for world in (1, 2, 4, 8):
        seen = sorted(e for r in range(world) for e in shard_episodes(850, r, world))
        assert seen == list(range(850))


def _grad_run(rank, world, port, out):
    import torch.distributed as dist
    from orbit_b200.evaluation import shard_episodes
    from orbit_b200.training import allreduce_gradients
    if world > 1:
        os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    params = [torch.nn.Parameter(torch.randn(7, 3, generator=g)), torch.nn.Parameter(torch.randn(5, generator=g)),
              torch.nn.Parameter(torch.randn(2, generator=g))]
    tasks_per_batch = 5
    for t in shard_episodes(tasks_per_batch, rank, world):
        x = torch.randn(3, generator=torch.Generator().manual_seed(100 + t))
        loss = ((params[0] @ x).sum() * (t + 1) + (params[1] ** 2).sum() * t) / tasks_per_batch    # params[2]: never used
        loss.backward()
    n = allreduce_gradients(params)
    if rank == 0:
        out.put((n, [None if p.grad is None else p.grad.clone() for p in params]))
    if world > 1:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_single_process():
    """DP meta-training: tasks of one optimiser step dealt over 2 ranks + one gradient all-reduce == all tasks on one rank."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    _grad_run(0, 1, 0, q)
    n1, single = q.get()
    port = _free_port()
    procs = [ctx.Process(target=_grad_run, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n2, double = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert n1 == n2 == 7 * 3 + 5 + 2
    for a, b in zip(single, double):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
