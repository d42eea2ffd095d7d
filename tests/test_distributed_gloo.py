"""world_size-2 gloo test (CPU) of the multi-GPU eval plumbing: impressions sharded by rank with the
global pair index preserved, per-rank metric partial sums (here produced by the CPU oracle — the
GPU box produces them with lime_rank_metrics / lime_metrics_reduce), one SUM all-reduce, global means
equal to the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lime_cikm25_b200 import parallel, synth
from oracle import lime_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _per_impression(imp, scores):
    rows = []
    for i in range(imp.num_impressions):
        s, y = scores[imp.cand_off[i]:imp.cand_off[i + 1]], imp.labels[imp.cand_off[i]:imp.cand_off[i + 1]]
        rows.append(O.impression_metrics(O.rank_impression(s), y))
    return np.asarray(rows, np.float64)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    imp = synth.make_impressions(64, 300, seed=9)
    scores = np.random.default_rng(3).standard_normal(imp.num_pairs).astype(np.float32)
    scores[::7] = 0.0                                              # ties
    mine, base, total = parallel.shard_impressions(imp, rank, world)
    assert total == imp.num_pairs
    per = _per_impression(mine, scores[base:base + mine.num_pairs])
    sums = torch.tensor(list(per.sum(0)) + [float(per.shape[0])], dtype=torch.float64)
    means = parallel.all_reduce_sums(sums)
    # the reference's mini-batch tail rule must not depend on the shard: global pair index is kept
    q.put((rank, means, base, mine.num_pairs))
    dist.barrier()
    dist.destroy_process_group()


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_from_env(backend="gloo")
    from lime_cikm25_b200 import trainer
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(7, 5)), torch.nn.Parameter(torch.randn(11)),
              torch.nn.Parameter(torch.randn(3, 3)), torch.nn.Parameter(torch.randn(2), requires_grad=False)]
    params[0].grad = torch.full((7, 5), float(rank + 1))
    params[1].grad = torch.arange(11.0) * (rank + 1)
    if rank == 0:
        params[2].grad = torch.ones(3, 3)            # a parameter that got no gradient on rank 1
    nbytes = trainer.allreduce_gradients(params)
    # numpy, not torch tensors: a tensor travels through the queue as a file descriptor served by THIS process, which may have
    # exited before the parent asks for it (ConnectionResetError)
    q.put((rank, params[0].grad.numpy().copy(), params[1].grad.numpy().copy(),
           None if params[2].grad is None else params[2].grad.numpy().copy(), nbytes))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_gloo_world2():
    """The data-parallel gradient exchange of the training path: mean over ranks, same flat layout on every
    rank even when a parameter has no gradient somewhere."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g0, g1, g2, nbytes in out:
        assert np.allclose(g0, np.full((7, 5), 1.5)) and np.allclose(g1, np.arange(11.0) * 1.5)
        assert nbytes == (35 + 11 + 9) * 4
    assert np.allclose(out[0][3], np.full((3, 3), 0.5)) and out[1][3] is None


def test_sharded_metric_reduction_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    imp = synth.make_impressions(64, 300, seed=9)
    scores = np.random.default_rng(3).standard_normal(imp.num_pairs).astype(np.float32)
    scores[::7] = 0.0
    want = _per_impression(imp, scores).mean(0)
    for rank, means, base, n in out:
        assert np.allclose(means, want, rtol=0, atol=1e-12)
    assert out[0][2] == 0 and out[1][2] == out[0][3] and out[0][3] + out[1][3] == imp.num_pairs


def test_dp_launcher_dry_run_world2():
    """lime_cikm25_b200.main (reference main.py:28 + trainer.py:246-426) with --world_size 2 on CPU: mp.spawn, gloo
    rendezvous on 127.0.0.1, the DistributedSampler-equivalent partition covers every sample each epoch, the dev
    shards cover the impression set, and both ranks leave through barrier + destroy_process_group (exit code 0)."""
    from lime_cikm25_b200 import main as launcher
    assert launcher.main(["--world_size", "2", "--dry-run", "--epoch", "2", "--master-port", str(_free_port()),
                          "--synthetic-train", "101", "--synthetic-dev", "37"]) == 0


def test_dp_launcher_behaviors_feed_the_dataset():
    """The synthetic behaviours have the tuple layout Train_Dataset reads (dataset.py:41-76): negative sampling and the
    padded history seconds work on them, and the single-process dry run partitions trivially."""
    from lime_cikm25_b200 import dataset as D, main as launcher
    beh = launcher.make_train_behaviors(20, 500, 50, seed=1)
    np.random.seed(0)
    s, f, l = D.negative_sampling(beh, 4)
    assert s.shape == (20, 5) and (s[:, 0] == [b[3] for b in beh]).all()
    assert all(len(D.pad_history_seconds(b[9], 50)) == 50 for b in beh)
    args = launcher.parse_args(["--dry-run", "--epoch", "1", "--synthetic-train", "20", "--synthetic-dev", "8"])
    hist = launcher.run_worker(0, 1, args)
    assert hist == [(1, 20, 8)]
