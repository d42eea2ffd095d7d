"""Data-parallel training launcher behind ``--world_size`` (reference main.py:15-43 + trainer.py:246-426).

    python -m lime_cikm25_b200.main --world_size 8 --epoch 2              # spawns 8 ranks (mp.spawn, as main.py:28)
    torchrun --nproc-per-node 8 -m lime_cikm25_b200.main --world_size 8   # or joins the group torchrun describes

Every rank (one process per GPU) runs the SAME loop the reference's Trainer.train runs (trainer.py:81-233), with the
three differences a working data-parallel version needs (the reference's own distributed_train is stale, SURVEY.md
section 2.1):

  * the epoch's samples are partitioned by a DistributedSampler-equivalent seeded permutation (dataset.epoch_order,
    trainer.py:293-295) and gathered ON THE DEVICE (dataset.DeviceTrainSet);
  * gradients are averaged with one NCCL all-reduce before the clip (trainer.allreduce_gradients);
  * the dev evaluation is SHARDED by impression over all ranks (parallel.shard_impressions +
    util.evaluate_impressions(sharded=True), one all-reduce of five sums) instead of running on rank 0 while the others
    wait at a barrier (trainer.py:342-411); rank 0 alone writes files.

Shutdown is a barrier + destroy_process_group on every rank (the reference kills the workers with SIGKILL,
trainer.py:411-426).  Without dataset files (this container has none; corpus.py preprocessing is out of scope) the
launcher trains on a synthetic corpus of the MIND shape: ``--synthetic-news / --synthetic-train / --synthetic-dev``.
``--dry-run`` exercises rendezvous, sampler, sharding, collectives and shutdown without a device (gloo; CPU tests).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--world_size", type=int, default=1, help="config.py:88")
    ap.add_argument("--epoch", type=int, default=2, help="config.py:56")
    ap.add_argument("--batch_size", type=int, default=32, help="per-rank mini-batch (config.py:57 divided by world_size, trainer.py:252)")
    ap.add_argument("--negative_sample_num", type=int, default=4)
    ap.add_argument("--early_stopping_epoch", type=int, default=5)
    ap.add_argument("--dev_criterion", default="avg", choices=["auc", "mrr", "ndcg5", "ndcg10", "avg"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--vocabulary_size", type=int, default=8000)
    ap.add_argument("--synthetic-news", type=int, default=2000)
    ap.add_argument("--synthetic-train", type=int, default=512, help="training behaviours")
    ap.add_argument("--synthetic-dev", type=int, default=256, help="dev impressions")
    ap.add_argument("--max-steps", type=int, default=0, help="stop every epoch after this many steps (0 = whole epoch)")
    ap.add_argument("--bf16", action="store_true", help="bf16 GEMM mode of the differentiable path")
    ap.add_argument("--result-dir", default="", help="rank 0 writes the dev log and the best checkpoint here")
    ap.add_argument("--master-port", type=int, default=29531)
    ap.add_argument("--dry-run", action="store_true", help="host logic only (no model, no device): gloo")
    return ap.parse_args(argv)


def make_train_behaviors(n, news_num, max_history, seed):
    """Synthetic training behaviours in the tuple layout Train_Dataset reads (dataset.py:41-76, 105-141):
    [0] user, [1] history indices [H], [2] history mask [H], [3] positive, [4] negatives, [6] freshness,
    [7] positive lifetime, [8] negative lifetimes, [9] / [10] history freshness / lifetime lists."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        hl = int(rng.integers(0, max_history + 1))
        hist = np.zeros(max_history, np.int64)
        hist[:hl] = rng.integers(1, news_num, size=hl)
        mask = np.arange(max_history) < hl
        nneg = int(rng.integers(1, 12))
        out.append((int(rng.integers(0, 50000)), hist, mask, int(rng.integers(1, news_num)),
                    rng.integers(1, news_num, size=nneg).tolist(), None, float(np.exp(rng.uniform(0, 14))),
                    float(np.exp(rng.uniform(6, 13))), np.exp(rng.uniform(6, 13, size=nneg)).tolist(),
                    np.exp(rng.uniform(0, 14, size=hl)).tolist(), np.exp(rng.uniform(6, 13, size=hl)).tolist()))
    return out


def criterion(name, auc, mrr, ndcg5, ndcg10):
    return {"auc": auc, "mrr": mrr, "ndcg5": ndcg5, "ndcg10": ndcg10, "avg": (auc + mrr + ndcg5 + ndcg10) / 4.0}[name]


def run_worker(rank, world, args, log=print):
    """One rank.  Returns the per-epoch dev metrics (identical on every rank)."""
    from . import dataset as D, parallel, synth
    device_ok = torch.cuda.is_available() and not args.dry_run
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(args.master_port))
    if world > 1 and not dist.is_initialized():
        if device_ok:
            torch.cuda.set_device(rank % torch.cuda.device_count())
        dist.init_process_group(backend="nccl" if device_ok else "gloo", rank=rank, world_size=world)
    if not device_ok and not args.dry_run:
        raise SystemExit("lime_cikm25_b200.main: no CUDA device (the B200 path has no CPU fallback); --dry-run checks the host logic")
    dev = torch.device("cuda", torch.cuda.current_device()) if device_ok else torch.device("cpu")

    news = synth.make_news_table(args.synthetic_news, vocabulary_size=args.vocabulary_size, seed=1)
    behaviors = make_train_behaviors(args.synthetic_train, news.news_num, 50, seed=2)
    dev_imp = synth.make_impressions(args.synthetic_dev, news.news_num, seed=3)
    my_imp, pair_base, total_pairs = parallel.shard_impressions(dev_imp, rank, world)
    history = []

    if args.dry_run:
        # host logic only: the sampler partitions every epoch exactly, the dev shards cover the set, the collectives
        # every rank must enter are entered, and the group shuts down cleanly
        for e in range(1, args.epoch + 1):
            mine = D.epoch_order(len(behaviors), args.seed, e, rank, world)
            seen = torch.zeros(len(behaviors), dtype=torch.int64)
            seen[mine] += 1
            t = torch.tensor([float(my_imp.num_impressions), float(my_imp.num_pairs)], dtype=torch.float64)
            if world > 1:
                dist.all_reduce(seen)
                dist.all_reduce(t)
            assert int(seen.min()) >= 1 and int(seen.sum()) == ((len(behaviors) + world - 1) // world) * world
            assert int(t[0]) == dev_imp.num_impressions and int(t[1]) == dev_imp.num_pairs == total_pairs
            history.append((e, int(mine.numel()), my_imp.num_impressions))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return history

    import lime_cikm25_b200 as L
    from . import autograd, engine, util
    from .config import default_config
    from .trainer import Trainer
    cfg = default_config(vocabulary_size=args.vocabulary_size, batch_size=args.batch_size, word_embedding_init="skip")
    cfg.seed = args.seed
    torch.manual_seed(args.seed)                      # identical initial weights on every rank (DDP broadcasts rank 0's)
    model = L.Model(cfg)
    model.initialize()
    synth.synthetic_parameters(model, seed=args.seed)
    model = model.to(dev)
    autograd.set_bf16(bool(args.bf16))
    trainer = Trainer(model, cfg)
    tables = D.DeviceNewsTables(news, dev)
    train_set = D.DeviceTrainSet(tables, behaviors, 50, args.negative_sample_num)
    dimp = engine.DeviceImpressions(my_imp, dev)
    best, best_epoch, not_increase = -1.0, 0, 0
    for e in range(1, args.epoch + 1):
        np.random.seed(args.seed * 1000 + e)          # the reference draws from numpy's global generator (dataset.py:41-76): same draws on every rank
        train_set.negative_sampling()
        order = D.epoch_order(len(train_set), args.seed, e, rank, world).to(dev)
        model.train()
        t0, loss_sum, steps = time.perf_counter(), 0.0, 0
        for lo in range(0, order.numel(), args.batch_size):
            batch = train_set.batch(order[lo:lo + args.batch_size])
            loss_sum += float(trainer.step(batch))
            steps += 1
            if args.max_steps and steps >= args.max_steps:
                break
        torch.cuda.synchronize()
        train_s = time.perf_counter() - t0
        # ---- sharded dev evaluation: every rank scores its impressions on its own cache replica ----
        model.eval()
        with torch.no_grad():
            cache = util.build_news_cache(model, news, dev)
            auc, mrr, ndcg5, ndcg10 = util.evaluate_impressions(model, cache, dimp, cfg.batch_size, pair_base, total_pairs,
                                                                sharded=world > 1)
            del cache
        crit = criterion(args.dev_criterion, auc, mrr, ndcg5, ndcg10)
        history.append((e, loss_sum / max(steps, 1), auc, mrr, ndcg5, ndcg10))
        if rank == 0:
            log("epoch %d: %d steps/rank, loss %.4f, %.1f samples/s (all ranks) | dev auc %.4f mrr %.4f ndcg5 %.4f ndcg10 %.4f"
                % (e, steps, loss_sum / max(steps, 1), steps * args.batch_size * world / train_s, auc, mrr, ndcg5, ndcg10))
        if crit >= best:
            best, best_epoch, not_increase = crit, e, 0
            if rank == 0 and args.result_dir:
                os.makedirs(args.result_dir, exist_ok=True)
                torch.save({model.model_name: model.state_dict()}, os.path.join(args.result_dir, model.model_name))
        else:
            not_increase += 1
        if not_increase == args.early_stopping_epoch:
            break
    if rank == 0 and args.result_dir:
        with open(os.path.join(args.result_dir, "dev_log.txt"), "w", encoding="utf-8") as f:
            f.write("Epoch\tAUC\tMRR\tnDCG@5\tnDCG@10\n")
            for h in history:
                f.write("%d\t%.4f\t%.4f\t%.4f\t%.4f\n" % (h[0], h[2], h[3], h[4], h[5]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return history


def _spawned(rank, world, args):
    os.environ["RANK"], os.environ["WORLD_SIZE"], os.environ["LOCAL_RANK"] = str(rank), str(world), str(rank)
    run_worker(rank, world, args)


def main(argv=None):
    args = parse_args(argv)
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:          # under torchrun: one rank per process already
        run_worker(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), args)
    elif args.world_size > 1:
        import torch.multiprocessing as mp
        mp.spawn(_spawned, args=(args.world_size, args), nprocs=args.world_size, join=True)
    else:
        run_worker(0, 1, args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
