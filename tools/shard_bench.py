#!/usr/bin/env python3
"""Id-range sharded EquSolver (fpie_b200/shard.py) against one GPU on the same system.

  torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/shard_bench.py [--size 4096] [--iters 2000] [--depth 16]

Every rank builds the same ring-mask system on its GPU (EquProcessor reset -> core.system()), rank 0 times the
single-GPU gather solver, then all ranks time the sharded one; prints one JSON line with both and the parity of the
sharded state against the single-GPU state."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-poisson-image-editing_b200")]
import numpy as np
import torch
import torch.distributed as dist

import fpie_b200
from fpie_b200 import shard, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--depth", type=int, default=16)
    ap.add_argument("--labelling", default="rowmajor", choices=["rowmajor", "shuffled"])
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    src, mask, tgt = synth.make_problem("ring", args.size, args.size, seed=0)
    proc = fpie_b200.EquProcessor("avg", mode="gather", device=dev)
    n = proc.reset(src, mask, tgt)
    A, X, B = proc.core.system()
    del proc
    if args.labelling == "shuffled":
        perm = np.concatenate([[0], 1 + np.random.default_rng(1).permutation(n - 1)])
        inv = np.argsort(perm)
        A, X, B = perm[A[inv]].astype(np.int32), X[inv], B[inv]
    single = fpie_b200.EquSolver(256, device=dev, mode="gather")
    single.reset(n, A, X, B)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    single.sweeps_async(200)
    torch.cuda.synchronize()
    ev[0].record()
    single.sweeps_async(args.iters)
    ev[1].record()
    torch.cuda.synchronize()
    single_ms = ev[0].elapsed_time(ev[1])
    want = single.state()
    info1 = single.info()
    del single
    solver = shard.make_sharded_equ_solver(dist, depth=args.depth, device=dev)
    t0 = time.perf_counter()
    solver.reset(n, A, X, B)
    reset_s = time.perf_counter() - t0
    solver.sweeps(200 - 200 % args.depth)  # warm-up: whole intervals, so the timed sweeps start after an exchange
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    solver.sweeps(args.iters)
    ev[1].record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev[0].elapsed_time(ev[1])], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    solver.sweeps(200 % args.depth)
    got = solver.state()
    ghosts = torch.tensor([solver.plan.ghosts], device="cuda")
    dist.all_reduce(ghosts, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({
            "what": f"EquSolver gather path, {args.size}^2 ring mask ({args.labelling} ids), {args.iters} sweeps",
            "unknowns": n - 1, "n_gpus": world, "depth": args.depth,
            "single_gpu": {"gupd_per_s": (n - 1) * args.iters / single_ms / 1e6, "ms": single_ms, "table": info1["table"],
                           "path": info1["path"]},
            "sharded": {"gupd_per_s": (n - 1) * args.iters / float(ms) / 1e6, "ms": float(ms),
                        "exchanges": solver.exchanges, "max_ghosts_per_rank": int(ghosts),
                        "bytes_sent_rank0": solver.bytes_sent, "reset_s": reset_s,
                        "table": solver.core.solver.info()["table"], "path": solver.core.solver.info()["path"]},
            "speedup": single_ms / float(ms),
            "state_bit_exact": bool(np.array_equal(got, want)),
        }))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
