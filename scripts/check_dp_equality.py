#!/usr/bin/env python
"""N-GPU == 1-GPU on the REAL engine (``engine.TrainStep``: captured step graph + NCCL all-reduce of the flat bucket +
fused clip/AdamW), SURVEY.md 8(e) parity target.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/check_dp_equality.py [--dtype fp32|bf16] [--steps 4]

Every rank holds one contiguous half (``dp.shard_ranges``) of the same global batch, collated with ``lg_inc="bonds"`` (so
that sharding does not change the graph), and trains with ``loss_scale = 1/world``.  Rank 0 also trains an identical
model on the whole batch with a single-rank ``TrainStep``.  After every step the reduced flat gradient bucket and the
parameters must agree: rtol 1e-4 in fp32; in bf16 within 2e-2 of each tensor's scale (summation order differs between the
two shardings).  All ranks must hold bit-identical buckets / parameters.  Prints one JSON line; exit code 1 on mismatch."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gnn_elasticity_predictor_b200 as pkg  # noqa: E402
from gnn_elasticity_predictor_b200 import dp, engine  # noqa: E402
from gnn_elasticity_predictor_b200.synthetic import collate, make_crystal, zscore_targets  # noqa: E402

ARCH = dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=256, layers=2, heads=4)


def build(dev, cd):
    torch.manual_seed(42)
    m = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.0, **ARCH), 2).to(dev)
    m.base.compute_dtype = cd
    m.train()
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", choices=["fp32", "bf16"], default="fp32")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--graphs-per-rank", type=int, default=16)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = dp.init_from_env("nccl")
    if world < 2:
        raise SystemExit("launch with torchrun --nproc-per-node 2 (or more)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cd = torch.float32 if args.dtype == "fp32" else torch.bfloat16

    gen = torch.Generator().manual_seed(0)
    graphs = [make_crystal(10, 6, gen) for _ in range(args.graphs_per_rank * world)]   # equal weights -> equal shards
    lo, hi = dp.shard_ranges([int(g.lg_edge_index.size(1)) for g in graphs], world)[rank]
    assert hi - lo == args.graphs_per_rank
    shard = collate(graphs[lo:hi], lg_inc="bonds").to(dev)
    tz_shard = zscore_targets(shard.y, shard.num_graphs)
    model = build(dev, cd)
    # phase 1 (steps 0 .. steps-1): lr = 0, the weights stand still, so EVERY step -- the eager warm-up step and the
    # replays of the captured graph with the collective inside -- must reproduce the single-rank gradient to rtol;
    # phase 2 (2 more steps): lr = 1e-3 on both sides, the parameter updates are compared
    step = engine.TrainStep(model, lr=0.0, weight_decay=1e-4, loss_scale=1.0 / world, graph=not args.no_graph,
                            graph_warmup=1)
    assert step.world == world

    single = full = tz_full = None
    if rank == 0:
        full = collate(graphs, lg_inc="bonds").to(dev)
        tz_full = zscore_targets(full.y, full.num_graphs)
        single = engine.TrainStep(build(dev, cd), lr=0.0, weight_decay=1e-4, loss_scale=1.0, graph=not args.no_graph,
                                  graph_warmup=1, data_parallel=False)
        assert single.world == 1

    rtol = 1e-4 if cd == torch.float32 else 2e-2
    worst = {"grad": 0.0, "param": 0.0, "loss": 0.0}
    ok = True
    for s in range(args.steps + 2):
        if s == args.steps:
            step.opt.set_lr(1e-3)
            if single is not None:
                single.opt.set_lr(1e-3)
        loss, _, _ = step.step(shard, tz_shard)
        torch.cuda.synchronize()
        # every rank holds the same reduced bucket / parameters, bit for bit
        for name, t in (("grad", step.bucket.flat), ("param", step.opt.flat_params)):
            gathered = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(gathered, t)
            if not all(torch.equal(gathered[0], g) for g in gathered[1:]):
                ok = False
                print(f"rank {rank}: {name} differs between ranks at step {s}", file=sys.stderr)
        losses = [torch.zeros_like(loss) for _ in range(world)]
        dist.all_gather(losses, loss.detach().clone())
        if rank == 0:
            loss1, _, _ = single.step(full, tz_full)
            torch.cuda.synchronize()
            mean_loss = float(torch.stack(losses).mean())
            worst["loss"] = max(worst["loss"], abs(mean_loss - float(loss1)) / max(abs(float(loss1)), 1e-12))
            ga, gb = step.bucket.flat, single.bucket.flat
            pa, pb = step.opt.flat_params, single.opt.flat_params
            gmax = float(gb.abs().max())
            for p, off in zip(single.bucket.params, single.bucket.offsets):
                n = p.numel()
                x, y = ga[off:off + n].double(), gb[off:off + n].double()
                scale = float(y.abs().max())
                if scale <= 1e-3 * gmax:        # e.g. lin_key.bias: true gradient exactly zero, rounding noise only
                    continue
                err = float((x - y).abs().max()) / scale
                worst["grad"] = max(worst["grad"], err)
                # parameters move ~lr per step (Adam's first steps are ~lr * sign(g)): an element whose gradient is within
                # rounding of zero may flip, so the update is compared in RMS over the tensor, in units of lr
                perr = float((pa[off:off + n].double() - pb[off:off + n].double()).pow(2).mean().sqrt()) / 1e-3
                worst["param"] = max(worst["param"], perr)
                pbound = 2e-2 if cd == torch.float32 else 0.5
                # once the weights move, the two runs' parameters differ by the (bounded) update error: gradients are
                # then only held to 10x the tolerance
                if err > (10 * rtol if s > args.steps else rtol) or perr > pbound:
                    ok = False
                    print(f"step {s} tensor @{off} (n={n}): grad rel {err:.3e} (<= {rtol:g}), param rms err / lr "
                          f"{perr:.3e} (<= {pbound:g})", file=sys.stderr)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check": "dp_equality", "world": world, "dtype": args.dtype, "steps": args.steps,
                          "graphs_per_rank": args.graphs_per_rank, "graph_replays": step.replays,
                          "allreduce_in_graph": bool(step.allreduce_in_graph and not args.no_graph),
                          "rtol_grad": rtol, "worst_rel_grad": worst["grad"], "worst_param_err_over_lr": worst["param"],
                          "worst_rel_loss": worst["loss"], "ok": bool(flag.item())}), flush=True)
    good = bool(flag.item())
    sys.stdout.flush()
    if not good:
        os._exit(1)
    dp.shutdown([step] + ([single] if single is not None else []))
    sys.exit(0)


if __name__ == "__main__":
    main()
