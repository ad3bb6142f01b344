"""Generate ``tests/golden/*.pt`` from the reference's OWN model classes.  TEST INFRASTRUCTURE ONLY.

Run in the build container (where ``/root/reference`` is mounted):

    python oracle/gen_golden.py

It imports ``/root/reference/scripts/train.py`` unmodified with the PyG shim first on ``sys.path``
(``oracle.load_reference_train_module``), instantiates ``AlignnRegressor`` / ``HeteroAlignnRegressor`` /
``EdgeUpdateBlock`` / ``NodeUpdateBlock`` under fixed seeds, runs them in fp32 on CPU on small seeded
synthetic batches and stores inputs, ``state_dict``, outputs, the training loss and every parameter
gradient.  The fixtures travel to the GPU box (``/root/reference`` does not), where they pin both the
own-code oracle (``oracle/model_ref.py``) and the CUDA path.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from gnn_elasticity_predictor_b200.synthetic import synthetic_batch, zscore_targets  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

MODEL_CASES = {
    # name: (ctor kwargs, batch kwargs)
    "smoke_arch": (  # reference tests/smoke.py:33-41,117-122 -> node 6 / edge 8 / angle 7 / H 32 / h 1 / 1 layer
        dict(node_dim=6, edge_dim=8, angle_dim=7, global_dim=289, target_dim=2, hidden=32, layers=1, heads=1, dropout=0.0),
        dict(n_graphs=3, atoms=6, k=4, seed=11, lg_inc="pyg", node_dim=6, edge_dim=8, angle_dim=7)),
    "default_dims_h32": (  # default feature dims (206/36/11/289), 2 layers, 4 heads
        dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=32, layers=2, heads=4, dropout=0.0),
        dict(n_graphs=4, atoms=8, k=4, seed=12, lg_inc="pyg")),
    "default_dims_h32_bonds": (
        dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=32, layers=2, heads=4, dropout=0.0),
        dict(n_graphs=4, atoms=8, k=4, seed=13, lg_inc="bonds")),
    "dups_selfloops_h64": (  # duplicate bonds + self loops (periodic images), generic-vs-fast head split
        dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=64, layers=1, heads=4, dropout=0.0),
        dict(n_graphs=3, atoms=7, k=4, seed=14, lg_inc="bonds", dups=True)),
    "default_arch_h256": (  # the default width / heads (train.py:1084-1086): the width the tensor-core kernels serve
        dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=256, layers=1, heads=4, dropout=0.0),
        dict(n_graphs=6, atoms=10, k=6, seed=16, lg_inc="pyg")),
    "odd_width_h48": (  # hidden/8 not a power of two -> generic kernels
        dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=48, layers=1, heads=3, dropout=0.0),
        dict(n_graphs=2, atoms=7, k=4, seed=15, lg_inc="pyg")),
}


def _batch_tensors(b):
    return {k: v.clone() for k, v in b.tensors().items()}


def make_model_case(ref, name, ctor, bkw):
    torch.manual_seed(1234)
    base = ref.AlignnRegressor(**ctor)
    model = ref.HeteroAlignnRegressor(base, ctor["target_dim"])
    model.train()  # dropout is 0: train == eval, but exercises the training code path
    batch = synthetic_batch(**bkw)
    mean, logvar = model(batch)
    target_z = zscore_targets(batch.y, batch.num_graphs)
    # loss of train_epoch_hetero (train.py:655-681), --log-sigma-l2 0.1, floor -2.9, no sample weights
    lv = torch.clamp(logvar, min=ref.MIN_LOGVAR_FLOOR)
    nll = 0.5 * (lv + (mean - target_z).pow(2) / torch.exp(lv))
    loss = nll.mean(dim=1).mean() + 0.1 * (0.5 * lv).pow(2).mean()
    loss.backward()
    embed = model.embed(batch).detach()
    plain = base(batch).detach()
    out = {
        "name": name, "ctor": ctor, "batch_kwargs": bkw, "num_graphs": batch.num_graphs,
        "batch": _batch_tensors(batch),
        "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
        "mean": mean.detach(), "logvar": logvar.detach(), "loss": loss.detach(), "embed": embed,
        "plain_output": plain,
        "grads": {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None},
    }
    return out


def make_block_case(ref, hidden, heads, n_nodes, n_edges, seed):
    """EdgeUpdateBlock / NodeUpdateBlock on a random multigraph (duplicates, self loops, empty rows)."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    eb = ref.EdgeUpdateBlock(hidden, heads, 0.0)
    nb = ref.NodeUpdateBlock(hidden, hidden, heads, 0.0)
    # last quarter of the nodes receives no edges (empty rows); sources cover everything
    dst = torch.randint(0, max(1, (3 * n_nodes) // 4), (n_edges,), generator=g)
    src = torch.randint(0, n_nodes, (n_edges,), generator=g)
    src[:3] = dst[:3]                      # self loops
    src[3:6], dst[3:6] = src[0:3], dst[0:3]  # duplicate edges
    index = torch.stack([src, dst])
    x = torch.randn(n_nodes, hidden, generator=g).requires_grad_(True)
    ea = torch.randn(n_edges, hidden, generator=g).requires_grad_(True)
    gout = torch.randn(n_nodes, hidden, generator=g)
    out = {}
    for tag, blk, args in (("edge_block", eb, (x, index, ea)), ("node_block", nb, (x, index, ea))):
        x.grad = ea.grad = None
        blk.zero_grad()
        y = blk(*args)
        y.backward(gout)
        out[tag] = {
            "state_dict": {k: v.detach().clone() for k, v in blk.state_dict().items()},
            "y": y.detach().clone(), "dx": x.grad.clone(), "dedge": ea.grad.clone(),
            "grads": {k: p.grad.detach().clone() for k, p in blk.named_parameters()},
        }
    out.update(hidden=hidden, heads=heads, index=index, x=x.detach().clone(), edge_attr=ea.detach().clone(),
               gout=gout)
    return out


def main():
    ref = oracle.load_reference_train_module()
    if ref is None:
        raise SystemExit("/root/reference is not mounted: goldens can only be regenerated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = set(sys.argv[1:])          # `python oracle/gen_golden.py default_arch_h256`: just the named model cases
    for name, (ctor, bkw) in MODEL_CASES.items():
        if only and name not in only:
            continue
        case = make_model_case(ref, name, ctor, bkw)
        path = os.path.join(GOLDEN_DIR, f"model_{name}.pt")
        torch.save(case, path)
        print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB  loss={float(case['loss']):.6f}")
    for hidden, heads, n, e, seed in () if only else ((64, 4, 40, 300, 21), (32, 1, 24, 150, 22), (128, 4, 12, 60, 23), (48, 3, 20, 90, 24)):
        case = make_block_case(ref, hidden, heads, n, e, seed)
        path = os.path.join(GOLDEN_DIR, f"blocks_h{hidden}_heads{heads}.pt")
        torch.save(case, path)
        print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
