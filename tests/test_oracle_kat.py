"""Known-answer tests that pin the PyG shim (the reference holds no vectors for this path)."""
import ctypes
import math
import os
import subprocess

import pytest
import torch

import oracle

oracle.install_shim()
from torch_geometric.nn import TransformerConv, global_mean_pool  # noqa: E402  (the shim)
from torch_geometric.data import Batch, Data  # noqa: E402

torch.manual_seed(0)


def _conv(hidden=8, heads=2, dtype=torch.float64):
    torch.manual_seed(3)
    return TransformerConv(hidden, hidden // heads, heads=heads, edge_dim=hidden, beta=True).to(dtype)


def _agg_only(conv, x, index, ea):
    """aggregate before the beta gate, recovered from the shim's pieces"""
    H, C = conv.heads, conv.out_channels
    q = conv.lin_query(x).view(-1, H, C); k = conv.lin_key(x).view(-1, H, C); v = conv.lin_value(x).view(-1, H, C)
    e = conv.lin_edge(ea).view(-1, H, C)
    return q, k, v, e


def test_kat1_single_edge_and_isolated_node():
    conv = _conv()
    x = torch.randn(3, 8, dtype=torch.float64)
    ea = torch.randn(1, 8, dtype=torch.float64)
    index = torch.tensor([[0], [1]])
    out = conv(x, index, ea)
    q, k, v, e = _agg_only(conv, x, index, ea)
    agg = torch.zeros(3, 8, dtype=torch.float64)
    agg[1] = (v[0] + e[0]).reshape(-1)            # alpha == 1 for the only in-edge
    xr = conv.lin_skip(x)
    beta = torch.sigmoid(conv.lin_beta(torch.cat([agg, xr, agg - xr], -1)))
    expect = beta * xr + (1 - beta) * agg
    assert torch.allclose(out, expect, atol=1e-12)
    # nodes 0 and 2 have no in-edges: agg = 0 -> out = beta * x_r
    assert torch.allclose(out[2], (beta * xr)[2], atol=1e-12)


def test_kat2_duplicate_edges_equal_one_edge():
    conv = _conv()
    x = torch.randn(4, 8, dtype=torch.float64)
    ea1 = torch.randn(1, 8, dtype=torch.float64)
    one = conv(x, torch.tensor([[2], [0]]), ea1)
    two = conv(x, torch.tensor([[2, 2], [0, 0]]), ea1.repeat(2, 1))
    assert torch.allclose(one, two, atol=1e-12)


def test_kat3_edge_permutation_invariance():
    conv = _conv()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 8, dtype=torch.float64, generator=g)
    index = torch.randint(0, 6, (2, 20), generator=g)
    ea = torch.randn(20, 8, dtype=torch.float64, generator=g)
    perm = torch.randperm(20, generator=g)
    assert torch.allclose(conv(x, index, ea), conv(x, index[:, perm], ea[perm]), atol=1e-12)


def test_kat4_gradcheck_fp64():
    conv = _conv(hidden=4, heads=2)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(4, 4, dtype=torch.float64, generator=g, requires_grad=True)
    ea = torch.randn(7, 4, dtype=torch.float64, generator=g, requires_grad=True)
    index = torch.randint(0, 4, (2, 7), generator=g)
    assert torch.autograd.gradcheck(lambda a, b: conv(a, index, b), (x, ea), eps=1e-6, atol=1e-5)


def test_dense_masked_attention_crosscheck():
    """On a simple graph (no duplicate edges) the conv equals dense masked multi-head attention."""
    conv = _conv(hidden=8, heads=2)
    n = 5
    g = torch.Generator().manual_seed(7)
    adj = torch.rand(n, n, generator=g) < 0.5          # adj[i, j]: edge j -> i
    adj[3] = False                                      # node 3 has no in-edges
    dst, src = adj.nonzero(as_tuple=True)
    index = torch.stack([src, dst])
    x = torch.randn(n, 8, dtype=torch.float64, generator=g)
    ea = torch.randn(index.size(1), 8, dtype=torch.float64, generator=g)
    out = conv(x, index, ea)
    q, k, v, e = _agg_only(conv, x, index, ea)
    H, C = 2, 4
    e_dense = torch.zeros(n, n, H, C, dtype=torch.float64)
    e_dense[dst, src] = e
    logits = torch.einsum("ihc,ijhc->ijh", q, k.unsqueeze(0) + e_dense) / math.sqrt(C)
    logits = logits.masked_fill(~adj.unsqueeze(-1), float("-inf"))
    alpha = torch.softmax(logits, dim=1)
    alpha = torch.nan_to_num(alpha, nan=0.0)           # rows without in-edges
    agg = torch.einsum("ijh,ijhc->ihc", alpha, v.unsqueeze(0) + e_dense).reshape(n, H * C)
    xr = conv.lin_skip(x)
    beta = torch.sigmoid(conv.lin_beta(torch.cat([agg, xr, agg - xr], -1)))
    assert torch.allclose(out, beta * xr + (1 - beta) * agg, atol=1e-10)


def test_global_mean_pool_and_collate_rules():
    d1 = Data(x=torch.ones(2, 3), edge_index=torch.tensor([[0, 1], [1, 0]]), edge_attr=torch.zeros(2, 1))
    d1.lg_edge_index = torch.tensor([[0], [1]]); d1.lg_edge_attr = torch.zeros(1, 1)
    d1.global_x = torch.zeros(59, 1); d1.y = torch.tensor([1.0, 2.0]); d1.material_id = "a"
    d2 = Data(x=2 * torch.ones(3, 3), edge_index=torch.tensor([[0, 1, 2], [1, 2, 0]]), edge_attr=torch.zeros(3, 1))
    d2.lg_edge_index = torch.tensor([[0, 1], [1, 2]]); d2.lg_edge_attr = torch.zeros(2, 1)
    d2.global_x = torch.zeros(59, 1); d2.y = torch.tensor([3.0, 4.0]); d2.material_id = "b"
    b = Batch.from_data_list([d1, d2])
    assert b.num_graphs == 2 and b.batch.tolist() == [0, 0, 1, 1, 1]
    assert b.edge_index.tolist() == [[0, 1, 2, 3, 4], [1, 0, 3, 4, 2]]
    # PyG default __inc__: any key containing "index" is offset by num_nodes (atoms), also lg_edge_index
    assert b.lg_edge_index.tolist() == [[0, 2, 3], [1, 3, 4]]
    assert b.global_x.shape == (118, 1) and b.y.shape == (4,) and b.material_id == ["a", "b"]
    pooled = global_mean_pool(b.x, b.batch)
    assert torch.equal(pooled, torch.tensor([[1.0] * 3, [2.0] * 3]))


# ---- independent plain-C restatement agrees with the shim ------------------------------------------------
def _c_lib():
    path = os.path.join(oracle.ORACLE_DIR, "_ref", "libconv_ref.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", oracle.ORACLE_DIR], check=True, capture_output=True)
    return ctypes.CDLL(path)


def _dp(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("hidden,heads,n,e", [(8, 2, 6, 25), (48, 3, 9, 60), (32, 1, 5, 0)])
def test_c_restatement_matches_shim(hidden, heads, n, e):
    lib = _c_lib()
    conv = _conv(hidden, heads)
    g = torch.Generator().manual_seed(hidden + e)
    x = torch.randn(n, hidden, dtype=torch.float64, generator=g)
    index = torch.randint(0, n, (2, e), generator=g)
    ea = torch.randn(e, hidden, dtype=torch.float64, generator=g)
    q, k, v, ee = (t.reshape(-1, hidden).contiguous() for t in _agg_only(conv, x, index, ea))
    agg = torch.empty(n, hidden, dtype=torch.float64)
    src, dst = index[0].contiguous(), index[1].contiguous()
    rc = lib.ref_conv_core_fwd(_dp(q.detach()), _dp(k.detach()), _dp(v.detach()), _dp(ee.detach()), _dp(src), _dp(dst),
                               ctypes.c_int64(e), ctypes.c_int64(n), heads, hidden // heads, _dp(agg))
    assert rc == 0
    xr = conv.lin_skip(x)
    beta = torch.sigmoid(conv.lin_beta(torch.cat([agg, xr, agg - xr], -1)))
    expect = conv(x, index, ea)
    assert torch.allclose(beta * xr + (1 - beta) * agg, expect, atol=1e-11)
    # block tail: x + relu(LN(out))
    ln = torch.nn.LayerNorm(hidden).double()
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5); ln.bias.uniform_(-0.5, 0.5)
    y = torch.empty(n, hidden, dtype=torch.float64)
    wb = conv.lin_beta.weight.detach().reshape(-1).contiguous()
    rc = lib.ref_gate_ln_relu_res(_dp(agg), _dp(xr.detach().contiguous()), _dp(x), _dp(wb), _dp(ln.weight.detach()),
                                  _dp(ln.bias.detach()), ctypes.c_int64(n), hidden, _dp(y))
    assert rc == 0
    assert torch.allclose(y, x + torch.relu(ln(expect)), atol=1e-10)


@pytest.mark.parametrize("n_nodes,n_edges", [(1, 0), (7, 50), (1000, 20000)])
def test_c_stable_sort_matches_torch(n_nodes, n_edges):
    lib = _c_lib()
    g = torch.Generator().manual_seed(n_edges)
    key = torch.randint(0, n_nodes, (n_edges,), generator=g)
    rowptr = torch.empty(n_nodes + 1, dtype=torch.int32)
    eid = torch.empty(max(n_edges, 1), dtype=torch.int32)
    assert lib.ref_stable_sort_by_key(_dp(key), ctypes.c_int64(n_edges), ctypes.c_int64(n_nodes), _dp(rowptr), _dp(eid)) == 0
    _, perm = torch.sort(key, stable=True)
    assert torch.equal(eid[:n_edges].long(), perm)
    counts = torch.bincount(key, minlength=n_nodes)
    assert torch.equal(rowptr.long(), torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]))
