"""Synthetic crystal generator + collate: index conventions of fetch.py and PyG's default Batch rules."""
import pytest
import torch

import oracle
from gnn_elasticity_predictor_b200.synthetic import (collate, make_crystal, ring_topology, synthetic_batch,
                                                      zscore_targets)

oracle.install_shim()
from torch_geometric.data import Batch, Data  # noqa: E402  (the shim restating PyG's collate)


@pytest.mark.parametrize("atoms,k", [(6, 4), (16, 12), (32, 12), (40, 16)])
def test_ring_topology_counts_and_order(atoms, k):
    ei, lg = ring_topology(atoms, k)
    e = atoms * k
    assert ei.shape == (2, e) and lg.shape == (2, e * (k - 1))
    # source-major emission (fetch.py:389-396) and e1-major line graph (fetch.py:421-444)
    assert torch.equal(ei[0], torch.sort(ei[0], stable=True).values)
    assert torch.equal(lg[0], torch.sort(lg[0], stable=True).values)
    # every bond has its reverse; K-regular
    pairs = set(map(tuple, ei.t().tolist()))
    assert all((j, i) in pairs for i, j in pairs)
    assert torch.equal(torch.bincount(ei[1], minlength=atoms), torch.full((atoms,), k))
    # line-graph edge (i->j) -> (j->k'): shares the middle atom, never the exact reverse bond
    e1, e2 = lg
    assert torch.equal(ei[1][e1], ei[0][e2])
    assert not bool(((ei[1][e2] == ei[0][e1]) & (ei[0][e2] == ei[1][e1])).any())


def test_dups_variant_has_duplicates_and_self_loops():
    g = torch.Generator().manual_seed(0)
    ei, lg = ring_topology(20, 6, dups=True, gen=g)
    assert ei.size(1) > 120
    assert bool((ei[0] == ei[1]).any())
    assert len(set(map(tuple, ei.t().tolist()))) < ei.size(1)
    assert torch.equal(ei[0], torch.sort(ei[0], stable=True).values)
    assert int(lg.max()) < ei.size(1)


def test_collate_offsets_pyg_vs_bonds_and_vectorised_equivalence():
    gen = torch.Generator().manual_seed(0)
    graphs = [make_crystal(6, 4, gen) for _ in range(3)]
    pyg, bonds = collate(graphs, "pyg"), collate(graphs, "bonds")
    e = graphs[0].edge_index.size(1)
    l = graphs[0].lg_edge_index.size(1)
    assert torch.equal(pyg.edge_index, bonds.edge_index)
    assert torch.equal(pyg.lg_edge_index[:, l:2 * l], graphs[1].lg_edge_index + 6)      # atoms (the quirk)
    assert torch.equal(bonds.lg_edge_index[:, l:2 * l], graphs[1].lg_edge_index + e)    # bonds
    assert pyg.global_x.shape == (3 * 59, 1) and pyg.sg_one_hot.shape == (3 * 230, 1) and pyg.y.shape == (6,)
    assert pyg.batch.tolist() == [0] * 6 + [1] * 6 + [2] * 6
    for mode in ("pyg", "bonds"):
        v = synthetic_batch(3, 6, 4, seed=0, lg_inc=mode)
        ref = collate(graphs, mode)
        assert torch.equal(v.edge_index, ref.edge_index) and torch.equal(v.lg_edge_index, ref.lg_edge_index)
        assert v.x.shape == ref.x.shape and v.lg_edge_attr.shape == ref.lg_edge_attr.shape


def test_collate_matches_pyg_default_batch_rules():
    """Our 'pyg' collate is exactly what PyG's Batch.from_data_list does to the reference's Data objects."""
    gen = torch.Generator().manual_seed(1)
    graphs = [make_crystal(5 + i, 4, gen) for i in range(3)]
    datas = []
    for gph in graphs:
        d = Data(x=gph.x, edge_index=gph.edge_index, edge_attr=gph.edge_attr)
        d.lg_edge_index, d.lg_edge_attr = gph.lg_edge_index, gph.lg_edge_attr
        d.global_x, d.sg_one_hot, d.y = gph.global_x, gph.sg_one_hot, gph.y
        datas.append(d)
    pb = Batch.from_data_list(datas)
    ours = collate(graphs, "pyg")
    for key in ("x", "edge_index", "edge_attr", "lg_edge_index", "lg_edge_attr", "global_x", "sg_one_hot", "batch", "y"):
        assert torch.equal(getattr(pb, key), getattr(ours, key)), key
    assert pb.num_graphs == ours.num_graphs


def test_baseline_config_sizes():
    b = synthetic_batch(64, 16, 12, seed=0)           # BASELINE config 1
    assert b.sizes == {"B": 64, "N": 1024, "E": 12288, "L": 135168}
    z = zscore_targets(b.y, b.num_graphs)
    assert z.shape == (64, 2) and abs(float(z.mean())) < 0.5
    # determinism
    b2 = synthetic_batch(64, 16, 12, seed=0)
    assert torch.equal(b.x, b2.x) and torch.equal(b.lg_edge_attr, b2.lg_edge_attr)
