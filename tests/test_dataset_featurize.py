"""SURVEY.md 8(f) rows N2 (device-resident dataset + on-GPU collate) and N3 (bond features + line graph on the device).

CPU part: the oracle restatement of the reference featuriser (oracle/linegraph_ref.py) against the golden vectors that the
reference's OWN ``build_graph_from_structure`` produced (oracle/gen_golden_linegraph.py), and host-side behaviour.
GPU part (``-m gpu``): the C-ABI kernels against the oracle / the host collate -- bit-exact for every index and every
copied value, 1e-6 (float32 rounding of float64 transcendentals) for the basis features."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import linegraph_ref
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import batching, dataset, featurize

DEV = torch.device("cuda", 0) if torch.cuda.is_available() else None
CASES = ["a", "b", "c", "d"]
KEYS = ("edge_index", "edge_attr", "lg_edge_index", "lg_edge_attr")


def _edges(g):
    return [(int(i), int(j), tuple(int(v) for v in im)) for i, j, im in
            zip(g["bond_src"].tolist(), g["bond_dst"].tolist(), g["bond_image"].tolist())]


# ---- CPU: oracle pinned by the reference's own outputs -------------------------------------------------------------------
@pytest.mark.parametrize("tag", CASES)
def test_oracle_featuriser_matches_reference_golden(tag):
    g = load_golden(f"linegraph_{tag}.pt")
    rc, rg, ac, ag = linegraph_ref.basis()
    mine = linegraph_ref.build_bond_and_line_graph(g["frac"].numpy(), g["lattice"].numpy(), g["en"].tolist(), _edges(g),
                                                   rc, rg, ac, ag)
    for k in KEYS:
        assert torch.equal(mine[k], g[k]), k
    e, l = g["edge_index"].size(1), g["lg_edge_index"].size(1)
    assert g["edge_attr"].shape == (e, 36) and g["lg_edge_attr"].shape == (l, 11)          # edge 36, angle 11


def test_oracle_line_graph_known_answers():
    """One atom bonded to its own +x / -x images: bond 0 = (+x), bond 1 = (-x).  Each bond's only continuation that is not
    its exact reverse is itself (same direction): angle between j->i (reverse) and j->k = pi."""
    frac, lat = np.zeros((1, 3)), np.eye(3) * 2.0
    edges = [(0, 0, (1, 0, 0)), (0, 0, (-1, 0, 0))]
    rc, rg, ac, ag = linegraph_ref.basis()
    out = linegraph_ref.build_bond_and_line_graph(frac, lat, [1.5], edges, rc, rg, ac, ag)
    assert out["lg_edge_index"].tolist() == [[0, 1], [0, 1]]
    assert torch.allclose(out["lg_edge_attr"][:, 8], torch.full((2,), float(np.pi)))
    assert torch.allclose(out["lg_edge_attr"][:, 9], torch.full((2,), -1.0))
    assert torch.allclose(out["edge_attr"][:, 33:], torch.tensor([[1.0, 0, 0], [-1.0, 0, 0]]))
    assert float(out["edge_attr"][0, 32]) == 0.0                                           # |EN_i - EN_i|
    # distance 2.0 sits between RBF centres: the largest response is the nearest centre
    assert int(out["edge_attr"][0, :32].argmax()) == int(np.abs(rc - 2.0).argmin())


def test_store_and_featuriser_refuse_cpu():
    gen = torch.Generator().manual_seed(0)
    graphs = [pkg.make_crystal(8, 6, gen)]
    with pytest.raises(RuntimeError, match="no CPU"):
        dataset.DeviceGraphStore(graphs, "cpu")
    z = torch.zeros(1, 3, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="no CPU"):
        featurize.build_bond_and_line_graph(z, torch.eye(3, dtype=torch.float64)[None], torch.zeros(1), torch.zeros(1).long(),
                                            torch.zeros(1).long(), torch.zeros(1, 3).int(), *featurize.default_basis())


def test_default_basis_matches_reference_formulae():
    rc, rg, ac, ag = linegraph_ref.basis()
    t_rc, t_rg, t_ac, t_ag = featurize.default_basis()
    assert np.array_equal(t_rc.numpy(), rc) and np.array_equal(t_ac.numpy(), ac) and t_rg == rg and t_ag == ag


# ---- GPU: featuriser ------------------------------------------------------------------------------------------------------
def _gpu_featurise(frac, lattice, en, src, dst, img, **kw):
    return featurize.build_bond_and_line_graph(frac.to(DEV), lattice.to(DEV), en.to(DEV), src.to(DEV), dst.to(DEV),
                                               img.to(DEV), *featurize.default_basis(), **kw)


def _close(a, b):
    return torch.allclose(a.cpu().double(), b.double(), rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_gpu_featuriser_matches_reference_golden(tag):
    g = load_golden(f"linegraph_{tag}.pt")
    out = _gpu_featurise(g["frac"], g["lattice"].reshape(1, 3, 3), g["en"], g["bond_src"], g["bond_dst"], g["bond_image"])
    assert torch.equal(out["edge_index"].cpu(), g["edge_index"])
    assert torch.equal(out["lg_edge_index"].cpu(), g["lg_edge_index"])                      # bit-exact, incl. emission order
    assert _close(out["edge_attr"], g["edge_attr"]) and _close(out["lg_edge_attr"], g["lg_edge_attr"])
    # almost every float32 is the identical bit pattern (float64 math, one final rounding)
    same = (out["lg_edge_attr"].cpu() == g["lg_edge_attr"]).float().mean()
    assert float(same) > 0.99


@pytest.mark.gpu
def test_gpu_featuriser_batched_structures_and_larger_cells():
    """Several structures concatenated (global atom ids, per-structure lattice, LOCAL line-graph ids) == each on its own."""
    rc, rg, ac, ag = linegraph_ref.basis()
    specs = [(5, 3, 3.0), (12, 4, 2.9), (1, 9, 4.4), (20, 6, 2.6)]
    fr, la, en, src, dst, img, agr, bptr, want = [], [], [], [], [], [], [], [0], []
    a_off = 0
    for gi, (na, seed, cut) in enumerate(specs):
        f, m, e_, edges = linegraph_ref.random_crystal(na, seed, cutoff=cut)
        want.append(linegraph_ref.build_bond_and_line_graph(f, m, e_, edges, rc, rg, ac, ag))
        fr.append(torch.tensor(f)); la.append(torch.tensor(m)); en.append(torch.tensor(e_, dtype=torch.float64))
        src.append(torch.tensor([e[0] for e in edges]) + a_off); dst.append(torch.tensor([e[1] for e in edges]) + a_off)
        img.append(torch.tensor([list(e[2]) for e in edges], dtype=torch.int32).reshape(-1, 3))
        agr.append(torch.full((na,), gi)); bptr.append(bptr[-1] + len(edges)); a_off += na
    out = _gpu_featurise(torch.cat(fr), torch.stack(la), torch.cat(en), torch.cat(src), torch.cat(dst), torch.cat(img),
                         atom_graph=torch.cat(agr), graph_bond_ptr=torch.tensor(bptr))
    assert torch.equal(out["lg_edge_index"].cpu(), torch.cat([w["lg_edge_index"] for w in want], dim=1))
    assert _close(out["edge_attr"], torch.cat([w["edge_attr"] for w in want]))
    assert _close(out["lg_edge_attr"], torch.cat([w["lg_edge_attr"] for w in want]))
    assert int(out["angle_ptr"][-1]) == sum(w["lg_edge_index"].size(1) for w in want)


@pytest.mark.gpu
def test_gpu_line_graph_of_ring_crystals_matches_synthetic_topology():
    """BASELINE-shaped input: K-regular ring crystals (synthetic.ring_topology restates fetch.py's index rules); the device
    builder must emit the same lg_edge_index, in the same order, for the whole batch at once."""
    atoms, k, n_graphs = 32, 12, 16
    ei, lg = pkg.synthetic.ring_topology(atoms, k)
    half = k // 2
    offs = torch.tensor([d for d in range(1, half + 1)] + [-d for d in range(1, half + 1)])
    img = torch.zeros(ei.size(1), 3, dtype=torch.int32)
    img[:, 0] = offs.repeat(atoms).int()                       # the ring offset stands in for the periodic image
    gid = torch.arange(n_graphs)
    src = (ei[0][None] + gid[:, None] * atoms).reshape(-1)
    dst = (ei[1][None] + gid[:, None] * atoms).reshape(-1)
    gen = torch.Generator().manual_seed(0)
    frac = torch.rand(n_graphs * atoms, 3, generator=gen, dtype=torch.float64)
    lat = (torch.eye(3, dtype=torch.float64) * 5.0).repeat(n_graphs, 1, 1)
    out = _gpu_featurise(frac, lat, torch.ones(n_graphs * atoms, dtype=torch.float64), src, dst, img.repeat(n_graphs, 1),
                         atom_graph=gid.repeat_interleave(atoms), graph_bond_ptr=torch.arange(n_graphs + 1) * ei.size(1))
    want = lg.repeat(1, n_graphs)
    assert torch.equal(out["lg_edge_index"].cpu(), want)
    assert out["lg_edge_attr"].shape == (n_graphs * lg.size(1), 11)
    a = out["lg_edge_attr"][:, 8]
    assert torch.allclose(out["lg_edge_attr"][:, 9], torch.cos(a), atol=1e-6) and float(a.min()) >= 0 and float(a.max()) <= 3.1416


# ---- GPU: device-resident dataset + collate ---------------------------------------------------------------------------------
def _graphs(sizes, k=6, seed=0, dups=False):
    gen = torch.Generator().manual_seed(seed)
    return [pkg.make_crystal(a, k, gen, dups=dups) for a in sizes]


def _assert_same_batch(got, want, ids=None):
    for name in ("x", "edge_index", "edge_attr", "lg_edge_index", "lg_edge_attr", "global_x", "sg_one_hot", "batch", "y"):
        a, b = getattr(got, name).cpu(), getattr(want, name)
        assert a.shape == b.shape and a.dtype == b.dtype, name
        assert torch.equal(a, b), name
    assert got.num_graphs == want.num_graphs and got.lg_active_rows == want.lg_active_rows
    assert got.source_sorted == want.source_sorted
    if ids is not None:
        assert got.train_idx.cpu().tolist() == list(ids)


@pytest.mark.gpu
@pytest.mark.parametrize("lg_inc", ["pyg", "bonds"])
def test_device_collate_is_bit_exact(lg_inc):
    graphs = _graphs([7, 9, 8, 13, 7, 10, 21, 8], dups=True)
    store = dataset.DeviceGraphStore(graphs, DEV, lg_inc=lg_inc)
    for ids in ([0, 1, 2, 3, 4, 5, 6, 7], [6, 2, 2, 5], [3], [7, 0]):
        got = store.collate(ids, validate=True)
        want = pkg.collate([graphs[i] for i in ids], lg_inc=lg_inc)
        _assert_same_batch(got, want, ids)
        assert got.seg_ptr[0].cpu().tolist() == [0] + np.cumsum([graphs[i].x.size(0) for i in ids]).tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("lg_inc", ["pyg", "bonds"])
def test_device_collate_into_shape_bucket_equals_pad_batch(lg_inc):
    graphs = _graphs([7, 9, 8, 13, 11], seed=3)
    store = dataset.DeviceGraphStore(graphs, DEV, lg_inc=lg_inc)
    ids = [4, 0, 3]
    got = store.collate(ids, pad_to_bucket=True, align=64, validate=True)
    want, mask = batching.pad_batch(pkg.collate([graphs[i] for i in ids], lg_inc=lg_inc), align=64)
    _assert_same_batch(got, want)
    assert got.padded and torch.equal(got.loss_mask.cpu(), mask)
    assert got.train_idx.cpu().tolist() == ids + [-1] * (want.num_graphs - len(ids))


@pytest.mark.gpu
def test_device_collate_config2_shape_and_errors():
    """Full BASELINE config-2 batch (256 x 32-atom cells, 1.08 M angles) drawn from a 320-graph store."""
    host = pkg.synthetic_batch(1, 32, 12, seed=0)
    gen = torch.Generator().manual_seed(1)
    graphs = [pkg.make_crystal(32, 12, gen) for _ in range(320)]
    store = dataset.DeviceGraphStore(graphs, DEV)
    ids = torch.randperm(320, generator=gen)[:256].tolist()
    got = store.collate(ids, validate=True)
    want = pkg.collate([graphs[i] for i in ids])
    _assert_same_batch(got, want, ids)
    assert got.sizes == {"B": 256, "N": 8192, "E": 98304, "L": 1081344} and host.x.size(1) == got.x.size(1)
    with pytest.raises(IndexError):
        store.collate([0, 320])
    with pytest.raises(ValueError):
        store.collate([0, 1], shape={"N": 8, "E": 8, "L": 8, "B": 4})
    # device-side guard (ids given on the device only): status bit 0, nothing written
    bad = torch.tensor([0, 999], device=DEV)
    store.collate([0, 1], ids_device=bad)
    assert int(store.status.item()) & 1


@pytest.mark.gpu
def test_standardisation_applied_once_matches_reference_per_sample_rule():
    """train.py:200-217 standardises x[:, :6], x[:, 6:] and global_x per sample on every __getitem__; the store does it once."""
    graphs = _graphs([8, 9], seed=5)
    nd, gd = graphs[0].x.size(1), graphs[0].global_x.numel()
    gen = torch.Generator().manual_seed(9)
    st = {"scalar_mean": torch.randn(6, generator=gen), "scalar_std": torch.rand(6, generator=gen) + 0.5,
          "embed_mean": torch.randn(nd - 6, generator=gen), "embed_std": torch.rand(nd - 6, generator=gen) + 0.5,
          "global_mean": torch.randn(gd, generator=gen), "global_std": torch.rand(gd, generator=gen) + 0.5}
    store = dataset.DeviceGraphStore(graphs, DEV, standardize=st)
    got = store.collate([1, 0])
    xs, gs = [], []
    for i in (1, 0):
        x = graphs[i].x.clone()
        x[:, :6] = (x[:, :6] - st["scalar_mean"]) / st["scalar_std"]
        x[:, 6:] = (x[:, 6:] - st["embed_mean"]) / st["embed_std"]
        xs.append(x)
        gs.append(((graphs[i].global_x.reshape(-1) - st["global_mean"]) / st["global_std"]).reshape(-1, 1))
    assert torch.equal(got.x.cpu(), torch.cat(xs)) and torch.equal(got.global_x.cpu(), torch.cat(gs))


@pytest.mark.gpu
def test_train_step_on_device_collated_bucket_matches_host_batch():
    """A TrainStep fed from the store (bucket-padded on the device) produces the same loss as the host-collated batch."""
    from gnn_elasticity_predictor_b200 import engine
    graphs = _graphs([16, 20, 16, 24], k=8, seed=2)
    store = dataset.DeviceGraphStore(graphs, DEV)
    ids = [2, 0, 3]
    losses = []
    for src in ("store", "host"):
        torch.manual_seed(0)
        m = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(206, 36, 11, 289, 2, 64, 2, 4, 0.0), 2).to(DEV)
        m.train()
        ts = engine.TrainStep(m, graph=False)
        if src == "store":
            b = store.collate(ids, pad_to_bucket=True, align=64)
            tz = pkg.zscore_targets(b.y, b.num_graphs)
            losses.append(float(ts.step(b, tz, mask=b.loss_mask)[0]))
        else:
            b = pkg.collate([graphs[i] for i in ids]).to(DEV)
            tz = pkg.zscore_targets(b.y, b.num_graphs)
            losses.append(float(ts.step(b, tz)[0]))
    assert abs(losses[0] - losses[1]) <= 1e-5 * max(1.0, abs(losses[1]))


# ---- N4: ensemble post-processing ---------------------------------------------------------------------------------------------
def test_oracle_ensemble_post_matches_reference_golden():
    """oracle/ensemble_ref.py against outputs of the reference's own ensemble_collect / conformal_calibration /
    apply_conformal_intervals / LogTransformer (oracle/gen_golden_ensemble.py)."""
    from oracle import ensemble_ref
    g = load_golden("ensemble_post.pt")
    mz, sz = ensemble_ref.moments_batched(g["member_means"], g["member_logvars"], g["batch_sizes"])
    assert torch.equal(mz, g["mean_z"]) and torch.equal(sz, g["std_z"])
    tz = ensemble_ref.to_z(g["targets"], g["log_means"], g["log_stds"])
    for method in ("scaled", "absolute"):
        conf = ensemble_ref.calibration(mz, sz, tz, 0.1, method)
        assert torch.equal(conf["q"], g[method]["q"]) and conf["method"] == g[method]["method"]
        mo, lo, hi = ensemble_ref.intervals(mz, sz, conf["q"], method, g["log_means"], g["log_stds"])
        assert torch.equal(mo, g[method]["mean"]) and torch.equal(lo, g[method]["lower"]) and torch.equal(hi, g[method]["upper"])
        _, lo_z, hi_z = ensemble_ref.intervals(mz, sz, conf["q"], method)
        assert torch.equal(lo_z, g[method]["lower_z"]) and torch.equal(hi_z, g[method]["upper_z"])
    assert bool((g["member_logvars"] < -2.9).any())                      # the floor is exercised


def test_ensemble_post_refuses_cpu():
    from gnn_elasticity_predictor_b200 import ensemble
    with pytest.raises(RuntimeError, match="no CPU"):
        ensemble.ensemble_post(torch.zeros(2, 3, 2), torch.zeros(2, 3, 2))


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["scaled", "absolute"])
def test_gpu_ensemble_post_matches_reference_golden(method):
    from gnn_elasticity_predictor_b200 import ensemble
    g = load_golden("ensemble_post.pt")
    mu, lv = g["member_means"].to(DEV), g["member_logvars"].to(DEV)
    out = ensemble.ensemble_post(mu, lv, q=g[method]["q"], method=method, log_means=g["log_means"], log_stds=g["log_stds"])
    close = lambda a, b: torch.allclose(a.cpu(), b, rtol=2e-6, atol=2e-6)     # noqa: E731
    assert close(out["mean_z"], g["mean_z"]) and close(out["std_z"], g["std_z"])
    assert close(out["mean"], g[method]["mean"]) and close(out["lower"], g[method]["lower"]) and close(out["upper"], g[method]["upper"])
    z = ensemble.ensemble_post(mu, lv, q=g[method]["q"], method=method)                       # z-space intervals
    assert close(z["lower"], g[method]["lower_z"]) and close(z["upper"], g[method]["upper_z"])
    # same moments as the torch composition used by EnsemblePredictor, and the calibration quantile on the device
    mz, vz, sz = ensemble.ensemble_moments(mu, lv)
    assert close(out["mean_z"], mz.cpu()) and close(out["var_z"], vz.cpu())
    tz = ((torch.log(g["targets"]) - g["log_means"]) / g["log_stds"]).to(DEV)
    conf = ensemble.conformal_calibration(out["mean_z"], out["std_z"], tz, 0.1, method)
    assert torch.allclose(conf["q"].cpu(), g[method]["q"], rtol=1e-5, atol=1e-6)
    # homoscedastic members: var = spread of the means only, intervals fall back to "absolute"
    h = ensemble.ensemble_post(mu, None, q=g[method]["q"], method=method)
    want_var = mu.pow(2).mean(0) - mu.mean(0).pow(2)
    assert close(h["var_z"], want_var.cpu()) and close(h["upper"], (mu.mean(0).cpu() + g[method]["q"]))


@pytest.mark.gpu
def test_device_collate_and_featuriser_degenerate_inputs():
    """Graphs without angles / without bonds, an empty selection, and a structure whose atoms have no bonds at all."""
    from gnn_elasticity_predictor_b200.synthetic import CrystalGraph
    gen = torch.Generator().manual_seed(0)
    full = pkg.make_crystal(8, 6, gen)
    lone = CrystalGraph(x=torch.randn(2, 206, generator=gen), edge_index=torch.tensor([[0], [1]]),
                        edge_attr=torch.rand(1, 36, generator=gen), lg_edge_index=torch.zeros(2, 0, dtype=torch.long),
                        lg_edge_attr=torch.zeros(0, 11), global_x=torch.randn(59, 1, generator=gen),
                        sg_one_hot=torch.zeros(230, 1), y=torch.tensor([10.0, 20.0]))
    bare = CrystalGraph(x=torch.randn(1, 206, generator=gen), edge_index=torch.zeros(2, 0, dtype=torch.long),
                        edge_attr=torch.zeros(0, 36), lg_edge_index=torch.zeros(2, 0, dtype=torch.long),
                        lg_edge_attr=torch.zeros(0, 11), global_x=torch.randn(59, 1, generator=gen),
                        sg_one_hot=torch.zeros(230, 1), y=torch.tensor([1.0, 2.0]))
    graphs = [full, lone, bare, full]
    for lg_inc in ("pyg", "bonds"):
        store = dataset.DeviceGraphStore(graphs, DEV, lg_inc=lg_inc)
        for ids in ([1, 2], [2], [0, 2, 1, 3], [2, 2]):
            _assert_same_batch(store.collate(ids, validate=True), pkg.collate([graphs[i] for i in ids], lg_inc=lg_inc), ids)
        got = store.collate([2, 1], pad_to_bucket=True, align=32, validate=True)
        want, _ = batching.pad_batch(pkg.collate([graphs[2], graphs[1]], lg_inc=lg_inc), align=32)
        _assert_same_batch(got, want)
        empty = store.collate([], validate=True)
        assert empty.sizes == {"B": 0, "N": 0, "E": 0, "L": 0}
    # featuriser: no bonds at all; and atoms whose bonds have no continuation (a single directed bond)
    z3 = torch.zeros(2, 3, dtype=torch.float64)
    lat = torch.eye(3, dtype=torch.float64)[None] * 3.0
    none = _gpu_featurise(z3, lat, torch.ones(2, dtype=torch.float64), torch.zeros(0).long(), torch.zeros(0).long(),
                          torch.zeros(0, 3).int())
    assert none["lg_edge_index"].shape == (2, 0) and none["edge_attr"].shape == (0, 36)
    frac = torch.tensor([[0.0, 0, 0], [0.5, 0, 0]], dtype=torch.float64)
    one = _gpu_featurise(frac, lat, torch.ones(2, dtype=torch.float64), torch.tensor([0]), torch.tensor([1]),
                         torch.zeros(1, 3).int())
    assert one["lg_edge_index"].shape == (2, 0) and int(one["angle_ptr"][-1]) == 0
    assert torch.allclose(one["edge_attr"][0, 33:].cpu(), torch.tensor([1.0, 0.0, 0.0]))
    pair = _gpu_featurise(frac, lat, torch.ones(2, dtype=torch.float64), torch.tensor([0, 1]), torch.tensor([1, 0]),
                          torch.zeros(2, 3).int())                       # i->j and its exact reverse: both continuations skipped
    assert pair["lg_edge_index"].shape == (2, 0)


@pytest.mark.gpu
def test_device_collate_into_existing_buffers():
    """collate(out=...) writes into a given batch (e.g. the input buffers of a captured CUDA graph) and refuses selections
    whose plan hints differ from what those buffers were captured with."""
    graphs = _graphs([8, 8, 8, 8, 9], k=6, seed=4)
    store = dataset.DeviceGraphStore(graphs, DEV)
    first = store.collate([0, 1, 2])
    ptrs = {k: v.data_ptr() for k, v in first.tensors().items()}
    again = store.collate([3, 1, 0], out=first)
    assert again is first and {k: v.data_ptr() for k, v in first.tensors().items()} == ptrs
    _assert_same_batch(first, pkg.collate([graphs[i] for i in (3, 1, 0)]), [3, 1, 0])
    with pytest.raises(ValueError):
        store.collate([0, 1, 4], out=first)            # 9-atom graph: other sizes
    with pytest.raises(ValueError):
        store.collate([0, 1], out=first)               # smaller selection would need padding the destination was not made with
