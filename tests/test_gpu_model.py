"""GPU parity of the drop-in modules against the golden vectors (made by the reference's own classes) and
against the oracle on fresh batches, fp32 and bf16 autocast; end-to-end properties at BASELINE sizes."""
import glob
import os

import pytest
import torch

from conftest import GOLDEN_DIR, Bag, check_per_tensor, grad_err, grads_gmax, load_golden, reference_amp_grads, rel_err
from oracle import model_ref
import gnn_elasticity_predictor_b200 as pkg

pytestmark = pytest.mark.gpu
DEV = "cuda"
MODEL_GOLDENS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN_DIR, "model_*.pt")))
BLOCK_GOLDENS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN_DIR, "blocks_*.pt")))


@pytest.mark.parametrize("name", MODEL_GOLDENS)
def test_model_matches_golden_fp32(name):
    g = load_golden(name)
    ctor = g["ctor"]
    model = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(**ctor), ctor["target_dim"]).to(DEV)
    model.load_state_dict(g["state_dict"], strict=True)
    model.train()
    batch = Bag(g["batch"], g["num_graphs"]).to(DEV)
    mean, logvar = model(batch)
    assert rel_err(mean.cpu(), g["mean"]) < 1e-5 and rel_err(logvar.cpu(), g["logvar"]) < 1e-5
    loss = pkg.gaussian_nll_loss(mean, logvar, pkg.zscore_targets(batch.y, batch.num_graphs))
    assert rel_err(loss.cpu(), g["loss"]) < 1e-5
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(grads) == set(g["grads"])
    gmax = grads_gmax(g["grads"])
    for k, want in g["grads"].items():
        assert grad_err(grads[k], want, gmax) < 1e-4, k
    assert rel_err(model.embed(batch).cpu(), g["embed"]) < 1e-5
    assert rel_err(model.base(batch).cpu(), g["plain_output"]) < 1e-5


@pytest.mark.parametrize("name", BLOCK_GOLDENS)
def test_blocks_match_golden_fp32(name):
    g = load_golden(name)
    hidden, heads = g["hidden"], g["heads"]
    for tag, blk in (("edge_block", pkg.EdgeUpdateBlock(hidden, heads, 0.0)),
                     ("node_block", pkg.NodeUpdateBlock(hidden, hidden, heads, 0.0))):
        blk = blk.to(DEV)
        blk.load_state_dict(g[tag]["state_dict"], strict=True)
        x = g["x"].to(DEV).requires_grad_(True)
        ea = g["edge_attr"].to(DEV).requires_grad_(True)
        y = blk(x, g["index"].to(DEV), ea)         # reference call signature, plan built internally
        y.backward(g["gout"].to(DEV))
        assert rel_err(y.cpu(), g[tag]["y"]) < 1e-5, tag
        assert rel_err(x.grad.cpu(), g[tag]["dx"]) < 1e-4, tag
        assert rel_err(ea.grad.cpu(), g[tag]["dedge"]) < 1e-4, tag
        gmax = grads_gmax(g[tag]["grads"])
        for k, p in blk.named_parameters():
            assert grad_err(p.grad, g[tag]["grads"][k], gmax) < 1e-4, (tag, k)


def test_transformer_conv_pyg_signature_matches_oracle():
    import oracle
    oracle.install_shim()
    from torch_geometric.nn import TransformerConv as RefConv
    torch.manual_seed(0)
    ref = RefConv(64, 16, heads=4, edge_dim=64, beta=True)
    ours = pkg.TransformerConv(64, 16, heads=4, edge_dim=64, beta=True).to(DEV)
    ours.load_state_dict(ref.state_dict(), strict=True)
    x, ea = torch.randn(30, 64), torch.randn(200, 64)
    index = torch.randint(0, 30, (2, 200))
    assert rel_err(ours(x.to(DEV), index.to(DEV), ea.to(DEV)).cpu(), ref(x, index, ea)) < 1e-5


def _pair(hidden, layers, heads, seed=42):
    ref = model_ref.build_hetero(hidden=hidden, layers=layers, heads=heads, seed=seed)
    ours = pkg.HeteroAlignnRegressor(
        pkg.AlignnRegressor(206, 36, 11, 289, 2, hidden, layers, heads, 0.0), 2).to(DEV)
    ours.load_state_dict(ref.state_dict(), strict=True)
    return ref, ours


def _as_double(batch):
    import copy
    b = copy.copy(batch)
    for k in ("x", "edge_attr", "lg_edge_attr", "global_x", "sg_one_hot"):
        setattr(b, k, getattr(batch, k).double())
    return b


def _loss_and_grads(model, batch, autocast=False):
    model.zero_grad()
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else torch.autocast("cuda", enabled=False)
    with ctx:
        mean, logvar = model(batch)
        loss = pkg.gaussian_nll_loss(mean.float(), logvar.float(), pkg.zscore_targets(batch.y, batch.num_graphs))
    loss.backward()
    return mean, logvar, loss, {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("lg_inc", ["pyg", "bonds"])
def test_config1_default_arch_fp32_vs_oracle(lg_inc):
    """BASELINE config 1: 64 crystals x 16 atoms x 12 neighbours, default arch (H=256, 4+4 layers, 4 heads)."""
    ref, ours = _pair(256, 4, 4)
    batch = pkg.synthetic_batch(64, 16, 12, seed=0, lg_inc=lg_inc)
    assert batch.sizes == {"B": 64, "N": 1024, "E": 12288, "L": 135168}
    # target = the oracle evaluated in fp64 (the fp32 CPU oracle itself is only good to ~2e-4 on some
    # gradients at this size: see profiles/r01_parity_report.txt)
    ref = ref.double()
    b64 = _as_double(batch)
    r_mean, r_logvar = ref(b64)
    r_loss = model_ref.gaussian_nll_loss(r_mean, r_logvar, pkg.zscore_targets(batch.y, 64).double())
    r_loss.backward()
    mean, logvar, loss, grads = _loss_and_grads(ours, batch.to(DEV))
    assert rel_err(mean, r_mean) < 1e-5 and rel_err(logvar, r_logvar) < 1e-5
    assert rel_err(loss, r_loss) < 1e-5
    want = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
    gmax = grads_gmax(want)
    for k, w in want.items():
        assert grad_err(grads[k], w, gmax) < 1e-4, k


# Tensors whose TRUE gradient is exactly zero: the segment softmax is invariant to adding one vector to every key of a
# row, so d loss / d lin_key.bias == 0 and every implementation (the fp32 CPU oracle vs the fp64 oracle included) produces
# rounding noise only.  They are held to an absolute bound relative to the largest gradient of the model instead.
ZERO_TRUE_GRADIENT = {"conv.lin_key.bias": (("abs", 2e-3), "true gradient is exactly zero (softmax shift invariance)")}


@pytest.mark.parametrize("seed", [1, 2])
def test_config1_default_arch_bf16_autocast_vs_oracle(seed):
    """bf16 autocast regime (reference ``train.py:632-636``).  Target = the oracle in fp64.  Tolerance: rel 2e-2 on
    outputs and loss.  Gradients: |err| <= 2e-2 of the gradient scale, OR no worse than 1.5x the error the
    reference's own bf16-autocast run (oracle on CPU under ``torch.autocast('cpu', bfloat16)``) makes on that
    tensor against the same fp64 target -- at this size the reference's AMP itself misses 2e-2 on
    ``feat_proj.0.weight`` (3.0e-2 / 3.8e-2 of the gradient scale for seeds 1 / 2); plus direction (cosine)."""
    import copy
    ref, ours = _pair(256, 4, 4)
    batch = pkg.synthetic_batch(64, 16, 12, seed=seed, lg_inc="pyg")
    tz = pkg.zscore_targets(batch.y, 64)
    ref64 = copy.deepcopy(ref).double()
    r_mean, r_logvar = ref64(_as_double(batch))
    r_loss = model_ref.gaussian_nll_loss(r_mean, r_logvar, tz.double())
    r_loss.backward()
    want = {k: p.grad for k, p in ref64.named_parameters() if p.grad is not None}
    gmax = grads_gmax(want)
    # the reference's own AMP run on the same batch (CPU bf16 autocast of the oracle): the per-tensor yardstick
    amp = reference_amp_grads(ref, batch, tz, model_ref.gaussian_nll_loss)
    mean, logvar, loss, grads = _loss_and_grads(ours, batch.to(DEV), autocast=True)
    assert rel_err(mean, r_mean) < 2e-2 and rel_err(logvar, r_logvar) < 2e-2
    assert rel_err(loss, r_loss) < 2e-2
    # per tensor, relative to the tensor's OWN scale: 2e-2, else -- by name, printed -- no worse than 1.5x what the
    # reference's own AMP run measures on that tensor against the same fp64 target
    rep = check_per_tensor(grads, want, 2e-2, ZERO_TRUE_GRADIENT, label=f"config1_bf16_seed{seed}", amp=amp)
    for k, (rel, share) in rep.items():
        if rel > 2e-2 and share > 1e-3 and not k.endswith("conv.lin_key.bias"):
            cos = float(torch.nn.functional.cosine_similarity(grads[k].double().flatten(), want[k].cpu().flatten(), dim=0))
            assert cos > 0.95, (k, cos)


def test_smoke_arch_eval_mode_and_no_grad():
    g = load_golden("model_smoke_arch.pt")
    model = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(**g["ctor"]), 2).to(DEV)
    model.load_state_dict(g["state_dict"])
    model.eval()
    with torch.no_grad():
        mean, logvar = model(Bag(g["batch"], g["num_graphs"]).to(DEV))
    assert rel_err(mean.cpu(), g["mean"]) < 1e-5 and rel_err(logvar.cpu(), g["logvar"]) < 1e-5


def test_empty_inputs_follow_reference_guards():
    """train.py:313-314,331-332,548-556: empty line graph / missing attributes fall through."""
    ref, ours = _pair(32, 1, 4)
    batch = pkg.synthetic_batch(3, 6, 4, seed=2)
    batch.lg_edge_index = torch.zeros(2, 0, dtype=torch.long)
    batch.lg_edge_attr = torch.zeros(0, 11)
    r_mean, r_logvar = ref(batch)
    mean, logvar = ours(batch.to(DEV))
    assert rel_err(mean.cpu(), r_mean) < 1e-5 and rel_err(logvar.cpu(), r_logvar) < 1e-5


def test_dropout_training_mode_runs_and_is_seed_reproducible():
    ours = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(206, 36, 11, 289, 2, 64, 2, 4, 0.15), 2).to(DEV)
    ours.train()
    batch = pkg.synthetic_batch(8, 16, 12, seed=3).to(DEV)
    torch.manual_seed(5)
    a = ours(batch)[0]
    torch.manual_seed(5)
    b = ours(batch)[0]
    torch.manual_seed(6)
    c = ours(batch)[0]
    assert torch.equal(a, b) and not torch.equal(a, c)
    ours.eval()
    assert torch.equal(ours(batch)[0], ours(batch)[0])


def test_h256_golden_made_by_the_reference_classes_bf16_tensor_core_path():
    """The H=256 / 4-head fixture made by the reference's own classes (oracle/gen_golden.py), run through the bf16
    regime = the tensor-core kernels the benchmark times (the fp32 parametrisation above runs the CUDA-core family)."""
    g = load_golden("model_default_arch_h256.pt")
    ctor = g["ctor"]
    model = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(**ctor), ctor["target_dim"]).to(DEV)
    model.load_state_dict(g["state_dict"], strict=True)
    model.train()
    batch = Bag(g["batch"], g["num_graphs"]).to(DEV)
    mean, logvar, loss, grads = _loss_and_grads(model, batch, autocast=True)
    assert rel_err(mean, g["mean"]) < 2e-2 and rel_err(logvar, g["logvar"]) < 2e-2
    assert rel_err(loss, g["loss"]) < 2e-2
    # yardstick for exceptions: the reference's own bf16 autocast on the same fixture (the oracle restatement is
    # bit-identical to the reference classes that made it: tests/test_oracle_golden.py)
    ref = model_ref.HeteroAlignnRegressor(model_ref.AlignnRegressor(**ctor), ctor["target_dim"])
    ref.load_state_dict(g["state_dict"], strict=True)
    cpu_batch = Bag(g["batch"], g["num_graphs"])
    amp = reference_amp_grads(ref, cpu_batch, pkg.zscore_targets(cpu_batch.y, cpu_batch.num_graphs),
                              model_ref.gaussian_nll_loss)
    check_per_tensor(grads, g["grads"], 2e-2, ZERO_TRUE_GRADIENT, label="golden_h256_bf16", amp=amp)


@pytest.mark.parametrize("lg_inc", ["pyg", "bonds"])
def test_config2_bf16_forward_loss_and_every_gradient_vs_fp64_oracle(lg_inc):
    """BASELINE config 2 (256 x 32-atom cells x 12 neighbours: N=8 192, E=98 304, L=1 081 344; H=256, 4+4 layers, 4 heads)
    -- the exact workload bench.py times -- forward, loss and every parameter gradient against the oracle in fp64."""
    ref, ours = _pair(256, 4, 4)
    batch = pkg.synthetic_batch(256, 32, 12, seed=0, lg_inc=lg_inc)
    assert batch.sizes == {"B": 256, "N": 8192, "E": 98304, "L": 1081344}
    tz = pkg.zscore_targets(batch.y, 256)
    torch.set_num_threads(max(torch.get_num_threads(), (os.cpu_count() or 1)))
    ref = ref.double()
    r_mean, r_logvar = ref(_as_double(batch))
    r_loss = model_ref.gaussian_nll_loss(r_mean, r_logvar, tz.double())
    r_loss.backward()
    want = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
    mean, logvar, loss, grads = _loss_and_grads(ours, batch.to(DEV), autocast=True)
    assert rel_err(mean, r_mean) < 2e-2 and rel_err(logvar, r_logvar) < 2e-2
    assert rel_err(loss, r_loss) < 2e-2
    # exceptions to 2e-2 per tensor: by name, printed, and no worse than 1.5x the reference's own bf16 autocast (oracle on
    # the host cores under torch.autocast('cpu', bfloat16)) on the same batch against the same fp64 target
    amp = reference_amp_grads(ref.float(), batch, tz, model_ref.gaussian_nll_loss)
    check_per_tensor(grads, want, 2e-2, ZERO_TRUE_GRADIENT, label=f"config2_bf16_{lg_inc}", amp=amp)
    # determinism of the hand-written path at this size (atomics-free): two runs agree bit for bit on the loss
    _, _, loss2, g2 = _loss_and_grads(ours, batch.to(DEV), autocast=True)
    assert torch.equal(loss, loss2)
    for k in ("base.edge_blocks.0.conv.lin_edge.weight", "base.node_encoder.0.weight"):
        assert rel_err(grads[k], g2[k]) < 1e-5, k     # cuBLAS split-K may reorder; ours may not


# ---- streaming path (hidden = 256): blocks against the fp64 oracle, and against the materialised path ---------
def _multigraph(n, e, seed, hub=0):
    g = torch.Generator().manual_seed(seed)
    dst = torch.randint(0, max(1, (3 * n) // 4), (e,), generator=g)      # last quarter of the rows stays empty
    src = torch.randint(0, n, (e,), generator=g)
    src[:3] = dst[:3]                                                     # self loops
    src[3:6], dst[3:6] = src[0:3].clone(), dst[0:3].clone()              # duplicate edges
    if hub:
        dst[-hub:] = 1                                                    # one long row
    return torch.stack([src, dst])


@pytest.mark.parametrize("heads", [1, 2, 4])
@pytest.mark.parametrize("kind", ["edge_block", "node_block"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_streaming_blocks_h256_vs_fp64_oracle(heads, kind, dtype):
    hidden, n, e = 256, 97, 1500
    torch.manual_seed(heads)
    if kind == "edge_block":
        ref = model_ref.EdgeUpdateBlock(hidden, heads, 0.0).double()
        ours = pkg.EdgeUpdateBlock(hidden, heads, 0.0).to(DEV)
    else:
        ref = model_ref.NodeUpdateBlock(hidden, hidden, heads, 0.0).double()
        ours = pkg.NodeUpdateBlock(hidden, hidden, heads, 0.0).to(DEV)
    with torch.no_grad():
        ref.norm.weight.uniform_(0.5, 1.5); ref.norm.bias.uniform_(-0.3, 0.3)
    ours.load_state_dict({k: v.float() for k, v in ref.state_dict().items()}, strict=True)
    ours.streaming = True
    index = _multigraph(n, e, 7 + heads, hub=600)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(n, hidden, generator=g)
    ea = torch.randn(e, hidden, generator=g)
    gout = torch.randn(n, hidden, generator=g)
    xr, er = x.double().requires_grad_(True), ea.double().requires_grad_(True)
    yr = ref(xr, index, er)
    yr.backward(gout.double())
    xo, eo = x.to(DEV).requires_grad_(True), ea.to(DEV).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
        yo = ours(xo, index.to(DEV), eo)
    yo.backward(gout.to(DEV))
    want = {k: p.grad for k, p in ref.named_parameters()}
    gmax = grads_gmax(want)
    if dtype == torch.float32:
        assert rel_err(yo, yr) < 1e-5
        assert rel_err(xo.grad, xr.grad) < 1e-4 and rel_err(eo.grad, er.grad) < 1e-4
        for k, p in ours.named_parameters():
            assert grad_err(p.grad, want[k], gmax) < 1e-4, k
        return
    # bf16: rel 2e-2, or no worse than 1.5x the error of the reference's OWN bf16-autocast run (oracle block on
    # CPU under torch.autocast('cpu', bfloat16)) against the same fp64 target
    import copy
    amp = copy.deepcopy(ref).float()
    amp.zero_grad()
    xa, ea_a = x.clone().requires_grad_(True), ea.clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ya = amp(xa, index, ea_a)
    ya.backward(gout)
    tol = lambda got, w: max(2e-2, 1.5 * rel_err(got, w))   # noqa: E731
    assert rel_err(yo, yr) < tol(ya, yr)
    assert rel_err(xo.grad, xr.grad) < tol(xa.grad, xr.grad), (rel_err(xo.grad, xr.grad), rel_err(xa.grad, xr.grad))
    assert rel_err(eo.grad, er.grad) < tol(ea_a.grad, er.grad), (rel_err(eo.grad, er.grad), rel_err(ea_a.grad, er.grad))
    amp_grads = {k: p.grad for k, p in amp.named_parameters()}
    for k, p in ours.named_parameters():
        err = float((p.grad.double().cpu() - want[k]).abs().max())
        amp_err = float((amp_grads[k].double() - want[k]).abs().max())
        # parameter gradients sum ~1e5 bf16-rounded terms: same order as the reference's own AMP error (factor 3)
        assert err < max(2e-2 * gmax, 3.0 * amp_err), (k, err / gmax, amp_err / gmax)


def test_streaming_and_materialised_paths_agree_on_the_model():
    torch.manual_seed(3)
    a = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(206, 36, 11, 289, 2, 256, 2, 4, 0.0), 2).to(DEV)
    batch = pkg.synthetic_batch(6, 10, 6, seed=4, lg_inc="bonds", dups=True).to(DEV)
    outs = {}
    for flag in (True, False):
        a.base.streaming = flag
        for blk in list(a.base.edge_blocks) + list(a.base.node_blocks):
            blk.streaming = flag
        _, _, loss, grads = _loss_and_grads(a, batch)
        outs[flag] = (loss.detach().clone(), grads)
    assert rel_err(outs[True][0], outs[False][0]) < 1e-5
    gmax = grads_gmax(outs[False][1])
    for k, w in outs[False][1].items():
        assert grad_err(outs[True][1][k], w, gmax) < 1e-4, k


def test_streaming_dropout_statistics():
    """attention dropout inside the streaming kernels: unbiased in expectation, reproducible per key"""
    torch.manual_seed(0)
    blk = pkg.EdgeUpdateBlock(256, 4, 0.3).to(DEV)
    blk.streaming = True
    n, e = 64, 6000
    index = _multigraph(n, e, 5).to(DEV)
    x, ea = torch.randn(n, 256, device=DEV), torch.randn(e, 256, device=DEV)
    blk.eval()
    y_eval = blk(x, index, ea)
    blk.train()
    torch.manual_seed(1); y1 = blk(x, index, ea)
    torch.manual_seed(1); y2 = blk(x, index, ea)
    torch.manual_seed(2); y3 = blk(x, index, ea)
    assert torch.equal(y1, y2) and not torch.equal(y1, y3) and not torch.equal(y1, y_eval)
    ys = []
    for s in range(24):
        torch.manual_seed(100 + s)
        ys.append(blk(x, index, ea))
    # block output = x + dropout(relu(LN(.))): its mean over many masks stays close to the eval output scale
    mean_train = torch.stack(ys).mean(0)
    assert float((mean_train - x).mean()) == pytest.approx(float((y_eval - x).mean()), rel=0.2)
