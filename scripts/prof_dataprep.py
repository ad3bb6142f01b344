"""CUDA-event timings of the data-side kernels (device collate, bond features + line graph) at BASELINE config-2 size.
usage: python scripts/prof_dataprep.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import dataset, featurize

dev = torch.device("cuda", 0)
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6650.0


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


gen = torch.Generator().manual_seed(1)
graphs = [pkg.make_crystal(32, 12, gen) for _ in range(512)]
store = dataset.DeviceGraphStore(graphs, dev)
ids = torch.randperm(512, generator=gen)[:256].tolist()
ids_dev = torch.tensor(ids, device=dev)
b = store.collate(ids)
nbytes = b.nbytes()
ms = timeit(lambda: store.collate(ids, ids_device=ids_dev))
print(f"device collate   config2 (256 graphs, N={b.sizes['N']}, E={b.sizes['E']}, L={b.sizes['L']}): {ms * 1e3:8.1f} us  "
      f"batch {nbytes / 1e6:.1f} MB -> read+write {2 * nbytes / ms / 1e6:7.0f} GB/s ({2 * nbytes / ms / 1e6 / PEAK:.2f} of measured HBM peak {PEAK:.0f}); "
      f"store {store.nbytes() / 1e6:.0f} MB resident")
ms_p = timeit(lambda: store.collate(ids, ids_device=ids_dev, pad_to_bucket=True))
print(f"device collate   into shape bucket: {ms_p * 1e3:8.1f} us")
host = pkg.collate([graphs[i] for i in ids]).pin_memory()
ms_h = timeit(lambda: host.to(dev, non_blocking=True))
print(f"host batch H2D   (pinned, {nbytes / 1e6:.1f} MB): {ms_h * 1e3:8.1f} us  ({nbytes / ms_h / 1e6:.1f} GB/s PCIe)")

# featuriser: 256 ring crystals x 32 atoms x 12 bonds, random geometry
atoms, k, ng = 32, 12, 256
ei, lg = pkg.synthetic.ring_topology(atoms, k)
half = k // 2
offs = torch.tensor([d for d in range(1, half + 1)] + [-d for d in range(1, half + 1)])
img = torch.zeros(ei.size(1), 3, dtype=torch.int32); img[:, 0] = offs.repeat(atoms).int()
gid = torch.arange(ng)
src = (ei[0][None] + gid[:, None] * atoms).reshape(-1).to(dev); dst = (ei[1][None] + gid[:, None] * atoms).reshape(-1).to(dev)
frac = torch.rand(ng * atoms, 3, generator=gen, dtype=torch.float64).to(dev)
lat = (torch.eye(3, dtype=torch.float64) * 5.0).repeat(ng, 1, 1).to(dev)
en = torch.ones(ng * atoms, dtype=torch.float64, device=dev)
imgs = img.repeat(ng, 1).to(dev); ag = gid.repeat_interleave(atoms).to(dev); bp = (torch.arange(ng + 1) * ei.size(1)).to(dev)
basis = featurize.default_basis()
f = lambda: featurize.build_bond_and_line_graph(frac, lat, en, src, dst, imgs, *basis, atom_graph=ag, graph_bond_ptr=bp)
out = f()
ms_f = timeit(f)
L = out["lg_edge_index"].size(1)
print(f"bond features + line graph  (E={src.numel()}, L={L}): {ms_f * 1e3:8.1f} us incl. one host sync  -> "
      f"{L / ms_f / 1e3:.0f} M angles/s, output {(L * (16 + 44) + src.numel() * 144) / 1e6:.0f} MB")
