"""Golden vectors for the bond / line-graph featuriser, produced by the reference's OWN ``build_graph_from_structure``
(``/root/reference/scripts/fetch.py:318-612``, executed from where it lies, unmodified).

pymatgen is not installed in this image, so the *objects* the function receives are duck-typed stand-ins (a structure with
``frac_coords`` / ``lattice.get_cartesian_coords`` / ``sites``; an element-property table; a fixed neighbour list); every
number in the stored outputs -- RBF bond features, direction vectors, line-graph indices, angle features -- is computed by
the reference's code.  Run from the repo root in the build container:

    python oracle/gen_golden_linegraph.py        # writes tests/golden/linegraph_*.pt

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import linegraph_ref  # noqa: E402


def load_fetch_module():
    oracle.install_shim()
    for name in ("pymatgen", "pymatgen.analysis", "pymatgen.analysis.structure_matcher", "pymatgen.core",
                 "pymatgen.symmetry", "pymatgen.symmetry.analyzer", "pymatgen.analysis.local_env"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pymatgen.analysis.structure_matcher"].StructureMatcher = lambda *a, **k: object()
    spec = importlib.util.spec_from_file_location("_reference_scripts_fetch",
                                                  os.path.join(oracle.REFERENCE_ROOT, "scripts", "fetch.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


class _Lattice:
    def __init__(self, m):
        self.matrix = np.asarray(m, dtype=float)
        self.a, self.b, self.c = (float(np.linalg.norm(self.matrix[i])) for i in range(3))

    def get_cartesian_coords(self, frac):
        return np.dot(np.asarray(frac, dtype=float), self.matrix)      # pymatgen Lattice.get_cartesian_coords


class _Specie:
    def __init__(self, k):
        self.symbol = f"X{k}"


class _Site:
    def __init__(self, k):
        self.specie = _Specie(k)


class _Structure:
    def __init__(self, frac, lattice):
        self.frac_coords = np.asarray(frac, dtype=float)
        self.lattice = _Lattice(lattice)
        self.sites = [_Site(k) for k in range(len(frac))]
        self.composition = None

    def __len__(self):
        return len(self.frac_coords)

    def copy(self):
        return self

    def add_oxidation_state_by_guess(self):
        raise RuntimeError("not available")


class _Doc:
    def __init__(self, structure):
        self.structure = structure
        self.formula_pretty = "X"
        self.material_id = "golden"
        self.k_vrh, self.g_vrh = 100.0, 50.0


def main():
    fetch = load_fetch_module()
    out_dir = os.path.join(ROOT, "tests", "golden")
    rbf_centers, rbf_gamma, angle_centers, angle_gamma = linegraph_ref.basis()
    cases = [("a", 3, 11, 3.0), ("b", 6, 5, 3.2), ("c", 1, 7, 4.6), ("d", 9, 23, 2.6)]
    for tag, n_atoms, seed, cutoff in cases:
        frac, lattice, en, edges = linegraph_ref.random_crystal(n_atoms, seed, cutoff=cutoff)
        structure = _Structure(frac, lattice)
        fetch._composition_and_prototype = lambda s: ("", "")
        fetch._element_props = lambda sym: (1, 1, 1, float(en[int(sym[1:])]), 1.0, 1.0)
        fetch._neighbors_edges = lambda s, nn_method, cutoff, fallback_cutoff=7.5: (list(edges), "cutoff")
        fetch._spacegroup_and_density = lambda s: (1, 1.0, (1.0, 1.0, 1.0, 90.0, 90.0, 90.0))
        fetch._metric_tensor_and_globals = lambda s, sg, dens: ([0.0] * 8, np.zeros(230))
        ge = fetch.build_graph_from_structure(_Doc(structure), "cutoff", cutoff, rbf_centers, rbf_gamma, angle_centers,
                                              angle_gamma, guess_oxidation=False)
        data = fetch.to_pyg_data(ge)
        blob = {
            "frac": torch.tensor(frac, dtype=torch.float64), "lattice": torch.tensor(lattice, dtype=torch.float64),
            "en": torch.tensor(en, dtype=torch.float64),
            "bond_src": torch.tensor([e[0] for e in edges], dtype=torch.long),
            "bond_dst": torch.tensor([e[1] for e in edges], dtype=torch.long),
            "bond_image": torch.tensor([list(e[2]) for e in edges], dtype=torch.int32).reshape(-1, 3),
            "cutoff": cutoff,
            "edge_index": data.edge_index.reshape(2, -1), "edge_attr": data.edge_attr.reshape(-1, 36),
            "lg_edge_index": data.lg_edge_index.reshape(2, -1), "lg_edge_attr": data.lg_edge_attr.reshape(-1, 11),
        }
        mine = linegraph_ref.build_bond_and_line_graph(frac, lattice, en, edges, rbf_centers, rbf_gamma, angle_centers,
                                                       angle_gamma)
        for k in ("edge_index", "edge_attr", "lg_edge_index", "lg_edge_attr"):
            assert torch.equal(mine[k], blob[k]), f"restatement differs from the reference on {k} (case {tag})"
        path = os.path.join(out_dir, f"linegraph_{tag}.pt")
        torch.save(blob, path)
        print(f"{path}: atoms {n_atoms}, bonds {len(edges)}, angles {blob['lg_edge_index'].size(1)}, "
              f"{os.path.getsize(path) / 1024:.0f} KB; restatement bit-identical")


if __name__ == "__main__":
    main()
