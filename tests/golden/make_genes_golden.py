"""Generates tests/golden/genes.npz by running the UNMODIFIED ``gen_model.read_genes`` of the four MOEA/D benchmark
zips (``/root/reference/test/benchmarks/MOEAD/*.zip``, extracted to a temporary directory by
``oracle/ref_harness.RefMoead``).  Run in the build container only:  python tests/golden/make_genes_golden.py

Per family: ``<family>_genes [T, N+E]`` float64 and what the reference returned / left on its model for each gene
vector: ``_point [T,4]`` float32, ``_y [T,N]``, ``_section [T,E]``, ``_d [T,ndof]``, ``_axial [T,E]``, ``_ratio [T,E]``.
Gene vectors: uniform random; heights scaled down so the d_min / pair fixes fire; exact .5 ties of ``round(g*4)``;
all zeros and all ones.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_harness  # noqa: E402

T = 40


def gene_vectors(N, E, seed):
    rng = np.random.RandomState(seed)
    out = []
    for t in range(T):
        g = rng.rand(N + E)
        if t % 5 == 1:
            g[:N] *= 0.05
        if t % 7 == 2:
            g[N:] = np.round(g[N:] * 8) / 8
        if t == 3:
            g[:] = 0
        if t == 4:
            g[:] = 1
        out.append(g)
    return np.array(out, dtype=np.float64)


def main():
    data = {}
    for k, fam in enumerate(ref_harness.MOEAD_ZIPS):
        ref = ref_harness.RefMoead(fam)
        N, E = len(ref.gen.model.nodes), len(ref.gen.model.elements)
        genes = gene_vectors(N, E, 100 + k)
        rec = [ref.read_genes(g) for g in genes]
        data[fam + "_genes"] = genes
        for key in ("point", "y", "section", "d", "axial", "ratio"):
            data["%s_%s" % (fam, key)] = np.array([r[key] for r in rec])
        data[fam + "_int_obj"] = np.array([ref.int_obj1, ref.int_obj2], dtype=np.float32)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "genes.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
