"""times tfem_read_genes (one launch = one MOEA/D population) with CUDA events; writes one JSON line per family"""
import json
import sys

import torch

sys.path.insert(0, ".")
from mop_truss_marl_b200 import genes  # noqa: E402

POP = {"small_bridge": 18000, "small_roof": 12800, "large_bridge": 12000, "large_roof": 24000}   # (n_neighbors+1) * 20
for fam, B in POP.items():
    ev = genes.GeneEvaluator(fam)
    g = torch.rand(B, ev.N + ev.E, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ev.read_genes(g)
    ts = []
    for i in range(20):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ev.read_genes(g); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sum(ts) / len(ts)
    bytes_per = 8 * (ev.N + ev.E) + 16 + 4        # genes in, point + status out
    print(json.dumps({"family": fam, "individuals": B, "ms_per_launch": ms, "evaluations_per_s": B / ms * 1e3,
                      "algorithmic_GBps": B * bytes_per / ms / 1e6, "l2": "flushed between launches"}))
