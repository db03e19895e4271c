"""FEM env-step kernel alone (CUDA events, L2 flushed between steps), for A/B builds: TFEM_LIB=<lib> python scripts/fem_time.py
[family batch]...  Development helper; prints one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    cases = sys.argv[1:] or ["small_bridge", "4096"]
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for fam, B in zip(cases[0::2], cases[1::2]):
        r = bench.fem_leg(fam, int(B), 0, 1, dev, flush, "fem", steps=100, warmup=10)
        print(json.dumps({"lib": os.environ.get("TFEM_LIB", ""), "family": fam, "B": int(B), "ms": r["ms_per_step"],
                          "frac": r["roofline_fem"]["frac"], "bad": r["status_nonzero_envs"]}), flush=True)


if __name__ == "__main__":
    main()
