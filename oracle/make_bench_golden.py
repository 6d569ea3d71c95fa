"""fp64 oracle loss of the synthetic batches bench.py times -> tests/golden/bench_loss.json.

TEST / BENCH INFRASTRUCTURE.  bench.py asserts the loss its timed step computes against these values
(1e-5 relative), so a fast run that computes something else cannot pass as a measurement.

    python -m oracle.make_bench_golden

config2_fp32[seed]  16 crops 256x256 fp32, synth.make_case(16, 256, 256, seed) -- rank r uses seed 1 + r
config3_bf16        64 crops 256x256 rounded to bf16, synth.make_case(64, 256, 256, seed=2); L1 sum and row
                    count per image, so any contiguous shard (and the global mean) can be checked
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ssl_oracle as oracle  # noqa: E402
from ssl_b200 import synth  # noqa: E402

KS, KW, SIGMA, EPS = 25, 9, 0.004, 1e-10


def per_image_l1(sr, gt, mask):
    out = []
    for i in range(sr.shape[0]):
        m = mask[i, 0].numpy()
        s = oracle.rows(sr[i].numpy().astype(np.float64), m, KS, KW, SIGMA, True, EPS)
        t = oracle.rows(gt[i].numpy().astype(np.float64), m, KS, KW, SIGMA, True, EPS)
        out.append((float(np.abs(s - t).sum()), int(s.shape[0])))
    return out


def main():
    res = {"ks": KS, "kw": KW, "sigma": SIGMA, "config2_fp32": {}, "config3_bf16": {}}
    for seed in range(1, 9):
        sr, gt, mask = synth.make_case(16, 256, 256, seed=seed, density=0.114)
        per = per_image_l1(sr, gt, mask)
        n = sum(p[1] for p in per)
        res["config2_fp32"][str(seed)] = sum(p[0] for p in per) / (n * KS * KS)
        res.setdefault("config2_rows", {})[str(seed)] = n
        print("config2 seed", seed, res["config2_fp32"][str(seed)], n, flush=True)
    sr, gt, mask = synth.make_case(64, 256, 256, seed=2, density=0.114)
    per = per_image_l1(sr.bfloat16().float(), gt.bfloat16().float(), mask)
    res["config3_bf16"] = {"l1_sum_per_image": [p[0] for p in per], "rows_per_image": [p[1] for p in per],
                           "loss": sum(p[0] for p in per) / (sum(p[1] for p in per) * KS * KS)}
    print("config3", res["config3_bf16"]["loss"], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "bench_loss.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
