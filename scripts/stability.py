#!/usr/bin/env python
"""The reference's `stability` binary on the GPU engine: same flags and defaults as src/bin/stability.rs:48-83 (accuracy-runner), same CSV schema.

    python scripts/stability.py --function inv --scenario well-conditioned --n 10000 --k-min 5 --k-max 200 --k-step 5 --output out.csv

The published curves (results/*_{inv,exp}_{well,ill}-conditioned.csv) were made with --n 10000."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from two_pass_lanczos_b200 import experiments  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--function", required=True, choices=experiments.FUNCTIONS)
    ap.add_argument("--scenario", required=True, choices=experiments.SCENARIOS)
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--k-min", type=int, default=5)
    ap.add_argument("--k-max", type=int, default=200)
    ap.add_argument("--k-step", type=int, default=5)
    ap.add_argument("--output", required=True)
    a = ap.parse_args()
    rows = experiments.run_accuracy(a.function, a.scenario, a.n, a.k_min, a.k_max, a.k_step)
    experiments.write_csv(a.output, experiments.ACCURACY_COLUMNS, rows)
    print(f"{len(rows)} rows -> {a.output}")


if __name__ == "__main__":
    main()
