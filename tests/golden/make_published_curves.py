"""Copies the NUMBERS the reference publishes for its accuracy / orthogonality harness into one JSON fixture.

    python tests/golden/make_published_curves.py        (needs /root/reference; run in the build container only)

Sources (outputs of the reference itself, produced by src/bin/stability.rs:196-323 and src/bin/orthogonality.rs:148-232 with
b = StdRng::seed_from_u64(42) uniforms on diagonal spectra, n = 10 000): results/accuracy_{inv,exp}_{well,ill}-conditioned.csv
and results/orthogonality_{inv,exp}_{well,ill}-conditioned.csv.  Nothing under /root/reference is read at test time: the
tests load tests/golden/published_curves.json.
"""
import csv
import json
import os

REF = "/root/reference/results"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "published_curves.json")


def main() -> None:
    doc = {"n": 10000, "seed": 42, "accuracy": {}, "orthogonality": {}}
    for func in ("inv", "exp"):
        for scen in ("well", "ill"):
            name = f"{func}_{scen}-conditioned.csv"
            with open(os.path.join(REF, "accuracy_" + name)) as fh:
                rows = [[int(r["k"]), float(r["relative_error_standard"]), float(r["relative_error_two_pass"]),
                         float(r["relative_solution_deviation"])] for r in csv.DictReader(fh)]
            doc["accuracy"][f"{func}_{scen}"] = {
                "source": "results/accuracy_" + name,
                "columns": ["k", "relative_error_standard", "relative_error_two_pass", "relative_solution_deviation"],
                "rows": rows}
            with open(os.path.join(REF, "orthogonality_" + name)) as fh:
                rows = [[int(r["k"]), float(r["ortho_loss_standard"]), float(r["ortho_loss_regenerated"]),
                         float(r["basis_drift_fro"]), float(r["solution_deviation_l2"])] for r in csv.DictReader(fh)]
            doc["orthogonality"][f"{func}_{scen}"] = {
                "source": "results/orthogonality_" + name,
                "columns": ["k", "ortho_loss_standard", "ortho_loss_regenerated", "basis_drift_fro", "solution_deviation_l2"],
                "rows": rows}
    with open(OUT, "w") as fh:
        json.dump(doc, fh, indent=0)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
