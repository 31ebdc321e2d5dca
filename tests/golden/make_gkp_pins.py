"""Pins for the GKP finite-squeezing noise channel, produced from
the REFERENCE's own notebook and data.  Run in the build container only (needs /root/reference):

    python tests/golden/make_gkp_pins.py          # writes tests/golden/gkp_pins.json

1. The analytic gate-error model: the functions ``db2eps``, ``analytical_gate_error``,
   ``gate_error_I`` and ``gate_error_P`` are cut out of ``plot_data.ipynb`` (cell with
   "Grover error estimate") and EXECUTED as they stand; their values at the squeezing levels of
   the data are stored.
2. The randomised-benchmarking data ``data/gkp_rb.dat`` (22 060 samples from the reference's CV
   simulation) is refitted exactly as the notebook does (``process_rb_samples`` and
   ``fidelity_analysis``, also executed from the notebook source): error rate r(dB) with its
   standard error, and the notebook's own residual (r_fit - analytic) / r_err.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PAPER = "/root/reference/impact_of_finite_squeezing_on_near-term_quantum_computations_using_gkp_qubits"


def notebook_functions(names):
    """Source of the top-level functions `names`, taken verbatim from the notebook."""
    nb = json.load(open(os.path.join(PAPER, "plot_data.ipynb")))
    text = "\n".join("".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code")
    out = []
    for name in names:
        m = re.search(rf"^def {name}\(.*?(?=^\S)", text, re.S | re.M)
        assert m, name
        out.append(m.group(0))
    return "\n".join(out)


def main():
    from scipy.optimize import curve_fit
    from scipy.special import erf
    ns = {"np": np, "erf": erf, "curve_fit": curve_fit, "print": lambda *a, **k: None}
    exec(notebook_functions(["db2eps", "analytical_gate_error", "gate_error_I", "gate_error_P",
                             "process_rb_samples", "fidelity_analysis"]), ns)
    samples = json.load(open(os.path.join(PAPER, "data", "gkp_rb.dat")))
    dbs = sorted({round(s["db"], 6) for s in samples})
    rows = []
    for db in dbs:
        sub = [s for s in samples if abs(s["db"] - db) < 1e-5]
        depths = sorted({s["depth"] for s in sub})
        row = {"db": db, "samples": len(sub), "depths": depths,
               "gate_error_I": float(ns["gate_error_I"](db)), "gate_error_P": float(ns["gate_error_P"](db)),
               "flip_k2": float(ns["analytical_gate_error"](db, 2)), "flip_k3": float(ns["analytical_gate_error"](db, 3)),
               "eps": float(ns["db2eps"](db))}
        row["analytic_mean"] = 0.5 * (row["gate_error_I"] + row["gate_error_P"])
        try:
            fid, _pur = ns["process_rb_samples"](sub)
            res = ns["fidelity_analysis"](fid)
            row.update({"r_fit": float(res["r"]), "r_err": float(res["r_err"]), "p_fit": float(res["p"]),
                        "mean_fidelity": [float(v) for v in fid["means"]],
                        "fidelity_sem": [float(v) for v in fid["errors"]],
                        "residual_sigma": float((res["r"] - row["analytic_mean"]) / res["r_err"])})
        except Exception as exc:                                  # a level the notebook's fit cannot handle
            row["fit_error"] = repr(exc)
        rows.append(row)
        print(row["db"], row["samples"], row.get("r_fit"), row.get("r_err"), row["analytic_mean"], row.get("residual_sigma"))
    out = {"source": "plot_data.ipynb functions executed verbatim + data/gkp_rb.dat refitted as the notebook does",
           "levels": rows}
    with open(os.path.join(HERE, "gkp_pins.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote gkp_pins.json")


if __name__ == "__main__":
    sys.exit(main())
