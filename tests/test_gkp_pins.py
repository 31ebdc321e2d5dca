"""Row a13 pins: the GKP finite-squeezing Pauli channel of this package
against (1) the analytic gate-error values of the reference's notebook, executed verbatim by
tests/golden/make_gkp_pins.py, and (2) the reference's randomised-benchmarking DATA
(PAPER/data/gkp_rb.dat, 22 060 samples of its CV simulation), refitted by the same script exactly
as plot_data.ipynb:188-200 does."""
import json
import os

import numpy as np
import pytest

from emu_backend import emu
from quantum_computations_b200 import channels, gates, workloads
from quantum_computations_b200.batched import BatchedSimulator

HERE = os.path.dirname(os.path.abspath(__file__))
PINS = json.load(open(os.path.join(HERE, "golden", "gkp_pins.json")))["levels"]

# the eight values SURVEY.md / BASELINE.md quote for mean(gate_error_I, gate_error_P), 5.83 .. 11.67 dB
QUOTED = [0.194, 0.140, 0.0949, 0.0600, 0.0350, 0.0187, 0.00895, 0.00382]


def test_analytic_error_values_match_the_notebook():
    assert len(PINS) == 9
    for row, quoted in zip(PINS, QUOTED):
        assert abs(row["analytic_mean"] - quoted) <= 0.5 * 10.0 ** np.floor(np.log10(quoted) - 2)   # as printed
    for row in PINS:
        nm = channels.GKPNoise(row["db"])
        assert abs(nm.epsilon - row["eps"]) <= 1e-14 * row["eps"]
        assert abs(nm.flip_probability(2) - row["flip_k2"]) <= 1e-12 * row["flip_k2"]
        assert abs(nm.flip_probability(3) - row["flip_k3"]) <= 1e-12 * row["flip_k3"]
        assert abs(nm.gate_error_I() - row["gate_error_I"]) <= 1e-12 * row["gate_error_I"]
        assert abs(nm.gate_error_P() - row["gate_error_P"]) <= 1e-12 * row["gate_error_P"]
        # what the channel does to a qubit: X flip with e(2); Z flip with e(2) (I/H type) or e(3) (P type)
        (px, pz), = nm.flips_for(gates.H(0))
        assert (px, pz) == (nm.flip_probability(2), nm.flip_probability(2))
        (px, pz), = nm.flips_for(gates.P(0))
        assert (px, pz) == (nm.flip_probability(2), nm.flip_probability(3))


def _fit_error_rate(depths, means, errors):
    """plot_data.ipynb:188-200: a p^m + 1/4, r = (1 - p)(1 - 2^-2)."""
    from scipy.optimize import curve_fit

    def exp_decay(m, a, p):
        return a * p ** m + 1 / 4
    p_guess = (4 / 3 * (means[-1] - 1 / 4)) ** (1 / depths[-1]) if means[-1] > 0.25 else 0.5
    popt, pcov = curve_fit(exp_decay, depths, means, sigma=errors, absolute_sigma=True, p0=[3 / 4, p_guess])
    return (1 - popt[1]) * 0.75, np.sqrt(pcov[1, 1]) * 0.75


def test_rb_decay_of_the_channel_lands_in_the_band_of_the_reference_data():
    """Two-qubit RB (the reference's random_circ generator) with this package's channel at every
    squeezing level of the data, same depths, fitted like the notebook: the error rate agrees
    with the fit of the reference's CV data within 4 sigma -- the notebook's own worst residual
    between that data and its analytic estimate is 3.9 sigma (11.67 dB)."""
    pytest.importorskip("scipy")
    worst = 0.0
    notebook_worst = max(abs(r["residual_sigma"]) for r in PINS)
    assert 3.5 < notebook_worst < 4.5
    for row in PINS:
        rng = np.random.default_rng(int(row["db"] * 1000))
        sim = BatchedSimulator(2, channels.GKPNoise(row["db"]), backend=emu())
        means, errors = [], []
        for depth in row["depths"]:
            circuits = [workloads.rb_random_circuit(2, depth, rng) for _ in range(200)]
            f = sim.run(circuits)["fidelity"]
            means.append(float(f.mean()))
            errors.append(float(f.std() / np.sqrt(len(f))))
        r, r_err = _fit_error_rate(row["depths"], means, errors)
        z = (r - row["r_fit"]) / np.hypot(r_err, row["r_err"])
        worst = max(worst, abs(z))
        assert abs(z) < 4.0, (row["db"], r, r_err, row["r_fit"], row["r_err"], z)
    assert worst < 4.0
