"""Pin the CPU oracle (oracle/dense_ref.py, oracle/strided.py) against every
golden vector produced by the real reference (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from golden.specs import as_oracle_ops, from_spec
from oracle import dense_ref, gkp_noise, strided
from parity_cases import load, mk, rel_err
from quantum_computations_b200 import workloads
from quantum_computations_b200.states import State

TOL = 1e-12
ORACLES = [dense_ref, strided]


@pytest.mark.parametrize("oracle", ORACLES, ids=["dense", "strided"])
def test_single_and_density_gates(oracle):
    for name in ("single_gates.npz", "density_gates.npz"):
        meta, z = load(name)
        for rec in meta:
            gate = mk(rec["gate"], z)
            got = oracle.apply_matrix(z[rec["in"]], gate.indices, gate.matrix)
            assert rel_err(got, z[rec["out"]]) < TOL, rec["gate"]


@pytest.mark.parametrize("oracle", ORACLES, ids=["dense", "strided"])
def test_measure(oracle):
    meta, z = load("measure.npz")
    for rec in meta:
        if "stream" in rec:
            np.random.seed(rec["seed"])
            got = [oracle.measure(z[rec["in"]], rec["q"], 0.0, 0.0)[1] for _ in rec["stream"]]
            assert got == rec["stream"]
            continue
        if rec["seed"] is not None:
            np.random.seed(rec["seed"])
        out, s = oracle.measure(z[rec["in"]], rec["q"], rec["theta"], rec["phi"], rec["forced"])
        assert s == rec["s"]
        assert rel_err(out, z[rec["out"]]) < TOL


def test_draw_outcome_matches_numpy_choice():
    rng = np.random.default_rng(7)
    for seed in range(300):
        p0 = float(rng.uniform())
        np.random.seed(seed)
        want = int(np.random.choice([0, 1], p=[p0, 1 - p0]))
        np.random.seed(seed)
        u = np.random.random_sample()
        assert dense_ref.draw_outcome(p0, 1 - p0, u) == want
    with pytest.raises(ValueError):
        dense_ref.draw_outcome(0.5, 0.6, 0.1)


@pytest.mark.parametrize("oracle", ORACLES, ids=["dense", "strided"])
def test_insert(oracle):
    meta, z = load("insert.npz")
    for rec in meta:
        got = oracle.insert_qubit(z[rec["in"]], rec["pos"], State[rec["state"]].get())
        assert rel_err(got, z[rec["out"]]) < TOL


@pytest.mark.parametrize("oracle", ORACLES, ids=["dense", "strided"])
def test_kraus(oracle):
    meta, z = load("kraus.npz")
    for rec in meta:
        got = oracle.apply_kraus(z[rec["in"]], rec["indices"], list(z[rec["kraus"]]))
        assert rel_err(got, z[rec["out"]]) < TOL
        if "px" in rec:      # the analytic model's numbers are the ones the fixture was made with
            assert abs(gkp_noise.flip_probability(rec["db"], rec["kx"]) - rec["px"]) < 1e-15


def test_circuits_strided():
    meta, z = load("circuits.npz")
    for rec in meta:
        circ = workloads.sv_random_circuit(rec["n"], rec["depth"], rec["seed"])
        psi0 = np.zeros(2 ** rec["n"]); psi0[0] = 1.0
        got, _ = strided.run(as_oracle_ops(circ), psi0)
        assert rel_err(got, z[rec["out"]]) < TOL


def test_circuits_from_random_kets():
    """C4 generator from random initial kets (every output amplitude distinct), N = 10 and 11:
    both oracle modules against the reference's own output."""
    meta, z = load("circuits_rand.npz")
    for rec in meta:
        assert rec["distinct_amplitudes"] == 2 ** rec["n"]
        circ = workloads.sv_random_circuit(rec["n"], rec["depth"], rec["seed"])
        got, _ = strided.run(as_oracle_ops(circ), z[rec["in"]])
        assert rel_err(got, z[rec["out"]]) < TOL
        if rec["n"] <= 10:
            got, _ = dense_ref.run(as_oracle_ops(circ), z[rec["in"]])
            assert rel_err(got, z[rec["out"]]) < TOL


def test_circuits_dense_small():
    meta, z = load("circuits.npz")
    for rec in meta:
        if rec["n"] > 8:
            continue
        circ = workloads.sv_random_circuit(rec["n"], rec["depth"], rec["seed"])
        psi0 = np.zeros(2 ** rec["n"]); psi0[0] = 1.0
        got, _ = dense_ref.run(as_oracle_ops(circ), psi0)
        assert rel_err(got, z[rec["out"]]) < TOL


@pytest.mark.parametrize("oracle", ORACLES, ids=["dense", "strided"])
def test_grover_and_sim_measure(oracle):
    meta, z = load("grover.npz")
    for rec in meta:
        circ = [mk(s, z) for s in rec["circuit"]]
        if rec["kind"] == "full":
            start = np.ones((1,))
        else:
            start = dense_ref.product_state([State[s].get() for s in rec["init"]])
        got, _ = oracle.run(as_oracle_ops(circ), start)
        assert rel_err(got, z[rec["out"]]) < TOL
    meta, z = load("sim_measure.npz")
    circ = [mk(s, z) for s in meta["circuit"]]
    for run in meta["runs"]:
        np.random.seed(run["seed"])
        state, results = np.ones((1,)), []
        for gate in circ:
            if type(gate).__name__ == "ClassicalControl":
                if not gate.eval(results):
                    continue
                gate = gate.gate
            state, out = oracle.run(as_oracle_ops([gate]), state)
            results += out
        assert results == run["results"]
        assert rel_err(state, z[run["out"]]) < TOL


def test_metrics_and_noisy_grover():
    meta, z = load("metrics.npz")
    for rec in meta:
        n = rec["n"]
        a, b, ra, rb = (z[f"{k}{n}"] for k in ("a", "b", "ra", "rb"))
        assert abs(dense_ref.fidelity(a, b) - rec["f_kk"]) < 1e-13
        assert abs(dense_ref.fidelity(a, rb) - rec["f_kr"]) < 1e-13
        assert abs(dense_ref.fidelity(ra, b) - rec["f_rk"]) < 1e-13
        assert abs(dense_ref.fidelity(ra, rb) - rec["f_rr"]) < 1e-10
        assert abs(dense_ref.purity(ra) - rec["purity"]) < 1e-13
    meta, z = load("noisy_grover.npz")
    from quantum_computations_b200 import channels
    for rec in meta[:3]:
        circ = channels.GKPNoise(rec["db"]).noisy([mk(s, z) for s in rec["circuit"]])
        psi0 = dense_ref.product_state([State[s].get() for s in rec["init"]]).astype(np.complex128)
        got, _ = strided.run(as_oracle_ops(circ), np.outer(psi0, psi0.conj()))
        assert rel_err(got, z[rec["out"]]) < TOL


def test_dense_equals_strided_random():
    rng = np.random.default_rng(5)
    from parity_cases import random_circuit
    for _ in range(10):
        n = int(rng.integers(2, 7))
        psi = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
        circ = random_circuit(n, 20, rng)
        a, _ = dense_ref.run(as_oracle_ops(circ), psi)
        b, _ = strided.run(as_oracle_ops(circ), psi)
        assert rel_err(a, b) < TOL


def test_gkp_analytic_values():
    """analytic mean of I-type and P-type gate errors."""
    dbs = [5.83, 6.67, 7.50, 8.33, 9.17, 10.0, 10.83, 11.67]
    want = [0.194, 0.140, 0.0949, 0.0600, 0.0350, 0.0187, 0.00895, 0.00382]
    for db, w in zip(dbs, want):
        exact_db = 5 + round((db - 5) / (10 / 12)) * (10 / 12)     # np.linspace(5, 15, 13) grid
        got = (gkp_noise.gate_error_I(exact_db) + gkp_noise.gate_error_P(exact_db)) / 2
        assert abs(got - w) / w < 0.01, (db, got, w)
