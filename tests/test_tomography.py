"""Process tomography: the package's tomography module against
fixtures produced by the reference's own tomography.py (tests/golden/make_golden.py,
gen_tomography), and the closed loop through the engine: fit the Kraus operators of a
noisy gate simulated on the (emulated) device and compare with the channel that was
applied."""
import numpy as np
import pytest

import parity_cases as pc
from quantum_computations_b200 import channels, gates, tomography as tomo
from quantum_computations_b200 import numpy_quantum as npq


@pytest.fixture(scope="module")
def golden():
    return pc.load("tomography.npz")


def test_bases_match_reference(golden):
    meta, arrays = golden
    for N in (1, 2):
        assert np.array_equal(np.stack(tomo.state_basis(N)), arrays[f"state_basis{N}"])
        assert np.allclose(np.stack(tomo.pure_state_basis_kets(N)), arrays[f"pure_kets{N}"], rtol=0, atol=1e-15)
        assert np.allclose(np.stack(tomo.operator_basis(N)), arrays[f"operator_basis{N}"], rtol=0, atol=1e-15)


def test_process_chi_and_kraus_match_reference(golden):
    meta, arrays = golden
    for case in meta:
        tag, N = case["tag"], case["N"]
        inputs, outputs = list(arrays[tag + "_inputs"]), list(arrays[tag + "_outputs"])
        M = tomo.process_matrix(inputs, outputs)
        assert np.abs(M - arrays[tag + "_M"]).max() < 1e-12
        chi = tomo.chi_matrix(M, N)                      # the reference's convention (default)
        assert np.abs(chi - arrays[tag + "_chi"]).max() < 1e-12
        D, Ks = tomo.krauss_operators(chi, N)
        assert np.abs(D - arrays[tag + "_D"]).max() < 1e-12
        # Kraus sets are unique only up to unitary mixing: compare the maps they define
        keep = D > 1e-12
        assert int(keep.sum()) == case["n_fitted"]
        fitted = [np.sqrt(d) * k for d, k, f in zip(D, Ks, keep) if f]
        out = tomo.quantum_channel(fitted)(arrays[tag + "_probe"])
        assert np.abs(out - arrays[tag + "_probe_out"]).max() < 1e-12       # same numbers as the reference ...
        truth = tomo.superoperator(list(arrays[tag + "_kraus"]))
        assert np.abs(tomo.superoperator(fitted) - truth.T).max() < 1e-12   # ... which describe the transposed map
        # faithful convention: the Kraus operators of the map itself
        Df, Kf = tomo.krauss_operators(tomo.chi_matrix(M, N, strict=True, faithful=True), N)
        good = [np.sqrt(d) * k for d, k in zip(Df, Kf) if d > 1e-12]
        assert np.abs(tomo.superoperator(good) - truth).max() < 1e-12


def test_process_tomography_end_to_end_on_host(golden):
    meta, arrays = golden
    for case in meta:
        ks = list(arrays[case["tag"] + "_kraus"])
        process = tomo.quantum_channel(ks, ket_input=True, return_input=True)
        fitted = tomo.process_tomography(process, case["N"], strict=True, faithful=True)
        assert len(fitted) == case["n_fitted"]
        assert np.abs(tomo.superoperator(fitted) - tomo.superoperator(ks)).max() < 1e-12
        same_as_ref = tomo.process_tomography(process, case["N"])
        out = tomo.quantum_channel(same_as_ref)(arrays[case["tag"] + "_probe"])
        assert np.abs(out - arrays[case["tag"] + "_probe_out"]).max() < 1e-12
        D, Ks = tomo.process_tomography(process, case["N"], normalised=True, full_output=True)
        assert len(Ks) == 4 ** case["N"] and abs(D.sum() - 2 ** case["N"]) < 1e-10


def test_error_behaviour():
    rho = npq.ket2dm(npq.ZERO)
    with pytest.raises(ValueError):
        tomo.process_matrix([rho], [rho, rho])
    with pytest.raises(ValueError):                      # two inputs cannot span a 4-dimensional space
        tomo.process_matrix([rho, rho], [rho, rho])
    not_tp = tomo.quantum_channel([0.5 * npq.IDTY], ket_input=True, return_input=True)
    with pytest.raises(ValueError):
        tomo.process_tomography(not_tp, 1, strict=True, faithful=True)


@pytest.mark.parametrize("db", [8.0, 12.0])
def test_fit_of_a_noisy_gate_through_the_engine(emu_backend, db):
    """H with its GKP noise channel on one qubit and CZ with both channels on two qubits,
    simulated as density matrices by the engine: tomography must return Kraus operators
    of exactly the map  noise . U (.) U^dagger."""
    noise = channels.GKPNoise(db)
    for circuit, n in (([gates.H(0)], 1), ([gates.H(1), gates.CZ(0, 1)], 2)):
        fitted = tomo.circuit_kraus(circuit, n, noise=noise, backend=emu_backend)
        want = np.eye(4 ** n, dtype=complex)
        for g in circuit:
            u = np.asarray(npq.expand_gate(np.asarray(g.matrix, dtype=complex), n, list(g.indices)))
            want = np.kron(u, np.conjugate(u)) @ want
            for q, (px, pz) in zip(g.indices, noise.flips_for(g)):
                ks = [np.asarray(npq.expand_gate(np.asarray(k, dtype=complex), n, [q])) for k in noise.pauli_kraus(px, pz)]
                want = tomo.superoperator(ks) @ want
        assert np.abs(tomo.superoperator(fitted) - want).max() < 1e-12
        assert 1 < len(fitted) <= 4 ** n
