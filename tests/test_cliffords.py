"""Clifford table: the known answers printed by the reference
(PAPER/average_clifford_fidelity.py:141-142: 720 classes, Cayley diameter 7) and
uniform Clifford RB through the batched executor on the emulator backend."""
import numpy as np

from quantum_computations_b200 import channels, cliffords
from quantum_computations_b200.batched import BatchedSimulator


def test_720_classes_and_diameter_7():
    unitaries, depths = cliffords.two_qubit_cliffords_mod_paulis()
    assert len(unitaries) == 720
    assert max(depths) == 7
    keys = {cliffords.symplectic_rep(u).tobytes() for u in unitaries}
    assert len(keys) == 720
    for u in unitaries[::37]:
        assert np.allclose(u @ u.conj().T, np.eye(4))
        s = cliffords.symplectic_rep(u)
        omega = np.kron(np.eye(2, dtype=int), np.array([[0, 1], [1, 0]]))
        assert np.array_equal((s.T @ omega @ s) % 2, omega)          # symplectic


def test_clifford_rb_decays_to_one_quarter(emu_backend):
    rng = np.random.default_rng(11)
    seqs = cliffords.clifford_rb_sequences(40, 6, rng)
    clean = BatchedSimulator(2, None, backend=emu_backend).run(seqs)
    assert np.allclose(clean["fidelity"], 1.0, atol=1e-12)           # inverse brings |00> back
    noisy = BatchedSimulator(2, channels.GKPNoise(6.0), backend=emu_backend).run(seqs)
    f = noisy["fidelity"].mean()
    assert 0.2 < f < 0.9                                              # decaying towards 1/4
    long_seqs = cliffords.clifford_rb_sequences(40, 40, rng)
    f_long = BatchedSimulator(2, channels.GKPNoise(6.0), backend=emu_backend).run(long_seqs)["fidelity"].mean()
    assert abs(f_long - 0.25) < 0.03                                  # PAPER/plot_data.ipynb:188-189 asymptote
