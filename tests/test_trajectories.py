"""Pauli-trajectory executor against the density-matrix path:
the shot average of |psi><psi| must converge to the noisy rho (checked on the
emulator backend with a fixed seed and a statistical tolerance)."""
import numpy as np

from quantum_computations_b200 import channels, gates, trajectories
from quantum_computations_b200.simulator import Simulator
from quantum_computations_b200.states import State


def test_trajectory_average_matches_density_matrix(emu_backend):
    noise = channels.GKPNoise(7.0)                       # strong noise: errors in most shots
    circ = [gates.H(0), gates.CZ(0, 1), gates.T(1), gates.H(1), gates.CZ(1, 2), gates.P(2), gates.H(2),
            gates.SWAP(0, 2), gates.H(0)]
    init = [State.ZERO] * 3
    rho0 = np.zeros((8, 8), dtype=np.complex128)
    rho0[0, 0] = 1.0
    rho = Simulator(noise.noisy(circ), backend=emu_backend).run(rho0)
    ideal = Simulator(circ, backend=emu_backend).run(init)
    shots = 3000
    res = trajectories.run_trajectories(circ, noise, init, shots, np.random.default_rng(5), backend=emu_backend,
                                        observable=ideal)
    want_p = np.real(np.diagonal(rho))
    want_f = np.real(np.vdot(ideal, rho @ ideal))
    # binomial standard error for 3000 shots is <= 0.0092; allow 4 sigma
    assert np.max(np.abs(res["probabilities"] - want_p)) < 0.04
    assert abs(res["fidelity"] - want_f) < 0.04
    assert abs(res["probabilities"].sum() - 1.0) < 1e-12


def test_noiseless_limit_is_exact(emu_backend):
    noise = channels.GKPNoise(40.0)                      # flip probabilities ~ 0
    circ = [gates.H(0), gates.CX(0, 1)]
    res = trajectories.run_trajectories(circ, noise, [State.ZERO] * 2, 5, np.random.default_rng(1),
                                        backend=emu_backend)
    assert np.allclose(res["probabilities"], [0.5, 0, 0, 0.5], atol=1e-12)
