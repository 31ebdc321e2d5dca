"""Parity tests proper: the CUDA path (through the C ABI of libqsim_b200.so)
against the golden vectors of the real reference and the CPU oracle.  Same
checks as tests/test_emulator_parity.py, plus sizes only a GPU can hold, where
parity is established through size-independent properties."""
import numpy as np
import pytest

import parity_cases as pc
from golden.specs import as_oracle_ops
from oracle import strided
from quantum_computations_b200 import engine, gates, workloads
from quantum_computations_b200.simulator import Simulator
from quantum_computations_b200.states import State

pytestmark = pytest.mark.gpu


def test_library_is_cuda(cuda_backend):
    assert cuda_backend.lib.qsim_has_cuda() == 1
    before = engine.launch_count(cuda_backend)
    gates.H(0).apply(np.array([1.0, 0.0]))
    assert engine.launch_count(cuda_backend) > before


def test_single_gates(cuda_backend):
    pc.check_single_gates(cuda_backend)


def test_density_gates(cuda_backend):
    pc.check_density_gates(cuda_backend)


def test_measure(cuda_backend):
    pc.check_measure(cuda_backend)


def test_insert(cuda_backend):
    pc.check_insert(cuda_backend)


def test_kraus(cuda_backend):
    pc.check_kraus(cuda_backend)


def test_metrics(cuda_backend):
    pc.check_metrics(cuda_backend)


def test_sv_circuits(cuda_backend):
    pc.check_sv_circuits(cuda_backend)
    pc.check_sv_circuits(cuda_backend, plan_options=dict(tile_bits=7, low_bits=2, max_group=2))
    pc.check_sv_circuits(cuda_backend, plan_options=dict(tile_bits=11, low_bits=5, max_group=1))


def test_grover(cuda_backend):
    pc.check_grover(cuda_backend)
    pc.check_noisy_grover(cuda_backend)


def test_sim_measure(cuda_backend):
    pc.check_sim_measure(cuda_backend)


def test_rb(cuda_backend):
    pc.check_rb(cuda_backend)


def test_random_vs_oracle_small_tiles(cuda_backend):
    pc.check_random_vs_oracle(cuda_backend, trials=60, n_range=(1, 12), tile_range=(5, 9), seed=21)


def test_random_vs_oracle_default_tiles(cuda_backend):
    pc.check_random_vs_oracle(cuda_backend, trials=12, n_range=(12, 18), tile_range=(10, 13), seed=22)


@pytest.mark.parametrize("n", [3, 15, 16, 17, 21])
def test_product_state_matches_kron(cuda_backend, n):
    """parse_state(list) (DV/simulator.py:26): the two product-state kernels (plain below
    16 qubits, table-driven from 16 on) against numpy's kron of the same kets."""
    rng = np.random.default_rng(n)
    kets = rng.normal(size=(n, 2)) + 1j * rng.normal(size=(n, 2))
    ref = np.ones(1, dtype=np.complex128)
    for k in kets:
        ref = np.kron(ref, k)
    got = engine.DeviceState.product(list(kets), cuda_backend).to_numpy()
    assert pc.rel_err(got, ref) < 1e-13


def test_tomography_of_a_noisy_gate(cuda_backend):
    """SURVEY 8f rank 4 on the device: fit the Kraus operators of H and CZ with their GKP noise
    channels from density-matrix simulations of the 4^N probe states."""
    from quantum_computations_b200 import channels, tomography as tomo
    from quantum_computations_b200 import numpy_quantum as npq
    noise = channels.GKPNoise(10.0)
    for circuit, n in (([gates.H(0)], 1), ([gates.H(1), gates.CZ(0, 1)], 2)):
        fitted = tomo.circuit_kraus(circuit, n, noise=noise, backend=cuda_backend)
        want = np.eye(4 ** n, dtype=complex)
        for g in circuit:
            u = np.asarray(npq.expand_gate(np.asarray(g.matrix, dtype=complex), n, list(g.indices)))
            want = np.kron(u, np.conjugate(u)) @ want
            for q, (px, pz) in zip(g.indices, noise.flips_for(g)):
                ks = [np.asarray(npq.expand_gate(np.asarray(k, dtype=complex), n, [q])) for k in noise.pauli_kraus(px, pz)]
                want = tomo.superoperator(ks) @ want
        assert np.abs(tomo.superoperator(fitted) - want).max() < 1e-12


def test_edge_cases(cuda_backend):
    pc.check_edge_cases(cuda_backend)


def test_swap_pack_unpack(cuda_backend):
    pc.check_swap_pack_unpack(cuda_backend)


def test_plan_cache(cuda_backend):
    pc.check_plan_cache(cuda_backend)


def test_deferred_tails(cuda_backend):
    pc.check_deferred_tails(cuda_backend, trials=30, seed=6)


def test_dm_layers(cuda_backend):
    pc.check_dm_layers_vs_oracle(cuda_backend, n=4, depth=6, seed=12)
    pc.check_dm_layers_vs_oracle(cuda_backend, n=6, depth=4, seed=12)
    pc.check_dm_layers_vs_oracle(cuda_backend, n=8, depth=2, seed=12)


@pytest.mark.parametrize("cta_log2", [0, 7, 8])
def test_sv_22q_vs_oracle(cuda_backend, cta_log2):
    """Config C4's generator at a size the strided oracle still finishes in seconds, on both CTA
    sizes of the tile pass (1024 tiles: every CTA loops over several)."""
    n, depth = 22, 3
    circ = workloads.sv_random_circuit(n, depth, 30)
    got = Simulator(circ, backend=cuda_backend, plan_options={"cta_log2": cta_log2}).run([State.ZERO] * n)
    psi0 = np.zeros(2 ** n, dtype=np.complex128)
    psi0[0] = 1.0
    ref, _ = strided.run(as_oracle_ops(circ), psi0)
    assert pc.rel_err(got, ref) < pc.RTOL


@pytest.mark.parametrize("n", [26, 30])
def test_circuit_then_inverse_returns_zero_state(cuda_backend, n):
    """No oracle reaches this size: U^-1 U |0> must be |0> and the norm must not
    drift."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < (16 << n) * 1.2:
        pytest.skip("not enough device memory")
    depth = 6
    circ = workloads.sv_random_circuit(n, depth, 30)
    both = circ + workloads.inverse_circuit(circ)
    mid = Simulator(circ, backend=cuda_backend).run([State.ZERO] * n, return_device=True)
    assert abs(mid.norm() - 1.0) < 1e-12
    # probability mass must have left |0...0> before the inverse brings it back
    amp0 = mid.buf[0].item()
    assert abs(amp0) < 0.9
    del mid
    out = Simulator(both, backend=cuda_backend).run([State.ZERO] * n, return_device=True)
    assert abs(out.norm() - 1.0) < 1e-12
    assert abs(out.buf[0].item() - 1.0) < 1e-10
    rest = torch.linalg.vector_norm(out.buf[1:]).item()
    assert rest < 1e-10


def test_planner_features(cuda_backend):
    """Paired dense blocks, quarter-turn split, widened steps, both CTA sizes (round 2)."""
    pc.check_planner_features(cuda_backend)


def test_trajectories_batched(cuda_backend):
    """SURVEY 8f rank 1 on the device, against the CPU oracle."""
    pc.check_trajectories(cuda_backend, shots=20000)


def test_layered_noise(cuda_backend):
    """SURVEY 8f rank 2 on the device, against the CPU oracle."""
    pc.check_layered_noise(cuda_backend)


def test_clifford_rb(cuda_backend):
    """SURVEY 8f rank 3 on the device, against the CPU oracle."""
    pc.check_clifford_rb(cuda_backend)


def test_nonblocking_run(cuda_backend):
    pc.check_nonblocking_run(cuda_backend)
