"""Package + planner + tile index arithmetic on the HOST EMULATOR backend
(tests/emu_backend.py) against the golden vectors and the oracle.  The same
checks run on the real CUDA backend in tests/test_gpu_parity.py."""
import parity_cases as pc


def test_single_gates(emu_backend):
    pc.check_single_gates(emu_backend)


def test_density_gates(emu_backend):
    pc.check_density_gates(emu_backend)


def test_measure(emu_backend):
    pc.check_measure(emu_backend)


def test_insert(emu_backend):
    pc.check_insert(emu_backend)


def test_kraus(emu_backend):
    pc.check_kraus(emu_backend)


def test_metrics(emu_backend):
    pc.check_metrics(emu_backend)


def test_sv_circuits(emu_backend):
    pc.check_sv_circuits(emu_backend, max_n=12)
    pc.check_sv_circuits(emu_backend, plan_options=dict(tile_bits=7, low_bits=2, max_group=2), max_n=8)


def test_grover(emu_backend):
    pc.check_grover(emu_backend)
    pc.check_noisy_grover(emu_backend)


def test_sim_measure(emu_backend):
    pc.check_sim_measure(emu_backend)


def test_rb(emu_backend):
    pc.check_rb(emu_backend)


def test_random_vs_oracle_small_tiles(emu_backend):
    pc.check_random_vs_oracle(emu_backend, trials=40, n_range=(1, 11), tile_range=(5, 8), seed=11)


def test_random_vs_oracle_default_tiles(emu_backend):
    pc.check_random_vs_oracle(emu_backend, trials=6, n_range=(12, 14), tile_range=(9, 12), seed=12)


def test_edge_cases(emu_backend):
    pc.check_edge_cases(emu_backend)


def test_swap_pack_unpack(emu_backend):
    pc.check_swap_pack_unpack(emu_backend)


def test_plan_cache(emu_backend):
    pc.check_plan_cache(emu_backend)


def test_deferred_tails(emu_backend):
    pc.check_deferred_tails(emu_backend, trials=30, seed=5)


def test_dm_layers(emu_backend):
    pc.check_dm_layers_vs_oracle(emu_backend, n=4, depth=6, seed=12)
    pc.check_dm_layers_vs_oracle(emu_backend, n=6, depth=3, seed=12)


def test_planner_features(emu_backend):
    pc.check_planner_features(emu_backend)


def test_trajectories_batched(emu_backend):
    pc.check_trajectories(emu_backend, shots=3000)


def test_layered_noise(emu_backend):
    pc.check_layered_noise(emu_backend)


def test_clifford_rb(emu_backend):
    pc.check_clifford_rb(emu_backend)


def test_nonblocking_run(emu_backend):
    pc.check_nonblocking_run(emu_backend)
