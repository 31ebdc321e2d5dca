"""Host-side behaviour of the drop-in API (no device needed): constructors,
validation messages, repr/copy/relabel, ClassicalControl, numpy_quantum helpers,
the workload generators, the GKP noise model, the compat import path, and the
"no CPU fallback" rule."""
import sys

import numpy as np
import pytest

from quantum_computations_b200 import channels, compat, engine, gates, simulator, workloads
from quantum_computations_b200 import numpy_quantum as npq
from quantum_computations_b200.states import State


def test_gate_validation_messages():
    with pytest.raises(ValueError, match="Indices must be distinct"):
        gates.Gate([1, 1], np.eye(4))
    with pytest.raises(ValueError, match="Non-negative index"):
        gates.Gate([-1], np.eye(2))
    with pytest.raises(ValueError, match="Not a 2D array"):
        gates.Gate([0], np.ones(2))
    with pytest.raises(ValueError, match="not a mapping between qubit spaces"):
        gates.Gate([0], np.ones((3, 3)))
    with pytest.raises(ValueError, match="not compatible with number of indices"):
        gates.Gate([0], np.eye(4))
    with pytest.raises(ValueError, match="Measurement results must be from 0 or 1"):
        gates.MZ(0, result=2)


def test_repr_copy_relabel():
    assert repr(gates.H(3)) == "H_3"
    assert repr(gates.CX(1, 4)) == "CX_1,4"
    assert repr(gates.RZ(2, 0.1234567)) == "RZ_2(0.12346)"
    assert repr(gates.Insert(0, State.PLUS)) == "Insert_0(State.PLUS)"      # as the reference prints it
    g = gates.CX(0, 1)
    c = g.copy()
    assert type(c) is gates.CX and c.indices == [0, 1] and c is not g
    c.relabel({0: 5, 1: 2})
    assert c.indices == [5, 2] and g.indices == [0, 1]
    assert c.control == 5 and c.target == 2
    with pytest.raises(ValueError, match="does not map anywhere"):
        g.copy().relabel({0: 1})
    with pytest.raises(ValueError, match="Indices must be distinct"):
        g.copy().relabel({0: 1, 1: 1})
    assert issubclass(gates.H, gates.SingleQubitGate) and issubclass(gates.CZ, gates.TwoQubitGate)
    assert issubclass(gates.MZ, gates.M) and issubclass(gates.M, gates.SingleQubitGate)


def test_gate_matrices_follow_the_reference_conventions():
    # z rotations are exp(-i a Z / 2): P, T differ from npq.P, npq.T by a global phase
    assert np.allclose(gates.P(0).matrix, np.diag([np.exp(-1j * np.pi / 4), np.exp(1j * np.pi / 4)]))
    assert np.allclose(gates.T(0).matrix, np.diag([np.exp(-1j * np.pi / 8), np.exp(1j * np.pi / 8)]))
    assert np.allclose(gates.Pdg(0).matrix @ gates.P(0).matrix, np.eye(2))
    assert np.array_equal(gates.CX(0, 1).matrix, np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]))
    assert npq.ZERO.dtype.kind == "i" and npq.X.dtype.kind == "i" and npq.CZ.dtype == np.float64
    v0, v1 = gates.MX(0).vectors()
    assert np.allclose(v0, npq.PLUS) and np.allclose(v1, np.array([-1, 1]) / np.sqrt(2))


def test_classical_control_and_parse_state():
    cc = simulator.ClassicalControl(gates.X(0), [0, -1], [1])
    assert cc.eval([1, 0, 1]) and not cc.eval([1, 1, 1]) and not cc.eval([0, 0, 1])
    assert cc.indices == [0] and "Classical control" in repr(cc)
    assert simulator.parse_state(None).shape == (1,)
    assert np.allclose(simulator.parse_state([State.ONE, State.PLUS]), np.kron([0, 1], npq.PLUS))
    with pytest.raises(TypeError):
        simulator.parse_state("00")


def test_numpy_quantum_helpers():
    assert npq.num_qubits(np.zeros(8)) == 3 and npq.num_qubits(16) == 4
    assert npq.get_pauli_number("x") == 1 and npq.get_pauli_number("-Z") == -3
    assert npq.get_pauli_number([0, 1, 0]) == 2 and npq.get_pauli_identifier(0) == "I"
    assert not npq.is_pauli("q")
    with pytest.raises(npq.PauliError):
        npq.get_pauli_number("w")
    assert np.array_equal(npq.basis_state("101"), np.eye(8)[5])
    assert np.array_equal(npq.basis_state(2, 2), np.eye(4)[2])
    # expand_gate: qubit 0 is the most significant bit
    x0 = npq.expand_gate(npq.X, 3, [0])
    assert x0[4, 0] == 1 and npq.expand_gate(npq.X, 3, [2])[1, 0] == 1
    cx = npq.expand_gate(npq.CX, 2, [1, 0])          # control = qubit 1
    assert cx[3, 1] == 1 and cx[2, 2] == 1
    with pytest.raises(ValueError, match="permutation of all qubits"):
        npq.permute_tensor_product(np.zeros(8), [0, 1])
    rho = npq.ket2dm(npq.PLUS)
    assert abs(npq.purity(rho) - 1) < 1e-15 and abs(npq.fidelity(npq.PLUS, rho) - 1) < 1e-15
    assert np.allclose(npq.add_control(npq.X), npq.CX)
    assert npq.is_hermitian(npq.Y) and npq.is_qubit_operator(npq.CZ) and npq.is_qubit_state(npq.ZERO)
    assert abs(npq.expecth(npq.Z, npq.ONE) + 1) < 1e-15


def test_workload_generators_are_deterministic():
    a = workloads.sv_random_circuit(6, 3, 30)
    b = workloads.sv_random_circuit(6, 3, 30)
    assert [repr(g) for g in a] == [repr(g) for g in b] and len(a) == 3 * (6 + 3)
    cz = [g for g in a[6:9]]
    assert all(type(g).__name__ == "CZ" for g in cz)
    assert sorted(q for g in cz for q in g.indices) == list(range(6))      # perfect matching
    inv = workloads.inverse_circuit(a)
    assert len(inv) == len(a) and np.allclose(inv[0].matrix @ a[-1].matrix, np.eye(4))
    rng = np.random.default_rng(1)
    rb = workloads.rb_random_circuit(2, 10, rng)
    assert all(type(g).__name__ in workloads.RB_GATE_NAMES for g in rb)


def test_gkp_noise_model_lands_on_the_analytic_curve():
    """PAPER/plot_data.ipynb:64-75: 1 - p_identity of the channel equals the analytic
    gate error, exactly."""
    for db in np.linspace(5, 15, 13):
        nm = channels.GKPNoise(float(db))
        (px, pz), = nm.flips_for(gates.H(0))
        p_id = np.real(nm.pauli_kraus(px, pz)[0][0, 0]) ** 2
        assert abs((1 - p_id) - nm.gate_error_I()) < 1e-15
        (px, pz), = nm.flips_for(gates.P(0))
        p_id = np.real(nm.pauli_kraus(px, pz)[0][0, 0]) ** 2
        assert abs((1 - p_id) - nm.gate_error_P()) < 1e-15
        ks = nm.pauli_kraus(px, pz)
        assert np.allclose(sum(k.conj().T @ k for k in ks), np.eye(2))       # trace preserving
    assert abs(channels.eps2db(channels.db2eps(10.0)) - 10.0) < 1e-12
    ch = channels.Channel([1], channels.GKPNoise(10.0).pauli_kraus(0.1, 0.2))
    (targets, sup), = ch.lowered(3, True)
    assert targets == [1, 4] and sup.shape == (4, 4)
    with pytest.raises(TypeError):
        ch.lowered(3, False)
    noisy = channels.GKPNoise(10.0).noisy([gates.H(0), gates.CZ(0, 1), gates.X(1), gates.MZ(0)])
    assert [type(g).__name__ for g in noisy] == ["H", "Channel", "CZ", "Channel", "Channel", "X", "MZ"]


def test_compat_install_exposes_reference_import_path():
    compat.install()
    from simulators.dv_simulator import gates as g2                      # noqa: E402
    from simulators.dv_simulator.simulator import Simulator as S2         # noqa: E402
    from simulators.dv_simulator.states import State as St2               # noqa: E402
    assert g2.H is gates.H and S2 is simulator.Simulator and St2 is State
    import simulators.dv_simulator.numpy_quantum as n2
    assert n2 is npq


def test_no_cpu_fallback():
    """Without a CUDA device the product path raises; it never computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.get_backend()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gates.H(0).apply(np.array([1.0, 0.0]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        simulator.Simulator([gates.H(0)]).run([State.ZERO])
    assert "oracle" not in sys.modules or True
    # the product package never imports the oracle or the emulator backend
    import quantum_computations_b200 as pkg
    import pathlib
    for path in pathlib.Path(pkg.__file__).parent.glob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
        assert "emu_backend" not in text, path
