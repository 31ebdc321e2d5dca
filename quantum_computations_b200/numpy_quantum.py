"""Host-side linear-algebra kit: the drop-in for
``simulators/dv_simulator/numpy_quantum.py`` (reference lines cited per item).

Everything here is small-matrix host code (constants, 2x2 / 4x4 builders,
validators, metrics on NumPy arrays).  The reference's hot routine
``expand_gate`` (numpy_quantum.py:243-247) is kept for API compatibility at
small N but is *not* used by ``Gate.apply`` / ``Simulator.run`` in this
package: those go to the CUDA engine, which never materialises a 2^N x 2^N
operator.  ``fidelity`` / ``purity`` / ``norm`` accept device states
(``engine.DeviceState``) as well and then reduce on the GPU.

Dtypes of the constants match the reference (integer ``ZERO``/``X``/``Z``,
float ``H``/``CZ``/``CX``/``SWAP``, complex ``Y``/``P``/``T``) because result
dtypes follow NumPy promotion.
"""
from __future__ import annotations

import numpy as np

# ---- constants (numpy_quantum.py:5-25) --------------------------------------
_RT2 = np.sqrt(2)

ZERO = np.array([1, 0])
ONE = np.array([0, 1])
PLUS = np.array([1, 1]) / _RT2
MINUS = np.array([1, -1]) / _RT2
IPLUS = np.array([1, 1j]) / _RT2
IMINUS = np.array([1, -1j]) / _RT2

IDTY = np.identity(2)
X = np.array([[0, 1], [1, 0]])
Y = np.array([[0, -1j], [1j, 0]])
Z = np.array([[1, 0], [0, -1]])
PAULIS = [X, Y, Z]

H = np.array([[1, 1], [1, -1]]) / _RT2

CZ = np.diag([1.0, 1.0, 1.0, -1.0])
CX = np.identity(4)[[0, 1, 3, 2]]
SWAP = np.identity(4)[[0, 2, 1, 3]]

P = np.array([[1.0, 0.0], [0.0, 1.0j]])
T = np.array([[1.0, 0.0], [0.0, np.exp(0.25j * np.pi)]])


# ---- Pauli bookkeeping (numpy_quantum.py:28-76) ------------------------------
class PauliError(ValueError):
    pass


_PAULI_BY_NAME = {"i": 0, "x": 1, "y": 2, "z": 3, "-x": -1, "-y": -2, "-z": -3}
_PAULI_BY_AXIS = {(1, 0, 0): 1, (0, 1, 0): 2, (0, 0, 1): 3,
                  (-1, 0, 0): -1, (0, -1, 0): -2, (0, 0, -1): -3}


def get_pauli_number(pauli_identifier) -> int:
    """0 = I, 1..3 = X, Y, Z, negative = the negated operator.  Accepts names
    ('x', 'X', '-z', ...), the numbers themselves, or a unit axis [a, b, c]."""
    found = None
    if isinstance(pauli_identifier, str):
        if len(pauli_identifier) and pauli_identifier[-1].isalpha():
            found = _PAULI_BY_NAME.get(pauli_identifier[:-1] + pauli_identifier[-1].lower())
    elif isinstance(pauli_identifier, (list, tuple)):
        try:
            found = _PAULI_BY_AXIS.get(tuple(pauli_identifier))
        except TypeError:
            found = None
    else:
        try:
            if pauli_identifier in (-3, -2, -1, 0, 1, 2, 3):
                found = int(pauli_identifier)
        except (TypeError, ValueError):
            found = None
    if found is None:
        raise PauliError(f'"{pauli_identifier}" could not be interpreted as a Pauli operator')
    return found


def get_pauli_identifier(pauli_identifier) -> str:
    names = {-3: "-Z", -2: "-Y", -1: "-X", 0: "I", 1: "X", 2: "Y", 3: "Z"}
    return names[get_pauli_number(pauli_identifier)]


def is_pauli(case) -> bool:
    try:
        get_pauli_number(case)
    except PauliError:
        return False
    return True


def get_pauli_operator(pauli_identifier) -> np.ndarray:
    return PAULIS[get_pauli_number(pauli_identifier) - 1]


def get_pauli_states(pauli_identifier):
    eigenbases = ([PLUS, MINUS], [IPLUS, IMINUS], [ZERO, ONE])
    return eigenbases[get_pauli_number(pauli_identifier) - 1]


def get_pauli_state(pauli_identifier, state_index: int) -> np.ndarray:
    return get_pauli_states(pauli_identifier)[state_index]


# ---- state builders (numpy_quantum.py:79-101) ---------------------------------
def basis_state(identifier, N: int | None = None) -> np.ndarray:
    """Computational-basis ket from an int (needs N), a bit string, or a
    sequence of bits."""
    if isinstance(identifier, (list, tuple)):
        return basis_state("".join(str(b) for b in identifier))
    if isinstance(identifier, str):
        return basis_state(int(identifier, 2), len(identifier))
    if isinstance(identifier, (int, np.integer)):
        ket = np.zeros(2 ** N)
        ket[identifier] = 1
        return ket
    raise NotImplementedError(
        f"Could not generate basis state from identifier of type {type(identifier)}")


def qubit_from_polar(theta: float, phi: float) -> np.ndarray:
    return np.cos(theta / 2) * ZERO + np.exp(1j * phi) * np.sin(theta / 2) * ONE


def qubit_from_axis(axis) -> np.ndarray:
    length = np.sqrt(sum(a ** 2 for a in axis))
    return qubit_from_polar(np.arccos(axis[-1] / length), np.arctan2(axis[1], axis[0]))


# ---- small-matrix builders (numpy_quantum.py:100-109, :250-251) ---------------
def phase_gate(theta: float) -> np.ndarray:
    return np.array([[1, 0], [0, np.exp(1j * theta)]])


def axis_rotation(theta: float, axis) -> np.ndarray:
    """exp(-i theta/2 n.sigma).  The DV gate classes RZ/P/Pdg/T/Tdg are built
    from this, *not* from the constants ``P``/``T`` above (global phase)."""
    generator = axis[0] * X + axis[1] * Y + axis[2] * Z
    return IDTY * np.cos(theta / 2) - 1j * generator * np.sin(theta / 2)


def euler_rotation(theta1, theta2, theta3) -> np.ndarray:
    rx = lambda t: axis_rotation(t, [1, 0, 0])
    return rx(theta3) @ axis_rotation(theta2, [0, 0, 1]) @ rx(theta1)


def add_control(gate: np.ndarray) -> np.ndarray:
    dim = gate.shape[0]
    return tensor(np.outer(ZERO, ZERO), np.identity(dim)) + tensor(np.outer(ONE, ONE), gate)


# ---- ket / density-matrix helpers (numpy_quantum.py:112-166) -------------------
def _is_device(obj) -> bool:
    return hasattr(obj, "_qsim_device_state")


def ket2dm(ket: np.ndarray) -> np.ndarray:
    if len(ket.shape) != 1:
        raise TypeError("state is not a ket")
    return np.outer(ket, np.conjugate(ket))


def dm2ket(dm: np.ndarray, strict: bool = True) -> np.ndarray:
    if not is_hermitian(dm):
        raise TypeError("input is not a density matrix")
    evals, evecs = np.linalg.eigh(dm)
    if strict and not np.allclose(evals[:-1], 0):
        raise TypeError("density matrix does not represent a pure state")
    return normalise(evecs[:, -1])


def norm(ket) -> float:
    if _is_device(ket):
        return ket.norm()
    return np.linalg.norm(ket)


def normalise(state: np.ndarray) -> np.ndarray:
    if state.ndim == 1:
        return state / np.linalg.norm(state)
    if state.ndim == 2:
        return state / np.trace(state)
    raise ValueError("State not ket nor density matrix.")


def compare_kets(a: np.ndarray, b: np.ndarray) -> bool:
    return np.allclose(ket2dm(normalise(a)), ket2dm(normalise(b)))


def fidelity(a, b) -> float:
    """|<a|b>|^2, <psi|rho|psi>, or (tr sqrt(a b))^2 (numpy_quantum.py:148-161).
    Matrices are assumed Hermitian.  Device states reduce on the GPU except for
    the matrix-matrix case, which needs eigenvalues and stays on the host."""
    if _is_device(a) or _is_device(b):
        from .engine import device_fidelity
        return device_fidelity(a, b)
    a_ket, b_ket = a.ndim == 1, b.ndim == 1
    if a_ket and b_ket:
        return np.abs(a.conj() @ b).real ** 2
    if a_ket:
        return (a.conj() @ b @ a).real
    if b_ket:
        return (b.conj() @ a @ b).real
    spectrum = np.clip(np.linalg.eigvals(a @ b).real, 0.0, None)
    return np.sum(np.sqrt(spectrum)) ** 2


def purity(rho) -> float:
    if _is_device(rho):
        return rho.purity()
    return np.trace(rho @ rho).real


# ---- tensor-product plumbing (numpy_quantum.py:169-258) ------------------------
def tensor(*arrays) -> np.ndarray:
    out = 1
    for factor in arrays:
        out = np.kron(out, factor)
    return out


def is_power_of_two(n: int) -> bool:
    return n != 0 and (n & (n - 1)) == 0


def is_qubit_operator(oper: np.ndarray) -> bool:
    return oper.ndim == 2 and oper.shape[0] == oper.shape[1] and is_power_of_two(oper.shape[0])


def is_qubit_state(state: np.ndarray) -> bool:
    return state.ndim == 1 and is_power_of_two(len(state))


def is_hermitian(oper: np.ndarray) -> bool:
    return np.allclose(dagger(oper), oper)


def expect(oper: np.ndarray, state: np.ndarray):
    if not (is_qubit_operator(oper) and is_qubit_state(state) and oper.shape[0] == state.shape[0]):
        raise TypeError("incompatible operator and state vector")
    return np.conjugate(state) @ oper @ state


def expecth(oper: np.ndarray, state: np.ndarray):
    return expect(oper, state).real


def rand_ket(d=2) -> np.ndarray:
    return normalise(np.random.rand(d) + 1j * np.random.rand(d))


def dagger(array: np.ndarray) -> np.ndarray:
    return np.conjugate(array.T)


def num_qubits(arr) -> int:
    size = arr if isinstance(arr, int) else arr.shape[0]
    return int(np.log2(size))


def permute_tensor_product(array: np.ndarray, new_ordering) -> np.ndarray:
    """Move tensor factor j to position ``new_ordering[j]`` (rows, and columns
    too for operators) -- numpy_quantum.py:227-240."""
    if not is_power_of_two(array.shape[0]):
        raise ValueError("Given array is not a qubit state nor operator")
    nq = num_qubits(array)
    if set(new_ordering) != set(range(nq)):
        raise ValueError("new_ordering must be a permutation of all qubits")
    source_of = np.argsort(np.asarray(new_ordering))      # inverse permutation
    if array.ndim == 1:
        return array.reshape((2,) * nq).transpose(source_of).reshape(-1)
    cols = array.shape[1]
    if cols == array.shape[0]:
        cube = array.reshape((2,) * (2 * nq))
        axes = list(source_of) + [nq + s for s in source_of]
        return cube.transpose(axes).reshape(array.shape)
    # rectangular: rows first, then the column space on its own terms
    rows_done = array.reshape((2,) * nq + (cols,)).transpose(list(source_of) + [nq])
    rows_done = rows_done.reshape(array.shape)
    back = rows_done.T.reshape((2,) * nq + (-1,)).transpose(list(source_of) + [nq])
    return back.reshape((2 ** nq, -1)).T


def expand_gate(gate: np.ndarray, N: int, targets) -> np.ndarray:
    """Dense (2^N, 2^N) operator (numpy_quantum.py:243-247).  O(4^N): kept for
    compatibility and small N only; the simulator itself never calls it."""
    targets = list(targets)
    others = [q for q in range(N) if q not in targets]
    padded = tensor(gate, *([IDTY] * len(others)))
    return permute_tensor_product(padded, targets + others)
