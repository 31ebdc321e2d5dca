"""Gate objects -- the drop-in for ``simulators/dv_simulator/gates.py``.

Class names, hierarchy, constructor signatures, ``.indices`` / ``.matrix`` /
``.copy()`` / ``.relabel()`` / ``__repr__`` and the validation messages follow
the reference (gates.py:7-194), because the reference's other packages dispatch
on these exact types.  What changes is ``apply``: instead
of expanding the gate to a dense 2^N x 2^N operator (gates.py:44-54 ->
numpy_quantum.py:243-247) it hands the small matrix to the CUDA engine.

``apply`` accepts either a NumPy array (reference behaviour: a *new* array comes
back, dtype following NumPy promotion; costs a host<->device round trip) or an
``engine.DeviceState`` (stays on the GPU; this is what ``Simulator.run`` uses).
"""
from __future__ import annotations

import numpy as np

from . import numpy_quantum as npq
from .states import State

REPR_DIGITS = 5


def _check_indices(indices) -> None:
    if len(set(indices)) != len(indices):
        raise ValueError("Indices must be distinct.")
    if min(indices) < 0:
        raise ValueError("Non-negative index")


class Gate:
    """A matrix acting on the listed qubits; the first tensor factor of the
    matrix acts on ``indices[0]`` (qubit 0 = most significant index bit)."""

    def __init__(self, indices: list[int], matrix: np.ndarray | None):
        _check_indices(indices)
        if matrix is not None:
            if matrix.ndim != 2:
                raise ValueError("Not a 2D array.")
            if not all(npq.is_power_of_two(extent) for extent in matrix.shape):
                raise ValueError("Given matrix is not a mapping between qubit spaces.")
            if matrix.shape[1] != 2 ** len(indices):
                raise ValueError("Dimensions of given matrix is not compatible with number of indices.")
        self.indices = indices
        self.matrix = matrix

    def __repr__(self):
        return f"{type(self).__name__}_" + ",".join(str(i) for i in self.indices)

    def copy(self) -> "Gate":
        """Shallow copy of the same concrete type."""
        twin = object.__new__(type(self))
        twin.__dict__.update(self.__dict__)
        return twin

    def relabel(self, mapping: dict):
        """Rename qubits in place: i -> mapping[i]."""
        renamed = []
        for i in self.indices:
            target = mapping.get(i, None)
            if target is None:
                raise ValueError(f"Index {i} does not map anywhere.")
            renamed.append(target)
        _check_indices(renamed)
        self.indices = renamed

    # -- engine hooks -----------------------------------------------------------------
    @property
    def _fusable(self) -> bool:
        """True if ``Simulator.run`` may batch this gate into a fused plan."""
        return self.matrix is not None

    def lowered(self, num_qubits: int, is_density: bool):
        """``[(targets, matrix), ...]`` in buffer-qubit numbering.  A density
        matrix lives on the device as its row-major vec, where U rho U^dagger is
        U on row qubit q and conj(U) on column qubit q + N."""
        if self.matrix is None:
            raise ValueError(f"Matrix representation not given for {self}.")
        if self.matrix.shape[0] != self.matrix.shape[1]:
            raise ValueError(f"{self}: only square matrices can be applied to a state.")
        if max(self.indices) >= num_qubits:
            raise ValueError("new_ordering must be a permutation of all qubits")
        ops = [(list(self.indices), self.matrix)]
        if is_density:
            ops.append(([i + num_qubits for i in self.indices], np.conjugate(self.matrix)))
        return ops

    def result_dtype(self, num_qubits: int, state_dtype) -> np.dtype:
        """dtype of ``expand_gate(...) @ state`` in the reference: the identity
        padding is float64, so only a gate covering every qubit keeps an integer
        matrix dtype."""
        mdt = self.matrix.dtype
        if num_qubits > len(self.indices):
            mdt = np.result_type(mdt, np.float64)
        return np.result_type(mdt, state_dtype)

    def apply(self, state):
        from . import engine
        if isinstance(state, engine.DeviceState):
            ops = self.lowered(state.num_qubits, state.ndim == 2)
            engine.apply_lowered(state, ops)
            state.host_dtype = self.result_dtype(state.num_qubits, state.host_dtype)
            return state
        if state.ndim not in (1, 2):
            raise ValueError("State has wrong dimensions.")
        dev = engine.DeviceState.from_numpy(state)
        return self.apply(dev).to_numpy()


class SingleQubitGate(Gate):
    def __init__(self, index: int, matrix):
        Gate.__init__(self, [index], matrix)


class TwoQubitGate(Gate):
    def __init__(self, index1: int, index2: int, matrix):
        Gate.__init__(self, [index1, index2], matrix)


# ---- the fixed-matrix zoo (gates.py:67-134), generated from a table ---------------------------
# Every entry becomes a class of that name whose constructor takes the qubit index (or the
# two indices) and nothing else, as in the reference; GKP/ dispatches on these types by name.
def _z_rotation(angle: float) -> np.ndarray:
    """diag(e^{-i a/2}, e^{+i a/2}): RZ, and P / Pdg / T / Tdg as its quarter and eighth turns."""
    return npq.axis_rotation(angle, [0, 0, 1])


def _one_qubit_class(name: str, build):
    def __init__(self, index):
        SingleQubitGate.__init__(self, index, build())
    return type(name, (SingleQubitGate,), {"__init__": __init__, "__module__": __name__, "__qualname__": name})


_ONE_QUBIT_TABLE = {
    "I": lambda: npq.IDTY, "X": lambda: npq.X, "Y": lambda: npq.Y, "Z": lambda: npq.Z, "H": lambda: npq.H,
    "P": lambda: _z_rotation(np.pi / 2), "Pdg": lambda: _z_rotation(-np.pi / 2),
    "T": lambda: _z_rotation(np.pi / 4), "Tdg": lambda: _z_rotation(-np.pi / 4),
}
I, X, Y, Z, H, P, Pdg, T, Tdg = (_one_qubit_class(_name, _build) for _name, _build in _ONE_QUBIT_TABLE.items())



class CX(TwoQubitGate):
    """Controlled X; the argument names are part of the surface (keyword calls)."""

    def __init__(self, control, target):
        TwoQubitGate.__init__(self, control, target, npq.CX)

    control = property(lambda self: self.indices[0])
    target = property(lambda self: self.indices[1])


class CZ(TwoQubitGate):
    def __init__(self, index1, index2):
        TwoQubitGate.__init__(self, index1, index2, npq.CZ)


class SWAP(TwoQubitGate):
    def __init__(self, index1, index2):
        TwoQubitGate.__init__(self, index1, index2, npq.SWAP)


class RZ(SingleQubitGate):
    def __init__(self, index, angle: float):
        SingleQubitGate.__init__(self, index, _z_rotation(angle))
        self.angle = angle

    def __repr__(self):
        return f"{Gate.__repr__(self)}({round(self.angle, REPR_DIGITS)})"


# ---- register-resizing operations (gates.py:136-194) ------------------------------------------
class Insert(SingleQubitGate):
    """Grow the register by one qubit, prepared in ``state``, at position ``index``."""

    def __init__(self, index: int, state: State):
        SingleQubitGate.__init__(self, index, state.get().reshape((1, 2)))
        self.state = state

    def __repr__(self):
        return f"{Gate.__repr__(self)}({self.state})"

    def apply(self, state):
        from . import engine
        vec = self.matrix[0, :]
        if isinstance(state, engine.DeviceState):
            return engine.insert(state, self.indices[0], vec)
        if state.ndim not in (1, 2):
            raise ValueError("State has wrong dimensions.")
        return engine.insert(engine.DeviceState.from_numpy(state), self.indices[0], vec).to_numpy()


class M(SingleQubitGate):
    """Projective measurement along the axis (theta, phi); removes the qubit and
    returns ``(state, outcome)``.  The contraction vectors are the *un-conjugated*
    ``Rz(phi) Ry(theta) e_s`` and the outcome is drawn from NumPy's global legacy
    generator, both exactly as in the reference (gates.py:169-183)."""

    def __init__(self, index: int, theta: float, phi: float, *, result: int = None):
        SingleQubitGate.__init__(self, index, None)
        if result is not None and result not in [0, 1]:
            raise ValueError(f"Measurement results must be from 0 or 1 but {result} was given.")
        self.theta = theta
        self.phi = phi
        self.result = result

    def vectors(self):
        rot = npq.axis_rotation(self.phi, [0, 0, 1]) @ npq.axis_rotation(self.theta, [0, 1, 0])
        return rot @ npq.ZERO, rot @ npq.ONE

    def apply(self, state):
        from . import engine
        v0, v1 = self.vectors()
        if isinstance(state, engine.DeviceState):
            return engine.measure(state, self.indices[0], v0, v1, self.result)
        if state.ndim not in (1, 2):
            raise ValueError("State has wrong dimensions.")
        out, s = engine.measure(engine.DeviceState.from_numpy(state), self.indices[0], v0, v1, self.result)
        return out.to_numpy(), s


class MZ(M):
    def __init__(self, index, *, result=None):
        M.__init__(self, index, 0.0, 0.0, result=result)


class MX(M):
    def __init__(self, index, *, result=None):
        M.__init__(self, index, np.pi / 2, 0.0, result=result)
