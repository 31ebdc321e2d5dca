"""Circuit executor -- the drop-in for ``simulators/dv_simulator/simulator.py``.

``Simulator(circuit).run(initial_state)`` keeps the reference's contract
(simulator.py:30-53: sequential semantics, ``results`` filled by measurements,
``ClassicalControl`` feed-forward, initial state ``None`` / ndarray /
``list[State]``, a NumPy array back) but executes differently: the state is
uploaded once, maximal runs of matrix gates are compiled into fused tile passes
by the native planner and run by the CUDA kernels, measurements / insertions
split the runs, and the result is downloaded once at the end.
"""
from __future__ import annotations

import numpy as np

from .gates import Gate, Insert, M
from .numpy_quantum import tensor
from .states import State


class ClassicalControl:
    """Apply ``gate`` only if the listed earlier measurement results are all 1
    (``positive_indices``) / all 0 (``negative_indices``).  Indices address
    ``Simulator.results`` and may be negative (simulator.py:6-17)."""

    def __init__(self, gate: Gate, positive_indices: list[int] = [], negative_indices: list[int] = []):
        self.gate = gate
        self.indices = gate.indices
        self._pos = positive_indices
        self._neg = negative_indices

    def __repr__(self):
        return f"Classical control: {self.gate}"

    def eval(self, observables: list[bool]) -> bool:
        wanted_on = all(observables[i] for i in self._pos)
        wanted_off = not any(observables[i] for i in self._neg)
        return wanted_on and wanted_off


def parse_state(state) -> np.ndarray:
    """Host-side form of the initial state (simulator.py:19-28).  ``Simulator.run``
    itself builds ``list[State]`` inputs directly on the device instead."""
    if state is None:
        return np.ones((1,))
    if isinstance(state, np.ndarray):
        return state
    if isinstance(state, list) and all(isinstance(item, State) for item in state):
        return tensor(*(s.get() for s in state))
    raise TypeError("Unsupported input type")


def _is_fusable(gate) -> bool:
    """Plain matrix gates whose ``apply`` is the stock one can be batched into a
    fused plan; anything else (M, Insert, user subclasses overriding ``apply``)
    is applied on its own."""
    return isinstance(gate, Gate) and type(gate).apply is Gate.apply and gate._fusable


class Simulator:
    def __init__(self, circuit: list[Gate], rng_seed: int = None, *, backend=None, plan_options=None):
        self.circuit: list[Gate] = circuit
        self.results: list[int] = None
        # kept for signature compatibility; like the reference, measurements
        # draw from NumPy's global generator, not from this one (gates.py:183)
        self._rng = np.random.default_rng(rng_seed)
        self._backend = backend
        self._plan_options = plan_options
        self.last_stats: list[dict] = []
        self._plans: dict = {}              # compiled plans of earlier runs, by content of the segment

    # -- helpers -----------------------------------------------------------------------
    def _initial(self, initial_state):
        from . import engine
        backend = self._backend or engine.get_backend()
        if isinstance(initial_state, engine.DeviceState):
            return initial_state
        if isinstance(initial_state, list) and initial_state and \
                all(isinstance(item, State) for item in initial_state):
            return engine.DeviceState.product([s.get() for s in initial_state], backend)
        return engine.DeviceState.from_numpy(parse_state(initial_state), backend)

    def _flush(self, state, segment):
        from . import engine
        if not segment:
            return
        nq, dens = state.num_qubits, state.ndim == 2
        ops = []
        dtype = state.host_dtype
        for gate in segment:
            ops.extend(gate.lowered(nq, dens))
            dtype = gate.result_dtype(nq, dtype)
        self.last_stats.append(self._apply_cached(state, ops))
        state.host_dtype = dtype
        segment.clear()

    _PLAN_CACHE_SIZE = 16

    def _apply_cached(self, state, ops) -> dict:
        """Like ``engine.apply_lowered``, but a plan compiled by an earlier ``run`` of this
        simulator is reused when the segment is the same gate for gate (matrices compared
        by value through a 128-bit digest, so mutating a gate between runs is safe)."""
        import hashlib
        from . import engine
        if not ops:
            return {}
        h = hashlib.blake2b(digest_size=16)
        h.update(repr((state.n_bits, id(state.backend), sorted((self._plan_options or {}).items()))).encode())
        for targets, matrix in ops:
            h.update(bytes(targets) if max(targets) < 256 else repr(targets).encode())
            h.update(b"|")
            h.update(np.ascontiguousarray(matrix, dtype=np.complex128).data)
        key = h.digest()
        plan = self._plans.get(key)
        if plan is None:
            plan = engine.Plan(state.backend, state.n_bits, ops, self._plan_options)
            if len(self._plans) >= self._PLAN_CACHE_SIZE:
                self._plans.pop(next(iter(self._plans)))
            self._plans[key] = plan
        plan.execute(state.buf)
        return plan.stats

    # -- public ---------------------------------------------------------------------------
    def run(self, initial_state=None, *, out: np.ndarray | None = None, return_device: bool = False,
            block: bool = True):
        """Run the circuit.  Additive keyword arguments: ``out`` (a preallocated,
        ideally pinned, complex128 host buffer to receive the final state),
        ``return_device`` (hand back the ``DeviceState`` without any download) and
        ``block=False`` (needs ``out``): return an ``engine.PendingState`` as soon as the work is
        queued -- the device-to-host copy runs on a side stream, so the next ``run`` computes
        while this result crosses PCIe; ``.result()`` waits for it."""
        from . import engine
        self.results = []
        self.last_stats = []
        state = self._initial(initial_state)
        if isinstance(initial_state, engine.DeviceState):
            state = state.copy()

        segment: list[Gate] = []
        for gate in self.circuit:
            if isinstance(gate, ClassicalControl):
                # feed-forward needs the measurement results so far; every M
                # before this point has already been executed (it ends a segment)
                if not gate.eval(self.results):
                    continue
                gate = gate.gate
            if _is_fusable(gate) and state.n_bits >= 1:
                segment.append(gate)
                continue
            self._flush(state, segment)
            if isinstance(gate, (M, Insert)):
                output = gate.apply(state)
            else:
                # foreign gate type with its own apply(): give it what the
                # reference would, a NumPy array, and take the result back
                output = gate.apply(state.to_numpy())
                if isinstance(output, tuple):
                    output = (engine.DeviceState.from_numpy(output[0], state.backend), output[1])
                else:
                    output = engine.DeviceState.from_numpy(output, state.backend)
            if isinstance(output, tuple):
                state = output[0]
                self.results.append(output[1])
            else:
                state = output
        self._flush(state, segment)

        if return_device:
            return state
        if not block:
            if out is None:
                raise ValueError("block=False needs a preallocated `out` buffer")
            return state.to_numpy_async(out)
        return state.to_numpy(out=out)
