"""Process tomography of N-qubit channels (N <= 2 in practice) and the bridge to the
device engine: fit the Kraus operators of whatever a (noisy) circuit does.

Same surface as the reference's ``tomography.py`` (impact_of_finite_squeezing...,
:14-215: ``quantum_channel``, ``state_basis``, ``pure_state_basis_kets``,
``operator_basis``, ``process_matrix``, ``chi_matrix``, ``krauss_operators``,
``eval_process``, ``process_tomography``; the reference's spelling of "krauss" is kept),
with the same conventions:

* density matrices are flattened row-major: ``vec(rho)[a * d + b] = rho[a, b]``, and the
  process matrix ``M`` acts as ``vec(E(rho)) = M vec(rho)``  (tomography.py:82-85);
* the operator basis is the Hermitian Pauli basis scaled to unit Hilbert-Schmidt norm,
  ``P_m = sigma_{m1} (x) ... / sqrt(2)^N``  (tomography.py:63-72), and the chi matrix is
  defined by ``E(rho) = sum_mn chi[m, n] P_m rho P_n``  (tomography.py:109-121).

The linear algebra is formulated differently from the reference.  With row-major vec,
``vec(P_m rho P_n) = (P_m (x) P_n^T) vec(rho)``, and the d^2 matrices ``P_m (x) P_n^T`` are
orthonormal under the Hilbert-Schmidt product, so

    chi[m, n] = < P_m (x) P_n^T , M >  =  sum_ij conj(P_m (x) P_n^T)[i, j] M[i, j]

-- one contraction, no pseudo-inverse of a d^4-sized array.  The process matrix is the
least-squares solution of ``M A = B`` (columns of A, B: flattened inputs, outputs).

One quirk of the reference is kept, because parity comes first, and can be switched
off.  Its ``lambda_inv`` is indexed (m, n, input basis state, output component) but is
contracted with the process matrix, which is indexed (output component, input basis
state) (tomography.py:109-125): the chi it returns is the chi of the TRANSPOSED map.
For maps with a symmetric process matrix (Pauli channels, real symmetric unitaries) the
two coincide; for a general channel the reference's Kraus operators do not reproduce
the channel (tests/test_tomography.py measures 0.15-0.38 on random channels).
``faithful=False`` (the default of the functions that mirror the reference) reproduces
the reference's numbers; ``faithful=True`` gives the chi and Kraus operators of the map
itself, and is what ``circuit_kraus`` uses.

Everything here is small dense host algebra (4^N x 4^N); the heavy part of a tomography
run is evaluating the process on the 4^N probe states, and ``circuit_process`` does that
on the GPU through ``Simulator`` (density-matrix path, noise channels included).
"""
from __future__ import annotations

from itertools import product

import numpy as np

from . import numpy_quantum as npq


# ---- channels from Kraus operators (tomography.py:14-41) ------------------------------------
def quantum_channel(Ks, *, ket_input: bool = False, return_input: bool = False, normalise: bool = False):
    """``rho -> sum_i [d_i] K_i rho K_i^dagger`` as a callable.  ``Ks`` is a list of
    full-size Kraus operators or a ``(weights, operators)`` pair.  Options as in the
    reference: kets as input, ``(input, output)`` pairs as output, trace renormalisation."""
    if isinstance(Ks, tuple) and len(Ks) == 2 and isinstance(Ks[1], list):
        weights, opers = list(Ks[0]), [np.asarray(k) for k in Ks[1]]
    else:
        opers = [np.asarray(k) for k in Ks]
        weights = [1.0] * len(opers)

    def channel(state):
        rho = npq.ket2dm(state) if ket_input else state
        out = 0
        for w, k in zip(weights, opers):
            out = out + w * (k @ rho @ npq.dagger(k))
        if normalise:
            out = npq.normalise(out)
        return (rho, out) if return_input else out

    return channel


# ---- bases ---------------------------------------------------------------------------------------
def _computational_kets(N: int) -> list[np.ndarray]:
    dim = 2 ** N
    return [np.eye(dim, dtype=complex)[i] for i in range(dim)]


def state_basis(N: int) -> list[np.ndarray]:
    """|n><m| for all computational n, m, n-major: the basis row-major flattening refers to."""
    kets = _computational_kets(N)
    return [np.outer(n, m) for n in kets for m in kets]


def pure_state_basis_kets(N: int) -> list[np.ndarray]:
    """4^N kets whose projectors span the operator space: the computational kets, then for
    every pair n < m the superpositions (n + m)/sqrt2 and (n + i m)/sqrt2."""
    kets = _computational_kets(N)
    out = list(kets)
    for a in range(len(kets)):
        for b in range(a + 1, len(kets)):
            out.append((kets[a] + kets[b]) / np.sqrt(2))
            out.append((kets[a] + 1j * kets[b]) / np.sqrt(2))
    return out


def operator_basis(N: int) -> list[np.ndarray]:
    """Pauli strings over N qubits (I, X, Y, Z; first qubit slowest), each factor / sqrt2."""
    single = [np.asarray(p, dtype=complex) / np.sqrt(2) for p in (npq.IDTY, npq.X, npq.Y, npq.Z)]
    out = []
    for combo in product(range(4), repeat=N):
        m = np.ones((1, 1), dtype=complex)
        for c in combo:
            m = np.kron(m, single[c])
        out.append(m)
    return out


# ---- process matrix, chi matrix, Kraus operators ----------------------------------------------------
def process_matrix(inputs: list[np.ndarray], outputs: list[np.ndarray]) -> np.ndarray:
    """Superoperator ``M`` with ``vec(out) = M vec(in)`` from sampled (in, out) density
    matrices; least squares when oversampled, ``ValueError`` when the inputs do not span
    the operator space (tomography.py:78-106)."""
    if len(inputs) != len(outputs):
        raise ValueError("Inconsistent number of inputs to outputs.")
    A = np.stack([np.asarray(r).reshape(-1) for r in inputs], axis=1)
    B = np.stack([np.asarray(r).reshape(-1) for r in outputs], axis=1)
    sing = np.linalg.svd(A, compute_uv=False)
    cutoff = max(A.shape) * np.finfo(A.dtype).eps * sing.max()
    if int(np.sum(sing > cutoff)) < A.shape[0] or A.shape[1] < A.shape[0]:
        raise ValueError("Insufficiently sampled input set.")
    # M A = B  <=>  A^T M^T = B^T
    Mt, *_ = np.linalg.lstsq(A.T, B.T, rcond=None)
    return Mt.T


def _chi_basis(N: int) -> np.ndarray:
    """B[m, n] = P_m (x) P_n^T, shape (d, d, d, d) with d = 4^N."""
    ops = np.stack(operator_basis(N))
    d = ops.shape[0]
    # (P (x) Q)[(a, b), (c, d)] = P[a, c] Q[b, d]  with  Q = P_n^T
    return np.einsum("mac,ndb->mnabcd", ops, ops).reshape(d, d, d, d)


def chi_matrix(process_matrix: np.ndarray, N: int, *, strict: bool = False, faithful: bool = False) -> np.ndarray:
    """chi with ``E(rho) = sum_mn chi[m, n] P_m rho P_n`` (tomography.py:124-142) -- of the
    map itself with ``faithful``, of the transposed map otherwise (the reference's result,
    see the module docstring).  ``strict`` raises unless chi is Hermitian and satisfies the
    reference's trace-preservation identity."""
    basis = _chi_basis(N)
    M = np.asarray(process_matrix)
    chi = np.einsum("mnij,ij->mn", np.conjugate(basis), M if faithful else M.T)
    if strict:
        if not np.allclose(chi, chi.conj().T):
            raise ValueError("Chi matrix not trace preserving (TP)")
        ops = operator_basis(N)
        total = sum(chi[n, m] * (ops[m] @ ops[n]) for n in range(len(ops)) for m in range(len(ops)))
        if not np.allclose(total, np.eye(total.shape[0])):
            raise ValueError("Chi matrix not trace preserving (TP)")
    return chi


def krauss_operators(chi: np.ndarray, N: int) -> tuple[np.ndarray, list[np.ndarray]]:
    """Eigen-decomposition of a Hermitian chi: eigenvalues D and unit-norm operators K_k =
    sum_m U[m, k] P_m, so that ``E(rho) = sum_k D_k K_k rho K_k^dagger`` (tomography.py:145-156)."""
    vals, vecs = np.linalg.eigh(chi)
    ops = np.stack(operator_basis(N))
    return vals, [np.tensordot(vecs[:, k], ops, axes=(0, 0)) for k in range(vecs.shape[1])]


def eval_process(process, N: int, ket_input: bool):
    """Run ``process`` on the probe states; it returns ``(input rho, output rho)`` pairs
    (tomography.py:163-171)."""
    inputs, outputs = [], []
    for ket in pure_state_basis_kets(N):
        rho_in, rho_out = process(ket) if ket_input else process(npq.ket2dm(ket))
        inputs.append(rho_in)
        outputs.append(rho_out)
    return inputs, outputs


def process_tomography(process, N: int, *, ket_input: bool = True, normalised: bool = False,
                       full_output: bool = False, strict: bool = False, cutoff: float = 1e-12,
                       faithful: bool = False):
    """Kraus operators of an N-qubit CPTP ``process`` (tomography.py:187-215): ``[sqrt(D_k) K_k]``
    for the eigenvalues above ``cutoff``, or ``(D, [K_k])`` with ``normalised``.  ``faithful``:
    see ``chi_matrix``."""
    M = process_matrix(*eval_process(process, N, ket_input))
    chi = chi_matrix(M, N, strict=strict, faithful=faithful)
    if not np.allclose(chi, npq.dagger(chi)):
        raise ValueError("Process is not a CPTP map!")
    vals, opers = krauss_operators(chi, N)
    if not full_output:
        keep = vals > cutoff
        vals = vals[keep]
        opers = [k for k, flag in zip(opers, keep) if flag]
    if normalised:
        return vals, opers
    return [np.sqrt(v) * k for v, k in zip(vals, opers)]


# ---- the bridge to the engine ---------------------------------------------------------------------
def superoperator(Ks) -> np.ndarray:
    """Row-major process matrix ``sum_i K_i (x) conj(K_i)`` of a Kraus set."""
    return sum(np.kron(np.asarray(k), np.conjugate(np.asarray(k))) for k in Ks)


def circuit_process(circuit, num_qubits: int, *, noise=None, backend=None):
    """A tomography ``process`` (kets in, ``(rho_in, rho_out)`` out) that runs ``circuit`` on
    the device through ``Simulator`` as a density-matrix simulation; with ``noise`` (a
    ``channels.GKPNoise``) every gate is followed by its noise channels, which is the
    per-gate channel the paper fits."""
    from .simulator import Simulator
    gates = []
    for g in circuit:
        gates.append(g)
        if noise is not None:
            gates.extend(noise.channels_after(g))
    sim = Simulator(gates, backend=backend)

    def process(ket):
        rho = npq.ket2dm(np.asarray(ket, dtype=np.complex128))
        if rho.shape != (2 ** num_qubits, 2 ** num_qubits):
            raise ValueError("probe state has the wrong size")
        return rho, np.asarray(sim.run(rho), dtype=np.complex128)

    return process


def circuit_kraus(circuit, num_qubits: int, *, noise=None, backend=None, cutoff: float = 1e-12):
    """Kraus operators of what ``circuit`` (with ``noise``) does, fitted from a device
    simulation of the 4^N probe states: ``sum_k K_k rho K_k^dagger`` reproduces the
    simulated map."""
    process = circuit_process(circuit, num_qubits, noise=noise, backend=backend)
    return process_tomography(process, num_qubits, faithful=True, cutoff=cutoff)
