// Element-wise bodies of the non-tiled kernels, shared by the device kernels
// (grid-stride loops in kernels.cu) and the host emulator of the CPU tests.
// Bit convention: reference qubit q of an n-qubit register is index bit n-1-q.
#pragma once
#include "tile_exec.h"

QS_HD qs_c128 qs_cmul(qs_c128 a, qs_c128 b) {
  qs_c128 o;
  o.x = a.x * b.x - a.y * b.y;
  o.y = a.x * b.y + a.y * b.x;
  return o;
}

QS_HD uint64_t qs_insert_bit(uint64_t r, int pos, uint64_t bit) {
  const uint64_t low = r & ((1ull << pos) - 1ull);
  return ((r >> pos) << (pos + 1)) | (bit << pos) | low;
}

// Index of travelling amplitude r of a k-bit exchange: the k bits at positions pos[]
// (ascending) are fixed to val[], the other bits are r's, in order.
struct QsBitSel { int k; int pos[8]; int val[8]; };
QS_HD uint64_t qs_deposit(uint64_t r, const QsBitSel& sel) {
  for (int i = 0; i < sel.k; ++i) r = qs_insert_bit(r, sel.pos[i], (uint64_t)sel.val[i]);
  return r;
}

// parse_state for list[State] (DV/simulator.py:26): amplitude i of the product
// of n single-qubit kets; amps = n x 2 complex, qubit q first.
QS_HD qs_c128 qs_product_amp(const double* amps, int n, uint64_t i) {
  qs_c128 v; v.x = 1.0; v.y = 0.0;
  for (int q = 0; q < n; ++q) {
    const int bit = (int)((i >> (n - 1 - q)) & 1ull);
    qs_c128 a; a.x = amps[4 * q + 2 * bit]; a.y = amps[4 * q + 2 * bit + 1];
    v = qs_cmul(v, a);
  }
  return v;
}

// (I .. bra .. I) psi at reduced index r (DV/gates.py:173-181): bra is NOT conjugated.
QS_HD qs_c128 qs_contract(const qs_c128* in, int pos, uint64_t r, const double* bra) {
  const qs_c128 a0 = in[qs_insert_bit(r, pos, 0)];
  const qs_c128 a1 = in[qs_insert_bit(r, pos, 1)];
  qs_c128 o;
  o.x = bra[0] * a0.x - bra[1] * a0.y + bra[2] * a1.x - bra[3] * a1.y;
  o.y = bra[0] * a0.y + bra[1] * a0.x + bra[2] * a1.y + bra[3] * a1.x;
  return o;
}

// Insert.apply (DV/gates.py:145-153): out index i has the new qubit at bit `pos`.
QS_HD qs_c128 qs_insert_amp(const qs_c128* in, int pos, uint64_t i, const double* amp) {
  const uint64_t bit = (i >> pos) & 1ull;
  const uint64_t low = i & ((1ull << pos) - 1ull);
  const uint64_t r = ((i >> (pos + 1)) << pos) | low;
  qs_c128 a; a.x = amp[2 * bit]; a.y = amp[2 * bit + 1];
  return qs_cmul(in[r], a);
}

// Generic k-qubit matrix, out of place: out[i] = sum_c M[row(i), c] in[i with targets <- c].
// bits[f] = index bit of matrix factor f (f = 0 most significant).
QS_HD qs_c128 qs_generic_amp(const qs_c128* in, const double* mat, const int* bits, int k, uint64_t i) {
  const int dim = 1 << k;
  int row = 0;
  uint64_t cleared = i;
  for (int f = 0; f < k; ++f) {
    row |= (int)((i >> bits[f]) & 1ull) << (k - 1 - f);
    cleared &= ~(1ull << bits[f]);
  }
  double re = 0.0, im = 0.0;
  for (int c = 0; c < dim; ++c) {
    uint64_t src = cleared;
    for (int f = 0; f < k; ++f) src |= (uint64_t)((c >> (k - 1 - f)) & 1) << bits[f];
    const qs_c128 a = in[src];
    const double mr = mat[2 * (row * dim + c)], mi = mat[2 * (row * dim + c) + 1];
    re += mr * a.x - mi * a.y;
    im += mr * a.y + mi * a.x;
  }
  qs_c128 o; o.x = re; o.y = im;
  return o;
}

// One RB sequence (PAPER/randomised_benchmarking.py:65-75, DV part): vec(rho)
// times a dim^2 x dim^2 superoperator per opcode, ideal ket times a dim x dim
// unitary per opcode; then fidelity <psi|rho|psi> and purity tr(rho rho).
template <int DIM>
QS_HD void qs_rb_sequence(const uint16_t* codes, int64_t len, const double* superops,
                          const double* unitaries, const double* rho0, const double* psi0,
                          double* out_fid, double* out_pur, double* out_rho) {
  constexpr int D2 = DIM * DIM;
  qs_c128 rho[D2], psi[DIM];
  for (int e = 0; e < D2; ++e) { rho[e].x = rho0[2 * e]; rho[e].y = rho0[2 * e + 1]; }
  for (int e = 0; e < DIM; ++e) { psi[e].x = psi0[2 * e]; psi[e].y = psi0[2 * e + 1]; }
  for (int64_t t = 0; t < len; ++t) {
    const double* S = superops + (size_t)codes[t] * 2 * D2 * D2;
    const double* U = unitaries + (size_t)codes[t] * 2 * DIM * DIM;
    qs_c128 nr[D2];
    for (int r = 0; r < D2; ++r) {
      double re = 0.0, im = 0.0;
      for (int c = 0; c < D2; ++c) {
        const double mr = S[2 * (r * D2 + c)], mi = S[2 * (r * D2 + c) + 1];
        re += mr * rho[c].x - mi * rho[c].y;
        im += mr * rho[c].y + mi * rho[c].x;
      }
      nr[r].x = re; nr[r].y = im;
    }
    for (int e = 0; e < D2; ++e) rho[e] = nr[e];
    qs_c128 np_[DIM];
    for (int r = 0; r < DIM; ++r) {
      double re = 0.0, im = 0.0;
      for (int c = 0; c < DIM; ++c) {
        const double mr = U[2 * (r * DIM + c)], mi = U[2 * (r * DIM + c) + 1];
        re += mr * psi[c].x - mi * psi[c].y;
        im += mr * psi[c].y + mi * psi[c].x;
      }
      np_[r].x = re; np_[r].y = im;
    }
    for (int e = 0; e < DIM; ++e) psi[e] = np_[e];
  }
  // fidelity = Re sum_ij conj(psi_i) rho_ij psi_j  (npq.fidelity ket/matrix branch)
  double fid = 0.0;
  for (int i = 0; i < DIM; ++i)
    for (int j = 0; j < DIM; ++j) {
      const qs_c128 rp = qs_cmul(rho[i * DIM + j], psi[j]);
      fid += psi[i].x * rp.x + psi[i].y * rp.y;
    }
  // purity = Re tr(rho rho) = Re sum_ij rho_ij rho_ji
  double pur = 0.0;
  for (int i = 0; i < DIM; ++i)
    for (int j = 0; j < DIM; ++j) {
      const qs_c128 a = rho[i * DIM + j], b = rho[j * DIM + i];
      pur += a.x * b.x - a.y * b.y;
    }
  *out_fid = fid;
  *out_pur = pur;
  if (out_rho)
    for (int e = 0; e < D2; ++e) { out_rho[2 * e] = rho[e].x; out_rho[2 * e + 1] = rho[e].y; }
}


// ---- Pauli-trajectory batch (trajectories.py; mechanism of GKP/simulator.py:26-55 at the DV
//      level): one shot = the circuit on a small ket with, after every gate and on each of its
//      qubits, an X flip and then a Z flip where the shot's flip bits say so -------------------
// A gate of the shared circuit: k = 1 or 2 qubits on index bits b0 (matrix factor 0, the
// most significant bit of the row index) and b1; `mat` = offset of its row-major complex
// matrix in the matrix table (in complex numbers).
struct QsTrajOp { int32_t k, b0, b1, mat; };

// Row r of (Z^z X^x M): X on a qubit swaps the rows that differ in its bit, Z negates the
// rows where its bit is 1.  flips = {x0, z0[, x1, z1]} for the gate's qubits in order.
QS_HD void qs_traj_rows(const QsTrajOp& op, const double* mats, const uint8_t* flips, qs_c128* M) {
  const int dim = 1 << op.k;
  int xmask = 0, zmask = 0;
  for (int f = 0; f < op.k; ++f) {
    xmask |= (flips[2 * f] ? 1 : 0) << (op.k - 1 - f);
    zmask |= (flips[2 * f + 1] ? 1 : 0) << (op.k - 1 - f);
  }
  const double* src = mats + 2 * (size_t)op.mat;
  for (int r = 0; r < dim; ++r) {
    const double sgn = (qs_par((uint32_t)(r & zmask)) ? -1.0 : 1.0);
    for (int c = 0; c < dim; ++c) {
      M[r * dim + c].x = sgn * src[2 * ((r ^ xmask) * dim + c)];
      M[r * dim + c].y = sgn * src[2 * ((r ^ xmask) * dim + c) + 1];
    }
  }
}

// Work items tid, tid + nthreads, ... of one gate on the 2^n amplitudes in `state`.
QS_HD void qs_traj_apply(qs_c128* state, int n, const QsTrajOp& op, const qs_c128* M, uint32_t tid, uint32_t nthreads) {
  if (op.k == 1) {
    const uint64_t bit = 1ull << op.b0;
    for (uint64_t w = tid; w < (1ull << (n - 1)); w += nthreads) {
      const uint64_t i0 = qs_insert_bit(w, op.b0, 0);
      const qs_c128 a0 = state[i0], a1 = state[i0 | bit];
      qs_c128 o0, o1;
      o0.x = M[0].x * a0.x - M[0].y * a0.y + M[1].x * a1.x - M[1].y * a1.y;
      o0.y = M[0].x * a0.y + M[0].y * a0.x + M[1].x * a1.y + M[1].y * a1.x;
      o1.x = M[2].x * a0.x - M[2].y * a0.y + M[3].x * a1.x - M[3].y * a1.y;
      o1.y = M[2].x * a0.y + M[2].y * a0.x + M[3].x * a1.y + M[3].y * a1.x;
      state[i0] = o0;
      state[i0 | bit] = o1;
    }
  } else {
    const int lo = op.b0 < op.b1 ? op.b0 : op.b1, hi = op.b0 < op.b1 ? op.b1 : op.b0;
    for (uint64_t w = tid; w < (1ull << (n - 2)); w += nthreads) {
      const uint64_t base = qs_insert_bit(qs_insert_bit(w, lo, 0), hi, 0);
      uint64_t idx[4];
      qs_c128 a[4];
      for (int r = 0; r < 4; ++r) {
        idx[r] = base | ((uint64_t)((r >> 1) & 1) << op.b0) | ((uint64_t)(r & 1) << op.b1);
        a[r] = state[idx[r]];
      }
      for (int r = 0; r < 4; ++r) {
        double re = 0.0, im = 0.0;
        for (int c = 0; c < 4; ++c) {
          re += M[4 * r + c].x * a[c].x - M[4 * r + c].y * a[c].y;
          im += M[4 * r + c].x * a[c].y + M[4 * r + c].y * a[c].x;
        }
        state[idx[r]].x = re;
        state[idx[r]].y = im;
      }
    }
  }
}
