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
