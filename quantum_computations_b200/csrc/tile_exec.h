// Per-thread phase functions of the tile pass.  The same code is compiled for
// the device (called from k_tile_pass with __syncthreads() between phases) and
// for the host emulator (tests/: every phase is looped over all thread ids),
// so the index arithmetic that the CPU tests exercise is the arithmetic the
// GPU runs.
#pragma once
#include "plan.h"

#if defined(__CUDACC__)
#define QS_HD __host__ __device__ __forceinline__
#else
#define QS_HD inline
#endif

struct qs_c128 { double x, y; };   // layout-compatible with double2 / complex128

QS_HD int qs_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}

// Shared-memory slot of local index j: XOR-fold every 3-bit chunk of j into the
// low 3 bits.  A slot is 16 B = 4 banks, a 128-bit access is served 8 lanes at
// a time, so 8 lanes are conflict-free iff their slots differ mod 8.  With the
// fold, any three index bits whose positions differ mod 3 enumerate all 8
// residues; the planner orders QsStep::fpos so the three fastest thread bits
// land on such positions.  The map is linear over GF(2):
// swz(a ^ b) == swz(a) ^ swz(b).
QS_HD uint32_t qs_swz(uint32_t j) {
  return j ^ ((j >> 3) & 7u) ^ ((j >> 6) & 7u) ^ ((j >> 9) & 7u) ^ ((j >> 12) & 7u);
}

// Deposit the low `count` bits of v at positions pos[0..count).
QS_HD uint32_t qs_scatter8(uint32_t v, const uint8_t* pos, int count) {
  uint32_t out = 0;
  for (int b = 0; b < count; ++b) out |= ((v >> b) & 1u) << pos[b];
  return out;
}

QS_HD uint64_t qs_scatter64(uint32_t v, const uint8_t* pos, int count) {
  uint64_t out = 0;
  for (int b = 0; b < count; ++b) out |= (uint64_t)((v >> b) & 1u) << pos[b];
  return out;
}

// Global index of the tile's local index 0: spread the tile number over the
// index bits that are NOT tile bits (tile_bits ascending).
QS_HD uint64_t qs_tile_base(const QsPass& P, uint64_t tile) {
  uint64_t base = tile;
  for (uint32_t l = 0; l < P.T; ++l) {
    const uint32_t pos = P.tile_bits[l];
    const uint64_t low = base & ((1ull << pos) - 1ull);
    base = ((base >> pos) << (pos + 1)) | low;
  }
  return base;
}

// Per tile and per step: the part of the step's sign block that is uniform over
// the tile (gsign) or linear in the local index (zmask).
QS_HD void qs_sign_prepare(const QsPass& P, int s, uint64_t base, uint32_t* zmask, uint32_t* gsign) {
  const QsStep& st = P.steps[s];
  const uint8_t* pr = P.pairs + 2 * (uint32_t)st.pair_off;
  uint32_t g = 0, z = 0;
  for (int i = 0; i < st.n_oo; ++i, pr += 2)
    g ^= (uint32_t)((base >> pr[0]) & (base >> pr[1]) & 1ull);
  for (int i = 0; i < st.n_lo; ++i, pr += 2)
    z ^= (uint32_t)((base >> pr[1]) & 1ull) << pr[0];
  *zmask = z;
  *gsign = g;
}

// Sign (0/1) of local index j under step s's block.
QS_HD uint32_t qs_sign_of(const QsPass& P, const QsStep& st, uint32_t j, uint32_t zmask, uint32_t gsign) {
  uint32_t sg = gsign ^ (uint32_t)(qs_popc(j & zmask) & 1);
  const uint8_t* pr = P.pairs + 2 * ((uint32_t)st.pair_off + st.n_oo + st.n_lo);
  for (int i = 0; i < st.n_ll; ++i, pr += 2) sg ^= (j >> pr[0]) & (j >> pr[1]) & 1u;
  return sg;
}

// ---- phase: global -> shared ------------------------------------------------
QS_HD void qs_phase_load(const QsPass& P, const qs_c128* state, qs_c128* tile,
                         uint64_t base, uint32_t tid, uint32_t nthr_log2) {
  const uint32_t T = P.T;
  const uint32_t nthr = 1u << nthr_log2;
  const uint32_t size = 1u << T;
  const int lo_bits = (int)(T < nthr_log2 ? T : nthr_log2);
  const uint64_t glo = qs_scatter64(tid, P.tile_bits, lo_bits);
  const uint32_t slo = qs_swz(tid);
  for (uint32_t i = 0, j = tid; j < size; ++i, j += nthr) {
    const uint64_t ghi = qs_scatter64(i, P.tile_bits + nthr_log2, (int)(T - lo_bits));
    tile[slo ^ qs_swz(i << nthr_log2)] = state[base | glo | ghi];
  }
}

// ---- phase: shared -> global, with the pass's final sign block ---------------
QS_HD void qs_phase_store(const QsPass& P, qs_c128* state, const qs_c128* tile,
                          uint64_t base, uint32_t tid, uint32_t nthr_log2,
                          uint32_t zmask, uint32_t gsign) {
  const uint32_t T = P.T;
  const uint32_t nthr = 1u << nthr_log2;
  const uint32_t size = 1u << T;
  const QsStep& st = P.steps[P.nsteps - 1];
  const bool has_sign = (st.n_oo | st.n_lo | st.n_ll) != 0;
  const int lo_bits = (int)(T < nthr_log2 ? T : nthr_log2);
  const uint64_t glo = qs_scatter64(tid, P.tile_bits, lo_bits);
  const uint32_t slo = qs_swz(tid);
  for (uint32_t i = 0, j = tid; j < size; ++i, j += nthr) {
    const uint64_t ghi = qs_scatter64(i, P.tile_bits + nthr_log2, (int)(T - lo_bits));
    qs_c128 v = tile[slo ^ qs_swz(i << nthr_log2)];
    if (has_sign && qs_sign_of(P, st, j, zmask, gsign)) { v.x = -v.x; v.y = -v.y; }
    state[base | glo | ghi] = v;
  }
}

// 2x2 complex matrix m (row-major, interleaved re/im: 8 doubles) on (a0, a1).
QS_HD void qs_mat2(const double* __restrict__ m, qs_c128& a0, qs_c128& a1) {
  const double m00r = m[0], m00i = m[1], m01r = m[2], m01i = m[3];
  const double m10r = m[4], m10i = m[5], m11r = m[6], m11i = m[7];
  const double x0 = a0.x, y0 = a0.y, x1 = a1.x, y1 = a1.y;
  a0.x = m00r * x0 - m00i * y0 + m01r * x1 - m01i * y1;
  a0.y = m00r * y0 + m00i * x0 + m01r * y1 + m01i * x1;
  a1.x = m10r * x0 - m10i * y0 + m11r * x1 - m11i * y1;
  a1.y = m10r * y0 + m10i * x0 + m11r * y1 + m11i * x1;
}

// ---- phase: one step ------------------------------------------------------------
// R group bits; every work item is the 2^R amplitudes that differ only in them.
template <int R>
QS_HD void qs_phase_step(const QsPass& P, int s, qs_c128* tile, uint32_t tid, uint32_t nthr_log2,
                         uint32_t zmask, uint32_t gsign) {
  const QsStep& st = P.steps[s];
  const uint32_t T = P.T;
  const uint32_t nfree = T - R;
  const uint32_t nwork = 1u << nfree;
  const uint32_t nthr = 1u << nthr_log2;
  const bool has_sign = (st.n_oo | st.n_lo | st.n_ll) != 0;
  constexpr int NA = 1 << R;

  // local-index offset of amplitude m of a work item (matrix factor f is the
  // (R-1-f)-th bit of m and sits at local position gpos[f])
  uint32_t dep[NA];
#pragma unroll
  for (int m = 0; m < NA; ++m) {
    uint32_t d = 0;
#pragma unroll
    for (int f = 0; f < R; ++f) d |= (uint32_t)((m >> (R - 1 - f)) & 1) << st.gpos[f];
    dep[m] = d;
  }

  const int lo_bits = (int)(nfree < nthr_log2 ? nfree : nthr_log2);
  const uint32_t jlo = qs_scatter8(tid, st.fpos, lo_bits);
  for (uint32_t i = 0, w = tid; w < nwork; ++i, w += nthr) {
    const uint32_t j0 = jlo | qs_scatter8(i, st.fpos + nthr_log2, (int)(nfree - lo_bits));
    const uint32_t s0 = qs_swz(j0);
    qs_c128 a[NA];
#pragma unroll
    for (int m = 0; m < NA; ++m) a[m] = tile[s0 ^ qs_swz(dep[m])];
    if (has_sign) {
#pragma unroll
      for (int m = 0; m < NA; ++m)
        if (qs_sign_of(P, st, j0 | dep[m], zmask, gsign)) { a[m].x = -a[m].x; a[m].y = -a[m].y; }
    }
    if (st.kind == QS_STEP_1Q) {
#pragma unroll
      for (int f = 0; f < R; ++f) {
        const double* mat = P.coef + st.coef_off + 8 * f;
        const int bit = 1 << (R - 1 - f);
#pragma unroll
        for (int m = 0; m < NA; ++m)
          if (!(m & bit)) qs_mat2(mat, a[m], a[m | bit]);
      }
    } else {
      const double* mat = P.coef + st.coef_off;
      qs_c128 o[NA];
#pragma unroll
      for (int row = 0; row < NA; ++row) {
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int c = 0; c < NA; ++c) {
          const double mr = mat[2 * (row * NA + c)], mi = mat[2 * (row * NA + c) + 1];
          re += mr * a[c].x - mi * a[c].y;
          im += mr * a[c].y + mi * a[c].x;
        }
        o[row].x = re; o[row].y = im;
      }
#pragma unroll
      for (int m = 0; m < NA; ++m) a[m] = o[m];
    }
#pragma unroll
    for (int m = 0; m < NA; ++m) tile[s0 ^ qs_swz(dep[m])] = a[m];
  }
}

// MAXR bounds the instantiated group sizes (and with them the register budget
// of the calling kernel).
template <int MAXR>
QS_HD void qs_phase_step_any(const QsPass& P, int s, qs_c128* tile, uint32_t tid, uint32_t nthr_log2,
                             uint32_t zmask, uint32_t gsign) {
  const int r = P.steps[s].r;
  if (r == 1) qs_phase_step<1>(P, s, tile, tid, nthr_log2, zmask, gsign);
  else if (r == 2) qs_phase_step<2>(P, s, tile, tid, nthr_log2, zmask, gsign);
  else if (r == 3) qs_phase_step<3>(P, s, tile, tid, nthr_log2, zmask, gsign);
  else if (MAXR >= 4 && r == 4) qs_phase_step<(MAXR >= 4 ? 4 : 1)>(P, s, tile, tid, nthr_log2, zmask, gsign);
}
