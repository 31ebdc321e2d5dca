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

QS_HD uint32_t qs_par(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__popc(v) & 1u;
#else
  return (uint32_t)__builtin_popcount(v) & 1u;
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

// Global-index contribution of the thread id (local bits 0..nthr_log2-1) and of
// the per-thread iteration counter i (the remaining local bits).  Both are the
// same for every tile of a pass, so the kernel computes them once.
QS_HD uint64_t qs_global_lo(const QsPass& P, uint32_t tid, uint32_t nthr_log2) {
  const int lo_bits = (int)(P.T < nthr_log2 ? P.T : nthr_log2);
  return qs_scatter64(tid, P.tile_bits, lo_bits);
}
QS_HD uint64_t qs_global_hi(const QsPass& P, uint32_t i, uint32_t nthr_log2) {
  if (P.T <= nthr_log2) return 0;
  return qs_scatter64(i, P.tile_bits + nthr_log2, (int)(P.T - nthr_log2));
}

// ---- sign blocks ------------------------------------------------------------------------
// Per tile and per step: the tile-uniform bit g and the linear mask z (plan.h).
QS_HD void qs_sign_prepare(const QsPass& P, int s, uint64_t base, uint32_t* zmask, uint32_t* gsign) {
  const QsStep& st = P.steps[s];
  const uint8_t* pr = P.pairs + 2 * (uint32_t)st.pair_off;
  uint32_t g = 0, z = st.zconst;
  for (int i = 0; i < st.n_oo; ++i, pr += 2)
    g ^= (uint32_t)((base >> pr[0]) & (base >> pr[1]) & 1ull);
  for (int i = 0; i < st.n_lo; ++i, pr += 2)
    z ^= (uint32_t)((base >> pr[1]) & 1ull) << pr[0];
  *zmask = z;
  *gsign = g;
}

// Q(x) for x = the scatter of the low `count` bits of v over pos[]: every coupled
// pair inside x counted once (the partner with the higher position).
QS_HD uint32_t qs_quad_scattered(const QsStep& st, uint32_t v, const uint8_t* pos, int count, uint32_t x) {
  uint32_t q = 0;
  for (int b = 0; b < count; ++b) {
    const uint32_t p = pos[b];
    const uint32_t up = ~((2u << p) - 1u);
    q ^= ((v >> b) & 1u) & qs_par(x & st.nsym[p] & up);
  }
  return q;
}

// XOR of the neighbour masks of the set bits of the scattered value.
QS_HD uint32_t qs_neigh_scattered(const QsStep& st, uint32_t v, const uint8_t* pos, int count) {
  uint32_t m = 0;
  for (int b = 0; b < count; ++b) m ^= (0u - ((v >> b) & 1u)) & st.nsym[pos[b]];
  return m;
}

// Reference implementation of the whole sign (used by the host emulator's
// self-check and by nothing on the hot path): g + z.j + Q(j).
QS_HD uint32_t qs_sign_slow(const QsStep& st, uint32_t T, uint32_t j, uint32_t zmask, uint32_t gsign) {
  uint32_t sg = gsign ^ qs_par(j & zmask);
  for (uint32_t p = 0; p < T; ++p)
    if ((j >> p) & 1u) sg ^= qs_par(j & st.nsym[p] & ~((2u << p) - 1u));
  return sg;
}

QS_HD void qs_flip(qs_c128& a, uint32_t sign_bit) {
  // sign_bit is 0 or 0x80000000: XOR it into the sign of both components
#if defined(__CUDA_ARCH__)
  a.x = __hiloint2double(__double2hiint(a.x) ^ (int)sign_bit, __double2loint(a.x));
  a.y = __hiloint2double(__double2hiint(a.y) ^ (int)sign_bit, __double2loint(a.y));
#else
  if (sign_bit) { a.x = -a.x; a.y = -a.y; }
#endif
}

// ---- phase: global -> shared ------------------------------------------------
// ghi[i] = qs_global_hi(P, i, ..) for i < 2^(T - nthr_log2)
QS_HD void qs_phase_load(const QsPass& P, const qs_c128* state, qs_c128* tile, uint64_t base, uint32_t tid,
                         uint32_t nthr_log2, uint64_t glo, const uint64_t* ghi) {
  const uint32_t nthr = 1u << nthr_log2;
  const uint32_t size = 1u << P.T;
  const uint32_t slo = qs_swz(tid);
  const uint64_t b0 = base | glo;
  for (uint32_t i = 0, j = tid; j < size; ++i, j += nthr)
    tile[slo ^ qs_swz(i << nthr_log2)] = state[b0 | ghi[i]];
}

// ---- phase: shared -> global, with the pass's final sign block ---------------
QS_HD void qs_phase_store(const QsPass& P, qs_c128* state, const qs_c128* tile, uint64_t base, uint32_t tid,
                          uint32_t nthr_log2, uint64_t glo, const uint64_t* ghi, uint32_t zmask,
                          uint32_t gsign) {
  const uint32_t T = P.T;
  const uint32_t nthr = 1u << nthr_log2;
  const uint32_t size = 1u << T;
  const QsStep& st = P.steps[P.nsteps - 1];
  const uint32_t slo = qs_swz(tid);
  const uint64_t b0 = base | glo;
  if (!st.has_sign) {
    for (uint32_t i = 0, j = tid; j < size; ++i, j += nthr)
      state[b0 | ghi[i]] = tile[slo ^ qs_swz(i << nthr_log2)];
    return;
  }
  // j = tid | (i << nthr_log2): split the quadratic form accordingly
  const int lo_bits = (int)(T < nthr_log2 ? T : nthr_log2);
  uint32_t qlo = gsign ^ qs_par(tid & zmask);
  for (int b = 0; b < lo_bits; ++b)
    qlo ^= ((tid >> b) & 1u) & qs_par(tid & st.nsym[b] & ~((2u << b) - 1u));
  for (uint32_t i = 0, j = tid; j < size; ++i, j += nthr) {
    const uint32_t jhi = i << nthr_log2;
    uint32_t q = qlo ^ qs_par(jhi & zmask);
    uint32_t neigh = 0;
    for (uint32_t p = nthr_log2; p < T; ++p) {
      const uint32_t on = (jhi >> p) & 1u;
      q ^= on & qs_par(jhi & st.nsym[p] & ~((2u << p) - 1u));
      neigh ^= (0u - on) & st.nsym[p];
    }
    q ^= qs_par(tid & neigh);
    qs_c128 v = tile[slo ^ qs_swz(jhi)];
    qs_flip(v, q << 31);
    state[b0 | ghi[i]] = v;
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
// Amplitude m of a work item has local index j0 ^ dep[m]; matrix factor f is bit
// (R-1-f) of m and sits at local position gpos[f].
//
// Sign of amplitude m (plan.h):  g + z.(j0^dep) + Q(j0^dep)
//   = [g + z.j0 + Q(j0)]  +  [z.dep + Q(dep)]  +  B(j0, dep)
// The first bracket is one bit per work item (and splits again into a per-thread
// and a per-iteration part), the second is a per-step table over m, and the
// bilinear term is parity(m & W) with W_f = parity(j0 & nsym[gpos[f]]).
template <int R, bool DENSE>
QS_HD void qs_phase_step(const QsPass& P, int s, qs_c128* tile, uint32_t tid, uint32_t nthr_log2,
                         uint32_t zmask, uint32_t gsign) {
  const QsStep& st = P.steps[s];
  const uint32_t T = P.T;
  const uint32_t nfree = T - R;
  const uint32_t nwork = 1u << nfree;
  const uint32_t nthr = 1u << nthr_log2;
  const bool has_sign = st.has_sign != 0;
  constexpr int NA = 1 << R;

  uint32_t sdep[NA];                       // swizzled slot offset of amplitude m
  uint32_t ng[R];                          // neighbour mask of group bit f
  uint32_t zg = 0;                         // z restricted to the group bits, in m-space
  uint32_t qg = 0;                         // bit m: Q(dep[m]) (pairs inside the group)
  {
    uint32_t gp[R];
#pragma unroll
    for (int f = 0; f < R; ++f) {
      gp[f] = st.gpos[f];
      ng[f] = has_sign ? (uint32_t)st.nsym[gp[f]] : 0u;
      zg |= ((zmask >> gp[f]) & 1u) << (R - 1 - f);
    }
#pragma unroll
    for (int m = 0; m < NA; ++m) {
      uint32_t d = 0, q = 0;
#pragma unroll
      for (int f = 0; f < R; ++f) {
        if ((m >> (R - 1 - f)) & 1) {
          d |= 1u << gp[f];
#pragma unroll
          for (int f2 = f + 1; f2 < R; ++f2)
            if ((m >> (R - 1 - f2)) & 1) q ^= (ng[f] >> gp[f2]) & 1u;
        }
      }
      sdep[m] = qs_swz(d);
      qg |= q << m;
    }
  }

  const int lo_bits = (int)(nfree < nthr_log2 ? nfree : nthr_log2);
  const int hi_bits = (int)(nfree - lo_bits);
  const uint32_t jlo = qs_scatter8(tid, st.fpos, lo_bits);
  uint32_t alo = 0;
  if (has_sign) alo = gsign ^ qs_par(jlo & zmask) ^ qs_quad_scattered(st, tid, st.fpos, lo_bits, jlo);

  for (uint32_t i = 0, w = tid; w < nwork; ++i, w += nthr) {
    const uint32_t jhi = qs_scatter8(i, st.fpos + nthr_log2, hi_bits);
    const uint32_t j0 = jlo | jhi;
    const uint32_t s0 = qs_swz(j0);
    qs_c128 a[NA];
#pragma unroll
    for (int m = 0; m < NA; ++m) a[m] = tile[s0 ^ sdep[m]];
    if (has_sign) {
      const uint32_t ahi = qs_par(jhi & zmask) ^ qs_quad_scattered(st, i, st.fpos + nthr_log2, hi_bits, jhi);
      const uint32_t cross = qs_par(jlo & qs_neigh_scattered(st, i, st.fpos + nthr_log2, hi_bits));
      const uint32_t A = alo ^ ahi ^ cross;
      uint32_t W = zg;
#pragma unroll
      for (int f = 0; f < R; ++f) W ^= qs_par(j0 & ng[f]) << (R - 1 - f);
      // signs of all 2^R amplitudes at once: bit m = A ^ qg_m ^ parity(m & W)
      uint32_t sg = qg ^ (0u - A);
#pragma unroll
      for (int b = 0; b < R; ++b) {
        uint32_t pat = 0;                  // bit m set iff bit b of m is set
#pragma unroll
        for (int m = 0; m < NA; ++m) pat |= (uint32_t)((m >> b) & 1) << m;
        sg ^= (0u - ((W >> b) & 1u)) & pat;
      }
#pragma unroll
      for (int m = 0; m < NA; ++m) qs_flip(a[m], (sg >> m) << 31);
    }
    if (!DENSE || st.kind == QS_STEP_1Q) {
#pragma unroll
      for (int f = 0; f < R; ++f) {
        const double* mat = P.coef + st.coef_off + 8 * f;
        const int bit = 1 << (R - 1 - f);
#pragma unroll
        for (int m = 0; m < NA; ++m)
          if (!(m & bit)) qs_mat2(mat, a[m], a[m | bit]);
      }
#pragma unroll
      for (int m = 0; m < NA; ++m) tile[s0 ^ sdep[m]] = a[m];
    } else {
      // dense 2^R x 2^R matrix: inputs are all in registers, so rows can be
      // written back one at a time (rolled loop keeps the code small)
      const double* mat = P.coef + st.coef_off;
#pragma unroll 1
      for (int row = 0; row < NA; ++row) {
        double re = 0.0, im = 0.0;
#pragma unroll
        for (int c = 0; c < NA; ++c) {
          const double mr = mat[2 * (row * NA + c)], mi = mat[2 * (row * NA + c) + 1];
          re += mr * a[c].x - mi * a[c].y;
          im += mr * a[c].y + mi * a[c].x;
        }
        uint32_t d = 0;
#pragma unroll
        for (int f = 0; f < R; ++f) d |= (uint32_t)((row >> (R - 1 - f)) & 1) << st.gpos[f];
        qs_c128 o; o.x = re; o.y = im;
        tile[s0 ^ qs_swz(d)] = o;
      }
    }
  }
}

// MAXR bounds the instantiated group sizes (and with them the register budget
// of the calling kernel); DENSE says whether dense (k >= 2) steps may occur.
template <int MAXR, bool DENSE>
QS_HD void qs_phase_step_any(const QsPass& P, int s, qs_c128* tile, uint32_t tid, uint32_t nthr_log2,
                             uint32_t zmask, uint32_t gsign) {
  const int r = P.steps[s].r;
  if (r == 1) qs_phase_step<1, false>(P, s, tile, tid, nthr_log2, zmask, gsign);
  else if (r == 2) qs_phase_step<2, DENSE>(P, s, tile, tid, nthr_log2, zmask, gsign);
  else if (r == 3) qs_phase_step<3, DENSE>(P, s, tile, tid, nthr_log2, zmask, gsign);
  else if (MAXR >= 4 && r == 4) qs_phase_step<(MAXR >= 4 ? 4 : 1), DENSE>(P, s, tile, tid, nthr_log2, zmask, gsign);
}
