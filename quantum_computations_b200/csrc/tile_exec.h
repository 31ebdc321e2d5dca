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

struct alignas(16) qs_c128 { double x, y; };   // complex128; 16-byte aligned so every access is one 128-bit op

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

// The same through the per-launch table (the deposit is OR-linear in the tile
// number): four lookups instead of a T-step loop.  Tile numbers beyond 24 bits
// fall back to the loop for the excess.
QS_HD uint64_t qs_tile_base_tab(const QsPass& P, const QsIoTab& io, uint64_t tile) {
  uint64_t base = io.base_tab[0][tile & 63u] | io.base_tab[1][(tile >> 6) & 63u] |
                  io.base_tab[2][(tile >> 12) & 63u] | io.base_tab[3][(tile >> 18) & 63u];
  if (tile >> 24) base |= qs_tile_base(P, (tile >> 24) << 24);
  return base;
}

// Q(x) of the final block: every coupled pair inside x counted once.
QS_HD uint32_t qs_fin_quad(const QsPass& P, uint32_t x) {
  uint32_t q = 0;
  for (uint32_t p = 0; p < P.T; ++p)
    q ^= ((x >> p) & 1u) & qs_par(x & P.fin_nsym[p] & ~((2u << p) - 1u));
  return q;
}
QS_HD uint32_t qs_fin_neigh(const QsPass& P, uint32_t x) {
  uint32_t m = 0;
  for (uint32_t p = 0; p < P.T; ++p) m ^= (0u - ((x >> p) & 1u)) & P.fin_nsym[p];
  return m;
}

// ---- per-launch tables (tile independent) -----------------------------------------------
// Entry e (0..80) of a step's table: 0..15 -> jA, 16..47 -> jB, 48..63 -> hi,
// 64..79 -> sdepb, 80 -> the scalar fields.
#define QS_TAB_ENTRIES 81
QS_HD void qs_build_step_tab(const QsPass& P, int s, int e, QsStepTab* tab, uint32_t nthr_log2) {
  const QsStep& st = P.steps[s];
  const int nfree = (int)P.T - st.r;
  const int lo_bits = nfree < (int)nthr_log2 ? nfree : (int)nthr_log2;
  if (e < 16) {
    const int c = lo_bits < 4 ? lo_bits : 4;
    tab->jA[e] = (uint16_t)qs_scatter8((uint32_t)e, st.fpos, c);
  } else if (e < 48) {
    const int c = lo_bits - 4 < 0 ? 0 : (lo_bits - 4 > 5 ? 5 : lo_bits - 4);
    tab->jB[e - 16] = (uint16_t)qs_scatter8((uint32_t)(e - 16), st.fpos + 4, c);
  } else if (e < 64) {
    const uint32_t jhi = qs_scatter8((uint32_t)(e - 48), st.fpos + nthr_log2, nfree - lo_bits);
    tab->hi[e - 48] = jhi | (qs_swz(jhi) << 16);
  } else if (e < 80) {
    // amplitude m: matrix factor f is bit (r-1-f) of m and sits at local position gpos[f]
    const int m = e - 64;
    uint32_t d = 0;
    for (int f = 0; f < st.r; ++f) d |= (uint32_t)((m >> (st.r - 1 - f)) & 1) << st.gpos[f];
    tab->sdepb[m] = qs_swz(d) << 4;
  } else {
    const int r = st.r;
    uint32_t qg = 0;
    bool all_rot = st.kind == QS_STEP_1Q;
    for (int f = 0; f < QS_MAX_R; ++f) {
      tab->ng[f] = f < r ? st.ng[f] : (uint16_t)0;
      tab->gpos[f] = f < r ? st.gpos[f] : (uint8_t)0;
      if (f < r && st.form[f] != QS_FORM_ROT) all_rot = false;
    }
    for (int m = 0; m < (1 << r); ++m) {
      uint32_t q = 0;
      for (int f = 0; f < r; ++f)
        if ((m >> (r - 1 - f)) & 1)
          for (int f2 = f + 1; f2 < r; ++f2)
            if ((m >> (r - 1 - f2)) & 1) q ^= ((uint32_t)st.ng[f] >> st.gpos[f2]) & 1u;
      qg |= q << m;
    }
    tab->qg = (uint16_t)qg;
    tab->all_rot = all_rot ? 1 : 0;
  }
}

QS_HD void qs_build_io_tab(const QsPass& P, uint32_t i, QsIoTab* io, uint32_t nthr_log2) {
  const uint32_t jhi = i << nthr_log2;
  io->gbyte[i] = ((P.T <= nthr_log2) ? 0ull : qs_scatter64(i, P.tile_bits + nthr_log2, (int)(P.T - nthr_log2))) << 4;
  io->sbyte[i] = qs_swz(jhi & ((1u << P.T) - 1u)) << 4;
  io->fin_neigh[i] = 0;
  if (P.fin_has_sign && jhi < (1u << P.T)) io->fin_neigh[i] = (uint16_t)qs_fin_neigh(P, jhi);
}
QS_HD void qs_build_base_tab(const QsPass& P, uint32_t e, QsIoTab* io) {   // e in [0, 256)
  io->base_tab[e >> 6][e & 63u] = qs_tile_base(P, (uint64_t)(e & 63u) << (6 * (e >> 6)));
}
// fin_q is a bit mask over i; build it with one thread (or sequentially on the host)
QS_HD uint64_t qs_build_fin_q(const QsPass& P, uint32_t nthr_log2) {
  uint64_t q = 0;
  if (!P.fin_has_sign) return 0;
  for (uint32_t i = 0; i < QS_MAX_ITER; ++i) {
    const uint32_t jhi = i << nthr_log2;
    if (jhi < (1u << P.T)) q |= (uint64_t)qs_fin_quad(P, jhi) << i;
  }
  return q;
}

// ---- per-tile sign data ---------------------------------------------------------------------
// Step s: Z's on the group bits for this tile, in m-space (bit r-1-f <-> factor f):
// local Z gates plus CZ's whose other bit is an outer bit that is 1 in this tile.
QS_HD uint32_t qs_step_zg(const QsPass& P, int s, uint64_t base) {
  const QsStep& st = P.steps[s];
  const uint8_t* pr = P.pairs + 2 * (uint32_t)st.pair_off;
  uint32_t z = st.zconst;
  for (int i = 0; i < st.n_lo; ++i, pr += 2) z ^= (uint32_t)((base >> pr[1]) & 1ull) << pr[0];
  uint32_t zg = 0;
  for (int f = 0; f < st.r; ++f) zg |= ((z >> st.gpos[f]) & 1u) << (st.r - 1 - f);
  return zg;
}
// Final block: tile-uniform bit g and linear mask z.
QS_HD void qs_fin_prepare(const QsPass& P, uint64_t base, uint32_t* zmask, uint32_t* gsign) {
  const uint8_t* pr = P.pairs + 2 * (uint32_t)P.fin_pair_off;
  uint32_t g = 0, z = P.fin_zconst;
  for (int i = 0; i < P.fin_n_oo; ++i, pr += 2) g ^= (uint32_t)((base >> pr[0]) & (base >> pr[1]) & 1ull);
  for (int i = 0; i < P.fin_n_lo; ++i, pr += 2) z ^= (uint32_t)((base >> pr[1]) & 1ull) << pr[0];
  *zmask = z;
  *gsign = g;
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
// `copy(dst, src)` moves one amplitude: a plain assignment on the host, a 16-byte
// cp.async on the device (so the next tile streams in while this one computes).
template <class Copy>
QS_HD void qs_phase_load(const QsPass& P, const qs_c128* state, qs_c128* tile, uint64_t base, uint32_t tid,
                         uint32_t nthr_log2, uint64_t glo, const QsIoTab& io, Copy copy) {
  const uint32_t niter = P.T > nthr_log2 ? 1u << (P.T - nthr_log2) : (tid < (1u << P.T) ? 1u : 0u);
  const uint32_t slob = qs_swz(tid) << 4;
  const char* g0 = reinterpret_cast<const char*>(state) + ((base | glo) << 4);   // disjoint bits: | == +
  char* t0 = reinterpret_cast<char*>(tile);
  for (uint32_t i = 0; i < niter; ++i) copy(t0 + (slob ^ io.sbyte[i]), g0 + io.gbyte[i]);
}

// ---- phase: shared -> global, with the pass's final sign block ---------------
// fin_qlo = Q(tid) of the final block (tile independent, computed once per launch).
QS_HD void qs_phase_store(const QsPass& P, qs_c128* state, const qs_c128* tile, uint64_t base, uint32_t tid,
                          uint32_t nthr_log2, uint64_t glo, const QsIoTab& io, uint32_t fin_qlo,
                          uint32_t zmask, uint32_t gsign) {
  const uint32_t niter = P.T > nthr_log2 ? 1u << (P.T - nthr_log2) : (tid < (1u << P.T) ? 1u : 0u);
  const uint32_t slob = qs_swz(tid) << 4;
  char* g0 = reinterpret_cast<char*>(state) + ((base | glo) << 4);
  const char* t0 = reinterpret_cast<const char*>(tile);
  if (!P.fin_has_sign) {
    for (uint32_t i = 0; i < niter; ++i)
      *reinterpret_cast<qs_c128*>(g0 + io.gbyte[i]) = *reinterpret_cast<const qs_c128*>(t0 + (slob ^ io.sbyte[i]));
    return;
  }
  // j = tid | (i << nthr_log2):  g + z.j + Q(tid) + Q(jhi) + B(tid, jhi)
  const uint32_t qlo = gsign ^ fin_qlo ^ qs_par(tid & zmask);
  const uint32_t zhi = zmask >> nthr_log2;
  for (uint32_t i = 0; i < niter; ++i) {
    const uint32_t q = qlo ^ qs_par(i & zhi) ^ (uint32_t)((io.fin_q >> i) & 1ull) ^ qs_par(tid & io.fin_neigh[i]);
    qs_c128 v = *reinterpret_cast<const qs_c128*>(t0 + (slob ^ io.sbyte[i]));
    qs_flip(v, q << 31);
    *reinterpret_cast<qs_c128*>(g0 + io.gbyte[i]) = v;
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

// diagonal / antidiagonal 2x2: two complex multiplies instead of four
QS_HD void qs_mat2_diag(const double* __restrict__ m, qs_c128& a0, qs_c128& a1) {
  const double x0 = a0.x, y0 = a0.y, x1 = a1.x, y1 = a1.y;
  a0.x = m[0] * x0 - m[1] * y0;
  a0.y = m[0] * y0 + m[1] * x0;
  a1.x = m[6] * x1 - m[7] * y1;
  a1.y = m[6] * y1 + m[7] * x1;
}
QS_HD void qs_mat2_anti(const double* __restrict__ m, qs_c128& a0, qs_c128& a1) {
  const double x0 = a0.x, y0 = a0.y, x1 = a1.x, y1 = a1.y;
  a0.x = m[2] * x1 - m[3] * y1;
  a0.y = m[2] * y1 + m[3] * x1;
  a1.x = m[4] * x0 - m[5] * y0;
  a1.y = m[4] * y0 + m[5] * x0;
}

// real rotation [[c, -s], [s, c]] on a pair of complex amplitudes
QS_HD void qs_mat2_rot(const double* __restrict__ m, qs_c128& a0, qs_c128& a1) {
  const double c = m[0], s = m[1];
  const double x0 = a0.x, y0 = a0.y, x1 = a1.x, y1 = a1.y;
  a0.x = c * x0 - s * x1;
  a0.y = c * y0 - s * y1;
  a1.x = s * x0 + c * x1;
  a1.y = s * y0 + c * y1;
}

// ---- phase: one step ------------------------------------------------------------
// R group bits; every work item is the 2^R amplitudes that differ only in them.
// Amplitude m of a work item has local index j0 ^ dep[m]; matrix factor f is bit
// (R-1-f) of m and sits at local position gpos[f].  Sign of amplitude m:
// parity(m & W) + qg(m) with W_f = z_f + parity(j0 & ng[f])   (plan.h).
template <int R, bool DENSE>
QS_HD void qs_phase_step(const QsPass& P, int s, qs_c128* tile, uint32_t tid, uint32_t nthr_log2,
                         uint32_t zg, const QsStepTab& tab, int debug_skip = 0) {
  const QsStep& st = P.steps[s];
  const uint32_t nwork = 1u << (P.T - R);
  const uint32_t nthr = 1u << nthr_log2;
  const bool has_sign = st.has_sign != 0;
  constexpr int NA = 1 << R;

  uint32_t sdb[NA];                        // byte offset of amplitude m inside the tile (swizzled)
#pragma unroll
  for (int m = 0; m < NA; ++m) sdb[m] = tab.sdepb[m];
  uint32_t ng[R];                          // in-tile CZ partners of group factor f
#pragma unroll
  for (int f = 0; f < R; ++f) ng[f] = has_sign ? (uint32_t)tab.ng[f] : 0u;
  const uint32_t qg = tab.qg;              // bit m: pairs inside the group
  const bool all_rot = tab.all_rot != 0;

  const uint32_t jlo = (uint32_t)tab.jA[tid & 15u] | (uint32_t)tab.jB[(tid >> 4) & 31u];
  const uint32_t slo = qs_swz(jlo);
  char* const t0 = reinterpret_cast<char*>(tile);

  // uniform trip count (a tile smaller than the CTA leaves the upper threads idle), so that
  // loop-invariant uniform loads can be hoisted
  uint32_t niter = nwork >> nthr_log2;
  if (niter == 0) {
    if (tid >= nwork) return;
    niter = 1;
  }
  for (uint32_t i = 0; i < niter; ++i) {
    const uint32_t hi = tab.hi[i];
    const uint32_t j0 = jlo | (hi & 0xffffu);
    const uint32_t s0b = (slo ^ (hi >> 16)) << 4;
    qs_c128 a[NA];
#pragma unroll
    for (int m = 0; m < NA; ++m) a[m] = *reinterpret_cast<const qs_c128*>(t0 + (s0b ^ sdb[m]));
    if (has_sign) {
      // W_f = z_f + parity(j0 & ng[f]); sign bit of amplitude m = qg_m ^ parity(m & W)
      uint32_t W = zg;
#pragma unroll
      for (int f = 0; f < R; ++f) W ^= qs_par(j0 & ng[f]) << (R - 1 - f);
      uint32_t sg = qg;
#pragma unroll
      for (int b = 0; b < R; ++b) {
        uint32_t pat = 0;                  // bit m set iff bit b of m is set
#pragma unroll
        for (int m = 0; m < NA; ++m) pat |= (uint32_t)((m >> b) & 1) << m;
        sg ^= (0u - ((W >> b) & 1u)) & pat;
      }
#pragma unroll
      for (int m = 0; m < NA; ++m) qs_flip(a[m], (sg << (31 - m)) & 0x80000000u);
    }
    if (debug_skip & 4) {
      // development: shared-memory round trip without the math
    } else if (all_rot) {
      // branch-free fast path: phase table, then one real rotation per group bit
      const double* ph = P.coef + st.ph_off;
#pragma unroll
      for (int m = 0; m < NA; ++m) {
        const double pr = ph[2 * m], pi = ph[2 * m + 1];
        const double x = a[m].x, y = a[m].y;
        a[m].x = pr * x - pi * y;
        a[m].y = pr * y + pi * x;
      }
#pragma unroll
      for (int f = 0; f < R; ++f) {
        const double* mat = P.coef + st.coef_off + 8 * f;
        const int bit = 1 << (R - 1 - f);
#pragma unroll
        for (int m = 0; m < NA; ++m)
          if (!(m & bit)) qs_mat2_rot(mat, a[m], a[m | bit]);
      }
    } else if (!DENSE || st.kind == QS_STEP_1Q) {
      if (st.has_phase) {
        const double* ph = P.coef + st.ph_off;
#pragma unroll
        for (int m = 0; m < NA; ++m) {
          const double pr = ph[2 * m], pi = ph[2 * m + 1];
          const double x = a[m].x, y = a[m].y;
          a[m].x = pr * x - pi * y;
          a[m].y = pr * y + pi * x;
        }
      }
#pragma unroll 1
      for (int f = 0; f < R; ++f) {        // rolled: the mixed-form path is rare, keep it small
        const double* mat = P.coef + st.coef_off + 8 * f;
        const uint32_t form = st.form[f];
        // pairs along factor f (f is a run-time value here): enumerate with a switch so
        // that the amplitude indices stay compile-time constants
        if (R >= 1 && f == 0) {
#pragma unroll
          for (int m = 0; m < NA; ++m)
            if (!(m & (1 << (R - 1)))) {
              if (form == QS_FORM_GENERAL) qs_mat2(mat, a[m], a[m | (1 << (R - 1))]);
              else if (form == QS_FORM_ROT) qs_mat2_rot(mat, a[m], a[m | (1 << (R - 1))]);
              else if (form == QS_FORM_DIAG) qs_mat2_diag(mat, a[m], a[m | (1 << (R - 1))]);
              else qs_mat2_anti(mat, a[m], a[m | (1 << (R - 1))]);
            }
        } else if (R >= 2 && f == 1) {
#pragma unroll
          for (int m = 0; m < NA; ++m)
            if (!(m & (1 << (R >= 2 ? R - 2 : 0)))) {
              if (form == QS_FORM_GENERAL) qs_mat2(mat, a[m], a[m | (1 << (R >= 2 ? R - 2 : 0))]);
              else if (form == QS_FORM_ROT) qs_mat2_rot(mat, a[m], a[m | (1 << (R >= 2 ? R - 2 : 0))]);
              else if (form == QS_FORM_DIAG) qs_mat2_diag(mat, a[m], a[m | (1 << (R >= 2 ? R - 2 : 0))]);
              else qs_mat2_anti(mat, a[m], a[m | (1 << (R >= 2 ? R - 2 : 0))]);
            }
        } else if (R >= 3 && f == 2) {
#pragma unroll
          for (int m = 0; m < NA; ++m)
            if (!(m & (1 << (R >= 3 ? R - 3 : 0)))) {
              if (form == QS_FORM_GENERAL) qs_mat2(mat, a[m], a[m | (1 << (R >= 3 ? R - 3 : 0))]);
              else if (form == QS_FORM_ROT) qs_mat2_rot(mat, a[m], a[m | (1 << (R >= 3 ? R - 3 : 0))]);
              else if (form == QS_FORM_DIAG) qs_mat2_diag(mat, a[m], a[m | (1 << (R >= 3 ? R - 3 : 0))]);
              else qs_mat2_anti(mat, a[m], a[m | (1 << (R >= 3 ? R - 3 : 0))]);
            }
        } else if (R >= 4 && f == 3) {
#pragma unroll
          for (int m = 0; m < NA; ++m)
            if (!(m & 1)) {
              if (form == QS_FORM_GENERAL) qs_mat2(mat, a[m], a[m | 1]);
              else if (form == QS_FORM_ROT) qs_mat2_rot(mat, a[m], a[m | 1]);
              else if (form == QS_FORM_DIAG) qs_mat2_diag(mat, a[m], a[m | 1]);
              else qs_mat2_anti(mat, a[m], a[m | 1]);
            }
        }
      }
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
        qs_c128 o; o.x = re; o.y = im;
        *reinterpret_cast<qs_c128*>(t0 + (s0b ^ tab.sdepb[row])) = o;
      }
      continue;
    }
#pragma unroll
    for (int m = 0; m < NA; ++m) *reinterpret_cast<qs_c128*>(t0 + (s0b ^ sdb[m])) = a[m];
  }
}

// MAXR bounds the instantiated group sizes (and with them the register budget
// of the calling kernel); DENSE says whether dense (k >= 2) steps may occur.
template <int MAXR, bool DENSE>
QS_HD void qs_phase_step_any(const QsPass& P, int s, qs_c128* tile, uint32_t tid, uint32_t nthr_log2,
                             uint32_t zmask, const QsStepTab& tab, int debug_skip = 0) {
  const int r = P.steps[s].r;
  if (r == 1) qs_phase_step<1, false>(P, s, tile, tid, nthr_log2, zmask, tab, debug_skip);
  else if (r == 2) qs_phase_step<2, DENSE>(P, s, tile, tid, nthr_log2, zmask, tab, debug_skip);
  else if (r == 3) qs_phase_step<3, DENSE>(P, s, tile, tid, nthr_log2, zmask, tab, debug_skip);
  else if (MAXR >= 4 && r == 4)
    qs_phase_step<(MAXR >= 4 ? 4 : 1), DENSE>(P, s, tile, tid, nthr_log2, zmask, tab, debug_skip);
}
