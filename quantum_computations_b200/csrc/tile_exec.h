// Per-thread phase functions of the tile pass.  The same code is compiled for
// the device (called from k_tile_pass) and for the host emulator (tests/: every
// phase is looped over all thread ids), so the index arithmetic that the CPU
// tests exercise is the arithmetic the GPU runs.
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

// Shared-memory slot (16 B) of local index j: TMA's 128-byte swizzle, i.e. the slot
// number inside a 128-byte row (bits 0..2) XOR the row number modulo 8 (bits 3..5).
// A 128-bit access is served 8 lanes at a time, so 8 lanes are conflict-free iff their
// slots differ mod 8: any three index bits among positions 0..5 whose positions differ
// mod 3 enumerate all 8 residues; the planner orders QsStep::fpos so that the three
// fastest thread bits land on such positions.  The map is linear over GF(2):
// swz(a ^ b) == swz(a) ^ swz(b).
QS_HD uint32_t qs_swz(uint32_t j) { return j ^ ((j >> 3) & 7u); }

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

// Q(x) of the final layer: every coupled pair inside x counted once.
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

// ---- per-launch thread table (tile independent) -----------------------------------------
// Entry e (0..47) of a step's table: 0..15 -> jA, 16..47 -> jB.
#define QS_TAB_ENTRIES 48
QS_HD void qs_build_step_tab(const QsPass& P, int s, int e, QsStepTab* tab, uint32_t nthr_log2) {
  const QsStep& st = P.steps[s];
  const int nfree = (int)P.T - st.r;
  const int lo_bits = nfree < (int)nthr_log2 ? nfree : (int)nthr_log2;
  if (e < 16) {
    const int c = lo_bits < 4 ? lo_bits : 4;
    tab->jA[e] = (uint16_t)qs_scatter8((uint32_t)e, st.fpos, c);
  } else {
    const int c = lo_bits - 4 < 0 ? 0 : (lo_bits - 4 > 5 ? 5 : lo_bits - 4);
    tab->jB[e - 16] = (uint16_t)qs_scatter8((uint32_t)(e - 16), st.fpos + 4, c);
  }
}
QS_HD uint32_t qs_thread_jlo(const QsStepTab& tab, uint32_t tid) {
  return (uint32_t)tab.jA[tid & 15u] | (uint32_t)tab.jB[(tid >> 4) & 31u];
}

// ---- per-tile sign data ---------------------------------------------------------------------
// Layer l: Z's on the local positions for this tile (local Z gates plus CZ's whose other
// bit is an outer bit that is 1 in this tile), packed as
//   bits 0..15  the Z's of the step's group bits in m-space (bit r-1-f <-> factor f)
//   bits 16..31 the Z's by local position (the final layer needs all of them).
QS_HD uint32_t qs_layer_z(const QsPass& P, int l, uint64_t base) {
  const QsLayer& L = P.layers[l];
  const QsStep& st = P.steps[L.step];
  const uint8_t* pr = P.pairs + 2 * (uint32_t)L.pair_off;
  uint32_t z = L.zconst;
  for (int i = 0; i < L.n_lo; ++i, pr += 2) z ^= (uint32_t)((base >> pr[1]) & 1ull) << pr[0];
  uint32_t zg = 0;
  for (int f = 0; f < st.r; ++f) zg |= ((z >> st.gpos[f]) & 1u) << (st.r - 1 - f);
  return zg | (z << 16);
}
// Final layer: tile-uniform bit g (pairs of outer bits).
QS_HD uint32_t qs_fin_g(const QsPass& P, uint64_t base) {
  const uint8_t* pr = P.pairs + 2 * (uint32_t)P.fin_pair_off;
  uint32_t g = 0;
  for (int i = 0; i < P.fin_n_oo; ++i, pr += 2) g ^= (uint32_t)((base >> pr[0]) & (base >> pr[1]) & 1ull);
  return g;
}

QS_HD void qs_flip(qs_c128& a, uint32_t mask) {
  // bit 31 of `mask` says whether to negate: XOR it into the sign of both components
#if defined(__CUDA_ARCH__)
  a.x = __hiloint2double(__double2hiint(a.x) ^ (int)(mask & 0x80000000u), __double2loint(a.x));
  a.y = __hiloint2double(__double2hiint(a.y) ^ (int)(mask & 0x80000000u), __double2loint(a.y));
#else
  if (mask & 0x80000000u) { a.x = -a.x; a.y = -a.y; }
#endif
}

// ---- plain tile fill / drain (host emulator; on the device only for tiny states and tile
//      sets that TMA cannot describe) ---------------------------------------------------------
QS_HD void qs_plain_load(const QsPass& P, const qs_c128* state, qs_c128* tile, uint64_t base, uint32_t tid,
                         uint32_t nthr) {
  for (uint32_t j = tid; j < (1u << P.T); j += nthr)
    tile[qs_swz(j)] = state[base | qs_scatter64(j, P.tile_bits, (int)P.T)];
}
QS_HD void qs_plain_store(const QsPass& P, qs_c128* state, const qs_c128* tile, uint64_t base, uint32_t tid,
                          uint32_t nthr) {
  for (uint32_t j = tid; j < (1u << P.T); j += nthr)
    state[base | qs_scatter64(j, P.tile_bits, (int)P.T)] = tile[qs_swz(j)];
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

QS_HD double qs_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return a * b + c;
#endif
}

// Rotations as shears: every operation updates one amplitude in place from the other, so
// the compiler needs no temporaries (the amplitudes live in fixed registers across the
// layer loop; a textbook 2x2 update costs one register move per FMA there).
//
// QS_FORM_TAN:  [[1, -t], [t, 1]] = [[1, 0], [t, 1]] . [[1, -u], [0, 1]] . diag(1, 1 + t^2),
//               u = t / (1 + t^2); the diagonal is folded into the layer's phase table.
QS_HD void qs_rot_tan(double u, double t, qs_c128& a0, qs_c128& a1) {
  a0.x = qs_fma(-u, a1.x, a0.x);
  a0.y = qs_fma(-u, a1.y, a0.y);
  a1.x = qs_fma(t, a0.x, a1.x);
  a1.y = qs_fma(t, a0.y, a1.y);
}
// QS_FORM_SHEAR3:  [[c, -s], [s, c]] = [[1, -p], [0, 1]] . [[1, 0], [q, 1]] . [[1, -p], [0, 1]],
//               p = tan(theta / 2), q = sin(theta): exact, no scale, any angle up to 90 degrees
QS_HD void qs_rot_shear3(double p, double q, qs_c128& a0, qs_c128& a1) {
  a0.x = qs_fma(-p, a1.x, a0.x);
  a0.y = qs_fma(-p, a1.y, a0.y);
  a1.x = qs_fma(q, a0.x, a1.x);
  a1.y = qs_fma(q, a0.y, a1.y);
  a0.x = qs_fma(-p, a1.x, a0.x);
  a0.y = qs_fma(-p, a1.y, a0.y);
}

// The 16 bytes of a layer that the step loop needs, fetched as one 128-bit (uniform) load.
struct QsLayerHot { uint32_t head, offs, ngp0, ngp1; };
QS_HD QsLayerHot qs_layer_hot(const QsPass& P, int l) {
#if defined(__CUDA_ARCH__)
  const uint4 v = *reinterpret_cast<const uint4*>(&P.layers[l]);
  QsLayerHot h; h.head = v.x; h.offs = v.y; h.ngp0 = v.z; h.ngp1 = v.w;
  return h;
#else
  const QsLayer& L = P.layers[l];
  QsLayerHot h; h.head = L.head; h.offs = (uint32_t)L.coef_off | ((uint32_t)L.ph_off << 16);
  h.ngp0 = L.ngp[0]; h.ngp1 = L.ngp[1];
  return h;
#endif
}
struct alignas(16) qs_d2 { double x, y; };
QS_HD qs_d2 qs_coef2(const QsPass& P, uint32_t off) {      // two doubles at an even offset: one 128-bit load
  return *reinterpret_cast<const qs_d2*>(P.coef + off);
}

// Sign words of the 2^R amplitudes of a work item for a layer: BIT 31 of S[m] says whether
// amplitude m is negated (the lower bits are garbage; qs_flip masks them).  (The pairs
// inside the group depend on m only; the planner folds them into the layer's phase table.)
//   W_f = z_f + parity(j0 & ng[f]);  amplitude m is negated iff parity(m & W) (+ the common bit)
template <int R>
QS_HD void qs_layer_sign(const QsPass& P, const QsLayerHot& H, uint32_t zl, uint32_t j0, uint32_t jlo, uint32_t i,
                         uint32_t fin_g, uint32_t fin_qlo, uint32_t* S) {
  constexpr int NA = 1 << R;
  uint32_t V[R];                       // V[b] bit 31 = bit b of W (bit b of m <-> factor R-1-b)
#pragma unroll
  for (int f = 0; f < R; ++f) {
    const uint32_t ng = (f & 1) ? ((f >> 1 ? H.ngp1 : H.ngp0) >> 16) : (f >> 1 ? H.ngp1 : H.ngp0);
    // j0 has no bits above position 15, so the packed partner of an even factor needs no mask;
    // the Z bit of the factor (bit R-1-f of zl) goes into the same popcount
    const uint32_t t = (j0 & ng) ^ (zl & (1u << (R - 1 - f)));
#if defined(__CUDA_ARCH__)
    V[R - 1 - f] = (uint32_t)__popc(t) << 31;
#else
    V[R - 1 - f] = (uint32_t)__builtin_popcount(t) << 31;
#endif
  }
  uint32_t C = 0u;
  if (H.head & QS_LH_FINAL) {
    // pairs that touch no group bit: g + z.j0 + Q(jlo) + Q(jhi) + B(jlo, jhi)
    const uint32_t c0 = fin_g ^ qs_par(j0 & (zl >> 16)) ^ fin_qlo ^ (((uint32_t)P.fin_qhi >> i) & 1u) ^
                        qs_par(jlo & (uint32_t)P.fin_neigh[i]);
    C = c0 << 31;
  }
#pragma unroll
  for (int m = 0; m < NA; ++m) {
    uint32_t v = C;
#pragma unroll
    for (int b = 0; b < R; ++b)
      if ((m >> b) & 1) v ^= V[b];
    S[m] = v;
  }
}

// ---- phase: one step ------------------------------------------------------------
// R group bits; every work item is the 2^R amplitudes that differ only in them.
// Amplitude m of a work item has local index j0 ^ dep[m]; group factor f is bit
// (R-1-f) of m and sits at local position gpos[f].  zm[l] = qs_layer_z of layer l for
// this tile; fin_g = qs_fin_g; fin_qlo = Q(jlo) of the final layer for this thread.
// ZASM (device only): read zm through a plain 32-bit shared address (inline PTX).  It changes
// nothing but the register allocation: measured on B200 the 16-amplitude instantiation needs it
// to stay under 128 registers without spills (2 CTAs/SM), the 8-amplitude one is 10 % faster
// without it.
template <int R, bool DENSE, bool ZASM>
QS_HD void qs_phase_step(const QsPass& P, int s, qs_c128* tile, uint32_t tid, uint32_t nthr_log2,
                         const uint32_t* zm, uint32_t fin_g, uint32_t fin_qlo, const QsStepTab& tab) {
  const QsStep& st = P.steps[s];
  const uint32_t nwork = 1u << (P.T - R);
  constexpr int NA = 1 << R;

  const uint32_t jlo = qs_thread_jlo(tab, tid);
  const uint32_t slo = qs_swz(jlo);
  char* const t0 = reinterpret_cast<char*>(tile);
#if defined(__CUDA_ARCH__)
  const uint32_t zm_s = ZASM ? (uint32_t)__cvta_generic_to_shared(zm) : 0u;
#endif

  // uniform trip count and no divergent exit, so that the compiler can keep the layer data on
  // the uniform datapath; a tile smaller than the CTA leaves the upper threads idle (they
  // recompute a lower thread's work item -- the thread table ignores their high id bits -- and
  // do not store)
  uint32_t niter = nwork >> nthr_log2;
  const bool active = niter != 0 || tid < nwork;
  if (niter == 0) niter = 1;
  const int l_end = (int)st.layer0 + (int)st.nlayers;
  for (uint32_t i = 0; i < niter; ++i) {
    const uint32_t hi = st.hi[i];
    const uint32_t j0 = jlo | (hi & 0xffffu);
    const uint32_t s0b = (slo ^ (hi >> 16)) << 4;
    qs_c128 a[NA];
#pragma unroll
    for (int m = 0; m < NA; ++m) a[m] = *reinterpret_cast<const qs_c128*>(t0 + (s0b ^ st.sdepb[m]));
    bool stored = false;
#pragma unroll 1
    for (int l = (int)st.layer0; l < l_end; ++l) {
      const QsLayerHot H = qs_layer_hot(P, l);
      const uint32_t head = H.head;
      if (!DENSE && (head & 0xffu) == (QS_LH_SIGN | QS_LH_PHASE)) {
        // By far the most frequent layer of a rotation step (sign block, phase table, two-shear
        // rotations; no final part): one test up front instead of one per part -- the branch
        // latencies of the per-part tests are a visible share of a layer at three warps per scheduler.
        uint32_t zl;
#if defined(__CUDA_ARCH__)
        if (ZASM) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(zl) : "r"(zm_s + 4u * (uint32_t)l));
        else zl = zm[l];
#else
        zl = zm[l];
#endif
        uint32_t S[NA];
        qs_layer_sign<R>(P, H, zl, j0, jlo, i, fin_g, fin_qlo, S);
#pragma unroll
        for (int m = 0; m < NA; ++m) qs_flip(a[m], S[m]);
        const uint32_t ph_off = H.offs >> 16;
#pragma unroll
        for (int m = 0; m < NA; ++m) {
          const qs_d2 ph = qs_coef2(P, ph_off + 2 * m);
          const double t1 = ph.y * a[m].y, t2 = ph.y * a[m].x;
          a[m].x = qs_fma(ph.x, a[m].x, -t1);
          a[m].y = qs_fma(ph.x, a[m].y, t2);
        }
        const uint32_t coef_off = H.offs & 0xffffu;
#pragma unroll
        for (int f = 0; f < R; ++f) {
          const int bit = 1 << (R - 1 - f);
          if (head & QS_LH_TAN(f)) {
            const qs_d2 cf = qs_coef2(P, coef_off + 2 * f);
#pragma unroll
            for (int m = 0; m < NA; ++m)
              if (!(m & bit)) qs_rot_tan(cf.x, cf.y, a[m], a[m | bit]);
          }
        }
        continue;
      }
      if (head & QS_LH_SIGN) {
        uint32_t zl;
#if defined(__CUDA_ARCH__)
        if (ZASM) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(zl) : "r"(zm_s + 4u * (uint32_t)l));
        else zl = zm[l];
#else
        zl = zm[l];
#endif
        uint32_t S[NA];
        qs_layer_sign<R>(P, H, zl, j0, jlo, i, fin_g, fin_qlo, S);
#pragma unroll
        for (int m = 0; m < NA; ++m) qs_flip(a[m], S[m]);
      }
      if (DENSE && R == 4 && (head & QS_LH_PAIR)) {
        // two dense 4x4 blocks in one round trip: A on factors (0,1) = bits 3,2 of m, then B on
        // factors (2,3) = bits 1,0 (density-matrix channels: one block per (q, q+N) pair)
        const double* const matA = P.coef + (H.offs & 0xffffu);
        const double* const matB = matA + 32;
        uint32_t S2[NA];
#pragma unroll
        for (int m = 0; m < NA; ++m) S2[m] = 0u;
        uint32_t cross2 = 0u;
        if (l + 1 < l_end) {                          // the pass's final sign layer
          const QsLayerHot H2 = qs_layer_hot(P, l + 1);
          qs_layer_sign<R>(P, H2, zm[l + 1], j0, jlo, i, fin_g, fin_qlo, S2);
          cross2 = P.layers[l + 1].cross;
        }
        const uint32_t cross1 = P.layers[l].cross;
#pragma unroll
        for (int m = 0; m < NA; ++m) qs_flip(a[m], (cross1 >> m) << 31);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const double* const mat = half ? matB : matA;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            // the four amplitudes of this block application: index e -> m
            qs_c128 in[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) in[e] = a[half ? (g * 4 + e) : (e * 4 + g)];
#pragma unroll
            for (int row = 0; row < 4; ++row) {
              double re = 0.0, im = 0.0;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const double mr = mat[2 * (row * 4 + c)], mi = mat[2 * (row * 4 + c) + 1];
                re += mr * in[c].x - mi * in[c].y;
                im += mr * in[c].y + mi * in[c].x;
              }
              qs_c128& o = a[half ? (g * 4 + row) : (row * 4 + g)];
              o.x = re; o.y = im;
            }
          }
        }
#pragma unroll
        for (int m = 0; m < NA; ++m) qs_flip(a[m], S2[m] ^ ((cross2 >> m) << 31));
        if (active) {
#pragma unroll
          for (int m = 0; m < NA; ++m) *reinterpret_cast<qs_c128*>(t0 + (s0b ^ st.sdepb[m])) = a[m];
        }
        stored = true;
        break;
      }
      if (DENSE && (head & QS_LH_DENSE)) {
        // dense 2^R x 2^R matrix: inputs are all in registers, so rows can be
        // written back one at a time (4x4: unrolled, the coefficients become uniform operands;
        // larger: a rolled loop keeps the code small)
        constexpr int ROW_UNROLL = R <= 2 ? NA : 1;
        const double* mat = P.coef + (H.offs & 0xffffu);
        // the only layer that may follow a dense one is the pass's final sign layer
        uint32_t S2[NA];
#pragma unroll
        for (int m = 0; m < NA; ++m) S2[m] = 0u;
        if (l + 1 < l_end) {
          const QsLayerHot H2 = qs_layer_hot(P, l + 1);
          qs_layer_sign<R>(P, H2, zm[l + 1], j0, jlo, i, fin_g, fin_qlo, S2);
        }
        uint32_t sg2 = 0u;               // bit `row` set: negate that output (a rolled loop cannot index S2)
#pragma unroll
        for (int m = 0; m < NA; ++m) sg2 |= (S2[m] >> 31) << m;
#pragma unroll ROW_UNROLL
        for (int row = 0; row < NA; ++row) {
          double re = 0.0, im = 0.0;
#pragma unroll
          for (int c = 0; c < NA; ++c) {
            const double mr = mat[2 * (row * NA + c)], mi = mat[2 * (row * NA + c) + 1];
            re += mr * a[c].x - mi * a[c].y;
            im += mr * a[c].y + mi * a[c].x;
          }
          qs_c128 o; o.x = re; o.y = im;
          qs_flip(o, sg2 << (31 - row));
          if (active) *reinterpret_cast<qs_c128*>(t0 + (s0b ^ st.sdepb[row])) = o;
        }
        stored = true;
        break;                               // a dense layer is the last gate layer of its step
      }
      if (head & QS_LH_PHASE) {
        const uint32_t ph_off = H.offs >> 16;
#pragma unroll
        for (int m = 0; m < NA; ++m) {
          // two products into temporaries, then both components updated in place
          const qs_d2 ph = qs_coef2(P, ph_off + 2 * m);
          const double t1 = ph.y * a[m].y, t2 = ph.y * a[m].x;
          a[m].x = qs_fma(ph.x, a[m].x, -t1);
          a[m].y = qs_fma(ph.x, a[m].y, t2);
        }
      }
      if (!(head & (QS_LH_GENERAL | QS_LH_SHEAR3ANY))) {
        // the common case: two-shear rotations only -- one test per factor
        const uint32_t coef_off = H.offs & 0xffffu;
#pragma unroll
        for (int f = 0; f < R; ++f) {
          const int bit = 1 << (R - 1 - f);
          if (head & QS_LH_TAN(f)) {
            const qs_d2 cf = qs_coef2(P, coef_off + 2 * f);
#pragma unroll
            for (int m = 0; m < NA; ++m)
              if (!(m & bit)) qs_rot_tan(cf.x, cf.y, a[m], a[m | bit]);
          }
        }
      } else if (!(head & QS_LH_GENERAL)) {
        const uint32_t coef_off = H.offs & 0xffffu;
#pragma unroll
        for (int f = 0; f < R; ++f) {
          const int bit = 1 << (R - 1 - f);
          if (head & QS_LH_TAN(f)) {
            const qs_d2 cf = qs_coef2(P, coef_off + 2 * f);
#pragma unroll
            for (int m = 0; m < NA; ++m)
              if (!(m & bit)) qs_rot_tan(cf.x, cf.y, a[m], a[m | bit]);
          } else if (head & QS_LH_SHEAR3(f)) {
            const qs_d2 cf = qs_coef2(P, coef_off + 2 * f);
#pragma unroll
            for (int m = 0; m < NA; ++m)
              if (!(m & bit)) qs_rot_shear3(cf.x, cf.y, a[m], a[m | bit]);
          }
        }
      } else {
        // general complex 2x2 per factor (non-unitary user matrices): rare, keep it small
        const uint32_t coef_off = H.offs & 0xffffu;
#pragma unroll 1
        for (int f = 0; f < R; ++f) {
          if (!(head & QS_LH_FULL(f))) continue;
          const double* mat = P.coef + coef_off + 8 * f;
          // pairs along factor f (f is a run-time value here): enumerate with a switch so
          // that the amplitude indices stay compile-time constants
          if (R >= 1 && f == 0) {
#pragma unroll
            for (int m = 0; m < NA; ++m)
              if (!(m & (1 << (R - 1)))) qs_mat2(mat, a[m], a[m | (1 << (R - 1))]);
          } else if (R >= 2 && f == 1) {
#pragma unroll
            for (int m = 0; m < NA; ++m)
              if (!(m & (1 << (R >= 2 ? R - 2 : 0)))) qs_mat2(mat, a[m], a[m | (1 << (R >= 2 ? R - 2 : 0))]);
          } else if (R >= 3 && f == 2) {
#pragma unroll
            for (int m = 0; m < NA; ++m)
              if (!(m & (1 << (R >= 3 ? R - 3 : 0)))) qs_mat2(mat, a[m], a[m | (1 << (R >= 3 ? R - 3 : 0))]);
          } else if (R >= 4 && f == 3) {
#pragma unroll
            for (int m = 0; m < NA; ++m)
              if (!(m & 1)) qs_mat2(mat, a[m], a[m | 1]);
          }
        }
      }
    }
    if (!stored && active) {
#pragma unroll
      for (int m = 0; m < NA; ++m) *reinterpret_cast<qs_c128*>(t0 + (s0b ^ st.sdepb[m])) = a[m];
    }
  }
}

// A dense step (its first layer is a QS_LAYER_DENSE layer) runs OUT OF LINE on the device: with
// the dense bodies inlined next to the rotation bodies ptxas takes the rotation steps' layer loop
// off the uniform datapath (a kernel that can take dense layers ran 24 % slower on a circuit
// without any).
#if defined(__CUDACC__)
#define QS_NOINLINE_DEVICE __noinline__
#else
#define QS_NOINLINE_DEVICE
#endif
template <int MAXR>
QS_NOINLINE_DEVICE
#if defined(__CUDACC__)
__host__ __device__
#else
inline
#endif
void qs_dense_step_any(const QsPass& P, int s, qs_c128* tile, uint32_t tid, uint32_t nthr_log2, const uint32_t* zm,
                       uint32_t fin_g, uint32_t fin_qlo, const QsStepTab& tab) {
  const int r = P.steps[s].r;
  if (r == 2) qs_phase_step<2, true, false>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
  else if (r == 3) qs_phase_step<3, true, false>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
  else if (MAXR >= 4 && r == 4)
    qs_phase_step<(MAXR >= 4 ? 4 : 2), true, false>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
}

// MAXR bounds the instantiated group sizes (and with them the register budget of the calling
// kernel).  DENSE: 0 = the pass has no dense layer; 1 = a few dense steps among rotation steps:
// the dense ones run out of line, the rotation steps keep the uniform datapath; 2 = mostly dense
// steps (density-matrix channels): everything inline, the dense bodies get the uniform operands.
template <int MAXR, int DENSE>
QS_HD void qs_phase_step_any(const QsPass& P, int s, qs_c128* tile, uint32_t tid, uint32_t nthr_log2,
                             const uint32_t* zm, uint32_t fin_g, uint32_t fin_qlo, const QsStepTab& tab) {
  constexpr bool ZASM = MAXR >= 4;
  const int r = P.steps[s].r;
  if (DENSE == 2) {
    if (r == 1) qs_phase_step<1, false, ZASM>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
    else if (r == 2) qs_phase_step<2, true, ZASM>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
    else if (r == 3) qs_phase_step<3, true, ZASM>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
    else if (MAXR >= 4 && r == 4)
      qs_phase_step<(MAXR >= 4 ? 4 : 1), true, ZASM>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
    return;
  }
  if (DENSE == 1 && P.layers[P.steps[s].layer0].kind == QS_LAYER_DENSE) {
    qs_dense_step_any<MAXR>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
    return;
  }
  if (r == 1) qs_phase_step<1, false, ZASM>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
  else if (r == 2) qs_phase_step<2, false, ZASM>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
  else if (r == 3) qs_phase_step<3, false, ZASM>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
  else if (MAXR >= 4 && r == 4)
    qs_phase_step<(MAXR >= 4 ? 4 : 1), false, ZASM>(P, s, tile, tid, nthr_log2, zm, fin_g, fin_qlo, tab);
}
