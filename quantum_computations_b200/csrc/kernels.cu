// sm_100a kernels and the device-facing half of the C ABI (include/qsim_b200.h).
//
// Everything here is memory-bound complex128 work: 16-byte (128-bit) loads and
// stores per amplitude, shared-memory staging of 2^T-amplitude tiles so several
// gates apply per HBM pass, persistent grids sized to the SM count.  No tensor
// cores: tcgen05 has no f64 kind and the 2x2/4x4 updates have no GEMM shape.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "elem_ops.h"
#include "planner.h"

namespace {

std::atomic<int64_t> g_launches{0};

struct DevCtx {
  bool ready = false;
  int sms = 0;
  double* d_partials = nullptr;   // [kReduceBlocks][2]
  double* d_out = nullptr;        // [2]
  double* h_out = nullptr;        // pinned [2]
  int occ[4] = {0, 0, 0, 0};      // resident CTAs/SM of the k_tile_pass variants at the last smem size
  int occ_smem = -1;
};

constexpr int kMaxDevices = 16;
constexpr int kReduceThreads = 256;
constexpr int kReduceBlocks = 148 * 8;
DevCtx g_ctx[kMaxDevices];

#define QS_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return qs::fail(QSIM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

int bind_device(const void* ptr, DevCtx** ctx) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return qs::fail(QSIM_ERR_CUDA, std::string("cudaPointerGetAttributes: ") + cudaGetErrorString(e));
  }
  if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)
    return qs::fail(QSIM_ERR_ARG, "buffer is not device memory (no CPU path exists in this library)");
  if (at.device < 0 || at.device >= kMaxDevices) return qs::fail(QSIM_ERR_ARG, "device index out of range");
  QS_CUDA(cudaSetDevice(at.device));
  DevCtx& c = g_ctx[at.device];
  if (!c.ready) {
    cudaDeviceProp prop;
    QS_CUDA(cudaGetDeviceProperties(&prop, at.device));
    c.sms = prop.multiProcessorCount;
    QS_CUDA(cudaMalloc(&c.d_partials, sizeof(double) * 2 * kReduceBlocks));
    QS_CUDA(cudaMalloc(&c.d_out, sizeof(double) * 2));
    QS_CUDA(cudaMallocHost(&c.h_out, sizeof(double) * 2));
    // stream-ordered allocations (generic-gate path) keep their memory between calls instead
    // of handing it back to the driver at every synchronisation
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, at.device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    c.ready = true;
  }
  *ctx = &c;
  return QSIM_OK;
}

// =====================================================================================
// Tile pass: the fused multi-gate kernel (plan.h / tile_exec.h)
// =====================================================================================
// 16-byte asynchronous global->shared copy (LDGSTS): no register staging, and the
// thread does not wait, so the next tile streams in while the current one computes.
struct CopyAsync16 {
  __device__ __forceinline__ void operator()(void* dst, const void* src) const {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
  }
};
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// One pass of the fused plan.  Persistent grid: every CTA loops over tiles.  With
// `dbuf` the dynamic shared memory holds two tile buffers and the loads of tile
// k+1 are issued before the steps of tile k run.
template <int MAXR, bool DENSE>
__global__ void __launch_bounds__(QS_THREADS, (QS_THREADS_LOG2 >= 9 ? (MAXR <= 3 ? 2 : 1)
                                               : QS_THREADS_LOG2 <= 7 ? 3 : (MAXR <= 3 ? 3 : 2)))
k_tile_pass(qs_c128* state, const __grid_constant__ QsPass P, uint64_t ntiles, int dbuf, int debug_skip) {
  extern __shared__ __align__(16) unsigned char qs_smem[];
  qs_c128* buf0 = reinterpret_cast<qs_c128*>(qs_smem);
  qs_c128* buf1 = buf0 + (dbuf ? (1u << P.T) : 0u);
  __shared__ QsStepTab s_tab[QS_MAX_STEPS];
  __shared__ QsIoTab s_io;
  __shared__ uint32_t s_zmask[QS_MAX_STEPS + 2];   // [nsteps], then final z, final g
  const uint32_t tid = threadIdx.x;
  const int nsteps = (int)P.nsteps;

  // tile-independent tables, once per launch (the grid is persistent)
  for (int e = (int)tid; e < nsteps * QS_TAB_ENTRIES; e += QS_THREADS)
    qs_build_step_tab(P, e / QS_TAB_ENTRIES, e % QS_TAB_ENTRIES, &s_tab[e / QS_TAB_ENTRIES], QS_THREADS_LOG2);
  if (tid < QS_MAX_ITER) qs_build_io_tab(P, tid, &s_io, QS_THREADS_LOG2);
  if (tid == QS_MAX_ITER) s_io.fin_q = qs_build_fin_q(P, QS_THREADS_LOG2);
  for (uint32_t e = tid; e < 256; e += QS_THREADS) qs_build_base_tab(P, e, &s_io);
  const int lo_bits = (int)(P.T < QS_THREADS_LOG2 ? P.T : QS_THREADS_LOG2);
  const uint64_t glo = qs_scatter64(tid, P.tile_bits, lo_bits);
  const uint32_t fin_qlo = P.fin_has_sign ? qs_fin_quad(P, tid & ((1u << P.T) - 1u)) : 0u;
  __syncthreads();

  uint64_t t = blockIdx.x;
  if (dbuf && t < ntiles)
    qs_phase_load(P, state, buf0, qs_tile_base_tab(P, s_io, t), tid, QS_THREADS_LOG2, glo, s_io, CopyAsync16());
  cp_async_commit();

  for (uint32_t k = 0; t < ntiles; t += gridDim.x, ++k) {
    qs_c128* cur = (dbuf && (k & 1u)) ? buf1 : buf0;
    qs_c128* nxt = (k & 1u) ? buf0 : buf1;
    const uint64_t base = qs_tile_base_tab(P, s_io, t);
    if ((int)tid < nsteps) {
      if (P.steps[tid].has_sign) s_zmask[tid] = qs_step_zg(P, (int)tid, base);
    } else if ((int)tid == nsteps && P.fin_has_sign) {
      qs_fin_prepare(P, base, &s_zmask[nsteps], &s_zmask[nsteps + 1]);
    }
    if (dbuf) {
      const uint64_t tn = t + gridDim.x;
      if (tn < ntiles)
        qs_phase_load(P, state, nxt, qs_tile_base_tab(P, s_io, tn), tid, QS_THREADS_LOG2, glo, s_io, CopyAsync16());
      cp_async_commit();
      cp_async_wait<1>();          // everything but the prefetch just issued has landed
    } else {
      if (!(debug_skip & 1)) qs_phase_load(P, state, cur, base, tid, QS_THREADS_LOG2, glo, s_io, CopyAsync16());
      cp_async_commit();
      cp_async_wait<0>();
    }
    __syncthreads();
    for (int s = 0; s < nsteps; ++s) {
      qs_phase_step_any<MAXR, DENSE>(P, s, cur, tid, QS_THREADS_LOG2, s_zmask[s], s_tab[s], debug_skip);
      // inside a run the next step touches only amplitudes of the same warp (plan.h)
      if (P.steps[s].block_sync) __syncthreads();
      else __syncwarp();
    }
    if (!(debug_skip & 2))
      qs_phase_store(P, state, cur, base, tid, QS_THREADS_LOG2, glo, s_io, fin_qlo, s_zmask[nsteps],
                     s_zmask[nsteps + 1]);
    __syncthreads();
  }
  cp_async_wait<0>();
}

typedef void (*TileKernel)(qs_c128*, const QsPass, uint64_t, int, int);

int launch_pass(DevCtx* ctx, const QsPass& P, qs_c128* state, int n, cudaStream_t stream) {
  if ((int)P.T > n) return qs::fail(QSIM_ERR_ARG, "pass tile larger than the state");
  const uint64_t ntiles = 1ull << (n - (int)P.T);
  // double-buffer when two tiles still leave room for three CTAs per SM (T <= 11);
  // QSIM_DBUF=0/1 overrides for experiments
  static const int dbuf_env = [] {
    const char* e = getenv("QSIM_DBUF");
    return e ? atoi(e) : -1;
  }();
  const int dbuf = dbuf_env >= 0 ? (dbuf_env != 0) : (P.T <= 11);
  const int smem = (int)(sizeof(qs_c128) << P.T) * (dbuf ? 2 : 1);
  int maxr = 1;
  bool dense = false;
  for (uint32_t s = 0; s < P.nsteps; ++s) {
    maxr = P.steps[s].r > maxr ? P.steps[s].r : maxr;
    dense |= P.steps[s].kind == QS_STEP_DENSE;
  }
  static const TileKernel variants[4] = {k_tile_pass<3, false>, k_tile_pass<3, true>,
                                         k_tile_pass<4, false>, k_tile_pass<4, true>};
  const int vi = (maxr <= 3 ? 0 : 2) + (dense ? 1 : 0);
  if (ctx->occ_smem != smem) {
    for (int v = 0; v < 4; ++v) {
      QS_CUDA(cudaFuncSetAttribute(variants[v], cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      QS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ[v], variants[v], QS_THREADS, smem));
    }
    ctx->occ_smem = smem;
  }
  const int occ = ctx->occ[vi];
  if (occ < 1) return qs::fail(QSIM_ERR_CUDA, "tile pass does not fit on an SM");
  uint64_t grid = (uint64_t)ctx->sms * (uint64_t)occ;
  if (grid > ntiles) grid = ntiles;
  // QSIM_DEBUG_SKIP (development only): bit 0 skips the global loads, bit 1 the global
  // stores of a pass, bit 2 the matrix arithmetic, to time the parts on their own
  // (results are garbage)
  static const int debug_skip = [] {
    const char* e = getenv("QSIM_DEBUG_SKIP");
    return e ? atoi(e) : 0;
  }();
  variants[vi]<<<(unsigned)grid, QS_THREADS, smem, stream>>>(state, P, ntiles, dbuf, debug_skip);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
  return QSIM_OK;
}

// =====================================================================================
// Simple streaming kernels
// =====================================================================================
// the 2 x n single-qubit amplitudes travel as a kernel parameter (no allocation, no copy)
struct ProductAmps { double a[4 * 40]; };

__global__ void k_init_product(qs_c128* state, int n, const __grid_constant__ ProductAmps amps, uint64_t count) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    state[i] = qs_product_amp(amps.a, n, i);
}

// n >= 16: the index splits into (o2 | o1 | tid) = (bits 16.., bits 8..15, bits 0..7) and the
// amplitude into three factors: one per thread (registers), a 256-entry table per block
// (shared memory) and one per o2 (uniform) -- two complex multiplications per amplitude
// instead of n, so the kernel runs at the HBM write rate.
__device__ __forceinline__ qs_c128 product_bits(const ProductAmps& amps, int n, uint32_t v, int first_bit, int nbits) {
  qs_c128 p; p.x = 1.0; p.y = 0.0;
  for (int b = 0; b < nbits; ++b) {
    const int q = n - 1 - (first_bit + b);
    const int bit = (int)((v >> b) & 1u);
    qs_c128 a; a.x = amps.a[4 * q + 2 * bit]; a.y = amps.a[4 * q + 2 * bit + 1];
    p = qs_cmul(p, a);
  }
  return p;
}

__global__ void __launch_bounds__(256) k_init_product_big(qs_c128* state, int n, const __grid_constant__ ProductAmps amps) {
  __shared__ qs_c128 t1[256];
  const uint32_t tid = threadIdx.x;
  const qs_c128 lo = product_bits(amps, n, tid, 0, 8);
  t1[tid] = product_bits(amps, n, tid, 8, 8);
  __syncthreads();
  const uint64_t n2 = 1ull << (n - 16);
  for (uint64_t o2 = blockIdx.x; o2 < n2; o2 += gridDim.x) {
    const qs_c128 pl = qs_cmul(product_bits(amps, n, (uint32_t)o2, 16, n - 16), lo);
    qs_c128* dst = state + (o2 << 16) + tid;
#pragma unroll 4
    for (uint32_t o1 = 0; o1 < 256; ++o1) dst[(uint64_t)o1 << 8] = qs_cmul(pl, t1[o1]);
  }
}

struct BraPair { double b[4]; };

__global__ void k_collapse(const qs_c128* __restrict__ in, qs_c128* __restrict__ out, int pos,
                           BraPair bra, double norm, uint64_t count) {
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < count;
       r += (uint64_t)gridDim.x * blockDim.x) {
    qs_c128 v = qs_contract(in, pos, r, bra.b);
    v.x /= norm; v.y /= norm;
    out[r] = v;
  }
}

__global__ void k_insert(const qs_c128* __restrict__ in, qs_c128* __restrict__ out, int pos,
                         BraPair amp, uint64_t count) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = qs_insert_amp(in, pos, i, amp.b);
}

struct BitList { int bits[10]; };

__global__ void k_generic(const qs_c128* __restrict__ in, qs_c128* __restrict__ out,
                          const double* __restrict__ mat, BitList bl, int k, uint64_t count) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = qs_generic_amp(in, mat, bl.bits, k, i);
}

__global__ void k_swap_pack(const qs_c128* __restrict__ shard, qs_c128* __restrict__ buf, QsBitSel sel,
                            uint64_t first, uint64_t count) {
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < count;
       r += (uint64_t)gridDim.x * blockDim.x)
    buf[r] = shard[qs_deposit(first + r, sel)];
}

__global__ void k_swap_unpack(qs_c128* __restrict__ shard, const qs_c128* __restrict__ buf, QsBitSel sel,
                              uint64_t first, uint64_t count) {
  for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < count;
       r += (uint64_t)gridDim.x * blockDim.x)
    shard[qs_deposit(first + r, sel)] = buf[r];
}

unsigned stream_grid(const DevCtx* ctx, uint64_t count, int threads) {
  uint64_t blocks = (count + threads - 1) / threads;
  const uint64_t cap = (uint64_t)ctx->sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// =====================================================================================
// Reductions: fixed grid, fixed tree -> bit-reproducible results run to run
// =====================================================================================
struct FnNorm2 {
  const qs_c128* a;
  __device__ void operator()(uint64_t i, double& s0, double& s1) const {
    const qs_c128 v = a[i];
    s0 += v.x * v.x + v.y * v.y;
  }
};
struct FnInner {   // sum conj(a_i) b_i
  const qs_c128 *a, *b;
  __device__ void operator()(uint64_t i, double& s0, double& s1) const {
    const qs_c128 u = a[i], v = b[i];
    s0 += u.x * v.x + u.y * v.y;
    s1 += u.x * v.y - u.y * v.x;
  }
};
struct FnMeasure {  // s0 += |bra0 . pair|^2, s1 += |bra1 . pair|^2
  const qs_c128* a;
  int pos;
  BraPair b0, b1;
  __device__ void operator()(uint64_t r, double& s0, double& s1) const {
    const qs_c128 u = qs_contract(a, pos, r, b0.b);
    const qs_c128 v = qs_contract(a, pos, r, b1.b);
    s0 += u.x * u.x + u.y * u.y;
    s1 += v.x * v.x + v.y * v.y;
  }
};
struct FnExpect {   // sum_ij conj(k_i) rho_ij k_j ; element e = i * dim + j
  const qs_c128 *ket, *rho;
  int n;
  __device__ void operator()(uint64_t e, double& s0, double& s1) const {
    const uint64_t i = e >> n, j = e & ((1ull << n) - 1ull);
    const qs_c128 t = qs_cmul(rho[e], ket[j]);
    const qs_c128 k = ket[i];
    s0 += k.x * t.x + k.y * t.y;
    s1 += k.x * t.y - k.y * t.x;
  }
};
struct FnPurity {   // sum_ij rho_ij rho_ji
  const qs_c128* rho;
  int n;
  __device__ void operator()(uint64_t e, double& s0, double& s1) const {
    const uint64_t i = e >> n, j = e & ((1ull << n) - 1ull);
    const qs_c128 t = qs_cmul(rho[e], rho[(j << n) | i]);
    s0 += t.x;
    s1 += t.y;
  }
};
struct FnTrace {
  const qs_c128* rho;
  int n;
  __device__ void operator()(uint64_t i, double& s0, double& s1) const {
    const qs_c128 v = rho[(i << n) | i];
    s0 += v.x;
    s1 += v.y;
  }
};

__device__ __forceinline__ void block_reduce2(double& s0, double& s1) {
  __shared__ double w0[kReduceThreads / 32], w1[kReduceThreads / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s0 += __shfl_down_sync(0xffffffffu, s0, off);
    s1 += __shfl_down_sync(0xffffffffu, s1, off);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { w0[warp] = s0; w1[warp] = s1; }
  __syncthreads();
  if (warp == 0) {
    s0 = lane < kReduceThreads / 32 ? w0[lane] : 0.0;
    s1 = lane < kReduceThreads / 32 ? w1[lane] : 0.0;
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
      s0 += __shfl_down_sync(0xffffffffu, s0, off);
      s1 += __shfl_down_sync(0xffffffffu, s1, off);
    }
  }
}

template <class F>
__global__ void __launch_bounds__(kReduceThreads) k_reduce(F f, uint64_t count, double* partials) {
  double s0 = 0.0, s1 = 0.0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count;
       i += (uint64_t)gridDim.x * blockDim.x)
    f(i, s0, s1);
  block_reduce2(s0, s1);
  if (threadIdx.x == 0) { partials[2 * blockIdx.x] = s0; partials[2 * blockIdx.x + 1] = s1; }
}

__global__ void __launch_bounds__(kReduceThreads) k_reduce_final(const double* partials, int nblocks, double* out) {
  double s0 = 0.0, s1 = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { s0 += partials[2 * i]; s1 += partials[2 * i + 1]; }
  block_reduce2(s0, s1);
  if (threadIdx.x == 0) { out[0] = s0; out[1] = s1; }
}

template <class F>
int reduce_to_host(DevCtx* ctx, F f, uint64_t count, double* out2, cudaStream_t stream) {
  uint64_t blocks = (count + kReduceThreads - 1) / kReduceThreads;
  if (blocks > (uint64_t)kReduceBlocks) blocks = kReduceBlocks;
  if (blocks < 1) blocks = 1;
  k_reduce<F><<<(unsigned)blocks, kReduceThreads, 0, stream>>>(f, count, ctx->d_partials);
  k_reduce_final<<<1, kReduceThreads, 0, stream>>>(ctx->d_partials, (int)blocks, ctx->d_out);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
  QS_CUDA(cudaMemcpyAsync(ctx->h_out, ctx->d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
  QS_CUDA(cudaStreamSynchronize(stream));
  out2[0] = ctx->h_out[0];
  out2[1] = ctx->h_out[1];
  return QSIM_OK;
}

// =====================================================================================
// Batched tiny-circuit executor
// =====================================================================================
template <int DIM>
__global__ void __launch_bounds__(128) k_rb_batch(int64_t n_seq, const uint16_t* __restrict__ codes,
                                                  const int64_t* __restrict__ offsets,
                                                  const double* __restrict__ superops,
                                                  const double* __restrict__ unitaries,
                                                  const double* __restrict__ rho0,
                                                  const double* __restrict__ psi0, double* out_fid,
                                                  double* out_pur, double* out_rho) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= n_seq) return;
  const int64_t lo = offsets[b], hi = offsets[b + 1];
  qs_rb_sequence<DIM>(codes + lo, hi - lo, superops, unitaries, rho0, psi0, out_fid + b, out_pur + b,
                      out_rho ? out_rho + (size_t)b * 2 * DIM * DIM : nullptr);
}

int run_generic(DevCtx* ctx, const qs::Op& op, qs_c128* state, qs_c128* scratch, int n, cudaStream_t stream) {
  if (!scratch) return qs::fail(QSIM_ERR_ARG, "a gate on more than 4 qubits needs a scratch buffer of 2^n amplitudes");
  const int dim = 1 << op.k;
  const size_t bytes = sizeof(double) * 2 * dim * dim;
  std::vector<double> flat(2 * (size_t)dim * dim);
  for (int e = 0; e < dim * dim; ++e) { flat[2 * e] = op.mat[e].real(); flat[2 * e + 1] = op.mat[e].imag(); }
  double* d_mat = nullptr;
  QS_CUDA(cudaMallocAsync(&d_mat, bytes, stream));
  QS_CUDA(cudaMemcpyAsync(d_mat, flat.data(), bytes, cudaMemcpyHostToDevice, stream));
  QS_CUDA(cudaStreamSynchronize(stream));     // flat goes out of scope; generic path is not the fast path
  BitList bl;
  for (int f = 0; f < op.k; ++f) bl.bits[f] = op.bits[f];
  const uint64_t count = 1ull << n;
  k_generic<<<stream_grid(ctx, count, 256), 256, 0, stream>>>(state, scratch, d_mat, bl, op.k, count);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
  QS_CUDA(cudaMemcpyAsync(state, scratch, sizeof(qs_c128) * count, cudaMemcpyDeviceToDevice, stream));
  QS_CUDA(cudaFreeAsync(d_mat, stream));
  return QSIM_OK;
}

int execute_plan(const qsim_plan* p, void* state, int n, void* scratch, void* stream) {
  if (!p || !state) return qs::fail(QSIM_ERR_ARG, "qsim_plan_execute: null argument");
  if (n != p->n) return qs::fail(QSIM_ERR_ARG, "qsim_plan_execute: plan was compiled for a different qubit count");
  DevCtx* ctx = nullptr;
  int rc = bind_device(state, &ctx);
  if (rc != QSIM_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  for (const qs::PlanItem& it : p->items) {
    rc = it.generic ? run_generic(ctx, it.op, (qs_c128*)state, (qs_c128*)scratch, n, st)
                    : launch_pass(ctx, it.pass, (qs_c128*)state, n, st);
    if (rc != QSIM_OK) return rc;
  }
  return QSIM_OK;
}

int one_gate(void* state, int n, const int* targets, int k, const double* matrix, void* scratch, void* stream) {
  qsim_circuit_t* c = nullptr;
  int rc = qsim_circuit_create(n, &c);
  if (rc != QSIM_OK) return rc;
  rc = qsim_circuit_add_matrix(c, k, targets, matrix);
  qsim_plan_t* p = nullptr;
  if (rc == QSIM_OK) rc = qsim_plan_compile(c, nullptr, &p);
  if (rc == QSIM_OK) rc = execute_plan(p, state, n, scratch, stream);
  qsim_plan_destroy(p);
  qsim_circuit_destroy(c);
  return rc;
}

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int qsim_has_cuda(void) { return 1; }

int64_t qsim_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int qsim_peer_alloc(int device, uint64_t bytes, void** out_ptr) {
  if (!out_ptr || bytes == 0) return qs::fail(QSIM_ERR_ARG, "qsim_peer_alloc: bad argument");
  QS_CUDA(cudaSetDevice(device));
  QS_CUDA(cudaMalloc(out_ptr, bytes));
  return QSIM_OK;
}

int qsim_peer_free(void* ptr) {
  if (ptr) QS_CUDA(cudaFree(ptr));
  return QSIM_OK;
}

int qsim_ipc_export(void* ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (!ptr || !handle64) return qs::fail(QSIM_ERR_ARG, "qsim_ipc_export: null argument");
  cudaIpcMemHandle_t h;
  QS_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  return QSIM_OK;
}

int qsim_ipc_import(int device, const unsigned char* handle64, void** out_ptr) {
  if (!handle64 || !out_ptr) return qs::fail(QSIM_ERR_ARG, "qsim_ipc_import: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  QS_CUDA(cudaSetDevice(device));
  QS_CUDA(cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return QSIM_OK;
}

int qsim_ipc_release(void* imported_ptr) {
  if (imported_ptr) QS_CUDA(cudaIpcCloseMemHandle(imported_ptr));
  return QSIM_OK;
}

int qsim_peer_copy(void* dst, const void* src, uint64_t bytes, void* stream) {
  if (!dst || !src) return qs::fail(QSIM_ERR_ARG, "qsim_peer_copy: null argument");
  QS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return QSIM_OK;
}

int qsim_plan_execute(const qsim_plan_t* p, void* state, int n_qubits, void* scratch, void* stream) {
  return execute_plan(p, state, n_qubits, scratch, stream);
}

int qsim_apply_matrix(void* state, int n_qubits, const int* targets, int k, const double* matrix,
                      void* scratch, void* stream) {
  if (!state || !targets || !matrix) return qs::fail(QSIM_ERR_ARG, "qsim_apply_matrix: null argument");
  return one_gate(state, n_qubits, targets, k, matrix, scratch, stream);
}

int qsim_apply_diagonal(void* state, int n_qubits, const int* targets, int k, const double* diag, void* stream) {
  if (!state || !targets || !diag) return qs::fail(QSIM_ERR_ARG, "qsim_apply_diagonal: null argument");
  if (k < 1 || k > QS_MAX_R) return qs::fail(QSIM_ERR_UNSUPPORTED, "qsim_apply_diagonal: k must be in [1, 4]");
  const int dim = 1 << k;
  std::vector<double> m(2 * (size_t)dim * dim, 0.0);
  for (int d = 0; d < dim; ++d) { m[2 * (d * dim + d)] = diag[2 * d]; m[2 * (d * dim + d) + 1] = diag[2 * d + 1]; }
  return one_gate(state, n_qubits, targets, k, m.data(), nullptr, stream);
}

int qsim_apply_permutation(void* state, int n_qubits, const int* targets, int k, const int* perm, void* stream) {
  if (!state || !targets || !perm) return qs::fail(QSIM_ERR_ARG, "qsim_apply_permutation: null argument");
  if (k < 1 || k > QS_MAX_R) return qs::fail(QSIM_ERR_UNSUPPORTED, "qsim_apply_permutation: k must be in [1, 4]");
  const int dim = 1 << k;
  std::vector<double> m(2 * (size_t)dim * dim, 0.0);
  std::vector<char> hit(dim, 0);
  for (int c = 0; c < dim; ++c) {
    if (perm[c] < 0 || perm[c] >= dim || hit[perm[c]]) return qs::fail(QSIM_ERR_ARG, "qsim_apply_permutation: not a permutation");
    hit[perm[c]] = 1;
    m[2 * (perm[c] * dim + c)] = 1.0;
  }
  return one_gate(state, n_qubits, targets, k, m.data(), nullptr, stream);
}

int qsim_apply_superop(void* vec_rho, int n_qubits, const int* targets, int k, const double* superop,
                       void* scratch, void* stream) {
  if (!vec_rho || !targets || !superop) return qs::fail(QSIM_ERR_ARG, "qsim_apply_superop: null argument");
  if (k < 1 || 2 * k > 10) return qs::fail(QSIM_ERR_UNSUPPORTED, "qsim_apply_superop: k must be in [1, 5]");
  std::vector<int> t(2 * k);
  for (int i = 0; i < k; ++i) { t[i] = targets[i]; t[k + i] = targets[i] + n_qubits; }
  return one_gate(vec_rho, 2 * n_qubits, t.data(), 2 * k, superop, scratch, stream);
}

int qsim_init_product(void* state, int n_qubits, const double* amps, void* stream) {
  if (!state || !amps || n_qubits < 1 || n_qubits > 40) return qs::fail(QSIM_ERR_ARG, "qsim_init_product: bad argument");
  DevCtx* ctx = nullptr;
  int rc = bind_device(state, &ctx);
  if (rc != QSIM_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  ProductAmps pa;
  memset(&pa, 0, sizeof(pa));
  memcpy(pa.a, amps, sizeof(double) * 4 * n_qubits);
  const uint64_t count = 1ull << n_qubits;
  if (n_qubits >= 16) {
    uint64_t blocks = count >> 16;
    if (blocks > (uint64_t)ctx->sms * 8) blocks = (uint64_t)ctx->sms * 8;
    k_init_product_big<<<(unsigned)blocks, 256, 0, st>>>((qs_c128*)state, n_qubits, pa);
  } else {
    k_init_product<<<stream_grid(ctx, count, 256), 256, 0, st>>>((qs_c128*)state, n_qubits, pa, count);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
  return QSIM_OK;
}

int qsim_measure_probs(const void* state, int n_qubits, int qubit, const double* bra0, const double* bra1,
                       double* out_norm2, void* stream) {
  if (!state || !bra0 || !bra1 || !out_norm2) return qs::fail(QSIM_ERR_ARG, "qsim_measure_probs: null argument");
  if (n_qubits < 1 || qubit < 0 || qubit >= n_qubits) return qs::fail(QSIM_ERR_ARG, "qsim_measure_probs: qubit out of range");
  DevCtx* ctx = nullptr;
  int rc = bind_device(state, &ctx);
  if (rc != QSIM_OK) return rc;
  FnMeasure f;
  f.a = (const qs_c128*)state;
  f.pos = n_qubits - 1 - qubit;
  memcpy(f.b0.b, bra0, sizeof(f.b0.b));
  memcpy(f.b1.b, bra1, sizeof(f.b1.b));
  return reduce_to_host(ctx, f, 1ull << (n_qubits - 1), out_norm2, (cudaStream_t)stream);
}

int qsim_collapse(const void* in, void* out, int n_qubits, int qubit, const double* bra, double norm, void* stream) {
  if (!in || !out || !bra) return qs::fail(QSIM_ERR_ARG, "qsim_collapse: null argument");
  if (n_qubits < 1 || qubit < 0 || qubit >= n_qubits) return qs::fail(QSIM_ERR_ARG, "qsim_collapse: qubit out of range");
  DevCtx* ctx = nullptr;
  int rc = bind_device(in, &ctx);
  if (rc != QSIM_OK) return rc;
  BraPair b;
  memcpy(b.b, bra, sizeof(b.b));
  const uint64_t count = 1ull << (n_qubits - 1);
  k_collapse<<<stream_grid(ctx, count, 256), 256, 0, (cudaStream_t)stream>>>(
      (const qs_c128*)in, (qs_c128*)out, n_qubits - 1 - qubit, b, norm, count);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
  return QSIM_OK;
}

int qsim_insert(const void* in, void* out, int n_qubits, int position, const double* amp, void* stream) {
  if (!in || !out || !amp) return qs::fail(QSIM_ERR_ARG, "qsim_insert: null argument");
  if (n_qubits < 0 || position < 0 || position > n_qubits) return qs::fail(QSIM_ERR_ARG, "qsim_insert: position out of range");
  DevCtx* ctx = nullptr;
  int rc = bind_device(in, &ctx);
  if (rc != QSIM_OK) return rc;
  BraPair a;
  memcpy(a.b, amp, sizeof(a.b));
  const uint64_t count = 1ull << (n_qubits + 1);
  // the new register has n+1 qubits; reference position p is index bit (n+1)-1-p
  k_insert<<<stream_grid(ctx, count, 256), 256, 0, (cudaStream_t)stream>>>(
      (const qs_c128*)in, (qs_c128*)out, n_qubits - position, a, count);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
  return QSIM_OK;
}

int qsim_reduce_norm2(const void* state, uint64_t n_amps, double* out, void* stream) {
  if (!state || !out) return qs::fail(QSIM_ERR_ARG, "qsim_reduce_norm2: null argument");
  DevCtx* ctx = nullptr;
  int rc = bind_device(state, &ctx);
  if (rc != QSIM_OK) return rc;
  FnNorm2 f{(const qs_c128*)state};
  double r[2];
  rc = reduce_to_host(ctx, f, n_amps, r, (cudaStream_t)stream);
  if (rc == QSIM_OK) *out = r[0];
  return rc;
}

int qsim_reduce_inner(const void* a, const void* b, uint64_t n_amps, double* out_re_im, void* stream) {
  if (!a || !b || !out_re_im) return qs::fail(QSIM_ERR_ARG, "qsim_reduce_inner: null argument");
  DevCtx* ctx = nullptr;
  int rc = bind_device(a, &ctx);
  if (rc != QSIM_OK) return rc;
  FnInner f{(const qs_c128*)a, (const qs_c128*)b};
  return reduce_to_host(ctx, f, n_amps, out_re_im, (cudaStream_t)stream);
}

int qsim_reduce_expect(const void* ket, const void* rho, int n_qubits, double* out_re_im, void* stream) {
  if (!ket || !rho || !out_re_im || n_qubits < 0 || n_qubits > 20) return qs::fail(QSIM_ERR_ARG, "qsim_reduce_expect: bad argument");
  DevCtx* ctx = nullptr;
  int rc = bind_device(rho, &ctx);
  if (rc != QSIM_OK) return rc;
  FnExpect f{(const qs_c128*)ket, (const qs_c128*)rho, n_qubits};
  return reduce_to_host(ctx, f, 1ull << (2 * n_qubits), out_re_im, (cudaStream_t)stream);
}

int qsim_reduce_purity(const void* rho, int n_qubits, double* out_re_im, void* stream) {
  if (!rho || !out_re_im || n_qubits < 0 || n_qubits > 20) return qs::fail(QSIM_ERR_ARG, "qsim_reduce_purity: bad argument");
  DevCtx* ctx = nullptr;
  int rc = bind_device(rho, &ctx);
  if (rc != QSIM_OK) return rc;
  FnPurity f{(const qs_c128*)rho, n_qubits};
  return reduce_to_host(ctx, f, 1ull << (2 * n_qubits), out_re_im, (cudaStream_t)stream);
}

int qsim_reduce_trace(const void* rho, int n_qubits, double* out_re_im, void* stream) {
  if (!rho || !out_re_im || n_qubits < 0 || n_qubits > 20) return qs::fail(QSIM_ERR_ARG, "qsim_reduce_trace: bad argument");
  DevCtx* ctx = nullptr;
  int rc = bind_device(rho, &ctx);
  if (rc != QSIM_OK) return rc;
  FnTrace f{(const qs_c128*)rho, n_qubits};
  return reduce_to_host(ctx, f, 1ull << n_qubits, out_re_im, (cudaStream_t)stream);
}

int qsim_rb_batch(int nq, int64_t n_seq, const uint16_t* opcodes, const int64_t* offsets, int n_opcodes,
                  const double* superops, const double* unitaries, const double* rho0, const double* psi0,
                  double* out_fidelity, double* out_purity, double* out_rho, void* stream) {
  if (!opcodes || !offsets || !superops || !unitaries || !rho0 || !psi0 || !out_fidelity || !out_purity)
    return qs::fail(QSIM_ERR_ARG, "qsim_rb_batch: null argument");
  if (nq < 1 || nq > 2) return qs::fail(QSIM_ERR_UNSUPPORTED, "qsim_rb_batch: nq must be 1 or 2");
  if (n_seq < 0 || n_opcodes < 1 || n_opcodes > 65536) return qs::fail(QSIM_ERR_ARG, "qsim_rb_batch: bad sizes");
  if (n_seq == 0) return QSIM_OK;
  DevCtx* ctx = nullptr;
  int rc = bind_device(out_fidelity, &ctx);
  if (rc != QSIM_OK) return rc;
  const unsigned blocks = (unsigned)((n_seq + 127) / 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (nq == 2)
    k_rb_batch<4><<<blocks, 128, 0, st>>>(n_seq, opcodes, offsets, superops, unitaries, rho0, psi0,
                                         out_fidelity, out_purity, out_rho);
  else
    k_rb_batch<2><<<blocks, 128, 0, st>>>(n_seq, opcodes, offsets, superops, unitaries, rho0, psi0,
                                         out_fidelity, out_purity, out_rho);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
  return QSIM_OK;
}

// fills `sel` (positions ascending) from reference-style local qubit numbers
static int swap_select(const char* who, int n_local, int nbits, const int* local_qubits, const int* bit_values,
                       uint64_t first, uint64_t count, QsBitSel* sel) {
  if (n_local < 1 || nbits < 1 || nbits > 8 || nbits > n_local || !local_qubits || !bit_values)
    return qs::fail(QSIM_ERR_ARG, std::string(who) + ": bad bit list");
  sel->k = nbits;
  for (int i = 0; i < nbits; ++i) {
    if (local_qubits[i] < 0 || local_qubits[i] >= n_local || (bit_values[i] | 1) != 1)
      return qs::fail(QSIM_ERR_ARG, std::string(who) + ": bad qubit or bit");
    sel->pos[i] = n_local - 1 - local_qubits[i];
    sel->val[i] = bit_values[i];
  }
  for (int i = 1; i < nbits; ++i)                 // insertion sort by position
    for (int j = i; j > 0 && sel->pos[j] < sel->pos[j - 1]; --j) {
      std::swap(sel->pos[j], sel->pos[j - 1]);
      std::swap(sel->val[j], sel->val[j - 1]);
    }
  for (int i = 1; i < nbits; ++i)
    if (sel->pos[i] == sel->pos[i - 1]) return qs::fail(QSIM_ERR_ARG, std::string(who) + ": repeated qubit");
  if (first + count > (1ull << (n_local - nbits)))
    return qs::fail(QSIM_ERR_ARG, std::string(who) + ": chunk out of range");
  return QSIM_OK;
}

int qsim_swap_pack(const void* shard, void* sendbuf, int n_local, int nbits, const int* local_qubits,
                   const int* bit_values, uint64_t first, uint64_t count, void* stream) {
  if (!shard || !sendbuf) return qs::fail(QSIM_ERR_ARG, "qsim_swap_pack: null argument");
  QsBitSel sel;
  int rc = swap_select("qsim_swap_pack", n_local, nbits, local_qubits, bit_values, first, count, &sel);
  if (rc != QSIM_OK) return rc;
  DevCtx* ctx = nullptr;
  rc = bind_device(shard, &ctx);
  if (rc != QSIM_OK) return rc;
  if (count == 0) return QSIM_OK;
  k_swap_pack<<<stream_grid(ctx, count, 256), 256, 0, (cudaStream_t)stream>>>(
      (const qs_c128*)shard, (qs_c128*)sendbuf, sel, first, count);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
  return QSIM_OK;
}

int qsim_swap_unpack(void* shard, const void* recvbuf, int n_local, int nbits, const int* local_qubits,
                     const int* bit_values, uint64_t first, uint64_t count, void* stream) {
  if (!shard || !recvbuf) return qs::fail(QSIM_ERR_ARG, "qsim_swap_unpack: null argument");
  QsBitSel sel;
  int rc = swap_select("qsim_swap_unpack", n_local, nbits, local_qubits, bit_values, first, count, &sel);
  if (rc != QSIM_OK) return rc;
  DevCtx* ctx = nullptr;
  rc = bind_device(shard, &ctx);
  if (rc != QSIM_OK) return rc;
  if (count == 0) return QSIM_OK;
  k_swap_unpack<<<stream_grid(ctx, count, 256), 256, 0, (cudaStream_t)stream>>>(
      (qs_c128*)shard, (const qs_c128*)recvbuf, sel, first, count);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
  return QSIM_OK;
}

}  // extern "C"
