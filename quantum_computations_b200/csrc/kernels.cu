// sm_100a kernels and the device-facing half of the C ABI (include/qsim_b200.h).
//
// Almost everything here is memory-bound complex128 work: TMA-staged 2^T-amplitude
// tiles in shared memory so that many gates apply per HBM pass, 16-byte (128-bit)
// accesses per amplitude, persistent grids sized to the SM count.  Tensor cores
// appear once: dense blocks on 5..10 qubits are FP64 MMAs (k_dense_block); tcgen05
// has no f64 kind and the 2x2 / 4x4 updates of the tile pass have no GEMM shape.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <algorithm>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "elem_ops.h"
#include "planner.h"

namespace {

std::atomic<int64_t> g_launches{0};

struct DevCtx {
  std::once_flag once;
  int init_rc = 0;                // cudaError_t of the one-time initialisation
  int sms = 0;
  // scratch of the reductions: one set per device, serialised by reduce_lock (a reduction
  // synchronises its stream anyway, so concurrent callers lose nothing)
  std::mutex reduce_lock;
  double* d_partials = nullptr;   // [kReduceBlocks][2]
  double* d_out = nullptr;        // [2]
  double* h_out = nullptr;        // pinned [2]
  std::mutex occ_lock;
  int occ[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // resident CTAs/SM of the k_tile_pass variants at the last smem size
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

// The device a buffer lives on, made current for the lifetime of this object (the caller's
// current device is restored afterwards: torch keeps its own idea of it).
struct Bound {
  DevCtx* ctx = nullptr;
  int device = -1;
  int prev = -1;
  ~Bound() {
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
  }
};

int bind_device(const void* ptr, Bound* b) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return qs::fail(QSIM_ERR_CUDA, std::string("cudaPointerGetAttributes: ") + cudaGetErrorString(e));
  }
  if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged)
    return qs::fail(QSIM_ERR_ARG, "buffer is not device memory (no CPU path exists in this library)");
  if (at.device < 0 || at.device >= kMaxDevices) return qs::fail(QSIM_ERR_ARG, "device index out of range");
  if (cudaGetDevice(&b->prev) != cudaSuccess) b->prev = -1;
  b->device = at.device;
  if (b->prev != at.device) QS_CUDA(cudaSetDevice(at.device));
  DevCtx& c = g_ctx[at.device];
  std::call_once(c.once, [&c, &at] {
    cudaDeviceProp prop;
    cudaError_t rc = cudaGetDeviceProperties(&prop, at.device);
    if (rc == cudaSuccess) c.sms = prop.multiProcessorCount;
    if (rc == cudaSuccess) rc = cudaMalloc(&c.d_partials, sizeof(double) * 2 * kReduceBlocks);
    if (rc == cudaSuccess) rc = cudaMalloc(&c.d_out, sizeof(double) * 2);
    if (rc == cudaSuccess) rc = cudaMallocHost(&c.h_out, sizeof(double) * 2);
    c.init_rc = (int)rc;
  });
  if (c.init_rc != 0)
    return qs::fail(QSIM_ERR_CUDA, std::string("device initialisation: ") + cudaGetErrorString((cudaError_t)c.init_rc));
  b->ctx = &c;
  return QSIM_OK;
}

// =====================================================================================
// Tile pass: the fused multi-gate kernel (plan.h / tile_exec.h)
// =====================================================================================
// TMA + mbarrier primitives (sm_90+ PTX; SASS: UTMALDG / UTMASTG / SYNCS).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Box `o` of the tile at `base`: its tile positions P.. are the bits of o.
__device__ __forceinline__ void tma_box_coords(const QsPass& P, const QsTmaGeom& G, uint64_t base, int o, int* c) {
  uint64_t b = base;
  for (int q = 0; q < (int)P.T - G.P; ++q) b |= (uint64_t)((o >> q) & 1) << P.tile_bits[G.P + q];
  c[0] = (int)(((b >> G.shift[0]) & G.mask[0]) * 2);       // dimension 0 counts doubles
#pragma unroll
  for (int i = 1; i < 5; ++i) c[i] = (int)((b >> G.shift[i]) & G.mask[i]);
}

// One pass of the fused plan.  Persistent grid: every CTA loops over tiles; TMA brings a
// tile into shared memory (one mbarrier per CTA), the steps run on it, TMA writes it back.
// The three CTAs of an SM are in different phases, so one CTA's fill and drain overlap
// the steps of the other two.
// TL2 = log2 threads per CTA (QsPass::cta_log2): 128-thread CTAs run three (smem-limited) or four
// to an SM, 256-thread CTAs two (16 amplitudes per thread) or three (8 per thread).
template <int MAXR, int DENSE, int TL2>
__global__ void __launch_bounds__(1 << TL2, (TL2 == 7 ? (MAXR <= 3 ? 6 : 4) : (MAXR <= 3 ? 3 : 2)))
k_tile_pass(qs_c128* state, const __grid_constant__ QsPass P, const __grid_constant__ CUtensorMap tmap,
            const __grid_constant__ QsTmaGeom G, uint64_t ntiles) {
  extern __shared__ __align__(1024) unsigned char qs_smem[];
  qs_c128* tile = reinterpret_cast<qs_c128*>(qs_smem);
  __shared__ QsStepTab s_tab[QS_MAX_STEPS];
  __shared__ uint32_t s_zm[QS_MAX_LAYERS + 1];     // per layer qs_layer_z; [QS_MAX_LAYERS] = final g
  __shared__ __align__(8) uint64_t s_bar;
  const uint32_t tid = threadIdx.x;
  const int nsteps = (int)P.nsteps;
  const int nlayers = (int)P.nlayers;
  const bool use_tma = G.use_tma != 0;

  // tile-independent thread tables, once per launch (the grid is persistent)
  for (int e = (int)tid; e < nsteps * QS_TAB_ENTRIES; e += (1 << TL2))
    qs_build_step_tab(P, e / QS_TAB_ENTRIES, e % QS_TAB_ENTRIES, &s_tab[e / QS_TAB_ENTRIES], TL2);
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t fin_qlo = P.has_final ? qs_fin_quad(P, qs_thread_jlo(s_tab[nsteps - 1], tid)) : 0u;
  uint32_t phase = 0;
  const uint32_t box_bytes = 16u << G.P;

  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    uint64_t base = 0;
    if (!use_tma || tid < 96) base = qs_tile_base(P, t);
    if (use_tma) {
      if (tid < 32) {
        tma_wait_read0();                        // my stores of the previous tile have left the buffer
        __syncwarp();
        if (tid == 0) mbar_expect_tx(&s_bar, 16u << P.T);
        __syncwarp();
        for (int o = (int)tid; o < G.nops; o += 32) {
          int c[5];
          tma_box_coords(P, G, base, o, c);
          tma_load_5d(qs_smem + (size_t)o * box_bytes, &tmap, &s_bar, c[0], c[1], c[2], c[3], c[4]);
        }
      }
    } else {
      qs_plain_load(P, state, tile, base, tid, (1u << TL2));
    }
    // per-tile sign data, one thread per layer (warps 1 and 2)
    if (tid >= 32 && (int)tid < 32 + nlayers) {
      const int l = (int)tid - 32;
      if (P.layers[l].flags & QS_LF_SIGN) s_zm[l] = qs_layer_z(P, l, base);
    } else if (tid == 95 && P.has_final) {
      s_zm[QS_MAX_LAYERS] = qs_fin_g(P, base);
    }
    __syncthreads();
    if (use_tma) {
      while (!mbar_try_wait(&s_bar, phase)) {}
      phase ^= 1u;
    }
    const uint32_t fin_g = s_zm[QS_MAX_LAYERS];
    for (int s = 0; s < nsteps; ++s) {
      qs_phase_step_any<MAXR, DENSE>(P, s, tile, tid, TL2, s_zm, fin_g, fin_qlo, s_tab[s]);
      // the tile goes back through the async proxy: order this thread's writes before it
      if (s == nsteps - 1 && use_tma) fence_async_smem();
      // inside a run the next step touches only amplitudes of the same warp (plan.h)
      if (P.steps[s].block_sync) __syncthreads();
      else __syncwarp();
    }
    if (use_tma) {
      if (tid < 32) {
        for (int o = (int)tid; o < G.nops; o += 32) {
          int c[5];
          tma_box_coords(P, G, base, o, c);
          tma_store_5d(&tmap, qs_smem + (size_t)o * box_bytes, c[0], c[1], c[2], c[3], c[4]);
        }
        tma_commit();
      }
    } else {
      qs_plain_store(P, state, tile, base, tid, (1u << TL2));
      __syncthreads();
    }
  }
  if (use_tma && tid < 32) tma_wait_read0();
}

typedef void (*TileKernel)(qs_c128*, const QsPass, const CUtensorMap, const QsTmaGeom, uint64_t);

PFN_cuTensorMapEncodeTiled tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return (PFN_cuTensorMapEncodeTiled)p;
  }();
  return fn;
}

// Describe the state to TMA so that one box covers the lowest tile positions of the pass
// (plan.h, QsTmaGeom).  Dimension 0 is the index-bit range [0, cut1) with a box of bits
// 0..2 (one 128-byte row, the swizzle span); dimensions 1..4 start at the four lowest runs
// of tile bits above bit 2 (a run longer than 8 bits is cut: boxes are at most 256 wide).
// Returns use_tma = 0 when the pass has to use plain loads.
void make_tma_geom(const QsPass& P, qs_c128* state, int n, QsTmaGeom* G, CUtensorMap* map) {
  memset(G, 0, sizeof(*G));
  memset(map, 0, sizeof(*map));
  static const bool off = qs::dev_knob("QSIM_NO_TMA", 0) != 0;
  PFN_cuTensorMapEncodeTiled enc = tensor_map_encoder();
  const int T = (int)P.T;
  if (off || !enc || n < 12 || n > 40 || T < 6) return;
  if (P.tile_bits[0] != 0 || P.tile_bits[1] != 1 || P.tile_bits[2] != 2) return;
  bool is_tile[64] = {false};
  for (int l = 0; l < T; ++l) is_tile[P.tile_bits[l]] = true;
  int cuts[4], ncuts = 0;
  {
    int run = 0;
    for (int b = 3; b < n && ncuts < 4; ++b) {
      if (is_tile[b] && (run == 0 || run == 8)) { cuts[ncuts++] = b; run = 1; }
      else if (is_tile[b]) ++run;
      else run = 0;
    }
    for (int b = n - 1; b > 3 && ncuts < 4; --b) {     // fillers: any other positions
      bool used = false;
      for (int i = 0; i < ncuts; ++i) used |= cuts[i] == b;
      if (!used) cuts[ncuts++] = b;
    }
    if (ncuts != 4) return;
    std::sort(cuts, cuts + 4);
  }
  const int start[6] = {0, cuts[0], cuts[1], cuts[2], cuts[3], n};
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  int covered = 3;
  bool prefix = true;         // the box must cover the LOWEST tile positions, in order
  for (int i = 0; i < 5; ++i) {
    const int lo = start[i], hi = start[i + 1];
    if (hi - lo > 31) return;
    int nb = 0;
    if (i == 0) {
      nb = 3;
      for (int b = 3; b < hi; ++b)
        if (is_tile[b]) return;                        // cannot happen: cut 1 is the first tile bit above 2
    } else {
      if (prefix)
        while (lo + nb < hi && is_tile[lo + nb] && nb < 8) ++nb;
      for (int b = lo + nb; b < hi; ++b)
        if (is_tile[b]) prefix = false;
      covered += nb;
    }
    G->shift[i] = lo;
    G->mask[i] = (uint32_t)((1ull << (hi - lo)) - 1ull);
    gdim[i] = (1ull << (hi - lo)) * (i == 0 ? 2 : 1);
    box[i] = (1u << nb) * (i == 0 ? 2 : 1);
    if (i > 0) gstride[i - 1] = 16ull << lo;
  }
  if (covered < 6 || T - covered > 6) return;
  if (enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, state, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return;
  G->P = covered;
  G->nops = 1 << (T - covered);
  G->use_tma = 1;
}

int launch_pass(DevCtx* ctx, const QsPass& P, qs_c128* state, int n, cudaStream_t stream) {
  if ((int)P.T > n) return qs::fail(QSIM_ERR_ARG, "pass tile larger than the state");
  const uint64_t ntiles = 1ull << (n - (int)P.T);
  const int smem = (int)(sizeof(qs_c128) << P.T);
  int maxr = 1;
  bool dense = false;
  for (uint32_t s = 0; s < P.nsteps; ++s) maxr = P.steps[s].r > maxr ? P.steps[s].r : maxr;
  int dense_steps = 0;
  for (uint32_t s = 0; s < P.nsteps; ++s) dense_steps += P.layers[P.steps[s].layer0].kind == QS_LAYER_DENSE ? 1 : 0;
  dense = dense_steps > 0;
  // dense mode of the kernel (tile_exec.h, qs_phase_step_any): mostly dense steps -> all inline
  int dense_mode = dense_steps == 0 ? 0 : (2 * dense_steps > (int)P.nsteps ? 2 : 1);
  static const int force_dense = qs::dev_knob("QSIM_FORCE_DENSE_VARIANT", 0);
  if (force_dense == 1 || force_dense == 2) dense_mode = std::max(dense_mode, force_dense);   // development knob
  // [cta size][MAXR 3 / 4][dense mode 0 / 1 / 2]
  static const TileKernel variants[12] = {
      k_tile_pass<3, 0, 8>, k_tile_pass<3, 1, 8>, k_tile_pass<3, 2, 8>,
      k_tile_pass<4, 0, 8>, k_tile_pass<4, 1, 8>, k_tile_pass<4, 2, 8>,
      k_tile_pass<3, 0, 7>, k_tile_pass<3, 1, 7>, k_tile_pass<3, 2, 7>,
      k_tile_pass<4, 0, 7>, k_tile_pass<4, 1, 7>, k_tile_pass<4, 2, 7>};
  if (P.cta_log2 != QS_THREADS_LOG2_MIN && P.cta_log2 != QS_THREADS_LOG2)
    return qs::fail(QSIM_ERR_ARG, "pass built for an unsupported CTA size");
  const int threads = 1 << P.cta_log2;
  const int vi = (P.cta_log2 == QS_THREADS_LOG2 ? 0 : 6) + (maxr <= 3 ? 0 : 3) + dense_mode;
  int occ;
  {
    std::lock_guard<std::mutex> guard(ctx->occ_lock);
    if (ctx->occ_smem != smem) {
      for (int v = 0; v < 12; ++v) {
        QS_CUDA(cudaFuncSetAttribute(variants[v], cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        QS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ[v], variants[v], v < 6 ? 256 : 128, smem));
      }
      ctx->occ_smem = smem;
    }
    occ = ctx->occ[vi];
  }
  if (occ < 1) return qs::fail(QSIM_ERR_CUDA, "tile pass does not fit on an SM");
  uint64_t grid = (uint64_t)ctx->sms * (uint64_t)occ;
  if (grid > ntiles) grid = ntiles;
  QsTmaGeom geom;
  CUtensorMap map;
  make_tma_geom(P, state, n, &geom, &map);
  variants[vi]<<<(unsigned)grid, threads, smem, stream>>>(state, P, map, geom, ntiles);
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

// One k-qubit exchange as ONE kernel over NVLink peer memory: gather, transfer and scatter
// fused.  With k rank qubits trading places with k local qubits, rank R swaps, with each of the
// 2^k - 1 partners R ^ d, the block of its shard whose k local bits spell the PARTNER's rank
// bits against the partner's block whose local bits spell R's -- in place on both sides.
// Every pair of amplitudes is handled by exactly one of the two ranks (alternating 512-byte
// runs: the lower rank takes the even ones), which loads its own element and the partner's
// (P2P load), and stores them crosswise (P2P store): each element crosses NVLink once, is
// read from HBM once and written once, and no staging buffer exists.  The caller brackets the
// launch with two stream-ordered barriers over the ranks (sharded.py).
struct ExchGeom {
  int k, n_local;
  int pos[8];                 // local bit positions, ascending
  int mine[8];                // this rank's value of the rank bit traded against pos[i]
  qs_c128* peer[256];         // peer[d]: partner shard for the bit pattern d (over pos[] order); [0] unused
  unsigned char higher[256];  // higher[d] = 1 if this rank's number is above the partner's
  int chunk_log2;             // > 0: work is dealt to the partners in chunks of 2^chunk_log2 pairs (round robin),
                              // so that all 2^k - 1 partners are being talked to at any time; 0: partner after partner
};
constexpr int kExchSplitBit = 5;

// U pairs per thread and trip: all 2 U loads (U of them over NVLink) are issued before the
// first store, which is what keeps enough bytes in flight to cover the NVLink round trip.
template <int U>
__global__ void __launch_bounds__(256)
k_exchange_p2p(qs_c128* __restrict__ shard, const __grid_constant__ ExchGeom G) {
  const int k = G.k;
  const uint64_t half = 1ull << (G.n_local - k - 1);          // amplitudes per partner that THIS rank moves
  const uint64_t total = half * ((1ull << k) - 1ull);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * U;
  for (uint64_t w0 = (uint64_t)blockIdx.x * blockDim.x * U + threadIdx.x; w0 < total; w0 += stride) {
    qs_c128* mine[U];
    qs_c128* remote[U];
    qs_c128 x[U], y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t w = w0 + (uint64_t)u * blockDim.x;
      mine[u] = nullptr;
      if (w >= total) continue;
      int d;
      uint64_t q;
      if (G.chunk_log2 > 0) {
        const uint64_t chunk = w >> G.chunk_log2, nd = (1ull << k) - 1ull;
        d = 1 + (int)(chunk % nd);
        q = ((chunk / nd) << G.chunk_log2) | (w & ((1ull << G.chunk_log2) - 1ull));
      } else {
        d = 1 + (int)(w / half);
        q = w % half;
      }
      // r: index inside the block; bit kExchSplitBit says which of the two ranks moves it
      const uint64_t low = q & ((1ull << kExchSplitBit) - 1ull);
      const uint64_t r = ((q >> kExchSplitBit) << (kExchSplitBit + 1)) | ((uint64_t)G.higher[d] << kExchSplitBit) | low;
      uint64_t mine_idx = r, theirs_idx = r;
#pragma unroll 1
      for (int i = 0; i < k; ++i) {
        const uint64_t lo_m = mine_idx & ((1ull << G.pos[i]) - 1ull);
        const uint64_t lo_t = theirs_idx & ((1ull << G.pos[i]) - 1ull);
        const uint64_t vd = (uint64_t)(G.mine[i] ^ ((d >> i) & 1));    // partner's rank bit
        mine_idx = ((mine_idx >> G.pos[i]) << (G.pos[i] + 1)) | (vd << G.pos[i]) | lo_m;
        theirs_idx = ((theirs_idx >> G.pos[i]) << (G.pos[i] + 1)) | ((uint64_t)G.mine[i] << G.pos[i]) | lo_t;
      }
      mine[u] = shard + mine_idx;
      remote[u] = G.peer[d] + theirs_idx;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (mine[u]) { y[u] = *remote[u]; x[u] = *mine[u]; }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (mine[u]) { *remote[u] = x[u]; *mine[u] = y[u]; }
  }
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
  std::lock_guard<std::mutex> guard(ctx->reduce_lock);
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

// =====================================================================================
// Pauli-trajectory batch: one CTA per shot, the ket (n <= 12 qubits) lives in shared memory
// =====================================================================================
// Every shot runs the same gate list; its random X / Z flips (bits prepared by the host from
// the caller's generator) are folded into the rows of each gate's matrix, so a noisy shot
// costs exactly what the noiseless circuit costs.  Outputs: |<obs|psi>|^2 per shot and the
// sum over shots of |amplitude|^2 (each CTA accumulates its shots in registers and adds once).
__global__ void __launch_bounds__(256)
k_traj_batch(int n, int64_t shots, int nops, const QsTrajOp* __restrict__ ops, const double* __restrict__ mats,
             const uint8_t* __restrict__ flips, int64_t flips_per_shot, const qs_c128* __restrict__ psi0,
             const qs_c128* __restrict__ obs, double* out_fid, double* out_prob, qs_c128* out_states) {
  extern __shared__ __align__(16) unsigned char qs_traj_smem[];
  qs_c128* state = reinterpret_cast<qs_c128*>(qs_traj_smem);
  __shared__ qs_c128 M[16];
  __shared__ double red[2 * 8];
  const uint32_t tid = threadIdx.x;
  const uint64_t dim = 1ull << n;
  double acc[16];                                   // n <= 12: at most 16 amplitudes per thread
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.0;
  for (int64_t shot = blockIdx.x; shot < shots; shot += gridDim.x) {
    for (uint64_t i = tid; i < dim; i += 256) state[i] = psi0[i];
    const uint8_t* fl = flips + shot * flips_per_shot;
    __syncthreads();
    for (int o = 0; o < nops; ++o) {
      const QsTrajOp op = ops[o];
      if (tid == 0) qs_traj_rows(op, mats, fl, M);
      fl += 2 * op.k;
      __syncthreads();
      qs_traj_apply(state, n, op, M, tid, 256);
      __syncthreads();
    }
    if (out_states)
      for (uint64_t i = tid; i < dim; i += 256) out_states[(uint64_t)shot * dim + i] = state[i];
    double fr = 0.0, fi = 0.0;
    int e = 0;
    for (uint64_t i = tid; i < dim; i += 256, ++e) {
      const qs_c128 v = state[i];
      acc[e] += v.x * v.x + v.y * v.y;
      if (obs) { const qs_c128 w = obs[i]; fr += w.x * v.x + w.y * v.y; fi += w.x * v.y - w.y * v.x; }
    }
    if (obs && out_fid) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        fr += __shfl_down_sync(0xffffffffu, fr, off);
        fi += __shfl_down_sync(0xffffffffu, fi, off);
      }
      if ((tid & 31u) == 0) { red[2 * (tid >> 5)] = fr; red[2 * (tid >> 5) + 1] = fi; }
      __syncthreads();
      if (tid == 0) {
        double sr = 0.0, si = 0.0;
        for (int w = 0; w < 8; ++w) { sr += red[2 * w]; si += red[2 * w + 1]; }
        out_fid[shot] = sr * sr + si * si;
      }
    }
    __syncthreads();
  }
  if (out_prob) {
    int e = 0;
    for (uint64_t i = tid; i < dim; i += 256, ++e) atomicAdd(out_prob + i, acc[e]);
  }
}

// =====================================================================================
// Dense k-qubit block, 5 <= k <= 10 (Gate.apply with any k, DV/gates.py:44-54): in place,
// one pass over the state, FP64 tensor-core MMA (DMMA.8x8x4)
// =====================================================================================
// A tile is the 2^k amplitudes of the gate's qubits times 8 "columns" (the three lowest
// index bits the gate does not touch, so rows of 8 amplitudes are 128 contiguous bytes
// whenever the gate leaves bits 0..2 alone).  With the complex matrix split into real and
// imaginary parts the update is the real GEMM
//     out_re = Mr in_re - Mi in_im,   out_im = Mr in_im + Mi in_re        (2^k x 2^k) x (2^k x 8)
// i.e. four m8n8k4 MMAs per 8-row block and 4-column chunk of M.  The whole tile is staged in
// shared memory first, every warp keeps its 8x8 output block in registers and writes it
// straight back to the addresses it came from: no scratch buffer, no second pass.
struct DenseGeom {
  int k, ncol, T;
  uint8_t gate_bits[10];     // index bit of matrix factor f (f = 0: most significant bit of the row index)
  uint8_t col_bits[3];       // ascending
  uint8_t tile_bits[16];     // ascending union of the two
};

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ uint64_t dense_index(const DenseGeom& G, uint64_t base, int row, int col) {
  uint64_t idx = base;
  for (int f = 0; f < G.k; ++f) idx |= (uint64_t)((row >> (G.k - 1 - f)) & 1) << G.gate_bits[f];
  for (int b = 0; b < G.ncol; ++b) idx |= (uint64_t)((col >> b) & 1) << G.col_bits[b];
  return idx;
}

__global__ void __launch_bounds__(256)
k_dense_block(qs_c128* state, const double* __restrict__ mat, const __grid_constant__ DenseGeom G, uint64_t ntiles) {
  extern __shared__ __align__(16) unsigned char qs_dense_smem[];
  qs_c128* tile = reinterpret_cast<qs_c128*>(qs_dense_smem);   // slot of (c, col): c * 8 + (col ^ ((c & 3) << 1))
  const int k = G.k, dim = 1 << k, ncols = 1 << G.ncol;
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (int)blockDim.x >> 5;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    uint64_t base = t;
    for (int l = 0; l < G.T; ++l) {
      const uint32_t pos = G.tile_bits[l];
      const uint64_t low = base & ((1ull << pos) - 1ull);
      base = ((base >> pos) << (pos + 1)) | low;
    }
    for (int e = tid; e < dim * 8; e += (int)blockDim.x) {
      const int c = e >> 3, col = e & 7;
      qs_c128 v; v.x = 0.0; v.y = 0.0;
      if (col < ncols) v = state[dense_index(G, base, c, col)];
      tile[c * 8 + (col ^ ((c & 3) << 1))] = v;
    }
    __syncthreads();
    for (int rb = warp; rb < dim / 8; rb += nwarps) {
      double re0 = 0.0, re1 = 0.0, im0 = 0.0, im1 = 0.0;
      const int row = rb * 8 + (lane >> 2);
      const double2* mrow = reinterpret_cast<const double2*>(mat) + ((size_t)row << k);
      const int colb = lane >> 2;
#pragma unroll 4
      for (int chunk = 0; chunk < dim / 4; ++chunk) {
        const int c = chunk * 4 + (lane & 3);
        const double2 mv = __ldg(mrow + c);                              // M[row][c]
        const qs_c128 b = tile[c * 8 + (colb ^ ((c & 3) << 1))];        // in[c][colb]
        dmma_8x8x4(re0, re1, mv.x, b.x);
        dmma_8x8x4(re0, re1, -mv.y, b.y);
        dmma_8x8x4(im0, im1, mv.x, b.y);
        dmma_8x8x4(im0, im1, mv.y, b.x);
      }
      const int col0 = 2 * (lane & 3);
      if (col0 < ncols) {
        qs_c128 o; o.x = re0; o.y = im0;
        state[dense_index(G, base, row, col0)] = o;
      }
      if (col0 + 1 < ncols) {
        qs_c128 o; o.x = re1; o.y = im1;
        state[dense_index(G, base, row, col0 + 1)] = o;
      }
    }
    __syncthreads();
  }
}

std::mutex g_dense_upload_lock;

int run_dense_block(DevCtx* ctx, const qs::PlanItem& it, qs_c128* state, int n, int device, cudaStream_t stream) {
  const qs::Op& op = it.op;
  const int k = op.k, dim = 1 << k;
  if (k < 5 || k > 10 || k > n) return qs::fail(QSIM_ERR_UNSUPPORTED, "dense block: k must be in [5, min(10, n)]");
  {
    // the matrix goes to the device once per plan (synchronously, like compiling the plan);
    // executions after the first launch without any host synchronisation
    std::lock_guard<std::mutex> guard(g_dense_upload_lock);
    if (!it.dev_mat || it.dev_index != device) {
      std::vector<double> flat(2 * (size_t)dim * dim);
      for (int e = 0; e < dim * dim; ++e) { flat[2 * e] = op.mat[e].real(); flat[2 * e + 1] = op.mat[e].imag(); }
      void* d = nullptr;
      QS_CUDA(cudaMalloc(&d, sizeof(double) * flat.size()));
      it.dev_mat = std::shared_ptr<void>(d, [](void* p) { cudaFree(p); });
      it.dev_index = device;
      QS_CUDA(cudaMemcpy(d, flat.data(), sizeof(double) * flat.size(), cudaMemcpyHostToDevice));
    }
  }
  DenseGeom G;
  memset(&G, 0, sizeof(G));
  G.k = k;
  uint64_t gmask = 0;
  for (int f = 0; f < k; ++f) { G.gate_bits[f] = (uint8_t)op.bits[f]; gmask |= 1ull << op.bits[f]; }
  for (int b = 0; b < n && G.ncol < 3; ++b)
    if (!(gmask >> b & 1)) G.col_bits[G.ncol++] = (uint8_t)b;
  uint64_t tmask = gmask;
  for (int i = 0; i < G.ncol; ++i) tmask |= 1ull << G.col_bits[i];
  for (int b = 0; b < n; ++b)
    if (tmask >> b & 1) G.tile_bits[G.T++] = (uint8_t)b;
  const uint64_t ntiles = 1ull << (n - G.T);
  const int smem = (int)sizeof(qs_c128) * dim * 8;
  const int threads = dim / 8 >= 8 ? 256 : 32 * (dim / 8);
  if (smem > 48 * 1024) QS_CUDA(cudaFuncSetAttribute(k_dense_block, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int occ = 0;
  QS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dense_block, threads, smem));
  if (occ < 1) return qs::fail(QSIM_ERR_CUDA, "dense block does not fit on an SM");
  uint64_t grid = (uint64_t)ctx->sms * (uint64_t)occ;
  if (grid > ntiles) grid = ntiles;
  k_dense_block<<<(unsigned)grid, threads, smem, stream>>>(state, (const double*)it.dev_mat.get(), G, ntiles);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
  return QSIM_OK;
}

int execute_plan(const qsim_plan* p, void* state, int n, void* scratch, void* stream) {
  if (!p || !state) return qs::fail(QSIM_ERR_ARG, "qsim_plan_execute: null argument");
  if (n != p->n) return qs::fail(QSIM_ERR_ARG, "qsim_plan_execute: plan was compiled for a different qubit count");
  Bound bound;
  int rc = bind_device(state, &bound);
  DevCtx* ctx = bound.ctx;
  if (rc != QSIM_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  for (const qs::PlanItem& it : p->items) {
    rc = it.generic ? run_dense_block(ctx, it, (qs_c128*)state, n, bound.device, st)
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

int qsim_ipc_export_ex(void* ptr, unsigned char* handle64, uint64_t* offset) {
  if (!ptr || !handle64 || !offset) return qs::fail(QSIM_ERR_ARG, "qsim_ipc_export_ex: null argument");
  cudaIpcMemHandle_t h;
  QS_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, 64);
  // the handle names the whole allocation: find its base through the driver
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn range = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    cudaGetLastError();
    return (RangeFn)p;
  }();
  if (!range) return qs::fail(QSIM_ERR_CUDA, "cuMemGetAddressRange is not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  if (range(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS)
    return qs::fail(QSIM_ERR_CUDA, "cuMemGetAddressRange failed");
  *offset = (uint64_t)((CUdeviceptr)ptr - base);
  return QSIM_OK;
}

int qsim_exchange_p2p(void* shard, void* const* peer_shards, int n_local, int nbits, const int* local_qubits,
                      const int* my_bits, const int* partner_is_lower, void* stream) {
  if (!shard || !peer_shards || !local_qubits || !my_bits || !partner_is_lower)
    return qs::fail(QSIM_ERR_ARG, "qsim_exchange_p2p: null argument");
  if (nbits < 1 || nbits > 8 || n_local - nbits - 1 < kExchSplitBit)
    return qs::fail(QSIM_ERR_ARG, "qsim_exchange_p2p: bad sizes (blocks must hold at least 64 amplitudes)");
  ExchGeom G;
  memset(&G, 0, sizeof(G));
  G.k = nbits;
  G.n_local = n_local;
  // order the bits by local position (ascending); pattern bit i of d stays with the caller's bit i
  int order[8];
  for (int i = 0; i < nbits; ++i) order[i] = i;
  for (int i = 1; i < nbits; ++i)
    for (int j = i; j > 0 && local_qubits[order[j]] > local_qubits[order[j - 1]]; --j) std::swap(order[j], order[j - 1]);
  for (int i = 0; i < nbits; ++i) {
    const int q = local_qubits[order[i]];
    if (q < 0 || q >= n_local || (my_bits[order[i]] | 1) != 1) return qs::fail(QSIM_ERR_ARG, "qsim_exchange_p2p: bad qubit or bit");
    G.pos[i] = n_local - 1 - q;                    // reference-style local qubit -> index bit
    G.mine[i] = my_bits[order[i]];
    if (i > 0 && G.pos[i] <= G.pos[i - 1]) return qs::fail(QSIM_ERR_ARG, "qsim_exchange_p2p: repeated qubit");
  }
  for (int d = 1; d < (1 << nbits); ++d) {
    // the caller's pattern d (bit i <-> its bit i) in sorted order
    int ds = 0;
    for (int i = 0; i < nbits; ++i) ds |= ((d >> order[i]) & 1) << i;
    if (!peer_shards[d]) return qs::fail(QSIM_ERR_ARG, "qsim_exchange_p2p: missing peer pointer");
    G.peer[ds] = (qs_c128*)peer_shards[d];
    G.higher[ds] = partner_is_lower[d] ? 1 : 0;
  }
  Bound bound;
  int rc = bind_device(shard, &bound);
  if (rc != QSIM_OK) return rc;
  const uint64_t total = (1ull << (n_local - nbits - 1)) * ((1ull << nbits) - 1ull);
  // development knobs (defaults measured on 2 and 8 B200s): pairs per thread and trip, CTAs per SM
  static const int batch = qs::dev_knob("QSIM_EXCH_BATCH", 4);
  static const int per_sm = qs::dev_knob("QSIM_EXCH_CTAS", 8);
  static const int chunk_log2 = qs::dev_knob("QSIM_EXCH_CHUNK_LOG2", 0);
  G.chunk_log2 = (chunk_log2 > 0 && chunk_log2 <= n_local - nbits - 1 && nbits > 1) ? chunk_log2 : 0;
  const int U = batch >= 8 ? 8 : batch >= 4 ? 4 : batch >= 2 ? 2 : 1;
  uint64_t blocks = (total + 256ull * U - 1) / (256ull * U);
  const uint64_t cap = (uint64_t)bound.ctx->sms * (uint64_t)(per_sm > 0 ? per_sm : 8);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (U == 8) k_exchange_p2p<8><<<(unsigned)blocks, 256, 0, st>>>((qs_c128*)shard, G);
  else if (U == 4) k_exchange_p2p<4><<<(unsigned)blocks, 256, 0, st>>>((qs_c128*)shard, G);
  else if (U == 2) k_exchange_p2p<2><<<(unsigned)blocks, 256, 0, st>>>((qs_c128*)shard, G);
  else k_exchange_p2p<1><<<(unsigned)blocks, 256, 0, st>>>((qs_c128*)shard, G);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
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
  Bound bound;
  int rc = bind_device(state, &bound);
  DevCtx* ctx = bound.ctx;
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
  Bound bound;
  int rc = bind_device(state, &bound);
  DevCtx* ctx = bound.ctx;
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
  Bound bound;
  int rc = bind_device(in, &bound);
  DevCtx* ctx = bound.ctx;
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
  Bound bound;
  int rc = bind_device(in, &bound);
  DevCtx* ctx = bound.ctx;
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
  Bound bound;
  int rc = bind_device(state, &bound);
  DevCtx* ctx = bound.ctx;
  if (rc != QSIM_OK) return rc;
  FnNorm2 f{(const qs_c128*)state};
  double r[2];
  rc = reduce_to_host(ctx, f, n_amps, r, (cudaStream_t)stream);
  if (rc == QSIM_OK) *out = r[0];
  return rc;
}

int qsim_reduce_inner(const void* a, const void* b, uint64_t n_amps, double* out_re_im, void* stream) {
  if (!a || !b || !out_re_im) return qs::fail(QSIM_ERR_ARG, "qsim_reduce_inner: null argument");
  Bound bound;
  int rc = bind_device(a, &bound);
  DevCtx* ctx = bound.ctx;
  if (rc != QSIM_OK) return rc;
  FnInner f{(const qs_c128*)a, (const qs_c128*)b};
  return reduce_to_host(ctx, f, n_amps, out_re_im, (cudaStream_t)stream);
}

int qsim_reduce_expect(const void* ket, const void* rho, int n_qubits, double* out_re_im, void* stream) {
  if (!ket || !rho || !out_re_im || n_qubits < 0 || n_qubits > 20) return qs::fail(QSIM_ERR_ARG, "qsim_reduce_expect: bad argument");
  Bound bound;
  int rc = bind_device(rho, &bound);
  DevCtx* ctx = bound.ctx;
  if (rc != QSIM_OK) return rc;
  FnExpect f{(const qs_c128*)ket, (const qs_c128*)rho, n_qubits};
  return reduce_to_host(ctx, f, 1ull << (2 * n_qubits), out_re_im, (cudaStream_t)stream);
}

int qsim_reduce_purity(const void* rho, int n_qubits, double* out_re_im, void* stream) {
  if (!rho || !out_re_im || n_qubits < 0 || n_qubits > 20) return qs::fail(QSIM_ERR_ARG, "qsim_reduce_purity: bad argument");
  Bound bound;
  int rc = bind_device(rho, &bound);
  DevCtx* ctx = bound.ctx;
  if (rc != QSIM_OK) return rc;
  FnPurity f{(const qs_c128*)rho, n_qubits};
  return reduce_to_host(ctx, f, 1ull << (2 * n_qubits), out_re_im, (cudaStream_t)stream);
}

int qsim_reduce_trace(const void* rho, int n_qubits, double* out_re_im, void* stream) {
  if (!rho || !out_re_im || n_qubits < 0 || n_qubits > 20) return qs::fail(QSIM_ERR_ARG, "qsim_reduce_trace: bad argument");
  Bound bound;
  int rc = bind_device(rho, &bound);
  DevCtx* ctx = bound.ctx;
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
  Bound bound;
  int rc = bind_device(out_fidelity, &bound);
  DevCtx* ctx = bound.ctx;
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

int qsim_traj_batch(int n_qubits, int64_t shots, int n_ops, const int32_t* ops, const double* matrices,
                    const uint8_t* flips, int64_t flips_per_shot, const double* psi0, const double* observable,
                    double* out_fidelity, double* out_prob_sum, double* out_states, void* stream) {
  if (!ops || !matrices || !flips || !psi0) return qs::fail(QSIM_ERR_ARG, "qsim_traj_batch: null argument");
  if (n_qubits < 1 || n_qubits > 12) return qs::fail(QSIM_ERR_UNSUPPORTED, "qsim_traj_batch: 1 <= n_qubits <= 12");
  if (shots < 0 || n_ops < 0 || flips_per_shot < 0) return qs::fail(QSIM_ERR_ARG, "qsim_traj_batch: bad sizes");
  if (shots == 0) return QSIM_OK;
  Bound bound;
  int rc = bind_device(psi0, &bound);
  if (rc != QSIM_OK) return rc;
  const int smem = (int)(sizeof(qs_c128) << n_qubits);
  if (smem > 48 * 1024) QS_CUDA(cudaFuncSetAttribute(k_traj_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int occ = 0;
  QS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_traj_batch, 256, smem));
  if (occ < 1) return qs::fail(QSIM_ERR_CUDA, "trajectory batch does not fit on an SM");
  int64_t grid = (int64_t)bound.ctx->sms * occ;
  if (grid > shots) grid = shots;
  k_traj_batch<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(
      n_qubits, shots, n_ops, (const QsTrajOp*)ops, matrices, flips, flips_per_shot, (const qs_c128*)psi0,
      (const qs_c128*)observable, out_fidelity, out_prob_sum, (qs_c128*)out_states);
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
  Bound bound;
  rc = bind_device(shard, &bound);
  DevCtx* ctx = bound.ctx;
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
  Bound bound;
  rc = bind_device(shard, &bound);
  DevCtx* ctx = bound.ctx;
  if (rc != QSIM_OK) return rc;
  if (count == 0) return QSIM_OK;
  k_swap_unpack<<<stream_grid(ctx, count, 256), 256, 0, (cudaStream_t)stream>>>(
      (qs_c128*)shard, (const qs_c128*)recvbuf, sel, first, count);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  QS_CUDA(cudaGetLastError());
  return QSIM_OK;
}

}  // extern "C"
