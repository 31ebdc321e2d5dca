// Development probe (not part of the product): TMA tile fill / store of a state-vector tile
// whose index bits are scattered, with the 128-byte shared-memory swizzle.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/tma_probe tools/tma_probe.cu
//   tools/_build/tma_probe [n=28]
//
// Checks (1) that the tile lands in shared memory where the kernel expects it
// (slot(j) = j ^ ((j >> 3) & 7), 16-byte slots), (2) that load + store round-trips the
// state, and times a pass that does nothing else: the memory-side ceiling of the
// tile pass for several tile-bit sets.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

struct TileMap {
  int T;               // tile bits
  int P;               // tile positions covered by one box
  int nops;            // 2^(T-P) boxes per tile
  int shift[5];        // first index bit of TMA dimension i
  uint32_t mask[5];    // extent-1 of dimension i (in amplitudes)
  uint8_t tile_bits[16];
};

struct c128 { double x, y; };

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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
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

__device__ __forceinline__ uint64_t tile_base(const TileMap& M, uint64_t tile) {
  uint64_t base = tile;
  for (int l = 0; l < M.T; ++l) {
    const uint32_t pos = M.tile_bits[l];
    const uint64_t low = base & ((1ull << pos) - 1ull);
    base = ((base >> pos) << (pos + 1)) | low;
  }
  return base;
}

// mode bit 0: verify the shared-memory layout (errors counted in *err)
// mode bit 1: add `work` shared-memory round trips per tile (LDS.128 + STS.128 per amplitude)
__global__ void __launch_bounds__(256, 3)
k_probe(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
        const __grid_constant__ TileMap M, uint64_t ntiles, int mode, int work, unsigned long long* err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  c128* tile = reinterpret_cast<c128*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  const uint32_t tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;
  const uint32_t box_bytes = 16u << M.P;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint64_t base = tile_base(M, t);
    if (tid < 32) {
      tma_wait_read0();                            // my stores of the previous tile have left the buffer
      __syncwarp();
      if (tid == 0) mbar_expect_tx(&bar, 16u << M.T);
      __syncwarp();
      for (int o = (int)tid; o < M.nops; o += 32) {
        uint64_t b = base;
        for (int q = 0; q < M.T - M.P; ++q) b |= (uint64_t)((o >> q) & 1) << M.tile_bits[M.P + q];
        tma_load_5d(smem_raw + (size_t)o * box_bytes, &map_in, &bar, (int)(((b >> M.shift[0]) & M.mask[0]) * 2),
                    (int)((b >> M.shift[1]) & M.mask[1]), (int)((b >> M.shift[2]) & M.mask[2]),
                    (int)((b >> M.shift[3]) & M.mask[3]), (int)((b >> M.shift[4]) & M.mask[4]));
      }
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    if (mode & 1) {
      for (uint32_t j = tid; j < (1u << M.T); j += 256) {
        uint64_t g = base;
        for (int l = 0; l < M.T; ++l) g |= (uint64_t)((j >> l) & 1) << M.tile_bits[l];
        const uint32_t slot = j ^ ((j >> 3) & 7u);
        const c128 v = tile[slot];
        if (v.x != (double)g || v.y != -(double)g) atomicAdd(err, 1ull);
      }
    }
    if (mode & 2) {
      for (int w = 0; w < work; ++w) {
        c128 a[8];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
          for (int m = 0; m < 8; ++m) a[m] = tile[((tid + 256 * (m + 8 * i)) ^ (uint32_t)w) & ((1u << M.T) - 1u)];
#pragma unroll
          for (int m = 0; m < 8; ++m) tile[((tid + 256 * (m + 8 * i)) ^ (uint32_t)w) & ((1u << M.T) - 1u)] = a[m];
        }
        __syncthreads();
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid < 32) {
      for (int o = (int)tid; o < M.nops; o += 32) {
        uint64_t b = base;
        for (int q = 0; q < M.T - M.P; ++q) b |= (uint64_t)((o >> q) & 1) << M.tile_bits[M.P + q];
        tma_store_5d(&map_out, smem_raw + (size_t)o * box_bytes, (int)(((b >> M.shift[0]) & M.mask[0]) * 2),
                     (int)((b >> M.shift[1]) & M.mask[1]), (int)((b >> M.shift[2]) & M.mask[2]),
                     (int)((b >> M.shift[3]) & M.mask[3]), (int)((b >> M.shift[4]) & M.mask[4]));
      }
      tma_commit();
    }
  }
  if (tid < 32) tma_wait_read0();
}

__global__ void k_fill(c128* s, uint64_t count) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
    s[i].x = (double)i;
    s[i].y = -(double)i;
  }
}
__global__ void k_check(const c128* s, uint64_t count, unsigned long long* err) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    if (s[i].x != (double)i || s[i].y != -(double)i) atomicAdd(err, 1ull);
}

static PFN_cuTensorMapEncodeTiled get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) {
    fprintf(stderr, "cuTensorMapEncodeTiled not found\n");
    exit(1);
  }
  return (PFN_cuTensorMapEncodeTiled)fn;
}

// Dimension ranges for a tile-bit set: range 0 = [0, cut1) with a box of bits 0..2, ranges
// 1..4 start at the four lowest runs of tile bits above bit 2.
static bool make_map(PFN_cuTensorMapEncodeTiled enc, void* state, int n, const std::vector<int>& bits, TileMap* M,
                     CUtensorMap* map) {
  memset(M, 0, sizeof(*M));
  M->T = (int)bits.size();
  for (int i = 0; i < M->T; ++i) M->tile_bits[i] = (uint8_t)bits[i];
  std::vector<char> is_tile(n + 1, 0);
  for (int b : bits) is_tile[b] = 1;
  if (!(is_tile[0] && is_tile[1] && is_tile[2])) return false;
  std::vector<int> cuts;          // first bits of ranges 1..4
  {
    int run = 0;
    for (int b = 3; b < n && (int)cuts.size() < 4; ++b) {
      if (is_tile[b] && (run == 0 || run == 8)) { cuts.push_back(b); run = 1; }
      else if (is_tile[b]) ++run;
      else run = 0;
    }
    // fillers: any unused positions above bit 3
    for (int b = n - 1; b > 3 && (int)cuts.size() < 4; --b)
      if (std::find(cuts.begin(), cuts.end(), b) == cuts.end()) cuts.push_back(b);
    std::sort(cuts.begin(), cuts.end());
  }
  if (cuts.size() != 4) return false;
  int start[6] = {0, cuts[0], cuts[1], cuts[2], cuts[3], n};
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  int P = 3;
  bool contiguous = true;   // box positions must be the lowest tile positions, in order
  for (int i = 0; i < 5; ++i) {
    const int lo = start[i], hi = start[i + 1];
    int nb = 0;
    if (i == 0) nb = 3;
    else {
      while (lo + nb < hi && is_tile[lo + nb] && nb < 8) ++nb;
      if (!contiguous) nb = 0;
      // a tile bit of this range that the box does not cover ends the covered prefix
      for (int b = lo + nb; b < hi; ++b)
        if (is_tile[b]) contiguous = false;
      P += nb;
    }
    if (i == 0)
      for (int b = 3; b < hi; ++b)
        if (is_tile[b]) return false;
    M->shift[i] = lo;
    M->mask[i] = (uint32_t)((1ull << (hi - lo)) - 1ull);
    gdim[i] = (1ull << (hi - lo)) * (i == 0 ? 2 : 1);
    box[i] = (1u << nb) * (i == 0 ? 2 : 1);
    if (i > 0) gstride[i - 1] = 16ull << lo;
  }
  M->P = P;
  M->nops = 1 << (M->T - P);
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, state, gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return false;
  }
  return true;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 28;
  const uint64_t count = 1ull << n;
  c128 *a = nullptr, *b = nullptr;
  CK(cudaMalloc(&a, count * 16));
  CK(cudaMalloc(&b, count * 16));
  unsigned long long* err = nullptr;
  CK(cudaMalloc(&err, 8));
  PFN_cuTensorMapEncodeTiled enc = get_encode();
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));

  std::vector<std::vector<int>> sets = {
      {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11},
      {0, 1, 2, 3, 7, 12, 13, 19, 22, 24, 25, 27},
      {0, 1, 2, 3, 5, 8, 11, 14, 17, 20, 23, 26},
      {0, 1, 2, 5, 8, 11, 14, 17, 20, 23, 25, 27},
      {0, 1, 2, 3, 16, 17, 18, 19, 20, 21, 22, 23},
      {0, 1, 2, 3, 20, 21, 22, 23, 24, 25, 26, 27},
      {0, 1, 2, 3, 4, 20, 21, 22, 23, 24, 25},          // T = 11
  };
  for (size_t si = 0; si < sets.size(); ++si) {
    std::vector<int> bits;
    for (int x : sets[si])
      if (x < n) bits.push_back(x);
    TileMap M;
    CUtensorMap map_in, map_out;
    if (!make_map(enc, a, n, bits, &M, &map_in) || !make_map(enc, b, n, bits, &M, &map_out)) {
      printf("set %zu: cannot build a tensor map\n", si);
      continue;
    }
    const int smem = 16 << M.T;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem + 1024));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_probe, 256, smem));
    const uint64_t ntiles = count >> M.T;
    const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)sms * occ);
    // correctness: layout + round trip
    k_fill<<<sms * 8, 256>>>(a, count);
    CK(cudaMemset(b, 0, count * 16));
    CK(cudaMemset(err, 0, 8));
    k_probe<<<grid, 256, smem>>>(map_in, map_out, M, ntiles, 1, 0, err);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    unsigned long long e1 = 0, e2 = 0;
    CK(cudaMemcpy(&e1, err, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemset(err, 0, 8));
    k_check<<<sms * 8, 256>>>(b, count, err);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&e2, err, 8, cudaMemcpyDeviceToHost));
    printf("set %zu: T=%d P=%d ops/tile=%d occ=%d layout_errors=%llu roundtrip_errors=%llu\n", si, M.T, M.P, M.nops,
           occ, e1, e2);
    // timing: copy pass and copy pass + shared-memory round trips
    for (int work : {0, 2, 4, 6}) {
      cudaEvent_t e0, e3;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e3));
      const int reps = 5;
      for (int w = 0; w < 2; ++w) k_probe<<<grid, 256, smem>>>(map_in, map_out, M, ntiles, work ? 2 : 0, work, err);
      CK(cudaEventRecord(e0));
      for (int r = 0; r < reps; ++r) k_probe<<<grid, 256, smem>>>(map_in, map_out, M, ntiles, work ? 2 : 0, work, err);
      CK(cudaEventRecord(e3));
      CK(cudaEventSynchronize(e3));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e3));
      ms /= reps;
      printf("   work=%d: %.3f ms/pass, %.1f GB/s (read+write)\n", work, ms, 2.0 * 16.0 * count / ms * 1e-6);
    }
  }
  // plain copy for reference
  {
    cudaEvent_t e0, e3;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e3));
    CK(cudaMemcpy(b, a, count * 16, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 5; ++r) CK(cudaMemcpyAsync(b, a, count * 16, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e3));
    CK(cudaEventSynchronize(e3));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e3));
    printf("cudaMemcpy D2D: %.3f ms, %.1f GB/s (read+write)\n", ms / 5, 2.0 * 16.0 * count / (ms / 5) * 1e-6);
  }
  return 0;
}
