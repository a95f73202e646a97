// Hand-written int8 x int8 -> int32 GEMM on the 5th-generation tensor cores (tcgen05, sm_100a) for the int8-sliced
// gemm_nonlop (ozaki.cu):     C(M x N, int32, column-major ldc) = A^T B,
// A: M rows of K int8 (K contiguous, row pitch lda), B: N rows of K int8 (row pitch ldb) -- both operands "K-major".
//
// One CTA computes a 128 x BN tile: TMA (cp.async.bulk.tensor, 128-byte swizzle) brings 128 x 128-byte A tiles and
// BN x 128-byte B tiles into a 4-stage shared-memory ring (full / empty mbarriers); one elected thread issues
// tcgen05.mma.cta_group::1.kind::i8 (M = 128, N = BN, K = 32 per instruction, 4 per stage) accumulating int32 in TMEM;
// tcgen05.commit releases ring slots and finally signals the epilogue, in which the 4 warps read their 32 TMEM lanes
// (tcgen05.ld 32x32b) and store coalesced columns of C.  Out-of-range rows of A / B are zero-filled by TMA; stores are
// predicated, so M, N need not be multiples of the tile.  K must be a multiple of 128 (the slicing kernels pad with zeros).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>

namespace abi {

namespace igemm_detail {

constexpr int BM = 128, BK = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major operand tile, 128-byte swizzle, rows of 128 bytes packed densely (8-row groups 1024 bytes apart)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  const uint64_t lo = (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16);              // start address, LBO = 1 (unused for swizzled K-major)
  const uint64_t hi = 64ull | (1ull << 14) | (2ull << 29);                            // SBO = 1024 B, version 1, SWIZZLE_128B
  return lo | (hi << 32);
}

// STAGES = 4, 1 CTA/SM: deepest ring for long K loops; STAGES = 2, 2 CTAs/SM: the second CTA's loads and MMAs cover the first
// one's prologue / epilogue when the K loop is short (opernlb: K = nprojs).  nsplit > 1: split-K with exact integer atomics
// (C zeroed by the caller) when there are too few output tiles to fill the chip.
template <int BN, int STAGES, int CTAS>
__global__ void __launch_bounds__(128, CTAS) k_igemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                         int32_t* __restrict__ C, long long ldc, int M, int N, int nkb_total, int tiles_n,
                                                         int tiles, int nsplit) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned operand ring
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK, STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x % tiles, z = blockIdx.x / tiles;
  const int tn = tile % tiles_n, tm = tile / tiles_n;
  const int m0 = tm * BM, n0 = tn * BN;
  const int kb_per = (nkb_total + nsplit - 1) / nsplit;
  const int kb0 = z * kb_per, nkb = max(0, min(nkb_total, kb0 + kb_per) - kb0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {   // TMEM allocation: BN columns of 32-bit accumulators (power of two >= 32), by one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer ----------------
    for (int kb = 0; kb < nkb; kb++) {
      const int s = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_expect_tx(&full[s], STAGE_BYTES);
      uint8_t* st = smem + s * STAGE_BYTES;
      tma_load_2d(st, &map_a, (kb0 + kb) * BK, m0, &full[s]);
      tma_load_2d(st + A_BYTES, &map_b, (kb0 + kb) * BK, n0, &full[s]);
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------- MMA issuer ----------------
    // instruction descriptor: C = S32, A = B = signed 8 bit, both K-major, N = BN, M = 128
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int kb = 0; kb < nkb; kb++) {
      const int s = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
      const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + A_BYTES);
#pragma unroll
      for (int k = 0; k < BK / 32; k++)    // 32 bytes of K per instruction: the start address advances by 2 (x16 bytes)
        umma_i8(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
      umma_commit(&empty[s]);              // slot reusable once these MMAs have read it
    }
    umma_commit(tmem_full);                // accumulator complete
  }
  __syncwarp();
  if (nkb == 0) {                            // empty K range of a split: nothing to add
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    return;
  }
  // ---------------- epilogue: all 4 warps, warp w owns TMEM lanes 32w .. 32w+31 = rows m0 + 32w + lane ----------------
  mbar_wait(tmem_full, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int m = m0 + warp * 32 + lane;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (m < M) {
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const int n = n0 + c0 + j;
        if (n < N) {
          if (nsplit == 1) C[(long long)n * ldc + m] = (int32_t)r[j];
          else atomicAdd(&C[(long long)n * ldc + m], (int32_t)r[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) {
      fprintf(stderr, "igemm_tc: cuTensorMapEncodeTiled not available\n"); abort();
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// rows x K int8 matrix, K contiguous, row pitch ld bytes; box = 128 bytes of K x box_rows rows, 128-byte swizzle
inline CUtensorMap make_map(const int8_t* base, long long rows, long long K, long long ld, int box_rows) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "igemm_tc: cuTensorMapEncodeTiled failed (%d)\n", (int)r); abort(); }
  return m;
}


template <int BN, int STAGES, int CTAS>
inline void launch_igemm(const CUtensorMap& ma, const CUtensorMap& mb, int32_t* C, long long ldc, int M, int N, int nkb, int nsplit,
                         cudaStream_t st) {
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN, tiles = tiles_m * tiles_n;
  const size_t smem = (size_t)STAGES * (BM * BK + BN * BK) + 1024 + 256;
  static bool done = false;
  auto chk = [](cudaError_t e, const char* what) { if (e != cudaSuccess) { fprintf(stderr, "igemm_tc: %s: %s\n", what, cudaGetErrorString(e)); abort(); } };
  if (!done) { chk(cudaFuncSetAttribute(k_igemm_tc<BN, STAGES, CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "shared-memory opt-in"); done = true; }
  if (nsplit > 1) chk(cudaMemsetAsync(C, 0, sizeof(int32_t) * (size_t)ldc * N, st), "memset");
  k_igemm_tc<BN, STAGES, CTAS><<<tiles * nsplit, 128, smem, st>>>(ma, mb, C, ldc, M, N, nkb, tiles_n, tiles, nsplit);
  chk(cudaGetLastError(), "launch");
}

}  // namespace igemm_detail

// C(M x N) = A^T B ; K % 128 == 0, lda % 16 == 0, ldb % 16 == 0, A / B 16-byte aligned.
// variant: 0 auto ; 1: 256-wide tiles, 4 stages, 1 CTA/SM ; 2: 256-wide, 2 stages, 2 CTAs/SM ; 3: 128-wide, 3 stages, 2 CTAs/SM
inline void igemm_tc(int M, int N, int K, const int8_t* A, long long lda, const int8_t* B, long long ldb, int32_t* C, long long ldc,
                     cudaStream_t st, int variant = 0) {
  using namespace igemm_detail;
  if (M == 0 || N == 0) return;
  if (K % BK != 0 || lda % 16 != 0 || ldb % 16 != 0) { fprintf(stderr, "igemm_tc: K must be a multiple of 128 and the pitches of 16\n"); abort(); }
  const int nkb = K / BK;
  if (variant == 0) variant = (N <= 128) ? 3 : (nkb >= 512 ? 1 : 2);
  const int bn = (variant == 3) ? 128 : 256;
  const CUtensorMap ma = make_map(A, M, K, lda, BM), mb = make_map(B, N, K, ldb, bn);
  const int tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
  const int slots = 148 * (variant == 1 ? 1 : 2);
  int nsplit = 1;
  if (tiles < slots && nkb >= 64) nsplit = std::min(std::max(1, slots / tiles), std::max(1, nkb / 32));
  if (variant == 1) launch_igemm<256, 4, 1>(ma, mb, C, ldc, M, N, nkb, nsplit, st);
  else if (variant == 2) launch_igemm<256, 2, 2>(ma, mb, C, ldc, M, N, nkb, nsplit, st);
  else launch_igemm<128, 3, 2>(ma, mb, C, ldc, M, N, nkb, nsplit, st);
}

}  // namespace abi
