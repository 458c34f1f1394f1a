// sf_diag.cu -- bring-up self-tests, kept in the library so a failing parity test can be bisected on the
// GPU box in one call: (1) what a swizzled TMA box load actually puts in shared memory, (2) one UMMA tile
// product from TMA-loaded operands read back through tcgen05.ld.
#include <string>

#include "sf_conv.cuh"

namespace {
using namespace sf;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<EncodeTiledFn>(p);
}

__global__ void __launch_bounds__(128) diag_tma_dump_kernel(const __grid_constant__ CUtensorMap map, int c0, int x0, int y0, int img,
                                                            int bytes, uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&bar), (uint32_t)bytes);
    tma_load_4d(smem_u32(smem), &map, smem_u32(&bar), c0, x0, y0, img);
  }
  mbar_wait(smem_u32(&bar), 0, nullptr, 7);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

// D[128][n] = A[128][64*kc] * B[n][64*kc]^T ; A and B row-major bf16 in global memory.
__global__ void __launch_bounds__(128) diag_umma_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                                                        float* d, int n, int kc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                 // 128 rows x 128 B
  uint8_t* sb = smem + 16384;         // n rows x 128 B (<= 32 KB)
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&full), 1);
    mbar_init(smem_u32(&done), 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  for (int k = 0; k < kc; ++k) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(smem_u32(&full), (uint32_t)(128 + n) * 128);
      tma_load_2d(smem_u32(sa), &amap, smem_u32(&full), k * 64, 0);
      tma_load_2d(smem_u32(sa) + 64 * 128, &amap, smem_u32(&full), k * 64, 64);
      for (int j = 0; j < n / 64; ++j) tma_load_2d(smem_u32(sb) + j * 64 * 128, &bmap, smem_u32(&full), k * 64, j * 64);
      mbar_wait(smem_u32(&full), k & 1, nullptr, 8);
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, (uint32_t)n);
      for (int kk = 0; kk < 4; ++kk)
        umma_bf16(tmem, make_sw128_desc(smem_u32(sa) + kk * 32), make_sw128_desc(smem_u32(sb) + kk * 32), idesc, (k | kk) ? 1u : 0u);
      umma_commit(smem_u32(&done));
      mbar_wait(smem_u32(&done), k & 1, nullptr, 9);
    }
    __syncthreads();
  }
  tc_fence_after();
  const int m = warp * 32 + lane;
  for (int j = 0; j < n / 16; ++j) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + j * 16, v);
    for (int i = 0; i < 16; ++i) d[(size_t)m * n + j * 16 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// Experiment: A is a [rows_a x 64] bf16 matrix (rows_a >= 128 + shift) loaded as ONE swizzled TMA box; the MMA reads
// 128 rows starting at row `shift` with 8-row-group stride `sbo_rows` rows.  D[m][n] should equal
// sum_k A[shift + (m/8)*sbo_rows + m%8][k] * B[n][k].  base_mode: 0 -> base_offset field 0; 1 -> (start>>7)&7.
__global__ void __launch_bounds__(128) diag_umma_shift_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                                                              float* d, int n, int rows_a, int shift, int sbo_rows, int base_mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                       // rows_a x 128 B (<= 64 KB)
  uint8_t* sb = smem + 65536;               // n rows x 128 B
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&full), 1);
    mbar_init(smem_u32(&done), 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&full), (uint32_t)(rows_a + n) * 128);
    for (int r = 0; r < rows_a; r += 64) tma_load_2d(smem_u32(sa) + r * 128, &amap, smem_u32(&full), 0, r);
    for (int j = 0; j < n / 64; ++j) tma_load_2d(smem_u32(sb) + j * 64 * 128, &bmap, smem_u32(&full), 0, j * 64);
    mbar_wait(smem_u32(&full), 0, nullptr, 8);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)n);
    const uint32_t a_start = smem_u32(sa) + shift * 128;
    const uint64_t base_off = base_mode ? (uint64_t)((a_start >> 7) & 7) : 0ull;
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t adesc = (uint64_t)(((a_start + kk * 32) & 0x3FFFFu) >> 4) | ((uint64_t)((sbo_rows * 128) >> 4) << 32) | (1ull << 46) |
                             (base_off << 49) | (2ull << 61);
      umma_bf16(tmem, adesc, make_sw128_desc(smem_u32(sb) + kk * 32), idesc, kk ? 1u : 0u);
    }
    umma_commit(smem_u32(&done));
    mbar_wait(smem_u32(&done), 0, nullptr, 9);
  }
  __syncthreads();
  tc_fence_after();
  const int m = warp * 32 + lane;
  for (int j = 0; j < n / 16; ++j) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + j * 16, v);
    for (int i = 0; i < 16; ++i) d[(size_t)m * n + j * 16 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// Experiment / bring-up for back-to-back GEMMs in an epilogue: the A operand of tcgen05.mma read from TENSOR MEMORY.
// Each thread (= TMEM lane = row m) packs its row of A [128 x 64] bf16 two elements per 32-bit column (element 2c in the low
// half of column c) and writes the 32 columns with tcgen05.st at column a_col; B [n x 64] comes from shared memory (TMA,
// SWIZZLE_128B).  D[m][j] = sum_k A[m][k] B[j][k] lands at column 0.  K step kk (16 elements) reads A columns a_col + 8 kk.
__global__ void __launch_bounds__(128) diag_umma_ts_kernel(const __nv_bfloat16* a, const __grid_constant__ CUtensorMap bmap, float* d, int n, int a_col) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sb = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&full), 1);
    mbar_init(smem_u32(&done), 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int m = warp * 32 + lane;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(a + (size_t)m * 64);     // 32 packed pairs
    float lo[16], hi[16];
    for (int i = 0; i < 16; ++i) { lo[i] = __uint_as_float(row[i]); hi[i] = __uint_as_float(row[16 + i]); }
    tmem_st16(lane_addr + a_col, lo);
    tmem_st16(lane_addr + a_col + 16, hi);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    mbar_expect_tx(smem_u32(&full), (uint32_t)n * 128);
    for (int j = 0; j < n / 64; ++j) tma_load_2d(smem_u32(sb) + j * 64 * 128, &bmap, smem_u32(&full), 0, j * 64);
    mbar_wait(smem_u32(&full), 0, nullptr, 8);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)n);
    for (int kk = 0; kk < 4; ++kk) umma_bf16_ts(tmem, tmem + a_col + 8 * kk, make_sw128_desc(smem_u32(sb) + kk * 32), idesc, kk ? 1u : 0u);
    umma_commit(smem_u32(&done));
    mbar_wait(smem_u32(&done), 0, nullptr, 9);
  }
  __syncthreads();
  tc_fence_after();
  for (int j = 0; j < n / 16; ++j) {
    float v[16];
    tmem_ld16(lane_addr + j * 16, v);
    for (int i = 0; i < 16; ++i) d[(size_t)m * n + j * 16 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}
}  // namespace

extern "C" {

int sf_diag_umma_shift(const void* a_bf16, const void* b_bf16, float* d, int n, int rows_a, int shift, int sbo_rows, int base_mode,
                       void* stream) {
  EncodeTiledFn enc = encode_fn();
  if (!enc || n % 64 || n > 256 || n <= 0 || rows_a % 64 || rows_a > 512) return SF_ERR_INVALID;
  CUtensorMap am, bm;
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  cuuint64_t stride[1] = {128};
  cuuint64_t adims[2] = {64, (cuuint64_t)rows_a};
  cuuint64_t bdims[2] = {64, (cuuint64_t)n};
  if (enc(&am, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a_bf16), adims, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return SF_ERR_CUDA;
  if (enc(&bm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(b_bf16), bdims, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return SF_ERR_CUDA;
  const int smem = 65536 + 32768 + 1024;
  cudaFuncSetAttribute(reinterpret_cast<const void*>(diag_umma_shift_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  diag_umma_shift_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(am, bm, d, n, rows_a, shift, sbo_rows, base_mode);
  return cudaGetLastError() == cudaSuccess ? SF_OK : SF_ERR_CUDA;
}

int sf_diag_umma_ts(const void* a_bf16, const void* b_bf16, float* d, int n, int a_col, void* stream) {
  EncodeTiledFn enc = encode_fn();
  if (!enc || n % 64 || n > 256 || n <= 0 || a_col < n || a_col + 32 > 512) return SF_ERR_INVALID;
  CUtensorMap bm;
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  cuuint64_t stride[1] = {128};
  cuuint64_t bdims[2] = {64, (cuuint64_t)n};
  if (enc(&bm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(b_bf16), bdims, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return SF_ERR_CUDA;
  const int smem = 32768 + 1024;
  cudaFuncSetAttribute(reinterpret_cast<const void*>(diag_umma_ts_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  diag_umma_ts_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(a_bf16), bm, d, n, a_col);
  return cudaGetLastError() == cudaSuccess ? SF_OK : SF_ERR_CUDA;
}

int sf_diag_tma_dump(const void* act_bf16, int n_images, int H, int W, int C, int img, int y0, int x0, int c0, int rows,
                     void* out_smem_copy, void* stream) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return SF_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_images};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, 8, (cuuint32_t)rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(act_bf16), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return SF_ERR_CUDA;
  const int bytes = rows * 8 * 128;
  cudaFuncSetAttribute(reinterpret_cast<const void*>(diag_tma_dump_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 1024);
  diag_tma_dump_kernel<<<1, 128, bytes + 1024, reinterpret_cast<cudaStream_t>(stream)>>>(m, c0, x0, y0, img, bytes,
                                                                                       reinterpret_cast<uint8_t*>(out_smem_copy));
  return cudaGetLastError() == cudaSuccess ? SF_OK : SF_ERR_CUDA;
}

int sf_diag_umma(const void* a_bf16, const void* b_bf16, float* d, int n, int k_chunks, void* stream) {
  EncodeTiledFn enc = encode_fn();
  if (!enc || n % 64 || n > 256 || n <= 0) return SF_ERR_INVALID;
  CUtensorMap am, bm;
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  cuuint64_t stride[1] = {(cuuint64_t)k_chunks * 64 * 2};
  cuuint64_t adims[2] = {(cuuint64_t)k_chunks * 64, 128};
  cuuint64_t bdims[2] = {(cuuint64_t)k_chunks * 64, (cuuint64_t)n};
  if (enc(&am, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a_bf16), adims, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return SF_ERR_CUDA;
  if (enc(&bm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(b_bf16), bdims, stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return SF_ERR_CUDA;
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(reinterpret_cast<const void*>(diag_umma_kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  diag_umma_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(am, bm, d, n, k_chunks);
  return cudaGetLastError() == cudaSuccess ? SF_OK : SF_ERR_CUDA;
}

}  // extern "C"
