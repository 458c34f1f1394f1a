// umma_rate.cu -- microbenchmark: cycles per tcgen05.mma (bf16, M=128 per CTA, K=16) as a function of N, operand source
// (A from shared memory vs A from TMEM) and CTA pairing (cta_group::1 vs ::2).  Answers one design question of the conv
// stage kernel: is the tensor pipe fed at full rate from shared memory for N = 64 / 128 / 256?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate scripts/umma_rate.cu && ./umma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 2000000000LL) return false;
  return true;
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t addr, uint32_t sbo = 1024u) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_bf16(uint32_t M, uint32_t N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24); }

template <int CTAS>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CTAS == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CTAS == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void commit(uint32_t bar) {
  if (CTAS == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// mode 0: A and B from shared memory; mode 1: A from TMEM.  shiftA: distinct A start per MMA (like the conv's tap windows).
template <int CTAS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int mode, int nacc, int iters, long long* out, int* status) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  // A: 64 KB region (room for shifted starts); B: up to 256 rows x 128 B = 32 KB
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + 64 * 1024;
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + (i & 7);   // small bf16 values
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const bool leader = (CTAS == 1) || cluster_rank() == 0;
  bool ok = true;
  if (warp == 0 && leader) {
    // the whole warp runs the loop converged and one elected lane issues (operands stay in uniform registers)
    const uint32_t idesc = idesc_bf16(128 * CTAS, (uint32_t)N);
    const uint64_t bdesc = sw128_desc(smem_u32(b_s));
    const uint32_t a0 = smem_u32(a_s);
    const uint32_t a_tmem = tmem + 504;     // columns [504, 512): a 128 x 16 bf16 A operand (the accumulators never reach it in mode 1)
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
      // 16 MMAs per iteration: 4 "taps" (A start shifted by whole 128-byte rows) x 4 K-steps of one 64-channel chunk
#pragma unroll
      for (int t = 0; t < 4; ++t) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // consecutive MMAs rotate over nacc independent accumulators (dependent MMAs on one accumulator serialise)
          const uint32_t d = tmem + (uint32_t)(((t * 4 + k) & (nacc - 1)) * N);
          if (mode == 2)        // the conv kernel's A operand: 8-row groups 18 pixel rows apart (halo box of a 16-wide tile), tap-shifted start
            mma_ss<CTAS>(d, sw128_desc(a0 + (t * 18 + t + 1) * 128, 18 * 128) + 2 * k, bdesc + (uint64_t)(t * 512) + 2 * k, idesc, it ? 1u : 0u);
          else if (mode == 0) mma_ss<CTAS>(d, sw128_desc(a0 + t * 1024 * 3) + 2 * k, bdesc + 2 * k, idesc, it ? 1u : 0u);
          else mma_ts<CTAS>(d, a_tmem, bdesc + 2 * k, idesc, it ? 1u : 0u);
        }
      }
      }
      __syncwarp();
    }
    if (elect_one()) commit<CTAS>(smem_u32(&bar));
    __syncwarp();
    ok = mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (!ok) atomicExch(status, 1);
  } else if (CTAS == 2 && threadIdx.x == 0) {
    ok = mbar_wait(smem_u32(&bar), 0);      // the multicast commit also arrives here
    if (!ok) atomicExch(status, 2);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int CTAS>
void run(int N, int mode, int nacc, int grid, int iters, long long* d_out, int* d_status) {
  const size_t smem = 97 * 1024 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaMemset(d_status, 0, sizeof(int)));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, rate_kernel<CTAS>, N, mode, nacc, iters, d_out, d_status));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long h[148 * 2] = {0}; int st = 0;
  CK(cudaMemcpy(h, d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&st, d_status, sizeof(int), cudaMemcpyDeviceToHost));
  const double mmas = 16.0 * iters;
  const double cyc = (double)h[0] / mmas;
  const double flops = 2.0 * 128 * CTAS * N * 16 * mmas * (grid / CTAS);
  printf("cta_group::%d  %s  N=%3d nacc=%d grid=%3d : %7.1f cycles/MMA (ideal %5.1f) -> tensor %5.1f %%   chip %7.1f TFLOP/s  [%s]\n", CTAS,
         mode == 1 ? "A=tmem" : mode == 2 ? "A=smem conv-window" : "A=smem", N, nacc, grid, cyc, N / 2.0, 100.0 * (N / 2.0) / cyc, flops / (ms * 1e-3) / 1e12, st ? "TIMEOUT" : "ok");
}

int main() {
  long long* d_out; int* d_status;
  CK(cudaMalloc(&d_out, sizeof(long long) * 512));
  CK(cudaMalloc(&d_status, sizeof(int)));
  const int iters = 4000;
  for (int grid : {1, 148})
    for (int mode = 0; mode < 2; ++mode)
      for (int N : {64, 128, 256})
        for (int nacc = 1; nacc * N <= (mode ? 256 : 512) && nacc <= 8; nacc *= 2) {
          if (grid == 148 && nacc * 2 * N <= (mode ? 256 : 512) && nacc < 8) continue;     // full grid: widest rotation only
          run<1>(N, mode, nacc, grid, iters, d_out, d_status);
        }
  for (int N : {64, 128, 256}) run<1>(N, 2, 2, 148, iters, d_out, d_status);
  for (int grid : {2, 148})
    for (int mode = 0; mode < 2; ++mode)
      for (int N : {64, 128, 256})
        for (int nacc = 1; nacc * N <= (mode ? 256 : 512) && nacc <= 8; nacc *= 2) {
          if (nacc * 2 * N <= (mode ? 256 : 512) && nacc < 8) continue;
          run<2>(N, mode, nacc, grid, iters, d_out, d_status);
        }
  return 0;
}
