// sf_elementwise.cuh -- the HBM-bound kernels of the path: squeeze-excite reduce / apply
// (res_models.py:150-165) and the NCHW fp32 <-> NHWC bf16 layout kernels. 128-bit accesses throughout.
#pragma once
#include <curand_kernel.h>

#include "sf_ptx.cuh"

namespace sf {

// scale[c] = sigmoid(fc2 . relu(fc1 . mean)), mean[c] = inv_n * sum_k partial[k][c]; one 256-thread block, fixed order.
template <int CH>
__device__ __forceinline__ void se_scale_from_partials(const float* partials, int n_partials, float inv_n, const float* fc1,
                                                       const float* fc2, float* scale_out, float* scratch) {
  constexpr int HID = CH / 8;
  float* mean_s = scratch;            // CH
  float* hid_s = scratch + CH;        // HID
  __syncthreads();
  if (threadIdx.x < CH) {
    float t = 0.0f;
    for (int k = 0; k < n_partials; ++k) t += partials[(size_t)k * CH + threadIdx.x];
    mean_s[threadIdx.x] = t * inv_n;
  }
  __syncthreads();
  {   // FC1: HID outputs, 256 / HID = 16 threads each (CH / 16 = 8 inputs per thread), shuffle-reduced
    constexpr int TPO = 256 / HID;
    const int o = threadIdx.x / TPO, part = threadIdx.x % TPO;
    float a = 0.0f;
    for (int c = part; c < CH; c += TPO) a = fmaf(fc1[o * CH + c], mean_s[c], a);
#pragma unroll
    for (int d = TPO / 2; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (part == 0) hid_s[o] = fmaxf(a, 0.0f);
  }
  __syncthreads();
  if (threadIdx.x < CH) {
    float a = 0.0f;
#pragma unroll
    for (int j = 0; j < HID; ++j) a = fmaf(fc2[threadIdx.x * HID + j], hid_s[j], a);
    scale_out[threadIdx.x] = __fdividef(1.0f, 1.0f + __expf(-a));
  }
}

// tiny kernel for the row-sharded path: scales from already all-reduced sums (one "partial" per sample)
template <int CH>
__global__ void __launch_bounds__(256) se_scale_kernel(const float* __restrict__ sums, int n_partials, float inv_n,
                                                       const float* __restrict__ fc1, const float* __restrict__ fc2,
                                                       float* __restrict__ scale_out) {
  __shared__ float scratch[CH + CH / 8];
  se_scale_from_partials<CH>(sums + (size_t)blockIdx.x * n_partials * CH, n_partials, inv_n, fc1, fc2, scale_out + (size_t)blockIdx.x * CH,
                             scratch);
}

// ---- SE step 1: per-(sample, channel) sums over the H*W pixels of a [img][H*W][CH] bf16 tensor --------
// grid = (blocks_per_image, n_active); block = 256 threads = 8 channel groups (16 ch = 32 B, one 256-bit load) x 32 pixel
// lanes, 4 independent loads in flight per thread.  Optional row window [row0, row1) (row sharding: own rows only).
template <int CH, bool X3>
__global__ void __launch_bounds__(256) se_reduce_kernel(const __nv_bfloat16* __restrict__ zh, const __nv_bfloat16* __restrict__ zl,
                                                        float* __restrict__ sums, const int* __restrict__ sample_id, int hw, int px0, int px1,
                                                        unsigned int* __restrict__ counters, float* __restrict__ scale_out,
                                                        const float* __restrict__ fc1, const float* __restrict__ fc2, float inv_n,
                                                        float* __restrict__ totals) {
  static_assert(CH == 128 || CH == 256, "SE layers of the prior network have 2C = 128 or 256 channels");
  constexpr int GROUPS = CH / 16;           // 8
  constexpr int LANES = 256 / GROUPS;       // 32 pixels per block iteration
  constexpr int UNROLL = 4;
  const int bi = blockIdx.y;
  const int sid = sample_id[bi];
  const int g = threadIdx.x % GROUPS, pl = threadIdx.x / GROUPS;
  const size_t base = (size_t)sid * hw * CH;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
  for (int px = px0 + blockIdx.x * LANES * UNROLL + pl; px < px1; px += gridDim.x * LANES * UNROLL) {
    uint32_t v[UNROLL][8];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int q = px + u * LANES;
      if (q < px1) ldg256(zh + base + (size_t)q * CH + g * 16, v[u]);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[u][i] = 0u;
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc[2 * i] += bf16_lo_f(v[u][i]); acc[2 * i + 1] += bf16_hi_f(v[u][i]); }
    if (X3) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int q = px + u * LANES;
        if (q < px1) {
          uint32_t w[8];
          ldg256(zl + base + (size_t)q * CH + g * 16, w);
#pragma unroll
          for (int i = 0; i < 8; ++i) { acc[2 * i] += bf16_lo_f(w[i]); acc[2 * i + 1] += bf16_hi_f(w[i]); }
        }
      }
    }
  }
  __shared__ float red[LANES][CH + 1];
#pragma unroll
  for (int i = 0; i < 16; ++i) red[pl][g * 16 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < CH) {
    float s = 0.0f;
#pragma unroll
    for (int l = 0; l < LANES; ++l) s += red[l][threadIdx.x];
    // per-block partial sums, combined in a fixed order below: deterministic (no float atomics)
    sums[((size_t)bi * gridDim.x + blockIdx.x) * CH + threadIdx.x] = s;
  }
  if (scale_out == nullptr && totals == nullptr) return;      // caller combines the partials itself
  // last block of this sample: mean -> FC(2C -> 2C/8) -> ReLU -> FC -> sigmoid, written once for se_apply to stream with
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(counters + bi, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (totals != nullptr) {
    // row sharding: this rank's band total per (sample, channel), partials added in a fixed order; the caller all-reduces the
    // [n_active][CH] totals across the ranks and then runs the scale + fold / apply step (sf_plan_se_finish)
    if (threadIdx.x < CH) {
      float t = 0.0f;
      for (int k = 0; k < (int)gridDim.x; ++k) t += sums[((size_t)bi * gridDim.x + k) * CH + threadIdx.x];
      totals[(size_t)bi * CH + threadIdx.x] = t;
    }
    if (threadIdx.x == 0) counters[bi] = 0;
    return;
  }
  se_scale_from_partials<CH>(sums + (size_t)bi * gridDim.x * CH, gridDim.x, inv_n, fc1, fc2, scale_out + (size_t)bi * CH, &red[0][0]);
  if (threadIdx.x == 0) counters[bi] = 0;      // ready for the next launch
}

// ---- SE step 2: y = z * scale (scale computed by se_reduce's last block / se_scale_kernel) ------------------
// pure streaming: 256-bit loads / stores, 4 in flight per thread.
template <int CH, bool X3>
__global__ void __launch_bounds__(256) se_apply_kernel(const __nv_bfloat16* __restrict__ zh, const __nv_bfloat16* __restrict__ zl,
                                                       __nv_bfloat16* __restrict__ yh, __nv_bfloat16* __restrict__ yl,
                                                       const float* __restrict__ scale, const int* __restrict__ sample_id, int hw) {
  constexpr int GROUPS = CH / 16;
  constexpr int LANES = 256 / GROUPS;
  constexpr int UNROLL = 4;
  const int bi = blockIdx.y;
  const int sid = sample_id[bi];
  const float* scale_s = scale + (size_t)bi * CH;
  const int g = threadIdx.x % GROUPS, pl = threadIdx.x / GROUPS;
  const size_t base = (size_t)sid * hw * CH;
  float sc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) sc[i] = scale_s[g * 16 + i];
  for (int px = blockIdx.x * LANES * UNROLL + pl; px < hw; px += gridDim.x * LANES * UNROLL) {
    uint32_t v[UNROLL][8];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int q = px + u * LANES;
      if (q < hw) ldg256(zh + base + (size_t)q * CH + g * 16, v[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int q = px + u * LANES;
      if (q >= hw) continue;
      const size_t off = base + (size_t)q * CH + g * 16;
      float f[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) { f[2 * i] = bf16_lo_f(v[u][i]); f[2 * i + 1] = bf16_hi_f(v[u][i]); }
      if (X3) {
        uint32_t w[8];
        ldg256(zl + off, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) { f[2 * i] += bf16_lo_f(w[i]); f[2 * i + 1] += bf16_hi_f(w[i]); }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] *= sc[i];
      uint32_t h[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
      stg256(yh + off, h);
      if (X3) {
        uint32_t l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) l[i] = pack_bf16x2(f[2 * i] - bf16_lo_f(h[i]), f[2 * i + 1] - bf16_hi_f(h[i]));
        stg256(yl + off, l);
      }
    }
  }
}

// ---- SE layer folded into a consumer's weights: w_scaled[sample][row][k] = bf16(w32[row][k] * scale[sample][c0(row) + k]) ---
// (residual rows of the split mode get the rounding residual of that product).  One thread = 8 input channels of one row.
__global__ void __launch_bounds__(256) se_fold_kernel(const float* __restrict__ w32, const int32_t* __restrict__ row_meta,
                                                      const float* __restrict__ scale, __nv_bfloat16* __restrict__ out, int rows, int CH) {
  const int bi = blockIdx.y;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int row = idx >> 3, g = idx & 7;
  if (row >= rows) return;
  const int meta = row_meta[row];
  const int c0 = meta & 0xffff;
  const bool lo = (meta >> 16) & 1;
  const float4* wp = reinterpret_cast<const float4*>(w32 + (size_t)row * 64 + g * 8);
  const float4* sp = reinterpret_cast<const float4*>(scale + (size_t)bi * CH + c0 + g * 8);
  const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1), s0 = sp[0], s1 = sp[1];
  const float v[8] = {w0.x * s0.x, w0.y * s0.y, w0.z * s0.z, w0.w * s0.w, w1.x * s1.x, w1.y * s1.y, w1.z * s1.z, w1.w * s1.w};
  uint32_t h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    if (lo) h[i] = pack_bf16x2(v[2 * i] - bf16_lo_f(h[i]), v[2 * i + 1] - bf16_hi_f(h[i]));
  }
  *reinterpret_cast<uint4*>(out + ((size_t)bi * rows + row) * 64 + g * 8) = make_uint4(h[0], h[1], h[2], h[3]);
}

// ---- standard-normal noise for every rsample call of a rollout, one launch --------------------------------------------
// Slot s (blockIdx.y) is filled exactly as the s-th of a sequence of torch normal_() calls on a numel-element CUDA float tensor
// would fill it: same thread -> element mapping (element idx + T * (4 k + i) comes from component i of thread idx's k-th
// curand_normal4 draw, T = gridDim.x * 256), same Philox4_32_10 subsequence (= thread index) and offset (offset0 + s * per_slot).
// curand_init(seed, idx, offset) followed by k curand_normal4 draws evaluates Philox at counter (offset / 4 + k) of subsequence
// idx (offset is a multiple of 4 here: ATen advances it by 4 per loop iteration); curand itself evaluates two extra blocks per
// thread (one in each skipahead, one look-ahead per draw), so the counters are formed directly.
// slot_list (optional): only the listed slots are filled -- a rollout that skips dead prior-net evaluations never reads the others
__global__ void __launch_bounds__(256) normal_slots_kernel(float* __restrict__ out, long long numel, unsigned long long seed,
                                                           unsigned long long offset0, unsigned int per_slot, int n_slots,
                                                           const int* __restrict__ slot_list) {
  const unsigned int idx = blockIdx.x * 256u + threadIdx.x;
  const uint2 key = make_uint2((unsigned int)seed, (unsigned int)(seed >> 32));
  const long long T = (long long)gridDim.x * 256;
  const long long rounded = ((numel - 1) / (T * 4) + 1) * T * 4;
  for (int si = blockIdx.y; si < n_slots; si += gridDim.y) {      // a block walks several slots: fewer, longer-lived blocks
    const int slot = slot_list ? slot_list[si] : si;
    unsigned long long ctr = (offset0 + (unsigned long long)slot * per_slot) >> 2;
    float* o = out + (long long)slot * numel;
    for (long long li = idx; li < rounded && li < numel; li += T * 4, ++ctr) {      // li >= numel: none of the 4 outputs is stored
      const uint4 x = curand_Philox4x32_10(make_uint4((unsigned int)ctr, (unsigned int)(ctr >> 32), idx, 0u), key);
      const float2 a = _curand_box_muller(x.x, x.y), b = _curand_box_muller(x.z, x.w);      // = curand_normal4
      const float v[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long l = li + T * i;
        if (l < numel) o[l] = v[i];
      }
    }
  }
}

// ---- NCHW fp32 -> NHWC bf16 (hi [+ lo]) : encoded observations entering the ODE loop ---------------------
// block = 256 threads handles 64 consecutive pixels x 64 channels of one image through a padded smem tile: 16 scalar loads
// in flight per thread (each warp-load = 128 contiguous bytes of one channel row), two 16-byte NHWC stores per thread.
constexpr int PACK_PX = 64;
template <bool X3>
__global__ void __launch_bounds__(256) pack_nchw_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dh,
                                                        __nv_bfloat16* __restrict__ dl, int C, int hw) {
  __shared__ float tile[64][PACK_PX + 1];
  const int img = blockIdx.z, cb = blockIdx.y * 64, p0 = blockIdx.x * PACK_PX;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float v[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float* row = src + ((size_t)img * C + cb + w + 8 * k) * hw + p0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int px = lane + 32 * h;
      v[k][h] = (p0 + px < hw) ? __ldg(row + px) : 0.0f;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    tile[w + 8 * k][lane] = v[k][0];
    tile[w + 8 * k][lane + 32] = v[k][1];
  }
  __syncthreads();
  // 64 pixels x 8 groups of 8 channels = 512 work items, two per thread
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int item = threadIdx.x + 256 * it;
    const int px = item >> 3, g = item & 7;
    if (p0 + px < hw) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = tile[g * 8 + i][px];
      uint32_t h[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
      const size_t off = ((size_t)img * hw + p0 + px) * C + cb + g * 8;
      *reinterpret_cast<uint4*>(dh + off) = make_uint4(h[0], h[1], h[2], h[3]);
      if (X3) {
        uint32_t l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) l[i] = pack_bf16x2(f[2 * i] - bf16_lo_f(h[i]), f[2 * i + 1] - bf16_hi_f(h[i]));
        *reinterpret_cast<uint4*>(dl + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
}

// ---- 2x2 max-pool and nearest 2x up-sampling on NHWC bf16 (hi [+ lo]) activations (SmallEncoder / SmallDecoder) -----------
// one thread = 8 channels (16 B) of one output pixel
template <bool X3>
__global__ void __launch_bounds__(256) maxpool2_kernel(const __nv_bfloat16* __restrict__ sh, const __nv_bfloat16* __restrict__ sl,
                                                       __nv_bfloat16* __restrict__ dh, __nv_bfloat16* __restrict__ dl, int n_img, int H,
                                                       int W, int C) {
  const int Ho = H / 2, Wo = W / 2, G = C / 8;
  const size_t total = (size_t)n_img * Ho * Wo * G;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    size_t r = idx / G;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const int img = (int)(r / Ho);
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const size_t off = (((size_t)img * H + 2 * yo + dy) * W + 2 * xo + dx) * C + g * 8;
        const uint4 v = *reinterpret_cast<const uint4*>(sh + off);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) { f[2 * i] = bf16_lo_f(w[i]); f[2 * i + 1] = bf16_hi_f(w[i]); }
        if (X3) {
          const uint4 u = *reinterpret_cast<const uint4*>(sl + off);
          const uint32_t x[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) { f[2 * i] += bf16_lo_f(x[i]); f[2 * i + 1] += bf16_hi_f(x[i]); }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], f[i]);
      }
    const size_t o = (((size_t)img * Ho + yo) * Wo + xo) * C + g * 8;
    uint32_t h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = pack_bf16x2(m[2 * i], m[2 * i + 1]);
    *reinterpret_cast<uint4*>(dh + o) = make_uint4(h[0], h[1], h[2], h[3]);
    if (X3) {
      uint32_t l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) l[i] = pack_bf16x2(m[2 * i] - bf16_lo_f(h[i]), m[2 * i + 1] - bf16_hi_f(h[i]));
      *reinterpret_cast<uint4*>(dl + o) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

// nearest-neighbour x2: pure copy of 16-byte channel groups (works per plane)
__global__ void __launch_bounds__(256) upsample2_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n_img, int H, int W, int G) {
  const int Ho = 2 * H, Wo = 2 * W;
  const size_t total = (size_t)n_img * Ho * Wo * G;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    size_t r = idx / G;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const int img = (int)(r / Ho);
    dst[idx] = src[(((size_t)img * H + yo / 2) * W + xo / 2) * G + g];
  }
}

// fp32 NHWC (recorded path states, gathered by slot) -> bf16 NHWC hi [+ lo]: the decoder's input
template <bool X3>
__global__ void __launch_bounds__(256) cast_nhwc_kernel(const float* __restrict__ src, const int* __restrict__ slots, __nv_bfloat16* __restrict__ dh,
                                                        __nv_bfloat16* __restrict__ dl, int n_out, size_t per_img) {   // per_img = H*W*C, multiple of 8
  const size_t groups = per_img / 8, total = (size_t)n_out * groups;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(idx / groups);
    const size_t g = idx % groups;
    const int slot = slots ? slots[o] : o;
    const float4* q = reinterpret_cast<const float4*>(src + (size_t)slot * per_img + g * 8);
    const float4 a = q[0], b = q[1];
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
    *reinterpret_cast<uint4*>(dh + (size_t)o * per_img + g * 8) = make_uint4(h[0], h[1], h[2], h[3]);
    if (X3) {
      uint32_t l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) l[i] = pack_bf16x2(f[2 * i] - bf16_lo_f(h[i]), f[2 * i + 1] - bf16_hi_f(h[i]));
      *reinterpret_cast<uint4*>(dl + (size_t)o * per_img + g * 8) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
}

// 8 channels (16 bytes per plane) of one NHWC pixel <-> fp32 registers; the residual plane is added / produced in the split mode
__device__ __forceinline__ void load8(const __nv_bfloat16* h, const __nv_bfloat16* l, size_t off, bool x3, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(h + off);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = bf16_lo_f(w[i]); f[2 * i + 1] = bf16_hi_f(w[i]); }
  if (x3) {
    const uint4 u = *reinterpret_cast<const uint4*>(l + off);
    const uint32_t x[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] += bf16_lo_f(x[i]); f[2 * i + 1] += bf16_hi_f(x[i]); }
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* h, __nv_bfloat16* l, size_t off, bool x3, const float (&f)[8]) {
  uint32_t a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(h + off) = make_uint4(a[0], a[1], a[2], a[3]);
  if (x3) {
    uint32_t b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = pack_bf16x2(f[2 * i] - bf16_lo_f(a[i]), f[2 * i + 1] - bf16_hi_f(a[i]));
    *reinterpret_cast<uint4*>(l + off) = make_uint4(b[0], b[1], b[2], b[3]);
  }
}


// ---- ConvNeXt front end: depthwise 7x7 conv + bias, then LayerNorm over the 64 channels (channels_last, eps 1e-6) --------
// Round 2 rewrite (the first version re-read every input from L1 per tap: 49 LDG + 98 LDS per 392 FMAs, 2.8 ms for 56 frames of
// 200 x 200 = 7 % of the fp32 rate; this one is 3-6x faster).  Block = 256 threads = 32 output columns x 8 channel groups of 8;
// every thread produces a strip of 16 output rows of its column.  The (16+6) x (32+6) pixel halo tile sits in shared memory as
// fp32 (hi + lo already summed; out-of-image pixels are zeros = the conv's padding), laid out per pixel as
// [half 0/1][group 0..7][4 floats] so that a quarter-warp's LDS.128 covers 128 contiguous bytes.  Per filter column kx a
// thread keeps the 7 x 8 weights of its channel group in registers and walks the 22 input rows once: each loaded value
// feeds up to 7 accumulator rows (128 accumulators / thread), i.e. ~400 shared-memory loads for 6272 FMAs (FMA-issue bound;
// 8-row strips were shared-memory-load bound: 0.90 ms vs the first version's 2.78 ms for 56 frames of 200 x 200).
constexpr int DW_TILE_W = 32, DW_TILE_H = 16, DW_TW = DW_TILE_W + 6, DW_TH = DW_TILE_H + 6;
constexpr int DW_SMEM_BYTES = (DW_TH * DW_TW * 64 + 49 * 64 + 3 * 64) * 4;
template <bool X3>
__global__ void __launch_bounds__(256, 1) dwconv7_ln_kernel(const __nv_bfloat16* __restrict__ sh, const __nv_bfloat16* __restrict__ sl,
                                                            __nv_bfloat16* __restrict__ dh, __nv_bfloat16* __restrict__ dl,
                                                            const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                            const float* __restrict__ ln_w, const float* __restrict__ ln_b, int H, int W) {
  extern __shared__ float dw_smem[];
  float* tile = dw_smem;                         // [DW_TH * DW_TW pixels][2][8][4]
  float* wsm = tile + DW_TH * DW_TW * 64;        // [tap][channel]
  float* vsm = wsm + 49 * 64;                    // conv bias | LN weight | LN bias
  for (int i = threadIdx.x; i < 49 * 64; i += 256) { const int c = i / 49, t = i % 49; wsm[t * 64 + c] = dw_w[c * 49 + t]; }   // weight [64][1][7][7]
  if (threadIdx.x < 64) {
    vsm[threadIdx.x] = dw_b[threadIdx.x];
    vsm[64 + threadIdx.x] = ln_w[threadIdx.x];
    vsm[128 + threadIdx.x] = ln_b[threadIdx.x];
  }
  const int x0 = blockIdx.x * DW_TILE_W - 3, y0 = blockIdx.y * DW_TILE_H - 3, img = blockIdx.z;
  // tile load: all 17 (x2 in the split mode) 16-byte loads of a thread are issued before the first one is consumed -- with one
  // block per SM nothing else hides their latency
  constexpr int DW_ITEMS = DW_TH * DW_TW * 8, DW_NIT = (DW_ITEMS + 255) / 256;
  uint4 vh[DW_NIT], vl[X3 ? DW_NIT : 1];
#pragma unroll
  for (int k = 0; k < DW_NIT; ++k) {
    const int i = threadIdx.x + 256 * k;
    const int gq = i & 7, pix = i >> 3;
    const int ty = pix / DW_TW, tx = pix - ty * DW_TW;
    const int yy = y0 + ty, xx = x0 + tx;
    const bool ok = i < DW_ITEMS && yy >= 0 && yy < H && xx >= 0 && xx < W;
    const size_t off = ok ? (((size_t)img * H + yy) * W + xx) * 64 + gq * 8 : 0;
    vh[k] = ok ? __ldg(reinterpret_cast<const uint4*>(sh + off)) : make_uint4(0u, 0u, 0u, 0u);
    if (X3) vl[k] = ok ? __ldg(reinterpret_cast<const uint4*>(sl + off)) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int k = 0; k < DW_NIT; ++k) {
    const int i = threadIdx.x + 256 * k;
    if (i < DW_ITEMS) {
      const int gq = i & 7, pix = i >> 3;
      const uint32_t a[4] = {vh[k].x, vh[k].y, vh[k].z, vh[k].w};
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) { f[2 * j] = bf16_lo_f(a[j]); f[2 * j + 1] = bf16_hi_f(a[j]); }
      if (X3) {
        const uint32_t b[4] = {vl[k].x, vl[k].y, vl[k].z, vl[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { f[2 * j] += bf16_lo_f(b[j]); f[2 * j + 1] += bf16_hi_f(b[j]); }
      }
      *reinterpret_cast<float4*>(tile + (pix * 2 + 0) * 32 + gq * 4) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(tile + (pix * 2 + 1) * 32 + gq * 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
  __syncthreads();
  const int g = threadIdx.x & 7, col = threadIdx.x >> 3;          // channel group, output column of the tile
  float acc[DW_TILE_H][8];
#pragma unroll
  for (int r = 0; r < DW_TILE_H; ++r)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[r][j] = vsm[g * 8 + j];
#pragma unroll 1
  for (int kx = 0; kx < 7; ++kx) {
    float w[7][8];
#pragma unroll
    for (int ky = 0; ky < 7; ++ky) {
      const float4 a = *reinterpret_cast<const float4*>(wsm + (ky * 7 + kx) * 64 + g * 8);
      const float4 b = *reinterpret_cast<const float4*>(wsm + (ky * 7 + kx) * 64 + g * 8 + 4);
      w[ky][0] = a.x; w[ky][1] = a.y; w[ky][2] = a.z; w[ky][3] = a.w; w[ky][4] = b.x; w[ky][5] = b.y; w[ky][6] = b.z; w[ky][7] = b.w;
    }
#pragma unroll
    for (int iy = 0; iy < DW_TH; ++iy) {
      const int pix = iy * DW_TW + col + kx;
      const float4 a = *reinterpret_cast<const float4*>(tile + (pix * 2 + 0) * 32 + g * 4);
      const float4 b = *reinterpret_cast<const float4*>(tile + (pix * 2 + 1) * 32 + g * 4);
      const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        const int oy = iy - ky;                                    // output row fed by input row iy through filter row ky
        if (oy >= 0 && oy < DW_TILE_H) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[oy][j] = fmaf(w[ky][j], f[j], acc[oy][j]);
        }
      }
    }
  }
  // LayerNorm across the pixel's 64 channels = 8 consecutive lanes (two-pass, biased variance), row by row
  const int x = blockIdx.x * DW_TILE_W + col;
#pragma unroll
  for (int r = 0; r < DW_TILE_H; ++r) {
    const int y = blockIdx.y * DW_TILE_H + r;
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[r][j];
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    const float mean = s * (1.0f / 64.0f);
    float q = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float dlt = acc[r][j] - mean; q = fmaf(dlt, dlt, q); }
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
    const float rstd = rsqrtf(q * (1.0f / 64.0f) + 1e-6f);
    if (x < W && y < H) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(vsm[64 + g * 8 + j], (acc[r][j] - mean) * rstd, vsm[128 + g * 8 + j]);
      store8(dh, dl, (((size_t)img * H + y) * W + x) * 64 + g * 8, X3, o);
    }
  }
}

// The same operator at any width CH (a multiple of 32; used for CH = 128, BASELINE config 5 at the module level): one warp per output
// pixel, lane = CH / 32 consecutive channels, the 49 taps read through L1 (a tap of a warp is CH * 2 contiguous bytes), the filter
// transposed in shared memory, LayerNorm by warp shuffles.  No halo tile: it would not fit shared memory at 128 channels.
template <int CH, bool X3>
__global__ void __launch_bounds__(256) dwconv7_ln_wide_kernel(const __nv_bfloat16* __restrict__ sh, const __nv_bfloat16* __restrict__ sl,
                                                              __nv_bfloat16* __restrict__ dh, __nv_bfloat16* __restrict__ dl,
                                                              const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                              const float* __restrict__ ln_w, const float* __restrict__ ln_b, int H, int W,
                                                              long long total) {
  static_assert(CH % 64 == 0 && CH <= 256, "2 or 4 ... channels per lane, loaded as 32-bit pairs");
  constexpr int PER = CH / 32;
  __shared__ float wsm[49 * CH];                  // [tap][channel]
  __shared__ float vsm[3 * CH];                   // conv bias | LN weight | LN bias
  for (int i = threadIdx.x; i < 49 * CH; i += 256) { const int c = i / 49, t = i % 49; wsm[t * CH + c] = dw_w[c * 49 + t]; }
  for (int i = threadIdx.x; i < CH; i += 256) { vsm[i] = dw_b[i]; vsm[CH + i] = ln_w[i]; vsm[2 * CH + i] = ln_b[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c0 = lane * PER;
  const long long hw = (long long)H * W;
  for (long long p = (long long)blockIdx.x * 8 + warp; p < total; p += (long long)gridDim.x * 8) {
    const long long img = p / hw;
    const int rem = (int)(p - img * hw), y = rem / W, x = rem - y * W;
    float acc[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) acc[j] = vsm[c0 + j];
    for (int ky = 0; ky < 7; ++ky) {
      const int yy = y + ky - 3;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < 7; ++kx) {
        const int xx = x + kx - 3;
        if (xx < 0 || xx >= W) continue;
        const size_t off = ((size_t)(img * H + yy) * W + xx) * CH + c0;
        const float* wt = wsm + (ky * 7 + kx) * CH + c0;
#pragma unroll
        for (int j = 0; j < PER; j += 2) {
          const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(sh + off + j));
          float f0 = bf16_lo_f(v), f1 = bf16_hi_f(v);
          if (X3) {
            const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(sl + off + j));
            f0 += bf16_lo_f(u); f1 += bf16_hi_f(u);
          }
          acc[j] = fmaf(wt[j], f0, acc[j]);
          acc[j + 1] = fmaf(wt[j + 1], f1, acc[j + 1]);
        }
      }
    }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < PER; ++j) s += acc[j];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    const float mean = s * (1.0f / CH);
    float q = 0.0f;
#pragma unroll
    for (int j = 0; j < PER; ++j) { const float dlt = acc[j] - mean; q = fmaf(dlt, dlt, q); }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
    const float rstd = rsqrtf(q * (1.0f / CH) + 1e-6f);
    const size_t o = (size_t)p * CH + c0;
#pragma unroll
    for (int j = 0; j < PER; j += 2) {
      const float o0 = fmaf(vsm[CH + c0 + j], (acc[j] - mean) * rstd, vsm[2 * CH + c0 + j]);
      const float o1 = fmaf(vsm[CH + c0 + j + 1], (acc[j + 1] - mean) * rstd, vsm[2 * CH + c0 + j + 1]);
      const uint32_t h = pack_bf16x2(o0, o1);
      *reinterpret_cast<uint32_t*>(dh + o + j) = h;
      if (X3) *reinterpret_cast<uint32_t*>(dl + o + j) = pack_bf16x2(o0 - bf16_lo_f(h), o1 - bf16_hi_f(h));
    }
  }
}

// ---- ASPP image pooling: per-image channel means of a 64-channel NHWC tensor (two deterministic levels), then
// bias[img] = proj_w . relu(pool_w . mean + pool_b) + proj_b  (BatchNorms folded on the host) -------------------------------
constexpr int POOL_PARTS = 32;
template <int CH, bool X3>
__global__ void __launch_bounds__(256) pool_partial_kernel(const __nv_bfloat16* __restrict__ sh, const __nv_bfloat16* __restrict__ sl,
                                                           float* __restrict__ partial, int hw) {   // partial [img][POOL_PARTS][CH]
  constexpr int GROUPS = CH / 8, LANES = 256 / GROUPS;          // 8 channels per thread, LANES pixels per block iteration
  const int g = threadIdx.x % GROUPS, pl = threadIdx.x / GROUPS, img = blockIdx.y;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int px = blockIdx.x * LANES + pl; px < hw; px += POOL_PARTS * LANES) {
    const size_t off = ((size_t)img * hw + px) * CH + g * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(sh + off);
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[2 * i] += bf16_lo_f(w4[i]); acc[2 * i + 1] += bf16_hi_f(w4[i]); }
    if (X3) {
      const uint4 u = *reinterpret_cast<const uint4*>(sl + off);
      const uint32_t x4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[2 * i] += bf16_lo_f(x4[i]); acc[2 * i + 1] += bf16_hi_f(x4[i]); }
    }
  }
  __shared__ float red[LANES][CH + 1];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[pl][g * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < CH) {
    float t = 0.0f;
#pragma unroll
    for (int l = 0; l < LANES; ++l) t += red[l][threadIdx.x];
    partial[((size_t)img * POOL_PARTS + blockIdx.x) * CH + threadIdx.x] = t;
  }
}
template <int CH>
__global__ void __launch_bounds__(128) pool_bias_kernel(const float* __restrict__ partial, float inv_n, const float* __restrict__ pool_w,
                                                        const float* __restrict__ pool_b, const float* __restrict__ proj_w,
                                                        const float* __restrict__ proj_b, float* __restrict__ out) {
  __shared__ float mean_s[CH], v_s[128];
  const int img = blockIdx.x, t = threadIdx.x;
  for (int c = t; c < CH; c += 128) {
    float a = 0.0f;
    for (int k = 0; k < POOL_PARTS; ++k) a += partial[((size_t)img * POOL_PARTS + k) * CH + c];
    mean_s[c] = a * inv_n;
  }
  __syncthreads();
  float a = pool_b[t];
  for (int c = 0; c < CH; ++c) a = fmaf(pool_w[t * CH + c], mean_s[c], a);
  v_s[t] = fmaxf(a, 0.0f);
  __syncthreads();
  float b = proj_b[t];
  for (int c = 0; c < 128; ++c) b = fmaf(proj_w[t * 128 + c], v_s[c], b);
  out[(size_t)img * 128 + t] = b;
}

// ---- NHWC fp32 (recorded path states) -> NCHW fp32 (decoder input), gathered by slot ----------------------
__global__ void __launch_bounds__(256) unpack_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                          const int* __restrict__ slots, int C, int hw) {
  __shared__ float tile[32][65];
  const int o = blockIdx.z, cb = blockIdx.y * 64, p0 = blockIdx.x * 32;
  const int slot = slots ? slots[o] : o;
  // read: 32 pixels x 64 channels, 16 float4 per pixel
  const int px = threadIdx.x >> 3, g = threadIdx.x & 7;
  if (p0 + px < hw) {
    const float4* q = reinterpret_cast<const float4*>(src + ((size_t)slot * hw + p0 + px) * C + cb + g * 8);
    const float4 a = q[0], b = q[1];
    float* t = &tile[px][g * 8];
    t[0] = a.x; t[1] = a.y; t[2] = a.z; t[3] = a.w; t[4] = b.x; t[5] = b.y; t[6] = b.z; t[7] = b.w;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = w; c < 64; c += 8)
    if (p0 + lane < hw) dst[((size_t)o * C + cb + c) * hw + p0 + lane] = tile[lane][c];
}

// ---- kernels of the BEV Decoder head (models/decoder.py:91-140) ------------------------------------------------
// Space to depth for the stride-2 convolutions: dst[img][i][j][(2 py + px) C + c] = src[img][2 i + py][2 j + px][c].  A
// stride-2 convolution then is a stride-1 convolution over the four phase images (channel blocks of dst), which the implicit
// GEMM stage kernel runs as ordinary chunks with shifted windows.  Pure copy of 16-byte groups, per plane.
__global__ void __launch_bounds__(256) space_to_depth2_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n_img, int H, int W, int G) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)n_img * Ho * Wo * 4 * G;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    size_t r = idx / G;
    const int ph = (int)(r % 4); r /= 4;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const int img = (int)(r / Ho);
    dst[idx] = src[(((size_t)img * H + 2 * yo + (ph >> 1)) * W + 2 * xo + (ph & 1)) * G + g];
  }
}

// UpsamplingAdd (convolutions.py:204-215) after its 1x1 conv + BN were applied at the LOW resolution (they commute with the
// bilinear interpolation, whose weights sum to one): dst = bilinear_x2(src, align_corners=False) + skip.
// One thread = 8 channels of one output pixel.  out[2k] = .25 in[k-1] + .75 in[k], out[2k+1] = .75 in[k] + .25 in[k+1], clamped.
template <bool X3>
__global__ void __launch_bounds__(256) bilinear_up2_add_kernel(const __nv_bfloat16* __restrict__ sh, const __nv_bfloat16* __restrict__ sl,
                                                               const __nv_bfloat16* __restrict__ kh, const __nv_bfloat16* __restrict__ kl,
                                                               __nv_bfloat16* __restrict__ dh, __nv_bfloat16* __restrict__ dl, int n_img, int H,
                                                               int W, int C) {
  const int Ho = 2 * H, Wo = 2 * W, G = C / 8;
  const size_t total = (size_t)n_img * Ho * Wo * G;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(idx % G);
    size_t r = idx / G;
    const int xo = (int)(r % Wo); r /= Wo;
    const int yo = (int)(r % Ho);
    const int img = (int)(r / Ho);
    // torch: s = max(0, (o + 0.5) / 2 - 0.5); i0 = floor(s); i1 = min(i0 + 1, n - 1); lambda = s - i0
    const float sy = fmaxf(0.0f, (yo + 0.5f) * 0.5f - 0.5f), sx = fmaxf(0.0f, (xo + 0.5f) * 0.5f - 0.5f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float ly = sy - y0, lx = sx - x0;
    float v00[8], v01[8], v10[8], v11[8], k[8], o[8];
    const size_t base = (size_t)img * H;
    load8(sh, sl, ((base + y0) * W + x0) * C + g * 8, X3, v00);
    load8(sh, sl, ((base + y0) * W + x1) * C + g * 8, X3, v01);
    load8(sh, sl, ((base + y1) * W + x0) * C + g * 8, X3, v10);
    load8(sh, sl, ((base + y1) * W + x1) * C + g * 8, X3, v11);
    const size_t off = (((size_t)img * Ho + yo) * Wo + xo) * C + g * 8;
    load8(kh, kl, off, X3, k);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float top = v00[i] + lx * (v01[i] - v00[i]), bot = v10[i] + lx * (v11[i] - v10[i]);
      o[i] = top + ly * (bot - top) + k[i];
    }
    store8(dh, dl, off, X3, o);
  }
}

// Output 1x1 convolution of a Decoder head (64 -> K <= 4 channels, + bias, optional sigmoid) on NHWC planes -> fp32 NCHW
// [img][K][H][W]; with mask != nullptr also the arg-max over the K channels per pixel (first maximum wins, like torch.argmax),
// written as uint8: the graded occupancy mask never needs the 64-channel tensor to leave the device.
template <bool X3>
__global__ void __launch_bounds__(256) head_1x1_kernel(const __nv_bfloat16* __restrict__ sh, const __nv_bfloat16* __restrict__ sl,
                                                       const float* __restrict__ w, const float* __restrict__ b, int K, int sigmoid_out,
                                                       float* __restrict__ out, unsigned char* __restrict__ mask, int n_img, int hw) {
  __shared__ float ws[4 * 64 + 4];
  for (int i = threadIdx.x; i < K * 64; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < K) ws[256 + threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  const size_t total = (size_t)n_img * hw;
  for (size_t px = (size_t)blockIdx.x * blockDim.x + threadIdx.x; px < total; px += (size_t)gridDim.x * blockDim.x) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float f[8];
      load8(sh, sl, px * 64 + g * 8, X3, f);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < K) acc[k] = fmaf(ws[k * 64 + g * 8 + i], f[i], acc[k]);
    }
    const size_t img = px / hw, p = px % hw;
    int best = 0;
    float bv = -INFINITY;
    for (int k = 0; k < K; ++k) {
      float v = acc[k] + ws[256 + k];
      if (sigmoid_out) v = 1.0f / (1.0f + expf(-v));
      out[(img * K + k) * hw + p] = v;
      if (v > bv) { bv = v; best = k; }
    }
    if (mask) mask[px] = (unsigned char)best;
  }
}

// ---- row sharding: the halo rows of up to 6 NHWC tensors <-> one flat byte buffer, both directions in ONE launch ------------
// Tensor t is [B][rows_local][row_bytes_t]; range r copies rows [row0[r], row0[r] + nrows) of every tensor and batch entry
// to / from flat[r] (layout: tensor-major, then batch, then the nrows * row_bytes_t bytes), 16 bytes per thread step.
struct HaloCopy {
  char* base[6];
  long long batch_stride[6], row_bytes[6];      // bytes
  char* flat[2];
  int row0[2];
  int n_tensors, B, nrows, to_flat;
};
__global__ void __launch_bounds__(256) halo_copy_kernel(const HaloCopy h) {
  unsigned units = 0;                                  // 16-byte units per direction; 32-bit index arithmetic (a flat buffer is < 64 GB)
  for (int t = 0; t < h.n_tensors; ++t) units += (unsigned)h.B * (unsigned)h.nrows * (unsigned)(h.row_bytes[t] >> 4);
  for (int r = 0; r < 2; ++r) {
    if (h.flat[r] == nullptr) continue;
    for (unsigned u = blockIdx.x * blockDim.x + threadIdx.x; u < units; u += gridDim.x * blockDim.x) {
      unsigned rem = u;
      int t = 0;
      for (; t < h.n_tensors - 1; ++t) {
        const unsigned sz = (unsigned)h.B * (unsigned)h.nrows * (unsigned)(h.row_bytes[t] >> 4);
        if (rem < sz) break;
        rem -= sz;
      }
      const unsigned chunk = (unsigned)h.nrows * (unsigned)(h.row_bytes[t] >> 4);
      const unsigned b = rem / chunk, inner = rem - b * chunk;
      uint4* g = reinterpret_cast<uint4*>(h.base[t] + (long long)b * h.batch_stride[t] + (long long)h.row0[r] * h.row_bytes[t] + ((long long)inner << 4));
      uint4* f = reinterpret_cast<uint4*>(h.flat[r]) + u;
      if (h.to_flat) *f = *g; else *g = *f;
    }
  }
}

}  // namespace sf
