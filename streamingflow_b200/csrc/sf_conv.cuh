// sf_conv.cuh -- the implicit-GEMM convolution stage kernel (tcgen05 + TMEM + TMA) and its fused epilogues.
//
// One launch = one conv stage of an event over all active samples:
//   M = 128 output pixels per tile (16 rows x 8 columns of the NHWC grid; TMEM lane m <-> pixel (m/8, m%8)),
//   N = up to 256 fp32 accumulator columns in TMEM, double buffered (2 x 256 of the SM's 512 columns),
//   K = sum over "chunks": 64 input channels of one activation buffer x RxR taps.
// A operand: for every horizontal tap dx the producer TMA-loads ONE box of (16+R-1) rows x 8 pixels x 64 ch
//   (128 B per pixel, SWIZZLE_128B; out-of-image pixels are zero-filled by TMA = the conv's zero padding).
//   A vertical tap dy is then just a 1024-byte (8 pixel-rows) offset of the UMMA descriptor into that box, so
//   every descriptor start stays 1024-byte aligned.  R loads serve R*R taps.
// B operand: packed weights [rows][64] bf16, one TMA tile of n*nrep rows per tap, streamed through its own ring.
// Warp roles: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane), warp 2 = TMEM allocator,
//   warps 4..7 = epilogue (one thread per pixel: its TMEM lane holds all output channels of that pixel, so
//   LayerNorm / softmax-mix / GRU blends are thread-local).  Persistent over tiles, static round-robin.
#pragma once
#include "sf_ptx.cuh"
#include "../../include/sf_b200.h"

namespace sf {

constexpr int TILE_H = 16;
constexpr int TILE_W = 8;
constexpr int KC = 64;            // channels per K-chunk (128 bytes of bf16)
constexpr int ROW_BYTES = 128;    // one pixel of one chunk
constexpr int VEC_MAX = 512;      // per-stage constant vector (bias / LN / gate weights), floats
constexpr int ACC_STAGE_COLS = 256;

struct ChunkK {
  int R, n, nrep, col, wrow, init, c0, img_sel;   // img_sel: 0 = sample id, 1 = the event's x image index
};

struct EpiArgs {
  const float* s_in;
  const float* s_base;
  float* s_out;
  float* a32;
  float* b32;
  float* path;
  const float* eps;
  float* x32;
  float* params32;
  __nv_bfloat16* out_h[4];
  __nv_bfloat16* out_l[4];
  const __nv_bfloat16* in_h[2];
  const __nv_bfloat16* in_l[2];
  int n_out;      // output channels of the stage's main output (64 or 128)
  int kind;       // event kind (0 derivative step, 1 jump)
};

struct alignas(64) StageParams {
  CUtensorMap amap[SF_MAX_CHUNKS];
  CUtensorMap wmap;
  ChunkK chunk[SF_MAX_CHUNKS];
  int nchunk;
  int H, W, tiles_x, tiles_y, n_active;
  const int* sample_id;
  const int* x_img;
  const int* rec_slot;
  const int* eps_slot;
  const float* dt;
  const float* vec;
  int nvec;
  int a_slot_bytes, b_slot_bytes, nA, nB;
  int acc_stages;
  int* err;
  EpiArgs e;
};

// ------------------------------------------------------------------------------------------------
// epilogue helpers (one thread = one pixel)
// ------------------------------------------------------------------------------------------------
template <bool X3>
__device__ __forceinline__ void store_act16(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, const float (&v)[16]) {
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
  uint4* ph = reinterpret_cast<uint4*>(hi + off);
  ph[0] = make_uint4(h[0], h[1], h[2], h[3]);
  ph[1] = make_uint4(h[4], h[5], h[6], h[7]);
  if (X3) {
    uint32_t l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) l[i] = pack_bf16x2(v[2 * i] - bf16_lo_f(h[i]), v[2 * i + 1] - bf16_hi_f(h[i]));
    uint4* pl = reinterpret_cast<uint4*>(lo + off);
    pl[0] = make_uint4(l[0], l[1], l[2], l[3]);
    pl[1] = make_uint4(l[4], l[5], l[6], l[7]);
  }
}
template <bool X3>
__device__ __forceinline__ void load_act16(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off, float (&v)[16]) {
  const uint4* ph = reinterpret_cast<const uint4*>(hi + off);
  uint4 a = ph[0], b = ph[1];
  uint32_t h[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[2 * i] = bf16_lo_f(h[i]); v[2 * i + 1] = bf16_hi_f(h[i]); }
  if (X3) {
    const uint4* pl = reinterpret_cast<const uint4*>(lo + off);
    uint4 c = pl[0], d = pl[1];
    uint32_t l[8] = {c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[2 * i] += bf16_lo_f(l[i]); v[2 * i + 1] += bf16_hi_f(l[i]); }
  }
}
__device__ __forceinline__ void load_f32x16(const float* p, float (&v)[16]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float4 t = q[i]; v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
}
__device__ __forceinline__ void store_f32x16(float* p, const float (&v)[16]) {
  float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void zero16(float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.0f;
}

// LayerNorm over the 64 channels of one pixel (convolutions.py:299-304: biased variance, eps 1e-6) + exact GELU.
__device__ __forceinline__ void ln_gelu64(float (&v)[64], const float* w, const float* b) {
  float mean = 0.0f;
#pragma unroll
  for (int i = 0; i < 64; ++i) mean += v[i];
  mean *= (1.0f / 64.0f);
  float var = 0.0f;
#pragma unroll
  for (int i = 0; i < 64; ++i) { float d = v[i] - mean; var += d * d; }
  var *= (1.0f / 64.0f);
  const float rstd = 1.0f / sqrtf(var + 1e-6f);
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = gelu_erf(w[i] * ((v[i] - mean) * rstd) + b[i]);
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float t[16];
    tmem_ld16(taddr + j * 16, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[j * 16 + i] = t[i];
  }
}

struct PixelCtx {
  int bi;        // index into the event's active list
  int sid;       // sample id (image index of per-sample buffers)
  int y, x;
  bool valid;
  size_t pix;    // (sid*H + y)*W + x
};

// ------------------------------------------------------------------------------------------------
// fused epilogues.  taddr = TMEM address of this warp's lane quadrant, column 0 of the accumulator stage.
// All tcgen05.ld are executed by every lane (they are warp-collective); global traffic is predicated.
// ------------------------------------------------------------------------------------------------
template <int EPI, bool X3>
__device__ __forceinline__ void run_epilogue(const StageParams& p, const float* vec, uint32_t taddr, const PixelCtx& c) {
  const EpiArgs& e = p.e;
  if constexpr (EPI == SF_EPI_GATES) {
    // columns: [0,64) u1 | [64,128) r1 | [128,192) u2 | [192,256) r2 ; vec = the four biases in column order
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
      const int ucol = g * 128, rcol = g * 128 + 64;
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        float u[16], r[16], s[16];
        tmem_ld16(taddr + ucol + j * 16, u);
        tmem_ld16(taddr + rcol + j * 16, r);
        if (c.valid) load_f32x16(e.s_in + c.pix * 64 + j * 16, s); else zero16(s);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          u[i] = sigmoidf_(u[i] + vec[ucol + j * 16 + i]);
          r[i] = (1.0f - sigmoidf_(r[i] + vec[rcol + j * 16 + i])) * s[i];
        }
        if (c.valid) {
          store_act16<X3>(e.out_h[g], e.out_l[g], c.pix * 64 + j * 16, u);           // u1 / u2
          store_act16<X3>(e.out_h[2 + g], e.out_l[2 + g], c.pix * 64 + j * 16, r);   // (1-r1)*s / (1-r2)*s
        }
      }
    }
  } else if constexpr (EPI == SF_EPI_PROPOSE) {
    // columns: [0,64) s~1 | [64,128) s~2 ; vec = [bias~1, bias~2]; in[0]=u1, in[1]=u2; out[0]=a, out[1]=h
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      float t1[16], t2[16], s[16], u1[16], u2[16];
      tmem_ld16(taddr + j * 16, t1);
      tmem_ld16(taddr + 64 + j * 16, t2);
      if (c.valid) {
        load_f32x16(e.s_in + c.pix * 64 + j * 16, s);
        load_act16<X3>(e.in_h[0], e.in_l[0], c.pix * 64 + j * 16, u1);
        load_act16<X3>(e.in_h[1], e.in_l[1], c.pix * 64 + j * 16, u2);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          t1[i] = (1.0f - u1[i]) * s[i] + u1[i] * (t1[i] + vec[j * 16 + i]);
          t2[i] = (1.0f - u2[i]) * s[i] + u2[i] * (t2[i] + vec[64 + j * 16 + i]);
        }
        store_f32x16(e.a32 + c.pix * 64 + j * 16, t1);
        store_act16<X3>(e.out_h[0], e.out_l[0], c.pix * 64 + j * 16, t1);
        store_act16<X3>(e.out_h[1], e.out_l[1], c.pix * 64 + j * 16, t2);
      }
    }
  } else if constexpr (EPI == SF_EPI_DECODE) {
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      float b[16];
      tmem_ld16(taddr + j * 16, b);
      if (c.valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] += vec[j * 16 + i];
        store_f32x16(e.b32 + c.pix * 64 + j * 16, b);
        store_act16<X3>(e.out_h[0], e.out_l[0], c.pix * 64 + j * 16, b);
      }
    }
  } else if constexpr (EPI == SF_EPI_LNGELU) {
    float v[64];
    tmem_ld64(taddr, v);
    ln_gelu64(v, vec, vec + 64);
    if (c.valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) t[i] = v[j * 16 + i];
        store_act16<X3>(e.out_h[0], e.out_l[0], c.pix * 64 + j * 16, t);
      }
    }
  } else if constexpr (EPI == SF_EPI_MIX) {
    // columns: [0,64) 3x3 trunk conv | [64,128) 1x1 projection of cat[a,b];
    // vec = [LN w (64), LN b (64), gate row 0 (64), gate row 1 (64)]
    float l0 = 0.0f, l1 = 0.0f;
    {
      float v[64];
      tmem_ld64(taddr, v);
      ln_gelu64(v, vec, vec + 64);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float pr[16];
        tmem_ld16(taddr + 64 + j * 16, pr);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float yv = v[j * 16 + i] + gelu_erf(pr[i]);
          l0 = fmaf(vec[128 + j * 16 + i], yv, l0);
          l1 = fmaf(vec[192 + j * 16 + i], yv, l1);
        }
      }
    }
    if (c.valid) {
      const float g0 = 1.0f / (1.0f + expf(l1 - l0));     // softmax over the two logits, channel 0
      const float g1 = 1.0f - g0;
      const float dt = p.dt[c.bi];
      const int slot = p.rec_slot ? p.rec_slot[c.bi] : -1;
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        float a[16], b[16], o[16];
        load_f32x16(e.a32 + c.pix * 64 + j * 16, a);
        load_f32x16(e.b32 + c.pix * 64 + j * 16, b);
        if (e.kind == 0) {
          float si[16], sb[16];
          load_f32x16(e.s_in + c.pix * 64 + j * 16, si);
          load_f32x16(e.s_base + c.pix * 64 + j * 16, sb);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = sb[i] + dt * ((b[i] * g0 + a[i] * g1) - si[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = b[i] * g0 + a[i] * g1;
        }
        store_f32x16(e.s_out + c.pix * 64 + j * 16, o);
        store_act16<X3>(e.out_h[0], e.out_l[0], c.pix * 64 + j * 16, o);
        if (slot >= 0)
          store_f32x16(e.path + ((size_t)slot * p.H * p.W + (size_t)c.y * p.W + c.x) * 64 + j * 16, o);
      }
    }
  } else if constexpr (EPI == SF_EPI_BIAS_LRELU) {
    const int n = e.n_out;
#pragma unroll 1
    for (int j = 0; j < n / 16; ++j) {
      float v[16];
      tmem_ld16(taddr + j * 16, v);
      if (c.valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = lrelu01(v[i] + vec[j * 16 + i]);
        store_act16<X3>(e.out_h[0], e.out_l[0], c.pix * n + j * 16, v);
      }
    }
  } else if constexpr (EPI == SF_EPI_RES_PROJ) {
    // columns: [0,128) conv_2 (BN folded) | [128,256) 1x1 projection ; vec = [bn bias (128), proj bias (128)]
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      float v[16], q[16];
      tmem_ld16(taddr + j * 16, v);
      tmem_ld16(taddr + 128 + j * 16, q);
      if (c.valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = lrelu01(v[i] + vec[j * 16 + i]) + (q[i] + vec[128 + j * 16 + i]);
        store_act16<X3>(e.out_h[0], e.out_l[0], c.pix * 128 + j * 16, v);
      }
    }
  } else if constexpr (EPI == SF_EPI_RES_ID) {
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      float v[16], r[16];
      tmem_ld16(taddr + j * 16, v);
      if (c.valid) {
        load_act16<X3>(e.in_h[0], e.in_l[0], c.pix * 128 + j * 16, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = lrelu01(v[i] + vec[j * 16 + i]) + r[i];
        store_act16<X3>(e.out_h[0], e.out_l[0], c.pix * 128 + j * 16, v);
      }
    }
  } else if constexpr (EPI == SF_EPI_SAMPLE) {
    // columns: [0,64) loc | [64,128) raw scale ; vec = conv bias (128); eps is NCHW [slot][64][H][W]
    const size_t hw = (size_t)p.H * p.W;
    const float* eps = nullptr;
    if (c.valid) eps = e.eps + (size_t)p.eps_slot[c.bi] * 64 * hw + (size_t)c.y * p.W + c.x;
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      float loc[16], raw[16];
      tmem_ld16(taddr + j * 16, loc);
      tmem_ld16(taddr + 64 + j * 16, raw);
      if (c.valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          loc[i] = lrelu01(loc[i] + vec[j * 16 + i]);
          raw[i] = lrelu01(raw[i] + vec[64 + j * 16 + i]);
        }
        if (e.params32) {
          store_f32x16(e.params32 + c.pix * 128 + j * 16, loc);
          store_f32x16(e.params32 + c.pix * 128 + 64 + j * 16, raw);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) loc[i] = loc[i] + (softplus_(raw[i]) + 1e-8f) * __ldg(eps + (size_t)(j * 16 + i) * hw);
        store_act16<X3>(e.out_h[0], e.out_l[0], c.pix * 64 + j * 16, loc);
        if (e.x32) store_f32x16(e.x32 + c.pix * 64 + j * 16, loc);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the stage kernel
// ------------------------------------------------------------------------------------------------
template <int EPI, bool X3>
__global__ void __launch_bounds__(256, 1) conv_stage_kernel(const __grid_constant__ StageParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nA = p.nA, nB = p.nB;
  uint8_t* a_base = smem;
  uint8_t* b_base = a_base + (size_t)nA * p.a_slot_bytes;
  float* vec_s = reinterpret_cast<float*>(b_base + (size_t)nB * p.b_slot_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(vec_s + VEC_MAX);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + nA;
  uint64_t* b_full = a_empty + nA;
  uint64_t* b_empty = b_full + nB;
  uint64_t* acc_full = b_empty + nB;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < nA; ++i) { mbar_init(smem_u32(a_full + i), 1); mbar_init(smem_u32(a_empty + i), 1); }
    for (int i = 0; i < nB; ++i) { mbar_init(smem_u32(b_full + i), 1); mbar_init(smem_u32(b_empty + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(acc_full + i), 1); mbar_init(smem_u32(acc_empty + i), 128); }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
  for (int i = threadIdx.x; i < p.nvec; i += blockDim.x) vec_s[i] = p.vec[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tpi = p.tiles_x * p.tiles_y;
  const int ntiles = p.n_active * tpi;
  const int acc2 = (p.acc_stages == 2);

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      for (int c = 0; c < p.nchunk; ++c) tma_prefetch_desc(&p.amap[c]);
      tma_prefetch_desc(&p.wmap);
      uint32_t ia = 0, ib = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int bi = tile / tpi, rem = tile - bi * tpi;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int y0 = ty * TILE_H, x0 = tx * TILE_W;
        const int sid = p.sample_id[bi], ximg = p.x_img[bi];
        for (int c = 0; c < p.nchunk; ++c) {
          const ChunkK ck = p.chunk[c];
          const int pad = (ck.R - 1) >> 1;
          const int img = ck.img_sel ? ximg : sid;
          const uint32_t a_bytes = (uint32_t)(TILE_H + ck.R - 1) * TILE_W * ROW_BYTES;
          const uint32_t b_rows = (uint32_t)ck.n * ck.nrep;
          for (int dx = 0; dx < ck.R; ++dx) {
            const uint32_t sa = ia % nA, pha = (ia / nA) & 1;
            mbar_wait(smem_u32(a_empty + sa), pha ^ 1, p.err, 1);
            mbar_expect_tx(smem_u32(a_full + sa), a_bytes);
            tma_load_4d(smem_u32(a_base + (size_t)sa * p.a_slot_bytes), &p.amap[c], smem_u32(a_full + sa), ck.c0,
                        x0 + dx - pad, y0 - pad, img);
            ++ia;
            for (int dy = 0; dy < ck.R; ++dy) {
              const uint32_t sb = ib % nB, phb = (ib / nB) & 1;
              mbar_wait(smem_u32(b_empty + sb), phb ^ 1, p.err, 2);
              mbar_expect_tx(smem_u32(b_full + sb), b_rows * ROW_BYTES);
              const int row0 = ck.wrow + (dx * ck.R + dy) * (int)b_rows;
              const uint32_t dst = smem_u32(b_base + (size_t)sb * p.b_slot_bytes);
              for (uint32_t j = 0; j < b_rows / 64; ++j)
                tma_load_2d(dst + j * 64 * ROW_BYTES, &p.wmap, smem_u32(b_full + sb), 0, row0 + (int)j * 64);
              ++ib;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      uint32_t ia = 0, ib = 0, it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t as = acc2 ? (it & 1) : 0;
        const uint32_t aph = acc2 ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(smem_u32(acc_empty + as), aph ^ 1, p.err, 3);
        tc_fence_after();
        const uint32_t d_base = tmem_base + as * ACC_STAGE_COLS;
        for (int c = 0; c < p.nchunk; ++c) {
          const ChunkK ck = p.chunk[c];
          const uint32_t idesc = make_idesc_bf16(128, (uint32_t)ck.n);
          uint32_t accumulate = ck.init ? 0u : 1u;
          for (int dx = 0; dx < ck.R; ++dx) {
            const uint32_t sa = ia % nA, pha = (ia / nA) & 1;
            mbar_wait(smem_u32(a_full + sa), pha, p.err, 4);
            tc_fence_after();
            const uint32_t a_slot = smem_u32(a_base + (size_t)sa * p.a_slot_bytes);
            for (int dy = 0; dy < ck.R; ++dy) {
              const uint32_t sb = ib % nB, phb = (ib / nB) & 1;
              mbar_wait(smem_u32(b_full + sb), phb, p.err, 5);
              tc_fence_after();
              const uint32_t a_addr = a_slot + (uint32_t)dy * TILE_W * ROW_BYTES;
              const uint32_t b_slot = smem_u32(b_base + (size_t)sb * p.b_slot_bytes);
              for (int rep = 0; rep < ck.nrep; ++rep) {
                const uint32_t b_addr = b_slot + (uint32_t)rep * ck.n * ROW_BYTES;
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // 4 x (K = 16 bf16 = 32 bytes) per 64-channel chunk
                  umma_bf16(d_base + ck.col, make_sw128_desc(a_addr + k * 32), make_sw128_desc(b_addr + k * 32), idesc,
                            accumulate);
                  accumulate = 1u;
                }
              }
              umma_commit(smem_u32(b_empty + sb));   // frees the weight tile once these MMAs retire
              ++ib;
            }
            umma_commit(smem_u32(a_empty + sa));
            ++ia;
          }
        }
        umma_commit(smem_u32(acc_full + as));
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> fused math -> global =====================
    const int q = warp - 4;
    const int m = q * 32 + lane;
    const int r = m >> 3, cx = m & 7;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const uint32_t as = acc2 ? (it & 1) : 0;
      const uint32_t aph = acc2 ? ((it >> 1) & 1) : (it & 1);
      const int bi = tile / tpi, rem = tile - bi * tpi;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      PixelCtx c;
      c.bi = bi;
      c.sid = p.sample_id[bi];
      c.y = ty * TILE_H + r;
      c.x = tx * TILE_W + cx;
      c.valid = (c.y < p.H) && (c.x < p.W);
      c.pix = ((size_t)c.sid * p.H + c.y) * p.W + c.x;
      mbar_wait(smem_u32(acc_full + as), aph, p.err, 6);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * ACC_STAGE_COLS + ((uint32_t)(q * 32) << 16);
      run_epilogue<EPI, X3>(p, vec_s, taddr, c);
      tc_fence_before();
      mbar_arrive(smem_u32(acc_empty + as));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace sf
