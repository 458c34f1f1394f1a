// sf_conv.cuh -- the implicit-GEMM convolution stage kernel (tcgen05 + TMEM + TMA) and its fused epilogues.
//
// One launch = one conv stage of an event over all active samples:
//   CTA tile = 16 rows x (8*MT) columns of the NHWC grid = MT "M-tiles" of 128 pixels (TMEM lane m <-> pixel (m/8, m%8));
//       MT = 2 for stages with <= 128 accumulator columns, 1 for the 256-column stages (gates, q2),
//   N = up to 256/MT fp32 accumulator columns per M-tile; the SM's 512 TMEM columns hold 2 pipeline stages x MT M-tiles,
//       each (stage, M-tile) slot owned by its own epilogue warpgroup,
//   K = sum over "chunks": 64 input channels of one activation buffer x RxR taps.
// A operand: ONE TMA box per chunk: the tile plus its halo, (16+R-1) rows x (8*MT+R-1) pixels x 64 ch (128 B per pixel,
//   SWIZZLE_128B; out-of-image pixels are zero-filled by TMA = the conv's zero padding).  Every tap (dy,dx) and M-tile is
//   the SAME box read through a UMMA descriptor whose start is shifted by (dy*WP + dx + 8*mt) pixel rows and whose 8-row
//   group stride is WP*128 B.  The 128-byte swizzle is a function of the absolute shared-memory address (verified on
//   B200: scripts/exp_shift.py, profiles/r01_exp_shifted_descriptors.txt), so such 128-byte-aligned starts read correctly
//   with base_offset = 0.  One load of (tile + halo) serves R*R taps: activation re-reads from L2 drop from 3.4x to ~1.3x.
// B operand: packed weights [rows][64] bf16 streamed through their own ring, one TMA tile per tap or per column of taps;
//   each weight tile feeds the MMAs of all MT M-tiles.  Tap order is rotated per tile position so that the CTAs do not all
//   pull the same weight tile from L2 at once.
// Warp roles: warps [0, 8*MT) = epilogue warpgroups (one thread per pixel: its TMEM lane holds all output channels of that
//   pixel, so LayerNorm / softmax-mix / GRU blends are thread-local), then TMA producer (one lane), MMA issuer (one lane),
//   TMEM allocator.  Persistent over work items, static round-robin: whole tiles, then (when the last wave would leave more
//   than half of the CTAs idle) single M-tiles.
// Row-paired taps (the 7x7 trunk at 64 channels): vertically adjacent taps share one MMA of twice the width; the epilogue
//   folds the second column block back one row (see the lngelu epilogue and the MMA issuer).
// Epilogue global accesses are 256-bit with L1::no_allocate; a stage whose epilogue tail no longer reads TMEM hands the
//   accumulator back early (mix).  Launches use programmatic dependent launch: the prologue runs under the previous stage's tail.
#pragma once
#include <type_traits>

#include "sf_ptx.cuh"
#include "../../include/sf_b200.h"

namespace sf {

constexpr int TILE_H = 16;
constexpr int TILE_W = 8;
constexpr int KC = 64;            // channels per K-chunk (128 bytes of bf16)
constexpr int ROW_BYTES = 128;    // one pixel of one chunk
constexpr int VEC_MAX = 512;      // per-stage constant vector (bias / LN / gate weights), floats
constexpr int WG_SCRATCH = 128;   // floats of shared scratch per epilogue warpgroup (behind the constant vector)
constexpr int WG_SCRATCH_PAIR = 1024;   // ... for a stage with row-paired taps: 2 buffers x 4 warps x 8 lanes x 16 floats
constexpr int TMEM_COLS = 512;
constexpr int MAX_RING = 16;

struct ChunkK {
  int R, n, nrep, col, wrow, init, c0, img_sel;   // img_sel: 0 = sample id, 1 = the event's x image index
  int tb;                                         // taps (along dy) per B tile: 1 or R
  int ox, oy;                                     // pixel offset of the input window (dilated taps)
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
  int out_cs[4], out_co[4];   // channel count of each output buffer and the first channel this launch writes
  int in_cs[2], in_co[2];     // same for the element-wise inputs
  int n_out;      // output channels written by this launch (64 or 128)
  int kind;       // event kind (0 derivative step, 1 jump)
  int act;        // bias_act / residual epilogues: 0 LeakyReLU(0.1), 1 tanh, 2 ReLU, 3 identity, 4 GELU
  int act_after_res;   // res_id: the activation is applied to conv + bias + residual (stage flag 2048)
  int deriv;           // propose: write u (s~ - s) instead of the blend (stage flag 4096)
  int pairs;      // gate pairs / proposals handled by this launch (2 at C = 64 in the dual cell, else 1)
  const float* res_scale;  // optional SE scales [active sample][res_scale_ch] multiplied into the residual input (res_id)
  int res_scale_ch;
  float* out32;   // optional fp32 NHWC copy of the bias_act output [image][H][W][n_out]
  const float* img_bias;   // optional per-image bias [image][n_out] (bias_act)
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
  int debug;                 // experiments (SF_DEBUG_STAGE): 1 = epilogues skipped, 2 = MMAs skipped, 4 = operand loads skipped
  int pair_rows;             // 1: vertically adjacent taps are paired into one MMA of twice the width (see conv_stage_kernel)
  int wg_scratch;            // floats of shared scratch per epilogue warpgroup
  int w_rows_per_sample;     // > 0: per-sample weights (SE layer folded in): active sample bi reads rows [bi * this, (bi+1) * this)
  int b2b_wrow;              // lngelu_b2b: first row of the 1x1 follow-up conv's weights [CG n x 64 k] (hi, then lo in the split mode)
  int b2b_bytes;             // ... and their size in shared memory (loaded once per CTA)
  int resident_b;            // > 0: the stage's whole packed weight matrix (this many bytes) is loaded once per CTA into the (single) weight
                             // slot and stays for the launch: no weight ring, the issuer addresses taps by their row in the matrix
  int* err;
  EpiArgs e;
};

// M-tiles per CTA tile: 256-column launches get 1, the rest 2 (accumulator slot = 256 / MT columns).  CG = hidden channels.
__host__ __device__ constexpr int mtiles_for(int epi, int CG) {
  return (epi == SF_EPI_GATES || epi == SF_EPI_RES_PROJ || epi == 12 /* SF_EPI_PW_B2B */ || (CG == 128 && (epi == SF_EPI_MIX || epi == SF_EPI_SAMPLE))) ? 1 : 2;
}
constexpr int ACC_STAGES = 2;
// epilogue warpgroups per accumulator slot: the fused pointwise pair (256 GELUs per pixel: ALU / MUFU bound) splits a slot's columns
// between two warpgroups (warps with equal warp % 4 share a TMEM lane quadrant)
__host__ __device__ constexpr int wgs_per_slot(int epi) { return epi == 12 /* SF_EPI_PW_B2B */ ? 2 : 1; }
__host__ __device__ constexpr int a_box_bytes(int R, int MT) { return (TILE_H + R - 1) * (TILE_W * MT + R - 1) * ROW_BYTES; }

// ------------------------------------------------------------------------------------------------
// epilogue helpers (one thread = one pixel; 16 channels at a time)
// ------------------------------------------------------------------------------------------------
template <bool X3>
__device__ __forceinline__ void store_act16(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, const float (&v)[16]) {
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
  stg256(hi + off, h);
  if (X3) {
    uint32_t l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) l[i] = pack_bf16x2(v[2 * i] - bf16_lo_f(h[i]), v[2 * i + 1] - bf16_hi_f(h[i]));
    stg256(lo + off, l);
  }
}
template <bool X3>
__device__ __forceinline__ void load_act16(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off, float (&v)[16]) {
  uint32_t h[8];
  ldg256(hi + off, h);
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[2 * i] = bf16_lo_f(h[i]); v[2 * i + 1] = bf16_hi_f(h[i]); }
  if (X3) {
    uint32_t l[8];
    ldg256(lo + off, l);
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[2 * i] += bf16_lo_f(l[i]); v[2 * i + 1] += bf16_hi_f(l[i]); }
  }
}
__device__ __forceinline__ void load_f32x16(const float* p, float (&v)[16]) {
  uint32_t a[8], b[8];
  ldg256(p, a);
  ldg256(p + 8, b);
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = __uint_as_float(a[i]); v[8 + i] = __uint_as_float(b[i]); }
}
__device__ __forceinline__ void store_f32x16(float* p, const float (&v)[16]) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = __float_as_uint(v[i]); b[i] = __float_as_uint(v[8 + i]); }
  stg256(p, a);
  stg256(p + 8, b);
}
// Internal epilogue ids: the bias_act / residual epilogues with a run-time activation code, an optional fp32 copy and an
// optional per-image bias (codec and refinement stages).  The ODE loop's own stages (LeakyReLU only) use the lean
// SF_EPI_BIAS_LRELU / SF_EPI_RES_ID instantiations; launch_stage picks the variant from the stage flags.
constexpr int SF_EPI_BIAS_ACT = 9;
constexpr int SF_EPI_RES_ID_ACT = 10;
// lngelu followed, inside the same epilogue, by a 1x1 convolution + LayerNorm + GELU (the Bottleblock's layers 0..5 in one
// launch): a back-to-back GEMM whose A operand is written to tensor memory by the epilogue threads (stage flag 1024)
constexpr int SF_EPI_LNGELU_B2B = 11;
// res_id whose conv is the ConvNeXt block's pwconv1 (1x1, C -> 4C): GELU and pwconv2 (4C -> C) run inside the epilogue as a
// back-to-back GEMM on tensor-memory A operands, so the 4C-channel intermediate never leaves the SM (stage flag 8192)
constexpr int SF_EPI_PW_B2B = 12;
constexpr int SF_EPI_KERNELS = 13;
__host__ __device__ constexpr bool epi_has_b2b(int epi) { return epi == SF_EPI_LNGELU_B2B || epi == SF_EPI_PW_B2B; }

// per-warpgroup state of the back-to-back GEMM (shared-window addresses; phase = parity of the group's completion barrier)
struct B2BCtx {
  uint32_t w_smem, w_full, done_bar, phase;
  bool w_ready;
  uint32_t peer_done;      // the completion barrier of the slot's other warpgroup (stages with two warpgroups per slot)
};

// v = act(v + b) on a 16-channel slice; the activation code is uniform per launch, so the switch is taken once per slice
template <bool ANYACT>
__device__ __forceinline__ void bias_act16(float (&v)[16], const float (&b)[16], int act) {
  if (!ANYACT || act == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = lrelu01(v[i] + b[i]);
  } else if (act == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i] + b[i]);
  } else if (act == 2) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + b[i], 0.0f);
  } else if (act == 4) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i] + b[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = v[i] + b[i];
  }
}
__device__ __forceinline__ void zero16(float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.0f;
}
// 16 floats of the per-stage constant vector from shared memory (128-bit broadcast LDS; vec = shared-window address)
__device__ __forceinline__ void vec16(uint32_t vec, int off, float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float4 t = lds128(vec + (uint32_t)(off + 4 * i) * 4u); v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
}

// LayerNorm statistics over the CG accumulator columns [col, col+CG) of this thread's pixel, two passes over TMEM
// (convolutions.py:299-304: biased variance, eps 1e-6).  Returns mean and 1/sqrt(var + eps).
template <int CG>
__device__ __forceinline__ void ln_stats(uint32_t taddr, float& mean, float& rstd) {
  float s = 0.0f;
#pragma unroll 1
  for (int j = 0; j < CG / 32; ++j) {
    float a[16], b[16];
    tmem_ld16x2(taddr + j * 32, taddr + j * 32 + 16, a, b);
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + b[i];
  }
  mean = s * (1.0f / CG);
  float q = 0.0f;
#pragma unroll 1
  for (int j = 0; j < CG / 32; ++j) {
    float a[16], b[16];
    tmem_ld16x2(taddr + j * 32, taddr + j * 32 + 16, a, b);
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float da = a[i] - mean, db = b[i] - mean; q = fmaf(da, da, q); q = fmaf(db, db, q); }
  }
  rstd = rsqrtf(q * (1.0f / CG) + 1e-6f);
}

struct PixelCtx {
  int bi;        // index into the event's active list
  int sid;       // sample id (image index of per-sample buffers)
  int y, x;
  bool valid;
  size_t pix;    // (sid*H + y)*W + x
  int wg, m;     // epilogue warpgroup index and the thread's index in it (= TMEM lane)
  int sub;       // which of the slot's warpgroups this is (stages with wgs_per_slot > 1)
};

// Row-paired taps (stage flag 512, 64-channel stages): column block 1 [64, 128) of lane m holds the partial sum that belongs to the
// pixel ONE ROW BELOW (lane m + 8).  Fold it in: block0[m] += block1[m - 8].  Inside a warp that is a shuffle by 8 lanes; the first
// row of a warp takes the last row of the warp above through the warpgroup's shared scratch (double-buffered per slice; one named
// barrier per slice).  The tile's first row (lane row 0) is a scratch row: it only feeds row 1.  Afterwards block 0 holds the
// complete accumulator and every epilogue runs unchanged.
__device__ __forceinline__ void fold_paired_rows(const StageParams& p, uint32_t vec, uint32_t taddr, const PixelCtx& c) {
  constexpr int N = 64;
  const int lane = c.m & 31, q = c.m >> 5;
  const uint32_t scratch = vec + (uint32_t)(VEC_MAX + c.wg * p.wg_scratch) * 4u;
#pragma unroll 1
  for (int j = 0; j < N / 16; ++j) {
    float v0[16], v1[16];
    tmem_ld16x2(taddr + j * 16, taddr + N + j * 16, v0, v1);
    const uint32_t buf = scratch + (uint32_t)((j & 1) * 512 + q * 128) * 4u;
    if (lane >= 24) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(buf + (uint32_t)((lane - 24) * 16 + 4 * i) * 4u), "f"(v1[4 * i]),
                     "f"(v1[4 * i + 1]), "f"(v1[4 * i + 2]), "f"(v1[4 * i + 3]) : "memory");
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + c.wg) : "memory");
    float up[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) up[i] = __shfl_up_sync(0xffffffffu, v1[i], 8);
    if (lane < 8) {
      if (q > 0) vec16(buf - 128u * 4u, lane * 16, up); else zero16(up);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v0[i] += up[i];
    tmem_st16(taddr + j * 16, v0);
  }
  tmem_st_wait();
}
// epilogues whose stage may keep its weights resident in shared memory (StageParams::resident_b).  The register-bound gates /
// propose / sample instantiations and the fused trunk (their weights never fit anyway) compile the path out: with it in, ptxas
// spilled 28 instead of 8 bytes in propose and the stage ran ~5 % slower.
__host__ __device__ constexpr bool epi_can_be_resident(int epi) {
  return !(epi == SF_EPI_GATES || epi == SF_EPI_PROPOSE || epi == SF_EPI_SAMPLE || epi == SF_EPI_LNGELU_B2B);
}
// epilogues that accept row-paired taps: one 64-column accumulator block at column 0 (the register-bound propose / mix / sample
// epilogues stay out: the extra code costs them spills)
__host__ __device__ constexpr bool epi_can_pair(int epi) {
  return epi == SF_EPI_LNGELU || epi == SF_EPI_LNGELU_B2B || epi == SF_EPI_DECODE || epi == SF_EPI_BIAS_LRELU || epi == SF_EPI_BIAS_ACT ||
         epi == SF_EPI_RES_ID || epi == SF_EPI_RES_ID_ACT;
}


// ------------------------------------------------------------------------------------------------
// fused epilogues.  taddr = TMEM address of this warp's lane quadrant, column 0 of the accumulator stage.
// All tcgen05.ld are executed by every lane (they are warp-collective); global traffic is predicated.
// ------------------------------------------------------------------------------------------------
// acc_empty: the accumulator stage's EMPTY barrier.  Epilogues whose tail no longer reads TMEM arrive on it themselves as soon
// as their last tcgen05.ld has completed (returning true), so the next tile's MMAs overlap that tail; otherwise the caller
// arrives after the epilogue returns.
template <int EPI, bool X3, int CG>
__device__ __forceinline__ bool run_epilogue(const StageParams& p, uint32_t vec, uint32_t taddr, const PixelCtx& c, uint32_t acc_empty, B2BCtx& b2b) {
  const EpiArgs& e = p.e;
  constexpr int NJ = CG / 16;                      // 16-channel slices of one CG-channel tensor
  const size_t pc = c.pix * CG;                    // this pixel in a CG-channel NHWC tensor
  if constexpr (CG == 64 && epi_can_pair(EPI)) {
    if (p.pair_rows) fold_paired_rows(p, vec, taddr, c);
  }
  if constexpr (EPI == SF_EPI_GATES) {
    // columns: pair g = [u_g (CG) | r_g (CG)]; CG = 64: two pairs (u1 r1 u2 r2) in one launch, CG = 128: one pair per launch.
    // vec = biases in column order; out[2g] = u_g, out[2g+1] = (1 - r_g) * s
    auto body = [&](auto npairs) {
      constexpr int NP = decltype(npairs)::value;
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        float s[16];
        if (c.valid) load_f32x16(e.s_in + pc + j * 16, s); else zero16(s);
#pragma unroll
        for (int g = 0; g < NP; ++g) {
          const int ucol = g * 2 * CG, rcol = ucol + CG;
          float u[16], r[16], bu[16], br[16];
          tmem_ld16x2(taddr + ucol + j * 16, taddr + rcol + j * 16, u, r);
          vec16(vec, ucol + j * 16, bu);
          vec16(vec, rcol + j * 16, br);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            u[i] = sigmoidf_(u[i] + bu[i]);
            r[i] = one_minus_sigmoidf_(r[i] + br[i]) * s[i];
          }
          if (c.valid) {
            store_act16<X3>(e.out_h[2 * g], e.out_l[2 * g], pc + j * 16, u);
            store_act16<X3>(e.out_h[2 * g + 1], e.out_l[2 * g + 1], pc + j * 16, r);
          }
        }
      }
    };
    if (CG == 64 && e.pairs == 2) body(std::integral_constant<int, (CG == 64 ? 2 : 1)>{}); else body(std::integral_constant<int, 1>{});
  } else if constexpr (EPI == SF_EPI_PROPOSE) {
    // columns: proposal k = [s~_k (CG)]; CG = 64: both GRUs in one launch, CG = 128: one per launch.  vec = biases;
    // in[k] = u_k; out[k] = (1-u_k) s + u_k s~_k; the first GRU's blend is also kept in fp32 (a32) when bound.
    auto body = [&](auto npairs) {
      constexpr int NP = decltype(npairs)::value;
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        float s[16], u[NP][16];
        if (c.valid) {
          load_f32x16(e.s_in + pc + j * 16, s);
#pragma unroll
          for (int k = 0; k < NP; ++k) load_act16<X3>(e.in_h[k], e.in_l[k], pc + j * 16, u[k]);
        } else {
          zero16(s);
#pragma unroll
          for (int k = 0; k < NP; ++k) zero16(u[k]);
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
          float t[16], b[16];
          tmem_ld16(taddr + k * CG + j * 16, t);
          vec16(vec, k * CG + j * 16, b);
          if (e.act == 2 || e.deriv) {
            // the plain ConvGRU cells (temporal_ode_bayes.py:14-61, 165-208): ReLU proposal (BatchNorm folded into the conv),
            // and for the ODE variant the derivative u (s~ - s) instead of the blended state
            const bool relu = e.act == 2, deriv = e.deriv != 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float st = t[i] + b[i];
              st = relu ? fmaxf(st, 0.0f) : st;
              t[i] = deriv ? u[k][i] * (st - s[i]) : fmaf(u[k][i], st - s[i], s[i]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = fmaf(u[k][i], (t[i] + b[i]) - s[i], s[i]);      // (1 - u) s + u s~
          }
          if (c.valid) {
            if (k == 0 && e.a32) store_f32x16(e.a32 + pc + j * 16, t);
            store_act16<X3>(e.out_h[k], e.out_l[k], pc + j * 16, t);
          }
        }
      }
    };
    if (CG == 64 && e.pairs == 2) body(std::integral_constant<int, (CG == 64 ? 2 : 1)>{}); else body(std::integral_constant<int, 1>{});
  } else if constexpr (EPI == SF_EPI_DECODE) {
#pragma unroll 1
    for (int j = 0; j < NJ; ++j) {
      float b[16], bb[16];
      tmem_ld16(taddr + j * 16, b);
      vec16(vec, j * 16, bb);
      if (c.valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] += bb[i];
        store_f32x16(e.b32 + pc + j * 16, b);
        store_act16<X3>(e.out_h[0], e.out_l[0], pc + j * 16, b);
      }
    }
  } else if constexpr (EPI == SF_EPI_LNGELU || EPI == SF_EPI_LNGELU_B2B) {
    float mean, rstd;
    ln_stats<CG>(taddr, mean, rstd);
    if constexpr (EPI == SF_EPI_LNGELU) {
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        float v[16], w[16], b[16];
        tmem_ld16(taddr + j * 16, v);
        vec16(vec, j * 16, w);
        vec16(vec, CG + j * 16, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = gelu_erf(fmaf(w[i], (v[i] - mean) * rstd, b[i]));
        if (c.valid) store_act16<X3>(e.out_h[0], e.out_l[0], pc + j * 16, v);
      }
    } else {
      // Back-to-back GEMM.  t1 = GELU(LN(conv)) never leaves the SM: every thread writes its pixel's CG channels as bf16 pairs
      // into the tensor-memory columns [CG, CG + CG/2) of its lane (block 1 is dead after the fold) -- the A operand [128 pixels
      // x CG] of the 1x1 convolution, whose weights [CG x CG] sit in shared memory for the whole launch.  One thread of the
      // warpgroup issues the MMAs (K = CG) into columns [0, CG) (every lane has read its conv result by then), everybody waits
      // for their completion barrier, and the second LayerNorm + GELU runs on the new accumulator.  Split mode: the residual
      // plane goes to [CG + CG/2, 2 CG) and three products are issued (t1h Wh + t1h Wl + t1l Wh).
      static_assert(CG == 64, "the fused trunk is built for 64 channels (128 columns per accumulator slot)");
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        float v[16], w[16], b[16];
        tmem_ld16(taddr + j * 16, v);
        vec16(vec, j * 16, w);
        vec16(vec, CG + j * 16, b);
        uint32_t h[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = gelu_erf(fmaf(w[i], (v[i] - mean) * rstd, b[i]));
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        tmem_st8(taddr + CG + j * 8, h);
        if (X3) {
          uint32_t l[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) l[i] = pack_bf16x2(v[2 * i] - bf16_lo_f(h[i]), v[2 * i + 1] - bf16_hi_f(h[i]));
          tmem_st8(taddr + CG + CG / 2 + j * 8, l);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + c.wg) : "memory");
      if (c.m == 0) {
        tc_fence_after();
        if (!b2b.w_ready) mbar_wait(b2b.w_full, 0, p.err, 7);
        constexpr uint32_t B_HI = 64u | (1u << 14) | (2u << 29);       // SBO = 1024 B, version 1, SWIZZLE_128B
        const uint32_t idesc = make_idesc_bf16(128, CG);
        const uint32_t w_lo = (b2b.w_smem & 0x3FFFFu) >> 4;
#pragma unroll
        for (uint32_t k = 0; k < CG / 16; ++k)
          umma_bf16_ts(taddr, taddr + CG + 8 * k, ((uint64_t)B_HI << 32) | (w_lo + 2 * k), idesc, k > 0 ? 1u : 0u);
        if (X3) {
          const uint32_t wl_lo = w_lo + ((CG * ROW_BYTES) >> 4);
#pragma unroll
          for (uint32_t k = 0; k < CG / 16; ++k)
            umma_bf16_ts(taddr, taddr + CG + 8 * k, ((uint64_t)B_HI << 32) | (wl_lo + 2 * k), idesc, 1u);
#pragma unroll
          for (uint32_t k = 0; k < CG / 16; ++k)
            umma_bf16_ts(taddr, taddr + CG + CG / 2 + 8 * k, ((uint64_t)B_HI << 32) | (w_lo + 2 * k), idesc, 1u);
        }
        umma_commit(b2b.done_bar);
      }
      b2b.w_ready = true;
      mbar_wait(b2b.done_bar, b2b.phase, p.err, 8);
      b2b.phase ^= 1;
      tc_fence_after();
      ln_stats<CG>(taddr, mean, rstd);
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        float v[16], w[16], b[16];
        tmem_ld16(taddr + j * 16, v);
        vec16(vec, 2 * CG + j * 16, w);
        vec16(vec, 3 * CG + j * 16, b);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = gelu_erf(fmaf(w[i], (v[i] - mean) * rstd, b[i]));
        if (c.valid) store_act16<X3>(e.out_h[0], e.out_l[0], pc + j * 16, v);
      }
    }
  } else if constexpr (EPI == SF_EPI_MIX) {
    // columns: [0,CG) 3x3 trunk conv | [CG,2CG) 1x1 projection of cat[a,b];
    // vec = [LN w (CG), LN b (CG), gate row 0 (CG), gate row 1 (CG)]
    float mean, rstd, l0 = 0.0f, l1 = 0.0f;
    ln_stats<CG>(taddr, mean, rstd);
#pragma unroll 1
    for (int j = 0; j < NJ; ++j) {
      float v[16], pr[16], w[16], b[16], g0w[16], g1w[16];
      tmem_ld16x2(taddr + j * 16, taddr + CG + j * 16, v, pr);
      vec16(vec, j * 16, w);
      vec16(vec, CG + j * 16, b);
      vec16(vec, 2 * CG + j * 16, g0w);
      vec16(vec, 3 * CG + j * 16, g1w);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float yv = gelu_erf(fmaf(w[i], (v[i] - mean) * rstd, b[i])) + gelu_erf(pr[i]);
        l0 = fmaf(g0w[i], yv, l0);
        l1 = fmaf(g1w[i], yv, l1);
      }
    }
    // the state update below works on global memory only: hand the accumulator back now
    tc_fence_before();
    mbar_arrive(acc_empty);
    if (c.valid) {
      const float g0 = __fdividef(1.0f, 1.0f + __expf(l1 - l0));     // softmax over the two logits, channel 0
      const float g1 = 1.0f - g0;
      const float dt = p.dt[c.bi];
      const int slot = p.rec_slot ? p.rec_slot[c.bi] : -1;
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        float a[16], b[16], o[16];
        load_f32x16(e.a32 + pc + j * 16, a);
        load_f32x16(e.b32 + pc + j * 16, b);
        if (e.kind == 0) {
          float si[16];
          load_f32x16(e.s_in + pc + j * 16, si);
          if (e.s_base == e.s_in) {            // Euler: the update starts from the state the cell read
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = si[i] + dt * ((b[i] * g0 + a[i] * g1) - si[i]);
          } else {                             // midpoint stage 2: base = s, cell state = k
            float sb[16];
            load_f32x16(e.s_base + pc + j * 16, sb);
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = sb[i] + dt * ((b[i] * g0 + a[i] * g1) - si[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = b[i] * g0 + a[i] * g1;
        }
        store_f32x16(e.s_out + pc + j * 16, o);
        store_act16<X3>(e.out_h[0], e.out_l[0], pc + j * 16, o);
        if (slot >= 0)
          store_f32x16(e.path + ((size_t)slot * p.H * p.W + (size_t)c.y * p.W + c.x) * CG + j * 16, o);
      }
    }
  } else if constexpr (EPI == SF_EPI_BIAS_LRELU || EPI == SF_EPI_BIAS_ACT) {
    constexpr bool ANYACT = EPI == SF_EPI_BIAS_ACT;
    const int n = e.n_out;
    const size_t o0 = c.pix * e.out_cs[0] + e.out_co[0];
#pragma unroll 1
    for (int j = 0; j < n / 16; ++j) {
      float v[16], b[16];
      tmem_ld16(taddr + j * 16, v);
      vec16(vec, j * 16, b);
      if (c.valid) {
        const int act = e.act;
        if (ANYACT && e.img_bias) {
          float ib[16];
          load_f32x16(e.img_bias + (size_t)c.sid * n + j * 16, ib);
#pragma unroll
          for (int i = 0; i < 16; ++i) b[i] += ib[i];
        }
        bias_act16<ANYACT>(v, b, act);
        store_act16<X3>(e.out_h[0], e.out_l[0], o0 + j * 16, v);
        if (ANYACT && e.out32) store_f32x16(e.out32 + c.pix * n + j * 16, v);
      }
    }
  } else if constexpr (EPI == SF_EPI_RES_PROJ) {
    // columns: [0,n) conv_2 (BN folded) | [n,2n) 1x1 projection, for n = n_out (64 or 128) output channels starting at
    // out_co; vec = [bn bias (n), proj bias (n)]
    const int n = e.n_out;
    const size_t o0 = c.pix * e.out_cs[0] + e.out_co[0];
#pragma unroll 1
    for (int j = 0; j < n / 16; ++j) {
      float v[16], q[16], b[16], pb[16];
      tmem_ld16x2(taddr + j * 16, taddr + n + j * 16, v, q);
      vec16(vec, j * 16, b);
      vec16(vec, n + j * 16, pb);
      if (c.valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = lrelu01(v[i] + b[i]) + (q[i] + pb[i]);
        store_act16<X3>(e.out_h[0], e.out_l[0], o0 + j * 16, v);
      }
    }
  } else if constexpr (EPI == SF_EPI_RES_ID || EPI == SF_EPI_RES_ID_ACT) {
    const size_t o0 = c.pix * e.out_cs[0] + e.out_co[0], i0 = c.pix * e.in_cs[0] + e.in_co[0];
    // residual = SE output z * scale[sample][channel] when the SE layer is folded into its consumers: the tile's (one sample's)
    // scales are staged once in the warpgroup's shared scratch (warpgroup-wide named barrier, id 1 + wg)
    const uint32_t sc_s = vec + (uint32_t)(VEC_MAX + c.wg * p.wg_scratch) * 4u;
    if (e.res_scale) {
      asm volatile("bar.sync %0, 128;" ::"r"(1 + c.wg) : "memory");          // readers of the previous tile are done
      if (c.m < e.n_out) {
        const float sv = __ldg(e.res_scale + (size_t)c.bi * e.res_scale_ch + e.in_co[0] + c.m);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(sc_s + (uint32_t)c.m * 4u), "f"(sv) : "memory");
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + c.wg) : "memory");
    }
#pragma unroll 1
    for (int j = 0; j < e.n_out / 16; ++j) {
      float v[16], r[16], b[16];
      if (c.valid) load_act16<X3>(e.in_h[0], e.in_l[0], i0 + j * 16, r); else zero16(r);
      if (e.res_scale) {
        float sc[16];
        vec16(sc_s, j * 16, sc);
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] *= sc[i];
      }
      tmem_ld16(taddr + j * 16, v);
      vec16(vec, j * 16, b);
      if (c.valid) {
        if (EPI == SF_EPI_RES_ID_ACT && e.act_after_res) {       // ResNet BasicBlock: act(conv + bias + residual)
#pragma unroll
          for (int i = 0; i < 16; ++i) { v[i] += r[i]; r[i] = 0.0f; }
        }
        bias_act16<EPI == SF_EPI_RES_ID_ACT>(v, b, e.act);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += r[i];
        store_act16<X3>(e.out_h[0], e.out_l[0], o0 + j * 16, v);
      }
    }
  } else if constexpr (EPI == SF_EPI_PW_B2B) {
    // ConvNeXt block, pointwise pair (convolutions.py:334-344): columns [0, 4C) = pwconv1 (1x1, C -> 4C) of this pixel.
    // The slot's TWO warpgroups each own one half of 2C columns (sub = 0 / 1): GELU(acc + b1) goes back into the columns it came
    // from as packed bf16 pairs (the first C of the half: a slice is read before the narrower slice that replaces it is written) --
    // the A operand [128 pixels x 2C] of that half of pwconv2, whose [C x 4C] weights (layer scale folded, four K-chunks of
    // [C rows x 64]) stay in shared memory for the whole launch.  One thread of the warpgroup issues the half's 2C/16 MMAs into the
    // half's upper C columns (dead once the half has been read) and commits to the warpgroup's barrier; both warpgroups then wait for
    // BOTH partial products and each finishes C/2 output channels: out = part0 + part1 + gamma b2 + residual (the block's input).
    // vec = [b1 (4C) | gamma * b2 (C)]; b2b.peer_done = the other warpgroup's barrier.
    static_assert(CG == 64 && !X3, "the fused pointwise pair is built for 64 channels, bf16 operands");
    const size_t o0 = c.pix * e.out_cs[0] + e.out_co[0], i0 = c.pix * e.in_cs[0] + e.in_co[0];
    constexpr uint32_t B_HI = 64u | (1u << 14) | (2u << 29);       // SBO = 1024 B, version 1, SWIZZLE_128B
    const uint32_t idesc = make_idesc_bf16(128, CG);
    const uint32_t w_lo = (b2b.w_smem & 0x3FFFFu) >> 4;
    const uint32_t half = (uint32_t)c.sub;
    const uint32_t base = taddr + half * 2 * CG;
#pragma unroll 1
    for (int j = 0; j < 2 * CG / 16; ++j) {
      float v[16], b[16];
      tmem_ld16(base + j * 16, v);
      vec16(vec, (int)half * 2 * CG + j * 16, b);
      uint32_t h[8];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i] + b[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
      tmem_st8(base + j * 8, h);
    }
    tmem_st_wait();
    tc_fence_before();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + c.wg) : "memory");
    if (c.m == 0) {
      tc_fence_after();
      if (!b2b.w_ready) mbar_wait(b2b.w_full, 0, p.err, 7);
#pragma unroll
      for (uint32_t k = 0; k < 2 * CG / 16; ++k)      // K-chunk 2 half + k / 4 (8 KB = 512 descriptor units apart), 16-wide step k % 4
        umma_bf16_ts(base + CG, base + 8 * k, ((uint64_t)B_HI << 32) | (w_lo + (2 * half + (k >> 2)) * ((CG * ROW_BYTES) >> 4) + 2 * (k & 3)), idesc,
                     k ? 1u : 0u);
      umma_commit(b2b.done_bar);
    }
    b2b.w_ready = true;
    mbar_wait(b2b.done_bar, b2b.phase, p.err, 8);
    mbar_wait(b2b.peer_done, b2b.phase, p.err, 9);
    b2b.phase ^= 1;
    tc_fence_after();
#pragma unroll 1
    for (int j = (int)half * (CG / 32); j < ((int)half + 1) * (CG / 32); ++j) {
      float v[16], q[16], r[16], b[16];
      if (c.valid) load_act16<X3>(e.in_h[0], e.in_l[0], i0 + j * 16, r); else zero16(r);
      tmem_ld16x2(taddr + CG + j * 16, taddr + 3 * CG + j * 16, v, q);
      vec16(vec, 4 * CG + j * 16, b);
      if (c.valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = ((v[i] + q[i]) + b[i]) + r[i];
        store_act16<X3>(e.out_h[0], e.out_l[0], o0 + j * 16, v);
      }
    }
  } else if constexpr (EPI == SF_EPI_SAMPLE) {
    // columns: [0,CG) loc | [CG,2CG) raw scale ; vec = conv bias (2CG); eps is NCHW [slot][CG][H][W]
    const size_t hw = (size_t)p.H * p.W;
    const float* eps = nullptr;
    if (c.valid) eps = e.eps + (size_t)p.eps_slot[c.bi] * CG * hw + (size_t)c.y * p.W + c.x;
#pragma unroll 1
    for (int j = 0; j < NJ; ++j) {
      float loc[16], raw[16], bl[16], br[16], ep[16];
      if (c.valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) ep[i] = __ldg(eps + (size_t)(j * 16 + i) * hw);
      } else {
        zero16(ep);
      }
      tmem_ld16x2(taddr + j * 16, taddr + CG + j * 16, loc, raw);
      vec16(vec, j * 16, bl);
      vec16(vec, CG + j * 16, br);
      if (c.valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          loc[i] = lrelu01(loc[i] + bl[i]);
          raw[i] = lrelu01(raw[i] + br[i]);
        }
        if (e.params32) {
          store_f32x16(e.params32 + c.pix * 2 * CG + j * 16, loc);
          store_f32x16(e.params32 + c.pix * 2 * CG + CG + j * 16, raw);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) loc[i] = loc[i] + (softplus_(raw[i]) + 1e-8f) * ep[i];
        store_act16<X3>(e.out_h[0], e.out_l[0], pc + j * 16, loc);
        if (e.x32) store_f32x16(e.x32 + pc + j * 16, loc);
      }
    }
  }
  return EPI == SF_EPI_MIX;
}

// ------------------------------------------------------------------------------------------------
// the stage kernel
// ------------------------------------------------------------------------------------------------
template <int EPI, bool X3, int CG>
__global__ void __launch_bounds__(128 + 128 * ACC_STAGES * mtiles_for(EPI, CG) * wgs_per_slot(EPI), 1) conv_stage_kernel(const __grid_constant__ StageParams p) {
  constexpr int S = ACC_STAGES;
  constexpr int MT = mtiles_for(EPI, CG);
  constexpr int WGS = wgs_per_slot(EPI);                // epilogue warpgroups per accumulator slot
  constexpr int NGROUPS = S * MT * WGS;                 // epilogue warpgroups
  constexpr uint32_t STAGE_COLS = TMEM_COLS / S;        // 256
  constexpr uint32_t SLOT_COLS = STAGE_COLS / MT;       // 256 or 128
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nA = p.nA, nB = p.nB;
  uint8_t* a_base = smem;
  uint8_t* b_base = a_base + (size_t)nA * p.a_slot_bytes;
  uint8_t* b2b_w = b_base + (size_t)nB * p.b_slot_bytes;                 // weights of the fused 1x1 follow-up conv (1024-aligned)
  float* vec_s = reinterpret_cast<float*>(b2b_w + (epi_has_b2b(EPI) ? p.b2b_bytes : 0));
  uint64_t* bars = reinterpret_cast<uint64_t*>(vec_s + VEC_MAX + p.wg_scratch * NGROUPS);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + nA;
  uint64_t* b_full = a_empty + nA;
  uint64_t* b_empty = b_full + nB;
  uint64_t* acc_full = b_empty + nB;
  uint64_t* acc_empty = acc_full + S;
  uint64_t* b2b_full = acc_empty + S;                 // 1 + NGROUPS barriers of the back-to-back GEMM (unused by the other epilogues)
  uint64_t* b2b_done = b2b_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b2b_done + NGROUPS);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp index, provably warp-uniform
  // warp roles: [0, 4*NGROUPS) epilogue warpgroups (warp % 4 = TMEM lane quadrant), then producer, MMA issuer, TMEM allocator
  constexpr int W_PROD = 4 * NGROUPS, W_MMA = 4 * NGROUPS + 1, W_ALLOC = 4 * NGROUPS + 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < nA; ++i) { mbar_init(smem_u32(a_full + i), 1); mbar_init(smem_u32(a_empty + i), 1); }
    for (int i = 0; i < nB; ++i) { mbar_init(smem_u32(b_full + i), 1); mbar_init(smem_u32(b_empty + i), 1); }
    for (int i = 0; i < S; ++i) { mbar_init(smem_u32(acc_full + i), 1); mbar_init(smem_u32(acc_empty + i), 128 * MT * WGS); }
    if (epi_has_b2b(EPI)) {
      mbar_init(smem_u32(b2b_full), 1);
      for (int i = 0; i < NGROUPS; ++i) mbar_init(smem_u32(b2b_done + i), 1);
    }
    mbar_fence_init();
  }
  if (warp == W_ALLOC) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  for (int i = threadIdx.x; i < p.nvec; i += blockDim.x) vec_s[i] = p.vec[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t vec_addr = smem_u32(vec_s);
  // programmatic dependent launch: everything above (barrier init, TMEM allocation, constant vector) overlapped the tail of the
  // previous kernel in the stream; from here on we read what it wrote.  The next kernel may start its own prologue as soon as
  // this grid's CTAs leave their SMs (its griddepcontrol.wait still waits for this whole grid to finish and flush).
  grid_dependency_wait();
  grid_launch_dependents();

  const int tpi = p.tiles_x * p.tiles_y;
  const int ntiles = p.n_active * tpi;
  // Work items.  Full waves of whole tiles first; when the last, partial wave would leave more than half of the CTAs idle, its
  // tiles are split into single M-tiles (half tiles) so that the tail costs half a tile time instead of a whole one.
  // item w < n_whole: tile w, all M-tiles; else tile n_whole + (w - n_whole) / MT, M-tile (w - n_whole) % MT only.
  const int n_whole = (MT > 1 && 2 * (ntiles % (int)gridDim.x) <= (int)gridDim.x) ? ntiles - ntiles % (int)gridDim.x : ntiles;
  const int nwork = n_whole + (ntiles - n_whole) * MT;
  // An M-tile that lies entirely to the right of the image (W = 200: the second M-tile of the 13th tile column; 50 x 50 latents:
  // of every 4th column) is masked out: no MMAs are issued for it and its epilogue warpgroup only hands the accumulator back.
  auto work_tile = [&](int w, uint32_t& mtmask) -> int {
    int tile;
    if (w < n_whole) {
      mtmask = (1u << MT) - 1u;
      tile = w;
    } else {
      const int h = w - n_whole;
      mtmask = 1u << (h % MT);
      tile = n_whole + h / MT;
    }
    if (MT > 1) {
      const int tx = (tile % tpi) % p.tiles_x;
#pragma unroll
      for (int mt = 1; mt < MT; ++mt)
        if ((tx * MT + mt) * TILE_W >= p.W) mtmask &= ~(1u << mt);
    }
    return tile;
  };
  const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty), b_full0 = smem_u32(b_full), b_empty0 = smem_u32(b_empty);
  const uint32_t a_smem0 = smem_u32(a_base), b_smem0 = smem_u32(b_base);

  if (warp == W_PROD) {
    {
      // ===================== TMA producer =====================
      // The whole warp runs this loop converged and ONE elected lane issues: all operands stay warp-uniform (uniform
      // registers), so no per-instruction R2UR waterfall.  Ring indices are counters: no div/mod per tap.
      if (elect_one()) {
        for (int c = 0; c < p.nchunk; ++c) tma_prefetch_desc(&p.amap[c]);
        tma_prefetch_desc(&p.wmap);
        if (epi_can_be_resident(EPI) && p.resident_b) {               // the stage's own weights: resident for the whole launch
          mbar_expect_tx(b_full0, (uint32_t)p.resident_b);
          for (int q = 0; q * 64 * ROW_BYTES < p.resident_b; ++q) tma_load_2d(b_smem0 + q * 64 * ROW_BYTES, &p.wmap, b_full0, 0, q * 64);
        }
        if (epi_has_b2b(EPI)) {          // the follow-up conv's weights: resident for the whole launch
          mbar_expect_tx(smem_u32(b2b_full), (uint32_t)p.b2b_bytes);
          for (int q = 0; q * 64 * ROW_BYTES < p.b2b_bytes; ++q)
            tma_load_2d(smem_u32(b2b_w) + q * 64 * ROW_BYTES, &p.wmap, smem_u32(b2b_full), 0, p.b2b_wrow + q * 64);
        }
      }
      uint32_t sa = 0, pa = 1, sb = 0, pb = 1;      // slot index and the parity to wait for on the EMPTY barrier
      for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        uint32_t mtmask;
        const int tile = work_tile(w, mtmask);
        const int bi = tile / tpi, rem = tile - bi * tpi;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        // row-paired taps: 15 output rows per tile; lane row 0 is the row above them (it only produces block-1 partial sums)
        const int y0 = p.pair_rows ? ty * (TILE_H - 1) - 1 : ty * TILE_H, x0 = tx * TILE_W * MT;
        const int sid = __shfl_sync(0xffffffffu, p.sample_id[bi], 0), ximg = __shfl_sync(0xffffffffu, p.x_img[bi], 0);
        for (int c = 0; c < p.nchunk; ++c) {
          const ChunkK ck = p.chunk[c];
          const int R = ck.R, pad = (R - 1) >> 1;
          const int img = ck.img_sel ? ximg : sid;
          // the tile + halo of this chunk: one box, all taps read it through shifted descriptors
          mbar_wait(a_empty0 + sa * 8, pa, p.err, 1);
          if (p.debug & 4) {
            if (elect_one()) mbar_arrive(a_full0 + sa * 8);
          } else if (elect_one()) {
            mbar_expect_tx(a_full0 + sa * 8, (uint32_t)a_box_bytes(R, MT));
            tma_load_4d(a_smem0 + sa * p.a_slot_bytes, &p.amap[c], a_full0 + sa * 8, ck.c0, x0 - pad + ck.ox, y0 - pad + ck.oy, img);
          }
          __syncwarp();
          if (++sa == (uint32_t)nA) { sa = 0; pa ^= 1; }
          const int tap_rows = ck.n * ck.nrep;                  // weight rows of one tap
          const int ngrp = (R + ck.tb - 1) / ck.tb;             // B tiles per dx column; the last one may hold fewer taps
          int dx = rem % R, g0 = (ck.tb == 1) ? (rem / R) % R : 0;     // tap rotation (same formula in the MMA warp)
          if (epi_can_be_resident(EPI) && p.resident_b) continue;      // weights already in shared memory
          for (int i = 0; i < R; ++i) {
            int gi = g0;
            for (int j = 0; j < ngrp; ++j) {
              const int ntap = min(ck.tb, R - gi * ck.tb);
              const int grp_rows = tap_rows * ntap;             // rows of this B tile
              mbar_wait(b_empty0 + sb * 8, pb, p.err, 2);
              if (p.debug & 4) {
                if (elect_one()) mbar_arrive(b_full0 + sb * 8);
              } else if (elect_one()) {
                mbar_expect_tx(b_full0 + sb * 8, (uint32_t)grp_rows * ROW_BYTES);
                int row = ck.wrow + (dx * R + gi * ck.tb) * tap_rows + bi * p.w_rows_per_sample;
                uint32_t dst = b_smem0 + sb * p.b_slot_bytes;
                for (int q = 0; q < grp_rows; q += 64, row += 64, dst += 64 * ROW_BYTES)
                  tma_load_2d(dst, &p.wmap, b_full0 + sb * 8, 0, row);
              }
              __syncwarp();
              if (++sb == (uint32_t)nB) { sb = 0; pb ^= 1; }
              if (++gi == ngrp) gi = 0;
            }
            if (++dx == R) dx = 0;
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    {
      // ===================== MMA issuer (converged warp, one elected lane issues) =====================
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, as = 0, pacc = 1;   // parities to wait for on the FULL barriers / acc EMPTY
      const bool resident = epi_can_be_resident(EPI) && p.resident_b != 0;
      if (resident && blockIdx.x < (unsigned)nwork) {              // the one-time weight load
        mbar_wait(b_full0, 0, p.err, 5);
        tc_fence_after();
      }
      for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        uint32_t mtmask;
        const int tile = work_tile(w, mtmask);
        const int rem = tile % tpi;
        mbar_wait(smem_u32(acc_empty + as), pacc, p.err, 3);
        tc_fence_after();
        const uint32_t d_base = tmem_base + as * STAGE_COLS;
        for (int c = 0; c < p.nchunk; ++c) {
          const ChunkK ck = p.chunk[c];
          const int R = ck.R;
          const uint32_t WP = TILE_W * MT + R - 1;                   // pixels per row of the halo box
          // descriptor high word: SBO = WP*128 B between 8-row groups, version 1 (bit 46), SWIZZLE_128B (2 << 61)
          const uint32_t a_hi = ((WP * ROW_BYTES) >> 4) | (1u << 14) | (2u << 29);
          constexpr uint32_t B_HI = 64u | (1u << 14) | (2u << 29);
          const uint32_t idesc = make_idesc_bf16(128, (uint32_t)ck.n);
          const uint32_t d_addr = d_base + ck.col;
          const uint32_t rep_lo = (uint32_t)(ck.n * ROW_BYTES) >> 4;
          const int ngrp = (R + ck.tb - 1) / ck.tb;
          const int g0 = (ck.tb == 1) ? (rem / R) % R : 0;
          uint32_t accumulate = ck.init ? 0u : 1u;
          mbar_wait(a_full0 + sa * 8, pa, p.err, 4);
          const uint32_t a_lo0 = ((a_smem0 + sa * p.a_slot_bytes) & 0x3FFFFu) >> 4;
          int dx = rem % R;
          for (int i = 0; i < R; ++i) {
            int gi = g0;
            for (int j = 0; j < ngrp; ++j) {
              if (!resident) {
                mbar_wait(b_full0 + sb * 8, pb, p.err, 5);
                tc_fence_after();
              }
              // streamed: the ring slot holds this group's rows; resident: the group starts at its row of the packed matrix
              uint32_t b_lo = resident ? ((b_smem0 + (uint32_t)(ck.wrow + (dx * R + gi * ck.tb) * ck.n * ck.nrep) * ROW_BYTES) & 0x3FFFFu) >> 4
                                       : ((b_smem0 + sb * p.b_slot_bytes) & 0x3FFFFu) >> 4;
              // first tap of this group: dy = gi*tb; one pixel row = 128 B = 8 descriptor units
              uint32_t a_lo = a_lo0 + ((uint32_t)(gi * ck.tb) * WP + (uint32_t)dx) * (ROW_BYTES >> 4);
              const int ntap = min(ck.tb, R - gi * ck.tb);
              if (p.debug & 2) {
                if (!resident && elect_one()) umma_commit(b_empty0 + sb * 8);
              } else if (p.pair_rows) {
                // Row-paired taps.  This B tile holds the taps dy = gi*tb .. gi*tb + ntap - 1 of column dx as ONE operand of
                // n * ntap rows per rep, ordered [dy_hi | dy_lo]; a single MMA per K step, on the window of dy_hi, produces the
                // dy_hi term of its own pixel in column block 0 and the dy_lo term of the pixel one row below in block 1 (the
                // epilogue folds block 1 back).  Twice the N per MMA for the 64-channel stages, whose issue cost is ~41 + N/2.
                // (a tile may hold several pairs: tb is even, so the pairs inside it start at even tap offsets)
                if (elect_one()) {
                  uint32_t acc = accumulate;
                  for (int t0 = 0; t0 < ntap; t0 += 2) {
                    const int m = min(2, ntap - t0);
                    const uint32_t a_pair = a_lo + (uint32_t)(t0 + m - 1) * WP * (ROW_BYTES >> 4);
                    const uint32_t idesc_pair = make_idesc_bf16(128, (uint32_t)(ck.n * m));
                    for (int rep = 0; rep < ck.nrep; ++rep, b_lo += rep_lo * m) {
#pragma unroll
                      for (uint32_t k = 0; k < 4; ++k) {
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                          if (MT > 1 && !((mtmask >> mt) & 1u)) continue;
                          umma_bf16(d_addr + mt * SLOT_COLS, ((uint64_t)a_hi << 32) | (a_pair + mt * (TILE_W * ROW_BYTES >> 4) + 2 * k),
                                    ((uint64_t)B_HI << 32) | (b_lo + 2 * k), idesc_pair, acc | (k > 0 ? 1u : 0u));
                        }
                      }
                      acc = 1u;
                    }
                  }
                  if (!resident) umma_commit(b_empty0 + sb * 8);
                }
              } else if (elect_one()) {
                uint32_t acc = accumulate;
                for (int t = 0; t < ntap; ++t, a_lo += WP * (ROW_BYTES >> 4)) {
                  for (int rep = 0; rep < ck.nrep; ++rep, b_lo += rep_lo) {
                    // 4 x (K = 16 bf16 = 32 bytes = 2 descriptor units) per 64-channel chunk; the M-tiles alternate so that
                    // consecutive MMAs hit different accumulators (no back-to-back dependency on one TMEM slot)
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k) {
#pragma unroll
                      for (int mt = 0; mt < MT; ++mt) {
                        if (MT > 1 && !((mtmask >> mt) & 1u)) continue;     // half tile: the other M-tile belongs to another CTA
                        umma_bf16(d_addr + mt * SLOT_COLS, ((uint64_t)a_hi << 32) | (a_lo + mt * (TILE_W * ROW_BYTES >> 4) + 2 * k),
                                  ((uint64_t)B_HI << 32) | (b_lo + 2 * k), idesc, acc | (k > 0 ? 1u : 0u));
                      }
                    }
                    acc = 1u;
                  }
                }
                if (!resident) umma_commit(b_empty0 + sb * 8);        // frees the weight tile once these MMAs retire
              }
              __syncwarp();
              accumulate = 1u;
              if (++sb == (uint32_t)nB) { sb = 0; pb ^= 1; }
              if (++gi == ngrp) gi = 0;
            }
            if (++dx == R) dx = 0;
          }
          if (elect_one()) umma_commit(a_empty0 + sa * 8);
          __syncwarp();
          if (++sa == (uint32_t)nA) { sa = 0; pa ^= 1; }
        }
        if (elect_one()) umma_commit(smem_u32(acc_full + as));
        __syncwarp();
        if (++as == (uint32_t)S) { as = 0; pacc ^= 1; }
      }
    }
  } else if (warp < 4 * NGROUPS) {
    // ===================== epilogue warpgroup g owns accumulator slot (stage g / MT, M-tile g % MT) =====================
    const int g = warp >> 2;
    const int st = g / (MT * WGS), mt = (g / WGS) % MT, sub = g % WGS;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int r = m >> 3, cx = m & 7;
    const uint32_t taddr = tmem_base + (uint32_t)st * STAGE_COLS + (uint32_t)mt * SLOT_COLS + ((uint32_t)(q * 32) << 16);
    uint32_t aph = 0;
    B2BCtx b2b{smem_u32(b2b_w), smem_u32(b2b_full), smem_u32(b2b_done + g), 0u, false, smem_u32(b2b_done + (g ^ (WGS - 1)))};
    // pixel of this thread in tile t (index into per-sample NHWC tensors), or -1 outside the image / past the last tile
    auto pixel_of = [&](int w, PixelCtx* out) -> long long {
      if (w >= nwork) return -1;
      uint32_t mtmask;
      const int t = work_tile(w, mtmask);
      if (!((mtmask >> mt) & 1u)) return -2;      // half tile owned by the other M-tile's warpgroup: nothing to do here
      const int bi = t / tpi, rem = t - bi * tpi;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int sid = p.sample_id[bi];
      const int y = p.pair_rows ? ty * (TILE_H - 1) - 1 + r : ty * TILE_H + r, x = (tx * MT + mt) * TILE_W + cx;
      const bool valid = (y < p.H) && (x < p.W) && !(p.pair_rows && r == 0);
      const size_t pix = ((size_t)sid * p.H + y) * p.W + x;
      if (out) { out->bi = bi; out->sid = sid; out->y = y; out->x = x; out->valid = valid; out->pix = pix; out->wg = g; out->m = m; out->sub = sub; }
      return valid ? (long long)pix : -1;
    };
    for (int w = blockIdx.x + st * gridDim.x; w < nwork; w += S * gridDim.x, aph ^= 1) {
      PixelCtx c;
      const bool mine = pixel_of(w, &c) != -2;
      mbar_wait(smem_u32(acc_full + st), aph, p.err, 6);
      tc_fence_after();
      bool released = false;
      if (mine && !(p.debug & 1)) released = run_epilogue<EPI, X3, CG>(p, vec_addr, taddr, c, smem_u32(acc_empty + st), b2b);
      if (!released) {
        tc_fence_before();
        mbar_arrive(smem_u32(acc_empty + st));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_ALLOC) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace sf
