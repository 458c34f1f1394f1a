// sf_plan.cu -- host side of libsf_b200.so: the plan object (buffer bindings, stage definitions, TMA
// descriptors), kernel launches and the extern "C" ABI declared in include/sf_b200.h.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "sf_conv.cuh"
#include "sf_elementwise.cuh"
#include "sf_peer.cuh"

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define SF_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return fail(SF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                    \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct ActBuf {
  void* hi = nullptr;
  void* lo = nullptr;
  int channels = 0;
  int n_images = 0;
};
struct Stage {
  bool defined = false;
  int epi = 0;
  std::vector<sf_chunk> chunks;
  const void* w = nullptr;
  int w_rows = 0;
  const float* vec = nullptr;
  int n_vec = 0;
  std::vector<int> io, io_off;
  int flags = 0, n_out = 0;
  CUtensorMap wmap;
  int a_slot = 0, b_slot = 0, nA = 0, nB = 0, smem = 0;
  std::vector<int> tb;     // per chunk: taps per B tile (1 or R)
  int b2b_wrow = 0, b2b_bytes = 0;   // fused 1x1 follow-up conv of a lngelu stage (flag 1024): its weights are the last rows of w
  int fixed_smem = 0;                // shared memory outside the operand rings
  int resident_b = 0;                // > 0: bytes of packed weights kept in shared memory for the whole launch instead of streamed per tile
  int mt_epi = 0;                    // epilogue id that decides the CTA tile shape (the fused pointwise pair, flag 8192, runs one M-tile)
  // SE layer folded into this stage's weights (sf_plan_define_stage_fold)
  int fold_se = -1;
  const float* w32 = nullptr;
  const int32_t* row_meta = nullptr;
  void* w_scaled = nullptr;
};
struct SeDef {
  bool defined = false;
  const float* fc1 = nullptr;
  const float* fc2 = nullptr;
  int in_buf = -1, out_buf = -1;
};

constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int BAR_AREA = 512;
// weight tile (one ring slot) size cap: a whole dx column of taps travels as one tile when it fits (SF_B_TILE_MAX overrides, KB)
const int B_TILE_MAX = [] { const char* v = getenv("SF_B_TILE_MAX"); return (v ? atoi(v) : 48) * 1024; }();

}  // namespace

struct sf_plan {
  sf_geometry g;
  int num_sms = 148;
  bool finalized = false;
  ActBuf act[SF_MAX_ACT_BUFS];
  void* f32[SF_F32_COUNT + 3] = {};
  Stage stage[SF_MAX_STAGES];
  SeDef se[2];
  std::vector<int> cell[2], prior;
  std::map<std::tuple<int, int, int, int>, CUtensorMap> amaps;   // (buf, plane, R, MT) -> activation tensor map
  int last_launches = 0;
};

namespace {

using namespace sf;

int encode_act_map(sf_plan* p, int buf, int plane, int R, int MT, CUtensorMap* out) {
  auto key = std::make_tuple(buf, plane, R, MT);
  auto it = p->amaps.find(key);
  if (it != p->amaps.end()) {
    *out = it->second;
    return SF_OK;
  }
  if (buf < 0 || buf >= SF_MAX_ACT_BUFS) return fail(SF_ERR_INVALID, "activation buffer id out of range");
  const ActBuf& a = p->act[buf];
  void* base = plane ? a.lo : a.hi;
  if (!base) return fail(SF_ERR_STATE, "activation buffer " + std::to_string(buf) + " plane " + std::to_string(plane) + " not bound");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(SF_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  const cuuint64_t C = a.channels, W = p->g.W, H = p->g.H, N = a.n_images;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)(TILE_W * MT + R - 1), (cuuint32_t)(TILE_H + R - 1), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SF_ERR_CUDA, "cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r));
  p->amaps[key] = m;
  *out = m;
  return SF_OK;
}

int encode_weight_map(const void* w, int rows, CUtensorMap* out) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(SF_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)KC, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)KC * 2};
  cuuint32_t box[2] = {(cuuint32_t)KC, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SF_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
  return SF_OK;
}

typedef void (*StageKernel)(const StageParams);
template <bool X3, int CG>
StageKernel kernel_for(int epi) {
  switch (epi) {
    case SF_EPI_GATES: return conv_stage_kernel<SF_EPI_GATES, X3, CG>;
    case SF_EPI_PROPOSE: return conv_stage_kernel<SF_EPI_PROPOSE, X3, CG>;
    case SF_EPI_DECODE: return conv_stage_kernel<SF_EPI_DECODE, X3, CG>;
    case SF_EPI_LNGELU: return conv_stage_kernel<SF_EPI_LNGELU, X3, CG>;
    case SF_EPI_MIX: return conv_stage_kernel<SF_EPI_MIX, X3, CG>;
    case SF_EPI_BIAS_LRELU: return conv_stage_kernel<SF_EPI_BIAS_LRELU, X3, CG>;
    case SF_EPI_RES_PROJ: return conv_stage_kernel<SF_EPI_RES_PROJ, X3, CG>;
    case SF_EPI_RES_ID: return conv_stage_kernel<SF_EPI_RES_ID, X3, CG>;
    case SF_EPI_SAMPLE: return conv_stage_kernel<SF_EPI_SAMPLE, X3, CG>;
    case SF_EPI_BIAS_ACT: return conv_stage_kernel<SF_EPI_BIAS_ACT, X3, CG>;
    case SF_EPI_RES_ID_ACT: return conv_stage_kernel<SF_EPI_RES_ID_ACT, X3, CG>;
    case SF_EPI_LNGELU_B2B:
      if constexpr (CG == 64) return conv_stage_kernel<SF_EPI_LNGELU_B2B, X3, CG>;
      else return nullptr;
    case SF_EPI_PW_B2B:
      if constexpr (CG == 64 && !X3) return conv_stage_kernel<SF_EPI_PW_B2B, X3, CG>;
      else return nullptr;
  }
  return nullptr;
}
StageKernel kernel_for(int epi, bool x3, int C) {
  if (C == 128) return x3 ? kernel_for<true, 128>(epi) : kernel_for<false, 128>(epi);
  return x3 ? kernel_for<true, 64>(epi) : kernel_for<false, 64>(epi);
}

// SF_PDL=0 in the environment turns programmatic dependent launch off (A/B measurements)
bool pdl_enabled() {
  static const bool on = [] { const char* v = getenv("SF_PDL"); return !(v && v[0] == '0'); }();
  return on;
}

int state_act_buf(int which) { return which; }   // activation buffers 0 and 1 mirror fp32 state buffers 0 and 1

int resolve_buf(const sf_event* ev, int buf) {
  if (buf == -1) return ev->x_buf;
  if (buf == -2) return state_act_buf(ev->s_in);
  if (buf == -3) return state_act_buf(ev->s_out);
  return buf;
}

float* se_scale_ptr(sf_plan* p, int which);

int launch_stage(sf_plan* p, int sidx, const sf_event* ev, const int32_t* table, cudaStream_t stream) {
  if (sidx < 0 || sidx >= SF_MAX_STAGES || !p->stage[sidx].defined) return fail(SF_ERR_STATE, "stage not defined");
  if (ev->n_active <= 0) return SF_OK;
  Stage& st = p->stage[sidx];
  const bool x3 = p->g.precision == SF_PREC_BF16X3;
  StageParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.nchunk = (int)st.chunks.size();
  for (int c = 0; c < sp.nchunk; ++c) {
    const sf_chunk& ck = st.chunks[c];
    const int buf = resolve_buf(ev, ck.buf);
    int rc = encode_act_map(p, buf, ck.plane, ck.R, sf::mtiles_for(st.mt_epi, p->g.C), &sp.amap[c]);
    if (rc) return rc;
    if (ck.c0 + KC > p->act[buf].channels) return fail(SF_ERR_INVALID, "chunk channel range exceeds buffer");
    sp.chunk[c] = ChunkK{ck.R, ck.n, ck.nrep, ck.col, ck.wrow, ck.init, ck.c0, ck.buf == -1 ? 1 : 0, st.tb[c], ck.ox, ck.oy};
  }
  sp.wmap = st.wmap;
  sp.H = p->g.H;
  sp.W = p->g.W;
  const int MT = sf::mtiles_for(st.mt_epi, p->g.C);
  sp.tiles_x = (p->g.W + TILE_W * MT - 1) / (TILE_W * MT);
  const bool pair = (st.flags & 512) != 0;
  const int tile_rows = pair ? TILE_H - 1 : TILE_H;          // row-paired taps: lane row 0 of a tile is the row above its 15 output rows
  sp.tiles_y = (p->g.H + tile_rows - 1) / tile_rows;
  sp.pair_rows = pair ? 1 : 0;
  sp.wg_scratch = pair ? sf::WG_SCRATCH_PAIR : sf::WG_SCRATCH;
  sp.n_active = ev->n_active;
  const int n = ev->n_active;
  const int32_t* rows = table + ev->table_off;
  sp.sample_id = rows;
  sp.x_img = rows + n;
  sp.rec_slot = rows + 2 * n;
  sp.eps_slot = rows + 3 * n;
  sp.dt = reinterpret_cast<const float*>(rows + 4 * n);
  sp.vec = st.vec;
  sp.nvec = st.n_vec;
  sp.a_slot_bytes = st.a_slot;
  sp.b_slot_bytes = st.b_slot;
  sp.nA = st.nA;
  sp.nB = st.nB;
  sp.w_rows_per_sample = st.fold_se >= 0 ? st.w_rows : 0;
  sp.resident_b = st.resident_b;
  sp.b2b_wrow = st.b2b_wrow;
  sp.b2b_bytes = st.b2b_bytes;
  static const int debug_stage = [] { const char* v = getenv("SF_DEBUG_STAGE"); return v ? atoi(v) : 0; }();    // read once, not per launch
  sp.debug = debug_stage;
  sp.err = reinterpret_cast<int*>(p->f32[SF_F32_COUNT]);
  EpiArgs& e = sp.e;
  e.kind = ev->kind;
  e.s_in = reinterpret_cast<const float*>(p->f32[SF_F32_STATE0 + ev->s_in]);
  e.s_base = reinterpret_cast<const float*>(p->f32[SF_F32_STATE0 + ev->s_base]);
  e.s_out = reinterpret_cast<float*>(p->f32[SF_F32_STATE0 + ev->s_out]);
  e.a32 = reinterpret_cast<float*>(p->f32[SF_F32_A]);
  e.b32 = reinterpret_cast<float*>(p->f32[SF_F32_B]);
  e.path = reinterpret_cast<float*>(p->f32[SF_F32_PATH]);
  e.eps = reinterpret_cast<const float*>(p->f32[SF_F32_EPS]);
  e.x32 = ev->want_f32 ? reinterpret_cast<float*>(p->f32[SF_F32_X]) : nullptr;
  e.params32 = ev->want_f32 ? reinterpret_cast<float*>(p->f32[SF_F32_PARAMS]) : nullptr;
  auto out = [&](int slot, int k) {
    const int buf = st.io[k];
    e.out_h[slot] = reinterpret_cast<__nv_bfloat16*>(p->act[buf].hi);
    e.out_l[slot] = reinterpret_cast<__nv_bfloat16*>(p->act[buf].lo);
    e.out_cs[slot] = p->act[buf].channels;
    e.out_co[slot] = st.io_off[k];
  };
  auto in = [&](int slot, int k) {
    const int buf = st.io[k];
    e.in_h[slot] = reinterpret_cast<const __nv_bfloat16*>(p->act[buf].hi);
    e.in_l[slot] = reinterpret_cast<const __nv_bfloat16*>(p->act[buf].lo);
    e.in_cs[slot] = p->act[buf].channels;
    e.in_co[slot] = st.io_off[k];
  };
  auto need_io = [&](size_t k) { return st.io.size() >= k; };
  for (int b : st.io)
    if (b < 0 || b >= SF_MAX_ACT_BUFS || !p->act[b].hi || (x3 && !p->act[b].lo))
      return fail(SF_ERR_STATE, "stage io buffer not bound");
  const int pairs = (st.flags & 32) ? 1 : 128 / p->g.C;          // gate pairs / proposals handled by one launch
  e.pairs = pairs;
  switch (st.epi) {
    case SF_EPI_GATES:                      // io = [u_0, gated_0, (u_1, gated_1)]
      if (!need_io(2 * pairs)) return fail(SF_ERR_INVALID, "gates stage needs 2 io buffers per gate pair");
      for (int i = 0; i < 2 * pairs; ++i) out(i, i);
      break;
    case SF_EPI_PROPOSE:                    // io = [u_0, (u_1,) out_0, (out_1)]
      if (!need_io(2 * pairs)) return fail(SF_ERR_INVALID, "propose stage needs 2 io buffers per proposal");
      for (int i = 0; i < pairs; ++i) { in(i, i); out(i, pairs + i); }
      if (!(st.flags & 1)) e.a32 = nullptr;   // only the first GRU's blend is kept in fp32
      break;
    case SF_EPI_MIX: {
      const int buf = state_act_buf(ev->s_out);
      e.out_h[0] = reinterpret_cast<__nv_bfloat16*>(p->act[buf].hi);
      e.out_l[0] = reinterpret_cast<__nv_bfloat16*>(p->act[buf].lo);
      break;
    }
    case SF_EPI_RES_ID:
      if (!need_io(2)) return fail(SF_ERR_INVALID, "residual stage needs 2 io buffers");
      in(0, 0); out(0, 1);
      e.n_out = st.n_out;
      break;
    default:
      if (!need_io(1)) return fail(SF_ERR_INVALID, "stage needs an output buffer");
      out(0, 0);
      e.n_out = (st.epi == SF_EPI_BIAS_LRELU || st.epi == SF_EPI_RES_PROJ) ? st.n_out : p->act[st.io[0]].channels;
      break;
  }
  e.act = (st.flags >> 1) & 7;                                        // bias_act activation code
  e.act_after_res = (st.flags & 2048) ? 1 : 0;
  e.deriv = (st.flags & 4096) ? 1 : 0;
  e.out32 = (st.flags & 16) ? reinterpret_cast<float*>(p->f32[SF_F32_OUT]) : nullptr;
  e.img_bias = (st.flags & 64) ? reinterpret_cast<const float*>(p->f32[SF_F32_IMG_BIAS]) : nullptr;
  e.res_scale = nullptr;
  if (st.flags & 128) {
    if (!p->f32[SF_F32_SE_SUMS]) return fail(SF_ERR_STATE, "SE scratch not bound");
    e.res_scale = se_scale_ptr(p, (st.flags >> 8) & 1);
    e.res_scale_ch = 2 * p->g.C;
  }
  if ((st.flags & 64) && !e.img_bias) return fail(SF_ERR_STATE, "per-image bias requested but SF_F32_IMG_BIAS is not bound");
  if ((st.flags & 16) && !e.out32) return fail(SF_ERR_STATE, "fp32 output requested but SF_F32_OUT is not bound");
  if (st.epi == SF_EPI_MIX || st.epi == SF_EPI_GATES || st.epi == SF_EPI_PROPOSE)
    if (!e.s_in || !e.s_out || !e.s_base) return fail(SF_ERR_STATE, "state buffers not bound");
  if (st.epi == SF_EPI_SAMPLE && !e.eps) return fail(SF_ERR_STATE, "eps buffer not bound");
  const int ntiles = n * sp.tiles_x * sp.tiles_y;
  const int grid = ntiles < p->num_sms ? ntiles : p->num_sms;
  // the ODE loop's LeakyReLU-only stages run the lean instantiations; any other activation, an fp32 copy or a per-image bias
  // selects the general variant of the same epilogue
  int kepi = st.epi;
  const bool general = e.act != 0 || e.out32 || e.img_bias || e.act_after_res;
  if (kepi == SF_EPI_BIAS_LRELU && general) kepi = SF_EPI_BIAS_ACT;
  if (kepi == SF_EPI_RES_ID && (st.flags & 8192)) kepi = SF_EPI_PW_B2B;
  if (kepi == SF_EPI_RES_ID && general) kepi = SF_EPI_RES_ID_ACT;
  if (kepi == SF_EPI_LNGELU && st.b2b_bytes) kepi = SF_EPI_LNGELU_B2B;
  StageKernel k = kernel_for(kepi, x3, p->g.C);
  if (!k) return fail(SF_ERR_INVALID, "unknown epilogue");
  void* args[] = {&sp};
  // programmatic dependent launch: this kernel's prologue (barrier init, TMEM allocation, constant vector) may overlap the
  // tail of the previous kernel in the stream; it executes griddepcontrol.wait before touching anything that kernel wrote
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128 + 128 * sf::ACC_STAGES * MT * sf::wgs_per_slot(kepi));
  cfg.dynamicSmemBytes = (size_t)st.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  SF_CUDA(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(k), args));
  p->last_launches += 1;
  return SF_OK;
}

constexpr int SE_MAX_PARTIALS = SF_SE_MAX_PARTIALS;
// layout of the caller's SF_F32_SE_SUMS buffer (floats): partial sums [2][max_images][SF_SE_MAX_PARTIALS][2C] | scales [2][max_images][2C] |
// block counters [2][max_images] (uint32, must start zeroed)
float* se_scale_ptr(sf_plan* p, int which) {
  const size_t CH = 2 * p->g.C, B = p->g.max_images;
  return reinterpret_cast<float*>(p->f32[SF_F32_SE_SUMS]) + 2 * B * SE_MAX_PARTIALS * CH + (size_t)which * B * CH;
}
unsigned int* se_counter_ptr(sf_plan* p, int which) {
  const size_t CH = 2 * p->g.C, B = p->g.max_images;
  return reinterpret_cast<unsigned int*>(reinterpret_cast<float*>(p->f32[SF_F32_SE_SUMS]) + 2 * B * SE_MAX_PARTIALS * CH + 2 * B * CH) + (size_t)which * B;
}

// SE step 1 over the pixel window [px0, px1) of every active sample; returns the number of per-block partials (> 0) or < 0
int launch_se_reduce(sf_plan* p, int which, const sf_event* ev, const int32_t* table, int px0, int px1, bool fused_scale, cudaStream_t stream,
                     bool band_totals = false) {
  const SeDef& se = p->se[which];
  if (!se.defined) return fail(SF_ERR_STATE, "SE layer not defined");
  const bool x3 = p->g.precision == SF_PREC_BF16X3;
  const int CH = 2 * p->g.C;
  const int hw = p->g.H * p->g.W;
  float* sums = reinterpret_cast<float*>(p->f32[SF_F32_SE_SUMS]);
  if (!sums) return fail(SF_ERR_STATE, "SE sums buffer not bound");
  sums += (size_t)which * p->g.max_images * SE_MAX_PARTIALS * CH;
  const ActBuf& zi = p->act[se.in_buf];
  if (!zi.hi) return fail(SF_ERR_STATE, "SE buffers not bound");
  if (px0 < 0 || px1 > hw || px0 >= px1) return fail(SF_ERR_INVALID, "bad SE pixel window");
  const int* sid = table + ev->table_off;
  // one balanced wave: 4 resident 256-thread blocks per SM over all samples (8 samples x 74 blocks = 592 = 4 x 148), at least
  // ~1 iteration of 4 x 32 pixels per block
  int bpi = (4 * p->num_sms + ev->n_active - 1) / ev->n_active;
  const int cap = (px1 - px0 + 32 * 4 - 1) / (32 * 4);
  if (bpi > cap) bpi = cap;
  if (bpi > SE_MAX_PARTIALS) bpi = SE_MAX_PARTIALS;
  if (bpi < 1) bpi = 1;
  dim3 grid(bpi, ev->n_active);
  auto zh = reinterpret_cast<const __nv_bfloat16*>(zi.hi);
  auto zl = reinterpret_cast<const __nv_bfloat16*>(zi.lo);
  // scratch behind the partial sums of both layers: [2][max_images][CH] scales, then [2][max_images] block counters (zeroed once)
  float* scale = se_scale_ptr(p, which);
  unsigned int* counters = se_counter_ptr(p, which);
  float* scale_arg = fused_scale ? scale : nullptr;
  float* totals = band_totals ? scale : nullptr;          // band totals land in the scale slot [n_active][2C] until sf_plan_se_finish
  const float inv_n = 1.0f / (float)hw;
  if (CH == 256) {
    if (x3) se_reduce_kernel<256, true><<<grid, 256, 0, stream>>>(zh, zl, sums, sid, hw, px0, px1, counters, scale_arg, se.fc1, se.fc2, inv_n, totals);
    else se_reduce_kernel<256, false><<<grid, 256, 0, stream>>>(zh, zl, sums, sid, hw, px0, px1, counters, scale_arg, se.fc1, se.fc2, inv_n, totals);
  } else {
    if (x3) se_reduce_kernel<128, true><<<grid, 256, 0, stream>>>(zh, zl, sums, sid, hw, px0, px1, counters, scale_arg, se.fc1, se.fc2, inv_n, totals);
    else se_reduce_kernel<128, false><<<grid, 256, 0, stream>>>(zh, zl, sums, sid, hw, px0, px1, counters, scale_arg, se.fc1, se.fc2, inv_n, totals);
  }
  SF_CUDA(cudaGetLastError());
  p->last_launches += 1;
  return bpi;
}

// SE step 2: mean = (sum of n_partials partial sums) * inv_n; y = z * sigmoid(fc2 relu(fc1 mean))
int launch_se_apply(sf_plan* p, int which, const sf_event* ev, const int32_t* table, int n_partials, float inv_n, cudaStream_t stream) {
  const SeDef& se = p->se[which];
  if (!se.defined) return fail(SF_ERR_STATE, "SE layer not defined");
  const bool x3 = p->g.precision == SF_PREC_BF16X3;
  const int CH = 2 * p->g.C;
  const int hw = p->g.H * p->g.W;
  float* sums = reinterpret_cast<float*>(p->f32[SF_F32_SE_SUMS]);
  if (!sums) return fail(SF_ERR_STATE, "SE sums buffer not bound");
  sums += (size_t)which * p->g.max_images * SE_MAX_PARTIALS * CH;
  const ActBuf& zi = p->act[se.in_buf];
  const ActBuf& yo = p->act[se.out_buf];
  if (!zi.hi || !yo.hi) return fail(SF_ERR_STATE, "SE buffers not bound");
  if (n_partials < 0 || n_partials > SE_MAX_PARTIALS) return fail(SF_ERR_INVALID, "bad SE partial count");
  const int* sid = table + ev->table_off;
  auto zh = reinterpret_cast<const __nv_bfloat16*>(zi.hi);
  auto zl = reinterpret_cast<const __nv_bfloat16*>(zi.lo);
  auto yh = reinterpret_cast<__nv_bfloat16*>(yo.hi);
  auto yl = reinterpret_cast<__nv_bfloat16*>(yo.lo);
  int bpa = (hw + 32 * 4 - 1) / (32 * 4);
  const int cap = (4 * p->num_sms + ev->n_active - 1) / ev->n_active;      // ~4 resident blocks per SM over all samples
  if (bpa > cap) bpa = cap;
  if (bpa < 1) bpa = 1;
  dim3 grid2(bpa, ev->n_active);
  float* scale = se_scale_ptr(p, which);
  if (n_partials > 0) {     // scales not yet computed by the reduce kernel (caller reduced the sums across GPUs)
    if (CH == 256) se_scale_kernel<256><<<ev->n_active, 256, 0, stream>>>(sums, n_partials, inv_n, se.fc1, se.fc2, scale);
    else se_scale_kernel<128><<<ev->n_active, 256, 0, stream>>>(sums, n_partials, inv_n, se.fc1, se.fc2, scale);
    p->last_launches += 1;
  }
  if (CH == 256) {
    if (x3) se_apply_kernel<256, true><<<grid2, 256, 0, stream>>>(zh, zl, yh, yl, scale, sid, hw);
    else se_apply_kernel<256, false><<<grid2, 256, 0, stream>>>(zh, zl, yh, yl, scale, sid, hw);
  } else {
    if (x3) se_apply_kernel<128, true><<<grid2, 256, 0, stream>>>(zh, zl, yh, yl, scale, sid, hw);
    else se_apply_kernel<128, false><<<grid2, 256, 0, stream>>>(zh, zl, yh, yl, scale, sid, hw);
  }
  SF_CUDA(cudaGetLastError());
  p->last_launches += 1;
  return SF_OK;
}

// SE layer folded into its consumers: reduce + scales (last block of the reduce), then the scaled per-sample weights of every
// stage registered for this layer.  No pass over the activation tensor.
int launch_se_fold(sf_plan* p, int which, const sf_event* ev, const int32_t* table, cudaStream_t stream) {
  if (ev->n_active <= 0) return SF_OK;
  const int hw = p->g.H * p->g.W;
  int n = launch_se_reduce(p, which, ev, table, 0, hw, true, stream);
  if (n < 0) return n;
  for (Stage& st : p->stage) {
    if (!st.defined || st.fold_se != which) continue;
    dim3 grid((st.w_rows * 8 + 255) / 256, ev->n_active);
    se_fold_kernel<<<grid, 256, 0, stream>>>(st.w32, st.row_meta, se_scale_ptr(p, which), reinterpret_cast<__nv_bfloat16*>(st.w_scaled),
                                             st.w_rows, 2 * p->g.C);
    SF_CUDA(cudaGetLastError());
    p->last_launches += 1;
  }
  return SF_OK;
}

int launch_se(sf_plan* p, int which, const sf_event* ev, const int32_t* table, cudaStream_t stream) {
  if (ev->n_active <= 0) return SF_OK;
  const int hw = p->g.H * p->g.W;
  int n = launch_se_reduce(p, which, ev, table, 0, hw, true, stream);
  if (n < 0) return n;
  return launch_se_apply(p, which, ev, table, 0, 1.0f / (float)hw, stream);
}

// Shared-memory layout of a stage's operand rings.
// Weight tiles: as many taps of a dx column per tile as the cap allows (fewer producer <-> issuer barrier round trips per MMA;
// measured: 48 KB tiles are worth 4 % of the rollout over 24 KB ones), shrunk until at least two weight slots fit next to
// the activation ring.  Activation ring: one slot = one chunk's tile + halo, 3 slots when they leave room, up to 4.
// Resident weights: when the whole packed matrix fits next to two activation slots it is loaded ONCE per CTA and stays for the
// launch (no per-tile weight stream from L2, no weight-ring handshakes): the dilated ASPP branches, the 64 -> 64 3x3 stages, mix,
// q1, q2.  Not for per-sample weights (SE layer folded in: the rows change with the tile's sample).  SF_RESIDENT_B=0 turns it off.
int layout_operand_rings(sf_plan* p, Stage& st, bool allow_resident) {
  static const bool resident_on = [] { const char* v = getenv("SF_RESIDENT_B"); return !(v && v[0] == '0'); }();
  const int fixed = st.fixed_smem, flags = st.flags;
  int a_slot = 0, b_slot = 0, nA = 0, nB = 0;
  for (int cap = B_TILE_MAX;; cap -= 8 * 1024) {
    st.tb.clear();
    a_slot = b_slot = 0;
    for (const sf_chunk& c : st.chunks) {
      const int tap_bytes = c.n * c.nrep * ROW_BYTES;
      int tb = cap / tap_bytes;
      if (tb < 1) tb = 1;
      if (tb > c.R) tb = c.R;
      if (flags & 512) tb = tb < 2 ? 2 : (tb & ~1);          // row-paired taps: a B tile = whole pairs of vertically adjacent taps
      st.tb.push_back(tb);
      const int a = (sf::a_box_bytes(c.R, sf::mtiles_for(st.mt_epi, p->g.C)) + 1023) & ~1023, b = tb * tap_bytes;
      a_slot = a > a_slot ? a : a_slot;
      b_slot = b > b_slot ? b : b_slot;
    }
    const int res_bytes = st.b2b_wrow * ROW_BYTES;           // every row of the chunks (a fused follow-up conv's rows have their own region)
    const int res_slot = (res_bytes + 1023) & ~1023;
    const int kepi_static = (st.epi == SF_EPI_LNGELU && st.b2b_bytes) ? sf::SF_EPI_LNGELU_B2B : st.epi;      // (the launch-time variants of bias_act / res_id can all be resident)
    if (allow_resident && resident_on && sf::epi_can_be_resident(kepi_static) && cap == B_TILE_MAX && res_bytes > 0 && fixed + res_slot + 2 * a_slot <= SMEM_BUDGET) {
      nA = (SMEM_BUDGET - fixed - res_slot) / a_slot;
      if (nA > 4) nA = 4;
      st.a_slot = a_slot; st.b_slot = res_slot; st.nA = nA; st.nB = 1;
      st.resident_b = res_bytes;
      st.smem = fixed + nA * a_slot + st.b_slot;
      return SF_OK;
    }
    nA = (fixed + 3 * a_slot + 2 * b_slot <= SMEM_BUDGET) ? 3 : 2;
    nB = (SMEM_BUDGET - fixed - nA * a_slot) / b_slot;
    if (nB > sf::MAX_RING) nB = sf::MAX_RING;
    if (nB >= 2) break;
    if (cap <= 8 * 1024) return fail(SF_ERR_INVALID, "stage does not fit in shared memory");
  }
  while (nA < 4 && fixed + (nA + 1) * a_slot + nB * b_slot <= SMEM_BUDGET) ++nA;
  st.a_slot = a_slot; st.b_slot = b_slot; st.nA = nA; st.nB = nB;
  st.resident_b = 0;
  st.smem = fixed + nA * a_slot + nB * b_slot;
  return SF_OK;
}

int run_item(sf_plan* p, int item, const sf_event* ev, const int32_t* table, cudaStream_t stream) {
  if (item >= 2000) return launch_se_fold(p, item - 2000, ev, table, stream);
  if (item >= 1000) return launch_se(p, item - 1000, ev, table, stream);
  return launch_stage(p, item, ev, table, stream);
}

}  // namespace

// ================================================================================================
// extern "C" ABI
// ================================================================================================
extern "C" {

int sf_abi_version(void) { return SF_ABI_VERSION; }
const char* sf_last_error(void) { return g_err.c_str(); }
// the other translation units of the library (sf_ode.cu) report their failures through the same thread-local string
void sf_internal_set_error(const char* msg) { g_err = msg ? msg : ""; }

int sf_device_supported(int device) {
  cudaDeviceProp prop;
  SF_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(SF_ERR_UNSUPPORTED, "libsf_b200 needs an sm_100 (Blackwell B200) device, found sm_" +
                                                            std::to_string(prop.major) + std::to_string(prop.minor));
  return SF_OK;
}

int sf_plan_create(const sf_geometry* g, sf_plan** out) {
  if (!g || !out) return fail(SF_ERR_INVALID, "null argument");
  if (g->C != 64 && g->C != 128) return fail(SF_ERR_INVALID, "hidden channels must be 64 or 128 (got " + std::to_string(g->C) + ")");
  if (g->H <= 0 || g->W <= 0 || g->max_images <= 0) return fail(SF_ERR_INVALID, "bad geometry");
  if (g->precision != SF_PREC_BF16 && g->precision != SF_PREC_BF16X3) return fail(SF_ERR_INVALID, "bad precision mode");
  int rc = sf_device_supported(g->device);
  if (rc) return rc;
  SF_CUDA(cudaSetDevice(g->device));
  sf_plan* p = new sf_plan();
  p->g = *g;
  cudaDeviceProp prop;
  SF_CUDA(cudaGetDeviceProperties(&prop, g->device));
  p->num_sms = prop.multiProcessorCount;
  *out = p;
  return SF_OK;
}

int sf_plan_destroy(sf_plan* p) {
  delete p;
  return SF_OK;
}

int sf_plan_bind_act(sf_plan* p, int buf, void* hi, void* lo, int channels, int n_images) {
  if (!p || buf < 0 || buf >= SF_MAX_ACT_BUFS) return fail(SF_ERR_INVALID, "bad activation buffer id");
  if (channels % KC != 0 || channels <= 0) return fail(SF_ERR_INVALID, "activation channels must be a multiple of 64");
  if ((reinterpret_cast<uintptr_t>(hi) & 15) || (reinterpret_cast<uintptr_t>(lo) & 15)) return fail(SF_ERR_INVALID, "activation buffers must be 16-byte aligned");
  p->act[buf] = ActBuf{hi, lo, channels, n_images};
  for (auto it = p->amaps.begin(); it != p->amaps.end();)
    it = (std::get<0>(it->first) == buf) ? p->amaps.erase(it) : std::next(it);
  return SF_OK;
}

int sf_plan_bind_f32(sf_plan* p, int slot, void* ptr) {
  if (!p || slot < 0 || slot > SF_F32_COUNT + 2) return fail(SF_ERR_INVALID, "bad fp32 slot");   // + error flag, SF_F32_OUT, SF_F32_IMG_BIAS
  p->f32[slot] = ptr;
  return SF_OK;
}

int sf_plan_define_stage(sf_plan* p, int stage, int epilogue, int n_chunks, const sf_chunk* chunks, const void* w_packed,
                         int w_rows, const float* vec, int n_vec, const int32_t* io_bufs, const int32_t* io_choff, int n_io, int flags) {
  if (!p || stage < 0 || stage >= SF_MAX_STAGES) return fail(SF_ERR_INVALID, "bad stage id");
  if (n_chunks <= 0 || n_chunks > SF_MAX_CHUNKS) return fail(SF_ERR_INVALID, "bad chunk count");
  if (n_vec > sf::VEC_MAX) return fail(SF_ERR_INVALID, "stage vector too long");
  Stage& st = p->stage[stage];
  st = Stage();
  st.epi = epilogue;
  // flag 8192 (res_id, C = 64, bf16): the stage's conv is the ConvNeXt block's pwconv1 (ONE 1x1 chunk, n = 256); GELU and pwconv2 run
  // in the epilogue as a back-to-back GEMM.  w_packed = [pwconv1 rows | pwconv2 as four K-chunks of [64 n][64 k] rows], vec = [b1 (256), b2 (64)]
  const bool pw = epilogue == SF_EPI_RES_ID && (flags & 8192);
  st.mt_epi = pw ? sf::SF_EPI_PW_B2B : epilogue;
  if (pw && (p->g.C != 64 || p->g.precision != SF_PREC_BF16 || n_chunks != 1 || chunks[0].R != 1 || chunks[0].n != 256 || chunks[0].col != 0 ||
             chunks[0].nrep != 1 || n_vec != 320 || w_rows < 512))
    return fail(SF_ERR_INVALID, "the fused pointwise pair needs 64 channels, bf16 operands, one 1x1 chunk with n = 256, a [b1 (256), b2 (64)] vector and pwconv2's 256 rows appended");
  st.chunks.assign(chunks, chunks + n_chunks);
  for (const sf_chunk& c : st.chunks) {
    if (!(c.R == 1 || c.R == 3 || c.R == 7)) return fail(SF_ERR_INVALID, "filter size must be 1, 3 or 7");
    if (c.n % 64 || c.n <= 0 || c.n > 256 || c.col < 0 || c.col + c.n > sf::TMEM_COLS / sf::ACC_STAGES / sf::mtiles_for(st.mt_epi, p->g.C)) return fail(SF_ERR_INVALID, "bad chunk N / column range");
    if (c.nrep < 1 || c.nrep > 2) return fail(SF_ERR_INVALID, "nrep must be 1 or 2");
    if (c.wrow < 0 || c.wrow + c.R * c.R * c.nrep * c.n > w_rows) return fail(SF_ERR_INVALID, "chunk weight rows exceed the packed matrix");
    if ((flags & 512) && (!sf::epi_can_pair(epilogue) || (flags & (128 | 8192)) || p->g.C != 64 ||
                          c.n != 64 || c.col != 0 || c.R < 2 || c.ox || c.oy))
      return fail(SF_ERR_INVALID, "row-paired taps need an epilogue with ONE 64-column accumulator block (lngelu, decode, bias_act, res_id), a 64-channel "
                                  "plan and undilated n = 64 chunks at column 0");
  }
  if (epilogue < 0 || epilogue > SF_EPI_SAMPLE) return fail(SF_ERR_INVALID, "unknown epilogue");
  // flag 1024: a 1x1 convolution + LayerNorm + GELU fused behind a lngelu stage (back-to-back GEMM in the epilogue); its
  // [C x C] weights (hi, then lo in the split mode) are the last rows of the packed matrix and stay in shared memory
  const bool b2b = (flags & 1024) != 0;
  const int b2b_rows = b2b ? p->g.C * (p->g.precision == SF_PREC_BF16X3 ? 2 : 1) : pw ? 256 : 0;
  if (b2b && (epilogue != SF_EPI_LNGELU || p->g.C != 64 || n_vec != 4 * p->g.C || w_rows < b2b_rows))
    return fail(SF_ERR_INVALID, "the fused 1x1 follow-up needs the lngelu epilogue, 64 channels, a [LN1 w, LN1 b, LN2 w, LN2 b] vector and its weights appended");
  for (const sf_chunk& c : st.chunks)
    if ((b2b || pw) && c.wrow + c.R * c.R * c.nrep * c.n > w_rows - b2b_rows) return fail(SF_ERR_INVALID, "chunk weight rows overlap the follow-up conv's rows");
  const int fixed = 1024 + (sf::VEC_MAX + 4 * ((flags & 512) ? sf::WG_SCRATCH_PAIR : sf::WG_SCRATCH)) * 4 + BAR_AREA + b2b_rows * ROW_BYTES;
  st.flags = flags;
  st.w_rows = w_rows;
  st.fixed_smem = fixed;
  st.b2b_bytes = b2b_rows * ROW_BYTES;
  st.b2b_wrow = w_rows - b2b_rows;
  if (int rc = layout_operand_rings(p, st, true)) return rc;
  st.w = w_packed;
  st.w_rows = w_rows;
  st.vec = vec;
  st.n_vec = n_vec;
  st.io.assign(io_bufs, io_bufs + n_io);
  if (io_choff) st.io_off.assign(io_choff, io_choff + n_io); else st.io_off.assign(n_io, 0);
  st.n_out = 0;
  for (const sf_chunk& c : st.chunks) if (c.col == 0 && c.n > st.n_out) st.n_out = c.n;
  int rc = encode_weight_map(w_packed, w_rows, &st.wmap);
  if (rc) return rc;
  st.defined = true;
  p->finalized = false;
  return SF_OK;
}

int sf_plan_define_stage_fold(sf_plan* p, int stage, int which_se, const float* w32, const int32_t* row_meta, void* w_scaled) {
  if (!p || stage < 0 || stage >= SF_MAX_STAGES || !p->stage[stage].defined) return fail(SF_ERR_INVALID, "bad stage");
  if (which_se < 0 || which_se > 1 || !w32 || !row_meta || !w_scaled) return fail(SF_ERR_INVALID, "bad fold arguments");
  Stage& st = p->stage[stage];
  st.fold_se = which_se;
  if (int rc = layout_operand_rings(p, st, false)) return rc;      // per-sample weights are streamed, never resident
  st.w32 = w32;
  st.row_meta = row_meta;
  st.w_scaled = w_scaled;
  // the stage streams its sample's rows of the scaled copy: one weight map over [max_images * w_rows][64]
  return encode_weight_map(w_scaled, st.w_rows * p->g.max_images, &st.wmap);
}

int sf_plan_define_se(sf_plan* p, int which, const float* fc1, const float* fc2, int in_buf, int out_buf) {
  if (!p || which < 0 || which > 1) return fail(SF_ERR_INVALID, "bad SE index");
  p->se[which] = SeDef{true, fc1, fc2, in_buf, out_buf};
  return SF_OK;
}

int sf_plan_define_event_graph(sf_plan* p, const int32_t* cell0, const int32_t* cell1, int n_cell, const int32_t* prior, int n_prior) {
  if (!p) return fail(SF_ERR_INVALID, "null plan");
  p->cell[0].assign(cell0, cell0 + n_cell);
  p->cell[1].assign(cell1, cell1 + n_cell);
  p->prior.assign(prior, prior + n_prior);
  return SF_OK;
}

int sf_plan_finalize(sf_plan* p) {
  if (!p) return fail(SF_ERR_INVALID, "null plan");
  int max_smem = 0;
  for (const Stage& st : p->stage)
    if (st.defined && st.smem > max_smem) max_smem = st.smem;
  for (int epi = 0; epi < SF_EPI_KERNELS; ++epi)
    for (int x3 = 0; x3 < 2; ++x3) {
      const void* k = reinterpret_cast<const void*>(kernel_for(epi, x3 != 0, p->g.C));
      if (k) SF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));      // null: no such variant at this width
    }
  p->finalized = true;
  return SF_OK;
}

int sf_plan_smem_bytes(sf_plan* p, int stage) {
  if (!p || stage < 0 || stage >= SF_MAX_STAGES || !p->stage[stage].defined) return fail(SF_ERR_INVALID, "bad stage");
  return p->stage[stage].smem;
}

int sf_plan_run_stage(sf_plan* p, int stage, const sf_event* ev, const int32_t* table, void* stream) {
  if (!p || !ev || !table) return fail(SF_ERR_INVALID, "null argument");
  if (!p->finalized) return fail(SF_ERR_STATE, "plan not finalised");
  return run_item(p, stage, ev, table, reinterpret_cast<cudaStream_t>(stream));
}

int sf_plan_run_events(sf_plan* p, const sf_event* evs, int n_events, const int32_t* table, void* stream) {
  if (!p || !evs || !table) return fail(SF_ERR_INVALID, "null argument");
  if (!p->finalized) return fail(SF_ERR_STATE, "plan not finalised");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  p->last_launches = 0;
  for (int i = 0; i < n_events; ++i) {
    const sf_event* ev = evs + i;
    if (ev->kind < 0 || ev->kind > 1) return fail(SF_ERR_INVALID, "bad event kind");
    if (ev->run_cell)
      for (int item : p->cell[ev->kind]) {
        int rc = run_item(p, item, ev, table, s);
        if (rc) return rc;
      }
    if (ev->run_prior)
      for (int item : p->prior) {
        int rc = run_item(p, item, ev, table, s);
        if (rc) return rc;
      }
  }
  return SF_OK;
}

int sf_plan_last_launches(sf_plan* p) { return p ? p->last_launches : 0; }

int sf_plan_se_reduce(sf_plan* p, int which, const sf_event* ev, const int32_t* table, int px0, int px1, void* stream) {
  if (!p || !ev || !table || which < 0 || which > 1) return fail(SF_ERR_INVALID, "bad argument");
  return launch_se_reduce(p, which, ev, table, px0, px1, false, reinterpret_cast<cudaStream_t>(stream));
}

int sf_plan_se_reduce_totals(sf_plan* p, int which, const sf_event* ev, const int32_t* table, int px0, int px1, void* stream) {
  if (!p || !ev || !table || which < 0 || which > 1) return fail(SF_ERR_INVALID, "bad argument");
  int n = launch_se_reduce(p, which, ev, table, px0, px1, false, reinterpret_cast<cudaStream_t>(stream), true);
  return n < 0 ? n : SF_OK;
}

int sf_plan_se_totals_ptr(sf_plan* p, int which, float** out) {
  if (!p || !out || which < 0 || which > 1 || !p->f32[SF_F32_SE_SUMS]) return fail(SF_ERR_INVALID, "bad argument");
  *out = se_scale_ptr(p, which);
  return SF_OK;
}

int sf_plan_se_finish(sf_plan* p, int which, const sf_event* ev, const int32_t* table, float inv_n, void* stream) {
  if (!p || !ev || !table || which < 0 || which > 1) return fail(SF_ERR_INVALID, "bad argument");
  if (ev->n_active <= 0) return SF_OK;
  const SeDef& se = p->se[which];
  if (!se.defined) return fail(SF_ERR_STATE, "SE layer not defined");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* scale = se_scale_ptr(p, which);       // holds the (all-reduced) channel totals [n_active][2C]; the scales replace them in place
  const int CH = 2 * p->g.C;
  if (CH == 256) se_scale_kernel<256><<<ev->n_active, 256, 0, s>>>(scale, 1, inv_n, se.fc1, se.fc2, scale);
  else se_scale_kernel<128><<<ev->n_active, 256, 0, s>>>(scale, 1, inv_n, se.fc1, se.fc2, scale);
  SF_CUDA(cudaGetLastError());
  p->last_launches += 1;
  bool folded = false;
  for (Stage& st : p->stage) {
    if (!st.defined || st.fold_se != which) continue;
    dim3 grid((st.w_rows * 8 + 255) / 256, ev->n_active);
    se_fold_kernel<<<grid, 256, 0, s>>>(st.w32, st.row_meta, scale, reinterpret_cast<__nv_bfloat16*>(st.w_scaled), st.w_rows, CH);
    SF_CUDA(cudaGetLastError());
    p->last_launches += 1;
    folded = true;
  }
  if (folded) return SF_OK;
  return launch_se_apply(p, which, ev, table, 0, inv_n, s);      // unfolded plan: y = z * scale
}

int sf_halo_copy(void* const* tensors, const long long* batch_stride_bytes, const long long* row_bytes, int n_tensors, int B, int nrows,
                 void* flat_a, int row0_a, void* flat_b, int row0_b, int to_flat, void* stream) {
  if (!tensors || !batch_stride_bytes || !row_bytes || n_tensors < 1 || n_tensors > 6 || B < 1 || nrows < 1 || (!flat_a && !flat_b))
    return fail(SF_ERR_INVALID, "bad halo copy arguments");
  HaloCopy h;
  memset(&h, 0, sizeof(h));
  long long total = 0;
  for (int t = 0; t < n_tensors; ++t) {
    if (!tensors[t] || (row_bytes[t] & 15) || (batch_stride_bytes[t] & 15) || (reinterpret_cast<uintptr_t>(tensors[t]) & 15))
      return fail(SF_ERR_INVALID, "halo tensors must be 16-byte aligned with row sizes that are multiples of 16 bytes");
    h.base[t] = reinterpret_cast<char*>(tensors[t]);
    h.batch_stride[t] = batch_stride_bytes[t];
    h.row_bytes[t] = row_bytes[t];
    total += (long long)B * nrows * row_bytes[t];
  }
  h.flat[0] = reinterpret_cast<char*>(flat_a); h.row0[0] = row0_a;
  h.flat[1] = reinterpret_cast<char*>(flat_b); h.row0[1] = row0_b;
  h.n_tensors = n_tensors; h.B = B; h.nrows = nrows; h.to_flat = to_flat;
  const int grid = (int)std::min<long long>(((total >> 4) + 255) / 256, 148 * 8);
  halo_copy_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(h);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

// ---- peer memory over NVLink (sf_peer.cuh): arena allocation / export / import, and the exchange kernels ----------------------
int sf_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!ptr || !handle64 || bytes == 0) return fail(SF_ERR_INVALID, "bad peer arena arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  SF_CUDA(cudaMalloc(&p, bytes));
  SF_CUDA(cudaMemset(p, 0, bytes));
  SF_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(SF_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return SF_OK;
}
int sf_peer_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) return fail(SF_ERR_INVALID, "bad peer handle arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  SF_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SF_OK;
}
int sf_peer_close(void* ptr) {
  if (!ptr) return SF_OK;
  SF_CUDA(cudaIpcCloseMemHandle(ptr));
  return SF_OK;
}
int sf_peer_free(void* ptr) {
  if (!ptr) return SF_OK;
  SF_CUDA(cudaFree(ptr));
  return SF_OK;
}

namespace {
int fill_peer_halo(sf::PeerHalo& a, void* const* tensors, const long long* batch_stride_bytes, const long long* row_bytes, int n_tensors, int B,
                   int nrows, void* flat_a, int row0_a, void* flat_b, int row0_b, long long parity_stride, unsigned* flag_a, unsigned* flag_b,
                   unsigned* seq, int* err, unsigned long long* trace, long long* total) {
  if (!tensors || !batch_stride_bytes || !row_bytes || n_tensors < 1 || n_tensors > 6 || B < 1 || nrows < 1 || (!flat_a && !flat_b) || !seq ||
      (flat_a && !flag_a) || (flat_b && !flag_b) || (parity_stride & 15))
    return fail(SF_ERR_INVALID, "bad peer halo arguments");
  memset(&a, 0, sizeof(a));
  *total = 0;
  for (int t = 0; t < n_tensors; ++t) {
    if (!tensors[t] || (row_bytes[t] & 15) || (batch_stride_bytes[t] & 15) || (reinterpret_cast<uintptr_t>(tensors[t]) & 15))
      return fail(SF_ERR_INVALID, "halo tensors must be 16-byte aligned with row sizes that are multiples of 16 bytes");
    a.h.base[t] = reinterpret_cast<char*>(tensors[t]);
    a.h.batch_stride[t] = batch_stride_bytes[t];
    a.h.row_bytes[t] = row_bytes[t];
    *total += (long long)B * nrows * row_bytes[t];
  }
  if (*total > parity_stride) return fail(SF_ERR_INVALID, "halo rows exceed the receive buffer");
  a.h.flat[0] = reinterpret_cast<char*>(flat_a); a.h.row0[0] = row0_a;
  a.h.flat[1] = reinterpret_cast<char*>(flat_b); a.h.row0[1] = row0_b;
  a.h.n_tensors = n_tensors; a.h.B = B; a.h.nrows = nrows; a.h.to_flat = 1;
  a.parity_stride = parity_stride;
  a.flag[0] = flag_a; a.flag[1] = flag_b;
  a.seq = seq; a.err = err; a.trace = trace;
  a.timeout_ns = 10000000000ull;
  return SF_OK;
}
}  // namespace

int sf_halo_push(void* const* tensors, const long long* batch_stride_bytes, const long long* row_bytes, int n_tensors, int B, int nrows,
                 void* peer_flat_a, int row0_a, void* peer_flat_b, int row0_b, long long parity_stride, unsigned* peer_flag_a,
                 unsigned* peer_flag_b, unsigned* seq, unsigned long long* trace, void* stream) {
  sf::PeerHalo a;
  long long total;
  if (int rc = fill_peer_halo(a, tensors, batch_stride_bytes, row_bytes, n_tensors, B, nrows, peer_flat_a, row0_a, peer_flat_b, row0_b,
                              parity_stride, peer_flag_a, peer_flag_b, seq, nullptr, trace, &total))
    return rc;
  const int grid = (int)std::min<long long>(((total >> 4) + 255) / 256, 148 * 4);
  sf::halo_push_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_halo_pull(void* const* tensors, const long long* batch_stride_bytes, const long long* row_bytes, int n_tensors, int B, int nrows,
                 void* flat_a, int row0_a, void* flat_b, int row0_b, long long parity_stride, unsigned* flag_a, unsigned* flag_b,
                 unsigned* seq, int* err, unsigned long long* trace, void* stream) {
  sf::PeerHalo a;
  long long total;
  if (!err) return fail(SF_ERR_INVALID, "sf_halo_pull needs an error word");
  if (int rc = fill_peer_halo(a, tensors, batch_stride_bytes, row_bytes, n_tensors, B, nrows, flat_a, row0_a, flat_b, row0_b, parity_stride,
                              flag_a, flag_b, seq, err, trace, &total))
    return rc;
  const int grid = (int)std::min<long long>(((total >> 4) + 255) / 256, 148 * 4);
  sf::halo_pull_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_peer_allreduce_f32(float* data, int n, int n_max, int rank, int world, void* const* slots, void* const* flags, unsigned* seq, int* err,
                          unsigned long long* trace, void* stream) {
  if (!data || n < 1 || n > n_max || world < 1 || world > sf::PEER_MAX || rank < 0 || rank >= world || !slots || !flags || !seq || !err)
    return fail(SF_ERR_INVALID, "bad peer all-reduce arguments");
  sf::PeerReduce a;
  memset(&a, 0, sizeof(a));
  a.data = data; a.n = n; a.n_max = n_max; a.rank = rank; a.world = world;
  for (int r = 0; r < world; ++r) {
    if (!slots[r] || !flags[r]) return fail(SF_ERR_INVALID, "peer all-reduce: missing arena pointer");
    a.slots[r] = reinterpret_cast<float*>(slots[r]);
    a.flags[r] = reinterpret_cast<unsigned*>(flags[r]);
  }
  a.seq = seq; a.err = err; a.trace = trace; a.timeout_ns = 10000000000ull;
  sf::peer_allreduce_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_plan_se_apply(sf_plan* p, int which, const sf_event* ev, const int32_t* table, int n_partials, float inv_n, void* stream) {
  if (!p || !ev || !table || which < 0 || which > 1) return fail(SF_ERR_INVALID, "bad argument");
  return launch_se_apply(p, which, ev, table, n_partials, inv_n, reinterpret_cast<cudaStream_t>(stream));
}

int sf_pack_nchw_f32(const float* src, void* dst_hi, void* dst_lo, int n_images, int C, int H, int W, void* stream) {
  if (!src || !dst_hi || C % 64 || n_images <= 0) return fail(SF_ERR_INVALID, "bad pack arguments");
  const int hw = H * W;
  dim3 grid((hw + PACK_PX - 1) / PACK_PX, C / 64, n_images);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dst_lo)
    pack_nchw_kernel<true><<<grid, 256, 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_hi), reinterpret_cast<__nv_bfloat16*>(dst_lo), C, hw);
  else
    pack_nchw_kernel<false><<<grid, 256, 0, s>>>(src, reinterpret_cast<__nv_bfloat16*>(dst_hi), nullptr, C, hw);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_normal_policy(long long numel, int device, int* grid, int* offset_per_slot) {
  if (numel <= 0 || !grid || !offset_per_slot) return fail(SF_ERR_INVALID, "bad normal policy arguments");
  // two cheap attribute queries, cached per device (cudaGetDeviceProperties costs milliseconds and this runs once per rollout)
  static int sms[64] = {0}, threads_per_sm[64] = {0};
  if (device < 0 || device >= 64) return fail(SF_ERR_INVALID, "bad device ordinal");
  if (!sms[device]) {
    int a = 0, b = 0;
    SF_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, device));
    SF_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrMaxThreadsPerMultiProcessor, device));
    threads_per_sm[device] = b;
    sms[device] = a;
  }
  // at::cuda::detail calc_execution_policy: block 256, unroll 4, grid = min(SMs * (maxThreadsPerSM / 256), ceil(numel / 256))
  const long long blocks_per_sm = threads_per_sm[device] / 256;
  long long g = (numel + 255) / 256;
  const long long cap = (long long)sms[device] * blocks_per_sm;
  if (g > cap) g = cap;
  *grid = (int)g;
  *offset_per_slot = (int)(((numel - 1) / (256 * g * 4) + 1) * 4);       // curand4_engine_calls = 4 per loop iteration
  return SF_OK;
}

int sf_normal_fill_slots(float* out, int n_slots, long long numel, unsigned long long seed, unsigned long long offset0, int grid,
                         int offset_per_slot, void* stream) {
  if (!out || n_slots <= 0 || numel <= 0 || grid <= 0 || offset_per_slot <= 0 || n_slots > 65535) return fail(SF_ERR_INVALID, "bad normal fill arguments");
  if ((offset0 & 3) || (offset_per_slot & 3)) return fail(SF_ERR_INVALID, "Philox offsets must be multiples of 4 (ATen's generator invariant)");
  // about 8 waves of resident blocks in total; each block row walks n_slots / grid.y slots
  int rows = (8 * 148 * 8 + grid - 1) / grid;
  if (rows > n_slots) rows = n_slots;
  normal_slots_kernel<<<dim3(grid, rows), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, numel, seed, offset0, (unsigned int)offset_per_slot, n_slots,
                                                                                            nullptr);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_normal_fill_slot_list(float* out, const int32_t* slot_list, int n_list, long long numel, unsigned long long seed, unsigned long long offset0,
                             int grid, int offset_per_slot, void* stream) {
  if (!out || !slot_list || n_list <= 0 || numel <= 0 || grid <= 0 || offset_per_slot <= 0 || n_list > 65535) return fail(SF_ERR_INVALID, "bad normal fill arguments");
  if ((offset0 & 3) || (offset_per_slot & 3)) return fail(SF_ERR_INVALID, "Philox offsets must be multiples of 4 (ATen's generator invariant)");
  int rows = (8 * 148 * 8 + grid - 1) / grid;
  if (rows > n_list) rows = n_list;
  normal_slots_kernel<<<dim3(grid, rows), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(out, numel, seed, offset0, (unsigned int)offset_per_slot, n_list,
                                                                                            slot_list);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_maxpool2(const void* src_hi, const void* src_lo, void* dst_hi, void* dst_lo, int n_images, int H, int W, int C, void* stream) {
  if (!src_hi || !dst_hi || C % 8 || (H & 1) || (W & 1) || n_images <= 0) return fail(SF_ERR_INVALID, "bad maxpool arguments");
  const size_t total = (size_t)n_images * (H / 2) * (W / 2) * (C / 8);
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto sh = reinterpret_cast<const __nv_bfloat16*>(src_hi); auto sl = reinterpret_cast<const __nv_bfloat16*>(src_lo);
  auto dh = reinterpret_cast<__nv_bfloat16*>(dst_hi); auto dl = reinterpret_cast<__nv_bfloat16*>(dst_lo);
  if (src_lo && dst_lo) maxpool2_kernel<true><<<grid, 256, 0, s>>>(sh, sl, dh, dl, n_images, H, W, C);
  else maxpool2_kernel<false><<<grid, 256, 0, s>>>(sh, sl, dh, dl, n_images, H, W, C);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_upsample2(const void* src, void* dst, int n_images, int H, int W, int C, void* stream) {
  if (!src || !dst || C % 8 || n_images <= 0) return fail(SF_ERR_INVALID, "bad upsample arguments");
  const size_t total = (size_t)n_images * (2 * H) * (2 * W) * (C / 8);
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  upsample2_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst),
                                                                           n_images, H, W, C / 8);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_space_to_depth2(const void* src, void* dst, int n_images, int H, int W, int C, void* stream) {
  if (!src || !dst || C % 8 || (H & 1) || (W & 1) || n_images <= 0) return fail(SF_ERR_INVALID, "bad space-to-depth arguments");
  const size_t total = (size_t)n_images * (H / 2) * (W / 2) * 4 * (C / 8);
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  space_to_depth2_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst),
                                                                                 n_images, H, W, C / 8);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_bilinear_up2_add(const void* src_hi, const void* src_lo, const void* skip_hi, const void* skip_lo, void* dst_hi, void* dst_lo,
                        int n_images, int H, int W, int C, void* stream) {
  if (!src_hi || !skip_hi || !dst_hi || C % 8 || n_images <= 0 || H <= 0 || W <= 0) return fail(SF_ERR_INVALID, "bad bilinear arguments");
  const bool x3 = src_lo && skip_lo && dst_lo;
  const size_t total = (size_t)n_images * (2 * H) * (2 * W) * (C / 8);
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto sh = reinterpret_cast<const __nv_bfloat16*>(src_hi); auto sl = reinterpret_cast<const __nv_bfloat16*>(src_lo);
  auto kh = reinterpret_cast<const __nv_bfloat16*>(skip_hi); auto kl = reinterpret_cast<const __nv_bfloat16*>(skip_lo);
  auto dh = reinterpret_cast<__nv_bfloat16*>(dst_hi); auto dl = reinterpret_cast<__nv_bfloat16*>(dst_lo);
  if (x3) bilinear_up2_add_kernel<true><<<grid, 256, 0, s>>>(sh, sl, kh, kl, dh, dl, n_images, H, W, C);
  else bilinear_up2_add_kernel<false><<<grid, 256, 0, s>>>(sh, sl, kh, kl, dh, dl, n_images, H, W, C);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_head_1x1(const void* src_hi, const void* src_lo, const float* w, const float* b, int K, int sigmoid_out, float* out, unsigned char* mask,
                int n_images, int H, int W, void* stream) {
  if (!src_hi || !w || !b || !out || K < 1 || K > 4 || n_images <= 0) return fail(SF_ERR_INVALID, "bad head arguments");
  const size_t total = (size_t)n_images * H * W;
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto sh = reinterpret_cast<const __nv_bfloat16*>(src_hi); auto sl = reinterpret_cast<const __nv_bfloat16*>(src_lo);
  if (src_lo) head_1x1_kernel<true><<<grid, 256, 0, s>>>(sh, sl, w, b, K, sigmoid_out, out, mask, n_images, H * W);
  else head_1x1_kernel<false><<<grid, 256, 0, s>>>(sh, sl, w, b, K, sigmoid_out, out, mask, n_images, H * W);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_cast_nhwc_f32(const float* src, const int32_t* slots, void* dst_hi, void* dst_lo, int n_out, int C, int H, int W, void* stream) {
  if (!src || !dst_hi || C % 8 || n_out <= 0) return fail(SF_ERR_INVALID, "bad cast arguments");
  const size_t per = (size_t)H * W * C, total = (size_t)n_out * per / 8;
  const int grid = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dst_lo) cast_nhwc_kernel<true><<<grid, 256, 0, s>>>(src, slots, reinterpret_cast<__nv_bfloat16*>(dst_hi), reinterpret_cast<__nv_bfloat16*>(dst_lo), n_out, per);
  else cast_nhwc_kernel<false><<<grid, 256, 0, s>>>(src, slots, reinterpret_cast<__nv_bfloat16*>(dst_hi), nullptr, n_out, per);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_dwconv7_ln(const void* src_hi, const void* src_lo, void* dst_hi, void* dst_lo, const float* dw_w, const float* dw_b,
                  const float* ln_w, const float* ln_b, int n_images, int C, int H, int W, void* stream) {
  if (!src_hi || !dst_hi || !dw_w || !dw_b || !ln_w || !ln_b || n_images <= 0) return fail(SF_ERR_INVALID, "bad dwconv arguments");
  if (C != 64 && C != 128) return fail(SF_ERR_INVALID, "depthwise 7x7 + LayerNorm is built for 64 or 128 channels");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto sh = reinterpret_cast<const __nv_bfloat16*>(src_hi); auto sl = reinterpret_cast<const __nv_bfloat16*>(src_lo);
  auto dh = reinterpret_cast<__nv_bfloat16*>(dst_hi); auto dl = reinterpret_cast<__nv_bfloat16*>(dst_lo);
  if (C == 128) {       // generic-width kernel: one warp per pixel
    const long long total = (long long)n_images * H * W;
    const int grid = (int)std::min<long long>((total + 7) / 8, 148 * 32);
    if (src_lo && dst_lo) dwconv7_ln_wide_kernel<128, true><<<grid, 256, 0, s>>>(sh, sl, dh, dl, dw_w, dw_b, ln_w, ln_b, H, W, total);
    else dwconv7_ln_wide_kernel<128, false><<<grid, 256, 0, s>>>(sh, sl, dh, dl, dw_w, dw_b, ln_w, ln_b, H, W, total);
    SF_CUDA(cudaGetLastError());
    return SF_OK;
  }
  dim3 grid((W + DW_TILE_W - 1) / DW_TILE_W, (H + DW_TILE_H - 1) / DW_TILE_H, n_images);
  // 146 KB of dynamic shared memory (fp32 halo tile + filter): opt in per call (a per-device attribute; the call is cheap)
  SF_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(src_lo && dst_lo ? dwconv7_ln_kernel<true> : dwconv7_ln_kernel<false>),
                               cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_BYTES));
  if (src_lo && dst_lo) dwconv7_ln_kernel<true><<<grid, 256, DW_SMEM_BYTES, s>>>(sh, sl, dh, dl, dw_w, dw_b, ln_w, ln_b, H, W);
  else dwconv7_ln_kernel<false><<<grid, 256, DW_SMEM_BYTES, s>>>(sh, sl, dh, dl, dw_w, dw_b, ln_w, ln_b, H, W);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_aspp_pool_bias(const void* src_hi, const void* src_lo, const float* pool_w, const float* pool_b, const float* proj_w,
                      const float* proj_b, float* scratch, float* out, int n_images, int C, int H, int W, void* stream) {
  if (!src_hi || !pool_w || !pool_b || !proj_w || !proj_b || !scratch || !out || n_images <= 0) return fail(SF_ERR_INVALID, "bad pool arguments");
  if (C != 64 && C != 128) return fail(SF_ERR_INVALID, "ASPP pooling is built for 64 or 128 input channels");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  auto sh = reinterpret_cast<const __nv_bfloat16*>(src_hi); auto sl = reinterpret_cast<const __nv_bfloat16*>(src_lo);
  dim3 grid(POOL_PARTS, n_images);
  const float inv_n = 1.0f / (float)(H * W);
  if (C == 64) {
    if (src_lo) pool_partial_kernel<64, true><<<grid, 256, 0, s>>>(sh, sl, scratch, H * W);
    else pool_partial_kernel<64, false><<<grid, 256, 0, s>>>(sh, sl, scratch, H * W);
    pool_bias_kernel<64><<<n_images, 128, 0, s>>>(scratch, inv_n, pool_w, pool_b, proj_w, proj_b, out);
  } else {
    if (src_lo) pool_partial_kernel<128, true><<<grid, 256, 0, s>>>(sh, sl, scratch, H * W);
    else pool_partial_kernel<128, false><<<grid, 256, 0, s>>>(sh, sl, scratch, H * W);
    pool_bias_kernel<128><<<n_images, 128, 0, s>>>(scratch, inv_n, pool_w, pool_b, proj_w, proj_b, out);
  }
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

int sf_unpack_nhwc_f32(const float* src, float* dst, const int32_t* slots, int n_out, int C, int H, int W, void* stream) {
  if (!src || !dst || C % 64 || n_out <= 0) return fail(SF_ERR_INVALID, "bad unpack arguments");
  const int hw = H * W;
  dim3 grid((hw + 31) / 32, C / 64, n_out);
  unpack_nhwc_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, slots, C, hw);
  SF_CUDA(cudaGetLastError());
  return SF_OK;
}

}  // extern "C"
