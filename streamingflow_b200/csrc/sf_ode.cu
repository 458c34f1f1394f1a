// sf_ode.cu -- the ODE head driven from include/sf_b200.h alone: weight packing (BatchNorm fold, cat[state, state] fold, tap
// order, hi/lo split), workspace carving, stage / SE / event-graph definition and the host step schedule, all as plain host
// C++ on top of the sf_plan_* entry points of sf_plan.cu.  A non-Python host needs nothing else to run the path; the Python
// engine (streamingflow_b200/engine.py) packs its weights through the same sf_pack_* calls.
//
// Reference lines restated here (parameter names: SURVEY.md 8a "state_dict contract"):
//   gates    conv_update_1/2, conv_reset_1/2                 temporal_ode_bayes.py:135-140,150-155
//   propose  conv_state_tilde_1/2 + GRU blend                :143-146,158-161
//   decode   conv_decoder_2                                  :121
//   trunk    trusting_gate.0.layers.0 (7x7) + LN + GELU + layers.3 (1x1) + LN + GELU      convolutions.py:356-361
//   mix      layers.6 (3x3) + LN + GELU, projection, 1x1 -> 2, softmax, mix, Euler / jump  :362-380, tob:124-131,446
//   q1..q5   p_model = ConvNet with eval-mode BatchNorm folded                              res_models.py:168-180
//   schedule start time :508, advance-to-observation :539-553, jump / record :562-581, advance-to-target :585-604,
//            output selection :606-622
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sf_b200.h"

extern "C" void sf_internal_set_error(const char* msg);      // sf_plan.cu

namespace sfo {

int ofail(int code, const std::string& msg) {      // sf_last_error() (sf_plan.cu) reports this file's failures too
  sf_internal_set_error(msg.c_str());
  return code;
}

// ---------------------------------------------------------------------------------------------------------------
// small host tensors
// ---------------------------------------------------------------------------------------------------------------
struct W4 {                    // conv weight [n][cin][R][R], torch layout
  int n = 0, cin = 0, R = 0;
  std::vector<float> d;
  float& at(int o, int c, int ky, int kx) { return d[(((size_t)o * cin + c) * R + ky) * R + kx]; }
  float at(int o, int c, int ky, int kx) const { return d[(((size_t)o * cin + c) * R + ky) * R + kx]; }
};
typedef std::vector<float> V1;

W4 in_slice(const W4& w, int c0, int c1) {       // w[:, c0:c1]
  W4 r;
  r.n = w.n; r.cin = c1 - c0; r.R = w.R;
  r.d.resize((size_t)r.n * r.cin * r.R * r.R);
  for (int o = 0; o < w.n; ++o)
    for (int c = c0; c < c1; ++c)
      memcpy(&r.at(o, c - c0, 0, 0), &w.d[(((size_t)o * w.cin + c) * w.R) * w.R], sizeof(float) * w.R * w.R);
  return r;
}
W4 out_slice(const W4& w, int o0, int o1) {      // w[o0:o1]
  W4 r;
  r.n = o1 - o0; r.cin = w.cin; r.R = w.R;
  const size_t per = (size_t)w.cin * w.R * w.R;
  r.d.assign(w.d.begin() + o0 * per, w.d.begin() + o1 * per);
  return r;
}
W4 cat_out(const std::vector<W4>& ws) {          // torch.cat(ws, 0)
  W4 r;
  r.cin = ws[0].cin; r.R = ws[0].R;
  for (const W4& w : ws) {
    r.n += w.n;
    r.d.insert(r.d.end(), w.d.begin(), w.d.end());
  }
  return r;
}
W4 add_w(const W4& a, const W4& b) {
  W4 r = a;
  for (size_t i = 0; i < r.d.size(); ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
V1 cat_v(const std::vector<V1>& vs) {
  V1 r;
  for (const V1& v : vs) r.insert(r.end(), v.begin(), v.end());
  return r;
}
V1 slice_v(const V1& v, int a, int b) { return V1(v.begin() + a, v.begin() + b); }

// round-to-nearest-even fp32 -> bf16, as torch's .to(torch.bfloat16)
uint16_t bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);     // NaN stays NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
float bf16_to_f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct TensorTable {
  std::map<std::string, std::pair<const float*, int64_t>> t;
  std::string missing;
  TensorTable(const sf_tensor* ts, int n, const std::string& strip) {
    for (int i = 0; i < n; ++i) {
      if (!ts[i].name || !ts[i].data) continue;
      std::string name = ts[i].name;
      if (!strip.empty()) {
        if (name.compare(0, strip.size(), strip) != 0) continue;
        name = name.substr(strip.size());
      }
      t[name] = std::make_pair(ts[i].data, ts[i].numel);
    }
  }
  bool has(const std::string& k) const { return t.count(k) != 0; }
  int64_t numel(const std::string& k) {
    auto it = t.find(k);
    if (it == t.end()) { if (missing.empty()) missing = k; return 0; }
    return it->second.second;
  }
  V1 vec(const std::string& k, int64_t expect = -1) {
    auto it = t.find(k);
    if (it == t.end() || (expect >= 0 && it->second.second != expect)) { if (missing.empty()) missing = k; return V1(expect > 0 ? (size_t)expect : 0, 0.0f); }
    return V1(it->second.first, it->second.first + it->second.second);
  }
  W4 conv(const std::string& k, int n, int cin, int R) {
    W4 w;
    w.n = n; w.cin = cin; w.R = R;
    w.d = vec(k, (int64_t)n * cin * R * R);
    return w;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// stage descriptions before / after packing
// ---------------------------------------------------------------------------------------------------------------
struct ChunkSrc {
  int buf, c0;
  W4 w;          // [n][64][R][R]
  int col, init, ox, oy;
};
struct StageBuild {
  std::string name;
  int se_layer = -1;
  int epilogue = 0, flags = 0;
  V1 vec;
  std::vector<int32_t> io, io_off;
  std::vector<ChunkSrc> chunks;
  int fold_se = -1;
  V1 b2b_w;      // [C][C] weights of the fused 1x1 follow-up conv ([n][k], K-major)
  int b2b_n = 0;
  // packed
  std::vector<sf_chunk> pchunks;
  std::vector<uint16_t> w;      // [rows][64] bf16
  int w_rows = 0;
  V1 w32;
  std::vector<int32_t> row_meta;
  V1 fc1, fc2;
  void add(int buf, const W4& w4, int col_, int init_, int c0 = 0) {      // one chunk per 64 input channels
    for (int i = 0; i < w4.cin / 64; ++i) chunks.push_back(ChunkSrc{buf, c0 + 64 * i, in_slice(w4, 64 * i, 64 * i + 64), col_, (init_ && i == 0) ? 1 : 0, 0, 0});
  }
};

constexpr int FLAG_KEEP_A32 = 1, FLAG_RES_SE_SCALE = 128, FLAG_PAIR_ROWS = 512, FLAG_B2B = 1024;

// rows of tap (dx, dy): taps[dx][dy][o][c] = w[o][c][dy][dx]
void tap_rows(const W4& w, int dx, int dy, bool lo, std::vector<uint16_t>& out) {
  for (int o = 0; o < w.n; ++o)
    for (int c = 0; c < 64; ++c) {
      const float v = w.at(o, c, dy, dx);
      const uint16_t h = bf16_rne(v);
      out.push_back(lo ? bf16_rne(v - bf16_to_f(h)) : h);
    }
}
void tap_rows_f32(const W4& w, int dx, int dy, V1& out) {
  for (int o = 0; o < w.n; ++o)
    for (int c = 0; c < 64; ++c) out.push_back(w.at(o, c, dy, dx));
}

std::vector<std::vector<int>> pair_groups(int R) {       // dy order of a dx column with row-paired taps: [1,0], [3,2], ..., [R-1]
  std::vector<std::vector<int>> g;
  for (int k = 0; k < R / 2; ++k) g.push_back({2 * k + 1, 2 * k});
  if (R % 2) g.push_back({R - 1});
  return g;
}

// the [rows][64] bf16 matrix the weight ring streams, in consumption order (see sf_chunk.wrow); x3: hi/lo split
void pack_stage(StageBuild& s, bool x3) {
  s.pchunks.clear();
  s.w.clear();
  int row = 0;
  const bool paired = (s.flags & FLAG_PAIR_ROWS) != 0;
  for (const ChunkSrc& ck : s.chunks) {
    const int n = ck.w.n, R = ck.w.R;
    auto desc = [&](int plane, int nrep, int init) {
      sf_chunk c;
      c.buf = ck.buf; c.plane = plane; c.c0 = ck.c0; c.R = R; c.n = n; c.nrep = nrep; c.col = ck.col; c.wrow = row; c.init = init; c.ox = ck.ox; c.oy = ck.oy;
      s.pchunks.push_back(c);
    };
    if (!paired) {
      desc(0, x3 ? 2 : 1, ck.init);
      for (int dx = 0; dx < R; ++dx)
        for (int dy = 0; dy < R; ++dy) {
          tap_rows(ck.w, dx, dy, false, s.w);
          if (x3) tap_rows(ck.w, dx, dy, true, s.w);
        }
      row += R * R * (x3 ? 2 : 1) * n;
      if (x3) {                                   // lo plane of the activations x hi weights
        desc(1, 1, 0);
        for (int dx = 0; dx < R; ++dx)
          for (int dy = 0; dy < R; ++dy) tap_rows(ck.w, dx, dy, false, s.w);
        row += R * R * n;
      }
    } else {                                      // chunk -> dx -> pair -> rep -> [tap_hi rows | tap_lo rows]
      const auto groups = pair_groups(R);
      desc(0, x3 ? 2 : 1, ck.init);
      for (int dx = 0; dx < R; ++dx)
        for (const auto& grp : groups)
          for (int rep = 0; rep < (x3 ? 2 : 1); ++rep)
            for (int dy : grp) tap_rows(ck.w, dx, dy, rep == 1, s.w);
      row += R * R * (x3 ? 2 : 1) * n;
      if (x3) {
        desc(1, 1, 0);
        for (int dx = 0; dx < R; ++dx)
          for (const auto& grp : groups)
            for (int dy : grp) tap_rows(ck.w, dx, dy, false, s.w);
        row += R * R * n;
      }
    }
  }
  if (s.flags & FLAG_B2B) {                       // the follow-up 1x1 conv's [n][k] weights: hi rows, then lo rows
    const int n = s.b2b_n;
    for (int rep = 0; rep < (x3 ? 2 : 1); ++rep)
      for (int o = 0; o < n; ++o)
        for (int k = 0; k < 64; ++k) {
          const float v = s.b2b_w[(size_t)o * 64 + k];
          const uint16_t h = bf16_rne(v);
          s.w.push_back(rep ? bf16_rne(v - bf16_to_f(h)) : h);
        }
    row += n * (x3 ? 2 : 1);
  }
  s.w_rows = row;
  // fp32 master + row metadata for a stage that gets an SE layer folded into its weights (se_fold_kernel's input)
  s.w32.clear();
  s.row_meta.clear();
  if (s.fold_se >= 0) {
    for (const ChunkSrc& ck : s.chunks) {
      const int n = ck.w.n, R = ck.w.R;
      for (int dx = 0; dx < R; ++dx)
        for (int dy = 0; dy < R; ++dy)
          for (int rep = 0; rep < (x3 ? 2 : 1); ++rep) {
            tap_rows_f32(ck.w, dx, dy, s.w32);
            s.row_meta.insert(s.row_meta.end(), n, ck.c0 | (rep ? (1 << 16) : 0));
          }
      if (x3)
        for (int dx = 0; dx < R; ++dx)
          for (int dy = 0; dy < R; ++dy) {
            tap_rows_f32(ck.w, dx, dy, s.w32);
            s.row_meta.insert(s.row_meta.end(), n, ck.c0);
          }
    }
  }
}

int isqrt_exact(int64_t v) {
  int r = (int)std::llround(std::sqrt((double)v));
  return (int64_t)r * r == v ? r : -1;
}

// eval-mode BatchNorm2d folded into the preceding bias-free conv: w' = w * g / sqrt(var + 1e-5), b' = beta - mean * scale
void bn_fold(TensorTable& T, const std::string& p, int n, int cin, int R, W4& w, V1& bias) {
  w = T.conv(p + ".conv.weight", n, cin, R);
  const V1 g = T.vec(p + ".norm.weight", n), var = T.vec(p + ".norm.running_var", n), beta = T.vec(p + ".norm.bias", n),
           mean = T.vec(p + ".norm.running_mean", n);
  bias.resize(n);
  const size_t per = (size_t)cin * R * R;
  for (int o = 0; o < n; ++o) {
    const float scale = g[o] / sqrtf(var[o] + 1e-5f);
    const float ms = mean[o] * scale;
    bias[o] = beta[o] - ms;
    for (size_t i = 0; i < per; ++i) w.d[o * per + i] = w.d[o * per + i] * scale;
  }
}

}  // namespace sfo
using namespace sfo;

struct sf_packed {
  std::vector<StageBuild> items;
  int C = 0;
};

namespace sfo {

int build_cell(TensorTable& T, const std::string& p, bool x3, int options, sf_packed* out) {
  const int C = (int)T.numel(p + "conv_decoder_2.bias");
  if (C != 64 && C != 128) return ofail(SF_ERR_INVALID, "cell weights '" + p + "*': hidden channels must be 64 or 128");
  out->C = C;
  auto g3 = [&](const std::string& k, int cin) { return T.conv(p + k + ".weight", C, cin, 3); };
  auto bias = [&](const std::string& k) { return T.vec(p + k + ".bias", C); };
  const W4 wu1 = g3("conv_update_1", 2 * C), wr1 = g3("conv_reset_1", 2 * C), wt1 = g3("conv_state_tilde_1", 2 * C);
  const W4 wu2 = g3("conv_update_2", 2 * C), wr2 = g3("conv_reset_2", 2 * C), wt2 = g3("conv_state_tilde_2", 2 * C);
  const V1 bu1 = bias("conv_update_1"), br1 = bias("conv_reset_1"), bu2 = bias("conv_update_2"), br2 = bias("conv_reset_2");
  auto x_of = [&](const W4& w) { return in_slice(w, 0, C); };
  auto s_of = [&](const W4& w) { return in_slice(w, C, 2 * C); };
  auto fold = [&](const W4& w) { return add_w(in_slice(w, 0, C), in_slice(w, C, 2 * C)); };   // gru_cell_2 sees cat[state, state] (tob:118)
  const int SX = SF_SRC_X, SS = SF_SRC_STATE_IN;
  std::vector<StageBuild>& L = out->items;
  auto stage = [&](const char* name, int epi, const V1& vec, std::vector<int32_t> io, int flags = 0) -> StageBuild& {
    L.emplace_back();
    StageBuild& s = L.back();
    s.name = name; s.epilogue = epi; s.vec = vec; s.io = io; s.io_off.assign(io.size(), 0); s.flags = flags;
    return s;
  };
  if (C == 64) {      // all four gates in one 256-column launch; both proposals in one launch
    StageBuild& gs = stage("gates", SF_EPI_GATES, cat_v({bu1, br1, bu2, br2}), {SF_BUF_U1, SF_BUF_G1, SF_BUF_U2, SF_BUF_G2});
    gs.add(SS, cat_out({s_of(wu1), s_of(wr1), fold(wu2), fold(wr2)}), 0, 1);
    gs.add(SX, cat_out({x_of(wu1), x_of(wr1)}), 0, 0);
    StageBuild& pr = stage("propose", SF_EPI_PROPOSE, cat_v({bias("conv_state_tilde_1"), bias("conv_state_tilde_2")}),
                           {SF_BUF_U1, SF_BUF_U2, SF_BUF_A, SF_BUF_HH}, FLAG_KEEP_A32);
    pr.add(SX, x_of(wt1), 0, 1);
    pr.add(SF_BUF_G1, s_of(wt1), 0, 0);
    pr.add(SS, x_of(wt2), 64, 1);
    pr.add(SF_BUF_G2, s_of(wt2), 64, 0);
  } else {            // 128 channels: one gate pair / one proposal per launch
    StageBuild& g1 = stage("gates_1", SF_EPI_GATES, cat_v({bu1, br1}), {SF_BUF_U1, SF_BUF_G1});
    g1.add(SS, cat_out({s_of(wu1), s_of(wr1)}), 0, 1);
    g1.add(SX, cat_out({x_of(wu1), x_of(wr1)}), 0, 0);
    stage("gates_2", SF_EPI_GATES, cat_v({bu2, br2}), {SF_BUF_U2, SF_BUF_G2}).add(SS, cat_out({fold(wu2), fold(wr2)}), 0, 1);
    StageBuild& p1 = stage("propose_1", SF_EPI_PROPOSE, bias("conv_state_tilde_1"), {SF_BUF_U1, SF_BUF_A}, FLAG_KEEP_A32);
    p1.add(SX, x_of(wt1), 0, 1);
    p1.add(SF_BUF_G1, s_of(wt1), 0, 0);
    StageBuild& p2 = stage("propose_2", SF_EPI_PROPOSE, bias("conv_state_tilde_2"), {SF_BUF_U2, SF_BUF_HH});
    p2.add(SS, x_of(wt2), 0, 1);
    p2.add(SF_BUF_G2, s_of(wt2), 0, 0);
  }
  // 64 channels: the 3x3 stages with ONE 64-column accumulator block pair their vertically adjacent taps too (N = 128 MMAs)
  const bool pair3 = C == 64 && (options & SF_PACK_PAIR_3X3);
  stage("decode", SF_EPI_DECODE, bias("conv_decoder_2"), {SF_BUF_B}, pair3 ? FLAG_PAIR_ROWS : 0).add(SF_BUF_HH, g3("conv_decoder_2", C), 0, 1);
  const std::string t = "trusting_gate.0.";
  const W4 w7 = T.conv(p + t + "layers.0.weight", C, 2 * C, 7);
  auto ln = [&](int i, const char* wb) { return T.vec(p + t + "layers." + std::to_string(i) + "." + wb, C); };
  const bool pair = C == 64 && (options & SF_PACK_PAIR_ROWS), b2b = C == 64 && (options & SF_PACK_B2B);
  if (b2b) {          // 7x7 + LN + GELU + 1x1 + LN + GELU as ONE stage; vertically adjacent taps paired into N = 128 MMAs
    StageBuild& tr = stage("trunk", SF_EPI_LNGELU, cat_v({ln(1, "weight"), ln(1, "bias"), ln(4, "weight"), ln(4, "bias")}), {SF_BUF_T2},
                           FLAG_B2B | (pair ? FLAG_PAIR_ROWS : 0));
    tr.add(SF_BUF_A, in_slice(w7, 0, C), 0, 1);
    tr.add(SF_BUF_B, in_slice(w7, C, 2 * C), 0, 0);
    tr.b2b_w = T.vec(p + t + "layers.3.weight", (int64_t)C * C);
    tr.b2b_n = C;
  } else {
    StageBuild& t7 = stage("trunk7", SF_EPI_LNGELU, cat_v({ln(1, "weight"), ln(1, "bias")}), {SF_BUF_T1}, pair ? FLAG_PAIR_ROWS : 0);
    t7.add(SF_BUF_A, in_slice(w7, 0, C), 0, 1);
    t7.add(SF_BUF_B, in_slice(w7, C, 2 * C), 0, 0);
    stage("trunk1", SF_EPI_LNGELU, cat_v({ln(4, "weight"), ln(4, "bias")}), {SF_BUF_T2}).add(SF_BUF_T1, T.conv(p + t + "layers.3.weight", C, C, 1), 0, 1);
  }
  const W4 wp = T.conv(p + t + "projection.0.weight", C, 2 * C, 1);
  const V1 wg = T.vec(p + "trusting_gate.1.weight", 2 * C);
  StageBuild& mx = stage("mix", SF_EPI_MIX, cat_v({ln(7, "weight"), ln(7, "bias"), slice_v(wg, 0, C), slice_v(wg, C, 2 * C)}), {});
  mx.add(SF_BUF_T2, T.conv(p + t + "layers.6.weight", C, C, 3), 0, 1);
  mx.add(SF_BUF_A, in_slice(wp, 0, C), C, 1);
  mx.add(SF_BUF_B, in_slice(wp, C, 2 * C), C, 0);
  if (!T.missing.empty()) return ofail(SF_ERR_INVALID, "cell weights: tensor '" + p + "...' missing or of unexpected size: " + T.missing);
  for (StageBuild& s : L) pack_stage(s, x3);
  return SF_OK;
}

int build_prior(TensorTable& T, const std::string& p, bool x3, int options, sf_packed* out) {
  const std::string m = p + "model.";
  const int C = (int)T.numel(m + "0.layers.conv_1.norm.weight");
  if (C != 64 && C != 128) return ofail(SF_ERR_INVALID, "p_model weights '" + p + "*': hidden channels must be 64 or 128");
  out->C = C;
  const bool fold_se = (options & SF_PACK_FOLD_SE) != 0;
  const int Y1 = fold_se ? SF_BUF_Z1 : SF_BUF_Y1, Y2 = fold_se ? SF_BUF_Z2 : SF_BUF_Y2;
  const int halves = (2 * C) / 128, SO = SF_SRC_STATE_OUT;
  std::vector<StageBuild>& L = out->items;
  auto stage = [&](const std::string& name, int epi, const V1& vec, std::vector<int32_t> io, std::vector<int32_t> io_off, int flags = 0) -> StageBuild& {
    L.emplace_back();
    StageBuild& s = L.back();
    s.name = name; s.epilogue = epi; s.vec = vec; s.io = io; s.io_off = io_off; s.flags = flags;
    return s;
  };
  auto se = [&](int which, int idx) {
    L.emplace_back();
    StageBuild& s = L.back();
    s.name = which ? "se1" : "se0";
    s.se_layer = which;
    s.fc1 = T.vec(m + std::to_string(idx) + ".fc.0.weight", (int64_t)(2 * C / 8) * 2 * C);
    s.fc2 = T.vec(m + std::to_string(idx) + ".fc.2.weight", (int64_t)(2 * C / 8) * 2 * C);
  };
  auto suffix = [&](int h) { return halves > 1 ? std::string(1, "ab"[h]) : std::string(); };
  W4 w1, w2, w3, w4;
  V1 b1, b2, b3, b4;
  bn_fold(T, m + "0.layers.conv_1", C, C, 3, w1, b1);
  bn_fold(T, m + "0.layers.conv_2", 2 * C, C, 3, w2, b2);
  stage("q1", SF_EPI_BIAS_LRELU, b1, {SF_BUF_Q1}, {0}, (C == 64 && (options & SF_PACK_PAIR_3X3)) ? FLAG_PAIR_ROWS : 0).add(SO, w1, 0, 1);
  const W4 wpj = T.conv(m + "0.projection.weight", 2 * C, C, 1);
  const V1 bpj = T.vec(m + "0.projection.bias", 2 * C);
  for (int h = 0; h < halves; ++h) {
    StageBuild& s = stage("q2" + suffix(h), SF_EPI_RES_PROJ, cat_v({slice_v(b2, 128 * h, 128 * h + 128), slice_v(bpj, 128 * h, 128 * h + 128)}), {SF_BUF_Z1}, {128 * h});
    s.add(SF_BUF_Q1, out_slice(w2, 128 * h, 128 * h + 128), 0, 1);
    s.add(SO, out_slice(wpj, 128 * h, 128 * h + 128), 128, 1);
  }
  se(0, 1);
  bn_fold(T, m + "2.layers.conv_1", 2 * C, 2 * C, 3, w3, b3);
  bn_fold(T, m + "2.layers.conv_2", 2 * C, 2 * C, 3, w4, b4);
  for (int h = 0; h < halves; ++h) {
    StageBuild& s = stage("q3" + suffix(h), SF_EPI_BIAS_LRELU, slice_v(b3, 128 * h, 128 * h + 128), {SF_BUF_Q3}, {128 * h});
    s.add(Y1, out_slice(w3, 128 * h, 128 * h + 128), 0, 1);
    s.fold_se = fold_se ? 0 : -1;
  }
  for (int h = 0; h < halves; ++h)
    stage("q4" + suffix(h), SF_EPI_RES_ID, slice_v(b4, 128 * h, 128 * h + 128), {Y1, SF_BUF_Z2}, {128 * h, 128 * h}, fold_se ? FLAG_RES_SE_SCALE : 0)
        .add(SF_BUF_Q3, out_slice(w4, 128 * h, 128 * h + 128), 0, 1);
  se(1, 3);
  StageBuild& q5 = stage("q5", SF_EPI_SAMPLE, T.vec(m + "4.conv.bias", 2 * C), {SF_BUF_X}, {0});
  q5.add(Y2, T.conv(m + "4.conv.weight", 2 * C, 2 * C, 3), 0, 1);
  q5.fold_se = fold_se ? 1 : -1;
  if (!T.missing.empty()) return ofail(SF_ERR_INVALID, "p_model weights: tensor missing or of unexpected size: " + T.missing);
  for (StageBuild& s : L)
    if (s.se_layer < 0) pack_stage(s, x3);
  return SF_OK;
}

size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// buffer widths in units of C (engine.py _BUF_CHANNELS)
int buf_mult(int buf) {
  switch (buf) {
    case SF_BUF_Z1: case SF_BUF_Y1: case SF_BUF_Q3: case SF_BUF_Z2: case SF_BUF_Y2: return 2;
    case SF_BUF_OBS: case SF_BUF_ZERO: return 0;          // sized separately
    default: return 1;
  }
}

struct Carve {                      // bump allocator over the caller's workspace (base may be null: size query)
  char* base;
  size_t off = 0;
  explicit Carve(void* b) : base(reinterpret_cast<char*>(b)) {}
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += align_up(bytes);
    return p;
  }
};

}  // namespace sfo
using namespace sfo;

struct sf_ode {
  sf_geometry g;
  sf_ode_options o;
  sf_plan* plan = nullptr;
  void* act_hi[SF_BUF_COUNT] = {};
  void* act_lo[SF_BUF_COUNT] = {};
  int act_ch[SF_BUF_COUNT] = {};
  int act_n[SF_BUF_COUNT] = {};
  void* tensors[9] = {};
  size_t tensor_bytes[9] = {};
  void* a32 = nullptr;
  void* b32 = nullptr;
  void* se = nullptr;
  size_t se_bytes = 0;
  size_t used = 0;
};

namespace sfo {

struct PackedSet {
  std::unique_ptr<sf_packed> cell[2], prior;
};

// Lays the workspace out; with ode != nullptr also records the addresses.  packed may be null (size query: weight sizes are a
// function of C / precision / options only, computed from a zero-weight pack).
size_t layout(const sf_geometry& g, const sf_ode_options& o, const PackedSet* ps, void* base, sf_ode* ode, std::vector<void*>* wptrs) {
  Carve cv(base);
  const bool x3 = g.precision == SF_PREC_BF16X3, fold = (o.pack_options & SF_PACK_FOLD_SE) != 0;
  const size_t B = g.max_images, HW = (size_t)g.H * g.W, C = g.C;
  for (int buf = 0; buf < SF_BUF_COUNT; ++buf) {
    size_t n = B, ch = C * buf_mult(buf);
    if (buf == SF_BUF_OBS) { n = o.obs_images > 0 ? o.obs_images : 1; ch = C; }
    if (buf == SF_BUF_ZERO) { n = 1; ch = C; }
    if (fold && (buf == SF_BUF_Y1 || buf == SF_BUF_Y2)) continue;       // SE outputs are never materialised when the layers are folded
    void* hi = cv.take(n * HW * ch * 2);
    void* lo = x3 ? cv.take(n * HW * ch * 2) : nullptr;
    if (ode) { ode->act_hi[buf] = hi; ode->act_lo[buf] = lo; ode->act_ch[buf] = (int)ch; ode->act_n[buf] = (int)n; }
  }
  auto f32 = [&](size_t elems) { return cv.take(elems * 4); };
  void* s0 = f32(B * HW * C);
  void* s1 = f32(B * HW * C);
  void* a32 = f32(B * HW * C);
  void* b32 = f32(B * HW * C);
  const size_t se_elems = 2 * B * SF_SE_MAX_PARTIALS * 2 * C + 2 * B * 2 * C + 2 * B;
  void* se = f32(se_elems);
  void* x32 = f32(B * HW * C);
  void* params = f32(B * HW * 2 * C);
  void* err = cv.take(256);
  const size_t path_n = o.path_slots > 0 ? o.path_slots : 1, eps_n = o.eps_slots > 0 ? o.eps_slots : 1;
  void* path = f32(path_n * HW * C);
  void* eps = f32(eps_n * HW * C);
  if (ode) {
    void* t[9] = {ode->act_hi[SF_BUF_OBS], ode->act_lo[SF_BUF_OBS], eps, path, s0, s1, x32, params, err};
    const size_t obs_b = (size_t)ode->act_n[SF_BUF_OBS] * HW * C * 2;
    size_t tb[9] = {obs_b, x3 ? obs_b : 0, eps_n * HW * C * 4, path_n * HW * C * 4, B * HW * C * 4, B * HW * C * 4, B * HW * C * 4, B * HW * 2 * C * 4, 4};
    for (int i = 0; i < 9; ++i) { ode->tensors[i] = t[i]; ode->tensor_bytes[i] = tb[i]; }
    ode->a32 = a32; ode->b32 = b32; ode->se = se; ode->se_bytes = se_elems * 4;
  }
  // weights: per stage w, vec [, w32, row_meta, scaled]; per SE layer fc1, fc2
  if (ps) {
    const sf_packed* sets[3] = {ps->cell[0].get(), ps->cell[1].get(), ps->prior.get()};
    for (const sf_packed* s : sets)
      for (const StageBuild& st : s->items) {
        if (st.se_layer >= 0) {
          void* a = cv.take(st.fc1.size() * 4);
          void* b = cv.take(st.fc2.size() * 4);
          if (wptrs) { wptrs->push_back(a); wptrs->push_back(b); }
          continue;
        }
        void* w = cv.take(st.w.size() * 2);
        void* v = cv.take(st.vec.size() * 4 + 4);
        if (wptrs) { wptrs->push_back(w); wptrs->push_back(v); }
        if (st.fold_se >= 0) {
          void* w32 = cv.take(st.w32.size() * 4);
          void* meta = cv.take(st.row_meta.size() * 4);
          void* scaled = cv.take(B * st.w.size() * 2);
          if (wptrs) { wptrs->push_back(w32); wptrs->push_back(meta); wptrs->push_back(scaled); }
        }
      }
  }
  return cv.off;
}

// zero weights of the right shapes: the packed sizes depend only on C, the precision and the options
int pack_all(const sf_tensor* tensors, int n, const std::string& prefix, const sf_geometry& g, const sf_ode_options& o, PackedSet* ps) {
  const char* cells[2] = {"gru_c.", "gru_obs.gru_d."};
  for (int k = 0; k < 2; ++k) {
    sf_packed* p = nullptr;
    int rc = sf_pack_cell_weights(tensors, n, (prefix + cells[k]).c_str(), g.precision, o.pack_options, &p);
    if (rc) return rc;
    ps->cell[k].reset(p);
    if (p->C != g.C) return ofail(SF_ERR_INVALID, "the cell weights have " + std::to_string(p->C) + " hidden channels, the geometry says " + std::to_string(g.C));
  }
  sf_packed* p = nullptr;
  int rc = sf_pack_pmodel_weights(tensors, n, (prefix + "p_model.").c_str(), g.precision, o.pack_options, &p);
  if (rc) return rc;
  ps->prior.reset(p);
  if (p->C != g.C) return ofail(SF_ERR_INVALID, "the p_model weights have " + std::to_string(p->C) + " hidden channels, the geometry says " + std::to_string(g.C));
  return SF_OK;
}

struct ZeroWeights {                 // a complete, zero-valued parameter set of width C (size queries)
  std::vector<std::string> names;
  std::vector<std::vector<float>> data;
  std::vector<sf_tensor> t;
  explicit ZeroWeights(int C) {
    auto add = [&](const std::string& n, size_t numel) { names.push_back(n); data.emplace_back(numel, 0.0f); };
    for (const char* cell : {"gru_c.", "gru_obs.gru_d."}) {
      const std::string p = cell;
      for (const char* k : {"conv_update_1", "conv_reset_1", "conv_state_tilde_1", "conv_update_2", "conv_reset_2", "conv_state_tilde_2"}) {
        add(p + k + ".weight", (size_t)C * 2 * C * 9);
        add(p + k + ".bias", C);
      }
      add(p + "conv_decoder_2.weight", (size_t)C * C * 9);
      add(p + "conv_decoder_2.bias", C);
      const std::string t0 = p + "trusting_gate.0.";
      add(t0 + "layers.0.weight", (size_t)C * 2 * C * 49);
      for (int i : {1, 4, 7}) { add(t0 + "layers." + std::to_string(i) + ".weight", C); add(t0 + "layers." + std::to_string(i) + ".bias", C); }
      add(t0 + "layers.3.weight", (size_t)C * C);
      add(t0 + "layers.6.weight", (size_t)C * C * 9);
      add(t0 + "projection.0.weight", (size_t)C * 2 * C);
      add(p + "trusting_gate.1.weight", 2 * C);
    }
    const std::string m = "p_model.model.";
    auto block = [&](const std::string& b, int n, int cin) {
      add(b + ".conv.weight", (size_t)n * cin * 9);
      for (const char* k : {"weight", "bias", "running_mean", "running_var"}) add(b + ".norm." + k, n);
    };
    block(m + "0.layers.conv_1", C, C);
    block(m + "0.layers.conv_2", 2 * C, C);
    add(m + "0.projection.weight", (size_t)2 * C * C);
    add(m + "0.projection.bias", 2 * C);
    for (int i : {1, 3}) { add(m + std::to_string(i) + ".fc.0.weight", (size_t)(2 * C / 8) * 2 * C); add(m + std::to_string(i) + ".fc.2.weight", (size_t)(2 * C / 8) * 2 * C); }
    block(m + "2.layers.conv_1", 2 * C, 2 * C);
    block(m + "2.layers.conv_2", 2 * C, 2 * C);
    add(m + "4.conv.weight", (size_t)2 * C * 2 * C * 9);
    add(m + "4.conv.bias", 2 * C);
    for (size_t i = 0; i < names.size(); ++i) t.push_back(sf_tensor{names[i].c_str(), data[i].data(), (int64_t)data[i].size()});
  }
};

sf_ode_options default_options(const sf_ode_options* o) {
  sf_ode_options r;
  r.path_slots = 1; r.obs_images = 1; r.eps_slots = 1; r.pack_options = SF_PACK_DEFAULT;
  if (o) r = *o;
  return r;
}

#define ODE_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return ofail(SF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)
// a failing sf_plan_* call has already set the library's error string: pass its code through
#define ODE_RC(call)            \
  do {                          \
    int rc_ = (call);           \
    if (rc_ < 0) return rc_;    \
  } while (0)

}  // namespace sfo
using namespace sfo;

extern "C" {

int sf_pack_cell_weights(const sf_tensor* tensors, int n_tensors, const char* prefix, int precision, int options, sf_packed** out) {
  if (!tensors || !out || n_tensors <= 0) return ofail(SF_ERR_INVALID, "sf_pack_cell_weights: null argument");
  if (precision != SF_PREC_BF16 && precision != SF_PREC_BF16X3) return ofail(SF_ERR_INVALID, "bad precision mode");
  TensorTable T(tensors, n_tensors, "");
  std::unique_ptr<sf_packed> p(new sf_packed());
  int rc = build_cell(T, prefix ? prefix : "", precision == SF_PREC_BF16X3, options, p.get());
  if (rc) return rc;
  *out = p.release();
  return SF_OK;
}

int sf_pack_pmodel_weights(const sf_tensor* tensors, int n_tensors, const char* prefix, int precision, int options, sf_packed** out) {
  if (!tensors || !out || n_tensors <= 0) return ofail(SF_ERR_INVALID, "sf_pack_pmodel_weights: null argument");
  if (precision != SF_PREC_BF16 && precision != SF_PREC_BF16X3) return ofail(SF_ERR_INVALID, "bad precision mode");
  TensorTable T(tensors, n_tensors, "");
  std::unique_ptr<sf_packed> p(new sf_packed());
  int rc = build_prior(T, prefix ? prefix : "", precision == SF_PREC_BF16X3, options, p.get());
  if (rc) return rc;
  *out = p.release();
  return SF_OK;
}

int sf_packed_count(const sf_packed* p) { return p ? (int)p->items.size() : ofail(SF_ERR_INVALID, "null packed object"); }

int sf_packed_get(const sf_packed* p, int i, sf_stage_desc* d) {
  if (!p || !d || i < 0 || i >= (int)p->items.size()) return ofail(SF_ERR_INVALID, "sf_packed_get: bad index");
  const StageBuild& s = p->items[i];
  memset(d, 0, sizeof(*d));
  d->name = s.name.c_str();
  d->se_layer = s.se_layer;
  d->fold_se = -1;
  if (s.se_layer >= 0) {
    d->fc1 = s.fc1.data();
    d->fc2 = s.fc2.data();
    d->n_fc = (int32_t)s.fc1.size();
    return SF_OK;
  }
  d->epilogue = s.epilogue; d->flags = s.flags;
  d->n_chunks = (int)s.pchunks.size(); d->chunks = s.pchunks.data();
  d->w = s.w.data(); d->w_rows = s.w_rows;
  d->vec = s.vec.data(); d->n_vec = (int)s.vec.size();
  d->n_io = (int)s.io.size(); d->io = s.io.data(); d->io_off = s.io_off.data();
  d->fold_se = s.fold_se;
  d->w32 = s.fold_se >= 0 ? s.w32.data() : nullptr;
  d->row_meta = s.fold_se >= 0 ? s.row_meta.data() : nullptr;
  return SF_OK;
}

int sf_packed_free(sf_packed* p) {
  delete p;
  return SF_OK;
}

int sf_ode_query_workspace(const sf_geometry* g, const sf_ode_options* o, size_t* bytes) {
  if (!g || !bytes) return ofail(SF_ERR_INVALID, "sf_ode_query_workspace: null argument");
  if (g->C != 64 && g->C != 128) return ofail(SF_ERR_INVALID, "hidden channels must be 64 or 128");
  const sf_ode_options opt = default_options(o);
  ZeroWeights z(g->C);
  PackedSet ps;
  int rc = pack_all(z.t.data(), (int)z.t.size(), "", *g, opt, &ps);
  if (rc) return rc;
  *bytes = layout(*g, opt, &ps, nullptr, nullptr, nullptr);
  return SF_OK;
}

int sf_ode_create(const sf_geometry* g, const sf_ode_options* o, const sf_tensor* tensors, int n_tensors, const char* prefix, void* workspace,
                  size_t workspace_bytes, sf_ode** out) {
  if (!g || !tensors || !workspace || !out) return ofail(SF_ERR_INVALID, "sf_ode_create: null argument");
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return ofail(SF_ERR_INVALID, "the workspace must be 256-byte aligned");
  const sf_ode_options opt = default_options(o);
  PackedSet ps;
  ODE_RC(pack_all(tensors, n_tensors, prefix ? prefix : "", *g, opt, &ps));
  std::unique_ptr<sf_ode> ode(new sf_ode());
  ode->g = *g;
  ode->o = opt;
  std::vector<void*> wp;
  ode->used = layout(*g, opt, &ps, workspace, ode.get(), &wp);
  if (ode->used > workspace_bytes)
    return ofail(SF_ERR_INVALID, "workspace too small: " + std::to_string(workspace_bytes) + " bytes given, " + std::to_string(ode->used) + " needed");
  ODE_RC(sf_plan_create(g, &ode->plan));
  std::unique_ptr<sf_plan, int (*)(sf_plan*)> plan_guard(ode->plan, sf_plan_destroy);       // destroyed on any early return
  ODE_CUDA(cudaSetDevice(g->device));
  ODE_CUDA(cudaMemset(workspace, 0, ode->used));        // zero state, ZERO buffer, SE block counters, error word
  for (int buf = 0; buf < SF_BUF_COUNT; ++buf)
    if (ode->act_hi[buf]) ODE_RC(sf_plan_bind_act(ode->plan, buf, ode->act_hi[buf], ode->act_lo[buf], ode->act_ch[buf], ode->act_n[buf]));
  ODE_RC(sf_plan_bind_f32(ode->plan, SF_F32_STATE0, ode->tensors[SF_ODE_STATE0]));
  ODE_RC(sf_plan_bind_f32(ode->plan, SF_F32_STATE1, ode->tensors[SF_ODE_STATE1]));
  ODE_RC(sf_plan_bind_f32(ode->plan, SF_F32_A, ode->a32));
  ODE_RC(sf_plan_bind_f32(ode->plan, SF_F32_B, ode->b32));
  ODE_RC(sf_plan_bind_f32(ode->plan, SF_F32_SE_SUMS, ode->se));
  ODE_RC(sf_plan_bind_f32(ode->plan, SF_F32_X, ode->tensors[SF_ODE_X32]));
  ODE_RC(sf_plan_bind_f32(ode->plan, SF_F32_PARAMS, ode->tensors[SF_ODE_PARAMS32]));
  ODE_RC(sf_plan_bind_f32(ode->plan, SF_F32_ERRFLAG, ode->tensors[SF_ODE_ERRFLAG]));
  ODE_RC(sf_plan_bind_f32(ode->plan, SF_F32_PATH, ode->tensors[SF_ODE_PATH]));
  ODE_RC(sf_plan_bind_f32(ode->plan, SF_F32_EPS, ode->tensors[SF_ODE_EPS]));
  // upload the packed weights and define the stages: slots [0, n_cell) derivative cell, [n_cell, 2 n_cell) jump cell, then the prior net
  const bool fold = (opt.pack_options & SF_PACK_FOLD_SE) != 0;
  size_t wi = 0;
  int slot = 0;
  std::vector<int32_t> graph[3];
  const sf_packed* sets[3] = {ps.cell[0].get(), ps.cell[1].get(), ps.prior.get()};
  for (int k = 0; k < 3; ++k)
    for (const StageBuild& st : sets[k]->items) {
      if (st.se_layer >= 0) {
        void* fc1 = wp[wi++];
        void* fc2 = wp[wi++];
        ODE_CUDA(cudaMemcpy(fc1, st.fc1.data(), st.fc1.size() * 4, cudaMemcpyHostToDevice));
        ODE_CUDA(cudaMemcpy(fc2, st.fc2.data(), st.fc2.size() * 4, cudaMemcpyHostToDevice));
        const int in_buf = st.se_layer ? SF_BUF_Z2 : SF_BUF_Z1, out_buf = st.se_layer ? SF_BUF_Y2 : SF_BUF_Y1;
        ODE_RC(sf_plan_define_se(ode->plan, st.se_layer, reinterpret_cast<const float*>(fc1), reinterpret_cast<const float*>(fc2), in_buf, out_buf));
        graph[k].push_back((fold ? 2000 : 1000) + st.se_layer);
        continue;
      }
      void* w = wp[wi++];
      void* v = wp[wi++];
      ODE_CUDA(cudaMemcpy(w, st.w.data(), st.w.size() * 2, cudaMemcpyHostToDevice));
      if (!st.vec.empty()) ODE_CUDA(cudaMemcpy(v, st.vec.data(), st.vec.size() * 4, cudaMemcpyHostToDevice));
      if (slot >= SF_MAX_STAGES) return ofail(SF_ERR_INVALID, "too many stages");
      ODE_RC(sf_plan_define_stage(ode->plan, slot, st.epilogue, (int)st.pchunks.size(), st.pchunks.data(), w, st.w_rows, reinterpret_cast<const float*>(v),
                                  (int)st.vec.size(), st.io.data(), st.io_off.data(), (int)st.io.size(), st.flags));
      if (st.fold_se >= 0) {
        void* w32 = wp[wi++];
        void* meta = wp[wi++];
        void* scaled = wp[wi++];
        ODE_CUDA(cudaMemcpy(w32, st.w32.data(), st.w32.size() * 4, cudaMemcpyHostToDevice));
        ODE_CUDA(cudaMemcpy(meta, st.row_meta.data(), st.row_meta.size() * 4, cudaMemcpyHostToDevice));
        ODE_RC(sf_plan_define_stage_fold(ode->plan, slot, st.fold_se, reinterpret_cast<const float*>(w32), reinterpret_cast<const int32_t*>(meta), scaled));
      }
      graph[k].push_back(slot++);
    }
  if (graph[0].size() != graph[1].size()) return ofail(SF_ERR_STATE, "the two cells packed to different stage counts");
  ODE_RC(sf_plan_define_event_graph(ode->plan, graph[0].data(), graph[1].data(), (int)graph[0].size(), graph[2].data(), (int)graph[2].size()));
  ODE_RC(sf_plan_finalize(ode->plan));
  ODE_CUDA(cudaDeviceSynchronize());
  plan_guard.release();
  *out = ode.release();
  return SF_OK;
}

int sf_ode_destroy(sf_ode* ode) {
  if (!ode) return SF_OK;
  if (ode->plan) sf_plan_destroy(ode->plan);
  delete ode;
  return SF_OK;
}

int sf_ode_plan(sf_ode* ode, sf_plan** plan) {
  if (!ode || !plan) return ofail(SF_ERR_INVALID, "sf_ode_plan: null argument");
  *plan = ode->plan;
  return SF_OK;
}

int sf_ode_tensor(sf_ode* ode, int which, void** ptr, size_t* bytes) {
  if (!ode || which < 0 || which > SF_ODE_ERRFLAG) return ofail(SF_ERR_INVALID, "sf_ode_tensor: bad tensor id");
  if (ptr) *ptr = ode->tensors[which];
  if (bytes) *bytes = ode->tensor_bytes[which];
  return SF_OK;
}

int sf_ode_set_observations(sf_ode* ode, const float* obs_nchw, int first_image, int n_images, void* stream) {
  if (!ode || !obs_nchw || first_image < 0 || n_images <= 0 || first_image + n_images > ode->act_n[SF_BUF_OBS])
    return ofail(SF_ERR_INVALID, "sf_ode_set_observations: image range exceeds the OBS buffer");
  const size_t off = (size_t)first_image * ode->g.H * ode->g.W * ode->g.C * 2;
  char* hi = reinterpret_cast<char*>(ode->act_hi[SF_BUF_OBS]) + off;
  char* lo = ode->act_lo[SF_BUF_OBS] ? reinterpret_cast<char*>(ode->act_lo[SF_BUF_OBS]) + off : nullptr;
  return sf_pack_nchw_f32(obs_nchw, hi, lo, n_images, ode->g.C, ode->g.H, ode->g.W, stream);
}

int sf_ode_reset_state(sf_ode* ode, void* stream) {
  if (!ode) return ofail(SF_ERR_INVALID, "sf_ode_reset_state: null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const size_t n32 = ode->tensor_bytes[SF_ODE_STATE0], n16 = n32 / 2;
  ODE_CUDA(cudaMemsetAsync(ode->tensors[SF_ODE_STATE0], 0, n32, s));
  ODE_CUDA(cudaMemsetAsync(ode->tensors[SF_ODE_STATE1], 0, n32, s));
  for (int buf : {SF_BUF_S0, SF_BUF_S1}) {
    ODE_CUDA(cudaMemsetAsync(ode->act_hi[buf], 0, n16, s));
    if (ode->act_lo[buf]) ODE_CUDA(cudaMemsetAsync(ode->act_lo[buf], 0, n16, s));
  }
  return SF_OK;
}

int sf_ode_event(sf_ode* ode, const sf_event* ev, const int32_t* table, void* stream) {
  if (!ode) return ofail(SF_ERR_INVALID, "sf_ode_event: null argument");
  return sf_plan_run_events(ode->plan, ev, 1, table, stream);
}

int sf_ode_rollout(sf_ode* ode, const sf_event* evs, int n_events, const int32_t* table, void* stream) {
  if (!ode) return ofail(SF_ERR_INVALID, "sf_ode_rollout: null argument");
  return sf_plan_run_events(ode->plan, evs, n_events, table, stream);
}

int sf_ode_read_path(sf_ode* ode, const int32_t* slots, int n, float* out_nchw, void* stream) {
  if (!ode || !slots || !out_nchw || n <= 0) return ofail(SF_ERR_INVALID, "sf_ode_read_path: bad argument");
  return sf_unpack_nhwc_f32(reinterpret_cast<const float*>(ode->tensors[SF_ODE_PATH]), out_nchw, slots, n, ode->g.C, ode->g.H, ode->g.W, stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// host schedule: timestamps -> per-sample op lists -> batched events + event table (schedule.py / rollout.py restated in C++)
// ---------------------------------------------------------------------------------------------------------------
struct sf_rollout_plan {
  std::vector<sf_event> events;
  std::vector<int32_t> table, out_slots;
  sf_rollout_info info;
};

namespace sfo {

struct Op {
  int kind;      // 0 step, 1 jump
  double dt;
  int obs;
};

// One sample's op list and, per target, the op whose resulting state is the output frame.  T1 / T2: float or double -- the dtype of
// the caller's timestamp tensors: the reference keeps current_time as a python double but compares / subtracts it against 0-dim
// tensors of the stamps' own dtype (temporal_ode_bayes.py:540-545, 586-590); a fixed step (two python floats) stays double.
template <typename T1, typename T2>
void plan_sample(const double* obs, int n_obs, const double* targets, int n_targets, double delta_t, bool variable, std::vector<Op>& ops,
                 std::vector<int>& picks) {
  double now = obs[0];
  for (int i = 1; i < n_obs; ++i) now = obs[i] < now ? obs[i] : now;
  std::vector<double> stamps;
  std::vector<int> stamp_op;
  const double half = 0.5 * delta_t;
  for (int k = 0; k < n_obs; ++k) {
    const T1 t1 = (T1)obs[k];
    while ((T1)now <= (T1)(t1 - (T1)delta_t)) {
      double h;
      if (variable) {
        const T1 hh = (T1)(t1 - (T1)now);
        now = (double)(T1)((T1)now + hh);
        h = (double)hh;
      } else {
        h = delta_t;
        now = now + h;
      }
      ops.push_back(Op{0, h, -1});
    }
    ops.push_back(Op{1, 0.0, k});
    stamps.push_back(obs[k]);
    stamp_op.push_back((int)ops.size() - 1);
  }
  for (int j = 0; j < n_targets; ++j) {
    const T2 t2 = (T2)targets[j];
    while ((T2)now < t2) {
      double h;
      if (variable) {
        const T2 hh = (T2)(t2 - (T2)now);
        now = (double)(T2)((T2)now + hh);
        h = (double)hh;
      } else {
        h = delta_t;
        now = now + h;
      }
      ops.push_back(Op{0, h, -1});
      if ((T2)(t2 - (T2)half) < (T2)now && (T2)now < (T2)(t2 + (T2)half)) {
        stamps.push_back(now);
        stamp_op.push_back((int)ops.size() - 1);
      }
    }
  }
  // selection (:606-620): latest recorded stamp strictly inside (t - half, t + half), else the nearest one (first on ties)
  for (int j = 0; j < n_targets; ++j) {
    const double when = targets[j];
    int best = -1;
    for (int i = 0; i < (int)stamps.size(); ++i)
      if (when - half < stamps[i] && stamps[i] < when + half) best = i;
    if (best < 0) {
      double bd = 0.0;
      for (int i = 0; i < (int)stamps.size(); ++i) {
        const double d = std::fabs(stamps[i] - when);
        if (best < 0 || d < bd) { best = i; bd = d; }
      }
    }
    picks.push_back(stamp_op[best]);
  }
}

struct SampleEvent {
  int kind, x_buf, x_img, s_in, s_base, s_out, eps, rec, run_prior;
  double dt;
};

}  // namespace sfo
using namespace sfo;

extern "C" {

int sf_merge_observations(const double* camera_t, int n_cam, const double* lidar_t, int n_lidar, double* times, int32_t* source) {
  if ((!camera_t && n_cam > 0) || (!lidar_t && n_lidar > 0) || n_cam < 0 || n_lidar < 0 || !times || !source)
    return ofail(SF_ERR_INVALID, "sf_merge_observations: bad argument");
  std::vector<std::pair<double, int32_t>> items;
  for (int i = 0; i < n_cam; ++i) items.push_back(std::make_pair(camera_t[i], (int32_t)i));
  for (int i = 0; i < n_lidar; ++i) items.push_back(std::make_pair(lidar_t[i], (int32_t)(65536 + i)));
  std::stable_sort(items.begin(), items.end(), [](const std::pair<double, int32_t>& a, const std::pair<double, int32_t>& b) { return a.first < b.first; });
  for (size_t i = 0; i < items.size(); ++i) { times[i] = items[i].first; source[i] = items[i].second; }
  return (int)items.size();
}

int sf_rollout_plan_create(const double* obs_times, int n_obs, const double* targets, int n_targets, int B, double delta_t, int variable_step,
                           int solver, int impute, int obs_f32, int target_f32, int flags, sf_rollout_plan** out) {
  if (!obs_times || !targets || !out || n_obs <= 0 || n_targets <= 0 || B <= 0) return ofail(SF_ERR_INVALID, "sf_rollout_plan_create: bad argument");
  if (solver != 0 && solver != 1) return ofail(SF_ERR_INVALID, "Unknown solver");        // temporal_ode_bayes.py:386
  if (!(delta_t > 0.0)) return ofail(SF_ERR_INVALID, "delta_t must be positive");
  const bool all_prior = (flags & 1) != 0, keep_last = (flags & 2) != 0;
  std::unique_ptr<sf_rollout_plan> r(new sf_rollout_plan());
  memset(&r->info, 0, sizeof(r->info));
  std::vector<std::vector<SampleEvent>> per_sample(B);
  int eps = 0, n_path = 0;
  r->out_slots.resize((size_t)B * n_targets);
  for (int b = 0; b < B; ++b) {
    std::vector<Op> ops;
    std::vector<int> picks;
    const double* ob = obs_times + (size_t)b * n_obs;
    const double* tg = targets + (size_t)b * n_targets;
    if (obs_f32 && target_f32) plan_sample<float, float>(ob, n_obs, tg, n_targets, delta_t, variable_step != 0, ops, picks);
    else if (obs_f32) plan_sample<float, double>(ob, n_obs, tg, n_targets, delta_t, variable_step != 0, ops, picks);
    else if (target_f32) plan_sample<double, float>(ob, n_obs, tg, n_targets, delta_t, variable_step != 0, ops, picks);
    else plan_sample<double, double>(ob, n_obs, tg, n_targets, delta_t, variable_step != 0, ops, picks);
    std::map<int, int> picked;
    for (int j = 0; j < n_targets; ++j) {
      auto it = picked.find(picks[j]);
      if (it == picked.end()) it = picked.insert(std::make_pair(picks[j], n_path++)).first;
      r->out_slots[(size_t)b * n_targets + j] = it->second;
    }
    const int n_ops = (int)ops.size();
    for (int i = 0; i < n_ops; ++i) {
      const Op& op = ops[i];
      auto pk = picked.find(i);
      const int rec = pk == picked.end() ? -1 : pk->second;
      // the input sampled after this op is read only by a following ode_step (GRUObservationCell ignores p, tob:327-344)
      const bool live = (i + 1 < n_ops) ? ops[i + 1].kind == 0 : keep_last;
      const int prior = (impute && (live || all_prior)) ? 1 : 0;
      if (op.kind == 1) {
        per_sample[b].push_back(SampleEvent{1, SF_BUF_OBS, b * n_obs + op.obs, 0, 0, 0, eps, rec, prior, 0.0});
        eps += 1;
        r->info.n_jumps += 1;
      } else {
        const int xb = impute ? SF_BUF_X : SF_BUF_ZERO, xi = impute ? b : 0;
        if (solver == 0) {
          per_sample[b].push_back(SampleEvent{0, xb, xi, 0, 0, 0, eps, rec, prior, op.dt});
          eps += 1;
        } else {       // midpoint: k = s + dt/2 f(x, s); pk = infer(k) | s = s + dt f(pk, k); x = infer(s)      (tob:449-454)
          per_sample[b].push_back(SampleEvent{0, xb, xi, 0, 0, 1, eps, -1, 1, op.dt / 2});
          per_sample[b].push_back(SampleEvent{0, SF_BUF_X, b, 1, 0, 0, eps + 1, rec, prior, op.dt});
          eps += 2;
        }
        r->info.n_state_steps += 1;
      }
    }
  }
  r->info.n_eps = eps;
  r->info.n_path = n_path;
  size_t depth = 0;
  for (const auto& e : per_sample) depth = e.size() > depth ? e.size() : depth;
  for (size_t round = 0; round < depth; ++round) {
    // samples whose round-th events agree in (kind, x source, state buffers, prior) share one set of launches; groups keep the
    // order in which their first member appears (python dict order)
    std::vector<std::vector<int>> groups;
    std::vector<SampleEvent> keys;
    for (int b = 0; b < B; ++b) {
      if (round >= per_sample[b].size()) continue;
      const SampleEvent& e = per_sample[b][round];
      size_t gi = 0;
      for (; gi < keys.size(); ++gi) {
        const SampleEvent& k = keys[gi];
        if (k.kind == e.kind && k.x_buf == e.x_buf && k.s_in == e.s_in && k.s_base == e.s_base && k.s_out == e.s_out && k.run_prior == e.run_prior) break;
      }
      if (gi == keys.size()) { keys.push_back(e); groups.emplace_back(); }
      groups[gi].push_back(b);
    }
    for (size_t gi = 0; gi < groups.size(); ++gi) {
      const SampleEvent& k = keys[gi];
      const int n = (int)groups[gi].size();
      sf_event ev;
      ev.kind = k.kind; ev.n_active = n; ev.x_buf = k.x_buf; ev.s_in = k.s_in; ev.s_base = k.s_base; ev.s_out = k.s_out;
      ev.run_cell = 1; ev.run_prior = k.run_prior; ev.want_f32 = 0; ev.table_off = (int32_t)r->table.size();
      r->events.push_back(ev);
      r->info.n_cell_evals += n;
      r->info.n_prior_evals += n * k.run_prior;
      const size_t base = r->table.size();
      r->table.resize(base + 5 * (size_t)n);
      for (int i = 0; i < n; ++i) {
        const int b = groups[gi][i];
        const SampleEvent& e = per_sample[b][round];
        r->table[base + i] = b;
        r->table[base + n + i] = e.x_img;
        r->table[base + 2 * n + i] = e.rec;
        r->table[base + 3 * n + i] = e.eps;
        const float dtf = (float)e.dt;
        int32_t bits;
        memcpy(&bits, &dtf, 4);
        r->table[base + 4 * n + i] = bits;
      }
    }
  }
  r->info.n_events = (int)r->events.size();
  r->info.n_table = (int)r->table.size();
  *out = r.release();
  return SF_OK;
}

int sf_rollout_plan_info(const sf_rollout_plan* r, sf_rollout_info* info) {
  if (!r || !info) return ofail(SF_ERR_INVALID, "sf_rollout_plan_info: null argument");
  *info = r->info;
  return SF_OK;
}
const sf_event* sf_rollout_plan_events(const sf_rollout_plan* r) { return r ? r->events.data() : nullptr; }
const int32_t* sf_rollout_plan_table(const sf_rollout_plan* r) { return r ? r->table.data() : nullptr; }
const int32_t* sf_rollout_plan_out_slots(const sf_rollout_plan* r) { return r ? r->out_slots.data() : nullptr; }
int sf_rollout_plan_free(sf_rollout_plan* r) {
  delete r;
  return SF_OK;
}

}  // extern "C"
