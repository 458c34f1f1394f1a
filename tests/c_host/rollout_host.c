/* rollout_host.c -- the GRU-ODE-Bayes rollout driven from C: this file includes include/sf_b200.h and the CUDA runtime API and
 * nothing else of the repository.  It is what a non-Python host of the reference's hot path (temporal_ode_bayes.py:479-627 without
 * the encoder / decoder) looks like on top of libsf_b200.so:
 *
 *   weights (fp32, reference state_dict names) + encoded observations + noise + per-sample time stamps
 *     -> sf_ode_query_workspace / cudaMalloc / sf_ode_create      pack the weights, carve the workspace, define the stages
 *     -> sf_rollout_plan_create                                    the reference's step schedule, batched into events
 *     -> sf_ode_set_observations, noise upload, sf_ode_reset_state
 *     -> sf_ode_rollout                                            every stage launch of the whole rollout, one call
 *     -> sf_ode_read_path                                          the selected latent states, fp32 NCHW
 *
 * Input file (tests/test_gpu_c_host.py writes it): int32 {C, h, w, B, n_obs, n_targets, n_eps, n_tensors}; double times[B][n_obs];
 * double targets[B][n_targets]; n_tensors x {int32 name_len, int32 numel, name bytes, float data[numel]};
 * float obs[B * n_obs][C][h][w]; float eps[n_eps][C][h][w].  Output file: float selected[B][n_targets][C][h][w].
 *
 * gcc -std=c99 -I include -I $CUDA/include tests/c_host/rollout_host.c -L streamingflow_b200 -l:libsf_b200.so -lcudart */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "sf_b200.h"

#define SF(call)                                                                         \
  do {                                                                                   \
    int rc_ = (call);                                                                    \
    if (rc_ < 0) {                                                                       \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, sf_last_error());              \
      return 2;                                                                          \
    }                                                                                    \
  } while (0)
#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                        \
      return 3;                                                                          \
    }                                                                                    \
  } while (0)

static int read_exact(FILE* f, void* dst, size_t bytes) { return fread(dst, 1, bytes, f) == bytes ? 0 : -1; }

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s input.bin output.bin\n", argv[0]);
    return 1;
  }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int32_t hd[8];
  if (read_exact(f, hd, sizeof(hd))) return 1;
  const int C = hd[0], h = hd[1], w = hd[2], B = hd[3], n_obs = hd[4], n_targets = hd[5], n_eps = hd[6], n_tensors = hd[7];
  double* times = (double*)malloc(sizeof(double) * B * n_obs);
  double* targets = (double*)malloc(sizeof(double) * B * n_targets);
  if (read_exact(f, times, sizeof(double) * B * n_obs) || read_exact(f, targets, sizeof(double) * B * n_targets)) return 1;
  sf_tensor* tensors = (sf_tensor*)calloc(n_tensors, sizeof(sf_tensor));
  for (int i = 0; i < n_tensors; ++i) {
    int32_t meta[2];
    if (read_exact(f, meta, sizeof(meta))) return 1;
    char* name = (char*)calloc(meta[0] + 1, 1);
    float* data = (float*)malloc(sizeof(float) * meta[1]);
    if (read_exact(f, name, meta[0]) || read_exact(f, data, sizeof(float) * meta[1])) return 1;
    tensors[i].name = name;
    tensors[i].data = data;
    tensors[i].numel = meta[1];
  }
  const size_t frame = (size_t)C * h * w;
  float* obs = (float*)malloc(sizeof(float) * frame * B * n_obs);
  float* eps = (float*)malloc(sizeof(float) * frame * n_eps);
  if (read_exact(f, obs, sizeof(float) * frame * B * n_obs) || read_exact(f, eps, sizeof(float) * frame * n_eps)) return 1;
  fclose(f);

  CU(cudaSetDevice(0));
  SF(sf_device_supported(0));
  /* the reference's schedule (variable-step Euler, IMPUTE on: config.py defaults of Prediction_LC_ODE_Variable) */
  sf_rollout_plan* plan = NULL;
  SF(sf_rollout_plan_create(times, n_obs, targets, n_targets, B, 0.05, 1, 0, 1, 0, 0, 0, &plan));
  sf_rollout_info info;
  SF(sf_rollout_plan_info(plan, &info));
  if (info.n_eps != n_eps) {
    fprintf(stderr, "the schedule consumes %d noise tensors, the file holds %d\n", info.n_eps, n_eps);
    return 1;
  }

  sf_geometry geo = {B, h, w, C, SF_PREC_BF16, 0};
  sf_ode_options opt = {info.n_path, B * n_obs, info.n_eps, SF_PACK_DEFAULT};
  size_t ws_bytes = 0;
  SF(sf_ode_query_workspace(&geo, &opt, &ws_bytes));
  void* ws = NULL;
  CU(cudaMalloc(&ws, ws_bytes));
  sf_ode* ode = NULL;
  SF(sf_ode_create(&geo, &opt, tensors, n_tensors, "", ws, ws_bytes, &ode));

  cudaStream_t stream;
  CU(cudaStreamCreate(&stream));
  float* d_obs = NULL;
  CU(cudaMalloc((void**)&d_obs, sizeof(float) * frame * B * n_obs));
  CU(cudaMemcpyAsync(d_obs, obs, sizeof(float) * frame * B * n_obs, cudaMemcpyHostToDevice, stream));
  SF(sf_ode_set_observations(ode, d_obs, 0, B * n_obs, stream));
  void* d_eps = NULL;
  size_t eps_bytes = 0;
  SF(sf_ode_tensor(ode, SF_ODE_EPS, &d_eps, &eps_bytes));
  CU(cudaMemcpyAsync(d_eps, eps, sizeof(float) * frame * n_eps, cudaMemcpyHostToDevice, stream));
  SF(sf_ode_reset_state(ode, stream));

  int32_t* d_table = NULL;
  CU(cudaMalloc((void**)&d_table, sizeof(int32_t) * (info.n_table + B * n_targets)));
  CU(cudaMemcpyAsync(d_table, sf_rollout_plan_table(plan), sizeof(int32_t) * info.n_table, cudaMemcpyHostToDevice, stream));
  int32_t* d_slots = d_table + info.n_table;
  CU(cudaMemcpyAsync(d_slots, sf_rollout_plan_out_slots(plan), sizeof(int32_t) * B * n_targets, cudaMemcpyHostToDevice, stream));
  SF(sf_ode_rollout(ode, sf_rollout_plan_events(plan), info.n_events, d_table, stream));

  float* d_out = NULL;
  const size_t out_elems = frame * B * n_targets;
  CU(cudaMalloc((void**)&d_out, sizeof(float) * out_elems));
  SF(sf_ode_read_path(ode, d_slots, B * n_targets, d_out, stream));
  float* out = (float*)malloc(sizeof(float) * out_elems);
  CU(cudaMemcpyAsync(out, d_out, sizeof(float) * out_elems, cudaMemcpyDeviceToHost, stream));
  int32_t errflag = 0;
  void* d_err = NULL;
  SF(sf_ode_tensor(ode, SF_ODE_ERRFLAG, &d_err, NULL));
  CU(cudaMemcpyAsync(&errflag, d_err, sizeof(errflag), cudaMemcpyDeviceToHost, stream));
  CU(cudaStreamSynchronize(stream));
  if (errflag) {
    fprintf(stderr, "device-side pipeline timeout, code 0x%x\n", errflag);
    return 4;
  }
  f = fopen(argv[2], "wb");
  if (!f || fwrite(out, sizeof(float), out_elems, f) != out_elems) return 1;
  fclose(f);
  printf("rollout_host: %d samples, %d events (%d state-steps, %d jumps, %d prior-net evaluations), workspace %.1f MB\n", B, info.n_events,
         info.n_state_steps, info.n_jumps, info.n_prior_evals, ws_bytes / 1048576.0);

  SF(sf_ode_destroy(ode));
  SF(sf_rollout_plan_free(plan));
  CU(cudaFree(d_out));
  CU(cudaFree(d_table));
  CU(cudaFree(d_obs));
  CU(cudaFree(ws));
  CU(cudaStreamDestroy(stream));
  return 0;
}
