/*
 * sf_b200.h -- C ABI of libsf_b200.so: the B200 (sm_100a) implementation of StreamingFlow's
 * GRU-ODE-Bayes BEV integration path.
 *
 * The reference has NO native code on this path: every FLOP is a torch.nn call made from Python
 *   streamingflow/layers/temporal_ode_bayes.py:92-161   DualGRUODECell.forward  (derivative)
 *   streamingflow/layers/temporal_ode_bayes.py:239-305  DualGRUCell.forward     (observation jump)
 *   streamingflow/layers/temporal_ode_bayes.py:436-477  ode_step / infer_state
 *   streamingflow/layers/res_models.py:150-180          SELayer / ConvNet (= p_model)
 *   streamingflow/layers/convolutions.py:283-380        LayerNorm(channels_first) / Bottleblock
 * so the "FFI" a maintainer binds is this header, loaded with ctypes from the Python module that
 * keeps the reference's nn.Module interface.  Plain pointers and sizes only; no
 * torch types cross this boundary.  All device pointers are owned by the caller (torch allocator);
 * the library allocates no persistent device memory.  Every call returns 0 on success or a negative
 * sf_status; sf_last_error() gives the message for the calling thread's last failure.
 *
 * Data layout: every activation is NHWC ("pixel-major"): [image][y][x][channel], channel fastest,
 * stored as bf16 (optionally a second bf16 plane holding the rounding residual: split mode).
 * The master copy of the ODE state and the two GRU branch outputs stay fp32 NHWC.
 */
#ifndef SF_B200_H_
#define SF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SF_ABI_VERSION 3

typedef enum {
  SF_OK = 0,
  SF_ERR_INVALID = -1,      /* bad argument / unsupported shape  (reference: Python assert, temporal_ode_bayes.py:101,249,386) */
  SF_ERR_CUDA = -2,         /* a CUDA runtime / driver call failed */
  SF_ERR_STATE = -3,        /* plan not finalised, buffer not bound, ... */
  SF_ERR_UNSUPPORTED = -4   /* device is not sm_100 */
} sf_status;

/* precision modes of the tensor-core operands (accumulation is always fp32 in TMEM) */
#define SF_PREC_BF16 0      /* bf16 operands: rel 1e-2 contract                              */
#define SF_PREC_BF16X3 1    /* hi/lo split bf16, 3 products: fp32-class, rel 1e-4 contract   */

/* fused epilogues (one per conv stage of an event; SURVEY.md 7.4 kernel map) */
typedef enum {
  SF_EPI_GATES = 0,     /* sigmoid gates: writes u1,u2,(1-r1)*s,(1-r2)*s            temporal_ode_bayes.py:135-140,150-155 */
  SF_EPI_PROPOSE = 1,   /* proposal + GRU blend: a = (1-u1)s+u1*s~1, h = (1-u2)s+u2*s~2          :143-146,158-161 */
  SF_EPI_DECODE = 2,    /* b = conv_decoder_2(h) + bias                                            :121 */
  SF_EPI_LNGELU = 3,    /* LayerNorm(C)+GELU (7x7 trunk conv, then the 1x1)                 convolutions.py:356-361 */
  SF_EPI_MIX = 4,       /* LN+GELU, +GELU(proj), 1x1->2, softmax, mix, Euler / jump update   :362-380, tob:124-131,446 */
  SF_EPI_BIAS_LRELU = 5,/* folded-BN bias + LeakyReLU(0.1)                                   res_models.py:42-49 */
  SF_EPI_RES_PROJ = 6,  /* LReLU(acc0+b0) + (acc1+b1): ResBlock with 1x1 projection           res_models.py:75-79 */
  SF_EPI_RES_ID = 7,    /* LReLU(acc+b) + residual input                                      res_models.py:79 */
  SF_EPI_SAMPLE = 8     /* bias+LReLU, loc/raw split, loc + (softplus(raw)+1e-8)*eps           model_utils.py:81-85,107-108 */
} sf_epilogue;

/* One K-chunk of an implicit-GEMM stage: 64 input channels of one activation buffer, all RxR taps. */
typedef struct {
  int32_t buf;      /* activation buffer id (see sf_plan_bind_act); -1 = the event's x source            */
  int32_t plane;    /* 0 = hi plane, 1 = lo (residual) plane                                             */
  int32_t c0;       /* first channel of the chunk inside the buffer                                       */
  int32_t R;        /* filter taps per side: 1, 3 or 7 (zero padding (R-1)/2)                             */
  int32_t n;        /* output channels accumulated by this chunk (multiple of 64, <= 256)                 */
  int32_t nrep;     /* weight tiles per tap applied to the same A tile and the same columns (1 or 2)      */
  int32_t col;      /* first TMEM accumulator column                                                      */
  int32_t wrow;     /* first row of this chunk in the packed weight matrix; tap (dx,dy) rep r starts at
                       wrow + ((dx*R+dy)*nrep + r)*n                                                     */
  int32_t init;     /* 1: the chunk's first MMA overwrites its columns instead of accumulating           */
  int32_t ox, oy;   /* pixel offset of the chunk's input window (dilated taps are R = 1 chunks at (kx-1)*d, (ky-1)*d) */
} sf_chunk;

#define SF_MAX_CHUNKS 40
#define SF_MAX_ACT_BUFS 32
#define SF_MAX_STAGES 32
#define SF_SE_MAX_PARTIALS 160   /* per-block partial channel sums of one SE layer and sample (SF_F32_SE_SUMS layout) */

typedef struct {
  int32_t max_images;   /* capacity (samples) of the per-sample buffers                                  */
  int32_t H, W;         /* ODE grid (latent) height / width                                               */
  int32_t C;            /* hidden channels: 64 or 128                                                     */
  int32_t precision;    /* SF_PREC_*                                                                      */
  int32_t device;       /* CUDA device ordinal                                                            */
} sf_geometry;

/* fp32 side tensors, all NHWC [image][H][W][C] unless noted */
typedef enum {
  SF_F32_STATE0 = 0,    /* master ODE state, buffer 0                                                     */
  SF_F32_STATE1 = 1,    /* second state buffer (midpoint stage k)                                         */
  SF_F32_A = 2,         /* rnn_state1 (GRU-1 output)                                                      */
  SF_F32_B = 3,         /* rnn_state2 (decoder of GRU-2)                                                  */
  SF_F32_PATH = 4,      /* recorded states [slot][H][W][C]                                                */
  SF_F32_SE_SUMS = 5,   /* SE scratch: partial sums [2][max_images][SF_SE_MAX_PARTIALS][2C] | scales [2][max_images][2C] |
                           uint32 block counters [2][max_images]; zero-initialised by the caller                */
  SF_F32_EPS = 6,       /* standard-normal noise, NCHW [slot][C][H][W] (torch's generation order)         */
  SF_F32_X = 7,         /* optional fp32 copy of the sampled input x  (infer_state API)                   */
  SF_F32_PARAMS = 8,    /* optional fp32 p_model output [image][H][W][2C] (infer_state API)               */
  SF_F32_COUNT = 9
} sf_f32_slot;
#define SF_F32_ERRFLAG 9   /* int32 device word: a bounded pipeline wait that times out stores a code here before trapping */
#define SF_F32_OUT 10      /* optional fp32 NHWC copy of a bias_act stage's output (stage flag bit 4)                       */
#define SF_F32_IMG_BIAS 11 /* optional per-image bias [image][n_out] added by a bias_act stage (stage flag bit 6)            */

/* One event = one pass of {cell stages} + optional {prior-network stages} over the listed samples. */
typedef struct {
  int32_t kind;          /* 0: ODE derivative step with the gru_c weights; 1: observation jump (gru_obs)  */
  int32_t n_active;      /* samples taking part                                                           */
  int32_t x_buf;         /* activation buffer holding the cell's x input (sampled input or encoded obs)    */
  int32_t s_in;          /* state buffer (0/1) the cell reads                                              */
  int32_t s_base;        /* state buffer the Euler update starts from (== s_in except midpoint stage 2)    */
  int32_t s_out;         /* state buffer the new state is written to                                       */
  int32_t run_cell;      /* 0 skips the cell stages (infer_state-only call)                                */
  int32_t run_prior;     /* 1 runs p_model + sampling on s_out and refreshes the x buffer                  */
  int32_t want_f32;      /* 1 also writes SF_F32_X / SF_F32_PARAMS                                         */
  int32_t table_off;     /* offset (int32 elements) of this event's rows in the device event table:
                            5 rows of n_active: sample id, x image index, record slot (-1 none),
                            eps slot, dt (float bits)                                                      */
} sf_event;

typedef struct sf_plan sf_plan;

int sf_abi_version(void);
const char* sf_last_error(void);
/* returns SF_OK when the current device can run the library (compute capability 10.x) */
int sf_device_supported(int device);

int sf_plan_create(const sf_geometry* g, sf_plan** out);
int sf_plan_destroy(sf_plan* p);
/* activation buffer: bf16 NHWC [n_images][H][W][channels]; lo may be NULL unless precision is BF16X3 */
int sf_plan_bind_act(sf_plan* p, int buf, void* hi, void* lo, int channels, int n_images);
int sf_plan_bind_f32(sf_plan* p, int slot, void* ptr);
/* packed weights: bf16 [w_rows][64] device pointer; vec: fp32 device pointer (bias / LN / gate weights) */
/* io_bufs / io_choff: the epilogue's activation buffers and the first channel each launch touches in them, in the order
   the epilogue expects (gates: u_0, gated_0[, u_1, gated_1]; propose: u_0[, u_1], out_0[, out_1]; res_id: residual, out;
   others: out).  flags: bit 0 (propose) also keep the blend in the fp32 tensor SF_F32_A; bits 1-3 (bias_act) activation:
   0 LeakyReLU(0.1), 1 tanh, 2 ReLU, 3 identity, 4 GELU(erf); bit 4 (bias_act) also write the output to SF_F32_OUT in fp32;
   bit 5 (gates / propose at C = 64) a single gate pair / proposal (plain ConvGRU of the refinement) instead of two;
   bit 6 (bias_act) add the per-image bias SF_F32_IMG_BIAS[image][n_out] (ASPP pooling branch);
   bit 9 (C = 64, n = 64 chunks at column 0; epilogues with one 64-column accumulator block: lngelu, decode, bias_lrelu / bias_act,
   res_id) row-paired taps: the packed weights order each dx column as pairs of vertically
   adjacent taps [dy = 1 | 0], [3 | 2], ... (+ the last tap alone when R is odd), per rep [tap_hi rows | tap_lo rows]; the kernel
   issues one MMA of twice the width per pair and folds the second column block back one row in the epilogue;
   bit 7 (res_id) the residual input is multiplied by the per-sample channel scales of SE layer (flags >> 8) & 1 (the SE
   layer was folded into its consumers, see sf_plan_define_stage_fold);
   bit 12 (propose) the output is the GRU-ODE derivative u (s~ - s) instead of the blend; with activation code 2 in bits 1-3 the
   proposal is ReLU(conv + bias) (the plain SpatialGRUODECell / SpatialGRUCell, temporal_ode_bayes.py:14-61, 165-208);
   bit 11 (res_id) the activation is applied after the residual add: out = act(conv + bias + residual) (ResNet BasicBlock);
   bit 13 (res_id, C = 64, bf16) the ConvNeXt block's pointwise pair in ONE launch (convolutions.py:334-344): the stage's conv is
   pwconv1 (one 1x1 chunk, n = 256), the epilogue applies GELU, keeps the 256-channel result in tensor memory as the A operand of
   pwconv2 (its [64][256] weights -- layer scale folded -- appended to w_packed as four K-chunks of [64 n][64 k] rows) and adds the
   residual: out = pwconv2(GELU(pwconv1(x) + b1)) + b2 + residual; vec = [b1 (256), b2 (64)];
   bit 10 (lngelu, C = 64) a 1x1 convolution + LayerNorm + GELU fused behind the stage (convolutions.py:356-361 in ONE launch):
   its [C][64] weights (hi rows, then lo rows in BF16X3) are the LAST rows of w_packed and vec = [LN1 w, LN1 b, LN2 w, LN2 b];
   the epilogue keeps the first LN+GELU result in tensor memory as the A operand of a back-to-back GEMM.                */
int sf_plan_define_stage(sf_plan* p, int stage, int epilogue, int n_chunks, const sf_chunk* chunks,
                         const void* w_packed, int w_rows, const float* vec, int n_vec,
                         const int32_t* io_bufs, const int32_t* io_choff, int n_io, int flags);
/* SE layer weights (fp32 device): fc1 [2C/8][2C], fc2 [2C][2C/8] for the two SE layers */
int sf_plan_define_se(sf_plan* p, int which, const float* fc1, const float* fc2, int in_buf, int out_buf);
/* Folds SE layer `which_se` (res_models.py:150-165, y = z * scale[sample][channel]) into a stage that convolves y: once per
   event the K (input-channel) columns of the stage's weights are multiplied by the sample's scales, and the stage then reads z
   itself with its sample's weights -- the [image][H][W][2C] tensor is not re-read and re-written just to be scaled.
   w32: fp32 master of the packed matrix, [w_rows][64] in the packed row order (every row carries the unsplit fp32 weights);
   row_meta[w_rows]: bits 0-15 first channel of the row's chunk in the SE layer's channel space, bit 16 set for the residual
   ("lo") rows of the split mode; w_scaled: bf16 scratch [max_images][w_rows][64] the stage streams instead of w_packed.
   Event-graph item 2000 + which = SE reduce + scales + weight fold (no activation pass); 1000 + which = reduce + apply.  */
int sf_plan_define_stage_fold(sf_plan* p, int stage, int which_se, const float* w32, const int32_t* row_meta, void* w_scaled);
/* stage slots used by an event: cell stages for kind 0 / kind 1, then the prior-network stages */
int sf_plan_define_event_graph(sf_plan* p, const int32_t* cell0, const int32_t* cell1, int n_cell,
                               const int32_t* prior, int n_prior);
int sf_plan_finalize(sf_plan* p);
int sf_plan_smem_bytes(sf_plan* p, int stage);

/* runs ONE stage (unit tests / profiling); table points at the event's rows in device memory */
int sf_plan_run_stage(sf_plan* p, int stage, const sf_event* ev, const int32_t* table, void* stream);
int sf_plan_run_events(sf_plan* p, const sf_event* evs, int n_events, const int32_t* table, void* stream);
/* number of kernels the last sf_plan_run_events call launched */
int sf_plan_last_launches(sf_plan* p);

/* the two halves of an SE layer, for callers that reduce the channel sums across GPUs in between (row sharding):
   reduce over the pixel window [px0, px1) of each active sample -> returns the number of per-block partial sums written
   to SF_F32_SE_SUMS[which][sample][partial][2C]; apply with mean = (sum of the first n_partials partials) * inv_n
   (n_partials = 0: the scales were already produced by a whole-image sf_plan_run_* reduce)                          */
int sf_plan_se_reduce(sf_plan* p, int which, const sf_event* ev, const int32_t* table, int px0, int px1, void* stream);
int sf_plan_se_apply(sf_plan* p, int which, const sf_event* ev, const int32_t* table, int n_partials, float inv_n, void* stream);
/* Row sharding, one squeeze-excite layer in three steps with ONE collective in between: (1) this rank's band totals per
 * (active sample, channel) over the pixel window [px0, px1) land in a [n_active][2C] fp32 array inside the SE scratch
 * (sf_plan_se_totals_ptr returns its address); (2) the caller all-reduces that array across the ranks (ncclAllReduce, sum);
 * (3) sf_plan_se_finish turns the totals into scales in place (mean = total * inv_n, two FCs, sigmoid) and folds them into the
 * consuming stages' weights (sf_plan_define_stage_fold) -- or, for a plan without folded stages, applies them (y = z * scale). */
int sf_plan_se_reduce_totals(sf_plan* p, int which, const sf_event* ev, const int32_t* table, int px0, int px1, void* stream);
int sf_plan_se_totals_ptr(sf_plan* p, int which, float** out);
int sf_plan_se_finish(sf_plan* p, int which, const sf_event* ev, const int32_t* table, float inv_n, void* stream);
/* Row sharding, halo exchange staging: rows [row0, row0 + nrows) of up to 6 NHWC tensors [B][rows][row_bytes] <-> one flat byte
 * buffer per direction (tensor-major, then batch), both directions (a = towards the upper neighbour, b = lower; NULL skips one)
 * in ONE launch; to_flat = 1 packs, 0 unpacks.  The flat buffers are what ncclSend / ncclRecv move. */
int sf_halo_copy(void* const* tensors, const long long* batch_stride_bytes, const long long* row_bytes, int n_tensors, int B, int nrows,
                 void* flat_a, int row0_a, void* flat_b, int row0_b, int to_flat, void* stream);

/* Row sharding over NVLink peer memory -- the per-event exchanges without an NCCL call (streamingflow_b200/csrc/sf_peer.cuh).
 * The reference has no multi-GPU form of this path (the split is this library's own); these entry points are what a
 * host binds next to ncclSend / ncclRecv / ncclAllReduce.
 * sf_peer_alloc: a zeroed device buffer plus its 64-byte inter-process handle (cudaIpcGetMemHandle); the other ranks of the node
 * map it with sf_peer_open (cudaIpcOpenMemHandle, peer access enabled lazily) and unmap it with sf_peer_close before the owner
 * calls sf_peer_free.  Handles travel between the processes by any host channel (torch.distributed.all_gather_object here). */
int sf_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int sf_peer_open(const unsigned char* handle64, void** ptr);
int sf_peer_close(void* ptr);
int sf_peer_free(void* ptr);
/* sf_halo_push: sf_halo_copy(to_flat = 1) whose flat buffers live in the NEIGHBOURS' memory (a = upper, b = lower neighbour; NULL
 * skips one): the rows are stored straight into copy (sequence & 1) of the neighbour's receive buffer (copies parity_stride bytes
 * apart), then the neighbour's arrival counter is released (system scope) with this launch's sequence number.  seq: two unsigned
 * words of LOCAL device memory {launches completed, block ticket}, advanced by the kernel -- no per-call argument changes, so the
 * launch replays inside a CUDA graph.
 * sf_halo_pull: waits until the local counters flag_a / flag_b have reached its own sequence number, then unpacks the received
 * rows into the halos.  A wait gives up after 10 s and latches a non-zero code in *err (device memory, zero-initialised); later
 * waits return at once, the host checks *err after the rollout.
 * trace (all three exchange entry points; NULL = off): a [64][4] ring of %globaltimer stamps in local device memory, entry
 * (sequence & 63) = {launch start, wait done, launch end, -}: where an event's exchange time goes, without a profiler. */
int sf_halo_push(void* const* tensors, const long long* batch_stride_bytes, const long long* row_bytes, int n_tensors, int B, int nrows,
                 void* peer_flat_a, int row0_a, void* peer_flat_b, int row0_b, long long parity_stride, unsigned* peer_flag_a,
                 unsigned* peer_flag_b, unsigned* seq, unsigned long long* trace, void* stream);
int sf_halo_pull(void* const* tensors, const long long* batch_stride_bytes, const long long* row_bytes, int n_tensors, int B, int nrows,
                 void* flat_a, int row0_a, void* flat_b, int row0_b, long long parity_stride, unsigned* flag_a, unsigned* flag_b,
                 unsigned* seq, int* err, unsigned long long* trace, void* stream);
/* In-place sum of data[0..n) over the ranks of the node in ONE launch of one block: the vector is stored into slot `rank` of every
 * rank's slot arena (slots[r]: [2 copies][world][n_max] floats in rank r's memory), one counter per rank is released
 * (flags[r]: [world] unsigned in rank r's memory), the block waits for the world's contributions and adds the slots in rank order
 * (bit-identical result on every rank).  Replaces ncclAllReduce for the [n_active][2C] squeeze-excite sums. */
int sf_peer_allreduce_f32(float* data, int n, int n_max, int rank, int world, void* const* slots, void* const* flags, unsigned* seq, int* err,
                          unsigned long long* trace, void* stream);

/* layout kernels (HBM-bound, 128-bit vectorised) */
int sf_pack_nchw_f32(const float* src, void* dst_hi, void* dst_lo, int n_images, int C, int H, int W, void* stream);
int sf_unpack_nhwc_f32(const float* src, float* dst, const int32_t* slots, int n_out, int C, int H, int W, void* stream);
/* Noise for the rsample calls of a whole rollout in ONE launch, bit-identical to n_slots successive
   torch.empty([1,C,H,W], device="cuda").normal_() calls (model_utils.py:107-108 -> torch.distributions.Normal.rsample ->
   at::native::normal_ -> distribution_nullary_kernel with curand_normal4 on Philox4_32_10): slot i is filled exactly as the
   i-th call would fill it, with torch's launch geometry (256-thread blocks, unroll 4, grid = min(SMs * maxThreads/256,
   ceil(numel/256))) and Philox offset offset0 + i * offset_per_slot.  sf_normal_policy returns that grid and the per-call offset
   increment for a tensor of numel elements on `device`; the caller (which owns the torch generator) passes seed / offset0 and
   advances the generator by n_slots * offset_per_slot afterwards.                                                          */
int sf_normal_policy(long long numel, int device, int* grid, int* offset_per_slot);
int sf_normal_fill_slots(float* out, int n_slots, long long numel, unsigned long long seed, unsigned long long offset0, int grid,
                         int offset_per_slot, void* stream);
/* The same for a subset: only the slots listed in slot_list (device int32 [n_list], indices into the same slot numbering) are filled,
   each exactly as sf_normal_fill_slots would fill it.  A rollout that skips the prior-net evaluations nothing reads (the reference
   draws for them too, temporal_ode_bayes.py:463-477 after every op) leaves those slots untouched; the caller still advances the
   generator by the full slot count, so the live draws are the reference's.                                                        */
int sf_normal_fill_slot_list(float* out, const int32_t* slot_list, int n_list, long long numel, unsigned long long seed,
                             unsigned long long offset0, int grid, int offset_per_slot, void* stream);

/* SmallEncoder / SmallDecoder glue on NHWC bf16 planes: 2x2 max-pool (res_models.py:96-104), nearest x2 up-sampling
   (:134-147; call once per plane), fp32 NHWC (gathered by slot) -> bf16 hi [+ lo] */
int sf_maxpool2(const void* src_hi, const void* src_lo, void* dst_hi, void* dst_lo, int n_images, int H, int W, int C, void* stream);
int sf_upsample2(const void* src, void* dst, int n_images, int H, int W, int C, void* stream);
int sf_cast_nhwc_f32(const float* src, const int32_t* slots, void* dst_hi, void* dst_lo, int n_out, int C, int H, int W, void* stream);
/* post-ODE refinement glue on C-channel NHWC bf16 planes, C = 64 or 128 (FuturePredictionODE's in_channels):
   ConvNeXt depthwise 7x7 conv + bias + channels-last LayerNorm (convolutions.py:327-333); ASPP image-pooling branch folded into a
   per-image bias of the projection conv (convolutions.py:198-240):
   out[img][128] = proj_w[128][128] . relu(pool_w[128][C] . mean(img) + pool_b) + proj_b;  scratch: [n_images][32][C] floats */
int sf_dwconv7_ln(const void* src_hi, const void* src_lo, void* dst_hi, void* dst_lo, const float* dw_w, const float* dw_b,
                  const float* ln_w, const float* ln_b, int n_images, int C, int H, int W, void* stream);
int sf_aspp_pool_bias(const void* src_hi, const void* src_lo, const float* pool_w, const float* pool_b, const float* proj_w,
                      const float* proj_b, float* scratch, float* out, int n_images, int C, int H, int W, void* stream);

/* self-test kernels for bring-up: TMA tile dump and a single UMMA tile product */
int sf_diag_tma_dump(const void* act_bf16, int n_images, int H, int W, int C, int img, int y0, int x0, int c0,
                     int rows, void* out_smem_copy, void* stream);
int sf_diag_umma(const void* a_bf16, const void* b_bf16, float* d, int n, int k_chunks, void* stream);
/* experiment: 128 operand rows starting `shift` rows into a swizzled box, 8-row groups `sbo_rows` rows apart */
int sf_diag_umma_shift(const void* a_bf16, const void* b_bf16, float* d, int n, int rows_a, int shift, int sbo_rows, int base_mode,
                       void* stream);

/* BEV Decoder head (streamingflow/models/decoder.py:91-140), the kernels next to the conv stages:
 * space to depth for the stride-2 convolutions: dst[img][i][j][(2 py + px) C + c] = src[img][2i + py][2j + px][c] (one plane);
 * UpsamplingAdd with its 1x1 conv + BN hoisted to the low resolution: dst = bilinear_x2(src) + skip (planes hi [+ lo]);
 * a head's output 1x1 conv (64 -> K <= 4, + bias, optional sigmoid) -> fp32 NCHW [img][K][H][W] and, with mask != NULL,
 * the per-pixel arg-max over the K channels as uint8 [img][H][W] (trainer.py:230-231).                                  */
int sf_space_to_depth2(const void* src, void* dst, int n_images, int H, int W, int C, void* stream);
int sf_bilinear_up2_add(const void* src_hi, const void* src_lo, const void* skip_hi, const void* skip_lo, void* dst_hi, void* dst_lo,
                        int n_images, int H, int W, int C, void* stream);
int sf_head_1x1(const void* src_hi, const void* src_lo, const float* w, const float* b, int K, int sigmoid_out, float* out,
                unsigned char* mask, int n_images, int H, int W, void* stream);

/* bring-up of the back-to-back GEMM in the fused trunk epilogue: A [128 x 64] bf16 written to tensor memory by the threads
 * (tcgen05.st, two elements per column, at column a_col >= n), B [n x 64] from shared memory; d [128 x n] fp32 */
int sf_diag_umma_ts(const void* a_bf16, const void* b_bf16, float* d, int n, int a_col, void* stream);

/* =================================================================================================================
 * The ODE head driven from this header alone (SURVEY 8b's proposed exports: sf_query_workspace, sf_pack_cell_weights,
 * sf_pack_pmodel_weights, sf_event, sf_rollout -- the last two are sf_ode_event / sf_ode_rollout here because `sf_event` is the
 * event struct).  streamingflow_b200/csrc/sf_ode.cu; pure host code on top of the sf_plan_* entry points above.
 *
 *   weights (host fp32, reference state_dict names)  --sf_pack_*-->  stage descriptions + packed bf16 matrices (host)
 *   sf_ode_create: carves ONE caller-owned device allocation into the activation / fp32 / weight buffers, uploads the packed
 *       weights, defines every stage, the two SE layers and the event graph on an sf_plan and finalises it
 *   sf_rollout_plan_*: per-sample timestamps -> the batched event list + int32 event table (the host schedule of
 *       temporal_ode_bayes.py:508,539-622, incl. the reference's float32 / float64 time arithmetic and its 1-ulp micro-steps)
 *   sf_ode_rollout: enqueues every stage launch of the event list on `stream` (no sync, no allocation: graph-capturable)
 * ================================================================================================================= */

/* activation buffer ids of the ODE head's plan (the stage descriptions produced by sf_pack_* refer to them) */
enum {
  SF_BUF_S0 = 0, SF_BUF_S1, SF_BUF_X, SF_BUF_OBS, SF_BUF_ZERO, SF_BUF_U1, SF_BUF_U2, SF_BUF_G1, SF_BUF_G2, SF_BUF_A, SF_BUF_B,
  SF_BUF_HH, SF_BUF_T1, SF_BUF_T2, SF_BUF_Q1, SF_BUF_Z1, SF_BUF_Y1, SF_BUF_Q3, SF_BUF_Z2, SF_BUF_Y2, SF_BUF_COUNT
};
/* chunk sources resolved per event: the event's x buffer, the bf16 mirror of the state buffer the cell reads / writes */
#define SF_SRC_X (-1)
#define SF_SRC_STATE_IN (-2)
#define SF_SRC_STATE_OUT (-3)

/* one named fp32 tensor in HOST memory, contiguous, in torch's own layout (conv weights [out][in][ky][kx]); `name` is the
   reference's state_dict key below the ODE module, e.g. "gru_c.conv_update_1.weight" (temporal_ode_bayes.py:357-393) */
typedef struct {
  const char* name;
  const float* data;
  int64_t numel;
} sf_tensor;

/* read-only view of one item of a packed stage list.  se_layer >= 0: the item is not a conv stage but marks where squeeze-excite
   layer se_layer (res_models.py:150-165) sits in the launch order; then only name / se_layer / fc1 / fc2 are meaningful. */
typedef struct {
  const char* name;
  int32_t se_layer;
  int32_t epilogue, flags;
  int32_t n_chunks;
  const sf_chunk* chunks;
  const void* w;            /* bf16 [w_rows][64], the matrix sf_plan_define_stage streams                              */
  int32_t w_rows;
  const float* vec;         /* per-stage constants (biases / LayerNorm / gate weights)                                 */
  int32_t n_vec;
  int32_t n_io;
  const int32_t* io;
  const int32_t* io_off;
  int32_t fold_se;          /* >= 0: this stage consumes SE layer fold_se folded into its weights: w32 / row_meta are the
                               arguments of sf_plan_define_stage_fold                                                  */
  const float* w32;
  const int32_t* row_meta;
  const float* fc1;         /* SE marker: fc.0.weight [2C/8][2C], fc.2.weight [2C][2C/8], n_fc elements each           */
  const float* fc2;
  int32_t n_fc;
} sf_stage_desc;

typedef struct sf_packed sf_packed;
#define SF_PACK_PAIR_ROWS 1   /* C = 64: row-paired taps in the 7x7 trunk (stage flag bit 9)                           */
#define SF_PACK_B2B 2         /* C = 64: 7x7 + LN + GELU + 1x1 + LN + GELU as ONE stage (stage flag bit 10)            */
#define SF_PACK_FOLD_SE 4     /* prior net: SE layers folded into their consumers' weights                             */
#define SF_PACK_PAIR_3X3 8    /* C = 64: row-paired taps also in the 64 -> 64 3x3 stages (decode, q1); measured slower: not a default */
#define SF_PACK_DEFAULT (SF_PACK_PAIR_ROWS | SF_PACK_B2B | SF_PACK_FOLD_SE)
/* One dual-GRU cell (prefix "gru_c." = derivative, temporal_ode_bayes.py:92-161; "gru_obs.gru_d." = observation jump, :239-305):
   gates / propose / decode / trunk / mix stages with cat[state, state] of GRU-2 folded; channel width from the weights. */
int sf_pack_cell_weights(const sf_tensor* tensors, int n_tensors, const char* prefix, int precision, int options, sf_packed** out);
/* p_model = ConvNet(C, 2C) with eval-mode BatchNorm folded (prefix "p_model.", res_models.py:168-180): q1..q5 + the two SE markers */
int sf_pack_pmodel_weights(const sf_tensor* tensors, int n_tensors, const char* prefix, int precision, int options, sf_packed** out);
int sf_packed_count(const sf_packed* p);
int sf_packed_get(const sf_packed* p, int i, sf_stage_desc* out);
int sf_packed_free(sf_packed* p);

typedef struct {
  int32_t path_slots;     /* recorded states the PATH tensor holds                                                    */
  int32_t obs_images;     /* encoded observation frames the OBS buffer holds                                          */
  int32_t eps_slots;      /* noise tensors the EPS tensor holds                                                       */
  int32_t pack_options;   /* SF_PACK_* (NULL options: SF_PACK_DEFAULT)                                                 */
} sf_ode_options;

typedef enum {            /* device tensors inside the workspace (sf_ode_tensor)                                      */
  SF_ODE_OBS_HI = 0,      /* bf16 NHWC [obs_images][H][W][C]; fill with sf_pack_nchw_f32 or sf_ode_set_observations    */
  SF_ODE_OBS_LO = 1,      /* its residual plane (BF16X3), else NULL                                                   */
  SF_ODE_EPS = 2,         /* fp32 NCHW [eps_slots][C][H][W]; fill with sf_normal_fill_slots                           */
  SF_ODE_PATH = 3,        /* fp32 NHWC [path_slots][H][W][C]; read with sf_unpack_nhwc_f32 or sf_ode_read_path        */
  SF_ODE_STATE0 = 4,      /* fp32 NHWC [max_images][H][W][C]                                                          */
  SF_ODE_STATE1 = 5,
  SF_ODE_X32 = 6,         /* fp32 copy of the sampled input (events with want_f32)                                    */
  SF_ODE_PARAMS32 = 7,    /* fp32 p_model output [max_images][H][W][2C] (events with want_f32)                        */
  SF_ODE_ERRFLAG = 8      /* int32 word: non-zero after a device-side pipeline timeout                                */
} sf_ode_tensor_id;

typedef struct sf_ode sf_ode;
/* bytes of the single device allocation sf_ode_create carves up (the library allocates no device memory itself) */
int sf_ode_query_workspace(const sf_geometry* g, const sf_ode_options* o, size_t* bytes);
/* tensors: the ODE module's parameters (gru_c.*, gru_obs.gru_d.*, p_model.*; others are ignored) with `prefix` stripped from
   the front of each name (e.g. "gru_ode." when they come from FuturePredictionODE's state_dict).  Synchronous (uploads). */
int sf_ode_create(const sf_geometry* g, const sf_ode_options* o, const sf_tensor* tensors, int n_tensors, const char* prefix,
                  void* workspace, size_t workspace_bytes, sf_ode** out);
int sf_ode_destroy(sf_ode* ode);
int sf_ode_plan(sf_ode* ode, sf_plan** plan);                /* the underlying plan (stage-level access)              */
int sf_ode_tensor(sf_ode* ode, int which, void** ptr, size_t* bytes);
/* encoded observations, fp32 NCHW on the device -> images [first_image, first_image + n_images) of the OBS buffer */
int sf_ode_set_observations(sf_ode* ode, const float* obs_nchw, int first_image, int n_images, void* stream);
/* state buffers (fp32 masters and bf16 mirrors) of the first n_images samples back to zero (temporal_ode_bayes.py:505)  */
int sf_ode_reset_state(sf_ode* ode, void* stream);
/* ONE event (cell + state update + prior net + sampling) / a whole event list; `table` is the device copy of the event table.
   SURVEY 8b's finer-grained proposals are events with flags: sf_dual_gru_cell(mode = DERIV | JUMP) = kind 0 / 1 with run_prior = 0
   (DualGRUODECell.forward / GRUObservationCell.forward; the derivative itself is the update from a zero base buffer with dt = 1:
   s_base = a zeroed state buffer, s_in = the state, dt = 1 -> s_out = f(x, s)), sf_infer_state = run_cell = 0, run_prior = 1,
   want_f32 = 1 (fills SF_ODE_X32 / SF_ODE_PARAMS32), sf_event = one full event, sf_rollout = the list. */
int sf_ode_event(sf_ode* ode, const sf_event* ev, const int32_t* table, void* stream);
int sf_ode_rollout(sf_ode* ode, const sf_event* evs, int n_events, const int32_t* table, void* stream);
/* recorded states `slots` (device int32 [n]) -> fp32 NCHW [n][C][H][W] on the device */
int sf_ode_read_path(sf_ode* ode, const int32_t* slots, int n, float* out_nchw, void* stream);

/* Host schedule: B samples, each with n_obs observation times (already in processing order, future_prediction_ode.py:37-45) and
   n_targets target times.  obs_f32 / target_f32: the caller's timestamp tensors were float32 (the reference then compares and
   steps in float32, see schedule.py).  solver: 0 euler, 1 midpoint.  flags: bit 0 keep the dead prior-net evaluations (the
   reference's literal schedule), bit 1 keep the input sampled after the last op alive (streaming).
   obs image index of sample b's k-th observation = b * n_obs + k. */
/* Processing order of ONE sample's observations (future_prediction_ode.py:37-45: the camera frames go into a dict, then the LiDAR
   frames, then a STABLE sort by time -- camera wins ties, duplicates are all kept).  Writes n_cam + n_lidar entries: times[i] and
   source[i] = sensor * 65536 + index (sensor 0 = camera, 1 = LiDAR); returns the count.  lidar_t may be NULL (n_lidar = 0). */
int sf_merge_observations(const double* camera_t, int n_cam, const double* lidar_t, int n_lidar, double* times, int32_t* source);
typedef struct sf_rollout_plan sf_rollout_plan;
typedef struct {
  int32_t n_events, n_table, n_eps, n_path, n_state_steps, n_jumps, n_cell_evals, n_prior_evals;
} sf_rollout_info;
int sf_rollout_plan_create(const double* obs_times, int n_obs, const double* targets, int n_targets, int B, double delta_t,
                           int variable_step, int solver, int impute, int obs_f32, int target_f32, int flags, sf_rollout_plan** out);
int sf_rollout_plan_info(const sf_rollout_plan* r, sf_rollout_info* info);
const sf_event* sf_rollout_plan_events(const sf_rollout_plan* r);
const int32_t* sf_rollout_plan_table(const sf_rollout_plan* r);          /* host int32 [n_table]: upload before sf_ode_rollout */
const int32_t* sf_rollout_plan_out_slots(const sf_rollout_plan* r);      /* [B][n_targets] path slot of every output frame     */
int sf_rollout_plan_free(sf_rollout_plan* r);

#ifdef __cplusplus
}
#endif
#endif /* SF_B200_H_ */
