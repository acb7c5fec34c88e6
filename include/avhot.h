/* avhot.h - C ABI of libavhot.so: the B200 (sm_100a) audio-visual inference hot path.
 *
 * The reference (parshva-mehta/multimodal-emotion-detection) is pure Python and has
 * no FFI of its own; its "operator interface" for this path is the set of Python
 * callables cited beside each entry point below (paths relative to the reference
 * root).  The Python drop-in modules in multimodal-emotion-detection_b200/ keep
 * those signatures and bind these symbols with ctypes (see INTEGRATION.md for the
 * stub a reference maintainer would add).
 *
 * Conventions
 *  - every pointer argument is a DEVICE pointer unless its name ends in _host;
 *    fp32, row-major, contiguous, in the reference's own tensor layouts;
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and
 *    never synchronise the host.  Only *_create allocates (device memory owned by
 *    the handle, released by *_destroy); the forward calls use caller workspaces;
 *  - return value: AVHOT_OK or a negative avhot_status; avhot_last_error() gives the
 *    message for the calling thread.  The Python layer raises ValueError for
 *    AVHOT_ERR_INVALID / AVHOT_ERR_UNSUPPORTED (the reference raises ValueError /
 *    assert for bad shapes: src/encoders.py:115,242,636; src/attention.py:34,74-76;
 *    src/fusion.py:696) and RuntimeError for AVHOT_ERR_CUDA;
 *  - one host thread per device, like the reference's callers (src/train.py:133-136).
 */
#ifndef AVHOT_H_
#define AVHOT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVHOT_VERSION 100 /* 0.1.0 */

typedef void* avhot_stream_t; /* cudaStream_t */

typedef enum {
  AVHOT_OK = 0,
  AVHOT_ERR_INVALID = -1,     /* bad argument / shape                      */
  AVHOT_ERR_UNSUPPORTED = -2, /* valid in the reference, not built here    */
  AVHOT_ERR_CUDA = -3,        /* CUDA runtime error (see avhot_last_error) */
  AVHOT_ERR_WORKSPACE = -4    /* caller workspace too small                */
} avhot_status;

/* arithmetic of the dense contractions (GEMM operands); accumulation is fp32 and everything
 * outside the contractions is fp32 in every mode. */
typedef enum {
  AVHOT_PREC_FP32 = 0,   /* fp32 FMA on CUDA cores: the parity path (1e-4 relative)            */
  AVHOT_PREC_TC = 2      /* tcgen05 tensor cores, operands rounded once: TF32 for the
                            time-batched GEMMs, fp16 for the recurrent GEMM; stated tolerance
                            in DESIGN.md                                                       */
} avhot_precision;

int avhot_version(void);
const char* avhot_last_error(void);
/* cudaGetDeviceProperties of the current device: fails unless it is sm_100. */
int avhot_check_device(void);

/* ---- measurement hooks (bench.py): kernels launched so far by this library in this process, and
 * optional CUDA-event timing of each named kernel on its launching stream. ---------------------- */
long long avhot_launch_count(void);
int avhot_prof_enable(int on);
int avhot_prof_reset(void);
/* total device time (ms) and number of timed launches of kernel `name` since the last reset;
 * synchronises on the recorded events. */
int avhot_prof_read(const char* name, double* total_ms, int* launches);
/* newline-separated names seen since the last reset */
int avhot_prof_names(char* buf, int buf_bytes);

/* C[M,N] = act(A[M,K] W[N,K]^T + bias[N]) with the library's GEMM in `precision` (test entry for the
 * contraction every nn.Linear on the path lowers to); A, W, C dense row-major, K % 4 == 0. */
int avhot_gemm(const float* a, const float* w, const float* bias, float* c, int M, int N, int K,
               int relu, int precision, avhot_stream_t stream);

/* ---- Philox-4x32-10 spec (oracle/philox.py) ------------------------------------ */
/* keep[b*n_elems + e] = 1 iff words(e>>2, clip0+b, t, site)[e&3] >= floor(p*2^32).
 * Test/diagnostic entry: the masks the fused kernels apply at the nn.Dropout sites
 * of src/encoders.py:228,230; src/attention.py:51; src/fusion.py:448. */
int avhot_dropout_keep(uint8_t* keep, int B, int n_elems, uint64_t clip0, uint32_t t,
                       uint32_t site, float p, uint64_t seed, avhot_stream_t stream);
/* src/data.py:108-118  modality dropout: mask (B, M) fp32, 1 = available. */
int avhot_modality_mask(float* mask, int B, int M, uint64_t clip0, float p, uint64_t seed,
                        avhot_stream_t stream);

/* ---- synthetic RAVDESS-shaped inputs (oracle/synth.py), bit-exact with the oracle -- */
int avhot_synth_audio(float* wave, int B, int n_samples, uint64_t clip0, uint64_t seed,
                      avhot_stream_t stream);
/* exactly one of frames_f32 / frames_u8 may be NULL */
int avhot_synth_video(float* frames_f32, uint8_t* frames_u8, int B, int n_elems_per_clip,
                      uint64_t clip0, uint64_t seed, avhot_stream_t stream);

/* ---- a1: src/dataprocessing.py:198-211  load_raw_audio post-conditions ------------ */
/* out[b, :] = first min(len_b, n_out) samples of in[b, :], right zero padded to n_out,
 * divided by max|.| when that is > 0.  lengths may be NULL (= in_stride). */
int avhot_condition_audio(const float* in, const int32_t* lengths, float* out, int B,
                          int in_stride, int n_out, avhot_stream_t stream);
/* The same for 16-bit PCM as a .wav file holds it (sample k = k / 32768, the conversion librosa.load applies at
 * src/dataprocessing.py:196): bit-identical to avhot_condition_audio on the converted samples, half the bytes in. */
int avhot_condition_audio_i16(const int16_t* in, const int32_t* lengths, float* out, int B,
                              int in_stride, int n_out, avhot_stream_t stream);

/* ---- a2: log-mel + MFCC front-end (no reference code; spec in oracle/frontend.py) --- */
/* wave (B, n_samples) -> feats (B, 1 + n_samples/160, 104) = [64 log-mel | 40 MFCC]. */
int avhot_frontend(const float* wave, float* feats, int B, int n_samples,
                   avhot_stream_t stream);
/* Fused variant for the tensor-core pipeline: besides (or instead of: feats may be NULL) the fp32 features it
 * writes the LSTM's first GEMM operand directly - fp16, time-major rows (t, clip) with the batch padded to
 * Bp = B rounded up to 128, pitch = avhot_lstm_split_pitch() = 320 halves per row, hi = fp16(x), lo = fp16(x - hi),
 * features in blocks of 64 with [hi | lo | hi] per block:
 * [lm_hi(64) | lm_lo(64) | lm_hi(64) | mf_hi(40) | mf_lo(40) | mf_hi(40) | 0(8)] - see avhot_lstm_forward_split. */
int avhot_frontend_ex(const float* wave, float* feats, void* seq_split_f16, int B, int Bp, int pitch,
                      int n_samples, avhot_stream_t stream);
/* the fp32 tables the kernel uses, copied to HOST buffers (any may be NULL):
 * mel (64, 257), dct (40, 64), window (512). */
int avhot_frontend_tables_host(float* mel_host, float* dct_host, float* window_host);

/* ---- a3: src/encoders.py:15-155  SequenceEncoder(encoder_type='lstm') -------------- */
typedef struct avhot_lstm* avhot_lstm_t;
typedef struct {
  int input_dim, hidden_dim, num_layers, output_dim; /* ctor args, src/encoders.py:23-31 */
  const float* weight_ih[4];                         /* rnn.weight_ih_l{k} (4H, in_k)    */
  const float* weight_hh[4];                         /* rnn.weight_hh_l{k} (4H, H)       */
  const float* bias_ih[4];                           /* rnn.bias_ih_l{k}   (4H)          */
  const float* bias_hh[4];                           /* rnn.bias_hh_l{k}   (4H)          */
  const float* proj_weight;                          /* projection.weight (out, H)       */
  const float* proj_bias;                            /* projection.bias   (out)          */
} avhot_lstm_desc;
int avhot_lstm_create(const avhot_lstm_desc* desc, avhot_lstm_t* out, avhot_stream_t stream);
void avhot_lstm_destroy(avhot_lstm_t h);
size_t avhot_lstm_workspace_bytes(avhot_lstm_t h, int B, int T, int precision);
/* forward(sequence (B,T,input_dim), lengths) -> (B, output_dim).  src/encoders.py:137-155.
 * lengths: (B) int32 valid steps per clip (the packed-sequence branch, :139-147: h_n is the
 * state after step lengths[b]), or NULL (every clip runs T steps, :148-154). */
int avhot_lstm_forward(avhot_lstm_t h, const float* seq, const int32_t* lengths, float* out,
                       int B, int T, int precision, void* workspace, size_t workspace_bytes,
                       avhot_stream_t stream);

/* Tensor-core path with the input already in the split layout avhot_frontend_ex writes ((T, Bp, pitch) fp16):
 * same result as avhot_lstm_forward(..., AVHOT_PREC_TC, ...) on the fp32 features, minus one pass over them. */
int avhot_lstm_split_pitch(avhot_lstm_t h);
int avhot_lstm_forward_split(avhot_lstm_t h, const void* seq_split_f16, const int32_t* lengths, float* out,
                             int B, int T, void* workspace, size_t workspace_bytes, avhot_stream_t stream);

/* The same for an operand written by avhot_frontend_ex (input_dim 104 = 64 log-mel + 40 MFCC): its MFCC columns are the
 * front-end's DCT matrix times its log-mel columns, so layer 0 multiplies the log-mel k-blocks alone by
 * W_ih[:, :64] + W_ih[:, 64:] D (folded once per handle, fp32-class): K = 192 instead of 320.  Same result as
 * avhot_lstm_forward_split within fp32 rounding of the MFCC; AVHOT_LSTM_DCT_FOLD=0 routes it to that call. */
int avhot_lstm_forward_frontend(avhot_lstm_t h, const void* seq_split_f16, const int32_t* lengths, float* out,
                                int B, int T, void* workspace, size_t workspace_bytes, avhot_stream_t stream);

/* ---- src/encoders.py:65-77, :137-155  SequenceEncoder(encoder_type='gru') ---------- */
/* Same descriptor as the LSTM with nn.GRU's tensors: weight_ih_l{k} (3H, in_k), weight_hh_l{k} (3H, H), bias_ih /
 * bias_hh (3H), gate rows r, z, n.  fp32 path; lengths as in avhot_lstm_forward. */
typedef struct avhot_gru* avhot_gru_t;
int avhot_gru_create(const avhot_lstm_desc* desc, avhot_gru_t* out, avhot_stream_t stream);
void avhot_gru_destroy(avhot_gru_t h);
size_t avhot_gru_workspace_bytes(avhot_gru_t h, int B, int T);
int avhot_gru_forward(avhot_gru_t h, const float* seq, const int32_t* lengths, float* out, int B, int T,
                      void* workspace, size_t workspace_bytes, avhot_stream_t stream);

/* src/encoders.py:79-94, :157-168  SequenceEncoder(encoder_type='cnn'): Conv1d(k 5) - BatchNorm1d (eval) - ReLU - Dropout -
 * Conv1d(k 3) - BatchNorm1d - ReLU - mean over T - Dropout - Linear.  fp32 path.  z (B, passes, out): pass t of clip b
 * draws the masks of (seed, t0 + t, clip0 + b) at sites AVHOT_SITE_CNN0 ((H, T) activation) and AVHOT_SITE_CNN1
 * (pooled vector); passes = 1 with p = 0 is eval mode.  `lengths` is ignored by the reference for this type. */
typedef struct avhot_cnn* avhot_cnn_t;
typedef struct {
  int input_dim, hidden_dim, output_dim;
  const float *conv1_weight, *conv1_bias;                       /* conv1.weight (H, in, 5), conv1.bias (H) */
  const float *bn1_weight, *bn1_bias, *bn1_mean, *bn1_var;      /* bn1.weight / bias / running_mean / running_var */
  const float *conv2_weight, *conv2_bias;                       /* conv2.weight (H, H, 3), conv2.bias (H) */
  const float *bn2_weight, *bn2_bias, *bn2_mean, *bn2_var;
  const float *proj_weight, *proj_bias;                         /* projection.weight (out, H), projection.bias */
  float bn_eps;                                                 /* BatchNorm1d eps (1e-5) */
} avhot_cnn_desc;
int avhot_cnn_create(const avhot_cnn_desc* desc, avhot_cnn_t* out, avhot_stream_t stream);
void avhot_cnn_destroy(avhot_cnn_t h);
size_t avhot_cnn_workspace_bytes(avhot_cnn_t h, int B, int T, int passes);
int avhot_cnn_forward(avhot_cnn_t h, const float* seq, float* z, int B, int T, int passes, float p, uint64_t seed,
                      uint64_t clip0, uint32_t t0, void* workspace, size_t workspace_bytes, avhot_stream_t stream);

/* ---- a4: src/encoders.py:196-327  FrameEncoder(temporal_pooling='attention') -------- */
typedef struct avhot_frame* avhot_frame_t;
typedef struct {
  int frame_dim, hidden_dim, output_dim; /* ctor args, src/encoders.py:203-210 */
  const float* mlp_weight;               /* frame_mlp.0.weight (H, frame_dim)  */
  const float* mlp_bias;                 /* frame_mlp.0.bias   (H)             */
  const float* attn_weight;              /* attention.weight   (1, H); both NULL for      */
  const float* attn_bias;                /* attention.bias     (1)     average/max pooling */
  const float* ln_weight;                /* projection.0.weight (H)            */
  const float* ln_bias;                  /* projection.0.bias   (H)            */
  const float* proj_weight;              /* projection.1.weight (out, H)       */
  const float* proj_bias;                /* projection.1.bias   (out)          */
} avhot_frame_desc;
int avhot_frame_create(const avhot_frame_desc* desc, avhot_frame_t* out, avhot_stream_t stream);
void avhot_frame_destroy(avhot_frame_t h);
size_t avhot_frame_workspace_bytes(avhot_frame_t h, int B, int F, int precision);
/* frame_mlp[0:2] = Linear + ReLU (src/encoders.py:225-227,272): frames (B*F, frame_dim)
 * fp32, or u8 pixels k meaning k/255 (src/dataprocessing.py:261-265) when frames_u8 is
 * given instead -> hid (B*F, H).  Pass-invariant under MC-dropout. */
int avhot_frame_hidden(avhot_frame_t h, const float* frames_f32, const uint8_t* frames_u8,
                       float* hid, int B, int F, int precision, void* workspace,
                       size_t workspace_bytes, avhot_stream_t stream);
/* the rest of forward (src/encoders.py:228,274-297,320-326): dropout -> attention_pool ->
 * dropout -> LayerNorm -> Linear, for T passes: z (B, T, out).  T = 1 with p = 0 is eval
 * mode.  Pass t of clip b uses Philox masks (seed, t0 + t, clip0 + b).  frame_valid: (B, F)
 * bytes, 1 = valid frame (forward's `mask`, src/encoders.py:321-324), or NULL (the hot path,
 * src/train.py:136). */
int avhot_frame_tail(avhot_frame_t h, const float* hid, const uint8_t* frame_valid, float* z,
                     int B, int F, int T, float p, uint64_t seed, uint64_t clip0, uint32_t t0,
                     avhot_stream_t stream);
/* forward with temporal_pooling = 'average' / 'max' (src/encoders.py:276-293; `mask` = frame_valid: masked mean with
 * the count clamped to >= 1, masked max with invalid frames at finfo.min), or 'attention' (= avhot_frame_tail). */
enum { AVHOT_POOL_ATTENTION = 0, AVHOT_POOL_AVERAGE = 1, AVHOT_POOL_MAX = 2 };
int avhot_frame_tail_pool(avhot_frame_t h, const float* hid, const uint8_t* frame_valid, float* z,
                          int B, int F, int T, float p, uint64_t seed, uint64_t clip0, uint32_t t0,
                          int pooling, avhot_stream_t stream);
/* FrameEncoder.attention_pool (src/encoders.py:300-327) on its own: frames (B, F, hidden) already encoded ->
 * pooled (B, hidden) = softmax_f(attention(frames)) weighted sum; frame_valid as above or NULL. */
int avhot_frame_attention_pool(avhot_frame_t h, const float* frames, const uint8_t* frame_valid,
                               float* pooled, int B, int F, avhot_stream_t stream);
/* avhot_frame_tail with the global index of every clip given explicitly - clip_ids (B) uint32, device - instead of
 * clip0 + b: rows of scattered clips draw the masks they would draw at their place in the full batch (exact.py). */
int avhot_frame_tail_ids(avhot_frame_t h, const float* hid, const uint8_t* frame_valid, float* z,
                         int B, int F, int T, float p, uint64_t seed, const uint32_t* clip_ids, uint32_t t0,
                         avhot_stream_t stream);
/* The same up to and including projection[0] = LayerNorm (src/encoders.py:246-248): y (B, T, hidden) fp32.  The
 * tensor-core hot path feeds y to a HybridFusion handle whose proj_layers.video is the product of projection[1]
 * (src/encoders.py:249) and proj_layers.video (src/fusion.py:486) - two linear maps with nothing in between. */
int avhot_frame_tail_norm(avhot_frame_t h, const float* hid, const uint8_t* frame_valid, float* y,
                          int B, int F, int T, float p, uint64_t seed, uint64_t clip0, uint32_t t0,
                          avhot_stream_t stream);

/* ---- a5: src/attention.py:14-141  CrossModalAttention (general shapes) ------------- */
typedef struct avhot_xattn* avhot_xattn_t;
typedef struct {
  int query_dim, key_dim, hidden_dim, num_heads; /* ctor args, src/attention.py:31-32 */
  const float *q_ln_w, *q_ln_b, *k_ln_w, *k_ln_b, *v_ln_w, *v_ln_b;
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b;
} avhot_xattn_desc;
int avhot_xattn_create(const avhot_xattn_desc* desc, avhot_xattn_t* out, avhot_stream_t stream);
void avhot_xattn_destroy(avhot_xattn_t h);
/* forward(query (B,Tq,Dq), key (B,Tk,Dk), value (B,Tk,Dk), mask) -> out (B,Tq,D),
 * attn (B,heads,Tq,Tk).  invalid: (B,Tk) bytes, 1 = masked key, or NULL.  Dropout on the
 * attention weights when p > 0 (site AVHOT_SITE_ATTN + site_index). Tk <= 64. */
size_t avhot_xattn_workspace_bytes(avhot_xattn_t h, int B, int Tq, int Tk);
int avhot_xattn_forward(avhot_xattn_t h, const float* query, const float* key,
                        const float* value, const uint8_t* invalid, float* out, float* attn,
                        int B, int Tq, int Tk, float p, uint64_t seed, uint64_t clip0,
                        uint32_t t, uint32_t site_index, void* workspace,
                        size_t workspace_bytes, avhot_stream_t stream);

/* ---- a6 + a7: src/fusion.py:404-530 HybridFusion; src/train.py:78-85,153-164 head --- */
typedef struct avhot_fusion* avhot_fusion_t;
#define AVHOT_MAX_MODALITIES 4
typedef struct {
  int num_modalities, hidden_dim, num_heads, num_classes, gate_hidden;
  int modality_dim[AVHOT_MAX_MODALITIES];            /* modality_dims values, in key order */
  const float* proj_w[AVHOT_MAX_MODALITIES];         /* proj_layers.<m>.weight (D, dim_m)  */
  const float* proj_b[AVHOT_MAX_MODALITIES];
  avhot_xattn_desc attn[AVHOT_MAX_MODALITIES];       /* attn_layers.<m>.*                  */
  const float *pre_ln_w, *pre_ln_b, *post_ln_w, *post_ln_b;
  const float *gate0_w, *gate0_b, *gate3_w, *gate3_b; /* fusion_gate.0 (G,D), .3 (1,G)     */
  const float *cls_w, *cls_b;                         /* classifier (C, D)                 */
} avhot_fusion_desc;
typedef struct {
  int in_dim, hidden_dim, num_classes;               /* fusion_head: Linear-ReLU-Linear    */
  const float *w0, *b0, *w2, *b2;
} avhot_head_desc;
/* exactly one of fusion / head is non-NULL: hybrid (src/train.py:138-147 composition) or
 * the live concat baseline (src/train.py:153-164). */
int avhot_fusion_create(const avhot_fusion_desc* fusion, const avhot_head_desc* head,
                        avhot_fusion_t* out, avhot_stream_t stream);
void avhot_fusion_destroy(avhot_fusion_t h);
/* logits for R = B*T rows.  z[m] points at modality m's embeddings; row (b, t) reads
 * z[m] + (b * z_rows_per_clip[m] + (z_rows_per_clip[m] == 1 ? 0 : t)) * dim_m, so a
 * pass-invariant modality passes 1 row per clip and a stochastic one passes T.
 * mask (B, M) fp32 1 = available, or NULL.  Optional outputs (may be NULL):
 * fusion_weights (R, M), attn (M, R, heads, M), h_att (R, M, D).
 * Dropout (p > 0): attention weights (site ATTN+i) and gate hidden (site GATE). */
size_t avhot_fusion_workspace_bytes(avhot_fusion_t h, int B, int T);
int avhot_fusion_forward(avhot_fusion_t h, const float* const* z_host_array,
                         const int* z_rows_per_clip_host, const float* mask, float* logits,
                         float* fusion_weights, float* attn, float* h_att, int B, int T,
                         float p, uint64_t seed, uint64_t clip0, uint32_t t0, int precision,
                         void* workspace, size_t workspace_bytes, avhot_stream_t stream);
/* workspace (optional, avhot_fusion_workspace_bytes): lets the hybrid path compute the pass-invariant
 * half of modality 0 once per clip when z_rows_per_clip[0] == 1 (same results, fewer FLOPs). */
/* logits only, with per-row global clip indices clip_ids (B) uint32 (device) instead of clip0 + b (exact.py) */
int avhot_fusion_forward_ids(avhot_fusion_t h, const float* const* z_host_array,
                             const int* z_rows_per_clip_host, const float* mask, float* logits, int B, int T,
                             float p, uint64_t seed, const uint32_t* clip_ids, uint32_t t0, int precision,
                             void* workspace, size_t workspace_bytes, avhot_stream_t stream);

/* ---- SURVEY.md 8(f) "next": the other fusion heads behind build_fusion_model and the calibration metrics ------ */
typedef struct avhot_head* avhot_head_t;
/* src/fusion.py:19-138  EarlyFusion: mlp = Linear(sum dims -> D), LayerNorm, ReLU, Dropout, Linear(D -> D), LayerNorm,
 * ReLU, Dropout, Linear(D -> C)  (state-dict keys mlp.0, mlp.1, mlp.4, mlp.5, mlp.8) */
typedef struct {
  int num_modalities, hidden_dim, num_classes;
  int modality_dim[AVHOT_MAX_MODALITIES];
  const float *w0, *b0, *ln1_w, *ln1_b, *w4, *b4, *ln5_w, *ln5_b, *w8, *b8;
} avhot_early_desc;
/* src/fusion.py:141-248  LateFusion: classifiers.<m> = Linear(dim_m -> D), ReLU, Dropout, Linear(D -> C); fusion_logits (M) */
typedef struct {
  int num_modalities, hidden_dim, num_classes;
  int modality_dim[AVHOT_MAX_MODALITIES];
  const float* w0[AVHOT_MAX_MODALITIES];
  const float* b0[AVHOT_MAX_MODALITIES];
  const float* w3[AVHOT_MAX_MODALITIES];
  const float* b3[AVHOT_MAX_MODALITIES];
  const float* fusion_logits;
} avhot_late_desc;
int avhot_early_create(const avhot_early_desc* desc, avhot_head_t* out, avhot_stream_t stream);
int avhot_late_create(const avhot_late_desc* desc, avhot_head_t* out, avhot_stream_t stream);
void avhot_head_destroy(avhot_head_t h);
/* forward(modality_features, modality_mask) for R = B*T rows; z / z_rows_per_clip / mask as in avhot_fusion_forward.
 * present_mask: bit m set = modality m was given (LateFusion gives a missing modality zero logits, :222-225;
 * EarlyFusion callers pass zeros, :120-121).  per_modality_logits: (M, R, C), LateFusion only, may be NULL.
 * Dropout (p > 0): sites AVHOT_SITE_EARLY0 / EARLY1 (mlp.3, mlp.7) and AVHOT_SITE_LATE + m (classifiers.<m>.2). */
int avhot_head_forward(avhot_head_t h, const float* const* z_host_array, const int* z_rows_per_clip_host,
                       int present_mask, const float* mask, float* logits, float* per_modality_logits,
                       int B, int T, float p, uint64_t seed, uint64_t clip0, uint32_t t0,
                       avhot_stream_t stream);
/* src/fusion.py:589-655  LateFusionWithUncertainty (build_fusion_model('uncertainty')): per modality a classifier and a
 * scalar-uncertainty head - hidden_dim > 0: Dropout, Linear(d, D), ReLU, Dropout, Linear(D, C | 1); hidden_dim == 0:
 * Dropout, Linear(d, C | 1) - softplus on the uncertainty, then UncertaintyWeightedFusion (src/uncertainty.py:279-318):
 * w_m = (mask_m / (u_m + eps)) / (sum_m + eps).  Tensors with hidden_dim == 0: only *_w3 / *_b3 (the single Linear). */
typedef struct {
  int num_modalities, hidden_dim, num_classes;
  int modality_dim[AVHOT_MAX_MODALITIES];
  const float *cls_w0[AVHOT_MAX_MODALITIES], *cls_b0[AVHOT_MAX_MODALITIES];   /* classifiers.<m>.1  (D, d), (D)  */
  const float *cls_w3[AVHOT_MAX_MODALITIES], *cls_b3[AVHOT_MAX_MODALITIES];   /* classifiers.<m>.4  (C, D), (C)  */
  const float *unc_w0[AVHOT_MAX_MODALITIES], *unc_b0[AVHOT_MAX_MODALITIES];   /* uncert_heads.<m>.1 (D, d), (D)  */
  const float *unc_w3[AVHOT_MAX_MODALITIES], *unc_b3[AVHOT_MAX_MODALITIES];   /* uncert_heads.<m>.4 (1, D), (1)  */
  float epsilon;
} avhot_unc_desc;
int avhot_unc_create(const avhot_unc_desc* desc, avhot_head_t* out, avhot_stream_t stream);   /* avhot_head_destroy frees */
/* forward(encoded_features, modality_mask) for R = B*T rows: logits (R, C), per_modality_logits (R, M, C) or NULL,
 * fusion_weights (R, M) or NULL; mask (B, M) is required (:622).  A modality that is not given (bit clear in
 * present_mask) is zeros (:629).  Dropout sites AVHOT_SITE_UNC + 4 m + {0 classifier input, 1 classifier hidden,
 * 2 uncertainty-head input, 3 uncertainty-head hidden}. */
int avhot_unc_forward(avhot_head_t h, const float* const* z_host_array, const int* z_rows_per_clip_host,
                      int present_mask, const float* mask, float* logits, float* per_modality_logits,
                      float* fusion_weights, int B, int T, float p, uint64_t seed, uint64_t clip0, uint32_t t0,
                      avhot_stream_t stream);
/* UncertaintyWeightedFusion.forward on its own: logits (R, M, C), uncert (R, M), mask (R, M) -> fused (R, C), weights (R, M) */
int avhot_uncertainty_fuse(const float* logits, const float* uncert, const float* mask, float* fused, float* weights,
                           long long R, int M, int C, float eps, avhot_stream_t stream);
/* src/uncertainty.py:88-183, :512-544  calibration statistics, accumulated into acc (device, doubles, caller-zeroed):
 * [0,nb) bin counts, [nb,2nb) sum of confidences, [2nb,3nb) correct counts, [3nb] NLL sum (logits given),
 * [3nb+1] correct total, [3nb+2] samples.  Either logits (N, C) - confidence = max softmax, prediction = argmax,
 * NLL = cross-entropy - or (confidences, predictions) as CalibrationMetrics' static methods take them.
 * bin_edges: (num_bins + 1) fp32, torch.linspace(0, 1, num_bins + 1); last bin is closed on the right. */
int avhot_calibration_accumulate(const float* logits, const float* confidences, const int64_t* predictions,
                                 const int64_t* labels, long long N, int C, const float* bin_edges,
                                 int num_bins, double* acc, avhot_stream_t stream);

/* ---- src/encoders.py:96-113, :170-192  SequenceEncoder(encoder_type='transformer') ----------
 * input_proj -> + pos_embedding[min(t, max_len - 1)] -> num_layers x nn.TransformerEncoderLayer (post-norm, GELU,
 * nhead heads, dim_feedforward = ffn_dim) with key padding from `lengths` -> mean over the valid steps ->
 * projection.  seq (B, T, input_dim) fp32, lengths (B) int32 or NULL -> z (B, passes, output_dim).  Pass t draws the
 * masks of (seed, t0 + t, clip0 + b) at sites AVHOT_SITE_TFM + 3 l + {0 dropout1 (T, H), 1 dropout (T, ffn),
 * 2 dropout2 (T, H)} of layer l; p = 0 is the eval forward.  Weights are nn.Linear / nn.MultiheadAttention layouts:
 * in_proj (3H, H) = [q | k | v] rows.  precision: AVHOT_PREC_FP32, or AVHOT_PREC_TC = every nn.Linear on the tcgen05
 * TF32 GEMM (stated tolerance 2e-3), attention / LayerNorm / GELU / pooling in fp32 either way. */
#define AVHOT_TFM_MAX_LAYERS 4
typedef struct avhot_tfm* avhot_tfm_t;
typedef struct {
  int input_dim, hidden_dim, output_dim, num_layers, num_heads, ffn_dim, max_len;
  float ln_eps;
  const float *input_w, *input_b, *pos, *proj_w, *proj_b;
  const float *in_proj_w[AVHOT_TFM_MAX_LAYERS], *in_proj_b[AVHOT_TFM_MAX_LAYERS];
  const float *out_w[AVHOT_TFM_MAX_LAYERS], *out_b[AVHOT_TFM_MAX_LAYERS];
  const float *lin1_w[AVHOT_TFM_MAX_LAYERS], *lin1_b[AVHOT_TFM_MAX_LAYERS];
  const float *lin2_w[AVHOT_TFM_MAX_LAYERS], *lin2_b[AVHOT_TFM_MAX_LAYERS];
  const float *n1_w[AVHOT_TFM_MAX_LAYERS], *n1_b[AVHOT_TFM_MAX_LAYERS];
  const float *n2_w[AVHOT_TFM_MAX_LAYERS], *n2_b[AVHOT_TFM_MAX_LAYERS];
} avhot_tfm_desc;
int avhot_tfm_create(const avhot_tfm_desc* desc, avhot_tfm_t* out, avhot_stream_t stream);
void avhot_tfm_destroy(avhot_tfm_t h);
size_t avhot_tfm_workspace_bytes(avhot_tfm_t h, int B, int T, int passes);
int avhot_tfm_forward(avhot_tfm_t h, const float* seq, const int32_t* lengths, float* z, int B, int T, int passes,
                      float p, uint64_t seed, uint64_t clip0, uint32_t t0, int precision, void* workspace,
                      size_t workspace_bytes, avhot_stream_t stream);

/* ---- src/attention.py:143-274  TemporalAttention -------------------------------------------
 * forward(sequence (B, S, feature_dim), mask (B, S) 1 = valid or NULL) -> attended (B, S, hidden), weights
 * (B, heads, S, S) (may be NULL).  Masked keys score -inf, the context of masked queries is multiplied by 0
 * (:219-236); the reference never applies its two nn.Dropout modules (its forward tests `hasattr(self, "attn_drop")`,
 * an attribute that does not exist, :227, :240), so the module is deterministic. */
typedef struct avhot_tattn* avhot_tattn_t;
typedef struct {
  int feature_dim, hidden_dim, num_heads;
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b;
} avhot_tattn_desc;
int avhot_tattn_create(const avhot_tattn_desc* desc, avhot_tattn_t* out, avhot_stream_t stream);
void avhot_tattn_destroy(avhot_tattn_t h);
size_t avhot_tattn_workspace_bytes(avhot_tattn_t h, int B, int S);
int avhot_tattn_forward(avhot_tattn_t h, const float* seq, const uint8_t* valid, float* attended, float* weights, int B,
                        int S, void* workspace, size_t workspace_bytes, avhot_stream_t stream);
/* pool_sequence (:245-274): key importance = mean over heads and queries of the weights, renormalised with + 1e-9;
 * pooled (B, D) = importance-weighted sum of sequence (B, S, D). */
int avhot_tattn_pool(const float* seq, const float* weights, float* pooled, int B, int S, int D, int heads,
                     avhot_stream_t stream);

/* ---- src/attention.py:421-548  PairwiseModalityAttention, aggregation per target modality ----
 * out (B, H) = LayerNorm(self_feat + dropout(sum_j msgs[j])) * avail   (:527-545).  msgs_host_array: n_msgs <= 3 device
 * pointers (B, H), summed left to right; avail (B) bytes or NULL; dropout site = `site` (AVHOT_SITE_PAIR_MSG + m),
 * element = column.  The directional messages are CrossModalAttention outputs (avhot_xattn_forward). */
int avhot_pair_combine(const float* self_feat, const float* const* msgs_host_array, int n_msgs, const uint8_t* avail,
                       const float* ln_w, const float* ln_b, float ln_eps, float* out, int B, int H, float p,
                       uint64_t seed, uint64_t clip0, uint32_t t, uint32_t site, avhot_stream_t stream);


/* ---- a8: src/uncertainty.py:17-86  MCDropoutUncertainty statistics ------------------ */
/* logits (B, T, C) -> mean_logits (B, C), uncertainty (B) = mean_c Var_t softmax (population
 * variance, src/uncertainty.py:77-84), labels (B) = argmax mean_logits (first max wins,
 * like torch.argmax).  packed (B, C+1) = [mean_logits | uncertainty] for the all-gather
 * (any output may be NULL). */
int avhot_mc_reduce(const float* logits, float* mean_logits, float* uncertainty,
                    int32_t* labels, float* packed, int B, int T, int C,
                    avhot_stream_t stream);

/* ---- label exactness of AVHOT_PREC_TC (north star: labels identical to the reference, whose prediction is
 * argmax of the fp32 mean logits, src/uncertainty.py:77).  A clip whose top-2 margin of mean_logits is below
 * `margin` (absolute: the caller's bound on the error of a logit difference; exact.py derives it from the classifier's
 * weights, so it does not depend on the batch) is appended to a pending list: entry e < capacity gets entry_tag[e] =
 * tag (the caller's id of this call), entry_row[e] = b, entry_clip[e] = low 32 bits of clip0 + b.  state: 4 uint32 on
 * the device, zeroed by the caller when a window of calls starts: [0] entries appended (can exceed capacity),
 * [1] overflow flag, [2] entries before the latest call.  No host synchronisation. */
int avhot_flag_near_ties(const float* mean_logits, int B, int C, float margin, uint32_t tag,
                         uint64_t clip0, uint32_t* state, uint32_t* entry_tag, uint32_t* entry_row,
                         uint32_t* entry_clip, int capacity, avhot_stream_t stream);
/* rows entry_row[e] of src (B rows of row_bytes bytes, a multiple of 4) -> pool row e, for the entries the latest
 * avhot_flag_near_ties call appended: the inputs of the pending clips, kept for their fp32 re-evaluation */
int avhot_gather_flagged(const void* src, size_t row_bytes, int B, const uint32_t* state,
                         const uint32_t* entry_row, void* pool, int capacity, avhot_stream_t stream);
/* records of the n re-evaluated entries (src_packed (n, C+1), src_labels (n)) -> the result tensors of the call
 * whose tag is `tag` (rows entry_row[e] where entry_tag[e] == tag); any destination may be NULL */
int avhot_scatter_resolved(const uint32_t* entry_tag, const uint32_t* entry_row, int n, uint32_t tag,
                           const float* src_packed, const int32_t* src_labels, float* mean_logits,
                           float* uncertainty, int32_t* labels, float* packed, int C,
                           avhot_stream_t stream);

#define AVHOT_SITE_FRAME_MLP 0u
#define AVHOT_SITE_FRAME_POOL 1u
#define AVHOT_SITE_GATE 2u
#define AVHOT_SITE_MODMASK 3u
#define AVHOT_SITE_EARLY0 4u
#define AVHOT_SITE_EARLY1 5u
#define AVHOT_SITE_CNN0 6u
#define AVHOT_SITE_CNN1 7u
#define AVHOT_SITE_ATTN 16u
#define AVHOT_SITE_LATE 32u
#define AVHOT_SITE_UNC 48u
#define AVHOT_SITE_TFM 64u      /* + 3 layer + {0, 1, 2} */
#define AVHOT_SITE_PAIR_MSG 80u /* + target modality index */

#ifdef __cplusplus
}
#endif
#endif /* AVHOT_H_ */
