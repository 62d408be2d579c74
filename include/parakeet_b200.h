/* parakeet_b200 C-ABI: B200 (sm_100a) kernels for the Parakeet TTS hot path.
 *
 * The reference (PaddlePaddle/Parakeet) has no FFI of its own: its hot path is Python calling paddle.nn ops.
 * This header is the boundary a Parakeet maintainer would bind instead of those ops (ctypes stub in
 * INTEGRATION.md).  Every entry point names the reference call site (file:line under /root/reference) it replaces.
 *
 * Rules:
 *   - plain C: raw device pointers, sizes, a cudaStream_t passed as void*; no torch / C++ types;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library owns nothing but its code;
 *   - every function returns 0 (PK_OK) or a negative error code, never throws, never aborts, never falls back to
 *     a CPU path; pk_last_error() returns a thread-local message for the last failing call;
 *   - all work is enqueued on the caller's stream; no internal synchronisation unless documented.
 *
 * Number format of GEMM operands ("split-bf16"): an fp32 tensor v is carried as two bf16 planes
 * hi = bf16(v), lo = bf16(v - hi).  Producers in this library write both planes; pk_split_f32 converts.
 */
#ifndef PARAKEET_B200_H_
#define PARAKEET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PK_OK 0
#define PK_ERR_INVALID_ARG (-1)
#define PK_ERR_CUDA (-2)
#define PK_ERR_UNSUPPORTED (-3)

#define PK_ACT_NONE 0
#define PK_ACT_RELU 1
#define PK_ACT_TANH 2

typedef void* pk_stream_t; /* cudaStream_t */

/* Library version (major*10000 + minor*100 + patch). */
int pk_version(void);
/* Thread-local description of the last error returned on this thread ("" if none). */
const char* pk_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches claim). */
int64_t pk_launch_count(void);

/* fp32 -> split-bf16 planes; n elements (any layout, element-wise). */
int pk_split_f32(const float* x, void* hi, void* lo, int64_t n, pk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Generic channels-last Conv1D / Linear / batched-matmul on tcgen05 tensor cores.
 *
 *   y[z, t, n] = act( scale * sum_{tap, k} A[za, t + (tap - pad) * dil, a_col + k] * B[zb, n, b_col + tap*Kp + k]
 *                     + bias[n] ) + residual[z, t, n]          ; rows t >= lens[b] are written as 0
 *
 * with z = b * heads + h, za = b * a_bmul + h * a_hmul, a_col = a_col0 + h * a_colh (same for B), Kp = K rounded
 * up to 64.  Rows / channels outside the declared extents read as zero (conv zero padding).
 *
 * Replaces, in the reference: nn.Linear call sites of MultiHeadedAttention (modules/fastspeech2_transformer/
 * attention.py:44-47,74-78,131), the two matmuls (:153-154,126), MultiLayeredConv1d (multi_layer_conv.py:47-77),
 * predictor Conv1D stacks (duration_predictor.py:69-83, variance_predictor.py:59-76), feat_out
 * (models/fastspeech2/fastspeech2.py:271,457) and Postnet convs (modules/tacotron2/decoder.py:128-180).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct pk_operand {
  const void* hi;         /* bf16 plane */
  const void* lo;         /* bf16 plane, same layout */
  int64_t batch_stride;   /* elements between consecutive "batches" of the plane */
  int32_t ld;             /* elements between consecutive rows (multiple of 8) */
  int32_t rows;           /* rows per batch; rows outside [0, rows) read as zero */
  int32_t cols;           /* channels per row; channels outside [0, cols) read as zero */
  int32_t batches;        /* number of batches in the plane */
  int32_t bmul, hmul;     /* batch index used for (b, h) = b*bmul + h*hmul */
  int32_t col0, colh;     /* first channel used for head h = col0 + h*colh */
} pk_operand;

typedef struct pk_conv_gemm_args {
  pk_operand a;           /* activations: rows = time */
  pk_operand b;           /* weights [n][taps*Kp] (bmul = hmul = 0) or a per-(b,h) matrix [n][K] */
  int32_t batch, heads;   /* grid z = batch * heads */
  int32_t m;              /* output rows per batch (time steps) */
  int32_t n;              /* output channels */
  int32_t k;              /* input channels per tap */
  int32_t taps, dil, pad; /* Conv1D kernel size, dilation, left padding in taps ((taps-1)/2 for "same") */
  float scale;            /* applied to the accumulator before bias */
  const float* bias;      /* [n] or NULL */
  int32_t act;            /* PK_ACT_* */
  const float* residual;  /* fp32, indexed like y_f32, or NULL (added after act) */
  const int32_t* lens;    /* [batch] valid rows per batch or NULL */
  float* y_f32;           /* fp32 output or NULL */
  void* y_hi;             /* split-bf16 output planes or NULL (both or none) */
  void* y_lo;
  int64_t y_batch_stride; /* elements; y offset = b*y_batch_stride + h*y_head_stride + t*y_ld + n */
  int64_t y_head_stride;
  int32_t y_ld;
  int32_t passes;         /* 3 = split-bf16 (fp32-grade), 1 = plain bf16 (hi planes only) */
} pk_conv_gemm_args;

int pk_conv_gemm(const pk_conv_gemm_args* args, pk_stream_t stream);

/* pk_conv_gemm with a fused pair epilogue: the output row has n = 2 * channels columns, column c of the first half is
 * combined with column channels + c of the second half (bias and scale applied to both) and the GEMM result itself is
 * never written.  n <= 256, channels % 32 == 0, heads == 1; act / residual / lens of the base arguments must be unset.
 *   PK_EPI_GATE       z = tanh(a + res_a) * sigmoid(g + res_g) written as split planes y_hi / y_lo (batch, m, y_ld);
 *                     `residual` (fp32, batch stride / row pitch given, 2 * channels columns) may be NULL.  Fuses the gate of
 *                     waveflow.ResidualBlock (models/waveflow.py:277-281) into its dilated-conv GEMM.
 *   PK_EPI_WF_UPDATE  state (batch, m, channels) += a;  skip = skip_init ? g : skip + g;  the new state is also written as
 *                     split planes into buf_hi / buf_lo (batch, m, buf_ld) at column buf_col0 when given.  Fuses
 *                     `res, skip = split(out_proj(z))` (models/waveflow.py:282-294) and ResidualNet.add_input (:386-392). */
enum { PK_EPI_NONE = 0, PK_EPI_GATE = 1, PK_EPI_WF_UPDATE = 2 };
typedef struct pk_gemm_epilogue {
  int32_t mode;                  /* PK_EPI_* */
  int32_t channels;              /* C: n == 2 * C */
  const float* residual;         /* GATE */
  int64_t residual_batch_stride;
  int32_t residual_ld;
  int32_t skip_init;             /* WF_UPDATE */
  float* state;
  float* skip;
  void* buf_hi;
  void* buf_lo;
  int32_t buf_ld, buf_col0;
} pk_gemm_epilogue;
int pk_conv_gemm_ex(const pk_conv_gemm_args* args, const pk_gemm_epilogue* epilogue, pk_stream_t stream);

/* Same contract evaluated with plain fp32 FMAs (one thread per output element).  Debug / cross-check kernel for
 * the tensor-core path at sizes where the CPU oracle is too slow; not used by the models. */
int pk_conv_gemm_simt(const pk_conv_gemm_args* args, pk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Length regulator: integer repeat-expand, bit-exact row copies.
 *   y[b, cum[b,j] + r, :] = x[b, j, :] for 0 <= r < d[b,j];  rows >= sum_j d[b,j] are 0 up to t_out.
 * out_lens[b] = sum_j max(d[b,j],0)... (durations are clipped >= 0 upstream; negative d is rejected as in the
 * reference's reachable domain).  Replaces LengthRegulator.expand (modules/fastspeech2_predictor/
 * length_regulator.py:46-66): host numpy loop + 0/1-matrix matmul.
 * ------------------------------------------------------------------------------------------------------------ */
/* out_lens[b] = sum_j d[b, j] (int32), device-side; t_in tokens per utterance. */
int pk_length_regulator_lens(const int64_t* dur, int32_t batch, int32_t t_in, int32_t* out_lens, pk_stream_t stream);
/* x: fp32 (batch, t_in, c); y: fp32 (batch, t_out, c), optional split planes y_hi / y_lo (NULL to skip). */
int pk_length_regulate(const float* x, const int64_t* dur, int32_t batch, int32_t t_in, int32_t c, int32_t t_out,
                       float* y, void* y_hi, void* y_lo, pk_stream_t stream);


/* ------------------------------------------------------------------------------------------------------------
 * Parallel WaveGAN generator (reference: parakeet/models/parallel_wavegan/parallel_wavegan.py).
 * Activations are channels-last split-bf16 planes (B, T, C); conditioning c is (B, T, aux).
 * ------------------------------------------------------------------------------------------------------------ */
/* ConvInUpsampleNet.forward (:201-216) = conv_in (Conv1D aux->aux, k = 2*window+1, no padding, no bias) followed by
 * UpsampleNet.forward (:119-138): per scale s, nearest stretch x s (Stretch2D :48-63) then a (1, 2s+1) FIR with zero
 * padding s.  mel: device fp32 (batch, aux, frames + 2*window), channel-first like the reference;
 * conv_in_w: device fp32 [aux][aux][2*window+1]; fir: HOST fp32, the n_stages FIRs concatenated (2*s_k+1 taps each);
 * scales: HOST int32 [n_stages].  Outputs (either or BOTH may be NULL - then only conv_in_ws is produced, which is all the frame-rate path needs): c_f32 (batch, aux, T) channel-first fp32 and
 * c_hi/c_lo (batch, T, aux) channels-last split-bf16, T = frames * prod(scales).  frame_lens (device int32 [batch] or
 * NULL): valid frames per utterance of a ragged batch; each utterance is then upsampled exactly as if alone
 * (zero padding at its own end) and its samples past frame_lens[b]*hop are written as zero.
 * conv_in_ws: device fp32 workspace (batch, frames, aux) that receives the conv_in output (channels-last); aux % 8 == 0.
 * Two launches: conv_in per frame, then the fused stretch/FIR cascade. */
int pk_pwg_upsample(const float* mel, const float* conv_in_w, const float* fir, const int32_t* scales, int32_t n_stages,
                    int32_t batch, int32_t aux, int32_t frames, int32_t window, const int32_t* frame_lens,
                    float* conv_in_ws, float* c_f32, void* c_hi, void* c_lo, pk_stream_t stream);

/* first_conv (:401-402, :464): x[b,t,r] = w[r] * noise[b,t] + bias[r] for 64 residual channels, written as split
 * planes (batch, t, 64); rows t >= lens[b] are written as zero (lens may be NULL). */
int pk_pwg_first_conv(const float* noise, const float* w, const float* bias, const int32_t* lens, int32_t batch, int32_t t,
                      void* x_hi, void* x_lo, pk_stream_t stream);

/* One fused ResidualBlock.forward (:284-315) for residual = skip = 64 channels, gate = 128, kernel 3:
 *   h = conv_k3_dil(x) + b1 + conv1x1_aux(c);  z = tanh(h[:64]) * sigmoid(h[64:]);
 *   skip (+)= conv1x1_skip(z)   [its bias is added once, in pk_pwg_tail];  y = (conv1x1_out(z) + b_out + x) * sqrt(0.5)
 * w1: packed [128][5*64] K-major = taps 0..2 (64 ch each) then the aux weight zero-padded to 128 channels;
 * w2: packed [128][64], rows 0..63 = conv1x1_skip, rows 64..127 = conv1x1_out; bias1 [128]; bias2 = skip bias | out bias.
 * y must not alias x.  Rows t >= lens[b] of y are written as zero and tiles wholly past lens[b] are skipped
 * (their y rows must already be zero). */
typedef struct pk_pwg_layer_args {
  int32_t batch, t, dilation, aux_channels;
  const int32_t* lens;     /* [batch] valid samples or NULL */
  const void* x_hi;        /* input planes (batch, t, 64) */
  const void* x_lo;
  void* y_hi;              /* output planes (batch, t, 64) */
  void* y_lo;
  const void* c_hi;        /* conditioning planes (batch, t, aux_channels) */
  const void* c_lo;
  const void* w1_hi;
  const void* w1_lo;
  const void* w2_hi;
  const void* w2_lo;
  const float* bias1;      /* HOST pointer [128]: conv bias (gate a | gate g); copied into the kernel parameter block */
  const float* bias2;      /* HOST pointer [128]: conv1x1_skip bias (ignored, see pk_pwg_tail) | conv1x1_out bias */
  float* skip;             /* fp32 (batch, t, 64) running sum of skips */
  int32_t skip_init;       /* 1: overwrite (first layer), 0: accumulate */
  void* prof;              /* debug: NULL, or device uint64[64] phase-cycle counters accumulated by the kernel
                              ([0..1] producer, [8..14] MMA issuer, [16..22]/[24..30] epilogue halves, [32] tiles) */
} pk_pwg_layer_args;
int pk_pwg_residual_layer(const pk_pwg_layer_args* args, pk_stream_t stream);

/* ResidualBlock with FRAME-RATE CONDITIONING (the default residual-stack path of the Python model; PK_PWG_FRAME_COND=0
 * selects pk_pwg_residual_layer).  The upsampling network is linear and per channel, so conv1x1_aux(upsample(m'))[t, n] =
 * sum_j U[t, j] * P[j, n] with P = conv1x1_aux applied to m' = conv_in(mel) at FRAME rate.  Instead of the sample-rate
 * conditioning planes (1.2 GB at cfg 2) the kernel takes
 *   u_hi / u_lo: the COMPACT band table of U as split planes (u_rows, 64) from ONE allocation (lo after hi): a row holds
 *                U[t, j0 + k] in column k < 16 (zeros after), j0 = floor8(((t / 256) * 256) / hop - 2) (the K window of
 *                the 256-sample pair tile of t).  Rows [0, u_period): interior rows by t mod u_period; rows
 *                [u_start_row, +128): the first 128 samples of any utterance; rows [u_end_base + 384 b, +384): the half
 *                tiles of utterance b from 128 * ((len_b - 128) / 128) on (zero at and past len_b).  Built by
 *                parakeet_b200/models/_pwg_frame_cond.py: compact_band_tables (lookup: source_row).
 *   p_hi / p_lo: P as split planes (batch, p_rows, p_ld), one allocation, frames along the last axis (p_frames valid
 *                columns, zeros up to p_ld >= 64); this layer's 128 output channels are rows [p_row0, p_row0 + 128).
 * x / y planes: one allocation each (lo after hi) so that a tile of both planes is ONE 4-D TMA box.
 * w1 / w2 / bias1 / bias2 / skip / lens as in pk_pwg_layer_args (the aux columns of w1 are not read).  hop >= 256. */
typedef struct pk_pwg_layer_fc_args {
  int32_t batch, t, dilation, hop;
  const int32_t* lens;
  const void* x_hi;
  const void* x_lo;
  void* y_hi;
  void* y_lo;
  const void* u_hi;
  const void* u_lo;
  int32_t u_rows, u_period, u_start_row, u_end_base;
  int32_t p_rows, p_ld, p_frames, p_row0;
  const void* p_hi;
  const void* p_lo;
  const void* w1_hi;
  const void* w1_lo;
  const void* w2_hi;
  const void* w2_lo;
  const float* bias1;      /* HOST pointers, as in pk_pwg_layer_args */
  const float* bias2;
  float* skip;
  int32_t skip_init;
  void* prof;
} pk_pwg_layer_fc_args;
int pk_pwg_residual_layer_fc(const pk_pwg_layer_fc_args* args, pk_stream_t stream);

/* last_conv_layers (:429-440) on the scaled skip sum (:469-471):
 *   out[row] = w2 . relu(W1 relu((skip[row] + skip_bias) * scale) + b1) + b2
 * with skip fp32 (rows, 64), W1 [64 out][64 in], w2 [64]; out fp32 (rows).  skip_bias [64] (or NULL) is the sum of the
 * layers' conv1x1_skip biases: pk_pwg_residual_layer accumulates the skip branch WITHOUT its bias (bias2[0..63] is
 * ignored there) and the constant vector is added once here. */
int pk_pwg_tail(const float* skip, const float* skip_bias, const float* w1, const float* b1, const float* w2, const float* b2,
                float scale, int64_t rows, float* out, pk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * FastSpeech2 row-wise kernels (reference: parakeet/models/fastspeech2/fastspeech2.py and the files under parakeet/modules).
 * Shapes are (batch, t, channels) channels-last fp32 unless noted.  `lens` (device int32 [batch] or NULL) selects
 * the "independent utterances" mode used for batched inference: rows t >= lens[b] are written as zero so that every
 * utterance sees exactly the zero padding it would see alone; with lens == NULL padded rows are computed like any
 * other row, which is what the reference's batched training forward does.
 * ------------------------------------------------------------------------------------------------------------ */
/* Embedding(padding_idx) + ScaledPositionalEncoding (modules/fastspeech2_transformer/embedding.py:46-62,111-126;
 * encoder.py:103-105): y = (ids ? table[ids] (zeros for padding_idx) : x_in) + alpha[0] * PE. */
int pk_embed_pe(const int64_t* ids, const float* table, int32_t vocab, int32_t padding_idx, const float* x_in,
                const float* alpha, const int32_t* lens, int32_t batch, int32_t t, int32_t d, float* y, pk_stream_t stream);
/* nn.LayerNorm over the last dim (encoder_layer.py:55-56,85,107; encoder.py:143,191; modules/layer_norm.py:47-63).
 * Outputs: y fp32 and/or split planes (either may be NULL). */
int pk_layer_norm(const float* x, const float* gamma, const float* beta, float eps, const int32_t* lens, int32_t batch,
                  int32_t t, int32_t d, float* y, void* y_hi, void* y_lo, pk_stream_t stream);
/* masked_fill(min) -> softmax -> masked_fill(0) of attention.py:107-119 over keys: s fp32 (batch*heads, rows, ld),
 * keys >= key_lens[b] and the padding columns [keys, ld) get probability 0; output split planes, same layout. */
int pk_masked_softmax(const float* s, const int32_t* key_lens, int32_t batch, int32_t heads, int32_t rows, int32_t keys,
                      int32_t ld, void* p_hi, void* p_lo, pk_stream_t stream);
/* paddle.nn.functional.normalize(x, p=2, axis=1, epsilon) on x viewed as (outer, n, inner), norm over the middle axis: the
 * speaker / tone embedding normalisation of FastSpeech2 (fastspeech2.py:577,581,606,611; axis 1 of a (B, T, D) tone tensor is
 * TIME in the reference's batched forward).  x, y fp32 contiguous; in place allowed. */
int pk_l2_normalize(const float* x, int32_t outer, int32_t n, int32_t inner, float eps, float* y, pk_stream_t stream);
/* Fused scaled-dot-product attention of an FFT block (attention.py:88-131: scores = q k^T / sqrt(d_k), masked_fill(min) ->
 * softmax -> masked_fill(0), p_attn . v, heads merged) in one kernel: scores and probabilities stay in tensor memory.
 *   qkv planes (batch, t, 3 * heads * dk): [q | k | v] of the fused QKV projection, head h in columns h * dk of each third;
 *   vt planes (batch * heads, dk, tp): v transposed per head (pk_transpose_heads), columns >= t zero; both pairs of planes
 *   from one allocation each (lo after hi).  key_lens / row_lens: device int32 [batch] or NULL (keys >= key_lens[b] masked;
 *   query rows >= row_lens[b] written as zero).  ctx planes (batch, t, heads * dk).  dk in {64, 128, 192}. */
int pk_fused_attention(const void* qkv_hi, const void* qkv_lo, const void* vt_hi, const void* vt_lo, int32_t batch, int32_t t,
                       int32_t heads, int32_t dk, int32_t tp, const int32_t* key_lens, const int32_t* row_lens, float scale,
                       void* ctx_hi, void* ctx_lo, pk_stream_t stream);
/* (batch, t, ld_src)[.., col0 + h*dk + d] -> (batch*heads, dk, ld_dst)[.., d, t] (zero-filled for t in [t, ld_dst)):
 * the value matrix in K-major form for the P.V product (attention.py:126). */
int pk_transpose_heads(const void* src_hi, const void* src_lo, int32_t batch, int32_t t, int32_t ld_src, int32_t col0,
                       int32_t dk, int32_t heads, int32_t ld_dst, void* dst_hi, void* dst_lo, pk_stream_t stream);
/* DurationPredictor.inference post-op (duration_predictor.py:94-101): d = max(round_half_away(exp(x) - offset), 0),
 * padded tokens -> 0.  Outputs fp32 and/or int64 (either may be NULL). */
int pk_duration_post(const float* x, const int32_t* lens, int32_t batch, int32_t t, float offset, float* d_f32, int64_t* d_i64,
                     pk_stream_t stream);
/* LengthRegulator.forward alpha != 1 (length_regulator.py:85-88): out = int64(round_half_away(d * alpha)). */
int pk_duration_scale(const int64_t* d, float alpha, int64_t n, int64_t* out, pk_stream_t stream);
/* masked_fill(x, pad_mask, 0) for predictor outputs (variance_predictor.py:101-103): x (batch, t, inner) in place. */
int pk_mask_rows(float* x, const int32_t* lens, int32_t batch, int32_t t, int32_t inner, pk_stream_t stream);
/* hs + pitch_embed(p) + energy_embed(e) (fastspeech2.py:426-430): Conv1D(1 -> c, k, pad (k-1)/2) on the scalar tracks
 * pitch / energy (batch, t); wp/we are [c][k] (Paddle [c,1,k]), bp/be [c]. */
int pk_variance_embed_add(const float* hs, const float* pitch, const float* energy, const float* wp, const float* bp, int32_t kp,
                          const float* we, const float* be, int32_t ke, const int32_t* lens, int32_t batch, int32_t t, int32_t c,
                          float* y, pk_stream_t stream);
/* ZScore.forward / .inverse over the last dim (modules/normalizer.py:18-33): inverse == 0: (x - mu) / sigma,
 * inverse != 0: x * sigma + mu; n = total elements, c = channels. */
int pk_zscore(const float* x, const float* mu, const float* sigma, int32_t c, int64_t n, int32_t inverse, float* y,
              pk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * STFT / mel front-end: batched radix-2 FFT (one CTA per frame, no cuFFT) with fused epilogues.
 * Replaces STFT.forward / .power / .magnitude (parakeet/modules/audio.py:161-215: reflect pad + conv1d with the
 * windowed DFT matrix), MelScale.forward (:218-229), stft() of modules/stft_loss.py:20-67 and the numpy/librosa
 * feature path of data/get_feats.py:47-88 (log-mel) and :196-203 (frame energy).
 *   x (batch, t) fp32; window (n_fft) fp32 = scipy get_window(fftbins=True) centre-padded to n_fft (host builds it);
 *   twiddle (n_fft/2) float2 = (cos, -sin)(2 pi j / n_fft); center != 0: reflect padding n_fft/2, frames = 1 + t/hop.
 * Outputs (any may be NULL): re / im (batch, bins, frames); mag = sqrt(max(re^2+im^2, power_clip)) (power_clip < 0:
 * no clip) in layout 0 (batch, bins, frames) or 1 (batch, frames, bins); mel (batch, frames, n_mels) = mel_w (n_mels,
 * bins) . mag, optionally log10(max(., mel_clip)); energy (batch, frames) = sqrt(max(sum_k |X|^2, energy_clip)).
 * ------------------------------------------------------------------------------------------------------------ */
int pk_stft(const float* x, int32_t batch, int32_t t, const float* window, const void* twiddle, int32_t n_fft, int32_t hop,
            int32_t center, float* re, float* im, float* mag, int32_t mag_layout, float power_clip, const float* mel_w,
            int32_t n_mels, float* mel, int32_t mel_log10, float mel_clip, float* energy, float energy_clip, pk_stream_t stream);

/* Reductions behind SpectralConvergenceLoss / LogSTFTMagnitudeLoss (modules/stft_loss.py:70-118) on two magnitude
 * spectrograms of n elements: out3 = { sum (y-x)^2, sum y^2, sum |log max(y,eps) - log max(x,eps)| } (device fp32[3]). */
int pk_spectral_loss_sums(const float* x_mag, const float* y_mag, int64_t n, float eps, float* out3, pk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * WaveFlow inference (reference: parakeet/models/waveflow.py).  The per-row residual net of Flow.inverse (:515-556,
 * ResidualBlock.add_input :248-294) runs its three GEMMs per layer through pk_conv_gemm on a channels-last row
 * (batch, W, C): the 3-row causal buffer is a (batch, W, 3C) split-bf16 ring (slot = row mod 3) convolved along W with
 * 3 dilated taps and K = 3C; the kernels below are the row-wise glue.
 * ------------------------------------------------------------------------------------------------------------ */
/* One layer of waveflow.UpsampleNet.forward (:103-132): Conv2DTranspose(1,1,(3,2f),stride (1,f),padding (1,f/2)) over
 * (mel, time), trim of the last f columns if trim != 0, leaky_relu(slope).  x (batch, c, t_in) -> y (batch, c,
 * t_in*f - trim*f); w [3][2f] (Paddle [1,1,3,2f]), bias [1]. */
int pk_waveflow_upsample(const float* x, const float* w, const float* bias, int32_t batch, int32_t c, int32_t t_in, int32_t factor,
                         int32_t trim, float slope, float* y, pk_stream_t stream);
/* Flow.input_proj (:437-442, 1x1 conv 1 -> C) on one row: state[b,w,:] = w * x_row[b,w] + bias (fp32 (batch, W, C)) and
 * the same values as split planes into columns [col0, col0 + C) of a (batch, W, ld) ring buffer. */
int pk_waveflow_input_proj(const float* x_row, int64_t x_batch_stride, const float* w, const float* bias, int32_t batch,
                           int32_t width, int32_t c, float* state, void* buf_hi, void* buf_lo, int32_t ld, int32_t col0,
                           pk_stream_t stream);
/* tanh(h[:, :c]) * sigmoid(h[:, c:]) (waveflow.py:280-281, also parallel_wavegan.py:310-311): h fp32 (rows, 2c) ->
 * split planes (rows, c). */
int pk_gated_activation(const float* h, int64_t rows, int32_t c, void* z_hi, void* z_lo, pk_stream_t stream);
/* res / skip update of ResidualBlock.add_input (:283-286) + ResidualNet.add_input (:385-392): o fp32 (rows, 2c);
 * state += o[:, :c]; skip = skip_init ? o[:, c:] : skip + o[:, c:]; optional split copy of the new state into the next
 * layer's ring buffer (columns [col0, col0 + c) of a (rows, ld) plane). */
int pk_waveflow_layer_update(const float* o, int64_t rows, int32_t c, float* state, float* skip, int32_t skip_init, void* buf_hi,
                             void* buf_lo, int32_t ld, int32_t col0, pk_stream_t stream);
/* Flow._predict_row_parameters tail + _inverse_transform_row (:496-510): (logs, b) = output_proj(skip) (C -> 2, w [2][C]);
 * x_next[b,w] = (z_row[b,w] - b) * exp(-logs). */
int pk_waveflow_row_out(const float* skip, const float* w, const float* bias, const float* z_row, int64_t z_batch_stride,
                        int32_t batch, int32_t width, int32_t c, float* x_next, int64_t x_batch_stride, pk_stream_t stream);

/* One whole ResidualBlock.add_input (:248-285) as ONE CTA-pair kernel (channels == 64; the default layer path of
 * ConditionalWaveFlow.inverse, PK_WF_FUSED=0 selects the pk_conv_gemm_ex pair above):
 *   a | g = conv2d(3-row ring, dilation (1, 2^l)) + condition_proj(condition row) + bias1;  z = tanh(a) sigmoid(g);
 *   skip | res = out_proj(z) + bias2;  skip accumulator (=|+=) skip;  ring slot `slot` of the NEXT layer <- row + res.
 * buf planes (batch, width, 192) hold this layer's ring (slot s in columns [64 s, 64 s + 64)); `slot` is the newest row's slot
 * (the residual input).  cond planes: the condition row (batch, width, n_mels), batch stride cond_batch_stride elements.
 * Every pair of planes comes from ONE allocation (lo after hi).
 * w1 planes (128, 704): row n = gate channel (a: 0..63, g: 64..127); columns [192 tap + 64 s + c] = conv.weight[n, c, kh(s), tap]
 * for the ring slot s holding kernel row kh(s) at this row step, columns [576 + m] = condition_proj.weight[n, m] (zeros from
 * n_mels to 128).  w2 planes (128, 64): out_proj.weight with rows reordered to skip (0..63) | res (64..127).
 * bias1 / bias2: HOST pointers [128] (conv.bias + condition_proj.bias; out_proj.bias as skip | res), copied into the kernel
 * parameter block.  next_hi / next_lo: the next layer's ring planes, or NULL (last layer).  skip fp32 (batch, width, 64). */
typedef struct pk_waveflow_layer_args {
  int32_t batch, width, channels, n_mels, dilation, slot;
  const void* buf_hi;
  const void* buf_lo;
  const void* cond_hi;
  const void* cond_lo;
  int64_t cond_batch_stride;
  const void* w1_hi;
  const void* w1_lo;
  const void* w2_hi;
  const void* w2_lo;
  const float* bias1;
  const float* bias2;
  void* next_hi;
  void* next_lo;
  float* skip;
  int32_t skip_init;
  void* prof;              /* debug: NULL, or device uint64[8]: MMA-issuer cycles [0] issue, [1] wait data, [2] wait acc2, [3] wait z; [4] tiles */
} pk_waveflow_layer_args;
int pk_waveflow_layer(const pk_waveflow_layer_args* args, pk_stream_t stream);

/* All row steps i = 1 .. n_group-1 of one Flow.inverse (:515-556) in ONE persistent dataflow launch (the default path of
 * ConditionalWaveFlow.inverse for 64 and 128 channels; PK_WF_FUSED=layer selects one pk_waveflow_layer launch per layer and row).
 * Per row step: the n_layers ResidualBlock.add_input of pk_waveflow_layer, then on the completed skip sum
 * (logs, b) = output_proj(skip), x[:, i] = (z[:, i] - b) exp(-logs) (:496-510) and input_proj(x[:, i]) (:437-442) into the
 * first layer's ring slot i mod 3.  The caller provides row 0: x[:, 0] = z[:, 0], input_proj(x[:, 0]) in slot 0 of ring 0
 * (pk_waveflow_input_proj), all other ring contents zero, and `flags` (one uint32 per tile: (n_group - 1) * n_layers *
 * batch * ceil(width / 256)) zeroed.  z / x: fp32 (batch, n_group, width).  cond planes (batch, n_group, width, n_mels);
 * cond_rows[i] (HOST) = the condition row used at row step i.  ring / w1 / w2 / bias1 / bias2: HOST arrays indexed by layer
 * (w1: 3 * layer + variant, variant = i mod 3) of the per-layer pointers of pk_waveflow_layer_args (biases: HOST floats).
 * in_w / in_b [channels], out_w [2][channels], out_b [2]: HOST floats.
 * channels == 128 (examples/waveflow/config.py) runs the same dataflow with the channels as two blocks of 64: ring planes
 * (batch, width, 384) with slot s in columns [128 s, 128 s + 128); w1 planes (256, 1280): rows a0 | g0 | a1 | g1 (64 each: gate
 * channel a_k = conv output 64 k + c, g_k = 128 + 64 k + c), columns [128 (3 tap + s) + c] conv, [1152 + m] condition_proj;
 * w2 planes (256, 128): rows skip0 | res0 | skip1 | res1; bias1 / bias2 [256] in the same row orders; skip (batch, width, 128). */
typedef struct pk_waveflow_flow_args {
  int32_t batch, width, channels, n_mels, n_layers, n_group;
  const int32_t* cond_rows;
  void* const* ring_hi;
  void* const* ring_lo;
  const void* cond_hi;
  const void* cond_lo;
  const void* const* w1_hi;
  const void* const* w1_lo;
  const void* const* w2_hi;
  const void* const* w2_lo;
  const float* const* bias1;
  const float* const* bias2;
  const float* in_w;
  const float* in_b;
  const float* out_w;
  const float* out_b;
  const float* z;
  float* x;
  float* skip;
  uint32_t* flags;
  int64_t flags_len;
  void* prof;              /* debug: as in pk_waveflow_layer_args */
} pk_waveflow_flow_args;
int pk_waveflow_flow(const pk_waveflow_flow_args* args, pk_stream_t stream);

/* FastSpeech2Loss.forward with use_masking=True (models/fastspeech2/fastspeech2.py:701-812; DurationPredictorLoss
 * duration_predictor.py:140-184): out4 = { l1_loss = L1(before, ys) + L1(after, ys) over valid frames,
 * duration_loss = MSE(d_outs, log(ds + 1)), pitch_loss = MSE(p_outs, ps), energy_loss = MSE(e_outs, es) over valid tokens }.
 * before/after/ys (batch, l_max, odim) fp32; d_outs/p_outs/ps/e_outs/es (batch, t_max) fp32; ds int64; workspace12:
 * device fp32[12] scratch (zeroed by the call). */
int pk_fs2_loss(const float* before, const float* after, const float* ys, const int32_t* olens, int32_t l_max, int32_t odim,
                const float* d_outs, const int64_t* ds, const float* p_outs, const float* ps, const float* e_outs,
                const float* es, const int32_t* ilens, int32_t t_max, int32_t batch, float* workspace12, float* out4,
                pk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * FastSpeech2 training step (reference: FastSpeech2Updater.update_core, models/fastspeech2/fastspeech2_updater.py:51-99:
 * forward, FastSpeech2Loss, loss.backward(), optimizer.step(); DataParallel gradient averaging is one NCCL all-reduce
 * on the flat gradient buffer, issued by the host through torch.distributed).  GEMM-shaped gradients reuse
 * pk_conv_gemm: dgrad = conv with flipped taps, wgrad / attention gradients = NT matmuls on transposed split planes.
 * ------------------------------------------------------------------------------------------------------------ */
/* dst[z*dst_zstride + c*ld_dst + r] = src[z*src_zstride + (r + shift)*ld_src + c0 + c] for r in [0, r_out), c in [0, cols)
 * (0 where r + shift is outside [0, rows)); split planes.  Builds K-major operands (X^T, dY^T, K^T, dS^T ...). */
int pk_transpose_planes(const void* src_hi, const void* src_lo, int32_t z, int32_t rows, int64_t src_zstride, int32_t ld_src,
                        int32_t c0, int32_t cols, int32_t shift, int32_t r_out, void* dst_hi, void* dst_lo, int64_t dst_zstride,
                        int64_t ld_dst, pk_stream_t stream);
/* LayerNorm backward: dx (+)= d/dx, dgamma += sum dy*xhat, dbeta += sum dy (fp32 [d], accumulated atomically). */
int pk_layer_norm_bwd(const float* x, const float* gamma, const float* dy, float eps, int64_t rows, int32_t d, float* dx,
                      int32_t accumulate, float* dgamma, float* dbeta, pk_stream_t stream);
/* softmax backward: ds = scale * p * (dp - sum_k p dp) over the first `keys` columns of rows of pitch ld (rest -> 0). */
int pk_softmax_bwd(const void* p_hi, const void* p_lo, const float* dp, int64_t rows, int32_t keys, int32_t ld, float scale,
                   void* ds_hi, void* ds_lo, pk_stream_t stream);
/* out[c] += sum_rows x[row, c] (bias gradients). */
int pk_colsum(const float* x, int64_t rows, int32_t c, float* out, pk_stream_t stream);
/* pk_colsum on split planes (rows, ld): out[c] += sum_rows (x_hi + x_lo)[row, c] for c < cols (bias gradients from the split
 * copy of dY, which - unlike the fp32 dY of a residual stream - nobody updates in place afterwards). */
int pk_colsum_split(const void* x_hi, const void* x_lo, int64_t rows, int32_t cols, int32_t ld, float* out, pk_stream_t stream);
/* out[i] = sum_{s < slices} part[s * n + i] (fp32): the reduction of split-K partial products of a weight gradient; overwrites
 * `out` (unlike pk_colsum, which accumulates). */
int pk_sum_slices(const float* part, int32_t slices, int64_t n, float* out, pk_stream_t stream);
/* BatchNorm1D in training mode on (rows, c) (tacotron2/decoder.py:128-180 under model.train()): batch statistics
 * (biased variance), y = act(gamma * xhat + beta) with act in {PK_ACT_NONE, PK_ACT_TANH}, running statistics updated with
 * Paddle's momentum (running = momentum * running + (1 - momentum) * batch); saves mean / rstd for the backward. */
int pk_batch_norm_train(const float* x, int64_t rows, int32_t c, const float* gamma, const float* beta, float eps, int32_t act,
                        float momentum, float* run_mean, float* run_var, float* sums2c, float* y, void* y_hi, void* y_lo,
                        float* save_mean, float* save_rstd, pk_stream_t stream);
/* backward of the above (including the tanh): dx; afterwards sums2c = { dbeta[c], dgamma[c] }. */
int pk_batch_norm_bwd(const float* x, const float* dy, const float* y_act, const float* mean, const float* rstd, const float* gamma,
                      int32_t act, int64_t rows, int32_t c, float* sums2c, float* dx, pk_stream_t stream);
/* dx = dy * (y > 0), y given by the hi plane of the saved ReLU output; fp32 and/or split output. */
int pk_relu_bwd(const float* dy, const void* y_hi, int64_t n, float* dx, void* dx_hi, void* dx_lo, pk_stream_t stream);
/* y += a * x */
int pk_axpy(float a, const float* x, int64_t n, float* y, pk_stream_t stream);
/* gradients of l1_loss + duration_loss + pitch_loss + energy_loss (pk_fs2_loss) w.r.t. before, after, d_outs, p_outs, e_outs */
int pk_fs2_loss_bwd(const float* before, const float* after, const float* ys, const int32_t* olens, int32_t l_max, int32_t odim,
                    const float* d_outs, const int64_t* ds, const float* p_outs, const float* ps, const float* e_outs,
                    const float* es, const int32_t* ilens, int32_t t_max, int32_t batch, float* g_before, float* g_after,
                    float* g_d, float* g_p, float* g_e, pk_stream_t stream);
/* backward of pk_embed_pe: dtable[ids] += dx (ids may be NULL: positional encoding only), dalpha += sum dx * PE */
int pk_embed_pe_bwd(const int64_t* ids, const float* dx, int32_t vocab, int32_t padding_idx, int32_t batch, int32_t t, int32_t d,
                    float* dtable, float* dalpha, pk_stream_t stream);
/* backward of pk_length_regulate: dx[b, j, :] = sum over the frames of token j of dy[b, frame, :] */
int pk_length_regulate_bwd(const float* dy, const int64_t* dur, int32_t batch, int32_t t_in, int32_t c, int32_t t_out, float* dx,
                           pk_stream_t stream);
/* weight / bias gradients of pitch_embed / energy_embed (Conv1D(1 -> c, k) on a scalar track): dw [c][k], db [c] accumulated */
int pk_scalar_conv_wgrad(const float* dhs, const float* track, int32_t batch, int32_t t, int32_t c, int32_t k, float* dw, float* db,
                         pk_stream_t stream);
/* paddle.optimizer.Adam step on a flat buffer (training/optimizer.py:17-46): g * grad_scale (1/world for DataParallel),
 * lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), p -= lr_t * m / (sqrt(v) + eps * sqrt(1 - beta2^t)). */
/* nn.Dropout in training mode (upscale_in_train): y[i] = keep(i) ? x[i] / (1 - p) : 0, keep(i) = word (i & 3) of
 * Philox4x32-10(counter {i >> 2 low, high, site, step}, key {seed low, high}) >= p * 2^32.  Input fp32 x or split planes
 * (x_hi + x_lo), outputs fp32 and/or split planes; in place allowed.  The backward pass applies the same call to the gradient
 * (same seed / site / step regenerate the mask).  Reference: the Dropout layers of parakeet/modules/fastspeech2_transformer/
 * {embedding.py:79,126, attention.py:124, encoder_layer.py:101,108, multi_layer_conv.py:76}, fastspeech2_predictor/
 * {duration_predictor.py:82, variance_predictor.py:73}, tacotron2/decoder.py:144-180 (Postnet). */
int pk_dropout(const float* x, const void* x_hi, const void* x_lo, int64_t n, float p, uint64_t seed, uint32_t site, uint32_t step,
               const uint32_t* step_dev /* device uint32 added to `step` (NULL: none) - lets a CUDA graph replay draw fresh masks */,
               float* y, void* y_hi, void* y_lo, pk_stream_t stream);
int pk_adam(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
            int32_t step, float grad_scale, pk_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Parallel WaveGAN training step (reference: PWGUpdater.update_core, models/parallel_wavegan/parallel_wavegan_updater.py:76-153;
 * generator :445-472, PWGDiscriminator :554-614, MultiResolutionSTFTLoss modules/stft_loss.py:163-219).  Every Conv1D forward /
 * data gradient / weight gradient and the DFTs run through pk_conv_gemm; these are the element-wise pieces in between.
 * Shapes are channels-last (rows = batch * t) fp32 unless noted.
 * ------------------------------------------------------------------------------------------------------------ */
/* z = tanh(h[:, :c]) * sigmoid(h[:, c:])  (ResidualBlock :307-310); h (rows, 2c); outputs fp32 and/or split planes (rows, c). */
int pk_gate_fwd(const float* h, int64_t rows, int32_t c, float* z, void* z_hi, void* z_lo, pk_stream_t stream);
int pk_gate_bwd(const float* h, const float* dz, int64_t rows, int32_t c, float* dh, pk_stream_t stream);
/* nn.LeakyReLU(negative_slope) forward (fp32 and/or split planes out) and backward (x = the pre-activation). */
int pk_leaky_relu(const float* x, int64_t n, float slope, float* y, void* y_hi, void* y_lo, pk_stream_t stream);
int pk_leaky_relu_bwd(const float* x, const float* dy, int64_t n, float slope, float* dx, pk_stream_t stream);
/* nn.utils.weight_norm (dim 0): w[r, :] = g[r] * v[r, :] / ||v[r, :]||; backward: dg, dv from dw.  v (rows, inner). */
int pk_weight_norm_fwd(const float* v, const float* g, int32_t rows, int32_t inner, float* w, float* norm, pk_stream_t stream);
int pk_weight_norm_bwd(const float* v, const float* g, const float* dw, int32_t rows, int32_t inner, float* dg, float* dv, pk_stream_t stream);
/* MSELoss against a constant over x[i * ld + col], i < n: acc[0] += sum (x - target)^2 (device double); dx (or NULL) = coef * (x - target). */
int pk_mse_const(const float* x, int64_t n, int32_t ld, int32_t col, float target, double* acc, float* dx, float coef, pk_stream_t stream);
/* acc[0] += sum x^2 (device double): the global gradient norm of ClipGradByGlobalNorm. */
int pk_sq_sum(const float* x, int64_t n, double* acc, pk_stream_t stream);
/* pk_adam with ClipGradByGlobalNorm folded in: g <- g * clip / max(sqrt(*sqnorm), clip) (sqnorm NULL or clip <= 0: no clipping). */
int pk_adam_clip(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 int32_t step, const double* sqnorm, float clip_norm, pk_stream_t stream);
/* generator residual / skip update (:311-315, :466-468): so (rows, 128) = [skip | out]; skips (=|+=) skip; xo = (out + x) * sqrt(1/2)
 * as fp32 and split planes; and its backward: dso = [dskips | dxo * sqrt(1/2)], dx_res = dxo * sqrt(1/2). */
int pk_pwg_res_update(const float* so, const float* x, int64_t rows, float* skips, int32_t init, float* xo, void* xo_hi, void* xo_lo,
                      pk_stream_t stream);
int pk_pwg_res_update_bwd(const float* dskips, const float* dxo, int64_t rows, float* dso, float* dx_res, pk_stream_t stream);
/* one upsampling stage (Stretch2D nearest x s + Conv2D FIR of 2s+1 taps, zero pad s; :48-63,119-138) on x (rows, tin) -> (rows, tin*s),
 * and its backward: dx (rows, tin) and/or dfir[2s+1] (device double, accumulated). */
int pk_up_stage_fwd(const float* x, const float* fir, int64_t rows, int32_t tin, int32_t s, float* y, pk_stream_t stream);
int pk_up_stage_bwd(const float* x, const float* dy, const float* fir, int64_t rows, int32_t tin, int32_t s, float* dx, double* dfir,
                    pk_stream_t stream);
/* gradient of weight * (spectral convergence + log STFT magnitude) of ONE resolution (stft_loss.py:20-161) w.r.t. re / im of the
 * generated signal's STFT: x / y re, im (batch, bins, frames) from pk_stft; sums = the device fp32[3] of pk_spectral_loss_sums;
 * g (batch * frames, 2 * bins_p) row-major [re | im] (padding columns are not written: zero them once). */
int pk_stft_loss_grad(const float* xre, const float* xim, const float* yre, const float* yim, int32_t batch, int32_t bins, int32_t frames,
                      int32_t bins_p, const float* sums, float weight, float* g, pk_stream_t stream);
/* adjoint of framing (centre, reflect padding, window): dx[b, reflect(f * hop + k - n_fft / 2)] += frames_grad[b, f, k] * window[k]. */
int pk_frames_overlap_add(const float* frames_grad, const float* window, int32_t batch, int32_t frames, int32_t n_fft, int32_t hop,
                          int32_t t, float* dx, pk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PARAKEET_B200_H_ */
