/*
 * csm_b200.h -- C ABI of libcsm_b200.so: the CSM-1B frame-generation hot path of
 * zenoran/sesameai-tts as hand-written sm_100a CUDA kernels.
 *
 * Boundary rules:
 *   - plain C, POD arguments only; every pointer marked "dev" is a CUDA device pointer
 *     (torch ``tensor.data_ptr()``), ``stream`` is a ``cudaStream_t`` passed as void*;
 *   - the caller (PyTorch) owns every buffer: weights, workspace, inputs, outputs.  The
 *     library owns only CUDA graphs / descriptors inside ``csm_ctx``;
 *   - every call is stream ordered and never synchronises the device; a ctx is not thread-safe
 *     (the reference's KV caches are mutable module state too);
 *   - return value 0 = success, negative = error (``csm_last_error`` gives the text).  Nothing
 *     throws or exits across the ABI.  There is NO CPU fallback: without a CUDA device every
 *     compute call returns CSM_ERR_CUDA.
 *
 * Each entry point cites the reference interface (file:line under /root/reference) it replaces.
 */
#ifndef CSM_B200_H
#define CSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSM_B200_ABI_VERSION 1

enum {
  CSM_OK = 0,
  CSM_ERR_ARG = -1,       /* bad argument / unsupported shape                     */
  CSM_ERR_CUDA = -2,      /* CUDA runtime error (see csm_last_error)              */
  CSM_ERR_STATE = -3,     /* e.g. caches not enabled -> Python AssertionError     */
  CSM_ERR_OVERFLOW = -4,  /* KV cache would exceed max_seq_len                    */
  CSM_ERR_WORKSPACE = -5  /* workspace too small                                  */
};

/* One llama3_2 stack (sesameai/models.py:10-39). head_dim = dim / heads, must be 64 or 128. */
typedef struct csm_stack_config {
  int32_t layers, dim, heads, kv_heads, ff;
} csm_stack_config;

/* ModelArgs + the constants hard-wired in sesameai/models.py:10-39,90-96,127. */
typedef struct csm_config {
  csm_stack_config backbone;   /* llama-1B:   16, 2048, 32, 8, 8192 */
  csm_stack_config decoder;    /* llama-100M:  4, 1024,  8, 2, 8192 */
  int32_t text_vocab;          /* 128256 */
  int32_t audio_vocab;         /* 2051   */
  int32_t codebooks;           /* 32     */
  int32_t max_seq_len;         /* backbone KV slots, 2048 (models.py:17)                      */
  float norm_eps;              /* 1e-5 */
} csm_config;

/* bf16 device pointers to one layer's parameters, torch layout ([out,in] row-major). */
typedef struct csm_layer_weights {
  const void *q_proj, *k_proj, *v_proj, *output_proj; /* attn.{q,k,v,output}_proj.weight */
  const void *w1, *w2, *w3;                           /* mlp.w1 (gate), w2 (down), w3 (up) */
  const void *sa_norm, *mlp_norm;                     /* {sa,mlp}_norm.scale               */
} csm_layer_weights;

/* All parameters of ``Model`` (sesameai/models.py:110-118), bf16, on the device. */
typedef struct csm_weights {
  const void *text_embeddings;   /* [text_vocab, D]                                   */
  const void *audio_embeddings;  /* [audio_vocab*codebooks, D]                        */
  const void *projection;        /* [Dd, D]                                           */
  const void *codebook0_head;    /* [audio_vocab, D]                                  */
  const void *audio_head;        /* [codebooks-1, Dd, audio_vocab]  ([in,out] per slice, models.py:118,176) */
  const void *backbone_norm, *decoder_norm;       /* norm.scale                       */
  const void *backbone_rope, *decoder_rope;       /* Llama3ScaledRoPE cache [max_pos, hd/2, 2] (cos,sin), bf16 */
  int32_t backbone_rope_len, decoder_rope_len;    /* rows in the tables               */
  const csm_layer_weights *backbone_layers;       /* host array[backbone.layers]      */
  const csm_layer_weights *decoder_layers;        /* host array[decoder.layers]       */
} csm_weights;

typedef struct csm_ctx csm_ctx;

/* ---- life cycle --------------------------------------------------------------------------- */

int32_t csm_abi_version(void);
/* Kernels this library has launched in this process so far (a CUDA-graph replay counts its
 * kernel nodes); bench.py reports the difference over the timed region as ``gpu_launches``. */
uint64_t csm_launch_count(void);
/* Test hook: launch the frame path's kernels with (1, default; env CSM_PDL=0 disables) or without (0) the programmatic
 * dependent launch attribute.  Applies to launches and graph captures made after the call. */
void csm_debug_set_pdl(int32_t on);
const char *csm_last_error(void);

/* Bytes of device workspace csm_create needs for ``max_batch`` streams: KV caches (GQA-compact
 * [L][B][KV][slots][hd] bf16), packed weight copies, activation scratch.  Returns 0 on bad config. */
size_t csm_workspace_bytes(const csm_config *cfg, int32_t max_batch);

/* Replaces Model.setup_caches(max_batch_size) (sesameai/models.py:120-130; torchtune
 * TransformerDecoder.setup_caches).  ``workspace`` is a dev buffer of >= csm_workspace_bytes,
 * 256-byte aligned; weights are re-packed from ``w`` on ``stream`` (fused QKV, interleaved
 * gate/up, transposed audio heads).  No causal-mask tensors are built: masks are implicit. */
int32_t csm_create(const csm_config *cfg, const csm_weights *w, int32_t max_batch, void *workspace,
                   size_t workspace_bytes, void *stream, csm_ctx **out);
void csm_destroy(csm_ctx *ctx);

/* Replaces Model.reset_caches() (sesameai/models.py:186-188): rewinds the backbone and decoder
 * cache positions.  No memset is needed because attention never reads beyond the valid length. */
int32_t csm_reset_caches(csm_ctx *ctx);

/* Backbone positions currently held in the KV cache (torchtune KVCache.size); lane 0 of a lock-step batch. */
int32_t csm_cache_len(const csm_ctx *ctx);

/* Continuous batching (the reference loop sesameai/generator.py:283-294 serves one
 * stream and stops at its EOS frame, models.py:160 assumes a lock-step batch): a cache lane is one stream's
 * slice of the KV cache.  csm_lane_reset rewinds one lane (a finished stream leaves, a new one may join),
 * csm_lane_len is the number of positions it holds. */
int32_t csm_lane_reset(csm_ctx *ctx, int32_t lane);
int32_t csm_lane_len(const csm_ctx *ctx, int32_t lane);

/* Device-side errors of earlier stream-ordered calls, WITHOUT a CUDA call (the kernels mirror the code into
 * mapped host memory): 0 = none; 0x100-0x4ff = a wait inside the decode megakernel gave up (the launch
 * drained, its tokens are garbage, the context is still usable: reset and retry -- tts_service.py:500-514
 * retries on any exception); 0x801 = token id outside its embedding table (the reference raises IndexError,
 * sesameai/models.py:190-203); 0x802 = teacher-forced id out of range; 0x803 = input_pos outside the RoPE
 * table or different from the cache position (sesameai/models.py:154,158: mask row and cache slot coincide
 * only for sequential use).  Meaningful after the stream has been synchronised; ``clear`` resets it. */
int32_t csm_check_error(csm_ctx *ctx, int32_t clear);

/* ---- the hot path -------------------------------------------------------------------------- */

/* Optional per-call extras; zero-initialise for the plain reference behaviour. */
typedef struct csm_frame_opts {
  const void *noise;      /* dev bf16 [codebooks, B, audio_vocab] Exp(1) draws q, or NULL -> in-kernel RNG */
  uint64_t seed;          /* RNG seed when noise == NULL                                                   */
  uint64_t offset;        /* RNG stream offset (frame counter) when noise == NULL                          */
  const int32_t *forced;  /* dev [B, codebooks] teacher-forced tokens, or NULL                             */
  void *logits_out;       /* dev bf16 [codebooks, B, audio_vocab] raw head outputs, or NULL                */
  int32_t *sampled_out;   /* dev [B, codebooks] sampled tokens before forcing, or NULL                     */
  int32_t path;           /* CSM_PATH_*: which launch strategy runs the last prompt row + frame tail          */
  int32_t prefill;        /* CSM_PREFILL_*: how prompt rows [0, S-1) are processed                            */
  const int32_t *lanes;   /* HOST int32 [B] (read before the call returns): KV-cache lane of each batch row, all
                           * distinct, each < max_batch; NULL = row b on lane b (the reference's lock-step batch).
                           * Every lane keeps its own length, so streams may join (prefill a free lane with a
                           * B = 1 call), advance together in any subset, and leave (csm_lane_reset) per frame.   */
} csm_frame_opts;

enum {
  CSM_PREFILL_AUTO = 0,      /* tensor cores when B*(S-1) >= 64 rows, else the small-row kernels */
  CSM_PREFILL_SMALL_ROW = 1, /* GEMV-style kernels, 8 frames per stream per pass                 */
  CSM_PREFILL_TENSOR = 2     /* TMA + tcgen05 GEMMs over up to 4096 rows per pass                */
};

enum {
  CSM_PATH_AUTO = 0,   /* batch 1: persistent megakernel; otherwise the captured per-op CUDA graph */
  CSM_PATH_DIRECT = 1, /* per-op kernels launched one by one (debugging / graph-free)              */
  CSM_PATH_GRAPH = 2,  /* per-op kernels replayed from one CUDA graph per frame                    */
  CSM_PATH_MEGA = 3    /* one persistent kernel per frame (batch 1 only)                           */
};

/* Replaces Model.generate_frame(tokens, tokens_mask, input_pos, temperature, topk)
 * (sesameai/models.py:132-184): embeds S frames per stream, appends S positions to the backbone
 * KV cache, samples codebook 0 from the last position and runs the 31-step depth decoder.
 *   tokens      dev int64 [B, S, codebooks+1]
 *   tokens_mask dev uint8 (torch.bool) [B, S, codebooks+1]
 *   input_pos   dev int64 [B, S]
 *   out         dev int32 [B, codebooks]
 * CSM_ERR_STATE if B exceeds max_batch; CSM_ERR_OVERFLOW if the cache would pass max_seq_len
 * (torchtune KVCache.update's assert). */
int32_t csm_generate_frame(csm_ctx *ctx, const int64_t *tokens, const uint8_t *tokens_mask,
                           const int64_t *input_pos, int32_t B, int32_t S, float temperature, int32_t topk,
                           const csm_frame_opts *opts, int32_t *out, void *stream);

/* ---- Mimi codec decode --------------------------------------------------------------------- */

/* Replaces moshi's ``MimiModel.decode`` as the reference calls it through
 * ``Generator._audio_tokenizer.decode`` (sesameai/generator.py:116,299; tts_service.py:245):
 * codes int64 [B, K<=32, T] -> waveform fp32 [B, 1, 1920*T] at 24 kHz.  All weights are fp32 device
 * pointers in moshi's own tensor layouts, passed as one flat array indexed by MIMI_W_*:
 *   CODEBOOK0 + 2k     quantizer.rvq_{first|rest}...layers[..]._codebook.embedding_sum [2048,256]
 *   CODEBOOK0 + 2k + 1 ... .cluster_usage [2048]            (k = 0 semantic, 1..31 acoustic)
 *   RVQ_FIRST_PROJ / RVQ_REST_PROJ   output_proj.weight [512,256,1]
 *   UPSAMPLE           upsample ConvTranspose1d weight [512,1,4] (depthwise, stride 2)
 *   LAYER0 + 10*l + {0 in_proj_weight [1536,512], 1 out_proj.weight, 2 norm1.weight, 3 norm1.bias,
 *                    4 norm2.weight, 5 norm2.bias, 6 linear1.weight [2048,512], 7 linear2.weight
 *                    [512,2048], 8 layer_scale_1.scale, 9 layer_scale_2.scale}
 *   CONV0 (+1 bias)    SEANet decoder model.0 Conv1d [1024,512,7]
 *   STAGE0 + 6*s + {0 convtr.weight [C, C/2, 2r], 1 convtr.bias, 2 block.1 conv weight [C/4, C/2, 3],
 *                   3 its bias, 4 block.3 conv weight [C/2, C/4, 1], 5 its bias}, r = 8,6,5,4
 *   FINAL (+1 bias)    last Conv1d [1,64,3] */
enum {
  MIMI_W_CODEBOOK0 = 0,
  MIMI_W_RVQ_FIRST_PROJ = 64,
  MIMI_W_RVQ_REST_PROJ = 65,
  MIMI_W_UPSAMPLE = 66,
  MIMI_W_LAYER0 = 67,
  MIMI_W_CONV0 = 147,
  MIMI_W_STAGE0 = 149,
  MIMI_W_FINAL = 173,
  /* encode side (moshi ``MimiModel.encode``, reference sesameai/generator.py:86):
   *   ENC_CONV0 (+1 bias)  encoder.model.0 Conv1d [64,1,7]
   *   ENC_STAGE0 + 6*s + {0 block.1 conv weight [C/2,C,3], 1 bias, 2 block.3 conv weight [C,C/2,1], 3 bias,
   *                       4 strided conv weight [2C, C, 2r], 5 bias}, C = 64,128,256,512, r = 4,5,6,8
   *   ENC_FINAL (+1 bias)  encoder.model.14 Conv1d [512,1024,3]
   *   ENC_LAYER0 + 10*l    encoder_transformer layers, same 10 tensors as LAYER0
   *   DOWNSAMPLE           downsample Conv1d weight [512,512,4] (stride 2, replicate padding, no bias)
   *   RVQ_FIRST_INPROJ / RVQ_REST_INPROJ   input_proj.weight [256,512,1] */
  MIMI_W_ENC_CONV0 = 175,
  MIMI_W_ENC_STAGE0 = 177,
  MIMI_W_ENC_FINAL = 201,
  MIMI_W_ENC_LAYER0 = 203,
  MIMI_W_DOWNSAMPLE = 283,
  MIMI_W_RVQ_FIRST_INPROJ = 284,
  MIMI_W_RVQ_REST_INPROJ = 285,
  MIMI_W_COUNT = 286
};
typedef struct mimi_ctx mimi_ctx;
size_t mimi_workspace_bytes(int32_t max_frames);
/* Packs the weights (embedding = embedding_sum / clamp(cluster_usage), tap-major conv matrices)
 * into ``workspace`` on ``stream`` and synchronises that stream once. */
int32_t mimi_create(const void *const *weights, int32_t n_weights, int32_t max_frames, void *workspace,
                    size_t workspace_bytes, void *stream, mimi_ctx **out);
/* T may exceed max_frames: the decode then runs in windows of max_frames frames that carry their causal left
 * context (conv tails, transformer K/V of the last 249 positions), bit-identical to one pass. */
int32_t mimi_decode(mimi_ctx *ctx, const int64_t *codes, int32_t B, int32_t K, int32_t T, float *out, void *stream);

/* Stateful streaming decode.  The reference's generate_stream decodes every 10-frame
 * buffer statelessly (sesameai/generator.py:61,111-117,189-196), so each chunk starts from silence; a
 * mimi_stream carries the causal left context from chunk to chunk, and the concatenated chunks equal the
 * one-shot decode of the whole utterance bit for bit.  ``state`` is a caller-owned device buffer of
 * mimi_stream_state_bytes() bytes (256-byte aligned); one stream = one utterance at a time. */
typedef struct mimi_stream mimi_stream;
size_t mimi_stream_state_bytes(void);
int32_t mimi_stream_create(mimi_ctx *ctx, void *state, size_t state_bytes, void *stream, mimi_stream **out);
int32_t mimi_stream_reset(mimi_stream *s, void *stream);          /* start of a new utterance */
/* codes int64 [K, T] (the next T frames of the utterance) -> out fp32 [1920*T] */
int32_t mimi_decode_stream(mimi_stream *s, const int64_t *codes, int32_t K, int32_t T, float *out, void *stream);
void mimi_stream_destroy(mimi_stream *s);
/* Replaces moshi's ``MimiModel.encode`` (reference sesameai/generator.py:86: voice-prompt audio ->
 * codes): wav fp32 [B, L] at 24 kHz (zero-padded on the right to whole 1920-sample frames) ->
 * codes int64 [B, K, ceil(L/1920)], K <= 32 codebooks (nearest centroid per residual layer). */
int32_t mimi_encode(mimi_ctx *ctx, const float *wav, int32_t B, int64_t L, int32_t K, int64_t *codes, void *stream);
/* Test entry: the split-RVQ nearest-centroid search of mimi_encode alone, on a given 12.5 Hz latent
 * (fp32 [T, 512] time-major, T <= max_frames) -> codes int64 [K, T]. */
int32_t mimi_k_rvq_encode(mimi_ctx *ctx, const float *latent, int32_t T, int32_t K, int64_t *codes, void *stream);
void mimi_destroy(mimi_ctx *ctx);

/* ---- waveform post-processing ------------------------------------------ */

/* torchaudio.functional.resample(x, orig_freq, new_freq) with its default arguments (sinc_interp_hann,
 * lowpass_filter_width 6, rolloff 0.99), as the reference calls it around the watermarker
 * (sesameai/watermarking.py:35-39: 24 kHz -> 44.1 kHz -> 24 kHz; tts_service.py:254-256): x dev fp32 [n] ->
 * y dev fp32 [csm_post_resample_len(n, ..)] = ceil(new * n / orig) samples.  ``workspace``: dev buffer of
 * csm_post_resample_workspace_bytes (the polyphase kernel table, rebuilt per call in fp64). */
int64_t csm_post_resample_len(int64_t n, int32_t orig_freq, int32_t new_freq);
size_t csm_post_resample_workspace_bytes(int32_t orig_freq, int32_t new_freq);
int32_t csm_post_resample(const float *x, int64_t n, int32_t orig_freq, int32_t new_freq, float *y, void *workspace,
                          size_t workspace_bytes, void *stream);

/* tts_service.generate_audio_segment (tts_service.py:287-306) on the device: peak-normalise
 * (audio / max(|audio|.max(), 1e-6)), convert to 16-bit PCM (x * 32767, truncated), add start / end silence
 * and pydub's fade_in / fade_out (per-sample linear gain from -120 dB, floor like audioop.mul; all counts in
 * SAMPLES): audio dev fp32 [n] -> out dev int16 [start_silence + n + end_silence]; scratch4: 4 dev bytes. */
int32_t csm_post_pcm16_segment(const float *audio, int64_t n, int64_t start_silence, int64_t end_silence, int64_t fade_in,
                               int64_t fade_out, int16_t *out, void *scratch4, void *stream);

/* Profiling aid: dev uint64 [n_ctas][n_phases][16] buffer that thread 0 of every CTA of the decode
 * megakernel fills (4 %globaltimer stamps: phase start, inputs staged, partial sums done, end; 12
 * clock64 marks); NULL disables.  Returns the number of phases (0 if the megakernel is unavailable). */
int32_t csm_debug_set_trace(csm_ctx *ctx, void *dev_buffer);

/* Host-only (no GPU needed): the megakernel's phase table for a configuration, with pointers taken
 * relative to an imaginary workspace.  tests/ use it to check the invariants of the tagged hand-off
 * statically: every consumed vector names the phase that wrote it last, and no vector is rewritten
 * sooner than two phases after it was written. */
typedef struct csm_phase_info {
  int32_t type;            /* 0 GEMV, 1 embed, 2 backbone attention, 3 sample */
  int32_t epi;             /* GEMV: 0 plain, 1 + residual (in place), 2 SwiGLU, 3 RoPE + KV append */
  int32_t nb, K, rows, R, G, rot, ldx, ldo, split_row, attn_prologue, has_qkv_table;
  int32_t x_src[2], resid_src[2], q_src, logits_src;
  uint64_t t_x, t_out, t_out2, t_q, t_kv, t_logits, t_next; /* byte offsets of the tagged vectors (0: unused) */
  /* ordering of the depth decoder's plain KV-cache rows: before this phase every CTA releases (1) / acquires (2)
   * through done words tagged with phase done_src; pos_mode 0 = depth decoder (cache rows [0, pos0) were written
   * in earlier codebook steps of the same launch), 1 = backbone (rows of earlier launches) */
  int32_t kv_sync, done_src, pos0, pos_mode;
} csm_phase_info;
int32_t csm_debug_phase_table(const csm_config *cfg, int32_t n_ctas, int32_t with_qkv_table, csm_phase_info *out,
                              int32_t max_phases);

/* ---- single-kernel entry points for unit parity tests (tests/ only) ------------------------ */

/* sample_topk (sesameai/models.py:77-87) on logits bf16 [B, V] with noise bf16 [B, V] -> int32 [B]. */
int32_t csm_k_sample_topk(const void *logits, const void *noise, int32_t B, int32_t V, float temperature,
                          int32_t topk, int32_t *out, void *stream);

/* _embed_tokens + mask + sum (sesameai/models.py:155-157,193-203): -> bf16 [N, D]. */
int32_t csm_k_embed_frames(const int64_t *tokens, const uint8_t *mask, const void *text_emb, const void *audio_emb,
                           int32_t N, int32_t codebooks, int32_t audio_vocab, int32_t D, void *out, void *stream);

/* y[N, out] = bf16(x[N, in] @ W[out, in]^T), x/W/y bf16 (nn.Linear without bias). */
int32_t csm_k_linear(const void *x, const void *W, int32_t N, int32_t in, int32_t out, void *y, void *stream);

/* Tensor-core path of the same op (TMA + tcgen05.mma + TMEM accumulator), used for prompt prefill:
 * y[N, out] = bf16(x @ W^T); epi 1 adds ``resid`` [N, out] (bf16(y) + resid, like h + linear(...));
 * epi 2 treats output columns as interleaved (gate_i, up_i) pairs and writes
 * bf16(bf16(silu(gate)) * up) into y[N, out/2].  ``in`` must be a multiple of 64. */
int32_t csm_k_gemm_tc(const void *x, const void *W, int32_t N, int32_t in, int32_t out, void *y, int32_t epi,
                      const void *resid, void *stream);

/* The same GEMM with split-K enabled when the shape calls for it (N <= 512 rows and few output tiles: the decode
 * steps of large batches): ``part`` = dev fp32 scratch of 148 * 128 * 128 floats, ``counters`` = 148 dev uint32,
 * zero on entry (left zero on exit). */
int32_t csm_k_gemm_tc_splitk(const void *x, const void *W, int32_t N, int32_t in, int32_t out, void *y, int32_t epi,
                             const void *resid, void *part, void *counters, void *stream);

/* Prompt attention (head_dim 64) on the tensor cores, the kernel behind the tensor-core prefill
 * (replaces the per-row scaled_dot_product_attention of torchtune's MultiHeadAttention for prompt rows):
 * q [B*chunk, heads*64] (row n = b*chunk + t), caches [B, kv_heads, slots, 64], row_slot[n] = cache slot of row
 * n (ascending in t); row n attends to slots [0 .. row_slot[n]] of stream b; out [B*chunk, heads*64], bf16. */
int32_t csm_k_attn_prefill(const void *q, const void *k_cache, const void *v_cache, const int32_t *row_slot, int32_t B,
                           int32_t chunk, int32_t heads, int32_t kv_heads, int32_t slots, void *out, void *stream);

/* torchtune RMSNorm: bf16(bf16(x * rsqrt(mean x^2 + eps)) * scale), [N, D]. */
int32_t csm_k_rmsnorm(const void *x, const void *scale, int32_t N, int32_t D, float eps, void *y, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CSM_B200_H */
