// C ABI of libcsm_b200.so (include/csm_b200.h): context, workspace carving, weight packing,
// the generate_frame launch sequence and its CUDA-graph capture.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <new>
#include <vector>

#include "../../include/csm_b200.h"
#include "nvtx_ranges.h"
#include "lm_kernels.cuh"
#include "mega.cuh"
#include "gemm_tc.cuh"
#include "skinny.cuh"

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}
#define CU_TRY(expr)                                                                       \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) return set_err(CSM_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// kernels launched by this library in this process (graph replays add their node count)
static std::atomic<unsigned long long> g_launches{0};
#define COUNT_LAUNCH() (g_launches.fetch_add(1, std::memory_order_relaxed))
extern "C" uint64_t csm_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// Kernel launch of the frame path: with the programmatic-serialisation attribute (common.cuh: pdl_trigger /
// pdl_wait) the kernel may start while its predecessor on the stream drains; every kernel launched here waits
// for that predecessor before it touches activations.  Captured into the decode graphs as programmatic edges.
// CSM_PDL=0 launches plainly (measurement aid).
static std::atomic<int> g_pdl{-1};
static bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    v = !(getenv("CSM_PDL") && getenv("CSM_PDL")[0] == '0');
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
// test hook: contexts created afterwards capture their graphs with (1) / without (0) programmatic edges
extern "C" void csm_debug_set_pdl(int32_t on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }
// cluster_z > 1: the z extent of the grid is launched as one thread-block cluster (split-K through DSMEM)
template <typename... P, typename... A>
static void launch_kcd(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, dim3 cluster, A&&... args);
template <typename... P, typename... A>
static void launch_kc(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_z, A&&... args) {
  launch_kcd(kern, grid, block, smem, st, dim3(1, 1, cluster_z > 1 ? cluster_z : 1), static_cast<A&&>(args)...);
}
template <typename... P, typename... A>
static void launch_kcd(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, dim3 cluster, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster.x * cluster.y * cluster.z > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster.x; at[n].val.clusterDim.y = cluster.y; at[n].val.clusterDim.z = cluster.z;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  (void)cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);  // failures surface through cudaGetLastError()
}
template <typename... P, typename... A>
static void launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  launch_kc(kern, grid, block, smem, st, 1, static_cast<A&&>(args)...);
}

struct csm_ctx;
namespace tc { struct RopeKV; }
static int launch_gemm_tc(const bf16* X, long long ldx, int rows, int K, const bf16* W, int n_out, bf16* out, long long ldo,
                          int epi, const bf16* resid, cudaStream_t st, const csm_ctx* splitk = nullptr,
                          const tc::RopeKV* rk = nullptr, bool* rk_fused = nullptr);

// shared with mimi_api.cu
int csm_set_error(int code, const char* msg) { return set_err(code, "%s", msg); }
void csm_count_launches(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int32_t csm_abi_version(void) { return CSM_B200_ABI_VERSION; }
extern "C" const char* csm_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------
static const int PREFILL_CHUNK = 8;     // prompt frames per stream per small-row pass
static const int PREFILL_TC_ROWS = 4096;  // rows per tensor-core prefill pass
// prompt rows (B * (S-1)) from which the row-batched path is used (skinny GEMMs up to 32 rows, tcgen05 beyond): a
// 32-frame text prompt as four per-op passes of 8 rows cost 4.6 ms, and a serving loop with ragged admission pays that
// nearly every round (config 5 with 32 lanes: 9.4 -> 8.0 s).  Up to 8 rows the per-op pass is one launch chain anyway.
static const int PREFILL_TC_MIN = 9;
static const int DECODE_TC_MIN = 16;      // streams from which a decode step runs on the tcgen05 GEMM
static int skinny_max_rows() {  // rows up to which a linear layer runs on the skinny fragment-major GEMM (CSM_SKINNY_MAX_ROWS: experiments)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CSM_SKINNY_MAX_ROWS");
    v = e ? atoi(e) : 32;  // (33 .. 64 rows: the tcgen05 GEMM with cluster split-K is faster -- 64 streams 12.3 -> 9.3 ms per step)
    if (v < 0 || v > 64) v = 32;
  }
  return v;
}
#define SKINNY_MAX_ROWS (skinny_max_rows())
static const int TC_SPLIT_TILES = 148;    // split-K of the tcgen05 GEMM: splits x tiles never exceeds one CTA per SM
static const int TC_SPLIT_MAX_ROWS = 512; // ... and is only used for decode-sized row counts

struct StackDev {
  csm_stack_config c;
  int hd, slots;
  std::vector<bf16*> wqkv, wgu;              // packed (workspace)
  std::vector<const bf16*> wo, wd, sa, mlp;  // caller's tensors
  const bf16* norm;
  const bf16* rope;
  int rope_len;
  bf16 *kc, *vc;  // [layers][streams][kv][slots][hd]
  size_t kv_layer_stride;
  // activations
  bf16 *h, *q, *att, *act;
  bf16 *xn, *qkv;  // tensor-core prefill only (backbone)
  // megakernel: fragment-major packed matrices and tagged activation words
  std::vector<bf16*> fqkv, fo, fgu, fd;
  uint32_t *t_h, *t_q, *t_kv, *t_att, *t_act;
  int rs_h, rs_q, rs_kv, rs_att, rs_act;  // words between the mega::REP copies of each tagged vector
};

struct csm_ctx {
  csm_config cfg;
  int max_batch, max_rows, Vp;
  StackDev bb, dec;
  const bf16 *text_emb, *audio_emb, *proj, *c0_head;
  bf16* head_t;   // [C-1][Vp][Dd]
  bf16* f_head0;  // fragment-major stacked [codebook0_head (Vf rows) ; projection (Dd rows)] x D
  bf16* f_heads;  // fragment-major audio heads [C-1][Vf][Dd]
  int Vf;         // audio vocab rounded up to 16 rows (fragment-major row groups)
  uint32_t* t_logits;  // tagged logits [Vf]
  uint32_t* t_done;    // tagged per-CTA "KV rows released" words (mega::kv_step_sync)
  char* tagged_base;   // all tagged words live in [tagged_base, tagged_base + tagged_bytes): zeroed at create
  size_t tagged_bytes;
  bf16* proj_table;  // projection(audio_embeddings[cb*V + tok]) for cb < C-1: [(C-1)*V][Dd]
  bf16* qkv_table;   // first decoder layer's RoPE'd [q;k;v] of proj_table rows (position cb + 1): [(C-1)*V][qkv cols]
  bool qkv_table_ok;
  bf16* dec_in;   // [2B][D]
  bf16* logits;   // [B][Vp]
  int *row_stream, *row_pos, *row_slot;
  FrameParams* d_params;
  int* d_lane_meta;            // [2 * max_batch]: row -> lane, row -> positions held (continuous batching)
  float* tc_part;              // split-K partial tiles of the tcgen05 GEMM (decode steps of large batches)
  unsigned int* tc_counters;   // one arrival counter per output tile, zero between launches
  std::vector<int> lane_len;   // positions held by every cache lane (host truth; calls are stream ordered)
  int cache_len;               // == lane_len[0]: the reference's single counter for a lock-step batch
  bool enabled;
  cudaStream_t cap_stream;
  // persistent decode megakernel (batch 1)
  mega::Phase* d_phases;
  mega::Sync* d_sync;       // device status word (frame counter + sticky error), always initialised
  unsigned int* mega_att_part;  // tagged partials of the megakernel's split long-context attention
  float* att_part;              // row-batched decode at long context: (o, m, l) of the key ranges (k_attn_split64_mma)
  bool long_ctx;                // this call's decode rows see >= ATT_LONG_MIN keys (set by csm_generate_frame)
  unsigned int* h_error;    // mapped host mirror of the error word (owned by the ctx)
  int n_phases, mega_grid;
  bool mega_ok;
  bool frag_ok;  // the fragment-major packed matrices exist (setup_mega): the skinny batched-decode GEMM can run
  unsigned long long* trace;  // optional device buffer [n_phases][8] (csm_debug_set_trace)
  mega::PfTable pf_table;     // weight-prefetch schedule, passed in kernel-parameter space
  unsigned mega_keep;  // depth-decoder matrices loaded with the L2 evict-last policy: 4 bits per layer (qkv, o, gate/up, down)
  int mega_Rbb[4], mega_Rdec[4], mega_Rh0, mega_Rh;  // row-group heights (mega_pick_R): [qkv, o, gate/up, down] per stack, stacked head, audio heads
  std::map<int, cudaGraphExec_t> graphs;  // keyed by B
  std::map<int, unsigned long long> graph_nodes;
};

static int backbone_pass_tc(csm_ctx* x, int B, int chunk, cudaStream_t st);
static int frame_tail_tc(csm_ctx* x, int B, cudaStream_t st);

struct Carver {
  char* base;
  size_t off;
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

static bool valid_stack(const csm_stack_config& s) {
  if (s.layers < 1 || s.heads < 1 || s.kv_heads < 1 || s.dim % s.heads) return false;
  const int hd = s.dim / s.heads;
  return (hd == 64 || hd == 128) && s.heads % s.kv_heads == 0 && s.dim % 256 == 0 && s.ff % 256 == 0 &&
         s.dim <= 8192 && s.ff <= 8192;
}
static bool valid_cfg(const csm_config* c) {
  return c && valid_stack(c->backbone) && valid_stack(c->decoder) && c->codebooks >= 2 && c->codebooks <= 64 &&
         c->audio_vocab >= 2 && c->audio_vocab <= SAMPLE_MAXV && c->text_vocab >= 1 && c->max_seq_len >= c->codebooks;
}

// (with the first-layer [q;k;v] table the steps after the first one have no QKV phase in layer 0)
static int mega_phase_count(const csm_config& c, bool qkv_table = false) {
  return 1 + c.backbone.layers * 5 + 2 + (c.codebooks - 1) * (c.decoder.layers * 4 + 2) - (qkv_table ? c.codebooks - 2 : 0);
}

static void carve_stack(Carver& cv, StackDev& s, const csm_stack_config& c, int slots, int streams, int rows) {
  s.c = c;
  s.hd = c.dim / c.heads;
  s.slots = slots;
  const size_t qkv_rows = (size_t)(c.heads + 2 * c.kv_heads) * s.hd;
  s.wqkv.resize(c.layers);
  s.wgu.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) {
    s.wqkv[l] = cv.take<bf16>(qkv_rows * c.dim);
    s.wgu[l] = cv.take<bf16>((size_t)2 * c.ff * c.dim);
  }
  s.kv_layer_stride = (size_t)streams * c.kv_heads * slots * s.hd;
  s.kc = cv.take<bf16>(s.kv_layer_stride * c.layers);
  s.vc = cv.take<bf16>(s.kv_layer_stride * c.layers);
  s.h = cv.take<bf16>((size_t)rows * c.dim);
  s.q = cv.take<bf16>((size_t)rows * c.dim);
  s.att = cv.take<bf16>((size_t)rows * c.dim);
  s.act = cv.take<bf16>((size_t)rows * c.ff);
  s.fqkv.resize(c.layers); s.fo.resize(c.layers); s.fgu.resize(c.layers); s.fd.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) {
    s.fqkv[l] = cv.take<bf16>(qkv_rows * c.dim);
    s.fo[l] = cv.take<bf16>((size_t)c.dim * c.dim);
    s.fgu[l] = cv.take<bf16>((size_t)2 * c.ff * c.dim);
    s.fd[l] = cv.take<bf16>((size_t)c.dim * c.ff);
  }
}
// REP copies of a tagged vector of n words, (n + pad) words apart
static uint32_t* take_rep(Carver& cv, size_t n, int* rs) {
  const size_t stride = ((n + 63) & ~(size_t)63) + 64;
  *rs = (int)stride;
  return cv.take<uint32_t>(stride * mega::REP);
}
static void carve_tagged(Carver& cv, StackDev& s) {
  const csm_stack_config& c = s.c;
  const size_t krows = (size_t)c.kv_heads * s.hd;
  s.t_h = take_rep(cv, (size_t)2 * c.dim, &s.rs_h);
  s.t_q = take_rep(cv, (size_t)2 * c.dim, &s.rs_q);
  s.t_kv = take_rep(cv, (size_t)2 * 2 * krows, &s.rs_kv);
  s.t_att = take_rep(cv, (size_t)2 * c.dim, &s.rs_att);
  s.t_act = take_rep(cv, (size_t)2 * c.ff, &s.rs_act);
}

static size_t carve_all(csm_ctx* x, char* base) {
  Carver cv{base, 0};
  const csm_config& c = x->cfg;
  x->max_rows = x->max_batch * PREFILL_CHUNK > PREFILL_TC_ROWS ? x->max_batch * PREFILL_CHUNK : PREFILL_TC_ROWS;
  x->Vp = (c.audio_vocab + 7) & ~7;
  carve_stack(cv, x->bb, c.backbone, c.max_seq_len, x->max_batch, x->max_rows);
  x->bb.xn = cv.take<bf16>((size_t)x->max_rows * c.backbone.dim);
  x->bb.qkv = cv.take<bf16>((size_t)x->max_rows * (c.backbone.heads + 2 * c.backbone.kv_heads) * (c.backbone.dim / c.backbone.heads));
  carve_stack(cv, x->dec, c.decoder, c.codebooks, x->max_batch, 2 * x->max_batch);
  x->dec.xn = cv.take<bf16>((size_t)2 * x->max_batch * c.decoder.dim);
  x->dec.qkv = cv.take<bf16>((size_t)2 * x->max_batch * (c.decoder.heads + 2 * c.decoder.kv_heads) * (c.decoder.dim / c.decoder.heads));
  x->head_t = cv.take<bf16>((size_t)(c.codebooks - 1) * x->Vp * c.decoder.dim);
  x->Vf = (c.audio_vocab + 15) & ~15;
  x->f_head0 = cv.take<bf16>((size_t)(x->Vf + c.decoder.dim) * c.backbone.dim);
  x->f_heads = cv.take<bf16>((size_t)(c.codebooks - 1) * x->Vf * c.decoder.dim);
  x->proj_table = cv.take<bf16>((size_t)(c.codebooks - 1) * c.audio_vocab * c.decoder.dim);
  x->qkv_table = cv.take<bf16>((size_t)(c.codebooks - 1) * c.audio_vocab * (c.decoder.heads + 2 * c.decoder.kv_heads) * (c.decoder.dim / c.decoder.heads));
  x->dec_in = cv.take<bf16>((size_t)2 * x->max_batch * c.backbone.dim);
  x->logits = cv.take<bf16>((size_t)x->max_batch * x->Vp);
  x->row_stream = cv.take<int>(x->max_rows);
  x->row_pos = cv.take<int>(x->max_rows);
  x->row_slot = cv.take<int>(x->max_rows);
  x->d_params = cv.take<FrameParams>(1);
  x->d_lane_meta = cv.take<int>((size_t)2 * x->max_batch);
  x->tc_part = cv.take<float>((size_t)TC_SPLIT_TILES * tc::BM * tc::BN);
  x->tc_counters = cv.take<unsigned int>(TC_SPLIT_TILES);
  x->d_sync = cv.take<mega::Sync>(1);
  x->mega_att_part = cv.take<unsigned int>((size_t)mega::ATTN_SPLITS * c.backbone.heads * mega::ATTN_PART_WORDS);
  x->att_part = cv.take<float>((size_t)x->max_batch * c.backbone.heads * AS_SPLITS * AS_PW);
  x->d_phases = cv.take<mega::Phase>(mega_phase_count(c));
  cv.off = (cv.off + 255) & ~(size_t)255;
  const size_t t0 = cv.off;
  carve_tagged(cv, x->bb);
  carve_tagged(cv, x->dec);
  x->t_logits = cv.take<uint32_t>((size_t)x->Vf);
  x->t_done = cv.take<uint32_t>(1024);  // one word per CTA (kv_step_sync)
  cv.off = (cv.off + 255) & ~(size_t)255;
  x->tagged_base = base ? base + t0 : nullptr;
  x->tagged_bytes = cv.off - t0;
  return (cv.off + 255) & ~(size_t)255;
}

extern "C" size_t csm_workspace_bytes(const csm_config* cfg, int32_t max_batch) {
  if (!valid_cfg(cfg) || max_batch < 1) return 0;
  csm_ctx tmp;
  tmp.cfg = *cfg;
  tmp.max_batch = max_batch;
  return carve_all(&tmp, nullptr);
}

// ---------------------------------------------------------------------------------------------
template <int NB, int EPI, bool NORM>
static cudaError_t gemv_attr() {
  return cudaFuncSetAttribute(k_gemv<NB, EPI, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192 * 2);
}
template <int EPI, bool NORM>
static cudaError_t gemv_attrs() {
  cudaError_t e;
  if ((e = gemv_attr<1, EPI, NORM>()) != cudaSuccess) return e;
  if ((e = gemv_attr<2, EPI, NORM>()) != cudaSuccess) return e;
  if ((e = gemv_attr<4, EPI, NORM>()) != cudaSuccess) return e;
  return gemv_attr<8, EPI, NORM>();
}
// Function attributes are per device: remember which devices have been opted in.
static bool device_done(std::atomic<unsigned long long>& mask, bool mark) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;  // unknown: set the attribute again
  const unsigned long long bit = 1ull << dev;
  if (mark) {
    mask.fetch_or(bit);
    return true;
  }
  return (mask.load() & bit) != 0;
}
// opt every GEMV instantiation into > 48 KB of dynamic shared memory (once per device)
static cudaError_t init_kernel_attrs() {
  static std::atomic<unsigned long long> done{0};
  if (device_done(done, false)) return cudaSuccess;
  cudaError_t e;
  if ((e = gemv_attrs<EPI_PLAIN, false>()) != cudaSuccess) return e;
  if ((e = gemv_attrs<EPI_PLAIN, true>()) != cudaSuccess) return e;
  if ((e = gemv_attrs<EPI_RESID, false>()) != cudaSuccess) return e;
  if ((e = gemv_attrs<EPI_SWIGLU, true>()) != cudaSuccess) return e;
  if ((e = gemv_attrs<EPI_ROPE_KV, true>()) != cudaSuccess) return e;
  device_done(done, true);
  return cudaSuccess;
}

template <int NB, int EPI, bool NORM>
static cudaError_t launch_gemv_t(const GemvArgs& a, cudaStream_t st) {
  const int threads = 256, wpc = threads / 32;
  const int npairs = (a.rows + 1) / 2;
  int gx = (npairs + wpc - 1) / wpc;
  if (gx > 148 * 8) gx = 148 * 8;
  dim3 grid(gx, (a.N + NB - 1) / NB);
  const size_t smem = (size_t)NB * a.K * sizeof(bf16);
  launch_k(k_gemv<NB, EPI, NORM>, dim3(grid), dim3(threads), smem, st, a); COUNT_LAUNCH();
  return cudaGetLastError();
}
template <int EPI, bool NORM>
static cudaError_t launch_gemv(const GemvArgs& a, cudaStream_t st) {
  if (a.N <= 1) return launch_gemv_t<1, EPI, NORM>(a, st);
  if (a.N <= 2) return launch_gemv_t<2, EPI, NORM>(a, st);
  if (a.N <= 4) return launch_gemv_t<4, EPI, NORM>(a, st);
  return launch_gemv_t<8, EPI, NORM>(a, st);
}

struct RowMeta {
  const int *stream, *pos, *slot;
  int imp_B, imp_pos;
  int chunk;  // prompt passes: rows n = b * chunk + t (0: not a prompt pass)
};

// One transformer layer on N rows (torchtune TransformerSelfAttentionLayer, Appendix A.3-A.4).
static cudaError_t run_layer(csm_ctx* x, StackDev& s, int l, int N, const RowMeta& m, cudaStream_t st) {
  const csm_stack_config& c = s.c;
  cudaError_t e;
  GemvArgs a;
  memset(&a, 0, sizeof(a));
  a.eps = x->cfg.norm_eps;
  a.N = N;
  a.row_stream = m.stream; a.row_pos = m.pos; a.row_slot = m.slot; a.imp_B = m.imp_B; a.imp_pos = m.imp_pos;
  a.heads = c.heads; a.kv_heads = c.kv_heads; a.hd = s.hd; a.slots = s.slots;
  bf16* kc = s.kc + s.kv_layer_stride * l;
  bf16* vc = s.vc + s.kv_layer_stride * l;
  // K2: sa_norm + [q;k;v] + RoPE + KV append
  a.W = s.wqkv[l]; a.rows = (c.heads + 2 * c.kv_heads) * s.hd; a.K = c.dim;
  a.x = s.h; a.ldx = c.dim; a.norm_scale = s.sa[l];
  a.q_out = s.q; a.k_cache = kc; a.v_cache = vc; a.rope = s.rope;
  if ((e = launch_gemv<EPI_ROPE_KV, true>(a, st)) != cudaSuccess) return e;
  // K3: attention
  {
    dim3 grid(N, c.heads);
    const size_t smem = (size_t)s.slots * sizeof(float);
    const float scale = 1.0f / sqrtf((float)s.hd);
    if (s.hd == 64) {
      launch_k(k_attn_rows<64>, dim3(grid), dim3(128), smem, st, s.q, kc, vc, m.stream, m.slot, m.imp_B, m.imp_pos, c.heads, c.kv_heads,
                                               s.slots, scale, s.att);
    } else {
      launch_k(k_attn_rows<128>, dim3(grid), dim3(128), smem, st, s.q, kc, vc, m.stream, m.slot, m.imp_B, m.imp_pos, c.heads, c.kv_heads,
                                                s.slots, scale, s.att);
    }
    COUNT_LAUNCH();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  // K4: output_proj + residual (in place on h)
  a.W = s.wo[l]; a.rows = c.dim; a.K = c.dim; a.x = s.att; a.ldx = c.dim;
  a.out = s.h; a.ldo = c.dim; a.resid = s.h; a.ldr = c.dim;
  if ((e = launch_gemv<EPI_RESID, false>(a, st)) != cudaSuccess) return e;
  // K5: mlp_norm + interleaved gate/up + SiLU*mul
  a.W = s.wgu[l]; a.rows = 2 * c.ff; a.K = c.dim; a.x = s.h; a.ldx = c.dim; a.norm_scale = s.mlp[l];
  a.out = s.act; a.ldo = c.ff;
  if ((e = launch_gemv<EPI_SWIGLU, true>(a, st)) != cudaSuccess) return e;
  // K6: down + residual
  a.W = s.wd[l]; a.rows = c.dim; a.K = c.ff; a.x = s.act; a.ldx = c.ff;
  a.out = s.h; a.ldo = c.dim; a.resid = s.h; a.ldr = c.dim;
  return launch_gemv<EPI_RESID, false>(a, st);
}

// Backbone pass over `chunk` prompt frames per stream starting at P->s0 (rows n = b*chunk + t).
static cudaError_t backbone_pass(csm_ctx* x, int B, int chunk, cudaStream_t st) {
  const csm_config& c = x->cfg;
  const int N = B * chunk;
  launch_k(k_embed_pass, dim3(N), dim3(256), 0, st, x->d_params, x->text_emb, x->audio_emb, c.codebooks, c.audio_vocab, c.backbone.dim,
                                  chunk, x->bb.h, x->row_stream, x->row_pos, x->row_slot, c.text_vocab, x->bb.rope_len,
                                  x->d_sync); COUNT_LAUNCH();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  RowMeta m{x->row_stream, x->row_pos, x->row_slot, 0, 0};
  for (int l = 0; l < c.backbone.layers; ++l)
    if ((e = run_layer(x, x->bb, l, N, m, st)) != cudaSuccess) return e;
  return cudaSuccess;
}

// Everything after the backbone layers of the LAST prompt row: final norm, codebook-0 head + sample,
// then the 31-step depth decoder (sesameai/models.py:160-184).  Requires chunk == 1 rows (n == b).
static cudaError_t frame_tail(csm_ctx* x, int B, cudaStream_t st) {
  const csm_config& c = x->cfg;
  const int D = c.backbone.dim, Dd = c.decoder.dim, V = c.audio_vocab, C = c.codebooks;
  cudaError_t e;
  // last_h = backbone.norm(h)  -> decoder input rows [0, B)
  launch_k(k_rmsnorm, dim3(B), dim3(256), 0, st, x->bb.h, D, x->bb.norm, D, c.norm_eps, x->dec_in, D); COUNT_LAUNCH();
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  GemvArgs a;
  memset(&a, 0, sizeof(a));
  a.eps = c.norm_eps;
  // codebook0_head
  a.W = x->c0_head; a.rows = V; a.K = D; a.x = x->dec_in; a.ldx = D; a.N = B; a.out = x->logits; a.ldo = x->Vp;
  if ((e = launch_gemv<EPI_PLAIN, false>(a, st)) != cudaSuccess) return e;
  launch_k(k_sample_step, dim3(B), dim3(SAMPLE_THREADS), 0, st, x->d_params, x->logits, x->Vp, 0, V, C, x->audio_emb, D,
                                              x->dec_in + (size_t)B * D, x->d_sync); COUNT_LAUNCH();
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  for (int i = 1; i < C; ++i) {
    const int N = (i == 1) ? 2 * B : B;           // first step carries [last_h, embed(c0)]
    const int pos0 = (i == 1) ? 0 : i;
    // projection (sesameai/models.py:173)
    memset(&a, 0, sizeof(a));
    a.eps = c.norm_eps;
    a.W = x->proj; a.rows = Dd; a.K = D; a.x = x->dec_in; a.ldx = D; a.N = N; a.out = x->dec.h; a.ldo = Dd;
    if ((e = launch_gemv<EPI_PLAIN, false>(a, st)) != cudaSuccess) return e;
    RowMeta m{nullptr, nullptr, nullptr, B, pos0};
    for (int l = 0; l < c.decoder.layers; ++l)
      if ((e = run_layer(x, x->dec, l, N, m, st)) != cudaSuccess) return e;
    // decoder.norm + audio_head[i-1] on the last position's rows
    memset(&a, 0, sizeof(a));
    a.eps = c.norm_eps;
    a.W = x->head_t + (size_t)(i - 1) * x->Vp * Dd; a.rows = V; a.K = Dd;
    a.x = x->dec.h + (size_t)(N - B) * Dd; a.ldx = Dd; a.N = B; a.norm_scale = x->dec.norm;
    a.out = x->logits; a.ldo = x->Vp;
    if ((e = launch_gemv<EPI_PLAIN, true>(a, st)) != cudaSuccess) return e;
    launch_k(k_sample_step, dim3(B), dim3(SAMPLE_THREADS), 0, st, x->d_params, x->logits, x->Vp, i, V, C, x->audio_emb, D,
                                                (i + 1 < C) ? x->dec_in : nullptr, x->d_sync); COUNT_LAUNCH();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  return cudaSuccess;
}


// ---------------------------------------------------------------------------------------------
// Megakernel phase table for one batch-1 decode frame (mega.cuh).

// Row-group height of a fragment-major matrix: the R in {8, 16} that gives the busiest CTA the fewest
// rows (ties -> 16: half as many mma instructions and epilogue items).
static int mega_pick_R(int rows, int ncta) {
  const int g16 = (rows + 15) / 16, g8 = (rows + 7) / 8;
  const int m16 = ((g16 + ncta - 1) / ncta) * 16, m8 = ((g8 + ncta - 1) / ncta) * 8;
  return m8 < m16 ? 8 : 16;
}

// Fragment-major re-pack (see mega.cuh, gemv_groups); rows >= src_rows are zero.
//   R == 8 : dst[group][k block of 32][lane][16 B] = W[8G + lane/4][32 kb + 8 (lane%4) .. +8]        (B operand)
//   R == 16: dst[group][k block][half h][lane][16 B] = { W[16G + 2g][k0], W[16G + 2g + 1][k0], W[16G + 2g][k0 + 2],
//            W[16G + 2g + 1][k0 + 2] } (pairs of bf16), g = lane/4, k0 = 32 kb + 8 (lane%4) + 4 h      (A operand)
__global__ void k_pack_frag(const bf16* __restrict__ src, int src_rows, int K, int R, bf16* __restrict__ dst) {
  const size_t KB = K / 32, upb = (size_t)R * 4;  // 16-byte units per block
  const size_t groups = (src_rows + R - 1) / R, total = groups * KB * upb;
  for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < total; u += (size_t)gridDim.x * blockDim.x) {
    const size_t blk = u / upb;
    const int w = (int)(u % upb);
    const size_t g = blk / KB, kb = blk % KB;
    const int L = w & 31;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (R == 16) {
      const int h = w >> 5;
      const int r0 = (int)g * 16 + 2 * (L >> 2);
      const size_t k0 = kb * 32 + (size_t)(L & 3) * 8 + 4 * h;
      const uint32_t* a = reinterpret_cast<const uint32_t*>(src + (size_t)r0 * K + k0);
      const uint32_t* b = reinterpret_cast<const uint32_t*>(src + (size_t)(r0 + 1) * K + k0);
      if (r0 < src_rows) { v.x = a[0]; v.z = a[1]; }
      if (r0 + 1 < src_rows) { v.y = b[0]; v.w = b[1]; }
    } else {
      const int row = (int)g * 8 + (L >> 2);
      const size_t k = kb * 32 + (size_t)(L & 3) * 8;
      if (row < src_rows) v = *reinterpret_cast<const uint4*>(src + (size_t)row * K + k);
    }
    reinterpret_cast<uint4*>(dst)[u] = v;
  }
}
// RoPE at one fixed position on rows of [q;k;v] (in place; same arithmetic as the EPI_ROPE_KV epilogue)
__global__ void __launch_bounds__(256) k_rope_table(bf16* __restrict__ qkv, const bf16* __restrict__ rope, int pos, int heads,
                                                    int kv_heads, int hd) {
  const int total = (heads + 2 * kv_heads) * hd, rot = (heads + kv_heads) * hd;
  bf16* row = qkv + (size_t)blockIdx.x * total;
  for (int p = threadIdx.x; p < rot / 2; p += blockDim.x) {
    const int r0 = 2 * p;
    const __nv_bfloat162 in = *reinterpret_cast<const __nv_bfloat162*>(row + r0);
    const float y0 = __low2float(in), y1 = __high2float(in);
    const __nv_bfloat162 cs = *reinterpret_cast<const __nv_bfloat162*>(rope + ((size_t)pos * (hd / 2) + ((r0 % hd) >> 1)) * 2);
    const float c = __low2float(cs), s = __high2float(cs);
    *reinterpret_cast<__nv_bfloat162*>(row + r0) = __floats2bfloat162_rn(rbf(__fsub_rn(__fmul_rn(y0, c), __fmul_rn(y1, s))),
                                                                         rbf(__fadd_rn(__fmul_rn(y1, c), __fmul_rn(y0, s))));
  }
}

static void pack_frag(const bf16* src, int rows, int K, int R, bf16* dst, cudaStream_t st) {
  k_pack_frag<<<1024, 256, 0, st>>>(src, rows, K, R, dst); COUNT_LAUNCH();
}

static int ilog2(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}
static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

struct MegaBuild {
  std::vector<mega::Phase> v;
  int ncta;
  int rot;  // CTA that receives the next phase's first row group (round-robin continues across phases)
};

static mega::Phase gemv_phase_desc(MegaBuild& mb, const bf16* Wf, int rows, int K, int R, const uint32_t* t_x, int ldx, int nb,
                                   int epi, const bf16* norm_scale, float eps, uint32_t* t_out, int ldo) {
  mega::Phase ph;
  memset(&ph, 0, sizeof(ph));
  ph.type = mega::PH_GEMV; ph.epi = epi; ph.norm = norm_scale != nullptr; ph.nb = nb;
  ph.W = Wf; ph.rows = rows; ph.K = K; ph.R = R;
  ph.G = (rows + R - 1) / R;
  ph.rot = mb.rot; ph.gq = ph.G / mb.ncta; ph.gr = ph.G % mb.ncta;
  mb.rot = (mb.rot + ph.gr) % mb.ncta;
  ph.tb = R * K * 2 / mega::NW;
  ph.chunk = ph.tb < mega::SLOT_BYTES ? ph.tb : mega::SLOT_BYTES;
  ph.nch = ph.tb / ph.chunk;
  ph.kchunk = ph.chunk / (R * 2);
  ph.nblk = ph.chunk / (R * 64);
  ph.inv_K = (K & (K - 1)) == 0 ? 1.0f / (float)K : 0.f;
  ph.t_x = t_x; ph.ldx = ldx; ph.norm_scale = norm_scale; ph.eps = eps; ph.t_out = t_out; ph.ldo = ldo;
  return ph;
}

// src[n]: index of the phase that last wrote row n of the stack's residual stream (updated here)
static void stack_phases(csm_ctx* x, StackDev& s, const int* R4, int l, int nb, int pos_mode, int pos0, bool fused_attn,
                         MegaBuild& mb, int* src, int keep_mask = 0 /* bit 0 qkv, 1 o, 2 gate/up, 3 down: L2 evict-last */,
                         int qkv_from = -1 /* >= 0: q / k / v come from that (sample) phase's table gather, no QKV phase */) {
  const csm_stack_config& c = s.c;
  const float eps = x->cfg.norm_eps;
  bf16* kc = s.kc + s.kv_layer_stride * l;
  bf16* vc = s.vc + s.kv_layer_stride * l;
  auto with_attn = [&](mega::Phase ph) {
    ph.t_q = s.t_q; ph.t_kv = s.t_kv; ph.q_rs = s.rs_q; ph.kv_rs = s.rs_kv; ph.kc = kc; ph.vc = vc; ph.rope = s.rope; ph.heads = c.heads; ph.kv_heads = c.kv_heads;
    ph.hd = s.hd; ph.hd_shift = s.hd == 128 ? 7 : 6; ph.slots = s.slots;
    ph.grp_shift = ilog2(c.heads / c.kv_heads); ph.heads_shift = ilog2(c.heads); ph.pos_mode = pos_mode; ph.pos0 = pos0;
    return ph;
  };
  mega::Phase q = with_attn(gemv_phase_desc(mb, s.fqkv[l], (c.heads + 2 * c.kv_heads) * s.hd, c.dim, R4[0], s.t_h, c.dim, nb,
                                            EPI_ROPE_KV, s.sa[l], eps, nullptr, 0));
  q.x_src[0] = src[0]; q.x_src[1] = src[1];
  q.x_rs = s.rs_h;
  q.keep = keep_mask & 1;
  int iq = (int)mb.v.size();
  if (qkv_from >= 0) iq = qkv_from;
  else mb.v.push_back(q);
  int iatt = -1;
  if (!fused_attn) {
    mega::Phase a;
    memset(&a, 0, sizeof(a));
    a.type = mega::PH_ATTN; a.nb = nb; a.t_out = s.t_att; a.out_rs = s.rs_att;
    a = with_attn(a);
    a.q_src = iq;
    iatt = (int)mb.v.size();
    mb.v.push_back(a);
  }
  mega::Phase o = with_attn(gemv_phase_desc(mb, s.fo[l], c.dim, c.dim, R4[1], s.t_att, c.dim, nb, EPI_RESID, nullptr, eps,
                                            s.t_h, c.dim));
  o.attn_prologue = fused_attn ? 1 : 0;
  o.x_rs = s.rs_att; o.out_rs = s.rs_h;
  o.keep = (keep_mask >> 1) & 1;
  o.q_src = iq;
  o.x_src[0] = o.x_src[1] = iatt;
  o.resid_src[0] = src[0]; o.resid_src[1] = src[1];
  src[0] = src[1] = (int)mb.v.size();
  mb.v.push_back(o);
  mega::Phase g = gemv_phase_desc(mb, s.fgu[l], 2 * c.ff, c.dim, R4[2], s.t_h, c.dim, nb, EPI_SWIGLU, s.mlp[l], eps, s.t_act, c.ff);
  g.x_src[0] = g.x_src[1] = src[0];
  g.x_rs = s.rs_h; g.out_rs = s.rs_act;
  g.keep = (keep_mask >> 2) & 1;
  const int ig = (int)mb.v.size();
  mb.v.push_back(g);
  mega::Phase d = gemv_phase_desc(mb, s.fd[l], c.dim, c.ff, R4[3], s.t_act, c.ff, nb, EPI_RESID, nullptr, eps, s.t_h, c.dim);
  d.x_src[0] = d.x_src[1] = ig;
  d.x_rs = s.rs_act; d.out_rs = s.rs_h;
  d.keep = (keep_mask >> 3) & 1;
  d.resid_src[0] = d.resid_src[1] = src[0];
  src[0] = src[1] = (int)mb.v.size();
  mb.v.push_back(d);
}

static void build_mega_phases(csm_ctx* x, MegaBuild& mb) {
  const csm_config& c = x->cfg;
  const int D = c.backbone.dim, Dd = c.decoder.dim, V = c.audio_vocab, C = c.codebooks;
  const float eps = c.norm_eps;
  auto sample = [&](int cb, int logits_src, uint32_t* t_next) {
    mega::Phase s;
    memset(&s, 0, sizeof(s));
    s.type = mega::PH_SAMPLE; s.cb = cb; s.V = V; s.C = C; s.D = D; s.t_logits = x->t_logits; s.logits_src = logits_src;
    s.t_next = t_next; s.next_rs = x->dec.rs_h; s.next_table = x->proj_table; s.next_ld = Dd;
    if (x->qkv_table_ok && cb >= 1 && t_next) {  // feeds step cb + 1 >= 2 (one row): first-layer q / k / v by gather
      const StackDev& d = x->dec;
      s.qkv_table = x->qkv_table; s.t_q = d.t_q; s.t_kv = d.t_kv; s.q_rs = d.rs_q; s.kv_rs = d.rs_kv;
      s.kc = d.kc; s.vc = d.vc;  // layer 0
      s.heads = d.c.heads; s.kv_heads = d.c.kv_heads; s.hd = d.hd; s.hd_shift = d.hd == 128 ? 7 : 6; s.slots = d.slots;
      s.pos0 = cb + 1;
    }
    return s;
  };
  mega::Phase e;
  memset(&e, 0, sizeof(e));
  e.type = mega::PH_EMBED; e.V = V; e.C = C; e.D = D; e.audio_emb = x->audio_emb; e.text_emb = x->text_emb; e.t_out = x->bb.t_h; e.out_rs = x->bb.rs_h;
  e.TV = c.text_vocab; e.rope_len = x->bb.rope_len;
  mb.v.push_back(e);
  int src[2] = {0, 0};
  for (int l = 0; l < c.backbone.layers; ++l) stack_phases(x, x->bb, x->mega_Rbb, l, 1, mega::POS_BACKBONE, 0, false, mb, src);
  // one phase: logits of codebook 0 AND projection(last_h) (depth-decoder row 0), from the stacked matrix;
  // every later decoder input is a row of the projection(embedding) table, so no projection phase remains
  mega::Phase h0 = gemv_phase_desc(mb, x->f_head0, x->Vf + Dd, D, x->mega_Rh0, x->bb.t_h, D, 1, EPI_PLAIN, x->bb.norm, eps,
                                   x->t_logits, x->Vf);
  h0.x_src[0] = h0.x_src[1] = src[0];
  h0.t_out2 = x->dec.t_h; h0.split_row = x->Vf;
  h0.x_rs = x->bb.rs_h; h0.out_rs = 0; h0.out2_rs = x->dec.rs_h;
  const int ih0 = (int)mb.v.size();
  mb.v.push_back(h0);
  const int is0 = (int)mb.v.size();
  mb.v.push_back(sample(0, ih0, x->dec.t_h + Dd));
  int dsrc[2] = {ih0, is0};
  int kv_prev = -1;  // done_src of the previous codebook step's release / acquire
  for (int i = 1; i < C; ++i) {
    const int nb = (i == 1) ? 2 : 1, pos0 = (i == 1) ? 0 : i;
    const int ifirst = (int)mb.v.size();  // the step's first phase
    for (int l = 0; l < c.decoder.layers; ++l)
      stack_phases(x, x->dec, x->mega_Rdec, l, nb, mega::POS_FIXED, pos0, true, mb, dsrc, (x->mega_keep >> (4 * (l & 7))) & 15,
                   (x->qkv_table_ok && i >= 2 && l == 0) ? dsrc[0] : -1);
    // KV rows of steps <= i are read from the cache in step i + 1: released after the last layer's gate/up phase
    // (= before its down phase), acquired before the sampling phase (mega::kv_step_sync; flags sit on the NEXT phase)
    const bool kv_step = i + 1 < C && mb.ncta <= 1024;
    const int igu = (int)mb.v.size() - 2;
    if (kv_prev >= 0) { mb.v[ifirst].kv_sync = 3; mb.v[ifirst].done_src = kv_prev; mb.v[ifirst].t_done = x->t_done; }
    if (kv_step) { mb.v[igu + 1].kv_sync = 1 | (kv_prev << 8); mb.v[igu + 1].done_src = igu; mb.v[igu + 1].t_done = x->t_done; }
    mega::Phase h = gemv_phase_desc(mb, x->f_heads + (size_t)(i - 1) * x->Vf * Dd, x->Vf, Dd, x->mega_Rh,
                                    x->dec.t_h + (size_t)(nb - 1) * Dd, Dd, 1, EPI_PLAIN, x->dec.norm, eps, x->t_logits, x->Vf);
    h.x_src[0] = h.x_src[1] = dsrc[nb - 1];
    h.x_rs = x->dec.rs_h; h.out_rs = 0;
    const int ih = (int)mb.v.size();
    mb.v.push_back(h);
    const int is = (int)mb.v.size();
    mb.v.push_back(sample(i, ih, (i + 1 < C) ? x->dec.t_h : nullptr));
    if (kv_step) { mb.v[is].kv_sync = 2 | (igu << 8); mb.v[is].done_src = igu; mb.v[is].t_done = x->t_done; }
    kv_prev = kv_step ? igu : -1;
    dsrc[0] = is;
  }
}

static int setup_mega(csm_ctx* x, cudaStream_t st) {
  x->mega_ok = false;
  x->frag_ok = false;
  x->trace = nullptr;
  const csm_config& c = x->cfg;
  // fused small attention needs <= 32 cached keys; scores of the backbone attention sit in the x buffer
  if (c.codebooks > 32 || c.max_seq_len * 4 + (8 * mega::NCT + 3 * 128) * 4 > mega::XBUF_ELEMS * 2) return CSM_OK;
  if (x->dec.hd != 128 || 2 * c.decoder.dim > 2048 || c.decoder.kv_heads > 2) return CSM_OK;  // fused attention layout
  if (!is_pow2(c.decoder.heads) || !is_pow2(c.decoder.heads / c.decoder.kv_heads) || 2 * (c.decoder.heads / c.decoder.kv_heads) > 8)
    return CSM_OK;  // (activation row, head in group) columns must fit the 8-wide mma tile
  if (c.backbone.dim > 8 * mega::NCT || c.decoder.dim > 8 * mega::NCT) return CSM_OK;  // one norm unit per thread
  if (mega_phase_count(c) > 2046) return CSM_OK;  // 11-bit phase tags
  int dev = 0, sms = 0, coop = 0, occ = 0;
  CU_TRY(cudaGetDevice(&dev));
  CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CU_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  CU_TRY(cudaFuncSetAttribute(mega::k_frame_mega, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mega::SMEM_BYTES));
  CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mega::k_frame_mega, mega::NTHREADS, mega::SMEM_BYTES));
  if (!coop || occ < 1) return CSM_OK;
  // fragment-major copies of every matrix the frame touches
  const int D = c.backbone.dim, Dd = c.decoder.dim;
  auto pick_stack = [&](const StackDev& s, int* r4) {
    r4[0] = mega_pick_R((s.c.heads + 2 * s.c.kv_heads) * s.hd, sms);
    r4[1] = mega_pick_R(s.c.dim, sms);
    r4[2] = mega_pick_R(2 * s.c.ff, sms);
    r4[3] = mega_pick_R(s.c.dim, sms);
  };
  pick_stack(x->bb, x->mega_Rbb);
  pick_stack(x->dec, x->mega_Rdec);
  x->mega_Rh0 = mega_pick_R(x->Vf + Dd, sms);
  x->mega_Rh = mega_pick_R(x->Vf, sms);
  auto pack_stack_frag = [&](StackDev& s, const int* r4) {
    for (int l = 0; l < s.c.layers; ++l) {
      pack_frag(s.wqkv[l], (s.c.heads + 2 * s.c.kv_heads) * s.hd, s.c.dim, r4[0], s.fqkv[l], st);
      pack_frag(s.wo[l], s.c.dim, s.c.dim, r4[1], s.fo[l], st);
      pack_frag(s.wgu[l], 2 * s.c.ff, s.c.dim, r4[2], s.fgu[l], st);
      pack_frag(s.wd[l], s.c.dim, s.c.ff, r4[3], s.fd[l], st);
    }
  };
  pack_stack_frag(x->bb, x->mega_Rbb);
  pack_stack_frag(x->dec, x->mega_Rdec);
  CU_TRY(cudaMemsetAsync(x->f_head0, 0, (size_t)(x->Vf + Dd) * D * sizeof(bf16), st));
  pack_frag(x->c0_head, c.audio_vocab, D, x->mega_Rh0, x->f_head0, st);
  pack_frag(x->proj, Dd, D, x->mega_Rh0, x->f_head0 + (size_t)x->Vf * D, st);
  CU_TRY(cudaMemsetAsync(x->f_heads, 0, (size_t)(c.codebooks - 1) * x->Vf * Dd * sizeof(bf16), st));
  for (int i = 0; i + 1 < c.codebooks; ++i)
    pack_frag(x->head_t + (size_t)i * x->Vp * Dd, c.audio_vocab, Dd, x->mega_Rh, x->f_heads + (size_t)i * x->Vf * Dd, st);
  CU_TRY(cudaMemsetAsync(x->tagged_base, 0, x->tagged_bytes, st));  // stale tags of an earlier context must never match
  CU_TRY(cudaGetLastError());
  x->frag_ok = true;
  // First decoder layer's [q;k;v] of every projection(embedding) row, RoPE applied at the position the row is used
  // at (codebook cb's token enters the decoder at position cb + 1): rows of cb >= 1 replace a QKV phase per step.
  x->qkv_table_ok = false;
  if (!getenv("CSM_MEGA_NO_QKV_TABLE") && c.decoder.dim % 64 == 0 && (size_t)c.audio_vocab * Dd <= (size_t)x->max_rows * D &&
      (c.decoder.heads + 2 * c.decoder.kv_heads) * x->dec.hd <= 3 * 4 * mega::NCT /* the sample phase gathers <= 3 units per thread */) {
    const StackDev& d = x->dec;
    const int qkv_cols = (d.c.heads + 2 * d.c.kv_heads) * d.hd, V = c.audio_vocab;
    for (int cb = 1; cb + 1 < c.codebooks; ++cb) {
      const bf16* rows = x->proj_table + (size_t)cb * V * Dd;
      bf16* out = x->qkv_table + (size_t)cb * V * qkv_cols;
      launch_k(k_rmsnorm, dim3(V), dim3(256), 0, st, rows, Dd, d.sa[0], Dd, c.norm_eps, x->bb.xn, Dd); COUNT_LAUNCH();
      int rc = launch_gemm_tc(x->bb.xn, Dd, V, Dd, d.wqkv[0], qkv_cols, out, qkv_cols, tc::EPI_STORE, nullptr, st);
      if (rc != CSM_OK) return rc;
      k_rope_table<<<V, 256, 0, st>>>(out, d.rope, cb + 1, d.c.heads, d.c.kv_heads, d.hd); COUNT_LAUNCH();
    }
    CU_TRY(cudaGetLastError());
    x->qkv_table_ok = true;
  }

  // Which depth-decoder matrices stay in L2 across the 31 codebook steps (see mega.cuh, producer_loop).
  // Measured (profiles/r1_mega_l2_keep.txt): no subset helps -- the stream phases are bound by the per-chunk
  // consumer loop and the hand-off chain, not by HBM bandwidth -- so the default is none; CSM_MEGA_KEEP=<hex>
  // sets it for experiments (4 bits per layer: 1 qkv, 2 o, 4 gate/up, 8 down).
  x->mega_keep = 0;
  if (const char* e = getenv("CSM_MEGA_KEEP")) x->mega_keep = (unsigned)strtoul(e, nullptr, 16);
  if (x->mega_keep) {
    // evict-last lines only persist inside the L2 set-aside: open it as far as the device allows
    int persist_max = 0;
    CU_TRY(cudaDeviceGetAttribute(&persist_max, cudaDevAttrMaxPersistingL2CacheSize, dev));
    if (persist_max > 0) CU_TRY(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)persist_max));
  }
  MegaBuild mb;
  mb.ncta = sms;
  mb.rot = 0;
  build_mega_phases(x, mb);
  std::vector<mega::Phase>& v = mb.v;
  if ((int)v.size() != mega_phase_count(x->cfg, x->qkv_table_ok)) return set_err(CSM_ERR_ARG, "internal: phase count mismatch");
  memset(&x->pf_table, 0, sizeof(x->pf_table));
  x->pf_table.t_done = x->t_done;
  for (const mega::Phase& ph : v) {
    if (ph.type != mega::PH_GEMV) continue;
    if (x->pf_table.n >= mega::MAX_GEMV) return CSM_OK;  // too deep for the parameter-space table: per-op path only
    if ((ph.G + sms - 1) / sms > mega::MAX_LOCAL_GROUPS) return CSM_OK;
    if (ph.K % (32 * mega::NW) != 0) return CSM_OK;
    mega::PfDesc& d = x->pf_table.d[x->pf_table.n++];
    d.W = ph.W; d.G = ph.G; d.rot = ph.rot; d.group_bytes = ph.R * ph.K * 2;
    if (ph.tb % ph.chunk != 0 || ph.chunk % (ph.R * 64) != 0) return CSM_OK;  // slices must be whole chunks of whole blocks
    d.chunk_nch = ph.chunk | (ph.nch << 16) | (ph.keep ? (1 << 30) : 0);
  }
  CU_TRY(cudaMemcpyAsync(x->d_phases, v.data(), v.size() * sizeof(mega::Phase), cudaMemcpyHostToDevice, st));
  CU_TRY(cudaStreamSynchronize(st));  // v is host stack memory
  x->n_phases = (int)v.size();
  x->mega_grid = sms;
  x->mega_ok = true;
  return CSM_OK;
}

static int launch_mega(csm_ctx* x, const FrameParams& p, cudaStream_t st) {
  mega::k_mega_prepare<<<1, 96, 0, st>>>(x->d_params, p, x->d_sync, x->cfg.codebooks, x->cfg.audio_vocab, x->cfg.text_vocab,
                                         x->bb.rope_len); COUNT_LAUNCH();
  CU_TRY(cudaGetLastError());
  const mega::Phase* ph = x->d_phases;
  int n = x->n_phases;
  const FrameParams* dp = x->d_params;
  mega::Sync* sy = x->d_sync;
  unsigned long long* trc = x->trace;
  void* args[] = {(void*)&ph, (void*)&n, (void*)&dp, (void*)&sy, (void*)&trc, (void*)&x->pf_table};
  CU_TRY(cudaLaunchCooperativeKernel((const void*)mega::k_frame_mega, dim3(x->mega_grid), dim3(mega::NTHREADS), args,
                                     mega::SMEM_BYTES, st));
  COUNT_LAUNCH();
  return CSM_OK;
}

// ---------------------------------------------------------------------------------------------
static int pack_stack(StackDev& s, const csm_layer_weights* lw, cudaStream_t st) {
  const csm_stack_config& c = s.c;
  const size_t D8 = c.dim / 8;
  const size_t qn = (size_t)c.heads * s.hd * D8, kn = (size_t)c.kv_heads * s.hd * D8;
  s.wo.resize(c.layers); s.wd.resize(c.layers); s.sa.resize(c.layers); s.mlp.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) {
    const csm_layer_weights& w = lw[l];
    if (!w.q_proj || !w.k_proj || !w.v_proj || !w.output_proj || !w.w1 || !w.w2 || !w.w3 || !w.sa_norm || !w.mlp_norm)
      return set_err(CSM_ERR_ARG, "null layer weight pointer");
    k_copy_rows<<<256, 256, 0, st>>>((const bf16*)w.q_proj, s.wqkv[l], qn); COUNT_LAUNCH();
    k_copy_rows<<<256, 256, 0, st>>>((const bf16*)w.k_proj, s.wqkv[l] + qn * 8, kn); COUNT_LAUNCH();
    k_copy_rows<<<256, 256, 0, st>>>((const bf16*)w.v_proj, s.wqkv[l] + (qn + kn) * 8, kn); COUNT_LAUNCH();
    k_interleave_rows<<<512, 256, 0, st>>>((const bf16*)w.w1, (const bf16*)w.w3, s.wgu[l], c.ff, (int)D8); COUNT_LAUNCH();
    s.wo[l] = (const bf16*)w.output_proj; s.wd[l] = (const bf16*)w.w2;
    s.sa[l] = (const bf16*)w.sa_norm; s.mlp[l] = (const bf16*)w.mlp_norm;
  }
  CU_TRY(cudaGetLastError());
  return CSM_OK;
}

extern "C" int32_t csm_create(const csm_config* cfg, const csm_weights* w, int32_t max_batch, void* workspace,
                              size_t workspace_bytes, void* stream, csm_ctx** out) {
  if (!out) return set_err(CSM_ERR_ARG, "out is null");
  *out = nullptr;
  if (!valid_cfg(cfg)) return set_err(CSM_ERR_ARG, "unsupported csm_config (head_dim must be 64/128, dims multiples of 256)");
  if (!w || !workspace || max_batch < 1) return set_err(CSM_ERR_ARG, "null weights/workspace or max_batch < 1");
  NvtxRange nvtx_create("csm.create");
  int ndev = 0;
  CU_TRY(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return set_err(CSM_ERR_CUDA, "no CUDA device (libcsm_b200 has no CPU fallback)");
  if (((uintptr_t)workspace & 255) != 0) return set_err(CSM_ERR_ARG, "workspace must be 256-byte aligned");
  const size_t need = csm_workspace_bytes(cfg, max_batch);
  if (workspace_bytes < need) return set_err(CSM_ERR_WORKSPACE, "workspace too small");
  if (!w->text_embeddings || !w->audio_embeddings || !w->projection || !w->codebook0_head || !w->audio_head ||
      !w->backbone_norm || !w->decoder_norm || !w->backbone_rope || !w->decoder_rope || !w->backbone_layers ||
      !w->decoder_layers)
    return set_err(CSM_ERR_ARG, "null weight pointer");
  if (w->backbone_rope_len < cfg->max_seq_len || w->decoder_rope_len < cfg->codebooks)
    return set_err(CSM_ERR_ARG, "rope table shorter than the cache");
  CU_TRY(init_kernel_attrs());
  csm_ctx* x = new (std::nothrow) csm_ctx();
  if (!x) return set_err(CSM_ERR_ARG, "out of host memory");
  x->cfg = *cfg;
  x->max_batch = max_batch;
  carve_all(x, (char*)workspace);
  cudaStream_t st = (cudaStream_t)stream;
  x->text_emb = (const bf16*)w->text_embeddings;
  x->audio_emb = (const bf16*)w->audio_embeddings;
  x->proj = (const bf16*)w->projection;
  x->c0_head = (const bf16*)w->codebook0_head;
  x->bb.norm = (const bf16*)w->backbone_norm; x->bb.rope = (const bf16*)w->backbone_rope; x->bb.rope_len = w->backbone_rope_len;
  x->dec.norm = (const bf16*)w->decoder_norm; x->dec.rope = (const bf16*)w->decoder_rope; x->dec.rope_len = w->decoder_rope_len;
  int rc;
  if ((rc = pack_stack(x->bb, w->backbone_layers, st)) != CSM_OK || (rc = pack_stack(x->dec, w->decoder_layers, st)) != CSM_OK) {
    delete x;
    return rc;
  }
  {
    const int K = cfg->decoder.dim, V = cfg->audio_vocab;
    dim3 grid((x->Vp + 31) / 32, (K + 31) / 32, cfg->codebooks - 1), block(32, 8);
    k_transpose_heads<<<grid, block, 0, st>>>((const bf16*)w->audio_head, x->head_t, K, V, x->Vp); COUNT_LAUNCH();
  }
  {
    // projection(embedding) table on the tensor cores: [(C-1)*V, D] x [Dd, D]^T
    if (cfg->backbone.dim % 64 == 0) {
      rc = launch_gemm_tc(x->audio_emb, cfg->backbone.dim, (cfg->codebooks - 1) * cfg->audio_vocab, cfg->backbone.dim, x->proj,
                          cfg->decoder.dim, x->proj_table, cfg->decoder.dim, tc::EPI_STORE, nullptr, st);
      if (rc != CSM_OK) {
        delete x;
        return rc;
      }
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&x->cap_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete x;
    return set_err(CSM_ERR_CUDA, "csm_create: %s", cudaGetErrorString(e));
  }
  {
    // status word: frame counter 0, no error, mirror in mapped host memory
    unsigned int* dmirror = nullptr;
    e = cudaHostAlloc((void**)&x->h_error, sizeof(unsigned int), cudaHostAllocMapped);
    if (e == cudaSuccess) {
      *x->h_error = 0;
      e = cudaHostGetDevicePointer((void**)&dmirror, x->h_error, 0);
    }
    mega::Sync init;
    init.seq = 0; init.error = 0; init.host_error = dmirror; init.att_part = x->mega_att_part;
    if (e == cudaSuccess)
      e = cudaMemsetAsync(x->mega_att_part, 0, (size_t)mega::ATTN_SPLITS * x->cfg.backbone.heads * mega::ATTN_PART_WORDS * sizeof(unsigned int), st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(x->d_sync, &init, sizeof(init), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // ``init`` is stack memory
    if (e != cudaSuccess) {
      csm_destroy(x);
      return set_err(CSM_ERR_CUDA, "csm_create: %s", cudaGetErrorString(e));
    }
  }
  e = cudaMemsetAsync(x->tc_counters, 0, TC_SPLIT_TILES * sizeof(unsigned int), st);
  if (e != cudaSuccess) {
    csm_destroy(x);
    return set_err(CSM_ERR_CUDA, "csm_create: %s", cudaGetErrorString(e));
  }
  x->cache_len = 0;
  x->lane_len.assign(max_batch, 0);
  x->enabled = true;
  if ((rc = setup_mega(x, st)) != CSM_OK) {
    csm_destroy(x);
    return rc;
  }
  *out = x;
  return CSM_OK;
}

extern "C" void csm_destroy(csm_ctx* x) {
  if (!x) return;
  for (auto& kv : x->graphs) cudaGraphExecDestroy(kv.second);
  if (x->cap_stream) cudaStreamDestroy(x->cap_stream);
  if (x->h_error) cudaFreeHost(x->h_error);
  delete x;
}

extern "C" int32_t csm_reset_caches(csm_ctx* x) {
  if (!x || !x->enabled) return set_err(CSM_ERR_STATE, "caches are not enabled");
  x->cache_len = 0;
  x->lane_len.assign(x->max_batch, 0);
  return CSM_OK;
}
extern "C" int32_t csm_lane_reset(csm_ctx* x, int32_t lane) {
  if (!x || !x->enabled) return set_err(CSM_ERR_STATE, "caches are not enabled");
  if (lane < 0 || lane >= x->max_batch) return set_err(CSM_ERR_ARG, "lane out of range");
  x->lane_len[lane] = 0;
  if (lane == 0) x->cache_len = 0;
  return CSM_OK;
}
extern "C" int32_t csm_lane_len(const csm_ctx* x, int32_t lane) {
  return (x && lane >= 0 && lane < x->max_batch) ? x->lane_len[lane] : -1;
}
extern "C" int32_t csm_cache_len(const csm_ctx* x) { return x ? x->cache_len : -1; }

extern "C" int32_t csm_check_error(csm_ctx* x, int32_t clear) {
  if (!x || !x->h_error) return 0;
  const unsigned int code = *reinterpret_cast<volatile unsigned int*>(x->h_error);
  if (code && clear) *reinterpret_cast<volatile unsigned int*>(x->h_error) = 0;
  return (int32_t)code;
}

// decode steps run row-batched (RMSNorm / linear / RoPE kernels over all streams) from 16 streams on the
// tcgen05 GEMM, and from 2 streams when the skinny GEMM has its fragment-major weights
static bool rows_path(const csm_ctx* x, int B) {
  if (x->cfg.backbone.dim % 64 || x->cfg.decoder.dim % 64) return false;
  return B >= DECODE_TC_MIN || (x->frag_ok && B >= 2 && 2 * B <= SKINNY_MAX_ROWS);
}

static int get_graph(csm_ctx* x, int B, cudaGraphExec_t* out) {
  const int key = B | (x->long_ctx ? 1 << 24 : 0);  // the long-context step uses other attention kernels
  auto it = x->graphs.find(key);
  if (it != x->graphs.end()) {
    *out = it->second;
    return CSM_OK;
  }
  cudaGraph_t g = nullptr;
  const unsigned long long before = g_launches.load();
  CU_TRY(cudaStreamBeginCapture(x->cap_stream, cudaStreamCaptureModeThreadLocal));
  cudaError_t e = cudaSuccess;
  int rc_tc = CSM_OK;
  if (rows_path(x, B)) {
    rc_tc = backbone_pass_tc(x, B, 1, x->cap_stream);
    if (rc_tc == CSM_OK) rc_tc = frame_tail_tc(x, B, x->cap_stream);
  } else {
    e = backbone_pass(x, B, 1, x->cap_stream);
    if (e == cudaSuccess) e = frame_tail(x, B, x->cap_stream);
  }
  cudaError_t e2 = cudaStreamEndCapture(x->cap_stream, &g);
  if (rc_tc != CSM_OK) {
    if (g) cudaGraphDestroy(g);
    return rc_tc;
  }
  x->graph_nodes[key] = g_launches.load() - before;  // captured, not executed
  g_launches.store(before);
  if (e != cudaSuccess || e2 != cudaSuccess) {
    if (g) cudaGraphDestroy(g);
    return set_err(CSM_ERR_CUDA, "graph capture: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
  }
  cudaGraphExec_t ge = nullptr;
  e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return set_err(CSM_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(e));
  x->graphs[key] = ge;
  *out = ge;
  return CSM_OK;
}


extern "C" int32_t csm_generate_frame(csm_ctx* x, const int64_t* tokens, const uint8_t* tokens_mask,
                                      const int64_t* input_pos, int32_t B, int32_t S, float temperature, int32_t topk,
                                      const csm_frame_opts* opts, int32_t* out, void* stream) {
  if (!x || !x->enabled) return set_err(CSM_ERR_STATE, "backbone caches are not enabled");
  if (!tokens || !tokens_mask || !input_pos || !out || B < 1 || S < 1) return set_err(CSM_ERR_ARG, "bad tokens/mask/pos/out");
  if (B > x->max_batch) return set_err(CSM_ERR_STATE, "batch size exceeds the batch the caches were set up for");
  if (!(temperature > 0.f) || topk < 1) return set_err(CSM_ERR_ARG, "temperature must be > 0 and topk >= 1");
  NvtxRange nvtx_frame("csm.generate_frame");
  cudaStream_t st = (cudaStream_t)stream;
  // cache lanes of the batch rows: identity for the reference's lock-step batch, any distinct lanes for a
  // continuous-batching caller; every lane keeps its own length
  const int32_t* lanes = opts ? opts->lanes : nullptr;
  std::vector<int> lane(B);
  bool uniform = lanes == nullptr;
  for (int b = 0; b < B; ++b) {
    lane[b] = lanes ? lanes[b] : b;
    if (lane[b] < 0 || lane[b] >= x->max_batch) return set_err(CSM_ERR_ARG, "lane out of range");
    for (int a = 0; a < b; ++a)
      if (lane[a] == lane[b]) return set_err(CSM_ERR_ARG, "two batch rows name the same cache lane");
    if (x->lane_len[lane[b]] + S > x->cfg.max_seq_len)
      return set_err(CSM_ERR_OVERFLOW, "KV cache overflow (cache_pos + seq_len > max_seq_len)");
    if (x->lane_len[lane[b]] != x->lane_len[lane[0]]) uniform = false;
  }
  {
    // decode rows at a long context (a voice prompt) take the flash-decoding attention kernels
    static const int long_min = getenv("CSM_ATT_LONG_MIN") ? atoi(getenv("CSM_ATT_LONG_MIN")) : 256;
    int mx = 0;
    for (int b = 0; b < B; ++b) mx = x->lane_len[lane[b]] > mx ? x->lane_len[lane[b]] : mx;
    x->long_ctx = long_min > 0 && mx + S >= long_min;
  }
  FrameParams p;
  memset(&p, 0, sizeof(p));
  p.tokens = tokens; p.mask = tokens_mask; p.pos = input_pos; p.out = out;
  p.temperature = temperature; p.topk = topk; p.B = B; p.S = S; p.cache_len = x->lane_len[lane[0]];
  if (!uniform) {
    for (int off = 0; off < B; off += 256) {
      LaneChunk ch;
      ch.n = B - off < 256 ? B - off : 256; ch.off = off; ch.B = B;
      for (int i = 0; i < ch.n; ++i) {
        ch.lane[i] = lane[off + i];
        ch.len[i] = x->lane_len[lane[off + i]];
      }
      k_set_lanes<<<1, 256, 0, st>>>(x->d_lane_meta, ch); COUNT_LAUNCH();
    }
    CU_TRY(cudaGetLastError());
    p.lane_meta = x->d_lane_meta;
  }
  auto advance = [&]() {
    for (int b = 0; b < B; ++b) x->lane_len[lane[b]] += S;
    x->cache_len = x->lane_len[0];
  };
  int path = 0;
  if (opts) {
    p.noise = (const bf16*)opts->noise; p.forced = opts->forced; p.logits_out = (bf16*)opts->logits_out;
    p.sampled_out = opts->sampled_out; p.seed = opts->seed; p.offset = opts->offset;
    path = opts->path;
  }
  if (path == CSM_PATH_AUTO) path = (B == 1 && x->mega_ok) ? CSM_PATH_MEGA : CSM_PATH_GRAPH;
  if (path == CSM_PATH_MEGA && (B != 1 || !x->mega_ok)) return set_err(CSM_ERR_ARG, "megakernel path needs batch 1");
  // prompt rows [0, S-1): row-batched passes (skinny / tcgen05 GEMMs, tiled attention) from PREFILL_TC_MIN rows on,
  // else per-op small-row passes of up to PREFILL_CHUNK frames per stream
  int prefill_path = opts ? opts->prefill : 0;
  if (prefill_path == CSM_PREFILL_AUTO)
  {
    static const int tc_min = getenv("CSM_PREFILL_TC_MIN") ? atoi(getenv("CSM_PREFILL_TC_MIN")) : PREFILL_TC_MIN;
    prefill_path = ((long long)B * (S - 1) >= tc_min && x->cfg.backbone.dim % 64 == 0) ? CSM_PREFILL_TENSOR : CSM_PREFILL_SMALL_ROW;
  }
  const int per_pass = prefill_path == CSM_PREFILL_TENSOR ? (PREFILL_TC_ROWS / B > 0 ? PREFILL_TC_ROWS / B : 1) : PREFILL_CHUNK;
  for (int s0 = 0; s0 < S - 1; s0 += per_pass) {
    NvtxRange nvtx_prefill("csm.prefill");
    const int chunk = (S - 1 - s0) < per_pass ? (S - 1 - s0) : per_pass;
    p.s0 = s0;
    launch_k(k_set_params, dim3(1), dim3(1), 0, st, x->d_params, p, s0 == 0 ? x->d_sync : nullptr); COUNT_LAUNCH();
    CU_TRY(cudaGetLastError());
    if (prefill_path == CSM_PREFILL_TENSOR) {
      int rc = backbone_pass_tc(x, B, chunk, st);
      if (rc != CSM_OK) return rc;
    } else {
      CU_TRY(backbone_pass(x, B, chunk, st));
    }
  }
  // last row + frame tail: persistent megakernel (batch 1) or the captured per-op graph
  p.s0 = S - 1;
  if (path == CSM_PATH_MEGA) {
    NvtxRange nvtx_mega("csm.decode.mega");
    int rc = launch_mega(x, p, st);
    if (rc != CSM_OK) return rc;
    advance();
    return CSM_OK;
  }
  NvtxRange nvtx_graph("csm.decode.graph");
  launch_k(k_set_params, dim3(1), dim3(1), 0, st, x->d_params, p, S == 1 ? x->d_sync : nullptr); COUNT_LAUNCH();
  CU_TRY(cudaGetLastError());
  if (path == CSM_PATH_DIRECT) {
    if (rows_path(x, B)) {
      int rc = backbone_pass_tc(x, B, 1, st);
      if (rc == CSM_OK) rc = frame_tail_tc(x, B, st);
      if (rc != CSM_OK) return rc;
    } else {
      CU_TRY(backbone_pass(x, B, 1, st));
      CU_TRY(frame_tail(x, B, st));
    }
  } else {
    cudaGraphExec_t ge;
    int rc = get_graph(x, B, &ge);
    if (rc != CSM_OK) return rc;
    CU_TRY(cudaGraphLaunch(ge, st));
    g_launches.fetch_add(x->graph_nodes[B | (x->long_ctx ? 1 << 24 : 0)], std::memory_order_relaxed);
  }
  advance();
  return CSM_OK;
}

extern "C" int32_t csm_debug_phase_table(const csm_config* cfg, int32_t n_ctas, int32_t with_qkv_table, csm_phase_info* out,
                                         int32_t max_phases) {
  if (!valid_cfg(cfg) || n_ctas < 1 || !out) return set_err(CSM_ERR_ARG, "bad phase_table arguments");
  csm_ctx x;
  x.cfg = *cfg;
  x.max_batch = 1;
  char* const base = reinterpret_cast<char*>((uintptr_t)1 << 32);  // imaginary workspace: pointers are only compared
  carve_all(&x, base);
  for (StackDev* s : {&x.bb, &x.dec}) {
    s->wo.assign(s->c.layers, nullptr); s->wd.assign(s->c.layers, nullptr);
    s->sa.assign(s->c.layers, reinterpret_cast<const bf16*>(base)); s->mlp.assign(s->c.layers, reinterpret_cast<const bf16*>(base));
    s->norm = reinterpret_cast<const bf16*>(base); s->rope = nullptr;
  }
  auto pick_stack = [&](const StackDev& s, int* r4) {
    r4[0] = mega_pick_R((s.c.heads + 2 * s.c.kv_heads) * s.hd, n_ctas);
    r4[1] = mega_pick_R(s.c.dim, n_ctas);
    r4[2] = mega_pick_R(2 * s.c.ff, n_ctas);
    r4[3] = mega_pick_R(s.c.dim, n_ctas);
  };
  pick_stack(x.bb, x.mega_Rbb);
  pick_stack(x.dec, x.mega_Rdec);
  x.mega_Rh0 = mega_pick_R(x.Vf + cfg->decoder.dim, n_ctas);
  x.mega_Rh = mega_pick_R(x.Vf, n_ctas);
  x.mega_keep = 0;
  x.bb.rope_len = x.dec.rope_len = 0;
  x.qkv_table_ok = with_qkv_table != 0;
  x.text_emb = x.audio_emb = nullptr;
  MegaBuild mb;
  mb.ncta = n_ctas;
  mb.rot = 0;
  build_mega_phases(&x, mb);
  const int n = (int)mb.v.size();
  if (n != mega_phase_count(*cfg, x.qkv_table_ok)) return set_err(CSM_ERR_ARG, "internal: phase count mismatch");
  auto off = [&](const void* p) { return p ? (uint64_t)(reinterpret_cast<const char*>(p) - base) : (uint64_t)0; };
  for (int i = 0; i < n && i < max_phases; ++i) {
    const mega::Phase& ph = mb.v[i];
    csm_phase_info& o = out[i];
    memset(&o, 0, sizeof(o));
    o.type = ph.type; o.epi = ph.epi; o.nb = ph.nb; o.K = ph.K; o.rows = ph.rows; o.R = ph.R; o.G = ph.G; o.rot = ph.rot;
    o.ldx = ph.ldx; o.ldo = ph.ldo; o.split_row = ph.split_row; o.attn_prologue = ph.attn_prologue; o.has_qkv_table = ph.qkv_table != nullptr;
    o.x_src[0] = ph.x_src[0]; o.x_src[1] = ph.x_src[1]; o.resid_src[0] = ph.resid_src[0]; o.resid_src[1] = ph.resid_src[1];
    o.q_src = ph.q_src; o.logits_src = ph.logits_src;
    o.t_x = off(ph.t_x); o.t_out = off(ph.t_out); o.t_out2 = off(ph.t_out2); o.t_q = off(ph.t_q); o.t_kv = off(ph.t_kv);
    o.t_logits = off(ph.t_logits); o.t_next = off(ph.t_next);
    o.kv_sync = ph.kv_sync & 255; o.done_src = ph.done_src; o.pos0 = ph.pos0; o.pos_mode = ph.pos_mode;
  }
  return n;
}

extern "C" int32_t csm_debug_set_trace(csm_ctx* x, void* dev_buffer) {
  if (!x) return set_err(CSM_ERR_ARG, "null ctx");
  x->trace = (unsigned long long*)dev_buffer;
  return x->mega_ok ? x->n_phases : 0;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 GEMM launcher: Y[rows, n_out] = X[rows, K] . W[n_out, K]^T  (gemm_tc.cuh)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (encode_tiled_fn)p;
  }
  return fn;
}
static int make_map_bf16(CUtensorMap* m, const bf16* base, long long rows, long long K, long long ld, int box_rows) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) return set_err(CSM_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(CSM_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return CSM_OK;
}
// rk / rk_fused: the caller's RoPE + KV-append step; *rk_fused says whether this launch did it in its epilogue (only
// the cluster split-K path can) -- otherwise the projection is in ``out`` and the caller runs k_rope_kv_rows
static int launch_gemm_tc(const bf16* X, long long ldx, int rows, int K, const bf16* W, int n_out, bf16* out, long long ldo,
                          int epi, const bf16* resid, cudaStream_t st, const csm_ctx* splitk, const tc::RopeKV* rk, bool* rk_fused) {
  if (rk_fused) *rk_fused = false;
  if (K % tc::BK || rows < 1 || n_out < 1) return set_err(CSM_ERR_ARG, "gemm_tc: K must be a multiple of 64");
  static std::atomic<unsigned long long> attr{0};
  if (!device_done(attr, false)) {
    CU_TRY(cudaFuncSetAttribute(tc::k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES));
    device_done(attr, true);
  }
  CUtensorMap mx, mw;
  int rc;
  if ((rc = make_map_bf16(&mx, X, rows, K, ldx, tc::BM)) != CSM_OK) return rc;
  if ((rc = make_map_bf16(&mw, W, n_out, K, K, tc::BN)) != CSM_OK) return rc;
  tc::Args a;
  a.out = out; a.ldo = ldo; a.resid = resid ? resid : out; a.rows = rows; a.n_out = n_out; a.K = K; a.epi = epi;
  a.part = nullptr; a.counters = nullptr; a.ldp = 0; a.cluster = 0;
  dim3 grid((n_out + tc::BN - 1) / tc::BN, (rows + tc::BM - 1) / tc::BM);
  // decode steps of large batches: few tiles, long K -> split K over the idle SMs (see gemm_tc.cuh)
  const int tiles = (int)(grid.x * grid.y), num_kb = K / tc::BK;
  // (only the long-K projections: with K = 1024 .. 2048 a tile is 16 .. 32 k blocks and the second pass costs more than it saves)
  // Measured on B200 (profiles/r2_splitk.txt): the decode step of 64 / 128 / 256 streams takes 13.4 / 15.9 / 19.8 ms
  // with split-K and 13.3 / 15.5 / 19.1 ms without -- the step is bound by its ~1200 dependent launches, not by the
  // 16-CTA projections -- so it is OFF unless CSM_TC_SPLITK=1 (the unit-test entry always enables it).
  static const bool splitk_on = getenv("CSM_TC_SPLITK") != nullptr;
  if (splitk && (splitk_on || splitk->max_batch == 0) && rows <= TC_SPLIT_MAX_ROWS && 2 * tiles <= TC_SPLIT_TILES && num_kb >= 64) {
    int splits = 1;
    while (splits * 2 <= 8 && splits * 2 * tiles <= TC_SPLIT_TILES && num_kb % (splits * 2) == 0 && num_kb / (splits * 2) >= 2)
      splits *= 2;
    if (splits > 1) {
      grid.z = splits;
      a.part = splitk->tc_part; a.counters = splitk->tc_counters; a.ldp = (long long)grid.x * tc::BN;
    }
  }
  // Cluster split-K (default for decode-sized GEMMs): a projection of a 256-stream decode step is 16 .. 34 tiles of
  // 16 .. 128 k blocks -- a handful of CTAs, each a chain of dependent TMA round trips (the 16-tile K = 8192 down
  // projection took 62 us, profiles/r2_decode_B256_shapes.txt).  The k blocks of a tile are split over a thread-block
  // cluster of 2 / 4 / 8 CTAs that add their partial tiles up through distributed shared memory.
  static const bool cluster_on = !(getenv("CSM_TC_CLUSTER") && getenv("CSM_TC_CLUSTER")[0] == '0');
  int cluster_z = 1;
  // (K = 1024 projections, 16 k blocks: under ncu the cluster's fixed cost cancels the shorter k loop, in the replayed
  //  graph splitting them too is 2-4 % faster per step at 128 / 256 streams: profiles/r2_cluster_splitk.txt)
  static const int cluster_min_kb = getenv("CSM_TC_CLUSTER_MIN_KB") ? atoi(getenv("CSM_TC_CLUSTER_MIN_KB")) : 16;
  if (cluster_on && grid.z == 1 && rows <= TC_SPLIT_MAX_ROWS && num_kb >= cluster_min_kb) {
    int sp = 8;
    while (sp > 1 && (sp * tiles > TC_SPLIT_TILES || num_kb % sp || num_kb / sp < 4)) sp >>= 1;
    if (sp > 1) {
      cluster_z = sp;
      grid.z = sp;
      a.cluster = 1;
    }
  }
  static const bool old_kernel = getenv("CSM_TC_ONE_TILE") != nullptr;  // measurement aid: round 1's one-tile-per-CTA kernel
  static std::atomic<int> nsm_cached{0};
  if (!nsm_cached.load(std::memory_order_relaxed)) {
    int dev = 0, n = 0;
    CU_TRY(cudaGetDevice(&dev));
    CU_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    nsm_cached.store(n > 0 ? n : 148, std::memory_order_relaxed);
  }
  const int nsm = nsm_cached.load(std::memory_order_relaxed);
  // (Tried for the 256-tile gate/up of a 256-stream step: one CTA taking two stacked tiles against one W box per
  //  k block, 128 CTAs x 48 KB per k block -- correct, and slower than the persistent kernel: 14.4 vs 12.8 ms per step.)
  // RoPE + KV append of the [q;k;v] projection in the GEMM epilogue (tc::EPI_ROPE_KV: the persistent kernel and the
  // cluster split-K reduce both walk the tile four columns per thread).  Correct (the GPU suite passes with it on) and
  // measured SLOWER than the separate k_rope_kv_rows launch in both places -- 256-stream decode step 12.7 vs 12.1 ms,
  // 32 x 1568-frame prefill 173.6 vs 169.2 ms: scattered cache stores and index arithmetic inside the GEMM's epilogue
  // versus a cheap, perfectly parallel kernel hidden by the dependent launch.  Off unless CSM_TC_ROPE_FUSE=1 (persistent
  // kernel) / 2 (both).
  static const int rope_fuse = getenv("CSM_TC_ROPE_FUSE") ? atoi(getenv("CSM_TC_ROPE_FUSE")) : 0;
  if (rk && rk_fused && rope_fuse > 0 && epi == tc::EPI_STORE && n_out % 4 == 0 && !old_kernel && a.part == nullptr &&
      ((a.cluster && rope_fuse > 1) || (!a.cluster && tiles > nsm))) {
    a.rk = *rk;
    a.epi = tc::EPI_ROPE_KV;
    *rk_fused = true;
  }
  // (one tile per CTA also when every tile gets its own SM: its ring is six stages deep, the persistent kernel's four)
  if (grid.z > 1 || old_kernel || tiles <= nsm) {
    launch_kc(tc::k_gemm_tc, dim3(grid), dim3(tc::THREADS), tc::SMEM_BYTES, st, cluster_z, mx, mw, a); COUNT_LAUNCH();
  } else {
    // persistent kernel: one CTA per SM walks the tiles; 128 x 256 tiles when the output is wide enough (CSM_TC_BN=128: off)
    static std::atomic<unsigned long long> attr_p{0};
    if (!device_done(attr_p, false)) {
      CU_TRY(cudaFuncSetAttribute(tc::k_gemm_tc_p<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::PCfg<128>::SMEM));
      CU_TRY(cudaFuncSetAttribute(tc::k_gemm_tc_p<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::PCfg<256>::SMEM));
      device_done(attr_p, true);
    }
    static const bool wide_on = !(getenv("CSM_TC_BN") && atoi(getenv("CSM_TC_BN")) == 128);
    // (prompt-sized row counts only: for the 256-tile gate/up of a 256-stream decode step the narrow tile is 1 % faster)
    if (wide_on && n_out >= 512 && rows > TC_SPLIT_MAX_ROWS) {
      if ((rc = make_map_bf16(&mw, W, n_out, K, K, 256)) != CSM_OK) return rc;
      const int tiles_w = (int)(((n_out + 255) / 256) * grid.y);
      launch_k(tc::k_gemm_tc_p<256>, dim3(tiles_w < nsm ? tiles_w : nsm), dim3(tc::P_THREADS), tc::PCfg<256>::SMEM, st, mx, mw, a); COUNT_LAUNCH();
    } else {
      launch_k(tc::k_gemm_tc_p<128>, dim3(tiles < nsm ? tiles : nsm), dim3(tc::P_THREADS), tc::PCfg<128>::SMEM, st, mx, mw, a); COUNT_LAUNCH();
    }
  }
  CU_TRY(cudaGetLastError());
  return CSM_OK;
}

// Batched decode (<= 64 rows): one CTA per fragment-major row group (skinny.cuh).  n_out = valid rows.
template <bool NORM>
static void launch_skinny_t(const sk::Args& a, cudaStream_t st) {
  if (a.N <= 8) launch_k(sk::k_skinny<1, NORM>, dim3(a.G), dim3(256), 0, st, a);
  else if (a.N <= 16) launch_k(sk::k_skinny<2, NORM>, dim3(a.G), dim3(256), 0, st, a);
  else if (a.N <= 32) launch_k(sk::k_skinny<4, NORM>, dim3(a.G), dim3(256), 0, st, a);
  else launch_k(sk::k_skinny<8, NORM>, dim3(a.G), dim3(256), 0, st, a);
}
// norm_scale != null: X holds the un-normalised rows and the kernel applies torchtune's RMSNorm on the fly
static int launch_skinny(const bf16* Wf, int R, const bf16* X, long long ldx, int N, int K, int n_out, bf16* out, long long ldo,
                         int epi, const bf16* resid, cudaStream_t st, const bf16* norm_scale = nullptr, float eps = 0.f,
                         const sk::Args* rope_kv = nullptr) {
  sk::Args a;
  memset(&a, 0, sizeof(a));
  if (rope_kv) a = *rope_kv;  // the RoPE / KV-append fields of the fused [q;k;v] projection
  a.Wf = Wf; a.R = R; a.K = K; a.n_out = n_out; a.G = (n_out + R - 1) / R; a.X = X; a.ldx = ldx; a.N = N; a.out = out; a.ldo = ldo;
  a.resid = resid ? resid : out; a.epi = epi; a.norm_scale = norm_scale; a.eps = eps;
  if (norm_scale) launch_skinny_t<true>(a, st);
  else launch_skinny_t<false>(a, st);
  COUNT_LAUNCH();
  CU_TRY(cudaGetLastError());
  return CSM_OK;
}
// a linear layer of the row-batched path: skinny kernel for few rows, tcgen05 GEMM otherwise
static bool skinny_ok(const csm_ctx* x, const bf16* Wf, int N, int K) {
  return x->frag_ok && Wf && N <= SKINNY_MAX_ROWS && K % 256 == 0;
}
static int linear_rows(csm_ctx* x, const bf16* W, const bf16* Wf, int R, const bf16* X, long long ldx, int N, int K, int n_out,
                       bf16* out, long long ldo, int epi, const bf16* resid, cudaStream_t st, const tc::RopeKV* rk = nullptr,
                       bool* rk_fused = nullptr) {
  if (rk_fused) *rk_fused = false;
  if (skinny_ok(x, Wf, N, K)) return launch_skinny(Wf, R, X, ldx, N, K, n_out, out, ldo, epi, resid, st);
  return launch_gemm_tc(X, ldx, N, K, W, n_out, out, ldo, epi, resid, st, x, rk, rk_fused);
}
// RMSNorm of N rows: a warp per row when the row fits in registers (D = 1024 / 2048, 16-byte aligned), else a CTA per row
static void launch_rmsnorm_rows(const bf16* x, int D, const bf16* scale, float eps, bf16* y, int N, cudaStream_t st) {
  const bool vec = ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)scale) & 15) == 0);
  if (vec && D == 1024) launch_k(k_rmsnorm_rows<4>, dim3((N + 7) / 8), dim3(256), 0, st, x, D, scale, eps, y, D, N);
  else if (vec && D == 2048) launch_k(k_rmsnorm_rows<8>, dim3((N + 7) / 8), dim3(256), 0, st, x, D, scale, eps, y, D, N);
  else launch_k(k_rmsnorm, dim3(N), dim3(256), 0, st, x, D, scale, D, eps, y, D);
  COUNT_LAUNCH();
}
// rows up to which the skinny kernel normalises its activation rows itself (CSM_SKINNY_NORM_ROWS: measurement aid)
static int skinny_norm_rows() {
  static const int n = [] {
    const char* e = getenv("CSM_SKINNY_NORM_ROWS");
    const int v = e ? atoi(e) : 16;
    return v < 0 ? 0 : (v > 64 ? 64 : v);
  }();
  return n;
}
// RMSNorm(H rows) followed by a linear layer: fused into the skinny kernel for few rows, else k_rmsnorm into
// ``xn`` (skipped when ``xn_ready``: an earlier call of the same pair normalised already) and the tcgen05 GEMM
static int norm_linear_rows(csm_ctx* x, const bf16* H, const bf16* scale, float eps, bf16* xn, bool xn_ready, const bf16* W,
                            const bf16* Wf, int R, int N, int K, int n_out, bf16* out, long long ldo, int epi, cudaStream_t st,
                            const tc::RopeKV* rk = nullptr, bool* rk_fused = nullptr) {
  if (rk_fused) *rk_fused = false;
  // (every CTA normalises all rows itself: cheaper than a launch up to 16 rows, measured slower at 32)
  if (skinny_ok(x, Wf, N, K) && N <= skinny_norm_rows()) return launch_skinny(Wf, R, H, K, N, K, n_out, out, ldo, epi, nullptr, st, scale, eps);
  if (!xn_ready) {
    launch_rmsnorm_rows(H, K, scale, eps, xn, N, st);
  }
  return linear_rows(x, W, Wf, R, xn, K, N, K, n_out, out, ldo, epi, nullptr, st, rk, rk_fused);
}

// All layers of one stack on N rows with the tcgen05 GEMM: per layer RMSNorm -> GEMM [q;k;v] -> RoPE +
// KV append -> attention -> GEMM O (+res) -> RMSNorm -> GEMM gate/up (SwiGLU epilogue) -> GEMM down
// (+res); same rounding points as the small-row path.
static int stack_pass_tc(csm_ctx* x, StackDev& s, int N, const RowMeta& m, cudaStream_t st) {
  const csm_stack_config& k = s.c;
  const int D = k.dim, qkv_cols = (k.heads + 2 * k.kv_heads) * s.hd;
  const float eps = x->cfg.norm_eps;
  const int* R4 = (&s == &x->bb) ? x->mega_Rbb : x->mega_Rdec;  // row-group heights of the fragment-major copies
  int rc;
  for (int l = 0; l < k.layers; ++l) {
    bf16* kc = s.kc + s.kv_layer_stride * l;
    bf16* vc = s.vc + s.kv_layer_stride * l;
    if (skinny_ok(x, s.fqkv[l], N, D)) {
      // few rows: RMSNorm (up to 16 rows), [q;k;v], RoPE and the KV append in ONE skinny launch
      sk::Args ra;
      memset(&ra, 0, sizeof(ra));
      ra.rope = s.rope; ra.row_stream = m.stream; ra.row_pos = m.pos; ra.row_slot = m.slot; ra.imp_B = m.imp_B; ra.imp_pos = m.imp_pos;
      ra.heads = k.heads; ra.kv_heads = k.kv_heads; ra.hd = s.hd; ra.slots = s.slots; ra.k_cache = kc; ra.v_cache = vc;
      const bool fuse_norm = N <= skinny_norm_rows();
      if (!fuse_norm) {
        launch_rmsnorm_rows(s.h, D, s.sa[l], eps, s.xn, N, st);
      }
      if ((rc = launch_skinny(s.fqkv[l], R4[0], fuse_norm ? s.h : s.xn, D, N, D, qkv_cols, s.q, 0, sk::EPI_ROPE_KV, nullptr, st,
                              fuse_norm ? s.sa[l] : nullptr, eps, &ra)) != CSM_OK) return rc;
    } else {
      tc::RopeKV rk;
      rk.rope = s.rope; rk.row_stream = m.stream; rk.row_pos = m.pos; rk.row_slot = m.slot; rk.imp_B = m.imp_B; rk.imp_pos = m.imp_pos;
      rk.heads = k.heads; rk.kv_heads = k.kv_heads; rk.hd = s.hd; rk.slots = s.slots; rk.q_out = s.q; rk.k_cache = kc; rk.v_cache = vc;
      bool fused = false;  // decode-sized row counts: RoPE + KV append run in the projection's cluster split-K epilogue
      if ((rc = norm_linear_rows(x, s.h, s.sa[l], eps, s.xn, false, s.wqkv[l], s.fqkv[l], R4[0], N, D, qkv_cols, s.qkv, qkv_cols,
                                 tc::EPI_STORE, st, &rk, &fused)) != CSM_OK) return rc;
      if (!fused) {
        launch_k(k_rope_kv_rows, dim3(N), dim3(256), 0, st, s.qkv, s.rope, m.stream, m.pos, m.slot, m.imp_B, m.imp_pos, k.heads, k.kv_heads, s.hd,
                 s.slots, s.q, kc, vc); COUNT_LAUNCH();
      }
    }
    {
      dim3 grid(N, k.heads);
      const size_t smem = (size_t)s.slots * sizeof(float);
      const float scale = 1.0f / sqrtf((float)s.hd);
      if (s.hd == 64 && m.chunk >= 16 && m.stream) {
        // prompt rows: tiled tensor-core attention, one CTA per (64 rows of a stream, q-head)
        dim3 fgrid((m.chunk + 63) / 64, k.heads, N / m.chunk);
        launch_k(k_attn_flash64, dim3(fgrid), dim3(128), 0, st, s.q, kc, vc, m.slot, m.chunk, k.heads, k.kv_heads, s.slots, scale, s.att);
      } else if (s.hd == 64 && x->long_ctx && &s == &x->bb && m.chunk <= 1 && N <= x->max_batch && k.heads % k.kv_heads == 0 &&
                 k.heads / k.kv_heads <= 8) {
        // decode rows at a long context: key ranges over CTAs (row, KV head, range), then the ranges are combined
        static std::atomic<unsigned long long> attr_am{0};
        if (!device_done(attr_am, false)) {
          CU_TRY(cudaFuncSetAttribute(k_attn_split64_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AM_SMEM));
          device_done(attr_am, true);
        }
        launch_k(k_attn_split64_mma, dim3(N, k.kv_heads, AS_SPLITS), dim3(128), AM_SMEM, st, s.q, kc, vc, m.stream, m.slot, m.imp_B,
                 m.imp_pos, k.heads, k.kv_heads, s.slots, scale, x->att_part);
        COUNT_LAUNCH();
        launch_k(k_attn_combine64, dim3(N, k.heads), dim3(64), 0, st, x->att_part, k.heads, s.att);
      } else if (s.hd == 64) {
        launch_k(k_attn_rows<64>, dim3(grid), dim3(128), smem, st, s.q, kc, vc, m.stream, m.slot, m.imp_B, m.imp_pos, k.heads, k.kv_heads, s.slots,
                                                 scale, s.att);
      } else if (s.hd == 128 && s.slots <= 32 && k.heads % k.kv_heads == 0 && k.heads / k.kv_heads <= 8) {
        // depth decoder: one CTA per (row, KV head), its q-heads share the staged K / V rows
        launch_k(k_attn_dec, dim3(N, k.kv_heads), dim3(32 * (k.heads / k.kv_heads)), 0, st, s.q, kc, vc, m.stream, m.slot, m.imp_B,
                 m.imp_pos, k.heads, k.kv_heads, s.slots, scale, s.att);
      } else {
        launch_k(k_attn_rows<128>, dim3(grid), dim3(128), smem, st, s.q, kc, vc, m.stream, m.slot, m.imp_B, m.imp_pos, k.heads, k.kv_heads, s.slots,
                                                  scale, s.att);
      }
      COUNT_LAUNCH();
    }
    CU_TRY(cudaGetLastError());
    if ((rc = linear_rows(x, s.wo[l], s.fo[l], R4[1], s.att, D, N, D, D, s.h, D, tc::EPI_ADD_RESID, s.h, st)) != CSM_OK) return rc;
    if ((rc = norm_linear_rows(x, s.h, s.mlp[l], eps, s.xn, false, s.wgu[l], s.fgu[l], R4[2], N, D, 2 * k.ff, s.act, k.ff,
                               tc::EPI_SWIGLU_PAIRS, st)) != CSM_OK) return rc;
    if ((rc = linear_rows(x, s.wd[l], s.fd[l], R4[3], s.act, k.ff, N, k.ff, D, s.h, D, tc::EPI_ADD_RESID, s.h, st)) != CSM_OK) return rc;
  }
  CU_TRY(cudaGetLastError());
  return CSM_OK;
}

// Backbone pass over `chunk` frames per stream on the tensor cores (rows n = b*chunk + t).
static int backbone_pass_tc(csm_ctx* x, int B, int chunk, cudaStream_t st) {
  const csm_config& c = x->cfg;
  const int N = B * chunk;
  launch_k(k_embed_pass, dim3(N), dim3(256), 0, st, x->d_params, x->text_emb, x->audio_emb, c.codebooks, c.audio_vocab, c.backbone.dim, chunk,
                                  x->bb.h, x->row_stream, x->row_pos, x->row_slot, c.text_vocab, x->bb.rope_len, x->d_sync);
  COUNT_LAUNCH();
  CU_TRY(cudaGetLastError());
  RowMeta m{x->row_stream, x->row_pos, x->row_slot, 0, 0, chunk};
  return stack_pass_tc(x, x->bb, N, m, st);
}

// frame_tail for many streams (B >= DECODE_TC_MIN): every linear layer is a tcgen05 GEMM over the B
// (or 2B) rows, the depth decoder's inputs after step 1 come from the projection(embedding) table.
static int frame_tail_tc(csm_ctx* x, int B, cudaStream_t st) {
  const csm_config& c = x->cfg;
  const int D = c.backbone.dim, Dd = c.decoder.dim, V = c.audio_vocab, C = c.codebooks;
  int rc;
  // last_h = backbone.norm(h) feeds the codebook-0 head and the projection (the fragment-major copy stacks
  // [codebook0_head padded to Vf rows ; projection]: two row-group ranges of it)
  if ((rc = norm_linear_rows(x, x->bb.h, x->bb.norm, c.norm_eps, x->dec_in, false, x->c0_head, x->f_head0, x->mega_Rh0, B, D, V,
                             x->logits, x->Vp, tc::EPI_STORE, st)) != CSM_OK) return rc;
  if ((rc = norm_linear_rows(x, x->bb.h, x->bb.norm, c.norm_eps, x->dec_in, true, x->proj, x->f_head0 + (size_t)x->Vf * D, x->mega_Rh0, B,
                             D, Dd, x->dec.h, Dd, tc::EPI_STORE, st)) != CSM_OK) return rc;
  launch_k(k_sample_step, dim3(B), dim3(SAMPLE_THREADS), 0, st, x->d_params, x->logits, x->Vp, 0, V, C, x->proj_table, Dd,
                                              x->dec.h + (size_t)B * Dd, x->d_sync); COUNT_LAUNCH();
  CU_TRY(cudaGetLastError());
  for (int i = 1; i < C; ++i) {
    const int N = (i == 1) ? 2 * B : B, pos0 = (i == 1) ? 0 : i;
    RowMeta m{nullptr, nullptr, nullptr, B, pos0};
    if ((rc = stack_pass_tc(x, x->dec, N, m, st)) != CSM_OK) return rc;
    if ((rc = norm_linear_rows(x, x->dec.h + (size_t)(N - B) * Dd, x->dec.norm, c.norm_eps, x->dec.xn, false,
                               x->head_t + (size_t)(i - 1) * x->Vp * Dd, x->f_heads + (size_t)(i - 1) * x->Vf * Dd, x->mega_Rh, B, Dd, V,
                               x->logits, x->Vp, tc::EPI_STORE, st)) != CSM_OK) return rc;
    launch_k(k_sample_step, dim3(B), dim3(SAMPLE_THREADS), 0, st, x->d_params, x->logits, x->Vp, i, V, C, x->proj_table, Dd,
                                                (i + 1 < C) ? x->dec.h : nullptr, x->d_sync); COUNT_LAUNCH();
    CU_TRY(cudaGetLastError());
  }
  return CSM_OK;
}

extern "C" int32_t csm_k_gemm_tc(const void* xin, const void* W, int32_t N, int32_t in, int32_t outf, void* y, int32_t epi,
                                 const void* resid, void* stream) {
  if (!xin || !W || !y || N < 1 || in % 64 || outf < 1 || epi < 0 || epi > 2) return set_err(CSM_ERR_ARG, "bad gemm_tc arguments");
  const long long ldo = epi == tc::EPI_SWIGLU_PAIRS ? outf / 2 : outf;
  return launch_gemm_tc((const bf16*)xin, in, N, in, (const bf16*)W, outf, (bf16*)y, ldo, epi, (const bf16*)resid,
                        (cudaStream_t)stream);
}

extern "C" int32_t csm_k_gemm_tc_splitk(const void* xin, const void* W, int32_t N, int32_t in, int32_t outf, void* y, int32_t epi,
                                        const void* resid, void* part, void* counters, void* stream) {
  if (!xin || !W || !y || !part || !counters || N < 1 || in % 64 || outf < 1 || epi < 0 || epi > 2)
    return set_err(CSM_ERR_ARG, "bad gemm_tc arguments");
  csm_ctx tmp;
  tmp.max_batch = 0;  // marks the unit-test context: split-K always on
  tmp.tc_part = (float*)part;
  tmp.tc_counters = (unsigned int*)counters;
  const long long ldo = epi == tc::EPI_SWIGLU_PAIRS ? outf / 2 : outf;
  return launch_gemm_tc((const bf16*)xin, in, N, in, (const bf16*)W, outf, (bf16*)y, ldo, epi, (const bf16*)resid,
                        (cudaStream_t)stream, &tmp);
}

extern "C" int32_t csm_k_attn_prefill(const void* q, const void* k_cache, const void* v_cache, const int32_t* row_slot, int32_t B,
                                      int32_t chunk, int32_t heads, int32_t kv_heads, int32_t slots, void* out, void* stream) {
  if (!q || !k_cache || !v_cache || !row_slot || !out || B < 1 || chunk < 1 || heads < 1 || kv_heads < 1 || heads % kv_heads ||
      slots < 1)
    return set_err(CSM_ERR_ARG, "bad attn_prefill arguments");
  dim3 grid((chunk + 63) / 64, heads, B);
  launch_k(k_attn_flash64, dim3(grid), dim3(128), 0, (cudaStream_t)stream, (const bf16*)q, (const bf16*)k_cache, (const bf16*)v_cache, row_slot, chunk,
                                                         heads, kv_heads, slots, 0.125f, (bf16*)out);
  COUNT_LAUNCH();
  CU_TRY(cudaGetLastError());
  return CSM_OK;
}

// ---- unit-test entry points -----------------------------------------------------------------
extern "C" int32_t csm_k_sample_topk(const void* logits, const void* noise, int32_t B, int32_t V, float temperature,
                                     int32_t topk, int32_t* out, void* stream) {
  if (!logits || !out || B < 1 || V < 1 || V > SAMPLE_MAXV) return set_err(CSM_ERR_ARG, "bad sample_topk arguments");
  k_sample_only<<<B, SAMPLE_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)logits, (const bf16*)noise, V, temperature,
                                                                topk, out); COUNT_LAUNCH();
  CU_TRY(cudaGetLastError());
  return CSM_OK;
}

extern "C" int32_t csm_k_embed_frames(const int64_t* tokens, const uint8_t* mask, const void* text_emb,
                                      const void* audio_emb, int32_t N, int32_t codebooks, int32_t audio_vocab, int32_t D,
                                      void* out, void* stream) {
  if (!tokens || !mask || !text_emb || !audio_emb || !out || N < 1 || D % 8) return set_err(CSM_ERR_ARG, "bad embed arguments");
  k_embed_frames<<<N, 256, 0, (cudaStream_t)stream>>>(tokens, mask, (const bf16*)text_emb, (const bf16*)audio_emb,
                                                      codebooks, audio_vocab, D, (bf16*)out); COUNT_LAUNCH();
  CU_TRY(cudaGetLastError());
  return CSM_OK;
}

extern "C" int32_t csm_k_linear(const void* xin, const void* W, int32_t N, int32_t in, int32_t outf, void* y,
                                void* stream) {
  if (!xin || !W || !y || N < 1 || in % 256 || in > 8192 || outf < 1) return set_err(CSM_ERR_ARG, "bad linear arguments");
  CU_TRY(init_kernel_attrs());
  GemvArgs a;
  memset(&a, 0, sizeof(a));
  a.W = (const bf16*)W; a.rows = outf; a.K = in; a.x = (const bf16*)xin; a.ldx = in; a.N = N;
  a.out = (bf16*)y; a.ldo = outf;
  CU_TRY((launch_gemv<EPI_PLAIN, false>(a, (cudaStream_t)stream)));
  return CSM_OK;
}

extern "C" int32_t csm_k_rmsnorm(const void* xin, const void* scale, int32_t N, int32_t D, float eps, void* y,
                                 void* stream) {
  if (!xin || !scale || !y || N < 1 || D < 1) return set_err(CSM_ERR_ARG, "bad rmsnorm arguments");
  launch_k(k_rmsnorm, dim3(N), dim3(256), 0, (cudaStream_t)stream, (const bf16*)xin, D, (const bf16*)scale, D, eps, (bf16*)y, D); COUNT_LAUNCH();
  CU_TRY(cudaGetLastError());
  return CSM_OK;
}
