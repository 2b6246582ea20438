// C ABI of libfitclip_b200 (include/fitclip_b200.h): model handle, weight loading, tower assembly, similarity/rank.
#include <algorithm>
#include <atomic>
#include <stdarg.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/fitclip_b200.h"
#include "gemm.cuh"
#include "kernels.cuh"

namespace fc {

// ---------------------------------------------------------------- errors / bookkeeping
static thread_local char g_err[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in `%s`", static_cast<int>(e), cudaGetErrorString(e), file, line, what);
  return FC_ERR_CUDA;
}
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---------------------------------------------------------------- profiler
struct ProfRec {
  cudaEvent_t e0, e1;
  int kind, tag;
  int64_t m, n, k;
  double flops, bytes;
};
static std::vector<ProfRec> g_prof;
static int g_prof_used = 0;
static bool g_prof_on = false;

ProfScope::ProfScope(cudaStream_t s, int kind, int tag, int64_t m, int64_t n, int64_t k, double flops, double bytes)
    : slot(-1), stream(s) {
  if (!g_prof_on || g_prof_used >= static_cast<int>(g_prof.size())) return;
  slot = g_prof_used++;
  ProfRec& r = g_prof[slot];
  r.kind = kind; r.tag = tag; r.m = m; r.n = n; r.k = k; r.flops = flops; r.bytes = bytes;
  cudaEventRecord(r.e0, stream);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].e1, stream);
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int check_arch() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    FC_CUDA(cudaGetDevice(&dev));
    FC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    ok = (major == 10) ? 1 : 0;
  }
  if (!ok) {
    set_error("libfitclip_b200 needs a compute-capability 10.x (sm_100a) device; there is no fallback path");
    return FC_ERR_ARCH;
  }
  return FC_OK;
}

// ---------------------------------------------------------------- model
struct Block {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *qkv_b, *out_b, *fc_b, *proj_b;
  bf16 *out_w, *proj_w;
  // ln_1 is folded into in_proj and ln_2 into c_fc (fold_ln_weights): fp32 masters as loaded, then the folded bf16
  // weight, its per-output column sums and the folded bias, rebuilt by finalize() whenever a parameter changes.
  float *qkv_w32, *fc_w32;
  bf16 *qkv_w, *fc_w;
  float *qkv_cs, *qkv_bf, *fc_cs, *fc_bf;
};

struct Slot {
  std::string name;
  void* dst;
  int64_t numel;
  bool as_bf16;
  bool loaded;
  int cols = 0, ld = 0;  // bf16 copies stored with a padded row pitch (cols of the source, ld of the copy); 0 = dense
};

}  // namespace fc

struct fc_model {
  fc_config cfg;
  int device = 0;
  // geometry
  int L_img = 0, grid = 0, patch_dim = 0;  // patch_dim: 3*P*P rounded up to a multiple of 8 (row pitch of the patch matrix)
  int patch_cols = 0;                      // 3*P*P
  int maxF = 0, maxC = 0;
  // arena
  uint8_t* arena = nullptr;
  int64_t arena_bytes = 0;
  std::vector<fc::Slot> slots;
  // vision params
  fc::bf16* conv_w = nullptr;
  float *cls = nullptr, *vpos = nullptr, *ln_pre_g = nullptr, *ln_pre_b = nullptr, *ln_post_g = nullptr,
        *ln_post_b = nullptr, *vproj = nullptr;
  std::vector<fc::Block> vblocks;
  // text params
  float *tok = nullptr, *tpos = nullptr, *ln_final_g = nullptr, *ln_final_b = nullptr, *tproj = nullptr;
  std::vector<fc::Block> tblocks;
  bool dirty = true;  // folded weights need (re)building
  // workspace
  fc::bf16 *x = nullptr, *y = nullptr, *big = nullptr;
  float *feat = nullptr, *stats_a = nullptr, *stats_b = nullptr;
  int* err_flag = nullptr;
  int64_t workspace_bytes = 0;
};

namespace fc {

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct ArenaPlan {
  int64_t off = 0;
  int64_t take(int64_t bytes) {
    const int64_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  }
};

// Wa: width of the attention (3 Wa rows of in_proj, Wa columns of out_proj); == W except for padded narrow heads
static void plan_blocks(fc_model* m, ArenaPlan& plan, std::vector<Block>& blocks, const std::string& prefix, int layers,
                        int W, int Wa, bool assign) {
  blocks.resize(layers);
  for (int i = 0; i < layers; ++i) {
    Block& b = blocks[i];
    const std::string p = prefix + "resblocks." + std::to_string(i) + ".";
    struct Item {
      const char* name;
      void** dst;
      int64_t numel;
      bool bf;
    } items[] = {
        {"attn.in_proj_weight", reinterpret_cast<void**>(&b.qkv_w32), int64_t(3) * Wa * W, false},
        {"attn.in_proj_bias", reinterpret_cast<void**>(&b.qkv_b), int64_t(3) * Wa, false},
        {"attn.out_proj.weight", reinterpret_cast<void**>(&b.out_w), int64_t(W) * Wa, true},
        {"attn.out_proj.bias", reinterpret_cast<void**>(&b.out_b), W, false},
        {"ln_1.weight", reinterpret_cast<void**>(&b.ln1_g), W, false},
        {"ln_1.bias", reinterpret_cast<void**>(&b.ln1_b), W, false},
        {"ln_2.weight", reinterpret_cast<void**>(&b.ln2_g), W, false},
        {"ln_2.bias", reinterpret_cast<void**>(&b.ln2_b), W, false},
        {"mlp.c_fc.weight", reinterpret_cast<void**>(&b.fc_w32), int64_t(4) * W * W, false},
        {"mlp.c_fc.bias", reinterpret_cast<void**>(&b.fc_b), int64_t(4) * W, false},
        {"mlp.c_proj.weight", reinterpret_cast<void**>(&b.proj_w), int64_t(4) * W * W, true},
        {"mlp.c_proj.bias", reinterpret_cast<void**>(&b.proj_b), W, false},
    };
    for (auto& it : items) {
      const int64_t o = plan.take(it.numel * (it.bf ? 2 : 4));
      if (assign) {
        *it.dst = m->arena + o;
        m->slots.push_back({p + it.name, *it.dst, it.numel, it.bf, false});
      }
    }
    struct Derived {
      void** dst;
      int64_t bytes;
    } derived[] = {
        {reinterpret_cast<void**>(&b.qkv_w), int64_t(3) * Wa * W * 2}, {reinterpret_cast<void**>(&b.qkv_cs), int64_t(3) * Wa * 4},
        {reinterpret_cast<void**>(&b.qkv_bf), int64_t(3) * Wa * 4},    {reinterpret_cast<void**>(&b.fc_w), int64_t(4) * W * W * 2},
        {reinterpret_cast<void**>(&b.fc_cs), int64_t(4) * W * 4},     {reinterpret_cast<void**>(&b.fc_bf), int64_t(4) * W * 4},
    };
    for (auto& d : derived) {
      const int64_t o = plan.take(d.bytes);
      if (assign) *d.dst = m->arena + o;
    }
  }
}

// (Re)build the LayerNorm-folded weights of every block; called lazily by the encoders after parameters changed.
static int finalize_blocks(const std::vector<Block>& blocks, int W, int Wa, cudaStream_t s) {
  int rc;
  for (const Block& b : blocks) {
    if ((rc = fold_ln_weights(b.qkv_w32, b.ln1_g, b.ln1_b, b.qkv_b, b.qkv_w, b.qkv_cs, b.qkv_bf, 3 * Wa, W, s))) return rc;
    if ((rc = fold_ln_weights(b.fc_w32, b.ln2_g, b.ln2_b, b.fc_b, b.fc_w, b.fc_cs, b.fc_bf, 4 * W, W, s))) return rc;
  }
  return FC_OK;
}

// Two passes over the same plan: first to size the arena, then (assign) to hand out pointers.
static void plan_model(fc_model* m, ArenaPlan& plan, bool assign) {
  const fc_config& c = m->cfg;
  const int W = c.vision_width, Wt = c.transformer_width, E = c.embed_dim;
  const int Wa = c.vision_attn_width > 0 ? c.vision_attn_width : W;
  auto one = [&](const std::string& name, void** dst, int64_t numel, bool bf) {
    const int64_t o = plan.take(numel * (bf ? 2 : 4));
    if (assign) {
      *dst = m->arena + o;
      m->slots.push_back({name, *dst, numel, bf, false});
    }
  };
  {  // conv1.weight (W, 3, P, P) -> bf16 (W, patch_dim) with zero padding columns when 3*P*P is not a multiple of 8
    const int64_t o = plan.take(int64_t(W) * m->patch_dim * 2);
    if (assign) {
      m->conv_w = reinterpret_cast<bf16*>(m->arena + o);
      Slot sl{"visual.conv1.weight", m->conv_w, int64_t(W) * m->patch_cols, true, false};
      if (m->patch_cols != m->patch_dim) {
        sl.cols = m->patch_cols;
        sl.ld = m->patch_dim;
      }
      m->slots.push_back(sl);
    }
  }
  one("visual.class_embedding", reinterpret_cast<void**>(&m->cls), W, false);
  one("visual.positional_embedding", reinterpret_cast<void**>(&m->vpos), int64_t(m->L_img) * W, false);
  if (c.vision_tower == FC_TOWER_OPENAI) {  // timm's VisionTransformer has no LayerNorm in front of its blocks
    one("visual.ln_pre.weight", reinterpret_cast<void**>(&m->ln_pre_g), W, false);
    one("visual.ln_pre.bias", reinterpret_cast<void**>(&m->ln_pre_b), W, false);
  }
  one("visual.ln_post.weight", reinterpret_cast<void**>(&m->ln_post_g), W, false);
  one("visual.ln_post.bias", reinterpret_cast<void**>(&m->ln_post_b), W, false);
  one("visual.proj", reinterpret_cast<void**>(&m->vproj), int64_t(W) * E, false);
  plan_blocks(m, plan, m->vblocks, "visual.transformer.", c.vision_layers, W, Wa, assign);
  one("token_embedding.weight", reinterpret_cast<void**>(&m->tok), int64_t(c.vocab_size) * Wt, false);
  one("positional_embedding", reinterpret_cast<void**>(&m->tpos), int64_t(c.context_length) * Wt, false);
  one("ln_final.weight", reinterpret_cast<void**>(&m->ln_final_g), Wt, false);
  one("ln_final.bias", reinterpret_cast<void**>(&m->ln_final_b), Wt, false);
  one("text_projection", reinterpret_cast<void**>(&m->tproj), int64_t(Wt) * E, false);
  plan_blocks(m, plan, m->tblocks, "transformer.", c.transformer_layers, Wt, Wt, assign);

  // workspace: x (residual stream), y (LayerNorm out / attention out), big (patches | qkv | MLP hidden), features
  const int64_t vtok = int64_t(m->maxF) * m->L_img, ttok = int64_t(m->maxC) * c.context_length;
  const int64_t x_el = std::max(vtok * std::max(W, Wa), ttok * Wt);  // y also holds the attention output (width Wa)
  const int64_t big_el = std::max(std::max(vtok * std::max(4 * W, 3 * Wa), ttok * 4 * Wt),
                                  int64_t(m->maxF) * m->grid * m->grid * m->patch_dim);
  const int64_t ws0 = plan.off;
  auto ws = [&](void** dst, int64_t bytes) {
    const int64_t o = plan.take(bytes);
    if (assign) *dst = m->arena + o;
  };
  ws(reinterpret_cast<void**>(&m->x), x_el * 2);
  ws(reinterpret_cast<void**>(&m->y), x_el * 2);
  ws(reinterpret_cast<void**>(&m->big), big_el * 2);
  ws(reinterpret_cast<void**>(&m->feat), int64_t(std::max(m->maxF, m->maxC)) * E * 4);
  const int64_t stats_bytes = std::max(vtok * (W / 64), ttok * (Wt / 64)) * 2 * 4;
  ws(reinterpret_cast<void**>(&m->stats_a), stats_bytes);
  ws(reinterpret_cast<void**>(&m->stats_b), stats_bytes);
  ws(reinterpret_cast<void**>(&m->err_flag), 256);
  m->workspace_bytes = plan.off - ws0;
}

// The residual stream x arrives with its row statistics in `stats_a` ([rows, W/64, 2] partial sums / sums of squares).
// act / eps: EPI_LN_BIAS_QGELU and 1e-5 for the OpenAI towers, EPI_LN_BIAS_GELU and 1e-6 for timm's VisionTransformer.
static int run_blocks(const std::vector<Block>& blocks, bf16* x, bf16* y, bf16* big, float* stats_a, float* stats_b,
                      int64_t seqs, int L, int W, int heads, int causal, int act, float eps, cudaStream_t s) {
  const int Wa = heads * 64;  // attention width (== W unless narrow heads were padded to 64-wide slots)
  const int64_t rows64 = seqs * L;
  FC_REQUIRE(rows64 < (int64_t(1) << 31), "too many tokens in one pass");
  const int rows = static_cast<int>(rows64);
  const int parts = W / 64;
  int rc;
  for (const Block& b : blocks) {
    // x = x + out_proj(attention(ln_1(x)))                      (slip.py:383); ln_1 folded into the QKV GEMM
    GemmParams p;
    p.M = rows; p.N = 3 * Wa; p.K = W; p.C = big; p.ldc = 3 * Wa; p.bias = b.qkv_bf;
    p.ln_stats = stats_a; p.ln_parts = parts; p.colsum = b.qkv_cs; p.ln_eps = eps;
    if ((rc = gemm_bf16_tn(EPI_LN_BIAS, x, W, b.qkv_w, W, p, s))) return rc;
    if ((rc = attention_bf16(big, y, seqs, L, heads, causal, s))) return rc;
    p = GemmParams();
    p.M = rows; p.N = W; p.K = Wa; p.C = x; p.ldc = W; p.bias = b.out_b; p.resid = x; p.ldr = W; p.stats_out = stats_b;
    if ((rc = gemm_bf16_tn(EPI_BIAS_RESID, y, Wa, b.out_w, Wa, p, s))) return rc;
    // x = x + c_proj(quickgelu(c_fc(ln_2(x))))                  (slip.py:384); ln_2 folded into the fc1 GEMM
    p = GemmParams();
    p.M = rows; p.N = 4 * W; p.K = W; p.C = big; p.ldc = 4 * W; p.bias = b.fc_bf;
    p.ln_stats = stats_b; p.ln_parts = parts; p.colsum = b.fc_cs; p.ln_eps = eps;
    if ((rc = gemm_bf16_tn(act, x, W, b.fc_w, W, p, s))) return rc;
    p = GemmParams();
    p.M = rows; p.N = W; p.K = 4 * W; p.C = x; p.ldc = W; p.bias = b.proj_b; p.resid = x; p.ldr = W; p.stats_out = stats_a;
    if ((rc = gemm_bf16_tn(EPI_BIAS_RESID, big, 4 * W, b.proj_w, 4 * W, p, s))) return rc;
  }
  return FC_OK;
}

static int finalize(fc_model* m, cudaStream_t s) {
  if (!m->dirty) return FC_OK;
  int rc;
  const int Wa = m->cfg.vision_attn_width > 0 ? m->cfg.vision_attn_width : m->cfg.vision_width;
  if ((rc = finalize_blocks(m->vblocks, m->cfg.vision_width, Wa, s))) return rc;
  if ((rc = finalize_blocks(m->tblocks, m->cfg.transformer_width, m->cfg.transformer_width, s))) return rc;
  m->dirty = false;
  return FC_OK;
}

// CLIP.encode_image for F frames -> un-normalised features feat[F, E]   ([3P] VisionTransformer.forward)
// `frames` == nullptr: the patch matrix of the F frames is already in m->big (fc_encode_video_uint8)
static int vision_pass(fc_model* m, const void* frames, int dtype, int64_t F, float* feat, cudaStream_t s) {
  const fc_config& c = m->cfg;
  const int W = c.vision_width, L = m->L_img, G = m->grid;
  int rc;
  if (frames && (rc = im2col_patches(frames, dtype, m->big, F, c.image_resolution, c.vision_patch_size, s))) return rc;
  GemmParams p;
  p.M = static_cast<int>(F * G * G); p.N = W; p.K = m->patch_dim; p.C = m->x; p.ldc = W;
  p.pos = m->vpos; p.patches_per_frame = G * G;
  if ((rc = gemm_bf16_tn(EPI_PATCH, m->big, m->patch_dim, m->conv_w, m->patch_dim, p, s))) return rc;
  if ((rc = cls_rows(m->x, m->cls, m->vpos, F, L, W, s))) return rc;
  const bool timm = c.vision_tower == FC_TOWER_TIMM;
  const float eps = timm ? 1e-6f : 1e-5f;
  if (timm) {  // blocks start on the embedded tokens themselves: only their row statistics are needed
    if ((rc = row_stats_bf16(m->x, W, F * L, W, m->stats_a, s))) return rc;
  } else if ((rc = layernorm_bf16(m->x, W, m->x, W, m->ln_pre_g, m->ln_pre_b, F * L, W, eps, m->stats_a, s))) {
    return rc;
  }
  const int Wa = c.vision_attn_width > 0 ? c.vision_attn_width : W;
  if ((rc = run_blocks(m->vblocks, m->x, m->y, m->big, m->stats_a, m->stats_b, F, L, W, Wa / 64, 0,
                       timm ? EPI_LN_BIAS_GELU : EPI_LN_BIAS_QGELU, eps, s)))
    return rc;
  return head_project(m->x, nullptr, m->ln_post_g, m->ln_post_b, m->vproj, feat, F, L, W, c.embed_dim, eps, s);
}

// CLIP.encode_text for C captions -> un-normalised features feat[C, E]   (slip.py:468-480)
static int text_pass(fc_model* m, const int32_t* ids, int64_t C, float* feat, cudaStream_t s) {
  const fc_config& c = m->cfg;
  const int W = c.transformer_width, L = c.context_length;
  int rc;
  if ((rc = text_embed(ids, m->tok, m->tpos, m->x, C, L, W, c.vocab_size, m->err_flag, m->stats_a, s))) return rc;
  if ((rc = run_blocks(m->tblocks, m->x, m->y, m->big, m->stats_a, m->stats_b, C, L, W, c.transformer_heads, 1,
                       EPI_LN_BIAS_QGELU, 1e-5f, s)))
    return rc;
  return head_project(m->x, ids, m->ln_final_g, m->ln_final_b, m->tproj, feat, C, L, W, c.embed_dim, 1e-5f, s);
}

}  // namespace fc

using namespace fc;

// =================================================================================================== C ABI
extern "C" {

int fc_version(void) { return 101; }

size_t fc_last_error(char* buf, size_t cap) {
  const size_t n = strlen(g_err);
  if (buf && cap) {
    const size_t k = n < cap - 1 ? n : cap - 1;
    memcpy(buf, g_err, k);
    buf[k] = 0;
  }
  return n;
}

int64_t fc_launch_count(void) { return g_launches.load(); }

int fc_profile_start(int32_t max_records) {
  FC_REQUIRE(max_records > 0 && max_records <= (1 << 20), "fc_profile_start: bad capacity");
  while (static_cast<int>(g_prof.size()) < max_records) {
    ProfRec r{};
    FC_CUDA(cudaEventCreate(&r.e0));
    FC_CUDA(cudaEventCreate(&r.e1));
    g_prof.push_back(r);
  }
  g_prof_used = 0;
  g_prof_on = true;
  return FC_OK;
}

int fc_profile_stop(fc_profile_record* out, int32_t cap) {
  g_prof_on = false;
  FC_CUDA(cudaDeviceSynchronize());
  int n = 0;
  for (int i = 0; i < g_prof_used; ++i) {
    const ProfRec& r = g_prof[i];
    float ms = 0.f;
    FC_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    // aggregate by (kind, tag, n, k)
    int j = 0;
    for (; j < n; ++j)
      if (out[j].kind == r.kind && out[j].tag == r.tag && out[j].n == r.n && out[j].k == r.k) break;
    if (j == n) {
      if (n >= cap) continue;
      out[n] = fc_profile_record{r.kind, r.tag, r.n, r.k, 0, 0.0, 0.0, 0.0, 0.0};
      ++n;
    }
    out[j].launches += 1;
    out[j].ms += ms;
    out[j].flops += r.flops;
    out[j].bytes += r.bytes;
    out[j].rows += static_cast<double>(r.m);
  }
  g_prof_used = 0;
  return n;
}

int fc_model_create(const fc_config* cfg, fc_model** out) {
  FC_REQUIRE(cfg && out, "fc_model_create: null argument");
  int rc = check_arch();
  if (rc) return rc;
  const fc_config& c = *cfg;
  FC_REQUIRE(c.embed_dim > 0 && c.image_resolution > 0 && c.vision_layers > 0 && c.vision_patch_size > 0 &&
                 c.context_length > 0 && c.vocab_size > 0 && c.transformer_layers > 0,
             "fc_model_create: non-positive geometry");
  FC_REQUIRE(c.vision_width % 64 == 0 && c.transformer_width % 64 == 0 && c.vision_width <= 1024 &&
                 c.transformer_width <= 1024,
             "fc_model_create: widths must be multiples of 64 and <= 1024 (got %d / %d)", c.vision_width,
             c.transformer_width);
  FC_REQUIRE(c.transformer_heads * 64 == c.transformer_width, "fc_model_create: text head dim must be 64");
  FC_REQUIRE(c.vision_attn_width == 0 || (c.vision_attn_width % 64 == 0 && c.vision_attn_width >= c.vision_width &&
                                          c.vision_attn_width <= 1024),
             "fc_model_create: vision_attn_width %d must be a multiple of 64 in [vision_width, 1024]", c.vision_attn_width);
  FC_REQUIRE(c.vision_tower == FC_TOWER_OPENAI || c.vision_tower == FC_TOWER_TIMM,
             "fc_model_create: unknown vision tower %d", c.vision_tower);
  FC_REQUIRE(c.image_resolution % c.vision_patch_size == 0,
             "fc_model_create: image resolution %d is not a multiple of the patch size %d", c.image_resolution,
             c.vision_patch_size);
  fc_model* m = new fc_model();
  m->cfg = c;
  FC_CUDA(cudaGetDevice(&m->device));
  m->grid = c.image_resolution / c.vision_patch_size;
  m->L_img = m->grid * m->grid + 1;
  m->patch_cols = 3 * c.vision_patch_size * c.vision_patch_size;
  m->patch_dim = (m->patch_cols + 7) / 8 * 8;
  if (m->L_img > 768 || c.context_length > 768) {
    set_error("fc_model_create: sequence length above 768 tokens is not supported (image %d, text %d)", m->L_img,
              c.context_length);
    delete m;
    return FC_ERR_INVALID;
  }
  m->maxF = c.max_frames_per_pass > 0 ? c.max_frames_per_pass : 512;
  m->maxC = c.max_texts_per_pass > 0 ? c.max_texts_per_pass : 1024;
  ArenaPlan sizing;
  plan_model(m, sizing, false);
  m->arena_bytes = sizing.off;
  cudaError_t e = cudaMalloc(&m->arena, m->arena_bytes);
  if (e != cudaSuccess) {
    delete m;
    set_error("fc_model_create: cudaMalloc of %lld bytes failed: %s", static_cast<long long>(sizing.off),
              cudaGetErrorString(e));
    return FC_ERR_NOMEM;
  }
  ArenaPlan assign;
  m->slots.clear();
  plan_model(m, assign, true);
  cudaMemset(m->err_flag, 0, 256);
  *out = m;
  return FC_OK;
}

int fc_model_destroy(fc_model* m) {
  if (!m) return FC_OK;
  if (m->arena) cudaFree(m->arena);
  delete m;
  return FC_OK;
}

int fc_model_set_param(fc_model* m, const char* name, const float* data, int64_t numel, void* stream) {
  FC_REQUIRE(m && name && data, "fc_model_set_param: null argument");
  if (strcmp(name, "logit_scale") == 0) return FC_OK;  // unused by the encoder (clip_video_text_encoder.py:75-77)
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (Slot& sl : m->slots) {
    if (sl.name == name) {
      FC_REQUIRE(sl.numel == numel, "fc_model_set_param: %s has %lld elements, expected %lld", name,
                 static_cast<long long>(numel), static_cast<long long>(sl.numel));
      if (sl.as_bf16) {
        int rc = sl.ld ? f32_to_bf16_padded(data, static_cast<bf16*>(sl.dst), numel / sl.cols, sl.cols, sl.ld, s)
                       : f32_to_bf16(data, static_cast<bf16*>(sl.dst), numel, s);
        if (rc) return rc;
      } else {
        FC_CUDA(cudaMemcpyAsync(sl.dst, data, numel * 4, cudaMemcpyDeviceToDevice, s));
      }
      sl.loaded = true;
      m->dirty = true;
      return FC_OK;
    }
  }
  set_error("fc_model_set_param: unexpected parameter name \"%s\"", name);
  return FC_ERR_INVALID;
}

int fc_model_ready(fc_model* m) {
  if (!m) return 0;
  for (const Slot& sl : m->slots)
    if (!sl.loaded) {
      set_error("missing parameter \"%s\"", sl.name.c_str());
      return 0;
    }
  return 1;
}

int64_t fc_model_workspace_bytes(const fc_model* m) { return m ? m->workspace_bytes : 0; }

int fc_encode_video(fc_model* m, const void* frames, int dtype, int64_t videos, int32_t T, float* out_video,
                    float* out_frames, void* stream) {
  FC_REQUIRE(m && ((out_video && frames) || videos == 0), "fc_encode_video: null argument");
  FC_REQUIRE(videos >= 0 && T >= 1, "fc_encode_video: bad shape videos=%lld frames_per_video=%d",
             static_cast<long long>(videos), T);
  FC_REQUIRE(T <= m->maxF, "fc_encode_video: frames_per_video=%d exceeds max_frames_per_pass=%d", T, m->maxF);
  if (!fc_model_ready(m)) return FC_ERR_STATE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int frc = finalize(m, s)) return frc;
  const fc_config& c = m->cfg;
  const int64_t frame_elems = int64_t(3) * c.image_resolution * c.image_resolution;
  const int64_t esz = dtype == FC_F32 ? 4 : 2;
  // equal-sized passes (a short last pass would pay the same launch latencies and wave tails as a full one)
  const int64_t max_vids = m->maxF / T;
  const int64_t n_passes = (videos + max_vids - 1) / max_vids;
  const int64_t vids_per_pass = n_passes ? (videos + n_passes - 1) / n_passes : 1;
  for (int64_t v0 = 0; v0 < videos; v0 += vids_per_pass) {
    const int64_t nv = std::min(vids_per_pass, videos - v0);
    const uint8_t* src = static_cast<const uint8_t*>(frames) + v0 * T * frame_elems * esz;
    int rc = vision_pass(m, src, dtype, nv * T, m->feat, s);
    if (rc) return rc;
    if (out_frames)
      FC_CUDA(cudaMemcpyAsync(out_frames + v0 * T * c.embed_dim, m->feat, nv * T * c.embed_dim * 4,
                              cudaMemcpyDeviceToDevice, s));
    rc = pool_normalize(m->feat, out_video + v0 * c.embed_dim, nullptr, nv, T, c.embed_dim, 1.f, s);
    if (rc) return rc;
  }
  return FC_OK;
}

int fc_preprocess_to_patches(const uint8_t* frames, int64_t n, int32_t H, int32_t W, int32_t size, int32_t patch,
                             const float* mean, const float* std, void* patches, int64_t ldp, int32_t interpolation,
                             void* stream) {
  FC_REQUIRE(ldp > 0 && ldp < (int64_t(1) << 31), "fc_preprocess_to_patches: bad row stride");
  return preprocess_to_patches(frames, n, H, W, size, patch, mean, std, static_cast<bf16*>(patches),
                               static_cast<int>(ldp), interpolation, static_cast<cudaStream_t>(stream));
}

int fc_encode_video_uint8(fc_model* m, const uint8_t* frames, int64_t videos, int32_t T, int32_t H, int32_t W,
                          const float* mean, const float* std, int32_t interpolation, float* out_video, float* out_frames,
                          void* stream) {
  FC_REQUIRE(m && ((out_video && frames) || videos == 0) && mean && std, "fc_encode_video_uint8: null argument");
  FC_REQUIRE(videos >= 0 && T >= 1 && H > 0 && W > 0, "fc_encode_video_uint8: bad shape videos=%lld T=%d H=%d W=%d",
             static_cast<long long>(videos), T, H, W);
  FC_REQUIRE(T <= m->maxF, "fc_encode_video_uint8: frames_per_video=%d exceeds max_frames_per_pass=%d", T, m->maxF);
  if (!fc_model_ready(m)) return FC_ERR_STATE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int frc = finalize(m, s)) return frc;
  const fc_config& c = m->cfg;
  const int64_t frame_bytes = int64_t(3) * H * W;
  const int64_t max_vids = m->maxF / T;
  const int64_t n_passes = (videos + max_vids - 1) / max_vids;
  const int64_t vids_per_pass = n_passes ? (videos + n_passes - 1) / n_passes : 1;
  for (int64_t v0 = 0; v0 < videos; v0 += vids_per_pass) {
    const int64_t nv = std::min(vids_per_pass, videos - v0);
    const int64_t F = nv * T;
    // eval transform fused into the patch gather: uint8 -> /255 -> bicubic resize -> crop -> normalise -> bf16 patch rows
    if (m->patch_dim != m->patch_cols)  // padded patch rows (ViT-L/14: 588 -> 592): the pad columns must be zero
      FC_CUDA(cudaMemsetAsync(m->big, 0, static_cast<size_t>(F) * m->grid * m->grid * m->patch_dim * sizeof(bf16), s));
    int rc = preprocess_to_patches(frames + v0 * T * frame_bytes, F, H, W, c.image_resolution, c.vision_patch_size, mean,
                                   std, m->big, m->patch_dim, interpolation, s);
    if (rc) return rc;
    if ((rc = vision_pass(m, nullptr, FC_BF16, F, m->feat, s))) return rc;
    if (out_frames)
      FC_CUDA(cudaMemcpyAsync(out_frames + v0 * T * c.embed_dim, m->feat, F * c.embed_dim * 4, cudaMemcpyDeviceToDevice, s));
    if ((rc = pool_normalize(m->feat, out_video + v0 * c.embed_dim, nullptr, nv, T, c.embed_dim, 1.f, s))) return rc;
  }
  return FC_OK;
}

int fc_encode_text(fc_model* m, const int32_t* ids, int64_t texts, float* out_text, void* stream) {
  FC_REQUIRE(m && ((out_text && ids) || texts == 0), "fc_encode_text: null argument");
  FC_REQUIRE(texts >= 0, "fc_encode_text: negative count");
  if (!fc_model_ready(m)) return FC_ERR_STATE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int frc = finalize(m, s)) return frc;
  const fc_config& c = m->cfg;
  const int64_t n_passes = (texts + m->maxC - 1) / m->maxC;
  const int64_t per_pass = n_passes ? (texts + n_passes - 1) / n_passes : 1;
  for (int64_t c0 = 0; c0 < texts; c0 += per_pass) {
    const int64_t n = std::min<int64_t>(per_pass, texts - c0);
    int rc = text_pass(m, ids + c0 * c.context_length, n, m->feat, s);
    if (rc) return rc;
    rc = pool_normalize(m->feat, out_text + c0 * c.embed_dim, nullptr, n, 1, c.embed_dim, 1.f, s);
    if (rc) return rc;
  }
  return FC_OK;
}

int fc_model_check(fc_model* m, void* stream) {
  FC_REQUIRE(m, "fc_model_check: null model");
  FC_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  int flag = 0;
  FC_CUDA(cudaMemcpy(&flag, m->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    cudaMemset(m->err_flag, 0, sizeof(int));
    set_error("token id outside [0, vocab_size) seen by fc_encode_text");
    return FC_ERR_INVALID;
  }
  return FC_OK;
}

int fc_preprocess_frames(const uint8_t* frames, int64_t n, int32_t H, int32_t W, int32_t size, const float* mean,
                         const float* std, void* out, int out_dtype, int32_t interpolation, void* stream) {
  int rc = check_arch();
  if (rc) return rc;
  return preprocess_frames(frames, n, H, W, size, mean, std, out, out_dtype, interpolation,
                           static_cast<cudaStream_t>(stream));
}

int fc_pool_normalize(const float* x, float* out, void* out_bf16, int64_t rows_out, int32_t T, int32_t D, float scale,
                      void* stream) {
  return pool_normalize(x, out, static_cast<bf16*>(out_bf16), rows_out, T, D, scale,
                        static_cast<cudaStream_t>(stream));
}

int fc_wise_lerp(const float* p1, const float* p2, float* out, void* out_bf16, int64_t n, double w, void* stream) {
  return wise_lerp(p1, p2, out, static_cast<bf16*>(out_bf16), n, w, static_cast<cudaStream_t>(stream));
}

// ---- similarity: workspace = [A' (nt, terms*dim) bf16 | B' (nv, terms*dim) bf16], each 256-byte aligned
static int64_t sim_a_bytes(int64_t nt, int dim, int terms) { return align_up(nt * terms * dim * 2, 256); }

int64_t fc_sim_workspace_bytes(int64_t nt, int64_t nv, int32_t dim, int32_t terms) {
  return sim_a_bytes(nt, dim, terms) + align_up(nv * terms * dim * 2, 256);
}

static int sim_args(const void* ws, int64_t nt, int64_t nv, int dim, int terms) {
  FC_REQUIRE(ws, "similarity: null workspace");
  FC_REQUIRE(terms == 1 || terms == 3, "similarity: terms must be 1 or 3");
  FC_REQUIRE(dim % 8 == 0 && nt > 0 && nv > 0 && nt < (int64_t(1) << 31) && nv < (int64_t(1) << 31),
             "similarity: bad shape nt=%lld nv=%lld dim=%d", static_cast<long long>(nt), static_cast<long long>(nv),
             dim);
  return check_arch();
}

int fc_sim_prepare(const float* text_emb, const float* video_emb, int64_t nt, int64_t nv, int32_t dim, int32_t terms,
                   void* workspace, void* stream) {
  int rc = sim_args(workspace, nt, nv, dim, terms);
  if (rc) return rc;
  FC_REQUIRE(text_emb && video_emb, "fc_sim_prepare: null embeddings");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bf16* a = static_cast<bf16*>(workspace);
  bf16* b = reinterpret_cast<bf16*>(static_cast<uint8_t*>(workspace) + sim_a_bytes(nt, dim, terms));
  if ((rc = split_bf16(text_emb, a, nt, dim, 0, terms, s))) return rc;
  return split_bf16(video_emb, b, nv, dim, 1, terms, s);
}

static int sim_gemm(int epi, const void* workspace, int64_t nt, int64_t nv, int dim, int terms, GemmParams& p,
                    void* stream) {
  int rc = sim_args(workspace, nt, nv, dim, terms);
  if (rc) return rc;
  const bf16* a = static_cast<const bf16*>(workspace);
  const bf16* b = reinterpret_cast<const bf16*>(static_cast<const uint8_t*>(workspace) + sim_a_bytes(nt, dim, terms));
  p.M = static_cast<int>(nt);
  p.N = static_cast<int>(nv);
  p.K = terms * dim;
  return gemm_bf16_tn(epi, a, p.K, b, p.K, p, static_cast<cudaStream_t>(stream));
}

int fc_sim_target_scores(const void* workspace, int64_t nt, int64_t nv, int32_t dim, int32_t terms,
                         const int32_t* target, int32_t col_offset, float* tscore, void* stream) {
  GemmParams p;
  p.target = target;
  p.tscore_out = tscore;
  p.col_offset = col_offset;
  return sim_gemm(EPI_TARGET, workspace, nt, nv, dim, terms, p, stream);
}

int fc_sim_count(const void* workspace, int64_t nt, int64_t nv, int32_t dim, int32_t terms, const int32_t* target,
                 int32_t col_offset, const float* tscore, int32_t* counts, void* stream) {
  GemmParams p;
  p.target = target;
  p.target_score = tscore;
  p.counts = counts;
  p.col_offset = col_offset;
  return sim_gemm(EPI_COUNT, workspace, nt, nv, dim, terms, p, stream);
}

int fc_sim_scores(const void* workspace, int64_t nt, int64_t nv, int32_t dim, int32_t terms, float alpha,
                  float* scores, int64_t ld, void* stream) {
  FC_REQUIRE(scores && ld >= nv, "fc_sim_scores: bad output");
  GemmParams p;
  p.C = scores;
  p.ldc = ld;
  p.alpha = alpha;
  return sim_gemm(EPI_F32, workspace, nt, nv, dim, terms, p, stream);
}

int fc_rank_from_scores(const float* scores, int64_t ld, int64_t rows, int64_t cols, const int32_t* target,
                        int64_t* ranks, void* stream) {
  return rank_from_scores(scores, ld, rows, cols, target, ranks, static_cast<cudaStream_t>(stream));
}
int fc_counts_to_ranks(const int32_t* counts, int64_t* ranks, int64_t n, void* stream) {
  return counts_to_ranks(counts, ranks, n, static_cast<cudaStream_t>(stream));
}
int fc_metrics_from_ranks(const int64_t* ranks, int64_t n, int64_t num_candidates, float* recall_1_5_10,
                          int64_t* median_rank, float* mean_rank, void* stream) {
  return metrics_from_ranks(ranks, n, num_candidates, recall_1_5_10, median_rank, mean_rank,
                            static_cast<cudaStream_t>(stream));
}
int fc_topk_rows(const float* scores, int64_t ld, int64_t rows, int64_t cols, int32_t k, float* values,
                 int32_t* indices, void* stream) {
  return topk_rows(scores, ld, rows, cols, k, values, indices, static_cast<cudaStream_t>(stream));
}
int fc_nce_loss(const float* scores, int64_t ld, int32_t B, float* workspace, float* out, void* stream) {
  return nce_loss(scores, ld, B, workspace, out, static_cast<cudaStream_t>(stream));
}
int fc_ts_nce_loss(const float* scores, const float* teacher_scores, int64_t ld, int32_t B, float* workspace,
                   float* out, void* stream) {
  return ts_nce_loss(scores, teacher_scores, ld, B, workspace, out, static_cast<cudaStream_t>(stream));
}

int fc_gemm_bf16(int epilogue, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                 const float* bias, const void* resid, int64_t ldr, float alpha, int32_t M, int32_t N, int32_t K,
                 void* stream) {
  int rc = check_arch();
  if (rc) return rc;
  FC_REQUIRE(epilogue == FC_EPI_BIAS || epilogue == FC_EPI_BIAS_QGELU || epilogue == FC_EPI_BIAS_RESID ||
                 epilogue == FC_EPI_F32,
             "fc_gemm_bf16: epilogue %d is not exposed", epilogue);
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = bias;
  p.resid = static_cast<const bf16*>(resid); p.ldr = ldr; p.alpha = alpha;
  return gemm_bf16_tn(epilogue, static_cast<const bf16*>(A), lda, static_cast<const bf16*>(B), ldb, p,
                      static_cast<cudaStream_t>(stream));
}

int fc_layernorm_bf16(const void* x, void* y, const float* gamma, const float* beta, int64_t rows, int32_t D,
                      float eps, void* stream) {
  return layernorm_bf16(static_cast<const bf16*>(x), D, static_cast<bf16*>(y), D, gamma, beta, rows, D, eps, nullptr,
                        static_cast<cudaStream_t>(stream));
}

int fc_attention_bf16(const void* qkv, void* out, int64_t seqs, int32_t L, int32_t heads, int32_t causal,
                      void* stream) {
  return attention_bf16(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), seqs, L, heads, causal,
                        static_cast<cudaStream_t>(stream));
}

}  // extern "C"
