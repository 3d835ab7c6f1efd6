// UNet2DCondition / AutoencoderKL op graphs over the sm_100a kernels (see model.h).
//
// Architecture follows SURVEY.md Appendix A (diffusers-0.16 modules as used by
// gyre/pipeline/unet/core.py:274 and gyre/pipeline/unified_pipeline.py:1523-1536); parameter keys are the
// diffusers state-dict names so that real checkpoints drop in.  Internal activation layout: token-major
// NHWC fp16 ([B, H*W, C]); a 1x1 conv and a Linear are the same GEMM, the SD1 "conv projection" and the SD2
// "linear projection" transformer variants coincide, and no permute exists anywhere inside the network.
#include "model.h"

#include <cstring>

#include "common.cuh"
#include "ops.h"

namespace gyre {

// ------------------------------------------------------------------------------------------ Exec
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void* Exec::alloc_p(size_t bytes) {
  bytes = align_up(bytes, 1024);
  const size_t off = bottom;
  bottom += bytes;
  if (bottom + top > peak) peak = bottom + top;
  if (!dry && bottom + top > cap) overflow = true;
  return dry ? nullptr : base + off;
}

void* Exec::alloc_s(size_t bytes) {
  bytes = align_up(bytes, 1024);
  top += bytes;
  if (bottom + top > peak) peak = bottom + top;
  if (!dry && bottom + top > cap) {
    overflow = true;
    return base;   // never used: callers check ex.overflow before launching
  }
  return dry ? nullptr : base + cap - top;
}

#define EX_CHECK(ex)                                                                         \
  do {                                                                                       \
    if ((ex).overflow) {                                                                     \
      set_last_error("workspace too small: need > %zu bytes, have %zu", (ex).peak, (ex).cap); \
      return -4;                                                                             \
    }                                                                                        \
  } while (0)
// launch helper: skipped entirely in a dry (sizing) run
#define RUN(ex, call)          \
  do {                         \
    EX_CHECK(ex);              \
    if (!(ex).dry) GYRE_TRY(call); \
  } while (0)

// ------------------------------------------------------------------------------------------ Model base
constexpr size_t kStreamKBytes = 48u << 20;
constexpr int kStreamKFlags = 256;
constexpr size_t kGnPreFloats = 4u << 20;     // conv-produced GroupNorm partials (16 MB: 8 x 1024 x 1024 images in the VAE)
constexpr size_t kGnPreStats = 16u << 10;

Model::Model() {
  cudaGetDevice(&device_);
  sk_ws_ = dalloc(kStreamKBytes);
  sk_flags_ = static_cast<int*>(dalloc(kStreamKFlags * sizeof(int)));   // dalloc zero-fills
  sk_ws_bytes_ = sk_ws_ ? kStreamKBytes : 0;
  gn_pre_buf_ = static_cast<float*>(dalloc((kGnPreFloats + kGnPreStats) * sizeof(float)));
  gn_pre_floats_ = gn_pre_buf_ ? kGnPreFloats : 0;
}

Model::~Model() {
  for (void* p : allocs_) cudaFree(p);
}

void* Model::dalloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes > 0 ? bytes : 16) != cudaSuccess) {
    alloc_failed_ = true;
    cudaGetLastError();
    return nullptr;
  }
  cudaMemset(p, 0, bytes);
  allocs_.push_back(p);
  return p;
}

void Model::reg(const std::string& key, int kind, void* dst, std::initializer_list<int64_t> shape, int ld) {
  ParamSlot s;
  s.kind = kind;
  s.dst = dst;
  s.ndim = static_cast<int>(shape.size());
  int i = 0;
  for (int64_t v : shape) s.shape[i++] = v;
  s.ld = ld;
  slots_[key] = s;
}

void Model::reg_norm(const std::string& p, int C, NormW* n) {
  n->C = C;
  n->g = static_cast<float*>(dalloc(sizeof(float) * C));
  n->b = static_cast<float*>(dalloc(sizeof(float) * C));
  reg(p + ".weight", P_F32, n->g, {C});
  reg(p + ".bias", P_F32, n->b, {C});
}

void Model::reg_linear(const std::string& p, int N, int K, bool bias, LinW* l) {
  l->N = N;
  l->K = K;
  l->w = static_cast<__half*>(dalloc(sizeof(__half) * N * K));
  reg(p + ".weight", P_LINEAR, l->w, {N, K}, K);
  if (bias) {
    l->bias = static_cast<float*>(dalloc(sizeof(float) * N));
    reg(p + ".bias", P_F32, l->bias, {N});
  }
}

void Model::reg_conv3(const std::string& p, int Cin, int Cout, Conv3W* c, bool upsampler) {
  c->Cin = Cin;
  c->Cout = Cout;
  c->wp = static_cast<__half*>(dalloc(sizeof(__half) * conv3x3_packed_elems(Cin, Cout)));
  c->bias = static_cast<float*>(dalloc(sizeof(float) * Cout));
  reg(p + ".weight", upsampler ? P_CONV3_UP : P_CONV3, c->wp, {Cout, Cin, 3, 3});
  if (upsampler) {
    c->wp_up = static_cast<__half*>(dalloc(sizeof(__half) * upconv3x3_packed_elems(Cin, Cout)));
    slots_[p + ".weight"].dst2 = c->wp_up;
  }
  reg(p + ".bias", P_F32, c->bias, {Cout});
}

// Upsample2D (nearest 2x) + conv3x3.  Folded path (tunable UPCONV_FOLD): four 2x2 phase convolutions read the
// low-res tensor directly - the 4x larger upsampled tensor is never written and 5/9 of the MACs disappear.
// The scratch buffer is reserved either way so that the workspace size does not depend on the tunable.
int Model::upsample_conv(Exec& ex, const Conv3W& c, const __half* x, int B, int H, int W, __half* out) {
  __half* up = ex.s16(static_cast<size_t>(B) * 4 * H * W * c.Cin);
  Epilogue e;
  e.out = out;
  e.ldo = c.Cout;
  e.bias = c.bias;
  if (tunable(TUNE_UPCONV_FOLD) && c.wp_up != nullptr && c.Cout % 8 == 0) {
    RUN(ex, upconv2x_f16(x, c.Cin, B, H, W, c.Cin, c.wp_up, c.Cout, e, ex.st));
  } else {
    RUN(ex, upsample2x_nhwc(x, B, H, W, c.Cin, up, ex.st));
    RUN(ex, conv3x3_f16(up, c.Cin, B, 2 * H, 2 * W, c.Cin, c.wp, c.Cout, 1, 1, e, ex.st));
  }
  return 0;
}

void Model::reg_resnet(const std::string& p, int cin, int cout, bool temb, ResnetW* r) {
  r->cin = cin;
  r->cout = cout;
  reg_norm(p + ".norm1", cin, &r->n1);
  reg_conv3(p + ".conv1", cin, cout, &r->c1);
  reg_norm(p + ".norm2", cout, &r->n2);
  reg_conv3(p + ".conv2", cout, cout, &r->c2);
  r->has_sc = cin != cout;
  if (r->has_sc) reg_linear(p + ".conv_shortcut", cout, cin, true, &r->sc);
  if (temb) {
    // rows [temb_off, temb_off+cout) of the fused projection; registered once temb_proj_ is allocated
    r->temb_off = temb_total_;
    temb_total_ += cout;
  }
}

int Model::ensure_device() {
  int cur = -1;
  GYRE_CHECK_CUDA(cudaGetDevice(&cur));
  GYRE_REQUIRE(cur == device_, "handle belongs to device %d but device %d is current", device_, cur);
  return 0;
}

int Model::load(const char* key, const void* data, int dtype, const int64_t* shape, int ndim, cudaStream_t st) {
  GYRE_TRY(ensure_device());
  GYRE_REQUIRE(key && data && shape, "load_weight: null argument");
  GYRE_REQUIRE(dtype == 0 || dtype == 1, "load_weight(%s): dtype %d (want fp16=0 / fp32=1)", key, dtype);
  auto it = slots_.find(key);
  GYRE_REQUIRE(it != slots_.end(), "load_weight: unknown parameter '%s'", key);
  ParamSlot& s = it->second;
  // a Linear weight may arrive as a 1x1 conv [N, K, 1, 1] (SD1 proj_in / conv_shortcut) or as [N, K]
  bool ok = false;
  if (ndim == s.ndim) {
    ok = true;
    for (int i = 0; i < ndim; ++i) ok = ok && shape[i] == s.shape[i];
  } else if ((s.kind == P_LINEAR || s.kind == P_F32MAT) && s.ndim == 2 && ndim == 4) {
    // [N, K, 1, 1] (1x1 conv) or, generally, a conv weight whose trailing dims flatten to K (ViT patch embedding)
    ok = shape[0] == s.shape[0] && shape[1] * shape[2] * shape[3] == s.shape[1];
  }
  GYRE_REQUIRE(ok, "load_weight(%s): shape mismatch (got ndim %d [%lld,%lld,..], want ndim %d [%lld,%lld,..])", key,
               ndim, (long long)shape[0], (long long)(ndim > 1 ? shape[1] : 0), s.ndim, (long long)s.shape[0],
               (long long)(s.ndim > 1 ? s.shape[1] : 0));
  switch (s.kind) {
    case P_F32:
      GYRE_TRY(cast_to_f32(data, dtype, s.shape[0], static_cast<float*>(s.dst), st));
      break;
    case P_F32MAT:
      GYRE_TRY(cast_to_f32(data, dtype, s.shape[0] * s.shape[1], static_cast<float*>(s.dst), st));
      break;
    case P_LINEAR:
      GYRE_TRY(cast_to_f16(data, dtype, s.shape[0], static_cast<int>(s.shape[1]), static_cast<__half*>(s.dst), s.ld, st));
      break;
    case P_CONV3:
    case P_CONV3_UP:
      GYRE_TRY(pack_conv3x3(data, dtype, static_cast<int>(s.shape[1]), static_cast<int>(s.shape[0]),
                            static_cast<__half*>(s.dst), st));
      if (s.kind == P_CONV3_UP)
        GYRE_TRY(pack_upconv3x3(data, dtype, static_cast<int>(s.shape[1]), static_cast<int>(s.shape[0]),
                                static_cast<__half*>(s.dst2), st));
      break;
    case P_GEGLU_W:
      GYRE_TRY(pack_geglu(data, dtype, static_cast<int>(s.shape[0] / 2), static_cast<int>(s.shape[1]), nullptr, 0,
                          static_cast<__half*>(s.dst), nullptr, st));
      break;
    case P_GEGLU_B:
      GYRE_TRY(pack_geglu(nullptr, 0, static_cast<int>(s.shape[0] / 2), 1, data, dtype, nullptr,
                          static_cast<float*>(s.dst), st));
      break;
    default:
      GYRE_REQUIRE(false, "load_weight(%s): bad slot", key);
  }
  s.loaded = true;
  on_load(key);
  return 0;
}

int Model::finalize() {
  GYRE_REQUIRE(!alloc_failed_, "device allocation failed while creating the model");
  int missing = 0;
  std::string first;
  for (auto& kv : slots_)
    if (!kv.second.loaded) {
      if (missing == 0 || kv.first < first) first = kv.first;
      ++missing;
    }
  GYRE_REQUIRE(missing == 0, "finalize: %d parameter(s) not loaded, e.g. '%s'", missing, first.c_str());
  return 0;
}

// ------------------------------------------------------------------------------------------ shared blocks
// pointer offset that stays null in a dry run
template <typename T>
static inline T* off(T* p, size_t n) { return p ? p + n : nullptr; }

Epilogue Model::ep_out(__half* out, int ldo, const float* bias, const __half* residual, int ldr, int act) const {
  Epilogue e;
  e.sk_ws = sk_ws_;
  e.sk_ws_bytes = sk_ws_bytes_;
  e.sk_flags = sk_flags_;
  e.sk_flags_count = sk_flags_ ? kStreamKFlags : 0;
  e.out = out;
  e.ldo = ldo;
  e.bias = bias;
  e.residual = residual;
  e.ldr = ldr;
  e.act = act;
  return e;
}

void Model::gn_offer(Epilogue& e, int B, int H, int W, int Cout, int stride, int pad) {
  gn_pre_.x = nullptr;
  if (gn_pre_buf_ == nullptr || e.out == nullptr || e.act != ACT_NONE || e.ldo != Cout) return;
  const int parts = conv3x3_gn_parts(B, H, W, Cout, stride, pad, groups_);
  if (parts <= 0) return;
  const int Ho = conv3x3_out_extent(H, stride, pad), Wo = conv3x3_out_extent(W, stride, pad);
  if (!groupnorm_pre_ok(Cout, Ho * Wo, groups_)) return;
  if (static_cast<size_t>(B) * parts * groups_ * 2 > gn_pre_floats_ || static_cast<size_t>(B) * groups_ * 2 > kGnPreStats) return;
  e.gn_out = gn_pre_buf_;
  e.gn_groups = groups_;
  e.gn_nparts = parts;
  gn_pre_ = GnPre{static_cast<const __half*>(e.out), B, Ho * Wo, Cout, parts};
}

int Model::gnorm(Exec& ex, const __half* x1, int C1, const __half* x2, int C2, int B, int HW, float eps, const NormW& n,
                 bool silu, __half* out) {
  const GnPre pre = gn_pre_;
  gn_pre_.x = nullptr;
  float* scratch = ex.s32(gn_partials_floats(B, HW, groups_));
  if (!ex.dry && pre.x != nullptr && pre.x == x1 && C2 == 0 && pre.C == C1 && pre.B == B && pre.HW == HW) {
    RUN(ex, groupnorm_nhwc_pre(x1, C1, B, HW, groups_, eps, n.g, n.b, silu, out, gn_pre_buf_, pre.nparts,
                               gn_pre_buf_ + gn_pre_floats_, ex.st));
    return 0;
  }
  RUN(ex, groupnorm_nhwc(x1, C1, x2, C2, B, HW, groups_, eps, n.g, n.b, silu, out, scratch, ex.st));
  return 0;
}

// ResnetBlock2D: GN+SiLU -> conv3x3 (+bias +temb) -> GN+SiLU -> conv3x3 (+bias) + shortcut(x)
// The input may be the channel concatenation x1 ++ x2 (UNet up path): GroupNorm and the 1x1 shortcut read
// both sources directly, so the concatenation is never written to memory.
int Model::resnet(Exec& ex, const ResnetW& r, const __half* x1, int C1, const __half* x2, int C2, int B, int H, int W,
                  float eps, const __half* temb_all, int temb_ld, __half* out) {
  const int HW = H * W;
  const size_t rows = static_cast<size_t>(B) * HW;
  GYRE_REQUIRE(C1 + C2 == r.cin, "resnet: input channels %d+%d != %d", C1, C2, r.cin);
  __half* t1 = ex.s16(rows * r.cin);
  GYRE_TRY(gnorm(ex, x1, C1, x2, C2, B, HW, eps, r.n1, true, t1));
  __half* t2 = ex.s16(rows * r.cout);
  {
    Epilogue e = ep_out(t2, r.cout, r.c1.bias);
    if (r.temb_off >= 0) {
      e.rowgroup_bias = off(temb_all, r.temb_off);
      e.rgb_ld = temb_ld;
      e.rows_per_group = HW;
    }
    gn_offer(e, B, H, W, r.cout, 1, 1);
    RUN(ex, conv3x3_f16(t1, r.cin, B, H, W, r.cin, r.c1.wp, r.cout, 1, 1, e, ex.st));
  }
  __half* t3 = ex.s16(rows * r.cout);
  GYRE_TRY(gnorm(ex, t2, r.cout, nullptr, 0, B, HW, eps, r.n2, true, t3));
  const __half* res = x1;
  if (r.has_sc) {
    __half* t4 = ex.s16(rows * r.cout);
    Epilogue e = ep_out(t4, r.cout, r.sc.bias);
    RUN(ex, gemm2_f16(x1, C1, C1, x2, C2, C2, r.sc.w, r.cin, static_cast<int>(rows), r.cout, e, ex.st));
    res = t4;
  } else {
    GYRE_REQUIRE(C2 == 0, "resnet: concatenated input needs a shortcut conv");
  }
  {
    Epilogue e = ep_out(out, r.cout, r.c2.bias, res, r.cout);
    gn_offer(e, B, H, W, r.cout, 1, 1);     // whoever normalises `out` next (a transformer's GroupNorm, the next resnet)
    RUN(ex, conv3x3_f16(t3, r.cout, B, H, W, r.cout, r.c2.wp, r.cout, 1, 1, e, ex.st));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ UNet
UNetModel::UNetModel(const gyre_b200_unet_config& cfg) : cfg_(cfg) {
  ln_fuse_ = tunable(TUNE_LN_FUSE) != 0;
  groups_ = cfg.norm_num_groups;
  const int L = cfg.num_levels;
  const int* ch = cfg.block_out_channels;
  temb_dim_ = ch[0] * 4;
  reg_conv3("conv_in", cfg.in_channels, ch[0], &conv_in_);
  reg_linear("time_embedding.linear_1", temb_dim_, ch[0], true, &time1_);
  reg_linear("time_embedding.linear_2", temb_dim_, temb_dim_, true, &time2_);

  if (cfg.addition_embed_dim > 0) {
    reg_linear("add_embedding.linear_1", temb_dim_, cfg.addition_embed_dim, true, &add1_);
    reg_linear("add_embedding.linear_2", temb_dim_, temb_dim_, true, &add2_);
  }

  auto add_transformer = [&](const std::string& p, int C, int heads, int depth) {
    tblocks_.emplace_back();
    TransformerW* t = &tblocks_.back();
    t->C = C;
    t->heads = heads;
    const int ctx = cfg.cross_attention_dim;
    reg_norm(p + ".norm", C, &t->gn);
    reg_linear(p + ".proj_in", C, C, true, &t->proj_in);
    reg_linear(p + ".proj_out", C, C, true, &t->proj_out);
    t->blocks.resize(depth < 1 ? 1 : depth);
    for (size_t bi = 0; bi < t->blocks.size(); ++bi) {
      TBlockW* k = &t->blocks[bi];
      k->flat = n_tblocks_flat_++;
      const std::string b = p + ".transformer_blocks." + std::to_string(bi);
      reg_norm(b + ".norm1", C, &k->ln1);
      reg_norm(b + ".norm2", C, &k->ln2);
      reg_norm(b + ".norm3", C, &k->ln3);
      // attn1: fused [to_q ; to_k ; to_v] -> one GEMM with N = 3C
      k->qkv.N = 3 * C;
      k->qkv.K = C;
      k->qkv.w = static_cast<__half*>(dalloc(sizeof(__half) * 3 * C * C));
      reg(b + ".attn1.to_q.weight", P_LINEAR, k->qkv.w, {C, C}, C);
      reg(b + ".attn1.to_k.weight", P_LINEAR, k->qkv.w ? k->qkv.w + static_cast<size_t>(C) * C : nullptr, {C, C}, C);
      reg(b + ".attn1.to_v.weight", P_LINEAR, k->qkv.w ? k->qkv.w + static_cast<size_t>(2) * C * C : nullptr, {C, C}, C);
      reg_linear(b + ".attn1.to_out.0", C, C, true, &k->o1);
      // attn2: q from tokens, fused [to_k ; to_v] from the text context
      reg_linear(b + ".attn2.to_q", C, C, false, &k->q2);
      k->kv2.N = 2 * C;
      k->kv2.K = ctx;
      k->kv2.w = static_cast<__half*>(dalloc(sizeof(__half) * 2 * C * ctx));
      reg(b + ".attn2.to_k.weight", P_LINEAR, k->kv2.w, {C, ctx}, ctx);
      reg(b + ".attn2.to_v.weight", P_LINEAR, k->kv2.w ? k->kv2.w + static_cast<size_t>(C) * ctx : nullptr, {C, ctx}, ctx);
      reg_linear(b + ".attn2.to_out.0", C, C, true, &k->o2);
      // GEGLU feed-forward
      k->geglu.N = 8 * C;
      k->geglu.K = C;
      k->geglu.w = static_cast<__half*>(dalloc(sizeof(__half) * 8 * C * C));
      k->geglu.bias = static_cast<float*>(dalloc(sizeof(float) * 8 * C));
      reg(b + ".ff.net.0.proj.weight", P_GEGLU_W, k->geglu.w, {8 * C, C});
      reg(b + ".ff.net.0.proj.bias", P_GEGLU_B, k->geglu.bias, {8 * C});
      reg_linear(b + ".ff.net.2", C, 4 * C, true, &k->ff2);
      if (ln_fuse_) {
        auto derived = [&](LnFoldW* f, int N) {
          f->w = static_cast<__half*>(dalloc(sizeof(__half) * N * C));
          f->colsum = static_cast<float*>(dalloc(sizeof(float) * N));
          f->bias = static_cast<float*>(dalloc(sizeof(float) * N));
        };
        derived(&k->qkv_ln, 3 * C);
        derived(&k->q2_ln, C);
        derived(&k->geglu_ln, 8 * C);
      }
    }
  };
  auto depth_of = [&](int lvl) { return cfg.transformer_depth[lvl] > 0 ? cfg.transformer_depth[lvl] : 1; };
  auto add_resnet = [&](const std::string& p, int cin, int cout) {
    resnets_.emplace_back();
    reg_resnet(p, cin, cout, true, &resnets_.back());
  };
  // reserve so that emplace_back never moves already-registered blocks (slots hold raw pointers into
  // device memory only, but ResnetW/TransformerW copies must stay stable for the lambdas above)
  resnets_.reserve(64);
  tblocks_.reserve(64);
  downs_.reserve(8);
  ups_.reserve(8);

  std::vector<int> skips{ch[0]};
  int cin = ch[0];
  for (int i = 0; i < L; ++i) {
    for (int j = 0; j < cfg.layers_per_block; ++j) {
      const std::string p = "down_blocks." + std::to_string(i);
      add_resnet(p + ".resnets." + std::to_string(j), cin, ch[i]);
      cin = ch[i];
      if (cfg.attn_levels[i]) add_transformer(p + ".attentions." + std::to_string(j), ch[i], cfg.num_heads[i], depth_of(i));
      skips.push_back(ch[i]);
    }
    if (i < L - 1) {
      downs_.emplace_back();
      reg_conv3("down_blocks." + std::to_string(i) + ".downsamplers.0.conv", ch[i], ch[i], &downs_.back());
      skips.push_back(ch[i]);
    }
  }
  add_resnet("mid_block.resnets.0", cin, cin);
  add_transformer("mid_block.attentions.0", cin, cfg.num_heads[L - 1], depth_of(L - 1));
  add_resnet("mid_block.resnets.1", cin, cin);
  if (cfg.controlnet) {
    // controlnet/models.py:41-94 (conditioning embedding) and :216-262 (one zero-initialised 1x1 conv per skip + mid)
    const int ec[4] = {16, 32, 96, 256};
    const int cc = cfg.conditioning_channels > 0 ? cfg.conditioning_channels : 3;
    cn_embed_.resize(8);
    reg_conv3("controlnet_cond_embedding.conv_in", cc, ec[0], &cn_embed_[0]);
    for (int i = 0; i < 3; ++i) {
      reg_conv3("controlnet_cond_embedding.blocks." + std::to_string(2 * i), ec[i], ec[i], &cn_embed_[1 + 2 * i]);
      reg_conv3("controlnet_cond_embedding.blocks." + std::to_string(2 * i + 1), ec[i], ec[i + 1], &cn_embed_[2 + 2 * i]);
    }
    reg_conv3("controlnet_cond_embedding.conv_out", ec[3], ch[0], &cn_embed_[7]);
    cn_down_.resize(skips.size());
    for (size_t k = 0; k < skips.size(); ++k)
      reg_linear("controlnet_down_blocks." + std::to_string(k), skips[k], skips[k], true, &cn_down_[k]);
    reg_linear("controlnet_mid_block", cin, cin, true, &cn_mid_);
  }
  for (int i = 0; i < (cfg.controlnet ? 0 : L); ++i) {
    const int lvl = L - 1 - i;
    const int c = ch[lvl];
    for (int j = 0; j < cfg.layers_per_block + 1; ++j) {
      const int s = skips.back();
      skips.pop_back();
      const std::string p = "up_blocks." + std::to_string(i);
      add_resnet(p + ".resnets." + std::to_string(j), cin + s, c);
      cin = c;
      if (cfg.attn_levels[lvl]) add_transformer(p + ".attentions." + std::to_string(j), c, cfg.num_heads[lvl], depth_of(lvl));
    }
    if (i < L - 1) {
      ups_.emplace_back();
      reg_conv3("up_blocks." + std::to_string(i) + ".upsamplers.0.conv", c, c, &ups_.back(), true);
    }
  }
  if (!cfg.controlnet) {
    reg_norm("conv_norm_out", ch[0], &norm_out_);
    reg_conv3("conv_out", ch[0], cfg.out_channels, &conv_out_);
  }

  // fused time_emb_proj: one [sum(cout), temb_dim] GEMM per forward instead of 22 M=batch GEMMs
  temb_proj_.N = temb_total_;
  temb_proj_.K = temb_dim_;
  temb_proj_.w = static_cast<__half*>(dalloc(sizeof(__half) * temb_total_ * temb_dim_));
  temb_proj_.bias = static_cast<float*>(dalloc(sizeof(float) * temb_total_));
  {
    // re-walk the resnet names in the same order to register the slices
    size_t idx = 0;
    auto reg_temb = [&](const std::string& p) {
      const ResnetW& r = resnets_[idx++];
      reg(p + ".time_emb_proj.weight", P_LINEAR,
          temb_proj_.w ? temb_proj_.w + static_cast<size_t>(r.temb_off) * temb_dim_ : nullptr, {r.cout, temb_dim_},
          temb_dim_);
      reg(p + ".time_emb_proj.bias", P_F32, temb_proj_.bias ? temb_proj_.bias + r.temb_off : nullptr, {r.cout});
    };
    for (int i = 0; i < L; ++i)
      for (int j = 0; j < cfg.layers_per_block; ++j)
        reg_temb("down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j));
    reg_temb("mid_block.resnets.0");
    reg_temb("mid_block.resnets.1");
    for (int i = 0; i < (cfg.controlnet ? 0 : L); ++i)
      for (int j = 0; j < cfg.layers_per_block + 1; ++j)
        reg_temb("up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j));
  }
}

UNetModel::~UNetModel() {
  for (__half* p : kv_cache_)
    if (p) cudaFree(p);
}

void UNetModel::on_load(const std::string& key) {
  if (key.find(".transformer_blocks.") != std::string::npos) ln_dirty_ = true;
}

// Derives, for every Linear that follows a LayerNorm (attn1 q|k|v, attn2 to_q, the GEGLU projection), the gamma-scaled
// weight copy, its column sums and the beta-folded bias.  Runs once after a (re)load, before the next forward.
int UNetModel::refold_layernorms(cudaStream_t st) {
  for (TransformerW& t : tblocks_)
    for (TBlockW& k : t.blocks) {
      const int C = t.C;
      GYRE_TRY(ln_fold_linear(k.qkv.w, 3 * C, C, k.ln1.g, k.ln1.b, nullptr, k.qkv_ln.w, k.qkv_ln.colsum, k.qkv_ln.bias, st));
      GYRE_TRY(ln_fold_linear(k.q2.w, C, C, k.ln2.g, k.ln2.b, nullptr, k.q2_ln.w, k.q2_ln.colsum, k.q2_ln.bias, st));
      // the GEGLU weight is packed (value / gate rows interleaved per tile) and so is its bias: row n of one is row n
      // of the other, which is all the fold needs
      GYRE_TRY(ln_fold_linear(k.geglu.w, 8 * C, C, k.ln3.g, k.ln3.b, k.geglu.bias, k.geglu_ln.w, k.geglu_ln.colsum,
                              k.geglu_ln.bias, st));
    }
  ln_dirty_ = false;
  return 0;
}

int UNetModel::set_context(const __half* ctx, int B, int L, cudaStream_t st) {
  GYRE_TRY(ensure_device());
  if (ctx == nullptr) {
    ctx_B_ = ctx_L_ = 0;
    return 0;
  }
  GYRE_REQUIRE(B > 0 && L > 0, "set_context: empty context");
  kv_cache_.resize(n_tblocks_flat_, nullptr);
  kv_cache_elems_.resize(n_tblocks_flat_, 0);
  ctx_B_ = ctx_L_ = 0;
  for (const TransformerW& t : tblocks_) {
    for (const TBlockW& k : t.blocks) {
      const size_t i = static_cast<size_t>(k.flat);
      const size_t need = static_cast<size_t>(B) * L * 2 * t.C;
      if (kv_cache_elems_[i] < need) {
        if (kv_cache_[i]) cudaFree(kv_cache_[i]);   // synchronises: no forward can still be reading it
        kv_cache_[i] = nullptr;
        kv_cache_elems_[i] = 0;
        GYRE_CHECK_CUDA(cudaMalloc(&kv_cache_[i], need * sizeof(__half)));
        kv_cache_elems_[i] = need;
      }
      GYRE_TRY(gemm_f16(ctx, k.kv2.K, k.kv2.w, k.kv2.K, B * L, 2 * t.C, k.kv2.K, ep_out(kv_cache_[i], 2 * t.C), st));
    }
  }
  ctx_B_ = B;
  ctx_L_ = L;
  return 0;
}

// Transformer2DModel + BasicTransformerBlock (nonfree/tome_unet.py:114-136)
int UNetModel::transformer(Exec& ex, const TransformerW& t, const __half* x, int B, int HW, const __half* ctx, int L,
                           int r, __half* out, bool share) {
  // `share` (CFG-parallel input, UNetModel::forward): the two halves of the batch are the SAME samples with different
  // text contexts, so everything up to the first cross-attention - GroupNorm, proj_in, the whole self-attention, attn2's
  // query projection - is computed once on `x` = [B / 2] samples; the halves part ways at the cross-attention's K / V.
  const int C = t.C;
  const int Bs = share ? B / 2 : B;             // samples of the shared prefix
  const int M = B * HW, Ms = Bs * HW;
  const int nh = share ? 2 : 1;                 // launches that fan the shared rows out to both halves
  const size_t n = static_cast<size_t>(M) * C;
  const size_t ns = static_cast<size_t>(Ms) * C;
  const int d = C / t.heads;
  const float scale = 1.0f / sqrtf(static_cast<float>(d));
  __half* tn = ex.s16(n);
  GYRE_TRY(gnorm(ex, x, C, nullptr, 0, Bs, HW, 1e-6f, t.gn, false, tn));
  __half* h = ex.s16(n);
  // LayerNorm folded into the GEMMs around it: the GEMM that PRODUCES a residual-stream tensor also emits
  // per-row (sum, sum of squares) partials; a tiny kernel turns them into (mean, rstd); the GEMM that CONSUMES the
  // normalised tensor reads the RAW rows against gamma-scaled weights and applies rstd * (acc - mean * colsum) + bias'
  // in its epilogue.  The standalone LayerNorm pass (one read + one write of the tensor, three times per block) is gone.
  // The scratch is reserved whenever the model was built with the folded weights, so the workspace size does not
  // depend on the run-time tunable.
  const bool fuse = ln_fuse_ && tunable(TUNE_LN_FUSE) != 0 && C > 32;
  const int parts_full = ln_fuse_ ? gemm_rowstat_parts(M, C) : 0;
  const int parts_s = ln_fuse_ ? gemm_rowstat_parts(Ms, C) : 0;       // what a producer launched on Ms rows writes
  const int parts_max = parts_full > parts_s ? parts_full : parts_s;
  float2* ln_part = ln_fuse_ ? reinterpret_cast<float2*>(ex.s32(static_cast<size_t>(2) * parts_max * M)) : nullptr;
  float2* ln_stat = ln_fuse_ ? reinterpret_cast<float2*>(ex.s32(static_cast<size_t>(2) * M)) : nullptr;
  // State of the statistics the last producer(s) left: `cur_parts` partials per row, laid out [cur_parts][cur_rows].
  int cur_parts = 0, cur_rows = 0;
  // producer epilogue: also leave the row statistics of what it writes (rows [row0, row0 + launch rows) of `rows` in all)
  auto with_stats = [&](Epilogue e, int launch_rows, int rows, int row0) {
    if (fuse) {
      e.rowstat_out = ln_part + row0;
      e.rowstat_ld = rows;
      cur_parts = gemm_rowstat_parts(launch_rows, C);
      cur_rows = rows;
    }
    return e;
  };
  // up to 8 partials per row (C = 320 / 640: two / four 160-column tiles x two column halves) the consumer folds them
  // itself; wider rows go through the finalize kernel
  auto finalize_stats = [&]() -> int {
    if (fuse && cur_parts > 8) RUN(ex, ln_finalize_rows(ln_part, cur_parts, cur_rows, C, 1e-5f, ln_stat, ex.st));
    return 0;
  };
  // consumer epilogue: LayerNorm(x) @ W^T from the raw rows
  auto ln_ep = [&](Epilogue e, const LnFoldW& f) {
    const bool direct = cur_parts <= 8;
    e.bias = f.bias;
    e.ln_colsum = f.colsum;
    e.ln_rowstat = direct ? ln_part : ln_stat;
    e.ln_parts = direct ? cur_parts : 0;
    e.ln_inv_c = 1.0f / static_cast<float>(C);
    e.ln_eps = 1e-5f;
    return e;
  };
  (void)parts_s;
  RUN(ex, gemm_f16(tn, C, t.proj_in.w, C, Ms, C, C, with_stats(ep_out(h, C, t.proj_in.bias), Ms, Ms, 0), ex.st));
  GYRE_TRY(finalize_stats());
  __half* nrm = tn;   // the GroupNorm output is dead after proj_in: reuse it for the LayerNorm outputs
  __half* qkv = ex.s16(n * 3);
  __half* o = ex.s16(n);
  __half* h2 = ex.s16(n);
  __half* g = ex.s16(n * 4);
  __half* kv_new = ex.s16(static_cast<size_t>(B) * L * 2 * C);   // reserved even when the bound context is used
  __half *km = nullptr, *vm = nullptr;
  void* tws = nullptr;
  size_t tome_bytes = 0;
  const int rr = r > 0 ? (r < HW / 2 ? r : HW / 2) : 0;
  if (rr > 0 || ex.dry) {
    // the sizing (dry) run reserves the ToMe scratch for the smallest merge (largest K/V), so that a workspace sized
    // once serves every r
    GYRE_REQUIRE(ex.dry || t.blocks.size() == 1, "ToMe is wired for transformer depth 1");
    GYRE_TRY(tome_workspace_bytes(B, HW, C, &tome_bytes));
    tws = ex.alloc_s(tome_bytes);
    const int nk_max = ex.dry ? HW : HW - rr;
    km = ex.s16(static_cast<size_t>(B) * nk_max * C);
    vm = ex.s16(static_cast<size_t>(B) * nk_max * C);
  }
  // h holds the residual stream on entry to every block and again on exit (h -> h2 -> h -> h2 -> copy-free swap)
  for (size_t bi = 0; bi < t.blocks.size(); ++bi) {
    const TBlockW& k = t.blocks[bi];
    // rows / samples of this block's self-attention part: the shared prefix exists in the first block only
    const bool sh = share && bi == 0;
    const int Bp = sh ? Bs : B, Mp = sh ? Ms : M;
    // ---- self-attention
    if (fuse) {
      RUN(ex, gemm_f16(h, C, k.qkv_ln.w, C, Mp, 3 * C, C, ln_ep(ep_out(qkv, 3 * C), k.qkv_ln), ex.st));
    } else {
      RUN(ex, layernorm_rows(h, Mp, C, 1e-5f, k.ln1.g, k.ln1.b, nrm, ex.st));
      RUN(ex, gemm_f16(nrm, C, k.qkv.w, C, Mp, 3 * C, C, ep_out(qkv, 3 * C), ex.st));
    }
    if (rr > 0) {
      // ToMe (nonfree/tome_memory_efficient_cross_attention.py:28-50): merge K and V with one plan built from K
      const int nk = HW - rr;
      RUN(ex, tome_merge_kv(off(qkv, C), off(qkv, 2 * C), 3 * C, Bp, HW, C, rr, km, vm, tws, tome_bytes, ex.st));
      RUN(ex, attention_f16(qkv, 3 * C, km, C, vm, C, Bp, t.heads, HW, nk, d, scale, o, C, ex.st));
    } else {
      RUN(ex, attention_f16(qkv, 3 * C, off(qkv, C), 3 * C, off(qkv, 2 * C), 3 * C, Bp, t.heads, HW, HW, d, scale, o, C, ex.st));
    }
    RUN(ex, gemm_f16(o, C, k.o1.w, C, Mp, C, C, with_stats(ep_out(h2, C, k.o1.bias, h, C), Mp, Mp, 0), ex.st));
    GYRE_TRY(finalize_stats());
    // ---- cross-attention: the query projection still belongs to the shared prefix
    __half* q = qkv;   // dead after self-attention
    if (fuse) {
      RUN(ex, gemm_f16(h2, C, k.q2_ln.w, C, Mp, C, C, ln_ep(ep_out(q, C), k.q2_ln), ex.st));
    } else {
      RUN(ex, layernorm_rows(h2, Mp, C, 1e-5f, k.ln2.g, k.ln2.b, nrm, ex.st));
      RUN(ex, gemm_f16(nrm, C, k.q2.w, C, Mp, C, C, ep_out(q, C), ex.st));
    }
    const __half* kv;
    if (ctx != nullptr || ex.dry) {
      RUN(ex, gemm_f16(ctx, k.kv2.K, k.kv2.w, k.kv2.K, B * L, 2 * C, k.kv2.K, ep_out(kv_new, 2 * C), ex.st));
      kv = kv_new;
    } else {
      kv = kv_cache_[static_cast<size_t>(k.flat)];
    }
    // shared prefix: the same queries against each half's keys / values, then attn2's output projection with the same
    // residual rows - from here on the two halves are different tensors
    const int fan = sh ? nh : 1;
    for (int hf = 0; hf < fan; ++hf) {
      const size_t kvo = static_cast<size_t>(hf) * Bp * L * 2 * C, oo = static_cast<size_t>(hf) * Mp * C;
      RUN(ex, attention_f16(q, C, off(kv, kvo), 2 * C, off(kv, kvo + C), 2 * C, Bp, t.heads, HW, L, d, scale, off(o, oo), C,
                            ex.st));
    }
    for (int hf = 0; hf < fan; ++hf) {
      const size_t oo = static_cast<size_t>(hf) * Mp * C;
      RUN(ex, gemm_f16(off(o, oo), C, k.o2.w, C, Mp, C, C,
                       with_stats(ep_out(off(h, oo), C, k.o2.bias, h2, C), Mp, fan * Mp, hf * Mp), ex.st));   // h <- h2 + attn2
    }
    GYRE_TRY(finalize_stats());
    // ---- GEGLU feed-forward
    const bool fuse_ff = fuse && tunable(TUNE_GELU_FAST) != 0;     // the folded GEGLU image carries the fast GELU only
    if (fuse_ff) {
      RUN(ex, gemm_f16(h, C, k.geglu_ln.w, C, M, 8 * C, C,
                       ln_ep(ep_out(g, 4 * C, nullptr, nullptr, 0, ACT_GEGLU), k.geglu_ln), ex.st));
    } else {
      RUN(ex, layernorm_rows(h, M, C, 1e-5f, k.ln3.g, k.ln3.b, nrm, ex.st));
      RUN(ex, gemm_f16(nrm, C, k.geglu.w, C, M, 8 * C, C, ep_out(g, 4 * C, k.geglu.bias, nullptr, 0, ACT_GEGLU), ex.st));
    }
    // h2 <- h + ff; with another block behind it, its norm1 needs the statistics of what this GEMM writes
    const bool more = bi + 1 < t.blocks.size();
    Epilogue e_ff2 = ep_out(h2, C, k.ff2.bias, h, C);
    RUN(ex, gemm_f16(g, 4 * C, k.ff2.w, 4 * C, M, C, 4 * C, more ? with_stats(e_ff2, M, M, 0) : e_ff2, ex.st));
    if (more) GYRE_TRY(finalize_stats());
    std::swap(h, h2);   // the block's output becomes the next block's residual stream
  }
  h2 = h;
  // proj_out + the transformer's input as residual (shared prefix: the same input rows for both halves)
  for (int hf = 0; hf < nh; ++hf) {
    const size_t oo = static_cast<size_t>(hf) * Ms * C;
    RUN(ex, gemm_f16(off(h2, oo), C, t.proj_out.w, C, Ms, C, C, ep_out(off(out, oo), C, t.proj_out.bias, x, C), ex.st));
  }
  (void)ns;
  return 0;
}

int UNetModel::num_skips() const {
  const int nl = cfg_.num_levels;
  return 1 + nl * cfg_.layers_per_block + (nl - 1);
}

int UNetModel::set_control_residuals(const __half* const* down, int n_down, const __half* mid) {
  if (n_down == 0 && mid == nullptr) {
    ctrl_down_.clear();
    ctrl_mid_ = nullptr;
    return 0;
  }
  GYRE_REQUIRE(n_down == 0 || n_down == num_skips(), "unet_set_control_residuals: %d down residuals given, the model has %d skips",
               n_down, num_skips());
  GYRE_REQUIRE(n_down == 0 || down != nullptr, "unet_set_control_residuals: null residual list");
  ctrl_down_.assign(down, down + n_down);
  for (const __half* p : ctrl_down_) GYRE_REQUIRE(p != nullptr, "unet_set_control_residuals: null residual tensor");
  ctrl_mid_ = mid;
  return 0;
}

int UNetModel::set_adapter_states(const __half* const* states, int n) {
  if (n == 0) {
    adapter_.clear();
    return 0;
  }
  GYRE_REQUIRE(n == cfg_.num_levels && states != nullptr, "unet_set_adapter_states: %d states given, the model has %d down blocks",
               n, cfg_.num_levels);
  adapter_.assign(states, states + n);
  for (const __half* p : adapter_) GYRE_REQUIRE(p != nullptr, "unet_set_adapter_states: null state tensor");
  return 0;
}

int UNetModel::forward(Exec& ex, const __half* sample, const int64_t* t, const __half* ctx, const __half* add_cond, int B,
                       int H, int W, int L, const int32_t* tome_r, __half* out, const ControlNetIO* cn) {
  GYRE_REQUIRE(B > 0 && H > 0 && W > 0 && L > 0, "unet_forward: empty problem");
  gn_forget();
  GYRE_REQUIRE(ex.dry || (cfg_.controlnet != 0) == (cn != nullptr),
               "unet_forward: a ControlNet handle runs through gyre_b200_controlnet_forward (and only it)");
  if (!ex.dry)
    GYRE_REQUIRE((cfg_.addition_embed_dim > 0) == (add_cond != nullptr),
                 "unet_forward: this model %s an additional conditioning vector (addition_embed_dim = %d)",
                 cfg_.addition_embed_dim > 0 ? "needs" : "does not take", cfg_.addition_embed_dim);
  if (!ex.dry && ctx == nullptr)
    GYRE_REQUIRE(ctx_B_ == B && ctx_L_ == L, "unet_forward: no context given and the bound one is [%d, %d], need [%d, %d]",
                 ctx_B_, ctx_L_, B, L);
  const int nl = cfg_.num_levels;
  GYRE_REQUIRE(H % (1 << (nl - 1)) == 0 && W % (1 << (nl - 1)) == 0,
               "unet_forward: latent %dx%d must be divisible by %d", H, W, 1 << (nl - 1));
  const int* ch = cfg_.block_out_channels;
  const float eps = cfg_.norm_eps;
  if (!ex.dry) GYRE_TRY(ensure_device());
  if (!ex.dry && ln_fuse_ && ln_dirty_) GYRE_TRY(refold_layernorms(ex.st));

  // ---- time embedding: sinusoid -> Linear -> SiLU -> Linear ; every consumer applies SiLU first, so the
  // second Linear's epilogue applies it once; then ONE fused GEMM produces all 22 resnet projections.
  __half* te0 = ex.p16(static_cast<size_t>(B) * ch[0]);
  RUN(ex, timestep_embed(t, B, ch[0], te0, ex.st));
  __half* te1 = ex.p16(static_cast<size_t>(B) * temb_dim_);
  RUN(ex, gemm_f16(te0, ch[0], time1_.w, ch[0], B, temb_dim_, ch[0],
                   ep_out(te1, temb_dim_, time1_.bias, nullptr, 0, ACT_SILU), ex.st));
  __half* te2 = ex.p16(static_cast<size_t>(B) * temb_dim_);
  if (cfg_.addition_embed_dim > 0) {
    // text_time conditioning: emb = time_embedding(t) + add_embedding(add_cond); SiLU comes after the sum
    const int ad = cfg_.addition_embed_dim;
    __half* traw = ex.p16(static_cast<size_t>(B) * temb_dim_);
    RUN(ex, gemm_f16(te1, temb_dim_, time2_.w, temb_dim_, B, temb_dim_, temb_dim_, ep_out(traw, temb_dim_, time2_.bias),
                     ex.st));
    __half* a1 = ex.p16(static_cast<size_t>(B) * temb_dim_);
    RUN(ex, gemm_f16(add_cond, ad, add1_.w, ad, B, temb_dim_, ad, ep_out(a1, temb_dim_, add1_.bias, nullptr, 0, ACT_SILU),
                     ex.st));
    __half* esum = ex.p16(static_cast<size_t>(B) * temb_dim_);
    RUN(ex, gemm_f16(a1, temb_dim_, add2_.w, temb_dim_, B, temb_dim_, temb_dim_,
                     ep_out(esum, temb_dim_, add2_.bias, traw, temb_dim_), ex.st));
    RUN(ex, silu_f16(esum, static_cast<int64_t>(B) * temb_dim_, te2, ex.st));
  } else {
    RUN(ex, gemm_f16(te1, temb_dim_, time2_.w, temb_dim_, B, temb_dim_, temb_dim_,
                     ep_out(te2, temb_dim_, time2_.bias, nullptr, 0, ACT_SILU), ex.st));
  }
  __half* temb_all = ex.p16(static_cast<size_t>(B) * temb_total_);
  RUN(ex, gemm_f16(te2, temb_dim_, temb_proj_.w, temb_dim_, B, temb_total_, temb_dim_,
                   ep_out(temb_all, temb_total_, temb_proj_.bias), ex.st));

  // ---- conv_in
  int h_ = H, w_ = W;
  // conv_in on the tensor cores too: the 4/9-channel latent is staged NHWC with a zero-padded channel pitch of
  // 8/16 (TMA needs 16-byte pixel rows); the padded K slice costs nothing measurable
  const int cin8 = (cfg_.in_channels + 7) & ~7;
  // CFG-parallel batches ([x ; x], same timesteps: set_cfg_duplicate): conv_in, the first resnet and the first
  // transformer up to its cross-attention see identical rows in both halves - they run ONCE, on the first half.  The
  // text_time conditioning of SDXL-style models differs between the halves (it enters through the time embedding), a
  // ControlNet's conditioning image may too: no sharing there.  The buffers keep their full-batch size either way.
  const bool share = cfg_dup_ && tunable(TUNE_CFG_SHARE) != 0 && B % 2 == 0 && cfg_.addition_embed_dim == 0 &&
                     !cfg_.controlnet && cfg_.attn_levels[0] && cfg_.layers_per_block >= 1;
  const int Bs = share ? B / 2 : B;
  __half* x_nhwc = ex.p16(static_cast<size_t>(B) * H * W * cin8);
  RUN(ex, nchw_to_nhwc_f16(sample, Bs, cfg_.in_channels, H, W, x_nhwc, cin8, ex.st));
  __half* hcur = ex.p16(static_cast<size_t>(B) * H * W * ch[0]);
  const __half* cond_emb = nullptr;
  if (cfg_.controlnet) {
    // ControlNetConditioningEmbedding.forward (controlnet/models.py:84-94) on the tensor-core conv kernel: the image is
    // staged NHWC with its channel pitch padded to 8, every conv but the last applies SiLU in its epilogue, every second
    // block conv has stride 2 (8H -> H); the result is added to conv_in's output through that conv's residual input
    const int cc = cfg_.conditioning_channels > 0 ? cfg_.conditioning_channels : 3;
    const int cc8 = (cc + 7) & ~7;
    int eh = 8 * H, ew = 8 * W;
    __half* e0 = ex.p16(static_cast<size_t>(B) * eh * ew * cc8);
    RUN(ex, nchw_to_nhwc_f16(cn ? cn->cond : nullptr, B, cc, eh, ew, e0, cc8, ex.st));
    const __half* ecur = e0;
    int ecin = cc8;
    for (int i = 0; i < 8; ++i) {
      const Conv3W& c = cn_embed_[i];
      const int stride = (i >= 1 && i <= 6 && (i % 2 == 0)) ? 2 : 1;       // blocks.1 / .3 / .5
      const int oh = stride == 2 ? (eh - 1) / 2 + 1 : eh, ow = stride == 2 ? (ew - 1) / 2 + 1 : ew;
      __half* eo = ex.p16(static_cast<size_t>(B) * oh * ow * c.Cout);
      RUN(ex, conv3x3_f16(ecur, ecin, B, eh, ew, ecin, c.wp, c.Cout, stride, 1,
                          ep_out(eo, c.Cout, c.bias, nullptr, 0, i < 7 ? ACT_SILU : ACT_NONE), ex.st));
      ecur = eo;
      ecin = c.Cout;
      eh = oh;
      ew = ow;
    }
    GYRE_REQUIRE(eh == H && ew == W, "controlnet: the conditioning image must be 8x the latent size");
    cond_emb = ecur;
  }
  {
    Epilogue e = ep_out(hcur, ch[0], conv_in_.bias, cond_emb, cond_emb ? ch[0] : 0);
    gn_offer(e, Bs, H, W, ch[0], 1, 1);      // the first resnet's norm1
    RUN(ex, conv3x3_f16(x_nhwc, cin8, Bs, H, W, cin8, conv_in_.wp, ch[0], 1, 1, e, ex.st));
  }
  if (share && !ex.dry) {
    // conv_in's output is also the first skip tensor, which the up path reads at the full batch: duplicate it
    const size_t half = static_cast<size_t>(Bs) * H * W * ch[0];
    GYRE_CHECK_CUDA(cudaMemcpyAsync(hcur + half, hcur, half * sizeof(__half), cudaMemcpyDeviceToDevice, ex.st));
  }

  struct Skip { __half* p; int C; int hw; };
  std::vector<Skip> skips;
  skips.push_back({hcur, ch[0], H * W});
  size_t ri = 0, ti = 0, di = 0, ui = 0;
  int ccur = ch[0];
  auto r_of = [&](size_t idx) { return tome_r ? tome_r[idx] : 0; };

  for (int i = 0; i < nl; ++i) {
    for (int j = 0; j < cfg_.layers_per_block; ++j) {
      ex.reset_scratch();
      // the very first resnet + transformer of a shared CFG batch work on the first half (the time-embedding rows of the
      // two halves are equal too); the transformer fans out to the full batch at its cross-attention
      const bool sh = share && i == 0 && j == 0;
      __half* o = ex.p16(static_cast<size_t>(B) * h_ * w_ * ch[i]);
      GYRE_TRY(resnet(ex, resnets_[ri++], hcur, ccur, nullptr, 0, sh ? Bs : B, h_, w_, eps, temb_all, temb_total_, o));
      hcur = o;
      ccur = ch[i];
      if (cfg_.attn_levels[i]) {
        ex.reset_scratch();
        __half* o2 = ex.p16(static_cast<size_t>(B) * h_ * w_ * ccur);
        GYRE_TRY(transformer(ex, tblocks_[ti], hcur, B, h_ * w_, ctx, L, r_of(ti), o2, sh));
        ++ti;
        hcur = o2;
      }
      skips.push_back({hcur, ccur, h_ * w_});
    }
    // T2I-adapter: the block's last hidden state (== its last skip tensor) takes the level's state in place
    if (!ex.dry && !adapter_.empty()) {
      RUN(ex, add_nchw_to_nhwc_f16(hcur, adapter_[i], B, ccur, h_ * w_, ex.st));
      gn_forget();   // the tensor a convolution left statistics for has just changed
    }
    if (i < nl - 1) {
      ex.reset_scratch();
      const int ho = (h_ - 1) / 2 + 1, wo = (w_ - 1) / 2 + 1;
      __half* o = ex.p16(static_cast<size_t>(B) * ho * wo * ccur);
      {
        Epilogue e = ep_out(o, ccur, downs_[di].bias);
        gn_offer(e, B, h_, w_, ccur, 2, 1);   // the next level's first resnet
        RUN(ex, conv3x3_f16(hcur, ccur, B, h_, w_, ccur, downs_[di].wp, ccur, 2, 1, e, ex.st));
      }
      ++di;
      hcur = o;
      h_ = ho;
      w_ = wo;
      skips.push_back({hcur, ccur, h_ * w_});
    }
  }
  // ---- ControlNet residuals: skip_i += residual_i once the down path no longer reads the skip tensors
  if (!ex.dry && !ctrl_down_.empty()) {
    GYRE_REQUIRE(ctrl_down_.size() == skips.size(), "unet_forward: %zu control residuals for %zu skips", ctrl_down_.size(),
                 skips.size());
    for (size_t k = 0; k < skips.size(); ++k) {
      // the last skip is also the mid block's input and diffusers adds the residual to the skip only: done after mid
      if (k + 1 == skips.size()) continue;
      RUN(ex, add_nchw_to_nhwc_f16(skips[k].p, ctrl_down_[k], B, skips[k].C, skips[k].hw, ex.st));
    }
    gn_forget();
  }
  // ---- mid
  {
    ex.reset_scratch();
    __half* o = ex.p16(static_cast<size_t>(B) * h_ * w_ * ccur);
    GYRE_TRY(resnet(ex, resnets_[ri++], hcur, ccur, nullptr, 0, B, h_, w_, eps, temb_all, temb_total_, o));
    ex.reset_scratch();
    __half* o2 = ex.p16(static_cast<size_t>(B) * h_ * w_ * ccur);
    GYRE_TRY(transformer(ex, tblocks_[ti], o, B, h_ * w_, ctx, L, r_of(ti), o2));
    ++ti;
    ex.reset_scratch();
    __half* o3 = ex.p16(static_cast<size_t>(B) * h_ * w_ * ccur);
    GYRE_TRY(resnet(ex, resnets_[ri++], o2, ccur, nullptr, 0, B, h_, w_, eps, temb_all, temb_total_, o3));
    hcur = o3;
    if (!ex.dry && ctrl_mid_ != nullptr) {
      RUN(ex, add_nchw_to_nhwc_f16(o3, ctrl_mid_, B, ccur, h_ * w_, ex.st));
      gn_forget();
    }
  }
  if (!ex.dry && !ctrl_down_.empty()) {
    // the deepest skip was the mid block's input: it takes its residual only now that the mid block has read it
    const Skip& ls = skips.back();
    RUN(ex, add_nchw_to_nhwc_f16(ls.p, ctrl_down_.back(), B, ls.C, ls.hw, ex.st));
    gn_forget();
  }
  // residuals are per call (core.py passes them with every UNet invocation)
  if (!ex.dry) {
    ctrl_down_.clear();
    ctrl_mid_ = nullptr;
    adapter_.clear();
  }
  if (cfg_.controlnet) {
    // ---- zero convolutions (controlnet/models.py:519-535): one 1x1 conv per skip tensor and one for the mid output,
    // written back in the NCHW layout gyre_b200_unet_set_control_residuals takes
    GYRE_REQUIRE(ex.dry || (cn->n_down == static_cast<int>(skips.size()) && cn->down_out && cn->mid_out),
                 "controlnet_forward: %d output tensors given, the model has %zu skips", cn ? cn->n_down : 0, skips.size());
    for (size_t k = 0; k <= skips.size(); ++k) {
      const bool mid = k == skips.size();
      const __half* src = mid ? hcur : skips[k].p;
      const int C = mid ? ccur : skips[k].C;
      const int hw = mid ? h_ * w_ : skips[k].hw;
      const LinW& z = mid ? cn_mid_ : cn_down_[k];
      ex.reset_scratch();
      __half* tmp = ex.s16(static_cast<size_t>(B) * hw * C);
      RUN(ex, gemm_f16(src, C, z.w, C, B * hw, C, C, ep_out(tmp, C, z.bias), ex.st));
      // [B, hw, C] -> [B, C, hw]: the spatial extent only matters as a product here
      RUN(ex, nhwc_to_nchw_f16(tmp, C, B, C, hw, 1, mid ? cn->mid_out : cn->down_out[k], ex.st));
    }
    return 0;
  }
  // ---- up
  for (int i = 0; i < nl; ++i) {
    const int lvl = nl - 1 - i;
    const int c = ch[lvl];
    for (int j = 0; j < cfg_.layers_per_block + 1; ++j) {
      const Skip s = skips.back();
      skips.pop_back();
      ex.reset_scratch();
      __half* o = ex.p16(static_cast<size_t>(B) * h_ * w_ * c);
      GYRE_TRY(resnet(ex, resnets_[ri++], hcur, ccur, s.p, s.C, B, h_, w_, eps, temb_all, temb_total_, o));
      hcur = o;
      ccur = c;
      if (cfg_.attn_levels[lvl]) {
        ex.reset_scratch();
        __half* o2 = ex.p16(static_cast<size_t>(B) * h_ * w_ * c);
        GYRE_TRY(transformer(ex, tblocks_[ti], hcur, B, h_ * w_, ctx, L, r_of(ti), o2));
        ++ti;
        hcur = o2;
      }
    }
    if (i < nl - 1) {
      ex.reset_scratch();
      __half* o = ex.p16(static_cast<size_t>(B) * 4 * h_ * w_ * c);
      GYRE_TRY(upsample_conv(ex, ups_[ui], hcur, B, h_, w_, o));
      h_ *= 2;
      w_ *= 2;
      ++ui;
      hcur = o;
    }
  }
  // ---- out
  ex.reset_scratch();
  {
    const size_t rows = static_cast<size_t>(B) * h_ * w_;
    __half* tn = ex.s16(rows * ccur);
    GYRE_TRY(gnorm(ex, hcur, ccur, nullptr, 0, B, h_ * w_, eps, norm_out_, true, tn));
    const int oc = cfg_.out_channels;
    __half* eps_nhwc = ex.s16(rows * oc);
    RUN(ex, conv3x3_f16(tn, ccur, B, h_, w_, ccur, conv_out_.wp, oc, 1, 1, ep_out(eps_nhwc, oc, conv_out_.bias),
                        ex.st));
    RUN(ex, nhwc_to_nchw_f16(eps_nhwc, oc, B, oc, h_, w_, out, ex.st));
  }
  EX_CHECK(ex);
  return 0;
}

// ------------------------------------------------------------------------------------------ VAE
VAEModel::VAEModel(const gyre_b200_vae_config& cfg) : cfg_(cfg) {
  groups_ = cfg.norm_num_groups;
  const int L = cfg.num_levels;
  const int* ch = cfg.block_out_channels;
  const int z = cfg.latent_channels;
  const int top = ch[L - 1];
  dec_res_.reserve(32);
  enc_res_.reserve(32);
  dec_ups_.reserve(8);
  enc_downs_.reserve(8);

  auto reg_attn = [&](const std::string& p, int C, VaeAttnW* a) {
    a->C = C;
    reg_norm(p + ".group_norm", C, &a->gn);
    a->qk.N = 2 * C;
    a->qk.K = C;
    a->qk.w = static_cast<__half*>(dalloc(sizeof(__half) * 2 * C * C));
    a->qk.bias = static_cast<float*>(dalloc(sizeof(float) * 2 * C));
    reg(p + ".query.weight", P_LINEAR, a->qk.w, {C, C}, C);
    reg(p + ".key.weight", P_LINEAR, a->qk.w ? a->qk.w + static_cast<size_t>(C) * C : nullptr, {C, C}, C);
    reg(p + ".query.bias", P_F32, a->qk.bias, {C});
    reg(p + ".key.bias", P_F32, a->qk.bias ? a->qk.bias + C : nullptr, {C});
    a->v.N = C;
    a->v.K = C;
    a->v.w = static_cast<__half*>(dalloc(sizeof(__half) * C * C));
    a->v_bias = static_cast<float*>(dalloc(sizeof(float) * C));
    reg(p + ".value.weight", P_LINEAR, a->v.w, {C, C}, C);
    reg(p + ".value.bias", P_F32, a->v_bias, {C});
    reg_linear(p + ".proj_attn", C, C, true, &a->proj);
  };

  // ---- decoder
  post_quant_.Cin = z;
  post_quant_.Cout = z;
  post_quant_.w = static_cast<float*>(dalloc(sizeof(float) * z * z));
  post_quant_.bias = static_cast<float*>(dalloc(sizeof(float) * z));
  reg("post_quant_conv.weight", P_F32MAT, post_quant_.w, {z, z});
  reg("post_quant_conv.bias", P_F32, post_quant_.bias, {z});
  reg_conv3("decoder.conv_in", z, top, &dec_conv_in_);
  reg_resnet("decoder.mid_block.resnets.0", top, top, false, &dec_mid_[0]);
  reg_attn("decoder.mid_block.attentions.0", top, &dec_attn_);
  reg_resnet("decoder.mid_block.resnets.1", top, top, false, &dec_mid_[1]);
  int cin = top;
  for (int i = 0; i < L; ++i) {
    const int c = ch[L - 1 - i];
    for (int j = 0; j < cfg.layers_per_block + 1; ++j) {
      dec_res_.emplace_back();
      reg_resnet("decoder.up_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), cin, c, false,
                 &dec_res_.back());
      cin = c;
    }
    if (i < L - 1) {
      dec_ups_.emplace_back();
      reg_conv3("decoder.up_blocks." + std::to_string(i) + ".upsamplers.0.conv", c, c, &dec_ups_.back(), true);
    }
  }
  reg_norm("decoder.conv_norm_out", ch[0], &dec_norm_out_);
  reg_conv3("decoder.conv_out", ch[0], cfg.out_channels, &dec_conv_out_);

  // ---- encoder
  reg_conv3("encoder.conv_in", cfg.in_channels, ch[0], &enc_conv_in_);
  cin = ch[0];
  for (int i = 0; i < L; ++i) {
    for (int j = 0; j < cfg.layers_per_block; ++j) {
      enc_res_.emplace_back();
      reg_resnet("encoder.down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), cin, ch[i], false,
                 &enc_res_.back());
      cin = ch[i];
    }
    if (i < L - 1) {
      enc_downs_.emplace_back();
      reg_conv3("encoder.down_blocks." + std::to_string(i) + ".downsamplers.0.conv", ch[i], ch[i], &enc_downs_.back());
    }
  }
  reg_resnet("encoder.mid_block.resnets.0", cin, cin, false, &enc_mid_[0]);
  reg_attn("encoder.mid_block.attentions.0", cin, &enc_attn_);
  reg_resnet("encoder.mid_block.resnets.1", cin, cin, false, &enc_mid_[1]);
  reg_norm("encoder.conv_norm_out", cin, &enc_norm_out_);
  reg_conv3("encoder.conv_out", cin, 2 * z, &enc_conv_out_);
  quant_.Cin = 2 * z;
  quant_.Cout = 2 * z;
  quant_.w = static_cast<float*>(dalloc(sizeof(float) * 4 * z * z));
  quant_.bias = static_cast<float*>(dalloc(sizeof(float) * 2 * z));
  reg("quant_conv.weight", P_F32MAT, quant_.w, {2 * z, 2 * z});
  reg("quant_conv.bias", P_F32, quant_.bias, {2 * z});
}

// Legacy single-head AttentionBlock: d = C = 512 exceeds the flash kernel's TMEM budget, so
// the scores are materialised: S = Q K^T (fp32) -> row softmax -> P (fp16) -> O = P V.  V^T is produced
// directly by a GEMM (W_v X^T); its bias is added after P.V because softmax rows sum to one.
int VAEModel::attn(Exec& ex, const VaeAttnW& a, const __half* x, int B, int HW, __half* out) {
  const int C = a.C;
  const int M = B * HW;
  const size_t n = static_cast<size_t>(M) * C;
  __half* tn = ex.s16(n);
  GYRE_TRY(gnorm(ex, x, C, nullptr, 0, B, HW, 1e-6f, a.gn, false, tn));
  __half* qk = ex.s16(n * 2);
  RUN(ex, gemm_f16(tn, C, a.qk.w, C, M, 2 * C, C, ep_out(qk, 2 * C, a.qk.bias), ex.st));
  __half* o = ex.s16(n);
  __half* vt = ex.s16(static_cast<size_t>(C) * HW);
  // query rows per pass: the fp32 scores + fp16 probabilities of one pass stay under ~0.75 GB however large the image is
  // (HW = 4096 at 512 x 512: one pass; 65536 at 2048 x 2048: 2048 rows per pass instead of a 26 GB score matrix)
  int qc = HW;
  {
    const size_t budget_rows = (size_t(768) << 20) / (static_cast<size_t>(HW) * 6);
    if (static_cast<size_t>(qc) > budget_rows) qc = static_cast<int>(budget_rows < 128 ? 128 : (budget_rows & ~size_t(127)));
  }
  float* s = ex.s32(static_cast<size_t>(qc) * HW);
  __half* p = ex.s16(static_cast<size_t>(qc) * HW);
  const float scale = 1.0f / sqrtf(static_cast<float>(C));
  for (int b = 0; b < B; ++b) {
    const __half* tb = off(tn, static_cast<size_t>(b) * HW * C);
    const __half* qb = off(qk, static_cast<size_t>(b) * HW * 2 * C);
    RUN(ex, gemm_f16(a.v.w, C, tb, C, C, HW, C, ep_out(vt, HW), ex.st));                      // V^T [C, HW]
    for (int q0 = 0; q0 < HW; q0 += qc) {
      const int nq = HW - q0 < qc ? HW - q0 : qc;
      {
        Epilogue e;
        e.out = s;
        e.ldo = HW;
        e.out_mode = OUT_F32;
        RUN(ex, gemm_f16(off(qb, static_cast<size_t>(q0) * 2 * C), 2 * C, off(qb, C), 2 * C, nq, HW, C, e, ex.st));   // S = Q K^T
      }
      RUN(ex, softmax_rows_f32(s, nq, HW, scale, p, HW, ex.st));
      RUN(ex, gemm_f16(p, HW, vt, HW, nq, C, HW,
                       ep_out(off(o, (static_cast<size_t>(b) * HW + q0) * C), C, a.v_bias), ex.st));
    }
  }
  RUN(ex, gemm_f16(o, C, a.proj.w, C, M, C, C, ep_out(out, C, a.proj.bias, x, C), ex.st));
  return 0;
}

int VAEModel::decode(Exec& ex, const __half* z, int B, int h, int w, bool postprocess, __half* img, uint8_t* img_u8) {
  GYRE_REQUIRE(B > 0 && h > 0 && w > 0, "vae_decode: empty problem");
  GYRE_REQUIRE(img != nullptr || img_u8 != nullptr, "vae_decode: no output requested");
  if (!ex.dry) GYRE_TRY(ensure_device());
  const int L = cfg_.num_levels;
  const int* ch = cfg_.block_out_channels;
  const int zc = cfg_.latent_channels;
  const int top = ch[L - 1];
  const float eps = 1e-6f;
  const size_t px = static_cast<size_t>(B) * h * w;
  // block outputs are simply stacked in the persistent region (the decoder is a chain; the dry run sizes
  // the workspace, a few GB at batch 8 / 512x512 -- small against 180 GB of HBM)
  __half* z_nhwc = ex.p16(px * zc);
  RUN(ex, nchw_to_nhwc_f16(z, B, zc, h, w, z_nhwc, zc, ex.st));
  const int zc8 = (zc + 7) & ~7;
  __half* z2 = ex.p16(px * zc8);
  RUN(ex, conv1x1_small(z_nhwc, static_cast<int64_t>(px), zc, post_quant_.w, post_quant_.bias, zc, z2, zc8, ex.st));
  __half* cur = ex.p16(px * top);
  gn_forget();
  {
    Epilogue e = ep_out(cur, top, dec_conv_in_.bias);
    gn_offer(e, B, h, w, top, 1, 1);
    RUN(ex, conv3x3_f16(z2, zc8, B, h, w, zc8, dec_conv_in_.wp, top, 1, 1, e, ex.st));
  }
  int H = h, W = w, C = top;
  auto block_out = [&](size_t elems) { return ex.p16(elems); };
  {
    ex.reset_scratch();
    __half* o = block_out(px * top);
    GYRE_TRY(resnet(ex, dec_mid_[0], cur, C, nullptr, 0, B, H, W, eps, nullptr, 0, o));
    ex.reset_scratch();
    __half* o2 = block_out(px * top);
    GYRE_TRY(attn(ex, dec_attn_, o, B, H * W, o2));
    ex.reset_scratch();
    __half* o3 = block_out(px * top);
    GYRE_TRY(resnet(ex, dec_mid_[1], o2, C, nullptr, 0, B, H, W, eps, nullptr, 0, o3));
    cur = o3;
  }
  size_t ri = 0;
  for (int i = 0; i < L; ++i) {
    const int c = ch[L - 1 - i];
    for (int j = 0; j < cfg_.layers_per_block + 1; ++j) {
      ex.reset_scratch();
      __half* o = block_out(static_cast<size_t>(B) * H * W * c);
      GYRE_TRY(resnet(ex, dec_res_[ri++], cur, C, nullptr, 0, B, H, W, eps, nullptr, 0, o));
      cur = o;
      C = c;
    }
    if (i < L - 1) {
      ex.reset_scratch();
      __half* o = block_out(static_cast<size_t>(B) * 4 * H * W * c);
      GYRE_TRY(upsample_conv(ex, dec_ups_[i], cur, B, H, W, o));
      H *= 2;
      W *= 2;
      cur = o;
    }
  }
  ex.reset_scratch();
  {
    const size_t rows = static_cast<size_t>(B) * H * W;
    __half* tn = ex.s16(rows * C);
    GYRE_TRY(gnorm(ex, cur, C, nullptr, 0, B, H * W, eps, dec_norm_out_, true, tn));
    const int oc = cfg_.out_channels;
    GYRE_REQUIRE(oc == 3, "vae_decode: out_channels must be 3");
    __half* rgb = ex.s16(rows * 4);
    RUN(ex, conv3x3_f16(tn, C, B, H, W, C, dec_conv_out_.wp, oc, 1, 1, ep_out(rgb, 4, dec_conv_out_.bias), ex.st));
    RUN(ex, vae_tail(rgb, 4, B, H, W, postprocess, img, img_u8, ex.st));
  }
  EX_CHECK(ex);
  return 0;
}

int VAEModel::encode(Exec& ex, const __half* img, int B, int H, int W, __half* moments) {
  GYRE_REQUIRE(B > 0 && H > 0 && W > 0, "vae_encode: empty problem");
  const int L = cfg_.num_levels;
  GYRE_REQUIRE(H % (1 << (L - 1)) == 0 && W % (1 << (L - 1)) == 0, "vae_encode: image %dx%d not divisible by %d", H, W,
               1 << (L - 1));
  if (!ex.dry) GYRE_TRY(ensure_device());
  const int* ch = cfg_.block_out_channels;
  const int zc = cfg_.latent_channels;
  const int ic = cfg_.in_channels;
  const float eps = 1e-6f;
  const int ic8 = (ic + 7) & ~7;
  __half* x = ex.p16(static_cast<size_t>(B) * H * W * ic8);
  RUN(ex, nchw_to_nhwc_f16(img, B, ic, H, W, x, ic8, ex.st));
  __half* cur = ex.p16(static_cast<size_t>(B) * H * W * ch[0]);
  gn_forget();
  {
    Epilogue e = ep_out(cur, ch[0], enc_conv_in_.bias);
    gn_offer(e, B, H, W, ch[0], 1, 1);
    RUN(ex, conv3x3_f16(x, ic8, B, H, W, ic8, enc_conv_in_.wp, ch[0], 1, 1, e, ex.st));
  }
  int C = ch[0];
  size_t ri = 0;
  for (int i = 0; i < L; ++i) {
    for (int j = 0; j < cfg_.layers_per_block; ++j) {
      ex.reset_scratch();
      __half* o = ex.p16(static_cast<size_t>(B) * H * W * ch[i]);
      GYRE_TRY(resnet(ex, enc_res_[ri++], cur, C, nullptr, 0, B, H, W, eps, nullptr, 0, o));
      cur = o;
      C = ch[i];
    }
    if (i < L - 1) {
      ex.reset_scratch();
      // Downsample2D(padding=0): F.pad(x, (0,1,0,1)) + conv3x3 stride 2 -- the asymmetric zero pad is the
      // TMA out-of-bounds fill on the right/bottom edge
      const int ho = (H + 1 - 3) / 2 + 1, wo = (W + 1 - 3) / 2 + 1;
      __half* o = ex.p16(static_cast<size_t>(B) * ho * wo * C);
      {
        Epilogue e = ep_out(o, C, enc_downs_[i].bias);
        gn_offer(e, B, H, W, C, 2, 0);
        RUN(ex, conv3x3_f16(cur, C, B, H, W, C, enc_downs_[i].wp, C, 2, 0, e, ex.st));
      }
      cur = o;
      H = ho;
      W = wo;
    }
  }
  {
    ex.reset_scratch();
    __half* o = ex.p16(static_cast<size_t>(B) * H * W * C);
    GYRE_TRY(resnet(ex, enc_mid_[0], cur, C, nullptr, 0, B, H, W, eps, nullptr, 0, o));
    ex.reset_scratch();
    __half* o2 = ex.p16(static_cast<size_t>(B) * H * W * C);
    GYRE_TRY(attn(ex, enc_attn_, o, B, H * W, o2));
    ex.reset_scratch();
    __half* o3 = ex.p16(static_cast<size_t>(B) * H * W * C);
    GYRE_TRY(resnet(ex, enc_mid_[1], o2, C, nullptr, 0, B, H, W, eps, nullptr, 0, o3));
    cur = o3;
  }
  ex.reset_scratch();
  {
    const size_t rows = static_cast<size_t>(B) * H * W;
    __half* tn = ex.s16(rows * C);
    GYRE_TRY(gnorm(ex, cur, C, nullptr, 0, B, H * W, eps, enc_norm_out_, true, tn));
    __half* m0 = ex.s16(rows * 2 * zc);
    RUN(ex, conv3x3_f16(tn, C, B, H, W, C, enc_conv_out_.wp, 2 * zc, 1, 1, ep_out(m0, 2 * zc, enc_conv_out_.bias),
                        ex.st));
    __half* m1 = ex.s16(rows * 2 * zc);
    RUN(ex, conv1x1_small(m0, static_cast<int64_t>(rows), 2 * zc, quant_.w, quant_.bias, 2 * zc, m1, 2 * zc, ex.st));
    RUN(ex, nhwc_to_nchw_f16(m1, 2 * zc, B, 2 * zc, H, W, moments, ex.st));
  }
  EX_CHECK(ex);
  return 0;
}

// ------------------------------------------------------------------------------------------ CLIP text encoder
// transformers CLIPTextModel (CLIPTextTransformer): embeddings -> num_layers x [LN1 -> causal self-attention
// (q, k, v, out projections with bias) -> +res ; LN2 -> fc1 -> quick_gelu -> fc2 -> +res] -> final LayerNorm.
ClipTextModel::ClipTextModel(const gyre_b200_clip_config& cfg) : cfg_(cfg) {
  const int C = cfg.hidden_size, F = cfg.intermediate_size;
  tok_emb_ = static_cast<__half*>(dalloc(sizeof(__half) * static_cast<size_t>(cfg.vocab_size) * C));
  pos_emb_ = static_cast<__half*>(dalloc(sizeof(__half) * static_cast<size_t>(cfg.max_positions) * C));
  reg("text_model.embeddings.token_embedding.weight", P_LINEAR, tok_emb_, {cfg.vocab_size, C}, C);
  reg("text_model.embeddings.position_embedding.weight", P_LINEAR, pos_emb_, {cfg.max_positions, C}, C);
  layers_.resize(cfg.num_layers);
  for (int i = 0; i < cfg.num_layers; ++i) {
    ClipLayerW* l = &layers_[i];
    const std::string p = "text_model.encoder.layers." + std::to_string(i);
    reg_norm(p + ".layer_norm1", C, &l->ln1);
    reg_norm(p + ".layer_norm2", C, &l->ln2);
    // fused [q ; k ; v] projection with its fused bias
    l->qkv.N = 3 * C;
    l->qkv.K = C;
    l->qkv.w = static_cast<__half*>(dalloc(sizeof(__half) * 3 * C * C));
    l->qkv.bias = static_cast<float*>(dalloc(sizeof(float) * 3 * C));
    const char* names[3] = {"q_proj", "k_proj", "v_proj"};
    for (int j = 0; j < 3; ++j) {
      reg(p + ".self_attn." + names[j] + ".weight", P_LINEAR, l->qkv.w ? l->qkv.w + static_cast<size_t>(j) * C * C : nullptr,
          {C, C}, C);
      reg(p + ".self_attn." + names[j] + ".bias", P_F32, l->qkv.bias ? l->qkv.bias + j * C : nullptr, {C});
    }
    reg_linear(p + ".self_attn.out_proj", C, C, true, &l->out);
    reg_linear(p + ".mlp.fc1", F, C, true, &l->fc1);
    reg_linear(p + ".mlp.fc2", C, F, true, &l->fc2);
  }
  reg_norm("text_model.final_layer_norm", C, &final_ln_);
}

int ClipTextModel::forward(Exec& ex, const int64_t* ids, int B, int L, int skip_last, bool final_ln, __half* out) {
  GYRE_REQUIRE(B > 0 && L > 0 && L <= cfg_.max_positions && L <= 128, "clip_forward: sequence length %d (max %d)", L,
               cfg_.max_positions);
  GYRE_REQUIRE(skip_last >= 0 && skip_last <= cfg_.num_layers, "clip_forward: skip_last %d", skip_last);
  if (!ex.dry) GYRE_TRY(ensure_device());
  const int C = cfg_.hidden_size, F = cfg_.intermediate_size, H = cfg_.num_heads;
  const int d = C / H;
  const int M = B * L;
  const size_t n = static_cast<size_t>(M) * C;
  const float eps = cfg_.layer_norm_eps;
  const float scale = 1.0f / sqrtf(static_cast<float>(d));
  __half* h = ex.p16(n);
  __half* h2 = ex.p16(n);
  __half* nrm = ex.p16(n);
  __half* qkv = ex.p16(n * 3);
  __half* att = ex.p16(n);
  __half* mid = ex.p16(static_cast<size_t>(M) * F);
  RUN(ex, embed_tokens(ids, tok_emb_, pos_emb_, B, L, C, cfg_.vocab_size, h, ex.st));
  const int act = cfg_.hidden_act == 0 ? ACT_QUICKGELU : ACT_GELU;
  for (int i = 0; i < cfg_.num_layers - skip_last; ++i) {
    const ClipLayerW& l = layers_[i];
    RUN(ex, layernorm_rows(h, M, C, eps, l.ln1.g, l.ln1.b, nrm, ex.st));
    RUN(ex, gemm_f16(nrm, C, l.qkv.w, C, M, 3 * C, C, ep_out(qkv, 3 * C, l.qkv.bias), ex.st));
    RUN(ex, causal_attention_short(qkv, B, L, H, d, scale, att, ex.st));
    RUN(ex, gemm_f16(att, C, l.out.w, C, M, C, C, ep_out(h2, C, l.out.bias, h, C), ex.st));          // h2 = h + attn
    RUN(ex, layernorm_rows(h2, M, C, eps, l.ln2.g, l.ln2.b, nrm, ex.st));
    RUN(ex, gemm_f16(nrm, C, l.fc1.w, C, M, F, C, ep_out(mid, F, l.fc1.bias, nullptr, 0, act), ex.st));
    RUN(ex, gemm_f16(mid, F, l.fc2.w, F, M, C, F, ep_out(h, C, l.fc2.bias, h2, C), ex.st));           // h = h2 + mlp
  }
  if (final_ln) {
    RUN(ex, layernorm_rows(h, M, C, eps, final_ln_.g, final_ln_.b, out, ex.st));
  } else if (!ex.dry) {
    GYRE_CHECK_CUDA(cudaMemcpyAsync(out, h, n * sizeof(__half), cudaMemcpyDeviceToDevice, ex.st));
  }
  EX_CHECK(ex);
  return 0;
}

// ------------------------------------------------------------------------------------------ CLIP vision tower
// transformers CLIPVisionModel (CLIPVisionTransformer): patch embedding (Conv2d(3, C, P, stride P) == a GEMM over
// flattened patches) + class token + position embedding -> pre-LayerNorm -> num_layers x [LN1 -> full self-attention ->
// +res ; LN2 -> fc1 -> quick_gelu -> fc2 -> +res] -> post-LayerNorm of the class token (pooled output) -> visual
// projection -> cosine similarity against the checker's concept embeddings (safety_checkers.py:32-37).
ClipVisionModel::ClipVisionModel(const gyre_b200_clip_vision_config& cfg) : cfg_(cfg) {
  const int C = cfg.hidden_size, F = cfg.intermediate_size, P = cfg.patch_size;
  const int ntok = (cfg.image_size / P) * (cfg.image_size / P) + 1;
  const int K = 3 * P * P;
  Kp_ = (K + 7) & ~7;
  patch_.N = C;
  patch_.K = Kp_;
  patch_.w = static_cast<__half*>(dalloc(sizeof(__half) * static_cast<size_t>(C) * Kp_));     // zero-filled: the pad stays 0
  reg("vision_model.embeddings.patch_embedding.weight", P_LINEAR, patch_.w, {C, K}, Kp_);
  class_emb_ = static_cast<float*>(dalloc(sizeof(float) * C));
  reg("vision_model.embeddings.class_embedding", P_F32, class_emb_, {C});
  pos_emb_ = static_cast<__half*>(dalloc(sizeof(__half) * static_cast<size_t>(ntok) * C));
  reg("vision_model.embeddings.position_embedding.weight", P_LINEAR, pos_emb_, {ntok, C}, C);
  reg_norm("vision_model.pre_layrnorm", C, &pre_ln_);          // the checkpoint key really is spelled like this
  reg_norm("vision_model.post_layernorm", C, &post_ln_);
  layers_.resize(cfg.num_layers);
  for (int i = 0; i < cfg.num_layers; ++i) {
    ClipLayerW* l = &layers_[i];
    const std::string p = "vision_model.encoder.layers." + std::to_string(i);
    reg_norm(p + ".layer_norm1", C, &l->ln1);
    reg_norm(p + ".layer_norm2", C, &l->ln2);
    l->qkv.N = 3 * C;
    l->qkv.K = C;
    l->qkv.w = static_cast<__half*>(dalloc(sizeof(__half) * 3 * C * C));
    l->qkv.bias = static_cast<float*>(dalloc(sizeof(float) * 3 * C));
    const char* names[3] = {"q_proj", "k_proj", "v_proj"};
    for (int j = 0; j < 3; ++j) {
      reg(p + ".self_attn." + names[j] + ".weight", P_LINEAR, l->qkv.w ? l->qkv.w + static_cast<size_t>(j) * C * C : nullptr,
          {C, C}, C);
      reg(p + ".self_attn." + names[j] + ".bias", P_F32, l->qkv.bias ? l->qkv.bias + j * C : nullptr, {C});
    }
    reg_linear(p + ".self_attn.out_proj", C, C, true, &l->out);
    reg_linear(p + ".mlp.fc1", F, C, true, &l->fc1);
    reg_linear(p + ".mlp.fc2", C, F, true, &l->fc2);
  }
  if (cfg.projection_dim > 0) reg_linear("visual_projection", cfg.projection_dim, C, false, &proj_);
  if (cfg.num_concepts <= 0 || cfg.projection_dim <= 0) return;       // tower only (style adapter's CLIP vision model)
  const int ne = cfg.num_special + cfg.num_concepts, D = cfg.projection_dim;
  embeds_ = static_cast<float*>(dalloc(sizeof(float) * static_cast<size_t>(ne) * D));
  thresholds_ = static_cast<float*>(dalloc(sizeof(float) * ne));
  reg("special_care_embeds", P_F32MAT, embeds_, {cfg.num_special, D});
  reg("concept_embeds", P_F32MAT, embeds_ ? embeds_ + static_cast<size_t>(cfg.num_special) * D : nullptr, {cfg.num_concepts, D});
  reg("special_care_embeds_weights", P_F32, thresholds_, {cfg.num_special});
  reg("concept_embeds_weights", P_F32, thresholds_ ? thresholds_ + cfg.num_special : nullptr, {cfg.num_concepts});
}

int ClipVisionModel::forward(Exec& ex, const __half* pixel_values, int B, __half* image_embeds, float* scores, __half* hidden,
                             int skip_last, bool hidden_only) {
  GYRE_REQUIRE(B > 0, "clip vision: empty batch");
  GYRE_REQUIRE(skip_last >= 0 && skip_last <= cfg_.num_layers, "clip vision: skip_last %d", skip_last);
  GYRE_REQUIRE(hidden_only || (cfg_.num_concepts > 0 && cfg_.projection_dim > 0),
               "safety_scores: this handle holds the vision tower only (no projection / concept embeddings)");
  if (!ex.dry) GYRE_TRY(ensure_device());
  const int C = cfg_.hidden_size, F = cfg_.intermediate_size, H = cfg_.num_heads, P = cfg_.patch_size, S = cfg_.image_size;
  const int np = (S / P) * (S / P), N = np + 1;
  const int d = C / H;
  const int M = B * N;
  const size_t n = static_cast<size_t>(M) * C;
  const float eps = cfg_.layer_norm_eps;
  const float scale = 1.0f / sqrtf(static_cast<float>(d));
  __half* patches_in = ex.p16(static_cast<size_t>(B) * np * Kp_);
  __half* patches = ex.p16(static_cast<size_t>(B) * np * C);
  __half* h = ex.p16(n);
  __half* h2 = ex.p16(n);
  __half* nrm = ex.p16(n);
  __half* qkv = ex.p16(n * 3);
  __half* att = ex.p16(n);
  __half* mid = ex.p16(static_cast<size_t>(M) * F);
  __half* cls = ex.p16(static_cast<size_t>(B) * C);
  __half* pooled = ex.p16(static_cast<size_t>(B) * C);
  __half* emb_local = ex.p16(static_cast<size_t>(B) * cfg_.projection_dim);
  RUN(ex, patchify(pixel_values, B, S, P, Kp_, patches_in, ex.st));
  RUN(ex, gemm_f16(patches_in, Kp_, patch_.w, Kp_, B * np, C, Kp_, ep_out(patches, C), ex.st));
  RUN(ex, vision_embed(patches, class_emb_, pos_emb_, B, N, C, h2, ex.st));
  RUN(ex, layernorm_rows(h2, M, C, eps, pre_ln_.g, pre_ln_.b, h, ex.st));
  const int act = cfg_.hidden_act == 0 ? ACT_QUICKGELU : ACT_GELU;
  for (int i = 0; i < cfg_.num_layers - (hidden_only ? skip_last : 0); ++i) {
    const ClipLayerW& l = layers_[i];
    RUN(ex, layernorm_rows(h, M, C, eps, l.ln1.g, l.ln1.b, nrm, ex.st));
    RUN(ex, gemm_f16(nrm, C, l.qkv.w, C, M, 3 * C, C, ep_out(qkv, 3 * C, l.qkv.bias), ex.st));
    RUN(ex, attention_f16(qkv, 3 * C, off(qkv, C), 3 * C, off(qkv, 2 * C), 3 * C, B, H, N, N, d, scale, att, C, ex.st));
    RUN(ex, gemm_f16(att, C, l.out.w, C, M, C, C, ep_out(h2, C, l.out.bias, h, C), ex.st));          // h2 = h + attn
    RUN(ex, layernorm_rows(h2, M, C, eps, l.ln2.g, l.ln2.b, nrm, ex.st));
    RUN(ex, gemm_f16(nrm, C, l.fc1.w, C, M, F, C, ep_out(mid, F, l.fc1.bias, nullptr, 0, act), ex.st));
    RUN(ex, gemm_f16(mid, F, l.fc2.w, F, M, C, F, ep_out(h, C, l.fc2.bias, h2, C), ex.st));           // h = h2 + mlp
  }
  if (hidden_only) {
    if (!ex.dry) GYRE_CHECK_CUDA(cudaMemcpyAsync(hidden, h, n * sizeof(__half), cudaMemcpyDeviceToDevice, ex.st));
    EX_CHECK(ex);
    return 0;
  }
  // pooled_output = post_layernorm(last_hidden_state[:, 0]): gather the class-token rows, normalise, project
  if (!ex.dry)
    GYRE_CHECK_CUDA(cudaMemcpy2DAsync(cls, static_cast<size_t>(C) * sizeof(__half), h, static_cast<size_t>(N) * C * sizeof(__half),
                                      static_cast<size_t>(C) * sizeof(__half), B, cudaMemcpyDeviceToDevice, ex.st));
  RUN(ex, layernorm_rows(cls, B, C, eps, post_ln_.g, post_ln_.b, pooled, ex.st));
  __half* emb_out = image_embeds ? image_embeds : emb_local;
  RUN(ex, gemm_f16(pooled, C, proj_.w, C, B, cfg_.projection_dim, C, ep_out(emb_out, cfg_.projection_dim), ex.st));
  if (scores || ex.dry)
    RUN(ex, cosine_scores(emb_out, B, cfg_.projection_dim, embeds_, cfg_.num_special + cfg_.num_concepts, scores, ex.st));
  EX_CHECK(ex);
  return 0;
}

// ------------------------------------------------------------------------------------------ style T2I-adapter
// StyleAdapter.forward (adapter.py:186-199): [x ; style_embedding] -> ln_pre -> n_layers x ResidualAttentionBlock
// (x + MHA(ln_1 x); x + c_proj(quick_gelu(c_fc(ln_2 x))), nn.MultiheadAttention = packed q|k|v in_proj + out_proj, attention
// across the tokens of one sample) -> ln_post of the last num_token tokens -> @ proj.
StyleAdapterModel::StyleAdapterModel(const gyre_b200_style_adapter_config& cfg) : cfg_(cfg) {
  const int D = cfg.width;
  layers_.resize(cfg.n_layers);
  for (int i = 0; i < cfg.n_layers; ++i) {
    ClipLayerW* l = &layers_[i];
    const std::string p = "transformer_layes." + std::to_string(i);          // (sic)
    reg_norm(p + ".ln_1", D, &l->ln1);
    reg_norm(p + ".ln_2", D, &l->ln2);
    l->qkv.N = 3 * D;
    l->qkv.K = D;
    l->qkv.w = static_cast<__half*>(dalloc(sizeof(__half) * 3 * static_cast<size_t>(D) * D));
    l->qkv.bias = static_cast<float*>(dalloc(sizeof(float) * 3 * D));
    reg(p + ".attn.in_proj_weight", P_LINEAR, l->qkv.w, {3 * D, D}, D);
    reg(p + ".attn.in_proj_bias", P_F32, l->qkv.bias, {3 * D});
    reg_linear(p + ".attn.out_proj", D, D, true, &l->out);
    reg_linear(p + ".mlp.c_fc", 4 * D, D, true, &l->fc1);
    reg_linear(p + ".mlp.c_proj", D, 4 * D, true, &l->fc2);
  }
  style_emb_ = static_cast<__half*>(dalloc(sizeof(__half) * static_cast<size_t>(cfg.num_token) * D));
  reg("style_embedding", P_LINEAR, style_emb_, {cfg.num_token, D}, D);
  reg_norm("ln_pre", D, &ln_pre_);
  reg_norm("ln_post", D, &ln_post_);
  proj_.N = cfg.context_dim;
  proj_.K = D;
  proj_.w = static_cast<__half*>(dalloc(sizeof(__half) * static_cast<size_t>(cfg.context_dim) * D));
  reg("proj", P_LINEAR, proj_.w, {cfg.context_dim, D}, D);
}

int StyleAdapterModel::forward(Exec& ex, const __half* x, int B, int L, __half* out) {
  GYRE_REQUIRE(B > 0 && L > 0, "style_adapter_forward: empty input");
  if (!ex.dry) GYRE_TRY(ensure_device());
  const int D = cfg_.width, T = cfg_.num_token, H = cfg_.num_head, F = 4 * D;
  const int N = L + T, M = B * N, d = D / H;
  const size_t n = static_cast<size_t>(M) * D;
  const float eps = 1e-5f;
  const float scale = 1.0f / sqrtf(static_cast<float>(d));
  __half* cat = ex.p16(n);
  __half* h = ex.p16(n);
  __half* h2 = ex.p16(n);
  __half* nrm = ex.p16(n);
  __half* qkv = ex.p16(n * 3);
  __half* att = ex.p16(n);
  __half* mid = ex.p16(static_cast<size_t>(M) * F);
  __half* tail = ex.p16(static_cast<size_t>(B) * T * D);
  __half* tail_n = ex.p16(static_cast<size_t>(B) * T * D);
  if (!ex.dry) {
    const size_t row = static_cast<size_t>(D) * sizeof(__half);
    GYRE_CHECK_CUDA(cudaMemcpy2DAsync(cat, N * row, x, L * row, L * row, B, cudaMemcpyDeviceToDevice, ex.st));
    for (int b = 0; b < B; ++b)
      GYRE_CHECK_CUDA(cudaMemcpyAsync(cat + (static_cast<size_t>(b) * N + L) * D, style_emb_, T * row, cudaMemcpyDeviceToDevice,
                                      ex.st));
  }
  RUN(ex, layernorm_rows(cat, M, D, eps, ln_pre_.g, ln_pre_.b, h, ex.st));
  for (int i = 0; i < cfg_.n_layers; ++i) {
    const ClipLayerW& l = layers_[i];
    RUN(ex, layernorm_rows(h, M, D, eps, l.ln1.g, l.ln1.b, nrm, ex.st));
    RUN(ex, gemm_f16(nrm, D, l.qkv.w, D, M, 3 * D, D, ep_out(qkv, 3 * D, l.qkv.bias), ex.st));
    RUN(ex, attention_f16(qkv, 3 * D, off(qkv, D), 3 * D, off(qkv, 2 * D), 3 * D, B, H, N, N, d, scale, att, D, ex.st));
    RUN(ex, gemm_f16(att, D, l.out.w, D, M, D, D, ep_out(h2, D, l.out.bias, h, D), ex.st));
    RUN(ex, layernorm_rows(h2, M, D, eps, l.ln2.g, l.ln2.b, nrm, ex.st));
    RUN(ex, gemm_f16(nrm, D, l.fc1.w, D, M, F, D, ep_out(mid, F, l.fc1.bias, nullptr, 0, ACT_QUICKGELU), ex.st));
    RUN(ex, gemm_f16(mid, F, l.fc2.w, F, M, D, F, ep_out(h, D, l.fc2.bias, h2, D), ex.st));
  }
  if (!ex.dry) {
    const size_t row = static_cast<size_t>(D) * sizeof(__half);
    GYRE_CHECK_CUDA(cudaMemcpy2DAsync(tail, T * row, h + static_cast<size_t>(L) * D, N * row, T * row, B, cudaMemcpyDeviceToDevice,
                                      ex.st));
  }
  RUN(ex, layernorm_rows(tail, B * T, D, eps, ln_post_.g, ln_post_.b, tail_n, ex.st));
  RUN(ex, gemm_f16(tail_n, D, proj_.w, D, B * T, cfg_.context_dim, D, ep_out(out, cfg_.context_dim), ex.st));
  EX_CHECK(ex);
  return 0;
}

// ------------------------------------------------------------------------------------------ T2I-adapter encoder
// gyre/pipeline/t2i_adapter/adapter.py: PixelUnshuffle(8) -> conv_in -> per level `nums_rb` ResnetBlocks
// ([Downsample] -> [in_conv] -> block1 3x3 -> ReLU -> block2 -> + skip) -> one feature map per level.
AdapterModel::AdapterModel(const gyre_b200_adapter_config& cfg) : cfg_(cfg) {
  const int* ch = cfg.channels;
  if (cfg.light) {
    // Adapter_light (adapter.py:202-263): state-dict names body.{level}.in_conv / .body.{j}.block1 / .block2 / .out_conv
    light_.resize(cfg.num_levels);
    for (int i = 0; i < cfg.num_levels; ++i) {
      AdapterLightW* e = &light_[i];
      const std::string p = "body." + std::to_string(i);
      e->in_c = i == 0 ? cfg.cin : ch[i - 1];
      e->inter_c = ch[i] / 4;
      e->out_c = ch[i];
      reg_linear(p + ".in_conv", e->inter_c, e->in_c, true, &e->in1);
      e->b1.resize(cfg.nums_rb);
      e->b2.resize(cfg.nums_rb);
      for (int j = 0; j < cfg.nums_rb; ++j) {
        reg_conv3(p + ".body." + std::to_string(j) + ".block1", e->inter_c, e->inter_c, &e->b1[j]);
        reg_conv3(p + ".body." + std::to_string(j) + ".block2", e->inter_c, e->inter_c, &e->b2[j]);
      }
      reg_linear(p + ".out_conv", e->out_c, e->inter_c, true, &e->out1);
    }
    return;
  }
  reg_conv3("conv_in", cfg.cin, ch[0], &conv_in_);
  body_.resize(static_cast<size_t>(cfg.num_levels) * cfg.nums_rb);
  for (int i = 0; i < cfg.num_levels; ++i)
    for (int j = 0; j < cfg.nums_rb; ++j) {
      AdapterBlockW* b = &body_[static_cast<size_t>(i) * cfg.nums_rb + j];
      const std::string p = "body." + std::to_string(i * cfg.nums_rb + j);
      b->down = i != 0 && j == 0;
      b->in_c = b->down ? ch[i - 1] : ch[i];
      b->out_c = ch[i];
      b->has_in = b->in_c != b->out_c || !cfg.sk;
      b->has_skep = !cfg.sk;
      if (b->down && cfg.use_conv) reg_conv3(p + ".down_opt.op", b->in_c, b->in_c, &b->down3);
      if (b->has_in) {
        if (cfg.ksize == 3) reg_conv3(p + ".in_conv", b->in_c, b->out_c, &b->in3);
        else reg_linear(p + ".in_conv", b->out_c, b->in_c, true, &b->in1);
      }
      reg_conv3(p + ".block1", b->out_c, b->out_c, &b->b1);
      if (cfg.ksize == 3) reg_conv3(p + ".block2", b->out_c, b->out_c, &b->b2_3);
      else reg_linear(p + ".block2", b->out_c, b->out_c, true, &b->b2_1);
      if (b->has_skep) {
        // upstream declares skep on in_c channels but applies it after in_conv (adapter.py:76-78, 91-97): the widths
        // must agree, as they do in every configuration upstream can run
        if (cfg.ksize == 3) reg_conv3(p + ".skep", b->in_c, b->out_c, &b->sk3);
        else reg_linear(p + ".skep", b->out_c, b->in_c, true, &b->sk1);
      }
    }
}

int AdapterModel::forward_light(Exec& ex, const __half* image, int B, int H, int W, __half* const* features) {
  const int cimg = cfg_.cin / 64;
  int h = H / 8, w = W / 8;
  __half* x = ex.p16(static_cast<size_t>(B) * h * w * cfg_.cin);
  RUN(ex, pixel_unshuffle8_nchw_to_nhwc(image, B, cimg, H, W, x, ex.st));
  for (int i = 0; i < cfg_.num_levels; ++i) {
    const AdapterLightW& e = light_[i];
    if (i > 0) {
      GYRE_REQUIRE(h >= 2 && w >= 2, "adapter: feature map %dx%d too small to pool", h, w);
      __half* d = ex.p16(static_cast<size_t>(B) * (h / 2) * (w / 2) * e.in_c);
      RUN(ex, avg_pool2x2_nhwc(x, B, h, w, e.in_c, d, ex.st));
      x = d;
      h /= 2;
      w /= 2;
    }
    const size_t n = static_cast<size_t>(B) * h * w * e.inter_c;
    __half* t = ex.p16(n);
    RUN(ex, gemm_f16(x, e.in_c, e.in1.w, e.in_c, B * h * w, e.inter_c, e.in_c, ep_out(t, e.inter_c, e.in1.bias), ex.st));
    for (int j = 0; j < cfg_.nums_rb; ++j) {
      __half* h1 = ex.p16(n);
      RUN(ex, conv3x3_f16(t, e.inter_c, B, h, w, e.inter_c, e.b1[j].wp, e.inter_c, 1, 1,
                          ep_out(h1, e.inter_c, e.b1[j].bias, nullptr, 0, ACT_RELU), ex.st));
      __half* o = ex.p16(n);
      RUN(ex, conv3x3_f16(h1, e.inter_c, B, h, w, e.inter_c, e.b2[j].wp, e.inter_c, 1, 1,
                          ep_out(o, e.inter_c, e.b2[j].bias, t, e.inter_c), ex.st));
      t = o;
    }
    __half* f = ex.p16(static_cast<size_t>(B) * h * w * e.out_c);
    RUN(ex, gemm_f16(t, e.inter_c, e.out1.w, e.inter_c, B * h * w, e.out_c, e.inter_c, ep_out(f, e.out_c, e.out1.bias), ex.st));
    RUN(ex, nhwc_to_nchw_f16(f, e.out_c, B, e.out_c, h, w, features[i], ex.st));
    x = f;
  }
  return 0;
}

int AdapterModel::forward(Exec& ex, const __half* image, int B, int H, int W, __half* const* features) {
  GYRE_REQUIRE(B > 0 && H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0, "adapter_forward: image %dx%d must be a multiple of 8", H, W);
  if (!ex.dry) GYRE_TRY(ensure_device());
  if (cfg_.light) return forward_light(ex, image, B, H, W, features);
  const int* ch = cfg_.channels;
  const int cimg = cfg_.cin / 64;
  int h = H / 8, w = W / 8;
  __half* x0 = ex.p16(static_cast<size_t>(B) * h * w * cfg_.cin);
  RUN(ex, pixel_unshuffle8_nchw_to_nhwc(image, B, cimg, H, W, x0, ex.st));
  __half* x = ex.p16(static_cast<size_t>(B) * h * w * ch[0]);
  RUN(ex, conv3x3_f16(x0, cfg_.cin, B, h, w, cfg_.cin, conv_in_.wp, ch[0], 1, 1, ep_out(x, ch[0], conv_in_.bias), ex.st));
  // one op of the ksize-1 / ksize-3 pair: 1x1 conv == GEMM over the pixels
  auto conv_k = [&](const Conv3W& c3, const LinW& c1, const __half* src, int cin, int cout, __half* dst,
                    const __half* residual) -> int {
    if (cfg_.ksize == 3)
      RUN(ex, conv3x3_f16(src, cin, B, h, w, cin, c3.wp, cout, 1, 1, ep_out(dst, cout, c3.bias, residual, residual ? cout : 0),
                          ex.st));
    else
      RUN(ex, gemm_f16(src, cin, c1.w, cin, B * h * w, cout, cin, ep_out(dst, cout, c1.bias, residual, residual ? cout : 0),
                       ex.st));
    return 0;
  };
  for (int i = 0; i < cfg_.num_levels; ++i) {
    for (int j = 0; j < cfg_.nums_rb; ++j) {
      const AdapterBlockW& b = body_[static_cast<size_t>(i) * cfg_.nums_rb + j];
      if (b.down) {
        GYRE_REQUIRE(b.has_skep == false || b.in_c == b.out_c, "adapter: sk = 0 needs equal widths on consecutive levels");
        if (cfg_.use_conv) {
          const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
          __half* d = ex.p16(static_cast<size_t>(B) * ho * wo * b.in_c);
          RUN(ex, conv3x3_f16(x, b.in_c, B, h, w, b.in_c, b.down3.wp, b.in_c, 2, 1, ep_out(d, b.in_c, b.down3.bias), ex.st));
          x = d;
          h = ho;
          w = wo;
        } else {
          GYRE_REQUIRE(h >= 2 && w >= 2, "adapter: feature map %dx%d too small to pool", h, w);
          __half* d = ex.p16(static_cast<size_t>(B) * (h / 2) * (w / 2) * b.in_c);
          RUN(ex, avg_pool2x2_nhwc(x, B, h, w, b.in_c, d, ex.st));
          x = d;
          h /= 2;
          w /= 2;
        }
      }
      const size_t n = static_cast<size_t>(B) * h * w * b.out_c;
      if (b.has_in) {
        __half* t = ex.p16(n);
        GYRE_TRY(conv_k(b.in3, b.in1, x, b.in_c, b.out_c, t, nullptr));
        x = t;
      }
      __half* h1 = ex.p16(n);
      RUN(ex, conv3x3_f16(x, b.out_c, B, h, w, b.out_c, b.b1.wp, b.out_c, 1, 1,
                          ep_out(h1, b.out_c, b.b1.bias, nullptr, 0, ACT_RELU), ex.st));
      const __half* skip = x;
      if (b.has_skep) {
        __half* sk = ex.p16(n);
        GYRE_TRY(conv_k(b.sk3, b.sk1, x, b.out_c, b.out_c, sk, nullptr));
        skip = sk;
      }
      __half* o = ex.p16(n);
      GYRE_TRY(conv_k(b.b2_3, b.b2_1, h1, b.out_c, b.out_c, o, skip));
      x = o;
    }
    RUN(ex, nhwc_to_nchw_f16(x, ch[i], B, ch[i], h, w, features[i], ex.st));
  }
  return 0;
}

}  // namespace gyre
