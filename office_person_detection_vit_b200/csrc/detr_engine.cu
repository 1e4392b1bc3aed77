// DETR-ResNet-50 detector engine: weight packing, workspace layout and the launch plan that chains the kernels of
// tc_gemm.cu (tcgen05 GEMM / implicit-GEMM convolution), attention.cu and detr_kernels.cu into one forward pass.
//
// Arithmetic replaced (third-party `transformers`, the code the reference's removed ViTDetector drove; file:line map in
// oracle/detr_oracle.py): DetrImageProcessor preprocessing, ResNet-50 + DetrFrozenBatchNorm2d, input_projection,
// DetrSinePositionEmbedding, 6 encoder + 6 decoder layers, class / box heads.
//
// Numerics (DESIGN.md "numerics"): frozen BN is folded into the convolution weights in float32 and the product is
// rounded once to bf16; every stored activation is bf16; accumulation, LayerNorm statistics, softmax, the final
// LayerNorm input path and both heads are float32.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <list>
#include <map>
#include <string>
#include <vector>

#include "detr_kernels.h"
#include "opd_common.h"
#include "tc_gemm.h"

namespace opd {
namespace {

constexpr int kD = 256, kHeads = 8, kFFN = 2048, kEnc = 6, kDec = 6, kQueries = 100, kClasses = 92;
constexpr int kStageDepth[4] = {3, 4, 6, 3};
constexpr int kStageWidth[4] = {256, 512, 1024, 2048};
constexpr float kBnEps = 1e-5f;

using bf16 = __nv_bfloat16;

struct ConvW {
  bf16* w = nullptr;      // [cout, k, k, cin]
  float* bias = nullptr;  // folded BN shift
  int cin = 0, cout = 0, k = 1, stride = 1;
};
struct LinW {
  bf16* w = nullptr;   // [n, k]
  float* b = nullptr;  // [n]
  int n = 0, k = 0;
};
struct LnW {
  float* g = nullptr;
  float* b = nullptr;
};
struct BlockW {
  bool has_shortcut = false;
  ConvW shortcut, c0, c1, c2;
  bf16* w3sc = nullptr;      // [width, mid + cin] = [layer.2 | shortcut] weights (first block of stages 1, 3 and 4)
  float* bias3sc = nullptr;  // layer.2 shift + shortcut shift
};
struct EncW {
  LinW qk, v, o, fc1, fc2;
  LnW ln1, ln2;
};
struct DecW {
  LinW sqk, sv, so, cq, co, fc1, fc2;
  LnW ln1, ln2, ln3;
};

struct Tap {
  const void* ptr;
  int64_t rows, cols;
  int is_f32;
};

struct Step {
  std::function<int(cudaStream_t)> run;
  int kind;          // OPD_STEP_* (include/opd_b200.h)
  double flops;      // algorithmic: 2 * M * N * K of the contraction (0 for bandwidth-bound kernels)
  double bytes;      // algorithmic: operands read once + result written once
  std::string name;
};

struct FrameGroup {   // n frames of one size (a batch may mix sizes: opd_detr_forward_mixed)
  int n, H0, W0;
  bool operator==(const FrameGroup& o) const { return n == o.n && H0 == o.H0 && W0 == o.W0; }
};

struct Plan {
  int B = 0, H0 = 0, W0 = 0;
  std::vector<FrameGroup> groups;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  std::vector<Step> steps;
  std::vector<Step> once;   // frame-independent prologue (decoder layer 0's self-attention block): run before the first step only
  bool once_done = false;
  std::map<std::string, Tap> taps;
};

}  // namespace
}  // namespace opd

struct opd_detr {
  int device = 0;
  int debug = 0;
  int fuse_shortcut = 1;   // stage-1 first block: projection shortcut computed inside the fused tail
  int fuse_tail = 1;   // 0: run the 3x3 and the 1x1 expansion of stages 1-2 as separate kernels (A/B comparison)
  int do_resize = 1;   // 0: frames are fed at their own size (DetrImageProcessor(do_resize=False))
  std::vector<void*> allocs;
  opd::ConvW stem;   // packed as a 4x1 convolution over the 64-channel space-to-depth layout (im2col fallback path)
  opd::bf16* stem_taps = nullptr;   // [64, 16 taps x 16 lanes]: 4x4 convolution over the 16-lane space-to-depth tensor
  int stem_halo = 1;
  std::vector<opd::BlockW> blocks;
  opd::LinW input_proj;
  opd::EncW enc[opd::kEnc];
  opd::DecW dec[opd::kDec];
  opd::LinW cross_k_all, cross_v_all;   // the 6 decoder layers' encoder_attn k_proj / v_proj stacked on N
  opd::LnW dec_norm;
  float* qpos = nullptr;   // [100, 256]
  opd::HeadWeights heads{};
  // per-call arguments read by the plan's steps
  const uint8_t* cur_frames = nullptr;
  int cur_bgr = 1;
  std::vector<const uint8_t*> cur_group_frames;   // mixed-size batches: one frame block per group
  std::vector<int> cur_group_bgr;
  float* cur_logits = nullptr;
  float* cur_boxes = nullptr;
  // Launch plans, most recently used first, one per (B, H0, W0, workspace): a stream that alternates between batch shapes (a short
  // last batch, frames of two sizes) re-encodes no tensor maps.  `plan` is the plan of the last forward (taps, profile).
  std::list<opd::Plan> plans;
  opd::Plan* plan = nullptr;
  // outputs of the frame-independent prologue, per batch size (owned by the model, not the caller's workspace; never freed before
  // opd_detr_destroy: a CUDA graph captured from an evicted plan may still read them)
  std::map<int, void*> dec0_bufs;
  void drop_plans() {
    plans.clear();
    plan = nullptr;
  }
};
constexpr size_t kMaxPlans = 8;

namespace opd {
namespace {

// ------------------------------------------------------------------------------------------------------------
// weight loading
// ------------------------------------------------------------------------------------------------------------
struct Loader {
  std::map<std::string, std::pair<const float*, int64_t>> t;
  opd_detr* m;
  int rc = OPD_OK;

  const float* get(const std::string& name, int64_t numel) {
    auto it = t.find(name);
    if (it == t.end()) {
      if (rc == OPD_OK) rc = fail(OPD_ERR_INVALID, "detr weights: tensor '%s' is missing", name.c_str());
      return nullptr;
    }
    if (it->second.second != numel) {
      if (rc == OPD_OK)
        rc = fail(OPD_ERR_INVALID, "detr weights: tensor '%s' has %lld elements, expected %lld", name.c_str(),
                  (long long)it->second.second, (long long)numel);
      return nullptr;
    }
    return it->second.first;
  }
  template <typename T>
  T* upload(const std::vector<T>& host) {
    if (rc != OPD_OK) return nullptr;
    void* d = nullptr;
    if (cudaMalloc(&d, host.size() * sizeof(T)) != cudaSuccess) {
      rc = fail(OPD_ERR_NOMEM, "detr weights: cudaMalloc of %zu bytes failed", host.size() * sizeof(T));
      return nullptr;
    }
    m->allocs.push_back(d);
    if (cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
      rc = fail(OPD_ERR_CUDA, "detr weights: upload failed");
      return nullptr;
    }
    return static_cast<T*>(d);
  }
  float* upload_f32(const float* src, int64_t n) {
    if (!src) return nullptr;
    return upload(std::vector<float>(src, src + n));
  }

  // fold frozen BN: scale = gamma * rsqrt(var + eps); w' = w * scale; shift = beta - mean * scale   (float32)
  bool folded(const std::string& prefix, int cin, int cout, int k, std::vector<float>* wf, std::vector<float>* shift) {
    const float* w = get(prefix + ".convolution.weight", (int64_t)cout * cin * k * k);
    const float* g = get(prefix + ".normalization.weight", cout);
    const float* b = get(prefix + ".normalization.bias", cout);
    const float* mu = get(prefix + ".normalization.running_mean", cout);
    const float* var = get(prefix + ".normalization.running_var", cout);
    if (rc != OPD_OK) return false;
    wf->resize((size_t)cout * cin * k * k);
    shift->resize(cout);
    for (int n = 0; n < cout; ++n) {
      const float scale = g[n] * (1.0f / sqrtf(var[n] + kBnEps));
      (*shift)[n] = b[n] - mu[n] * scale;
      const float* src = w + (size_t)n * cin * k * k;
      float* dst = wf->data() + (size_t)n * cin * k * k;
      for (int i = 0; i < cin * k * k; ++i) dst[i] = src[i] * scale;
    }
    return true;
  }

  ConvW conv(const std::string& prefix, int cin, int cout, int k, int stride) {
    ConvW c;
    c.cin = cin; c.cout = cout; c.k = k; c.stride = stride;
    std::vector<float> wf, shift;
    if (!folded(prefix, cin, cout, k, &wf, &shift)) return c;
    std::vector<bf16> packed((size_t)cout * k * k * cin);   // OIHW -> O,kh,kw,I
    for (int n = 0; n < cout; ++n)
      for (int ci = 0; ci < cin; ++ci)
        for (int r = 0; r < k; ++r)
          for (int s = 0; s < k; ++s)
            packed[(((size_t)n * k + r) * k + s) * cin + ci] =
                __float2bfloat16(wf[(((size_t)n * cin + ci) * k + r) * k + s]);
    c.w = upload(packed);
    c.bias = upload(shift);
    last_packed = packed;
    last_shift = shift;
    return c;
  }
  std::vector<bf16> last_packed;   // host copies of the most recent conv() (for fused weight layouts)
  std::vector<float> last_shift;

  // 7x7 / stride 2 / pad 3 stem over RGB  ->  4x1 convolution over the 64-channel layout written by K1:
  //   channel = kw4 * 16 + (dy * 2 + dx) * 3 + c,  tap row kh4;  original tap (r, s): kh4 = (r+1)/2, dy = (r+1)&1 (same for s)
  ConvW stem_conv(const std::string& prefix) {
    ConvW c;
    c.cin = 64; c.cout = 64; c.k = 4; c.stride = 1;
    std::vector<float> wf, shift;
    if (!folded(prefix, 3, 64, 7, &wf, &shift)) return c;
    std::vector<bf16> packed((size_t)64 * 4 * 64, __float2bfloat16(0.f));
    for (int n = 0; n < 64; ++n)
      for (int ci = 0; ci < 3; ++ci)
        for (int r = 0; r < 7; ++r)
          for (int s = 0; s < 7; ++s) {
            const int kh4 = (r + 1) / 2, dy = (r + 1) & 1, kw4 = (s + 1) / 2, dx = (s + 1) & 1;
            const int ch = kw4 * 16 + (dy * 2 + dx) * 3 + ci;
            packed[((size_t)n * 4 + kh4) * 64 + ch] = __float2bfloat16(wf[(((size_t)n * 3 + ci) * 7 + r) * 7 + s]);
          }
    c.w = upload(packed);
    c.bias = upload(shift);
    // tap layout of stem_conv.cu: column (kh4 * 4 + kw4) * 16 + (dy * 2 + dx) * 3 + ci
    std::vector<bf16> taps((size_t)64 * 256, __float2bfloat16(0.f));
    for (int n = 0; n < 64; ++n)
      for (int ci = 0; ci < 3; ++ci)
        for (int r = 0; r < 7; ++r)
          for (int s = 0; s < 7; ++s) {
            const int kh4 = (r + 1) / 2, dy = (r + 1) & 1, kw4 = (s + 1) / 2, dx = (s + 1) & 1;
            taps[(size_t)n * 256 + (kh4 * 4 + kw4) * 16 + (dy * 2 + dx) * 3 + ci] =
                __float2bfloat16(wf[(((size_t)n * 3 + ci) * 7 + r) * 7 + s]);
          }
    m->stem_taps = upload(taps);
    return c;
  }

  // rows of several [n_i, k] Linear layers stacked on N
  LinW linear(const std::vector<std::string>& prefixes, int n_each, int k) {
    LinW l;
    l.n = n_each * (int)prefixes.size();
    l.k = k;
    std::vector<bf16> w((size_t)l.n * k);
    std::vector<float> b(l.n);
    for (size_t i = 0; i < prefixes.size(); ++i) {
      const float* ws = get(prefixes[i] + ".weight", (int64_t)n_each * k);
      const float* bs = get(prefixes[i] + ".bias", n_each);
      if (rc != OPD_OK) return l;
      for (size_t j = 0; j < (size_t)n_each * k; ++j) w[i * (size_t)n_each * k + j] = __float2bfloat16(ws[j]);
      for (int j = 0; j < n_each; ++j) b[i * n_each + j] = bs[j];
    }
    l.w = upload(w);
    l.b = upload(b);
    return l;
  }
  LnW layer_norm(const std::string& prefix) {
    LnW l;
    l.g = upload_f32(get(prefix + ".weight", kD), kD);
    l.b = upload_f32(get(prefix + ".bias", kD), kD);
    return l;
  }
  float* transposed(const std::string& name, int n, int k) {   // [n,k] -> [k,n] float32
    const float* src = get(name, (int64_t)n * k);
    if (!src) return nullptr;
    std::vector<float> t((size_t)n * k);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < k; ++j) t[(size_t)j * n + i] = src[(size_t)i * k + j];
    return upload(t);
  }
};

int load_weights(opd_detr* m, const opd_tensor_f32* tensors, int n_tensors) {
  Loader L;
  L.m = m;
  for (int i = 0; i < n_tensors; ++i) {
    OPD_REQUIRE(tensors[i].name && tensors[i].data, "detr weights: tensor %d has a NULL name or data pointer", i);
    L.t[tensors[i].name] = {tensors[i].data, tensors[i].numel};
  }
  const std::string bb = "model.backbone.model.";
  m->stem = L.stem_conv(bb + "embedder.embedder");
  int cin = 64;
  for (int s = 0; s < 4; ++s) {
    const int width = kStageWidth[s], mid = width / 4;
    for (int l = 0; l < kStageDepth[s]; ++l) {
      const int stride = (l == 0 && s > 0) ? 2 : 1;
      const std::string p = bb + "encoder.stages." + std::to_string(s) + ".layers." + std::to_string(l);
      BlockW b;
      b.has_shortcut = l == 0;
      std::vector<bf16> sc_w;
      std::vector<float> sc_b;
      if (l == 0) {
        b.shortcut = L.conv(p + ".shortcut", cin, width, 1, stride);
        sc_w = L.last_packed;
        sc_b = L.last_shift;
      }
      b.c0 = L.conv(p + ".layer.0", cin, mid, 1, 1);
      b.c1 = L.conv(p + ".layer.1", mid, mid, 3, stride);   // v1.5: the stride sits on the 3x3
      b.c2 = L.conv(p + ".layer.2", mid, width, 1, 1);
      if (l == 0 && cin == 64 && mid == 64 && stride == 1 && L.rc == OPD_OK) {
        // fused projection shortcut (tc_bottleneck_halo.cu): one [width, 128] operand, K = [layer.2 input | block input]
        std::vector<bf16> cat((size_t)width * 128);
        std::vector<float> bias(width);
        for (int n = 0; n < width; ++n) {
          for (int k = 0; k < 64; ++k) {
            cat[(size_t)n * 128 + k] = L.last_packed[(size_t)n * 64 + k];
            cat[(size_t)n * 128 + 64 + k] = sc_w[(size_t)n * 64 + k];
          }
          bias[n] = L.last_shift[n] + sc_b[n];
        }
        b.w3sc = L.upload(cat);
        b.bias3sc = L.upload(bias);
      }
      if (l == 0 && mid >= 256 && L.rc == OPD_OK) {
        // stages 3-4: [layer.2 | shortcut] along K for the dual-source GEMM (tc_gemm.cu gemm_plan_linear_plus_shortcut)
        const int kk = mid + cin;
        std::vector<bf16> cat((size_t)width * kk);
        std::vector<float> bias(width);
        for (int n = 0; n < width; ++n) {
          for (int k = 0; k < mid; ++k) cat[(size_t)n * kk + k] = L.last_packed[(size_t)n * mid + k];
          for (int k = 0; k < cin; ++k) cat[(size_t)n * kk + mid + k] = sc_w[(size_t)n * cin + k];
          bias[n] = L.last_shift[n] + sc_b[n];
        }
        b.w3sc = L.upload(cat);
        b.bias3sc = L.upload(bias);
      }
      m->blocks.push_back(b);
      cin = width;
    }
  }
  m->input_proj = L.linear({"model.input_projection"}, kD, 2048);
  for (int i = 0; i < kEnc; ++i) {
    const std::string p = "model.encoder.layers." + std::to_string(i);
    EncW& e = m->enc[i];
    e.qk = L.linear({p + ".self_attn.q_proj", p + ".self_attn.k_proj"}, kD, kD);
    e.v = L.linear({p + ".self_attn.v_proj"}, kD, kD);
    e.o = L.linear({p + ".self_attn.o_proj"}, kD, kD);
    e.ln1 = L.layer_norm(p + ".self_attn_layer_norm");
    e.fc1 = L.linear({p + ".mlp.fc1"}, kFFN, kD);
    e.fc2 = L.linear({p + ".mlp.fc2"}, kD, kFFN);
    e.ln2 = L.layer_norm(p + ".final_layer_norm");
  }
  std::vector<std::string> ck, cv;
  for (int i = 0; i < kDec; ++i) {
    const std::string p = "model.decoder.layers." + std::to_string(i);
    DecW& d = m->dec[i];
    d.sqk = L.linear({p + ".self_attn.q_proj", p + ".self_attn.k_proj"}, kD, kD);
    d.sv = L.linear({p + ".self_attn.v_proj"}, kD, kD);
    d.so = L.linear({p + ".self_attn.o_proj"}, kD, kD);
    d.ln1 = L.layer_norm(p + ".self_attn_layer_norm");
    d.cq = L.linear({p + ".encoder_attn.q_proj"}, kD, kD);
    d.co = L.linear({p + ".encoder_attn.o_proj"}, kD, kD);
    d.ln2 = L.layer_norm(p + ".encoder_attn_layer_norm");
    d.fc1 = L.linear({p + ".mlp.fc1"}, kFFN, kD);
    d.fc2 = L.linear({p + ".mlp.fc2"}, kD, kFFN);
    d.ln3 = L.layer_norm(p + ".final_layer_norm");
    ck.push_back(p + ".encoder_attn.k_proj");
    cv.push_back(p + ".encoder_attn.v_proj");
  }
  m->cross_k_all = L.linear(ck, kD, kD);
  m->cross_v_all = L.linear(cv, kD, kD);
  m->dec_norm = L.layer_norm("model.decoder.layernorm");
  m->qpos = L.upload_f32(L.get("model.query_position_embeddings.weight", (int64_t)kQueries * kD), (int64_t)kQueries * kD);
  m->heads.wc_t = L.transposed("class_labels_classifier.weight", kClasses, kD);
  m->heads.bc = L.upload_f32(L.get("class_labels_classifier.bias", kClasses), kClasses);
  m->heads.w0_t = L.transposed("bbox_predictor.layers.0.weight", kD, kD);
  m->heads.b0 = L.upload_f32(L.get("bbox_predictor.layers.0.bias", kD), kD);
  m->heads.w1_t = L.transposed("bbox_predictor.layers.1.weight", kD, kD);
  m->heads.b1 = L.upload_f32(L.get("bbox_predictor.layers.1.bias", kD), kD);
  m->heads.w2 = L.upload_f32(L.get("bbox_predictor.layers.2.weight", 4 * kD), 4 * kD);
  m->heads.b2 = L.upload_f32(L.get("bbox_predictor.layers.2.bias", 4), 4);
  return L.rc;
}

// ------------------------------------------------------------------------------------------------------------
// shapes
// ------------------------------------------------------------------------------------------------------------
// transformers/image_transforms.py:206-242 get_size_with_aspect_ratio(size=800, max_size=1333)
void resized_size(int h, int w, int* oh, int* ow) {
  double size = 800.0;
  const double max_size = 1333.0;
  const double mn = (double)(h < w ? h : w), mx = (double)(h < w ? w : h);
  bool has_raw = false;
  double raw = 0.0;
  if (mx / mn * size > max_size) {
    raw = max_size * mn / mx;
    has_raw = true;
    size = (double)(long long)nearbyint(raw);   // python round(): half to even
  }
  const int isz = (int)size;
  if ((h <= w && h == isz) || (w <= h && w == isz)) {
    *oh = h; *ow = w;
    return;
  }
  if (w < h) {
    *ow = isz;
    *oh = (int)(has_raw ? raw * h / w : size * h / w);
  } else {
    *oh = isz;
    *ow = (int)(has_raw ? raw * w / h : size * w / h);
  }
}

// Separable uint8 antialias-bilinear resize tables, restating ATen's CPU kernel (aten/src/ATen/native/cpu/
// UpSampleKernel.cpp: _compute_indices_min_size_weights_aa + _compute_index_ranges_int16_weights; what torchvision's
// resize on a uint8 tensor, i.e. transformers' DetrImageProcessor, runs): double-precision triangle-filter weights,
// normalised, scaled to int16 with the largest precision that keeps the maximum weight below 2^15.
struct ResizeTable {
  std::vector<int32_t> x0;
  std::vector<int16_t> w;   // [out, ksize]
  int ksize = 0, precision = 0;
};
ResizeTable resize_table(int in_size, int out_size) {
  ResizeTable t;
  const double scale = (double)in_size / (double)out_size;
  const double support = scale >= 1.0 ? scale : 1.0;   // interp_size (2) * 0.5 * scale
  t.ksize = (int)std::ceil(support) * 2 + 1;
  const double invscale = scale >= 1.0 ? 1.0 / scale : 1.0;
  std::vector<double> wt((size_t)out_size * t.ksize, 0.0);
  t.x0.resize(out_size);
  double wt_max = 0.0;
  for (int i = 0; i < out_size; ++i) {
    const double center = scale * (i + 0.5);
    long long xmin = (long long)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    long long xsize = (long long)(center + support + 0.5);
    if (xsize > in_size) xsize = in_size;
    xsize -= xmin;
    if (xsize < 0) xsize = 0;
    if (xsize > t.ksize) xsize = t.ksize;
    double total = 0.0;
    double* wp = wt.data() + (size_t)i * t.ksize;
    for (int j = 0; j < xsize; ++j) {
      double x = ((double)(j + xmin) - center + 0.5) * invscale;
      if (x < 0.0) x = -x;
      const double w = x < 1.0 ? 1.0 - x : 0.0;
      wp[j] = w;
      total += w;
    }
    if (total != 0.0)
      for (int j = 0; j < xsize; ++j) {
        wp[j] /= total;
        if (wp[j] > wt_max) wt_max = wp[j];
      }
    t.x0[i] = (int32_t)xmin;
  }
  for (t.precision = 0; t.precision < 22; ++t.precision) {
    const int next_value = (int)(0.5 + wt_max * (double)(1 << (t.precision + 1)));
    if (next_value >= (1 << 15)) break;
  }
  t.w.resize(wt.size());
  for (size_t i = 0; i < wt.size(); ++i) {
    const double v = wt[i] * (double)(1 << t.precision);
    t.w[i] = (int16_t)(v < 0 ? (int)(-0.5 + v) : (int)(0.5 + v));
  }
  return t;
}

inline int conv_out(int x, int k, int stride, int pad) { return (x + 2 * pad - k) / stride + 1; }

struct Shapes {
  int Hin, Win;        // model input
  int Hs, Ws;          // stem output (= space-to-depth grid)
  int Hp, Wp;          // after max pooling
  int h[4], w[4];      // stage outputs
};
Shapes shapes_for(int H0, int W0, bool do_resize = true) {
  Shapes s;
  s.Hin = H0;
  s.Win = W0;
  if (do_resize) resized_size(H0, W0, &s.Hin, &s.Win);
  s.Hs = conv_out(s.Hin, 7, 2, 3);
  s.Ws = conv_out(s.Win, 7, 2, 3);
  s.Hp = conv_out(s.Hs, 3, 2, 1);
  s.Wp = conv_out(s.Ws, 3, 2, 1);
  int hh = s.Hp, ww = s.Wp;
  for (int i = 0; i < 4; ++i) {
    if (i > 0) {
      hh = conv_out(hh, 3, 2, 1);
      ww = conv_out(ww, 3, 2, 1);
    }
    s.h[i] = hh;
    s.w[i] = ww;
  }
  return s;
}

// ------------------------------------------------------------------------------------------------------------
// workspace arena: named slots are reused (backbone ping-pong) unless the engine is in debug mode
// ------------------------------------------------------------------------------------------------------------
struct Arena {
  uint8_t* base;
  size_t off = 0;
  void* take(size_t bytes) {
    off = (off + 1023) & ~(size_t)1023;
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

// torch.nn.functional.interpolate(mask, size=(out,)) in its default "nearest" mode: source index of output index i
// (modeling_detr.py:281-283 downsamples pixel_mask to the feature map this way)
inline int nearest_src(int i, int in_size, int out_size) {
  const float scale = (float)in_size / (float)out_size;
  const int src = (int)floorf((float)i * scale);
  return src < in_size - 1 ? src : in_size - 1;
}

// Builds the launch plan.  With ws == nullptr only the workspace size is computed (no tensor maps are encoded).
// `groups`: the batch, group after group; more than one frame size -> every frame is resized on its own (800 / 1333 rule), placed
// in the top-left corner of a canvas of the largest model input size and the rest of the path runs with DetrImageProcessor's
// pixel_mask semantics (zero padding after normalisation, mask-aware sine embedding, key-padding mask in the attentions).
int build_plan(opd_detr* m, const std::vector<FrameGroup>& groups, void* ws, Plan* plan, size_t* bytes_out) {
  const bool dry = ws == nullptr;
  int B = 0;
  for (const FrameGroup& g : groups) B += g.n;
  // model input size of every group and of the batch canvas
  std::vector<std::pair<int, int>> gin(groups.size());
  int Hc = 0, Wc = 0;
  for (size_t i = 0; i < groups.size(); ++i) {
    gin[i] = {groups[i].H0, groups[i].W0};
    if (m->do_resize) resized_size(groups[i].H0, groups[i].W0, &gin[i].first, &gin[i].second);
    Hc = std::max(Hc, gin[i].first);
    Wc = std::max(Wc, gin[i].second);
  }
  const bool mixed = groups.size() > 1;
  const int H0 = mixed ? Hc : groups[0].H0, W0 = mixed ? Wc : groups[0].W0;
  const Shapes sh = shapes_for(H0, W0, !mixed && m->do_resize != 0);
  OPD_REQUIRE(sh.Hs >= 4 && sh.h[3] >= 1 && sh.w[3] >= 1, "detr: frames of %dx%d are too small", H0, W0);
  OPD_REQUIRE(sh.Hs == (sh.Hin + 1) / 2 && sh.Ws == (sh.Win + 1) / 2, "detr: unexpected stem geometry");
  Arena A{static_cast<uint8_t*>(ws)};
  auto& steps = plan->steps;
  auto& taps = plan->taps;
  const bool dbg = m->debug != 0;

  auto act_bytes = [&](long long rows, int ch) { return (size_t)rows * ch * sizeof(bf16); };
  std::vector<Step>* sink = &steps;   // &plan->once while the frame-independent prologue is being planned
  // Direction in which the previous launch walked its rows (false: ascending).  A GEMM layer walks the other way, so that it starts
  // with the rows its producer wrote last, which are still in L2 (opd_set_option("gemm_reverse", 0): always ascending).
  bool prev_reversed = false;
  auto add = [&](int kind, const std::string& name, double flops, double bytes, std::function<int(cudaStream_t)> fn) {
    sink->push_back(Step{std::move(fn), kind, flops, bytes, name});
    prev_reversed = false;
  };
  const long long Ms = (long long)B * sh.Hs * sh.Ws, Mp = (long long)B * sh.Hp * sh.Wp;

  // ---- backbone buffers: 3 big ping-pong slots + 2 mid slots (each sized for its largest user) ----
  size_t big_bytes = act_bytes(Ms, 64), mid_bytes = 0;
  {
    int hh = sh.Hp, ww = sh.Wp;
    for (int s = 0; s < 4; ++s) {
      const int width = kStageWidth[s], mid = width / 4;
      const long long m_in = (long long)B * hh * ww;
      const long long m_out = (long long)B * sh.h[s] * sh.w[s];
      big_bytes = std::max(big_bytes, act_bytes(m_out, width));
      mid_bytes = std::max(mid_bytes, act_bytes(m_in, mid));   // layer.0 of the first block runs at the input size
      hh = sh.h[s];
      ww = sh.w[s];
    }
  }
  void* big[3] = {nullptr, nullptr, nullptr};
  void* mids[2] = {nullptr, nullptr};
  if (!dbg) {
    for (auto& b : big) b = A.take(big_bytes);
    for (auto& b : mids) b = A.take(mid_bytes);
  }
  auto big_slot = [&](int i, size_t bytes) { return dbg ? A.take(bytes) : big[i]; };
  auto mid_slot = [&](int i, size_t bytes) { return dbg ? A.take(bytes) : mids[i]; };

  // ---- uint8 antialias resize to the model input size (camera frames: 720x1280 -> 750x1333) ----
  const bool needs_resize = sh.Hin != H0 || sh.Win != W0;
  uint8_t* resized = nullptr;
  if (needs_resize) {
    const ResizeTable tx = resize_table(W0, sh.Win), ty = resize_table(H0, sh.Hin);
    uint8_t* tmp = static_cast<uint8_t*>(A.take((size_t)B * H0 * sh.Win * 3));
    resized = static_cast<uint8_t*>(A.take((size_t)B * sh.Hin * sh.Win * 3));
    int16_t* wx = static_cast<int16_t*>(A.take(tx.w.size() * sizeof(int16_t)));
    int32_t* x0 = static_cast<int32_t*>(A.take(tx.x0.size() * sizeof(int32_t)));
    int16_t* wy = static_cast<int16_t*>(A.take(ty.w.size() * sizeof(int16_t)));
    int32_t* y0 = static_cast<int32_t*>(A.take(ty.x0.size() * sizeof(int32_t)));
    if (!dry) {
      OPD_CUDA_OK(cudaMemcpy(wx, tx.w.data(), tx.w.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
      OPD_CUDA_OK(cudaMemcpy(x0, tx.x0.data(), tx.x0.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      OPD_CUDA_OK(cudaMemcpy(wy, ty.w.data(), ty.w.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
      OPD_CUDA_OK(cudaMemcpy(y0, ty.x0.data(), ty.x0.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      const int kx = tx.ksize, px = tx.precision, ky = ty.ksize, py = ty.precision;
      add(OPD_STEP_ELEMENTWISE, "resize", 0.0, 3.0 * B * ((double)H0 * W0 + 2.0 * H0 * sh.Win + (double)sh.Hin * sh.Win),
          [=](cudaStream_t s) {
            return launch_resize_u8(m->cur_frames, B, H0, W0, m->cur_bgr, tmp, resized, sh.Hin, sh.Win, wx, x0, kx, px, wy,
                                    y0, ky, py, s);
          });
      taps["resized_u8"] = {resized, (long long)B * sh.Hin * sh.Win, 3, 2};
    }
  }

  // ---- mixed-size batch: every group is resized (or copied) into the top-left corner of its frames of one uint8 canvas ----
  int32_t* valid_hw = nullptr;   // [B, 2] picture size of every frame inside the canvas (device)
  if (mixed) {
    OPD_REQUIRE(m->stem_halo != 0, "detr: mixed-size batches need the space-to-depth stem path");
    const long long frame_stride = (long long)Hc * Wc * 3, row_pitch = (long long)Wc * 3;
    resized = static_cast<uint8_t*>(A.take((size_t)B * frame_stride));
    valid_hw = static_cast<int32_t*>(A.take((size_t)B * 2 * sizeof(int32_t)));
    std::vector<int32_t> hv;
    int b0 = 0;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
      const FrameGroup g = groups[gi];
      const int H1 = gin[gi].first, W1 = gin[gi].second;
      for (int i = 0; i < g.n; ++i) {
        hv.push_back(H1);
        hv.push_back(W1);
      }
      uint8_t* dst = resized + (size_t)b0 * frame_stride;
      if (H1 != g.H0 || W1 != g.W0) {
        const ResizeTable tx = resize_table(g.W0, W1), ty = resize_table(g.H0, H1);
        uint8_t* tmp = static_cast<uint8_t*>(A.take((size_t)g.n * g.H0 * W1 * 3));
        int16_t* wx = static_cast<int16_t*>(A.take(tx.w.size() * sizeof(int16_t)));
        int32_t* x0 = static_cast<int32_t*>(A.take(tx.x0.size() * sizeof(int32_t)));
        int16_t* wy = static_cast<int16_t*>(A.take(ty.w.size() * sizeof(int16_t)));
        int32_t* y0 = static_cast<int32_t*>(A.take(ty.x0.size() * sizeof(int32_t)));
        if (!dry) {
          OPD_CUDA_OK(cudaMemcpy(wx, tx.w.data(), tx.w.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
          OPD_CUDA_OK(cudaMemcpy(x0, tx.x0.data(), tx.x0.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
          OPD_CUDA_OK(cudaMemcpy(wy, ty.w.data(), ty.w.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
          OPD_CUDA_OK(cudaMemcpy(y0, ty.x0.data(), ty.x0.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
          const int kx = tx.ksize, px = tx.precision, ky = ty.ksize, py = ty.precision;
          add(OPD_STEP_ELEMENTWISE, "resize", 0.0, 3.0 * g.n * ((double)g.H0 * g.W0 + 2.0 * g.H0 * W1 + (double)H1 * W1),
              [=](cudaStream_t st) {
                return launch_resize_u8(m->cur_group_frames[gi], g.n, g.H0, g.W0, m->cur_group_bgr[gi], tmp, dst, H1, W1, wx, x0, kx, px,
                                        wy, y0, ky, py, st, frame_stride, row_pitch);
              });
        }
      } else if (!dry) {
        add(OPD_STEP_ELEMENTWISE, "copy_to_canvas", 0.0, 6.0 * g.n * g.H0 * g.W0, [=](cudaStream_t st) {
          return launch_copy_into_canvas(m->cur_group_frames[gi], g.n, g.H0, g.W0, m->cur_group_bgr[gi], dst, frame_stride, row_pitch, st);
        });
      }
      b0 += g.n;
    }
    if (!dry) {
      OPD_CUDA_OK(cudaMemcpy(valid_hw, hv.data(), hv.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
      taps["resized_u8"] = {resized, (long long)B * Hc * Wc, 3, 2};
    }
  }

  // ---- K1 preprocess: -> S [B, Hs, Ws, 16] (stem_conv.cu) or, on the im2col fallback path, X2 [B, Hs, Ws, 64] ----
  const bool stem_halo = m->stem_halo != 0;
  bf16* x2 = static_cast<bf16*>(big_slot(0, act_bytes(Ms, stem_halo ? 16 : 64)));
  if (!dry) {
    const uint8_t* fixed_src = resized;
    if (stem_halo) {
      add(OPD_STEP_ELEMENTWISE, "preprocess", 0.0, (double)B * sh.Hin * sh.Win * 3 + (double)act_bytes(Ms, 16),
          [m, B, sh, x2, fixed_src, valid_hw](cudaStream_t s) {
            return fixed_src ? launch_preprocess_s2d(fixed_src, B, sh.Hin, sh.Win, 0, x2, s, valid_hw)
                             : launch_preprocess_s2d(m->cur_frames, B, sh.Hin, sh.Win, m->cur_bgr, x2, s);
          });
    } else {
      add(OPD_STEP_ELEMENTWISE, "preprocess", 0.0, (double)B * sh.Hin * sh.Win * 3 + (double)act_bytes(Ms, 64),
          [m, B, sh, x2, fixed_src](cudaStream_t s) {
            return fixed_src ? launch_preprocess(fixed_src, B, sh.Hin, sh.Win, 0, x2, s)
                             : launch_preprocess(m->cur_frames, B, sh.Hin, sh.Win, m->cur_bgr, x2, s);
          });
    }
    taps["x2"] = {x2, Ms, stem_halo ? 16 : 64, 0};
  }

  std::string cur_name = "gemm";
  auto add_gemm = [&](const GemmPlan& gp_in) {
    GemmPlan gp = gp_in;
    gp.reverse = g_option_gemm_reverse.load() && !prev_reversed;
    const bool rev = gp.reverse != 0;
    // algorithmic traffic: A once (im2col: the input tensor once), W once, D (and D2 / residual) once
    const double a_bytes = gp.im2col ? 2.0 * gp.g.B * gp.g.H * gp.g.W * gp.g.C / (gp.g.stride * gp.g.stride > 1 && gp.g.KH == 1 ? gp.g.stride * gp.g.stride : 1)
                                     : 2.0 * gp.M * gp.K;
    const double bytes = a_bytes + 2.0 * gp.N * gp.K + 2.0 * gp.M * gp.N * (1 + (gp.has_d2 ? 1 : 0) + (gp.residual ? 1 : 0));
    add(gp.im2col ? OPD_STEP_CONV : OPD_STEP_GEMM, cur_name, 2.0 * gp.M * gp.N * gp.K, bytes,
        [gp](cudaStream_t s) { return gemm_launch(gp, s); });
    prev_reversed = rev;
  };
  auto conv = [&](const bf16* x, int H, int W, const ConvW& c, bf16* y, int epi, const bf16* res) -> int {
    if (dry) return OPD_OK;
    GemmPlan gp;
    if (c.k == 1 && c.stride == 1) {
      const int M = B * H * W;
      if (int rc = gemm_plan_linear(&gp, x, c.cin, c.w, y, c.cout, M, c.cout, c.cin, epi, c.bias, res, c.cout, nullptr,
                                    nullptr, nullptr, nullptr, 0))
        return rc;
    } else {
      const int pad = c.k / 2;
      ConvGeom g{B, H, W, c.cin, c.k, c.k, c.stride, pad, pad, conv_out(H, c.k, c.stride, pad), conv_out(W, c.k, c.stride, pad)};
      if (int rc = gemm_plan_conv(&gp, x, g, c.w, y, c.cout, epi, c.bias, res)) return rc;
    }
    add_gemm(gp);
    return OPD_OK;
  };

  // ---- K2 stem: 4x1 convolution over X2, pad 2 rows above / 1 below ----
  bf16* stem_out = static_cast<bf16*>(big_slot(1, act_bytes(Ms, 64)));
  // max pooling fused into the stem's epilogue (stem_conv.cu): the stem output never reaches memory.  Debug plans keep the
  // two kernels (the "stem" tap) unless opd_set_option("stem_pool", 2) forces the fused kernel.
  const int sp_opt = g_option_stem_pool.load();
  const bool fuse_pool = stem_halo && (sp_opt == 2 || (sp_opt == 1 && !dbg));
  if (!dry) {
    if (fuse_pool) {
      // issued below, once the pooled buffer is known
    } else if (stem_halo) {
      StemPlan sp;
      if (int rc = stem_plan(&sp, x2, B, sh.Hs, sh.Ws, m->stem_taps, m->stem.bias, stem_out)) return rc;
      add(OPD_STEP_CONV, "stem", 2.0 * Ms * 64 * 147, (double)act_bytes(Ms, 16) + (double)act_bytes(Ms, 64),
          [sp](cudaStream_t s) { return stem_launch(sp, s); });
    } else {
      GemmPlan gp;
      cur_name = "stem";
      ConvGeom g{B, sh.Hs, sh.Ws, 64, 4, 1, 1, 2, 0, sh.Hs, sh.Ws};
      if (int rc = gemm_plan_conv(&gp, x2, g, m->stem.w, stem_out, 64, EPI_BIAS_RELU, m->stem.bias, nullptr)) return rc;
      add_gemm(gp);
    }
    if (!fuse_pool) taps["stem"] = {stem_out, Ms, 64, 0};
  }
  // ---- K3 max pooling ----
  bf16* x = static_cast<bf16*>(big_slot(2, act_bytes(Mp, 64)));
  if (!dry) {
    if (fuse_pool) {
      StemPlan sp;
      if (int rc = stem_plan(&sp, x2, B, sh.Hs, sh.Ws, m->stem_taps, m->stem.bias, stem_out, x)) return rc;
      OPD_REQUIRE(sp.P == sh.Hp && sp.Q == sh.Wp, "stem + pool: pooled size %dx%d, expected %dx%d", sp.P, sp.Q, sh.Hp, sh.Wp);
      add(OPD_STEP_CONV, "stem+maxpool", 2.0 * Ms * 64 * 147, (double)act_bytes(Ms, 16) + (double)act_bytes(Mp, 64),
          [sp](cudaStream_t s) { return stem_launch(sp, s); });
    } else {
      add(OPD_STEP_ELEMENTWISE, "maxpool", 0.0, (double)act_bytes(Ms, 64) + (double)act_bytes(Mp, 64),
          [B, sh, stem_out, x](cudaStream_t s) { return launch_maxpool(stem_out, B, sh.Hs, sh.Ws, 64, x, sh.Hp, sh.Wp, s); });
    }
    taps["pool"] = {x, Mp, 64, 0};
  }
  int cur = 2, hh = sh.Hp, ww = sh.Wp, bi = 0;
  for (int s = 0; s < 4; ++s) {
    const int width = kStageWidth[s], mid = width / 4;
    for (int l = 0; l < kStageDepth[s]; ++l, ++bi) {
      const BlockW& bw = m->blocks[bi];
      const int stride = bw.c1.stride;
      const int ho = conv_out(hh, 3, stride, 1), wo = conv_out(ww, 3, stride, 1);
      const long long m_in = (long long)B * hh * ww, m_out = (long long)B * ho * wo;
      const bf16* res = x;
      const std::string bname = "stage" + std::to_string(s) + "." + std::to_string(l);
      const bool fuse_sc = bw.has_shortcut && bw.w3sc && mid == 64 && m->fuse_tail && m->fuse_shortcut && g_option_bneck_halo.load();
      // stages 3-4: the strided projection shortcut is a second A source of the 1x1 expansion GEMM (no shortcut tensor)
      const bool dual_sc = bw.has_shortcut && bw.w3sc && mid >= 256 && m->fuse_shortcut;
      if (bw.has_shortcut && !fuse_sc && !dual_sc) {
        cur_name = bname + ".shortcut";
        bf16* sc = static_cast<bf16*>(big_slot((cur + 1) % 3, act_bytes(m_out, width)));
        if (int rc = conv(x, hh, ww, bw.shortcut, sc, EPI_BIAS, nullptr)) return rc;
        res = sc;
      }
      bf16* m1 = static_cast<bf16*>(mid_slot(0, act_bytes(m_in, mid)));
      bf16* m2 = static_cast<bf16*>(mid_slot(1, act_bytes(m_out, mid)));
      bf16* out = static_cast<bf16*>(big_slot((cur + 2) % 3, act_bytes(m_out, width)));
      cur_name = bname + ".conv1x1a";
      if (int rc = conv(x, hh, ww, bw.c0, m1, EPI_BIAS_RELU, nullptr)) return rc;
      if (mid <= 128 && m->fuse_tail) {
        // stages 1-2: 3x3 convolution + 1x1 expansion + residual in one kernel (tc_bottleneck.cu)
        if (!dry) {
          BneckPlan bp;
          ConvGeom g{B, hh, ww, mid, 3, 3, stride, 1, 1, ho, wo};
          if (fuse_sc) {
            if (int rc = bneck_halo_plan(&bp, m1, g, bw.c1.w, bw.c1.bias, bw.w3sc, bw.bias3sc, width, nullptr, out, x)) return rc;
            const double flops = 2.0 * m_out * mid * (9.0 * mid + 2.0 * width);
            const double bytes = 2.0 * m_in * mid * 2 + 2.0 * m_out * width + 2.0 * mid * (9.0 * mid + 2.0 * width);
            add(OPD_STEP_CONV, bname + ".tail(3x3+1x1b+shortcut)", flops, bytes, [bp](cudaStream_t s) { return bneck_launch(bp, s); });
          } else {
            if (int rc = bneck_plan(&bp, m1, g, bw.c1.w, bw.c1.bias, bw.c2.w, bw.c2.bias, width, res, out)) return rc;
            const double flops = 2.0 * m_out * mid * (9.0 * mid + width);
            const double bytes = 2.0 * m_in * mid + 4.0 * m_out * width + 2.0 * mid * (9.0 * mid + width);
            add(OPD_STEP_CONV, bname + ".tail(3x3+1x1b)", flops, bytes, [bp](cudaStream_t s) { return bneck_launch(bp, s); });
          }
        }
      } else {
        cur_name = bname + ".conv3x3";
        if (int rc = conv(m1, hh, ww, bw.c1, m2, EPI_BIAS_RELU, nullptr)) return rc;
        if (dual_sc) {
          cur_name = bname + ".conv1x1b+shortcut";
          if (!dry) {
            GemmPlan gp;
            ConvGeom g2{B, hh, ww, bw.shortcut.cin, 1, 1, stride, 0, 0, ho, wo};
            if (int rc = gemm_plan_linear_plus_shortcut(&gp, m2, mid, x, g2, bw.w3sc, out, width, EPI_BIAS_RELU, bw.bias3sc)) return rc;
            add(OPD_STEP_GEMM, cur_name, 2.0 * gp.M * gp.N * gp.K,
                2.0 * gp.M * mid + 2.0 * gp.M * g2.C + 2.0 * gp.N * gp.K + 2.0 * gp.M * gp.N,
                [gp](cudaStream_t st) { return gemm_launch(gp, st); });
          }
        } else {
          cur_name = bname + ".conv1x1b";
          if (int rc = conv(m2, ho, wo, bw.c2, out, EPI_BIAS_RES_RELU, res)) return rc;
        }
      }
      if (!dry) taps["stage" + std::to_string(s) + "." + std::to_string(l)] = {out, m_out, width, 0};
      x = out;
      cur = (cur + 2) % 3;
      hh = ho;
      ww = wo;
    }
  }

  // ---- transformer ----
  const int S = hh * ww;
  const int M = B * S, Mq = B * kQueries;
  // sine position table: [S, 256] shared by the batch, or [B, S, 256] (one per frame: it depends on the frame's mask) when padded
  const int pos_rows = mixed ? B * S : S;
  float* pos = static_cast<float*>(A.take((size_t)pos_rows * kD * sizeof(float)));
  const int mask_words = (S + 31) / 32 + 4;   // 32-key words per frame (+ slack for the tiles' whole-word reads past S)
  uint32_t* key_mask = mixed ? static_cast<uint32_t*>(A.take((size_t)B * mask_words * sizeof(uint32_t))) : nullptr;
  int32_t* fvalid = mixed ? static_cast<int32_t*>(A.take((size_t)B * 2 * sizeof(int32_t))) : nullptr;
  bf16* ex = static_cast<bf16*>(A.take(act_bytes(M, kD)));     // encoder stream
  bf16* ex1 = static_cast<bf16*>(A.take(act_bytes(M, kD)));    // after attention + LN
  bf16* exp_ = static_cast<bf16*>(A.take(act_bytes(M, kD)));   // stream + pos (q / k input)
  bf16* eqk = static_cast<bf16*>(A.take(act_bytes(M, 2 * kD)));
  bf16* ev = static_cast<bf16*>(A.take(act_bytes(M, kD)));
  bf16* eo = static_cast<bf16*>(A.take(act_bytes(M, kD)));
  bf16* ef = static_cast<bf16*>(A.take(act_bytes(M, kFFN)));
  bf16* memk = static_cast<bf16*>(A.take(act_bytes(M, kDec * kD)));
  bf16* memv = static_cast<bf16*>(A.take(act_bytes(M, kDec * kD)));
  bf16* dy = static_cast<bf16*>(A.take(act_bytes(Mq, kD)));
  bf16* dy1 = static_cast<bf16*>(A.take(act_bytes(Mq, kD)));
  bf16* dy2 = static_cast<bf16*>(A.take(act_bytes(Mq, kD)));
  bf16* dyp = static_cast<bf16*>(A.take(act_bytes(Mq, kD)));
  bf16* dqk = static_cast<bf16*>(A.take(act_bytes(Mq, 2 * kD)));
  bf16* dv = static_cast<bf16*>(A.take(act_bytes(Mq, kD)));
  bf16* dq = static_cast<bf16*>(A.take(act_bytes(Mq, kD)));
  bf16* dob = static_cast<bf16*>(A.take(act_bytes(Mq, kD)));
  bf16* df = static_cast<bf16*>(A.take(act_bytes(Mq, kFFN)));
  bf16* dout = static_cast<bf16*>(A.take(act_bytes(Mq, kD)));
  // layer outputs: written in place over the stream buffer, or one buffer per layer when debugging (taps)
  bf16* enc_out[kEnc];
  bf16* dec_out[kDec];
  for (int i = 0; i < kEnc; ++i) enc_out[i] = dbg ? static_cast<bf16*>(A.take(act_bytes(M, kD))) : ex;
  for (int i = 0; i < kDec; ++i) dec_out[i] = dbg ? static_cast<bf16*>(A.take(act_bytes(Mq, kD))) : dy;

  *bytes_out = (A.off + 1023) & ~(size_t)1023;
  if (dry) return OPD_OK;

  auto linear = [&](const bf16* a, int rows, const LinW& w, bf16* d, int epi, const bf16* res, const LnW* ln, bf16* d2,
                    const float* posv, int pos_rows, int pos_row0 = 0) -> int {
    GemmPlan gp;
    if (int rc = gemm_plan_linear(&gp, a, w.k, w.w, d, w.n, rows, w.n, w.k, epi, w.b, res, kD, ln ? ln->g : nullptr,
                                  ln ? ln->b : nullptr, d2, posv, pos_rows))
      return rc;
    gp.pos_row0 = pos_row0;   // a GEMM over rows [r0, r0 + rows) of the token matrix: pos row of its row 0
    add_gemm(gp);
    return OPD_OK;
  };
  // fc1 + ReLU + fc2 + residual + LayerNorm (+ pos) of one layer as ONE kernel (tc_mlp.cu): the [rows, 2048] hidden tensor stays on chip
  // (benchmarks/mlp_microbench.py, cta_group::2 pairs: 138 against 163 us at 67 200 rows, 39 against 56 us at 6 400; option 0 = two launches)
  const int mlp_opt = g_option_mlp_fused.load();
  // below 16 row tiles the two launches win: they spread fc1's 2048 columns over many SMs, the fused kernel has one CTA per row tile
  // (batch 1: 1.63 -> 1.75 ms per frame with it)
  auto use_mlp = [&](int rows) { return mlp_opt == 2 || (mlp_opt == 1 && (rows + 127) / 128 >= 16); };
  auto mlp = [&](const bf16* xa, int rows, const LinW& fc1, const LinW& fc2, const LnW& ln, bf16* d, bf16* d2, const float* posv,
                 int pos_rows_) -> int {
    OPD_REQUIRE(fc1.n == kFFN && fc1.k == kD && fc2.n == kD && fc2.k == kFFN, "detr: unexpected feed-forward shape");
    MlpPlan mp;
    if (int rc = mlp_plan(&mp, xa, fc1.w, fc1.b, fc2.w, fc2.b, ln.g, ln.b, d, d2, posv, pos_rows_, rows)) return rc;
    add(OPD_STEP_GEMM, cur_name, 4.0 * rows * kD * kFFN, 2.0 * rows * kD * (d2 ? 3 : 2) + 4.0 * kD * kFFN,
        [mp](cudaStream_t s) { return mlp_launch(mp, s); });
    prev_reversed = false;   // ascending row blocks
    return OPD_OK;
  };
  int attn_rc = OPD_OK;
  if (mixed) {
    OPD_REQUIRE(g_option_attention_tc.load(), "detr: mixed-size batches need the tcgen05 attention kernel (key-padding mask)");
    // pixel_mask -> feature mask by nearest interpolation (modeling_detr.py:281-283): frame b keeps the feature cells whose source
    // pixel lies inside its picture - a top-left rectangle (fh, fw); keys outside it are masked in every attention over the map
    std::vector<int32_t> fv;
    std::vector<uint32_t> words((size_t)B * mask_words, 0u);
    int b = 0;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
      int fh = 0, fw = 0;
      for (int y = 0; y < hh; ++y) fh += nearest_src(y, Hc, hh) < gin[gi].first;
      for (int x = 0; x < ww; ++x) fw += nearest_src(x, Wc, ww) < gin[gi].second;
      OPD_REQUIRE(fh >= 1 && fw >= 1, "detr: a frame of the mixed batch has no valid feature cell");
      for (int i = 0; i < groups[gi].n; ++i, ++b) {
        fv.push_back(fh);
        fv.push_back(fw);
        for (int y = 0; y < fh; ++y)
          for (int x = 0; x < fw; ++x) {
            const int k = y * ww + x;
            words[(size_t)b * mask_words + k / 32] |= 1u << (k % 32);
          }
      }
    }
    OPD_CUDA_OK(cudaMemcpy(fvalid, fv.data(), fv.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    OPD_CUDA_OK(cudaMemcpy(key_mask, words.data(), words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }
  auto attn = [&](const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, bf16* o, int Lq, int Lk) {
    const double flops = 4.0 * B * kHeads * (double)Lq * Lk * 32, bytes = 2.0 * B * kD * (2.0 * Lq + 2.0 * Lk);
    if (g_option_attention_tc.load()) {
      AttnPlan ap;
      attn_rc = attn_plan(&ap, q, ldq, k, ldk, v, ldv, o, kD, B, kHeads, Lq, Lk);
      if (mixed && Lk == S) {   // attention over the feature map (encoder self-attention, decoder cross-attention)
        ap.key_mask = key_mask;
        ap.key_mask_stride = mask_words;
      }
      add(OPD_STEP_ATTENTION, cur_name, flops, bytes, [ap](cudaStream_t s) { return attn_launch(ap, s); });
    } else {
      add(OPD_STEP_ATTENTION, cur_name, flops, bytes,
          [=](cudaStream_t s) { return launch_attention(q, ldq, k, ldk, v, ldv, o, kD, B, kHeads, Lq, Lk, s); });
    }
  };

  if (mixed)
    add(OPD_STEP_ELEMENTWISE, "pos_embed", 0.0, 4.0 * B * S * kD,
        [pos, B, hh, ww, fvalid](cudaStream_t s) { return launch_pos_embed_masked(pos, B, hh, ww, fvalid, s); });
  else
    add(OPD_STEP_ELEMENTWISE, "pos_embed", 0.0, 4.0 * S * kD, [pos, hh, ww](cudaStream_t s) { return launch_pos_embed(pos, hh, ww, s); });
  cur_name = "input_proj";
  taps["pos"] = {pos, pos_rows, kD, 1};
  // input_projection (+ pos for the first layer's q / k input)
  if (int rc = linear(x, M, m->input_proj, ex, EPI_BIAS, nullptr, nullptr, exp_, pos, pos_rows)) return rc;
  taps["enc_in"] = {ex, M, kD, 0};

  bf16* xin = ex;
  for (int i = 0; i < kEnc; ++i) {
    const EncW& e = m->enc[i];
    bf16* xout = enc_out[i];   // == xin unless debugging
    const std::string ln = "enc" + std::to_string(i);
    cur_name = ln + ".qk";
    if (int rc = linear(exp_, M, e.qk, eqk, EPI_BIAS, nullptr, nullptr, nullptr, nullptr, 0)) return rc;
    cur_name = ln + ".v";
    if (int rc = linear(xin, M, e.v, ev, EPI_BIAS, nullptr, nullptr, nullptr, nullptr, 0)) return rc;
    cur_name = ln + ".attn";
    attn(eqk, 2 * kD, eqk + kD, 2 * kD, ev, kD, eo, S, S);
    cur_name = ln + ".o+ln";
    if (int rc = linear(eo, M, e.o, ex1, EPI_BIAS_RES_LN, xin, &e.ln1, nullptr, nullptr, 0)) return rc;
    // (Measured and dropped: running the FFN in L2-sized row chunks, so that fc2 reads the hidden tensor from L2, costs more in
    // ramp-up / drain of the eight small launches than it gains: 0.19 -> 0.22 ms per layer.)
    // layer output overwrites the layer input stream (its last reader, the o_proj residual, has completed)
    if (use_mlp(M)) {
      cur_name = ln + ".mlp+ln";
      if (int rc = mlp(ex1, M, e.fc1, e.fc2, e.ln2, xout, exp_, pos, pos_rows)) return rc;
    } else {
      cur_name = ln + ".fc1";
      if (int rc = linear(ex1, M, e.fc1, ef, EPI_BIAS_RELU, nullptr, nullptr, nullptr, nullptr, 0)) return rc;
      cur_name = ln + ".fc2+ln";
      if (int rc = linear(ef, M, e.fc2, xout, EPI_BIAS_RES_LN, ex1, &e.ln2, exp_, pos, pos_rows)) return rc;
    }
    taps["enc" + std::to_string(i)] = {xout, M, kD, 0};
    xin = xout;
  }
  cur_name = "dec.cross_kv";
  const bf16* memory = xin;        // encoder output
  const bf16* memory_pos = exp_;   // + pos: keys of every cross attention
  if (int rc = linear(memory_pos, M, m->cross_k_all, memk, EPI_BIAS, nullptr, nullptr, nullptr, nullptr, 0)) return rc;
  if (int rc = linear(memory, M, m->cross_v_all, memv, EPI_BIAS, nullptr, nullptr, nullptr, nullptr, 0)) return rc;

  // The object queries enter layer 0 as zeros (+ the learned query positions), so layer 0's self-attention block and its cross-
  // attention query projection do not depend on the frame: decoder_init, sqk, sv, attention, so + LayerNorm and cq of layer 0 -
  // six launches - run ONCE per plan into buffers owned by the model (same kernels, same inputs: bit-identical results), and
  // the step starts at layer 0's cross attention.  opd_set_option("dec0_const", 0) plans them into every step (A/B, tests).
  const bool const0 = g_option_dec0_const.load() != 0;
  bf16 *c_y = dy, *c_yp = dyp, *c_y1 = dy1, *c_q = dq;   // layer-0 inputs / outputs of the prologue
  if (const0) {
    const size_t each = (act_bytes(Mq, kD) + 1023) & ~(size_t)1023;
    void*& buf = m->dec0_bufs[B];   // contents depend on the weights and B only: plans of the same batch size share it
    if (!buf) OPD_CUDA_OK(cudaMalloc(&buf, 5 * each));
    uint8_t* cb = static_cast<uint8_t*>(buf);
    c_y = reinterpret_cast<bf16*>(cb);               // zeros
    bf16* c_yp0 = reinterpret_cast<bf16*>(cb + each);   // query positions
    c_y1 = reinterpret_cast<bf16*>(cb + 2 * each);   // after self attention + LayerNorm
    c_yp = reinterpret_cast<bf16*>(cb + 3 * each);   // ... + query positions
    c_q = reinterpret_cast<bf16*>(cb + 4 * each);    // cross-attention queries
    sink = &plan->once;
    add(OPD_STEP_ELEMENTWISE, "decoder_init", 0.0, 4.0 * Mq * kD,
        [m, c_y, c_yp0, B](cudaStream_t s) { return launch_decoder_init(c_y, c_yp0, m->qpos, B, kQueries, s); });
    const DecW& d = m->dec[0];
    cur_name = "dec0.const";
    if (int rc = linear(c_yp0, Mq, d.sqk, dqk, EPI_BIAS, nullptr, nullptr, nullptr, nullptr, 0)) return rc;
    if (int rc = linear(c_y, Mq, d.sv, dv, EPI_BIAS, nullptr, nullptr, nullptr, nullptr, 0)) return rc;
    attn(dqk, 2 * kD, dqk + kD, 2 * kD, dv, kD, dob, kQueries, kQueries);
    if (int rc = linear(dob, Mq, d.so, c_y1, EPI_BIAS_RES_LN, c_y, &d.ln1, c_yp, m->qpos, kQueries)) return rc;
    if (int rc = linear(c_yp, Mq, d.cq, c_q, EPI_BIAS, nullptr, nullptr, nullptr, nullptr, 0)) return rc;
    sink = &steps;
  } else {
    add(OPD_STEP_ELEMENTWISE, "decoder_init", 0.0, 4.0 * Mq * kD,
        [m, dy, dyp, B](cudaStream_t s) { return launch_decoder_init(dy, dyp, m->qpos, B, kQueries, s); });
  }
  bf16* yin = dy;
  for (int i = 0; i < kDec; ++i) {
    const DecW& d = m->dec[i];
    bf16* yout = dec_out[i];
    cur_name = "dec" + std::to_string(i);
    const bool pre = const0 && i == 0;   // this layer's self-attention block is in the prologue
    bf16* y1 = pre ? c_y1 : dy1;
    bf16* q = pre ? c_q : dq;
    if (!pre) {
      // self attention
      if (int rc = linear(dyp, Mq, d.sqk, dqk, EPI_BIAS, nullptr, nullptr, nullptr, nullptr, 0)) return rc;
      if (int rc = linear(yin, Mq, d.sv, dv, EPI_BIAS, nullptr, nullptr, nullptr, nullptr, 0)) return rc;
      attn(dqk, 2 * kD, dqk + kD, 2 * kD, dv, kD, dob, kQueries, kQueries);
      if (int rc = linear(dob, Mq, d.so, y1, EPI_BIAS_RES_LN, yin, &d.ln1, dyp, m->qpos, kQueries)) return rc;
      // cross attention
      if (int rc = linear(dyp, Mq, d.cq, q, EPI_BIAS, nullptr, nullptr, nullptr, nullptr, 0)) return rc;
    }
    attn(q, kD, memk + i * kD, kDec * kD, memv + i * kD, kDec * kD, dob, kQueries, S);
    if (int rc = linear(dob, Mq, d.co, dy2, EPI_BIAS_RES_LN, y1, &d.ln2, nullptr, nullptr, 0)) return rc;
    // FFN
    if (use_mlp(Mq)) {
      if (int rc = mlp(dy2, Mq, d.fc1, d.fc2, d.ln3, yout, dyp, m->qpos, kQueries)) return rc;
    } else {
      if (int rc = linear(dy2, Mq, d.fc1, df, EPI_BIAS_RELU, nullptr, nullptr, nullptr, nullptr, 0)) return rc;
      if (int rc = linear(df, Mq, d.fc2, yout, EPI_BIAS_RES_LN, dy2, &d.ln3, dyp, m->qpos, kQueries)) return rc;
    }
    taps["dec" + std::to_string(i)] = {yout, Mq, kD, 0};
    yin = yout;
  }
  add(OPD_STEP_ELEMENTWISE, "dec.final_ln", 0.0, 4.0 * Mq * kD,
      [m, yin, dout, Mq](cudaStream_t s) { return launch_layernorm(yin, m->dec_norm.g, m->dec_norm.b, dout, Mq, s); });
  taps["dec_out"] = {dout, Mq, kD, 0};
  if (attn_rc != OPD_OK) return attn_rc;
  add(OPD_STEP_HEADS, "heads", 2.0 * Mq * kD * (kClasses + 2 * kD + 4), 2.0 * Mq * kD + 4.0 * Mq * (kClasses + 4),
      [m, dout, Mq](cudaStream_t s) { return launch_heads(dout, m->heads, m->cur_logits, m->cur_boxes, Mq, s); });
  return OPD_OK;
}

}  // namespace
}  // namespace opd

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

int opd_detr_create(const opd_tensor_f32* tensors, int32_t n_tensors, int32_t device, opd_detr** out) {
  OPD_REQUIRE(tensors && n_tensors > 0 && out, "opd_detr_create: NULL argument");
  opd::DeviceGuard guard(device);
  OPD_CUDA_OK(guard.err);
  opd_detr* m = new opd_detr();
  m->device = device;
  const int rc = opd::load_weights(m, tensors, n_tensors);
  if (rc != OPD_OK) {
    opd_detr_destroy(m);
    return rc;
  }
  OPD_CUDA_OK(cudaDeviceSynchronize());
  *out = m;
  return OPD_OK;
}

void opd_detr_destroy(opd_detr* m) {
  if (!m) return;
  for (void* p : m->allocs) cudaFree(p);
  for (auto& kv : m->dec0_bufs)
    if (kv.second) cudaFree(kv.second);
  delete m;
}

int opd_detr_set_debug(opd_detr* m, int32_t debug) {
  OPD_REQUIRE(m, "opd_detr_set_debug: NULL handle");
  m->debug = debug;
  m->drop_plans();
  return OPD_OK;
}

int opd_detr_set_fusion(opd_detr* m, int32_t fuse_bottleneck_tail) {
  OPD_REQUIRE(m, "opd_detr_set_fusion: NULL handle");
  m->fuse_tail = fuse_bottleneck_tail & 1;
  m->stem_halo = (fuse_bottleneck_tail & 2) ? 0 : 1;
  m->fuse_shortcut = (fuse_bottleneck_tail & 4) ? 0 : 1;   // bit 2 set: separate shortcut kernel   // bit 1 set: im2col stem over the 64-lane layout (A/B comparison)
  m->drop_plans();
  return OPD_OK;
}

int opd_detr_set_resize(opd_detr* m, int32_t do_resize) {
  OPD_REQUIRE(m, "opd_detr_set_resize: NULL handle");
  m->do_resize = do_resize;
  m->drop_plans();
  return OPD_OK;
}

int opd_detr_input_shape(int32_t H0, int32_t W0, int32_t* H_in, int32_t* W_in, int32_t* h_feat, int32_t* w_feat) {
  OPD_REQUIRE(H0 > 0 && W0 > 0, "opd_detr_input_shape: bad frame size %dx%d", H0, W0);
  const opd::Shapes s = opd::shapes_for(H0, W0);
  if (H_in) *H_in = s.Hin;
  if (W_in) *W_in = s.Win;
  if (h_feat) *h_feat = s.h[3];
  if (w_feat) *w_feat = s.w[3];
  return OPD_OK;
}

int opd_detr_workspace_bytes(const opd_detr* m, int32_t B, int32_t H0, int32_t W0, size_t* bytes) {
  OPD_REQUIRE(m && bytes && B > 0 && H0 > 0 && W0 > 0, "opd_detr_workspace_bytes: bad argument");
  opd::Plan scratch;
  return opd::build_plan(const_cast<opd_detr*>(m), {opd::FrameGroup{B, H0, W0}}, nullptr, &scratch, bytes);
}

namespace {
int groups_from_abi(const opd_frame_group* groups, int32_t n_groups, std::vector<opd::FrameGroup>* out, bool need_frames) {
  OPD_REQUIRE(groups && n_groups >= 1 && n_groups <= 64, "detr: 1 .. 64 frame groups (got %d)", n_groups);
  long long B = 0;
  for (int i = 0; i < n_groups; ++i) {
    OPD_REQUIRE(groups[i].n > 0 && groups[i].H0 > 0 && groups[i].W0 > 0 && (!need_frames || groups[i].frames_dev),
                "detr: frame group %d is empty or has no frames", i);
    out->push_back(opd::FrameGroup{groups[i].n, groups[i].H0, groups[i].W0});
    B += groups[i].n;
  }
  OPD_REQUIRE(B * opd::kQueries < (1 << 24), "detr: bad batch %lld", B);
  return OPD_OK;
}

// finds or builds the plan of this batch description (most recently used first) and makes it current
int select_plan(opd_detr* m, const std::vector<opd::FrameGroup>& groups, void* workspace_dev, size_t workspace_bytes) {
  auto hit = m->plans.begin();
  for (; hit != m->plans.end(); ++hit)
    if (hit->groups == groups && hit->ws == workspace_dev && !hit->steps.empty()) break;
  if (hit != m->plans.end()) {
    m->plans.splice(m->plans.begin(), m->plans, hit);   // list nodes do not move: m->plan stays valid
  } else {
    m->plan = nullptr;
    m->plans.emplace_front();
    opd::Plan& fresh = m->plans.front();
    size_t need = 0;
    int rc = opd::build_plan(m, groups, workspace_dev, &fresh, &need);
    if (rc == OPD_OK && need > workspace_bytes)
      rc = opd::fail(OPD_ERR_INVALID, "opd_detr_forward: workspace of %zu bytes, %zu needed", workspace_bytes, need);
    if (rc != OPD_OK) {
      m->plans.pop_front();
      return rc;
    }
    fresh.groups = groups;
    fresh.B = 0;
    for (const auto& g : groups) fresh.B += g.n;
    fresh.H0 = groups[0].H0; fresh.W0 = groups[0].W0; fresh.ws = workspace_dev; fresh.ws_bytes = need;
    while (m->plans.size() > kMaxPlans) m->plans.pop_back();
  }
  m->plan = &m->plans.front();
  return OPD_OK;
}

int run_plan(opd_detr* m, cudaStream_t s) {
  opd::Plan& p = *m->plan;
  if (!p.once_done) {   // frame-independent prologue: on the same stream, ahead of the first step of this plan
    for (auto& step : p.once)
      if (int rc = step.run(s)) return rc;
    p.once_done = true;
  }
  for (auto& step : p.steps)
    if (int rc = step.run(s)) {   // name the layer: a launch error otherwise only carries the kernel's source line
      const std::string why = opd::last_error_ref();
      return opd::fail(rc, "step '%s': %s", step.name.c_str(), why.c_str());
    }
  return OPD_OK;
}
}  // namespace

int opd_detr_workspace_bytes_mixed(const opd_detr* m, const opd_frame_group* groups, int32_t n_groups, size_t* bytes) {
  OPD_REQUIRE(m && bytes, "opd_detr_workspace_bytes_mixed: NULL argument");
  std::vector<opd::FrameGroup> g;
  if (int rc = groups_from_abi(groups, n_groups, &g, false)) return rc;
  opd::Plan scratch;
  return opd::build_plan(const_cast<opd_detr*>(m), g, nullptr, &scratch, bytes);
}

int opd_detr_forward_mixed(opd_detr* m, const opd_frame_group* groups, int32_t n_groups, void* workspace_dev, size_t workspace_bytes,
                           float* logits_dev, float* boxes_dev, void* stream) {
  OPD_REQUIRE(m && workspace_dev && logits_dev && boxes_dev, "opd_detr_forward_mixed: NULL argument");
  OPD_REQUIRE((reinterpret_cast<uintptr_t>(workspace_dev) & 1023) == 0, "opd_detr_forward_mixed: workspace must be 1024-byte aligned");
  std::vector<opd::FrameGroup> g;
  if (int rc = groups_from_abi(groups, n_groups, &g, true)) return rc;
  opd::DeviceGuard guard(m->device);
  OPD_CUDA_OK(guard.err);
  if (int rc = select_plan(m, g, workspace_dev, workspace_bytes)) return rc;
  m->cur_group_frames.clear();
  m->cur_group_bgr.clear();
  for (int i = 0; i < n_groups; ++i) {
    m->cur_group_frames.push_back(groups[i].frames_dev);
    m->cur_group_bgr.push_back(groups[i].frames_are_bgr);
  }
  m->cur_frames = groups[0].frames_dev;   // a one-group batch runs the plain path
  m->cur_bgr = groups[0].frames_are_bgr;
  m->cur_logits = logits_dev;
  m->cur_boxes = boxes_dev;
  return run_plan(m, static_cast<cudaStream_t>(stream));
}

int opd_detr_forward(opd_detr* m, const uint8_t* frames_dev, int32_t B, int32_t H0, int32_t W0, int32_t frames_are_bgr,
                     void* workspace_dev, size_t workspace_bytes, float* logits_dev, float* boxes_dev, void* stream) {
  OPD_REQUIRE(m && frames_dev && workspace_dev && logits_dev && boxes_dev, "opd_detr_forward: NULL argument");
  OPD_REQUIRE(B > 0 && (long long)B * opd::kQueries < (1 << 24), "opd_detr_forward: bad batch %d", B);
  OPD_REQUIRE((reinterpret_cast<uintptr_t>(workspace_dev) & 1023) == 0, "opd_detr_forward: workspace must be 1024-byte aligned");
  opd::DeviceGuard guard(m->device);   // the handle's device, whatever the caller's current device is
  OPD_CUDA_OK(guard.err);
  if (int rc = select_plan(m, {opd::FrameGroup{B, H0, W0}}, workspace_dev, workspace_bytes)) return rc;
  m->cur_frames = frames_dev;
  m->cur_bgr = frames_are_bgr;
  m->cur_logits = logits_dev;
  m->cur_boxes = boxes_dev;
  return run_plan(m, static_cast<cudaStream_t>(stream));
}

int opd_detr_profile(opd_detr* m, void* stream, int32_t max_steps, int32_t* n_steps, int32_t* kinds, double* flops,
                     double* bytes, float* ms, char* names, int32_t name_stride) {
  OPD_REQUIRE(m && n_steps, "opd_detr_profile: NULL argument");
  OPD_REQUIRE(m->plan && !m->plan->steps.empty() && m->cur_frames, "opd_detr_profile: run opd_detr_forward first");
  opd::Plan& p = *m->plan;
  opd::DeviceGuard guard(m->device);
  OPD_CUDA_OK(guard.err);
  const int n = (int)p.steps.size();
  *n_steps = n;
  if (max_steps <= 0) return OPD_OK;
  OPD_REQUIRE(max_steps >= n && kinds && flops && bytes && ms, "opd_detr_profile: %d steps, room for %d", n, max_steps);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) OPD_CUDA_OK(cudaEventCreate(&e));
  int rc = OPD_OK;
  OPD_CUDA_OK(cudaEventRecord(ev[0], s));
  for (int i = 0; i < n && rc == OPD_OK; ++i) {
    rc = p.steps[i].run(s);
    if (cudaEventRecord(ev[i + 1], s) != cudaSuccess && rc == OPD_OK) rc = opd::fail(OPD_ERR_CUDA, "cudaEventRecord failed");
  }
  if (rc == OPD_OK && cudaStreamSynchronize(s) != cudaSuccess) rc = opd::fail(OPD_ERR_CUDA, "profile: stream sync failed");
  for (int i = 0; i < n && rc == OPD_OK; ++i) {
    kinds[i] = p.steps[i].kind;
    flops[i] = p.steps[i].flops;
    bytes[i] = p.steps[i].bytes;
    cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
    if (names && name_stride > 0) {
      std::strncpy(names + (size_t)i * name_stride, p.steps[i].name.c_str(), name_stride - 1);
      names[(size_t)i * name_stride + name_stride - 1] = 0;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

int opd_detr_tap(const opd_detr* m, const char* name, const void** ptr_dev, int64_t* rows, int64_t* cols, int32_t* is_f32) {
  OPD_REQUIRE(m && name && ptr_dev && rows && cols && is_f32, "opd_detr_tap: NULL argument");
  OPD_REQUIRE(m->plan, "opd_detr_tap: run a forward first");
  auto it = m->plan->taps.find(name);
  OPD_REQUIRE(it != m->plan->taps.end(), "opd_detr_tap: no activation named '%s' (run a forward first)", name);
  *ptr_dev = it->second.ptr;
  *rows = it->second.rows;
  *cols = it->second.cols;
  *is_f32 = it->second.is_f32;
  return OPD_OK;
}

int opd_detr_tap_copy(const opd_detr* m, const char* name, void* dst_dev, size_t bytes, void* stream) {
  const void* src = nullptr;
  int64_t rows = 0, cols = 0;
  int32_t f32 = 0;
  if (int rc = opd_detr_tap(m, name, &src, &rows, &cols, &f32)) return rc;
  const size_t esz = f32 == 1 ? 4 : (f32 == 2 ? 1 : 2);
  OPD_REQUIRE(dst_dev && bytes == (size_t)rows * cols * esz, "opd_detr_tap_copy: '%s' is %lld x %lld of %zu-byte elements", name,
              (long long)rows, (long long)cols, esz);
  OPD_CUDA_OK(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return OPD_OK;
}

int opd_detr_postprocess(const float* logits_dev, const float* boxes_dev, int32_t B, int32_t Q, int32_t C, int32_t H0,
                         int32_t W0, float threshold, int32_t person_label, float* scores_dev, int32_t* labels_dev,
                         float* xyxy_dev, double* det_xywh_dev, float* det_score_dev, double* det_foot_dev,
                         int32_t* det_query_dev, int32_t* n_keep_dev, int32_t* det_slot_dev, int32_t slot_base,
                         void* stream) {
  OPD_REQUIRE(logits_dev && boxes_dev && scores_dev && labels_dev && xyxy_dev && det_xywh_dev && det_score_dev &&
                  det_foot_dev && det_query_dev && n_keep_dev,
              "opd_detr_postprocess: NULL argument");
  OPD_REQUIRE(B > 0 && Q > 0 && C > 1, "opd_detr_postprocess: bad shape B=%d Q=%d C=%d", B, Q, C);
  return opd::launch_postprocess(logits_dev, boxes_dev, B, Q, C, H0, W0, threshold, person_label, scores_dev, labels_dev,
                                 xyxy_dev, det_xywh_dev, det_score_dev, det_foot_dev, det_query_dev, n_keep_dev, det_slot_dev,
                                 slot_base, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
