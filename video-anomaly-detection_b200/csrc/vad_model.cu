// Model-level entry points of libvad_b200.so: ONE C call per reference method (include/vad_b200.h, "Model-level entry
// points").  Each call walks the layer schedule of models/autoencoder.py / models/video_autoencoder.py and enqueues the
// fused layer kernels (vad_conv_layer & co.) on the caller's stream; activations, score partials and the ConvLSTM step
// counters come out of the caller's workspace, so a call allocates nothing, never synchronises and is re-entrant per
// (stream, workspace).  The same schedule code runs in a "dry" mode that only measures the workspace.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "vad_internal.h"

namespace {
using namespace vad;

constexpr float kLeaky = 0.2f, kRelu = 0.0f, kIdent = 1.0f;
constexpr size_t kAlign = 256;
inline size_t align_up(size_t n) { return (n + kAlign - 1) & ~(kAlign - 1); }

// ---- workspace ---------------------------------------------------------------------------------------------------
// Two kinds of regions: "persist" (lives for the whole call: score partials, ConvLSTM sequences / state / counters) and
// two ping-pong regions for layer activations (layer i reads one and writes the other; the stream orders the reuse).
struct Plan {
  size_t persist = 0;
  size_t ping[2] = {0, 0};
  size_t total() const { return persist + ping[0] + ping[1]; }
};

struct Arena {
  bool dry = true;
  char* base = nullptr;
  Plan plan;            // dry: being measured; real: the measured plan (region sizes)
  size_t persist_off = 0;

  void* persist(size_t n) {
    n = align_up(n);
    const size_t off = persist_off;
    persist_off += n;
    if (dry) {
      plan.persist = persist_off;
      return nullptr;
    }
    return base + off;
  }
  void* ping(int which, size_t n) {
    if (dry) {
      plan.ping[which] = std::max(plan.ping[which], align_up(n));
      return nullptr;
    }
    return base + plan.persist + (which ? plan.ping[0] : 0);
  }
};

// ---- per-layer profiling (vad_profile_enable / vad_profile_dump) ----------------------------------------------------
struct ProfEntry {
  std::string name;
  cudaEvent_t e0, e1;
};
std::mutex g_prof_mutex;
bool g_prof_on = false;
std::vector<ProfEntry> g_prof;

struct ProfScope {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t stream;
  const char* name;
  ProfScope(const char* n, cudaStream_t s) : stream(s), name(n) {
    if (!g_prof_on) return;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
      e0 = e1 = nullptr;
      return;
    }
    cudaEventRecord(e0, stream);
  }
  void cancel() {  // nothing was launched after all
    if (!e0) return;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    e0 = e1 = nullptr;
  }
  ~ProfScope() {
    if (!e0) return;
    cudaEventRecord(e1, stream);
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof.push_back({name, e0, e1});
  }
};

struct Ctx {
  Arena A;
  cudaStream_t stream = nullptr;
  bool dry() const { return A.dry; }
};

#define VAD_TRY(expr)              \
  do {                             \
    const int rc_ = (expr);        \
    if (rc_ != VAD_OK) return rc_; \
  } while (0)

inline bool weights_ok(const vad_gemm_weights& w) { return w.w && w.bias && w.ntaps > 0 && w.ctap > 0 && w.n_total > 0; }
inline bool first_ok(const vad_first_weights& w) { return w.w && w.bias && w.cout > 0; }
inline bool hw_ok(int H, int W) { return H > 0 && W > 0 && H % 16 == 0 && W % 16 == 0; }

vad_conv_desc base_desc(const vad_gemm_weights& w, const void* src, int B, int H, int W, int epi, float slope) {
  vad_conv_desc d;
  std::memset(&d, 0, sizeof(d));
  d.src0 = src;
  d.c0 = w.ctap;
  d.T0 = d.T1 = 1;
  d.B = B; d.H = H; d.W = W;
  d.ntaps = w.ntaps;
  d.weight = w.w;
  d.weight_kx = w.w_kx;
  d.bias = w.bias;
  d.w_ctap = w.ctap;
  d.n_total = w.n_total;
  d.cout = w.cout;
  d.epilogue = epi;
  d.slope = slope;
  return d;
}

// conv 3x3 (or 1x1) + folded BN + activation (+ 2x2 max-pool): bf16 NHWC -> bf16 NHWC
int conv(Ctx& c, const char* name, const vad_gemm_weights& w, const void* src, int B, int H, int W, void* out,
         float slope, bool pool) {
  if (c.dry()) return VAD_OK;
  vad_conv_desc d = base_desc(w, src, B, H, W, pool ? VAD_EPI_POOL : VAD_EPI_STORE, slope);
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  d.out = out;
  d.out_frame_stride = static_cast<long long>(Ho) * Wo * w.n_total;
  d.out_cpitch = w.n_total;
  ProfScope p(name, c.stream);
  if (w.w_pair && w.bias_pair && w.ntaps == 9 && w.ctap == 32 && w.n_total <= 64 && W % 2 == 0 && W >= 32 && H >= 16) {
    // pixel-pair view (include/vad_b200.h `pair_fold`): N = 2*Cout lifts the layer off the N = 32 A-operand ceiling
    vad_conv_desc dp = d;
    dp.c0 = 64;
    dp.W = W / 2;
    dp.weight = w.w_pair;
    dp.weight_kx = nullptr;
    dp.bias = w.bias_pair;
    dp.w_ctap = 64;
    dp.n_total = 2 * w.n_total;
    dp.cout = 2 * w.n_total;
    dp.pair_fold = 1;
    dp.out_cpitch = pool ? w.n_total : 2 * w.n_total;
    const int rc = vad_conv_layer(&dp, c.stream);
    if (rc != VAD_ERR_UNSUPPORTED) return rc;
  }
  return vad_conv_layer(&d, c.stream);
}

// transposed conv k2 s2 + folded BN + ReLU: bf16 NHWC [B,H,W,Cin] -> [B,2H,2W,Cout]
int convt(Ctx& c, const char* name, const vad_gemm_weights& w, const void* src, int B, int H, int W, void* out) {
  if (c.dry()) return VAD_OK;
  vad_conv_desc d = base_desc(w, src, B, H, W, VAD_EPI_CONVT, kRelu);
  d.out = out;
  d.out_frame_stride = 4LL * H * W * w.cout;
  d.out_cpitch = w.cout;
  ProfScope p(name, c.stream);
  return vad_conv_layer(&d, c.stream);
}

int first_conv(Ctx& c, const vad_first_weights& w, const float* x, int B, int H, int W, bool pool, void* out) {
  if (c.dry()) return VAD_OK;
  ProfScope p("first_conv", c.stream);
  if (pool && w.w_pf && w.cout == 32) return vad_first_conv_pool(x, w.w_pf, w.bias, kLeaky, B, H, W, out, c.stream);
  if (w.w_tc && w.cout == 32) return vad_first_conv_tc(x, w.w_tc, w.bias, kLeaky, pool ? 1 : 0, B, H, W, out, c.stream);
  return vad_first_conv(x, w.w, w.bias, w.cout, kLeaky, pool ? 1 : 0, B, H, W, out, c.stream);
}

// ---- scoring outputs --------------------------------------------------------------------------------------------------
struct ScoreOut {
  const float* x = nullptr;  // model input the reconstruction is compared with
  float* recon = nullptr;
  float* score = nullptr;
  float* minmax = nullptr;
  float* heat = nullptr;
  float* partials = nullptr;  // workspace
  size_t partials_bytes = 0;
};

// upper bound of the per-tile partials of any score kernel for `frames` frames of Ho x Wo output pixels: every tiling in
// use covers at least 8 x 6 valid output pixels per tile (kx kernel: 16 x 6; fused image tail: 14 x 30; others: 128 px)
size_t partials_bound(int frames, int Ho, int Wo) {
  return static_cast<size_t>(frames) * ((Ho + 7) / 8) * ((Wo + 5) / 6) * 64;
}

int prepare_score(Ctx& c, ScoreOut& s, int frames, int Ho, int Wo) {
  s.partials_bytes = partials_bound(frames, Ho, Wo);
  s.partials = static_cast<float*>(c.A.persist(s.partials_bytes));
  if (!s.score) s.score = static_cast<float*>(c.A.persist(static_cast<size_t>(frames) * 4));  // finalize needs a target
  return VAD_OK;
}

int finalize(Ctx& c, const ScoreOut& s, int frames, int tiles, int Ho, int Wo) {
  if (tiles <= 0 || tiles % frames != 0) return VAD_ERR_ARG;
  ProfScope p("finalize", c.stream);
  return vad_score_finalize(s.partials, frames, 4 * (tiles / frames), Ho, Wo, s.score, s.minmax, c.stream);
}

// last decoder layer (3x3 conv or k2s2 transposed conv to 3 channels) + tanh + fused (x - recon)^2 reduction
int score_layer(Ctx& c, const char* name, const vad_gemm_weights& w, const void* src, int frames, int H, int W, int epi,
                const ScoreOut& s, int Ho, int Wo) {
  if (c.dry()) return VAD_OK;
  vad_conv_desc d = base_desc(w, src, frames, H, W, epi, kIdent);
  d.x = s.x; d.recon = s.recon; d.heat = s.heat; d.partials = s.partials;
  const int tiles = vad_conv_layer_tiles(&d);
  if (tiles <= 0) return tiles < 0 ? tiles : VAD_ERR_ARG;
  if (static_cast<size_t>(tiles) * 64 > s.partials_bytes) return VAD_ERR_WORKSPACE;
  {
    ProfScope p(name, c.stream);
    VAD_TRY(vad_conv_layer(&d, c.stream));
  }
  return finalize(c, s, frames, tiles, Ho, Wo);
}

// ---- image model ------------------------------------------------------------------------------------------------------
// Encoder.forward (models/autoencoder.py:81-86): x fp32 [B,3,H,W] -> bf16 NHWC [B,H/16,W/16,latent] in ping region *which
int image_encode(Ctx& c, const vad_image_model& m, const float* x, int B, int H, int W, void** z, int* which_out) {
  static const char* names[7] = {"enc1.3", "enc2.0", "enc2.3", "enc3.0", "enc3.3", "enc4.0", "enc4.3"};
  int which = 0;
  int h = H, w = W;
  void* cur = nullptr;
  int first_blk = 0;
  const vad_gemm_weights& w13 = m.enc[0];
  const bool fuse1 = !(m.flags & VAD_FLAG_NO_FUSED_ENC1) && m.enc1_0.w_tc && m.enc1_0.cout == 32 && w13.w_pair &&
                     w13.bias_pair && w13.ntaps == 9 && w13.ctap == 32 && w13.n_total == 32;
  if (fuse1) {
    // enc1.0 + enc1.3 + pool in one kernel: the 32-channel full-resolution tensor never reaches HBM
    cur = c.A.ping(which, static_cast<size_t>(B) * (H / 2) * (W / 2) * 32 * 2);
    if (!c.dry()) {
      ProfScope p("enc1.0+1.3", c.stream);
      VAD_TRY(vad_enc1_fused(x, m.enc1_0.w_tc, m.enc1_0.bias, w13.w_pair, w13.bias_pair, kLeaky, B, H, W, cur, c.stream));
    }
    h /= 2;
    w /= 2;
    first_blk = 1;
  } else {
    cur = c.A.ping(which, static_cast<size_t>(B) * H * W * 32 * 2);
    VAD_TRY(first_conv(c, m.enc1_0, x, B, H, W, false, cur));
  }
  for (int blk = first_blk; blk < 4; ++blk) {
    if (blk > 0) {
      const vad_gemm_weights& w0 = m.enc[2 * blk - 1];
      void* nxt = c.A.ping(which ^ 1, static_cast<size_t>(B) * h * w * w0.n_total * 2);
      VAD_TRY(conv(c, names[2 * blk - 1], w0, cur, B, h, w, nxt, kLeaky, false));
      cur = nxt;
      which ^= 1;
    }
    const vad_gemm_weights& w3 = m.enc[2 * blk];
    void* nxt = c.A.ping(which ^ 1, static_cast<size_t>(B) * (h / 2) * (w / 2) * w3.n_total * 2);
    VAD_TRY(conv(c, names[2 * blk], w3, cur, B, h, w, nxt, kLeaky, true));
    cur = nxt;
    which ^= 1;
    h /= 2;
    w /= 2;
  }
  *z = cur;
  *which_out = which;
  return VAD_OK;
}

// Decoder.forward (:141-146) with the error reduction of get_reconstruction_error (:214-221) fused onto its last layer.
// z bf16 NHWC [B,h,w,latent] living in ping region `which` (-1: elsewhere).
int image_decode_score(Ctx& c, const vad_image_model& m, const void* z, int which, int B, int h, int w, ScoreOut& s) {
  static const char* names[8] = {"dec1.0", "dec1.3", "dec2.0", "dec2.3", "dec3.0", "dec3.3", "dec4.0", "dec4.3+score"};
  const int Ho = 16 * h, Wo = 16 * w;
  VAD_TRY(prepare_score(c, s, B, Ho, Wo));
  const vad_gemm_weights& w40 = m.dec[6];
  const vad_gemm_weights& w43 = m.dec[7];
  const bool fuse = !(m.flags & VAD_FLAG_NO_FUSED_TAIL) && w43.w_kx && w40.ctap == 32 && w40.n_total == 128 &&
                    w40.cout == 32 && w43.ctap == 32 && w43.n_total == 16;
  const void* cur = z;
  int nxt_which = which == 0 ? 1 : 0;
  for (int blk = 0; blk < 4; ++blk) {
    const vad_gemm_weights& wt = m.dec[2 * blk];
    if (blk == 3 && fuse) {
      // dec4.0 + dec4.3 + score in one kernel: the 32-channel full-resolution tensor never reaches HBM
      if (c.dry()) return VAD_OK;
      vad_conv_desc d = base_desc(wt, cur, B, h, w, VAD_EPI_CONVT, kRelu);
      d.x = s.x; d.recon = s.recon; d.heat = s.heat; d.partials = s.partials;
      const int tiles = vad_convt_conv_score_tiles(&d);
      if (tiles <= 0) return tiles < 0 ? tiles : VAD_ERR_ARG;
      if (static_cast<size_t>(tiles) * 64 > s.partials_bytes) return VAD_ERR_WORKSPACE;
      {
        ProfScope p("dec4.0+4.3+score", c.stream);
        VAD_TRY(vad_convt_conv_score(&d, w43.w_kx, w43.bias, c.stream));
      }
      return finalize(c, s, B, tiles, Ho, Wo);
    }
    void* up = c.A.ping(nxt_which, static_cast<size_t>(B) * 4 * h * w * wt.cout * 2);
    VAD_TRY(convt(c, names[2 * blk], wt, cur, B, h, w, up));
    cur = up;
    nxt_which ^= 1;
    h *= 2;
    w *= 2;
    if (blk < 3) {
      const vad_gemm_weights& wc = m.dec[2 * blk + 1];
      void* nxt = c.A.ping(nxt_which, static_cast<size_t>(B) * h * w * wc.n_total * 2);
      VAD_TRY(conv(c, names[2 * blk + 1], wc, cur, B, h, w, nxt, kRelu, false));
      cur = nxt;
      nxt_which ^= 1;
    }
  }
  return score_layer(c, names[7], w43, cur, B, h, w, VAD_EPI_TANH_SCORE, s, Ho, Wo);
}

int image_forward_impl(Ctx& c, const vad_image_model& m, const float* x, int B, int H, int W, float* recon,
                       float* latent, float* score, float* minmax, float* heat) {
  void* z = nullptr;
  int which = 0;
  VAD_TRY(image_encode(c, m, x, B, H, W, &z, &which));
  const int h = H / 16, w = W / 16;
  if (latent && !c.dry()) {
    ProfScope p("latent_out", c.stream);
    VAD_TRY(vad_nhwc_bf16_to_nchw_f32(z, B, h, w, m.enc[6].n_total, latent, c.stream));
  }
  if (!recon && !score && !minmax && !heat) return VAD_OK;  // get_latent / Encoder.forward
  ScoreOut s;
  s.x = x; s.recon = recon; s.score = score; s.minmax = minmax; s.heat = heat;
  return image_decode_score(c, m, z, which, B, h, w, s);
}

int image_decode_impl(Ctx& c, const vad_image_model& m, const float* zf, int B, int h, int w, float* recon) {
  const int latent = m.dec[0].ctap;
  void* z = c.A.ping(0, static_cast<size_t>(B) * h * w * latent * 2);
  const size_t xbytes = static_cast<size_t>(B) * 3 * (16 * h) * (16 * w) * 4;
  float* xzero = static_cast<float*>(c.A.persist(xbytes));  // the fused score epilogue needs an input to compare with
  if (!c.dry()) {
    VAD_TRY(vad_nchw_f32_to_nhwc_bf16(zf, B, latent, h, w, z, c.stream));
    VAD_TRY(static_cast<int>(cudaMemsetAsync(xzero, 0, xbytes, c.stream)));
  }
  ScoreOut s;
  s.x = xzero; s.recon = recon;
  return image_decode_score(c, m, z, 0, B, h, w, s);
}

// ---- video model --------------------------------------------------------------------------------------------------------
// VideoEncoder.forward (models/video_autoencoder.py:217-231): frames fp32 [F,3,H,W] -> bf16 NHWC [F,H/16,W/16,latent]
int video_encode(Ctx& c, const vad_video_model& m, const float* x, int F, int H, int W, void** z, int* which_out) {
  static const char* names[3] = {"encoder.4", "encoder.8", "encoder.12"};
  int which = 0;
  int h = H / 2, w = W / 2;
  void* cur = c.A.ping(which, static_cast<size_t>(F) * h * w * 32 * 2);
  VAD_TRY(first_conv(c, m.enc0, x, F, H, W, true, cur));
  for (int i = 0; i < 3; ++i) {
    const vad_gemm_weights& wt = m.enc[i];
    void* nxt = c.A.ping(which ^ 1, static_cast<size_t>(F) * (h / 2) * (w / 2) * wt.n_total * 2);
    VAD_TRY(conv(c, names[i], wt, cur, F, h, w, nxt, kLeaky, true));
    cur = nxt;
    which ^= 1;
    h /= 2;
    w /= 2;
  }
  *z = cur;
  *which_out = which;
  return VAD_OK;
}

vad_conv_desc lstm_desc(const vad_gemm_weights& wt, const void* src, void* hseq, float* cst, int B, int T, int h, int w,
                        void* counters) {
  const int hid = wt.cout, cin = wt.ctap - hid;
  vad_conv_desc d;
  std::memset(&d, 0, sizeof(d));
  d.src0 = src; d.src1 = hseq; d.out = hseq;
  d.c0 = cin; d.c1 = hid; d.T0 = T; d.T1 = T;
  d.B = B; d.H = h; d.W = w; d.ntaps = 9;
  d.weight = wt.w; d.bias = wt.bias; d.w_ctap = wt.ctap;
  d.n_total = wt.n_total; d.cout = hid; d.epilogue = VAD_EPI_LSTM; d.slope = kIdent;
  d.out_frame_stride = static_cast<long long>(T) * h * w * hid;
  d.out_cpitch = hid;
  d.c_state = cst;
  d.scratch = counters;
  return d;
}

// ConvLSTM.forward (:127-172), zero initial state: seq bf16 [B,T,h,w,C] -> last layer's hidden sequence bf16
// [B,T,h,w,hid] (persist region) and its final cell state fp32 [B,h,w,hid].
int video_convlstm(Ctx& c, const vad_video_model& m, const void* seq, int B, int T, int h, int w, void** out_seq,
                   float** out_c) {
  const int L = m.lstm_layers;
  if (L <= 0 || L > VAD_MAX_LSTM_LAYERS) return VAD_ERR_ARG;
  void* hseq[VAD_MAX_LSTM_LAYERS];
  float* cst[VAD_MAX_LSTM_LAYERS];
  for (int l = 0; l < L; ++l) {
    const int hid = m.lstm[l].cout;
    hseq[l] = c.A.persist(static_cast<size_t>(B) * T * h * w * hid * 2);
    cst[l] = static_cast<float*>(c.A.persist(static_cast<size_t>(B) * h * w * hid * 4));
  }
  // Clips are independent, so a batch whose recurrent tiles exceed the SMs runs as groups of clips that each fit the
  // persistent kernels (one launch per group) instead of falling back to one launch per time step.
  int group = B;
  {
    const vad_gemm_weights& w0 = m.lstm[0];
    const int sms = sm_count();
    int fit = 0;
    for (int b = 1; b <= B; ++b) {
      if (lstm_persistent_tiles(b, h, w, w0.ctap - w0.cout, w0.cout, w0.n_total) > sms) break;
      fit = b;
    }
    if (fit > 0 && fit < B) {
      const int ngroups = (B + fit - 1) / fit;
      group = (B + ngroups - 1) / ngroups;
      if ((group & 1) && group + 1 <= fit) ++group;  // 8x8 latents pack two clips per tile
    }
  }
  const int ngroups = (B + group - 1) / group;
  unsigned int* counters = static_cast<unsigned int*>(c.A.persist(static_cast<size_t>(ngroups) * L * 16));
  if (c.dry()) {
    *out_seq = nullptr;
    *out_c = nullptr;
    return VAD_OK;
  }
  char name[48];
  for (int g = 0; g < ngroups; ++g) {
    const int b0 = g * group, bc = std::min(group, B - b0);
    const void* cur = static_cast<const char*>(seq) + static_cast<size_t>(b0) * T * h * w * (m.lstm[0].ctap - m.lstm[0].cout) * 2;
    int l = 0;
    while (l < L) {
      const int hid = m.lstm[l].cout;
      void* hs = static_cast<char*>(hseq[l]) + static_cast<size_t>(b0) * T * h * w * hid * 2;
      float* cs = cst[l] + static_cast<size_t>(b0) * h * w * hid;
      unsigned int* cnt = counters + (static_cast<size_t>(g) * L + l) * 4;
      vad_conv_desc d = lstm_desc(m.lstm[l], cur, hs, cs, bc, T, h, w, cnt);
      if (!(m.flags & VAD_FLAG_NO_LSTM_WAVEFRONT) && l + 1 < L) {
        // two layers as one wavefront launch (layer 2's step t runs next to layer 1's step t+1)
        const int hid2 = m.lstm[l + 1].cout;
        void* hs2 = static_cast<char*>(hseq[l + 1]) + static_cast<size_t>(b0) * T * h * w * hid2 * 2;
        float* cs2 = cst[l + 1] + static_cast<size_t>(b0) * h * w * hid2;
        vad_conv_desc d2 = lstm_desc(m.lstm[l + 1], hs, hs2, cs2, bc, T, h, w, nullptr);
        std::snprintf(name, sizeof(name), "convlstm.%d+%d", l, l + 1);
        int rc;
        {
          ProfScope p(name, c.stream);
          rc = vad_convlstm2_sequence(&d, &d2, T, c.stream);
          if (rc == VAD_ERR_UNSUPPORTED) p.cancel();
        }
        if (rc == VAD_OK) {
          cur = hs2;
          l += 2;
          continue;
        }
        if (rc != VAD_ERR_UNSUPPORTED) return rc;
      }
      std::snprintf(name, sizeof(name), "convlstm.%d", l);
      {
        ProfScope p(name, c.stream);
        VAD_TRY(vad_convlstm_sequence(&d, T, c.stream));
      }
      cur = hs;
      l += 1;
    }
  }
  *out_seq = hseq[L - 1];
  *out_c = cst[L - 1];
  return VAD_OK;
}

// proj (1x1 conv, only when lstm_hidden_dim != latent_dim; :311-312,346-349)
int video_project(Ctx& c, const vad_video_model& m, const void* seq, int F, int h, int w, const void** out, int* which) {
  if (!m.has_proj) {
    *out = seq;
    return VAD_OK;
  }
  const int dst = (*which == 0) ? 1 : 0;
  void* o = c.A.ping(dst, static_cast<size_t>(F) * h * w * m.proj.n_total * 2);
  VAD_TRY(conv(c, "proj", m.proj, seq, F, h, w, o, kIdent, false));
  *out = o;
  *which = dst;
  return VAD_OK;
}

// VideoDecoder.forward (:263-276) with the error reduction (:371-384) fused onto the last layer(s).
int video_decode_score(Ctx& c, const vad_video_model& m, const void* z, int which, int F, int h, int w, ScoreOut& s) {
  static const char* names[3] = {"decoder.0", "decoder.3", "decoder.6"};
  const int Ho = 16 * h, Wo = 16 * w;
  VAD_TRY(prepare_score(c, s, F, Ho, Wo));
  const vad_gemm_weights& w6 = m.dec[2];
  const vad_gemm_weights& w9 = m.dec[3];
  const bool fuse = !(m.flags & VAD_FLAG_NO_FUSED_TAIL) && w6.ctap == 64 && w6.n_total == 128 && w6.cout == 32 &&
                    w9.ctap == 32 && w9.n_total == 16;
  const void* cur = z;
  int nxt_which = which == 0 ? 1 : 0;
  const int n_plain = fuse ? 2 : 3;
  for (int i = 0; i < n_plain; ++i) {
    const vad_gemm_weights& wt = m.dec[i];
    void* up = c.A.ping(nxt_which, static_cast<size_t>(F) * 4 * h * w * wt.cout * 2);
    VAD_TRY(convt(c, names[i], wt, cur, F, h, w, up));
    cur = up;
    nxt_which ^= 1;
    h *= 2;
    w *= 2;
  }
  if (!fuse) return score_layer(c, "decoder.9+score", w9, cur, F, h, w, VAD_EPI_CONVT_TANH_SCORE, s, Ho, Wo);
  // decoder.6 + decoder.9 + score in one kernel: the 32-channel half-resolution tensor never reaches HBM
  if (c.dry()) return VAD_OK;
  vad_conv_desc d = base_desc(w6, cur, F, h, w, VAD_EPI_CONVT, kRelu);
  d.x = s.x; d.recon = s.recon; d.heat = s.heat; d.partials = s.partials;
  const int tiles = vad_convt2_score_tiles(&d);
  if (tiles <= 0) return tiles < 0 ? tiles : VAD_ERR_ARG;
  if (static_cast<size_t>(tiles) * 64 > s.partials_bytes) return VAD_ERR_WORKSPACE;
  {
    ProfScope p("decoder.6+9+score", c.stream);
    VAD_TRY(vad_convt2_score(&d, w9.w, w9.bias, c.stream));
  }
  return finalize(c, s, F, tiles, Ho, Wo);
}

int video_score_latents_impl(Ctx& c, const vad_video_model& m, const void* z, int z_which, const float* x, int B, int T,
                             int h, int w, float* recon, float* score, float* minmax, float* heat) {
  void* seq = nullptr;
  float* cl = nullptr;
  VAD_TRY(video_convlstm(c, m, z, B, T, h, w, &seq, &cl));
  const void* zp = seq;
  int which = -1;
  (void)z_which;  // the LSTM output lives in the persist region, so both ping regions are free again
  VAD_TRY(video_project(c, m, seq, B * T, h, w, &zp, &which));
  ScoreOut s;
  s.x = x; s.recon = recon; s.score = score; s.minmax = minmax; s.heat = heat;
  return video_decode_score(c, m, zp, which, B * T, h, w, s);
}

int video_forward_impl(Ctx& c, const vad_video_model& m, const float* x, int B, int T, int H, int W, float* recon,
                       float* score, float* minmax, float* heat) {
  void* z = nullptr;
  int which = 0;
  VAD_TRY(video_encode(c, m, x, B * T, H, W, &z, &which));
  return video_score_latents_impl(c, m, z, which, x, B, T, H / 16, W / 16, recon, score, minmax, heat);
}

int video_decode_impl(Ctx& c, const vad_video_model& m, const float* zf, int F, int h, int w, float* recon) {
  const int latent = m.dec[0].ctap;
  void* z = c.A.ping(0, static_cast<size_t>(F) * h * w * latent * 2);
  const size_t xbytes = static_cast<size_t>(F) * 3 * (16 * h) * (16 * w) * 4;
  float* xzero = static_cast<float*>(c.A.persist(xbytes));
  if (!c.dry()) {
    VAD_TRY(vad_nchw_f32_to_nhwc_bf16(zf, F, latent, h, w, z, c.stream));
    VAD_TRY(static_cast<int>(cudaMemsetAsync(xzero, 0, xbytes, c.stream)));
  }
  ScoreOut s;
  s.x = xzero; s.recon = recon;
  return video_decode_score(c, m, z, 0, F, h, w, s);
}

// ---- uint8 HWC frames in, (optionally) normalised uint8 heat maps out: the decoder-facing form of the two forwards ----
// (ToTensor + Normalize on the device: utils/dataset.py:65-70, utils/video_dataset.py:62-66; create_heatmap's
// normalisation: evaluate_video.py:56-57.)  The fp32 frames / heat maps the kernels work on live in the workspace.
struct U8Io {
  float* x = nullptr;     // normalised frames fp32 [F,3,H,W] (workspace)
  float* heat = nullptr;  // fp32 heat map actually written (caller's, or workspace when only heat_u8 is wanted)
  float* minmax = nullptr;
};

int u8_prepare(Ctx& c, const uint8_t* frames, int F, int H, int W, float* heat, float* minmax, uint8_t* heat_u8, U8Io& io) {
  io.x = static_cast<float*>(c.A.persist(static_cast<size_t>(F) * 3 * H * W * 4));
  io.heat = heat;
  io.minmax = minmax;
  if (heat_u8) {
    if (!io.heat) io.heat = static_cast<float*>(c.A.persist(static_cast<size_t>(F) * H * W * 4));
    if (!io.minmax) io.minmax = static_cast<float*>(c.A.persist(static_cast<size_t>(F) * 2 * 4));
  }
  if (c.dry()) return VAD_OK;
  ProfScope p("u8_to_f32", c.stream);
  return vad_u8_hwc_to_f32_nchw(frames, F, H, W, io.x, c.stream);
}

int u8_finish(Ctx& c, const U8Io& io, int F, int H, int W, uint8_t* heat_u8) {
  if (!heat_u8 || c.dry()) return VAD_OK;
  ProfScope p("heat_u8", c.stream);
  return vad_heatmap_u8(io.heat, io.minmax, F, H, W, heat_u8, c.stream);
}

// fp32 NCHW <-> NHWC (ConvLSTM cell state at the sub-module boundary)
__global__ void f32_nchw_to_nhwc_kernel(const float* __restrict__ src, long long total, int HW, int C,
                                        float* __restrict__ dst) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;  // NHWC order
  if (i >= total) return;
  const int ch = static_cast<int>(i % C);
  const long long np = i / C;
  const long long n = np / HW;
  const int p = static_cast<int>(np - n * HW);
  dst[i] = src[(n * C + ch) * HW + p];
}
__global__ void f32_nhwc_to_nchw_kernel(const float* __restrict__ src, long long total, int HW, int C,
                                        float* __restrict__ dst) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;  // NCHW order
  if (i >= total) return;
  const long long n = i / (static_cast<long long>(C) * HW);
  const long long rem = i - n * C * HW;
  const int ch = static_cast<int>(rem / HW);
  const int p = static_cast<int>(rem - static_cast<long long>(ch) * HW);
  dst[i] = src[(n * HW + p) * C + ch];
}
int f32_nchw_to_nhwc(const float* src, int N, int C, int HW, float* dst, cudaStream_t stream) {
  const long long total = static_cast<long long>(N) * C * HW;
  f32_nchw_to_nhwc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(src, total, HW, C, dst);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}
int f32_nhwc_to_nchw(const float* src, int N, int C, int HW, float* dst, cudaStream_t stream) {
  const long long total = static_cast<long long>(N) * C * HW;
  f32_nhwc_to_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(src, total, HW, C, dst);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int convlstm_forward_impl(Ctx& c, const vad_video_model& m, const float* x, int B, int T, int h, int w, float* out,
                          float* h_last, float* c_last) {
  const int cin = m.lstm[0].ctap - m.lstm[0].cout;
  const int hid = m.lstm[m.lstm_layers - 1].cout;
  void* seq = c.A.ping(0, static_cast<size_t>(B) * T * h * w * cin * 2);
  if (!c.dry()) VAD_TRY(vad_nchw_f32_to_nhwc_bf16(x, B * T, cin, h, w, seq, c.stream));
  void* hs = nullptr;
  float* cl = nullptr;
  VAD_TRY(video_convlstm(c, m, seq, B, T, h, w, &hs, &cl));
  if (c.dry()) return VAD_OK;
  VAD_TRY(vad_nhwc_bf16_to_nchw_f32(hs, B * T, h, w, hid, out, c.stream));
  const size_t frame = static_cast<size_t>(hid) * h * w * 4;
  if (h_last)
    VAD_TRY(static_cast<int>(cudaMemcpy2DAsync(h_last, frame, out + static_cast<size_t>(T - 1) * hid * h * w, frame * T,
                                               frame, B, cudaMemcpyDeviceToDevice, c.stream)));
  if (c_last) VAD_TRY(f32_nhwc_to_nchw(cl, B, hid, h * w, c_last, c.stream));
  return VAD_OK;
}

int convlstm_cell_impl(Ctx& c, const vad_gemm_weights& wt, const float* x, const float* h_cur, const float* c_cur, int B,
                       int h, int w, float* h_next, float* c_next) {
  const int hid = wt.cout, cin = wt.ctap - hid;
  const size_t px = static_cast<size_t>(B) * h * w;
  void* xb = c.A.persist(px * cin * 2);
  void* hb = c.A.persist(px * hid * 2);
  void* hn = c.A.persist(px * hid * 2);
  float* cs = static_cast<float*>(c.A.persist(px * hid * 4));
  if (c.dry()) return VAD_OK;
  VAD_TRY(vad_nchw_f32_to_nhwc_bf16(x, B, cin, h, w, xb, c.stream));
  VAD_TRY(vad_nchw_f32_to_nhwc_bf16(h_cur, B, hid, h, w, hb, c.stream));
  VAD_TRY(f32_nchw_to_nhwc(c_cur, B, hid, h * w, cs, c.stream));
  vad_conv_desc d;
  std::memset(&d, 0, sizeof(d));
  d.src0 = xb; d.src1 = hb; d.out = hn;
  d.c0 = cin; d.c1 = hid; d.T0 = d.T1 = 1;
  d.B = B; d.H = h; d.W = w; d.ntaps = 9;
  d.weight = wt.w; d.bias = wt.bias; d.w_ctap = wt.ctap;
  d.n_total = wt.n_total; d.cout = hid; d.epilogue = VAD_EPI_LSTM; d.slope = kIdent;
  d.out_frame_stride = static_cast<long long>(h) * w * hid;
  d.out_cpitch = hid;
  d.c_state = cs;
  d.lstm_first = 0;
  VAD_TRY(vad_conv_layer(&d, c.stream));
  VAD_TRY(vad_nhwc_bf16_to_nchw_f32(hn, B, h, w, hid, h_next, c.stream));
  return f32_nhwc_to_nchw(cs, B, hid, h * w, c_next, c.stream);
}

// ---- dry run + real run of one schedule ---------------------------------------------------------------------------------
template <typename Fn>
size_t measure(Fn&& fn) {
  Ctx c;
  c.A.dry = true;
  if (fn(c) != VAD_OK) return 0;
  const size_t total = c.A.plan.total();
  return total ? total : kAlign;
}

template <typename Fn>
int run(Fn&& fn, void* ws, size_t ws_bytes, vad_stream_t stream) {
  Ctx dry;
  dry.A.dry = true;
  VAD_TRY(fn(dry));
  if (dry.A.plan.total() > 0 && (!ws || reinterpret_cast<uintptr_t>(ws) % kAlign != 0)) return VAD_ERR_ARG;
  if (dry.A.plan.total() > ws_bytes) return VAD_ERR_WORKSPACE;
  Ctx c;
  c.A.dry = false;
  c.A.base = static_cast<char*>(ws);
  c.A.plan = dry.A.plan;
  c.stream = static_cast<cudaStream_t>(stream);
  return fn(c);
}

bool image_model_ok(const vad_image_model* m, bool need_enc, bool need_dec) {
  if (!m) return false;
  if (need_enc) {
    if (!m->has_encoder || !first_ok(m->enc1_0)) return false;
    for (const auto& w : m->enc)
      if (!weights_ok(w)) return false;
  }
  if (need_dec) {
    if (!m->has_decoder) return false;
    for (const auto& w : m->dec)
      if (!weights_ok(w)) return false;
  }
  return true;
}

bool video_model_ok(const vad_video_model* m, bool need_enc, bool need_lstm, bool need_dec) {
  if (!m) return false;
  if (need_enc) {
    if (!m->has_encoder || !first_ok(m->enc0)) return false;
    for (const auto& w : m->enc)
      if (!weights_ok(w)) return false;
  }
  if (need_lstm) {
    if (m->lstm_layers <= 0 || m->lstm_layers > VAD_MAX_LSTM_LAYERS) return false;
    for (int l = 0; l < m->lstm_layers; ++l)
      if (!weights_ok(m->lstm[l]) || m->lstm[l].ctap <= m->lstm[l].cout) return false;
    if (m->has_proj && !weights_ok(m->proj)) return false;
  }
  if (need_dec) {
    if (!m->has_decoder) return false;
    for (const auto& w : m->dec)
      if (!weights_ok(w)) return false;
  }
  return true;
}

}  // namespace

extern "C" {

size_t vad_image_workspace_bytes(const vad_image_model* m, int op, int B, int H, int W) {
  if (B <= 0) return 0;
  switch (op) {
    case VAD_OP_FORWARD:
      if (!image_model_ok(m, true, true) || !hw_ok(H, W)) return 0;
      return measure([&](Ctx& c) {
        float dummy;  // "every output requested" is the largest plan
        return image_forward_impl(c, *m, nullptr, B, H, W, &dummy, &dummy, nullptr, &dummy, &dummy);
      });
    case VAD_OP_ENCODE:
      if (!image_model_ok(m, true, false) || !hw_ok(H, W)) return 0;
      return measure([&](Ctx& c) {
        float dummy;
        return image_forward_impl(c, *m, nullptr, B, H, W, nullptr, &dummy, nullptr, nullptr, nullptr);
      });
    case VAD_OP_DECODE:  // H, W are the LATENT extent here
      if (!image_model_ok(m, false, true) || H <= 0 || W <= 0) return 0;
      return measure([&](Ctx& c) { return image_decode_impl(c, *m, nullptr, B, H, W, nullptr); });
    case VAD_OP_FORWARD_U8:
      if (!image_model_ok(m, true, true) || !hw_ok(H, W)) return 0;
      return measure([&](Ctx& c) {
        float dummy;
        uint8_t dummy8;
        U8Io io;
        VAD_TRY(u8_prepare(c, nullptr, B, H, W, nullptr, nullptr, &dummy8, io));  // heat map + min/max in the workspace
        return image_forward_impl(c, *m, nullptr, B, H, W, &dummy, &dummy, nullptr, &dummy, &dummy);
      });
    default:
      return 0;
  }
}

int vad_image_forward(const vad_image_model* m, const float* x, int B, int H, int W, float* recon, float* latent,
                      float* score, float* minmax, float* heat, void* ws, size_t ws_bytes, vad_stream_t stream) {
  const bool need_dec = recon || score || minmax || heat;
  if (!x || B <= 0 || (!need_dec && !latent)) return VAD_ERR_ARG;
  if (!image_model_ok(m, true, need_dec)) return VAD_ERR_ARG;
  if (!hw_ok(H, W)) return VAD_ERR_SHAPE;
  return run([&](Ctx& c) { return image_forward_impl(c, *m, x, B, H, W, recon, latent, score, minmax, heat); }, ws,
             ws_bytes, stream);
}

int vad_image_decode(const vad_image_model* m, const float* z, int B, int h, int w, float* recon, void* ws,
                     size_t ws_bytes, vad_stream_t stream) {
  if (!z || !recon || B <= 0 || h <= 0 || w <= 0) return VAD_ERR_ARG;
  if (!image_model_ok(m, false, true)) return VAD_ERR_ARG;
  return run([&](Ctx& c) { return image_decode_impl(c, *m, z, B, h, w, recon); }, ws, ws_bytes, stream);
}

size_t vad_video_workspace_bytes(const vad_video_model* m, int op, int B, int T, int H, int W) {
  if (B <= 0 || T <= 0) return 0;
  float dummy;
  switch (op) {
    case VAD_OP_FORWARD:
      if (!video_model_ok(m, true, true, true) || !hw_ok(H, W)) return 0;
      return measure([&](Ctx& c) { return video_forward_impl(c, *m, nullptr, B, T, H, W, &dummy, nullptr, &dummy, &dummy); });
    case VAD_OP_ENCODE: {  // B*T frames
      if (!video_model_ok(m, true, false, false) || !hw_ok(H, W)) return 0;
      return measure([&](Ctx& c) {
        void* z;
        int which;
        return video_encode(c, *m, nullptr, B * T, H, W, &z, &which);
      });
    }
    case VAD_OP_DECODE:  // B*T frames, H x W = latent extent
      if (!video_model_ok(m, false, false, true) || H <= 0 || W <= 0) return 0;
      return measure([&](Ctx& c) { return video_decode_impl(c, *m, nullptr, B * T, H, W, nullptr); });
    case VAD_OP_CONVLSTM:  // H x W = latent extent
      if (!video_model_ok(m, false, true, false) || H <= 0 || W <= 0) return 0;
      return measure([&](Ctx& c) { return convlstm_forward_impl(c, *m, nullptr, B, T, H, W, nullptr, nullptr, nullptr); });
    case VAD_OP_FORWARD_U8:
      if (!video_model_ok(m, true, true, true) || !hw_ok(H, W)) return 0;
      return measure([&](Ctx& c) {
        uint8_t dummy8;
        U8Io io;
        VAD_TRY(u8_prepare(c, nullptr, B * T, H, W, nullptr, nullptr, &dummy8, io));
        return video_forward_impl(c, *m, nullptr, B, T, H, W, &dummy, nullptr, &dummy, &dummy);
      });
    case VAD_OP_SCORE_LATENTS:  // H x W = latent extent
      if (!video_model_ok(m, false, true, true) || H <= 0 || W <= 0) return 0;
      return measure([&](Ctx& c) {
        return video_score_latents_impl(c, *m, nullptr, -1, nullptr, B, T, H, W, &dummy, nullptr, &dummy, &dummy);
      });
    default:
      return 0;
  }
}

int vad_video_forward(const vad_video_model* m, const float* x, int B, int T, int H, int W, float* recon, float* score,
                      float* minmax, float* heat, void* ws, size_t ws_bytes, vad_stream_t stream) {
  if (!x || B <= 0 || T <= 0 || (!recon && !score && !minmax && !heat)) return VAD_ERR_ARG;
  if (!video_model_ok(m, true, true, true)) return VAD_ERR_ARG;
  if (!hw_ok(H, W)) return VAD_ERR_SHAPE;
  return run([&](Ctx& c) { return video_forward_impl(c, *m, x, B, T, H, W, recon, score, minmax, heat); }, ws, ws_bytes,
             stream);
}

int vad_image_forward_u8(const vad_image_model* m, const uint8_t* frames, int B, int H, int W, float* recon, float* latent,
                         float* score, float* minmax, float* heat, uint8_t* heat_u8, void* ws, size_t ws_bytes,
                         vad_stream_t stream) {
  if (!frames || B <= 0 || (!recon && !latent && !score && !minmax && !heat && !heat_u8)) return VAD_ERR_ARG;
  if (!image_model_ok(m, true, recon || score || minmax || heat || heat_u8)) return VAD_ERR_ARG;
  if (!hw_ok(H, W)) return VAD_ERR_SHAPE;
  return run(
      [&](Ctx& c) {
        U8Io io;
        VAD_TRY(u8_prepare(c, frames, B, H, W, heat, minmax, heat_u8, io));
        VAD_TRY(image_forward_impl(c, *m, io.x, B, H, W, recon, latent, score, io.minmax, io.heat));
        return u8_finish(c, io, B, H, W, heat_u8);
      },
      ws, ws_bytes, stream);
}

int vad_video_forward_u8(const vad_video_model* m, const uint8_t* frames, int B, int T, int H, int W, float* recon,
                         float* score, float* minmax, float* heat, uint8_t* heat_u8, void* ws, size_t ws_bytes,
                         vad_stream_t stream) {
  if (!frames || B <= 0 || T <= 0 || (!recon && !score && !minmax && !heat && !heat_u8)) return VAD_ERR_ARG;
  if (!video_model_ok(m, true, true, true)) return VAD_ERR_ARG;
  if (!hw_ok(H, W)) return VAD_ERR_SHAPE;
  return run(
      [&](Ctx& c) {
        U8Io io;
        VAD_TRY(u8_prepare(c, frames, B * T, H, W, heat, minmax, heat_u8, io));
        VAD_TRY(video_forward_impl(c, *m, io.x, B, T, H, W, recon, score, io.minmax, io.heat));
        return u8_finish(c, io, B * T, H, W, heat_u8);
      },
      ws, ws_bytes, stream);
}

int vad_video_encode(const vad_video_model* m, const float* x, int F, int H, int W, float* latent, void* latent_bf16,
                     void* ws, size_t ws_bytes, vad_stream_t stream) {
  if (!x || F <= 0 || (!latent && !latent_bf16)) return VAD_ERR_ARG;
  if (!video_model_ok(m, true, false, false)) return VAD_ERR_ARG;
  if (!hw_ok(H, W)) return VAD_ERR_SHAPE;
  return run(
      [&](Ctx& c) {
        void* z = nullptr;
        int which = 0;
        VAD_TRY(video_encode(c, *m, x, F, H, W, &z, &which));
        if (c.dry()) return static_cast<int>(VAD_OK);
        const int h = H / 16, w = W / 16, C = m->enc[2].n_total;
        if (latent) VAD_TRY(vad_nhwc_bf16_to_nchw_f32(z, F, h, w, C, latent, c.stream));
        if (latent_bf16)
          VAD_TRY(static_cast<int>(cudaMemcpyAsync(latent_bf16, z, static_cast<size_t>(F) * h * w * C * 2,
                                                   cudaMemcpyDeviceToDevice, c.stream)));
        return static_cast<int>(VAD_OK);
      },
      ws, ws_bytes, stream);
}

int vad_video_score_latents(const vad_video_model* m, const void* z_bf16, const float* x, int B, int T, int h, int w,
                            float* recon, float* score, float* minmax, float* heat, void* ws, size_t ws_bytes,
                            vad_stream_t stream) {
  if (!z_bf16 || !x || B <= 0 || T <= 0 || h <= 0 || w <= 0 || (!recon && !score && !minmax && !heat)) return VAD_ERR_ARG;
  if (!video_model_ok(m, false, true, true)) return VAD_ERR_ARG;
  return run(
      [&](Ctx& c) { return video_score_latents_impl(c, *m, z_bf16, -1, x, B, T, h, w, recon, score, minmax, heat); }, ws,
      ws_bytes, stream);
}

int vad_video_decode(const vad_video_model* m, const float* z, int F, int h, int w, float* recon, void* ws,
                     size_t ws_bytes, vad_stream_t stream) {
  if (!z || !recon || F <= 0 || h <= 0 || w <= 0) return VAD_ERR_ARG;
  if (!video_model_ok(m, false, false, true)) return VAD_ERR_ARG;
  return run([&](Ctx& c) { return video_decode_impl(c, *m, z, F, h, w, recon); }, ws, ws_bytes, stream);
}

int vad_convlstm_forward(const vad_video_model* m, const float* x, int B, int T, int h, int w, float* out,
                         float* h_last, float* c_last, void* ws, size_t ws_bytes, vad_stream_t stream) {
  if (!x || !out || B <= 0 || T <= 0 || h <= 0 || w <= 0) return VAD_ERR_ARG;
  if (!video_model_ok(m, false, true, false)) return VAD_ERR_ARG;
  return run([&](Ctx& c) { return convlstm_forward_impl(c, *m, x, B, T, h, w, out, h_last, c_last); }, ws, ws_bytes,
             stream);
}

size_t vad_convlstm_cell_workspace_bytes(const vad_gemm_weights* cell, int B, int h, int w) {
  if (!cell || !weights_ok(*cell) || cell->ctap <= cell->cout || B <= 0 || h <= 0 || w <= 0) return 0;
  return measure([&](Ctx& c) { return convlstm_cell_impl(c, *cell, nullptr, nullptr, nullptr, B, h, w, nullptr, nullptr); });
}

int vad_convlstm_cell(const vad_gemm_weights* cell, const float* x, const float* h_cur, const float* c_cur, int B,
                      int h, int w, float* h_next, float* c_next, void* ws, size_t ws_bytes, vad_stream_t stream) {
  if (!cell || !weights_ok(*cell) || cell->ctap <= cell->cout) return VAD_ERR_ARG;
  if (!x || !h_cur || !c_cur || !h_next || !c_next || B <= 0 || h <= 0 || w <= 0) return VAD_ERR_ARG;
  return run([&](Ctx& c) { return convlstm_cell_impl(c, *cell, x, h_cur, c_cur, B, h, w, h_next, c_next); }, ws,
             ws_bytes, stream);
}

int vad_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  const int prev = g_prof_on ? 1 : 0;
  g_prof_on = on != 0;
  return prev;
}

int vad_profile_dump(char* buf, size_t buf_bytes) {
  if (!buf || buf_bytes == 0) return VAD_ERR_ARG;
  std::vector<ProfEntry> log;
  {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    log.swap(g_prof);
  }
  size_t off = 0;
  buf[0] = 0;
  int rc = VAD_OK;
  for (auto& e : log) {
    float ms = 0.f;
    cudaError_t err = cudaEventSynchronize(e.e1);
    if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e.e0, e.e1);
    if (err != cudaSuccess) rc = static_cast<int>(err);
    cudaEventDestroy(e.e0);
    cudaEventDestroy(e.e1);
    if (rc == VAD_OK) {
      const int n = std::snprintf(buf + off, buf_bytes - off, "%s\t%.6f\n", e.name.c_str(), ms);
      if (n < 0 || static_cast<size_t>(n) >= buf_bytes - off) rc = VAD_ERR_WORKSPACE;
      else off += static_cast<size_t>(n);
    }
  }
  return rc == VAD_OK ? static_cast<int>(off) : rc;
}

}  // extern "C"
