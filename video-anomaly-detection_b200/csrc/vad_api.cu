// C-ABI entry points of libvad_b200.so (see include/vad_b200.h) and the small CUDA-core kernels around the
// tcgen05 GEMM: first 3-channel convolution, standalone scoring pass, score finalisation, layout helpers.
#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "vad_internal.h"
#include "vad_ptx.cuh"

namespace vad {

// ------------------------------------------------------------------------------------------------ bookkeeping
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// Host-mapped slot a timed-out mbarrier wait writes before trapping (readable even after the context is poisoned).
// All per-device state (the trap slot's device symbol, the SM count, constant-memory tables) is cached per device index,
// so a process that drives several GPUs gets each of them set up (the caller makes the tensors' device current).
constexpr int kMaxDevices = 64;
static int current_device_index() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}
static std::mutex g_setup_mutex;
static unsigned long long* g_trap_host = nullptr;
static std::atomic<bool> g_trap_ready[kMaxDevices];
static void ensure_trap_slot() {
  const int dev = current_device_index();
  if (g_trap_ready[dev].load(std::memory_order_acquire)) return;
  std::lock_guard<std::mutex> lock(g_setup_mutex);
  if (g_trap_ready[dev].load(std::memory_order_relaxed)) return;
  if (!g_trap_host) {  // one host-mapped record shared by all devices (portable pinned memory)
    unsigned long long* h = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&h), 4 * sizeof(unsigned long long),
                      cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
      std::memset(h, 0, 4 * sizeof(unsigned long long));
      g_trap_host = h;
    }
  }
  if (g_trap_host) {
    unsigned long long* d = nullptr;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), g_trap_host, 0) == cudaSuccess) set_trap_slot(d);
  }
  g_trap_ready[dev].store(true, std::memory_order_release);
}

int sm_count() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device_index();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int floor_log2(int v) {
  int l = 0;
  while ((2 << l) <= v) ++l;
  return l;
}
static int ceil_log2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// 128 accumulator rows = TN frames x TH rows x TW columns (powers of two; partial tiles are zero-filled by TMA on
// load and masked on store).  Score layers force TN = 1 so that every tile lies inside one frame.
TileGeom pick_tile_geometry(int B, int H, int W, bool single_frame_tiles) {
  TileGeom g;
  g.lgTW = floor_log2(W) < 4 ? floor_log2(W) : 4;
  const int rem = 7 - g.lgTW;
  int lgTH = floor_log2(H);
  const int cap = (g.lgTW == 4) ? 3 : rem;
  g.lgTH = lgTH < cap ? lgTH : cap;
  int lgTN = rem - g.lgTH;
  if (single_frame_tiles) lgTN = 0;
  const int need = ceil_log2(B);
  g.lgTN = lgTN < need ? lgTN : need;
  g.tiles_w = (W + (1 << g.lgTW) - 1) >> g.lgTW;
  g.tiles_h = (H + (1 << g.lgTH) - 1) >> g.lgTH;
  g.tiles_b = (B + (1 << g.lgTN) - 1) >> g.lgTN;
  return g;
}

static int encode_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int T, int B, int CK,
                          const TileGeom& g) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return VAD_ERR_DRIVER;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)T * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)CK, 1u << g.lgTW, 1u << g.lgTH, 1u, 1u << g.lgTN};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VAD_OK : VAD_ERR_DRIVER;
}

// generic bf16 5-D map whose innermost box extent is one swizzle span (64 ch -> 128B swizzle, 32 ch -> 64B)
static int encode_map5(CUtensorMap* m, const void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                       const cuuint32_t* box, int inner_channels) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return VAD_ERR_DRIVER;
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  inner_channels == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VAD_OK : VAD_ERR_DRIVER;
}

static int encode_act_map_box(CUtensorMap* m, const void* base, int C, int W, int H, int T, int B, int CK, int box_w,
                              int box_h, int box_n) {
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                           (cuuint64_t)T * H * W * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)CK, (cuuint32_t)box_w, (cuuint32_t)box_h, 1u, (cuuint32_t)box_n};
  return encode_map5(m, base, dims, strides, box, CK);
}

// Tuning / bring-up switches (environment, read once):
//   VAD_HALO = 0 off | 1 one (TW+2)-wide patch, descriptor base offset 0 | 2 same, base offset (addr>>7)&7 |
//              3 three TW-wide patches shifted by dx (only whole-row descriptor shifts)
//   VAD_TMA_STORE = 0 direct epilogue stores | 1 staged TMA stores
static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}
static int halo_mode_setting() {
  static int v = env_int("VAD_HALO", 1);
  return v;
}
// VAD_DUAL_MMA bit mask: 1 halo kernel, 2 first conv, 4 streaming kernel (short k-loops only)
static int dual_mma_setting() {
  static int v = env_int("VAD_DUAL_MMA", 1);
  return v;
}
// VAD_KX: 0 = never use the kx-merged kernel | 1 (default) = the 3-channel score layer only (for wider outputs the
// 3x TMEM read of the merged accumulator costs more than the saved MMAs: TMEM -> registers runs at 64 B/clk/SM) |
// 2 = also Cout = 32 | 3 = also Cout = 64
static int g_kx_override = -1;  // vad_debug_set_kx
static int kx_setting() {
  static int v = env_int("VAD_KX", 1);
  return g_kx_override >= 0 ? g_kx_override : v;
}
// VAD_HS: 0 = off | 1 (default) = wide 3x3 layers (N multiple of 128; Cin >= 128, or Cin = 64 without pooling — measured:
// enc3.0 0.147 -> 0.135 ms, but the pooled 64 -> 128 video layer is faster on the resident-weights kernel) use the
// patch + streamed-weights kernel | 2 = every Cin = 64 layer too
static int hs_setting() {
  static int v = env_int("VAD_HS", 1);
  return v;
}
// VAD_TOKEN bit mask: the two MMA issuers pass a token (strict tile order) — 1 halo kernel, 2 first conv, 8 kx (score)
// kernel.  Default: on for tiles with a long MMA sequence (64-channel chunks: 36 MMAs, where an ordered, uninterrupted
// stream matters), off for the short ones (free-running issuers overlap their per-tile latencies better).  The token is
// forced on when the patch ring has an odd number of slots: a slot is then waited for by both issuers in turn, and
// only the token's ordering keeps either from being two barrier phases ahead.
static int token_setting() {
  static int v = env_int("VAD_TOKEN", 0);
  return v;
}
// VAD_POOL_DIRECT: 1 (default) pooled outputs are stored straight from registers | 0 staged + TMA store
static int pool_direct_setting() {
  static int v = env_int("VAD_POOL_DIRECT", 1);
  return v;
}
// VAD_PDL_ALL: 1 (default) every conv kernel is launched with programmatic stream serialisation (its prologue overlaps
// the previous kernel's tail) | 0 plain stream order
static int pdl_all_setting() {
  static int v = env_int("VAD_PDL_ALL", 1);
  return v;
}
static int tma_store_setting() {
  static int v = env_int("VAD_TMA_STORE", 1);
  return v;
}

static int encode_weight_map(CUtensorMap* m, const void* base, int K, int N, int CK, int BN) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return VAD_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)CK, (cuuint32_t)BN};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VAD_OK : VAD_ERR_DRIVER;
}

// ------------------------------------------------------------------------------------------------ first conv
// fp32 NCHW 3-channel input -> bf16 NHWC, conv3x3 pad 1 + folded BN + LeakyReLU (+ 2x2 max-pool).
// K = 27 is too thin for the tensor pipe to matter and the layer is HBM-bound; CUDA cores.
template <int COUT, bool POOL>
__global__ void __launch_bounds__(256) first_conv_kernel(const float* __restrict__ x, const float* __restrict__ wgt,
                                                         const float* __restrict__ bias, float slope, int B, int H,
                                                         int W, __nv_bfloat16* __restrict__ out, int tiles_x,
                                                         int tiles_y) {
  // One thread = one conv output pixel of a 32x8 tile.  POOL: the four pixels of a 2x2 window sit in four
  // adjacent lanes (max by two shuffles); each of them then stores a quarter of the pooled pixel's channels.
  constexpr int TX = 32, TY = 8;
  constexpr int IW = TX + 2, IH = TY + 2;
  static_assert(COUT == 32, "pooled store splits 32 channels over 4 lanes");
  __shared__ float s_in[3][IH][IW + 1];
  __shared__ float4 s_w[27][COUT / 4];
  __shared__ float s_b[COUT];

  const int tile = blockIdx.x;
  const int tx_i = tile % tiles_x;
  const int ty_i = (tile / tiles_x) % tiles_y;
  const int f = tile / (tiles_x * tiles_y);
  const int x0 = tx_i * TX, y0 = ty_i * TY;  // conv-resolution tile origin
  const float* xf = x + static_cast<long long>(f) * 3 * H * W;

  for (int i = threadIdx.x; i < 27 * (COUT / 4); i += 256)
    s_w[i / (COUT / 4)][i % (COUT / 4)] = reinterpret_cast<const float4*>(wgt)[i];
  if (threadIdx.x < COUT) s_b[threadIdx.x] = bias[threadIdx.x];
  for (int i = threadIdx.x; i < 3 * IH * IW; i += 256) {
    const int c = i / (IH * IW);
    const int rem = i - c * (IH * IW);
    const int iy = rem / IW, ix = rem - iy * IW;
    const int gy = y0 - 1 + iy, gx = x0 - 1 + ix;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg(xf + (static_cast<long long>(c) * H + gy) * W + gx);
    s_in[c][iy][ix] = v;
  }
  __syncthreads();

  int px, py, sub = 0;
  if constexpr (POOL) {
    const int quad = threadIdx.x >> 2;
    sub = threadIdx.x & 3;
    px = 2 * (quad & 15) + (sub & 1);
    py = 2 * (quad >> 4) + (sub >> 1);
  } else {
    px = threadIdx.x & 31;
    py = threadIdx.x >> 5;
  }
  float acc[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) acc[j] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float v = s_in[ci][py + ky][px + kx];
        const int k = (ky * 3 + kx) * 3 + ci;
#pragma unroll
        for (int j4 = 0; j4 < COUT / 4; ++j4) {
          const float4 w4 = s_w[k][j4];
          acc[4 * j4 + 0] = fmaf(v, w4.x, acc[4 * j4 + 0]);
          acc[4 * j4 + 1] = fmaf(v, w4.y, acc[4 * j4 + 1]);
          acc[4 * j4 + 2] = fmaf(v, w4.z, acc[4 * j4 + 2]);
          acc[4 * j4 + 3] = fmaf(v, w4.w, acc[4 * j4 + 3]);
        }
      }
  if constexpr (POOL) {
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
      acc[j] = fmaxf(acc[j], __shfl_xor_sync(0xffffffffu, acc[j], 1));
      acc[j] = fmaxf(acc[j], __shfl_xor_sync(0xffffffffu, acc[j], 2));
    }
  }
#pragma unroll
  for (int j = 0; j < COUT; ++j) {
    const float v = acc[j] + s_b[j];
    acc[j] = v > 0.f ? v : v * slope;
  }
  const int gy = y0 + py, gx = x0 + px;
  if (gy < H && gx < W) {
    if constexpr (POOL) {
      const int Ho = H >> 1, Wo = W >> 1;
      uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<long long>(f) * Ho + (gy >> 1)) * Wo + (gx >> 1)) * COUT +
                                            sub * 8);
      uint4 val;
      if (sub == 0) val = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
      else if (sub == 1) val = make_uint4(pack_bf16x2(acc[8], acc[9]), pack_bf16x2(acc[10], acc[11]), pack_bf16x2(acc[12], acc[13]), pack_bf16x2(acc[14], acc[15]));
      else if (sub == 2) val = make_uint4(pack_bf16x2(acc[16], acc[17]), pack_bf16x2(acc[18], acc[19]), pack_bf16x2(acc[20], acc[21]), pack_bf16x2(acc[22], acc[23]));
      else val = make_uint4(pack_bf16x2(acc[24], acc[25]), pack_bf16x2(acc[26], acc[27]), pack_bf16x2(acc[28], acc[29]), pack_bf16x2(acc[30], acc[31]));
      *dst = val;
    } else {
      uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<long long>(f) * H + gy) * W + gx) * COUT);
#pragma unroll
      for (int j8 = 0; j8 < COUT / 8; ++j8)
        dst[j8] = make_uint4(pack_bf16x2(acc[j8 * 8 + 0], acc[j8 * 8 + 1]), pack_bf16x2(acc[j8 * 8 + 2], acc[j8 * 8 + 3]),
                             pack_bf16x2(acc[j8 * 8 + 4], acc[j8 * 8 + 5]), pack_bf16x2(acc[j8 * 8 + 6], acc[j8 * 8 + 7]));
    }
  }
}

// ------------------------------------------------------------------------------------------------ scoring kernels
// Standalone pass: one block reduces a contiguous pixel range of one frame; float4 streaming loads of the three
// channel planes of x and recon; writes optional heat map; per-block (sum, min, max) partial.
__global__ void __launch_bounds__(256) score_partial_kernel(const float* __restrict__ x,
                                                            const float* __restrict__ recon, long long plane,
                                                            int blocks_per_frame, long long chunk,
                                                            float* __restrict__ heat, float* __restrict__ partials) {
  const int f = blockIdx.x / blocks_per_frame;
  const int bi = blockIdx.x % blocks_per_frame;
  const long long p0 = bi * chunk;
  long long p1 = p0 + chunk;
  if (p1 > plane) p1 = plane;
  const float* xf = x + static_cast<long long>(f) * 3 * plane;
  const float* rf = recon + static_cast<long long>(f) * 3 * plane;
  float ssum = 0.f, smin = INFINITY, smax = -INFINITY;
  for (long long p = p0 + threadIdx.x * 4LL; p < p1; p += 256 * 4) {
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 a = __ldcs(reinterpret_cast<const float4*>(xf + c * plane + p));
      const float4 b = __ldcs(reinterpret_cast<const float4*>(rf + c * plane + p));
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      e.x += d0 * d0; e.y += d1 * d1; e.z += d2 * d2; e.w += d3 * d3;
    }
    ssum += (e.x + e.y) + (e.z + e.w);
    smin = fminf(smin, fminf(fminf(e.x, e.y), fminf(e.z, e.w)));
    smax = fmaxf(smax, fmaxf(fmaxf(e.x, e.y), fmaxf(e.z, e.w)));
    if (heat) {
      const float k = 1.f / 3.f;
      __stcs(reinterpret_cast<float4*>(heat + static_cast<long long>(f) * plane + p),
             make_float4(e.x * k, e.y * k, e.z * k, e.w * k));
    }
  }
  __shared__ float red[8][3];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
    smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[warp][0] = ssum; red[warp][1] = smin; red[warp][2] = smax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, mn = INFINITY, mx = -INFINITY;
    for (int i = 0; i < 8; ++i) { s += red[i][0]; mn = fminf(mn, red[i][1]); mx = fmaxf(mx, red[i][2]); }
    *reinterpret_cast<float4*>(partials + static_cast<long long>(blockIdx.x) * 4) =
        make_float4(s, mn * (1.f / 3.f), mx * (1.f / 3.f), 0.f);
  }
}

// One warp per frame folds that frame's tile partials in a fixed order (bitwise reproducible run to run).
__global__ void __launch_bounds__(128) score_finalize_kernel(const float* __restrict__ partials, int frames,
                                                             int tiles_per_frame, float inv_count,
                                                             float* __restrict__ score, float* __restrict__ minmax) {
  const int f = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= frames) return;
  const float4* p = reinterpret_cast<const float4*>(partials) + static_cast<long long>(f) * tiles_per_frame;
  float s = 0.f, mn = INFINITY, mx = -INFINITY;
  for (int i = lane; i < tiles_per_frame; i += 32) {
    const float4 v = p[i];
    s += v.x;
    mn = fminf(mn, v.y);
    mx = fmaxf(mx, v.z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    score[f] = s * inv_count;
    if (minmax) { minmax[2 * f] = mn; minmax[2 * f + 1] = mx; }
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, long long total, int HW, int C,
                                    float* __restrict__ dst) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;  // index in NCHW order
  if (i >= total) return;
  const long long n = i / (static_cast<long long>(C) * HW);
  const long long rem = i - n * C * HW;
  const int c = static_cast<int>(rem / HW);
  const int p = static_cast<int>(rem - static_cast<long long>(c) * HW);
  dst[i] = __bfloat162float(src[(n * HW + p) * C + c]);
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, long long total, int HW, int C,
                                    __nv_bfloat16* __restrict__ dst) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;  // index in NHWC order
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const long long np = i / C;
  const long long n = np / HW;
  const int p = static_cast<int>(np - n * HW);
  dst[i] = __float2bfloat16_rn(src[(n * C + c) * HW + p]);
}

__global__ void heatmap_u8_kernel(const float* __restrict__ heat, const float* __restrict__ minmax, long long plane,
                                  long long total, uint8_t* __restrict__ out) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= total) return;
  const long long f = i / plane;
  const float mn = minmax[2 * f], mx = minmax[2 * f + 1];
  // numpy float32 semantics of evaluate_video.py:56-57, one correctly rounded operation each (no contraction, no
  // reciprocal tricks): the bytes are bit-exact against `((e - e.min()) / (e.max() - e.min() + 1e-8) * 255).astype(uint8)`
  const float norm = __fdiv_rn(__fsub_rn(heat[i], mn), __fadd_rn(__fsub_rn(mx, mn), 1e-8f));
  out[i] = static_cast<uint8_t>(__fmul_rn(norm, 255.f));
}

// ------------------------------------------------------------------------------------------------ frame I/O helpers
// the host-side numpy / torchvision / cv2 steps either side of the scoring path.

// uint8 HWC RGB frames -> fp32 NCHW in [-1,1]: torchvision ToTensor (x/255) + Normalize(mean .5, std .5)
// (reference utils/dataset.py:65-70, utils/video_dataset.py:62-66,356-360).  One thread = 4 consecutive pixels of a row:
// 12 input bytes, three float4 stores.
__global__ void __launch_bounds__(256) u8_hwc_to_f32_nchw_kernel(const uint8_t* __restrict__ src, long long plane,
                                                                 long long quads, float* __restrict__ dst) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;  // quad index over N * plane / 4
  if (i >= quads) return;
  const long long per_frame = plane >> 2;
  const long long n = i / per_frame;
  const long long p = (i - n * per_frame) << 2;
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + (n * plane + p) * 3);  // 12 bytes, 4-byte aligned
  const uint32_t a = __ldg(s32), b = __ldg(s32 + 1), c = __ldg(s32 + 2);
  const uint8_t px[12] = {uint8_t(a), uint8_t(a >> 8), uint8_t(a >> 16), uint8_t(a >> 24), uint8_t(b), uint8_t(b >> 8),
                          uint8_t(b >> 16), uint8_t(b >> 24), uint8_t(c), uint8_t(c >> 8), uint8_t(c >> 16),
                          uint8_t(c >> 24)};
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float4 v;
    v.x = (static_cast<float>(px[ch]) / 255.f - 0.5f) / 0.5f;
    v.y = (static_cast<float>(px[3 + ch]) / 255.f - 0.5f) / 0.5f;
    v.z = (static_cast<float>(px[6 + ch]) / 255.f - 0.5f) / 0.5f;
    v.w = (static_cast<float>(px[9 + ch]) / 255.f - 0.5f) / 0.5f;
    *reinterpret_cast<float4*>(dst + (n * 3 + ch) * plane + p) = v;
  }
}

// fp32 NCHW in [-1,1] -> uint8 HWC: `denormalize` (reference evaluate_video.py:40-49): x*0.5+0.5, clamp, *255, truncate
__global__ void __launch_bounds__(256) f32_nchw_to_u8_hwc_kernel(const float* __restrict__ src, long long plane,
                                                                 long long total_px, uint8_t* __restrict__ dst) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;  // pixel index over N * plane
  if (i >= total_px) return;
  const long long n = i / plane, p = i - n * plane;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float v = __ldg(src + (n * 3 + ch) * plane + p) * 0.5f + 0.5f;
    v = fminf(fmaxf(v, 0.f), 1.f);
    dst[i * 3 + ch] = static_cast<uint8_t>(v * 255.f);
  }
}

// cv2.COLORMAP_JET as RGB (r | g << 8 | b << 16), generated with cv2.applyColorMap(arange(256)) — reference
// create_heatmap (evaluate_video.py:52-66): normalise per frame, uint8 (truncation), JET, BGR -> RGB
__constant__ uint32_t c_jet_rgb[256] = {
    0x800000u, 0x840000u, 0x880000u, 0x8c0000u, 0x900000u, 0x940000u, 0x980000u, 0x9c0000u,
    0xa00000u, 0xa40000u, 0xa80000u, 0xac0000u, 0xb00000u, 0xb40000u, 0xb80000u, 0xbc0000u,
    0xc00000u, 0xc40000u, 0xc80000u, 0xcc0000u, 0xd00000u, 0xd40000u, 0xd80000u, 0xdc0000u,
    0xe00000u, 0xe40000u, 0xe80000u, 0xec0000u, 0xf00000u, 0xf40000u, 0xf80000u, 0xfc0000u,
    0xff0000u, 0xff0400u, 0xff0800u, 0xff0c00u, 0xff1000u, 0xff1400u, 0xff1800u, 0xff1c00u,
    0xff2000u, 0xff2400u, 0xff2800u, 0xff2c00u, 0xff3000u, 0xff3400u, 0xff3800u, 0xff3c00u,
    0xff4000u, 0xff4400u, 0xff4800u, 0xff4c00u, 0xff5000u, 0xff5400u, 0xff5800u, 0xff5c00u,
    0xff6000u, 0xff6400u, 0xff6800u, 0xff6c00u, 0xff7000u, 0xff7400u, 0xff7800u, 0xff7c00u,
    0xff8000u, 0xff8400u, 0xff8800u, 0xff8c00u, 0xff9000u, 0xff9400u, 0xff9800u, 0xff9c00u,
    0xffa000u, 0xffa400u, 0xffa800u, 0xffac00u, 0xffb000u, 0xffb400u, 0xffb800u, 0xffbc00u,
    0xffc000u, 0xffc400u, 0xffc800u, 0xffcc00u, 0xffd000u, 0xffd400u, 0xffd800u, 0xffdc00u,
    0xffe000u, 0xffe400u, 0xffe800u, 0xffec00u, 0xfff000u, 0xfff400u, 0xfff800u, 0xfffc00u,
    0xfeff02u, 0xfaff06u, 0xf6ff0au, 0xf2ff0eu, 0xeeff12u, 0xeaff16u, 0xe6ff1au, 0xe2ff1eu,
    0xdeff22u, 0xdaff26u, 0xd6ff2au, 0xd2ff2eu, 0xceff32u, 0xcaff36u, 0xc6ff3au, 0xc2ff3eu,
    0xbeff42u, 0xbaff46u, 0xb6ff4au, 0xb2ff4eu, 0xaeff52u, 0xaaff56u, 0xa6ff5au, 0xa2ff5eu,
    0x9eff62u, 0x9aff66u, 0x96ff6au, 0x92ff6eu, 0x8eff72u, 0x8aff76u, 0x86ff7au, 0x82ff7eu,
    0x7eff82u, 0x7aff86u, 0x76ff8au, 0x72ff8eu, 0x6eff92u, 0x6aff96u, 0x66ff9au, 0x62ff9eu,
    0x5effa2u, 0x5affa6u, 0x56ffaau, 0x52ffaeu, 0x4effb2u, 0x4affb6u, 0x46ffbau, 0x42ffbeu,
    0x3effc2u, 0x3affc6u, 0x36ffcau, 0x32ffceu, 0x2effd2u, 0x2affd6u, 0x26ffdau, 0x22ffdeu,
    0x1effe2u, 0x1affe6u, 0x16ffeau, 0x12ffeeu, 0x0efff2u, 0x0afff6u, 0x06fffau, 0x01fffeu,
    0x00fcffu, 0x00f8ffu, 0x00f4ffu, 0x00f0ffu, 0x00ecffu, 0x00e8ffu, 0x00e4ffu, 0x00e0ffu,
    0x00dcffu, 0x00d8ffu, 0x00d4ffu, 0x00d0ffu, 0x00ccffu, 0x00c8ffu, 0x00c4ffu, 0x00c0ffu,
    0x00bcffu, 0x00b8ffu, 0x00b4ffu, 0x00b0ffu, 0x00acffu, 0x00a8ffu, 0x00a4ffu, 0x00a0ffu,
    0x009cffu, 0x0098ffu, 0x0094ffu, 0x0090ffu, 0x008cffu, 0x0088ffu, 0x0084ffu, 0x0080ffu,
    0x007cffu, 0x0078ffu, 0x0074ffu, 0x0070ffu, 0x006cffu, 0x0068ffu, 0x0064ffu, 0x0060ffu,
    0x005cffu, 0x0058ffu, 0x0054ffu, 0x0050ffu, 0x004cffu, 0x0048ffu, 0x0044ffu, 0x0040ffu,
    0x003cffu, 0x0038ffu, 0x0034ffu, 0x0030ffu, 0x002cffu, 0x0028ffu, 0x0024ffu, 0x0020ffu,
    0x001cffu, 0x0018ffu, 0x0014ffu, 0x0010ffu, 0x000cffu, 0x0008ffu, 0x0004ffu, 0x0000ffu,
    0x0000fcu, 0x0000f8u, 0x0000f4u, 0x0000f0u, 0x0000ecu, 0x0000e8u, 0x0000e4u, 0x0000e0u,
    0x0000dcu, 0x0000d8u, 0x0000d4u, 0x0000d0u, 0x0000ccu, 0x0000c8u, 0x0000c4u, 0x0000c0u,
    0x0000bcu, 0x0000b8u, 0x0000b4u, 0x0000b0u, 0x0000acu, 0x0000a8u, 0x0000a4u, 0x0000a0u,
    0x00009cu, 0x000098u, 0x000094u, 0x000090u, 0x00008cu, 0x000088u, 0x000084u, 0x000080u,
};

__global__ void __launch_bounds__(256) heatmap_jet_kernel(const float* __restrict__ heat, const float* __restrict__ minmax,
                                                          long long plane, long long total, uint8_t* __restrict__ out) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= total) return;
  const long long f = i / plane;
  const float mn = minmax[2 * f], mx = minmax[2 * f + 1];
  const float norm = __fdiv_rn(__fsub_rn(heat[i], mn), __fadd_rn(__fsub_rn(mx, mn), 1e-8f));  // as heatmap_u8_kernel
  const uint32_t c = c_jet_rgb[static_cast<uint8_t>(__fmul_rn(norm, 255.f))];
  out[i * 3 + 0] = static_cast<uint8_t>(c);
  out[i * 3 + 1] = static_cast<uint8_t>(c >> 8);
  out[i * 3 + 2] = static_cast<uint8_t>(c >> 16);
}

// Side-by-side panel of the reference's video output (evaluate_video.py:355-364, generate_visualizations :279-286):
// np.hstack([denormalize(frame), denormalize(reconstruction), create_heatmap(error_map)]) -> uint8 [F][H][3W][3].
// (create_heatmap's cv2.resize to (256, 256) is the identity at the reference's image size; other sizes stay with cv2.)
__global__ void __launch_bounds__(256) compose_panel_kernel(const float* __restrict__ x, const float* __restrict__ recon,
                                                            const float* __restrict__ heat,
                                                            const float* __restrict__ minmax, int W, long long plane,
                                                            long long total_px, uint8_t* __restrict__ out) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;  // pixel index over F * plane
  if (i >= total_px) return;
  const long long n = i / plane, p = i - n * plane;
  const long long row = p / W;
  const int col = static_cast<int>(p - row * W);
  uint8_t* o = out + ((n * (plane / W) + row) * 3LL * W + col) * 3;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    float v = __fadd_rn(__fmul_rn(__ldg(x + (n * 3 + ch) * plane + p), 0.5f), 0.5f);   // denormalize: evaluate_video.py:40-49
    v = fminf(fmaxf(v, 0.f), 1.f);
    o[ch] = static_cast<uint8_t>(__fmul_rn(v, 255.f));
    float r = __fadd_rn(__fmul_rn(__ldg(recon + (n * 3 + ch) * plane + p), 0.5f), 0.5f);
    r = fminf(fmaxf(r, 0.f), 1.f);
    o[3 * W + ch] = static_cast<uint8_t>(__fmul_rn(r, 255.f));
  }
  const float mn = minmax[2 * n], mx = minmax[2 * n + 1];
  const float norm = __fdiv_rn(__fsub_rn(heat[i], mn), __fadd_rn(__fsub_rn(mx, mn), 1e-8f));  // as heatmap_u8_kernel
  const uint32_t c = c_jet_rgb[static_cast<uint8_t>(__fmul_rn(norm, 255.f))];
  o[6 * W + 0] = static_cast<uint8_t>(c);
  o[6 * W + 1] = static_cast<uint8_t>(c >> 8);
  o[6 * W + 2] = static_cast<uint8_t>(c >> 16);
}

// ------------------------------------------------------------------------------------------------ SSIM
// SSIMLoss.forward of the reference (utils/losses.py:51-93): Gaussian-weighted (11x11, sigma 1.5, zero padding) local
// means / variances / covariance per channel, ssim = (2 mu_p mu_t + C1)(2 s_pt + C2) / ((mu_p^2 + mu_t^2 + C1)(s_pp +
// s_tt + C2)).  The Gaussian is separable: one block = one 32x32 output tile of one (frame, channel) plane; the two
// inputs come in with a 5-pixel halo, a horizontal pass forms the five row-filtered moments (p, t, p^2, t^2, p t) in
// shared memory and the vertical pass finishes them.  HBM traffic = the two images once (+ halo) and the optional map.
constexpr int kSsimTile = 32, kSsimR = 5, kSsimIn = kSsimTile + 2 * kSsimR;  // 42
__constant__ float c_ssim_gauss[11];

__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                   int H, int W, int tiles_x, int tiles_y, float* __restrict__ map,
                                                   float* __restrict__ partials) {
  __shared__ float s_p[kSsimIn][kSsimIn + 1];
  __shared__ float s_t[kSsimIn][kSsimIn + 1];
  __shared__ float s_h[5][kSsimIn][kSsimTile + 1];
  __shared__ float s_red[8];
  const int tile = blockIdx.x;
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y;
  const long long plane_idx = tile / (tiles_x * tiles_y);  // frame * 3 + channel
  const long long plane = static_cast<long long>(H) * W;
  const float* pp = pred + plane_idx * plane;
  const float* tp = target + plane_idx * plane;
  const int x0 = tx * kSsimTile - kSsimR, y0 = ty * kSsimTile - kSsimR;
  for (int i = threadIdx.x; i < kSsimIn * kSsimIn; i += 256) {
    const int iy = i / kSsimIn, ix = i - iy * kSsimIn;
    const int gy = y0 + iy, gx = x0 + ix;
    const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
    s_p[iy][ix] = in ? __ldg(pp + static_cast<long long>(gy) * W + gx) : 0.f;
    s_t[iy][ix] = in ? __ldg(tp + static_cast<long long>(gy) * W + gx) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSsimIn * kSsimTile; i += 256) {  // horizontal pass
    const int iy = i / kSsimTile, ox = i - iy * kSsimTile;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = c_ssim_gauss[k], p = s_p[iy][ox + k], t = s_t[iy][ox + k];
      a0 = fmaf(g, p, a0);
      a1 = fmaf(g, t, a1);
      a2 = fmaf(g, p * p, a2);
      a3 = fmaf(g, t * t, a3);
      a4 = fmaf(g, p * t, a4);
    }
    s_h[0][iy][ox] = a0; s_h[1][iy][ox] = a1; s_h[2][iy][ox] = a2; s_h[3][iy][ox] = a3; s_h[4][iy][ox] = a4;
  }
  __syncthreads();
  float acc = 0.f;
  for (int i = threadIdx.x; i < kSsimTile * kSsimTile; i += 256) {  // vertical pass + SSIM
    const int oy = i / kSsimTile, ox = i - oy * kSsimTile;
    const int gy = ty * kSsimTile + oy, gx = tx * kSsimTile + ox;
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float g = c_ssim_gauss[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] = fmaf(g, s_h[q][oy + k][ox], m[q]);
    }
    const float mu_pp = m[0] * m[0], mu_tt = m[1] * m[1], mu_pt = m[0] * m[1];
    const float s_pp = m[2] - mu_pp, s_tt = m[3] - mu_tt, s_pt = m[4] - mu_pt;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float v = ((2.f * mu_pt + C1) * (2.f * s_pt + C2)) / ((mu_pp + mu_tt + C1) * (s_pp + s_tt + C2));
    if (gy < H && gx < W) {
      acc += v;
      if (map) map[plane_idx * plane + static_cast<long long>(gy) * W + gx] = v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += s_red[i];
    partials[blockIdx.x] = s;
  }
}

// loss[f] = 1 - (sum of the frame's 3 * tiles partials) / (3*H*W), in a fixed order
__global__ void __launch_bounds__(128) ssim_finalize_kernel(const float* __restrict__ partials, int frames, int per_frame,
                                                            float inv_count, float* __restrict__ loss) {
  const int f = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= frames) return;
  float s = 0.f;
  for (int i = lane; i < per_frame; i += 32) s += partials[static_cast<long long>(f) * per_frame + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) loss[f] = 1.f - s * inv_count;
}

}  // namespace vad

// ================================================================================================ C ABI
using namespace vad;

extern "C" {

const char* vad_error_string(int code) {
  switch (code) {
    case VAD_OK: return "ok";
    case VAD_ERR_ARG: return "bad argument (null pointer or non-positive size)";
    case VAD_ERR_SHAPE: return "unsupported shape (H, W must be multiples of 16; channels multiples of 32)";
    case VAD_ERR_UNSUPPORTED: return "no kernel instantiation for this layer configuration";
    case VAD_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed (needs an sm_100 driver)";
    case VAD_ERR_WORKSPACE: return "workspace too small";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown error";
  }
}

int vad_version(void) { return 100; }

unsigned long long vad_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

static long long* g_timeline = nullptr;
int vad_debug_set_timeline(long long* device_buf) {  // up to 5 roles x 64 tiles x 16 events; NULL disables
  g_timeline = device_buf;
  return VAD_OK;
}

int vad_debug_set_kx(int mode) {  // -1 = back to the VAD_KX environment setting; returns the previous effective mode
  const int prev = kx_setting();
  g_kx_override = mode;
  return prev;
}

static int g_lstm_mode_override = -1;  // vad_debug_set_lstm_mode
int vad_debug_set_lstm_mode(int mode) {  // 0 one launch per step, 1 persistent sequence kernel, -1 environment default
  const int prev = g_lstm_mode_override;
  g_lstm_mode_override = mode;
  return prev;
}

int vad_debug_last_trap(unsigned long long out[4]) {
  if (!g_trap_host || !out) return VAD_ERR_ARG;
  for (int i = 0; i < 4; ++i) out[i] = g_trap_host[i];
  return VAD_OK;
}

int vad_conv_m_tiles(int B, int H, int W, int force_single_frame_tiles) {
  if (B <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  return pick_tile_geometry(B, H, W, force_single_frame_tiles != 0).m_tiles();
}

}  // extern "C"

namespace {
struct ConvLaunch {
  ConvArgs a;
  int CK, BN, epi, grid;
  bool use_halo;
  bool use_kx;
  bool use_hs;
};

int launch_built(const ConvLaunch& L, cudaStream_t stream) {
  if (L.use_hs) return launch_conv_hs(L.epi, L.a, L.grid, stream);
  if (L.use_kx) return launch_conv_kx(L.CK, L.BN, L.epi, L.a, L.grid, stream);
  if (L.use_halo) return launch_conv_halo(L.CK, L.BN, L.epi, L.a, L.grid, stream);
  return launch_conv_umma(L.CK, L.BN, L.epi, L.a, L.grid, stream);
}

// Validates a layer description and fills the kernel argument block (tensor maps, tile geometry, kernel choice).
int build_conv(const vad_conv_desc* d, ConvLaunch& L) {
  if (!d || !d->src0 || !d->weight || !d->bias) return VAD_ERR_ARG;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->n_total <= 0) return VAD_ERR_ARG;
  if (d->ntaps != 9 && d->ntaps != 1) return VAD_ERR_ARG;
  if (d->c0 <= 0 || d->c0 % 32 != 0 || d->c1 < 0 || d->c1 % 32 != 0) return VAD_ERR_SHAPE;
  if (d->c1 > 0 && !d->src1) return VAD_ERR_ARG;
  const int epi = d->epilogue;
  const bool is_score = (epi == VAD_EPI_TANH_SCORE || epi == VAD_EPI_CONVT_TANH_SCORE);
  const int CK = (d->c0 % 64 == 0 && d->c1 % 64 == 0) ? 64 : 32;
  int BN;
  switch (epi) {
    case VAD_EPI_STORE:
    case VAD_EPI_POOL:
      if (d->n_total % 32 != 0) return VAD_ERR_SHAPE;
      BN = d->n_total >= 256 && d->n_total % 256 == 0 ? 256 : d->n_total % 128 == 0 ? 128 : d->n_total % 64 == 0 ? 64 : 32;
      if (CK == 32 && BN > 64) BN = 64;
      if (!d->out) return VAD_ERR_ARG;
      if (epi == VAD_EPI_POOL && ((d->H | (d->pair_fold ? 0 : d->W)) & 1)) return VAD_ERR_SHAPE;
      if (d->pair_fold && (d->ntaps != 9 || d->c0 != 64 || d->c1 != 0 || d->n_total > 128 || d->H < 16 || d->W < 16 ||
                           (d->T0 > 1)))
        return VAD_ERR_UNSUPPORTED;  // (pair view: 2 x 32 input channels, halo kernel shapes only)
      break;
    case VAD_EPI_CONVT:
      if (d->n_total % 128 != 0 || d->cout * 4 != d->n_total || d->cout % 32 != 0) return VAD_ERR_SHAPE;
      BN = 128;
      if (!d->out) return VAD_ERR_ARG;
      break;
    case VAD_EPI_LSTM:
      if (d->n_total % 128 != 0 || d->cout * 4 != d->n_total) return VAD_ERR_SHAPE;
      BN = 128;
      if (!d->out || !d->c_state) return VAD_ERR_ARG;
      break;
    case VAD_EPI_TANH_SCORE:
    case VAD_EPI_CONVT_TANH_SCORE:
      if (d->n_total != 16 || d->cout != 3) return VAD_ERR_SHAPE;
      BN = 16;
      if (!d->x || !d->partials) return VAD_ERR_ARG;
      break;
    default:
      return VAD_ERR_ARG;
  }

  if (d->n_total > 1024) return VAD_ERR_SHAPE;  // bias slab in shared memory (ConvLSTM hidden <= 256, ConvT Cout <= 256)

  ensure_trap_slot();
  ConvArgs& a = L.a;
  std::memset(&a, 0, sizeof(a));
  TileGeom g = pick_tile_geometry(d->B, d->H, d->W, is_score);

  // ---- kx-merged halo kernel: narrow 3x3 layers whose caller supplied the [kx*Cout + co][ky*Cin + ci] weight matrix
  const bool halo_shape = d->ntaps == 9 && d->c1 == 0 && (d->c0 == 32 || d->c0 == 64) && d->n_total == BN &&
                          (d->T0 <= 1) && d->H >= 16 && d->W >= 16 &&
                          (epi == VAD_EPI_STORE || epi == VAD_EPI_POOL || epi == VAD_EPI_TANH_SCORE);
  bool use_kx = kx_setting() != 0 && d->weight_kx != nullptr && halo_shape && BN <= (kx_setting() >= 3 ? 64 : kx_setting() == 2 ? 32 : 16) &&
                kx_fixed_smem_bytes(CK, BN, epi) > 0;
  if (use_kx) {
    g.lgTW = 3; g.lgTH = 4; g.lgTN = 0;
    g.tiles_w = (d->W + 5) / 6;
    g.tiles_h = (d->H + 15) >> 4;
    g.tiles_b = d->B;
    const int patch = 8 * 18 * CK * 2;
    int stages = 8;
    while (stages >= 2 && kx_fixed_smem_bytes(CK, BN, epi) + stages * patch > 227 * 1024 - 4096) --stages;
    // (An odd ring depth is fine although the two MMA issuers then alternate on a slot: the turn token orders their
    // issues, so an issuer reaches the wait for use k+1 of a slot only after use k has been issued — it can never be
    // two barrier phases ahead.  The first conv's free-running converter warps have no such ordering; see there.)
    if (stages < 2) {
      use_kx = false;
      g = pick_tile_geometry(d->B, d->H, d->W, is_score);
    } else {
      a.halo_stages = stages;
    }
  }

  // ---- patch + streamed-weights kernel for wide 3x3 layers: pairs of 8x16 tiles, N tiles of 128
  bool use_hs = !d->pair_fold && hs_setting() != 0 && d->ntaps == 9 && d->c1 == 0 && d->c0 % 64 == 0 &&
                (d->c0 >= 128 || hs_setting() >= 2 || epi == VAD_EPI_STORE) && d->n_total % 128 == 0 && (d->T0 <= 1) &&
                d->H >= 8 &&
                d->W % 16 == 0 && (epi == VAD_EPI_STORE || epi == VAD_EPI_POOL) && hs_patch_stages(epi) >= 2;
  if (use_hs) {
    BN = 128;
    g.lgTW = 3; g.lgTH = 4; g.lgTN = 0;
    g.tiles_w = d->W >> 3;
    g.tiles_h = (d->H + 15) >> 4;
    g.tiles_b = d->B;
    a.halo_stages = hs_patch_stages(epi);
  }

  // ---- halo kernel: 3x3 conv, one source of 32/64 channels, all of N in one tile, frames of at least one tile
  const int halo_mode = halo_mode_setting();
  bool use_halo = !use_kx && !use_hs && halo_mode != 0 && d->ntaps == 9 && d->c1 == 0 && (d->c0 == 32 || d->c0 == 64) &&
                  d->n_total == BN && (d->T0 <= 1) && d->H >= 16 && d->W >= 16 &&
                  (epi == VAD_EPI_STORE || epi == VAD_EPI_POOL || epi == VAD_EPI_TANH_SCORE);
  int halo_box_w = 0, halo_box_h = 0;
  if (use_halo) {
    const bool three = (halo_mode == 3);
    g.lgTW = three ? 4 : 3;
    g.lgTH = three ? 3 : 4;
    g.lgTN = 0;
    g.tiles_w = (d->W + (1 << g.lgTW) - 1) >> g.lgTW;
    g.tiles_h = (d->H + (1 << g.lgTH) - 1) >> g.lgTH;
    g.tiles_b = d->B;
    const int TW = 1 << g.lgTW, TH = 1 << g.lgTH;
    a.halo_npatch = three ? 3 : 1;
    a.halo_pw = halo_box_w = three ? TW : TW + 2;
    a.halo_ph = halo_box_h = TH + 2;
    a.halo_patch_bytes = (a.halo_pw * a.halo_ph * CK * 2 + 1023) / 1024 * 1024;
    a.halo_sbo_rows = three ? 8 : a.halo_pw;
    a.halo_base_mode = (halo_mode == 2) ? 1 : 0;
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
      a.tap_patch[tap] = three ? kx : 0;
      a.tap_row[tap] = three ? ky * a.halo_pw : ky * a.halo_pw + kx;
    }
    const int per_stage = a.halo_patch_bytes * a.halo_npatch;
    int stages = 8;
    while (stages >= 2 && halo_smem_bytes(CK, BN, epi, per_stage, stages) > 227 * 1024 - 4096) --stages;
    if (stages < 2 || halo_smem_bytes(CK, BN, epi, per_stage, stages) == 0) {
      use_halo = false;
      g = pick_tile_geometry(d->B, d->H, d->W, is_score);
    } else {
      a.halo_stages = stages;
    }
  }

  int rc;
  if (use_hs) {
    rc = encode_act_map_box(&a.mapA0, d->src0, d->c0, d->W, d->H, 1, d->B, 64, 18, 18, 1);
  } else if (use_kx) {
    rc = encode_act_map_box(&a.mapA0, d->src0, d->c0, d->W, d->H, 1, d->B, CK, 8, 18, 1);
  } else if (use_halo) {
    // box {CK, PW, PH, 1, 1}: PW/PH need not be powers of two
    rc = encode_act_map_box(&a.mapA0, d->src0, d->c0, d->W, d->H, 1, d->B, CK, halo_box_w, halo_box_h, 1);
  } else {
    rc = encode_act_map(&a.mapA0, d->src0, d->c0, d->W, d->H, d->T0 > 0 ? d->T0 : 1, d->B, CK, g);
  }
  if (rc != VAD_OK) return rc;
  if (d->c1 > 0) {
    rc = encode_act_map(&a.mapA1, d->src1, d->c1, d->W, d->H, d->T1 > 0 ? d->T1 : 1, d->B, CK, g);
    if (rc != VAD_OK) return rc;
  }
  const int w_ctap = d->w_ctap > 0 ? d->w_ctap : d->c0 + d->c1;
  if (w_ctap < d->c0 + d->c1) return VAD_ERR_ARG;
  const int K = d->ntaps * w_ctap;
  if (use_kx) {  // [kx*Cout + co (zero rows up to the MMA N)][ky*Cin + ci]
    const int nm = kx_mma_columns(BN, epi);
    rc = encode_weight_map(&a.mapB, d->weight_kx, 3 * d->c0, nm, CK, nm);
  } else {
    rc = encode_weight_map(&a.mapB, d->weight, K, d->n_total, CK, BN);
  }
  if (rc != VAD_OK) return rc;
  a.chunks0 = d->c0 / CK;
  a.chunks1 = d->c1 / CK;
  a.ntaps = d->ntaps;
  a.w_ctap = use_kx ? d->c0 : w_ctap;
  a.pair = use_hs ? 2 : 1;
  a.w_step = use_kx ? 6 : (1 << g.lgTW);
  a.tw_valid = use_kx ? 6 : (1 << g.lgTW);
  a.tA0 = d->t0;
  a.tA1 = d->t1;
  a.B = d->B; a.H = d->H; a.W = d->W;
  a.lgTW = g.lgTW; a.lgTH = g.lgTH; a.lgTN = g.lgTN;
  a.tiles_w = g.tiles_w; a.tiles_h = g.tiles_h; a.tiles_b = g.tiles_b;
  a.n_tiles = d->n_total / BN;
  a.total_tiles = g.m_tiles() * a.n_tiles;
  a.bias = d->bias;
  a.slope = d->slope;
  a.out = d->out;
  a.out_fs = d->out_frame_stride;
  a.out_cp = d->out_cpitch;
  a.cout = d->cout;
  a.c_state = d->c_state;
  a.lstm_first = d->lstm_first;
  a.x = d->x; a.recon = d->recon; a.heat = d->heat; a.partials = d->partials;
  a.pair_fold = d->pair_fold ? 1 : 0;
  if (d->pair_fold && !use_halo) return VAD_ERR_UNSUPPORTED;  // only the resident-weights patch kernel skips the zero K steps

  // ---- output tensor map: the epilogue stages each bf16 chunk in swizzled smem and TMA-stores it
  if (tma_store_setting() && (epi == VAD_EPI_STORE || epi == VAD_EPI_POOL || epi == VAD_EPI_CONVT || epi == VAD_EPI_LSTM)) {
    const int TW = 1 << g.lgTW, TH = 1 << g.lgTH, TN = 1 << g.lgTN;
    const long long cp = d->out_cpitch;
    const bool aligned = (reinterpret_cast<uintptr_t>(d->out) % 16 == 0) && cp % 8 == 0 && d->out_frame_stride % 8 == 0;
    if (aligned && epi == VAD_EPI_STORE) {
      a.out_chunk = (BN % 64 == 0) ? 64 : 32;
      cuuint64_t dims[5] = {(cuuint64_t)d->n_total, (cuuint64_t)d->W, (cuuint64_t)d->H, 1, (cuuint64_t)d->B};
      cuuint64_t st[4] = {(cuuint64_t)cp * 2, (cuuint64_t)d->W * cp * 2, (cuuint64_t)d->H * d->W * cp * 2,
                          (cuuint64_t)d->out_frame_stride * 2};
      cuuint32_t box[5] = {(cuuint32_t)a.out_chunk, (cuuint32_t)(use_kx ? 6 : TW), (cuuint32_t)TH, 1, (cuuint32_t)TN};
      a.tma_store = encode_map5(&a.mapOut, d->out, dims, st, box, a.out_chunk) == VAD_OK;
    } else if (aligned && epi == VAD_EPI_LSTM) {
      a.out_chunk = 32;
      cuuint64_t dims[5] = {(cuuint64_t)d->cout, (cuuint64_t)d->W, (cuuint64_t)d->H, 1, (cuuint64_t)d->B};
      cuuint64_t st[4] = {(cuuint64_t)cp * 2, (cuuint64_t)d->W * cp * 2, (cuuint64_t)d->H * d->W * cp * 2,
                          (cuuint64_t)d->out_frame_stride * 2};
      cuuint32_t box[5] = {32, (cuuint32_t)TW, (cuuint32_t)TH, 1, (cuuint32_t)TN};
      a.tma_store = encode_map5(&a.mapOut, d->out, dims, st, box, 32) == VAD_OK;
    } else if (aligned && epi == VAD_EPI_POOL && TW >= 2 && TH >= 2 && !pool_direct_setting() && !d->pair_fold) {
      a.out_chunk = (BN % 64 == 0) ? 64 : 32;
      const int Wo = d->W / 2, Ho = d->H / 2;
      cuuint64_t dims[5] = {(cuuint64_t)d->n_total, (cuuint64_t)Wo, (cuuint64_t)Ho, 1, (cuuint64_t)d->B};
      cuuint64_t st[4] = {(cuuint64_t)cp * 2, (cuuint64_t)Wo * cp * 2, (cuuint64_t)Ho * Wo * cp * 2,
                          (cuuint64_t)d->out_frame_stride * 2};
      cuuint32_t box[5] = {(cuuint32_t)a.out_chunk, (cuuint32_t)(use_kx ? 3 : TW / 2), (cuuint32_t)(TH / 2), 1, (cuuint32_t)TN};
      a.tma_store = encode_map5(&a.mapOut, d->out, dims, st, box, a.out_chunk) == VAD_OK;
    } else if (aligned && epi == VAD_EPI_CONVT) {
      // pixel shuffle as a 5-D map {co, dj, w, di, b*H + h}; needs dense frames and tiles whose rows are
      // consecutive in the merged (b, h) dimension
      const bool dense = d->out_frame_stride == 4LL * d->H * d->W * cp;
      const bool rows_ok = (TN == 1 && d->H % TH == 0) || (TN > 1 && TH == d->H);
      if (dense && rows_ok) {
        a.out_chunk = (d->cout % 64 == 0) ? 64 : 32;
        cuuint64_t dims[5] = {(cuuint64_t)d->cout, 2, (cuuint64_t)d->W, 2, (cuuint64_t)d->B * d->H};
        cuuint64_t st[4] = {(cuuint64_t)cp * 2, (cuuint64_t)2 * cp * 2, (cuuint64_t)2 * d->W * cp * 2,
                            (cuuint64_t)4 * d->W * cp * 2};
        cuuint32_t box[5] = {(cuuint32_t)a.out_chunk, 1, (cuuint32_t)TW, 1, (cuuint32_t)(TH * TN)};
        a.tma_store = encode_map5(&a.mapOut, d->out, dims, st, box, a.out_chunk) == VAD_OK;
      } else if (TN == 1) {
        // frame heights that are not a multiple of the tile height (720p: 90 and 45 rows against 16-row tiles): the merged
        // (b, h) dimension would let a tile's overhanging rows land in the next frame, and the direct-store fallback costs
        // 30-40 % of these HBM-bound layers.  Two maps instead, one per output row parity di, with h and b as separate
        // dimensions: {co, dj, w, h, b}, base = out + di * (one output row).
        a.out_chunk = (d->cout % 64 == 0) ? 64 : 32;
        cuuint64_t dims[5] = {(cuuint64_t)d->cout, 2, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->B};
        cuuint64_t st[4] = {(cuuint64_t)cp * 2, (cuuint64_t)2 * cp * 2, (cuuint64_t)4 * d->W * cp * 2,
                            (cuuint64_t)d->out_frame_stride * 2};
        cuuint32_t box[5] = {(cuuint32_t)a.out_chunk, 1, (cuuint32_t)TW, (cuuint32_t)TH, 1};
        const __nv_bfloat16* row1 = static_cast<const __nv_bfloat16*>(d->out) + 2LL * d->W * cp;
        a.convt_split = encode_map5(&a.mapOut, d->out, dims, st, box, a.out_chunk) == VAD_OK &&
                        encode_map5(&a.mapOut2, row1, dims, st, box, a.out_chunk) == VAD_OK;
        a.tma_store = a.convt_split;
      }
    }
  }

  a.timeline = g_timeline;
  a.dbg = env_int("VAD_DBG", 0);
  a.pdl = pdl_all_setting() ? 1 : 0;
  a.dual_mma = (dual_mma_setting() & ((use_halo || use_kx) ? 1 : 4)) != 0;
  a.token = (token_setting() & (use_kx ? 8 : 1)) != 0 || ((use_halo || use_kx) && ((a.halo_stages & 1) || CK == 64));
  L.use_kx = use_kx;
  L.use_hs = use_hs;
  L.CK = CK;
  L.BN = BN;
  L.epi = epi;
  L.use_halo = use_halo;
  L.grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
  if (use_hs) L.grid = (a.total_tiles >> 1) < sm_count() ? (a.total_tiles >> 1) : sm_count();
  // Transposed convolutions on the streaming kernel keep their weight tiles resident (see conv_umma_kernel): needs one
  // tap, one source, a k loop that fits the ring (5 slots of 32 KB for CK = 64, BN = 128) and a grid that is a multiple
  // of n_tiles, so that every tile of a CTA has the same n0.  VAD_CONVT_RESIDENT=0 switches it off.
  a.b_resident = 0;
  static const int resident_env = env_int("VAD_CONVT_RESIDENT", 1);
  if (resident_env && !use_hs && !use_kx && !use_halo && epi == VAD_EPI_CONVT && d->ntaps == 1 && a.chunks1 == 0 && CK == 64 &&
      BN == 128 && a.ntaps * a.chunks0 <= 4 && L.grid >= a.n_tiles && a.n_tiles <= 8) {
    L.grid = (L.grid / a.n_tiles) * a.n_tiles;
    a.b_resident = 1;
  }
  return VAD_OK;
}
}  // namespace

extern "C" {

int vad_conv_layer(const vad_conv_desc* d, vad_stream_t stream_) {
  ConvLaunch L;
  const int rc = build_conv(d, L);
  if (rc != VAD_OK) return rc;
  return launch_built(L, static_cast<cudaStream_t>(stream_));
}

int vad_conv_layer_tiles(const vad_conv_desc* d) {
  ConvLaunch L;
  const int rc = build_conv(d, L);
  if (rc != VAD_OK) return rc;
  return L.a.total_tiles / L.a.n_tiles;
}

// ConvT(64->32)+ReLU -> ConvT(32->3)+tanh -> score in one kernel (convt2_score_kernel).  `d` describes the FIRST
// transposed convolution plus the score outputs; weight2 / bias2 are the second one's [16][32] / [16].
static int build_convt2_score(const vad_conv_desc* d, const void* weight2, const float* bias2, ConvArgs& a, int& grid) {
  if (!d || !d->src0 || !d->weight || !d->bias || !weight2 || !bias2 || !d->x || !d->partials) return VAD_ERR_ARG;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0) return VAD_ERR_ARG;
  if (d->ntaps != 1 || d->c1 != 0) return VAD_ERR_ARG;
  if (d->c0 != 64 || d->n_total != 128 || d->cout != 32) return VAD_ERR_UNSUPPORTED;
  if (d->w_ctap != 0 && d->w_ctap != 64) return VAD_ERR_ARG;
  ensure_trap_slot();
  std::memset(&a, 0, sizeof(a));
  const TileGeom g = pick_tile_geometry(d->B, d->H, d->W, true);
  int rc = encode_act_map(&a.mapA0, d->src0, 64, d->W, d->H, d->T0 > 0 ? d->T0 : 1, d->B, 64, g);
  if (rc != VAD_OK) return rc;
  rc = encode_weight_map(&a.mapB, d->weight, 64, 128, 64, 128);
  if (rc != VAD_OK) return rc;
  rc = encode_weight_map(&a.mapA1, weight2, 32, 16, 32, 16);  // (mapA1 carries the second layer's weights here)
  if (rc != VAD_OK) return rc;
  a.chunks0 = 1;
  a.ntaps = 1;
  a.w_ctap = 64;
  a.pair = 1;
  a.w_step = a.tw_valid = 1 << g.lgTW;
  a.tA0 = d->t0;
  a.B = d->B; a.H = d->H; a.W = d->W;
  a.lgTW = g.lgTW; a.lgTH = g.lgTH; a.lgTN = g.lgTN;
  a.tiles_w = g.tiles_w; a.tiles_h = g.tiles_h; a.tiles_b = g.tiles_b;
  a.n_tiles = 1;
  a.total_tiles = g.m_tiles();
  a.bias = d->bias;
  a.bias2 = bias2;
  a.slope = d->slope;
  a.cout = 32;
  a.x = d->x; a.recon = d->recon; a.heat = d->heat; a.partials = d->partials;
  a.dbg = env_int("VAD_DBG", 0);
  a.pdl = pdl_all_setting() ? 1 : 0;
  grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
  return VAD_OK;
}

int vad_convt2_score(const vad_conv_desc* d, const void* weight2, const float* bias2, vad_stream_t stream_) {
  ConvArgs a;
  int grid = 0;
  const int rc = build_convt2_score(d, weight2, bias2, a, grid);
  if (rc != VAD_OK) return rc;
  return launch_convt2_score(a, grid, static_cast<cudaStream_t>(stream_));
}

int vad_convt2_score_tiles(const vad_conv_desc* d) {
  if (!d || d->B <= 0 || d->H <= 0 || d->W <= 0) return VAD_ERR_ARG;
  if (d->ntaps != 1 || d->c1 != 0) return VAD_ERR_ARG;
  if (d->c0 != 64 || d->n_total != 128 || d->cout != 32) return VAD_ERR_UNSUPPORTED;
  return pick_tile_geometry(d->B, d->H, d->W, true).m_tiles();
}

// ConvT(32->32)+ReLU -> Conv3x3(32->3)+tanh -> score in one kernel (convt_conv_score_kernel).  `d` describes the
// transposed convolution plus the score outputs; weight2_kx / bias2 are the 3x3 conv's kx-folded [16][96] / [16].
static bool convt_conv_score_shape_ok(const vad_conv_desc* d) {
  return d->c0 == 32 && d->n_total == 128 && d->cout == 32;
}
static void convt_conv_score_tiles(const vad_conv_desc* d, int& tiles_w, int& tiles_h) {
  tiles_h = (2 * d->H + 1 + 13) / 14;  // tile th scores output rows [14*th - 1, 14*th + 13)
  tiles_w = (2 * d->W + 1 + 29) / 30;  // tile tw scores output columns [30*tw - 1, 30*tw + 29)
}
static int build_convt_conv_score(const vad_conv_desc* d, const void* weight2_kx, const float* bias2, ConvArgs& a,
                                  int& grid) {
  if (!d || !d->src0 || !d->weight || !d->bias || !weight2_kx || !bias2 || !d->x || !d->partials) return VAD_ERR_ARG;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0) return VAD_ERR_ARG;
  if (d->ntaps != 1 || d->c1 != 0 || d->T0 > 1) return VAD_ERR_ARG;
  if (!convt_conv_score_shape_ok(d)) return VAD_ERR_UNSUPPORTED;
  if (d->w_ctap != 0 && d->w_ctap != 32) return VAD_ERR_ARG;
  if (12LL * d->H * d->W >= 2147483647LL) return VAD_ERR_SHAPE;  // the kernel indexes a frame's three planes with ints
  ensure_trap_slot();
  std::memset(&a, 0, sizeof(a));
  int rc = encode_act_map_box(&a.mapA0, d->src0, 32, d->W, d->H, 1, d->B, 32, 16, 8, 1);
  if (rc != VAD_OK) return rc;
  rc = encode_weight_map(&a.mapB, d->weight, 32, 128, 32, 128);
  if (rc != VAD_OK) return rc;
  rc = encode_weight_map(&a.mapA1, weight2_kx, 96, 16, 32, 16);  // (mapA1 carries the 3x3 conv's weights here)
  if (rc != VAD_OK) return rc;
  a.chunks0 = 1;
  a.ntaps = 1;
  a.w_ctap = 32;
  a.pair = 1;
  a.B = d->B; a.H = d->H; a.W = d->W;
  convt_conv_score_tiles(d, a.tiles_w, a.tiles_h);
  a.tiles_b = d->B;
  a.n_tiles = 1;
  a.total_tiles = a.tiles_w * a.tiles_h * a.tiles_b;
  a.bias = d->bias;
  a.bias2 = bias2;
  a.slope = d->slope;
  a.cout = 32;
  a.x = d->x; a.recon = d->recon; a.heat = d->heat; a.partials = d->partials;
  a.dbg = env_int("VAD_DBG", 0);
  a.pdl = pdl_all_setting() ? 1 : 0;
  grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
  return VAD_OK;
}

int vad_convt_conv_score(const vad_conv_desc* d, const void* weight2_kx, const float* bias2, vad_stream_t stream_) {
  ConvArgs a;
  int grid = 0;
  const int rc = build_convt_conv_score(d, weight2_kx, bias2, a, grid);
  if (rc != VAD_OK) return rc;
  return launch_convt_conv_score(a, grid, static_cast<cudaStream_t>(stream_));
}

int vad_convt_conv_score_tiles(const vad_conv_desc* d) {
  if (!d || d->B <= 0 || d->H <= 0 || d->W <= 0) return VAD_ERR_ARG;
  if (d->ntaps != 1 || d->c1 != 0) return VAD_ERR_ARG;
  if (!convt_conv_score_shape_ok(d)) return VAD_ERR_UNSUPPORTED;
  int tw, th;
  convt_conv_score_tiles(d, tw, th);
  return tw * th * d->B;
}

// Validates one ConvLSTM layer description and fills the kernel argument block; `patch_ok` tells whether the arguments
// were set up for the persistent patch kernel (every tile has its own resident CTA, tile shapes apply, mode 1).
static int build_lstm_layer(const vad_conv_desc* d, int T, ConvLaunch& L, bool& patch_ok, int& seq,
                            bool allow_patch = true) {
  // d describes a generic step t >= 1: src0 = layer input sequence [B][T][h][w][c0], src1 = out = hidden sequence
  // [B][T][h][w][hid] (step t reads h_{t-1} from it and writes h_t into it), c_state fp32 [B][h][w][hid].
  patch_ok = false;
  if (!d || T <= 0 || d->epilogue != VAD_EPI_LSTM || !d->src1 || d->src1 != d->out || d->c1 != d->cout) return VAD_ERR_ARG;
  if (d->T0 != T || d->T1 != T) return VAD_ERR_ARG;
  int rc = build_conv(d, L);
  if (rc != VAD_OK) return rc;
  ConvArgs& a = L.a;
  const long long step_elems = static_cast<long long>(d->H) * d->W * d->out_cpitch;
  if (d->out_frame_stride != step_elems * T) return VAD_ERR_ARG;
  if (a.tma_store) {  // output map over the whole sequence; the step index becomes the T coordinate of the store
    const int TW = 1 << a.lgTW, TH = 1 << a.lgTH, TN = 1 << a.lgTN;
    const long long cp = d->out_cpitch;
    cuuint64_t dims[5] = {(cuuint64_t)d->cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)T, (cuuint64_t)d->B};
    cuuint64_t st[4] = {(cuuint64_t)cp * 2, (cuuint64_t)d->W * cp * 2, (cuuint64_t)step_elems * 2,
                        (cuuint64_t)d->out_frame_stride * 2};
    cuuint32_t box[5] = {32, (cuuint32_t)TW, (cuuint32_t)TH, 1, (cuuint32_t)TN};
    if (encode_map5(&a.mapOut, d->out, dims, st, box, 32) != VAD_OK) a.tma_store = 0;
  }
  // One persistent launch for the whole sequence when every tile gets its own resident CTA (cell state in registers,
  // grid-wide step counter); VAD_LSTM_SEQ=0 keeps one launch per step.
  static const int seq_env = env_int("VAD_LSTM_SEQ", 1);
  seq = g_lstm_mode_override >= 0 ? g_lstm_mode_override : seq_env;
  // Patch variant (A operand through one patch per chunk, weights through their own ring): 8x8 frames (two per tile) or
  // tiles of 8 x 16 pixels inside larger frames.  VAD_LSTM_SEQ=2 disables it (streaming sequence kernel instead).
  if (allow_patch && seq == 1 && a.tma_store && L.CK == 64 && T <= 4096 && !L.use_halo && !L.use_kx && !L.use_hs) {
    const bool geo1 = d->H == 8 && d->W == 8;
    const bool geo2 = !geo1 && d->H >= 8 && d->W >= 8;
    TileGeom g;
    g.lgTW = 3;
    g.lgTH = geo1 ? 3 : 4;
    g.lgTN = geo1 ? 1 : 0;
    g.tiles_w = (d->W + 7) >> 3;
    g.tiles_h = (d->H + (1 << g.lgTH) - 1) >> g.lgTH;
    g.tiles_b = (d->B + (1 << g.lgTN) - 1) >> g.lgTN;
    if ((geo1 || geo2) && g.m_tiles() * a.n_tiles <= sm_count()) {
      const long long cp = d->out_cpitch;
      auto patch_map = [&](CUtensorMap* m, const void* base, int C) -> int {
        if (!geo1) return encode_act_map_box(m, base, C, d->W, d->H, T, d->B, 64, 10, 18, 1);
        // dims {C, W, B, H, T}: the frame index sits between W and H so that the patch lands as [y][frame][x]
        cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)d->W, (cuuint64_t)d->B, (cuuint64_t)d->H, (cuuint64_t)T};
        cuuint64_t st[4] = {(cuuint64_t)C * 2, (cuuint64_t)T * d->H * d->W * C * 2, (cuuint64_t)d->W * C * 2,
                            (cuuint64_t)d->H * d->W * C * 2};
        cuuint32_t box[5] = {64, 10, 2, 10, 1};
        return encode_map5(m, base, dims, st, box, 64);
      };
      rc = patch_map(&a.mapA0, d->src0, d->c0);
      if (rc == VAD_OK) rc = patch_map(&a.mapA1, d->src1, d->c1);
      if (rc == VAD_OK) {
        cuuint64_t dims[5] = {(cuuint64_t)d->cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)T, (cuuint64_t)d->B};
        cuuint64_t st[4] = {(cuuint64_t)cp * 2, (cuuint64_t)d->W * cp * 2, (cuuint64_t)step_elems * 2,
                            (cuuint64_t)d->out_frame_stride * 2};
        cuuint32_t box[5] = {32, 8, 1u << g.lgTH, 1, 1u << g.lgTN};
        rc = encode_map5(&a.mapOut, d->out, dims, st, box, 32);
      }
      if (rc != VAD_OK) return rc;
      a.lgTW = g.lgTW; a.lgTH = g.lgTH; a.lgTN = g.lgTN;
      a.tiles_w = g.tiles_w; a.tiles_h = g.tiles_h; a.tiles_b = g.tiles_b;
      a.total_tiles = g.m_tiles() * a.n_tiles;
      a.w_step = 8; a.tw_valid = 8; a.pair = 1;
      a.row_perm = geo1 ? 2 : 0;
      a.out = d->out;
      patch_ok = true;
    }
  }
  return VAD_OK;
}

// Both layers of a two-layer ConvLSTM in ONE persistent launch, as a wavefront: the CTA alternates between layer 1's
// step t+1 and layer 2's step t, so one layer's recurrence chain (gates -> store -> publish -> acquire -> patch load)
// runs under the other's MMAs.  d1 / d2 as for vad_convlstm_sequence, with d2->src0 == d1->out.  VAD_ERR_UNSUPPORTED when
// the shapes do not fit the persistent patch kernel (callers then run the layers one after the other).
int vad_convlstm2_sequence(const vad_conv_desc* d1, const vad_conv_desc* d2, int T, vad_stream_t stream_) {
  if (!d1 || !d2) return VAD_ERR_ARG;
  if (d2->src0 != d1->out || d2->c0 != d1->cout || d1->B != d2->B || d1->H != d2->H || d1->W != d2->W) return VAD_ERR_ARG;
  static const int lstm2_env = env_int("VAD_LSTM2", 1);
  if (!lstm2_env) return VAD_ERR_UNSUPPORTED;
  ConvLaunch L1, L2;
  bool ok1 = false, ok2 = false;
  int seq = 0;
  int rc = build_lstm_layer(d1, T, L1, ok1, seq);
  if (rc != VAD_OK) return rc;
  rc = build_lstm_layer(d2, T, L2, ok2, seq);
  if (rc != VAD_OK) return rc;
  if (!ok1 || !ok2 || L1.a.total_tiles != L2.a.total_tiles || L1.a.n_tiles != L2.a.n_tiles) return VAD_ERR_UNSUPPORTED;
  return launch_convlstm2_patch(L1.a, L2.a, T, L1.a.total_tiles, static_cast<unsigned int*>(d1->scratch),
                                static_cast<cudaStream_t>(stream_));
}

int vad_convlstm_sequence(const vad_conv_desc* d, int T, vad_stream_t stream_) {
  ConvLaunch L;
  bool patch_ok = false;
  int seq = 0;
  int rc = build_lstm_layer(d, T, L, patch_ok, seq);
  if (rc != VAD_OK) return rc;
  ConvArgs& a = L.a;
  const long long step_elems = static_cast<long long>(d->H) * d->W * d->out_cpitch;
  const int chunks1 = a.chunks1;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  unsigned int* counters = static_cast<unsigned int*>(d->scratch);
  if (patch_ok) {
    rc = launch_convlstm_patch(a, T, a.total_tiles, counters, stream);
    if (rc != VAD_ERR_UNSUPPORTED) return rc;  // (unsupported = the grid cannot be co-resident here: per-step launches)
    rc = build_lstm_layer(d, T, L, patch_ok, seq, /*allow_patch=*/false);  // back to the per-step tile geometry / maps
    if (rc != VAD_OK) return rc;
  } else if (seq && a.tma_store && a.total_tiles <= sm_count() && T <= 4096 && !L.use_halo && !L.use_kx && !L.use_hs) {
    a.out = d->out;
    rc = launch_convlstm_seq(L.CK, a, T, a.total_tiles, counters, stream);
    if (rc != VAD_ERR_UNSUPPORTED) return rc;
  }
  static const int pdl = env_int("VAD_PDL", 1);  // 0: plain stream order between the steps
  for (int t = 0; t < T; ++t) {
    // Step t >= 1 (pdl 2) starts its x half of the K loop BEFORE waiting for the previous launch, and every kernel
    // releases its dependents as soon as it starts, so a whole run of step kernels can be resident while whatever
    // produced the x sequence (the encoder's last layer; layer 1 for layer 2) is still running.  Step 0 is therefore a
    // plain stream-ordered launch: it — and with it every later step — starts only after ALL earlier work completed.
    a.pdl = (pdl && t > 0) ? 2 : 0;
    a.tA0 = t;
    a.tA1 = t > 0 ? t - 1 : 0;
    a.chunks1 = t > 0 ? chunks1 : 0;  // step 0: h_{-1} = 0, skip the h half of K
    a.lstm_first = t == 0;
    a.out_t = t;
    a.out = reinterpret_cast<__nv_bfloat16*>(d->out) + t * step_elems;  // direct-store fallback path
    rc = launch_built(L, stream);
    if (rc != VAD_OK) return rc;
  }
  return VAD_OK;
}

int vad_first_conv(const float* x, const float* weight, const float* bias, int cout, float slope, int pool, int B,
                   int H, int W, void* out, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !weight || !bias || !out || B <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  if (cout != 32) return VAD_ERR_UNSUPPORTED;
  if (pool && ((H | W) & 1)) return VAD_ERR_SHAPE;
  const int tiles_x = (W + 31) / 32, tiles_y = (H + 7) / 8;
  const long long blocks = static_cast<long long>(tiles_x) * tiles_y * B;
  if (blocks > 0x7fffffffLL) return VAD_ERR_SHAPE;
  if (pool)
    first_conv_kernel<32, true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        x, weight, bias, slope, B, H, W, reinterpret_cast<__nv_bfloat16*>(out), tiles_x, tiles_y);
  else
    first_conv_kernel<32, false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        x, weight, bias, slope, B, H, W, reinterpret_cast<__nv_bfloat16*>(out), tiles_x, tiles_y);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int vad_first_conv_tc(const float* x, const void* weight, const float* bias, float slope, int pool, int B, int H,
                      int W, void* out, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !weight || !bias || !out || B <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  if (pool && ((H | W) & 1)) return VAD_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(weight) % 16 != 0 || reinterpret_cast<uintptr_t>(out) % 16 != 0) return VAD_ERR_ARG;
  ensure_trap_slot();
  ConvArgs a;
  std::memset(&a, 0, sizeof(a));
  a.lgTW = 4; a.lgTH = 3; a.lgTN = 0;  // 8 x 16 pixel tiles inside one frame
  a.w_step = 16; a.tw_valid = 16;
  a.row_perm = 0;
  a.pair = 1;
  a.tiles_w = (W + 15) / 16;
  a.tiles_h = (H + 7) / 8;
  a.tiles_b = B;
  a.n_tiles = 1;
  const long long tiles = static_cast<long long>(a.tiles_w) * a.tiles_h * B;
  if (tiles > 0x7fffffffLL) return VAD_ERR_SHAPE;
  a.total_tiles = static_cast<int>(tiles);
  a.B = B; a.H = H; a.W = W;
  a.bias = bias;
  a.slope = slope;
  a.x = x;
  a.w_first = weight;
  a.dbg = env_int("VAD_DBG", 0);
  a.dual_mma = (dual_mma_setting() & 2) != 0;
  a.token = (token_setting() & 2) != 0;
  a.pdl = pdl_all_setting() ? 1 : 0;
  a.timeline = g_timeline;
  a.out = out;
  a.cout = 32;
  a.out_cp = 32;
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  a.out_fs = static_cast<long long>(Ho) * Wo * 32;
  a.out_chunk = 32;
  {
    // fp32 NCHW input as a 4-D map {W, H, 3, B}; box = 24 x 10 x 3 patch starting at column w0-4 (16-byte aligned
    // start; columns 3..20 are used), zero fill outside the frame = conv padding
    EncodeTiledFn fn = encode_fn();
    if (!fn) return VAD_ERR_DRIVER;
    if (reinterpret_cast<uintptr_t>(x) % 16 != 0 || W % 4 != 0) return VAD_ERR_SHAPE;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
    cuuint64_t st[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
    cuuint32_t box[4] = {24, 10, 3, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (fn(&a.mapA0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, st, box, es,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return VAD_ERR_DRIVER;
  }
  {
    cuuint64_t dims[5] = {32, (cuuint64_t)Wo, (cuuint64_t)Ho, 1, (cuuint64_t)B};
    cuuint64_t st[4] = {64, (cuuint64_t)Wo * 64, (cuuint64_t)Ho * Wo * 64, (cuuint64_t)Ho * Wo * 64};
    cuuint32_t box[5] = {32, pool ? 8u : 16u, pool ? 4u : 8u, 1, 1};
    const int rc = encode_map5(&a.mapOut, out, dims, st, box, 32);
    if (rc != VAD_OK) return rc;
    a.tma_store = (pool && pool_direct_setting()) ? 0 : 1;
  }
  const int grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
  return launch_conv_first(pool ? VAD_EPI_POOL : VAD_EPI_STORE, a, grid, stream);
}

int vad_first_conv_pool(const float* x, const void* weight_pf, const float* bias, float slope, int B, int H, int W,
                        void* out, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !weight_pf || !bias || !out || B <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  if ((H | W) & 1) return VAD_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(weight_pf) % 16 != 0 || reinterpret_cast<uintptr_t>(out) % 16 != 0) return VAD_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(x) % 16 != 0 || W % 4 != 0) return VAD_ERR_SHAPE;
  ensure_trap_slot();
  ConvArgs a;
  std::memset(&a, 0, sizeof(a));
  const int Hp = H / 2, Wp = W / 2;
  a.pair = 1;
  a.n_tiles = 1;
  a.tiles_w = (Wp + 15) / 16;  // tiles of 8 x 16 POOLED pixels
  a.tiles_h = (Hp + 7) / 8;
  a.tiles_b = B;
  const long long tiles = static_cast<long long>(a.tiles_w) * a.tiles_h * B;
  if (tiles > 0x7fffffffLL) return VAD_ERR_SHAPE;
  a.total_tiles = static_cast<int>(tiles);
  a.B = B; a.H = H; a.W = W;
  a.bias = bias;
  a.slope = slope;
  a.x = x;
  a.w_first = weight_pf;
  a.out = out;
  a.cout = 32;
  a.dbg = env_int("VAD_DBG", 0);
  a.pdl = pdl_all_setting() ? 1 : 0;
  {
    // fp32 NCHW input as a 4-D map {W, H, 3, B}; box = 40 x 18 x 3 patch starting at column 32*tw - 4 (16-byte aligned
    // start; columns 3..36 are used), zero fill outside the frame = conv padding
    EncodeTiledFn fn = encode_fn();
    if (!fn) return VAD_ERR_DRIVER;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
    cuuint64_t st[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
    cuuint32_t box[4] = {40, 18, 3, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (fn(&a.mapA0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, st, box, es,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return VAD_ERR_DRIVER;
  }
  {
    // pooled output bf16 NHWC [B][Hp][Wp][32] as a 5-D map {32, Wp, Hp, 1, B}; one store = one 8 x 16 pixel tile
    cuuint64_t dims[5] = {32, (cuuint64_t)Wp, (cuuint64_t)Hp, 1, (cuuint64_t)B};
    cuuint64_t st[4] = {64, (cuuint64_t)Wp * 64, (cuuint64_t)Hp * Wp * 64, (cuuint64_t)Hp * Wp * 64};
    cuuint32_t box[5] = {32, 16, 8, 1, 1};
    const int rc = encode_map5(&a.mapOut, out, dims, st, box, 32);
    if (rc != VAD_OK) return rc;
    a.tma_store = 1;
  }
  const int grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
  return launch_conv_first_pool(a, grid, stream);
}

int vad_enc1_fused(const float* x, const void* w_first, const float* bias1, const void* w_pair, const float* bias2,
                   float slope, int B, int H, int W, void* out, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !w_first || !bias1 || !w_pair || !bias2 || !out || B <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  if (H % 16 != 0 || W % 16 != 0) return VAD_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(w_first) % 16 != 0 || reinterpret_cast<uintptr_t>(out) % 16 != 0 ||
      reinterpret_cast<uintptr_t>(x) % 16 != 0)
    return VAD_ERR_ARG;
  ensure_trap_slot();
  ConvArgs a;
  std::memset(&a, 0, sizeof(a));
  a.pair = 1;
  a.n_tiles = 1;
  a.tiles_w = W / 16;  // tiles of 16 x 16 output pixels of the second conv (8 x 8 pooled pixels)
  a.tiles_h = H / 16;
  a.tiles_b = B;
  const long long tiles = static_cast<long long>(a.tiles_w) * a.tiles_h * B;
  if (tiles > 0x7fffffffLL) return VAD_ERR_SHAPE;
  a.total_tiles = static_cast<int>(tiles);
  a.B = B; a.H = H; a.W = W;
  a.bias = bias1;
  a.bias2 = bias2;
  a.slope = slope;
  a.x = x;
  a.w_first = w_first;
  a.out = out;
  a.cout = 32;
  a.pair_fold = 1;
  a.timeline = g_timeline;
  a.dbg = env_int("VAD_DBG", 0);
  a.pdl = pdl_all_setting() ? 1 : 0;
  {
    // fp32 NCHW input as a 4-D map {W, H, 3, B}; box = 28 x 20 x 3 patch from (16*th - 2, 16*tw - 4): the 16 x 16 tile
    // plus the halos of both convolutions (24 columns used; 28 keeps the converter's reads bank-conflict free), 16-byte
    // aligned start, zero fill outside the frame = the first conv's padding
    EncodeTiledFn fn = encode_fn();
    if (!fn) return VAD_ERR_DRIVER;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
    cuuint64_t st[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
    cuuint32_t box[4] = {28, 20, 3, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (fn(&a.mapA0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, st, box, es,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return VAD_ERR_DRIVER;
  }
  const int rc = encode_weight_map(&a.mapB, w_pair, 9 * 64, 64, 64, 64);  // nine [64 n][64 k] slabs of the pair kernel
  if (rc != VAD_OK) return rc;
  const int grid = a.total_tiles < sm_count() ? a.total_tiles : sm_count();
  return launch_enc1_fused(a, grid, stream);
}

int vad_score_finalize(const float* partials, int frames, int tiles_per_frame, int H, int W, float* score,
                       float* minmax, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!partials || !score || frames <= 0 || tiles_per_frame <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  const float inv = 1.0f / (3.0f * static_cast<float>(H) * static_cast<float>(W));
  score_finalize_kernel<<<(frames + 3) / 4, 128, 0, stream>>>(partials, frames, tiles_per_frame, inv, score, minmax);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

static int score_blocks_per_frame(int H, int W) {
  const long long plane = static_cast<long long>(H) * W;
  long long bpf = plane / 8192;
  if (bpf < 1) bpf = 1;
  if (bpf > 4096) bpf = 4096;
  return static_cast<int>(bpf);
}

size_t vad_score_scratch_bytes(int frames, int H, int W) {
  if (frames <= 0 || H <= 0 || W <= 0) return 0;
  return static_cast<size_t>(frames) * score_blocks_per_frame(H, W) * 4 * sizeof(float);
}

int vad_score(const float* x, const float* recon, int frames, int H, int W, float* score, float* minmax, float* heat,
              void* scratch, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !recon || !score || !scratch || frames <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  const long long plane = static_cast<long long>(H) * W;
  if (plane % 4 != 0) return VAD_ERR_SHAPE;
  const int bpf = score_blocks_per_frame(H, W);
  long long chunk = (plane + bpf - 1) / bpf;
  chunk = (chunk + 3) / 4 * 4;
  const long long blocks = static_cast<long long>(frames) * bpf;
  if (blocks > 0x7fffffffLL) return VAD_ERR_SHAPE;
  score_partial_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, recon, plane, bpf, chunk, heat,
                                                                         reinterpret_cast<float*>(scratch));
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  return vad_score_finalize(reinterpret_cast<const float*>(scratch), frames, bpf, H, W, score, minmax, stream_);
}

int vad_nhwc_bf16_to_nchw_f32(const void* src, int N, int H, int W, int C, float* dst, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || !dst || N <= 0 || H <= 0 || W <= 0 || C <= 0) return VAD_ERR_ARG;
  const long long total = static_cast<long long>(N) * C * H * W;
  nhwc_to_nchw_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), total, H * W, C, dst);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int vad_nchw_f32_to_nhwc_bf16(const float* src, int N, int C, int H, int W, void* dst, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || !dst || N <= 0 || H <= 0 || W <= 0 || C <= 0) return VAD_ERR_ARG;
  const long long total = static_cast<long long>(N) * C * H * W;
  nchw_to_nhwc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      src, total, H * W, C, reinterpret_cast<__nv_bfloat16*>(dst));
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int vad_heatmap_u8(const float* heat, const float* minmax, int frames, int H, int W, uint8_t* out,
                   vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!heat || !minmax || !out || frames <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  const long long plane = static_cast<long long>(H) * W;
  const long long total = plane * frames;
  heatmap_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(heat, minmax, plane, total, out);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int vad_u8_hwc_to_f32_nchw(const uint8_t* src, int frames, int H, int W, float* dst, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || !dst || frames <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  const long long plane = static_cast<long long>(H) * W;
  if (plane % 4 != 0 || reinterpret_cast<uintptr_t>(src) % 4 != 0 || reinterpret_cast<uintptr_t>(dst) % 16 != 0)
    return VAD_ERR_SHAPE;
  const long long quads = plane / 4 * frames;
  u8_hwc_to_f32_nchw_kernel<<<static_cast<unsigned>((quads + 255) / 256), 256, 0, stream>>>(src, plane, quads, dst);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int vad_f32_nchw_to_u8_hwc(const float* src, int frames, int H, int W, uint8_t* dst, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!src || !dst || frames <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  const long long plane = static_cast<long long>(H) * W;
  const long long total = plane * frames;
  f32_nchw_to_u8_hwc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(src, plane, total, dst);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int vad_heatmap_jet_rgb(const float* heat, const float* minmax, int frames, int H, int W, uint8_t* out,
                        vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!heat || !minmax || !out || frames <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  const long long plane = static_cast<long long>(H) * W;
  const long long total = plane * frames;
  heatmap_jet_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(heat, minmax, plane, total, out);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int vad_compose_panel(const float* x, const float* recon, const float* heat, const float* minmax, int frames, int H,
                      int W, uint8_t* out, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !recon || !heat || !minmax || !out || frames <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  const long long plane = static_cast<long long>(H) * W;
  const long long total = plane * frames;
  compose_panel_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(x, recon, heat, minmax, W, plane,
                                                                                      total, out);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

size_t vad_ssim_scratch_bytes(int frames, int H, int W) {
  if (frames <= 0 || H <= 0 || W <= 0) return 0;
  const size_t tiles = static_cast<size_t>((W + kSsimTile - 1) / kSsimTile) * ((H + kSsimTile - 1) / kSsimTile);
  return tiles * 3 * static_cast<size_t>(frames) * sizeof(float);
}

int vad_ssim_loss(const float* pred, const float* target, int frames, int H, int W, float* loss, float* ssim_map,
                  void* scratch, vad_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!pred || !target || !loss || !scratch || frames <= 0 || H <= 0 || W <= 0) return VAD_ERR_ARG;
  static std::atomic<bool> window_ready[kMaxDevices];  // __constant__ memory is per device
  const int dev_index = current_device_index();
  if (!window_ready[dev_index].load(std::memory_order_acquire)) {  // the reference's window: normalised exp(-x^2 / (2 * 1.5^2)), x = -5..5 (utils/losses.py:38-41)
    float g[11], sum = 0.f;
    for (int i = 0; i < 11; ++i) {
      const float x = static_cast<float>(i - 5);
      g[i] = std::exp(-(x * x) / (2.f * 1.5f * 1.5f));
      sum += g[i];
    }
    for (int i = 0; i < 11; ++i) g[i] /= sum;
    cudaError_t e = cudaMemcpyToSymbol(c_ssim_gauss, g, sizeof(g));
    if (e != cudaSuccess) return static_cast<int>(e);
    window_ready[dev_index].store(true, std::memory_order_release);
  }
  const int tiles_x = (W + kSsimTile - 1) / kSsimTile, tiles_y = (H + kSsimTile - 1) / kSsimTile;
  const long long blocks = static_cast<long long>(tiles_x) * tiles_y * 3 * frames;
  if (blocks > 0x7fffffffLL) return VAD_ERR_SHAPE;
  ssim_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(pred, target, H, W, tiles_x, tiles_y, ssim_map,
                                                               reinterpret_cast<float*>(scratch));
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  const float inv = 1.0f / (3.0f * static_cast<float>(H) * static_cast<float>(W));
  ssim_finalize_kernel<<<(frames + 3) / 4, 128, 0, stream>>>(reinterpret_cast<const float*>(scratch), frames,
                                                            tiles_x * tiles_y * 3, inv, loss);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"



namespace vad {
int lstm_persistent_tiles(int B, int H, int W, int c0, int c1, int n_total) {
  const int n_tiles = n_total / 128 > 0 ? n_total / 128 : 1;
  static const int seq_env = env_int("VAD_LSTM_SEQ", 1);
  const int mode = g_lstm_mode_override >= 0 ? g_lstm_mode_override : seq_env;
  if (mode == 1 && c0 % 64 == 0 && c1 % 64 == 0 && H >= 8 && W >= 8) {  // patch kernel (build_lstm_layer)
    const bool geo1 = H == 8 && W == 8;
    const int tiles_w = (W + 7) >> 3;
    const int tiles_h = geo1 ? 1 : (H + 15) >> 4;
    const int tiles_b = geo1 ? (B + 1) >> 1 : B;
    return tiles_w * tiles_h * tiles_b * n_tiles;
  }
  return pick_tile_geometry(B, H, W, false).m_tiles() * n_tiles;
}
}  // namespace vad
