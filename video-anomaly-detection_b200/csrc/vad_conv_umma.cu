// Implicit-GEMM convolution / transposed convolution / ConvLSTM gate kernels for sm_100a.
//
//   D[128 pixels, BN] (fp32, TMEM)  +=  A[128 pixels, CK channels of one tap] (bf16, smem via TMA, 5-D tiled map
//                                        over [B][T][H][W][C] with out-of-bounds zero fill = conv padding)
//                                     x  B[BN, CK] (bf16, smem via TMA, weights K-major)
//
// Seven single-layer kernels share one set of epilogues; three more fuse neighbouring layers (section markers below, in
// file order):
//   epilogue              — STORE / POOL / CONVT / LSTM / TANH_SCORE / CONVT_TANH_SCORE on one finished accumulator
//                           tile; EpiLane (per-lane constants), TileIter (division-free tile walking), epilogue_loop.
//   conv_umma_kernel      — "streaming": one (A tile, B tile) pair per tap and channel chunk through an mbarrier ring.
//                           1-tap GEMMs (ConvT / 1x1), ConvLSTM steps launched one by one, shapes nothing else takes.
//   conv_halo_kernel      — 3x3 conv, Cin in {32, 64}: nine weight slabs resident in smem, ONE input patch (1-pixel
//                           halo) per tile read by all nine taps through shifted shared-memory descriptors.
//   conv_kx_kernel        — the same with the three horizontal taps folded into N (the 3-channel score layer).
//   conv_hs_kernel        — wide 3x3 layers: patches for pairs of tiles (M = 256) + streamed weight tiles.
//   convlstm_seq_kernel   — all T steps of a ConvLSTM layer in one launch (cell state in registers), streamed A tiles.
//   convlstm_patch_kernel — the same with A patches, a separate weight ring / producer and an 8-warp gate epilogue.
//   convlstm2_patch_kernel — both ConvLSTM layers in one launch as a wavefront (layer 2 one step behind layer 1).
//   conv_first_kernel     — 3 -> 32 first conv from the fp32 NCHW input: TMA patch -> converter warps (im2col, bf16,
//                           they also issue the MMAs) -> epilogue.
//   convt2_score_kernel   — video decoder tail: ConvT(64->32)+ReLU -> ConvT(32->3)+tanh -> score; the two GEMMs are
//                           chained through shared memory (epilogue 1 writes the second GEMM's A operands).
//   convt_conv_score_kernel — image decoder tail: ConvT(32->32)+ReLU -> conv3x3(32->3)+tanh -> score; the transposed
//                           conv's output lives as a 16 x 32 pixel patch in shared memory.
// All are persistent and warp-specialised (TMA producer(s), tcgen05.mma issuer(s), TMEM allocator, 1-6 epilogue groups
// of four warps: TMEM -> registers -> bias/activation/pool/pixel-shuffle/gates/score -> registers or swizzled smem ->
// global / TMA store); 2-4 accumulator stages in TMEM let epilogues overlap the next tiles' MMAs.  Host-side launchers
// and the (kernel, tile shape) dispatch tables are at the end of the file; layer -> kernel selection is in vad_api.cu.
//
// Replaces (reference file:line): nn.Conv2d 3x3 models/autoencoder.py:39-76,107-137; nn.ConvTranspose2d k2 s2
// models/autoencoder.py:104-134 and models/video_autoencoder.py:244-259; ConvLSTMCell.forward
// models/video_autoencoder.py:54-85; the error reduction models/autoencoder.py:214-221 and
// models/video_autoencoder.py:371-384 (fused onto the last decoder layer).
#include <atomic>

#include "vad_internal.h"
#include "vad_ptx.cuh"

namespace vad {

constexpr int kEpiWarp0 = 4;  // warps 0..3: TMA producer, MMA issuer, TMEM allocator, spare
constexpr int kTileM = 128;
constexpr int kMaxAccStages = 8;
constexpr int kStagingBuf = 16384;  // one staged output chunk: 128 rows x 128 B
constexpr int kMaxBias = 512;         // bias slab of the kernels whose whole N fits one or a few tiles
constexpr int kMaxBiasStream = 1024;  // streaming kernel (ConvT with 4*Cout columns, ConvLSTM gates with 4*hidden columns)
constexpr int kSmemBudget = 227 * 1024 - 4096;  // dynamic smem we allow ourselves (static smem + slack kept free)

__host__ __device__ constexpr bool epi_uses_staging(int epi) {
  return epi == VAD_EPI_STORE || epi == VAD_EPI_POOL || epi == VAD_EPI_CONVT || epi == VAD_EPI_LSTM;
}
// Epilogue warps come in groups of four (one warp per TMEM lane quarter); group g takes this CTA's tiles g, g+G, ...
// and owns TMEM accumulator stage g.  The narrowest tiles (N <= 32: full-resolution layers, thousands of tiny tiles
// per SM; N <= 64) are bound by epilogue latency and get four groups, 128-wide tiles two, and 256-wide tiles (MMA-bound) one
// group plus the smem for a deeper operand ring.
// (transposed convolutions have K <= 256: a couple of MMAs per tile against a 128-column epilogue -> four groups too)
__host__ __device__ constexpr int epi_groups(int bn, int epi = -1) {
  return bn >= 256 ? 1 : ((bn <= 64 || epi == VAD_EPI_CONVT) ? 4 : 2);
}
__host__ __device__ constexpr int acc_stages_for(int groups) { return groups < 2 ? 2 : groups; }
__host__ __device__ constexpr int block_threads(int bn, int epi = -1) { return 128 + 128 * epi_groups(bn, epi); }
// staged-output buffers per epilogue group
__host__ __device__ constexpr int staging_bufs(int bn, int epi) {
  // the narrow N=32 tiles and the ConvLSTM epilogue (8 KB chunks) double-buffer, everything else has one 16 KB buffer
  return !epi_uses_staging(epi) ? 0 : (epi == VAD_EPI_LSTM ? 2 : ((bn >= 128 || bn == 64) ? 1 : 2));
}
// one staged chunk: 128 rows x (64 ch = 128 B | 32 ch = 64 B)
__host__ __device__ constexpr int staging_buf_bytes(int bn, int epi) {
  return (bn == 32 || epi == VAD_EPI_LSTM) ? kStagingBuf / 2 : kStagingBuf;
}
__host__ __device__ constexpr int staging_group_bytes(int bn, int epi) {
  return staging_bufs(bn, epi) * staging_buf_bytes(bn, epi);
}
__host__ __device__ constexpr int staging_bytes(int bn, int epi, int groups = 0) {
  return (groups ? groups : epi_groups(bn, epi)) * staging_group_bytes(bn, epi);
}
__host__ __device__ constexpr uint32_t tmem_cols_for(int bn, int groups = 0, int epi = -1) {
  const int c = acc_stages_for(groups ? groups : epi_groups(bn, epi)) * bn;
  return (c <= 32) ? 32u : (c <= 64) ? 64u : (c <= 128) ? 128u : (c <= 256) ? 256u : 512u;
}

// ---- kx-merged halo mode ("kx kernel") ---------------------------------------------------------------------------
// A tcgen05.mma with M=128, K=16 costs max(N/2, 32 + N/4) cycles (tools/umma_bench.cu): below N=128 it is bound by
// streaming the 4 KB A slab from shared memory, not by math.  For the narrow 3x3 layers (Cout = 32, and the 3-channel
// last conv) the three horizontal taps are therefore folded into the N extent: D[pixel q][kx][co] = sum over (ky, ci)
// of in[q + ky row shift][ci] * w[co][ky][kx][ci] — three row-shifted MMAs per K step instead of nine — and the
// epilogue forms out[x] = D[x-1][0] + D[x][1] + D[x+1][2] with two lane shuffles.  One tile = 16 rows x 8 patch
// columns (x0-1 .. x0+6) -> 6 valid output columns.
constexpr int kKxValid = 6;  // valid output columns of a kx tile (8 accumulator columns minus the two halo columns)
__host__ __device__ constexpr bool is_score_epi(int epi) {
  return epi == VAD_EPI_TANH_SCORE || epi == VAD_EPI_CONVT_TANH_SCORE;
}
__host__ __device__ constexpr int kx_mma_n(int bn, int epi) { return is_score_epi(epi) ? 16 : 3 * bn; }
__host__ __device__ constexpr int kx_acc_stride(int bn, int epi) {
  return kx_mma_n(bn, epi) <= 16 ? 16 : (kx_mma_n(bn, epi) <= 128 ? 128 : 256);
}
// epilogue groups: the score epilogue is light on registers and bound by latency -> six groups (896 threads)
__host__ __device__ constexpr int kx_groups(int bn, int epi) {
  return is_score_epi(epi) ? 6 : (kx_acc_stride(bn, epi) * 4 <= 512 ? 4 : 2);
}
__host__ __device__ constexpr uint32_t kx_tmem_cols(int bn, int epi) {
  const int c = kx_groups(bn, epi) * kx_acc_stride(bn, epi);
  return (c <= 32) ? 32u : (c <= 64) ? 64u : (c <= 128) ? 128u : (c <= 256) ? 256u : 512u;
}

template <int CK, int BN, int EPI>
struct Cfg {
  static constexpr int kRowBytes = CK * 2;            // one swizzle span per row: 128 B or 64 B
  static constexpr int kABytes = kTileM * kRowBytes;  // 16 KB / 8 KB
  static constexpr int kBBytes = BN * kRowBytes;      // multiple of 1024 for BN >= 16
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = staging_bytes(BN, EPI);
  static constexpr int kStagesRaw = (kSmemBudget - 1024 - 2048 - kStagingBytes) / kStageBytes;  // (-2048: bias slab)
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024;  // +1024 alignment slack
  static constexpr uint32_t kLayout = (CK == 64) ? 2u : 4u;                        // SWIZZLE_128B : SWIZZLE_64B
  static constexpr uint32_t kSBO = 8 * kRowBytes;                                  // bytes between 8-row groups
  static constexpr uint32_t kTmemCols = tmem_cols_for(BN, 0, EPI);
  static_assert(CK == 64 || CK == 32, "K chunk must be one 128B or 64B swizzle span");
  static_assert(kBBytes % 1024 == 0, "B stage must keep 1024B alignment");
  static_assert(kStages >= 2, "pipeline too shallow");
};

struct TileCoord {
  int n0;          // first GEMM column of this tile
  int b0, h0, w0;  // first frame / row / column of this tile
  int m_tile;
};

// LeakyReLU family with 0 <= slope <= 1 (0 = ReLU, 1 = identity): act(v) = max(v, slope*v)
// Bring-up timeline: CTA 0 stamps clock64 at role events of its first 64 tiles (role 0 producer, 1 MMA, 2/3 epilogue
// group 0/1 leader).  Compiled in always; one predictable branch per event when disabled.
// The stamps cost real time in the single-warp role loops (a constant load, a predicate chain and a clock read per
// event even when disabled), so they only exist in -DVAD_TIMELINE builds (`python build.py --timeline`).
__device__ __forceinline__ void tl_stamp(const ConvArgs& a, int role, int n, int ev) {
#ifdef VAD_TIMELINE
  if (a.timeline != nullptr && blockIdx.x == 0 && n < 64) a.timeline[(role * 64 + n) * 16 + ev] = clock64();
#endif
}

// Tile walker for the persistent loops: tile index = ((tb*tiles_h + th)*tiles_w + tw)*n_tiles + n_t advances by a
// fixed step; the step is decomposed once (integer divisions) and then added component-wise with carries, so the
// per-tile cost is a handful of adds instead of three division sequences per role.
struct TileIter {
  int tile, step;
  int m, n_t, tw, th, tb;  // tw counts tile PAIRS when a.pair == 2
  int dm, dn, dw, dh, db;
  __device__ __forceinline__ TileIter(const ConvArgs& a, int first, int step_) : tile(first), step(step_) {
    const int tiles_wp = a.tiles_w / a.pair;
    int v = first;
    m = v % a.pair; v /= a.pair;
    n_t = v % a.n_tiles; v /= a.n_tiles;
    tw = v % tiles_wp; v /= tiles_wp;
    th = v % a.tiles_h; tb = v / a.tiles_h;
    v = step_;
    dm = v % a.pair; v /= a.pair;
    dn = v % a.n_tiles; v /= a.n_tiles;
    dw = v % tiles_wp; v /= tiles_wp;
    dh = v % a.tiles_h; db = v / a.tiles_h;
  }
  __device__ __forceinline__ void next(const ConvArgs& a) {
    tile += step;
    m += dm;
    int c = m >= a.pair;
    m -= c ? a.pair : 0;
    n_t += dn + c;
    c = n_t >= a.n_tiles;
    n_t -= c ? a.n_tiles : 0;
    tw += dw + c;
    const int tiles_wp = a.pair == 2 ? a.tiles_w >> 1 : a.tiles_w;
    c = tw >= tiles_wp;
    tw -= c ? tiles_wp : 0;
    th += dh + c;
    c = th >= a.tiles_h;
    th -= c ? a.tiles_h : 0;
    tb += db + c;
  }
  __device__ __forceinline__ TileCoord coord(const ConvArgs& a, int bn) const {
    TileCoord t;
    const int twi = tw * a.pair + m;
    t.n0 = n_t * bn;
    t.w0 = twi * a.w_step;
    t.h0 = th << a.lgTH;
    t.b0 = tb << a.lgTN;
    t.m_tile = (tb * a.tiles_h + th) * a.tiles_w + twi;
    return t;
  }
};

__device__ __forceinline__ float act_fn(float v, float slope) { return fmaxf(v, v * slope); }
__device__ __forceinline__ float sigmoid_fn(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }
// tanh(v) = 1 - 2 / (1 + e^{2v}) on the fast exp / reciprocal units: absolute error ~1e-7 (fp32 rounding of the
// subtraction), saturates correctly to +-1 for large |v|
__device__ __forceinline__ float tanh_fn(float v) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * v)); }

// Byte offset of 16-byte chunk `c16` of row `row` in a staged tile whose rows are one swizzle span wide.
__device__ __forceinline__ uint32_t staged_off(int row, int c16, int out_chunk) {
  return out_chunk == 64 ? static_cast<uint32_t>(row * 128 + ((c16 ^ (row & 7)) << 4))
                         : static_cast<uint32_t>(row * 64 + ((c16 ^ ((row >> 1) & 3)) << 4));
}

// ---------------------------------------------------------------------------------------------------- epilogue
// Runs on the four epilogue warps for one finished accumulator tile.  `stg_i` counts staged chunks (ring index).
// Tile-invariant facts about one epilogue lane (= one accumulator row), computed once before the tile loop so that
// the per-tile code does not redo the index arithmetic.
struct EpiLane {
  int ww, hh, bb;  // pixel of this accumulator row inside the tile (column, row, frame)
  int srow;        // row of this pixel in a staged [frame][h][w] output tile
  int mx, my;      // lane xor masks that reach the horizontal / vertical neighbour pixel (2x2 pooling)
  bool row_ok;     // the row belongs to the tile at all
};
__device__ __forceinline__ EpiLane make_epi_lane(const ConvArgs& a, int q, int lane) {
  const int r = q * 32 + lane;
  const int TW = 1 << a.lgTW, TH = 1 << a.lgTH;
  EpiLane L;
  if (a.row_perm == 2) {
    // ConvLSTM patch kernel on 8x8 frames, two frames per tile: row = ((y*2 + frame)*8 + x) — the order in which one
    // patch [y][frame][x] serves all nine taps with a constant 8-row-group stride
    L.ww = r & 7;
    L.bb = (r >> 3) & 1;
    L.hh = r >> 4;
    L.mx = 1;
    L.my = 16;
  } else if (a.row_perm) {
    // first conv (8 x 16 tiles): A row = (hh >> 1)*32 + (ww & 3)*8 + (hh & 1)*4 + (ww >> 2) — the order in which the
    // im2col converter can write its rows without shared-memory bank conflicts
    L.hh = ((r >> 5) << 1) | ((r >> 2) & 1);
    L.ww = ((r & 3) << 2) | ((r >> 3) & 3);
    L.bb = 0;
    L.mx = 8;
    L.my = 4;
  } else {
    L.ww = r & (TW - 1);
    L.hh = (r >> a.lgTW) & (TH - 1);
    L.bb = r >> (a.lgTW + a.lgTH);
    L.mx = 1;
    L.my = TW;
  }
  L.srow = ((L.bb << a.lgTH) + L.hh) * TW + L.ww;
  L.row_ok = (r < (1 << (a.lgTW + a.lgTH + a.lgTN))) && (L.ww < a.tw_valid);
  return L;
}

// 32 accumulator columns [lc, lc+32) of this thread's row.  KX == 3: the three kx column groups (BN columns apart) are
// combined across neighbouring rows: out[ww] = D[ww][kx=0] + D[ww+1][kx=1] + D[ww+2][kx=2] (rows = lanes).
template <int BN, int KX>
__device__ __forceinline__ void load_acc32(uint32_t tacc, int lc, uint32_t (&v)[32]) {
  if constexpr (KX == 1) {
    tmem_ld_x32(tacc + lc, v);
    tmem_ld_wait();
  } else {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t c0[16], c1[16], c2[16];
      tmem_ld_x16(tacc + lc + half * 16, c0);
      tmem_ld_x16(tacc + BN + lc + half * 16, c1);
      tmem_ld_x16(tacc + 2 * BN + lc + half * 16, c2);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float s1 = __shfl_down_sync(0xffffffffu, __uint_as_float(c1[j]), 1);
        const float s2 = __shfl_down_sync(0xffffffffu, __uint_as_float(c2[j]), 2);
        v[half * 16 + j] = __float_as_uint((__uint_as_float(c0[j]) + s1) + s2);
      }
    }
  }
}

template <int BN, int EPI, int KX = 1>
__device__ __forceinline__ void epilogue_tile(const ConvArgs& a, const TileCoord& t, const EpiLane& L, uint32_t tacc,
                                              int q, int lane, uint8_t* stg, const float* s_bias,
                                              float (*red_smem)[3], int& stg_i,
                                              uint32_t bar_id, const float* xpre, uint32_t acc_empty, int tl_role = -1,
                                              int tl_n = 0) {
  auto stamp = [&](int ev) {
#ifdef VAD_TIMELINE
    if (tl_role >= 0) tl_stamp(a, tl_role, tl_n, ev);
#endif
  };
  // Hand the TMEM accumulator stage back to the MMA warp as soon as its last column is in registers: the rest of
  // the epilogue (math, staging, stores) then overlaps the next tile's MMAs into the same stage.
  auto release_acc = [&]() {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_a(acc_empty);
  };
  const int TW = 1 << a.lgTW;
  const int ww = L.ww, hh = L.hh, bb = L.bb;
  const int fb = t.b0 + bb, h = t.h0 + hh, w = t.w0 + ww;
  const bool valid = L.row_ok && (fb < a.B) && (h < a.H) && (w < a.W);
  const bool leader = (q == 0 && lane == 0);
  constexpr int kBufs = staging_bufs(BN, EPI);

  if constexpr (EPI == VAD_EPI_STORE || EPI == VAD_EPI_POOL || EPI == VAD_EPI_CONVT) {
    if (a.tma_store) {
      const int OC = a.out_chunk;                             // 64 or 32
      const int n_chunks = (OC == 64) ? BN / 64 : BN / 32;    // (compile-time divisions)
      stamp(9);
#pragma unroll 1
      for (int oc = 0; oc < n_chunks; ++oc) {
        uint8_t* buf = stg + (kBufs > 1 ? (stg_i & 1) : 0) * staging_buf_bytes(BN, EPI);
        if (leader) {  // the TMA store that last read this buffer must have finished reading it
          if (kBufs > 1) bulk_wait_group_read<1>(); else bulk_wait_group_read<0>();
        }
        stamp(10);
        named_bar_sync(bar_id, 128);
        stamp(3);
#pragma unroll 1
        for (int sub = 0; sub < OC / 32; ++sub) {
          const int lc = oc * OC + sub * 32;
          uint32_t v[32];
          load_acc32<BN, KX>(tacc, lc, v);
          stamp(4);
          if (lc + 32 == BN) release_acc();
          if (a.dbg & 32) continue;  // ablation: no epilogue math / staging
          if constexpr (EPI == VAD_EPI_POOL) {
            // 2x2 max-pool as a reduce-scatter over the 4 lanes of a window: exchange halves with the horizontal
            // neighbour (lane^1), then quarters with the vertical neighbour (lane^TW); each lane ends up owning the
            // max of 8 channels and applies bias / activation / packing to those only.
            const bool up1 = (ww & 1) != 0, up2 = (hh & 1) != 0;
            float g[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float lo = __uint_as_float(v[j]), hi = __uint_as_float(v[j + 16]);
              const float recv = __shfl_xor_sync(0xffffffffu, up1 ? lo : hi, L.mx);
              g[j] = fmaxf(up1 ? hi : lo, recv);
            }
            float m[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float recv = __shfl_xor_sync(0xffffffffu, up2 ? g[j] : g[j + 8], L.my);
              m[j] = fmaxf(up2 ? g[j + 8] : g[j], recv);
            }
            const int part = (up1 ? 2 : 0) + (up2 ? 1 : 0);  // channels [8*part, 8*part+8) of this 32-column chunk
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + t.n0 + lc + part * 8);
            const float4 b0 = b4[0], b1 = b4[1];
            const int prow = (KX == 3) ? (hh >> 1) * (kKxValid / 2) + (ww >> 1)
                                       : ((bb << (a.lgTH - 1)) + (hh >> 1)) * (TW >> 1) + (ww >> 1);
            const uint4 val = make_uint4(
                pack_bf16x2(act_fn(m[0] + b0.x, a.slope), act_fn(m[1] + b0.y, a.slope)),
                pack_bf16x2(act_fn(m[2] + b0.z, a.slope), act_fn(m[3] + b0.w, a.slope)),
                pack_bf16x2(act_fn(m[4] + b1.x, a.slope), act_fn(m[5] + b1.y, a.slope)),
                pack_bf16x2(act_fn(m[6] + b1.z, a.slope), act_fn(m[7] + b1.w, a.slope)));
            if (KX == 1 || ww < kKxValid) sts128(buf + staged_off(prow, sub * 4 + part, OC), val);
          } else {
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + t.n0 + lc);
            uint32_t p[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = b4[j];
              p[2 * j] = pack_bf16x2(act_fn(__uint_as_float(v[4 * j]) + bv.x, a.slope),
                                     act_fn(__uint_as_float(v[4 * j + 1]) + bv.y, a.slope));
              p[2 * j + 1] = pack_bf16x2(act_fn(__uint_as_float(v[4 * j + 2]) + bv.z, a.slope),
                                         act_fn(__uint_as_float(v[4 * j + 3]) + bv.w, a.slope));
            }
            const int srow = (KX == 3) ? hh * kKxValid + ww : L.srow;
            if (KX == 1 || ww < kKxValid) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                sts128(buf + staged_off(srow, sub * 4 + j, OC), make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]));
            }
          }
        }
        stamp(5);
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
        stamp(6);
        named_bar_sync(bar_id, 128);
        stamp(7);
        if (leader && !(a.dbg & 64)) {  // (ablation bit 64: no TMA store)
          const int col = t.n0 + oc * OC;
          if constexpr (EPI == VAD_EPI_STORE) {
            tma_store_5d(&a.mapOut, buf, col, t.w0, t.h0, a.out_t, t.b0);
          } else if constexpr (EPI == VAD_EPI_POOL) {
            tma_store_5d(&a.mapOut, buf, col, t.w0 >> 1, t.h0 >> 1, 0, t.b0);
          } else {  // ConvT pixel shuffle: map dims {co, dj, w, di, b*H + h}
            const int quad = (col >= a.cout) + (col >= 2 * a.cout) + (col >= 3 * a.cout);
            if (a.convt_split)  // one map per output row parity, {co, dj, w, h, b}: rows past the frame are clipped
              tma_store_5d((quad >> 1) ? &a.mapOut2 : &a.mapOut, buf, col - quad * a.cout, quad & 1, t.w0, t.h0, t.b0);
            else
              tma_store_5d(&a.mapOut, buf, col - quad * a.cout, quad & 1, t.w0, quad >> 1, t.b0 * a.H + t.h0);
          }
          bulk_commit_group();
        }
        ++stg_i;
      }
    } else {
      // direct stores.  POOL: the default path — after the pooling butterfly the 32 lanes of a warp hold 32 different
      // 16-byte pieces of eight adjacent pooled pixels, i.e. 64-byte (or longer) contiguous runs: fully used sectors
      // straight from registers, no staging buffer, no group barrier, no proxy fence, no TMA store.  STORE / CONVT: the
      // fallback for tile shapes the output tensor map cannot express.
      __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(a.out);
      if (EPI == VAD_EPI_POOL && KX == 1 && a.pair_fold) {
        // pixel-pair folded layer: columns [0, BN/2) are the pair's left pixel, [BN/2, BN) its right pixel, so the
        // horizontal half of the 2x2 window is a max inside the lane; the vertical half is one exchange with the row
        // neighbour (reduce-scatter: each lane keeps 8 of every 16 channels).  a.W counts pairs = pooled columns.
        constexpr int CO = BN / 2;
        const bool up2 = (hh & 1) != 0;
        __nv_bfloat16* dst0 = outp + fb * a.out_fs + (static_cast<long long>(h >> 1) * a.W + w) * a.out_cp + (up2 ? 8 : 0);
#pragma unroll 1
        for (int c = 0; c < CO / 16; ++c) {
          uint32_t v0[16], v1[16];
          tmem_ld_x16(tacc + c * 16, v0);
          tmem_ld_x16(tacc + CO + c * 16, v1);
          tmem_ld_wait();
          if (c == CO / 16 - 1) release_acc();
          float g[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) g[j] = fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j]));
          float m[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float recv = __shfl_xor_sync(0xffffffffu, up2 ? g[j] : g[j + 8], L.my);
            m[j] = fmaxf(up2 ? g[j + 8] : g[j], recv);
          }
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c * 16 + (up2 ? 8 : 0));
          const float4 b0 = b4[0], b1 = b4[1];
          if (valid)
            *reinterpret_cast<uint4*>(dst0 + c * 16) = make_uint4(
                pack_bf16x2(act_fn(m[0] + b0.x, a.slope), act_fn(m[1] + b0.y, a.slope)),
                pack_bf16x2(act_fn(m[2] + b0.z, a.slope), act_fn(m[3] + b0.w, a.slope)),
                pack_bf16x2(act_fn(m[4] + b1.x, a.slope), act_fn(m[5] + b1.y, a.slope)),
                pack_bf16x2(act_fn(m[6] + b1.z, a.slope), act_fn(m[7] + b1.w, a.slope)));
        }
      } else if constexpr (EPI == VAD_EPI_POOL) {
        const bool up1 = (ww & 1) != 0, up2 = (hh & 1) != 0;
        const int part = (up1 ? 2 : 0) + (up2 ? 1 : 0);
        __nv_bfloat16* dst0 = outp + fb * a.out_fs +
                              (static_cast<long long>(h >> 1) * (a.W >> 1) + (w >> 1)) * a.out_cp + t.n0 + part * 8;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          load_acc32<BN, KX>(tacc, c * 32, v);
          if (c * 32 + 32 == BN) release_acc();
          float g[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float lo = __uint_as_float(v[j]), hi = __uint_as_float(v[j + 16]);
            const float recv = __shfl_xor_sync(0xffffffffu, up1 ? lo : hi, L.mx);
            g[j] = fmaxf(up1 ? hi : lo, recv);
          }
          float m[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float recv = __shfl_xor_sync(0xffffffffu, up2 ? g[j] : g[j + 8], L.my);
            m[j] = fmaxf(up2 ? g[j + 8] : g[j], recv);
          }
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + t.n0 + c * 32 + part * 8);
          const float4 b0 = b4[0], b1 = b4[1];
          if (valid)
            *reinterpret_cast<uint4*>(dst0 + c * 32) = make_uint4(
                pack_bf16x2(act_fn(m[0] + b0.x, a.slope), act_fn(m[1] + b0.y, a.slope)),
                pack_bf16x2(act_fn(m[2] + b0.z, a.slope), act_fn(m[3] + b0.w, a.slope)),
                pack_bf16x2(act_fn(m[4] + b1.x, a.slope), act_fn(m[5] + b1.y, a.slope)),
                pack_bf16x2(act_fn(m[6] + b1.z, a.slope), act_fn(m[7] + b1.w, a.slope)));
        }
      } else {
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        load_acc32<BN, KX>(tacc, c * 32, v);
        if (c * 32 + 32 == BN) release_acc();
        const int col = t.n0 + c * 32;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = act_fn(f[j] + s_bias[col + j], a.slope);
        uint32_t p[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) p[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
        if constexpr (EPI == VAD_EPI_STORE) {
          if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(outp + fb * a.out_fs +
                                                  (static_cast<long long>(h) * a.W + w) * a.out_cp + col);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
          }
        } else {  // VAD_EPI_CONVT: column = quad*cout + co, quad = di*2 + dj
          if (valid) {
            const int quad = (col >= a.cout) + (col >= 2 * a.cout) + (col >= 3 * a.cout);
            const int co = col - quad * a.cout;
            const int ho = 2 * h + (quad >> 1), wo = 2 * w + (quad & 1);
            uint4* dst = reinterpret_cast<uint4*>(outp + fb * a.out_fs +
                                                  (static_cast<long long>(ho) * (2 * a.W) + wo) * a.out_cp + co);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
          }
        }
      }
      }  // (STORE / CONVT direct path)
    }
  } else if constexpr (EPI == VAD_EPI_LSTM) {
    // columns of this tile: [gate g in i,f,g,o][32 hidden channels j0..j0+31]; weights/bias pre-permuted on host
    static_assert(EPI != VAD_EPI_LSTM || BN == 128, "LSTM tile is 4 gates x 32 channels");
    if (a.pdl == 2) pdl_wait();  // the cell state below was written by the previous step's launch (cheap once satisfied)
    const int hid = a.cout;
    const int j0 = (t.n0 >> 7) * 32;
    const long long pix = (static_cast<long long>(fb) * a.H + h) * a.W + w;
    float* cptr = a.c_state + pix * hid + j0;
    uint8_t* buf = stg + (stg_i & 1) * staging_buf_bytes(BN, EPI);
    if (a.tma_store) {
      if (leader) bulk_wait_group_read<1>();
      named_bar_sync(bar_id, 128);
    }
    __nv_bfloat16* hptr = reinterpret_cast<__nv_bfloat16*>(a.out) + fb * a.out_fs +
                          (static_cast<long long>(h) * a.W + w) * a.out_cp + j0;
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
      uint32_t gi[8], gf[8], gg[8], go[8];
      tmem_ld_x8(tacc + 0 + s * 8, gi);
      tmem_ld_x8(tacc + 32 + s * 8, gf);
      tmem_ld_x8(tacc + 64 + s * 8, gg);
      tmem_ld_x8(tacc + 96 + s * 8, go);
      tmem_ld_wait();
      if (s == 3) release_acc();
      float cprev[8];
      if (valid && !a.lstm_first) {
        const float4 c0 = *reinterpret_cast<const float4*>(cptr + s * 8);
        const float4 c1 = *reinterpret_cast<const float4*>(cptr + s * 8 + 4);
        cprev[0] = c0.x; cprev[1] = c0.y; cprev[2] = c0.z; cprev[3] = c0.w;
        cprev[4] = c1.x; cprev[5] = c1.y; cprev[6] = c1.z; cprev[7] = c1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) cprev[e] = 0.f;
      }
      float cn[8], hn[8];
      const float* bp = s_bias + t.n0 + s * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xi = __uint_as_float(gi[e]) + bp[e];
        const float xf = __uint_as_float(gf[e]) + bp[32 + e];
        const float xg = __uint_as_float(gg[e]) + bp[64 + e];
        const float xo = __uint_as_float(go[e]) + bp[96 + e];
        cn[e] = sigmoid_fn(xf) * cprev[e] + sigmoid_fn(xi) * tanh_fn(xg);
        hn[e] = sigmoid_fn(xo) * tanh_fn(cn[e]);
      }
      const uint4 hv = make_uint4(pack_bf16x2(hn[0], hn[1]), pack_bf16x2(hn[2], hn[3]), pack_bf16x2(hn[4], hn[5]),
                                  pack_bf16x2(hn[6], hn[7]));
      if (valid) {
        *reinterpret_cast<float4*>(cptr + s * 8) = make_float4(cn[0], cn[1], cn[2], cn[3]);
        *reinterpret_cast<float4*>(cptr + s * 8 + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
        if (!a.tma_store) *reinterpret_cast<uint4*>(hptr + s * 8) = hv;
      }
      if (a.tma_store) sts128(buf + staged_off(L.srow, s, 32), hv);
    }
    if (a.tma_store) {
      fence_proxy_async_smem();
      named_bar_sync(bar_id, 128);
      if (leader) {
        tma_store_5d(&a.mapOut, buf, j0, t.w0, t.h0, a.out_t, t.b0);
        bulk_commit_group();
      }
      ++stg_i;
    }
  } else {
    // ------------------------------------------------ fused tanh + reconstruction-error reduction (N tile = 16)
    static_assert((EPI != VAD_EPI_TANH_SCORE && EPI != VAD_EPI_CONVT_TANH_SCORE) || BN == 16, "score tile is 16 wide");
    uint32_t v[16];
    tmem_ld_x16(tacc, v);
    tmem_ld_wait();
    release_acc();
    if constexpr (KX == 3) {  // columns [kx][co]: fold the three horizontal taps from the neighbouring rows
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float s1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[3 + ch]), 1);
        const float s2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[6 + ch]), 2);
        v[ch] = __float_as_uint((__uint_as_float(v[ch]) + s1) + s2);
      }
    }
    float ssum = 0.f, smin = INFINITY, smax = -INFINITY;
    if constexpr (EPI == VAD_EPI_TANH_SCORE) {
      if (valid) {
        const long long plane = static_cast<long long>(a.H) * a.W;
        const long long off = static_cast<long long>(fb) * 3 * plane + static_cast<long long>(h) * a.W + w;
        float sq = 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float rec = tanh_fn(__uint_as_float(v[ch]) + s_bias[ch]);
          const float d = xpre[ch] - rec;
          sq += d * d;
          if (a.recon) a.recon[off + ch * plane] = rec;
        }
        if (a.heat) a.heat[static_cast<long long>(fb) * plane + static_cast<long long>(h) * a.W + w] = sq * (1.f / 3.f);
        ssum = sq; smin = sq; smax = sq;
      }
    } else {
      if (valid) {
        const int Ho = 2 * a.H, Wo = 2 * a.W;
        const long long plane = static_cast<long long>(Ho) * Wo;
#pragma unroll
        for (int di = 0; di < 2; ++di) {
          const long long off = static_cast<long long>(fb) * 3 * plane + static_cast<long long>(2 * h + di) * Wo + 2 * w;
          float sq0 = 0.f, sq1 = 0.f;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float r0 = tanh_fn(__uint_as_float(v[(di * 2 + 0) * 3 + ch]) + s_bias[(di * 2 + 0) * 3 + ch]);
            const float r1 = tanh_fn(__uint_as_float(v[(di * 2 + 1) * 3 + ch]) + s_bias[(di * 2 + 1) * 3 + ch]);
            const float d0 = xpre[(di * 3 + ch) * 2] - r0, d1 = xpre[(di * 3 + ch) * 2 + 1] - r1;
            sq0 += d0 * d0;
            sq1 += d1 * d1;
            if (a.recon) *reinterpret_cast<float2*>(a.recon + off + ch * plane) = make_float2(r0, r1);
          }
          if (a.heat)
            *reinterpret_cast<float2*>(a.heat + static_cast<long long>(fb) * plane +
                                       static_cast<long long>(2 * h + di) * Wo + 2 * w) =
                make_float2(sq0 * (1.f / 3.f), sq1 * (1.f / 3.f));
          ssum += sq0 + sq1;
          smin = fminf(smin, fminf(sq0, sq1));
          smax = fmaxf(smax, fmaxf(sq0, sq1));
        }
      }
    }
    // reduction in a fixed order (deterministic): lanes of a warp here, the four warp partials of a tile and the tiles
    // of a frame in score_finalize_kernel — no cross-warp traffic or barrier in the tile loop
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
      smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
      smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    }
    if (lane == 0)
      *reinterpret_cast<float4*>(a.partials + (static_cast<long long>(t.m_tile) * 4 + q) * 4) =
          make_float4(ssum, smin * (1.f / 3.f), smax * (1.f / 3.f), 0.f);
  }
}

// Loads the model-input pixels a score epilogue will compare against (issued before the accumulator wait so the HBM
// latency overlaps the MMAs of the tile).
template <int EPI>
__device__ __forceinline__ void prefetch_x(const ConvArgs& a, const TileCoord& t, const EpiLane& L, float* xpre) {
  if constexpr (EPI == VAD_EPI_TANH_SCORE || EPI == VAD_EPI_CONVT_TANH_SCORE) {
    const int fb = t.b0 + L.bb, h = t.h0 + L.hh, w = t.w0 + L.ww;
    const bool valid = L.row_ok && (fb < a.B) && (h < a.H) && (w < a.W);
    if constexpr (EPI == VAD_EPI_TANH_SCORE) {
      const long long plane = static_cast<long long>(a.H) * a.W;
      const long long off = static_cast<long long>(fb) * 3 * plane + static_cast<long long>(h) * a.W + w;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) xpre[ch] = valid ? __ldg(a.x + off + ch * plane) : 0.f;
    } else {
      const int Wo = 2 * a.W;
      const long long plane = 4LL * a.H * a.W;
#pragma unroll
      for (int di = 0; di < 2; ++di) {
        const long long off = static_cast<long long>(fb) * 3 * plane + static_cast<long long>(2 * h + di) * Wo + 2 * w;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float2 xv = valid ? __ldg(reinterpret_cast<const float2*>(a.x + off + ch * plane)) : make_float2(0.f, 0.f);
          xpre[(di * 3 + ch) * 2] = xv.x;
          xpre[(di * 3 + ch) * 2 + 1] = xv.y;
        }
      }
    }
  }
}

// Shared body of the epilogue warps.  Group g (warps 4+4g .. 7+4g) handles this CTA's tiles g, g+G, g+2G, ...;
// with G = 2 each group owns one TMEM accumulator stage.
// PAIR: the CTA works on pairs of horizontally adjacent tiles (tile indices 2u, 2u+1); group g takes half g of every
// pair and alternates between the accumulator stages g and g+2.
template <int BN, int EPI, int G = epi_groups(BN), int KX = 1, int STRIDE = BN, bool PAIR = false>
__device__ __forceinline__ void epilogue_loop(const ConvArgs& a, uint32_t tmem_base, int warp, int lane, uint8_t* stg,
                                              const float* s_bias, float (*red_smem)[4][3], uint64_t* acc_full_bar,
                                              uint64_t* acc_empty_bar) {
  const int g = (warp - kEpiWarp0) >> 2;
  const int q = warp & 3;  // TMEM lane quarter == warp_id % 4
  uint8_t* my_stg = stg + g * staging_group_bytes(BN, EPI);
  int stg_i = 0;
  // group g walks tiles g, g+G, ... of this CTA; the x pixels a score epilogue needs are fetched one tile ahead
  TileIter ti(a, PAIR ? 2 * blockIdx.x + g : blockIdx.x + g * gridDim.x, PAIR ? 2 * gridDim.x : G * gridDim.x);
  const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
  const EpiLane L = make_epi_lane(a, q, lane);
  float xcur[12], xnext[12];
  if (ti.tile < a.total_tiles) prefetch_x<EPI>(a, ti.coord(a, BN), L, xcur);
  for (int n = 0; ti.tile < a.total_tiles; ++n) {
    const int as = PAIR ? g + 2 * (n & 1) : ((G >= 2) ? g : (n & 1));
    const uint32_t aphase = (PAIR || G < 2) ? ((n >> 1) & 1) : (n & 1);
    const TileCoord t = ti.coord(a, BN);
    ti.next(a);
    if (ti.tile < a.total_tiles) prefetch_x<EPI>(a, ti.coord(a, BN), L, xnext);
    // timeline rows 2/3: leaders of groups 0/1, or (VAD_DBG & 4) warps 0 and 3 of group 0
    const bool tl = lane == 0 && ((a.dbg & 4) ? (g == 0 && (q == 0 || q == 3)) : (g < 2 && q == 0));
    const int tl_row = (a.dbg & 4) ? (q == 0 ? 2 : 3) : 2 + g;
    if (tl) tl_stamp(a, tl_row, n, 0);
    mbar_wait_a(accf0 + as * 8, aphase, 4u | (static_cast<uint32_t>(n) << 8) | (static_cast<uint32_t>(g) << 28));
    if (tl) tl_stamp(a, tl_row, n, 1);
    tc_fence_after();
    if (tl) tl_stamp(a, tl_row, n, 8);
    const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * STRIDE);
    epilogue_tile<BN, EPI, KX>(a, t, L, tacc, q, lane, my_stg, s_bias, red_smem[g], stg_i, 1u + g, xcur, acce0 + as * 8,
                               tl ? tl_row : -1, n);
    if (tl) tl_stamp(a, tl_row, n, 2);
#pragma unroll
    for (int j = 0; j < 12; ++j) xcur[j] = xnext[j];
  }
  if (q == 0 && lane == 0) bulk_wait_group<0>();  // all TMA stores of this group have landed before the CTA exits
}

template <int BN>
__device__ __forceinline__ void load_bias_smem(const ConvArgs& a, float* s_bias, int n_total) {
  for (int i = threadIdx.x; i < n_total && i < kMaxBiasStream; i += blockDim.x) s_bias[i] = a.bias[i];
}

// ---------------------------------------------------------------------------------------------------- streaming
template <int CK, int BN, int EPI>
__global__ void __launch_bounds__(block_threads(BN, EPI), 1) conv_umma_kernel(const __grid_constant__ ConvArgs a) {
  using C = Cfg<CK, BN, EPI>;
  constexpr int kAS = acc_stages_for(epi_groups(BN, EPI));
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[2 * C::kStages];   // (b_resident: up to 2 * kStages - k_iters activation slots)
  __shared__ uint64_t empty_bar[2 * C::kStages];
  __shared__ uint64_t acc_full_bar[kMaxAccStages];
  __shared__ uint64_t acc_empty_bar[kMaxAccStages];
  __shared__ uint64_t turn_bar[2];  // MMA issuer ping-pong token
  __shared__ uint64_t w_bar;        // b_resident: this CTA's weight tiles have landed
  __shared__ uint32_t tmem_base_slot;
  __shared__ float red_smem[kMaxAccStages][4][3];
  __shared__ __align__(16) float s_bias[kMaxBiasStream];

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + C::kStages * C::kStageBytes;

  const int rows_valid = 1 << (a.lgTW + a.lgTH + a.lgTN);  // <= 128
  const uint32_t tx_bytes = static_cast<uint32_t>(rows_valid * C::kRowBytes + C::kBBytes);
  const int chunks = a.chunks0 + a.chunks1;
  const int k_iters = a.ntaps * chunks;
  // Ring slots.  Normally slot s = [A tile | B tile] at s * kStageBytes.  With resident weights (b_resident) weight tile
  // k sits in the B half of slot k for the whole launch and EVERY other half carries an activation tile: the A halves of
  // all kStages slots plus the B halves of slots k_iters .. kStages-1 — twice the bytes in flight per SM, which is
  // what a latency-bound streaming layer needs (the 720p transposed convolutions: nothing but loads on the critical
  // path — ablating the whole epilogue and the stores changes nothing).
  const int ns = a.b_resident ? 2 * C::kStages - k_iters : C::kStages;
  auto slot_off = [&](int sl) -> uint32_t {
    return static_cast<uint32_t>(sl < C::kStages ? sl * C::kStageBytes : (k_iters + sl - C::kStages) * C::kStageBytes + C::kABytes);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.mapA0);
    if (a.chunks1 > 0) tma_prefetch_desc(&a.mapA1);
    tma_prefetch_desc(&a.mapB);
    if (a.tma_store) tma_prefetch_desc(&a.mapOut);
    if (a.convt_split) tma_prefetch_desc(&a.mapOut2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2 * C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&turn_bar[0], 1);
    mbar_init(&turn_bar[1], 1);
    mbar_init(&w_bar, 1);
    for (int i = 0; i < kMaxAccStages; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4);  // one arrive per warp of the owning epilogue group
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<C::kTmemCols>(&tmem_base_slot);
    tmem_relinquish();
  }
  load_bias_smem<BN>(a, s_bias, a.n_tiles * BN);
  // Programmatic dependent launch (a.pdl != 0: launched with the stream-serialisation attribute).  Everything above
  // touches only constants (weights' tensor maps, bias); from here on the previous kernel's output is read.
  //   pdl == 1: wait here.   pdl == 2 (ConvLSTM step t >= 1): only the h_{t-1} loads and the cell state depend on the
  //   previous launch — the producer waits before its first h tile, the epilogue before it reads c; the x half of the
  //   K loop (loaded and multiplied first) overlaps the previous step's tail.
  if (a.pdl) pdl_launch_dependents();
  if (a.pdl == 1) pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // Role loops are executed by the WHOLE warp (uniform control flow); one elected lane issues the TMA / MMA.
  if (warp == 0) {
    // ===================================================================== TMA producer
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t smem0 = smem_addr_once(smem);
    int stage = 0;
    uint32_t phase = 0;
    // K order: source-major (all taps of source 0, then all taps of source 1) — the weight tile of (tap, chunk) is
    // addressed explicitly, so any order is valid, and this one puts everything that does not depend on the previous
    // ConvLSTM step first
    bool waited = a.pdl != 2;
    // Resident weights (transposed convolutions: one tap, k_iters <= ring depth, grid a multiple of n_tiles so that this
    // CTA's tiles all have the same n0): weight tile k lives in the B half of ring slot k for the whole launch and the
    // ring's A halves carry the activations — half the L2 -> shared-memory traffic and half the TMA issues per tile
    // (the 720p video decoder's transposed convolutions ran at 0.35-0.45 of the HBM peak with L2 70 % busy).
    const uint32_t a_bytes = static_cast<uint32_t>(rows_valid * C::kRowBytes);
    if (a.b_resident && blockIdx.x < a.total_tiles) {
      TileIter t0(a, blockIdx.x, gridDim.x);
      const int n0 = t0.coord(a, BN).n0;
      if (elect_one()) {
        mbar_arrive_expect_tx(&w_bar, static_cast<uint32_t>(k_iters * C::kBBytes));
        for (int kq = 0; kq < k_iters; ++kq)
          tma_load_2d(smem + kq * C::kStageBytes + C::kABytes, &a.mapB, &w_bar, kq * CK, n0);
      }
      __syncwarp();
    }
    for (TileIter ti(a, blockIdx.x, gridDim.x); ti.tile < a.total_tiles; ti.next(a)) {
      const TileCoord t = ti.coord(a, BN);
      for (int src = 0; src < 2; ++src) {
        const int c_lo = src == 0 ? 0 : a.chunks0, c_hi = src == 0 ? a.chunks0 : chunks;
        if (c_lo == c_hi) continue;
        if (src == 1 && !waited) { pdl_wait(); waited = true; }
        int dy = (a.ntaps == 9) ? -1 : 0, dx = dy;
        for (int tap = 0; tap < a.ntaps; ++tap) {
          int kcol = tap * a.w_ctap + c_lo * CK;
          for (int c = c_lo; c < c_hi; ++c) {
            mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 1);
            if (elect_one()) {
              const uint32_t sa = smem0 + slot_off(stage);
              const uint32_t fb = full0 + stage * 8;
              mbar_arrive_expect_tx_a(fb, a.b_resident ? a_bytes : tx_bytes);
              if (src == 0)
                tma_load_5d_a(sa, &a.mapA0, fb, c * CK, t.w0 + dx, t.h0 + dy, a.tA0, t.b0);
              else
                tma_load_5d_a(sa, &a.mapA1, fb, (c - a.chunks0) * CK, t.w0 + dx, t.h0 + dy, a.tA1, t.b0);
              if (!a.b_resident) tma_load_2d_a(sa + C::kABytes, &a.mapB, fb, kcol, t.n0);
            }
            __syncwarp();
            kcol += CK;
            if (++stage == ns) { stage = 0; phase ^= 1u; }
          }
          if (++dx == 2) { dx = -1; ++dy; }  // next tap (3x3: row-major over (dy, dx) in -1..1)
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================================================================== MMA issuers (two warps, alternate tiles)
    // Two issuers are only safe while a tile's k-loop is SHORTER than the operand ring: the issuer that runs ahead
    // then never waits on a slot a full ring revolution early (mbarrier phase parity would alias and let it through
    // on stale data).  Long k-loops use warp 1 alone.
    const bool dual = a.dual_mma && k_iters < ns;
    const int mi = warp == 1 ? 0 : 1;
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, BN);
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const uint64_t da_base = umma_smem_desc(smem_u32(smem), C::kSBO, C::kLayout);
    const uint64_t db_base = umma_smem_desc(smem_u32(smem + C::kABytes), C::kSBO, C::kLayout);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    if (a.b_resident && blockIdx.x < a.total_tiles && (dual || mi == 0)) mbar_wait(&w_bar, 0, 5);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
      if (!dual && mi == 1) break;
      if (dual && (it & 1) != mi) {  // the other issuer's tile: only keep the ring position in step
        stage += k_iters % ns;
        phase ^= static_cast<uint32_t>((k_iters / ns) & 1);
        if (stage >= ns) { stage -= ns; phase ^= 1u; }
        continue;
      }
      const int as = it % kAS;
      mbar_wait_a(acce0 + as * 8, ((it / kAS) & 1) ^ 1u, 3);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait_a(full0 + stage * 8, phase, 2);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t soff = static_cast<uint64_t>(slot_off(stage) >> 4);
          const uint64_t da = da_base + soff;
          const uint64_t db = a.b_resident ? db_base + static_cast<uint64_t>(k * (C::kStageBytes >> 4)) : db_base + soff;
#pragma unroll
          for (int kk = 0; kk < CK / 16; ++kk) {
            // advance 16 bf16 = 32 B inside the swizzle span: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc,
                      (k > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit_a(empty0 + stage * 8);                       // frees the smem slot when these MMAs retire
          if (k == k_iters - 1) umma_commit_a(accf0 + as * 8);      // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == ns) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    epilogue_loop<BN, EPI, epi_groups(BN, EPI)>(a, tmem_base, warp, lane, stg, s_bias, red_smem, acc_full_bar, acc_empty_bar);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------- halo
constexpr int kHaloMaxStages = 8;

template <int CK, int BN, int EPI>
__global__ void __launch_bounds__(block_threads(BN), 1) conv_halo_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int kRowBytes = CK * 2;
  constexpr int kBBytes = BN * kRowBytes;  // one tap's weight slab
  constexpr uint32_t kLayout = (CK == 64) ? 2u : 4u;
  constexpr uint32_t kTmemCols = tmem_cols_for(BN);
  constexpr int kAS = acc_stages_for(epi_groups(BN));
  static_assert(kBBytes % 1024 == 0, "weight slab must keep 1024B alignment");
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_bar;
  __shared__ uint64_t full_bar[kHaloMaxStages];
  __shared__ uint64_t empty_bar[kHaloMaxStages];
  __shared__ uint64_t acc_full_bar[kMaxAccStages];
  __shared__ uint64_t acc_empty_bar[kMaxAccStages];
  __shared__ uint64_t turn_bar[2];  // MMA issuer ping-pong token
  __shared__ uint32_t tmem_base_slot;
  __shared__ float red_smem[kMaxAccStages][4][3];
  __shared__ __align__(16) float s_bias[kMaxBias];

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;                                   // 9 weight slabs, resident for the CTA's lifetime
  uint8_t* s_a = smem + 9 * kBBytes;                     // ring of input patches
  const int stage_bytes = a.halo_patch_bytes * a.halo_npatch;
  uint8_t* stg = s_a + a.halo_stages * stage_bytes;      // epilogue staging
  const uint32_t patch_tx = static_cast<uint32_t>(a.halo_npatch * a.halo_pw * a.halo_ph * kRowBytes);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.mapA0);
    tma_prefetch_desc(&a.mapB);
    if (a.tma_store) tma_prefetch_desc(&a.mapOut);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < a.halo_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&turn_bar[0], 1);
    mbar_init(&turn_bar[1], 1);
    for (int i = 0; i < kMaxAccStages; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<kTmemCols>(&tmem_base_slot);
    tmem_relinquish();
  }
  load_bias_smem<BN>(a, s_bias, BN);
  if (a.pdl) {  // programmatic dependent launch: everything above touched constants only
    pdl_launch_dependents();
    pdl_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer (whole warp, elected issue)
    if (elect_one()) {  // weights: nine [BN x CK] slabs, once
      mbar_arrive_expect_tx(&w_bar, 9u * kBBytes);
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(s_w + tap * kBBytes, &a.mapB, &w_bar, tap * a.w_ctap, 0);
    }
    __syncwarp();
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t sa0 = smem_addr_once(s_a);
    int stage = 0;
    uint32_t phase = 0;
    int pn = 0;
    for (TileIter ti(a, blockIdx.x, gridDim.x); ti.tile < a.total_tiles; ti.next(a), ++pn) {
      const TileCoord t = ti.coord(a, BN);
      if (lane == 0) tl_stamp(a, 0, pn, 0);
      mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 1);
      if (lane == 0) tl_stamp(a, 0, pn, 1);
      if (elect_one()) {
        const uint32_t sa = sa0 + stage * stage_bytes;
        if ((a.dbg & 128) && pn >= a.halo_stages) {  // ablation: no input traffic after the ring's first fill
          mbar_arrive_a(full0 + stage * 8);
        } else {
          mbar_arrive_expect_tx_a(full0 + stage * 8, patch_tx);
          for (int p = 0; p < a.halo_npatch; ++p)
            tma_load_5d_a(sa + p * a.halo_patch_bytes, &a.mapA0, full0 + stage * 8, 0, t.w0 - 1 + p, t.h0 - 1, a.tA0,
                          t.b0);
        }
      }
      __syncwarp();
      if (++stage == a.halo_stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================================================================== MMA issuers (two warps, alternate tiles)
    // The issuer is a single warp running scalar code whose per-tile latency (two barrier waits, 18 descriptor
    // updates, 18 issues, two commits) exceeds the tensor pipe's time for the tile; with two warps taking alternate
    // tiles one waits / prepares while the other's MMAs are in flight.  A token (turn_bar) makes them issue strictly
    // in tile order.
    const int mi = warp == 1 ? 0 : 1;
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, BN);
    const uint32_t sbo = static_cast<uint32_t>(a.halo_sbo_rows * kRowBytes);
    // per-tap operand start offsets in 16-byte units (the descriptor's address field), fixed for the whole kernel
    uint32_t tap_off[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
      tap_off[tap] = static_cast<uint32_t>(a.tap_patch[tap] * a.halo_patch_bytes + a.tap_row[tap] * kRowBytes) >> 4;
    const uint64_t da_hi = umma_smem_desc(0, sbo, kLayout);             // everything but the start address
    const uint64_t db0 = umma_smem_desc(smem_u32(s_w), 8 * kRowBytes, kLayout);
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const uint32_t turn0 = smem_addr_once(&turn_bar[0]);
    const uint32_t sa16_0 = (smem_u32(s_a) & 0x3FFFF) >> 4, stage16 = static_cast<uint32_t>(stage_bytes) >> 4;
    const bool dual = a.dual_mma != 0;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    mbar_wait(&w_bar, 0, 5);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
      if (!dual && mi == 1) break;
      if (!dual || (it & 1) == mi) {
        const int as = it % kAS;
        if (lane == 0) tl_stamp(a, 1, it, 0);
        mbar_wait_a(acce0 + as * 8, ((it / kAS) & 1) ^ 1u, 3);
        if (lane == 0) tl_stamp(a, 1, it, 1);
        mbar_wait_a(full0 + stage * 8, phase, 2);
        if (dual && a.token && it > 0) mbar_wait_a(turn0 + mi * 8, static_cast<uint32_t>(((it - 1) >> 1) & 1), 8);
        if (lane == 0) tl_stamp(a, 1, it, 2);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
          const uint64_t da0 = da_hi + static_cast<uint64_t>(sa16_0 + stage * stage16);
          // Pixel-pair folding (a.pair_fold: rows = pairs of horizontally adjacent pixels, CK = 2 x 32 channels,
          // N = 2 x Cout): the left neighbour pair only feeds the tap through its SECOND pixel (K columns 32..63) and
          // the right one through its FIRST (0..31) — the other half of their weight slabs is zero and is not issued:
          // 24 MMAs per 256 pixels instead of 36.
          uint32_t acc = 0u;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if ((a.dbg & 16) && tap > 0) break;  // ablation: one tap only
            const uint64_t da = da0 + static_cast<uint64_t>(tap_off[tap]);
            const uint64_t db = db0 + static_cast<uint64_t>((tap * kBBytes) >> 4);
#pragma unroll
            for (int kk = 0; kk < CK / 16; ++kk) {
              if (CK == 64 && a.pair_fold && ((tap % 3 == 0 && kk < 2) || (tap % 3 == 2 && kk >= 2))) continue;
              umma_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc, acc);
              acc = 1u;
            }
          }
          umma_commit_a(empty0 + stage * 8);
          umma_commit_a(accf0 + as * 8);
          if (dual && a.token) mbar_arrive_a(turn0 + (mi ^ 1) * 8);
        }
        __syncwarp();
        if (lane == 0) tl_stamp(a, 1, it, 3);
      }
      if (++stage == a.halo_stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp >= kEpiWarp0) {
    epilogue_loop<BN, EPI>(a, tmem_base, warp, lane, stg, s_bias, red_smem, acc_full_bar, acc_empty_bar);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------- kx-merged halo
// 3x3 conv with the horizontal taps folded into N (see kx_mma_n above).  Per tile: ONE TMA patch of 18 rows x 8
// columns x CK channels (rows h0-1 .. h0+16, columns w0-1 .. w0+6); A operand of vertical tap ky = the patch shifted
// down by ky rows (start address + ky*8 pixel rows, standard 8-row core groups); B = resident weight slab ky of
// [kx*Cout + co][ci].  3 * CK/16 MMAs per tile instead of 9 * CK/16, each with 3x the columns.
template <int CK, int BN, int EPI>
__global__ void __launch_bounds__(128 + 128 * kx_groups(BN, EPI), 1) conv_kx_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int kRowBytes = CK * 2;
  constexpr int NM = kx_mma_n(BN, EPI);       // MMA N extent
  constexpr int kBBytes = NM * kRowBytes;     // one vertical tap's weight slab
  constexpr int kPatchBytes = 8 * 18 * kRowBytes;
  constexpr uint32_t kLayout = (CK == 64) ? 2u : 4u;
  constexpr int G = kx_groups(BN, EPI);
  constexpr int kStride = kx_acc_stride(BN, EPI);
  constexpr uint32_t kTmemCols = kx_tmem_cols(BN, EPI);
  static_assert(kBBytes % 1024 == 0 && kPatchBytes % 1024 == 0, "operand slabs must keep 1024B alignment");
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_bar;
  __shared__ uint64_t full_bar[kHaloMaxStages];
  __shared__ uint64_t empty_bar[kHaloMaxStages];
  __shared__ uint64_t acc_full_bar[kMaxAccStages];
  __shared__ uint64_t acc_empty_bar[kMaxAccStages];
  __shared__ uint64_t turn_bar[2];  // MMA issuer ping-pong token
  __shared__ uint32_t tmem_base_slot;
  __shared__ float red_smem[kMaxAccStages][4][3];
  __shared__ __align__(16) float s_bias[kMaxBias];

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;                               // 3 weight slabs, resident for the CTA's lifetime
  uint8_t* s_a = smem + 3 * kBBytes;                 // ring of input patches
  uint8_t* stg = s_a + a.halo_stages * kPatchBytes;  // epilogue staging

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.mapA0);
    tma_prefetch_desc(&a.mapB);
    if (a.tma_store) tma_prefetch_desc(&a.mapOut);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < a.halo_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&turn_bar[0], 1);
    mbar_init(&turn_bar[1], 1);
    for (int i = 0; i < kMaxAccStages; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<kTmemCols>(&tmem_base_slot);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < BN; i += 128 + 128 * G) s_bias[i] = a.bias[i];
  if (a.pdl) {  // programmatic dependent launch: everything above touched constants only
    pdl_launch_dependents();
    pdl_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0 || (warp == 2 && (a.halo_stages & 1) == 0)) {
    // ===================================================================== TMA producers
    // warp 0, plus warp 2 (idle after the TMEM allocation) on alternate tiles when the ring depth is even (each slot
    // then always belongs to the same producer)
    const bool two = (a.halo_stages & 1) == 0;
    const int pi = warp == 0 ? 0 : 1;
    if (warp == 0) {
      if (elect_one()) {
        mbar_arrive_expect_tx(&w_bar, 3u * kBBytes);
        for (int ky = 0; ky < 3; ++ky) tma_load_2d(s_w + ky * kBBytes, &a.mapB, &w_bar, ky * a.w_ctap, 0);
      }
      __syncwarp();
    }
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t sa0 = smem_addr_once(s_a);
    const int pstep = two ? 2 : 1;
    int it = pi;
    for (TileIter ti(a, blockIdx.x + pi * gridDim.x, pstep * gridDim.x); ti.tile < a.total_tiles; ti.next(a), it += pstep) {
      const TileCoord t = ti.coord(a, BN);
      const int stage = it % a.halo_stages;
      const uint32_t phase = static_cast<uint32_t>((it / a.halo_stages) & 1);
      mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 1);
      if (elect_one()) {
        mbar_arrive_expect_tx_a(full0 + stage * 8, kPatchBytes);
        tma_load_5d_a(sa0 + stage * kPatchBytes, &a.mapA0, full0 + stage * 8, 0, t.w0 - 1, t.h0 - 1, a.tA0, t.b0);
      }
      __syncwarp();
    }
  } else if (warp == 1 || warp == 3) {
    // ===================================================================== MMA issuers (alternate tiles)
    const int mi = warp == 1 ? 0 : 1;
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, NM);
    const uint64_t da_hi = umma_smem_desc(0, 8 * kRowBytes, kLayout);
    const uint64_t db0 = umma_smem_desc(smem_u32(s_w), 8 * kRowBytes, kLayout);
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const uint32_t turn0 = smem_addr_once(&turn_bar[0]);
    const uint32_t sa16_0 = (smem_u32(s_a) & 0x3FFFF) >> 4;
    const bool dual = a.dual_mma != 0;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    mbar_wait(&w_bar, 0, 5);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
      if (!dual && mi == 1) break;
      if (!dual || (it & 1) == mi) {
        const int as = it % G;
        if (lane == 0) tl_stamp(a, 1, it, 0);
        mbar_wait_a(acce0 + as * 8, ((it / G) & 1) ^ 1u, 3);
        if (lane == 0) tl_stamp(a, 1, it, 1);
        mbar_wait_a(full0 + stage * 8, phase, 2);
        // ping-pong: start issuing only after the other issuer has issued its whole tile, so that this warp's waits
        // overlap the other's MMAs and the tensor pipe sees one uninterrupted stream (without the token both warps
        // interleave their MMAs, block on the same queue and then sit in their waits at the same time)
        if (dual && a.token && it > 0) mbar_wait_a(turn0 + mi * 8, static_cast<uint32_t>(((it - 1) >> 1) & 1), 8);
        if (lane == 0) tl_stamp(a, 1, it, 2);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kStride);
          const uint32_t sa16 = sa16_0 + stage * (kPatchBytes >> 4);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint64_t da = da_hi | static_cast<uint64_t>(sa16 + ((ky * 8 * kRowBytes) >> 4));
            const uint64_t db = db0 + static_cast<uint64_t>((ky * kBBytes) >> 4);
#pragma unroll
            for (int kk = 0; kk < CK / 16; ++kk)
              umma_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc,
                        (ky > 0 || kk > 0) ? 1u : 0u);
          }
          umma_commit_a(empty0 + stage * 8);
          umma_commit_a(accf0 + as * 8);
          if (dual && a.token) mbar_arrive_a(turn0 + (mi ^ 1) * 8);
        }
        __syncwarp();
        if (lane == 0) tl_stamp(a, 1, it, 3);
      }
      if (++stage == a.halo_stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp >= kEpiWarp0) {
    epilogue_loop<BN, EPI, G, 3, kStride>(a, tmem_base, warp, lane, stg, s_bias, red_smem, acc_full_bar, acc_empty_bar);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------- halo + streamed B
// Wide 3x3 layers (Cin >= 128, N tile 128): the nine weight slabs of all channel chunks no longer fit in shared memory,
// and the tap-per-stage streaming kernel re-reads every input pixel nine times from L2 (it runs at the L2 bandwidth
// limit, not at the tensor pipe's).  Here a CTA works on PAIRS of 8x16 tiles (a 16x16 pixel block, M = 256): per
// 64-channel chunk ONE TMA patch of 18x18 pixels serves all nine taps of both tiles through shifted descriptors, and
// every streamed [128 x 64] weight tile is used by both tiles before its ring slot is released — per 128 output
// pixels 185 KB of operand traffic instead of 576 KB.
constexpr int kHsBRing = 4;                       // weight tiles in flight (16 KB each); 6 + two patch slots is slower
constexpr int kHsPatchBytes = 18 * 18 * 128;      // 41472: one chunk's patch, rows = pixels (y*18 + x), 128 B each
constexpr int kHsPatchPitch = 41984;              // 1024-aligned ring pitch
constexpr int kHsBBytes = 128 * 128;              // [128 n][64 k] bf16

template <int EPI>
__global__ void __launch_bounds__(128 + 256, 1) conv_hs_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int BN = 128;
  constexpr int kRowBytes = 128;
  constexpr uint32_t kLayout = 2u;  // SWIZZLE_128B
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t p_full[4], p_empty[4];
  __shared__ uint64_t b_full[kHsBRing], b_empty[kHsBRing];
  __shared__ uint64_t acc_full_bar[kMaxAccStages];
  __shared__ uint64_t acc_empty_bar[kMaxAccStages];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float red_smem[kMaxAccStages][4][3];
  __shared__ __align__(16) float s_bias[kMaxBias];

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_b = smem;                                 // weight ring
  uint8_t* s_p = smem + kHsBRing * kHsBBytes;          // patch ring (a.halo_stages slots)
  uint8_t* stg = s_p + a.halo_stages * kHsPatchPitch;  // epilogue staging
  const int chunks = a.chunks0;
  const int units = a.total_tiles >> 1;                // (tile pair, n tile) work items

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.mapA0);
    tma_prefetch_desc(&a.mapB);
    if (a.tma_store) tma_prefetch_desc(&a.mapOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&p_full[i], 1);
      mbar_init(&p_empty[i], 1);
    }
    for (int i = 0; i < kHsBRing; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < kMaxAccStages; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<512>(&tmem_base_slot);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < a.n_tiles * BN && i < kMaxBias; i += 384) s_bias[i] = a.bias[i];
  if (a.pdl) {  // programmatic dependent launch: everything above touched constants only
    pdl_launch_dependents();
    pdl_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    // two streams over the same (unit, chunk) sequence: patches run one step ahead of the weight tiles so that a
    // chunk's patch is already in flight while the previous chunk's nine weight tiles are being issued
    const uint32_t pf0 = smem_addr_once(&p_full[0]), pe0 = smem_addr_once(&p_empty[0]);
    const uint32_t bf0 = smem_addr_once(&b_full[0]), be0 = smem_addr_once(&b_empty[0]);
    const uint32_t sp0 = smem_addr_once(s_p), sb0 = smem_addr_once(s_b);
    TileIter tp(a, 2 * blockIdx.x, 2 * gridDim.x);  // patch stream position (unit) ...
    int cp = 0;                                     // ... and chunk
    int ps = 0;
    uint32_t pphase = 0;
    int bs = 0;
    uint32_t bphase = 0;
    auto issue_patch = [&]() {
      const TileCoord t = tp.coord(a, BN);
      mbar_wait_a(pe0 + ps * 8, pphase ^ 1u, 1);
      if (elect_one()) {
        mbar_arrive_expect_tx_a(pf0 + ps * 8, kHsPatchBytes);
        tma_load_5d_a(sp0 + ps * kHsPatchPitch, &a.mapA0, pf0 + ps * 8, cp * 64, t.w0 - 1, t.h0 - 1, a.tA0, t.b0);
      }
      __syncwarp();
      if (++ps == a.halo_stages) { ps = 0; pphase ^= 1u; }
      if (++cp == chunks) { cp = 0; tp.next(a); }
    };
    if (tp.tile < a.total_tiles) issue_patch();
    for (TileIter tb_(a, 2 * blockIdx.x, 2 * gridDim.x); tb_.tile < a.total_tiles; tb_.next(a)) {
      const int n0 = tb_.n_t * BN;
      for (int c = 0; c < chunks; ++c) {
        if (tp.tile < a.total_tiles) issue_patch();
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait_a(be0 + bs * 8, bphase ^ 1u, 9);
          if (elect_one()) {
            mbar_arrive_expect_tx_a(bf0 + bs * 8, kHsBBytes);
            tma_load_2d_a(sb0 + bs * kHsBBytes, &a.mapB, bf0 + bs * 8, tap * a.w_ctap + c * 64, n0);
          }
          __syncwarp();
          if (++bs == kHsBRing) { bs = 0; bphase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, BN);
    const uint32_t pf0 = smem_addr_once(&p_full[0]), pe0 = smem_addr_once(&p_empty[0]);
    const uint32_t bf0 = smem_addr_once(&b_full[0]), be0 = smem_addr_once(&b_empty[0]);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const uint64_t da_base = umma_smem_desc(smem_u32(s_p), 18 * kRowBytes, kLayout);  // 8-row groups one patch row apart
    const uint64_t db_base = umma_smem_desc(smem_u32(s_b), 8 * kRowBytes, kLayout);
    int ps = 0;
    uint32_t pphase = 0;
    int bs = 0;
    uint32_t bphase = 0;
    int j = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++j) {
      const int as0 = 2 * (j & 1);
      const uint32_t aph = static_cast<uint32_t>(((j >> 1) & 1) ^ 1);
      mbar_wait_a(acce0 + as0 * 8, aph, 3);
      mbar_wait_a(acce0 + (as0 + 1) * 8, aph, 3);
      tc_fence_after();
      const uint32_t d0 = tmem_base + static_cast<uint32_t>(as0 * BN);
      for (int c = 0; c < chunks; ++c) {
        mbar_wait_a(pf0 + ps * 8, pphase, 2);
        const uint64_t da_c = da_base + static_cast<uint64_t>(ps * (kHsPatchPitch >> 4));
        int ky = 0, kx = 0;
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait_a(bf0 + bs * 8, bphase, 10);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t db = db_base + static_cast<uint64_t>(bs * (kHsBBytes >> 4));
            const uint64_t da_t = da_c + static_cast<uint64_t>(((ky * 18 + kx) * kRowBytes) >> 4);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16(d0 + static_cast<uint32_t>(m * BN), da_t + static_cast<uint64_t>(m * ((8 * kRowBytes) >> 4) + kk * 2),
                          db + static_cast<uint64_t>(kk * 2), idesc, (c > 0 || tap > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit_a(be0 + bs * 8);
            if (tap == 8) umma_commit_a(pe0 + ps * 8);
            if (tap == 8 && c == chunks - 1) {
              umma_commit_a(accf0 + as0 * 8);
              umma_commit_a(accf0 + (as0 + 1) * 8);
            }
          }
          __syncwarp();
          if (++bs == kHsBRing) { bs = 0; bphase ^= 1u; }
          if (++kx == 3) { kx = 0; ++ky; }
        }
        if (++ps == a.halo_stages) { ps = 0; pphase ^= 1u; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    epilogue_loop<BN, EPI, 2, 1, BN, true>(a, tmem_base, warp, lane, stg, s_bias, red_smem, acc_full_bar, acc_empty_bar);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------- ConvLSTM sequence
// One launch for all T steps of a ConvLSTM layer (reference ConvLSTM.forward time loop, models/video_autoencoder.py:
// 153-167) when every (pixel tile, 32-channel tile) fits one CTA per SM: the CTA keeps its tile for the whole sequence,
// the CELL STATE LIVES IN REGISTERS (32 fp32 per epilogue thread) and only h_t goes to memory (it is the next step's
// MMA operand for this CTA and its neighbours, and the layer's output).  Steps are separated by a grid-wide counter:
// a CTA publishes h_t (TMA store complete -> release-add), the producer of every CTA acquires `t * gridDim.x` before
// its first h_{t-1} load.  The x half of step t+1's K loop (input sequence, independent of the recurrence) is loaded
// and multiplied while step t's epilogue and the exchange are still in flight.
template <int CK>
__global__ void __launch_bounds__(256, 1) convlstm_seq_kernel(const __grid_constant__ ConvArgs a, int T,
                                                              unsigned int* __restrict__ step_counter) {
  constexpr int BN = 128;
  using C = Cfg<CK, BN, VAD_EPI_LSTM>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[C::kStages];
  __shared__ uint64_t empty_bar[C::kStages];
  __shared__ uint64_t acc_full_bar[2];
  __shared__ uint64_t acc_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[BN];

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + C::kStages * C::kStageBytes;  // one 8 KB staged h tile

  const int rows_valid = 1 << (a.lgTW + a.lgTH + a.lgTN);
  const uint32_t tx_bytes = static_cast<uint32_t>(rows_valid * C::kRowBytes + C::kBBytes);
  const TileCoord tc = TileIter(a, blockIdx.x, gridDim.x).coord(a, BN);  // this CTA's tile, for every step

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.mapA0);
    tma_prefetch_desc(&a.mapA1);
    tma_prefetch_desc(&a.mapB);
    tma_prefetch_desc(&a.mapOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<256>(&tmem_base_slot);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < BN; i += 256) s_bias[i] = a.bias[tc.n0 + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t smem0 = smem_addr_once(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < T; ++t) {
      for (int src = 0; src < 2; ++src) {
        if (src == 1) {
          if (t == 0) break;  // h_{-1} = 0: the h half of K is skipped
          // every CTA has published h_{t-1} (the 3x3 halo reaches into the neighbours' tiles)
          const unsigned int target = static_cast<unsigned int>(t) * gridDim.x;
          if (lane == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(step_counter) < target) {
              if (clock64() - t0 > VAD_WAIT_TIMEOUT_CLOCKS) {
                if (g_vad_trap_slot) {
                  g_vad_trap_slot[0] = 11;
                  g_vad_trap_slot[1] = blockIdx.x;
                  g_vad_trap_slot[2] = static_cast<unsigned long long>(t);
                  g_vad_trap_slot[3] = target;
                  __threadfence_system();
                }
                __trap();
              }
            }
          }
          __syncwarp();
          fence_proxy_async_all();  // the acquired data is read through the async proxy (TMA) below
        }
        const int n_chunks = src == 0 ? a.chunks0 : a.chunks1;
        const int kbase = src == 0 ? 0 : a.chunks0 * CK;
        int dy = -1, dx = -1;
        for (int tap = 0; tap < 9; ++tap) {
          int kcol = tap * a.w_ctap + kbase;
          for (int c = 0; c < n_chunks; ++c) {
            mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 1);
            if (elect_one()) {
              const uint32_t sa = smem0 + stage * C::kStageBytes;
              const uint32_t fb = full0 + stage * 8;
              mbar_arrive_expect_tx_a(fb, tx_bytes);
              if (src == 0)
                tma_load_5d_a(sa, &a.mapA0, fb, c * CK, tc.w0 + dx, tc.h0 + dy, t, tc.b0);
              else
                tma_load_5d_a(sa, &a.mapA1, fb, c * CK, tc.w0 + dx, tc.h0 + dy, t - 1, tc.b0);
              tma_load_2d_a(sa + C::kABytes, &a.mapB, fb, kcol, tc.n0);
            }
            __syncwarp();
            kcol += CK;
            if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
          }
          if (++dx == 2) { dx = -1; ++dy; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, BN);
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const uint64_t da_base = umma_smem_desc(smem_u32(smem), C::kSBO, C::kLayout);
    const uint64_t db_base = umma_smem_desc(smem_u32(smem + C::kABytes), C::kSBO, C::kLayout);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < T; ++t) {
      const int k_iters = 9 * (a.chunks0 + (t > 0 ? a.chunks1 : 0));
      const int as = t & 1;
      mbar_wait_a(acce0 + as * 8, static_cast<uint32_t>(((t >> 1) & 1) ^ 1), 3);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait_a(full0 + stage * 8, phase, 2);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t soff = static_cast<uint64_t>(stage * (C::kStageBytes >> 4));
#pragma unroll
          for (int kk = 0; kk < CK / 16; ++kk)
            umma_bf16(d_tmem, da_base + soff + static_cast<uint64_t>(kk * 2), db_base + soff + static_cast<uint64_t>(kk * 2),
                      idesc, (k > 0 || kk > 0) ? 1u : 0u);
          umma_commit_a(empty0 + stage * 8);
          if (k == k_iters - 1) umma_commit_a(accf0 + as * 8);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================================== epilogue: gates, c (registers), h
    const int q = warp & 3;
    const EpiLane L = make_epi_lane(a, q, lane);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const bool leader = (q == 0 && lane == 0);
    const int fb = tc.b0 + L.bb, h = tc.h0 + L.hh, w = tc.w0 + L.ww;
    const bool valid = L.row_ok && (fb < a.B) && (h < a.H) && (w < a.W);
    const int j0 = (tc.n0 >> 7) * 32;  // first hidden channel of this tile
    float c[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) c[e] = 0.f;
    for (int t = 0; t < T; ++t) {
      const int as = t & 1;
      mbar_wait_a(accf0 + as * 8, static_cast<uint32_t>((t >> 1) & 1), 4u | (static_cast<uint32_t>(t) << 8));
      tc_fence_after();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
      // (the staging buffer is free: the leader waited for the previous store's completion before publishing h_{t-1},
      //  and every warp passed the barrier below after that)
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        uint32_t gi[8], gf[8], gg[8], go[8];
        tmem_ld_x8(tacc + 0 + s * 8, gi);
        tmem_ld_x8(tacc + 32 + s * 8, gf);
        tmem_ld_x8(tacc + 64 + s * 8, gg);
        tmem_ld_x8(tacc + 96 + s * 8, go);
        tmem_ld_wait();
        if (s == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(acce0 + as * 8);
        }
        float hn[8];
        const float* bp = s_bias + s * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xi = __uint_as_float(gi[e]) + bp[e];
          const float xf = __uint_as_float(gf[e]) + bp[32 + e];
          const float xg = __uint_as_float(gg[e]) + bp[64 + e];
          const float xo = __uint_as_float(go[e]) + bp[96 + e];
          const float cn = sigmoid_fn(xf) * c[s * 8 + e] + sigmoid_fn(xi) * tanh_fn(xg);
          c[s * 8 + e] = cn;
          hn[e] = sigmoid_fn(xo) * tanh_fn(cn);
        }
        sts128(stg + staged_off(L.srow, s, 32), make_uint4(pack_bf16x2(hn[0], hn[1]), pack_bf16x2(hn[2], hn[3]),
                                                            pack_bf16x2(hn[4], hn[5]), pack_bf16x2(hn[6], hn[7])));
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (leader) {
        tma_store_5d(&a.mapOut, stg, j0, tc.w0, tc.h0, t, tc.b0);
        bulk_commit_group();
        bulk_wait_group_read<0>();
      }
      named_bar_sync(1, 128);  // nobody overwrites the staging buffer before the store has read it
      if (leader) {
        bulk_wait_group<0>();  // h_t of this tile is in global memory ...
        __threadfence();
        red_release_gpu_add(step_counter, 1u);  // ... and published
      }
    }
    if (valid && a.c_state != nullptr) {  // final cell state (ConvLSTM.forward returns it; VideoAutoencoder drops it)
      float* cptr = a.c_state + ((static_cast<long long>(fb) * a.H + h) * a.W + w) * a.cout + j0;
#pragma unroll
      for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(cptr + e) = make_float4(c[e], c[e + 1], c[e + 2], c[e + 3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------- ConvLSTM, patches
// The sequence kernel above streams one (A tile, B tile) pair per tap and chunk: 1.15 MB per CTA and step, and with
// ~190 KB of shared memory in flight against ~1.2 us of TMA latency that alone takes ~7 us per step (clusters that
// multicast the weights did not help: the limit is bytes in flight per SM, not L2 reads).  Here the A side uses
// patches like the other 3x3 kernels — per 64-channel chunk ONE TMA patch serves all nine taps through shifted
// descriptors (A traffic 576 -> 102 KB per step) — and the weight tiles stream through their own 8-slot ring, fed
// by a second producer warp that never waits for the recurrence (weights do not depend on it).
//   geometry 1 (8x8 frames, two frames per tile): patch [y 10][frame 2][x 10], rows ((y*2+f)*8 + x), map dims
//               {C, W, B, H, T};   geometry 2 (tile 8 wide x 16 tall inside one frame): patch [y 18][x 10].
constexpr int kLpBRing = 8;               // (10 slots + 2 patch slots measured the same: the step is bound by the serial
                                          //  chain h MMAs -> gates -> store -> publish -> acquire -> patch load, not by the stream)
constexpr int kLpPatchPitch = 26 * 1024;  // >= 200 rows x 128 B
constexpr int kLpPatches = 3;

__global__ void __launch_bounds__(384, 1) convlstm_patch_kernel(const __grid_constant__ ConvArgs a, int T,
                                                                unsigned int* __restrict__ step_counter) {
  constexpr int BN = 128, CK = 64;
  constexpr int kRowBytes = 128;
  constexpr int kBBytes = BN * kRowBytes;  // 16 KB
  constexpr uint32_t kLayout = 2u;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t p_full[kLpPatches], p_empty[kLpPatches];
  __shared__ uint64_t b_full[kLpBRing], b_empty[kLpBRing];
  __shared__ uint64_t acc_full_bar[2];
  __shared__ uint64_t acc_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[BN];

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_b = smem;
  uint8_t* s_p = smem + kLpBRing * kBBytes;
  uint8_t* stg = s_p + kLpPatches * kLpPatchPitch;

  const bool geo1 = a.row_perm == 2;
  const uint32_t patch_tx = static_cast<uint32_t>((geo1 ? 200 : 180) * kRowBytes);
  const int ky_rows = geo1 ? 20 : 10;  // patch rows per image row
  const TileCoord tc = TileIter(a, blockIdx.x, gridDim.x).coord(a, BN);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.mapA0);
    tma_prefetch_desc(&a.mapA1);
    tma_prefetch_desc(&a.mapB);
    tma_prefetch_desc(&a.mapOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kLpPatches; ++i) {
      mbar_init(&p_full[i], 1);
      mbar_init(&p_empty[i], 1);
    }
    for (int i = 0; i < kLpBRing; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 8);  // eight epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<256>(&tmem_base_slot);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < BN; i += 384) s_bias[i] = a.bias[tc.n0 + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================================================================== patch producer (x_t, then h_{t-1})
    const uint32_t pf0 = smem_addr_once(&p_full[0]), pe0 = smem_addr_once(&p_empty[0]);
    const uint32_t sp0 = smem_addr_once(s_p);
    int ps = 0;
    uint32_t pphase = 0;
    for (int t = 0; t < T; ++t) {
      for (int src = 0; src < 2; ++src) {
        if (src == 1) {
          if (t == 0) break;
          const unsigned int target = static_cast<unsigned int>(t) * gridDim.x;
          if (lane == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(step_counter) < target) {
              if (clock64() - t0 > VAD_WAIT_TIMEOUT_CLOCKS) {
                if (g_vad_trap_slot) {
                  g_vad_trap_slot[0] = 12;
                  g_vad_trap_slot[1] = blockIdx.x;
                  g_vad_trap_slot[2] = static_cast<unsigned long long>(t);
                  g_vad_trap_slot[3] = target;
                  __threadfence_system();
                }
                __trap();
              }
            }
          }
          __syncwarp();
          fence_proxy_async_all();
        }
        const int n_chunks = src == 0 ? a.chunks0 : a.chunks1;
        const void* map = src == 0 ? static_cast<const void*>(&a.mapA0) : static_cast<const void*>(&a.mapA1);
        const int tt = src == 0 ? t : t - 1;
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait_a(pe0 + ps * 8, pphase ^ 1u, 1);
          if (elect_one()) {
            mbar_arrive_expect_tx_a(pf0 + ps * 8, patch_tx);
            if (geo1)  // map dims {C, W, B, H, T}
              tma_load_5d_a(sp0 + ps * kLpPatchPitch, map, pf0 + ps * 8, c * CK, tc.w0 - 1, tc.b0, tc.h0 - 1, tt);
            else       // map dims {C, W, H, T, B}
              tma_load_5d_a(sp0 + ps * kLpPatchPitch, map, pf0 + ps * 8, c * CK, tc.w0 - 1, tc.h0 - 1, tt, tc.b0);
          }
          __syncwarp();
          if (++ps == kLpPatches) { ps = 0; pphase ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================================== weight producer (independent of the steps)
    const uint32_t bf0 = smem_addr_once(&b_full[0]), be0 = smem_addr_once(&b_empty[0]);
    const uint32_t sb0 = smem_addr_once(s_b);
    int bs = 0;
    uint32_t bphase = 0;
    for (int t = 0; t < T; ++t) {
      for (int src = 0; src < 2; ++src) {
        if (src == 1 && t == 0) break;
        const int n_chunks = src == 0 ? a.chunks0 : a.chunks1;
        const int kbase = src == 0 ? 0 : a.chunks0 * CK;
        for (int c = 0; c < n_chunks; ++c) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait_a(be0 + bs * 8, bphase ^ 1u, 9);
            if (elect_one()) {
              mbar_arrive_expect_tx_a(bf0 + bs * 8, kBBytes);
              tma_load_2d_a(sb0 + bs * kBBytes, &a.mapB, bf0 + bs * 8, tap * a.w_ctap + kbase + c * CK, tc.n0);
            }
            __syncwarp();
            if (++bs == kLpBRing) { bs = 0; bphase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, BN);
    const uint32_t pf0 = smem_addr_once(&p_full[0]), pe0 = smem_addr_once(&p_empty[0]);
    const uint32_t bf0 = smem_addr_once(&b_full[0]), be0 = smem_addr_once(&b_empty[0]);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const uint64_t da_base = umma_smem_desc(smem_u32(s_p), 10 * kRowBytes, kLayout);  // 8-row groups 10 patch rows apart
    const uint64_t db_base = umma_smem_desc(smem_u32(s_b), 8 * kRowBytes, kLayout);
    int ps = 0, bs = 0;
    uint32_t pphase = 0, bphase = 0;
    for (int t = 0; t < T; ++t) {
      const int n_patches = a.chunks0 + (t > 0 ? a.chunks1 : 0);
      const int as = t & 1;
      mbar_wait_a(acce0 + as * 8, static_cast<uint32_t>(((t >> 1) & 1) ^ 1), 3);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      for (int p = 0; p < n_patches; ++p) {
        mbar_wait_a(pf0 + ps * 8, pphase, 2);
        const uint64_t da_p = da_base + static_cast<uint64_t>(ps * (kLpPatchPitch >> 4));
        int ky = 0, kx = 0;
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait_a(bf0 + bs * 8, bphase, 10);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = da_p + static_cast<uint64_t>(((ky * ky_rows + kx) * kRowBytes) >> 4);
            const uint64_t db = db_base + static_cast<uint64_t>(bs * (kBBytes >> 4));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc,
                        (p > 0 || tap > 0 || kk > 0) ? 1u : 0u);
            umma_commit_a(be0 + bs * 8);
            if (tap == 8) umma_commit_a(pe0 + ps * 8);
            if (tap == 8 && p == n_patches - 1) umma_commit_a(accf0 + as * 8);
          }
          __syncwarp();
          if (++bs == kLpBRing) { bs = 0; bphase ^= 1u; }
          if (++kx == 3) { kx = 0; ++ky; }
        }
        if (++ps == kLpPatches) { ps = 0; pphase ^= 1u; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================================== epilogue: gates, c (registers), h
    // The epilogue sits on the recurrence's critical path (h_t must be published before any neighbour can start the
    // h half of step t+1), so eight warps share it: warps 4-7 take hidden channels 0..15 of the tile, warps 8-11
    // channels 16..31 (a warp may read the TMEM lane quarter warp_id % 4, any columns).
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;  // 0 | 1: which 16 channels
    const EpiLane L = make_epi_lane(a, q, lane);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const bool leader = (warp == kEpiWarp0 && lane == 0);
    const int fb = tc.b0 + L.bb, h = tc.h0 + L.hh, w = tc.w0 + L.ww;
    const bool valid = L.row_ok && (fb < a.B) && (h < a.H) && (w < a.W);
    const int j0 = (tc.n0 >> 7) * 32;
    float c[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) c[e] = 0.f;
    for (int t = 0; t < T; ++t) {
      const int as = t & 1;
      mbar_wait_a(accf0 + as * 8, static_cast<uint32_t>((t >> 1) & 1), 4u | (static_cast<uint32_t>(t) << 8));
      tc_fence_after();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int s = half * 2 + s2;  // 8-channel group of the tile
        uint32_t gi[8], gf[8], gg[8], go[8];
        tmem_ld_x8(tacc + 0 + s * 8, gi);
        tmem_ld_x8(tacc + 32 + s * 8, gf);
        tmem_ld_x8(tacc + 64 + s * 8, gg);
        tmem_ld_x8(tacc + 96 + s * 8, go);
        tmem_ld_wait();
        if (s2 == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(acce0 + as * 8);
        }
        float hn[8];
        const float* bp = s_bias + s * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xi = __uint_as_float(gi[e]) + bp[e];
          const float xf = __uint_as_float(gf[e]) + bp[32 + e];
          const float xg = __uint_as_float(gg[e]) + bp[64 + e];
          const float xo = __uint_as_float(go[e]) + bp[96 + e];
          const float cn = sigmoid_fn(xf) * c[s2 * 8 + e] + sigmoid_fn(xi) * tanh_fn(xg);
          c[s2 * 8 + e] = cn;
          hn[e] = sigmoid_fn(xo) * tanh_fn(cn);
        }
        sts128(stg + staged_off(L.srow, s, 32), make_uint4(pack_bf16x2(hn[0], hn[1]), pack_bf16x2(hn[2], hn[3]),
                                                            pack_bf16x2(hn[4], hn[5]), pack_bf16x2(hn[6], hn[7])));
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (leader) {  // publishing h_t is on the critical path of every neighbour: do it before the group barrier
        tma_store_5d(&a.mapOut, stg, j0, tc.w0, tc.h0, t, tc.b0);
        bulk_commit_group();
        bulk_wait_group<0>();  // h_t of this tile is in global memory ...
        __threadfence();
        red_release_gpu_add(step_counter, 1u);  // ... and published
      }
      named_bar_sync(1, 256);  // nobody overwrites the staging buffer before the store has read it
    }
    if (valid && a.c_state != nullptr) {
      float* cptr = a.c_state + ((static_cast<long long>(fb) * a.H + h) * a.W + w) * a.cout + j0 + half * 16;
#pragma unroll
      for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(cptr + e) = make_float4(c[e], c[e + 1], c[e + 2], c[e + 3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------- ConvLSTM, two layers, wavefront
// Both layers of the video model's ConvLSTM (reference ConvLSTM.forward, models/video_autoencoder.py:153-167: layer-outer,
// time-inner loops) in one persistent launch.  Layer 2's step t only needs layer 1's h_t, so the two recurrences are
// independent chains one step apart; the CTA (one pixel tile x one 32-channel tile of BOTH layers) walks the flat item
// list  L1(0), L1(1), L2(0), L1(2), L2(1), ..., L1(T-1), L2(T-2), L2(T-1)  with the same roles as convlstm_patch_kernel.
// While one layer's serial chain (gates -> h store -> publish -> neighbours acquire -> patch load) is in flight the
// tensor pipe works on the other layer's item, which is what the single-layer kernel could not hide (9.3 us per step
// against 4.7 us of MMAs).  One step counter per layer; layer 2's x patches wait for layer 1's counter.
__device__ __forceinline__ void lstm2_item(int i, int T, int& layer, int& t) {
  if (i == 0) { layer = 0; t = 0; }
  else if (i == 2 * T - 1) { layer = 1; t = T - 1; }
  else if (i & 1) { layer = 0; t = (i + 1) >> 1; }
  else { layer = 1; t = (i >> 1) - 1; }
}

__device__ __forceinline__ void lstm_wait_counter(const unsigned int* counter, unsigned int target, int lane, int tag) {
  if (lane == 0) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(counter) < target) {
      if (clock64() - t0 > VAD_WAIT_TIMEOUT_CLOCKS) {
        if (g_vad_trap_slot) {
          g_vad_trap_slot[0] = static_cast<unsigned long long>(tag);
          g_vad_trap_slot[1] = blockIdx.x;
          g_vad_trap_slot[2] = target;
          g_vad_trap_slot[3] = ld_acquire_gpu(counter);
          __threadfence_system();
        }
        __trap();
      }
    }
  }
  __syncwarp();
  fence_proxy_async_all();  // the TMA loads that follow read what other CTAs published
}

__global__ void __launch_bounds__(384, 1) convlstm2_patch_kernel(const __grid_constant__ ConvArgs a1,
                                                                 const __grid_constant__ ConvArgs a2, int T,
                                                                 unsigned int* __restrict__ counters) {
  constexpr int BN = 128, CK = 64;
  constexpr int kRowBytes = 128;
  constexpr int kBBytes = BN * kRowBytes;  // 16 KB
  constexpr uint32_t kLayout = 2u;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t p_full[kLpPatches], p_empty[kLpPatches];
  __shared__ uint64_t b_full[kLpBRing], b_empty[kLpBRing];
  __shared__ uint64_t acc_full_bar[2];
  __shared__ uint64_t acc_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[2][BN];

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_b = smem;
  uint8_t* s_p = smem + kLpBRing * kBBytes;
  uint8_t* stg = s_p + kLpPatches * kLpPatchPitch;

  const bool geo1 = a1.row_perm == 2;
  const uint32_t patch_tx = static_cast<uint32_t>((geo1 ? 200 : 180) * kRowBytes);
  const int ky_rows = geo1 ? 20 : 10;  // patch rows per image row
  const TileCoord tc = TileIter(a1, blockIdx.x, gridDim.x).coord(a1, BN);
  const int n_items = 2 * T;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a1.mapA0); tma_prefetch_desc(&a1.mapA1); tma_prefetch_desc(&a1.mapB); tma_prefetch_desc(&a1.mapOut);
    tma_prefetch_desc(&a2.mapA0); tma_prefetch_desc(&a2.mapA1); tma_prefetch_desc(&a2.mapB); tma_prefetch_desc(&a2.mapOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kLpPatches; ++i) {
      mbar_init(&p_full[i], 1);
      mbar_init(&p_empty[i], 1);
    }
    for (int i = 0; i < kLpBRing; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 8);  // eight epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<256>(&tmem_base_slot);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 2 * BN; i += 384) s_bias[i >> 7][i & 127] = (i < BN ? a1.bias : a2.bias)[tc.n0 + (i & 127)];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================================================================== patch producer (x half, then h half)
    const uint32_t pf0 = smem_addr_once(&p_full[0]), pe0 = smem_addr_once(&p_empty[0]);
    const uint32_t sp0 = smem_addr_once(s_p);
    int ps = 0;
    uint32_t pphase = 0;
    for (int i = 0; i < n_items; ++i) {
      int layer, t;
      lstm2_item(i, T, layer, t);
      const ConvArgs& a = layer == 0 ? a1 : a2;
      for (int src = 0; src < 2; ++src) {
        if (src == 1 && t == 0) break;
        // layer 2's input x_t is layer 1's h_t: published once counter 0 has reached (t+1) * grid;
        // a layer's own h_{t-1}: its counter at t * grid
        if (src == 0 && layer == 1) lstm_wait_counter(counters, static_cast<unsigned int>(t + 1) * gridDim.x, lane, 13);
        if (src == 1) lstm_wait_counter(counters + layer, static_cast<unsigned int>(t) * gridDim.x, lane, 12);
        const int n_chunks = src == 0 ? a.chunks0 : a.chunks1;
        const void* map = src == 0 ? static_cast<const void*>(&a.mapA0) : static_cast<const void*>(&a.mapA1);
        const int tt = src == 0 ? t : t - 1;
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait_a(pe0 + ps * 8, pphase ^ 1u, 1);
          if (elect_one()) {
            mbar_arrive_expect_tx_a(pf0 + ps * 8, patch_tx);
            if (geo1)  // map dims {C, W, B, H, T}
              tma_load_5d_a(sp0 + ps * kLpPatchPitch, map, pf0 + ps * 8, c * CK, tc.w0 - 1, tc.b0, tc.h0 - 1, tt);
            else       // map dims {C, W, H, T, B}
              tma_load_5d_a(sp0 + ps * kLpPatchPitch, map, pf0 + ps * 8, c * CK, tc.w0 - 1, tc.h0 - 1, tt, tc.b0);
          }
          __syncwarp();
          if (++ps == kLpPatches) { ps = 0; pphase ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================================== weight producer (independent of the steps)
    const uint32_t bf0 = smem_addr_once(&b_full[0]), be0 = smem_addr_once(&b_empty[0]);
    const uint32_t sb0 = smem_addr_once(s_b);
    int bs = 0;
    uint32_t bphase = 0;
    for (int i = 0; i < n_items; ++i) {
      int layer, t;
      lstm2_item(i, T, layer, t);
      const ConvArgs& a = layer == 0 ? a1 : a2;
      for (int src = 0; src < 2; ++src) {
        if (src == 1 && t == 0) break;
        const int n_chunks = src == 0 ? a.chunks0 : a.chunks1;
        const int kbase = src == 0 ? 0 : a.chunks0 * CK;
        for (int c = 0; c < n_chunks; ++c) {
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait_a(be0 + bs * 8, bphase ^ 1u, 9);
            if (elect_one()) {
              mbar_arrive_expect_tx_a(bf0 + bs * 8, kBBytes);
              tma_load_2d_a(sb0 + bs * kBBytes, &a.mapB, bf0 + bs * 8, tap * a.w_ctap + kbase + c * CK, tc.n0);
            }
            __syncwarp();
            if (++bs == kLpBRing) { bs = 0; bphase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, BN);
    const uint32_t pf0 = smem_addr_once(&p_full[0]), pe0 = smem_addr_once(&p_empty[0]);
    const uint32_t bf0 = smem_addr_once(&b_full[0]), be0 = smem_addr_once(&b_empty[0]);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const uint64_t da_base = umma_smem_desc(smem_u32(s_p), 10 * kRowBytes, kLayout);  // 8-row groups 10 patch rows apart
    const uint64_t db_base = umma_smem_desc(smem_u32(s_b), 8 * kRowBytes, kLayout);
    int ps = 0, bs = 0;
    uint32_t pphase = 0, bphase = 0;
    for (int i = 0; i < n_items; ++i) {
      int layer, t;
      lstm2_item(i, T, layer, t);
      const ConvArgs& a = layer == 0 ? a1 : a2;
      const int n_patches = a.chunks0 + (t > 0 ? a.chunks1 : 0);
      const int as = i & 1;
      mbar_wait_a(acce0 + as * 8, static_cast<uint32_t>(((i >> 1) & 1) ^ 1), 3);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
      for (int p = 0; p < n_patches; ++p) {
        mbar_wait_a(pf0 + ps * 8, pphase, 2);
        const uint64_t da_p = da_base + static_cast<uint64_t>(ps * (kLpPatchPitch >> 4));
        int ky = 0, kx = 0;
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait_a(bf0 + bs * 8, bphase, 10);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = da_p + static_cast<uint64_t>(((ky * ky_rows + kx) * kRowBytes) >> 4);
            const uint64_t db = db_base + static_cast<uint64_t>(bs * (kBBytes >> 4));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc,
                        (p > 0 || tap > 0 || kk > 0) ? 1u : 0u);
            umma_commit_a(be0 + bs * 8);
            if (tap == 8) umma_commit_a(pe0 + ps * 8);
            if (tap == 8 && p == n_patches - 1) umma_commit_a(accf0 + as * 8);
          }
          __syncwarp();
          if (++bs == kLpBRing) { bs = 0; bphase ^= 1u; }
          if (++kx == 3) { kx = 0; ++ky; }
        }
        if (++ps == kLpPatches) { ps = 0; pphase ^= 1u; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================================== epilogue: gates, c (registers), h
    const int q = warp & 3;
    const int half = (warp - kEpiWarp0) >> 2;  // 0 | 1: which 16 channels
    const EpiLane L = make_epi_lane(a1, q, lane);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const bool leader = (warp == kEpiWarp0 && lane == 0);
    const int fb = tc.b0 + L.bb, h = tc.h0 + L.hh, w = tc.w0 + L.ww;
    const bool valid = L.row_ok && (fb < a1.B) && (h < a1.H) && (w < a1.W);
    const int j0 = (tc.n0 >> 7) * 32;
    float c1[16], c2[16];  // cell state of this lane's pixel, 16 channels, both layers
#pragma unroll
    for (int e = 0; e < 16; ++e) c1[e] = c2[e] = 0.f;
    for (int i = 0; i < n_items; ++i) {
      int layer, t;
      lstm2_item(i, T, layer, t);
      const int as = i & 1;
      mbar_wait_a(accf0 + as * 8, static_cast<uint32_t>((i >> 1) & 1), 4u | (static_cast<uint32_t>(i) << 8));
      tc_fence_after();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int s = half * 2 + s2;  // 8-channel group of the tile
        uint32_t gi[8], gf[8], gg[8], go[8];
        tmem_ld_x8(tacc + 0 + s * 8, gi);
        tmem_ld_x8(tacc + 32 + s * 8, gf);
        tmem_ld_x8(tacc + 64 + s * 8, gg);
        tmem_ld_x8(tacc + 96 + s * 8, go);
        tmem_ld_wait();
        if (s2 == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(acce0 + as * 8);
        }
        float hn[8];
        const float* bp = s_bias[layer] + s * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xi = __uint_as_float(gi[e]) + bp[e];
          const float xf = __uint_as_float(gf[e]) + bp[32 + e];
          const float xg = __uint_as_float(gg[e]) + bp[64 + e];
          const float xo = __uint_as_float(go[e]) + bp[96 + e];
          const float cprev = layer == 0 ? c1[s2 * 8 + e] : c2[s2 * 8 + e];
          const float cn = sigmoid_fn(xf) * cprev + sigmoid_fn(xi) * tanh_fn(xg);
          if (layer == 0) c1[s2 * 8 + e] = cn; else c2[s2 * 8 + e] = cn;
          hn[e] = sigmoid_fn(xo) * tanh_fn(cn);
        }
        sts128(stg + staged_off(L.srow, s, 32), make_uint4(pack_bf16x2(hn[0], hn[1]), pack_bf16x2(hn[2], hn[3]),
                                                            pack_bf16x2(hn[4], hn[5]), pack_bf16x2(hn[6], hn[7])));
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (leader) {  // publishing h_t is on the critical path of every neighbour: do it before the group barrier
        tma_store_5d(layer == 0 ? &a1.mapOut : &a2.mapOut, stg, j0, tc.w0, tc.h0, t, tc.b0);
        bulk_commit_group();
        bulk_wait_group<0>();  // h_t of this tile is in global memory ...
        __threadfence();
        red_release_gpu_add(counters + layer, 1u);  // ... and published
      }
      named_bar_sync(1, 256);  // nobody overwrites the staging buffer before the store has read it
    }
    if (valid) {
      const long long pix = (static_cast<long long>(fb) * a1.H + h) * a1.W + w;
      if (a1.c_state != nullptr) {
        float* cptr = a1.c_state + pix * a1.cout + j0 + half * 16;
#pragma unroll
        for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(cptr + e) = make_float4(c1[e], c1[e + 1], c1[e + 2], c1[e + 3]);
      }
      if (a2.c_state != nullptr) {
        float* cptr = a2.c_state + pix * a2.cout + j0 + half * 16;
#pragma unroll
        for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(cptr + e) = make_float4(c2[e], c2[e + 1], c2[e + 2], c2[e + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------- first conv
// 3 -> 32 channel 3x3 conv straight from the fp32 NCHW model input.  K = 27 (padded to 32).  TMA brings the fp32
// input patch of a tile (3 channels x 10 rows x 24 columns, zero-filled outside the frame = conv padding) into smem;
// four converter warps turn it into the bf16 im2col A tile [128 pixels][32] in swizzled smem (27 shared loads with
// immediate offsets per pixel); the MMA (two K=16 steps against the resident 32x32 weight tile) and the epilogue are
// the same as everywhere else.  Replaces models/autoencoder.py:39-41 and models/video_autoencoder.py:193-196 (with
// the 2x2 max-pool fused for the video encoder).
constexpr int kFirstStages = 8;
constexpr int kFirstGroupsC = 4;                // epilogue groups
constexpr int kConvWarps = 4;                   // converter warps; each converts whole tiles (every 4th)
constexpr int kFirstConvWarp0 = 4 + 4 * kFirstGroupsC;
constexpr int kFirstThreads = 32 * (kFirstConvWarp0 + kConvWarps);  // roles | epilogue groups | converters
// Every use of a ring slot must be waited for by the SAME warp (a warp that is two phases ahead of a barrier passes
// its parity wait spuriously): slot = tile % kFirstStages, converter = tile % kConvWarps, MMA issuer = tile % 2.
static_assert(kFirstStages % kConvWarps == 0 && kFirstStages % 2 == 0, "ring slots must map to fixed warps");
// fp32 patch: 24 columns x 10 rows x 3 channels starting at column w0-4: TMA needs the box's first byte 16-byte
// aligned in the innermost dimension, so the 1-pixel left halo is fetched as part of an aligned group of four
constexpr int kPatchW = 24, kPatchH = 10, kPatchX0 = 4;
constexpr int kPatchBytes = 3 * kPatchH * kPatchW * 4;         // 2880
constexpr int kPatchStride = 3072;                             // ring pitch

template <int EPI>
__global__ void __launch_bounds__(kFirstThreads, 1) conv_first_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int BN = 32;
  constexpr int kFirstGroups = kFirstGroupsC;
  constexpr int kAS = acc_stages_for(kFirstGroups);
  constexpr int kABytes = kTileM * 64;  // 128 rows x 32 bf16
  constexpr uint32_t kTmemCols = tmem_cols_for(BN, kFirstGroups);
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t patch_full[kFirstStages];
  __shared__ uint64_t patch_empty[kFirstStages];
  __shared__ uint64_t full_bar[kFirstStages];
  __shared__ uint64_t empty_bar[kFirstStages];
  __shared__ uint64_t acc_full_bar[kMaxAccStages];
  __shared__ uint64_t acc_empty_bar[kMaxAccStages];
  __shared__ uint64_t turn_bar[2];  // MMA issuer ping-pong token
  __shared__ uint32_t tmem_base_slot;
  __shared__ float red_smem[kMaxAccStages][4][3];
  __shared__ __align__(16) float s_bias[kMaxBias];

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;                                   // [32 n][32 k] bf16, 64B-swizzled (2 KB slot)
  uint8_t* s_a = smem + 2048;                            // ring of A tiles
  uint8_t* s_p = s_a + kFirstStages * kABytes;           // ring of fp32 input patches
  uint8_t* stg = s_p + kFirstStages * kPatchStride;      // epilogue staging (1024-aligned: 8*3072 = 24 KB)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.mapA0);
    if (a.tma_store) tma_prefetch_desc(&a.mapOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kFirstStages; ++i) {
      mbar_init(&patch_full[i], 1);
      mbar_init(&patch_empty[i], 1);  // one arrive by the converter warp that consumed the patch
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&turn_bar[0], 1);
    mbar_init(&turn_bar[1], 1);
    for (int i = 0; i < kMaxAccStages; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<kTmemCols>(&tmem_base_slot);
    tmem_relinquish();
  }
  // weights: global bf16 [32][32] row-major -> swizzled smem (16-byte chunks)
  if (threadIdx.x < 128) {
    const int n = threadIdx.x >> 2, c16 = threadIdx.x & 3;
    const uint4 v = reinterpret_cast<const uint4*>(a.w_first)[threadIdx.x];
    *reinterpret_cast<uint4*>(s_w + staged_off(n, c16, 32)) = v;
  }
  for (int i = threadIdx.x; i < BN; i += kFirstThreads) s_bias[i] = a.bias[i];
  fence_proxy_async_smem();  // s_w is read by the tensor core (async proxy)
  if (a.pdl) {  // programmatic dependent launch: everything above touched constants only
    pdl_launch_dependents();
    pdl_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0 || warp == 2) {
    // ===================================================================== TMA: fp32 input patches
    // two producer warps take alternate tiles (warp 2 is free once TMEM is allocated): the per-tile loop of a single
    // producer (tile walk, one barrier wait, one TMA issue: ~400 cycles of scalar latency) would otherwise cap the
    // whole kernel.  Ring slots alternate with the tiles, so each slot always belongs to the same producer.
    const int pi = warp == 0 ? 0 : 1;
    const uint32_t pfull0 = smem_addr_once(&patch_full[0]), pempty0 = smem_addr_once(&patch_empty[0]);
    const uint32_t sp0 = smem_addr_once(s_p);
    int it = pi;
    for (TileIter ti(a, blockIdx.x + pi * gridDim.x, 2 * gridDim.x); ti.tile < a.total_tiles; ti.next(a), it += 2) {
      const TileCoord t = ti.coord(a, BN);
      const int stage = it % kFirstStages;
      const uint32_t phase = (it / kFirstStages) & 1;
      mbar_wait_a(pempty0 + stage * 8, phase ^ 1u, 7);
      if (elect_one()) {
        mbar_arrive_expect_tx_a(pfull0 + stage * 8, kPatchBytes);
        tma_load_4d_a(sp0 + stage * kPatchStride, &a.mapA0, pfull0 + stage * 8, t.w0 - kPatchX0, t.h0 - 1, 0, t.b0);
      }
      __syncwarp();
    }
  } else if (warp >= kFirstConvWarp0) {
    // ===================================================================== converters: fp32 patch -> bf16 im2col rows
    // One warp converts a whole tile (8 rows x 16 columns).  A lane owns one column and four rows, {0,1,4,5} + 2*rg:
    // rows 2 apart are 48 floats apart in the patch (pitch 24), so the two half-warps read disjoint banks (scalar LDS,
    // one wavefront each), and at every store the eight lanes of a quarter warp write eight consecutive A rows, which
    // the 64-byte swizzle spreads over all bank groups: 72 + 64 shared-memory wavefronts per tile (the earlier
    // 4-pixels-in-a-row layout needed ~450 because of 4-way store conflicts).  K order k = (ky*3 + kx)*3 + ci.
    const int cw = warp - kFirstConvWarp0;
    const int col = lane & 15, rg = lane >> 4;
    const uint32_t pfull0 = smem_addr_once(&patch_full[0]), pempty0 = smem_addr_once(&patch_empty[0]);
    const uint32_t empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, BN);
    const uint64_t db = umma_smem_desc(smem_u32(s_w), 512, 4u);
    const uint64_t da_base = umma_smem_desc(smem_u32(s_a), 512, 4u);
    static_assert(kConvWarps == kAS, "tile -> converter warp -> accumulator stage -> epilogue group is one chain");
    int it = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
      if ((it & (kConvWarps - 1)) != cw) continue;
      const int stage = it % kFirstStages;
      const uint32_t phase = (it / kFirstStages) & 1;
      if (lane == 0) tl_stamp(a, 0, it, 0);
      mbar_wait_a(pfull0 + stage * 8, phase, 6);
      mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 1);
      if (lane == 0) tl_stamp(a, 0, it, 1);
      uint8_t* sa = s_a + stage * kABytes;
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        if (a.dbg & 256) break;  // ablation: no conversion work
        const int base_row = 4 * hb + 2 * rg;  // tile rows base_row, base_row + 1 <- patch rows base_row .. base_row + 3
        // patch column of pixel `col`, tap kx: col + (kPatchX0 - 1) + kx
        const float* pp = reinterpret_cast<const float*>(s_p + stage * kPatchStride) + base_row * kPatchW + col +
                          (kPatchX0 - 1);
        float f[3][4][3];  // [ci][patch row][kx]
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int pr = 0; pr < 4; ++pr)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) f[ci][pr][kx] = pp[ci * (kPatchH * kPatchW) + pr * kPatchW + kx];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float v[28];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
              for (int ci = 0; ci < 3; ++ci) v[(ky * 3 + kx) * 3 + ci] = f[ci][j + ky][kx];
          v[27] = 0.f;
          const int r = (base_row + j) * 16 + col;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            sts128(sa + staged_off(r, c, 32),
                   make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                              pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7])));
          sts128(sa + staged_off(r, 3, 32), make_uint4(pack_bf16x2(v[24], v[25]), pack_bf16x2(v[26], 0.f), 0u, 0u));
        }
      }
      fence_proxy_async_smem();  // every lane's A rows -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive_a(pempty0 + stage * 8);  // patch slot may be refilled
      // The converter issues its tile's two MMAs itself (tile -> converter warp -> accumulator stage -> epilogue group
      // is a fixed 1:1:1:1 chain, it % 4): four issuers instead of two and one hand-off less per tile — the separate
      // issuer warps' per-tile loop (two barrier waits, issue, two commits) was what capped this kernel.
      const int as = it % kAS;
      mbar_wait_a(acce0 + as * 8, static_cast<uint32_t>(((it / kAS) & 1) ^ 1), 3);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        const uint64_t da = da_base + static_cast<uint64_t>(stage * (kABytes >> 4));
        umma_bf16(d_tmem, da, db, idesc, 0u);
        umma_bf16(d_tmem, da + 2, db + 2, idesc, 1u);
        umma_commit_a(empty0 + stage * 8);   // A slot free once the MMAs have read it
        umma_commit_a(accf0 + as * 8);       // accumulator ready for the epilogue group
      }
      __syncwarp();
      if (lane == 0) tl_stamp(a, 0, it, 4);
    }
  } else if (warp >= kEpiWarp0) {
    epilogue_loop<BN, EPI, kFirstGroups>(a, tmem_base, warp, lane, stg, s_bias, red_smem, acc_full_bar, acc_empty_bar);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------- first conv + 2x2 max-pool, folded
// The video encoder's first layer (reference models/video_autoencoder.py:193-196: Conv2d(3,32,3,padding=1) + BN +
// LeakyReLU + MaxPool2d(2,2)) with the pooling window folded into the GEMM's N extent: one accumulator ROW is one POOLED
// output pixel, its 128 COLUMNS are the four conv outputs of the 2x2 window x 32 channels, and its K = 48 operand is the
// 4 x 4 x 3 input window those four convolutions read (weights zero outside each position's 3x3 sub-window):
//     D[pooled px][pos*32 + co] = sum_{wy,wx,ci} x[2py - 1 + wy][2px - 1 + wx][ci] * Wpf[pos*32 + co][(wy*4 + wx)*3 + ci]
//     out[pooled px][co] = act(max_pos D[.][pos*32 + co] + bias[co])
// Against conv_first_kernel<POOL> (one row per INPUT pixel, pooling by lane shuffles) this is, per 512 input pixels:
// 3 N=128 MMAs (192 cycles) instead of 8 N=32 ones (320), 6144 im2col values instead of 13824, and an epilogue that takes
// the maximum of four column groups inside the lane (96 FMNMX, no shuffles / selects) before bias + activation touch a
// quarter of the values — the SM-side work that bounded the old kernel (0.29 of the HBM peak at 720p) drops ~5x.
// Tile = 8 x 16 pooled pixels = 16 x 32 input pixels; TMA brings the fp32 patch (3 ch x 18 rows x 40 columns from column
// 32 tw - 4: 16-byte aligned box start, zero fill outside the frame = conv padding).
constexpr int kPfStages = 8;   // fp32 patch ring (TMA -> converter): two slots per converter warp hide the load latency
constexpr int kPfAStages = 4;  // A-tile ring (converter -> MMA): one slot per converter warp
constexpr int kPfGroups = 4;                  // epilogue groups == accumulator stages (128 TMEM columns each)
constexpr int kPfConvWarps = 4;               // converter warps; each converts whole tiles (every 4th) and issues their MMAs
constexpr int kPfConvWarp0 = 4 + 4 * kPfGroups;
constexpr int kPfThreads = 32 * (kPfConvWarp0 + kPfConvWarps);
constexpr int kPfPatchW = 40, kPfPatchH = 18, kPfPatchX0 = 4;
constexpr int kPfPatchBytes = 3 * kPfPatchH * kPfPatchW * 4;  // 8640
constexpr int kPfPatchPitch = 8704;                           // ring pitch (multiple of 128)
constexpr int kPfABytes = kTileM * 128;                       // 128 rows x 64 bf16 (48 used), SWIZZLE_128B
constexpr int kPfWBytes = 128 * 128;                          // [128 n][64 k] bf16
constexpr int kPfStgBytes = kTileM * 64;                      // one staged output tile: 128 pooled pixels x 32 ch bf16
constexpr int kPfSmemBytes = 1024 + kPfWBytes + kPfAStages * kPfABytes + kPfGroups * 2 * kPfStgBytes + kPfStages * kPfPatchPitch;
static_assert(kPfSmemBytes <= kSmemBudget, "conv_first_pool_kernel shared memory");
static_assert(kPfStages % kPfConvWarps == 0 && kPfStages % 2 == 0 && kPfAStages % kPfConvWarps == 0,
              "ring slots must map to fixed warps");

__global__ void __launch_bounds__(kPfThreads, 1) conv_first_pool_kernel(const __grid_constant__ ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t patch_full[kPfStages];
  __shared__ uint64_t patch_empty[kPfStages];
  __shared__ uint64_t empty_bar[kPfAStages];
  __shared__ uint64_t acc_full_bar[kPfGroups];
  __shared__ uint64_t acc_empty_bar[kPfGroups];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[32];

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w = smem;                              // [128 n][64 k] bf16, 128B-swizzled
  uint8_t* s_a = s_w + kPfWBytes;                   // ring of A tiles
  uint8_t* s_o = s_a + kPfAStages * kPfABytes;       // staged output tiles: two per epilogue group (1024-aligned)
  uint8_t* s_p = s_o + kPfGroups * 2 * kPfStgBytes;  // ring of fp32 input patches

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.mapA0);
    tma_prefetch_desc(&a.mapOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kPfStages; ++i) {
      mbar_init(&patch_full[i], 1);
      mbar_init(&patch_empty[i], 1);  // one arrive by the converter warp that consumed the patch
    }
    for (int i = 0; i < kPfAStages; ++i) mbar_init(&empty_bar[i], 1);  // A slot free: committed by the MMAs that read it
    for (int i = 0; i < kPfGroups; ++i) {
      mbar_init(&acc_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<512>(&tmem_base_slot);
    tmem_relinquish();
  }
  // weights: global bf16 [128][64] row-major -> swizzled smem (16-byte chunks)
  for (int i = threadIdx.x; i < 128 * 8; i += kPfThreads) {
    const int n = i >> 3, c16 = i & 7;
    *reinterpret_cast<uint4*>(s_w + staged_off(n, c16, 64)) = reinterpret_cast<const uint4*>(a.w_first)[i];
  }
  if (threadIdx.x < 32) s_bias[threadIdx.x] = a.bias[threadIdx.x];
  fence_proxy_async_smem();  // s_w is read by the tensor core (async proxy)
  if (a.pdl) {  // programmatic dependent launch: everything above touched constants only
    pdl_launch_dependents();
    pdl_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0 || warp == 2) {
    // ===================================================================== TMA: fp32 input patches (two producers, alternate tiles)
    const int pi = warp == 0 ? 0 : 1;
    const uint32_t pfull0 = smem_addr_once(&patch_full[0]), pempty0 = smem_addr_once(&patch_empty[0]);
    const uint32_t sp0 = smem_addr_once(s_p);
    int it = pi;
    for (TileIter ti(a, blockIdx.x + pi * gridDim.x, 2 * gridDim.x); ti.tile < a.total_tiles; ti.next(a), it += 2) {
      const int stage = it % kPfStages;
      const uint32_t phase = (it / kPfStages) & 1;
      mbar_wait_a(pempty0 + stage * 8, phase ^ 1u, 7);
      if (elect_one()) {
        if (a.dbg & 128) {  // ablation: no input loads
          mbar_arrive_a(pfull0 + stage * 8);
        } else {
          mbar_arrive_expect_tx_a(pfull0 + stage * 8, kPfPatchBytes);
          tma_load_4d_a(sp0 + stage * kPfPatchPitch, &a.mapA0, pfull0 + stage * 8, 32 * ti.tw - kPfPatchX0,
                        16 * ti.th - 1, 0, ti.tb);
        }
      }
      __syncwarp();
    }
  } else if (warp >= kPfConvWarp0) {
    // ===================================================================== converters: fp32 patch -> bf16 window rows
    // A lane owns one pooled column and four pooled rows {hb*2 + (lane >> 4)}: its 4 x 4 x 3 window is 12 rows of six
    // consecutive floats (three 8-byte loads each; a half-warp reads 32 consecutive floats: conflict-free), packed in K
    // order k = (wy*4 + wx)*3 + ci into six 16-byte chunks of A row pr*16 + pc (eight consecutive rows per quarter-warp
    // store: the 128-byte swizzle spreads them over all banks).
    const int cw = warp - kPfConvWarp0;
    const int pc = lane & 15, rg = lane >> 4;
    const uint32_t pfull0 = smem_addr_once(&patch_full[0]), pempty0 = smem_addr_once(&patch_empty[0]);
    const uint32_t empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t accf0 = smem_addr_once(&acc_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, 128);
    const uint64_t db = umma_smem_desc(smem_u32(s_w), 1024, 2u);
    const uint64_t da_base = umma_smem_desc(smem_u32(s_a), 1024, 2u);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++it) {
      if ((it & (kPfConvWarps - 1)) != cw) continue;
      const int stage = it % kPfStages;
      const uint32_t phase = (it / kPfStages) & 1;
      const int astage = it % kPfAStages;
      mbar_wait_a(pfull0 + stage * 8, phase, 6);
      mbar_wait_a(empty0 + astage * 8, static_cast<uint32_t>(((it / kPfAStages) & 1) ^ 1), 1);
      uint8_t* sa = s_a + astage * kPfABytes;
      const float* patch = reinterpret_cast<const float*>(s_p + stage * kPfPatchPitch);
#pragma unroll 1
      for (int hb = 0; hb < 4; ++hb) {
        if (a.dbg & 256) break;  // ablation: no conversion work
        const int pr = hb * 2 + rg;
        // window of pooled pixel (pr, pc): patch rows 2pr .. 2pr+3, patch columns 3 + 2pc .. 6 + 2pc
        const float* pp = patch + (2 * pr) * kPfPatchW + 2 + 2 * pc;
        float v[48];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int wy = 0; wy < 4; ++wy) {
            const float2* q2 = reinterpret_cast<const float2*>(pp + ci * (kPfPatchH * kPfPatchW) + wy * kPfPatchW);
            const float2 f0 = q2[0], f1 = q2[1], f2 = q2[2];
            v[(wy * 4 + 0) * 3 + ci] = f0.y;
            v[(wy * 4 + 1) * 3 + ci] = f1.x;
            v[(wy * 4 + 2) * 3 + ci] = f1.y;
            v[(wy * 4 + 3) * 3 + ci] = f2.x;
          }
        const int r = pr * 16 + pc;
#pragma unroll
        for (int c = 0; c < 6; ++c)
          sts128(sa + staged_off(r, c, 64),
                 make_uint4(pack_bf16x2(v[8 * c], v[8 * c + 1]), pack_bf16x2(v[8 * c + 2], v[8 * c + 3]),
                            pack_bf16x2(v[8 * c + 4], v[8 * c + 5]), pack_bf16x2(v[8 * c + 6], v[8 * c + 7])));
      }
      fence_proxy_async_smem();  // every lane's A rows -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive_a(pempty0 + stage * 8);  // patch slot may be refilled
      // the converter issues its tile's three MMAs itself (tile -> converter warp -> accumulator stage -> epilogue group
      // is a fixed 1:1:1:1 chain, it % 4)
      const int as = it % kPfGroups;
      mbar_wait_a(acce0 + as * 8, static_cast<uint32_t>(((it / kPfGroups) & 1) ^ 1), 3);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * 128);
        const uint64_t da = da_base + static_cast<uint64_t>(astage * (kPfABytes >> 4));
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)  // K = 48: the last 16 columns of the 64-wide rows are never read
          umma_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc, kk > 0 ? 1u : 0u);
        umma_commit_a(empty0 + astage * 8);  // A slot free once the MMAs have read it
        umma_commit_a(accf0 + as * 8);       // accumulator ready for the epilogue group
      }
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================================== epilogue groups: max over the window, bias, act
    const int g = (warp - kEpiWarp0) >> 2;
    const int q = warp & 3;  // TMEM lane quarter == warp_id % 4
    const int r = q * 32 + lane;  // accumulator row = pooled pixel (r >> 4, r & 15) of the tile = staged row
    const uint32_t accf = smem_addr_once(&acc_full_bar[g]), acce = smem_addr_once(&acc_empty_bar[g]);
    const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g * 128);
    // The pooled tile (128 pixels x 64 B) is staged in swizzled smem and written by ONE TMA store per tile (clipped by
    // the hardware at partial tiles): direct 16-byte stores at a 64-byte lane stride cost more than the whole rest of
    // the epilogue (ablation: 0.53 -> 0.34 ms at 720p without them).
    const bool leader = (q == 0 && lane == 0);
    const uint32_t bar_id = 1 + g;
    uint32_t ph = 0;
    int n = 0;
    for (TileIter ti(a, blockIdx.x + g * gridDim.x, kPfGroups * gridDim.x); ti.tile < a.total_tiles;
         ti.next(a), ph ^= 1u, ++n) {
      uint8_t* stg = s_o + (g * 2 + (n & 1)) * kPfStgBytes;
      mbar_wait_a(accf, ph, 4);
      tc_fence_after();
      if (a.dbg & 32) {  // ablation: no epilogue work at all
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(acce);
        continue;
      }
      if (leader) bulk_wait_group_read<1>();  // the store that last used this staging buffer has read it
      named_bar_sync(bar_id, 128);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t m[16], t0[16], t1[16];
        tmem_ld_x16(tacc + 0 * 32 + half * 16, m);
        tmem_ld_x16(tacc + 1 * 32 + half * 16, t0);
        tmem_ld_wait();
        tmem_ld_x16(tacc + 2 * 32 + half * 16, t1);
#pragma unroll
        for (int j = 0; j < 16; ++j) m[j] = __float_as_uint(fmaxf(__uint_as_float(m[j]), __uint_as_float(t0[j])));
        tmem_ld_wait();
        tmem_ld_x16(tacc + 3 * 32 + half * 16, t0);
#pragma unroll
        for (int j = 0; j < 16; ++j) m[j] = __float_as_uint(fmaxf(__uint_as_float(m[j]), __uint_as_float(t1[j])));
        tmem_ld_wait();
        if (half == 1) {  // every column of this stage has been read: hand it back to the converters
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(acce);
        }
        uint32_t p[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v0 = fmaxf(__uint_as_float(m[2 * j]), __uint_as_float(t0[2 * j])) + s_bias[half * 16 + 2 * j];
          const float v1 = fmaxf(__uint_as_float(m[2 * j + 1]), __uint_as_float(t0[2 * j + 1])) + s_bias[half * 16 + 2 * j + 1];
          p[j] = pack_bf16x2(act_fn(v0, a.slope), act_fn(v1, a.slope));
        }
        sts128(stg + staged_off(r, half * 2, 32), make_uint4(p[0], p[1], p[2], p[3]));
        sts128(stg + staged_off(r, half * 2 + 1, 32), make_uint4(p[4], p[5], p[6], p[7]));
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
      named_bar_sync(bar_id, 128);
      if (leader && !(a.dbg & 64)) {  // (ablation bit 64: no output store)
        tma_store_5d(&a.mapOut, stg, 0, 16 * ti.tw, 8 * ti.th, 0, ti.tb);
        bulk_commit_group();
      }
    }
    if (leader) bulk_wait_group_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------- enc1.0 -> enc1.3 (+ pool), fused (image encoder)
// The first block of the image encoder (reference models/autoencoder.py:38-45: Conv2d(3,32,3,p=1) + BN + LeakyReLU,
// Conv2d(32,32,3,p=1) + BN + LeakyReLU, MaxPool2d(2,2)) in ONE kernel: the 32-channel full-resolution tensor between
// the two convolutions (1.07 GB at batch 256, written by enc1.0 and read back by enc1.3) stays in shared memory.
// Per tile = 16 x 16 output pixels of enc1.3 (8 x 8 pooled pixels):
//   TMA      fp32 input patch, 3 ch x 20 rows x 24 columns from (16 th - 2, 16 tw - 4), zero fill outside the frame
//   convert  four warps write the bf16 im2col rows of the 18 x 20 = 360 pixels enc1.3 needs from enc1.0 (its 1-pixel
//            halo included): three [128 px][32 k] operands, K order k = (ky*3 + kx)*3 + ci as conv_first_kernel
//   stage A  D1[j][128 px][32 ch] = A1[j] · W1^T                                   (3 x 2 MMAs, N = 32)
//   ep A     bias + LeakyReLU, bf16 -> the PIXEL-PAIR patch of enc1.3's input in smem: 18 rows x 10 pairs x 128 B with
//            the SWIZZLE_128B pattern on absolute address bits; pixels outside the image are written as zeros (enc1.3's
//            padding)
//   stage B  the pixel-pair folded 3x3 conv of conv_halo_kernel<64, 64> (pair_fold): D2[16 rows x 8 pairs][2 x 32] from
//            nine row-shifted views of the patch against the resident pair weights, 24 MMAs (N = 64), into D1's columns
//   ep B     2x2 max-pool (horizontal half inside the lane, vertical half with the row-neighbour lane), bias,
//            LeakyReLU, 16-byte stores of the pooled bf16 NHWC tile.
// Both GEMM stages issue the same MMAs in the same order as the two single-layer kernels, and the intermediate is
// rounded to bf16 exactly where enc1.0 would have stored it: outputs are bit-identical to the two-layer path (tested).
constexpr int kE1Groups = 3;                       // pipeline stages: TMEM accumulator stage + pair patch per tile in flight
constexpr int kE1XStages = 4;                      // fp32 input patch ring
constexpr int kE1AStages = 2;                      // im2col operand ring (three A1 tiles per slot)
constexpr int kE1ConvWarps = 4;
// warps: 0 TMA, 1 stage-A issuer, 2 TMEM, 3 stage-B issuer | 4..15 epilogue A: set j = (warp - 4) / 4 reads operand j's
// accumulator of EVERY tile | 16..19 epilogue B of every tile | 20..23 converters.  (Fixed roles instead of one group
// of four warps walking a tile through both epilogues: a single warp runs such serial code at ~0.3 instructions per
// clock, so the group's chain A -> ep A (3 x 128 rows) -> B -> ep B took ~5000 cycles and three groups could not keep
// the tensor pipe busy.)
constexpr int kE1EpiBWarp0 = 4 + 12;
constexpr int kE1ConvWarp0 = kE1EpiBWarp0 + 4;
constexpr int kE1Threads = 32 * (kE1ConvWarp0 + kE1ConvWarps);
constexpr int kE1XW = 28, kE1XH = 20;              // (24 columns are used; 28 makes three patch rows 84 = 20 mod 32 floats,
                                                   //  so that the two row bands a converter warp reads hit disjoint banks)
constexpr int kE1XBytes = 3 * kE1XH * kE1XW * 4;   // 6720
constexpr int kE1XPitch = 6784;                    // multiple of 128
constexpr int kE1PatchW = 20, kE1PatchH = 18, kE1PatchPx = kE1PatchW * kE1PatchH;  // enc1.0 outputs per tile: 360
constexpr int kE1A1Bytes = 3 * kTileM * 64;        // three [128][32] bf16 tiles, SWIZZLE_64B: 24 KB
constexpr int kE1W1Bytes = 32 * 64;                // [32 n][32 k] bf16
constexpr int kE1W2Bytes = 9 * 64 * 128;           // nine [64 n][64 k] bf16 slabs, SWIZZLE_128B: 72 KB
constexpr int kE1PatchBytes = 23 * 1024;           // 18 x 10 pair rows x 128 B = 23040 -> 1024-aligned pitch
constexpr int kE1SmemBytes = 1024 + kE1W2Bytes + kE1W1Bytes + kE1AStages * kE1A1Bytes + kE1Groups * kE1PatchBytes +
                             kE1XStages * kE1XPitch;
static_assert(kE1SmemBytes <= kSmemBudget, "enc1_fused_kernel shared memory");

__global__ void __launch_bounds__(kE1Threads, 1) enc1_fused_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int G = kE1Groups;
#ifdef VAD_TIMELINE
  int tl_n = 0;  // tiles this role has walked (timeline row; tools/timeline_enc1.py)
#define E1_STAMP(role, ev, cond) do { if (cond) tl_stamp(a, role, tl_n, ev); } while (0)
#define E1_NEXT() (++tl_n)
#else
#define E1_STAMP(role, ev, cond) do { } while (0)
#define E1_NEXT() do { } while (0)
#endif
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_bar;
  __shared__ uint64_t x_full[kE1XStages];
  __shared__ uint64_t x_empty[kE1XStages];
  __shared__ uint64_t a1_full[kE1AStages];
  __shared__ uint64_t a1_empty[kE1AStages];
  __shared__ uint64_t d1_full_bar[G];
  __shared__ uint64_t a2_ready_bar[G];
  __shared__ uint64_t d2_full_bar[G];
  __shared__ uint64_t acc_empty_bar[G];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[64];  // [0,32) enc1.0, [32,64) enc1.3

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w2 = smem;                                  // pair weights, resident
  uint8_t* s_w1 = s_w2 + kE1W2Bytes;                     // first-conv weights [32][32], 64B-swizzled
  uint8_t* s_a1 = s_w1 + kE1W1Bytes;                     // im2col ring (1024-aligned: 72 KB + 2 KB)
  uint8_t* s_p = s_a1 + kE1AStages * kE1A1Bytes;         // pair patches, one per epilogue group (1024-aligned)
  uint8_t* s_x = s_p + G * kE1PatchBytes;                // fp32 input patch ring

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.mapA0);
    tma_prefetch_desc(&a.mapB);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < kE1XStages; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], kE1ConvWarps);
    }
    for (int i = 0; i < kE1AStages; ++i) {
      mbar_init(&a1_full[i], kE1ConvWarps);
      mbar_init(&a1_empty[i], 1);
    }
    for (int i = 0; i < G; ++i) {
      mbar_init(&d1_full_bar[i], 1);
      mbar_init(&a2_ready_bar[i], 12);  // every epilogue-A warp has written its rows of the patch
      mbar_init(&d2_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4);  // the four epilogue-B warps have read the stage
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<512>(&tmem_base_slot);
    tmem_relinquish();
  }
  if (threadIdx.x < 128) {  // first-conv weights: global bf16 [32][32] row-major -> swizzled smem (16-byte chunks)
    const int n = threadIdx.x >> 2, c16 = threadIdx.x & 3;
    *reinterpret_cast<uint4*>(s_w1 + staged_off(n, c16, 32)) = reinterpret_cast<const uint4*>(a.w_first)[threadIdx.x];
  }
  if (threadIdx.x < 64) s_bias[threadIdx.x] = threadIdx.x < 32 ? a.bias[threadIdx.x] : a.bias2[threadIdx.x - 32];
  fence_proxy_async_smem();  // s_w1 is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (warp == 0 && elect_one()) {  // pair weights: nine [64 x 64] slabs, once (constants: may be fetched before the PDL wait)
    mbar_arrive_expect_tx(&w_bar, static_cast<uint32_t>(kE1W2Bytes));
    for (int tap = 0; tap < 9; ++tap) tma_load_2d(s_w2 + tap * (64 * 128), &a.mapB, &w_bar, tap * 64, 0);
  }
  if (a.pdl) {
    pdl_launch_dependents();
    pdl_wait();
  }

  if (warp == 0) {
    // ===================================================================== TMA producer: fp32 input patches
    const uint32_t xf0 = smem_addr_once(&x_full[0]), xe0 = smem_addr_once(&x_empty[0]);
    const uint32_t sx0 = smem_addr_once(s_x);
    int stage = 0;
    uint32_t phase = 0;
    for (TileIter ti(a, blockIdx.x, gridDim.x); ti.tile < a.total_tiles; ti.next(a)) {
      mbar_wait_a(xe0 + stage * 8, phase ^ 1u, 1);
      if (elect_one()) {
        mbar_arrive_expect_tx_a(xf0 + stage * 8, kE1XBytes);
        tma_load_4d_a(sx0 + stage * kE1XPitch, &a.mapA0, xf0 + stage * 8, 16 * ti.tw - 4, 16 * ti.th - 2, 0, ti.tb);
      }
      __syncwarp();
      if (++stage == kE1XStages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp >= kE1ConvWarp0) {
    // ===================================================================== converters: fp32 patch -> bf16 im2col rows
    // Converter thread L = cw*32 + lane < 120 owns patch column c = L % 20 and the three patch rows 3*(L / 20) .. +2:
    // 5 input rows x 3 columns x 3 channels = 45 shared loads (explicit LDS) for its three im2col rows instead of 81.
    // Patch pixel (r, c) is image pixel (16 th - 1 + r, 16 tw - 2 + c); its tap (ky, kx) reads input patch element
    // (r + ky, c + 1 + kx); it becomes row p % 128 of operand p / 128, p = r*20 + c.  A warp's lanes read consecutive
    // floats of at most two row bands 84 floats apart: conflict-free; they write consecutive A rows: conflict-free.
    const int cw = warp - kE1ConvWarp0;
    const int cl = cw * 32 + lane;
    const int band = cl / kE1PatchW, cc0 = cl - band * kE1PatchW;
    const bool conv_on = cl < 6 * kE1PatchW;
    const uint32_t xf0 = smem_addr_once(&x_full[0]), xe0 = smem_addr_once(&x_empty[0]);
    const uint32_t af0 = smem_addr_once(&a1_full[0]), ae0 = smem_addr_once(&a1_empty[0]);
    const uint32_t sx_lane = smem_addr_once(s_x) + static_cast<uint32_t>((3 * band * kE1XW + cc0 + 1) * 4);
    int xs = 0, as = 0;
    uint32_t xph = 0, aph = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      E1_STAMP(0, 0, cl == 0);
      mbar_wait_a(xf0 + xs * 8, xph, 6);
      E1_STAMP(0, 1, cl == 0);
      mbar_wait_a(ae0 + as * 8, aph ^ 1u, 1);
      E1_STAMP(0, 2, cl == 0);
      uint8_t* sa = s_a1 + as * kE1A1Bytes;
      if (conv_on && !(a.dbg & 256)) {
        const uint32_t xa = sx_lane + static_cast<uint32_t>(xs * kE1XPitch);
        float f[3][5][3];  // [ci][input row][kx]
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int ir = 0; ir < 5; ++ir)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              f[ci][ir][kx] = lds_f32(xa + static_cast<uint32_t>((ci * (kE1XH * kE1XW) + ir * kE1XW + kx) * 4));
#pragma unroll
        for (int jr = 0; jr < 3; ++jr) {
          const int p = (3 * band + jr) * kE1PatchW + cc0;
          const int m = p & 127;
          uint8_t* st = sa + (p >> 7) * (kTileM * 64);
          float v[28];
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
              for (int ci = 0; ci < 3; ++ci) v[(ky * 3 + kx) * 3 + ci] = f[ci][jr + ky][kx];
          v[27] = 0.f;
#pragma unroll
          for (int cc = 0; cc < 3; ++cc)
            sts128(st + staged_off(m, cc, 32),
                   make_uint4(pack_bf16x2(v[8 * cc], v[8 * cc + 1]), pack_bf16x2(v[8 * cc + 2], v[8 * cc + 3]),
                              pack_bf16x2(v[8 * cc + 4], v[8 * cc + 5]), pack_bf16x2(v[8 * cc + 6], v[8 * cc + 7])));
          sts128(st + staged_off(m, 3, 32), make_uint4(pack_bf16x2(v[24], v[25]), pack_bf16x2(v[26], 0.f), 0u, 0u));
        }
      }
      fence_proxy_async_smem();  // every lane's A rows -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_a(af0 + as * 8);  // this warp's quarter of the three operands is in place
        mbar_arrive_a(xe0 + xs * 8);  // ... and it is done with the input patch
      }
      E1_STAMP(0, 3, cl == 0);
      E1_NEXT();
      if (++xs == kE1XStages) { xs = 0; xph ^= 1u; }
      if (++as == kE1AStages) { as = 0; aph ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================================================================== stage-A MMA issuer (enc1.0)
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, 32);
    const uint32_t af0 = smem_addr_once(&a1_full[0]), ae0 = smem_addr_once(&a1_empty[0]);
    const uint32_t d1f0 = smem_addr_once(&d1_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const uint64_t da_base = umma_smem_desc(smem_u32(s_a1), 512, 4u);
    const uint64_t db = umma_smem_desc(smem_u32(s_w1), 512, 4u);
    int as = 0, g = 0, jg = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      E1_STAMP(1, 0, lane == 0);
      mbar_wait_a(acce0 + g * 8, static_cast<uint32_t>(jg & 1) ^ 1u, 3);
      E1_STAMP(1, 1, lane == 0);
      mbar_wait_a(af0 + as * 8, aph, 2);
      E1_STAMP(1, 2, lane == 0);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const uint64_t da = da_base + static_cast<uint64_t>((as * kE1A1Bytes + j * (kTileM * 64)) >> 4);
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(g * 128 + j * 32);
          umma_bf16(d_tmem, da, db, idesc, 0u);
          umma_bf16(d_tmem, da + 2, db + 2, idesc, 1u);
        }
        umma_commit_a(ae0 + as * 8);
        umma_commit_a(d1f0 + g * 8);
      }
      __syncwarp();
      E1_STAMP(1, 3, lane == 0);
      E1_NEXT();
      if (++as == kE1AStages) { as = 0; aph ^= 1u; }
      if (++g == G) { g = 0; ++jg; }
    }
  } else if (warp == 3) {
    // ===================================================================== stage-B MMA issuer (pair-folded enc1.3)
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, 64);
    const uint32_t a2r0 = smem_addr_once(&a2_ready_bar[0]), d2f0 = smem_addr_once(&d2_full_bar[0]);
    const uint64_t da_hi = umma_smem_desc(0, 10 * 128, 2u);  // 8-pair row groups, 10 pair rows (one patch row) apart
    const uint64_t db0 = umma_smem_desc(smem_u32(s_w2), 1024, 2u);
    const uint32_t sp16 = (smem_u32(s_p) & 0x3FFFF) >> 4;
    int g = 0, jg = 0;
    mbar_wait(&w_bar, 0, 5);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      E1_STAMP(3, 0, lane == 0);
      mbar_wait_a(a2r0 + g * 8, static_cast<uint32_t>(jg & 1), 6);
      E1_STAMP(3, 1, lane == 0);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p16 = sp16 + static_cast<uint32_t>(g * (kE1PatchBytes >> 4));
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(g * 128);
        uint32_t acc = 0u;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          if ((a.dbg & 16) && tap > 0) break;  // ablation: one tap only
          const uint64_t da = da_hi | static_cast<uint64_t>(p16 + ((((tap / 3) * 10 + tap % 3) * 128) >> 4));
          const uint64_t db = db0 + static_cast<uint64_t>((tap * 64 * 128) >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if ((tap % 3 == 0 && kk < 2) || (tap % 3 == 2 && kk >= 2)) continue;  // structurally zero K steps
            umma_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc, acc);
            acc = 1u;
          }
        }
        umma_commit_a(d2f0 + g * 8);
      }
      __syncwarp();
      E1_STAMP(3, 2, lane == 0);
      E1_NEXT();
      if (++g == G) { g = 0; ++jg; }
    }
  } else if (warp >= kEpiWarp0 && warp < kE1EpiBWarp0) {
    // ===================================================================== epilogue A: enc1.0's bias + LeakyReLU -> pair patch
    const int j = (warp - kEpiWarp0) >> 2;  // which of the three [128 px][32 ch] accumulators this warp set reads
    const int q = warp & 3;                 // TMEM lane quarter == warp_id % 4
    const int p = j * 128 + q * 32 + lane;  // patch pixel of this thread (tile-invariant)
    const bool p_ok = p < kE1PatchPx;
    const int pr = p / kE1PatchW, pc = p - pr * kE1PatchW;
    const uint32_t poff0 = staged_off(pr * 10 + (pc >> 1), (pc & 1) * 4, 64);  // chunk 0 of this pixel's half pair row
    const int prow7 = (pr * 10 + (pc >> 1)) & 7;
    const uint32_t d1f0 = smem_addr_once(&d1_full_bar[0]), a2r0 = smem_addr_once(&a2_ready_bar[0]);
    const uint32_t tacc0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(j * 32);
    float bias[32];
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) bias[jj] = s_bias[jj];
    int g = 0;
    uint32_t ph = 0;
    for (TileIter ti(a, blockIdx.x, gridDim.x); ti.tile < a.total_tiles; ti.next(a)) {
      const int gy = 16 * ti.th - 1 + pr, gx = 16 * ti.tw - 2 + pc;  // image coordinates of this thread's patch pixel
      const bool inside = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
      uint8_t* patch = s_p + g * kE1PatchBytes;
      E1_STAMP(2, 0, warp == kEpiWarp0 && lane == 0);
      mbar_wait_a(d1f0 + g * 8, ph, 4);
      E1_STAMP(2, 1, warp == kEpiWarp0 && lane == 0);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_x32(tacc0 + static_cast<uint32_t>(g * 128), v);  // (warp-collective: also the lanes past pixel 359)
      tmem_ld_wait();
      E1_STAMP(2, 2, warp == kEpiWarp0 && lane == 0);
      if (p_ok && !(a.dbg & 32)) {
        uint32_t pk[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj)
          pk[jj] = pack_bf16x2(act_fn(__uint_as_float(v[2 * jj]) + bias[2 * jj], a.slope),
                               act_fn(__uint_as_float(v[2 * jj + 1]) + bias[2 * jj + 1], a.slope));
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          // chunk (pc & 1) * 4 + jj of the pair row, 128-byte swizzle: XOR the chunk index with the row's low three bits
          const uint32_t off = (poff0 & ~0x70u) | (((((pc & 1) * 4 + jj) ^ prow7) & 7) << 4);
          sts128(patch + off, inside ? make_uint4(pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3])
                                     : make_uint4(0u, 0u, 0u, 0u));  // zeros outside the image: enc1.3's padding
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a2r0 + g * 8);
      E1_STAMP(2, 3, warp == kEpiWarp0 && lane == 0);
      E1_NEXT();
      if (++g == G) { g = 0; ph ^= 1u; }
    }
  } else if (warp >= kE1EpiBWarp0 && warp < kE1ConvWarp0) {
    // ===================================================================== epilogue B: 2x2 max-pool, bias, LeakyReLU, store
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int hh = r >> 3, ww = r & 7;  // this accumulator row's (row, pixel pair) inside the 16 x 8 tile
    const bool up2 = (hh & 1) != 0;
    const uint32_t d2f0 = smem_addr_once(&d2_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const uint32_t tacc0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int Hp = a.H >> 1, Wp = a.W >> 1;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(a.out);
    int g = 0;
    uint32_t ph = 0;
    for (TileIter ti(a, blockIdx.x, gridDim.x); ti.tile < a.total_tiles; ti.next(a)) {
      const uint32_t tacc = tacc0 + static_cast<uint32_t>(g * 128);
      const int oh = 8 * ti.th + (hh >> 1), ow = 8 * ti.tw + ww;
      const bool ok = oh < Hp && ow < Wp && !(a.dbg & 64);
      __nv_bfloat16* dst0 = outp + ((static_cast<long long>(ti.tb) * Hp + oh) * Wp + ow) * 32 + (up2 ? 8 : 0);
      E1_STAMP(4, 0, r == 0);
      mbar_wait_a(d2f0 + g * 8, ph, 7);
      E1_STAMP(4, 1, r == 0);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v0[16], v1[16];
        tmem_ld_x16(tacc + c * 16, v0);
        tmem_ld_x16(tacc + 32 + c * 16, v1);
        tmem_ld_wait();
        if (c == 1) {  // the stage's TMEM columns (and its patch) are free for a later tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(acce0 + g * 8);
          E1_STAMP(4, 2, r == 0);
        }
        float gmax[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) gmax[jj] = fmaxf(__uint_as_float(v0[jj]), __uint_as_float(v1[jj]));
        float mm[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float recv = __shfl_xor_sync(0xffffffffu, up2 ? gmax[jj] : gmax[jj + 8], 8);
          mm[jj] = fmaxf(up2 ? gmax[jj + 8] : gmax[jj], recv);
        }
        const float4* b4 = reinterpret_cast<const float4*>(s_bias + 32 + c * 16 + (up2 ? 8 : 0));
        const float4 b0 = b4[0], b1 = b4[1];
        if (ok)
          *reinterpret_cast<uint4*>(dst0 + c * 16) = make_uint4(
              pack_bf16x2(act_fn(mm[0] + b0.x, a.slope), act_fn(mm[1] + b0.y, a.slope)),
              pack_bf16x2(act_fn(mm[2] + b0.z, a.slope), act_fn(mm[3] + b0.w, a.slope)),
              pack_bf16x2(act_fn(mm[4] + b1.x, a.slope), act_fn(mm[5] + b1.y, a.slope)),
              pack_bf16x2(act_fn(mm[6] + b1.z, a.slope), act_fn(mm[7] + b1.w, a.slope)));
      }
      E1_STAMP(4, 3, r == 0);
      E1_NEXT();
      if (++g == G) { g = 0; ph ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}
#undef E1_STAMP
#undef E1_NEXT

// ------------------------------------------------------------------------------- ConvT -> ConvT + tanh + score, fused
// The last two layers of the video decoder (reference models/video_autoencoder.py:252-259: ConvTranspose2d(64,32,2,2) +
// BN + ReLU, ConvTranspose2d(32,3,2,2) + Tanh) and the error reduction (:371-384) in ONE kernel.  A k2s2 transposed
// convolution has no halo: input pixel (h,w) alone determines the 4x4 block of output pixels (4h..4h+3, 4w..4w+3), so the
// 32-channel full-width intermediate (the largest tensor of the video model: 29.5 MB per 720p frame, written and read
// back by the layer-by-layer path) never has to exist in HBM.  Per tile of 128 input pixels:
//   stage 1  D1[128 px][4 taps x 32 ch] = A[128 x 64] · W1^T                     (4 MMAs, N = 128; A by TMA, W1 resident)
//   ep 1     TMEM -> registers -> +bias, ReLU, bf16 -> smem, one K-major SWIZZLE_64B [128 px][32 ch] operand per tap
//            (the same swizzled layout the staged TMA stores use — finding 1 of DESIGN.md §4: UMMA reads what TMA writes)
//   stage 2  D2[tap1][128 px][4 taps x 3 ch -> 16] = A2[tap1][128 x 32] · W2^T   (4 x 2 MMAs, N = 16; into D1's columns)
//   ep 2     TMEM -> registers -> tanh, (x - recon)^2 against the 4x4 block of the fp32 input (float4 per row and
//            channel, issued before stage 1 is even waited for), heat map float4 stores, per-tile sum / min / max.
// Roles: warp 0 TMA producer, warp 1 stage-1 issuer, warp 3 stage-2 issuer, warp 2 TMEM allocator, then kC2Groups
// epilogue groups of four warps; group g owns TMEM columns [128g, 128g+128), its own A2 buffers, and this CTA's tiles
// g, g+G, ...  Algorithmic HBM bytes per input pixel: 128 (A) + 192 (x) + 64 (heat) = 384 instead of 1408.
constexpr int kC2Groups = 3;
constexpr int kC2Stages = 6;
constexpr int kC2Threads = 128 + 128 * kC2Groups;
constexpr int kC2W1Bytes = 128 * 128;     // [4 taps x 32 ch][64 k] bf16, SWIZZLE_128B
constexpr int kC2W2Bytes = 1024;          // [4 taps x 3 ch -> 16 rows][32 k] bf16, SWIZZLE_64B
constexpr int kC2BiasBytes = 4096 + 1024 + 1024;  // the first ConvT's bias as a GEMM operand (convt_conv_score_kernel's
                                                  // scheme) + padding that keeps the input ring 1024-byte aligned
constexpr int kC2ABytes = kTileM * 128;   // one input tile: 128 pixels x 64 channels
constexpr int kC2TapBytes = kTileM * 64;  // one stage-2 operand: 128 pixels x 32 channels
constexpr int kC2SmemBytes = 1024 + kC2W1Bytes + kC2W2Bytes + kC2BiasBytes + kC2Stages * kC2ABytes + kC2Groups * 4 * kC2TapBytes;
static_assert((kC2W1Bytes + kC2W2Bytes + kC2BiasBytes) % 1024 == 0, "1024-byte aligned operands");
static_assert(kC2SmemBytes <= kSmemBudget, "convt2_score_kernel shared memory");

__global__ void __launch_bounds__(kC2Threads, 1) convt2_score_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int G = kC2Groups;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_bar;
  __shared__ uint64_t full_bar[kC2Stages];
  __shared__ uint64_t empty_bar[kC2Stages];
  __shared__ uint64_t d1_full_bar[G];   // stage-1 accumulator complete                  (MMA commit)
  __shared__ uint64_t a2_ready_bar[G];  // ep 1 done: A2 written, D1 columns read          (4 warps)
  __shared__ uint64_t d2_full_bar[G];   // stage-2 accumulators complete                   (MMA commit)
  __shared__ uint64_t acc_empty_bar[G];  // ep 2 has read D2: the group's columns are free  (4 warps)
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[128 + 16];

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w1 = smem;
  uint8_t* s_w2 = s_w1 + kC2W1Bytes;
  uint8_t* s_bb = s_w2 + kC2W2Bytes;   // bias operand B: [2][128][16 B]
  uint8_t* s_ba = s_bb + 4096;         // bias operand A: [2][8][16 B]
  uint8_t* s_a = s_w2 + kC2W2Bytes + kC2BiasBytes;
  uint8_t* s_a2 = s_a + kC2Stages * kC2ABytes;
  const int rows_valid = 1 << (a.lgTW + a.lgTH + a.lgTN);  // <= 128
  const uint32_t tx_bytes = static_cast<uint32_t>(rows_valid * 128);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.mapA0);
    tma_prefetch_desc(&a.mapA1);
    tma_prefetch_desc(&a.mapB);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < kC2Stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < G; ++i) {
      mbar_init(&d1_full_bar[i], 1);
      mbar_init(&a2_ready_bar[i], 4);
      mbar_init(&d2_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<512>(&tmem_base_slot);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 128 + 16; i += kC2Threads) s_bias[i] = i < 128 ? a.bias[i] : a.bias2[i - 128];
  if (threadIdx.x < 128) {  // D1 += ones · bias^T as one more K step of stage 1 (see convt_conv_score_kernel)
    const float b = a.bias[threadIdx.x];
    const __nv_bfloat16 hi = __float2bfloat16_rn(b);
    const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
    const uint32_t w = static_cast<uint32_t>(__bfloat16_as_ushort(hi)) | (static_cast<uint32_t>(__bfloat16_as_ushort(lo)) << 16);
    *reinterpret_cast<uint4*>(s_bb + threadIdx.x * 16) = make_uint4(w, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(s_bb + 2048 + threadIdx.x * 16) = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x < 16)
      *reinterpret_cast<uint4*>(s_ba + threadIdx.x * 16) = make_uint4(threadIdx.x < 8 ? 0x3F803F80u : 0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();  // the bias operands are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (warp == 0 && elect_one()) {  // both weight matrices, once (constants: may be fetched before the PDL wait)
    mbar_arrive_expect_tx(&w_bar, kC2W1Bytes + kC2W2Bytes);
    tma_load_2d(s_w1, &a.mapB, &w_bar, 0, 0);
    tma_load_2d(s_w2, &a.mapA1, &w_bar, 0, 0);
  }
  if (a.pdl) {
    pdl_launch_dependents();
    pdl_wait();
  }

  if (warp == 0) {
    // ===================================================================== TMA producer: one input tile per tile
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t sa0 = smem_addr_once(s_a);
    int stage = 0;
    uint32_t phase = 0;
    for (TileIter ti(a, blockIdx.x, gridDim.x); ti.tile < a.total_tiles; ti.next(a)) {
      const TileCoord t = ti.coord(a, 128);
      mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 1);
      if (elect_one()) {
        mbar_arrive_expect_tx_a(full0 + stage * 8, tx_bytes);
        tma_load_5d_a(sa0 + stage * kC2ABytes, &a.mapA0, full0 + stage * 8, 0, t.w0, t.h0, a.tA0, t.b0);
      }
      __syncwarp();
      if (++stage == kC2Stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================================================================== stage-1 MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, 128);
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t d1f0 = smem_addr_once(&d1_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const uint64_t da_base = umma_smem_desc(smem_u32(s_a), 1024, 2u);
    const uint64_t db = umma_smem_desc(smem_u32(s_w1), 1024, 2u);
    const uint64_t da_bias = umma_smem_desc_noswz(smem_u32(s_ba), 128, 0);     // every 8-row group: the same core matrix
    const uint64_t db_bias = umma_smem_desc_noswz(smem_u32(s_bb), 2048, 128);
    int stage = 0, g = 0, j = 0;
    uint32_t phase = 0;
    mbar_wait(&w_bar, 0, 5);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      mbar_wait_a(acce0 + g * 8, static_cast<uint32_t>(j & 1) ^ 1u, 3);
      mbar_wait_a(full0 + stage * 8, phase, 2);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = da_base + static_cast<uint64_t>(stage * (kC2ABytes >> 4));
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(g * 128);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc, kk > 0 ? 1u : 0u);
        umma_commit_a(empty0 + stage * 8);
        umma_bf16(d_tmem, da_bias, db_bias, idesc, 1u);  // + bias (constants only: after the input slot's release)
        umma_commit_a(d1f0 + g * 8);
      }
      __syncwarp();
      if (++stage == kC2Stages) { stage = 0; phase ^= 1u; }
      if (++g == G) { g = 0; ++j; }
    }
  } else if (warp == 3) {
    // ===================================================================== stage-2 MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, 16);
    const uint32_t a2r0 = smem_addr_once(&a2_ready_bar[0]), d2f0 = smem_addr_once(&d2_full_bar[0]);
    const uint64_t da_base = umma_smem_desc(smem_u32(s_a2), 512, 4u);
    const uint64_t db = umma_smem_desc(smem_u32(s_w2), 512, 4u);
    int g = 0, j = 0;
    mbar_wait(&w_bar, 0, 5);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      mbar_wait_a(a2r0 + g * 8, static_cast<uint32_t>(j & 1), 6);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 4; ++tap) {
          const uint64_t da = da_base + static_cast<uint64_t>(((g * 4 + tap) * kC2TapBytes) >> 4);
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(g * 128 + tap * 16);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            umma_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc, kk > 0 ? 1u : 0u);
        }
        umma_commit_a(d2f0 + g * 8);
      }
      __syncwarp();
      if (++g == G) { g = 0; ++j; }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================================== epilogue groups
    const int g = (warp - kEpiWarp0) >> 2;
    const int q = warp & 3;  // TMEM lane quarter == warp_id % 4
    const int r = q * 32 + lane;
    const EpiLane L = make_epi_lane(a, q, lane);
    const uint32_t d1f = smem_addr_once(&d1_full_bar[g]), a2r = smem_addr_once(&a2_ready_bar[g]);
    const uint32_t d2f = smem_addr_once(&d2_full_bar[g]), acce = smem_addr_once(&acc_empty_bar[g]);
    const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g * 128);
    uint8_t* my_a2 = s_a2 + g * 4 * kC2TapBytes;
    const bool relu = a.slope == 0.f;  // (the model's case: ReLU folded into the bf16 conversion)
    const int Wo = 4 * a.W;
    const long long plane = 16LL * a.H * a.W;
    const float* b2 = s_bias + 128;
    uint32_t ph = 0;
    for (TileIter ti(a, blockIdx.x + g * gridDim.x, G * gridDim.x); ti.tile < a.total_tiles; ti.next(a), ph ^= 1u) {
      const TileCoord t = ti.coord(a, 128);
      const int fb = t.b0 + L.bb, h = t.h0 + L.hh, w = t.w0 + L.ww;
      const bool valid = L.row_ok && (fb < a.B) && (h < a.H) && (w < a.W);
      const long long pix0 = static_cast<long long>(4 * h) * Wo + 4 * w;  // first pixel of this lane's 4x4 output block
      const long long xoff = static_cast<long long>(fb) * 3 * plane + pix0;
      // the 4x4x3 block of the model input this lane scores against: in flight while both GEMM stages run
      float4 xr[4][3];
#pragma unroll
      for (int ro = 0; ro < 4; ++ro)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          xr[ro][ch] = valid ? __ldg(reinterpret_cast<const float4*>(a.x + xoff + ch * plane + ro * Wo))
                             : make_float4(0.f, 0.f, 0.f, 0.f);

      // ---- ep 1: first ConvT's bias + ReLU, bf16, into the stage-2 operands (accumulator row r -> operand row r)
      mbar_wait_a(d1f, ph, 4);
      tc_fence_after();
#pragma unroll 1
      for (int tap = 0; tap < 4; ++tap) {
        uint32_t v[32];
        tmem_ld_x32(tacc + tap * 32, v);
        tmem_ld_wait();
        uint32_t p[16];  // (bias: added by the GEMM)
        if (relu) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj)
            p[jj] = pack_bf16x2_relu(__uint_as_float(v[2 * jj]), __uint_as_float(v[2 * jj + 1]));
        } else {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj)
            p[jj] = pack_bf16x2(act_fn(__uint_as_float(v[2 * jj]), a.slope), act_fn(__uint_as_float(v[2 * jj + 1]), a.slope));
        }
        uint8_t* buf = my_a2 + tap * kC2TapBytes;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          sts128(buf + staged_off(r, jj, 32), make_uint4(p[4 * jj], p[4 * jj + 1], p[4 * jj + 2], p[4 * jj + 3]));
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a2r);

      // ---- ep 2: second ConvT's bias + tanh, squared error against x, heat map, per-tile partials
      mbar_wait_a(d2f, ph, 7);
      tc_fence_after();
      float ssum = 0.f, smin = INFINITY, smax = -INFINITY;
#pragma unroll
      for (int d1i = 0; d1i < 2; ++d1i) {
        uint32_t v0[16], v1[16];  // first-layer taps (d1i, 0) and (d1i, 1): output columns 0-1 and 2-3 of rows 2*d1i, +1
        tmem_ld_x16(tacc + (d1i * 2 + 0) * 16, v0);
        tmem_ld_x16(tacc + (d1i * 2 + 1) * 16, v1);
        tmem_ld_wait();
        if (d1i == 1) {  // the group's TMEM columns are free for the next tile's stage 1
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(acce);
        }
#pragma unroll
        for (int d2i = 0; d2i < 2; ++d2i) {
          const int ro = d1i * 2 + d2i;
          float sq[4] = {0.f, 0.f, 0.f, 0.f};
          float rec[3][4];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const float xv[4] = {xr[ro][ch].x, xr[ro][ch].y, xr[ro][ch].z, xr[ro][ch].w};
#pragma unroll
            for (int cx = 0; cx < 4; ++cx) {
              const int idx = (d2i * 2 + (cx & 1)) * 3 + ch;
              const float acc = __uint_as_float((cx >> 1) ? v1[idx] : v0[idx]);
              rec[ch][cx] = tanh_fn(acc + b2[idx]);
              const float d = xv[cx] - rec[ch][cx];
              sq[cx] += d * d;
            }
          }
          if (valid) {
            const long long rowoff = pix0 + static_cast<long long>(ro) * Wo;
            if (a.recon) {
#pragma unroll
              for (int ch = 0; ch < 3; ++ch)
                *reinterpret_cast<float4*>(a.recon + static_cast<long long>(fb) * 3 * plane + ch * plane + rowoff) =
                    make_float4(rec[ch][0], rec[ch][1], rec[ch][2], rec[ch][3]);
            }
            if (a.heat)
              *reinterpret_cast<float4*>(a.heat + static_cast<long long>(fb) * plane + rowoff) =
                  make_float4(sq[0] * (1.f / 3.f), sq[1] * (1.f / 3.f), sq[2] * (1.f / 3.f), sq[3] * (1.f / 3.f));
            ssum += (sq[0] + sq[1]) + (sq[2] + sq[3]);
            smin = fminf(smin, fminf(fminf(sq[0], sq[1]), fminf(sq[2], sq[3])));
            smax = fmaxf(smax, fmaxf(fmaxf(sq[0], sq[1]), fmaxf(sq[2], sq[3])));
          }
        }
      }
      // fixed-order reduction (deterministic), as in the layer-by-layer score epilogue
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
        smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
      }
      if (lane == 0)
        *reinterpret_cast<float4*>(a.partials + (static_cast<long long>(t.m_tile) * 4 + q) * 4) =
            make_float4(ssum, smin * (1.f / 3.f), smax * (1.f / 3.f), 0.f);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------- ConvT -> conv3x3 + tanh + score, fused
// The last block of the image decoder (reference models/autoencoder.py:131-138: ConvTranspose2d(32,32,2,2) + BN + ReLU,
// Conv2d(32,3,3,padding=1) + Tanh) and the error reduction (:214-221) in ONE kernel: the 32-channel full-resolution
// intermediate (1.07 GB at batch 256, written by dec4.0 and read back by dec4.3) stays in shared memory.
// Per tile: an 8 x 16 block of input pixels (TMA, origin (7*th - 1, 15*tw - 1), zero fill outside) ->
//   stage 1  D1[128 px][4 taps x 32 ch] = A[128 x 32] · W1^T                      (2 MMAs, N = 128)
//   ep 1     bias + ReLU, bf16 -> a 16 x 32 pixel PATCH of the intermediate in smem (pixel p = y*32 + x at p*64 B with
//            the SWIZZLE_64B pattern on absolute address bits; pixels outside the image are written as zeros = the
//            3x3 conv's padding)
//   stage 2  four sub-tiles of 4 patch rows x 32 columns (128 consecutive patch pixels, dense 8-pixel row groups):
//            D2[s][p][kx*3 + co] = sum_ky A(patch + (4s + ky) rows) · W2[ky]^T — the kx-folded form of conv_kx_kernel
//            (3 x 2 MMAs each, N = 16), into D1's columns.  (Round 1 used five 16 x 8 sub-tiles at columns 6s: 30 MMAs
//            and five epilogue passes with 84 of 128 lanes scoring, strided 6-pixel global accesses.)
//   ep 2     out[y][x] = D2[x-1][kx=0] + D2[x][1] + D2[x+1][2] (two lane shuffles along the row a warp holds), tanh,
//            (x - recon)^2, heat, partials: 30 of 32 lanes score, 120-byte contiguous x / heat / recon accesses.
// Valid outputs of a tile: patch rows 1..14 x columns 1..30 = output rows [14*th - 1, 14*th + 13), columns
// [30*tw - 1, 30*tw + 29) — tiles overlap by one input pixel, 75-80 % of stage 1 and 82 % of stage 2 is useful work,
// which is cheap next to the 2.1 GB of HBM traffic saved per batch.
constexpr int kI2Groups = 4;
constexpr int kI2Stages = 7;                  // input-tile ring (8 until the bias operands took 5 KB)
constexpr int kI2Threads = 128 + 128 * kI2Groups;
constexpr int kI2W1Bytes = 128 * 64;          // [4 taps x 32 ch][32 k] bf16, SWIZZLE_64B
constexpr int kI2W2Bytes = 3 * 1024;          // three [16][32] slabs (ky), SWIZZLE_64B
constexpr int kI2BiasBytes = 4096 + 1024;     // the transposed conv's bias as a GEMM operand (no swizzle): B = [2 K
                                              // chunks][128 n][16 B] with k0 / k1 = bias hi / lo (bf16), A = one 8-row
                                              // core matrix of {1, 1, 0, ...} + a zero one, shared by all rows (SBO = 0)
constexpr int kI2ABytes = kTileM * 64;        // one input tile: 8 x 16 pixels x 32 channels
constexpr int kI2PatchBytes = 18 * 32 * 64;   // 16 patch rows + 2 rows only ever read into discarded accumulator rows
constexpr int kI2SmemBytes = 1024 + kI2W1Bytes + kI2W2Bytes + kI2BiasBytes + kI2Stages * kI2ABytes + kI2Groups * kI2PatchBytes;
static_assert(kI2SmemBytes <= kSmemBudget, "convt_conv_score_kernel shared memory");
static_assert((kI2W1Bytes + kI2W2Bytes + kI2BiasBytes) % 1024 == 0 && kI2PatchBytes % 1024 == 0, "1024-byte aligned operands");

__global__ void __launch_bounds__(kI2Threads, 1) convt_conv_score_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int G = kI2Groups;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t w_bar;
  __shared__ uint64_t full_bar[kI2Stages];
  __shared__ uint64_t empty_bar[kI2Stages];
  __shared__ uint64_t d1_full_bar[G];
  __shared__ uint64_t a2_ready_bar[G];
  __shared__ uint64_t d2_full_bar[G];
  __shared__ uint64_t acc_empty_bar[G];
  __shared__ uint32_t tmem_base_slot;
  __shared__ __align__(16) float s_bias[128 + 16];

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w1 = smem;
  uint8_t* s_w2 = s_w1 + kI2W1Bytes;
  uint8_t* s_bb = s_w2 + kI2W2Bytes;   // bias operand B: [2][128][16 B]
  uint8_t* s_ba = s_bb + 4096;         // bias operand A: [2][8][16 B]
  uint8_t* s_a = s_ba + 1024;
  uint8_t* s_p = s_a + kI2Stages * kI2ABytes;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.mapA0);
    tma_prefetch_desc(&a.mapA1);
    tma_prefetch_desc(&a.mapB);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(&w_bar, 1);
    for (int i = 0; i < kI2Stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < G; ++i) {
      mbar_init(&d1_full_bar[i], 1);
      mbar_init(&a2_ready_bar[i], 4);
      mbar_init(&d2_full_bar[i], 1);
      mbar_init(&acc_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<512>(&tmem_base_slot);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 128 + 16; i += kI2Threads) s_bias[i] = i < 128 ? a.bias[i] : a.bias2[i - 128];
  if (threadIdx.x < 128) {
    // D1 += ones · bias^T as a third K step of stage 1: bias[n] = hi + lo in bf16 (exact to 2^-17 relative) against two
    // columns of ones — epilogue 1 loses its 128 FADDs and 32 bias loads per lane (it is instruction-issue-bound)
    const float b = a.bias[threadIdx.x];
    const __nv_bfloat16 hi = __float2bfloat16_rn(b);
    const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
    const uint32_t w = static_cast<uint32_t>(__bfloat16_as_ushort(hi)) | (static_cast<uint32_t>(__bfloat16_as_ushort(lo)) << 16);
    *reinterpret_cast<uint4*>(s_bb + threadIdx.x * 16) = make_uint4(w, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(s_bb + 2048 + threadIdx.x * 16) = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x < 16)
      *reinterpret_cast<uint4*>(s_ba + threadIdx.x * 16) = make_uint4(threadIdx.x < 8 ? 0x3F803F80u : 0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();  // the bias operands are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (warp == 0 && elect_one()) {  // weights, once (constants: may be fetched before the PDL wait)
    mbar_arrive_expect_tx(&w_bar, kI2W1Bytes + kI2W2Bytes);
    tma_load_2d(s_w1, &a.mapB, &w_bar, 0, 0);
    for (int ky = 0; ky < 3; ++ky) tma_load_2d(s_w2 + ky * 1024, &a.mapA1, &w_bar, ky * 32, 0);
  }
  if (a.pdl) {
    pdl_launch_dependents();
    pdl_wait();
  }

  if (warp == 0) {
    // ===================================================================== TMA producer: one 8 x 16 input block per tile
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t sa0 = smem_addr_once(s_a);
    int stage = 0;
    uint32_t phase = 0;
    for (TileIter ti(a, blockIdx.x, gridDim.x); ti.tile < a.total_tiles; ti.next(a)) {
      mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 1);
      if (elect_one()) {
        mbar_arrive_expect_tx_a(full0 + stage * 8, kI2ABytes);
        tma_load_5d_a(sa0 + stage * kI2ABytes, &a.mapA0, full0 + stage * 8, 0, 15 * ti.tw - 1, 7 * ti.th - 1, 0, ti.tb);
      }
      __syncwarp();
      if (++stage == kI2Stages) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    // ===================================================================== stage-1 MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, 128);
    const uint32_t full0 = smem_addr_once(&full_bar[0]), empty0 = smem_addr_once(&empty_bar[0]);
    const uint32_t d1f0 = smem_addr_once(&d1_full_bar[0]), acce0 = smem_addr_once(&acc_empty_bar[0]);
    const uint64_t da_base = umma_smem_desc(smem_u32(s_a), 512, 4u);
    const uint64_t db = umma_smem_desc(smem_u32(s_w1), 512, 4u);
    const uint64_t da_bias = umma_smem_desc_noswz(smem_u32(s_ba), 128, 0);     // every 8-row group: the same core matrix
    const uint64_t db_bias = umma_smem_desc_noswz(smem_u32(s_bb), 2048, 128);
    int stage = 0, g = 0, j = 0;
    uint32_t phase = 0;
    mbar_wait(&w_bar, 0, 5);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      mbar_wait_a(acce0 + g * 8, static_cast<uint32_t>(j & 1) ^ 1u, 3);
      mbar_wait_a(full0 + stage * 8, phase, 2);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = da_base + static_cast<uint64_t>(stage * (kI2ABytes >> 4));
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(g * 128);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          umma_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc, kk > 0 ? 1u : 0u);
        umma_commit_a(empty0 + stage * 8);
        umma_bf16(d_tmem, da_bias, db_bias, idesc, 1u);  // + bias (reads constants only: after the input slot's release)
        umma_commit_a(d1f0 + g * 8);
      }
      __syncwarp();
      if (++stage == kI2Stages) { stage = 0; phase ^= 1u; }
      if (++g == G) { g = 0; ++j; }
    }
  } else if (warp == 3) {
    // ===================================================================== stage-2 MMA issuer (kx-folded 3x3 conv)
    constexpr uint32_t idesc = umma_idesc_bf16_f32(kTileM, 16);
    const uint32_t a2r0 = smem_addr_once(&a2_ready_bar[0]), d2f0 = smem_addr_once(&d2_full_bar[0]);
    const uint64_t da_hi = umma_smem_desc(0, 8 * 64, 4u);  // dense: 8-pixel groups follow each other along the patch rows
    const uint64_t db0 = umma_smem_desc(smem_u32(s_w2), 512, 4u);
    const uint32_t sp16 = (smem_u32(s_p) & 0x3FFFF) >> 4;
    int g = 0, j = 0;
    mbar_wait(&w_bar, 0, 5);
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      mbar_wait_a(a2r0 + g * 8, static_cast<uint32_t>(j & 1), 6);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t p16 = sp16 + static_cast<uint32_t>(g * (kI2PatchBytes >> 4));
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          // sub-tile s = patch rows 4s .. 4s+3, all 32 columns (128 consecutive patch pixels); tap ky starts ky rows down
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(g * 128 + s * 16);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint64_t da = da_hi | static_cast<uint64_t>(p16 + ((((4 * s + ky) * 32) * 64) >> 4));
            const uint64_t db = db0 + static_cast<uint64_t>((ky * 1024) >> 4);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              if ((a.dbg & 16) && (ky > 0 || kk > 0)) continue;  // ablation: one MMA per sub-tile
              umma_bf16(d_tmem, da + static_cast<uint64_t>(kk * 2), db + static_cast<uint64_t>(kk * 2), idesc,
                        (ky > 0 || kk > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit_a(d2f0 + g * 8);
      }
      __syncwarp();
      if (++g == G) { g = 0; ++j; }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================================================================== epilogue groups
    const int g = (warp - kEpiWarp0) >> 2;
    const int q = warp & 3;  // TMEM lane quarter == warp_id % 4
    const int r = q * 32 + lane;
    const uint32_t d1f = smem_addr_once(&d1_full_bar[g]), a2r = smem_addr_once(&a2_ready_bar[g]);
    const uint32_t d2f = smem_addr_once(&d2_full_bar[g]), acce = smem_addr_once(&acc_empty_bar[g]);
    const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g * 128);
    uint8_t* patch = s_p + g * kI2PatchBytes;
    const int Ho = 2 * a.H, Wo = 2 * a.W;
    const long long plane = static_cast<long long>(Ho) * Wo;
    const int plane_i = Ho * Wo;
    const bool want_recon = a.recon != nullptr, want_heat = a.heat != nullptr;
    const float* b2 = s_bias + 128;
    const bool relu = a.slope == 0.f;     // (the model's case: ReLU folded into the bf16 conversion)
    const int iy = r >> 4, ix = r & 15;   // ep 1: this accumulator row's input pixel inside the 8 x 16 block
    // ep 2: sub-tile s covers patch rows 4s .. 4s+3 x all 32 columns, so warp quarter q holds row 4s + q and a lane one
    // column: 30 of 32 lanes score a pixel (the kx fold reaches two lanes to the right) and x / heat / recon accesses
    // are 120-byte contiguous per warp
    uint32_t ph = 0;
    for (TileIter ti(a, blockIdx.x + g * gridDim.x, G * gridDim.x); ti.tile < a.total_tiles; ti.next(a), ph ^= 1u) {
      const int oy0 = 14 * ti.th - 2, ox0 = 30 * ti.tw - 2;  // output coordinates of patch pixel (0, 0)
      const int fb = ti.tb;
      const int m_tile = (ti.tb * a.tiles_h + ti.th) * a.tiles_w + ti.tw;
      // the model-input pixels this lane scores in the four sub-tiles (in flight while both GEMM stages run)
      // (one 64-bit pixel offset per tile; everything else is 32-bit index arithmetic from there — the host checks that
      // a frame's three planes fit an int — and the validity tests are done once, as a bit mask)
      const int oy = oy0 + 1 + q, ox = ox0 + 1 + lane;  // this lane's output pixel in sub-tile s: (oy + 4 s, ox)
      const bool col_ok = lane < 30 && ox >= 0 && ox < Wo;
      const long long pix = static_cast<long long>(oy) * Wo + ox;             // (never dereferenced when !ok)
      const float* xt = a.x + static_cast<long long>(fb) * 3 * plane + pix;
      uint32_t okmask = 0;
#pragma unroll
      for (int s = 0; s < 4; ++s)
        okmask |= (col_ok && 4 * s + q < 14 && oy + 4 * s >= 0 && oy + 4 * s < Ho) ? (1u << s) : 0u;
      if (a.dbg & 128) okmask = 0;  // ablation: no x loads (and nothing scored)
      float xs[4][3];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) xs[s][ch] = ((okmask >> s) & 1u) ? __ldg(xt + ch * plane_i + 4 * s * Wo) : 0.f;
      }

      // ---- ep 1: transposed conv's bias + ReLU -> bf16 patch (zeros outside the image: the 3x3 conv's padding).
      // Eight 16-column pieces (half a tap each), the next piece's TMEM load in flight while this one is packed.
      mbar_wait_a(d1f, ph, 4);
      tc_fence_after();
      {
        uint32_t va[16], vb[16];
        tmem_ld_x16(tacc, va);
#pragma unroll
        for (int hp = 0; hp < 8; ++hp) {
          uint32_t (&v)[16] = (hp & 1) ? vb : va;
          uint32_t (&vn)[16] = (hp & 1) ? va : vb;
          tmem_ld_wait();
          if (hp < 7) tmem_ld_x16(tacc + (hp + 1) * 16, vn);
          const int tap = hp >> 1;
          const int py = 2 * iy + (tap >> 1), px = 2 * ix + (tap & 1);
          const int gy = oy0 + py, gx = ox0 + px;
          const bool inside = gy >= 0 && gy < Ho && gx >= 0 && gx < Wo;
          uint32_t p[8];  // (bias: added by the GEMM)
          if (relu) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              p[jj] = pack_bf16x2_relu(__uint_as_float(v[2 * jj]), __uint_as_float(v[2 * jj + 1]));
          } else {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              p[jj] = pack_bf16x2(act_fn(__uint_as_float(v[2 * jj]), a.slope), act_fn(__uint_as_float(v[2 * jj + 1]), a.slope));
          }
          const int pp = py * 32 + px;
          if (a.dbg & 32) continue;  // ablation: no patch writes
#pragma unroll
          for (int jj = 0; jj < 2; ++jj)
            sts128(patch + staged_off(pp, (hp & 1) * 2 + jj, 32),
                   inside ? make_uint4(p[4 * jj], p[4 * jj + 1], p[4 * jj + 2], p[4 * jj + 3]) : make_uint4(0u, 0u, 0u, 0u));
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a2r);

      // ---- ep 2: fold the horizontal taps, bias + tanh, squared error against x, heat map, per-tile partials
      mbar_wait_a(d2f, ph, 7);
      tc_fence_after();
      float ssum = 0.f, smin = INFINITY, smax = -INFINITY;
      uint32_t wa[16], wb[16];
      tmem_ld_x16(tacc, wa);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        uint32_t (&v)[16] = (s & 1) ? wb : wa;
        uint32_t (&vn)[16] = (s & 1) ? wa : wb;
        tmem_ld_wait();
        if (s < 3) tmem_ld_x16(tacc + (s + 1) * 16, vn);  // next sub-tile's accumulator in flight during this one's math
        if (s == 3) {  // the group's TMEM columns are free for the next tile's stage 1
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(acce);
        }
        const bool ok = ((okmask >> s) & 1u) != 0u && !(a.dbg & 64);  // (ablation bit 64: no stores / reduction)
        if (a.dbg & 256) continue;  // ablation: no ep 2 math
        float sq = 0.f;
        float rec[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float s1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[3 + ch]), 1);
          const float s2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[6 + ch]), 2);
          rec[ch] = tanh_fn(((__uint_as_float(v[ch]) + s1) + s2) + b2[ch]);
          const float d = xs[s][ch] - rec[ch];
          sq += d * d;
        }
        if (ok) {
          if (want_recon) {
            float* rt = a.recon + static_cast<long long>(fb) * 3 * plane + pix;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) rt[ch * plane_i + 4 * s * Wo] = rec[ch];
          }
          if (want_heat) a.heat[static_cast<long long>(fb) * plane + pix + 4 * s * Wo] = sq * (1.f / 3.f);
          ssum += sq;
          smin = fminf(smin, sq);
          smax = fmaxf(smax, sq);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        smin = fminf(smin, __shfl_xor_sync(0xffffffffu, smin, o));
        smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
      }
      if (lane == 0)
        *reinterpret_cast<float4*>(a.partials + (static_cast<long long>(m_tile) * 4 + q) * 4) =
            make_float4(ssum, smin * (1.f / 3.f), smax * (1.f / 3.f), 0.f);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Launch with (a.pdl != 0) or without the programmatic-stream-serialisation attribute.  With it the kernel may start
// while its predecessor in the stream is still draining; every kernel launched this way runs its prologue (barrier
// init, TMEM allocation, descriptor prefetch, bias load — constants only) and then `griddepcontrol.wait`s before it
// touches anything the predecessor wrote.
// Per-device launch state.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device (per-context) setting, so
// the "already configured" caches are indexed by the current device; a second GPU used from the same process gets its
// own configuration call instead of failing its first >48 KB launch.
constexpr int kMaxDevices = 64;
static int current_device_index() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}
struct SmemConfig {
  std::atomic<int> bytes[kMaxDevices];  // largest dynamic shared-memory size configured so far (zero-initialised statics)
};
template <typename Kernel>
static int ensure_smem(Kernel kernel, SmemConfig& c, int smem) {
  std::atomic<int>& cur = c.bytes[current_device_index()];
  if (smem <= cur.load(std::memory_order_acquire)) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return static_cast<int>(e);
  int prev = cur.load(std::memory_order_relaxed);
  while (prev < smem && !cur.compare_exchange_weak(prev, smem, std::memory_order_release)) {}
  return 0;
}

// Cooperative launch: the driver either makes every CTA of the grid resident at once or fails the launch
// (cudaErrorCooperativeLaunchTooLarge); two such grids on different streams are serialised instead of interleaved.  The
// persistent ConvLSTM kernels spin on a grid-wide step counter, which is only safe under that guarantee.
template <typename Kernel, typename... Args>
static int launch_cooperative(Kernel kernel, int grid, int block, int smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  count_launch();
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e == cudaErrorCooperativeLaunchTooLarge) {
    (void)cudaGetLastError();  // not sticky: the caller falls back to one launch per step
    return VAD_ERR_UNSUPPORTED;
  }
  return static_cast<int>(e);
}

template <typename Kernel>
static int launch_conv_kernel(Kernel kernel, const ConvArgs& a, int grid, int block, int smem, cudaStream_t stream) {
  count_launch();
  if (a.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return static_cast<int>(cudaLaunchKernelEx(&cfg, kernel, a));
  }
  kernel<<<grid, block, smem, stream>>>(a);
  return static_cast<int>(cudaGetLastError());
}

template <int EPI>
static int launch_first_one(const ConvArgs& a, int grid, cudaStream_t stream) {
  constexpr int smem = 1024 + 2048 + kFirstStages * (kTileM * 64 + kPatchStride) + staging_bytes(32, EPI, kFirstGroupsC);
  static SmemConfig cfg;
  if (int e = ensure_smem(conv_first_kernel<EPI>, cfg, smem)) return e;
  return launch_conv_kernel(conv_first_kernel<EPI>, a, grid, kFirstThreads, smem, stream);
}

int launch_convt2_score(const ConvArgs& a, int grid, cudaStream_t stream) {
  static SmemConfig cfg;
  if (int e = ensure_smem(convt2_score_kernel, cfg, kC2SmemBytes)) return e;
  return launch_conv_kernel(convt2_score_kernel, a, grid, kC2Threads, kC2SmemBytes, stream);
}

int launch_convt_conv_score(const ConvArgs& a, int grid, cudaStream_t stream) {
  static SmemConfig cfg;
  if (int e = ensure_smem(convt_conv_score_kernel, cfg, kI2SmemBytes)) return e;
  return launch_conv_kernel(convt_conv_score_kernel, a, grid, kI2Threads, kI2SmemBytes, stream);
}

int launch_conv_first_pool(const ConvArgs& a, int grid, cudaStream_t stream) {
  static SmemConfig cfg;
  if (int e = ensure_smem(conv_first_pool_kernel, cfg, kPfSmemBytes)) return e;
  return launch_conv_kernel(conv_first_pool_kernel, a, grid, kPfThreads, kPfSmemBytes, stream);
}

int launch_enc1_fused(const ConvArgs& a, int grid, cudaStream_t stream) {
  static SmemConfig cfg;
  if (int e = ensure_smem(enc1_fused_kernel, cfg, kE1SmemBytes)) return e;
  return launch_conv_kernel(enc1_fused_kernel, a, grid, kE1Threads, kE1SmemBytes, stream);
}

int launch_conv_first(int EPI, const ConvArgs& a, int grid, cudaStream_t stream) {
  if (EPI == VAD_EPI_STORE) return launch_first_one<VAD_EPI_STORE>(a, grid, stream);
  if (EPI == VAD_EPI_POOL) return launch_first_one<VAD_EPI_POOL>(a, grid, stream);
  return VAD_ERR_UNSUPPORTED;
}

int set_trap_slot(unsigned long long* device_ptr) {
  return static_cast<int>(cudaMemcpyToSymbol(g_vad_trap_slot, &device_ptr, sizeof(device_ptr)));
}

// ------------------------------------------------------------------------------------------------ host dispatch
template <int CK, int BN, int EPI>
static int launch_one(const ConvArgs& a, int grid, cudaStream_t stream) {
  using C = Cfg<CK, BN, EPI>;
  static SmemConfig cfg;  // per instantiation
  if (int e = ensure_smem(conv_umma_kernel<CK, BN, EPI>, cfg, C::kSmemBytes)) return e;
  return launch_conv_kernel(conv_umma_kernel<CK, BN, EPI>, a, grid, block_threads(BN, EPI), C::kSmemBytes, stream);
}

#define VAD_CASE(ck, bn, epi) \
  if (CK == ck && BN == bn && EPI == epi) return launch_one<ck, bn, epi>(a, grid, stream);

int launch_conv_umma(int CK, int BN, int EPI, const ConvArgs& a, int grid, cudaStream_t stream) {
  // (K chunk, N tile, epilogue) instantiations used by the two models; anything else is VAD_ERR_UNSUPPORTED.
  VAD_CASE(32, 32, VAD_EPI_STORE)
  VAD_CASE(32, 64, VAD_EPI_STORE)
  VAD_CASE(64, 32, VAD_EPI_STORE)
  VAD_CASE(64, 64, VAD_EPI_STORE)
  VAD_CASE(64, 128, VAD_EPI_STORE)
  VAD_CASE(64, 256, VAD_EPI_STORE)
  VAD_CASE(32, 32, VAD_EPI_POOL)
  VAD_CASE(32, 64, VAD_EPI_POOL)
  VAD_CASE(64, 32, VAD_EPI_POOL)
  VAD_CASE(64, 64, VAD_EPI_POOL)
  VAD_CASE(64, 128, VAD_EPI_POOL)
  VAD_CASE(64, 256, VAD_EPI_POOL)
  VAD_CASE(32, 128, VAD_EPI_CONVT)
  VAD_CASE(64, 128, VAD_EPI_CONVT)
  VAD_CASE(32, 128, VAD_EPI_LSTM)
  VAD_CASE(64, 128, VAD_EPI_LSTM)
  VAD_CASE(32, 16, VAD_EPI_TANH_SCORE)
  VAD_CASE(32, 16, VAD_EPI_CONVT_TANH_SCORE)
  return VAD_ERR_UNSUPPORTED;
}
#undef VAD_CASE

template <int CK, int BN, int EPI>
static constexpr int halo_fixed_bytes() {
  return 1024 + 9 * BN * CK * 2 + staging_bytes(BN, EPI);
}

template <int CK, int BN, int EPI>
static int launch_halo_one(const ConvArgs& a, int grid, cudaStream_t stream) {
  const int smem = halo_fixed_bytes<CK, BN, EPI>() + a.halo_stages * a.halo_patch_bytes * a.halo_npatch;
  static SmemConfig cfg;  // per instantiation
  if (int e = ensure_smem(conv_halo_kernel<CK, BN, EPI>, cfg, smem)) return e;
  return launch_conv_kernel(conv_halo_kernel<CK, BN, EPI>, a, grid, block_threads(BN), smem, stream);
}

#define VAD_HALO_CASES(X)              \
  X(32, 32, VAD_EPI_STORE)             \
  X(32, 64, VAD_EPI_STORE)             \
  X(64, 64, VAD_EPI_STORE)             \
  X(64, 128, VAD_EPI_STORE)            \
  X(32, 32, VAD_EPI_POOL)              \
  X(32, 64, VAD_EPI_POOL)              \
  X(64, 64, VAD_EPI_POOL)              \
  X(64, 128, VAD_EPI_POOL)             \
  X(32, 16, VAD_EPI_TANH_SCORE)

int halo_smem_bytes(int CK, int BN, int EPI, int patch_bytes_total, int stages) {
#define X(ck, bn, epi) \
  if (CK == ck && BN == bn && EPI == epi) return halo_fixed_bytes<ck, bn, epi>() + stages * patch_bytes_total;
  VAD_HALO_CASES(X)
#undef X
  return 0;
}

int launch_conv_halo(int CK, int BN, int EPI, const ConvArgs& a, int grid, cudaStream_t stream) {
#define X(ck, bn, epi) \
  if (CK == ck && BN == bn && EPI == epi) return launch_halo_one<ck, bn, epi>(a, grid, stream);
  VAD_HALO_CASES(X)
#undef X
  return VAD_ERR_UNSUPPORTED;
}

template <int CK, int BN, int EPI>
static constexpr int kx_fixed_bytes() {
  return 1024 + 3 * kx_mma_n(BN, EPI) * CK * 2 + kx_groups(BN, EPI) * staging_group_bytes(BN, EPI);
}

template <int CK, int BN, int EPI>
static int launch_kx_one(const ConvArgs& a, int grid, cudaStream_t stream) {
  const int smem = kx_fixed_bytes<CK, BN, EPI>() + a.halo_stages * (8 * 18 * CK * 2);
  static SmemConfig cfg;
  if (int e = ensure_smem(conv_kx_kernel<CK, BN, EPI>, cfg, smem)) return e;
  return launch_conv_kernel(conv_kx_kernel<CK, BN, EPI>, a, grid, 128 + 128 * kx_groups(BN, EPI), smem, stream);
}

#define VAD_KX_CASES(X)                \
  X(32, 32, VAD_EPI_STORE)             \
  X(32, 32, VAD_EPI_POOL)              \
  X(32, 64, VAD_EPI_STORE)             \
  X(32, 64, VAD_EPI_POOL)              \
  X(64, 64, VAD_EPI_STORE)             \
  X(64, 64, VAD_EPI_POOL)              \
  X(32, 16, VAD_EPI_TANH_SCORE)

// dynamic smem of the kx kernel without the patch ring (0: configuration not instantiated)
int kx_fixed_smem_bytes(int CK, int BN, int EPI) {
#define X(ck, bn, epi) \
  if (CK == ck && BN == bn && EPI == epi) return kx_fixed_bytes<ck, bn, epi>();
  VAD_KX_CASES(X)
#undef X
  return 0;
}
int kx_mma_columns(int BN, int EPI) { return kx_mma_n(BN, EPI); }

int launch_conv_kx(int CK, int BN, int EPI, const ConvArgs& a, int grid, cudaStream_t stream) {
#define X(ck, bn, epi) \
  if (CK == ck && BN == bn && EPI == epi) return launch_kx_one<ck, bn, epi>(a, grid, stream);
  VAD_KX_CASES(X)
#undef X
  return VAD_ERR_UNSUPPORTED;
}

template <int EPI>
static int launch_hs_one(const ConvArgs& a, int grid, cudaStream_t stream) {
  const int smem = 1024 + kHsBRing * kHsBBytes + a.halo_stages * kHsPatchPitch + 2 * staging_group_bytes(128, EPI);
  static SmemConfig cfg;
  if (int e = ensure_smem(conv_hs_kernel<EPI>, cfg, smem)) return e;
  return launch_conv_kernel(conv_hs_kernel<EPI>, a, grid, 384, smem, stream);
}

// patch ring slots that fit next to the weight ring and the staging buffers
int hs_patch_stages(int EPI) {
  const int fixed = 1024 + kHsBRing * kHsBBytes + 2 * staging_group_bytes(128, EPI);
  int n = (kSmemBudget - fixed) / kHsPatchPitch;
  return n > 4 ? 4 : n;
}

int launch_conv_hs(int EPI, const ConvArgs& a, int grid, cudaStream_t stream) {
  if (EPI == VAD_EPI_STORE) return launch_hs_one<VAD_EPI_STORE>(a, grid, stream);
  if (EPI == VAD_EPI_POOL) return launch_hs_one<VAD_EPI_POOL>(a, grid, stream);
  return VAD_ERR_UNSUPPORTED;
}

// Step counters of the persistent ConvLSTM kernels.  Callers that own scratch memory (the model-level entry points:
// their workspace) pass `counters`; otherwise a slot of a rotating per-device pool is used (slot index taken atomically,
// so concurrent host threads never share one).  Either way the counters are zeroed on the call's own stream first.
__device__ unsigned int g_lstm_counters[256];
static std::atomic<unsigned int> g_lstm_next_slot{0};

static int lstm_counters(unsigned int* caller, int n, cudaStream_t stream, unsigned int** out) {
  unsigned int* c = caller;
  if (!c) {
    unsigned int* base = nullptr;
    cudaError_t e = cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_lstm_counters);
    if (e != cudaSuccess) return static_cast<int>(e);
    c = base + 2 * (g_lstm_next_slot.fetch_add(1, std::memory_order_relaxed) & 127u);
  }
  cudaError_t e = cudaMemsetAsync(c, 0, n * sizeof(unsigned int), stream);
  if (e != cudaSuccess) return static_cast<int>(e);
  *out = c;
  return 0;
}

template <int CK>
static int launch_lstm_seq_one(const ConvArgs& a, int T, int grid, unsigned int* counters, cudaStream_t stream) {
  using C = Cfg<CK, 128, VAD_EPI_LSTM>;
  constexpr int smem = C::kStages * C::kStageBytes + 8192 + 1024;
  static SmemConfig cfg;
  if (int e = ensure_smem(convlstm_seq_kernel<CK>, cfg, smem)) return e;
  unsigned int* counter = nullptr;
  if (int e = lstm_counters(counters, 1, stream, &counter)) return e;
  return launch_cooperative(convlstm_seq_kernel<CK>, grid, 256, smem, stream, a, T, counter);
}

int launch_convlstm_seq(int CK, const ConvArgs& a, int T, int grid, unsigned int* counters, cudaStream_t stream) {
  if (CK == 64) return launch_lstm_seq_one<64>(a, T, grid, counters, stream);
  if (CK == 32) return launch_lstm_seq_one<32>(a, T, grid, counters, stream);
  return VAD_ERR_UNSUPPORTED;
}

int launch_convlstm_patch(const ConvArgs& a, int T, int grid, unsigned int* counters, cudaStream_t stream) {
  constexpr int smem = kLpBRing * 128 * 128 + kLpPatches * kLpPatchPitch + 8192 + 1024;
  static_assert(smem <= kSmemBudget, "ConvLSTM patch kernel shared memory");
  static SmemConfig cfg;
  if (int e = ensure_smem(convlstm_patch_kernel, cfg, smem)) return e;
  unsigned int* counter = nullptr;
  if (int e = lstm_counters(counters, 1, stream, &counter)) return e;
  return launch_cooperative(convlstm_patch_kernel, grid, 384, smem, stream, a, T, counter);
}

int launch_convlstm2_patch(const ConvArgs& a1, const ConvArgs& a2, int T, int grid, unsigned int* counters,
                           cudaStream_t stream) {
  constexpr int smem = kLpBRing * 128 * 128 + kLpPatches * kLpPatchPitch + 8192 + 1024;
  static SmemConfig cfg;
  if (int e = ensure_smem(convlstm2_patch_kernel, cfg, smem)) return e;
  unsigned int* pair = nullptr;  // {layer 1, layer 2}
  if (int e = lstm_counters(counters, 2, stream, &pair)) return e;
  return launch_cooperative(convlstm2_patch_kernel, grid, 384, smem, stream, a1, a2, T, pair);
}

}  // namespace vad
