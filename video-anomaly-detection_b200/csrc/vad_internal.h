// Internal declarations shared by the .cu files of libvad_b200.so (not part of the public ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vad_b200.h"

namespace vad {

// Kernel argument block of the conv kernels (passed as a __grid_constant__ parameter; the tensor maps must stay
// 64-byte aligned).
struct alignas(64) ConvArgs {
  CUtensorMap mapA0;   // bf16 [B][T][H][W][C] source 0; box {CK, TW, TH, 1, TN} (streaming) or {CK, PW, PH, 1, 1} (halo)
  CUtensorMap mapA1;   // optional source 1 (ConvLSTM hidden state)
  CUtensorMap mapB;    // bf16 [n_total][K] weights, box {CK, BN}
  CUtensorMap mapOut;  // bf16 output, used when tma_store != 0 (box = one staged chunk of the output tile)
  CUtensorMap mapOut2;  // transposed conv with convt_split: the map of output rows 2h + 1 (mapOut: rows 2h)
  int chunks0, chunks1;  // CK-wide channel chunks per source
  int ntaps;
  int w_ctap;  // weight columns per tap
  int tA0, tA1;
  int B, H, W;
  int lgTW, lgTH, lgTN;
  int tiles_w, tiles_h, tiles_b;
  int w_step;    // columns between consecutive tiles (1 << lgTW, or 6 for the kx-merged kernel)
  int tw_valid;  // valid output columns per tile row (1 << lgTW, or 6)
  int pdl;       // programmatic dependent launch: 0 off, 1 wait before the first global read, 2 ConvLSTM step t >= 1
  int pair;      // 1, or 2: tiles are enumerated (and processed) as horizontally adjacent pairs
  int row_perm;  // 1: accumulator rows are in the first conv's permuted pixel order (make_epi_lane)
  int n_tiles;
  int total_tiles;
  const float* bias;
  const float* bias2;  // convt2_score_kernel: bias of the second transposed convolution (16 entries)
  float slope;
  void* out;
  long long out_fs;
  int out_cp;
  int cout;
  float* c_state;
  int lstm_first;
  const float* x;
  float* recon;
  float* heat;
  float* partials;
  int dbg;              // bring-up switches (VAD_DBG environment variable)
  int dual_mma;         // 1: two MMA-issuer warps take alternate tiles
  int token;            // 1: the two issuers pass a token (strict tile order on the tensor pipe)
  long long* timeline;  // optional [role 0..3][64 tiles][8 events] clock64 stamps of CTA 0 (vad_debug_set_timeline)
  const void* w_first;  // first conv: bf16 [32 n][32 k] weights, k = (ky*3+kx)*3+ci (27 real + 5 zero)
  int pair_fold;        // halo kernel: pixel-pair folded 3x3 layer (vad_conv_desc.pair_fold)
  int convt_split;      // ConvT TMA store through two maps {co, dj, w, h, b} (di folded into the base address): h and b stay
                        // separate dimensions, so tiles that overhang the frame's last rows are clipped by the hardware
  int b_resident;       // streaming kernel, single-tap layers: this CTA's weight tiles stay in shared memory (the grid is
                        // a multiple of n_tiles, so a CTA only ever sees one n tile); the ring carries activations only
  // epilogue staging / TMA store
  int tma_store;  // 1: stage the bf16 tile in swizzled smem and store it with TMA (coalesced, clipped by hardware)
  int out_chunk;  // channels per staged chunk: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
  int out_t;      // T coordinate of the TMA store (ConvLSTM sequence: the step index)
  // halo kernel (3x3 conv, weights resident in smem, one input patch per tile reused by all 9 taps)
  int halo_pw, halo_ph;       // patch width / height in pixels
  int halo_npatch;            // 1: one (TW+2)-wide patch; 3: three TW-wide patches shifted by dx
  int halo_patch_bytes;       // bytes of one patch in smem (multiple of 1024)
  int halo_stages;            // ring depth
  int halo_sbo_rows;          // patch rows between consecutive 8-row groups of the A operand
  int halo_base_mode;         // smem descriptor base-offset field: 0 -> 0, 1 -> (addr >> 7) & 7
  int tap_patch[9];           // which patch a tap reads
  int tap_row[9];             // first patch row (pixel index) of the tap's A operand
};

struct TileGeom {
  int lgTW, lgTH, lgTN;
  int tiles_w, tiles_h, tiles_b;
  int m_tiles() const { return tiles_w * tiles_h * tiles_b; }
};
TileGeom pick_tile_geometry(int B, int H, int W, bool single_frame_tiles);

int launch_conv_umma(int CK, int BN, int EPI, const ConvArgs& a, int grid, cudaStream_t stream);
int launch_conv_halo(int CK, int BN, int EPI, const ConvArgs& a, int grid, cudaStream_t stream);
int launch_conv_kx(int CK, int BN, int EPI, const ConvArgs& a, int grid, cudaStream_t stream);
int kx_fixed_smem_bytes(int CK, int BN, int EPI);
int kx_mma_columns(int BN, int EPI);
int launch_conv_hs(int EPI, const ConvArgs& a, int grid, cudaStream_t stream);
int hs_patch_stages(int EPI);
// persistent ConvLSTM kernels (cooperative launches).  `counters`: device scratch for the grid-wide step counters (one
// unsigned per layer; zeroed by the launcher on `stream`), or nullptr to take a slot of the library's rotating pool.
// VAD_ERR_UNSUPPORTED when the grid cannot be co-resident on this device (callers fall back to one launch per step).
int launch_convlstm_seq(int CK, const ConvArgs& a, int T, int grid, unsigned int* counters, cudaStream_t stream);
int launch_convlstm_patch(const ConvArgs& a, int T, int grid, unsigned int* counters, cudaStream_t stream);
int launch_convlstm2_patch(const ConvArgs& a1, const ConvArgs& a2, int T, int grid, unsigned int* counters,
                           cudaStream_t stream);
int launch_conv_first(int EPI, const ConvArgs& a, int grid, cudaStream_t stream);
int launch_conv_first_pool(const ConvArgs& a, int grid, cudaStream_t stream);
int launch_enc1_fused(const ConvArgs& a, int grid, cudaStream_t stream);
int launch_convt2_score(const ConvArgs& a, int grid, cudaStream_t stream);
int launch_convt_conv_score(const ConvArgs& a, int grid, cudaStream_t stream);
// dynamic shared memory the halo kernel needs for `stages` ring slots (0 if the configuration is not instantiated)
int halo_smem_bytes(int CK, int BN, int EPI, int patch_bytes_total, int stages);
int set_trap_slot(unsigned long long* device_ptr);
void count_launch();
int sm_count();
// M x N tiles (= CTAs) the persistent ConvLSTM path would use for a batch of B clips with an h x w latent: the patch
// kernel's geometry where it applies, the generic one otherwise.  The model-level schedule splits batches whose tiles
// exceed the SM count into resident-size groups of clips.
int lstm_persistent_tiles(int B, int H, int W, int c0, int c1, int n_total);

}  // namespace vad
