// Internal declarations shared by the .cu files of libvad_b200.so (not part of the public ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "vad_b200.h"

namespace vad {

// Kernel argument block of conv_umma_kernel (passed as a __grid_constant__ parameter; the three tensor maps
// must stay 64-byte aligned).
struct alignas(64) ConvArgs {
  CUtensorMap mapA0;  // bf16 [B][T][H][W][C] source 0, box {CK, TW, TH, 1, TN}
  CUtensorMap mapA1;  // optional source 1 (ConvLSTM hidden state)
  CUtensorMap mapB;   // bf16 [n_total][K] weights, box {CK, BN}
  int chunks0, chunks1;  // CK-wide channel chunks per source
  int ntaps;
  int w_ctap;  // weight columns per tap
  int tA0, tA1;
  int B, H, W;
  int lgTW, lgTH, lgTN;
  int tiles_w, tiles_h, tiles_b;
  int n_tiles;
  int total_tiles;
  const float* bias;
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
};

struct TileGeom {
  int lgTW, lgTH, lgTN;
  int tiles_w, tiles_h, tiles_b;
  int m_tiles() const { return tiles_w * tiles_h * tiles_b; }
};
TileGeom pick_tile_geometry(int B, int H, int W, bool single_frame_tiles);

int launch_conv_umma(int CK, int BN, int EPI, const ConvArgs& a, int grid, cudaStream_t stream);
void count_launch();
int sm_count();

}  // namespace vad
