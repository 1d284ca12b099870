/*
 * vad_b200.h — C ABI of the B200-native anomaly-scoring hot path.
 *
 * The reference (KuldeepChoksi/video-anomaly-detection) has no FFI: its hot path is the Python class API of
 * models/autoencoder.py and models/video_autoencoder.py.  This library is what the drop-in Python classes
 * (video-anomaly-detection_b200/models/*.py) bind with ctypes; every entry point names the reference lines it
 * replaces.  Conventions:
 *   - extern "C", plain pointers and sizes only; all data pointers are DEVICE pointers owned by the caller.
 *   - nothing here allocates device memory or synchronises; work is enqueued on `stream`.
 *   - return value: 0 = ok, negative = argument / configuration error (see vad_error_string), positive = cudaError_t.
 *   - activations between layers are bf16 NHWC ("pixel-major"); model inputs / outputs are fp32 NCHW like the reference.
 */
#ifndef VAD_B200_H_
#define VAD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vad_stream_t; /* cudaStream_t */

/* ---- error codes --------------------------------------------------------------------------------------------- */
#define VAD_OK 0
#define VAD_ERR_ARG (-1)          /* null pointer / bad size */
#define VAD_ERR_SHAPE (-2)        /* H or W not a multiple of 16, channel count unsupported */
#define VAD_ERR_UNSUPPORTED (-3)  /* no kernel instantiation for this (K-chunk, N-tile, epilogue) */
#define VAD_ERR_DRIVER (-4)       /* cuTensorMapEncodeTiled unavailable / failed */
#define VAD_ERR_WORKSPACE (-5)    /* workspace too small */

const char* vad_error_string(int code);
int vad_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches claim). */
unsigned long long vad_launch_count(void);
/* Bring-up aid: if a kernel's bounded mbarrier wait timed out (the kernel then traps), out = {wait-site tag,
 * blockIdx.x, threadIdx.x, parity}; all zero otherwise.  Readable even after the CUDA context reports an error. */
int vad_debug_last_trap(unsigned long long out[4]);
/* Tuning aid: which narrow 3x3 layers may use the kernel that folds the horizontal taps into N when `weight_kx` is
 * given: 0 none, 1 the 3-channel score layer (default), 2 also Cout = 32, 3 also Cout = 64; -1 restores the VAD_KX
 * environment setting.  Returns the previous mode. */
int vad_debug_set_kx(int mode);
/* Tuning / test aid for vad_convlstm_sequence: 0 = one launch per time step (chained with programmatic dependent
 * launch), 1 = one persistent launch per layer when the tiles fit the SMs (default; the patch kernel where its tile
 * shapes apply), 2 = the same but always the streaming sequence kernel, -1 = VAD_LSTM_SEQ environment setting.
 * Returns the previous override. */
int vad_debug_set_lstm_mode(int mode);
/* Bring-up aid: when set, CTA 0 of every vad_conv_layer kernel stamps clock64 at role events of its first 64 tiles
 * into device_buf[4 roles][64][16] (role 0 TMA producer, 1 MMA issuer, 2/3 epilogue group 0/1).  NULL disables. */
int vad_debug_set_timeline(long long* device_buf);

/* ---- one convolution-as-GEMM layer ----------------------------------------------------------------------------
 * Implicit-GEMM on tcgen05/TMEM fed by TMA.  M = B*H*W input pixels (tiles of 128), N = n_total, K = ntaps*(c0+c1).
 *   ntaps = 9 : nn.Conv2d(k=3, padding=1)            reference models/autoencoder.py:39-76,107-137,
 *                                                     models/video_autoencoder.py:46-52,193-211
 *   ntaps = 1 : nn.ConvTranspose2d(k=2, stride=2) as a GEMM with N = 4*Cout (column = (di*2+dj)*Cout + co)
 *               reference models/autoencoder.py:104-134, models/video_autoencoder.py:244-259;
 *               or a 1x1 conv (video_autoencoder.py:311-312 `proj`).
 * BatchNorm (eval) must already be folded into weight/bias by the caller (host-side prepare step).
 */
enum vad_epilogue {
  VAD_EPI_STORE = 0,            /* y = act(acc + bias) -> bf16 NHWC [B,H,W,n_total]                                   */
  VAD_EPI_POOL = 1,             /* y = act(maxpool2x2(acc) + bias) -> bf16 NHWC [B,H/2,W/2,n_total]                   */
  VAD_EPI_CONVT = 2,            /* pixel-shuffle store: bf16 NHWC [B,2H,2W,n_total/4], y = act(acc + bias)            */
  VAD_EPI_LSTM = 3,             /* ConvLSTM gate update (video_autoencoder.py:72-83): c' , h' ; N tile = 4 gates x 32  */
  VAD_EPI_TANH_SCORE = 4,       /* conv 3x3 -> 3 ch: recon = tanh(acc+bias); fused (x-recon)^2 reduction              */
  VAD_EPI_CONVT_TANH_SCORE = 5  /* convT k2s2 -> 3 ch: recon = tanh(acc+bias); fused (x-recon)^2 reduction            */
};

typedef struct vad_conv_desc {
  /* A operand: one or two bf16 NHWC sources viewed as [B][T][H][W][C] (second source = ConvLSTM hidden state) */
  const void* src0;
  const void* src1;
  int c0, c1;   /* channels per source; multiples of 32 (c1 = 0: no second source) */
  int T0, T1;   /* T extent of each source buffer (1 for plain frame batches) */
  int t0, t1;   /* time index read from each source */
  int B, H, W;  /* frames and INPUT spatial size */
  int ntaps;    /* 9 or 1 */
  /* B operand */
  const void* weight; /* bf16 [n_total][ntaps*w_ctap], K index = tap*w_ctap + channel (source 0 first) */
  const float* bias;  /* fp32 [n_total] */
  int w_ctap;         /* weight columns per tap (0 = c0+c1); lets ConvLSTM step 0 skip the h half (c1 = 0) */
  int n_total;        /* multiple of 16 */
  int cout;           /* real output channels (CONVT: n_total/4; *_SCORE: 3) */
  int epilogue;       /* enum vad_epilogue */
  float slope;        /* act(v) = v > 0 ? v : v*slope   (0.2 LeakyReLU, 0 ReLU, 1 identity) */
  /* outputs */
  void* out;                  /* bf16; STORE/POOL/CONVT/LSTM(h') */
  long long out_frame_stride; /* elements between consecutive frames b in `out` */
  int out_cpitch;             /* elements between consecutive pixels in `out` */
  float* c_state;             /* LSTM: fp32 [B,H,W,hid] cell state, updated in place */
  int lstm_first;             /* LSTM: 1 = step 0 (c_prev = 0, c_state not read) */
  const float* x;             /* *_SCORE: fp32 [B,3,Ho,Wo] model input */
  float* recon;               /* *_SCORE: optional fp32 [B,3,Ho,Wo] */
  float* heat;                /* *_SCORE: optional fp32 [B,Ho,Wo] per-pixel channel-mean squared error */
  float* partials;            /* *_SCORE: fp32 [m_tiles][4 warps][4]: per tile and 32-row quarter (sum of squares, min, max, -) */
  /* optional second layout of the same 3x3 weights for narrow layers (Cout <= 64, one source of 32/64 channels):
   * bf16 [3*Cout rows, zero-padded to a multiple of 16][3*Cin], row = kx*Cout + co, column = ky*Cin + ci
   * (*_SCORE: Cout = 3 -> 9 rows padded to 16).  When present the library may fold the horizontal taps into the
   * GEMM N extent (3 instead of 9 shifted MMAs per K step); NULL keeps the tap-per-MMA kernels. */
  const void* weight_kx;
  /* optional device scratch of >= 8 bytes for vad_convlstm_sequence / vad_convlstm2_sequence: the grid-wide step
   * counters of the persistent kernels live there (zeroed by the library on `stream`; must not be shared by two calls
   * in flight).  NULL: a slot of the library's own rotating pool is used. */
  void* scratch;
  /* Pixel-pair folding of a narrow 3x3 layer (Cin = 32; STORE or POOL epilogue).  A tcgen05.mma with N = 32 is bound
   * by streaming its A operand, not by math, so such a layer can be described on the PAIR view of its tensors instead:
   * input bf16 NHWC [B,H,W,32] seen as [B,H,W/2,64] (c0 = 64, W = W/2), n_total = 2*Cout with column p*Cout + co =
   * output pixel 2Q+p, channel co, and `weight` the 3x3 "pair" kernel [2*Cout][9*64] that holds w[co][ci][ky][kx] at
   * row p_out*Cout + co, column (ky*3 + kxp)*64 + p_in*32 + ci, kx = 2*(kxp-1) + p_in - p_out + 1 (zero where kx is
   * not in 0..2); bias repeated for both pixels.  STORE then writes [B,H,W/2,2*Cout] = the ordinary NHWC output;
   * POOL writes [B,H/2,W/2,Cout] (out_cpitch = Cout) with the horizontal half of the 2x2 window taken inside the
   * accumulator row.  pair_fold = 1 tells the kernel which K steps are structurally zero (skipped). */
  int pair_fold;
} vad_conv_desc;

int vad_conv_layer(const vad_conv_desc* d, vad_stream_t stream);
/* The last two layers of the video decoder and the error reduction in one kernel (reference
 * models/video_autoencoder.py:252-259 ConvTranspose2d(64,32,2,2)+BN+ReLU, ConvTranspose2d(32,3,2,2)+Tanh, and
 * :371-384): a k2s2 transposed convolution has no halo, so input pixel (h,w) alone determines the 4x4 output block
 * (4h.., 4w..) and the 32-channel intermediate never reaches HBM.  `d` describes the FIRST transposed convolution
 * (ntaps = 1, c0 = 64, n_total = 128, cout = 32, folded-BN weight / bias, slope) on a [B,H,W,64] bf16 input, plus the
 * score outputs for the [B,3,4H,4W] model input: x, partials ([vad_convt2_score_tiles(d)][4][4]), optional recon / heat.
 * weight2: bf16 [16][32], row = (di*2+dj)*3 + co (rows 12..15 zero); bias2: fp32 [16] in the same order.
 * Other channel widths: VAD_ERR_UNSUPPORTED (callers then run the two layers with vad_conv_layer). */
int vad_convt2_score(const vad_conv_desc* d, const void* weight2, const float* bias2, vad_stream_t stream);
int vad_convt2_score_tiles(const vad_conv_desc* d);
/* The last block of the image decoder and the error reduction in one kernel (reference models/autoencoder.py:131-138
 * ConvTranspose2d(32,32,2,2)+BN+ReLU, Conv2d(32,3,3,padding=1)+Tanh, and :214-221): the transposed convolution is
 * recomputed per tile with a one-pixel halo and its 32-channel full-resolution output stays in shared memory.
 * `d` describes the transposed convolution (ntaps = 1, c0 = 32, n_total = 128, cout = 32) on a [B,H,W,32] bf16 input
 * plus the score outputs for the [B,3,2H,2W] model input: x, partials ([vad_convt_conv_score_tiles(d)][4][4]),
 * optional recon / heat.  weight2_kx: the 3x3 conv's kx-folded layout (see vad_conv_desc.weight_kx) bf16 [16][96];
 * bias2: fp32 [16] (3 real entries).  Other channel widths: VAD_ERR_UNSUPPORTED. */
int vad_convt_conv_score(const vad_conv_desc* d, const void* weight2_kx, const float* bias2, vad_stream_t stream);
int vad_convt_conv_score_tiles(const vad_conv_desc* d);
/* One ConvLSTM layer over a whole sequence (reference ConvLSTM.forward time loop, models/video_autoencoder.py:153-167):
 * T launches of the VAD_EPI_LSTM layer with the tensor maps encoded once.  `d` describes a step t >= 1: src0 = input
 * sequence bf16 [B][T][h][w][c0] (T0 = T), src1 = out = hidden sequence bf16 [B][T][h][w][hid] (T1 = T, c1 = hid,
 * out_frame_stride = T*h*w*hid), c_state fp32 [B][h][w][hid]; zero initial state. */
int vad_convlstm_sequence(const vad_conv_desc* d, int T, vad_stream_t stream);
/* Both layers of a two-layer ConvLSTM in ONE persistent launch, as a wavefront (layer 2's step t only needs layer 1's
 * h_t): each CTA alternates between layer 1's step t+1 and layer 2's step t, so one layer's recurrence chain runs under
 * the other's MMAs.  d1 / d2 as for vad_convlstm_sequence with d2->src0 == d1->out; results are bit-identical to two
 * vad_convlstm_sequence calls.  VAD_ERR_UNSUPPORTED when the shapes do not fit the persistent patch kernel (the tiles
 * of a layer must all be resident at once) or VAD_LSTM2=0: callers then run the layers one after the other. */
int vad_convlstm2_sequence(const vad_conv_desc* d1, const vad_conv_desc* d2, int T, vad_stream_t stream);
/* number of M tiles (= rows of `partials` for a *_SCORE layer) vad_conv_layer will use for this description;
 * negative = the error vad_conv_layer would return.  Launches nothing. */
int vad_conv_layer_tiles(const vad_conv_desc* d);
/* number of 128-pixel tiles of the default tiling for a (B,H,W) input (upper bound helpers / tests) */
int vad_conv_m_tiles(int B, int H, int W, int force_single_frame_tiles);

/* ---- first layer: fp32 NCHW (3 ch) -> bf16 NHWC, conv3x3 + folded BN + LeakyReLU (+ 2x2 max-pool) --------------
 * reference models/autoencoder.py:39-41 (enc1.0, no pool), models/video_autoencoder.py:193-196 (encoder.0, pool) */
int vad_first_conv(const float* x, const float* weight /* fp32 [27][cout], k = (ky*3+kx)*3+ci */,
                   const float* bias, int cout, float slope, int pool, int B, int H, int W, void* out_bf16_nhwc,
                   vad_stream_t stream);

/* Tensor-core variant of the same layer (default): weight is bf16 [32][32], row = output channel, column
 * k = (ky*3+kx)*3+ci for k < 27 and zero for k >= 27; input pixels are rounded to bf16 by the im2col producer. */
int vad_first_conv_tc(const float* x, const void* weight_bf16, const float* bias, float slope, int pool, int B, int H,
                      int W, void* out_bf16_nhwc, vad_stream_t stream);

/* The pooled first layer (video encoder: conv3x3 + BN + LeakyReLU + 2x2 max-pool, models/video_autoencoder.py:193-196)
 * with the pooling window folded into the GEMM N extent (one accumulator row per POOLED pixel, K = the 4x4x3 input
 * window): weight_pf is bf16 [128][64], row = pos*32 + co with pos = py*2 + px the output's place in its 2x2 window,
 * column k = ((py+ky)*4 + (px+kx))*3 + ci holding w[co][ci][ky][kx] (zero elsewhere, columns 48..63 zero).
 * H, W even; out bf16 NHWC [B,H/2,W/2,32]. */
int vad_first_conv_pool(const float* x, const void* weight_pf, const float* bias, float slope, int B, int H, int W,
                        void* out_bf16_nhwc, vad_stream_t stream);

/* The first block of the image encoder in one kernel (models/autoencoder.py:38-45: conv3x3(3->32)+BN+LeakyReLU,
 * conv3x3(32->32)+BN+LeakyReLU, MaxPool2d(2,2)): the 32-channel full-resolution tensor between the two convolutions
 * stays in shared memory.  w_first: bf16 [32][32] as vad_first_conv_tc; w_pair / bias2: the pixel-pair folded second
 * conv, bf16 [64][9*64] / fp32 [>= 32] (vad_conv_desc.pair_fold; pack_conv3x3_pair).  H, W multiples of 16;
 * out bf16 NHWC [B,H/2,W/2,32].  Bit-identical to vad_first_conv_tc followed by the pair-folded VAD_EPI_POOL layer. */
int vad_enc1_fused(const float* x, const void* w_first, const float* bias1, const void* w_pair, const float* bias2,
                   float slope, int B, int H, int W, void* out_bf16_nhwc, vad_stream_t stream);

/* ---- scoring reduction ----------------------------------------------------------------------------------------
 * reference models/autoencoder.py:214-221, models/video_autoencoder.py:371-384, evaluate_video.py:56 (min/max) */
/* finalize the partials of a fused *_SCORE layer (tiles_per_frame = 4 x the frame's tiles: one entry per tile quarter)
 * or of vad_score: score[f] = sum/(3*H*W), minmax[f] = {min,max} of the map */
int vad_score_finalize(const float* partials, int frames, int tiles_per_frame, int H, int W, float* score,
                       float* minmax /* nullable [frames][2] */, vad_stream_t stream);
/* standalone (unfused) scoring pass: x, recon fp32 [N,3,H,W]; scratch >= vad_score_scratch_bytes(N,H,W) */
size_t vad_score_scratch_bytes(int frames, int H, int W);
int vad_score(const float* x, const float* recon, int frames, int H, int W, float* score, float* minmax,
              float* heat, void* scratch, vad_stream_t stream);

/* ---- layout helpers -------------------------------------------------------------------------------------------- */
/* bf16 NHWC [N,H,W,C] -> fp32 NCHW [N,C,H,W]  (get_latent: models/autoencoder.py:195-197) */
int vad_nhwc_bf16_to_nchw_f32(const void* src, int N, int H, int W, int C, float* dst, vad_stream_t stream);
/* fp32 NCHW -> bf16 NHWC (entry for ConvLSTM / encoder / decoder sub-module calls on fp32 tensors) */
int vad_nchw_f32_to_nhwc_bf16(const float* src, int N, int C, int H, int W, void* dst, vad_stream_t stream);
/* per-frame heat-map normalisation to uint8: (e-min)/(max-min+1e-8)*255, truncation — evaluate_video.py:56-57 */
int vad_heatmap_u8(const float* heat, const float* minmax, int frames, int H, int W, uint8_t* out,
                   vad_stream_t stream);

/* ---- frame I/O either side of the path --------------------------------------------------- */
/* uint8 HWC RGB [N,H,W,3] -> fp32 NCHW [N,3,H,W] in [-1,1]: ToTensor + Normalize(.5,.5) — utils/dataset.py:65-70,
 * utils/video_dataset.py:62-66,356-360 (the resize stays with the decoder); H*W multiple of 4 */
int vad_u8_hwc_to_f32_nchw(const uint8_t* src, int frames, int H, int W, float* dst, vad_stream_t stream);
/* fp32 NCHW [-1,1] -> uint8 HWC: `denormalize` — evaluate_video.py:40-49 */
int vad_f32_nchw_to_u8_hwc(const float* src, int frames, int H, int W, uint8_t* dst, vad_stream_t stream);
/* per-frame normalised error map -> JET-coloured RGB uint8 [N,H,W,3]: `create_heatmap` — evaluate_video.py:52-66 */
int vad_heatmap_jet_rgb(const float* heat, const float* minmax, int frames, int H, int W, uint8_t* out,
                        vad_stream_t stream);

/* the side-by-side panel of the reference's video output: np.hstack([denormalize(frame), denormalize(reconstruction),
 * create_heatmap(error_map)]) — evaluate_video.py:279-286,355-364 — as uint8 RGB [N,H,3W,3], byte-exact for frames of the
 * size create_heatmap resizes to (its cv2.resize is then the identity; other sizes stay with cv2 on the host).
 * x, recon fp32 [N,3,H,W]; heat fp32 [N,H,W]; minmax fp32 [N][2] (all from one vad_*_forward call). */
int vad_compose_panel(const float* x, const float* recon, const float* heat, const float* minmax, int frames, int H,
                      int W, uint8_t* out, vad_stream_t stream);

/* ---- SSIM as an alternative anomaly score ----------------------------------------------------- */
/* SSIMLoss.forward — utils/losses.py:51-93 — per frame: loss[f] = 1 - mean over (3,H,W) of the SSIM map between
 * pred[f] and target[f] (fp32 NCHW, 3 channels; 11x11 Gaussian window, sigma 1.5, zero padding).  The reference's batch
 * value is the mean of loss[]; CombinedLoss (:116-121) = (1-alpha)*MSE + alpha*loss with the MSE from vad_score.
 * ssim_map (nullable) receives the per-pixel map [frames,3,H,W]; scratch >= vad_ssim_scratch_bytes(frames,H,W). */
size_t vad_ssim_scratch_bytes(int frames, int H, int W);
int vad_ssim_loss(const float* pred, const float* target, int frames, int H, int W, float* loss, float* ssim_map,
                  void* scratch, vad_stream_t stream);

/* =================================================================================================================
 * Model-level entry points: ONE call per reference method.  These are what the drop-in Python classes bind; the
 * per-layer entry points above stay public for layer-level tests and tuning.
 *
 * Weights are the host-prepared GEMM operands (eval-mode BatchNorm folded, bf16 K-major; models/_prepare.py), passed as
 * device pointers.  Every call enqueues its whole layer schedule on `stream`, allocates nothing and never synchronises:
 * intermediate activations, score partials and the ConvLSTM step counters live in the caller-supplied workspace `ws`
 * (>= the matching *_workspace_bytes(); 256-byte aligned; must not be shared by two calls in flight on different streams).
 * All entry points are re-entrant per (stream, workspace) and use the CUDA device that is current on the calling thread.
 * ================================================================================================================= */
typedef struct vad_gemm_weights {
  const void* w;     /* bf16 [n_total][ntaps*ctap] (vad_conv_desc.weight) */
  const void* w_kx;  /* optional kx-folded layout of a narrow 3x3 layer (vad_conv_desc.weight_kx) */
  const float* bias; /* fp32 [n_total] */
  int ntaps, ctap, n_total, cout;
  /* optional pixel-pair folded form of a 3x3 layer with 32 input channels (vad_conv_desc.pair_fold): bf16
   * [2*n_total][9*64] and fp32 [2*n_total]; NULL: the layer runs on the ordinary view */
  const void* w_pair;
  const float* bias_pair;
} vad_gemm_weights;

typedef struct vad_first_weights {
  const float* w;    /* fp32 [27][cout] (vad_first_conv) */
  const void* w_tc;  /* bf16 [cout][32] (vad_first_conv_tc); NULL: CUDA-core kernel */
  const void* w_pf;  /* bf16 [128][64] (vad_first_conv_pool, cout = 32): used when the layer is pooled; NULL: w_tc path */
  const float* bias; /* fp32 [cout] */
  int cout;
} vad_first_weights;

#define VAD_FLAG_NO_FUSED_TAIL 1     /* run the decoder's last two layers one by one (test / tuning aid) */
#define VAD_FLAG_NO_LSTM_WAVEFRONT 2 /* one launch per ConvLSTM layer instead of the two-layer wavefront kernel */
#define VAD_FLAG_NO_FUSED_ENC1 4     /* image encoder: enc1.0 and enc1.3 as two launches instead of vad_enc1_fused */

/* ConvAutoencoder — reference models/autoencoder.py:24-221 */
typedef struct vad_image_model {
  int has_encoder, has_decoder, flags, reserved;
  vad_first_weights enc1_0; /* encoder.enc1.0 (+BN .1) */
  vad_gemm_weights enc[7];  /* enc1.3, enc2.0, enc2.3, enc3.0, enc3.3, enc4.0, enc4.3 */
  vad_gemm_weights dec[8];  /* dec1.0 (ConvT), dec1.3, dec2.0, dec2.3, dec3.0, dec3.3, dec4.0, dec4.3 (-> 3 ch, n_total 16) */
} vad_image_model;

/* VideoAutoencoder — reference models/video_autoencoder.py:24-384 */
#define VAD_MAX_LSTM_LAYERS 8
typedef struct vad_video_model {
  int has_encoder, lstm_layers, has_proj, has_decoder, flags, reserved;
  vad_first_weights enc0;                     /* encoder.encoder.0 (+BN .1, pooled) */
  vad_gemm_weights enc[3];                    /* encoder.encoder.4 / .8 / .12 */
  vad_gemm_weights lstm[VAD_MAX_LSTM_LAYERS]; /* convlstm.cells.i.conv, gate rows permuted (see _prepare.py) */
  vad_gemm_weights proj;                      /* 1x1 conv when lstm_hidden_dim != latent_dim */
  vad_gemm_weights dec[4];                    /* decoder.decoder.0 / .3 / .6 / .9 (-> 3 ch, n_total 16) */
} vad_video_model;

enum vad_op {
  VAD_OP_FORWARD = 0,       /* vad_image_forward / vad_video_forward */
  VAD_OP_ENCODE = 1,        /* vad_image_forward with only `latent` requested / vad_video_encode */
  VAD_OP_DECODE = 2,        /* vad_image_decode / vad_video_decode */
  VAD_OP_CONVLSTM = 3,      /* vad_convlstm_forward */
  VAD_OP_SCORE_LATENTS = 4, /* vad_video_score_latents */
  VAD_OP_FORWARD_U8 = 5     /* vad_image_forward_u8 / vad_video_forward_u8 */
};
/* Workspace the given entry point needs for this shape (T ignored by the image model); 0 = invalid arguments. */
size_t vad_image_workspace_bytes(const vad_image_model* m, int op, int B, int H, int W);
size_t vad_video_workspace_bytes(const vad_video_model* m, int op, int B, int T, int H, int W);

/* ConvAutoencoder.forward / get_latent / get_reconstruction_error in one pass — models/autoencoder.py:181-221.
 * x fp32 [B,3,H,W].  Every output is optional (NULL = not produced; the reconstruction then never reaches HBM):
 *   recon fp32 [B,3,H,W] · latent fp32 [B,latent,H/16,W/16] · score fp32 [B] (mean over C,H,W of (x-recon)^2) ·
 *   minmax fp32 [B][2] (min / max of the per-pixel map) · heat fp32 [B,H,W] (the `per_pixel=True` map).
 * Only `latent` requested: the encoder alone runs (Encoder.forward, :81-86). */
int vad_image_forward(const vad_image_model* m, const float* x, int B, int H, int W, float* recon, float* latent,
                      float* score, float* minmax, float* heat, void* ws, size_t ws_bytes, vad_stream_t stream);
/* The same from the decoder's uint8 frames: frames uint8 RGB [B,H,W,3] are normalised on the device exactly like the
 * reference datasets do on the host (ToTensor + Normalize(.5,.5): utils/dataset.py:65-70), so a caller uploads a quarter
 * of the bytes; heat_u8 (optional) uint8 [B,H,W] is create_heatmap's per-frame normalisation of the error map
 * (evaluate_video.py:56-57, bit-exact) — a quarter of the bytes to download.  Other outputs as vad_image_forward. */
int vad_image_forward_u8(const vad_image_model* m, const uint8_t* frames, int B, int H, int W, float* recon, float* latent,
                         float* score, float* minmax, float* heat, uint8_t* heat_u8, void* ws, size_t ws_bytes,
                         vad_stream_t stream);
/* Decoder.forward — models/autoencoder.py:141-146: z fp32 [B,latent,h,w] -> recon fp32 [B,3,16h,16w]. */
int vad_image_decode(const vad_image_model* m, const float* z, int B, int h, int w, float* recon, void* ws,
                     size_t ws_bytes, vad_stream_t stream);

/* VideoAutoencoder.forward / get_reconstruction_error in one pass — models/video_autoencoder.py:329-384.
 * x fp32 [B,T,3,H,W]; recon fp32 [B,T,3,H,W], score fp32 [B*T] per frame, minmax fp32 [B*T][2], heat fp32 [B*T,H,W]
 * (all optional).  The per-sequence score is the mean of a clip's frame scores (equal-sized frames).  The ConvLSTM
 * starts from the zero state for every call (:144-145); batches whose recurrent tiles exceed the SM count are run
 * in resident-size groups of clips. */
int vad_video_forward(const vad_video_model* m, const float* x, int B, int T, int H, int W, float* recon, float* score,
                      float* minmax, float* heat, void* ws, size_t ws_bytes, vad_stream_t stream);
/* ... from uint8 frames [B,T,H,W,3] (utils/video_dataset.py:62-66,141-144), optional uint8 heat maps [B*T,H,W]. */
int vad_video_forward_u8(const vad_video_model* m, const uint8_t* frames, int B, int T, int H, int W, float* recon,
                         float* score, float* minmax, float* heat, uint8_t* heat_u8, void* ws, size_t ws_bytes,
                         vad_stream_t stream);
/* VideoEncoder.forward — :217-231: frames fp32 [F,3,H,W] -> latent fp32 [F,latent,H/16,W/16] and / or the bf16 NHWC
 * [F,H/16,W/16,latent] tensor vad_video_score_latents consumes (either may be NULL). */
int vad_video_encode(const vad_video_model* m, const float* x, int F, int H, int W, float* latent, void* latent_bf16,
                     void* ws, size_t ws_bytes, vad_stream_t stream);
/* ConvLSTM -> proj -> decoder -> scoring from cached encoder features (overlap-aware streaming):
 * z bf16 NHWC [B,T,h,w,latent], x fp32 [B,T,3,16h,16w]; outputs as vad_video_forward. */
int vad_video_score_latents(const vad_video_model* m, const void* z_bf16, const float* x, int B, int T, int h, int w,
                            float* recon, float* score, float* minmax, float* heat, void* ws, size_t ws_bytes,
                            vad_stream_t stream);
/* VideoDecoder.forward — :263-276: z fp32 [F,latent,h,w] -> recon fp32 [F,3,16h,16w]. */
int vad_video_decode(const vad_video_model* m, const float* z, int F, int h, int w, float* recon, void* ws,
                     size_t ws_bytes, vad_stream_t stream);
/* ConvLSTM.forward (zero initial state, last layer's outputs) — :127-172: x fp32 [B,T,C,h,w] -> out fp32 [B,T,hid,h,w];
 * optional final state of the last layer: h_last, c_last fp32 [B,hid,h,w]. */
int vad_convlstm_forward(const vad_video_model* m, const float* x, int B, int T, int h, int w, float* out,
                         float* h_last, float* c_last, void* ws, size_t ws_bytes, vad_stream_t stream);
/* ConvLSTMCell.forward(x, (h, c)) -> (h_next, c_next) — :54-85.  One gate GEMM with the state update in its epilogue;
 * x fp32 [B,cin,h,w], h_cur / c_cur / h_next / c_next fp32 [B,hid,h,w] (h is rounded to bf16, the MMA operand type). */
size_t vad_convlstm_cell_workspace_bytes(const vad_gemm_weights* cell, int B, int h, int w);
int vad_convlstm_cell(const vad_gemm_weights* cell, const float* x, const float* h_cur, const float* c_cur, int B,
                      int h, int w, float* h_next, float* c_next, void* ws, size_t ws_bytes, vad_stream_t stream);

/* Per-layer timing of the model-level calls (bench.py's roofline): while enabled every layer launch is bracketed by
 * CUDA events on the call's stream.  vad_profile_dump synchronises those events, writes "name\tms\n" lines (in launch
 * order) into buf, clears the log and returns the number of bytes written (excluding the NUL), or a negative error. */
int vad_profile_enable(int on);
int vad_profile_dump(char* buf, size_t buf_bytes);

#ifdef __cplusplus
}
#endif
#endif /* VAD_B200_H_ */
