/*
 * deepsee_b200 C ABI  —  the drop-in boundary for DeepSEE's data-parallel hot path on B200.
 *
 * Every entry point takes raw DEVICE pointers, explicit shapes and a cudaStream_t (passed as
 * void*), enqueues its work on that stream and returns without synchronising the host.
 *   return 0   ok
 *   return <0  invalid argument / unsupported device (message via dsee_last_error())
 *   return >0  CUDA runtime error code (message via dsee_last_error())
 * The library never allocates, frees or retains user-visible device memory: inputs, outputs
 * and workspaces are owned by the caller (PyTorch's caching allocator).  No CPU fallback exists;
 * on a non-sm_100 device every compute entry point fails with -2.
 *
 * Layouts.  Feature maps inside the path are NHWC ("pixels x channels"); the reference's
 * NCHW fp32 tensors only appear at the two ends (stem input, image-head output, discriminator
 * input).  Tensor-core operands are 16-bit "split planes": value = hi + lo with
 * hi = fp16(value), lo = fp16(value - hi)  (3-pass product hi*hi + lo*hi + hi*lo ~ fp32
 * accuracy; 1-pass hi*hi ~ TF32 accuracy).
 *
 * "Replaces" cites the reference interface (file:line under mcbuehler/DeepSEE) that each entry
 * point stands in for.  The reference has no FFI of its own (it is 100 % Python calling ATen),
 * so the binding a maintainer adds is the ctypes stub shown in INTEGRATION.md.
 */
#ifndef DEEPSEE_B200_H
#define DEEPSEE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSEE_ABI_VERSION 4

/* ---- library ------------------------------------------------------------------------------ */
int dsee_version(void);
/* thread-local, valid until the next failing call on this thread */
const char* dsee_last_error(void);
/* number of kernels this library has launched from this process (all threads) */
int64_t dsee_launch_count(void);

/* ---- label maps (bit-exact integer work) --------------------------------------------------- */
/* Replaces data/preprocessor.py:35-41 (Preprocessor.preprocess_label):
 * out[b,c,y,x] = (label[b,0,y,x] == c) ? 1.0f : 0.0f ; label int64 [B,1,H,W], out fp32 [B,L,H,W].
 * A label outside [0,L) sets *bad_flag (device int) to 1 and writes zeros for that pixel. */
int dsee_onehot_from_labels(const int64_t* label, float* out, int B, int L, int H, int W,
                            int* bad_flag, void* stream);
/* Inverse: fp32 one-hot [B,L,H,W] -> uint8 label map [B,H,W].  Pixels that are not exactly
 * one-hot set *bad_flag to 1 (the conditional-norm kernels rely on the map being one-hot;
 * normalization.py:110-114,174-185 consume it as conv input / region mask). */
int dsee_labels_from_onehot(const float* onehot, uint8_t* labels, int B, int L, int H, int W,
                            int* bad_flag, void* stream);
/* Replaces F.interpolate(segmap, size, mode='nearest') normalization.py:110,174,261 and
 * encoder.py:38,126 on the label map: out[b,y,x] = in[b, floor(y*Hin/Hout), floor(x*Win/Wout)]. */
int dsee_resize_labels(const uint8_t* in, uint8_t* out, int B, int Hin, int Win, int Hout,
                       int Wout, void* stream);

/* ---- noise injection ---------------------------------------------------------------------------- */
/* NoiseInjection (normalization.py:299-304) draws a fresh N(0,1) tensor per forward.  Every entry
 * point that takes a `noise` tensor also takes a `noise_seed`: with noise == NULL and a non-zero
 * seed the tensor's elements are regenerated on the fly (Philox4x32-7 keyed by the seed, counter =
 * NHWC element index / 4, Box-Muller), identically in every kernel, and never touch HBM.
 * dsee_noise_fill materialises exactly that tensor (tests; n % 4 == 0). */
int dsee_noise_fill(unsigned long long seed, float* out, int64_t n, void* stream);

/* ---- conditional-norm operand builders ------------------------------------------------------ */
/* Replaces mlp_shared = Conv2d(L, nh, 3, pad 1) + ReLU over the one-hot map
 * (normalization.py:98-101,114 / 149-152,175 / 239-242,262) as a 9-tap table gather:
 *   actv[b,y,x,o] = relu(bias[o] + sum_tap table[tap][labels[b, yl+dy, xl+dx]][o]),
 *   (yl,xl) = (y >> ups, x >> ups)   [ups=1 reproduces F.interpolate(actv, size=out_size),
 *   normalization.py:188-189, for feature maps larger than max_fm_size]
 * table fp32 [9][L][nh] (= weight[o][l][ky][kx] transposed), labels uint8 [B,Hl,Wl],
 * out_hi/out_lo fp16 NHWC [B, Hl<<ups, Wl<<ups, nh]. out_lo may be NULL.
 * uniform_rows: optional scratch fp32 [L][nh]; when given, the call first fills it with the pre-activation
 * of a uniform 3x3 window per label (same additions, same order) and pixels whose window carries one
 * label - the interior of every region of a parse map - read one row instead of nine.  Bit-identical
 * results with or without it. */
int dsee_shared_mlp_fwd(const uint8_t* labels, const float* table, const float* bias,
                        void* out_hi, void* out_lo, int B, int Hl, int Wl, int ups, int L, int nh,
                        float* uniform_rows, void* stream);
/* Replaces style_map = (style[:,:,:,None,None] * seg[:,:,None]).sum(1)
 * (normalization.py:182-185 / 269-272): out[b,y,x,:] = style[b, labels[b,y,x], :].
 * style fp32 [B,L,d]; out_hi/out_lo fp16 NHWC [B,H,W,d]. */
int dsee_style_gather_fwd(const uint8_t* labels, const float* style, void* out_hi, void* out_lo,
                          int B, int H, int W, int L, int d, void* stream);

/* Input side of the path.
 * dsee_labels_u8: the dataloader's int64 label map -> the uint8 map every kernel consumes;
 *   *bad_flag = 1 if a label is outside [0, L) (the reference's scatter_ raises there,
 *   data/preprocessor.py:40).  The fp32 one-hot tensor is only materialised if a caller asks for it.
 * dsee_bicubic_clamp: Preprocessor.downsample_image (data/preprocessor.py:17-33):
 *   F.interpolate(hr, (Ho, Wo), mode='bicubic') (align_corners=False, A = -0.75, no antialias)
 *   followed by clamp(-1, 1); fp32 NCHW in and out. */
int dsee_labels_u8(const int64_t* label, uint8_t* out, int64_t n, int L, int* bad_flag, void* stream);
int dsee_bicubic_clamp(const float* in, float* out, int B, int C, int Hi, int Wi, int Ho, int Wo,
                       void* stream);

/* VGG19 perceptual loss (loss.py:104-119, architecture.py:151-181): its 3x3 convs run on
 * dsee_conv2d_tc / dsee_conv2d_direct_fwd with the ReLU fused (dsee_conv_epilogue.lrelu = 2;
 * dsee_act_bwd act = 3 in the backward pass); nn.MaxPool2d(2, 2) is this pair (NHWC fp32, C % 4 == 0;
 * the backward pass recomputes the arg-max from `in`, first maximum in window order like ATen). */
int dsee_maxpool2_fwd(const float* in, float* out, int B, int Hi, int Wi, int C, void* stream);
int dsee_maxpool2_bwd(const float* in, const float* dout, float* din, int B, int Hi, int Wi, int C,
                      void* stream);

/* Sync-BN statistics exchange over NVLink peer memory (replaces the reference's master/slave pipes,
 * sync_batchnorm/batchnorm.py:80-93 + comm.py): all-reduce(sum) of a vector of n <= maxn floats
 * across the `world` GPUs of one node in ONE kernel - push my vector into every peer's slot, release
 * a sequence flag, acquire-wait for all peers, sum in rank order from local memory.  peer_bufs[r] is
 * rank r's symmetric allocation (dsee_peer_exchange_bytes(world, ring, maxn) bytes, zero-initialised,
 * e.g. torch.distributed._symmetric_memory) as mapped in this process.  Every rank must issue the
 * same sequence of calls.  in == out is allowed.  Capturable in a CUDA graph. */
int64_t dsee_peer_exchange_bytes(int world, int ring, int maxn);
int dsee_peer_allreduce_small(const void* const* peer_bufs, int world, int rank, int ring, int maxn,
                              const float* in, float* out, int n, void* stream);

/* Noise epoch: a per-device 64-bit counter that every kernel regenerating NoiseInjection noise
 * from a seed folds into that seed.  Advancing it (a one-thread kernel on `stream`) makes launches
 * whose seeds are baked into a captured CUDA graph draw fresh noise on every replay; forward and
 * backward kernels launched between two advances see the same tensor.  Starts at 0. */
int dsee_noise_epoch_advance(void* stream);
int dsee_noise_epoch_set(unsigned long long value, void* stream);

/* ---- tensor-core operand preparation -------------------------------------------------------- */
/* Conv weight fp32 [N][C][3][3] (PyTorch layout; spectral normalisation already applied,
 * architecture.py:40-44) -> GEMM B-operand planes fp16 [N][9*C] with k = (ky*3+kx)*C + c,
 * multiplied by a power of two 2^e chosen so max|w|*2^e is in [2^13, 2^14) (keeps the lo plane
 * out of the fp16 subnormal range). inv_scale is a device float[3]: [0] receives 2^-e (what the
 * conv kernels read), [1] max|w|, [2] (transpose=1 only) max over rows of sum |w| along the row =
 * the factor that bounds the backward-data result (dsee_dgrad_modulate_bwd.w_l1).
 * transpose=1 builds the backward-data operand instead (autograd of F.conv2d wrt its input =
 * convolution of the output gradient with the transposed, 180-degree-rotated filter):
 * planes [C][9*N] with k = (8-tap)*N + n. */
int dsee_prep_conv_weight(const float* w, void* out_hi, void* out_lo, float* inv_scale, int N,
                          int C, int transpose, void* stream);
/* fp8 companion of the planes above for dsee_conv_operands.passes == 2 (same reference weight,
 * architecture.py:40-44): with ws = w * 2^e (inv_scale[0] = 2^-e must already hold the value
 * dsee_prep_conv_weight wrote for this weight) and hi = fp16(ws):
 *   out8[n][tap][0][c] = e4m3(ws * 2^-8)    (pairs with the activation's lo plane * 2^8)
 *   out8[n][tap][1][c] = e4m3(ws - hi)      (pairs with the activation itself)
 * out8: [N][9][2][C] bytes, C a multiple of 128. */
int dsee_prep_conv_weight_f8(const float* w, const float* inv_scale, void* out8, int N, int C,
                             void* stream);
/* Modulation weight of a SEAN layer with the style branch folded into per-image weights
 * (normalization.py:182-185,198-213): style_map[b,:,y,x] = style[b, label(y,x), :] makes
 *   conv(style_map, W_sty)[b,o,y,x] = sum_{tap,l} onehot[b,l,(y,x)+tap] * Ws[b,o,l,tap],
 *   Ws[b,o,l,tap] = sum_s W_sty[o,s,tap] * style[b,l,s]
 * so K1 reads the exact one-hot planes (Lp channels, Lp = 64) instead of a gathered 128-channel
 * style map: K per tap 256 -> 192.  wa fp32 [N][Ca][3][3] (the mlp_shared-activation columns, shared
 * by all images), ws fp32 [B][N][Ls][3][3] (Ls <= Lp label columns) -> planes [B*N][9*(Ca+Lp)],
 * k = tap*(Ca+Lp) + c (zeros for Ls <= c - Ca < Lp), one power-of-two scale for everything;
 * inv_scale: device float[2] = (2^-e, max|w|). */
int dsee_prep_mod_weight_batched(const float* wa, const float* ws, void* out_hi, void* out_lo,
                                 float* inv_scale, int B, int N, int Ca, int Ls, int Lp, void* stream);

/* fp32 NHWC [rows][C] -> fp16 split planes (used for tensors not produced by a fused epilogue). */
int dsee_split_f16(const float* in, void* out_hi, void* out_lo, int64_t n, void* stream);

/* fp32 NHWC [B,H,W,C] -> fp16 split planes [B,2H,2W,C] with the nn.Upsample(scale_factor=2) in
 * front of the encoder's up_conv / conv2 (encoder.py:94-95,153-154) materialised, and its
 * transpose for the backward pass: out[b,y,x,c] = sum of the 2x2 block of in [B,2Ho,2Wo,C]. */
int dsee_split_f16_ups2(const float* in, void* out_hi, void* out_lo, int B, int H, int W, int C,
                        void* stream);
int dsee_fold2x2(const float* in, float* out, int B, int Ho, int Wo, int C, void* stream);

/* ---- the fused tensor-core kernels ---------------------------------------------------------- */
typedef struct {
    /* geometry: stride-1 3x3 conv, zero padding 1, output size == input size */
    int B, H, W;
    /* A operand: 1 or 2 NHWC fp16 split-plane tensors concatenated along channels
     * (channels multiple of 64; a_channels[1] may be 0). lo planes may be NULL iff passes==1. */
    const void* a_hi[2];
    const void* a_lo[2];
    int a_channels[2];
    /* B operand from dsee_prep_conv_weight: [n_total][9*(a_channels[0]+a_channels[1])] */
    const void* w_hi;
    const void* w_lo;
    const float* w_inv_scale; /* device scalar */
    int n_total;
    int passes; /* 1: hi*hi (TF32-class accuracy)   3: hi*hi + lo*hi + hi*lo (fp32-class)
                   2: hi*hi + an fp8 correction GEMM for both operand-rounding terms (a8_lo * w8[0] +
                      a8_hi * w8[1], tcgen05 kind::f8f6f4 at twice the fp16 rate, accumulated into the
                      same TMEM accumulator): the cost of two fp16 passes, ~5 % of the 1-pass error.
                      dsee_conv3x3_fwd only; one A source. */
    int a_dtype; /* element type of the A planes: 0 fp16, 1 bf16 */
    int w_dtype; /* element type of the weight planes; must equal a_dtype (tcgen05 kind::f16
                    rejects mixed A/B formats) */
    const float* a_inv_scale; /* NULL, or device scalar 2^-e of pre-scaled A planes (gradient
                                 operands from dsee_grad_prep / dsee_spade_modulate_bwd) */
    /* passes == 2 only: e5m2 NHWC [B,H,W,C] planes written by dsee_spade_modulate_fwd
     * (a8_lo = (a - a_hi) * 2^8, a8_hi = a) and the e4m3 weight from dsee_prep_conv_weight_f8
     * ([n_total][9][2][C]: per tap the hi weights * 2^-8, then the lo residual of the fp16 plane). */
    const void* a8_lo;
    const void* a8_hi;
    const void* w8;
    /* 0: one weight matrix for all images.  > 0 (= n_total): per-image weights, the planes hold
     * [B * n_total][K] and image b reads rows b * w_batch_rows ... (the SEAN style branch folded into
     * per-image weights over the one-hot label planes, dsee_prep_mod_weight_batched) */
    int w_batch_rows;
    /* Sub-pixel form of a 3x3 conv over a nearest-2x-upsampled tensor (normalization.py:188-190,
     * 275-277: above max_fm_size the style branch convolves the upsampled mlp_shared activation):
     * the A planes hold the tensor at HALF resolution [B, H/2, W/2, C] and this call computes the
     * output pixels of one parity class (y % 2, x % 2) = (sub_py, sub_px) as a 2x2-tap conv over the
     * half-resolution planes: 4/9 of the FLOPs, and the upsampled tensor is never built.  The weight
     * planes are the class's collapsed filter [n_total][4 * C_total], k = (ty*2+tx)*C_total + c, with
     *   wc[ty][tx] = sum of w[ky][kx] over the taps with floor((sub_py + ky - 1) / 2) == ty + (sub_py - 1)
     * (likewise in x); H, W stay the OUTPUT size (even).  0 = ordinary 3x3 conv.
     * sub_py < 0: all four classes in ONE launch - the weight planes then hold the four collapsed
     * filters one after the other, [4 * n_total][4 * C_total] (class = py*2 + px), and the class is a
     * tile index next to the pixel tile.  dsee_spade_modulate_fwd only. */
    int a_sub, sub_py, sub_px;
} dsee_conv_operands;

/* Opt-in CTA-pair form of dsee_conv3x3_fwd (tcgen05 cta_group::2: two SMs of a TPC share one 256-row
 * weight box; used when n_total % 256 == 0 and H % 16 == 0).  on = 1 / 0 sets the process-wide mode,
 * on < 0 only queries; returns the previous mode.  Results are bit-identical to the single-CTA form
 * (same accumulation order per output element).  No reference counterpart (cuDNN picks its own
 * kernels). */
int dsee_conv_pair_mode(int on);

/* K2.  Replaces conv_0 / conv_1 of SPADEResnetBlock (architecture.py:34-35,98,122) plus what the
 * reference does to the conv output before the next layer reads it:
 *   - the residual add `out = x_s + dx` (architecture.py:127), the shortcut read through a folded
 *     nn.Upsample(scale_factor=2) (sr.py:57,69,87) when res_ups=1;
 *   - NoiseInjection terms (normalization.py:299-304): noise_middle on conv_0's output
 *     (architecture.py:111-112), noise_in + noise_skip on the shortcut (architecture.py:76-79,133-134);
 *   - the first pass of the next batch norm (sync_batchnorm/batchnorm.py:72-76): per-channel sum and
 *     sum of squares of the final output, as deterministic tile partials.
 *   out[b,y,x,n] = bias[n] + sum_{tap,c} A[b,y+dy,x+dx,c] * w[n,c,tap]
 *                  (+ residual[b, y>>res_ups, x>>res_ups, n]) (+ sum_i noise_w[i][n]*noise[i][b,y,x,n])
 * out fp32 NHWC [B,H,W,n_total]; residual fp32 NHWC [B,H>>res_ups,W>>res_ups,n_total] or NULL;
 * noise[i] fp32 NHWC [B,H,W,n_total] or NULL; stats_partial NULL or fp32
 * [dsee_conv3x3_stats_tiles()][n_total][2], reduced by dsee_bn_finalize. */
typedef struct {
    const float* bias;
    const float* residual;
    int res_ups;
    const float* noise[2];
    const float* noise_w[2];
    float* out;
    float* stats_partial;
    /* backward-data use (the same kernel run on the output gradient with the transposed,
     * flipped weight): multiply by LeakyReLU'(t), 1 where act_mask > 0 else 0.2
     * (architecture.py:147 backward). act_mask = the hi plane K1 wrote in the forward pass,
     * fp16 NHWC [B,H,W,n_total], or NULL. bias may be NULL (= 0). */
    const void* act_mask;
    /* optional device float: receives max |out| (what the next gradient-plane scale is chosen
     * from); zeroed by the call. */
    float* amax_out;
    /* activation fused after the bias: 1 = LeakyReLU(0.2) (discriminator.py:84-85), 2 = ReLU (VGG19) */
    int lrelu;
    /* when noise[i] is NULL and noise_seed[i] != 0 the noise tensor is regenerated in the kernel
     * from the counter-based generator (see dsee_noise_fill) instead of being read from HBM */
    unsigned long long noise_seed[2];
    /* optional second output for the tensor-core image head: fp16 hi / lo planes of
     * leaky_relu(out, 0.2) (sr.py:94: `F.leaky_relu(x, 2e-1)` in front of conv_img), NHWC
     * [B,H,W,n_total]; act16_lo may be NULL.  dsee_conv3x3_fwd only. */
    void* act16_hi;
    void* act16_lo;
} dsee_conv_epilogue;
int dsee_conv3x3_fwd(const dsee_conv_operands* ops, const dsee_conv_epilogue* epi, void* stream);
int dsee_conv3x3_stats_tiles(int B, int H, int W);

/* General KH x KW (<= 16 taps), stride 1 / 2 convolution on the same tcgen05 kernel: the style
 * encoder's 3x3 stride-1/2 layers (encoder.py:84-98,142-157) and the discriminator's 4x4
 * stride-1/2 layers (discriminator.py:84-96).  A stride-2 tap is a TMA box with element stride 2.
 *   transposed = 0:  out[b,yo,xo,n] = bias[n] + sum A[b, yo*stride - pad + ky, xo*stride - pad + kx, c]
 *                                               * w[n,c,ky,kx]          (out [B,Ho,Wo,n_total])
 *   transposed = 1:  backward-data of that conv: A = dY planes [B,Hi,Wi,Ci] (Hi, Wi = the forward
 *                    output size, Ci = forward Cout), out = dX [B,Ho,Wo,n_total] (the forward input
 *                    size / channels); stride 2 runs as 4 output-parity classes.
 * A channels only need to be a multiple of 8: TMA zero-fills the 64-channel K block beyond Ci and
 * dsee_prep_conv_weight_ex pads the weight K blocks with zeros.
 * Epilogue: bias, lrelu, act_mask (stride 1 only), amax_out, out; no residual / noise / stats. */
typedef struct {
    int B, Hi, Wi;
    const void* a_hi;
    const void* a_lo;
    int Ci;
    const float* a_inv_scale;
    int KH, KW, stride, pad;
    const void* w_hi;
    const void* w_lo;
    const float* w_inv_scale;
    int n_total;
    int passes;
    int transposed;
    int Ho, Wo;
} dsee_conv2d_tc_args;
int dsee_conv2d_tc(const dsee_conv2d_tc_args* args, const dsee_conv_epilogue* epi, void* stream);
/* Weights for dsee_conv2d_tc / dsee_conv2d_tc_wgrad: fp32 [N][C][KH][KW] -> scaled fp16 planes
 *   transpose = 0: [N][KH*KW*Cp], k = (ky*KW+kx)*Cp + c, Cp = C rounded up to 64 (zeros beyond C)
 *   transpose = 1: [C][KH*KW*Np], k = (ky*KW+kx)*Np + n, Np = N rounded up to 64 (not rotated:
 *                  the tap offsets of the transposed conv carry the rotation). */
int dsee_prep_conv_weight_ex(const float* w, void* out_hi, void* out_lo, float* inv_scale, int N,
                             int C, int KH, int KW, int transpose, void* stream);

/* dsee_conv3x3_wgrad2 with one result per image (the per-image weights above):
 * dw fp32 [B][n_total][c_total][3][3]; the pixel splits never straddle an image. */
int64_t dsee_conv3x3_wgrad_per_image_workspace_floats(int B, int H, int W, int n_total, int c_total);
int dsee_conv3x3_wgrad2_per_image(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale,
                                  const void* const* a_hi, const void* const* a_lo, const int* a_channels,
                                  int dtype, int B, int H, int W, int n_total, int passes,
                                  float* workspace, float* dw, void* stream);

/* Backward of the sub-pixel form (dsee_conv_operands.a_sub) for one parity class:
 * dsee_subpixel_wgrad: dwc[n][c][ty][tx] = sum over the class's output pixels (2Y+py, 2X+px) of
 *   dY[b,2Y+py,2X+px,n] * A[b, Y + ty + py - 1, X + tx + px - 1, c]   (A at half resolution, 1 or 2
 *   channel-concatenated sources; dY planes at full resolution [B,H,W,n_total]); dwc fp32
 *   [n_total][C_total][2][2]; workspace from dsee_subpixel_wgrad_workspace_floats.
 * dsee_subpixel_dgrad: gradient wrt the half-resolution A, all four classes in one GEMM:
 *   dA[b,Y,X,c] = sum_{class, tap, n} dY[b, 2(Y - ry) + py, 2(X - rx) + px, n] * wc[class][n][c][tap],
 *   weights as planes [C][16 * n_total] from dsee_prep_conv_weight_ex on a [C][n_total][4][4] tensor
 *   (KH index = class py*2+px, KW index = tap ty*2+tx); out fp32 [B,H/2,W/2,C]; optional amax_out. */
int64_t dsee_subpixel_wgrad_workspace_floats(int B, int H, int W, int n_total, int c_total);
int dsee_subpixel_wgrad(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale,
                        const void* const* a_hi, const void* const* a_lo, const int* a_channels, int B,
                        int H, int W, int n_total, int sub_py, int sub_px, int passes, float* workspace,
                        float* dwc, void* stream);
int dsee_subpixel_dgrad(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale, const void* w_hi,
                        const void* w_lo, const float* w_inv_scale, int B, int H, int W, int n_total,
                        int C, int passes, float* out, float* amax_out, void* stream);

/* Weight gradient of dsee_conv2d_tc: dY planes [B,Ho,Wo,n_total], activation planes
 * [B,Hi,Wi,Ci] (the forward input) -> dw fp32 [n_total][Cp][KH][KW], Cp = Ci rounded up to 64
 * (columns beyond Ci are zero).  workspace fp32 [dsee_conv2d_tc_wgrad_workspace_floats()]. */
int64_t dsee_conv2d_tc_wgrad_workspace_floats(int B, int Ho, int Wo, int n_total, int Ci, int KH,
                                              int KW);
int dsee_conv2d_tc_wgrad(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale,
                         const void* a_hi, const void* a_lo, const float* a_inv_scale, int B, int Ho,
                         int Wo, int Hi, int Wi, int n_total, int Ci, int KH, int KW, int stride,
                         int pad, int passes, float* workspace, float* dw, void* stream);

/* K1.  Replaces SPADE.forward (normalization.py:105-120), SEAN_Block.forward (:167-213) and
 * PureSEAN_Block.forward (:254-286) together with the following actvn (architecture.py:96,114,147):
 *   [gamma|beta][b,y,x,c] = conv3x3(A, w)        (n_total = 2*C; rows interleaved per 128 channels:
 *                                                  [g(0..127) | b(0..127) | g(128..255) | ...])
 *   xin  = x[b, y>>x_ups, x>>x_ups, c] (+ noise_w[c] * noise[b,y,x,c])     (normalization.py:299-304)
 *   xhat = xin * bn_scale[c] + bn_shift[c]                                  (batchnorm.py:66,78-93)
 *   t    = xhat * (gamma + gamma_bias[c]) + (beta + beta_bias[c])
 *   out  = leaky_relu(t, 0.2) as fp16 split planes NHWC [B,H,W,C]
 * gamma_bias carries the conv bias, the SEAN alpha blend of the two biases and the "+1"
 * (absent for PureSEAN, normalization.py:286). The alpha blend of the weights
 * (normalization.py:208-212) is folded into w by the caller. */
typedef struct {
    const float* x;      /* fp32 NHWC [B, H>>x_ups, W>>x_ups, C] */
    int x_ups;           /* 0 or 1 */
    const float* noise;  /* fp32 NHWC [B,H,W,C] or NULL */
    const float* noise_w;/* fp32 [C] or NULL */
    const float* bn_scale;
    const float* bn_shift;
    const float* gamma_bias;
    const float* beta_bias;
    void* out_hi;        /* fp16 NHWC [B,H,W,C] */
    void* out_lo;        /* may be NULL */
    int C;               /* multiple of 128 */
    /* optional (training): G = gamma + gamma_bias saved as fp16 split planes NHWC [B,H,W,C] so
     * the backward pass (dsee_spade_modulate_bwd_saved) does not re-run the gamma GEMM */
    void* g_hi;
    void* g_lo;
    unsigned long long noise_seed; /* used when noise == NULL and noise_w != NULL */
    /* optional: e5m2 NHWC [B,H,W,C] planes of the activation for a consumer that runs the fp8
     * correction (dsee_conv_operands.passes == 2): out8_lo = (t - fp16(t)) * 2^8, out8_hi = t */
    void* out8_lo;
    void* out8_hi;
} dsee_modulate_args;
int dsee_spade_modulate_fwd(const dsee_conv_operands* ops, const dsee_modulate_args* mod,
                            void* stream);

/* K1 backward.  Re-runs the gamma half of the modulation GEMM (weights: the gamma rows only,
 * n_total == C, prepared with dsee_prep_conv_weight) and, with G = gamma + gamma_bias,
 * xhat = xin*bn_scale + bn_shift, dt = dL/dt:
 *   dxhat = dt * G                                 -> fp32 NHWC [B,H,W,C]
 *   [dG | dB] = [dt * xhat | dt] * 2^e             -> fp16 split planes NHWC [B,H,W,2C], channels
 *                                                     interleaved per 128 like the forward weights
 *                                                     (A operand of the two following GEMMs);
 *                                                     2^e is chosen from *dt_amax (max |dt|, e.g.
 *                                                     dsee_conv_epilogue.amax_out of the kernel that
 *                                                     produced dt), 2^-e is written to *dgb_inv_scale
 *   partial[tile][c] = (sum dxhat, sum dxhat*xhat, sum dG, sum dB) over the tile's pixels
 *     (first two: batch-norm backward, batchnorm.py:78-93 differentiated; last two: gradients of
 *      gamma_bias / beta_bias). partial: fp32 [dsee_conv3x3_stats_tiles()][C][4]. */
typedef struct {
    const float* x;
    int x_ups;
    const float* noise;
    const float* noise_w;
    const float* bn_scale;
    const float* bn_shift;
    const float* gamma_bias;
    const float* dt;
    float* dxhat;
    void* dgb_hi;
    void* dgb_lo; /* may be NULL */
    float* partial;
    int C;
    const float* dt_amax;
    float* dgb_inv_scale;
} dsee_modulate_bwd_args;
int dsee_spade_modulate_bwd(const dsee_conv_operands* ops, const dsee_modulate_bwd_args* args,
                            void* stream);

/* Backward-data of a main conv (architecture.py:98,122) FUSED with K1's backward (autograd of
 * normalization.py:105-120,167-213,254-286 + architecture.py:147's LeakyReLU): `ops` are the gradient
 * planes of the conv output and the weight from dsee_prep_conv_weight(transpose=1); the accumulator
 * dt = conv_transpose(dY, W) * LeakyReLU'(t) never reaches HBM.  The epilogue reads the saved
 * activation's sign (act_mask), the saved G = gamma + gamma_bias planes and x, and writes
 *   dxhat fp32 NHWC = dt * G,   dgb planes [B,H,W,2C] = [dt * xhat | dt] (scaled fp16, channels
 *   interleaved per 128 like the modulation weight rows), partial [dsee_conv3x3_stats_tiles()][C][4]
 *   = tile sums of (dxhat, dxhat*xhat, dG, dB).
 * The plane scale comes from a bound available before the GEMM runs: |dt| <= dy_amax * w_l1 with
 * dy_amax = max|dY| (inv_scale[1] of dsee_grad_prep) and w_l1 = max over input channels of
 * sum_{n,tap} |W| (inv_scale[2] of dsee_prep_conv_weight(transpose=1)). */
typedef struct {
    const void* act_mask;       /* fp16 hi plane of the forward activation a = LeakyReLU(t) */
    const void* g_hi;           /* saved G planes (dsee_modulate_args.g_hi / g_lo) */
    const void* g_lo;           /* NULL for 1-pass */
    const float* x;             /* K1's x input (fp32 NHWC, at H >> x_ups) */
    int x_ups;
    const float* noise;         /* NoiseInjection on x: tensor, or NULL with noise_seed != 0 */
    unsigned long long noise_seed;
    const float* noise_w;
    const float* bn_scale;
    const float* bn_shift;
    const float* dy_amax;       /* device scalar */
    const float* w_l1;          /* device scalar */
    float* dxhat;
    void* dgb_hi;
    void* dgb_lo;
    float* dgb_inv_scale;       /* out: 2^-e of the dgb planes */
    float* partial;
    int C;
} dsee_dgrad_modbwd_args;
int dsee_dgrad_modulate_bwd(const dsee_conv_operands* ops, const dsee_dgrad_modbwd_args* args,
                            void* stream);

/* ---- generator backward (autograd of the fused kernels above) -------------------------------- */
/* Gradient tensors are fp32 NHWC in HBM; as tensor-core operands they are fp16 split planes like
 * the activations (tcgen05 kind::f16 needs A and B in the same format), multiplied by a per-tensor
 * power of two 2^e that places max|g| in [2^13, 2^14); the 2^-e travels with the planes as a device
 * scalar (dsee_conv_operands.a_inv_scale / dsee_conv3x3_wgrad's dy_inv_scale).
 *
 * Backward-data of K2 / K1's GEMM is dsee_conv3x3_fwd itself, run on the gradient planes with a
 * weight from dsee_prep_conv_weight(transpose=1); dsee_conv_epilogue.act_mask folds in LeakyReLU'. */

/* dY fp32 NHWC [npix][C] -> scaled fp16 split planes (inv_scale: device float[2], [0] receives
 * 2^-e, [1] is scratch), plus per-channel block partials of
 * (sum dY, sum dY*noise0, sum dY*noise1) = gradients of a conv bias (architecture.py:98,122) and of
 * NoiseInjection.weight (normalization.py:299-304).  partial fp32 [dsee_grad_prep_blocks()][C][nq],
 * nq = 1 + (noise0 or seed0 given) + (noise1 or seed1 given); reduce with dsee_reduce_partials.
 * amax_in: NULL (max|dY| is computed here, one extra read of dY), or a device scalar holding it
 * (dsee_bn_bwd.amax_out); either way inv_scale[1] holds max|dY| afterwards. */
int dsee_grad_prep_blocks(int64_t npix);
int dsee_grad_prep(const float* dy, void* out_hi, void* out_lo, float* inv_scale,
                   const float* noise0, const float* noise1, unsigned long long seed0,
                   unsigned long long seed1, int64_t npix, int C, float* partial, const float* amax_in,
                   void* stream);
/* out[k][c] = scale * sum_s partial[s][c][k]  (double accumulation, fixed order). */
int dsee_reduce_partials(const float* partial, int n, int C, int nq, float scale, float* out,
                         void* stream);
/* Weight gradient of a 3x3 conv (autograd of architecture.py:98,122 and normalization.py:116-117,
 * 198-201,283-284): dW[n][c][tap] = s * sum_{b,y,x} dY[b,y,x,n] * A[b,y+dy,x+dx,c], s = the product
 * of *dy_inv_scale and *a_inv_scale (device scalars, NULL = 1).
 * dY planes [B,H,W,n_total] (n_total % 128 == 0), A planes [B,H,W,c_total] (c_total = 64, 128 or a
 * multiple of 256); dtype 0 fp16 / 1 bf16, the same for both operands.  workspace fp32
 * [dsee_conv3x3_wgrad_workspace_floats()]; dw fp32 [n_total][c_total][3][3] (layout_nc9 = 1, the
 * PyTorch layout) or [n_total][9][c_total] (layout_nc9 = 0).  Deterministic split-K. */
int64_t dsee_conv3x3_wgrad_workspace_floats(int B, int H, int W, int n_total, int c_total);
int dsee_conv3x3_wgrad(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale,
                       const void* a_hi, const void* a_lo, const float* a_inv_scale, int dtype, int B,
                       int H, int W, int n_total, int c_total, int passes, float* workspace, float* dw,
                       int layout_nc9, void* stream);
/* The same with the activation operand given as up to two channel-concatenated sources (the
 * [actv | style_map] input of a SEAN modulation conv, normalization.py:198-201): one launch with a
 * 256-wide MMA instead of two 128-wide ones.  a_channels[i] % 64 == 0, a_channels[1] may be 0. */
int dsee_conv3x3_wgrad2(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale,
                        const void* const* a_hi, const void* const* a_lo, const int* a_channels,
                        int dtype, int B, int H, int W, int n_total, int passes, float* workspace,
                        float* dw, int layout_nc9, void* stream);
/* Batch-norm backward (differentiates batchnorm.py:78-93 / F.batch_norm in training mode) with the
 * folded 2x upsample transposed into a 2x2 sum and the identity shortcut's gradient added:
 *   xhat = (x[up] + noise_w*noise) * bn_scale + bn_shift
 *   dxin = bn_scale * (dxhat - inv_count*sums[0] - xhat * inv_count*sums[1])
 *   dx[b,y',x',c] = sum over the 2^ups x 2^ups block of (dxin + dskip)
 * sums fp32 [2][C] = (sum dxhat, sum dxhat*xhat) (pass inv_count = 0 for eval-mode statistics);
 * dskip fp32 NHWC [B,Hx<<ups,Wx<<ups,C] or NULL; nw_partial NULL or fp32
 * [dsee_bn_bwd_blocks()][C] = block partials of sum(dxin*noise) (gradient of noise_in.weight), or of
 * sum((dxin+dskip)*noise) when noise_grad_with_skip != 0 (the shortcut of architecture.py:76-79 adds
 * the same noise term, so its weight gradient can ride along instead of regenerating the noise in
 * dsee_grad_prep).  amax_out: NULL, or a device scalar that receives max|dx| (what dsee_grad_prep
 * would otherwise compute with a pass of its own over dx). */
int dsee_bn_bwd_blocks(int B, int Hx, int Wx);
int dsee_bn_bwd(const float* dxhat, const float* x, int x_ups, const float* noise,
                unsigned long long noise_seed, const float* noise_w, const float* bn_scale, const float* bn_shift,
                const float* sums, float inv_count, const float* dskip, int B, int Hx, int Wx, int C,
                float* dx, float* nw_partial, int noise_grad_with_skip, float* amax_out, void* stream);
/* Backward of dsee_shared_mlp_fwd: gradient of the 9-tap table and bias.  dsrc fp32 NHWC with row
 * stride ld, the actv gradient in columns [coff, coff+nh).  partial fp32
 * [dsee_shared_mlp_bwd_blocks()][9*L+1][nh]; dtable_dbias fp32 [9*L+1][nh] (last row = bias). */
int dsee_shared_mlp_bwd_blocks(int B, int Hl, int Wl);
int dsee_shared_mlp_bwd(const float* dsrc, int ld, int coff, const void* actv_hi,
                        const uint8_t* labels, int B, int Hl, int Wl, int ups, int L, int nh,
                        float* partial, float* dtable_dbias, void* stream);
/* The same gradient on the tensor cores: mlp_shared is a 3x3 conv over the one-hot map, so
 * d table = dsee_conv3x3_wgrad(G planes, one-hot planes) with
 *   G[b,yl,xl,o] = relu'(actv) * (sum over the 2^ups x 2^ups copies of dsrc[..., coff+o]),
 * emitted here as scaled fp16 split planes [B,Hl,Wl,nh] (scale from *dsrc_amax = max|dsrc|, e.g.
 * dsee_conv_epilogue.amax_out of the kernel that produced dsrc; 2^-e -> inv_scale[0]) together with
 * block partials fp32 [dsee_grad_prep_blocks(B*Hl*Wl)][nh] of sum G (the bias gradient).
 * dsee_onehot_planes writes the other operand: fp16 [npix][Lp] (Lp % 8 == 0, zeros beyond L). */
int dsee_actv_grad_prep(const float* dsrc, int ld, int coff, const void* actv_hi,
                        const float* dsrc_amax, int B, int Hl, int Wl, int ups, int nh, void* out_hi,
                        void* out_lo, float* inv_scale, float* partial, void* stream);
int dsee_onehot_planes(const uint8_t* labels, void* out, int64_t npix, int Lp, void* stream);
/* Backward of dsee_style_gather_fwd: dstyle[b,l,:] = sum_{p: labels[b,p]==l} dsrc[b,p,coff:coff+d].
 * workspace fp32 [B][dsee_region_pool_chunks(HW)][L][d]. */
int dsee_style_gather_bwd(const float* dsrc, int ld, int coff, const uint8_t* labels, float* dstyle,
                          float* workspace, int B, int HW, int L, int d, void* stream);
/* Backward of dsee_stem_fwd (weights only; the LR image needs no gradient).
 * partial fp32 [dsee_stem_bwd_blocks()][C][28]; dw_db fp32 [C][28] = 27 weight grads + bias grad. */
int dsee_stem_bwd_blocks(int B, int H, int W);
int dsee_stem_bwd(const float* x, const float* dy, int B, int H, int W, int C, float* partial,
                  float* dw_db, void* stream);
/* Backward of dsee_head_fwd.  out = the forward result, dout its gradient (NCHW [B,3,H,W]);
 * dx fp32 NHWC [B,H,W,C]; partial fp32 [dsee_head_bwd_blocks()][C][28];
 * dw_db fp32 [C][28]: [c][o*9+tap] = dW[o][c][tap], [c][27] = dbias[c] for c < 3. */
int dsee_head_bwd_blocks(int B, int H, int W);
int dsee_head_bwd(const float* x, const float* w, const float* out, const float* dout, int B, int H,
                  int W, int C, float* dx, float* partial, float* dw_db, void* stream);

/* K1 backward from the saved G planes (no GEMM; one streaming pass, HBM bound: reads x, dt, G,
 * writes dxhat and the dgb planes).  Same outputs as dsee_spade_modulate_bwd; partial fp32
 * [dsee_grad_prep_blocks(B*H*W)][C][4]. */
int dsee_spade_modulate_bwd_saved(const float* x, int x_ups, const float* noise,
                                  unsigned long long noise_seed, const float* noise_w,
                                  const float* bn_scale, const float* bn_shift, const void* g_hi,
                                  const void* g_lo, const float* dt, const float* dt_amax, int B, int H,
                                  int W, int C, float* dxhat, void* dgb_hi, void* dgb_lo,
                                  float* dgb_inv_scale, float* partial, void* stream);

/* ---- batch-norm statistics ------------------------------------------------------------------ */
/* Per-channel sum / sum of squares of x (+ noise) at the post-upsample resolution, written as
 * tile partials like the K2 epilogue does.  Replaces the statistics half of
 * SynchronizedBatchNorm2d.forward (batchnorm.py:72-76) for tensors K2 did not produce. */
int dsee_bn_stats(const float* x, int x_ups, const float* noise, unsigned long long noise_seed,
                  const float* noise_w, int B, int H, int W, int C, float* stats_partial,
                  int* n_partials, void* stream);
/* Reduces partials [n_partials][C][2] in a fixed order (double accumulation) and produces
 * bn_scale = 1/sqrt(var+eps), bn_shift = -mean*bn_scale; when running_mean/var are non-NULL also
 * updates them with momentum and the unbiased variance (batchnorm.py:84-93). count = number of
 * summed elements per channel; unbias_count = the n of var*n/(n-1) (differs from count when the
 * statistics of a 2x nearest-upsampled tensor are taken from its low-resolution source:
 * same mean and biased variance, n = 4*count); <= 0 means count. */
int dsee_bn_finalize(const float* stats_partial, int n_partials, int C, double count,
                     double unbias_count, float eps, float momentum, float* running_mean,
                     float* running_var, float* bn_scale, float* bn_shift, float* mean_out,
                     float* var_out, void* stream);
/* Eval mode: bn_scale/bn_shift from running statistics (batchnorm.py:65-68). */
int dsee_bn_eval_affine(const float* running_mean, const float* running_var, float eps, int C,
                        float* bn_scale, float* bn_shift, void* stream);

/* ---- generator ends ------------------------------------------------------------------------- */
/* Replaces DeepSEESR.initial (sr.py:31,65): conv 3->C 3x3 pad 1.
 * x fp32 NCHW [B,3,H,W]; w fp32 [C,3,3,3]; out fp32 NHWC [B,H,W,C]. */
int dsee_stem_fwd(const float* x, const float* w, const float* bias, float* out, int B, int H,
                  int W, int C, void* stream);
/* Replaces F.tanh(conv_img(F.leaky_relu(x, 0.2))) (sr.py:56,94-95).
 * x fp32 NHWC [B,H,W,C]; w fp32 [3,C,3,3]; out fp32 NCHW [B,3,H,W]. */
int dsee_head_fwd(const float* x, const float* w, const float* bias, float* out, int B, int H,
                  int W, int C, void* stream);

/* The same head on the tensor cores (sr.py:56,94-95): the last dsee_conv3x3_fwd of the generator writes
 * leaky_relu(x) as fp16 planes (dsee_conv_epilogue.act16_hi/lo), dsee_conv2d_tc multiplies them by the
 * 1x1 weight W27[tap*3+o][c] = w[o][c][tap] (27 rows padded to 32) into P fp32 NHWC [B,H,W,32], and
 *   dsee_head_gather_fwd:  out[b,o,y,x] = tanh(bias[o] + sum_tap P[b,y+dy,x+dx,tap*3+o])  (NCHW [B,3,H,W]).
 * Backward: dsee_head_scatter_bwd writes dP[b,y,x,tap*3+o] = (dout * (1 - out^2))[b,o,y-dy,x-dx]
 * (fp32 NHWC [B,H,W,32], columns 27..31 zero); the weight and data gradients are then
 * dsee_conv2d_tc_wgrad / dsee_conv2d_tc(transposed, act_mask = the hi plane) on dP's planes, and the
 * bias gradient is the column sum of the centre tap (columns 12..14). */
int dsee_head_gather_fwd(const float* P, const float* bias, float* out, int B, int H, int W, void* stream);
int dsee_head_scatter_bwd(const float* dout, const float* out, float* dP, int B, int H, int W, void* stream);

/* ---- style encoder / discriminator layers (fp32, NHWC) -------------------------------------- */
/* Replaces the nn.Conv2d calls of encoder.py:84-98,142-157 and discriminator.py:84-96.
 * x fp32 NHWC [B,Hi,Wi,Cin] read through an optional folded 2x nearest upsample (ups=1 replaces
 * the nn.Upsample(scale_factor=2) in front of up_conv / conv2, encoder.py:94-95,153-154);
 * w fp32 [KH][KW][Cin][Cout]; bias fp32 [Cout] or NULL; out fp32 NHWC [B,Ho,Wo,Cout],
 * Ho = ((Hi<<ups) + 2*pad - KH)/stride + 1. lrelu=1 fuses LeakyReLU(0.2) (discriminator.py:85). */
int dsee_conv2d_direct_fwd(const float* x, const float* w, const float* bias, float* out, int B,
                           int Hi, int Wi, int Cin, int Cout, int KH, int KW, int stride, int pad,
                           int ups, int lrelu, void* stream);
/* Replaces nn.InstanceNorm2d(affine=False) (normalization.py:48) + the following activation
 * (act: 0 none, 1 LeakyReLU(0.2), 2 tanh; encoder.py:25-26,86). x, out fp32 NHWC [B,HW,C];
 * mean, rstd fp32 [B,C] are kept for the backward pass; workspace:
 * dsee_instance_norm_workspace_bytes() bytes of per-chunk double partials (deterministic). */
int64_t dsee_instance_norm_workspace_bytes(int B, int HW, int C);
int dsee_instance_norm_fwd(const float* x, float* out, float* mean, float* rstd, void* workspace,
                           int B, int HW, int C, float eps, int act, void* stream);
/* Replaces AbtractStyleEncoder.extract_style_matrix (encoder.py:36-49):
 * style[b,l,c] = sum_{p : labels[b,p]==l} x[b,p,c] / HW.  workspace fp32
 * [B][dsee_region_pool_chunks(HW)][L][C]. */
int dsee_region_pool_chunks(int HW);
int dsee_region_pool_fwd(const float* x, const uint8_t* labels, float* style, float* workspace,
                         int B, int HW, int C, int L, void* stream);
/* fp32 NCHW [B,C,H,W] -> NHWC [B,H,W,Cp], channels C..Cp-1 zero. */
int dsee_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, int Cp, void* stream);
/* Replaces SRModel.discriminate's two torch.cat calls (sr_model.py:655-664):
 * out NHWC [2B,H,W,Cp] = [[onehot(labels) | fake] ; [onehot(labels) | real]], Cp >= L+3. */
int dsee_disc_input(const uint8_t* labels, const float* fake, const float* real, float* out, int B,
                    int L, int H, int W, int Cp, void* stream);
/* Replaces MultiscaleDiscriminator.downsample (discriminator.py:46-49), NHWC. */
int dsee_avgpool3s2_fwd(const float* in, float* out, int B, int Hi, int Wi, int C, void* stream);

/* ---- style encoder / discriminator backward (fp32, NHWC) ------------------------------------- */
/* dx = dy * act'(.) evaluated from the layer output: act 1 LeakyReLU(0.2), 2 tanh.  Used for the
 * LeakyReLU fused into dsee_conv2d_direct_fwd (discriminator.py:84-85). */
int dsee_act_bwd(const float* dy, const float* out, float* dx, int64_t n, int act, void* stream);
/* Backward-data of dsee_conv2d_direct_fwd (autograd of the nn.Conv2d calls at encoder.py:84-98,
 * 142-157 and discriminator.py:84-96): dx fp32 NHWC [B,Hi,Wi,Cin] at the pre-upsample resolution
 * (the folded 2x upsample is transposed into a 2x2 sum). Cin % 4 == 0. */
int dsee_conv2d_direct_dgrad(const float* dy, const float* w, float* dx, int B, int Hi, int Wi,
                             int Cin, int Cout, int KH, int KW, int stride, int pad, int ups,
                             void* stream);
/* Weight gradient, dw fp32 [KH][KW][Cin][Cout]; deterministic split over pixels through
 * workspace fp32 [dsee_conv2d_direct_wgrad_workspace_floats()]. */
int64_t dsee_conv2d_direct_wgrad_workspace_floats(int B, int Ho, int Wo, int Cin, int Cout, int KH,
                                                  int KW);
int dsee_conv2d_direct_wgrad(const float* x, const float* dy, float* dw, float* workspace, int B,
                             int Hi, int Wi, int Cin, int Cout, int KH, int KW, int stride, int pad,
                             int ups, void* stream);
/* out[c] = sum_p x[p][c] (conv bias gradient). workspace fp32 [dsee_channel_sum_chunks()][C]. */
int dsee_channel_sum_chunks(int64_t npix);
int dsee_channel_sum(const float* x, int64_t npix, int C, float* workspace, float* out, void* stream);
/* Backward of dsee_instance_norm_fwd: x = the forward INPUT, mean / rstd from the forward;
 * sums scratch fp32 [B][C][2]; workspace as for the forward. */
int dsee_instance_norm_bwd(const float* x, const float* dout, const float* mean, const float* rstd,
                           float* dx, float* sums, void* workspace, int B, int HW, int C, int act,
                           void* stream);
/* Backward of dsee_region_pool_fwd: dx[b,p,c] = dstyle[b, labels[b,p], c] / HW. */
int dsee_region_pool_bwd(const float* dstyle, const uint8_t* labels, float* dx, int B, int HW, int C,
                         int L, void* stream);
/* Backward of dsee_avgpool3s2_fwd. din fp32 NHWC [B,Hi,Wi,C]. */
int dsee_avgpool3s2_bwd(const float* dout, float* din, int B, int Hi, int Wi, int C, void* stream);
/* Backward of dsee_disc_input wrt the fake image: dfake NCHW [B,3,H,W] from dx NHWC [2B,H,W,Cp]. */
int dsee_disc_input_bwd(const float* dx, float* dfake, int B, int L, int H, int W, int Cp,
                        void* stream);

/* ---- parameter-side fusions ---------------------------------------------------------------------- */
/* Replaces torch.nn.utils.spectral_norm's compute_weight (torch/nn/utils/spectral_norm.py:92-113,
 * applied at architecture.py:40-44 and normalization.py:29-31) for a weight viewed as [N][K]:
 *   power_iteration (training): v <- normalize(W^T u), u <- normalize(W v)   (in place, eps inside max)
 *   sigma = u . (W v);  w_eff = W / sigma;  sigma2 = {sigma, 1/sigma} (device float[2])
 * workspace fp32 [dsee_spectral_workspace_floats(N,K)].  The backward treats u, v as constants like
 * torch does:  dW = (dW_eff - <dW_eff, W_eff> u v^T) / sigma   (workspace: 128 doubles). */
int64_t dsee_spectral_workspace_floats(int N, int K);
int dsee_spectral_weight_fwd(const float* w_orig, float* u, float* v, int N, int K, int power_iteration,
                             float eps, float* workspace, float* sigma2, float* w_eff, void* stream);
int dsee_spectral_weight_bwd(const float* dw_eff, const float* w_eff, const float* u, const float* v,
                             const float* sigma2, int N, int K, void* workspace, float* dw_orig,
                             void* stream);

/* The forward of dsee_spectral_weight_fwd for many layers at once (one launch per step for up to
 * DSEE_SN_MAX_BATCH layers instead of five launches per layer; more layers run in chunks).  items is
 * a HOST array; every pointer in it is a device pointer.  workspace: dsee_spectral_workspace_floats(N, K)
 * floats per item; u_saved / v_saved (optional, may be NULL): copies of this forward's u / v for
 * dsee_spectral_weight_bwd.  Arithmetic identical to the per-layer entry point. */
#define DSEE_SN_MAX_BATCH 24
typedef struct {
    const float* w_orig;
    float* u;
    float* v;
    float* w_eff;
    float* sigma2;
    float* workspace;
    float* u_saved;
    float* v_saved;
    int N, K;
} dsee_sn_item;
int dsee_spectral_weight_fwd_batched(const dsee_sn_item* items, int count, int power_iteration, float eps,
                                     void* stream);
/* Assembly of K1's fused modulation weight from the reference's separate convs
 * (normalization.py:116-119 SPADE, :198-213 SEAN with the sigmoid(alpha) blend, :283-286 PureSEAN):
 * rows interleaved per 128 channels [gamma | beta], columns [seg source (c1) | style source (c2)],
 *   Wm[gamma c] = [(1-a_g) Wg[c] | a_g Wsg[c]],  gamma_bias = (1-a_g) bg + a_g bsg (+1 if plus_one)
 * with a = sigmoid(alpha) for two sources, 0 for the seg source only, 1 for the style source only.
 * index 0 = gamma, 1 = beta.  The backward scatters dWm / d gamma_bias / d beta_bias to the sources
 * and reduces d alpha deterministically. */
typedef struct {
    const float* w_seg[2];   /* mlp_gamma / mlp_beta weights [C][c1][3][3] or NULL */
    const float* w_sty[2];   /* mlp_style_gamma / mlp_style_beta weights [C][c2][3][3] or NULL */
    const float* b_seg[2];
    const float* b_sty[2];
    const float* alpha[2];   /* alpha_gamma / alpha_beta device scalars (two sources only) */
    int C, c1, c2;
    int plus_one;
} dsee_modweight_args;
typedef struct {
    float* dw_seg[2];
    float* dw_sty[2];
    float* db_seg[2];
    float* db_sty[2];
    float* dalpha[2];
} dsee_modweight_grads;
int dsee_modweight_fwd(const dsee_modweight_args* args, float* wm, float* gamma_bias, float* beta_bias,
                       void* stream);
int64_t dsee_modweight_bwd_workspace_bytes(int C, int c1, int c2);
int dsee_modweight_bwd(const dsee_modweight_args* args, const float* dwm, const float* dgamma_bias,
                       const float* dbeta_bias, const dsee_modweight_grads* grads, void* workspace,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPSEE_B200_H */
