// Style-encoder / discriminator kernels: generic direct convolution (fp32, NHWC), instance norm,
// region-wise masked mean pooling, avg-pool pyramid and the fused discriminator input assembly.
// These layers are 0.3-1.5 % of the step's FLOPs and HBM / latency bound.
//
// Reference: encoder.py:36-49,84-98,142-157 ; discriminator.py:46-49,84-100 ;
//            normalization.py:19-54 (spectral conv + InstanceNorm2d(affine=False)).
#include "common.cuh"
#include "launch_count.h"
#include "../../include/deepsee_b200.h"

namespace dsee {

static inline int cdiv2(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// direct convolution: thread = 4 consecutive output pixels (along x) x 4 consecutive out channels
//   x   fp32 NHWC [B, Hi, Wi, Cin]  (read through an optional folded 2x nearest upsample)
//   w   fp32 [KH][KW][Cin][Cout]    (transposed from PyTorch's [Cout][Cin][KH][KW] by the caller)
//   out fp32 NHWC [B, Ho, Wo, Cout]
// ------------------------------------------------------------------------------------------------
constexpr int DC_PX = 4;

__global__ void __launch_bounds__(128)
direct_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                   const float* __restrict__ bias, float* __restrict__ out, int B, int Hi, int Wi,
                   int Cin, int Ho, int Wo, int Cout, int KH, int KW, int stride, int pad, int ups,
                   float lrelu_slope, int apply_act) {
    const int cq = Cout >> 2;
    const int wg = (Wo + DC_PX - 1) / DC_PX;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Ho * wg * cq) return;
    const int q = (int)(i % cq);
    int64_t r = i / cq;
    const int xg = (int)(r % wg);
    r /= wg;
    const int yo = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const int Hu = Hi << ups, Wu = Wi << ups;  // logical (upsampled) input size

    float4 acc[DC_PX];
    float4 bv = bias ? __ldg(reinterpret_cast<const float4*>(bias) + q) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int p = 0; p < DC_PX; ++p) acc[p] = bv;

    for (int ky = 0; ky < KH; ++ky) {
        const int yi = yo * stride - pad + ky;
        if (yi < 0 || yi >= Hu) continue;
        const float* xrow = x + ((size_t)b * Hi + (yi >> ups)) * Wi * Cin;
        for (int kx = 0; kx < KW; ++kx) {
            const float* wt = w + (size_t)(ky * KW + kx) * Cin * Cout + q * 4;
            const float* xp[DC_PX];
            bool ok[DC_PX];
#pragma unroll
            for (int p = 0; p < DC_PX; ++p) {
                const int xo = xg * DC_PX + p;
                const int xi = xo * stride - pad + kx;
                ok[p] = (xo < Wo) && xi >= 0 && xi < Wu;
                xp[p] = xrow + (size_t)(ok[p] ? (xi >> ups) : 0) * Cin;
            }
            int c = 0;
            if ((Cin & 3) == 0) {
                for (; c < Cin; c += 4) {
                    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)c * Cout));
                    const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(c + 1) * Cout));
                    const float4 w2 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(c + 2) * Cout));
                    const float4 w3 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(c + 3) * Cout));
#pragma unroll
                    for (int p = 0; p < DC_PX; ++p) {
                        if (!ok[p]) continue;
                        const float4 v = __ldg(reinterpret_cast<const float4*>(xp[p] + c));
                        acc[p].x += v.x * w0.x + v.y * w1.x + v.z * w2.x + v.w * w3.x;
                        acc[p].y += v.x * w0.y + v.y * w1.y + v.z * w2.y + v.w * w3.y;
                        acc[p].z += v.x * w0.z + v.y * w1.z + v.z * w2.z + v.w * w3.z;
                        acc[p].w += v.x * w0.w + v.y * w1.w + v.z * w2.w + v.w * w3.w;
                    }
                }
            }
            for (; c < Cin; ++c) {
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)c * Cout));
#pragma unroll
                for (int p = 0; p < DC_PX; ++p) {
                    if (!ok[p]) continue;
                    const float v = __ldg(xp[p] + c);
                    acc[p].x += v * w0.x;
                    acc[p].y += v * w0.y;
                    acc[p].z += v * w0.z;
                    acc[p].w += v * w0.w;
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < DC_PX; ++p) {
        const int xo = xg * DC_PX + p;
        if (xo >= Wo) continue;
        float4 a = acc[p];
        if (apply_act) {
            a.x = a.x > 0.f ? a.x : lrelu_slope * a.x;
            a.y = a.y > 0.f ? a.y : lrelu_slope * a.y;
            a.z = a.z > 0.f ? a.z : lrelu_slope * a.z;
            a.w = a.w > 0.f ? a.w : lrelu_slope * a.w;
        }
        *reinterpret_cast<float4*>(out + (((size_t)b * Ho + yo) * Wo + xo) * Cout + q * 4) = a;
    }
}

// Cout not a multiple of 4 (the discriminator's 1-channel prediction conv): thread = 1 output.
__global__ void direct_conv_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                         const float* __restrict__ bias, float* __restrict__ out,
                                         int B, int Hi, int Wi, int Cin, int Ho, int Wo, int Cout,
                                         int KH, int KW, int stride, int pad) {
    // warp per output pixel; lanes split Cin; all Cout (<= 4) accumulated per lane
    const int lane = threadIdx.x & 31;
    int64_t pix = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (pix >= (int64_t)B * Ho * Wo) return;
    const int xo = (int)(pix % Wo);
    const int yo = (int)((pix / Wo) % Ho);
    const int b = (int)(pix / ((int64_t)Wo * Ho));
    float acc[4] = {0, 0, 0, 0};
    for (int ky = 0; ky < KH; ++ky) {
        const int yi = yo * stride - pad + ky;
        if (yi < 0 || yi >= Hi) continue;
        for (int kx = 0; kx < KW; ++kx) {
            const int xi = xo * stride - pad + kx;
            if (xi < 0 || xi >= Wi) continue;
            const float* xp = x + (((size_t)b * Hi + yi) * Wi + xi) * Cin;
            const float* wt = w + (size_t)(ky * KW + kx) * Cin * Cout;
            for (int c = lane; c < Cin; c += 32) {
                const float v = __ldg(xp + c);
                for (int o = 0; o < Cout; ++o) acc[o] += v * __ldg(wt + (size_t)c * Cout + o);
            }
        }
    }
    for (int o = 0; o < Cout; ++o) {
        float a = acc[o];
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
        if (lane == 0) out[(size_t)pix * Cout + o] = a + (bias ? bias[o] : 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// instance norm (affine=False, eps inside sqrt, biased variance) + activation
// ------------------------------------------------------------------------------------------------
// grid (C/32, B), block 256 = 32 channels x 8 pixel lanes; fixed-order reduction.
__global__ void instnorm_stats_kernel(const float* __restrict__ x, int HW, int C, float eps,
                                      float* __restrict__ mean, float* __restrict__ rstd) {
    __shared__ double sh[8][32][2];
    const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl, b = blockIdx.y;
    double a = 0.0, q = 0.0;
    if (c < C) {
        const float* p = x + (size_t)b * HW * C + c;
        for (int i = g; i < HW; i += 8) {
            float v = __ldg(p + (size_t)i * C);
            a += v;
            q += (double)v * v;
        }
    }
    sh[g][cl][0] = a;
    sh[g][cl][1] = q;
    __syncthreads();
    if (g == 0 && c < C) {
        double A = 0, Q = 0;
        for (int k = 0; k < 8; ++k) {
            A += sh[k][cl][0];
            Q += sh[k][cl][1];
        }
        double m = A / HW;
        double var = Q / HW - m * m;
        if (var < 0) var = 0;
        mean[(size_t)b * C + c] = (float)m;
        rstd[(size_t)b * C + c] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// act: 0 none, 1 leaky relu (slope), 2 tanh
__global__ void instnorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                      const float* __restrict__ rstd, float* __restrict__ out,
                                      int64_t n4, int HW, int C, int act, float slope) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const int cq = C >> 2;
    const int q = (int)(i % cq);
    const int b = (int)(i / ((int64_t)cq * HW));
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean + (size_t)b * C) + q);
    const float4 r = __ldg(reinterpret_cast<const float4*>(rstd + (size_t)b * C) + q);
    float a[4] = {(v.x - m.x) * r.x, (v.y - m.y) * r.y, (v.z - m.z) * r.z, (v.w - m.w) * r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (act == 1) a[e] = a[e] > 0.f ? a[e] : slope * a[e];
        else if (act == 2) a[e] = tanhf(a[e]);
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(a[0], a[1], a[2], a[3]);
}

// ------------------------------------------------------------------------------------------------
// region-wise masked mean pooling (encoder.py:36-49): style[b,l,c] = sum_{p: label(p)=l} x[b,p,c] / HW
// grid (chunks, B); block = C threads (thread owns a channel column -> no smem conflicts).
// ------------------------------------------------------------------------------------------------
constexpr int POOL_CHUNK = 1024;
__global__ void region_pool_partial_kernel(const float* __restrict__ x, int ld, int coff,
                                           const uint8_t* __restrict__ labels, int HW, int C, int L,
                                           float* __restrict__ partial) {
    extern __shared__ float acc[];  // [L][C]
    const int c = threadIdx.x, b = blockIdx.y;
    for (int l = 0; l < L; ++l) acc[l * C + c] = 0.f;
    const int p0 = blockIdx.x * POOL_CHUNK;
    const int p1 = min(p0 + POOL_CHUNK, HW);
    const uint8_t* lb = labels + (size_t)b * HW;
    const float* xp = x + (size_t)b * HW * ld + coff + c;
    for (int p = p0; p < p1; ++p) {
        int l = lb[p];
        if (l < L) acc[l * C + c] += __ldg(xp + (size_t)p * ld);
    }
    float* out = partial + ((size_t)(b * gridDim.x + blockIdx.x) * L) * C + c;
    for (int l = 0; l < L; ++l) out[(size_t)l * C] = acc[l * C + c];
}

__global__ void region_pool_final_kernel(const float* __restrict__ partial, int chunks, int LC,
                                         float inv_hw, float* __restrict__ style) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= LC) return;
    double a = 0.0;
    for (int k = 0; k < chunks; ++k) a += partial[((size_t)(b * chunks + k)) * LC + i];
    style[(size_t)b * LC + i] = (float)(a * inv_hw);
}

// ------------------------------------------------------------------------------------------------
// layout helpers / discriminator input
// ------------------------------------------------------------------------------------------------
// fp32 NCHW [B,C,H,W] -> NHWC [B,H,W,Cp] (channels >= C zero-filled).
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int B,
                                    int C, int HW, int Cp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HW * Cp) return;
    const int c = (int)(i % Cp);
    const int64_t bp = i / Cp;
    const int p = (int)(bp % HW);
    const int b = (int)(bp / HW);
    out[i] = c < C ? in[((size_t)b * C + c) * HW + p] : 0.f;
}

// cat([onehot(labels), image], 1) for the fake and the real image, stacked on the batch dim
// (sr_model.py:655-664), written NHWC with the channel count padded to Cp (zeros):
//   out[b]     = [onehot(labels[b]) | fake[b]],   out[B + b] = [onehot(labels[b]) | real[b]]
__global__ void disc_input_kernel(const uint8_t* __restrict__ labels, const float* __restrict__ fake,
                                  const float* __restrict__ real, float* __restrict__ out, int B,
                                  int L, int HW, int Cp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)2 * B * HW * Cp) return;
    const int c = (int)(i % Cp);
    const int64_t bp = i / Cp;
    const int p = (int)(bp % HW);
    const int b2 = (int)(bp / HW);
    const int b = b2 % B;
    float v = 0.f;
    if (c < L) {
        v = (labels[(size_t)b * HW + p] == c) ? 1.f : 0.f;
    } else if (c < L + 3) {
        const float* img = b2 < B ? fake : real;
        v = img[((size_t)b * 3 + (c - L)) * HW + p];
    }
    out[i] = v;
}

// F.avg_pool2d(kernel 3, stride 2, padding 1, count_include_pad=False), NHWC (discriminator.py:46-49)
__global__ void avgpool3s2_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int Hi,
                                  int Wi, int C, int Ho, int Wo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Ho * Wo * C) return;
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int xo = (int)(r % Wo);
    r /= Wo;
    const int yo = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float s = 0.f;
    int cnt = 0;
    for (int ky = 0; ky < 3; ++ky) {
        const int yi = yo * 2 - 1 + ky;
        if (yi < 0 || yi >= Hi) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int xi = xo * 2 - 1 + kx;
            if (xi < 0 || xi >= Wi) continue;
            s += in[(((size_t)b * Hi + yi) * Wi + xi) * C + c];
            ++cnt;
        }
    }
    out[i] = s / (float)cnt;
}

}  // namespace dsee

using namespace dsee;

#define LAUNCH_END()               \
    count_launch();                \
    DSEE_CUDA(cudaGetLastError()); \
    return 0

extern "C" int dsee_conv2d_direct_fwd(const float* x, const float* w, const float* bias, float* out,
                                      int B, int Hi, int Wi, int Cin, int Cout, int KH, int KW,
                                      int stride, int pad, int ups, int lrelu, void* stream) {
    DSEE_CHECK_ARG(x && w && out && B > 0 && Hi > 0 && Wi > 0 && Cin > 0 && Cout > 0, "bad argument");
    DSEE_CHECK_ARG(stride >= 1 && (ups == 0 || ups == 1), "bad stride/ups");
    int rc = require_sm100();
    if (rc) return rc;
    const int Hu = Hi << ups, Wu = Wi << ups;
    const int Ho = (Hu + 2 * pad - KH) / stride + 1, Wo = (Wu + 2 * pad - KW) / stride + 1;
    DSEE_CHECK_ARG(Ho > 0 && Wo > 0, "empty output");
    if (Cout % 4 == 0) {
        int64_t n = (int64_t)B * Ho * ((Wo + DC_PX - 1) / DC_PX) * (Cout / 4);
        direct_conv_kernel<<<cdiv2(n, 128), 128, 0, (cudaStream_t)stream>>>(
            x, w, bias, out, B, Hi, Wi, Cin, Ho, Wo, Cout, KH, KW, stride, pad, ups, 0.2f, lrelu);
    } else {
        DSEE_CHECK_ARG(Cout <= 4 && ups == 0 && !lrelu, "Cout %% 4 != 0 only supported for Cout <= 4");
        int64_t n = (int64_t)B * Ho * Wo * 32;
        direct_conv_small_kernel<<<cdiv2(n, 256), 256, 0, (cudaStream_t)stream>>>(
            x, w, bias, out, B, Hi, Wi, Cin, Ho, Wo, Cout, KH, KW, stride, pad);
    }
    LAUNCH_END();
}

extern "C" int dsee_instance_norm_fwd(const float* x, float* out, float* mean, float* rstd, int B,
                                      int HW, int C, float eps, int act, void* stream) {
    DSEE_CHECK_ARG(x && out && mean && rstd && B > 0 && HW > 0 && C > 0 && C % 4 == 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    instnorm_stats_kernel<<<dim3(cdiv2(C, 32), B), 256, 0, st>>>(x, HW, C, eps, mean, rstd);
    count_launch();
    int64_t n4 = (int64_t)B * HW * C / 4;
    instnorm_apply_kernel<<<cdiv2(n4, 256), 256, 0, st>>>(x, mean, rstd, out, n4, HW, C, act, 0.2f);
    LAUNCH_END();
}

extern "C" int dsee_region_pool_chunks(int HW) { return (HW + POOL_CHUNK - 1) / POOL_CHUNK; }

extern "C" int dsee_region_pool_fwd(const float* x, const uint8_t* labels, float* style,
                                    float* workspace, int B, int HW, int C, int L, void* stream) {
    DSEE_CHECK_ARG(x && labels && style && workspace && B > 0 && HW > 0 && C > 0 && C <= 1024 && L > 0,
                   "bad argument");
    DSEE_CHECK_ARG((size_t)L * C * 4 <= 48 * 1024, "L*C too large for the shared-memory accumulator");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = dsee_region_pool_chunks(HW);
    region_pool_partial_kernel<<<dim3(chunks, B), C, (size_t)L * C * 4, st>>>(x, C, 0, labels, HW, C, L,
                                                                              workspace);
    count_launch();
    region_pool_final_kernel<<<dim3(cdiv2(L * C, 256), B), 256, 0, st>>>(workspace, chunks, L * C,
                                                                         1.0f / (float)HW, style);
    LAUNCH_END();
}

extern "C" int dsee_style_gather_bwd(const float* dsrc, int ld, int coff, const uint8_t* labels,
                                     float* dstyle, float* workspace, int B, int HW, int L, int d,
                                     void* stream) {
    DSEE_CHECK_ARG(dsrc && labels && dstyle && workspace && B > 0 && HW > 0 && d > 0 && d <= 1024 &&
                       L > 0 && coff >= 0 && coff + d <= ld,
                   "bad argument");
    DSEE_CHECK_ARG((size_t)L * d * 4 <= 48 * 1024, "L*d too large for the shared-memory accumulator");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = dsee_region_pool_chunks(HW);
    region_pool_partial_kernel<<<dim3(chunks, B), d, (size_t)L * d * 4, st>>>(dsrc, ld, coff, labels, HW,
                                                                              d, L, workspace);
    count_launch();
    region_pool_final_kernel<<<dim3(cdiv2(L * d, 256), B), 256, 0, st>>>(workspace, chunks, L * d, 1.0f,
                                                                         dstyle);
    LAUNCH_END();
}

extern "C" int dsee_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, int Cp,
                                 void* stream) {
    DSEE_CHECK_ARG(in && out && B > 0 && C > 0 && H > 0 && W > 0 && Cp >= C, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    int64_t n = (int64_t)B * H * W * Cp;
    nchw_to_nhwc_kernel<<<cdiv2(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, C, H * W, Cp);
    LAUNCH_END();
}

extern "C" int dsee_disc_input(const uint8_t* labels, const float* fake, const float* real,
                               float* out, int B, int L, int H, int W, int Cp, void* stream) {
    DSEE_CHECK_ARG(labels && fake && real && out && B > 0 && L > 0 && Cp >= L + 3, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    int64_t n = (int64_t)2 * B * H * W * Cp;
    disc_input_kernel<<<cdiv2(n, 256), 256, 0, (cudaStream_t)stream>>>(labels, fake, real, out, B, L,
                                                                       H * W, Cp);
    LAUNCH_END();
}

extern "C" int dsee_avgpool3s2_fwd(const float* in, float* out, int B, int Hi, int Wi, int C,
                                   void* stream) {
    DSEE_CHECK_ARG(in && out && B > 0 && Hi > 0 && Wi > 0 && C > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    const int Ho = (Hi + 2 - 3) / 2 + 1, Wo = (Wi + 2 - 3) / 2 + 1;
    int64_t n = (int64_t)B * Ho * Wo * C;
    avgpool3s2_kernel<<<cdiv2(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, Hi, Wi, C, Ho, Wo);
    LAUNCH_END();
}
