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
// grid (C/32, B, chunks), block 256 = 32 channels x 8 pixel lanes: per-chunk double partials
// (sum, sum of squares), then a fixed-order finalize -> deterministic, and enough blocks to fill the
// GPU even for the 32-channel layers.
constexpr int IN_CHUNK = 2048;  // pixels per chunk
__global__ void instnorm_stats_kernel(const float* __restrict__ x, int HW, int C,
                                      double* __restrict__ partial) {
    // block 256 = 8 channel quads (one 32-channel slab, float4 loads) x 32 pixel lanes
    __shared__ float sh[32][8][8];
    const int q = threadIdx.x & 7, g = threadIdx.x >> 3;
    const int c = blockIdx.x * 32 + q * 4, b = blockIdx.y;
    const int p0 = blockIdx.z * IN_CHUNK, p1 = min(p0 + IN_CHUNK, HW);
    float a[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
        const float* p = x + (size_t)b * HW * C + c;
#pragma unroll 4
        for (int i = p0 + g; i < p1; i += 32) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p + (size_t)i * C));
            a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
            s2[0] += v.x * v.x; s2[1] += v.y * v.y; s2[2] += v.z * v.z; s2[3] += v.w * v.w;
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        sh[g][q][e] = a[e];
        sh[g][q][4 + e] = s2[e];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        // thread = (channel of the slab, sum | sum of squares); fixed-order double accumulation
        const int cl = threadIdx.x & 31, k = threadIdx.x >> 5;
        if (blockIdx.x * 32 + cl < C) {
            double A = 0;
            for (int l = 0; l < 32; ++l) A += (double)sh[l][cl >> 2][k * 4 + (cl & 3)];
            partial[(((size_t)b * gridDim.z + blockIdx.z) * C + blockIdx.x * 32 + cl) * 2 + k] = A;
        }
    }
}
__global__ void instnorm_finalize_kernel(const double* __restrict__ partial, int chunks, int HW, int C,
                                         int BC, float eps, float* __restrict__ mean,
                                         float* __restrict__ rstd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BC) return;
    const int b = i / C, c = i % C;
    double A = 0, Q = 0;
    for (int k = 0; k < chunks; ++k) {
        const double* o = partial + (((size_t)b * chunks + k) * C + c) * 2;
        A += o[0];
        Q += o[1];
    }
    const double m = A / HW;
    double var = Q / HW - m * m;
    if (var < 0) var = 0;
    mean[i] = (float)m;
    rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

// act: 0 none, 1 leaky relu (slope), 2 tanh
__global__ void instnorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                      const float* __restrict__ rstd, float* __restrict__ out,
                                      int64_t n4, int HW, int C, int act, float slope) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const uint32_t cq = (uint32_t)C >> 2;
    const int q = (int)((uint32_t)i % cq);  // n4 < 2^31 (host wrapper)
    const int b = (int)((uint32_t)i / (cq * (uint32_t)HW));
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
constexpr int POOL_CHUNK = 256;
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
    int p = p0;
    for (; p + 4 <= p1; p += 4) {  // 4 independent loads in flight per thread
        const float v0 = __ldg(xp + (size_t)p * ld), v1 = __ldg(xp + (size_t)(p + 1) * ld);
        const float v2 = __ldg(xp + (size_t)(p + 2) * ld), v3 = __ldg(xp + (size_t)(p + 3) * ld);
        const int l0 = lb[p], l1 = lb[p + 1], l2 = lb[p + 2], l3 = lb[p + 3];
        if (l0 < L) acc[l0 * C + c] += v0;
        if (l1 < L) acc[l1 * C + c] += v1;
        if (l2 < L) acc[l2 * C + c] += v2;
        if (l3 < L) acc[l3 * C + c] += v3;
    }
    for (; p < p1; ++p) {
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
    const uint32_t iu = (uint32_t)i;  // < 2^31 elements (host wrapper): 32-bit div/mod
    const int c = (int)(iu % (uint32_t)Cp);
    const uint32_t bp = iu / (uint32_t)Cp;
    const int p = (int)(bp % (uint32_t)HW);
    const int b2 = (int)(bp / (uint32_t)HW);
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
    const uint32_t iu = (uint32_t)i;  // < 2^31 elements (host wrapper): 32-bit div/mod
    const int c = (int)(iu % (uint32_t)C);
    uint32_t r = iu / (uint32_t)C;
    const int xo = (int)(r % (uint32_t)Wo);
    r /= (uint32_t)Wo;
    const int yo = (int)(r % (uint32_t)Ho);
    const int b = (int)(r / (uint32_t)Ho);
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
            x, w, bias, out, B, Hi, Wi, Cin, Ho, Wo, Cout, KH, KW, stride, pad, ups, lrelu == 2 ? 0.f : 0.2f,
            lrelu);
    } else {
        DSEE_CHECK_ARG(Cout <= 4 && ups == 0 && !lrelu, "Cout %% 4 != 0 only supported for Cout <= 4");
        int64_t n = (int64_t)B * Ho * Wo * 32;
        direct_conv_small_kernel<<<cdiv2(n, 256), 256, 0, (cudaStream_t)stream>>>(
            x, w, bias, out, B, Hi, Wi, Cin, Ho, Wo, Cout, KH, KW, stride, pad);
    }
    LAUNCH_END();
}

extern "C" int64_t dsee_instance_norm_workspace_bytes(int B, int HW, int C) {
    return (int64_t)B * cdiv2(HW, IN_CHUNK) * C * 2 * (int64_t)sizeof(double);
}

extern "C" int dsee_instance_norm_fwd(const float* x, float* out, float* mean, float* rstd,
                                      void* workspace, int B, int HW, int C, float eps, int act,
                                      void* stream) {
    DSEE_CHECK_ARG(x && out && mean && rstd && workspace && B > 0 && HW > 0 && C > 0 && C % 4 == 0,
                   "bad argument");
    DSEE_CHECK_ARG((int64_t)B * HW * C < ((int64_t)1 << 31), "more than 2^31 elements");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = cdiv2(HW, IN_CHUNK);
    instnorm_stats_kernel<<<dim3(cdiv2(C, 32), B, chunks), 256, 0, st>>>(x, HW, C, (double*)workspace);
    count_launch();
    instnorm_finalize_kernel<<<cdiv2(B * C, 128), 128, 0, st>>>((const double*)workspace, chunks, HW, C,
                                                                B * C, eps, mean, rstd);
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
    DSEE_CHECK_ARG((int64_t)2 * B * H * W * Cp < ((int64_t)1 << 31), "more than 2^31 elements");
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
    DSEE_CHECK_ARG((int64_t)B * Hi * Wi * C < ((int64_t)1 << 31), "more than 2^31 elements");
    int rc = require_sm100();
    if (rc) return rc;
    const int Ho = (Hi + 2 - 3) / 2 + 1, Wo = (Wi + 2 - 3) / 2 + 1;
    int64_t n = (int64_t)B * Ho * Wo * C;
    avgpool3s2_kernel<<<cdiv2(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, Hi, Wi, C, Ho, Wo);
    LAUNCH_END();
}

// =================================================================================================
// backward of the style-encoder / discriminator layers (fp32, NHWC; deterministic reductions)
// =================================================================================================
namespace dsee {

// dx = dy * act'(.) from the layer OUTPUT: act 1 LeakyReLU / 3 ReLU (sign of out), 2 tanh (1 - out^2)
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ out,
                               float* __restrict__ dx, int64_t n, int act, float slope) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float o = out[i];
    dx[i] = dy[i] * (act != 2 ? (o > 0.f ? 1.f : slope) : (1.f - o * o));
}

// backward-data of direct_conv_kernel: thread = one (pre-upsample) input pixel x 4 input channels
//   dx[b,yi,xi,ci] = sum over the 2^ups x 2^ups upsampled copies (yu,xu), taps (ky,kx) with
//                    yo*stride - pad + ky == yu, and co:  dy[b,yo,xo,co] * w[ky][kx][ci][co]
__global__ void __launch_bounds__(128)
direct_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx,
                    int B, int Hi, int Wi, int Cin, int Ho, int Wo, int Cout, int KH, int KW,
                    int stride, int pad, int ups) {
    const int cq = Cin >> 2;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Hi * Wi * cq) return;
    const int q = (int)(i % cq);
    int64_t r = i / cq;
    const int xi = (int)(r % Wi);
    r /= Wi;
    const int yi = (int)(r % Hi);
    const int b = (int)(r / Hi);
    const int f = 1 << ups;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sy = 0; sy < f; ++sy)
        for (int sx = 0; sx < f; ++sx) {
            const int yu = yi * f + sy, xu = xi * f + sx;
            for (int ky = 0; ky < KH; ++ky) {
                const int ty = yu + pad - ky;
                if (ty < 0 || ty % stride != 0) continue;
                const int yo = ty / stride;
                if (yo >= Ho) continue;
                for (int kx = 0; kx < KW; ++kx) {
                    const int tx = xu + pad - kx;
                    if (tx < 0 || tx % stride != 0) continue;
                    const int xo = tx / stride;
                    if (xo >= Wo) continue;
                    const float* dp = dy + (((size_t)b * Ho + yo) * Wo + xo) * Cout;
                    const float* wt = w + ((size_t)(ky * KW + kx) * Cin + q * 4) * Cout;
                    int co = 0;
                    if ((Cout & 3) == 0) {
                        for (; co < Cout; co += 4) {
                            const float4 d = __ldg(reinterpret_cast<const float4*>(dp + co));
                            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + co));
                            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + Cout + co));
                            const float4 w2 = __ldg(reinterpret_cast<const float4*>(wt + 2 * Cout + co));
                            const float4 w3 = __ldg(reinterpret_cast<const float4*>(wt + 3 * Cout + co));
                            acc.x += d.x * w0.x + d.y * w0.y + d.z * w0.z + d.w * w0.w;
                            acc.y += d.x * w1.x + d.y * w1.y + d.z * w1.z + d.w * w1.w;
                            acc.z += d.x * w2.x + d.y * w2.y + d.z * w2.z + d.w * w2.w;
                            acc.w += d.x * w3.x + d.y * w3.y + d.z * w3.z + d.w * w3.w;
                        }
                    }
                    for (; co < Cout; ++co) {
                        const float d = __ldg(dp + co);
                        acc.x += d * __ldg(wt + co);
                        acc.y += d * __ldg(wt + Cout + co);
                        acc.z += d * __ldg(wt + 2 * Cout + co);
                        acc.w += d * __ldg(wt + 3 * Cout + co);
                    }
                }
            }
        }
    *reinterpret_cast<float4*>(dx + (((size_t)b * Hi + yi) * Wi + xi) * Cin + q * 4) = acc;
}

// weight gradient of direct_conv_kernel as a pixel-reduction GEMM per filter tap:
//   dw[tap][ci][co] = sum_p x[p @ tap][ci] * dy[p][co]
// block = 256 threads -> a 64(ci) x 64(co) tile, 4x4 per thread, 16 pixels per smem stage; the
// pixel range is split over `splits` blocks whose partial tiles a fixed-order kernel reduces.
constexpr int DW_T = 64, DW_PK = 16;
__global__ void __launch_bounds__(256)
direct_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                    float* __restrict__ partial, int B, int Hi, int Wi, int Cin, int Ho, int Wo,
                    int Cout, int KH, int KW, int stride, int pad, int ups, int splits, int ci_tiles,
                    int co_tiles) {
    __shared__ float xs[DW_PK][DW_T];
    __shared__ float ds[DW_PK][DW_T];
    int u = blockIdx.x;
    const int cot = u % co_tiles;
    u /= co_tiles;
    const int cit = u % ci_tiles;
    u /= ci_tiles;
    const int tap = u % (KH * KW);
    const int split = u / (KH * KW);
    const int ky = tap / KW, kx = tap % KW;
    const int ci0 = cit * DW_T, co0 = cot * DW_T;
    const int64_t npix = (int64_t)B * Ho * Wo;
    const int64_t pbeg = npix * split / splits, pend = npix * (split + 1) / splits;
    const int Hu = Hi << ups, Wu = Wi << ups;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    for (int64_t p0 = pbeg; p0 < pend; p0 += DW_PK) {
        for (int e = threadIdx.x; e < DW_PK * DW_T; e += 256) {
            const int pk = e / DW_T, c = e % DW_T;
            const int64_t p = p0 + pk;
            float xv = 0.f, dv = 0.f;
            if (p < pend) {
                const int xo = (int)(p % Wo);
                const int yo = (int)((p / Wo) % Ho);
                const int b = (int)(p / ((int64_t)Wo * Ho));
                const int yu = yo * stride - pad + ky, xu = xo * stride - pad + kx;
                if (yu >= 0 && yu < Hu && xu >= 0 && xu < Wu && ci0 + c < Cin)
                    xv = __ldg(x + (((size_t)b * Hi + (yu >> ups)) * Wi + (xu >> ups)) * Cin + ci0 + c);
                if (co0 + c < Cout) dv = __ldg(dy + (size_t)p * Cout + co0 + c);
            }
            xs[pk][c] = xv;
            ds[pk][c] = dv;
        }
        __syncthreads();
#pragma unroll
        for (int pk = 0; pk < DW_PK; ++pk) {
            const float4 a = *reinterpret_cast<const float4*>(&xs[pk][ty * 4]);
            const float4 d = *reinterpret_cast<const float4*>(&ds[pk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, dv4[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] += av[r] * dv4[c];
        }
        __syncthreads();
    }
    float* out = partial + ((size_t)split * KH * KW + tap) * Cin * Cout;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int ci = ci0 + ty * 4 + r;
        if (ci >= Cin) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int co = co0 + tx * 4 + c;
            if (co < Cout) out[(size_t)ci * Cout + co] = acc[r][c];
        }
    }
}

// out[i] = sum_s partial[s][i]  (fixed order)
__global__ void sum_splits_kernel(const float* __restrict__ partial, float* __restrict__ out, int splits,
                                  int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += partial[(size_t)s * n + i];
    out[i] = a;
}

// per-channel sums over pixels (bias gradient): grid (C/32, chunks); partial [chunks][C]
constexpr int CS_PIX = 2048;
__global__ void channel_sum_kernel(const float* __restrict__ x, int64_t npix, int C,
                                   float* __restrict__ partial) {
    __shared__ float sh[8][32];
    const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int64_t p0 = (int64_t)blockIdx.y * CS_PIX;
    const int64_t p1 = p0 + CS_PIX < npix ? p0 + CS_PIX : npix;
    float a = 0.f;
    if (c < C)
        for (int64_t p = p0 + g; p < p1; p += 8) a += __ldg(x + (size_t)p * C + c);
    sh[g][cl] = a;
    __syncthreads();
    if (g == 0 && c < C) {
        float t = 0.f;
        for (int j = 0; j < 8; ++j) t += sh[j][cl];
        partial[(size_t)blockIdx.y * C + c] = t;
    }
}

// instance-norm backward. y = (x - mean) * rstd, out = act(y):
//   g = dout * act'(y);  dx = rstd * (g - mean_p(g) - y * mean_p(g * y))
// stats kernel: grid (C/32, B), block 32 x 8 -> sums[b][c][2] = (sum g, sum g*y)
__device__ __forceinline__ float act_grad(float y, int act, float slope) {
    if (act == 1) return y > 0.f ? 1.f : slope;
    if (act == 2) {
        const float t = tanhf(y);
        return 1.f - t * t;
    }
    return 1.f;
}
__global__ void instnorm_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                          int HW, int C, int act, float slope,
                                          double* __restrict__ partial) {
    // same thread layout as instnorm_stats_kernel
    __shared__ float sh[32][8][8];
    const int q = threadIdx.x & 7, g = threadIdx.x >> 3;
    const int c = blockIdx.x * 32 + q * 4, b = blockIdx.y;
    const int p0 = blockIdx.z * IN_CHUNK, p1 = min(p0 + IN_CHUNK, HW);
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
        const float4 m4 = __ldg(reinterpret_cast<const float4*>(mean + (size_t)b * C + c));
        const float4 r4 = __ldg(reinterpret_cast<const float4*>(rstd + (size_t)b * C + c));
        const float m[4] = {m4.x, m4.y, m4.z, m4.w}, r[4] = {r4.x, r4.y, r4.z, r4.w};
        const size_t base = (size_t)b * HW * C + c;
#pragma unroll 2
        for (int i = p0 + g; i < p1; i += 32) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + base + (size_t)i * C));
            const float4 dv = __ldg(reinterpret_cast<const float4*>(dout + base + (size_t)i * C));
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float y = (xs[e] - m[e]) * r[e];
                const float gg = ds[e] * act_grad(y, act, slope);
                s0[e] += gg;
                s1[e] += gg * y;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        sh[g][q][e] = s0[e];
        sh[g][q][4 + e] = s1[e];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        const int cl = threadIdx.x & 31, k = threadIdx.x >> 5;
        if (blockIdx.x * 32 + cl < C) {
            double A = 0;
            for (int l = 0; l < 32; ++l) A += (double)sh[l][cl >> 2][k * 4 + (cl & 3)];
            partial[(((size_t)b * gridDim.z + blockIdx.z) * C + blockIdx.x * 32 + cl) * 2 + k] = A;
        }
    }
}
__global__ void instnorm_bwd_finalize_kernel(const double* __restrict__ partial, int chunks, int HW,
                                             int C, int BC, float* __restrict__ sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BC) return;
    const int b = i / C, c = i % C;
    double a = 0, q = 0;
    for (int k = 0; k < chunks; ++k) {
        const double* o = partial + (((size_t)b * chunks + k) * C + c) * 2;
        a += o[0];
        q += o[1];
    }
    sums[(size_t)i * 2] = (float)(a / HW);
    sums[(size_t)i * 2 + 1] = (float)(q / HW);
}
__global__ void instnorm_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dout,
                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                          const float* __restrict__ sums, float* __restrict__ dx,
                                          int64_t n, int HW, int C, int act, float slope) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)((uint32_t)i % (uint32_t)C);  // n < 2^31 (host wrapper)
    const int b = (int)((uint32_t)i / ((uint32_t)C * (uint32_t)HW));
    const size_t bc = (size_t)b * C + c;
    const float r = rstd[bc];
    const float y = (x[i] - mean[bc]) * r;
    const float g = dout[i] * act_grad(y, act, slope);
    dx[i] = r * (g - sums[bc * 2] - y * sums[bc * 2 + 1]);
}

// backward of region_pool: dx[b,p,c] = dstyle[b, labels[b,p], c] / HW
__global__ void region_pool_bwd_kernel(const float* __restrict__ dstyle, const uint8_t* __restrict__ labels,
                                       float* __restrict__ dx, int64_t n4, int HW, int C, int L,
                                       float inv_hw) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const int cq = C >> 2;
    const int q = (int)(i % cq);
    const int64_t bp = i / cq;
    const int b = (int)(bp / HW);
    const int l = labels[bp];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < L) {
        v = __ldg(reinterpret_cast<const float4*>(dstyle + ((size_t)b * L + l) * C) + q);
        v.x *= inv_hw; v.y *= inv_hw; v.z *= inv_hw; v.w *= inv_hw;
    }
    reinterpret_cast<float4*>(dx)[i] = v;
}

// nn.MaxPool2d(kernel_size=2, stride=2) of VGG19 (architecture.py:151-181 via torchvision), NHWC;
// the backward pass recomputes the arg-max from the input (first maximum in row-major window order,
// like ATen) instead of storing indices.  thread = (output pixel, 4 channels)
__global__ void maxpool2_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int Hi, int Wi,
                                int C4, int Ho, int Wo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Ho * Wo * C4) return;
    const uint32_t iu = (uint32_t)i;
    const int g = (int)(iu % (uint32_t)C4);
    uint32_t r = iu / (uint32_t)C4;
    const int xo = (int)(r % (uint32_t)Wo);
    r /= (uint32_t)Wo;
    const int yo = (int)(r % (uint32_t)Ho);
    const int b = (int)(r / (uint32_t)Ho);
    const float4* p = reinterpret_cast<const float4*>(in) + (((size_t)b * Hi + 2 * yo) * Wi + 2 * xo) * C4 + g;
    const float4 a = __ldg(p), bq = __ldg(p + C4), c = __ldg(p + (size_t)Wi * C4), d = __ldg(p + (size_t)Wi * C4 + C4);
    float4 m;
    m.x = fmaxf(fmaxf(a.x, bq.x), fmaxf(c.x, d.x));
    m.y = fmaxf(fmaxf(a.y, bq.y), fmaxf(c.y, d.y));
    m.z = fmaxf(fmaxf(a.z, bq.z), fmaxf(c.z, d.z));
    m.w = fmaxf(fmaxf(a.w, bq.w), fmaxf(c.w, d.w));
    reinterpret_cast<float4*>(out)[i] = m;
}

__device__ __forceinline__ void route_max(float a, float b, float c, float d, float g, float& ga, float& gb,
                                          float& gc, float& gd) {
    int k = 0;
    float m = a;
    if (b > m) { m = b; k = 1; }
    if (c > m) { m = c; k = 2; }
    if (d > m) { m = d; k = 3; }
    ga = k == 0 ? g : 0.f;
    gb = k == 1 ? g : 0.f;
    gc = k == 2 ? g : 0.f;
    gd = k == 3 ? g : 0.f;
}

__global__ void maxpool2_bwd_kernel(const float* __restrict__ in, const float* __restrict__ dout,
                                    float* __restrict__ din, int B, int Hi, int Wi, int C4, int Ho, int Wo) {
    // thread = (output pixel, 4 channels): writes the 2x2 window of din (windows do not overlap);
    // rows / columns beyond 2*Ho, 2*Wo (odd sizes) are zeroed by the host wrapper
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Ho * Wo * C4) return;
    const uint32_t iu = (uint32_t)i;
    const int g = (int)(iu % (uint32_t)C4);
    uint32_t r = iu / (uint32_t)C4;
    const int xo = (int)(r % (uint32_t)Wo);
    r /= (uint32_t)Wo;
    const int yo = (int)(r % (uint32_t)Ho);
    const int b = (int)(r / (uint32_t)Ho);
    const size_t o00 = (((size_t)b * Hi + 2 * yo) * Wi + 2 * xo) * C4 + g;
    const float4* p = reinterpret_cast<const float4*>(in) + o00;
    const float4 a = __ldg(p), bq = __ldg(p + C4), c = __ldg(p + (size_t)Wi * C4), d = __ldg(p + (size_t)Wi * C4 + C4);
    const float4 gq = __ldg(reinterpret_cast<const float4*>(dout) + i);
    float4 ga, gb, gc, gd;
    route_max(a.x, bq.x, c.x, d.x, gq.x, ga.x, gb.x, gc.x, gd.x);
    route_max(a.y, bq.y, c.y, d.y, gq.y, ga.y, gb.y, gc.y, gd.y);
    route_max(a.z, bq.z, c.z, d.z, gq.z, ga.z, gb.z, gc.z, gd.z);
    route_max(a.w, bq.w, c.w, d.w, gq.w, ga.w, gb.w, gc.w, gd.w);
    float4* q = reinterpret_cast<float4*>(din) + o00;
    q[0] = ga;
    q[C4] = gb;
    q[(size_t)Wi * C4] = gc;
    q[(size_t)Wi * C4 + C4] = gd;
}

// backward of avgpool3s2 (count_include_pad=False): din[yi,xi] = sum over outputs covering it of
// dout / count(output)
__global__ void avgpool3s2_bwd_kernel(const float* __restrict__ dout, float* __restrict__ din, int B,
                                      int Hi, int Wi, int C, int Ho, int Wo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Hi * Wi * C) return;
    const uint32_t iu = (uint32_t)i;  // < 2^31 elements (host wrapper): 32-bit div/mod
    const int c = (int)(iu % (uint32_t)C);
    uint32_t r = iu / (uint32_t)C;
    const int xi = (int)(r % (uint32_t)Wi);
    r /= (uint32_t)Wi;
    const int yi = (int)(r % (uint32_t)Hi);
    const int b = (int)(r / (uint32_t)Hi);
    float s = 0.f;
    for (int yo = (yi) / 2; yo <= (yi + 1) / 2; ++yo) {
        if (yo >= Ho) continue;
        const int y0 = max(yo * 2 - 1, 0), y1 = min(yo * 2 + 1, Hi - 1);
        for (int xo = (xi) / 2; xo <= (xi + 1) / 2; ++xo) {
            if (xo >= Wo) continue;
            const int x0 = max(xo * 2 - 1, 0), x1 = min(xo * 2 + 1, Wi - 1);
            const int cnt = (y1 - y0 + 1) * (x1 - x0 + 1);
            s += dout[(((size_t)b * Ho + yo) * Wo + xo) * C + c] / (float)cnt;
        }
    }
    din[i] = s;
}

// backward of disc_input wrt the fake image: dfake[b,c,p] = dx[b,p,L+c] (first B images)
__global__ void disc_input_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dfake, int B,
                                      int L, int HW, int Cp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * 3 * HW) return;
    const int p = (int)(i % HW);
    const int c = (int)((i / HW) % 3);
    const int b = (int)(i / ((int64_t)3 * HW));
    dfake[i] = dx[((size_t)b * HW + p) * Cp + L + c];
}

}  // namespace dsee

extern "C" int dsee_act_bwd(const float* dy, const float* out, float* dx, int64_t n, int act,
                            void* stream) {
    DSEE_CHECK_ARG(dy && out && dx && n > 0 && act >= 1 && act <= 3, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    act_bwd_kernel<<<cdiv2(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, out, dx, n, act,
                                                                    act == 3 ? 0.f : 0.2f);
    LAUNCH_END();
}

extern "C" int dsee_conv2d_direct_dgrad(const float* dy, const float* w, float* dx, int B, int Hi,
                                        int Wi, int Cin, int Cout, int KH, int KW, int stride, int pad,
                                        int ups, void* stream) {
    DSEE_CHECK_ARG(dy && w && dx && B > 0 && Hi > 0 && Wi > 0 && Cin > 0 && Cin % 4 == 0 && Cout > 0,
                   "bad argument (Cin must be a multiple of 4)");
    DSEE_CHECK_ARG(stride >= 1 && (ups == 0 || ups == 1), "bad stride/ups");
    int rc = require_sm100();
    if (rc) return rc;
    const int Hu = Hi << ups, Wu = Wi << ups;
    const int Ho = (Hu + 2 * pad - KH) / stride + 1, Wo = (Wu + 2 * pad - KW) / stride + 1;
    int64_t n = (int64_t)B * Hi * Wi * (Cin / 4);
    direct_dgrad_kernel<<<cdiv2(n, 128), 128, 0, (cudaStream_t)stream>>>(
        dy, w, dx, B, Hi, Wi, Cin, Ho, Wo, Cout, KH, KW, stride, pad, ups);
    LAUNCH_END();
}

static int direct_wgrad_splits(int64_t npix, int tiles) {
    int splits = (148 * 4 + tiles - 1) / tiles;
    const int64_t maxs = (npix + 255) / 256;  // at least 256 pixels per split
    if (splits > maxs) splits = (int)maxs;
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
    return splits;
}

extern "C" int64_t dsee_conv2d_direct_wgrad_workspace_floats(int B, int Ho, int Wo, int Cin, int Cout,
                                                             int KH, int KW) {
    const int tiles = KH * KW * ((Cin + DW_T - 1) / DW_T) * ((Cout + DW_T - 1) / DW_T);
    return (int64_t)direct_wgrad_splits((int64_t)B * Ho * Wo, tiles) * KH * KW * Cin * Cout;
}

extern "C" int dsee_conv2d_direct_wgrad(const float* x, const float* dy, float* dw, float* workspace,
                                        int B, int Hi, int Wi, int Cin, int Cout, int KH, int KW,
                                        int stride, int pad, int ups, void* stream) {
    DSEE_CHECK_ARG(x && dy && dw && workspace && B > 0 && Hi > 0 && Wi > 0 && Cin > 0 && Cout > 0,
                   "bad argument");
    DSEE_CHECK_ARG(stride >= 1 && (ups == 0 || ups == 1), "bad stride/ups");
    int rc = require_sm100();
    if (rc) return rc;
    const int Hu = Hi << ups, Wu = Wi << ups;
    const int Ho = (Hu + 2 * pad - KH) / stride + 1, Wo = (Wu + 2 * pad - KW) / stride + 1;
    const int ci_tiles = (Cin + DW_T - 1) / DW_T, co_tiles = (Cout + DW_T - 1) / DW_T;
    const int tiles = KH * KW * ci_tiles * co_tiles;
    const int splits = direct_wgrad_splits((int64_t)B * Ho * Wo, tiles);
    cudaStream_t st = (cudaStream_t)stream;
    direct_wgrad_kernel<<<tiles * splits, 256, 0, st>>>(x, dy, workspace, B, Hi, Wi, Cin, Ho, Wo, Cout,
                                                        KH, KW, stride, pad, ups, splits, ci_tiles,
                                                        co_tiles);
    count_launch();
    const int64_t n = (int64_t)KH * KW * Cin * Cout;
    sum_splits_kernel<<<cdiv2(n, 256), 256, 0, st>>>(workspace, dw, splits, n);
    LAUNCH_END();
}

extern "C" int dsee_channel_sum_chunks(int64_t npix) { return cdiv2(npix, CS_PIX); }

extern "C" int dsee_channel_sum(const float* x, int64_t npix, int C, float* workspace, float* out,
                                void* stream) {
    DSEE_CHECK_ARG(x && workspace && out && npix > 0 && C > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = dsee_channel_sum_chunks(npix);
    channel_sum_kernel<<<dim3(cdiv2(C, 32), chunks), 256, 0, st>>>(x, npix, C, workspace);
    count_launch();
    sum_splits_kernel<<<cdiv2(C, 256), 256, 0, st>>>(workspace, out, chunks, C);
    LAUNCH_END();
}

extern "C" int dsee_instance_norm_bwd(const float* x, const float* dout, const float* mean,
                                      const float* rstd, float* dx, float* sums, void* workspace, int B,
                                      int HW, int C, int act, void* stream) {
    DSEE_CHECK_ARG(x && dout && mean && rstd && dx && sums && workspace && B > 0 && HW > 0 && C > 0 &&
                       C % 4 == 0,
                   "bad argument");
    DSEE_CHECK_ARG(act >= 0 && act <= 2, "act must be 0, 1 or 2");
    DSEE_CHECK_ARG((int64_t)B * HW * C < ((int64_t)1 << 31), "more than 2^31 elements");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = cdiv2(HW, IN_CHUNK);
    instnorm_bwd_stats_kernel<<<dim3(cdiv2(C, 32), B, chunks), 256, 0, st>>>(
        x, dout, mean, rstd, HW, C, act, 0.2f, (double*)workspace);
    count_launch();
    instnorm_bwd_finalize_kernel<<<cdiv2(B * C, 128), 128, 0, st>>>((const double*)workspace, chunks, HW,
                                                                    C, B * C, sums);
    count_launch();
    const int64_t n = (int64_t)B * HW * C;
    instnorm_bwd_apply_kernel<<<cdiv2(n, 256), 256, 0, st>>>(x, dout, mean, rstd, sums, dx, n, HW, C,
                                                             act, 0.2f);
    LAUNCH_END();
}

extern "C" int dsee_region_pool_bwd(const float* dstyle, const uint8_t* labels, float* dx, int B,
                                    int HW, int C, int L, void* stream) {
    DSEE_CHECK_ARG(dstyle && labels && dx && B > 0 && HW > 0 && C > 0 && C % 4 == 0 && L > 0,
                   "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    const int64_t n4 = (int64_t)B * HW * (C / 4);
    region_pool_bwd_kernel<<<cdiv2(n4, 256), 256, 0, (cudaStream_t)stream>>>(dstyle, labels, dx, n4, HW,
                                                                             C, L, 1.0f / (float)HW);
    LAUNCH_END();
}

extern "C" int dsee_avgpool3s2_bwd(const float* dout, float* din, int B, int Hi, int Wi, int C,
                                   void* stream) {
    DSEE_CHECK_ARG(dout && din && B > 0 && Hi > 0 && Wi > 0 && C > 0, "bad argument");
    DSEE_CHECK_ARG((int64_t)B * Hi * Wi * C < ((int64_t)1 << 31), "more than 2^31 elements");
    int rc = require_sm100();
    if (rc) return rc;
    const int Ho = (Hi + 2 - 3) / 2 + 1, Wo = (Wi + 2 - 3) / 2 + 1;
    const int64_t n = (int64_t)B * Hi * Wi * C;
    avgpool3s2_bwd_kernel<<<cdiv2(n, 256), 256, 0, (cudaStream_t)stream>>>(dout, din, B, Hi, Wi, C, Ho,
                                                                           Wo);
    LAUNCH_END();
}

extern "C" int dsee_maxpool2_fwd(const float* in, float* out, int B, int Hi, int Wi, int C, void* stream) {
    DSEE_CHECK_ARG(in && out && B > 0 && Hi >= 2 && Wi >= 2 && C > 0 && C % 4 == 0, "bad argument");
    DSEE_CHECK_ARG((int64_t)B * Hi * Wi * C < ((int64_t)1 << 31), "more than 2^31 elements");
    int rc = require_sm100();
    if (rc) return rc;
    const int Ho = Hi / 2, Wo = Wi / 2;
    const int64_t n = (int64_t)B * Ho * Wo * (C / 4);
    maxpool2_kernel<<<cdiv2(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, Hi, Wi, C / 4, Ho, Wo);
    LAUNCH_END();
}

extern "C" int dsee_maxpool2_bwd(const float* in, const float* dout, float* din, int B, int Hi, int Wi, int C,
                                 void* stream) {
    DSEE_CHECK_ARG(in && dout && din && B > 0 && Hi >= 2 && Wi >= 2 && C > 0 && C % 4 == 0, "bad argument");
    DSEE_CHECK_ARG((int64_t)B * Hi * Wi * C < ((int64_t)1 << 31), "more than 2^31 elements");
    int rc = require_sm100();
    if (rc) return rc;
    const int Ho = Hi / 2, Wo = Wi / 2;
    if ((Hi & 1) || (Wi & 1))
        DSEE_CUDA(cudaMemsetAsync(din, 0, (size_t)B * Hi * Wi * C * sizeof(float), (cudaStream_t)stream));
    const int64_t n = (int64_t)B * Ho * Wo * (C / 4);
    maxpool2_bwd_kernel<<<cdiv2(n, 256), 256, 0, (cudaStream_t)stream>>>(in, dout, din, B, Hi, Wi, C / 4, Ho, Wo);
    LAUNCH_END();
}

extern "C" int dsee_disc_input_bwd(const float* dx, float* dfake, int B, int L, int H, int W, int Cp,
                                   void* stream) {
    DSEE_CHECK_ARG(dx && dfake && B > 0 && L > 0 && Cp >= L + 3, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    const int64_t n = (int64_t)B * 3 * H * W;
    disc_input_bwd_kernel<<<cdiv2(n, 256), 256, 0, (cudaStream_t)stream>>>(dx, dfake, B, L, H * W, Cp);
    LAUNCH_END();
}
