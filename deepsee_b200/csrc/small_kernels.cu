// HBM-bound helper kernels around the two tensor-core kernels: label-map integer work,
// conditional-norm operand builders (table gather instead of a conv over a one-hot map),
// weight / activation splitting into fp16 planes, batch-norm statistics, generator stem and head.
// All are simple coalesced / vectorised streaming kernels; none is reshaped into a GEMM.
#include "common.cuh"
#include "launch_count.h"
#include "../../include/deepsee_b200.h"

#include <math.h>
#include <cuda_fp8.h>

namespace dsee {

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ void split_f16(float a, __half& hi, __half& lo) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    hi = __float2half_rn(a);
    lo = __float2half_rn(a - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// ------------------------------------------------------------------------------------------------
// label maps
// ------------------------------------------------------------------------------------------------
__global__ void onehot_from_labels_kernel(const int64_t* __restrict__ label, float* __restrict__ out,
                                          int B, int L, int HW, int* bad) {
    // thread per (b, pixel); writes L planes (coalesced along pixels)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HW) return;
    int b = (int)(i / HW), px = (int)(i % HW);
    int64_t l = label[i];
    if (l < 0 || l >= L) {
        *bad = 1;
        l = -1;
    }
    float* o = out + (size_t)b * L * HW + px;
    for (int c = 0; c < L; ++c) o[(size_t)c * HW] = (c == l) ? 1.0f : 0.0f;
}

__global__ void labels_from_onehot_kernel(const float* __restrict__ oh, uint8_t* __restrict__ labels,
                                          int B, int L, int HW, int* bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * HW) return;
    int b = (int)(i / HW), px = (int)(i % HW);
    const float* p = oh + (size_t)b * L * HW + px;
    int found = -1, ones = 0;
    bool other = false;
    for (int c = 0; c < L; ++c) {
        float v = p[(size_t)c * HW];
        if (v == 1.0f) {
            ++ones;
            found = c;
        } else if (v != 0.0f) {
            other = true;
        }
    }
    if (ones != 1 || other) {
        *bad = 1;
        if (found < 0) found = 0;
    }
    labels[i] = (uint8_t)found;
}

// int64 label map (the dataloader's) -> uint8, flagging values outside [0, L) (the reference's
// scatter_ would raise on those, preprocessor.py:40)
__global__ void labels_u8_kernel(const int64_t* __restrict__ label, uint8_t* __restrict__ out, int64_t n,
                                 int L, int* bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t l = label[i];
    if (l < 0 || l >= L) {
        *bad = 1;
        l = 0;
    }
    out[i] = (uint8_t)l;
}

// F.interpolate(mode='bicubic', align_corners=False) + clamp(-1, 1) (preprocessor.py:29-32), NCHW fp32.
// Same arithmetic as ATen's upsample_bicubic2d: source index scale * (dst + 0.5) - 0.5 (scale =
// in / out as float), cubic convolution coefficients with A = -0.75, border taps clamped to the image.
__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
    const float A = -0.75f;
    c[0] = cubic2(t + 1.f, A);
    c[1] = cubic1(t, A);
    const float x2 = 1.f - t;
    c[2] = cubic1(x2, A);
    c[3] = cubic2(x2 + 1.f, A);
}
__global__ void bicubic_clamp_kernel(const float* __restrict__ in, float* __restrict__ out, int BC, int Hi,
                                     int Wi, int Ho, int Wo, float sh, float sw) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)BC * Ho * Wo) return;
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho), bc = (int)(i / ((int64_t)Wo * Ho));
    const float ry = sh * (oy + 0.5f) - 0.5f, rx = sw * (ox + 0.5f) - 0.5f;
    const int iy = (int)floorf(ry), ix = (int)floorf(rx);
    float cy[4], cx[4];
    cubic_coeffs(ry - iy, cy);
    cubic_coeffs(rx - ix, cx);
    const float* img = in + (size_t)bc * Hi * Wi;
    float rows[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int y = min(max(iy - 1 + k, 0), Hi - 1);
        const float* r = img + (size_t)y * Wi;
        const float v0 = r[min(max(ix - 1, 0), Wi - 1)], v1 = r[min(max(ix, 0), Wi - 1)];
        const float v2 = r[min(max(ix + 1, 0), Wi - 1)], v3 = r[min(max(ix + 2, 0), Wi - 1)];
        rows[k] = v0 * cx[0] + v1 * cx[1] + v2 * cx[2] + v3 * cx[3];
    }
    const float v = rows[0] * cy[0] + rows[1] * cy[1] + rows[2] * cy[2] + rows[3] * cy[3];
    out[i] = fminf(fmaxf(v, -1.f), 1.f);
}

__global__ void resize_labels_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                     int B, int Hin, int Win, int Hout, int Wout, float sh,
                                     float sw) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Hout * Wout) return;
    int x = (int)(i % Wout);
    int y = (int)((i / Wout) % Hout);
    int b = (int)(i / ((int64_t)Wout * Hout));
    // ATen upsample_nearest: src = min(floor(dst * scale), in - 1), scale = in / out (float)
    int ys = min((int)floorf(y * sh), Hin - 1);
    int xs = min((int)floorf(x * sw), Win - 1);
    out[i] = in[((size_t)b * Hin + ys) * Win + xs];
}

// ------------------------------------------------------------------------------------------------
// conditional-norm operand builders
// ------------------------------------------------------------------------------------------------
// actv = relu(bias + sum_tap table[tap][label(tap)]) ; thread = (pixel, 4 hidden channels)
// rows[l][o] = bias[o] + sum_{tap = 0..8} table[tap][l][o], added in the order the gather kernel uses:
// the pre-activation of a pixel whose whole 3x3 window carries label l (the interior of a region)
__global__ void shared_mlp_rows_kernel(const float* __restrict__ table, const float* __restrict__ bias,
                                       float* __restrict__ rows, int L, int nh) {
    const int l = blockIdx.x, g = threadIdx.x;  // g: group of 4 channels
    if (g >= (nh >> 2)) return;
    float4 acc = __ldg(reinterpret_cast<const float4*>(bias) + g);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(table + ((size_t)tap * L + l) * nh) + g);
        acc.x += t.x;
        acc.y += t.y;
        acc.z += t.z;
        acc.w += t.w;
    }
    reinterpret_cast<float4*>(rows + (size_t)l * nh)[g] = acc;
}

__global__ void shared_mlp_kernel(const uint8_t* __restrict__ labels, const float* __restrict__ table,
                                  const float* __restrict__ bias, const float* __restrict__ rows,
                                  __half* __restrict__ out_hi, __half* __restrict__ out_lo, int B, int Hl,
                                  int Wl, int ups, int L, int nh) {
    // 32-bit index arithmetic (the host wrapper checks the element count): five 64-bit div/mod per
    // thread cost more than the 8 bytes the thread produces
    const uint32_t groups = (uint32_t)nh >> 2;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int H = Hl << ups, W = Wl << ups;
    if (i >= (uint32_t)B * H * W * groups) return;
    const int g = (int)(i % groups);
    const uint32_t pix = i / groups;
    const int x = (int)(pix % (uint32_t)W);
    const uint32_t t2 = pix / (uint32_t)W;
    const int y = (int)(t2 % (uint32_t)H);
    const int b = (int)(t2 / (uint32_t)H);
    const int yl = y >> ups, xl = x >> ups;
    const uint8_t* lb = labels + (size_t)b * Hl * Wl;
    // the window's nine labels (255 = outside the image: zero padding)
    int lab[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int yy = yl + tap / 3 - 1, xx = xl + tap % 3 - 1;
        lab[tap] = (yy < 0 || yy >= Hl || xx < 0 || xx >= Wl) ? 255 : (int)lb[yy * Wl + xx];
    }
    bool uniform = rows != nullptr;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) uniform = uniform && lab[tap] == lab[4];
    float4 acc;
    if (uniform) {
        // interior of a region (most pixels of a face parse): one precomputed row instead of nine - the same
        // fp32 additions in the same order, so the result is bit-identical to the general path
        acc = __ldg(reinterpret_cast<const float4*>(rows + (size_t)lab[4] * nh) + g);
    } else {
        acc = __ldg(reinterpret_cast<const float4*>(bias) + g);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            if (lab[tap] == 255) continue;  // zero padding
            float4 t = __ldg(reinterpret_cast<const float4*>(table + ((size_t)tap * L + lab[tap]) * nh) + g);
            acc.x += t.x;
            acc.y += t.y;
            acc.z += t.z;
            acc.w += t.w;
        }
    }
    float a[4] = {fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f)};
    __half h[4], l4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_f16(a[e], h[e], l4[e]);
    size_t o = (size_t)pix * nh + (size_t)g * 4;
    *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(pack2(h[0], h[1]), pack2(h[2], h[3]));
    if (out_lo)
        *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(pack2(l4[0], l4[1]), pack2(l4[2], l4[3]));
}

// The same gather with four consecutive output pixels of a row per thread (W % 4 == 0): the 3 x 6 (ups = 0)
// or 3 x 4 (ups = 1: the four pixels share two low-resolution columns) label strip is loaded once, the
// index arithmetic is paid once per quad, and four 8-byte stores leave per thread.  Same additions in the
// same order as shared_mlp_kernel: bit-identical planes.
__global__ void shared_mlp_x4_kernel(const uint8_t* __restrict__ labels, const float* __restrict__ table,
                                     const float* __restrict__ bias, const float* __restrict__ rows,
                                     __half* __restrict__ out_hi, __half* __restrict__ out_lo, int B, int Hl,
                                     int Wl, int ups, int L, int nh) {
    const uint32_t groups = (uint32_t)nh >> 2;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int H = Hl << ups, W = Wl << ups, Wq = W >> 2;
    if (i >= (uint32_t)B * H * Wq * groups) return;
    const int g = (int)(i % groups);
    const uint32_t quad = i / groups;
    const int x0 = (int)(quad % (uint32_t)Wq) * 4;
    const uint32_t t2 = quad / (uint32_t)Wq;
    const int y = (int)(t2 % (uint32_t)H);
    const int b = (int)(t2 / (uint32_t)H);
    const int yl = y >> ups, xl0 = x0 >> ups;
    const int ncol = ups ? 2 : 4;   // low-resolution columns under the quad
    const uint8_t* lb = labels + (size_t)b * Hl * Wl;
    int lab[3][6];   // rows yl-1 .. yl+1, columns xl0-1 .. xl0+ncol (255 = outside the image: zero padding)
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const int yy = yl - 1 + r, xx = xl0 - 1 + c;
            lab[r][c] = (c >= ncol + 2 || yy < 0 || yy >= Hl || xx < 0 || xx >= Wl) ? 255 : (int)lb[yy * Wl + xx];
        }
    const float4 bias4 = __ldg(reinterpret_cast<const float4*>(bias) + g);
    const size_t o0 = ((size_t)(((size_t)b * H + y) * W + x0)) * nh + (size_t)g * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= ncol) break;
        bool uniform = rows != nullptr;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) uniform = uniform && lab[tap / 3][j + tap % 3] == lab[1][j + 1];
        float4 acc;
        if (uniform) {
            acc = __ldg(reinterpret_cast<const float4*>(rows + (size_t)lab[1][j + 1] * nh) + g);
        } else {
            acc = bias4;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int l = lab[tap / 3][j + tap % 3];
                if (l == 255) continue;  // zero padding
                const float4 t = __ldg(reinterpret_cast<const float4*>(table + ((size_t)tap * L + l) * nh) + g);
                acc.x += t.x;
                acc.y += t.y;
                acc.z += t.z;
                acc.w += t.w;
            }
        }
        const float a[4] = {fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f)};
        __half h[4], l4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_f16(a[e], h[e], l4[e]);
        const uint2 vh = make_uint2(pack2(h[0], h[1]), pack2(h[2], h[3]));
        const uint2 vl = make_uint2(pack2(l4[0], l4[1]), pack2(l4[2], l4[3]));
        const int reps = ups ? 2 : 1;   // ups: the low-resolution column covers two output pixels
        for (int k = 0; k < reps; ++k) {
            const size_t o = o0 + (size_t)(j * reps + k) * nh;
            *reinterpret_cast<uint2*>(out_hi + o) = vh;
            if (out_lo) *reinterpret_cast<uint2*>(out_lo + o) = vl;
        }
    }
}

__global__ void style_gather_kernel(const uint8_t* __restrict__ labels, const float* __restrict__ style,
                                    __half* __restrict__ out_hi, __half* __restrict__ out_lo, int B,
                                    int HW, int L, int d) {
    const uint32_t groups = (uint32_t)d >> 2;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // 32-bit: checked by the host wrapper
    if (i >= (uint32_t)B * HW * groups) return;
    const int g = (int)(i % groups);
    const uint32_t pix = i / groups;
    const int b = (int)(pix / (uint32_t)HW);
    int l = labels[pix];
    if (l >= L) l = L - 1;
    float4 s = __ldg(reinterpret_cast<const float4*>(style + ((size_t)b * L + l) * d) + g);
    float a[4] = {s.x, s.y, s.z, s.w};
    __half h[4], l4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_f16(a[e], h[e], l4[e]);
    size_t o = (size_t)pix * d + (size_t)g * 4;
    *reinterpret_cast<uint2*>(out_hi + o) = make_uint2(pack2(h[0], h[1]), pack2(h[2], h[3]));
    if (out_lo)
        *reinterpret_cast<uint2*>(out_lo + o) = make_uint2(pack2(l4[0], l4[1]), pack2(l4[2], l4[3]));
}

// ------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------
__global__ void amax_kernel(const float* __restrict__ w, int64_t n, float* scratch) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(scratch), __float_as_uint(m));
}

__global__ void prep_weight_kernel(const float* __restrict__ w, __half* __restrict__ out_hi,
                                   __half* __restrict__ out_lo, float* inv_scale, int N, int C,
                                   int transpose) {
    // out[n][tap][c] = w[n][c][tap] * scale      (transpose: out[c][8-tap][n])
    const float scale = pow2_scale_for(inv_scale[1], 14);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) inv_scale[0] = 1.f / scale;
    if (i >= (int64_t)N * 9 * C) return;
    const uint32_t iu = (uint32_t)i;  // < 2^31 elements (host wrapper): 32-bit div/mod
    int c = (int)(iu % (uint32_t)C);
    int tap = (int)((iu / (uint32_t)C) % 9u);
    int n = (int)(iu / (9u * (uint32_t)C));
    float v = w[((size_t)n * C + c) * 9 + tap] * scale;
    __half h, l;
    split_f16(v, h, l);
    const size_t o = transpose ? ((size_t)c * 9 + (8 - tap)) * N + n : (size_t)i;
    out_hi[o] = h;
    if (out_lo) out_lo[o] = l;
}

// Non-transposed form, one block per filter n: w[n] (C x 9 floats, contiguous) staged in shared memory,
// written tap-major as half2 pairs (coalesced on both sides).
__global__ void __launch_bounds__(256) prep_weight_rows_kernel(const float* __restrict__ w,
                                                               __half* __restrict__ out_hi,
                                                               __half* __restrict__ out_lo, float* inv_scale,
                                                               int C) {
    extern __shared__ float srow[];  // [C * 9]
    const float scale = pow2_scale_for(inv_scale[1], 14);
    if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[0] = 1.f / scale;
    const int n = blockIdx.x, len = C * 9;
    const float* r = w + (size_t)n * len;
    for (int i = threadIdx.x; i < len; i += blockDim.x) srow[i] = r[i];
    __syncthreads();
    const int hp = C >> 1;
    uint32_t* oh = reinterpret_cast<uint32_t*>(out_hi + (size_t)n * len);
    uint32_t* ol = out_lo ? reinterpret_cast<uint32_t*>(out_lo + (size_t)n * len) : nullptr;
    for (int e = threadIdx.x; e < 9 * hp; e += blockDim.x) {
        const int tap = e / hp, c = 2 * (e - tap * hp);
        uint32_t h, l;
        split_half2_sat(srow[c * 9 + tap] * scale, srow[(c + 1) * 9 + tap] * scale, h, l);
        oh[e] = h;
        if (ol) ol[e] = l;
    }
}

// fp8 companion planes of a prepared weight (dsee_conv_operands.passes == 2):
// out8[n][tap][0][c] = e4m3(ws * 2^-8), out8[n][tap][1][c] = e4m3(ws - fp16(ws)), ws = w * 2^e
__global__ void prep_weight_f8_kernel(const float* __restrict__ w, const float* __restrict__ inv_scale,
                                      uint8_t* __restrict__ out8, int N, int C) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint32_t)N * 9u * (uint32_t)C) return;
    const float scale = 1.f / __ldg(inv_scale);
    const int c = (int)(i % (uint32_t)C);
    const int tap = (int)((i / (uint32_t)C) % 9u);
    const int n = (int)(i / (9u * (uint32_t)C));
    const float v = fminf(fmaxf(w[((size_t)n * C + c) * 9 + tap] * scale, -65504.f), 65504.f);
    const float lo = v - __half2float(__float2half_rn(v));
    uint8_t* o = out8 + ((size_t)n * 9 + tap) * 2 * C + c;
    o[0] = __nv_cvt_float_to_fp8(v * (1.f / 256.f), __NV_SATFINITE, __NV_E4M3);
    o[C] = __nv_cvt_float_to_fp8(lo, __NV_SATFINITE, __NV_E4M3);
}

// Batched modulation weight (see dsee_prep_mod_weight_batched): out[b*N + n][tap][c],
// c < Ca: wa[n][c][tap];  Ca <= c < Ca+Ls: ws[b][n][c-Ca][tap];  else 0.
// One block per output row (b, n): the row's sources (contiguous in c-major / tap-minor order) are
// staged in shared memory with coalesced loads, then written out tap-major as half2 pairs.
__global__ void __launch_bounds__(256) prep_mod_weight_batched_kernel(
    const float* __restrict__ wa, const float* __restrict__ ws, __half* __restrict__ out_hi,
    __half* __restrict__ out_lo, float* inv_scale, int B, int N, int Ca, int Ls, int Lp) {
    extern __shared__ float srow[];  // [Ca * 9] from wa, then [Ls * 9] from ws
    const float scale = pow2_scale_for(inv_scale[1], 14);
    if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[0] = 1.f / scale;
    const int row = blockIdx.x;  // b * N + n
    const int n = row % N;
    const int na = Ca * 9, ns = Ls * 9;
    const float* ra = wa + (size_t)n * na;
    const float* rs = ws + (size_t)row * ns;
    for (int i = threadIdx.x; i < na; i += blockDim.x) srow[i] = ra[i];
    for (int i = threadIdx.x; i < ns; i += blockDim.x) srow[na + i] = rs[i];
    __syncthreads();
    const int Ct = Ca + Lp, hp = Ct >> 1;
    uint32_t* oh = reinterpret_cast<uint32_t*>(out_hi + (size_t)row * 9 * Ct);
    uint32_t* ol = out_lo ? reinterpret_cast<uint32_t*>(out_lo + (size_t)row * 9 * Ct) : nullptr;
    for (int e = threadIdx.x; e < 9 * hp; e += blockDim.x) {
        const int tap = e / hp, c = 2 * (e - tap * hp);
        float v[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int cc = c + k;   // (Ca and Ls + Ca boundaries are even or the pair straddles into zeros)
            v[k] = cc < Ca + Ls ? srow[cc * 9 + tap] * scale : 0.f;
        }
        uint32_t h, l;
        split_half2_sat(v[0], v[1], h, l);
        oh[e] = h;
        if (ol) ol[e] = l;
    }
}

// max over rows of sum_k |hi[row][k]| * inv_scale: one block per row of the prepared (scaled fp16)
// planes; bounds |sum_k a[k] * w[row][k]| <= max|a| * result.
__global__ void row_l1max_kernel(const __half* __restrict__ hi, int rowlen, float* inv_scale) {
    __shared__ float red[8];
    const __half* r = hi + (size_t)blockIdx.x * rowlen;
    float a = 0.f;
    for (int i = threadIdx.x; i < rowlen; i += blockDim.x) a += fabsf(__half2float(r[i]));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
        // 1 + 2^-9: the hi plane is w * scale rounded to 11 bits
        atomic_max_nonneg(inv_scale + 2, t * inv_scale[0] * 1.002f);
    }
}

__global__ void prep_weight_ex_kernel(const float* __restrict__ w, __half* __restrict__ out_hi,
                                      __half* __restrict__ out_lo, float* inv_scale, int N, int C,
                                      int T, int transpose) {
    // normal:    out[n][tap][c (padded to Cp)];   transpose: out[c][tap][n (padded to Np)]
    const float scale = pow2_scale_for(inv_scale[1], 14);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) inv_scale[0] = 1.f / scale;
    const int rows = transpose ? C : N, cols = transpose ? N : C;
    const int cp = (cols + 63) / 64 * 64;
    if (i >= (int64_t)rows * T * cp) return;
    const uint32_t iu = (uint32_t)i;  // < 2^31 elements (host wrapper): 32-bit div/mod
    const int cc = (int)(iu % (uint32_t)cp);
    const int tap = (int)((iu / (uint32_t)cp) % (uint32_t)T);
    const int r = (int)(iu / ((uint32_t)T * (uint32_t)cp));
    float v = 0.f;
    if (cc < cols) {
        const int n = transpose ? cc : r, c = transpose ? r : cc;
        v = w[((size_t)n * C + c) * T + tap] * scale;
    }
    __half h, l;
    split_f16(v, h, l);
    out_hi[i] = h;
    if (out_lo) out_lo[i] = l;
}

__global__ void split_kernel(const float* __restrict__ in, __half* __restrict__ hi,
                             __half* __restrict__ lo, int64_t n4) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    float a[4] = {v.x, v.y, v.z, v.w};
    __half h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_f16(a[e], h[e], l[e]);
    reinterpret_cast<uint2*>(hi)[i] = make_uint2(pack2(h[0], h[1]), pack2(h[2], h[3]));
    if (lo) reinterpret_cast<uint2*>(lo)[i] = make_uint2(pack2(l[0], l[1]), pack2(l[2], l[3]));
}

// fp32 NHWC [B,H,W,C] -> fp16 planes [B,2H,2W,C] (nearest 2x upsample materialised for the
// encoder's up_conv / conv2 input, encoder.py:94-95,153-154)
__global__ void split_ups_kernel(const float* __restrict__ in, __half* __restrict__ hi,
                                 __half* __restrict__ lo, int B, int H, int W, int C4) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * H * W * C4) return;
    const int g = (int)(i % C4);
    int64_t r = i / C4;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    float a[4] = {v.x, v.y, v.z, v.w};
    __half h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_f16(a[e], h[e], l[e]);
    const uint2 ph = make_uint2(pack2(h[0], h[1]), pack2(h[2], h[3]));
    const uint2 pl = make_uint2(pack2(l[0], l[1]), pack2(l[2], l[3]));
#pragma unroll
    for (int s2 = 0; s2 < 4; ++s2) {
        const size_t o = (((size_t)b * 2 * H + 2 * y + (s2 >> 1)) * 2 * W + 2 * x + (s2 & 1)) * C4 + g;
        reinterpret_cast<uint2*>(hi)[o] = ph;
        if (lo) reinterpret_cast<uint2*>(lo)[o] = pl;
    }
}

// dY [B,2Ho,2Wo,C] -> [B,Ho,Wo,C] summing 2x2 (transpose of the nearest upsample)
__global__ void fold2x2_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int Ho,
                               int Wo, int C4) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * Ho * Wo * C4) return;
    const int g = (int)(i % C4);
    int64_t r = i / C4;
    const int xo = (int)(r % Wo);
    r /= Wo;
    const int yo = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const int W = Wo * 2, H = Ho * 2;
    const float4* base = reinterpret_cast<const float4*>(in);
    float4 a = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int s2 = 0; s2 < 4; ++s2) {
        const size_t fp = ((size_t)b * H + yo * 2 + (s2 >> 1)) * W + xo * 2 + (s2 & 1);
        const float4 v = __ldg(base + fp * C4 + g);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = a;
}

// ------------------------------------------------------------------------------------------------
// batch-norm statistics
// ------------------------------------------------------------------------------------------------
constexpr int STAT_PIX = 256;  // pixels per partial

// block = 256 threads: thread t owns channels 4*(t % (C/4)) .. +3 and every (256/(C/4))-th pixel.
__global__ void noise_fill_kernel(unsigned long long seed, float* __restrict__ out, int64_t n4,
                                  const unsigned long long* epoch) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    seed = eff_noise_seed(seed, epoch);
    reinterpret_cast<float4*>(out)[i] = noise_normal4(seed, (unsigned long long)i);
}

template <bool HAS_NOISE>
__global__ void bn_stats_kernel(const float* __restrict__ x, int x_ups, const float* __restrict__ noise,
                                unsigned long long noise_seed, const float* __restrict__ noise_w, int B,
                                int H, int W, int C, float* __restrict__ partial,
                                const unsigned long long* epoch) {
    extern __shared__ float red[];  // [pix_lanes][C][2]
    if (HAS_NOISE) noise_seed = eff_noise_seed(noise_seed, epoch);
    const int cg = C >> 2;
    const int pix_lanes = blockDim.x / cg;
    const int g = threadIdx.x % cg, pl = threadIdx.x / cg;
    const int64_t npix = (int64_t)B * H * W;
    const int64_t p0 = (int64_t)blockIdx.x * STAT_PIX;
    const int Hx = H >> x_ups, Wx = W >> x_ups;
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    float4 nw = make_float4(0, 0, 0, 0);
    constexpr bool has_noise = HAS_NOISE;
    if (has_noise) nw = __ldg(reinterpret_cast<const float4*>(noise_w) + g);
    if (pl < pix_lanes) {
        const int iend = (int)(npix - p0 < STAT_PIX ? npix - p0 : STAT_PIX);
        // (b, yy, xx) of this lane's first pixel by one 32-bit division, then advanced incrementally
        // (per-pixel div/mod would cost more issue slots than the noise generator)
        const uint32_t pix0 = (uint32_t)(p0 + pl);
        int xx = (int)(pix0 % (uint32_t)W);
        int yy = (int)((pix0 / (uint32_t)W) % (uint32_t)H);
        int b = (int)(pix0 / ((uint32_t)W * (uint32_t)H));
        xx -= pix_lanes;
#pragma unroll(HAS_NOISE ? 1 : 4)
        for (int i = pl; i < iend; i += pix_lanes) {
            int64_t pix = p0 + i;
            xx += pix_lanes;
            while (xx >= W) {
                xx -= W;
                if (++yy == H) {
                    yy = 0;
                    ++b;
                }
            }
            size_t xp = (size_t)pix;
            if (x_ups) xp = ((size_t)b * Hx + (yy >> 1)) * Wx + (xx >> 1);
            float4 v = __ldg(reinterpret_cast<const float4*>(x + xp * C) + g);
            if (has_noise) {
                float4 nv = load_noise4(noise, noise_seed, (size_t)pix * C + g * 4);
                v.x += nw.x * nv.x;
                v.y += nw.y * nv.y;
                v.z += nw.z * nv.z;
                v.w += nw.w * nv.w;
            }
            s1[0] += v.x; s2[0] += v.x * v.x;
            s1[1] += v.y; s2[1] += v.y * v.y;
            s1[2] += v.z; s2[2] += v.z * v.z;
            s1[3] += v.w; s2[3] += v.w * v.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            red[((size_t)pl * C + g * 4 + e) * 2] = s1[e];
            red[((size_t)pl * C + g * 4 + e) * 2 + 1] = s2[e];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, q = 0.f;
        for (int l = 0; l < pix_lanes; ++l) {  // fixed order -> deterministic
            a += red[((size_t)l * C + c) * 2];
            q += red[((size_t)l * C + c) * 2 + 1];
        }
        partial[((size_t)blockIdx.x * C + c) * 2] = a;
        partial[((size_t)blockIdx.x * C + c) * 2 + 1] = q;
    }
}

// grid = C/32 blocks, block = 32 channels x 32 partial lanes; double accumulation, fixed order.
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ partial, int n_partials, int C,
                                   double count, double ucount, float eps, float momentum,
                                   float* running_mean,
                                   float* running_var, float* bn_scale, float* bn_shift,
                                   float* mean_out, float* var_out) {
    __shared__ double sh[32][32][2];
    const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double a = 0.0, q = 0.0;
    if (c < C) {
        for (int s = g; s < n_partials; s += 32) {
            float2 v = __ldg(reinterpret_cast<const float2*>(partial) + (size_t)s * C + c);
            a += (double)v.x;
            q += (double)v.y;
        }
    }
    sh[g][cl][0] = a;
    sh[g][cl][1] = q;
    __syncthreads();
    if (g == 0 && c < C) {
        double A = 0.0, Q = 0.0;
        for (int k = 0; k < 32; ++k) {
            A += sh[k][cl][0];
            Q += sh[k][cl][1];
        }
        double mean = A / count;
        double var = Q / count - mean * mean;  // biased (batchnorm.py:86-88)
        if (var < 0.0) var = 0.0;
        float rstd = (float)(1.0 / sqrt(var + (double)eps));
        bn_scale[c] = rstd;
        bn_shift[c] = (float)(-mean) * rstd;
        if (mean_out) mean_out[c] = (float)mean;
        if (var_out) var_out[c] = (float)var;
        if (running_mean) {
            double unbias = ucount > 1.0 ? var * ucount / (ucount - 1.0) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbias;
        }
    }
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ rm, const float* __restrict__ rv,
                                      float eps, int C, float* sc, float* sh) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float r = 1.0f / sqrtf(rv[c] + eps);
    sc[c] = r;
    sh[c] = -rm[c] * r;
}

// ------------------------------------------------------------------------------------------------
// generator stem / head
// ------------------------------------------------------------------------------------------------
// block per pixel, thread per 4 output channels; weights re-read through L1 (13.8 K floats)
__global__ void stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                            const float* __restrict__ bias, float* __restrict__ out, int B, int H,
                            int W, int C) {
    const int pix = blockIdx.x;
    const int xx = pix % W, yy = (pix / W) % H, b = pix / (W * H);
    __shared__ float patch[27];
    if (threadIdx.x < 27) {
        int ci = threadIdx.x / 9, tap = threadIdx.x % 9;
        int y2 = yy + tap / 3 - 1, x2 = xx + tap % 3 - 1;
        float v = 0.f;
        if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) v = x[(((size_t)b * 3 + ci) * H + y2) * W + x2];
        patch[threadIdx.x] = v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = bias[c];
        const float* wc = w + (size_t)c * 27;
#pragma unroll
        for (int k = 0; k < 27; ++k) acc += wc[k] * patch[k];
        out[(size_t)pix * C + c] = acc;
    }
}

// Image head: tanh(conv3x3(leaky_relu(x), 512 -> 3) + bias), NHWC fp32 in, NCHW fp32 out.
// HBM-bound by design (x is read once, 4 B per element): a block owns a 16 x 64 pixel tile and walks
// the channels in slabs of 8.  Each slab's halo tile (18 x 66 pixels, LeakyReLU applied) is staged in
// shared memory channel-major ([c][row][col]) through registers, double-buffered so the next slab's
// global loads are in flight while the current one is consumed.  A thread owns 4 horizontally
// adjacent pixels x 3 outputs; per channel and filter row it reads 6 consecutive inputs
// (LDS.128 + LDS.64, conflict-free) and the 27 weights of that channel as broadcast LDS.128.
constexpr int HT_H = 16, HT_W = 64, HT_CS = 8, HT_WP = 68;
constexpr int HT_NPIX = (HT_H + 2) * (HT_W + 2);
constexpr int HT_NLD = (2 * HT_NPIX + 255) / 256;
constexpr int HT_XS = HT_CS * (HT_H + 2) * HT_WP;  // floats per x buffer
constexpr int HT_SMEM = (2 * HT_XS + 2 * HT_CS * 28) * (int)sizeof(float);
__global__ void __launch_bounds__(256, 2)
head_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
            float* __restrict__ out, int B, int H, int W, int C) {
    extern __shared__ __align__(16) float hsm[];
    float* xs = hsm;                 // [2][HT_CS][HT_H + 2][HT_WP]
    float* ws = hsm + 2 * HT_XS;     // [2][HT_CS][28]: index (ky*3 + kx)*3 + o
    const int tiles_w = (W + HT_W - 1) / HT_W, tiles_h = (H + HT_H - 1) / HT_H;
    const int tile = blockIdx.x;
    const int w0 = (tile % tiles_w) * HT_W;
    const int h0 = ((tile / tiles_w) % tiles_h) * HT_H;
    const int b = tile / (tiles_w * tiles_h);
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int nslab = C / HT_CS;
    float4 stage[HT_NLD];

    // staging slots of this thread, decoded once: (row, col, channel half) of the halo tile packed in
    // one register each (the slab index is the only thing that changes between iterations)
    uint32_t slot[HT_NLD];
#pragma unroll
    for (int k = 0; k < HT_NLD; ++k) {
        const int e = t + k * 256;
        slot[k] = 0xffffffffu;
        if (e < 2 * HT_NPIX) {
            const int half = e / HT_NPIX, pix = e % HT_NPIX;
            const int r = pix / (HT_W + 2), cidx = pix % (HT_W + 2);
            const int y = h0 - 1 + r, xg = w0 - 1 + cidx;
            const uint32_t inside = (y >= 0 && y < H && xg >= 0 && xg < W) ? 1u : 0u;
            slot[k] = (uint32_t)r | ((uint32_t)cidx << 5) | ((uint32_t)half << 12) | (inside << 13);
        }
    }
    const float* xbase = x + (((size_t)b * H + (h0 - 1)) * W + (w0 - 1)) * C;  // halo origin (may be out of range)
    auto gload = [&](int slab) {
#pragma unroll
        for (int k = 0; k < HT_NLD; ++k) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const uint32_t sl = slot[k];
            if (sl != 0xffffffffu && (sl >> 13)) {
                const int r = sl & 31, cidx = (sl >> 5) & 127, half = (sl >> 12) & 1;
                v = __ldg(reinterpret_cast<const float4*>(xbase + ((ptrdiff_t)r * W + cidx) * C + slab * HT_CS +
                                                          half * 4));
                v.x = v.x > 0.f ? v.x : 0.2f * v.x;
                v.y = v.y > 0.f ? v.y : 0.2f * v.y;
                v.z = v.z > 0.f ? v.z : 0.2f * v.z;
                v.w = v.w > 0.f ? v.w : 0.2f * v.w;
            }
            stage[k] = v;
        }
    };
    auto sstore = [&](int buf, int slab) {
#pragma unroll
        for (int k = 0; k < HT_NLD; ++k) {
            const uint32_t sl = slot[k];
            if (sl != 0xffffffffu) {
                const int r = sl & 31, cidx = (sl >> 5) & 127, half = (sl >> 12) & 1;
                float* d = xs + buf * HT_XS + (half * 4) * (HT_H + 2) * HT_WP + r * HT_WP + cidx;
                d[0] = stage[k].x;
                d[(HT_H + 2) * HT_WP] = stage[k].y;
                d[2 * (HT_H + 2) * HT_WP] = stage[k].z;
                d[3 * (HT_H + 2) * HT_WP] = stage[k].w;
            }
        }
        if (t < HT_CS * 27) {
            const int c = t / 27, k = t % 27, tap = k / 3, o = k % 3;
            ws[(buf * HT_CS + c) * 28 + k] = __ldg(w + ((size_t)o * C + slab * HT_CS + c) * 9 + tap);
        }
    };

    float acc[4][3];
#pragma unroll
    for (int px = 0; px < 4; ++px) acc[px][0] = acc[px][1] = acc[px][2] = 0.f;

    gload(0);
    sstore(0, 0);
    __syncthreads();
    for (int slab = 0; slab < nslab; ++slab) {
        const int cur = slab & 1;
        if (slab + 1 < nslab) gload(slab + 1);
#pragma unroll 2
        for (int c = 0; c < HT_CS; ++c) {
            float wv[28];
            const float4* wp = reinterpret_cast<const float4*>(ws + (cur * HT_CS + c) * 28);
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const float4 q = wp[j];
                wv[4 * j] = q.x; wv[4 * j + 1] = q.y; wv[4 * j + 2] = q.z; wv[4 * j + 3] = q.w;
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float* row = xs + cur * HT_XS + (c * (HT_H + 2) + ty + ky) * HT_WP + 4 * tx;
                const float4 a4 = *reinterpret_cast<const float4*>(row);
                const float2 b2 = *reinterpret_cast<const float2*>(row + 4);
                const float v[6] = {a4.x, a4.y, a4.z, a4.w, b2.x, b2.y};
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int px = 0; px < 4; ++px)
#pragma unroll
                        for (int o = 0; o < 3; ++o) acc[px][o] += v[px + kx] * wv[(ky * 3 + kx) * 3 + o];
            }
        }
        if (slab + 1 < nslab) sstore(cur ^ 1, slab + 1);
        __syncthreads();
    }
    const int y = h0 + ty;
    if (y < H) {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const float bo = __ldg(bias + o);
            float* orow = out + (((size_t)b * 3 + o) * H + y) * W;
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                const int xg = w0 + 4 * tx + px;
                if (xg < W) orow[xg] = tanhf(acc[px][o] + bo);
            }
        }
    }
}

// ---- tensor-core image head: the 512 -> 3 conv as a 1x1 GEMM into 27 (tap, channel) partial products
// per input pixel (dsee_conv2d_tc, N = 32) and these two HBM-trivial kernels around it ------------
// out[b,c,y,x] = tanh(bias[c] + sum_tap P[b, y+dy, x+dx, tap*3 + c])   (P zero outside the image)
__global__ void head_gather_kernel(const float* __restrict__ P, const float* __restrict__ bias,
                                   float* __restrict__ out, int B, int H, int W) {
    const int64_t npix = (int64_t)B * H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const int64_t b = i / ((int64_t)W * H);
    float acc[3] = {__ldg(bias), __ldg(bias + 1), __ldg(bias + 2)};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const float* r = P + ((b * H + yy) * W + xx) * 32 + tap * 3;
        acc[0] += __ldg(r);
        acc[1] += __ldg(r + 1);
        acc[2] += __ldg(r + 2);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[((b * 3 + c) * H + y) * W + x] = tanhf(acc[c]);
}

// dP[b,y,x,tap*3+c] = dpre[b,c,y-dy,x-dx], dpre = dout * (1 - out^2); columns 27..31 zero
__global__ void head_scatter_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                    float* __restrict__ dP, int B, int H, int W) {
    const int64_t npix = (int64_t)B * H * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const int64_t b = i / ((int64_t)W * H);
    float v[32];
#pragma unroll
    for (int j = 27; j < 32; ++j) v[j] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int yy = y - (tap / 3 - 1), xx = x - (tap % 3 - 1);
        const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float d = 0.f;
            if (ok) {
                const int64_t e = ((b * 3 + c) * H + yy) * W + xx;
                const float o = __ldg(out + e);
                d = __ldg(dout + e) * (1.f - o * o);
            }
            v[tap * 3 + c] = d;
        }
    }
    float4* dst = reinterpret_cast<float4*>(dP + i * 32);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}

}  // namespace dsee

using namespace dsee;

#define LAUNCH_END()              \
    count_launch();               \
    DSEE_CUDA(cudaGetLastError()); \
    return 0

extern "C" int dsee_onehot_from_labels(const int64_t* label, float* out, int B, int L, int H, int W,
                                       int* bad_flag, void* stream) {
    DSEE_CHECK_ARG(label && out && bad_flag && B > 0 && L > 0 && H > 0 && W > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    int64_t n = (int64_t)B * H * W;
    onehot_from_labels_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(label, out, B, L, H * W,
                                                                              bad_flag);
    LAUNCH_END();
}

extern "C" int dsee_labels_from_onehot(const float* onehot, uint8_t* labels, int B, int L, int H,
                                       int W, int* bad_flag, void* stream) {
    DSEE_CHECK_ARG(onehot && labels && bad_flag && B > 0 && L > 0 && L <= 255 && H > 0 && W > 0,
                   "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    int64_t n = (int64_t)B * H * W;
    labels_from_onehot_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(onehot, labels, B, L,
                                                                              H * W, bad_flag);
    LAUNCH_END();
}

extern "C" int dsee_labels_u8(const int64_t* label, uint8_t* out, int64_t n, int L, int* bad_flag,
                              void* stream) {
    DSEE_CHECK_ARG(label && out && bad_flag && n > 0 && L > 0 && L <= 255, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    labels_u8_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(label, out, n, L, bad_flag);
    LAUNCH_END();
}

extern "C" int dsee_bicubic_clamp(const float* in, float* out, int B, int C, int Hi, int Wi, int Ho, int Wo,
                                  void* stream) {
    DSEE_CHECK_ARG(in && out && B > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    const int64_t n = (int64_t)B * C * Ho * Wo;
    bicubic_clamp_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
        in, out, B * C, Hi, Wi, Ho, Wo, (float)Hi / (float)Ho, (float)Wi / (float)Wo);
    LAUNCH_END();
}

extern "C" int dsee_resize_labels(const uint8_t* in, uint8_t* out, int B, int Hin, int Win, int Hout,
                                  int Wout, void* stream) {
    DSEE_CHECK_ARG(in && out && B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    int64_t n = (int64_t)B * Hout * Wout;
    resize_labels_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
        in, out, B, Hin, Win, Hout, Wout, (float)Hin / (float)Hout, (float)Win / (float)Wout);
    LAUNCH_END();
}

extern "C" int dsee_shared_mlp_fwd(const uint8_t* labels, const float* table, const float* bias,
                                   void* out_hi, void* out_lo, int B, int Hl, int Wl, int ups, int L,
                                   int nh, float* uniform_rows, void* stream) {
    DSEE_CHECK_ARG(labels && table && bias && out_hi, "NULL pointer");
    DSEE_CHECK_ARG(B > 0 && Hl > 0 && Wl > 0 && (ups == 0 || ups == 1) && L > 0 && L < 255 && nh % 4 == 0 &&
                       nh <= 4096,
                   "bad shape");
    int rc = require_sm100();
    if (rc) return rc;
    int64_t n = (int64_t)B * (Hl << ups) * (Wl << ups) * (nh / 4);
    DSEE_CHECK_ARG(n < ((int64_t)1 << 31), "more than 2^31 output quads");
    cudaStream_t st = (cudaStream_t)stream;
    if (uniform_rows) {
        shared_mlp_rows_kernel<<<L, (nh / 4 + 31) / 32 * 32, 0, st>>>(table, bias, uniform_rows, L, nh);
        count_launch();
    }
    if ((Wl << ups) % 4 == 0)
        shared_mlp_x4_kernel<<<cdiv(n / 4, 256), 256, 0, st>>>(labels, table, bias, uniform_rows, (__half*)out_hi,
                                                              (__half*)out_lo, B, Hl, Wl, ups, L, nh);
    else
        shared_mlp_kernel<<<cdiv(n, 256), 256, 0, st>>>(labels, table, bias, uniform_rows, (__half*)out_hi,
                                                       (__half*)out_lo, B, Hl, Wl, ups, L, nh);
    LAUNCH_END();
}

extern "C" int dsee_style_gather_fwd(const uint8_t* labels, const float* style, void* out_hi,
                                     void* out_lo, int B, int H, int W, int L, int d, void* stream) {
    DSEE_CHECK_ARG(labels && style && out_hi && B > 0 && H > 0 && W > 0 && L > 0 && d % 4 == 0,
                   "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    int64_t n = (int64_t)B * H * W * (d / 4);
    DSEE_CHECK_ARG(n < ((int64_t)1 << 31), "more than 2^31 output quads");
    style_gather_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
        labels, style, (__half*)out_hi, (__half*)out_lo, B, H * W, L, d);
    LAUNCH_END();
}

extern "C" int dsee_prep_conv_weight(const float* w, void* out_hi, void* out_lo, float* inv_scale,
                                     int N, int C, int transpose, void* stream) {
    DSEE_CHECK_ARG(w && out_hi && inv_scale && N > 0 && C > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t n = (int64_t)N * C * 9;
    DSEE_CHECK_ARG(n < ((int64_t)1 << 31), "weight tensor too large");
    DSEE_CUDA(cudaMemsetAsync(inv_scale, 0, 3 * sizeof(float), st));
    int blocks = cdiv(n, 256 * 8);
    if (blocks > 1024) blocks = 1024;
    amax_kernel<<<blocks, 256, 0, st>>>(w, n, inv_scale + 1);
    count_launch();
    if (!transpose && C % 2 == 0 && (size_t)C * 9 * sizeof(float) <= 48 * 1024)
        prep_weight_rows_kernel<<<N, 256, (size_t)C * 9 * sizeof(float), st>>>(w, (__half*)out_hi, (__half*)out_lo,
                                                                              inv_scale, C);
    else
        prep_weight_kernel<<<cdiv(n, 256), 256, 0, st>>>(w, (__half*)out_hi, (__half*)out_lo, inv_scale,
                                                         N, C, transpose);
    if (transpose) {
        count_launch();
        DSEE_CUDA(cudaGetLastError());
        row_l1max_kernel<<<C, 256, 0, st>>>((const __half*)out_hi, 9 * N, inv_scale);
    }
    LAUNCH_END();
}

extern "C" int dsee_prep_conv_weight_f8(const float* w, const float* inv_scale, void* out8, int N, int C,
                                        void* stream) {
    DSEE_CHECK_ARG(w && inv_scale && out8 && N > 0 && C > 0 && C % 128 == 0,
                   "bad argument (C must be a multiple of 128)");
    int rc = require_sm100();
    if (rc) return rc;
    const int64_t n = (int64_t)N * C * 9;
    DSEE_CHECK_ARG(n < ((int64_t)1 << 31), "weight tensor too large");
    prep_weight_f8_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w, inv_scale, (uint8_t*)out8, N, C);
    LAUNCH_END();
}

extern "C" int dsee_prep_mod_weight_batched(const float* wa, const float* ws, void* out_hi, void* out_lo,
                                            float* inv_scale, int B, int N, int Ca, int Ls, int Lp,
                                            void* stream) {
    DSEE_CHECK_ARG(wa && ws && out_hi && inv_scale, "NULL pointer");
    DSEE_CHECK_ARG(B > 0 && N > 0 && Ca > 0 && Ca % 64 == 0 && Ls > 0 && Ls <= Lp && Lp % 64 == 0,
                   "bad shape (Ca, Lp multiples of 64; Ls <= Lp)");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)B * N * 9 * (Ca + Lp);
    DSEE_CHECK_ARG(total < ((int64_t)1 << 31), "weight tensor too large");
    DSEE_CUDA(cudaMemsetAsync(inv_scale, 0, 2 * sizeof(float), st));
    const int64_t na = (int64_t)N * Ca * 9, ns = (int64_t)B * N * Ls * 9;
    int blocks = cdiv(na, 256 * 8);
    amax_kernel<<<blocks > 1024 ? 1024 : blocks, 256, 0, st>>>(wa, na, inv_scale + 1);
    count_launch();
    blocks = cdiv(ns, 256 * 8);
    amax_kernel<<<blocks > 1024 ? 1024 : blocks, 256, 0, st>>>(ws, ns, inv_scale + 1);
    count_launch();
    prep_mod_weight_batched_kernel<<<B * N, 256, (size_t)(Ca + Ls) * 9 * sizeof(float), st>>>(
        wa, ws, (__half*)out_hi, (__half*)out_lo, inv_scale, B, N, Ca, Ls, Lp);
    LAUNCH_END();
}

extern "C" int dsee_prep_conv_weight_ex(const float* w, void* out_hi, void* out_lo, float* inv_scale,
                                        int N, int C, int KH, int KW, int transpose, void* stream) {
    DSEE_CHECK_ARG(w && out_hi && inv_scale && N > 0 && C > 0 && KH > 0 && KW > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = KH * KW;
    const int64_t n = (int64_t)N * C * T;
    DSEE_CUDA(cudaMemsetAsync(inv_scale, 0, 2 * sizeof(float), st));
    int blocks = cdiv(n, 256 * 8);
    if (blocks > 1024) blocks = 1024;
    amax_kernel<<<blocks, 256, 0, st>>>(w, n, inv_scale + 1);
    count_launch();
    const int rows = transpose ? C : N, cols = transpose ? N : C;
    const int64_t total = (int64_t)rows * T * ((cols + 63) / 64 * 64);
    DSEE_CHECK_ARG(total < ((int64_t)1 << 31), "weight tensor too large");
    prep_weight_ex_kernel<<<cdiv(total, 256), 256, 0, st>>>(w, (__half*)out_hi, (__half*)out_lo,
                                                            inv_scale, N, C, T, transpose);
    LAUNCH_END();
}

extern "C" int dsee_split_f16(const float* in, void* out_hi, void* out_lo, int64_t n, void* stream) {
    DSEE_CHECK_ARG(in && out_hi && n > 0 && n % 4 == 0, "bad argument (n must be a multiple of 4)");
    int rc = require_sm100();
    if (rc) return rc;
    split_kernel<<<cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(in, (__half*)out_hi,
                                                                     (__half*)out_lo, n / 4);
    LAUNCH_END();
}

extern "C" int dsee_split_f16_ups2(const float* in, void* out_hi, void* out_lo, int B, int H, int W,
                                   int C, void* stream) {
    DSEE_CHECK_ARG(in && out_hi && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    const int64_t n = (int64_t)B * H * W * (C / 4);
    split_ups_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(in, (__half*)out_hi, (__half*)out_lo,
                                                                     B, H, W, C / 4);
    LAUNCH_END();
}

extern "C" int dsee_fold2x2(const float* in, float* out, int B, int Ho, int Wo, int C, void* stream) {
    DSEE_CHECK_ARG(in && out && B > 0 && Ho > 0 && Wo > 0 && C % 4 == 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    const int64_t n = (int64_t)B * Ho * Wo * (C / 4);
    fold2x2_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, Ho, Wo, C / 4);
    LAUNCH_END();
}

extern "C" int dsee_noise_fill(unsigned long long seed, float* out, int64_t n, void* stream) {
    DSEE_CHECK_ARG(out && n > 0 && n % 4 == 0, "bad argument (n must be a multiple of 4)");
    int rc = require_sm100();
    if (rc) return rc;
    noise_fill_kernel<<<cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(seed, out, n / 4, noise_epoch_ptr());
    LAUNCH_END();
}

extern "C" int dsee_bn_stats(const float* x, int x_ups, const float* noise, unsigned long long noise_seed,
                             const float* noise_w, int B, int H, int W, int C, float* stats_partial,
                             int* n_partials, void* stream) {
    DSEE_CHECK_ARG(n_partials != nullptr, "n_partials is NULL");
    int64_t npix = (int64_t)B * H * W;
    *n_partials = cdiv(npix, STAT_PIX);
    if (!stats_partial) return 0;  // size query
    DSEE_CHECK_ARG(x && C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0,
                   "C must divide 1024 (got %d)", C);
    DSEE_CHECK_ARG((noise != nullptr || noise_seed != 0) == (noise_w != nullptr), "noise/noise_w mismatch");
    DSEE_CHECK_ARG(npix < ((int64_t)1 << 31), "more than 2^31 pixels");
    int rc = require_sm100();
    if (rc) return rc;
    int pix_lanes = 256 / (C / 4);
    size_t sm = (size_t)pix_lanes * C * 2 * sizeof(float);
    auto kern = noise_w ? bn_stats_kernel<true> : bn_stats_kernel<false>;
    kern<<<*n_partials, 256, sm, (cudaStream_t)stream>>>(x, x_ups, noise, noise_seed, noise_w,
                                                                    B, H, W, C, stats_partial, noise_epoch_ptr());
    LAUNCH_END();
}

extern "C" int dsee_bn_finalize(const float* stats_partial, int n_partials, int C, double count,
                                double unbias_count, float eps, float momentum, float* running_mean,
                                float* running_var, float* bn_scale, float* bn_shift, float* mean_out,
                                float* var_out, void* stream) {
    if (unbias_count <= 0) unbias_count = count;
    DSEE_CHECK_ARG(stats_partial && n_partials > 0 && C > 0 && count > 0 && bn_scale && bn_shift,
                   "bad argument");
    DSEE_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "running stats mismatch");
    int rc = require_sm100();
    if (rc) return rc;
    bn_finalize_kernel<<<cdiv(C, 32), 1024, 0, (cudaStream_t)stream>>>(
        stats_partial, n_partials, C, count, unbias_count, eps, momentum, running_mean, running_var,
        bn_scale, bn_shift, mean_out, var_out);
    LAUNCH_END();
}

extern "C" int dsee_bn_eval_affine(const float* running_mean, const float* running_var, float eps,
                                   int C, float* bn_scale, float* bn_shift, void* stream) {
    DSEE_CHECK_ARG(running_mean && running_var && bn_scale && bn_shift && C > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    bn_eval_affine_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(running_mean, running_var,
                                                                          eps, C, bn_scale, bn_shift);
    LAUNCH_END();
}

extern "C" int dsee_stem_fwd(const float* x, const float* w, const float* bias, float* out, int B,
                             int H, int W, int C, void* stream) {
    DSEE_CHECK_ARG(x && w && bias && out && B > 0 && H > 0 && W > 0 && C > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    stem_kernel<<<B * H * W, 128, 0, (cudaStream_t)stream>>>(x, w, bias, out, B, H, W, C);
    LAUNCH_END();
}

extern "C" int dsee_head_gather_fwd(const float* P, const float* bias, float* out, int B, int H, int W,
                                    void* stream) {
    DSEE_CHECK_ARG(P && bias && out && B > 0 && H > 0 && W > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    const int64_t npix = (int64_t)B * H * W;
    head_gather_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(P, bias, out, B, H, W);
    LAUNCH_END();
}

extern "C" int dsee_head_scatter_bwd(const float* dout, const float* out, float* dP, int B, int H, int W,
                                     void* stream) {
    DSEE_CHECK_ARG(dout && out && dP && B > 0 && H > 0 && W > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    const int64_t npix = (int64_t)B * H * W;
    head_scatter_kernel<<<(unsigned)((npix + 127) / 128), 128, 0, (cudaStream_t)stream>>>(dout, out, dP, B, H, W);
    LAUNCH_END();
}

extern "C" int dsee_head_fwd(const float* x, const float* w, const float* bias, float* out, int B,
                             int H, int W, int C, void* stream) {
    DSEE_CHECK_ARG(x && w && bias && out && B > 0 && H > 0 && W > 0, "bad argument");
    DSEE_CHECK_ARG(C > 0 && C % HT_CS == 0, "C must be a multiple of %d (got %d)", HT_CS, C);
    int rc = require_sm100();
    if (rc) return rc;
    static bool configured[64] = {false};
    int dev = 0;
    DSEE_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        DSEE_CUDA(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM));
        configured[dev] = true;
    }
    const int64_t tiles = (int64_t)B * ((H + HT_H - 1) / HT_H) * ((W + HT_W - 1) / HT_W);
    head_kernel<<<(unsigned)tiles, 256, HT_SMEM, (cudaStream_t)stream>>>(x, w, bias, out, B, H, W, C);
    LAUNCH_END();
}
