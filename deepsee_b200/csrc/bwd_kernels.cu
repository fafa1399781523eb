// HBM-bound backward helpers of the generator path: gradient splitting + per-channel sums,
// batch-norm backward (with the folded 2x upsample transposed into a 2x2 sum), the scatter of the
// mlp_shared table gradient, stem / image-head backward.  Deterministic: every reduction goes
// through per-block partials that a fixed-order kernel sums.
#include "common.cuh"
#include "launch_count.h"
#include "../../include/deepsee_b200.h"

#include <cuda_bf16.h>

namespace dsee {

static inline int cdivb(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------
// dY (fp32 NHWC) -> fp16 split planes scaled by 2^e (max|dY| * 2^e in [2^13, 2^14)),
// + per-channel partial sums: sum dY, sum dY*noise_i
// block = 256 threads = (C/4 channel quads) x pixel lanes, GP_PIX pixels per block
// ------------------------------------------------------------------------------------------------
constexpr int GP_PIX = 256;
__global__ void grad_amax_kernel(const float* __restrict__ x, int64_t n4, float* __restrict__ amax) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f && !isinf(m) && !isnan(m)) atomic_max_nonneg(amax, m);
}

__global__ void grad_prep_kernel(const float* __restrict__ dy, __half* __restrict__ hi,
                                 __half* __restrict__ lo, float* __restrict__ inv_scale,
                                 const float* __restrict__ n0, const float* __restrict__ n1,
                                 int64_t npix, int C, float* __restrict__ partial, int nq) {
    extern __shared__ float red[];  // [lanes][C][nq]
    const float scale = pow2_scale_for(inv_scale[1], 14);
    if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[0] = 1.f / scale;
    const int cg = C >> 2;
    const int lanes = blockDim.x / cg;
    const int g = threadIdx.x % cg, pl = threadIdx.x / cg;
    const int64_t p0 = (int64_t)blockIdx.x * GP_PIX;
    float s[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    if (pl < lanes) {
        for (int i = pl; i < GP_PIX; i += lanes) {
            const int64_t pix = p0 + i;
            if (pix >= npix) break;
            const float4 v = __ldg(reinterpret_cast<const float4*>(dy + (size_t)pix * C) + g);
            const float a[4] = {v.x, v.y, v.z, v.w};
            uint32_t ph[2], plw[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float s0 = fminf(fmaxf(a[2 * e] * scale, -65504.f), 65504.f);
                const float s1 = fminf(fmaxf(a[2 * e + 1] * scale, -65504.f), 65504.f);
                const __half h0 = __float2half_rn(s0), h1 = __float2half_rn(s1);
                const __half l0 = __float2half_rn(s0 - __half2float(h0));
                const __half l1 = __float2half_rn(s1 - __half2float(h1));
                ph[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                plw[e] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            *reinterpret_cast<uint2*>(hi + (size_t)pix * C + g * 4) = make_uint2(ph[0], ph[1]);
            if (lo) *reinterpret_cast<uint2*>(lo + (size_t)pix * C + g * 4) = make_uint2(plw[0], plw[1]);
#pragma unroll
            for (int e = 0; e < 4; ++e) s[0][e] += a[e];
            if (n0) {
                const float4 nv = __ldg(reinterpret_cast<const float4*>(n0 + (size_t)pix * C) + g);
                s[1][0] += a[0] * nv.x; s[1][1] += a[1] * nv.y; s[1][2] += a[2] * nv.z; s[1][3] += a[3] * nv.w;
            }
            if (n1) {
                const float4 nv = __ldg(reinterpret_cast<const float4*>(n1 + (size_t)pix * C) + g);
                s[2][0] += a[0] * nv.x; s[2][1] += a[1] * nv.y; s[2][2] += a[2] * nv.z; s[2][3] += a[3] * nv.w;
            }
        }
        for (int k = 0; k < nq; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) red[((size_t)pl * C + g * 4 + e) * nq + k] = s[k][e];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * nq; i += blockDim.x) {
        float a = 0.f;
        for (int l = 0; l < lanes; ++l) a += red[(size_t)l * C * nq + i];
        partial[(size_t)blockIdx.x * C * nq + i] = a;
    }
}

// out[k][c] = sum_s partial[s][c][k] (double, fixed order). grid = C/32, block = 32 x 8.
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int n, int C, int nq,
                                       float scale, float* __restrict__ out) {
    __shared__ double sh[8][32];
    const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    for (int k = 0; k < nq; ++k) {
        double a = 0.0;
        if (c < C)
            for (int s = g; s < n; s += 8) a += (double)partial[((size_t)s * C + c) * nq + k];
        sh[g][cl] = a;
        __syncthreads();
        if (g == 0 && c < C) {
            double t = 0.0;
            for (int j = 0; j < 8; ++j) t += sh[j][cl];
            out[(size_t)k * C + c] = (float)(t * scale);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// batch-norm backward + transpose of the folded upsample + noise-weight gradient partials
//   dxin = bn_scale * (dxhat - m1 - xhat * m2),  m1 = sum(dxhat)/n, m2 = sum(dxhat*xhat)/n
//   dx[b,y',x',c] = sum over the 2^ups x 2^ups full-resolution pixels of (dxin + dskip)
// thread = (low-res pixel, 4 channels); block writes sum(dxin * noise) partials.
// ------------------------------------------------------------------------------------------------
constexpr int BB_PIX = 64;  // low-res pixels per block
__global__ void bn_bwd_kernel(const float* __restrict__ dxhat, const float* __restrict__ x, int ups,
                              const float* __restrict__ noise, const float* __restrict__ noise_w,
                              const float* __restrict__ sc, const float* __restrict__ sh,
                              const float* __restrict__ sums /*[2][C]: sum dxhat, sum dxhat*xhat*/,
                              float inv_count, const float* __restrict__ dskip, int B, int Hx,
                              int Wx, int C, float* __restrict__ dx,
                              float* __restrict__ nw_partial) {
    extern __shared__ float red[];  // [lanes][C]
    const int cg = C >> 2;
    const int lanes = blockDim.x / cg;
    const int g = threadIdx.x % cg, pl = threadIdx.x / cg;
    const int64_t npix = (int64_t)B * Hx * Wx;
    const int64_t p0 = (int64_t)blockIdx.x * BB_PIX;
    const int H = Hx << ups, W = Wx << ups;
    const int f = 1 << ups;
    float nacc[4] = {0, 0, 0, 0};
    if (pl < lanes) {
        const float4 scv = __ldg(reinterpret_cast<const float4*>(sc) + g);
        const float4 shv = __ldg(reinterpret_cast<const float4*>(sh) + g);
        float4 m1 = __ldg(reinterpret_cast<const float4*>(sums) + g);
        float4 m2 = __ldg(reinterpret_cast<const float4*>(sums + C) + g);
        m1.x *= inv_count; m1.y *= inv_count; m1.z *= inv_count; m1.w *= inv_count;
        m2.x *= inv_count; m2.y *= inv_count; m2.z *= inv_count; m2.w *= inv_count;
        float4 nw = make_float4(0, 0, 0, 0);
        if (noise) nw = __ldg(reinterpret_cast<const float4*>(noise_w) + g);
        for (int i = pl; i < BB_PIX; i += lanes) {
            const int64_t pix = p0 + i;
            if (pix >= npix) break;
            const int xx = (int)(pix % Wx);
            const int yy = (int)((pix / Wx) % Hx);
            const int b = (int)(pix / ((int64_t)Wx * Hx));
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)pix * C) + g);
            float4 acc = make_float4(0, 0, 0, 0);
            for (int sy = 0; sy < f; ++sy)
                for (int sx = 0; sx < f; ++sx) {
                    const size_t fp = ((size_t)b * H + (yy * f + sy)) * W + (xx * f + sx);
                    const float4 d = __ldg(reinterpret_cast<const float4*>(dxhat + fp * C) + g);
                    float4 xi = xv, nv = make_float4(0, 0, 0, 0);
                    if (noise) {
                        nv = __ldg(reinterpret_cast<const float4*>(noise + fp * C) + g);
                        xi.x += nw.x * nv.x; xi.y += nw.y * nv.y; xi.z += nw.z * nv.z; xi.w += nw.w * nv.w;
                    }
                    float4 r;
                    r.x = scv.x * (d.x - m1.x - (xi.x * scv.x + shv.x) * m2.x);
                    r.y = scv.y * (d.y - m1.y - (xi.y * scv.y + shv.y) * m2.y);
                    r.z = scv.z * (d.z - m1.z - (xi.z * scv.z + shv.z) * m2.z);
                    r.w = scv.w * (d.w - m1.w - (xi.w * scv.w + shv.w) * m2.w);
                    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
                    if (dskip) {
                        const float4 k = __ldg(reinterpret_cast<const float4*>(dskip + fp * C) + g);
                        acc.x += k.x; acc.y += k.y; acc.z += k.z; acc.w += k.w;
                    }
                    nacc[0] += r.x * nv.x; nacc[1] += r.y * nv.y; nacc[2] += r.z * nv.z; nacc[3] += r.w * nv.w;
                }
            reinterpret_cast<float4*>(dx + (size_t)pix * C)[g] = acc;
        }
        if (nw_partial)
#pragma unroll
            for (int e = 0; e < 4; ++e) red[(size_t)pl * C + g * 4 + e] = nacc[e];
    }
    if (nw_partial) {
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float a = 0.f;
            for (int l = 0; l < lanes; ++l) a += red[(size_t)l * C + c];
            nw_partial[(size_t)blockIdx.x * C + c] = a;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// mlp_shared backward: d table[tap][label(p+tap)][o] += relu'(actv[p,o]) * dactv[p,o], d bias
// grid = blocks over low-res pixels; block = nh threads (thread owns a hidden channel -> private
// smem column, no atomics); partial tables reduced afterwards.
// ------------------------------------------------------------------------------------------------
constexpr int SM_PIX = 512;
__global__ void shared_mlp_bwd_kernel(const float* __restrict__ dsrc, int ld, int coff,
                                      const __half* __restrict__ actv_hi,
                                      const uint8_t* __restrict__ labels, int B, int Hl, int Wl,
                                      int ups, int L, int nh, float* __restrict__ partial) {
    extern __shared__ float tab[];  // [9*L + 1][nh]  (last row: bias)
    const int o = threadIdx.x;
    const int rows = 9 * L + 1;
    for (int r = 0; r < rows; ++r) tab[r * nh + o] = 0.f;
    const int64_t npix = (int64_t)B * Hl * Wl;
    const int64_t p0 = (int64_t)blockIdx.x * SM_PIX;
    const int f = 1 << ups;
    const int H = Hl << ups, W = Wl << ups;
    for (int i = 0; i < SM_PIX; ++i) {
        const int64_t pix = p0 + i;
        if (pix >= npix) break;
        const int xl = (int)(pix % Wl);
        const int yl = (int)((pix / Wl) % Hl);
        const int b = (int)(pix / ((int64_t)Wl * Hl));
        // gradient wrt the low-res activation = sum over its f x f upsampled copies
        const size_t fp0 = ((size_t)b * H + yl * f) * W + xl * f;
        if (!(__half2float(actv_hi[fp0 * nh + o]) > 0.f)) continue;  // ReLU gate (same for all copies)
        float gsum = 0.f;
        for (int sy = 0; sy < f; ++sy)
            for (int sx = 0; sx < f; ++sx)
                gsum += __ldg(dsrc + (fp0 + (size_t)sy * W + sx) * ld + coff + o);
        const uint8_t* lb = labels + (size_t)b * Hl * Wl;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = yl + tap / 3 - 1, xx = xl + tap % 3 - 1;
            if (yy < 0 || yy >= Hl || xx < 0 || xx >= Wl) continue;
            const int l = lb[yy * Wl + xx];
            tab[(tap * L + l) * nh + o] += gsum;
        }
        tab[(9 * L) * nh + o] += gsum;
    }
    float* out = partial + (size_t)blockIdx.x * rows * nh;
    for (int r = 0; r < rows; ++r) out[r * nh + o] = tab[r * nh + o];
}

// ------------------------------------------------------------------------------------------------
// stem backward (weights only; the LR image needs no gradient): block partials over pixels
//   dW[c][27] = sum_p dY[p][c] * patch_p[27],  db[c] = sum_p dY[p][c]
// ------------------------------------------------------------------------------------------------
constexpr int ST_PIX = 64;
__global__ void stem_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B, int H,
                                int W, int C, float* __restrict__ partial /*[blocks][C][28]*/) {
    __shared__ float patch[ST_PIX][27];
    const int64_t npix = (int64_t)B * H * W;
    const int64_t p0 = (int64_t)blockIdx.x * ST_PIX;
    for (int i = threadIdx.x; i < ST_PIX * 27; i += blockDim.x) {
        const int pi = i / 27, k = i % 27;
        const int64_t pix = p0 + pi;
        float v = 0.f;
        if (pix < npix) {
            const int xx = (int)(pix % W), yy = (int)((pix / W) % H), b = (int)(pix / ((int64_t)W * H));
            const int ci = k / 9, tap = k % 9;
            const int y2 = yy + tap / 3 - 1, x2 = xx + tap % 3 - 1;
            if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) v = x[(((size_t)b * 3 + ci) * H + y2) * W + x2];
        }
        patch[pi][k] = v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc[28];
#pragma unroll
        for (int k = 0; k < 28; ++k) acc[k] = 0.f;
        for (int pi = 0; pi < ST_PIX; ++pi) {
            const int64_t pix = p0 + pi;
            if (pix >= npix) break;
            const float d = __ldg(dy + (size_t)pix * C + c);
#pragma unroll
            for (int k = 0; k < 27; ++k) acc[k] += d * patch[pi][k];
            acc[27] += d;
        }
        float* o = partial + ((size_t)blockIdx.x * C + c) * 28;
#pragma unroll
        for (int k = 0; k < 28; ++k) o[k] = acc[k];
    }
}

// ------------------------------------------------------------------------------------------------
// image-head backward.  out = tanh(conv(lrelu(x))):  dpre = dout * (1 - out^2)   [B,3,H,W] NCHW
//   dX[p][c]       = lrelu'(x[p][c]) * sum_{tap,o} dpre[o][p - tap] * w[o][c][tap]
//   dW[o][c][tap]  = sum_p dpre[o][p] * lrelu(x[p + tap][c]),   db[o] = sum_p dpre[o][p]
// ------------------------------------------------------------------------------------------------
__global__ void head_dpre_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                 float* __restrict__ dpre, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float o = out[i];
    dpre[i] = dout[i] * (1.f - o * o);
}

// thread = (pixel, 4 channels); weights [3][C][9] read through L1
__global__ void head_dx_kernel(const float* __restrict__ x, const float* __restrict__ dpre,
                               const float* __restrict__ w, float* __restrict__ dx, int B, int H, int W,
                               int C) {
    const int cg = C >> 2;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)B * H * W * cg) return;
    const int g = (int)(i % cg);
    const int64_t pix = i / cg;
    const int xx = (int)(pix % W), yy = (int)((pix / W) % H), b = (int)(pix / ((int64_t)W * H));
    float acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        // output pixel q = p - (tap offset) used x[p] with weight tap
        const int yq = yy - (tap / 3 - 1), xq = xx - (tap % 3 - 1);
        if (yq < 0 || yq >= H || xq < 0 || xq >= W) continue;
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            const float d = __ldg(dpre + (((size_t)b * 3 + o) * H + yq) * W + xq);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] += d * __ldg(w + ((size_t)o * C + g * 4 + e) * 9 + tap);
        }
    }
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)pix * C) + g);
    float4 r;
    r.x = acc[0] * (xv.x > 0.f ? 1.f : 0.2f);
    r.y = acc[1] * (xv.y > 0.f ? 1.f : 0.2f);
    r.z = acc[2] * (xv.z > 0.f ? 1.f : 0.2f);
    r.w = acc[3] * (xv.w > 0.f ? 1.f : 0.2f);
    reinterpret_cast<float4*>(dx + (size_t)pix * C)[g] = r;
}

// block = C threads over HD_PIX output pixels; partial [blocks][C][27 + (c < 3 ? bias : 0)]
constexpr int HD_PIX = 128;
__global__ void head_dw_kernel(const float* __restrict__ x, const float* __restrict__ dpre, int B, int H,
                               int W, int C, float* __restrict__ partial /*[blocks][C][28]*/) {
    const int c = threadIdx.x;
    const int64_t npix = (int64_t)B * H * W;
    const int64_t p0 = (int64_t)blockIdx.x * HD_PIX;
    float acc[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) acc[k] = 0.f;
    for (int i = 0; i < HD_PIX; ++i) {
        const int64_t pix = p0 + i;
        if (pix >= npix) break;
        const int xx = (int)(pix % W), yy = (int)((pix / W) % H), b = (int)(pix / ((int64_t)W * H));
        float d[3];
#pragma unroll
        for (int o = 0; o < 3; ++o) d[o] = __ldg(dpre + (((size_t)b * 3 + o) * H + yy) * W + xx);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int y2 = yy + tap / 3 - 1, x2 = xx + tap % 3 - 1;
            if (y2 < 0 || y2 >= H || x2 < 0 || x2 >= W) continue;
            float v = __ldg(x + (((size_t)b * H + y2) * W + x2) * C + c);
            v = v > 0.f ? v : 0.2f * v;
#pragma unroll
            for (int o = 0; o < 3; ++o) acc[o * 9 + tap] += d[o] * v;
        }
        if (c < 3) acc[27] += d[c];
    }
    float* o = partial + ((size_t)blockIdx.x * C + c) * 28;
#pragma unroll
    for (int k = 0; k < 28; ++k) o[k] = acc[k];
}

}  // namespace dsee

using namespace dsee;

#define LAUNCH_END()               \
    count_launch();                \
    DSEE_CUDA(cudaGetLastError()); \
    return 0

extern "C" int dsee_grad_prep_blocks(int64_t npix) { return cdivb(npix, GP_PIX); }

extern "C" int dsee_grad_prep(const float* dy, void* out_hi, void* out_lo, float* inv_scale,
                              const float* noise0, const float* noise1, int64_t npix, int C,
                              float* partial, void* stream) {
    DSEE_CHECK_ARG(dy && out_hi && inv_scale && partial && npix > 0, "bad argument");
    DSEE_CHECK_ARG(C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0, "C must divide 1024 (got %d)", C);
    DSEE_CHECK_ARG(!(noise1 && !noise0), "noise1 without noise0");
    int rc = require_sm100();
    if (rc) return rc;
    const int nq = 1 + (noise0 ? 1 : 0) + (noise1 ? 1 : 0);
    const int lanes = 256 / (C / 4);
    size_t sm = (size_t)lanes * C * nq * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    DSEE_CUDA(cudaMemsetAsync(inv_scale, 0, 2 * sizeof(float), st));
    const int64_t n4 = npix * C / 4;
    int ablocks = cdivb(n4, 256 * 8);
    if (ablocks > 148 * 8) ablocks = 148 * 8;
    grad_amax_kernel<<<ablocks, 256, 0, st>>>(dy, n4, inv_scale + 1);
    count_launch();
    grad_prep_kernel<<<cdivb(npix, GP_PIX), 256, sm, st>>>(dy, (__half*)out_hi, (__half*)out_lo,
                                                           inv_scale, noise0, noise1, npix, C, partial,
                                                           nq);
    LAUNCH_END();
}

extern "C" int dsee_reduce_partials(const float* partial, int n, int C, int nq, float scale,
                                    float* out, void* stream) {
    DSEE_CHECK_ARG(partial && out && n > 0 && C > 0 && nq > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    reduce_partials_kernel<<<cdivb(C, 32), 256, 0, (cudaStream_t)stream>>>(partial, n, C, nq, scale, out);
    LAUNCH_END();
}

extern "C" int dsee_bn_bwd_blocks(int B, int Hx, int Wx) { return cdivb((int64_t)B * Hx * Wx, BB_PIX); }

extern "C" int dsee_bn_bwd(const float* dxhat, const float* x, int x_ups, const float* noise,
                           const float* noise_w, const float* bn_scale, const float* bn_shift,
                           const float* sums, float inv_count, const float* dskip, int B, int Hx,
                           int Wx, int C, float* dx, float* nw_partial, void* stream) {
    DSEE_CHECK_ARG(dxhat && x && bn_scale && bn_shift && sums && dx, "NULL pointer");
    DSEE_CHECK_ARG(C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0, "C must divide 1024 (got %d)", C);
    DSEE_CHECK_ARG((noise == nullptr) == (noise_w == nullptr), "noise/noise_w mismatch");
    DSEE_CHECK_ARG(!nw_partial || noise, "nw_partial needs noise");
    int rc = require_sm100();
    if (rc) return rc;
    const int lanes = 256 / (C / 4);
    size_t sm = nw_partial ? (size_t)lanes * C * sizeof(float) : 0;
    bn_bwd_kernel<<<dsee_bn_bwd_blocks(B, Hx, Wx), 256, sm, (cudaStream_t)stream>>>(
        dxhat, x, x_ups, noise, noise_w, bn_scale, bn_shift, sums, inv_count, dskip, B, Hx, Wx, C, dx,
        nw_partial);
    LAUNCH_END();
}

extern "C" int dsee_shared_mlp_bwd_blocks(int B, int Hl, int Wl) {
    return cdivb((int64_t)B * Hl * Wl, SM_PIX);
}

extern "C" int dsee_shared_mlp_bwd(const float* dsrc, int ld, int coff, const void* actv_hi,
                                   const uint8_t* labels, int B, int Hl, int Wl, int ups, int L,
                                   int nh, float* partial, float* dtable_dbias, void* stream) {
    DSEE_CHECK_ARG(dsrc && actv_hi && labels && partial && dtable_dbias, "NULL pointer");
    DSEE_CHECK_ARG(nh > 0 && nh <= 1024 && (size_t)(9 * L + 1) * nh * 4 <= 200 * 1024, "table too large");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    static bool configured[64] = {false};
    int dev = 0;
    DSEE_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        DSEE_CUDA(cudaFuncSetAttribute(shared_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024));
        configured[dev] = true;
    }
    const int rows = 9 * L + 1;
    const int blocks = dsee_shared_mlp_bwd_blocks(B, Hl, Wl);
    shared_mlp_bwd_kernel<<<blocks, nh, (size_t)rows * nh * 4, st>>>(
        dsrc, ld, coff, (const __half*)actv_hi, labels, B, Hl, Wl, ups, L, nh, partial);
    count_launch();
    // [blocks][rows*nh][1] -> [rows*nh]
    reduce_partials_kernel<<<cdivb(rows * nh, 32), 256, 0, st>>>(partial, blocks, rows * nh, 1, 1.f,
                                                                  dtable_dbias);
    LAUNCH_END();
}

extern "C" int dsee_stem_bwd_blocks(int B, int H, int W) { return cdivb((int64_t)B * H * W, ST_PIX); }

extern "C" int dsee_stem_bwd(const float* x, const float* dy, int B, int H, int W, int C,
                             float* partial, float* dw_db, void* stream) {
    DSEE_CHECK_ARG(x && dy && partial && dw_db, "NULL pointer");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = dsee_stem_bwd_blocks(B, H, W);
    stem_bwd_kernel<<<blocks, 128, 0, st>>>(x, dy, B, H, W, C, partial);
    count_launch();
    // partial [blocks][C*28] -> dw_db [C*28]  (per channel: 27 weight grads then the bias grad)
    reduce_partials_kernel<<<cdivb(C * 28, 32), 256, 0, st>>>(partial, blocks, C * 28, 1, 1.f, dw_db);
    LAUNCH_END();
}

extern "C" int dsee_head_bwd_blocks(int B, int H, int W) { return cdivb((int64_t)B * H * W, HD_PIX); }

extern "C" int dsee_head_bwd(const float* x, const float* w, const float* out, const float* dout,
                             int B, int H, int W, int C, float* dpre, float* dx, float* partial,
                             float* dw_db, void* stream) {
    DSEE_CHECK_ARG(x && w && out && dout && dpre && dx && partial && dw_db, "NULL pointer");
    DSEE_CHECK_ARG(C % 4 == 0 && C <= 1024, "C must be a multiple of 4 and <= 1024");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)B * 3 * H * W;
    head_dpre_kernel<<<cdivb(n, 256), 256, 0, st>>>(dout, out, dpre, n);
    count_launch();
    const int64_t nx = (int64_t)B * H * W * (C / 4);
    head_dx_kernel<<<cdivb(nx, 256), 256, 0, st>>>(x, dpre, w, dx, B, H, W, C);
    count_launch();
    const int blocks = dsee_head_bwd_blocks(B, H, W);
    head_dw_kernel<<<blocks, C, 0, st>>>(x, dpre, B, H, W, C, partial);
    count_launch();
    reduce_partials_kernel<<<cdivb(C * 28, 32), 256, 0, st>>>(partial, blocks, C * 28, 1, 1.f, dw_db);
    LAUNCH_END();
}
