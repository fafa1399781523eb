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
                                 unsigned long long seed0, unsigned long long seed1, int64_t npix, int C,
                                 float* __restrict__ partial, int nq, const float* __restrict__ amax_in,
                                 const unsigned long long* epoch) {
    extern __shared__ float red[];  // [lanes][C][nq]
    seed0 = eff_noise_seed(seed0, epoch);
    seed1 = eff_noise_seed(seed1, epoch);
    // max|dY|: from the producer (amax_in) or from grad_amax_kernel (inv_scale[1]); inv_scale[1] is
    // only read by kernels launched after this one when it is written here
    const float amax = amax_in ? __ldg(amax_in) : inv_scale[1];
    const float scale = pow2_scale_for(amax, 14);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        inv_scale[0] = 1.f / scale;
        if (amax_in) inv_scale[1] = amax;
    }
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
            if (nq > 1) {
                const float4 nv = load_noise4(n0, seed0, (size_t)pix * C + g * 4);
                s[1][0] += a[0] * nv.x; s[1][1] += a[1] * nv.y; s[1][2] += a[2] * nv.z; s[1][3] += a[3] * nv.w;
            }
            if (nq > 2) {
                const float4 nv = load_noise4(n1, seed1, (size_t)pix * C + g * 4);
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

// ------------------------------------------------------------------------------------------------
// K1 backward from saved G = gamma + gamma_bias (fp16 planes): with xhat = xin*sc + sh,
//   dxhat = dt * G,  dG = dt * xhat,  dB = dt   (dG|dB as scaled fp16 planes, channels interleaved
//   per 128 like the modulation weight rows) + block partials (sum dxhat, sum dxhat*xhat, sum dG,
//   sum dB).  thread = (pixel lane, 4 channels); GP_PIX pixels per block.
// ------------------------------------------------------------------------------------------------
__global__ void modulate_bwd_saved_kernel(const float* __restrict__ x, int x_ups,
                                          const float* __restrict__ noise, unsigned long long noise_seed,
                                          const float* __restrict__ noise_w,
                                          const float* __restrict__ sc, const float* __restrict__ sh,
                                          const __half* __restrict__ g_hi, const __half* __restrict__ g_lo,
                                          const float* __restrict__ dt, const float* __restrict__ dt_amax,
                                          int B, int H, int W, int C, float* __restrict__ dxhat,
                                          __half* __restrict__ dgb_hi, __half* __restrict__ dgb_lo,
                                          float* __restrict__ dgb_inv_scale, float* __restrict__ partial,
                                          const unsigned long long* epoch) {
    extern __shared__ float red[];  // [lanes][C][4]
    noise_seed = eff_noise_seed(noise_seed, epoch);
    const int cg = C >> 2;
    const int lanes = blockDim.x / cg;
    const int g = threadIdx.x % cg, pl = threadIdx.x / cg;
    const float gscale = pow2_scale_for(__ldg(dt_amax), 10);  // 2^6 headroom for |xhat|
    if (blockIdx.x == 0 && threadIdx.x == 0) *dgb_inv_scale = 1.f / gscale;
    const int64_t npix = (int64_t)B * H * W;
    const int64_t p0 = (int64_t)blockIdx.x * GP_PIX;
    const int Hx = H >> x_ups, Wx = W >> x_ups;
    float s[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[k][e] = 0.f;
    if (pl < lanes) {
        const float4 scv = __ldg(reinterpret_cast<const float4*>(sc) + g);
        const float4 shv = __ldg(reinterpret_cast<const float4*>(sh) + g);
        float4 nw = make_float4(0, 0, 0, 0);
        const bool has_noise = noise_w != nullptr;
        if (has_noise) nw = __ldg(reinterpret_cast<const float4*>(noise_w) + g);
        const int c = g * 4;
        const int ng = (c >> 7) * 256 + (c & 127);  // interleaved position of channel c
        for (int i = pl; i < GP_PIX; i += lanes) {
            const int64_t pix = p0 + i;
            if (pix >= npix) break;
            size_t xp = (size_t)pix;
            if (x_ups) {  // 32-bit coordinates (npix < 2^31 is checked by the host wrapper)
                const uint32_t pu = (uint32_t)pix;
                const uint32_t xx = pu % (uint32_t)W, t2 = pu / (uint32_t)W;
                const uint32_t yy = t2 % (uint32_t)H, b = t2 / (uint32_t)H;
                xp = ((size_t)b * Hx + (yy >> 1)) * Wx + (xx >> 1);
            }
            float4 xv = __ldg(reinterpret_cast<const float4*>(x + xp * C) + g);
            if (has_noise) {
                const float4 nv = load_noise4(noise, noise_seed, (size_t)pix * C + g * 4);
                xv.x += nw.x * nv.x; xv.y += nw.y * nv.y; xv.z += nw.z * nv.z; xv.w += nw.w * nv.w;
            }
            const float xh[4] = {xv.x * scv.x + shv.x, xv.y * scv.y + shv.y, xv.z * scv.z + shv.z,
                                 xv.w * scv.w + shv.w};
            const float4 dv = __ldg(reinterpret_cast<const float4*>(dt + (size_t)pix * C) + g);
            const float d[4] = {dv.x, dv.y, dv.z, dv.w};
            const uint2 gh = __ldg(reinterpret_cast<const uint2*>(g_hi + (size_t)pix * C) + g);
            float G[4];
            G[0] = __half2float(__ushort_as_half((unsigned short)(gh.x & 0xffffu)));
            G[1] = __half2float(__ushort_as_half((unsigned short)(gh.x >> 16)));
            G[2] = __half2float(__ushort_as_half((unsigned short)(gh.y & 0xffffu)));
            G[3] = __half2float(__ushort_as_half((unsigned short)(gh.y >> 16)));
            if (g_lo) {
                const uint2 gl = __ldg(reinterpret_cast<const uint2*>(g_lo + (size_t)pix * C) + g);
                G[0] += __half2float(__ushort_as_half((unsigned short)(gl.x & 0xffffu)));
                G[1] += __half2float(__ushort_as_half((unsigned short)(gl.x >> 16)));
                G[2] += __half2float(__ushort_as_half((unsigned short)(gl.y & 0xffffu)));
                G[3] += __half2float(__ushort_as_half((unsigned short)(gl.y >> 16)));
            }
            float dxh[4], dG[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                dxh[e] = d[e] * G[e];
                dG[e] = d[e] * xh[e];
                s[0][e] += dxh[e];
                s[1][e] += dxh[e] * xh[e];
                s[2][e] += dG[e];
                s[3][e] += d[e];
            }
            reinterpret_cast<float4*>(dxhat + (size_t)pix * C)[g] = make_float4(dxh[0], dxh[1], dxh[2], dxh[3]);
            __half* rowh = dgb_hi + (size_t)pix * 2 * C + ng;
            __half* rowl = dgb_lo ? dgb_lo + (size_t)pix * 2 * C + ng : nullptr;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const float* src = half ? d : dG;
                uint32_t ph[2], plw[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float s0 = fminf(fmaxf(src[2 * e] * gscale, -65504.f), 65504.f);
                    const float s1 = fminf(fmaxf(src[2 * e + 1] * gscale, -65504.f), 65504.f);
                    const __half h0 = __float2half_rn(s0), h1 = __float2half_rn(s1);
                    const __half l0 = __float2half_rn(s0 - __half2float(h0));
                    const __half l1 = __float2half_rn(s1 - __half2float(h1));
                    ph[e] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                    plw[e] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                }
                *reinterpret_cast<uint2*>(rowh + half * 128) = make_uint2(ph[0], ph[1]);
                if (rowl) *reinterpret_cast<uint2*>(rowl + half * 128) = make_uint2(plw[0], plw[1]);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) red[((size_t)pl * C + g * 4 + e) * 4 + k] = s[k][e];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 4; i += blockDim.x) {
        float a = 0.f;
        for (int l = 0; l < lanes; ++l) a += red[(size_t)l * C * 4 + i];
        partial[(size_t)blockIdx.x * C * 4 + i] = a;
    }
}

// out[k][c] = sum_s partial[s][c][k] (double, fixed order).  Block = 8 channels x 128 row lanes: a warp
// reads 4 rows x (8 channels x NQ values) = whole 32 B x NQ segments with one vector load per thread,
// folds its 4 row lanes by shuffles, and the 32 warp partials are summed in warp order.
template <int NQ>
__global__ void __launch_bounds__(1024)
reduce_partials_kernel(const float* __restrict__ partial, int n, int C, float scale, float* __restrict__ out) {
    __shared__ double sh[32][8 * NQ];
    const int cl = threadIdx.x & 7, g = threadIdx.x >> 3;
    const int c = blockIdx.x * 8 + cl;
    double a[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) a[k] = 0.0;
    if (c < C) {
#pragma unroll 4
        for (int s = g; s < n; s += 128) {
            const float* p = partial + ((size_t)s * C + c) * NQ;
            float v[NQ];
            if (NQ == 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p));
                v[0] = t.x; v[1 % NQ] = t.y; v[2 % NQ] = t.z; v[3 % NQ] = t.w;
            } else if (NQ == 2) {
                const float2 t = __ldg(reinterpret_cast<const float2*>(p));
                v[0] = t.x; v[1 % NQ] = t.y;
            } else {
#pragma unroll
                for (int k = 0; k < NQ; ++k) v[k] = __ldg(p + k);
            }
#pragma unroll
            for (int k = 0; k < NQ; ++k) a[k] += (double)v[k];
        }
    }
#pragma unroll
    for (int k = 0; k < NQ; ++k) {   // the warp's 4 row lanes (lane bits 3, 4)
        a[k] += __shfl_xor_sync(0xffffffffu, a[k], 8);
        a[k] += __shfl_xor_sync(0xffffffffu, a[k], 16);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 8)
#pragma unroll
        for (int k = 0; k < NQ; ++k) sh[warp][lane * NQ + k] = a[k];
    __syncthreads();
    if (threadIdx.x < 8 * NQ) {
        const int cc = blockIdx.x * 8 + threadIdx.x / NQ, k = threadIdx.x % NQ;
        if (cc < C) {
            double t = 0.0;
            for (int j = 0; j < 32; ++j) t += sh[j][threadIdx.x];
            out[(size_t)k * C + cc] = (float)(t * scale);
        }
    }
}

static void launch_reduce_partials(const float* partial, int n, int C, int nq, float scale, float* out,
                                   cudaStream_t st) {
    const int blocks = (C + 7) / 8;
    switch (nq) {
        case 1: reduce_partials_kernel<1><<<blocks, 1024, 0, st>>>(partial, n, C, scale, out); break;
        case 2: reduce_partials_kernel<2><<<blocks, 1024, 0, st>>>(partial, n, C, scale, out); break;
        case 3: reduce_partials_kernel<3><<<blocks, 1024, 0, st>>>(partial, n, C, scale, out); break;
        default: reduce_partials_kernel<4><<<blocks, 1024, 0, st>>>(partial, n, C, scale, out); break;
    }
}

// ------------------------------------------------------------------------------------------------
// batch-norm backward + transpose of the folded upsample + noise-weight gradient partials
//   dxin = bn_scale * (dxhat - m1 - xhat * m2),  m1 = sum(dxhat)/n, m2 = sum(dxhat*xhat)/n
//   dx[b,y',x',c] = sum over the 2^ups x 2^ups full-resolution pixels of (dxin + dskip)
// thread = (low-res pixel, 4 channels); block writes sum(dxin * noise) partials.
// ------------------------------------------------------------------------------------------------
constexpr int BB_PIX = 64;  // low-res pixels per block
template <bool HAS_NOISE>
__global__ void __launch_bounds__(256, HAS_NOISE ? 3 : 4) bn_bwd_kernel(const float* __restrict__ dxhat, const float* __restrict__ x, int ups,
                              const float* __restrict__ noise, unsigned long long noise_seed,
                              const float* __restrict__ noise_w,
                              const float* __restrict__ sc, const float* __restrict__ sh,
                              const float* __restrict__ sums /*[2][C]: sum dxhat, sum dxhat*xhat*/,
                              float inv_count, const float* __restrict__ dskip, int B, int Hx,
                              int Wx, int C, float* __restrict__ dx,
                              float* __restrict__ nw_partial, int nw_with_skip,
                              float* __restrict__ amax_out, const unsigned long long* epoch) {
    extern __shared__ float red[];  // [lanes][C]
    if (HAS_NOISE) noise_seed = eff_noise_seed(noise_seed, epoch);
    const int cg = C >> 2;
    const int lanes = blockDim.x / cg;
    const int g = threadIdx.x % cg, pl = threadIdx.x / cg;
    const int64_t npix = (int64_t)B * Hx * Wx;
    const int64_t p0 = (int64_t)blockIdx.x * BB_PIX;
    const int H = Hx << ups, W = Wx << ups;
    const int f = 1 << ups;
    float nacc[4] = {0, 0, 0, 0};
    float tmax = 0.f;
    if (pl < lanes) {
        const float4 scv = __ldg(reinterpret_cast<const float4*>(sc) + g);
        const float4 shv = __ldg(reinterpret_cast<const float4*>(sh) + g);
        float4 m1 = __ldg(reinterpret_cast<const float4*>(sums) + g);
        float4 m2 = __ldg(reinterpret_cast<const float4*>(sums + C) + g);
        m1.x *= inv_count; m1.y *= inv_count; m1.z *= inv_count; m1.w *= inv_count;
        m2.x *= inv_count; m2.y *= inv_count; m2.z *= inv_count; m2.w *= inv_count;
        float4 nw = make_float4(0, 0, 0, 0);
        constexpr bool has_noise = HAS_NOISE;
        if (has_noise) nw = __ldg(reinterpret_cast<const float4*>(noise_w) + g);
        const int iend = (int)(npix - p0 < BB_PIX ? npix - p0 : BB_PIX);
        // (b, yy, xx) of this lane's first pixel by 32-bit division, then advanced incrementally:
        // 64-bit div/mod per pixel cost more issue slots than the memory traffic they index
        const uint32_t pix0 = (uint32_t)(p0 + pl);
        int xx = (int)(pix0 % (uint32_t)Wx);
        int yy = (int)((pix0 / (uint32_t)Wx) % (uint32_t)Hx);
        int b = (int)(pix0 / ((uint32_t)Wx * (uint32_t)Hx));
        xx -= lanes;
#pragma unroll(HAS_NOISE ? 1 : 2)
        for (int i = pl; i < iend; i += lanes) {
            const int64_t pix = p0 + i;
            xx += lanes;
            while (xx >= Wx) {
                xx -= Wx;
                if (++yy == Hx) {
                    yy = 0;
                    ++b;
                }
            }
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)pix * C) + g);
            float4 acc = make_float4(0, 0, 0, 0);
            for (int sy = 0; sy < f; ++sy)
                for (int sx = 0; sx < f; ++sx) {
                    const size_t fp = ((size_t)b * H + (yy * f + sy)) * W + (xx * f + sx);
                    const float4 d = __ldg(reinterpret_cast<const float4*>(dxhat + fp * C) + g);
                    float4 xi = xv, nv = make_float4(0, 0, 0, 0);
                    if (has_noise) {
                        nv = load_noise4(noise, noise_seed, fp * C + g * 4);
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
                        if (nw_with_skip) { r.x += k.x; r.y += k.y; r.z += k.z; r.w += k.w; }
                    }
                    nacc[0] += r.x * nv.x; nacc[1] += r.y * nv.y; nacc[2] += r.z * nv.z; nacc[3] += r.w * nv.w;
                }
            reinterpret_cast<float4*>(dx + (size_t)pix * C)[g] = acc;
            tmax = fmaxf(tmax, fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w))));
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
    if (amax_out) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
        if ((threadIdx.x & 31) == 0 && tmax > 0.f && !isinf(tmax) && !isnan(tmax)) atomic_max_nonneg(amax_out, tmax);
    }
}

// ------------------------------------------------------------------------------------------------
// mlp_shared backward on the tensor cores.  actv[p,o] = relu(bias[o] + sum_tap table[tap][label(p+tap)][o])
// is a 3x3 conv over the one-hot map, so d table = wgrad(G, onehot) with
//   G[p,o] = relu'(actv[p,o]) * (sum over the 2^ups x 2^ups upsampled copies of d actv)
// This kernel builds G as scaled fp16 planes (+ block partials of sum G = the bias gradient);
// onehot_planes_kernel builds the other operand; dsee_conv3x3_wgrad does the reduction.
// ------------------------------------------------------------------------------------------------
__global__ void actv_grad_prep_kernel(const float* __restrict__ dsrc, int ld, int coff,
                                      const __half* __restrict__ actv_hi, const float* __restrict__ amax,
                                      int B, int Hl, int Wl, int ups, int nh, __half* __restrict__ hi,
                                      __half* __restrict__ lo, float* __restrict__ inv_scale,
                                      float* __restrict__ partial) {
    extern __shared__ float red[];  // [lanes][nh]
    const int cg = nh >> 2;
    const int lanes = blockDim.x / cg;
    const int g = threadIdx.x % cg, pl = threadIdx.x / cg;
    const int f = 1 << ups;
    // |G| <= f*f * max|dsrc|
    const float scale = pow2_scale_for(__ldg(amax) * (float)(f * f), 14);
    if (blockIdx.x == 0 && threadIdx.x == 0) inv_scale[0] = 1.f / scale;
    const int64_t npix = (int64_t)B * Hl * Wl;
    const int64_t p0 = (int64_t)blockIdx.x * GP_PIX;
    const int H = Hl << ups, W = Wl << ups;
    float s[4] = {0, 0, 0, 0};
    if (pl < lanes) {
        for (int i = pl; i < GP_PIX; i += lanes) {
            const int64_t pix = p0 + i;
            if (pix >= npix) break;
            const uint32_t pu = (uint32_t)pix;  // 32-bit coordinates (host wrapper checks npix < 2^31)
            const int xl = (int)(pu % (uint32_t)Wl);
            const uint32_t t2 = pu / (uint32_t)Wl;
            const int yl = (int)(t2 % (uint32_t)Hl);
            const int b = (int)(t2 / (uint32_t)Hl);
            const size_t fp0 = ((size_t)b * H + yl * f) * W + xl * f;
            const uint2 av = __ldg(reinterpret_cast<const uint2*>(actv_hi + fp0 * nh) + g);
            float a[4] = {0, 0, 0, 0};
            for (int sy = 0; sy < f; ++sy)
                for (int sx = 0; sx < f; ++sx) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(
                        dsrc + (fp0 + (size_t)sy * W + sx) * ld + coff + g * 4));
                    a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
                }
            // ReLU gate: actv > 0  <=>  fp16 hi plane positive (sign clear, magnitude non-zero)
            const uint32_t m[4] = {av.x & 0xffffu, av.x >> 16, av.y & 0xffffu, av.y >> 16};
            uint32_t ph[2], plw[2];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (!(m[e] != 0 && m[e] < 0x8000u)) a[e] = 0.f;
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
            *reinterpret_cast<uint2*>(hi + (size_t)pix * nh + g * 4) = make_uint2(ph[0], ph[1]);
            if (lo) *reinterpret_cast<uint2*>(lo + (size_t)pix * nh + g * 4) = make_uint2(plw[0], plw[1]);
#pragma unroll
            for (int e = 0; e < 4; ++e) s[e] += a[e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) red[(size_t)pl * nh + g * 4 + e] = s[e];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nh; i += blockDim.x) {
        float a = 0.f;
        for (int l = 0; l < lanes; ++l) a += red[(size_t)l * nh + i];
        partial[(size_t)blockIdx.x * nh + i] = a;
    }
}

// one-hot of a uint8 label map as an fp16 NHWC plane with Lp channels (zeros beyond the label range)
__global__ void onehot_planes_kernel(const uint8_t* __restrict__ labels, __half* __restrict__ out,
                                     int64_t npix, int Lp8) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix * Lp8) return;
    const int g = (int)(i % Lp8);
    const int l = labels[i / Lp8] - g * 8;
    uint32_t w[4] = {0, 0, 0, 0};
    if (l >= 0 && l < 8) w[l >> 1] = (l & 1) ? 0x3C000000u : 0x00003C00u;  // fp16 1.0
    reinterpret_cast<uint4*>(out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
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
//   dW[o][c][tap]  = sum_p dpre[o][p - tap] * lrelu(x[p][c]),   db[o] = sum_p dpre[o][p]
// One pass over x: a persistent block walks 8x16 pixel tiles; the 27 values dpre[o][p - tap] of
// every tile pixel are staged in shared memory (read as warp broadcasts), a thread owns 2 channels
// with its 54 weights and 54 dW accumulators in registers.
// ------------------------------------------------------------------------------------------------
constexpr int HB_TH = 8, HB_TW = 16, HB_PX = HB_TH * HB_TW;
__global__ void __launch_bounds__(512)
head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ out,
                const float* __restrict__ dout, int B, int H, int W, int C, float* __restrict__ dx,
                float* __restrict__ partial /*[grid][C][28]*/) {
    __shared__ float patch[HB_PX][28];
    const int c2 = threadIdx.x * 2;
    float wr[2][27], acc[2][27];
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            wr[e][k] = __ldg(w + ((size_t)(k / 9) * C + c2 + e) * 9 + k % 9);
            acc[e][k] = 0.f;
        }
    float bacc = 0.f;
    const int tiles_w = (W + HB_TW - 1) / HB_TW, tiles_h = (H + HB_TH - 1) / HB_TH;
    const int ntiles = B * tiles_h * tiles_w;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int w0 = (tile % tiles_w) * HB_TW;
        const int h0 = ((tile / tiles_w) % tiles_h) * HB_TH;
        const int b = tile / (tiles_w * tiles_h);
        __syncthreads();
        for (int i = threadIdx.x; i < HB_PX * 27; i += blockDim.x) {
            const int pi = i / 27, k = i % 27;
            const int o = k / 9, tap = k % 9;
            const int yy = h0 + pi / HB_TW - (tap / 3 - 1), xx = w0 + pi % HB_TW - (tap % 3 - 1);
            float v = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const size_t idx = (((size_t)b * 3 + o) * H + yy) * W + xx;
                const float ov = __ldg(out + idx);
                v = __ldg(dout + idx) * (1.f - ov * ov);
            }
            patch[pi][k] = v;
        }
        __syncthreads();
        for (int pi = 0; pi < HB_PX; ++pi) {
            const int yy = h0 + pi / HB_TW, xx = w0 + pi % HB_TW;
            if (yy >= H || xx >= W) continue;  // block-uniform
            const size_t pix = ((size_t)b * H + yy) * W + xx;
            const float2 xv = __ldg(reinterpret_cast<const float2*>(x + pix * C + c2));
            const float v0 = xv.x > 0.f ? xv.x : 0.2f * xv.x, v1 = xv.y > 0.f ? xv.y : 0.2f * xv.y;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int k = 0; k < 27; ++k) {
                const float d = patch[pi][k];
                a0 += d * wr[0][k];
                a1 += d * wr[1][k];
                acc[0][k] += d * v0;
                acc[1][k] += d * v1;
            }
            *reinterpret_cast<float2*>(dx + pix * C + c2) =
                make_float2(a0 * (xv.x > 0.f ? 1.f : 0.2f), a1 * (xv.y > 0.f ? 1.f : 0.2f));
            if (threadIdx.x < 3) bacc += patch[pi][threadIdx.x * 9 + 4];  // centre tap = dpre[o][p]
        }
    }
    float* o = partial + (size_t)blockIdx.x * C * 28;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
#pragma unroll
        for (int k = 0; k < 27; ++k) o[(size_t)(c2 + e) * 28 + k] = acc[e][k];
        if (c2 + e >= 3) o[(size_t)(c2 + e) * 28 + 27] = 0.f;
    }
    if (threadIdx.x < 3) o[(size_t)threadIdx.x * 28 + 27] = bacc;
}

}  // namespace dsee

using namespace dsee;

#define LAUNCH_END()               \
    count_launch();                \
    DSEE_CUDA(cudaGetLastError()); \
    return 0

extern "C" int dsee_grad_prep_blocks(int64_t npix) { return cdivb(npix, GP_PIX); }

extern "C" int dsee_grad_prep(const float* dy, void* out_hi, void* out_lo, float* inv_scale,
                              const float* noise0, const float* noise1, unsigned long long seed0,
                              unsigned long long seed1, int64_t npix, int C, float* partial,
                              const float* amax_in, void* stream) {
    DSEE_CHECK_ARG(dy && out_hi && inv_scale && partial && npix > 0, "bad argument");
    DSEE_CHECK_ARG(C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0, "C must divide 1024 (got %d)", C);
    const bool has0 = noise0 || seed0, has1 = noise1 || seed1;
    DSEE_CHECK_ARG(!(has1 && !has0), "noise1 without noise0");
    int rc = require_sm100();
    if (rc) return rc;
    const int nq = 1 + (has0 ? 1 : 0) + (has1 ? 1 : 0);
    const int lanes = 256 / (C / 4);
    size_t sm = (size_t)lanes * C * nq * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (!amax_in) {
        DSEE_CUDA(cudaMemsetAsync(inv_scale, 0, 2 * sizeof(float), st));
        const int64_t n4 = npix * C / 4;
        int ablocks = cdivb(n4, 256 * 8);
        if (ablocks > 148 * 8) ablocks = 148 * 8;
        grad_amax_kernel<<<ablocks, 256, 0, st>>>(dy, n4, inv_scale + 1);
        count_launch();
    }
    grad_prep_kernel<<<cdivb(npix, GP_PIX), 256, sm, st>>>(dy, (__half*)out_hi, (__half*)out_lo,
                                                           inv_scale, noise0, noise1, seed0, seed1, npix,
                                                           C, partial, nq, amax_in, noise_epoch_ptr());
    LAUNCH_END();
}

extern "C" int dsee_spade_modulate_bwd_saved(const float* x, int x_ups, const float* noise,
                                             unsigned long long noise_seed, const float* noise_w,
                                             const float* bn_scale,
                                             const float* bn_shift, const void* g_hi, const void* g_lo,
                                             const float* dt, const float* dt_amax, int B, int H, int W,
                                             int C, float* dxhat, void* dgb_hi, void* dgb_lo,
                                             float* dgb_inv_scale, float* partial, void* stream) {
    DSEE_CHECK_ARG(x && bn_scale && bn_shift && g_hi && dt && dt_amax && dxhat && dgb_hi && dgb_inv_scale &&
                       partial,
                   "NULL pointer");
    DSEE_CHECK_ARG(C % 128 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0, "C must be 128, 256, 512 or 1024");
    DSEE_CHECK_ARG((x_ups == 0 || x_ups == 1) &&
                       (noise != nullptr || noise_seed != 0) == (noise_w != nullptr),
                   "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    const int64_t npix = (int64_t)B * H * W;
    DSEE_CHECK_ARG(npix < ((int64_t)1 << 31), "more than 2^31 pixels");
    const int lanes = 256 / (C / 4);
    const size_t sm = (size_t)lanes * C * 4 * sizeof(float);
    modulate_bwd_saved_kernel<<<cdivb(npix, GP_PIX), 256, sm, (cudaStream_t)stream>>>(
        x, x_ups, noise, noise_seed, noise_w, bn_scale, bn_shift, (const __half*)g_hi, (const __half*)g_lo, dt,
        dt_amax,
        B, H, W, C, dxhat, (__half*)dgb_hi, (__half*)dgb_lo, dgb_inv_scale, partial, noise_epoch_ptr());
    LAUNCH_END();
}

extern "C" int dsee_reduce_partials(const float* partial, int n, int C, int nq, float scale,
                                    float* out, void* stream) {
    DSEE_CHECK_ARG(partial && out && n > 0 && C > 0 && nq > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    DSEE_CHECK_ARG(nq <= 4, "at most 4 quantities per channel (got %d)", nq);
    launch_reduce_partials(partial, n, C, nq, scale, out, (cudaStream_t)stream);
    LAUNCH_END();
}

extern "C" int dsee_bn_bwd_blocks(int B, int Hx, int Wx) { return cdivb((int64_t)B * Hx * Wx, BB_PIX); }

extern "C" int dsee_bn_bwd(const float* dxhat, const float* x, int x_ups, const float* noise,
                           unsigned long long noise_seed, const float* noise_w, const float* bn_scale,
                           const float* bn_shift,
                           const float* sums, float inv_count, const float* dskip, int B, int Hx,
                           int Wx, int C, float* dx, float* nw_partial, int noise_grad_with_skip,
                           float* amax_out, void* stream) {
    DSEE_CHECK_ARG(dxhat && x && bn_scale && bn_shift && sums && dx, "NULL pointer");
    DSEE_CHECK_ARG(C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0, "C must divide 1024 (got %d)", C);
    DSEE_CHECK_ARG((noise != nullptr || noise_seed != 0) == (noise_w != nullptr), "noise/noise_w mismatch");
    DSEE_CHECK_ARG(!nw_partial || noise_w, "nw_partial needs noise");
    DSEE_CHECK_ARG((int64_t)B * Hx * Wx < ((int64_t)1 << 31), "more than 2^31 pixels");
    int rc = require_sm100();
    if (rc) return rc;
    const int lanes = 256 / (C / 4);
    size_t sm = nw_partial ? (size_t)lanes * C * sizeof(float) : 0;
    DSEE_CHECK_ARG(!noise_grad_with_skip || (dskip && nw_partial), "noise_grad_with_skip needs dskip and nw_partial");
    if (amax_out) DSEE_CUDA(cudaMemsetAsync(amax_out, 0, sizeof(float), (cudaStream_t)stream));
    auto kern = noise_w ? bn_bwd_kernel<true> : bn_bwd_kernel<false>;
    kern<<<dsee_bn_bwd_blocks(B, Hx, Wx), 256, sm, (cudaStream_t)stream>>>(
        dxhat, x, x_ups, noise, noise_seed, noise_w, bn_scale, bn_shift, sums, inv_count, dskip, B, Hx, Wx,
        C, dx, nw_partial, noise_grad_with_skip, amax_out, noise_epoch_ptr());
    LAUNCH_END();
}

extern "C" int dsee_actv_grad_prep(const float* dsrc, int ld, int coff, const void* actv_hi,
                                   const float* dsrc_amax, int B, int Hl, int Wl, int ups, int nh,
                                   void* out_hi, void* out_lo, float* inv_scale, float* partial,
                                   void* stream) {
    DSEE_CHECK_ARG(dsrc && actv_hi && dsrc_amax && out_hi && inv_scale && partial, "NULL pointer");
    DSEE_CHECK_ARG(nh % 4 == 0 && nh / 4 <= 256 && 256 % (nh / 4) == 0 && (ups == 0 || ups == 1) &&
                       ld % 4 == 0 && coff % 4 == 0,
                   "bad shape");
    int rc = require_sm100();
    if (rc) return rc;
    const int64_t npix = (int64_t)B * Hl * Wl;
    DSEE_CHECK_ARG(npix < ((int64_t)1 << 31), "more than 2^31 pixels");
    const int lanes = 256 / (nh / 4);
    actv_grad_prep_kernel<<<cdivb(npix, GP_PIX), 256, (size_t)lanes * nh * sizeof(float),
                            (cudaStream_t)stream>>>(dsrc, ld, coff, (const __half*)actv_hi, dsrc_amax,
                                                    B, Hl, Wl, ups, nh, (__half*)out_hi, (__half*)out_lo,
                                                    inv_scale, partial);
    LAUNCH_END();
}

extern "C" int dsee_onehot_planes(const uint8_t* labels, void* out, int64_t npix, int Lp, void* stream) {
    DSEE_CHECK_ARG(labels && out && npix > 0 && Lp > 0 && Lp % 8 == 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    onehot_planes_kernel<<<cdivb(npix * (Lp / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        labels, (__half*)out, npix, Lp / 8);
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
    launch_reduce_partials(partial, blocks, rows * nh, 1, 1.f, dtable_dbias, st);
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
    launch_reduce_partials(partial, blocks, C * 28, 1, 1.f, dw_db, st);
    LAUNCH_END();
}

extern "C" int dsee_head_bwd_blocks(int B, int H, int W) {
    const int ntiles = B * ((H + HB_TH - 1) / HB_TH) * ((W + HB_TW - 1) / HB_TW);
    return ntiles < 2 * 148 ? ntiles : 2 * 148;
}

extern "C" int dsee_head_bwd(const float* x, const float* w, const float* out, const float* dout,
                             int B, int H, int W, int C, float* dx, float* partial, float* dw_db,
                             void* stream) {
    DSEE_CHECK_ARG(x && w && out && dout && dx && partial && dw_db, "NULL pointer");
    DSEE_CHECK_ARG(C % 2 == 0 && C >= 4 && C <= 1024, "C must be even, 4..1024");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = dsee_head_bwd_blocks(B, H, W);
    head_bwd_kernel<<<blocks, C / 2, 0, st>>>(x, w, out, dout, B, H, W, C, dx, partial);
    count_launch();
    launch_reduce_partials(partial, blocks, C * 28, 1, 1.f, dw_db, st);
    LAUNCH_END();
}
