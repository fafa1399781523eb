// Parameter-side kernels: spectral normalisation of conv weights and the assembly of the fused
// modulation weight.  These are tiny (a 512 x 4608 matrix is 9.4 MB) but the reference does them
// with ~12 and ~25 ATen launches per layer per forward; fusing them keeps the training step
// GPU-bound instead of launch-bound.  Every reduction is two-stage with a fixed order.
//
// Reference: torch/nn/utils/spectral_norm.py:92-113 (as applied at architecture.py:40-44 and
// normalization.py:29-31); normalization.py:116-119, 198-213, 283-286 (gamma/beta convs and the
// sigmoid(alpha) blend).
#include "common.cuh"
#include <string.h>
#include "launch_count.h"
#include "../../include/deepsee_b200.h"

namespace dsee {

static inline int cdivw(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr int SN_SLICES = 8;

// partial[s][k] = sum_{n in slice s} W[n][k] * u[n]
__global__ void sn_wt_u_kernel(const float* __restrict__ w, const float* __restrict__ u, int N, int K,
                               float* __restrict__ partial) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int s = blockIdx.y;
    const int n0 = (int)((int64_t)N * s / SN_SLICES), n1 = (int)((int64_t)N * (s + 1) / SN_SLICES);
    float a = 0.f;
    for (int n = n0; n < n1; ++n) a += __ldg(w + (size_t)n * K + k) * __ldg(u + n);
    partial[(size_t)s * K + k] = a;
}

__device__ __forceinline__ double block_sum_double(double v, double* sh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double t = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) t += sh[i];
    __syncthreads();
    return t;
}

// v[k] = t[k] / max(||t||, eps), t = sum of the slices (one block)
__global__ void sn_finish_v_kernel(const float* __restrict__ partial, int K, float eps,
                                   float* __restrict__ v) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float t = 0.f;
        for (int s = 0; s < SN_SLICES; ++s) t += partial[(size_t)s * K + k];
        v[k] = t;
        acc += (double)t * t;
    }
    const double nrm = sqrt(block_sum_double(acc, sh));
    const float inv = 1.f / fmaxf((float)nrm, eps);
    for (int k = threadIdx.x; k < K; k += blockDim.x) v[k] *= inv;
}

// s[n] = sum_k W[n][k] * v[k]  (warp per row)
__global__ void sn_w_v_kernel(const float* __restrict__ w, const float* __restrict__ v, int N, int K,
                              float* __restrict__ s) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    float a = 0.f;
    for (int k = lane; k < K; k += 32) a += __ldg(w + (size_t)n * K + k) * __ldg(v + k);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) s[n] = a;
}

// training: u = s / max(||s||, eps); always: sigma = u . s; out[0] = sigma, out[1] = 1/sigma
__global__ void sn_finish_u_kernel(const float* __restrict__ s, int N, float eps, int update_u,
                                   float* __restrict__ u, float* __restrict__ sig) {
    __shared__ double sh[32];
    double acc = 0.0;
    if (update_u) {
        for (int n = threadIdx.x; n < N; n += blockDim.x) acc += (double)s[n] * s[n];
        const double nrm = sqrt(block_sum_double(acc, sh));
        const float inv = 1.f / fmaxf((float)nrm, eps);
        for (int n = threadIdx.x; n < N; n += blockDim.x) u[n] = s[n] * inv;
        __syncthreads();
    }
    acc = 0.0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) acc += (double)u[n] * s[n];
    const double sigma = block_sum_double(acc, sh);
    if (threadIdx.x == 0) {
        sig[0] = (float)sigma;
        sig[1] = (float)(1.0 / sigma);
    }
}

__global__ void scale_by_kernel(const float* __restrict__ in, const float* __restrict__ scale,
                                float* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] * __ldg(scale);
}

// ------------------------------------------------------------------------------------------------
// The same five steps for up to SN_MAX_BATCH layers per launch (blockIdx.z / .y / .x = layer): a
// network's spectral-norm layers are independent, and one launch per step per NETWORK instead of per
// LAYER removes ~250 launches of microsecond kernels from every training iteration.
// ------------------------------------------------------------------------------------------------
struct SnBatch {
    dsee_sn_item it[DSEE_SN_MAX_BATCH];
    int count;
};
__global__ void sn_wt_u_batched_kernel(const __grid_constant__ SnBatch b) {
    const dsee_sn_item& t = b.it[blockIdx.z];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= t.K) return;
    const int s = blockIdx.y;
    const int n0 = (int)((int64_t)t.N * s / SN_SLICES), n1 = (int)((int64_t)t.N * (s + 1) / SN_SLICES);
    float a = 0.f;
    for (int n = n0; n < n1; ++n) a += __ldg(t.w_orig + (size_t)n * t.K + k) * __ldg(t.u + n);
    t.workspace[(size_t)s * t.K + k] = a;
}
__global__ void sn_finish_v_batched_kernel(const __grid_constant__ SnBatch b, float eps) {
    const dsee_sn_item& t = b.it[blockIdx.x];
    __shared__ double sh[32];
    double acc = 0.0;
    for (int k = threadIdx.x; k < t.K; k += blockDim.x) {
        float x = 0.f;
        for (int s = 0; s < SN_SLICES; ++s) x += t.workspace[(size_t)s * t.K + k];
        t.v[k] = x;
        acc += (double)x * x;
    }
    const double nrm = sqrt(block_sum_double(acc, sh));
    const float inv = 1.f / fmaxf((float)nrm, eps);
    for (int k = threadIdx.x; k < t.K; k += blockDim.x) t.v[k] *= inv;
}
__global__ void sn_w_v_batched_kernel(const __grid_constant__ SnBatch b) {
    const dsee_sn_item& t = b.it[blockIdx.y];
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= t.N) return;
    float a = 0.f;
    for (int k = lane; k < t.K; k += 32) a += __ldg(t.w_orig + (size_t)n * t.K + k) * __ldg(t.v + k);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) t.workspace[(size_t)SN_SLICES * t.K + n] = a;
}
__global__ void sn_finish_u_batched_kernel(const __grid_constant__ SnBatch b, float eps, int update_u) {
    const dsee_sn_item& t = b.it[blockIdx.x];
    const float* s = t.workspace + (size_t)SN_SLICES * t.K;
    __shared__ double sh[32];
    double acc = 0.0;
    if (update_u) {
        for (int n = threadIdx.x; n < t.N; n += blockDim.x) acc += (double)s[n] * s[n];
        const double nrm = sqrt(block_sum_double(acc, sh));
        const float inv = 1.f / fmaxf((float)nrm, eps);
        for (int n = threadIdx.x; n < t.N; n += blockDim.x) t.u[n] = s[n] * inv;
        __syncthreads();
    }
    acc = 0.0;
    for (int n = threadIdx.x; n < t.N; n += blockDim.x) acc += (double)t.u[n] * s[n];
    const double sigma = block_sum_double(acc, sh);
    if (threadIdx.x == 0) {
        t.sigma2[0] = (float)sigma;
        t.sigma2[1] = (float)(1.0 / sigma);
    }
    // this forward's u / v for the backward pass (the buffers are overwritten by the next forward)
    if (t.u_saved)
        for (int n = threadIdx.x; n < t.N; n += blockDim.x) t.u_saved[n] = t.u[n];
    if (t.v_saved)
        for (int k = threadIdx.x; k < t.K; k += blockDim.x) t.v_saved[k] = t.v[k];
}
__global__ void scale_by_batched_kernel(const __grid_constant__ SnBatch b) {
    const dsee_sn_item& t = b.it[blockIdx.y];
    const int64_t n = (int64_t)t.N * t.K;
    const float sc = __ldg(t.sigma2 + 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        t.w_eff[i] = t.w_orig[i] * sc;
}

// backward of W_eff = W / sigma(W), sigma = u^T W v with u, v constants:
//   dW = (dW_eff - <dW_eff, W_eff> u v^T) / sigma
__global__ void dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                   double* __restrict__ partial) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        acc += (double)a[i] * b[i];
    const double t = block_sum_double(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void sn_bwd_kernel(const float* __restrict__ dweff, const float* __restrict__ u,
                              const float* __restrict__ v, const float* __restrict__ sig,
                              const double* __restrict__ partial, int nparts, int N, int K,
                              float* __restrict__ dw) {
    __shared__ float dot_s;
    if (threadIdx.x == 0) {
        double d = 0.0;
        for (int i = 0; i < nparts; ++i) d += partial[i];
        dot_s = (float)d;
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * K) return;
    const int n = (int)(i / K), k = (int)(i % K);
    dw[i] = (dweff[i] - dot_s * __ldg(u + n) * __ldg(v + k)) * __ldg(sig + 1);
}

// ------------------------------------------------------------------------------------------------
// modulation weight assembly.  Output rows interleaved per 128 channels [g(128) | b(128) | ...],
// columns [seg source (c1) | style source (c2)]:
//   Wm[g-row c][0:c1]      = (1 - a_g) * Wg[c]      Wm[g-row c][c1:] = a_g * Wsg[c]
//   gb[c] = (1 - a_g) * bg[c] + a_g * bsg[c] (+ 1)  (same for beta with a_b, no +1)
// a = sigmoid(alpha) when both sources exist, a = 0 with the seg source only (SPADE), a = 1 with the
// style source only (PureSEAN).
// ------------------------------------------------------------------------------------------------
struct CombineArgs {
    const float* w_seg[2];  // [gamma, beta], each [C][c1][9] or NULL
    const float* w_sty[2];  // each [C][c2][9] or NULL
    const float* b_seg[2];
    const float* b_sty[2];
    const float* alpha[2];  // device scalars or NULL
    int C, c1, c2;
    float plus_one;
};
__device__ __forceinline__ float blend_a(const CombineArgs& p, int gb) {
    if (p.w_seg[gb] && p.w_sty[gb]) return 1.f / (1.f + __expf(-__ldg(p.alpha[gb])));
    return p.w_sty[gb] ? 1.f : 0.f;
}
__global__ void combine_fwd_kernel(CombineArgs p, float* __restrict__ wm, float* __restrict__ gb,
                                   float* __restrict__ bb) {
    const int ct = p.c1 + p.c2;
    const int64_t total = (int64_t)2 * p.C * ct * 9;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) {
        const uint32_t iu = (uint32_t)i;  // total < 2^31 (host wrapper): 32-bit div/mod
        const int tap = (int)(iu % 9u);
        const int col = (int)((iu / 9u) % (uint32_t)ct);
        const int row = (int)(iu / (9u * (uint32_t)ct));
        const int which = (row >> 7) & 1;                 // 0 gamma, 1 beta
        const int c = ((row >> 8) << 7) + (row & 127);    // channel
        const float a = blend_a(p, which);
        float v;
        if (col < p.c1) v = (1.f - a) * __ldg(p.w_seg[which] + ((size_t)c * p.c1 + col) * 9 + tap);
        else v = a * __ldg(p.w_sty[which] + ((size_t)c * p.c2 + (col - p.c1)) * 9 + tap);
        wm[i] = v;
    }
    if (i < 2 * p.C) {
        const int which = (int)(i / p.C), c = (int)(i % p.C);
        const float a = blend_a(p, which);
        float v = 0.f;
        if (p.b_seg[which]) v += (1.f - a) * __ldg(p.b_seg[which] + c);
        if (p.b_sty[which]) v += a * __ldg(p.b_sty[which] + c);
        if (which == 0) gb[c] = v + p.plus_one;
        else bb[c] = v;
    }
}

// backward: source gradients (elementwise) + block partials of d a_g, d a_b (a = sigmoid(alpha))
struct CombineBwdOut {
    float* dw_seg[2];
    float* dw_sty[2];
    float* db_seg[2];
    float* db_sty[2];
};
__global__ void combine_bwd_kernel(CombineArgs p, const float* __restrict__ dwm,
                                   const float* __restrict__ dgb, const float* __restrict__ dbb,
                                   CombineBwdOut o, double* __restrict__ partial /*[blocks][2]*/) {
    __shared__ double sh[32];
    const int ct = p.c1 + p.c2;
    const int64_t total = (int64_t)2 * p.C * ct * 9;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double da[2] = {0.0, 0.0};
    if (i < total) {
        const uint32_t iu = (uint32_t)i;  // total < 2^31 (host wrapper): 32-bit div/mod
        const int tap = (int)(iu % 9u);
        const int col = (int)((iu / 9u) % (uint32_t)ct);
        const int row = (int)(iu / (9u * (uint32_t)ct));
        const int which = (row >> 7) & 1;
        const int c = ((row >> 8) << 7) + (row & 127);
        const float a = blend_a(p, which);
        const float g = dwm[i];
        if (col < p.c1) {
            const size_t j = ((size_t)c * p.c1 + col) * 9 + tap;
            o.dw_seg[which][j] = (1.f - a) * g;
            da[which] -= (double)g * __ldg(p.w_seg[which] + j);
        } else {
            const size_t j = ((size_t)c * p.c2 + (col - p.c1)) * 9 + tap;
            o.dw_sty[which][j] = a * g;
            da[which] += (double)g * __ldg(p.w_sty[which] + j);
        }
    }
    if (i < 2 * p.C) {
        const int which = (int)(i / p.C), c = (int)(i % p.C);
        const float a = blend_a(p, which);
        const float g = which == 0 ? dgb[c] : dbb[c];
        if (p.b_seg[which]) {
            o.db_seg[which][c] = (1.f - a) * g;
            da[which] -= (double)g * __ldg(p.b_seg[which] + c);
        }
        if (p.b_sty[which]) {
            o.db_sty[which][c] = a * g;
            da[which] += (double)g * __ldg(p.b_sty[which] + c);
        }
    }
    const double t0 = block_sum_double(da[0], sh);
    const double t1 = block_sum_double(da[1], sh);
    if (threadIdx.x == 0) {
        partial[(size_t)blockIdx.x * 2] = t0;
        partial[(size_t)blockIdx.x * 2 + 1] = t1;
    }
}
__global__ void combine_bwd_alpha_kernel(CombineArgs p, const double* __restrict__ partial, int nblocks,
                                         float* __restrict__ dalpha_g, float* __restrict__ dalpha_b) {
    // grid = 2 blocks (gamma, beta); fixed assignment of partials to threads -> deterministic
    __shared__ double sh[32];
    const int which = blockIdx.x;
    float* out = which == 0 ? dalpha_g : dalpha_b;
    if (!out) return;
    double d = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) d += partial[(size_t)i * 2 + which];
    d = block_sum_double(d, sh);
    if (threadIdx.x == 0) {
        const float a = blend_a(p, which);
        *out = (float)(d * (double)(a * (1.f - a)));
    }
}

}  // namespace dsee

using namespace dsee;

#define LAUNCH_END()               \
    count_launch();                \
    DSEE_CUDA(cudaGetLastError()); \
    return 0

extern "C" int64_t dsee_spectral_workspace_floats(int N, int K) { return (int64_t)SN_SLICES * K + N; }

extern "C" int dsee_spectral_weight_fwd(const float* w_orig, float* u, float* v, int N, int K,
                                        int power_iteration, float eps, float* workspace,
                                        float* sigma2, float* w_eff, void* stream) {
    DSEE_CHECK_ARG(w_orig && u && v && workspace && sigma2 && w_eff && N > 0 && K > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    float* s = workspace + (size_t)SN_SLICES * K;
    if (power_iteration) {
        sn_wt_u_kernel<<<dim3(cdivw(K, 256), SN_SLICES), 256, 0, st>>>(w_orig, u, N, K, workspace);
        count_launch();
        sn_finish_v_kernel<<<1, 1024, 0, st>>>(workspace, K, eps, v);
        count_launch();
    }
    sn_w_v_kernel<<<cdivw((int64_t)N * 32, 256), 256, 0, st>>>(w_orig, v, N, K, s);
    count_launch();
    sn_finish_u_kernel<<<1, 1024, 0, st>>>(s, N, eps, power_iteration, u, sigma2);
    count_launch();
    const int64_t n = (int64_t)N * K;
    scale_by_kernel<<<cdivw(n, 256), 256, 0, st>>>(w_orig, sigma2 + 1, w_eff, n);
    LAUNCH_END();
}

extern "C" int dsee_spectral_weight_fwd_batched(const dsee_sn_item* items, int count, int power_iteration,
                                                float eps, void* stream) {
    DSEE_CHECK_ARG(items && count > 0, "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    for (int base = 0; base < count; base += DSEE_SN_MAX_BATCH) {
        SnBatch b;
        memset(&b, 0, sizeof(b));
        b.count = count - base < DSEE_SN_MAX_BATCH ? count - base : DSEE_SN_MAX_BATCH;
        int maxN = 0, maxK = 0;
        int64_t maxNK = 0;
        for (int i = 0; i < b.count; ++i) {
            const dsee_sn_item& t = items[base + i];
            DSEE_CHECK_ARG(t.w_orig && t.u && t.v && t.w_eff && t.sigma2 && t.workspace && t.N > 0 && t.K > 0,
                           "bad spectral-norm item %d", base + i);
            b.it[i] = t;
            maxN = t.N > maxN ? t.N : maxN;
            maxK = t.K > maxK ? t.K : maxK;
            maxNK = (int64_t)t.N * t.K > maxNK ? (int64_t)t.N * t.K : maxNK;
        }
        if (power_iteration) {
            sn_wt_u_batched_kernel<<<dim3(cdivw(maxK, 256), SN_SLICES, b.count), 256, 0, st>>>(b);
            count_launch();
            sn_finish_v_batched_kernel<<<b.count, 1024, 0, st>>>(b, eps);
            count_launch();
        }
        sn_w_v_batched_kernel<<<dim3(cdivw((int64_t)maxN * 32, 256), b.count), 256, 0, st>>>(b);
        count_launch();
        sn_finish_u_batched_kernel<<<b.count, 1024, 0, st>>>(b, eps, power_iteration);
        count_launch();
        int blocks = cdivw(maxNK, 256 * 4);
        if (blocks > 2048) blocks = 2048;
        scale_by_batched_kernel<<<dim3(blocks, b.count), 256, 0, st>>>(b);
        count_launch();
        DSEE_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int dsee_spectral_weight_bwd(const float* dw_eff, const float* w_eff, const float* u,
                                        const float* v, const float* sigma2, int N, int K,
                                        void* workspace, float* dw_orig, void* stream) {
    DSEE_CHECK_ARG(dw_eff && w_eff && u && v && sigma2 && workspace && dw_orig && N > 0 && K > 0,
                   "bad argument");
    int rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)N * K;
    const int parts = 128;  // workspace: 128 doubles
    dot_partial_kernel<<<parts, 256, 0, st>>>(dw_eff, w_eff, n, (double*)workspace);
    count_launch();
    sn_bwd_kernel<<<cdivw(n, 256), 256, 0, st>>>(dw_eff, u, v, sigma2, (const double*)workspace, parts, N,
                                                 K, dw_orig);
    LAUNCH_END();
}

static int fill_combine(CombineArgs& p, const dsee_modweight_args* a) {
    DSEE_CHECK_ARG(a && a->C > 0 && a->C % 128 == 0 && a->c1 >= 0 && a->c2 >= 0 && a->c1 + a->c2 > 0,
                   "bad modulation-weight shape");
    for (int i = 0; i < 2; ++i) {
        DSEE_CHECK_ARG((a->c1 > 0) == (a->w_seg[i] != nullptr) && (a->c2 > 0) == (a->w_sty[i] != nullptr),
                       "source pointers must match c1 / c2");
        DSEE_CHECK_ARG(!(a->c1 > 0 && a->c2 > 0) || a->alpha[i], "two sources need alpha");
        p.w_seg[i] = a->w_seg[i];
        p.w_sty[i] = a->w_sty[i];
        p.b_seg[i] = a->b_seg[i];
        p.b_sty[i] = a->b_sty[i];
        p.alpha[i] = a->alpha[i];
    }
    p.C = a->C;
    p.c1 = a->c1;
    p.c2 = a->c2;
    p.plus_one = a->plus_one ? 1.f : 0.f;
    return 0;
}

extern "C" int dsee_modweight_fwd(const dsee_modweight_args* a, float* wm, float* gamma_bias,
                                  float* beta_bias, void* stream) {
    DSEE_CHECK_ARG(wm && gamma_bias && beta_bias, "NULL output");
    CombineArgs p;
    int rc = fill_combine(p, a);
    if (rc) return rc;
    rc = require_sm100();
    if (rc) return rc;
    const int64_t total = (int64_t)2 * p.C * (p.c1 + p.c2) * 9;
    DSEE_CHECK_ARG(total < ((int64_t)1 << 31), "modulation weight too large");
    combine_fwd_kernel<<<cdivw(total, 256), 256, 0, (cudaStream_t)stream>>>(p, wm, gamma_bias, beta_bias);
    LAUNCH_END();
}

extern "C" int64_t dsee_modweight_bwd_workspace_bytes(int C, int c1, int c2) {
    return (int64_t)cdivw((int64_t)2 * C * (c1 + c2) * 9, 256) * 2 * (int64_t)sizeof(double);
}

extern "C" int dsee_modweight_bwd(const dsee_modweight_args* a, const float* dwm, const float* dgb,
                                  const float* dbb, const dsee_modweight_grads* g, void* workspace,
                                  void* stream) {
    DSEE_CHECK_ARG(dwm && dgb && dbb && g && workspace, "NULL argument");
    CombineArgs p;
    int rc = fill_combine(p, a);
    if (rc) return rc;
    CombineBwdOut o;
    for (int i = 0; i < 2; ++i) {
        DSEE_CHECK_ARG((p.c1 == 0 || (g->dw_seg[i] && (!p.b_seg[i] || g->db_seg[i]))) &&
                           (p.c2 == 0 || (g->dw_sty[i] && (!p.b_sty[i] || g->db_sty[i]))),
                       "missing gradient output");
        o.dw_seg[i] = g->dw_seg[i];
        o.dw_sty[i] = g->dw_sty[i];
        o.db_seg[i] = g->db_seg[i];
        o.db_sty[i] = g->db_sty[i];
    }
    rc = require_sm100();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)2 * p.C * (p.c1 + p.c2) * 9;
    DSEE_CHECK_ARG(total < ((int64_t)1 << 31), "modulation weight too large");
    const int blocks = cdivw(total, 256);
    combine_bwd_kernel<<<blocks, 256, 0, st>>>(p, dwm, dgb, dbb, o, (double*)workspace);
    count_launch();
    if (g->dalpha[0] || g->dalpha[1]) {
        combine_bwd_alpha_kernel<<<2, 256, 0, st>>>(p, (const double*)workspace, blocks, g->dalpha[0],
                                                   g->dalpha[1]);
        count_launch();
    }
    DSEE_CUDA(cudaGetLastError());
    return 0;
}
