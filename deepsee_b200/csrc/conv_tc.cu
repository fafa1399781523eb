// Implicit-GEMM 3x3 convolution on 5th-gen tensor cores (tcgen05 + TMEM), operands fed by TMA.
//
//   D[pixel, n] = sum_{pass, tap, c}  A_plane(pass)[pixel + tap, c] * W_plane(pass)[n, tap, c]
//
// * M = 128 output pixels = an 8 x 16 spatial tile of one image. For filter tap (dy,dx) the A tile
//   is the same box shifted by (dy,dx): one 4-D TMA load {64 ch, 16 w, 8 h, 1 b} with signed
//   coordinates; the zero padding of the convolution is the TMA out-of-bounds fill. No im2col
//   buffer exists anywhere.
// * N = 256 output channels per tile, K block = 64 input channels of one tap (128 B rows,
//   128B-swizzled K-major operands for tcgen05.mma kind::f16, M=128 N=256 K=16).
// * fp32 accumulators live in TMEM, double-buffered (2 x 256 columns) so the epilogue of tile i
//   overlaps the main loop of tile i+1. Persistent CTAs, one per SM, static round-robin tiles.
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread) + TMEM
//   allocator, warps 2..9 = epilogue (two per TMEM lane quarter, alternating 32-column chunks;
//   thread == output pixel).
//
// Epilogue data layout: tcgen05.ld hands a thread one pixel row (32 channels) of the accumulator.
// Global memory wants the opposite (a warp instruction covering consecutive channels of few pixels),
// so EPI_CONV and EPI_DGRAD_MODBWD pass every 32 x 32 chunk through a per-warp padded shared-memory
// tile and continue with lane = (pixel sub-row r = lane/8, channel quad cq = lane%8): 8 steps of
// 4 pixels x 32 channels, every global access a float4 / 8-byte access of a fully used 128-byte
// (64-byte for fp16 planes) segment, per-channel sums accumulated in registers (lane owns 4 channels).
//
// Epilogues sharing the main loop:
//   EPI_CONV      K2: + bias (+ residual, optionally read through a folded 2x nearest upsample)
//                 -> fp32 NHWC, optional per-channel sum / sum-of-squares tile partials.
//   EPI_MODULATE  K1: the accumulator columns are [gamma(128) | beta(128)]; the epilogue applies
//                 batch-norm scale/shift, x_hat*(gamma)+beta, LeakyReLU(0.2) and emits the fp16
//                 split planes that the next conv consumes.
//
//   EPI_DGRAD_MODBWD  backward-data of a main conv fused with K1's backward: the accumulator is
//                 dt (gradient wrt the conditional norm's output, LeakyReLU' applied from the saved
//                 activation's sign); the epilogue multiplies by the saved G / x_hat and emits
//                 dx_hat (fp32), the [dG | dB] gradient planes and the per-channel sums batch-norm's
//                 backward needs - dt itself never reaches HBM.
// Reference semantics: architecture.py:75-130, normalization.py:105-120,167-213,254-286.
#include "common.cuh"
#include "../../include/deepsee_b200.h"
#include "launch_count.h"
#include <string.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>

namespace dsee {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;
constexpr int TILE_W = 16;
constexpr int TILE_H = 8;
constexpr int STAGES = 4;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 48 KB
constexpr int EPI_WARPS = 8;              // two per TMEM lane quarter, alternating 32-column chunks
constexpr int EPI_TILE_FLOATS = 32 * 32;  // per-warp 32 pixel x 32 channel transpose tile (tile_put_row / tile_get4)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ +
                           EPI_WARPS * EPI_TILE_FLOATS * 4;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;
constexpr int MAX_TAPS = 16;

enum { EPI_CONV = 0, EPI_MODULATE = 1, EPI_MODULATE_BWD = 2, EPI_DGRAD_MODBWD = 3 };

struct alignas(64) ConvParams {
    CUtensorMap tmA[4];  // [source*2 + plane]
    CUtensorMap tmB[2];  // [plane]
    // fp8 correction phase (passes == 2): A planes [0] = (a - a_hi) * 2^8, [1] = a (e5m2, bytes), and
    // the e4m3 weight [n][tap][2][C]; K blocks of 128 bytes
    CUtensorMap tmA8[2];
    CUtensorMap tmB8;
    int cb8;           // 128-channel blocks per fp8 plane (0 = no fp8 phase)
    int w_brows;       // per-image weights: row offset of image b is b * w_brows (0 = shared)
    int sub4;          // 1: sub-pixel form, all four parity classes in this launch: the class is a tile
                       // index (fastest after the N tile), taps [cls*4, cls*4+4), weight rows + cls*n_total
    int pair;          // 1: CTA-pair form (cta_group::2): 16 x 16-pixel tiles, 8 rows and 128 weight rows per CTA
    int m2;            // 1: a tile is 16 x 16 pixels = two 128-pixel halves sharing one <= 128-row weight
                       // box (two MMAs per K step into TMEM columns [0,128) and [128,256)): narrow-N
                       // GEMMs then move as few operand bytes per FLOP as the N = 256 tile
    uint32_t idesc8;
    int B, H, W;       // tile space: the output pixels this launch computes, per image
    int Hm, Wm;        // output tensor dims in memory; pixel (y,x) of the tile space lives at
    int o_step, o_offy, o_offx;  //   (y*o_step + o_offy, x*o_step + o_offx)
    int ntaps, a_step; // filter taps; A box origin = tile origin * a_step + tap offset
    int8_t tap_dy[MAX_TAPS], tap_dx[MAX_TAPS];
    int tap_k[MAX_TAPS];  // K offset of the tap's weight block
    int b_rows;        // weight rows per TMA box (= MMA N)
    int lrelu;         // EPI_CONV: activation fused after the bias: 1 LeakyReLU(0.2), 2 ReLU
    int tiles_w, tiles_h, n_tiles, num_tiles;
    int cb0, cb_total;  // 64-channel blocks in source 0 / in total
    int passes;
    uint32_t idesc;
    int n_total;
    const float* w_inv_scale;
    const float* a_inv_scale;  // NULL, or the 2^-e of pre-scaled A planes (gradient operands)
    // EPI_CONV
    const float* bias;
    const float* residual;
    int res_ups;
    const float* rnoise[2];
    const float* rnoise_w[2];
    unsigned long long rnoise_seed[2];
    float* out;
    float* stats_partial;
    const __half* act_mask;  // dgrad: multiply by LeakyReLU'(t) read off the saved activation's sign
    float* amax_out;         // optional: max |out| (atomic, non-negative float bits)
    // EPI_MODULATE
    const float* x;
    int x_ups;
    const float* noise;
    const float* noise_w;
    unsigned long long noise_seed;
    const unsigned long long* noise_epoch;  // device counter folded into every noise seed
    const float* bn_scale;
    const float* bn_shift;
    const float* gamma_bias;
    const float* beta_bias;
    __half* out_hi;
    __half* out_lo;
    __half* g_hi;  // optional: G = gamma + gamma_bias saved for the backward pass (fp16 planes)
    __half* g_lo;
    uint8_t* out8_lo;  // optional: e5m2 planes of the activation for a passes == 2 consumer
    uint8_t* out8_hi;
    int C;
    // EPI_MODULATE_BWD
    const float* dt;         // fp32 NHWC [B,H,W,C]: gradient wrt the pre-activation t
    float* dxhat;            // fp32 NHWC [B,H,W,C]
    __half* dgb_hi;          // fp16 NHWC [B,H,W,2C] * dgb scale, channels interleaved [dG(128)|dB(128)]
    __half* dgb_lo;
    const float* dt_amax;    // device scalar: max |dt| (sets the dgb plane scale)
    float* dgb_inv_scale;    // out: 2^-e of the dgb planes
    float* bwd_partial;      // [m_tiles*4][C][4] = sum dxhat, sum dxhat*xhat, sum dG, sum dB
    // EPI_DGRAD_MODBWD (backward-data GEMM whose epilogue is K1's backward)
    const __half* gs_hi;     // saved G = gamma + gamma_bias planes (fp16 NHWC [B,H,W,C])
    const __half* gs_lo;
    const float* dy_amax;    // device scalar: max |dY| of the gradient operand
    const float* w_l1;       // device scalar: max over input channels of sum |W| (bounds |dt|)
};

__device__ __forceinline__ void decode_tile(const ConvParams& p, int tile, int& b, int& h0, int& w0,
                                            int& nt, int& cls) {
    nt = tile % p.n_tiles;
    int mt = tile / p.n_tiles;
    cls = 0;
    if (p.sub4) {  // the four classes of one pixel tile run back to back: x and the sources stay in L2
        cls = mt & 3;
        mt >>= 2;
    }
    int tw = mt % p.tiles_w;
    mt /= p.tiles_w;
    int th = mt % p.tiles_h;
    b = mt / p.tiles_h;
    h0 = th * (TILE_H << (p.m2 | p.pair));
    w0 = tw * TILE_W;
}

__device__ __forceinline__ void store_split8(__half* hi, __half* lo, const float* a) {
    // 8 consecutive channels -> one 16-byte store per plane
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_half2_sat(a[2 * j], a[2 * j + 1], ph[j], pl[j]);
    *reinterpret_cast<uint4*>(hi) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    if (lo) *reinterpret_cast<uint4*>(lo) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// Column sums of a 32(lanes) x 32(values) register tile: afterwards lane l holds sum_lanes v[l].
// 31 shuffles instead of 32*5 (recursive halving: each step trades half of the live values).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
            // keep j (lower lanes) or j+half (upper lanes); send the other one across
            float keep = upper ? v[j + half] : v[j];
            float send = upper ? v[j] : v[j + half];
            float recv = __shfl_xor_sync(0xffffffffu, send, half);
            v[j] = keep + recv;
        }
    }
    return v[0];
}

// CTA2 = true: the kernel runs as 2-CTA clusters (one TPC each).  A pair owns a 16 x 16-pixel tile: CTA
// r computes rows h0 + 8 r .. of it (its own A boxes, its own TMEM accumulator, its own epilogue) and
// stages rows [128 r, 128 r + 128) of the 256-row weight box; the leader issues ONE
// tcgen05.mma.cta_group::2 (M = 256, N = 256) per K step that reads both CTAs' shared memory.  Per
// CTA and K block 32 KB come from L2 instead of 48 KB.  Barrier protocol: both producers
// arrive.expect_tx on the LEADER's full barrier (count 2) and their TMA loads complete_tx there; the
// leader's commits are multicast to both CTAs' empty / accumulator-full barriers; the peer's epilogue
// warps release the accumulator on the leader's accumulator-empty barrier (count 2 x EPI_WARPS).
template <int EPI, bool CTA2 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle needs 1024-byte aligned tiles
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full_bar = bars;                  // [STAGES]
    uint64_t* empty_bar = bars + STAGES;        // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int crank = CTA2 ? (int)cluster_ctarank() : 0;      // 0 = leader
    const int tile0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tstep = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.tmA[i]);
        tma_prefetch_desc(&p.tmB[0]);
        tma_prefetch_desc(&p.tmB[1]);
        if (p.cb8) {
            tma_prefetch_desc(&p.tmA8[0]);
            tma_prefetch_desc(&p.tmA8[1]);
            tma_prefetch_desc(&p.tmB8);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], CTA2 ? 2 : 1);   // pair: one arrive.expect_tx per producer
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], CTA2 ? 2 * EPI_WARPS : EPI_WARPS);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CTA2) {
            tmem_alloc_pair(tmem_slot, TMEM_COLS);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_slot, TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CTA2) cluster_sync_all();   // the peer's barriers exist before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // fp16 K iterations, then (passes == 2) the fp8 correction: per tap 2 planes x cb8 blocks of 128 B
    const int k16 = p.passes * p.ntaps * p.cb_total;
    const int k_iters = k16 + p.ntaps * 2 * p.cb8;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = tile0; tile < p.num_tiles; tile += tstep) {
                int b, h0, w0, nt, cls;
                decode_tile(p, tile, b, h0, w0, nt, cls);
                if (CTA2) h0 += crank * TILE_H;   // my 8 rows of the pair's 16 x 16-pixel tile
                // (cls != 0 only in the sub4 form; pair: my half of the weight rows)
                const int n0 = nt * BLOCK_N + cls * p.n_total + (CTA2 ? crank * (BLOCK_N / 2) : 0);
                for (int pass = 0; pass < p.passes; ++pass) {
                    const int pa = (pass == 1) ? 1 : 0;
                    const int pb = (pass == 2) ? 1 : 0;
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        const int dy = p.tap_dy[cls * 4 + tap], dx = p.tap_dx[cls * 4 + tap];
                        for (int cb = 0; cb < p.cb_total; ++cb, ++it) {
                            const int s = it % STAGES;
                            const uint32_t ph = (it / STAGES) & 1;
                            mbar_wait(&empty_bar[s], ph ^ 1);
                            uint8_t* sa = smem + s * STAGE_BYTES;
                            uint8_t* sb = sa + (A_BYTES << p.m2);
                            const int src = (cb >= p.cb0) ? 1 : 0;
                            const int cl = src ? cb - p.cb0 : cb;
                            if (CTA2) {
                                mbar_expect_tx_leader(&full_bar[s], A_BYTES + p.b_rows * BLOCK_K * 2);
                                tma_load_4d_pair(&p.tmA[src * 2 + pa], &full_bar[s], sa, cl * BLOCK_K,
                                                 w0 * p.a_step + dx, h0 * p.a_step + dy, b);
                                tma_load_2d_pair(&p.tmB[pb], &full_bar[s], sb, p.tap_k[tap] + cb * BLOCK_K, n0);
                                continue;
                            }
                            mbar_expect_tx(&full_bar[s], (A_BYTES << p.m2) + p.b_rows * BLOCK_K * 2);
                            tma_load_4d(&p.tmA[src * 2 + pa], &full_bar[s], sa, cl * BLOCK_K,
                                        w0 * p.a_step + dx, h0 * p.a_step + dy, b);
                            tma_load_2d(&p.tmB[pb], &full_bar[s], sb, p.tap_k[tap] + cb * BLOCK_K,
                                        n0 + b * p.w_brows);
                        }
                    }
                }
                if (p.cb8) {
                    // fp8 correction: [a_lo * 2^8 | a] x [w * 2^-8 ; w_lo], same stage geometry in bytes
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        const int dy = p.tap_dy[tap], dx = p.tap_dx[tap];
                        for (int cb = 0; cb < 2 * p.cb8; ++cb, ++it) {
                            const int s = it % STAGES;
                            const uint32_t ph = (it / STAGES) & 1;
                            mbar_wait(&empty_bar[s], ph ^ 1);
                            uint8_t* sa = smem + s * STAGE_BYTES;
                            uint8_t* sb = sa + A_BYTES;
                            const int pl = cb >= p.cb8 ? 1 : 0;
                            if (CTA2) {
                                mbar_expect_tx_leader(&full_bar[s], A_BYTES + p.b_rows * 128);
                                tma_load_4d_pair(&p.tmA8[pl], &full_bar[s], sa, (cb - pl * p.cb8) * 128,
                                                 w0 * p.a_step + dx, h0 * p.a_step + dy, b);
                                tma_load_2d_pair(&p.tmB8, &full_bar[s], sb, (tap * 2 * p.cb8 + cb) * 128, n0);
                                continue;
                            }
                            mbar_expect_tx(&full_bar[s], A_BYTES + p.b_rows * 128);
                            tma_load_4d(&p.tmA8[pl], &full_bar[s], sa, (cb - pl * p.cb8) * 128,
                                        w0 * p.a_step + dx, h0 * p.a_step + dy, b);
                            tma_load_2d(&p.tmB8, &full_bar[s], sb, (tap * 2 * p.cb8 + cb) * 128, n0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && crank == 0) {   // pair: the leader issues for both CTAs
            uint32_t it = 0;
            int lt = 0;
            for (int tile = tile0; tile < p.num_tiles; tile += tstep, ++lt) {
                const int as = lt & 1;
                const uint32_t aph = (lt >> 1) & 1;
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BLOCK_N;
                for (int kit = 0; kit < k_iters; ++kit, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                    const uint32_t sb = sa + (A_BYTES << p.m2);
                    const uint64_t da = umma_desc_sw128(sa, 1024);
                    const uint64_t db = umma_desc_sw128(sb, 1024);
                    if (CTA2) {
                        if (kit < k16) {
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k)
                                umma_f16_pair(tmem_d, da + 2 * k, db + 2 * k, p.idesc, (kit | k) != 0);
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_f8_pair(tmem_d, da + 2 * k, db + 2 * k, p.idesc8, 1u);
                        }
                        umma_commit_pair(&empty_bar[s]);
                        continue;
                    }
                    if (kit < k16) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; ++k) {
                            // advance 16 elements (32 B) along K inside the swizzle atom: +2 (16 B units)
                            umma_f16(tmem_d, da + 2 * k, db + 2 * k, p.idesc, (kit | k) != 0);
                        }
                        if (p.m2) {  // second pixel half of the 16 x 16 tile: A rows 128..255, same weights
                            const uint64_t da2 = umma_desc_sw128(sa + A_BYTES, 1024);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k)
                                umma_f16(tmem_d + 128, da2 + 2 * k, db + 2 * k, p.idesc, (kit | k) != 0);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)  // 32 fp8 elements = the same 32 B per step
                            umma_f8(tmem_d, da + 2 * k, db + 2 * k, p.idesc8, 1u);
                    }
                    umma_commit(&empty_bar[s]);  // frees the smem stage when these MMAs retire
                }
                // accumulator complete -> epilogue (of both CTAs in the pair form)
                if (CTA2) umma_commit_pair(&tfull_bar[as]);
                else umma_commit(&tfull_bar[as]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may touch
        const int eh = (warp - 2) >> 2;  // which of the quarter's two warps (chunk parity)
        constexpr int ESTEP = EPI_WARPS / 4;
        const int m = q * 32 + lane;
        const int ly = m / TILE_W, lx = m % TILE_W;
        const float inv_scale = __ldg(p.w_inv_scale) * (p.a_inv_scale ? __ldg(p.a_inv_scale) : 1.f);
        const unsigned long long nseed = eff_noise_seed(p.noise_seed, p.noise_epoch);
        const unsigned long long rseed0 = eff_noise_seed(p.rnoise_seed[0], p.noise_epoch);
        const unsigned long long rseed1 = eff_noise_seed(p.rnoise_seed[1], p.noise_epoch);
        int lt = 0;
        for (int tile = tile0; tile < p.num_tiles; tile += tstep, ++lt) {
            const int as = lt & 1;
            const uint32_t aph = (lt >> 1) & 1;
            int b, h0, w0, nt, cls;
            decode_tile(p, tile, b, h0, w0, nt, cls);
            if (CTA2) h0 += crank * TILE_H;
            // pixel tile index for the per-tile statistics slots (pair: two 8-row halves per tile)
            const int ptile = CTA2 ? (tile / p.n_tiles) * 2 + crank : tile / p.n_tiles;
            const int y = h0 + ly, x = w0 + lx;
            const bool valid = (y < p.H) && (x < p.W);
            mbar_wait(&tfull_bar[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N;

            float* T = reinterpret_cast<float*>(bars + 32) + (warp - 2) * EPI_TILE_FLOATS;
            const int er = lane >> 3, ecq = lane & 7;  // transposed layout: pixel sub-row, channel quad
            if (EPI == EPI_CONV) {
                const int n0 = nt * BLOCK_N;
                const int h0t = h0;
                const int Hr = p.H >> p.res_ups, Wr = p.W >> p.res_ups;
                float tmax = 0.f;
#pragma unroll 1
                for (int ch = eh; ch < BLOCK_N / 32; ch += ESTEP) {
                    // m2: TMEM columns [0,128) hold the tile's upper 8 pixel rows, [128,256) the lower 8
                    const int sub = p.m2 ? (ch >> 2) : 0;
                    const int n = n0 + (p.m2 ? (ch & 3) : ch) * 32;
                    if (n >= p.n_total) {  // warp-uniform
                        if (p.m2) continue;
                        break;
                    }
                    const int h0 = h0t + sub * TILE_H;
                    {
                        uint32_t v[32];
                        tmem_ld32(taddr + ch * 32, v);
                        tmem_ld_wait();
                        tile_put_row(T, lane, v);
                    }
                    __syncwarp();
                    const int nc = n + ecq * 4;  // this lane's 4 channels
                    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + nc));
                    float4 nw0 = make_float4(0.f, 0.f, 0.f, 0.f), nw1 = nw0;
                    if (p.rnoise_w[0]) nw0 = __ldg(reinterpret_cast<const float4*>(p.rnoise_w[0] + nc));
                    if (p.rnoise_w[1]) nw1 = __ldg(reinterpret_cast<const float4*>(p.rnoise_w[1] + nc));
                    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
                    // residual / mask loads of all 8 steps are issued before the first use
                    float4 rq[8];
                    uint2 mq[8];
                    bool ok[8];
#pragma unroll
                    for (int st = 0; st < 8; ++st) {
                        const int mm = q * 32 + st * 4 + er;
                        const int yy = h0 + mm / TILE_W, xx = w0 + mm % TILE_W;
                        ok[st] = (yy < p.H) && (xx < p.W);
                        rq[st] = make_float4(0.f, 0.f, 0.f, 0.f);
                        mq[st] = make_uint2(0u, 0u);
                        if (ok[st]) {
                            if (p.residual) {
                                const size_t rp = ((size_t)b * Hr + (yy >> p.res_ups)) * Wr + (xx >> p.res_ups);
                                rq[st] = __ldg(reinterpret_cast<const float4*>(p.residual + rp * p.n_total + nc));
                            }
                            if (p.act_mask) {
                                const size_t pix = ((size_t)b * p.Hm + (yy * p.o_step + p.o_offy)) * p.Wm +
                                                   (xx * p.o_step + p.o_offx);
                                mq[st] = __ldg(reinterpret_cast<const uint2*>(p.act_mask + pix * p.n_total + nc));
                            }
                        }
                    }
#pragma unroll
                    for (int st = 0; st < 8; ++st) {
                        if (!ok[st]) continue;
                        const int pi = st * 4 + er;
                        const int mm = q * 32 + pi;
                        const int yy = h0 + mm / TILE_W, xx = w0 + mm % TILE_W;
                        const float4 t4 = tile_get4(T, pi, ecq);
                        float o[4] = {t4.x * inv_scale + bias4.x, t4.y * inv_scale + bias4.y,
                                      t4.z * inv_scale + bias4.z, t4.w * inv_scale + bias4.w};
                        if (p.lrelu) {
                            const float slope = p.lrelu == 2 ? 0.f : 0.2f;
#pragma unroll
                            for (int e = 0; e < 4; ++e) o[e] = o[e] > 0.f ? o[e] : slope * o[e];
                        }
                        const size_t pix = ((size_t)b * p.Hm + (yy * p.o_step + p.o_offy)) * p.Wm +
                                           (xx * p.o_step + p.o_offx);
                        const size_t oe = pix * p.n_total + nc;
                        if (p.act_mask) {
                            const uint2 mv = mq[st];
                            // fp16 > 0 <=> the half, moved to the top of a 32-bit word, is a positive integer
                            const bool pos[4] = {(int)(mv.x << 16) > 0, (int)(mv.x & 0xffff0000u) > 0,
                                                 (int)(mv.y << 16) > 0, (int)(mv.y & 0xffff0000u) > 0};
#pragma unroll
                            for (int e = 0; e < 4; ++e) o[e] *= pos[e] ? 1.f : 0.2f;
                        }
                        o[0] += rq[st].x; o[1] += rq[st].y; o[2] += rq[st].z; o[3] += rq[st].w;
                        if (p.rnoise_w[0]) {
                            const float4 r4 = load_noise4(p.rnoise[0], rseed0, oe);
                            o[0] += nw0.x * r4.x; o[1] += nw0.y * r4.y; o[2] += nw0.z * r4.z; o[3] += nw0.w * r4.w;
                        }
                        if (p.rnoise_w[1]) {
                            const float4 r4 = load_noise4(p.rnoise[1], rseed1, oe);
                            o[0] += nw1.x * r4.x; o[1] += nw1.y * r4.y; o[2] += nw1.z * r4.z; o[3] += nw1.w * r4.w;
                        }
                        *reinterpret_cast<float4*>(p.out + oe) = make_float4(o[0], o[1], o[2], o[3]);
                        if (p.out_hi) {  // leaky_relu(out) as fp16 planes for the tensor-core image head
                            uint32_t hh[2], ll[2];
                            split_half2_sat(fmaxf(o[0], 0.2f * o[0]), fmaxf(o[1], 0.2f * o[1]), hh[0], ll[0]);
                            split_half2_sat(fmaxf(o[2], 0.2f * o[2]), fmaxf(o[3], 0.2f * o[3]), hh[1], ll[1]);
                            *reinterpret_cast<uint2*>(p.out_hi + oe) = make_uint2(hh[0], hh[1]);
                            if (p.out_lo) *reinterpret_cast<uint2*>(p.out_lo + oe) = make_uint2(ll[0], ll[1]);
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            tmax = fmaxf(tmax, fabsf(o[e]));
                            s1[e] += o[e];
                            s2[e] += o[e] * o[e];
                        }
                    }
                    if (p.stats_partial) {
                        // tile partial of sum / sum^2 per channel, one slot per (m-tile, quarter):
                        // fold the 4 pixel sub-rows (lane bits 3, 4), lanes 0..7 write 4 channels each
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 8);
                            s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 8);
                            s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
                            s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 16);
                        }
                        if (lane < 8) {
                            const size_t slot = (size_t)ptile * 4 + q;
                            float4* sp = reinterpret_cast<float4*>(p.stats_partial + (slot * p.n_total + nc) * 2);
                            sp[0] = make_float4(s1[0], s2[0], s1[1], s2[1]);
                            sp[1] = make_float4(s1[2], s2[2], s1[3], s2[3]);
                        }
                    }
                    __syncwarp();  // T is rewritten by the next chunk
                }
                if (p.amax_out) {
#pragma unroll
                    for (int sft = 16; sft >= 1; sft >>= 1)
                        tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, sft));
                    if (lane == 0 && tmax > 0.f && !isinf(tmax) && !isnan(tmax)) atomic_max_nonneg(p.amax_out, tmax);
                }
            } else if (EPI == EPI_MODULATE_BWD) {
                // gamma-only GEMM (n_total == C): recompute G = gamma + gamma_bias, then
                //   dG = dt * xhat, dB = dt, dxhat = dt * G   (+ the per-channel sums BN backward
                //   and the bias gradients need)
                const int c0 = nt * BLOCK_N;
                const size_t pix = ((size_t)b * p.H + y) * p.W + x;
                const int Hx = p.H >> p.x_ups, Wx = p.W >> p.x_ups;
                const size_t xp = ((size_t)b * Hx + (y >> p.x_ups)) * Wx + (x >> p.x_ups);
                const float* xrow = p.x + xp * p.C;
                const float* nrow = p.noise ? p.noise + pix * p.C : nullptr;
                const float* dtrow = p.dt + pix * p.C;
                float* dxrow = p.dxhat + pix * p.C;
                __half* ghrow = p.dgb_hi + pix * (size_t)(2 * p.C);
                __half* glrow = p.dgb_lo ? p.dgb_lo + pix * (size_t)(2 * p.C) : nullptr;
                // |dG| = |dt * xhat| <= amax(dt) * |xhat|: leave 2^6 of headroom for |xhat|
                // (store_split8 clamps beyond it)
                const float gscale = pow2_scale_for(__ldg(p.dt_amax), 10);
                if (tile == 0 && threadIdx.x == 64) *p.dgb_inv_scale = 1.f / gscale;
#pragma unroll 1
                for (int ch = eh; ch < BLOCK_N / 32; ch += ESTEP) {
                    const int c = c0 + ch * 32;
                    if (c >= p.C) break;
                    uint32_t g[32];
                    tmem_ld32(taddr + ch * 32, g);
                    tmem_ld_wait();
                    float s_dx[32], s_dxx[32], s_dg[32], s_db[32];
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        float xs[4] = {0, 0, 0, 0}, dts[4] = {0, 0, 0, 0};
                        if (valid) {
                            float4 xv = __ldg(reinterpret_cast<const float4*>(xrow + c) + j4);
                            float4 dv = __ldg(reinterpret_cast<const float4*>(dtrow + c) + j4);
                            xs[0] = xv.x; xs[1] = xv.y; xs[2] = xv.z; xs[3] = xv.w;
                            dts[0] = dv.x; dts[1] = dv.y; dts[2] = dv.z; dts[3] = dv.w;
                            if (nrow) {
                                float4 nv = __ldg(reinterpret_cast<const float4*>(nrow + c) + j4);
                                float4 wv = __ldg(reinterpret_cast<const float4*>(p.noise_w + c) + j4);
                                xs[0] += wv.x * nv.x; xs[1] += wv.y * nv.y;
                                xs[2] += wv.z * nv.z; xs[3] += wv.w * nv.w;
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = j4 * 4 + e, cc = c + j;
                            const float xh = xs[e] * __ldg(p.bn_scale + cc) + __ldg(p.bn_shift + cc);
                            const float G = __uint_as_float(g[j]) * inv_scale + __ldg(p.gamma_bias + cc);
                            const float dtv = dts[e];
                            s_dg[j] = valid ? dtv * xh : 0.f;
                            s_db[j] = dtv;
                            s_dx[j] = dtv * G;
                            s_dxx[j] = s_dx[j] * xh;
                        }
                    }
                    if (valid) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4)
                            reinterpret_cast<float4*>(dxrow + c)[j4] = make_float4(
                                s_dx[4 * j4], s_dx[4 * j4 + 1], s_dx[4 * j4 + 2], s_dx[4 * j4 + 3]);
                        // interleaved channel position of c: (c/128)*256 + c%128 (+128 for dB)
                        const int ng = (c >> 7) * 256 + (c & 127);
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const float* src = half ? s_db : s_dg;
                            __half* dh = ghrow + ng + half * 128;
                            __half* dl = glrow ? glrow + ng + half * 128 : nullptr;
#pragma unroll
                            for (int j8 = 0; j8 < 4; ++j8) {
                                float a8[8];
#pragma unroll
                                for (int e = 0; e < 8; ++e) a8[e] = src[j8 * 8 + e] * gscale;
                                store_split8(dh + j8 * 8, dl ? dl + j8 * 8 : nullptr, a8);
                            }
                        }
                    }
                    // tile partials (invalid pixels contribute zeros: dt was loaded as 0)
                    const float r0 = warp_transpose_reduce(s_dx, lane);
                    const float r1 = warp_transpose_reduce(s_dxx, lane);
                    const float r2 = warp_transpose_reduce(s_dg, lane);
                    const float r3 = warp_transpose_reduce(s_db, lane);
                    const size_t slot = (size_t)ptile * 4 + q;
                    *reinterpret_cast<float4*>(p.bwd_partial + (slot * p.C + c + lane) * 4) =
                        make_float4(r0, r1, r2, r3);
                }
            } else if (EPI == EPI_DGRAD_MODBWD) {
                // accumulator = conv_transpose(dY, W) for channels c0..c0+255 of the tile's pixels:
                //   dt = acc * lrelu'(t)   (sign of the saved activation),
                //   dxhat = dt * G, dG = dt * xhat, dB = dt  (+ the 4 per-channel tile sums)
                const int c0 = nt * BLOCK_N;
                const int Hx = p.H >> p.x_ups, Wx = p.W >> p.x_ups;
                const bool has_noise = p.noise_w != nullptr;
                // |dt| <= max|dY| * max_c sum_{n,tap} |W[n,c,tap]|: a bound known before the GEMM runs,
                // so the planes can be scaled here (2^6 of headroom for |xhat|, the split clamps)
                const float gscale = pow2_scale_for(__ldg(p.dy_amax) * __ldg(p.w_l1), 10);
                if (tile == 0 && threadIdx.x == 64) *p.dgb_inv_scale = 1.f / gscale;
#pragma unroll 1
                for (int ch = eh; ch < BLOCK_N / 32; ch += ESTEP) {
                    const int c = c0 + ch * 32;
                    if (c >= p.C) break;
                    {
                        uint32_t v[32];
                        tmem_ld32(taddr + ch * 32, v);
                        tmem_ld_wait();
                        tile_put_row(T, lane, v);
                    }
                    __syncwarp();
                    const int cc = c + ecq * 4;  // this lane's 4 channels
                    const float4 sc4 = __ldg(reinterpret_cast<const float4*>(p.bn_scale + cc));
                    const float4 sh4 = __ldg(reinterpret_cast<const float4*>(p.bn_shift + cc));
                    float4 nw4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (has_noise) nw4 = __ldg(reinterpret_cast<const float4*>(p.noise_w + cc));
                    const int ng = (cc >> 7) * 256 + (cc & 127);  // interleaved [dG(128) | dB(128)] position
                    float sm[4][4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int e = 0; e < 4; ++e) sm[k][e] = 0.f;
                    // all global loads of the chunk's 8 steps are issued before the first use (one
                    // memory round trip per chunk instead of eight)
                    float4 xq[8];
                    uint2 mq[8], gq[8], lq[8];
                    bool ok[8];
#pragma unroll
                    for (int st = 0; st < 8; ++st) {
                        const int mm = q * 32 + st * 4 + er;
                        const int yy = h0 + mm / TILE_W, xx = w0 + mm % TILE_W;
                        ok[st] = (yy < p.H) && (xx < p.W);
                        xq[st] = make_float4(0.f, 0.f, 0.f, 0.f);
                        mq[st] = gq[st] = lq[st] = make_uint2(0u, 0u);
                        if (ok[st]) {
                            const size_t pe = (((size_t)b * p.H + yy) * p.W + xx) * p.C + cc;
                            const size_t xp = ((size_t)b * Hx + (yy >> p.x_ups)) * Wx + (xx >> p.x_ups);
                            xq[st] = __ldg(reinterpret_cast<const float4*>(p.x + xp * p.C + cc));
                            mq[st] = __ldg(reinterpret_cast<const uint2*>(p.act_mask + pe));
                            gq[st] = __ldg(reinterpret_cast<const uint2*>(p.gs_hi + pe));
                            if (p.gs_lo) lq[st] = __ldg(reinterpret_cast<const uint2*>(p.gs_lo + pe));
                        }
                    }
#pragma unroll
                    for (int st = 0; st < 8; ++st) {
                        if (!ok[st]) continue;
                        const int pi = st * 4 + er;
                        const int mm = q * 32 + pi;
                        const int yy = h0 + mm / TILE_W, xx = w0 + mm % TILE_W;
                        const size_t pix = ((size_t)b * p.H + yy) * p.W + xx;
                        const size_t pe = pix * p.C + cc;
                        float4 xv = xq[st];
                        const uint2 mv = mq[st], gv = gq[st], lv = lq[st];
                        if (has_noise) {
                            const float4 nv = load_noise4(p.noise, nseed, pe);
                            xv.x += nw4.x * nv.x; xv.y += nw4.y * nv.y; xv.z += nw4.z * nv.z; xv.w += nw4.w * nv.w;
                        }
                        const float xh[4] = {xv.x * sc4.x + sh4.x, xv.y * sc4.y + sh4.y, xv.z * sc4.z + sh4.z,
                                             xv.w * sc4.w + sh4.w};
                        // fp16 > 0 <=> the half, moved to the top of a 32-bit word, is a positive integer
                        const bool pos[4] = {(int)(mv.x << 16) > 0, (int)(mv.x & 0xffff0000u) > 0,
                                             (int)(mv.y << 16) > 0, (int)(mv.y & 0xffff0000u) > 0};
                        const float2 ga = unpack_half2(gv.x), gb2 = unpack_half2(gv.y);
                        const float2 la = unpack_half2(lv.x), lb2 = unpack_half2(lv.y);
                        const float Gs[4] = {ga.x + la.x, ga.y + la.y, gb2.x + lb2.x, gb2.y + lb2.y};
                        const float4 t4 = tile_get4(T, pi, ecq);
                        const float tp[4] = {t4.x, t4.y, t4.z, t4.w};
                        float d[4], dxh[4], dG[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            d[e] = tp[e] * inv_scale * (pos[e] ? 1.f : 0.2f);
                            dxh[e] = d[e] * Gs[e];
                            dG[e] = d[e] * xh[e];
                            sm[0][e] += dxh[e];
                            sm[1][e] += dxh[e] * xh[e];
                            sm[2][e] += dG[e];
                            sm[3][e] += d[e];
                        }
                        *reinterpret_cast<float4*>(p.dxhat + pe) = make_float4(dxh[0], dxh[1], dxh[2], dxh[3]);
                        __half* rowh = p.dgb_hi + pix * (size_t)(2 * p.C) + ng;
                        __half* rowl = p.dgb_lo ? p.dgb_lo + pix * (size_t)(2 * p.C) + ng : nullptr;
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const float* src = half ? d : dG;
                            uint32_t ph[2], plw[2];
                            if (rowl) {
                                split_half2_sat(src[0] * gscale, src[1] * gscale, ph[0], plw[0]);
                                split_half2_sat(src[2] * gscale, src[3] * gscale, ph[1], plw[1]);
                                *reinterpret_cast<uint2*>(rowl + half * 128) = make_uint2(plw[0], plw[1]);
                            } else {
                                ph[0] = pack_half2_sat(src[0] * gscale, src[1] * gscale);
                                ph[1] = pack_half2_sat(src[2] * gscale, src[3] * gscale);
                            }
                            *reinterpret_cast<uint2*>(rowh + half * 128) = make_uint2(ph[0], ph[1]);
                        }
                    }
                    // fold the 4 pixel sub-rows; lanes 0..7 write their 4 channels x 4 sums
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            sm[k][e] += __shfl_xor_sync(0xffffffffu, sm[k][e], 8);
                            sm[k][e] += __shfl_xor_sync(0xffffffffu, sm[k][e], 16);
                        }
                    if (lane < 8) {
                        const size_t slot = (size_t)ptile * 4 + q;
                        float4* bp = reinterpret_cast<float4*>(p.bwd_partial + (slot * p.C + cc) * 4);
#pragma unroll
                        for (int e = 0; e < 4; ++e) bp[e] = make_float4(sm[0][e], sm[1][e], sm[2][e], sm[3][e]);
                    }
                    __syncwarp();  // T is rewritten by the next chunk
                }
            } else {
                // EPI_MODULATE: columns [0,128) gamma, [128,256) beta for channels nt*128 + j.
                // Both 32 x 32 chunks go through the transpose tile (gamma first, kept in registers).
                const int c0 = nt * 128;
                // tile pixel (yy, xx) lives at (yy * o_step + o_offy, xx * o_step + o_offx) of the [Hm, Wm]
                // output (o_step = 2: one parity class of the sub-pixel form; 1 otherwise)
                const int Hx = p.Hm >> p.x_ups, Wx = p.Wm >> p.x_ups;
                const int ooy = p.sub4 ? (cls >> 1) : p.o_offy, oox = p.sub4 ? (cls & 1) : p.o_offx;
                const bool has_noise = p.noise_w != nullptr;
#pragma unroll 1
                for (int ch = eh; ch < 4; ch += ESTEP) {
                    const int c = c0 + ch * 32;
                    float gt[8][4];
                    {
                        uint32_t v[32];
                        tmem_ld32(taddr + ch * 32, v);
                        tmem_ld_wait();
                        tile_put_row(T, lane, v);
                        __syncwarp();
#pragma unroll
                        for (int st = 0; st < 8; ++st) {
                            const float4 g4 = tile_get4(T, st * 4 + er, ecq);
                            gt[st][0] = g4.x; gt[st][1] = g4.y; gt[st][2] = g4.z; gt[st][3] = g4.w;
                        }
                        __syncwarp();
                        tmem_ld32(taddr + 128 + ch * 32, v);
                        tmem_ld_wait();
                        tile_put_row(T, lane, v);
                        __syncwarp();
                    }
                    const int cc = c + ecq * 4;  // this lane's 4 channels
                    const float4 sc4 = __ldg(reinterpret_cast<const float4*>(p.bn_scale + cc));
                    const float4 sh4 = __ldg(reinterpret_cast<const float4*>(p.bn_shift + cc));
                    const float4 gb4 = __ldg(reinterpret_cast<const float4*>(p.gamma_bias + cc));
                    const float4 bb4 = __ldg(reinterpret_cast<const float4*>(p.beta_bias + cc));
                    float4 nw4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (has_noise) nw4 = __ldg(reinterpret_cast<const float4*>(p.noise_w + cc));
                    const float scv[4] = {sc4.x, sc4.y, sc4.z, sc4.w}, shv[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
                    const float gbv[4] = {gb4.x, gb4.y, gb4.z, gb4.w}, bbv[4] = {bb4.x, bb4.y, bb4.z, bb4.w};
                    float4 xq[8];
#pragma unroll
                    for (int st = 0; st < 8; ++st) {  // x loads of all 8 steps issued up front
                        const int mm = q * 32 + st * 4 + er;
                        const int yy = h0 + mm / TILE_W, xx = w0 + mm % TILE_W;
                        xq[st] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (yy < p.H && xx < p.W) {
                            const int fy = yy * p.o_step + ooy, fx = xx * p.o_step + oox;
                            const size_t xp = ((size_t)b * Hx + (fy >> p.x_ups)) * Wx + (fx >> p.x_ups);
                            xq[st] = __ldg(reinterpret_cast<const float4*>(p.x + xp * p.C + cc));
                        }
                    }
#pragma unroll
                    for (int st = 0; st < 8; ++st) {
                        const int pi = st * 4 + er;
                        const int mm = q * 32 + pi;
                        const int yy = h0 + mm / TILE_W, xx = w0 + mm % TILE_W;
                        if (yy >= p.H || xx >= p.W) continue;
                        const size_t pix = ((size_t)b * p.Hm + (yy * p.o_step + ooy)) * p.Wm + (xx * p.o_step + oox);
                        const size_t pe = pix * p.C + cc;
                        float4 xv = xq[st];
                        if (has_noise) {
                            const float4 nv = load_noise4(p.noise, nseed, pe);
                            xv.x += nw4.x * nv.x; xv.y += nw4.y * nv.y; xv.z += nw4.z * nv.z; xv.w += nw4.w * nv.w;
                        }
                        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
                        const float4 t4 = tile_get4(T, pi, ecq);
                        const float tp[4] = {t4.x, t4.y, t4.z, t4.w};
                        float t[4], Gv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float xh = xs[e] * scv[e] + shv[e];
                            Gv[e] = gt[st][e] * inv_scale + gbv[e];
                            const float Bv = tp[e] * inv_scale + bbv[e];
                            const float tt = xh * Gv[e] + Bv;
                            t[e] = fmaxf(tt, 0.2f * tt);   // LeakyReLU(0.2)
                        }
                        // packed saturating conversions (pack_half2_sat): hi = fp16(t), lo = fp16(t - hi)
                        uint32_t ah[2], al[2] = {0u, 0u}, gh[2] = {0u, 0u}, gl[2] = {0u, 0u};
                        uint32_t q_lo = 0u, q_hi = 0u;
                        ah[0] = pack_half2_sat(t[0], t[1]);
                        ah[1] = pack_half2_sat(t[2], t[3]);
                        if (p.out_lo || p.out8_hi) {
                            const float2 f0 = unpack_half2(ah[0]), f1 = unpack_half2(ah[1]);
                            const float r[4] = {t[0] - f0.x, t[1] - f0.y, t[2] - f1.x, t[3] - f1.y};
                            if (p.out_lo) {
                                al[0] = pack_half2_sat(r[0], r[1]);
                                al[1] = pack_half2_sat(r[2], r[3]);
                            }
                            if (p.out8_hi) {
                                q_lo = pack_e5m2x2_sat(r[0] * 256.f, r[1] * 256.f) |
                                       (pack_e5m2x2_sat(r[2] * 256.f, r[3] * 256.f) << 16);
                                q_hi = pack_e5m2x2_sat(t[0], t[1]) | (pack_e5m2x2_sat(t[2], t[3]) << 16);
                            }
                        }
                        if (p.g_hi) {
                            if (p.g_lo) {
                                split_half2_sat(Gv[0], Gv[1], gh[0], gl[0]);
                                split_half2_sat(Gv[2], Gv[3], gh[1], gl[1]);
                            } else {
                                gh[0] = pack_half2_sat(Gv[0], Gv[1]);
                                gh[1] = pack_half2_sat(Gv[2], Gv[3]);
                            }
                        }
                        *reinterpret_cast<uint2*>(p.out_hi + pe) = make_uint2(ah[0], ah[1]);
                        if (p.out_lo) *reinterpret_cast<uint2*>(p.out_lo + pe) = make_uint2(al[0], al[1]);
                        if (p.g_hi) *reinterpret_cast<uint2*>(p.g_hi + pe) = make_uint2(gh[0], gh[1]);
                        if (p.g_lo) *reinterpret_cast<uint2*>(p.g_lo + pe) = make_uint2(gl[0], gl[1]);
                        if (p.out8_hi) {
                            *reinterpret_cast<uint32_t*>(p.out8_lo + pe) = q_lo;
                            *reinterpret_cast<uint32_t*>(p.out8_hi + pe) = q_hi;
                        }
                    }
                    __syncwarp();  // T is rewritten by the next chunk
                }
            }
            // all TMEM reads of this accumulator stage are complete (wait::ld above)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTA2) mbar_arrive_leader(&tempty_bar[as]);
                else mbar_arrive(&tempty_bar[as]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CTA2) cluster_sync_all();   // nobody leaves while the peer can still signal or read this CTA
    if (warp == 1) {
        if (CTA2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
        else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int fill_common(ConvParams& p, const dsee_conv_operands* ops, bool allow_f8 = false,
                       bool allow_sub = false, bool m2 = false, bool pair = false) {
    DSEE_CHECK_ARG(ops != nullptr, "conv operands are NULL");
    DSEE_CHECK_ARG(ops->B > 0 && ops->H > 0 && ops->W > 0, "bad geometry B=%d H=%d W=%d", ops->B,
                   ops->H, ops->W);
    DSEE_CHECK_ARG(ops->passes >= 1 && ops->passes <= 3, "passes must be 1, 2 or 3 (got %d)",
                   ops->passes);
    if (ops->passes == 2) {
        DSEE_CHECK_ARG(allow_f8, "passes == 2 (fp8 correction) is implemented for dsee_conv3x3_fwd only");
        DSEE_CHECK_ARG(ops->a8_lo && ops->a8_hi && ops->w8, "passes == 2 needs a8_lo, a8_hi and w8");
        DSEE_CHECK_ARG(ops->a_channels[1] == 0 && ops->a_channels[0] % 128 == 0 && ops->a_dtype == 0,
                       "passes == 2 needs one fp16 A source with a multiple of 128 channels");
    }
    DSEE_CHECK_ARG(ops->a_channels[0] > 0 && ops->a_channels[0] % BLOCK_K == 0 &&
                       ops->a_channels[1] >= 0 && ops->a_channels[1] % BLOCK_K == 0,
                   "A channel counts must be multiples of %d (got %d, %d)", BLOCK_K,
                   ops->a_channels[0], ops->a_channels[1]);
    DSEE_CHECK_ARG(ops->a_hi[0] && ops->w_hi && ops->w_inv_scale, "NULL operand pointer");
    DSEE_CHECK_ARG(ops->a_channels[1] == 0 || ops->a_hi[1], "second A source is NULL");
    if (ops->passes == 3) {
        DSEE_CHECK_ARG(ops->a_lo[0] && ops->w_lo && (ops->a_channels[1] == 0 || ops->a_lo[1]),
                       "passes=3 needs the lo planes");
    }
    DSEE_CHECK_ARG(ops->n_total > 0 && ops->n_total % 32 == 0, "n_total must be a multiple of 32");
    int rc = require_sm100();
    if (rc) return rc;

    p.B = ops->B;
    p.H = ops->H;
    p.W = ops->W;
    p.Hm = ops->H;
    p.Wm = ops->W;
    p.o_step = 1;
    p.a_step = 1;
    p.ntaps = 9;
    for (int t = 0; t < 9; ++t) {
        p.tap_dy[t] = (int8_t)(t / 3 - 1);
        p.tap_dx[t] = (int8_t)(t % 3 - 1);
        p.tap_k[t] = t * (ops->a_channels[0] + ops->a_channels[1]);
    }
    // MMA N = weight rows per TMA box: all of BLOCK_N, or the (16-aligned) n_total when it is smaller
    p.b_rows = ops->n_total < BLOCK_N ? (ops->n_total + 15) / 16 * 16 : BLOCK_N;
    DSEE_CHECK_ARG(ops->w_batch_rows == 0 || ops->w_batch_rows == ops->n_total,
                   "w_batch_rows must be 0 or n_total");
    DSEE_CHECK_ARG(ops->w_batch_rows == 0 || ops->n_total % BLOCK_N == 0,
                   "per-image weights need n_total to be a multiple of %d", BLOCK_N);
    p.w_brows = ops->w_batch_rows;
    p.noise_epoch = noise_epoch_ptr();
    p.m2 = m2 ? 1 : 0;
    p.pair = pair ? 1 : 0;
    p.tiles_w = (ops->W + TILE_W - 1) / TILE_W;
    p.tiles_h = (ops->H + (TILE_H << (p.m2 | p.pair)) - 1) / (TILE_H << (p.m2 | p.pair));
    p.n_tiles = (ops->n_total + BLOCK_N - 1) / BLOCK_N;
    p.num_tiles = p.B * p.tiles_h * p.tiles_w * p.n_tiles;
    p.cb0 = ops->a_channels[0] / BLOCK_K;
    p.cb_total = (ops->a_channels[0] + ops->a_channels[1]) / BLOCK_K;
    p.passes = ops->passes == 2 ? 1 : ops->passes;  // fp16 passes; passes == 2 adds the fp8 phase
    p.n_total = ops->n_total;
    p.w_inv_scale = ops->w_inv_scale;
    p.a_inv_scale = ops->a_inv_scale;
    // kind::f16 instruction descriptor: fp32 accumulate, fp16 A/B, K-major both, N=256, M=128
    DSEE_CHECK_ARG((ops->a_dtype == 0 || ops->a_dtype == 1) && ops->w_dtype == ops->a_dtype,
                   "operand dtypes must both be 0 (fp16) or both 1 (bf16): tcgen05 kind::f16 rejects "
                   "mixed A/B formats");
    p.idesc = (1u << 4) | ((uint32_t)ops->a_dtype << 7) | ((uint32_t)ops->w_dtype << 10) |
              ((uint32_t)(p.b_rows >> 3) << 17) | ((uint32_t)((BLOCK_M << p.pair) >> 4) << 24);

    const int sub = ops->a_sub ? 1 : 0;
    if (sub) {
        DSEE_CHECK_ARG(allow_sub, "a_sub (sub-pixel form) is implemented for dsee_spade_modulate_fwd only");
        const bool all4 = ops->sub_py < 0;
        DSEE_CHECK_ARG(ops->H % 2 == 0 && ops->W % 2 == 0 &&
                           (all4 || ((ops->sub_py | 1) == 1 && (ops->sub_px | 1) == 1)) && ops->passes != 2,
                       "sub-pixel form needs even H, W, a parity class in {0,1}^2 (or sub_py < 0 = all "
                       "four in one launch) and passes 1 or 3");
        // tile space = the class's output pixels = the half-resolution grid
        p.H = ops->H / 2;
        p.W = ops->W / 2;
        p.o_step = 2;
        p.o_offy = all4 ? 0 : ops->sub_py;
        p.o_offx = all4 ? 0 : ops->sub_px;
        p.sub4 = all4 ? 1 : 0;
        p.ntaps = 4;
        const int ctot = ops->a_channels[0] + ops->a_channels[1];
        for (int c = 0; c < (all4 ? 4 : 1); ++c) {
            const int py = all4 ? (c >> 1) : ops->sub_py, px = all4 ? (c & 1) : ops->sub_px;
            for (int t = 0; t < 4; ++t) {
                p.tap_dy[c * 4 + t] = (int8_t)(t / 2 + py - 1);
                p.tap_dx[c * 4 + t] = (int8_t)(t % 2 + px - 1);
                p.tap_k[c * 4 + t] = t * ctot;
            }
        }
        p.tiles_w = (p.W + TILE_W - 1) / TILE_W;
        p.tiles_h = (p.H + TILE_H - 1) / TILE_H;
        p.num_tiles = p.B * p.tiles_h * p.tiles_w * p.n_tiles * (all4 ? 4 : 1);
    }
    const int Ha = ops->H >> sub, Wa = ops->W >> sub;   // resolution of the A planes
    for (int src = 0; src < 2; ++src) {
        const int C = ops->a_channels[src];
        for (int pl = 0; pl < 2; ++pl) {
            const void* base = pl ? ops->a_lo[src] : ops->a_hi[src];
            if (C == 0 || base == nullptr) {
                // unused slot: alias a valid map so the prefetch stays legal
                p.tmA[src * 2 + pl] = p.tmA[0];
                continue;
            }
            uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wa, (uint64_t)Ha, (uint64_t)ops->B};
            uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)Wa * C * 2, (uint64_t)Ha * Wa * C * 2};
            uint32_t box[4] = {BLOCK_K, TILE_W, (uint32_t)(TILE_H << p.m2), 1};
            rc = encode_tmap_16b(&p.tmA[src * 2 + pl], base, 4, dims, strides, box, ops->a_dtype == 1);
            if (rc) return rc;
        }
    }
    if (ops->passes == 2) {
        const int C = ops->a_channels[0];
        p.cb8 = C / 128;
        // kind::f8f6f4: fp32 accumulate, A = e5m2, B = e4m3, K-major both
        p.idesc8 = (1u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(p.b_rows >> 3) << 17) |
                   ((uint32_t)((BLOCK_M << p.pair) >> 4) << 24);
        for (int pl = 0; pl < 2; ++pl) {
            uint64_t dims[4] = {(uint64_t)C, (uint64_t)ops->W, (uint64_t)ops->H, (uint64_t)ops->B};
            uint64_t strides[3] = {(uint64_t)C, (uint64_t)ops->W * C, (uint64_t)ops->H * ops->W * C};
            uint32_t box[4] = {128, TILE_W, TILE_H, 1};
            rc = encode_tmap_8b(&p.tmA8[pl], pl ? ops->a8_hi : ops->a8_lo, 4, dims, strides, box);
            if (rc) return rc;
        }
        uint64_t dims[2] = {(uint64_t)18 * C, (uint64_t)ops->n_total};
        uint64_t strides[1] = {(uint64_t)18 * C};
        uint32_t box[2] = {128, (uint32_t)(p.b_rows >> p.pair)};
        rc = encode_tmap_8b(&p.tmB8, ops->w8, 2, dims, strides, box);
        if (rc) return rc;
    }
    const uint64_t Ktot = (uint64_t)(sub ? 4 : 9) * (ops->a_channels[0] + ops->a_channels[1]);
    // the weight matrix is padded by the caller to a multiple of BLOCK_N rows? No: TMA zero-fills
    // rows >= n_total, and the epilogue never stores those columns.
    for (int pl = 0; pl < 2; ++pl) {
        const void* base = pl ? ops->w_lo : ops->w_hi;
        if (!base) {
            p.tmB[pl] = p.tmB[0];
            continue;
        }
        uint64_t dims[2] = {Ktot, (uint64_t)ops->n_total * (ops->w_batch_rows ? ops->B : 1) * (p.sub4 ? 4 : 1)};
        uint64_t strides[1] = {Ktot * 2};
        uint32_t box[2] = {BLOCK_K, (uint32_t)(p.b_rows >> p.pair)};
        rc = encode_tmap_16b(&p.tmB[pl], base, 2, dims, strides, box, ops->w_dtype == 1);
        if (rc) return rc;
    }
    if (p.pair) p.b_rows >>= 1;   // from here on: weight rows staged per CTA (the descriptors hold N = 256)
    return 0;
}

// CTA-pair launch: 2-CTA clusters, one per TPC
static int g_cta_pair = 0;   // dsee_conv_pair_mode

template <int EPI>
static int launch_pair(const ConvParams& p, cudaStream_t stream) {
    static bool configured[64] = {false};
    static int max_clusters[64] = {0};
    int dev = 0;
    DSEE_CUDA(cudaGetDevice(&dev));
    DSEE_CHECK_ARG(dev < 64, "device index %d out of range", dev);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (!configured[dev]) {
        DSEE_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<EPI, true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        int sms = 0;
        DSEE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        cfg.gridDim = dim3(sms / 2 * 2, 1, 1);
        int nc = 0;
        DSEE_CUDA(cudaOccupancyMaxActiveClusters(&nc, conv3x3_tc_kernel<EPI, true>, &cfg));
        DSEE_CHECK_ARG(nc > 0, "no 2-CTA cluster of the pair kernel fits on this device");
        max_clusters[dev] = nc < sms / 2 ? nc : sms / 2;
        configured[dev] = true;
    }
    const int clusters = p.num_tiles < max_clusters[dev] ? p.num_tiles : max_clusters[dev];
    cfg.gridDim = dim3(2 * clusters, 1, 1);
    DSEE_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<EPI, true>, p));
    count_launch();
    return 0;
}

template <int EPI>
static int launch(const ConvParams& p, cudaStream_t stream) {
    if constexpr (EPI == EPI_CONV) {
        if (p.pair) return launch_pair<EPI>(p, stream);
    }
    static bool configured[64] = {false};
    int dev = 0;
    DSEE_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        DSEE_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<EPI>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured[dev] = true;
    }
    int sms = 0;
    DSEE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = p.num_tiles < sms ? p.num_tiles : sms;
    conv3x3_tc_kernel<EPI><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(p);
    count_launch();
    DSEE_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace dsee

using namespace dsee;

extern "C" int dsee_conv_pair_mode(int on) {
    const int prev = g_cta_pair;
    if (on >= 0) g_cta_pair = on ? 1 : 0;
    return prev;
}

extern "C" int dsee_conv3x3_stats_tiles(int B, int H, int W) {
    return B * ((H + TILE_H - 1) / TILE_H) * ((W + TILE_W - 1) / TILE_W) * 4;
}

extern "C" int dsee_conv3x3_fwd(const dsee_conv_operands* ops, const dsee_conv_epilogue* epi,
                                void* stream) {
    ConvParams p;
    memset(&p, 0, sizeof(p));
    DSEE_CHECK_ARG(epi && epi->out, "epilogue/out is NULL");
    // narrow outputs (the modulation's backward-data GEMM, N = 128): 16 x 16-pixel tiles
    const bool m2 = ops && ops->n_total <= 128 && !epi->stats_partial && ops->passes != 2 &&
                    ops->w_batch_rows == 0;
    // wide outputs on 16-row-aligned images: CTA pairs sharing the weight box (opt-in)
    const bool pair = g_cta_pair && ops && ops->n_total % BLOCK_N == 0 && ops->H % (2 * TILE_H) == 0 &&
                      ops->w_batch_rows == 0 && !ops->a_sub;
    int rc = fill_common(p, ops, true, false, m2, pair);
    if (rc) return rc;
    DSEE_CHECK_ARG(epi->res_ups == 0 || epi->res_ups == 1, "res_ups must be 0 or 1");
    DSEE_CHECK_ARG(!epi->residual || epi->res_ups == 0 || (ops->H % 2 == 0 && ops->W % 2 == 0),
                   "folded upsample needs even H, W");
    for (int i = 0; i < 2; ++i) {
        DSEE_CHECK_ARG((epi->noise[i] != nullptr || epi->noise_seed[i] != 0) == (epi->noise_w[i] != nullptr),
                       "noise[%d] (tensor or seed) and noise_w[%d] must be given together", i, i);
        p.rnoise[i] = epi->noise[i];
        p.rnoise_w[i] = epi->noise_w[i];
        p.rnoise_seed[i] = epi->noise_seed[i];
    }
    p.bias = epi->bias;
    p.residual = epi->residual;
    p.res_ups = epi->res_ups;
    p.out = epi->out;
    p.stats_partial = epi->stats_partial;
    p.act_mask = (const __half*)epi->act_mask;
    p.lrelu = epi->lrelu;
    p.amax_out = epi->amax_out;
    DSEE_CHECK_ARG(epi->act16_hi || !epi->act16_lo, "act16_lo without act16_hi");
    p.out_hi = (__half*)epi->act16_hi;
    p.out_lo = (__half*)epi->act16_lo;
    if (p.amax_out) DSEE_CUDA(cudaMemsetAsync(p.amax_out, 0, sizeof(float), (cudaStream_t)stream));
    return launch<EPI_CONV>(p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// general KH x KW / strided convolution and its backward-data on the same kernel (EPI_CONV)
// ------------------------------------------------------------------------------------------------
static int round_up(int a, int b) { return (a + b - 1) / b * b; }

extern "C" int dsee_conv2d_tc(const dsee_conv2d_tc_args* a, const dsee_conv_epilogue* epi, void* stream) {
    DSEE_CHECK_ARG(a && epi && epi->out, "NULL argument");
    DSEE_CHECK_ARG(a->B > 0 && a->Hi > 0 && a->Wi > 0 && a->Ho > 0 && a->Wo > 0, "bad geometry");
    DSEE_CHECK_ARG(a->Ci > 0 && a->Ci % 8 == 0, "A channels must be a multiple of 8 (got %d)", a->Ci);
    DSEE_CHECK_ARG(a->KH > 0 && a->KW > 0 && a->KH * a->KW <= MAX_TAPS, "at most %d filter taps", MAX_TAPS);
    DSEE_CHECK_ARG(a->stride == 1 || a->stride == 2, "stride must be 1 or 2");
    DSEE_CHECK_ARG(a->passes == 1 || (a->passes == 3 && a->a_lo && a->w_lo), "passes must be 1, or 3 with lo planes");
    DSEE_CHECK_ARG(a->a_hi && a->w_hi && a->w_inv_scale, "NULL operand pointer");
    DSEE_CHECK_ARG(a->n_total > 0 && a->n_total % 32 == 0, "n_total must be a multiple of 32");
    DSEE_CHECK_ARG(!epi->residual && !epi->noise_w[0] && !epi->noise_w[1] && !epi->stats_partial &&
                       (!epi->act_mask || a->stride == 1),
                   "epilogue option not supported by the general conv");
    int rc = require_sm100();
    if (rc) return rc;
    const int T = a->KH * a->KW;
    const int cpad = round_up(a->Ci, BLOCK_K);  // K block per tap in the prepared weight
    ConvParams base;
    memset(&base, 0, sizeof(base));
    base.B = a->B;
    base.Hm = a->Ho;
    base.Wm = a->Wo;
    base.cb0 = base.cb_total = cpad / BLOCK_K;
    base.noise_epoch = noise_epoch_ptr();
    base.passes = a->passes;
    base.n_total = a->n_total;
    base.w_inv_scale = a->w_inv_scale;
    base.a_inv_scale = a->a_inv_scale;
    base.b_rows = round_up(a->n_total < BLOCK_N ? a->n_total : BLOCK_N, 16);
    base.idesc = (1u << 4) | ((uint32_t)(base.b_rows >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
    base.m2 = a->n_total <= 128 ? 1 : 0;   // narrow layers (32 ... 128 output channels): 16 x 16-pixel tiles
    base.n_tiles = (a->n_total + BLOCK_N - 1) / BLOCK_N;
    base.bias = epi->bias;
    base.out = epi->out;
    base.lrelu = epi->lrelu;
    base.act_mask = (const __half*)epi->act_mask;
    base.amax_out = epi->amax_out;
    if (base.amax_out) DSEE_CUDA(cudaMemsetAsync(base.amax_out, 0, sizeof(float), (cudaStream_t)stream));
    const int a_step = a->transposed ? 1 : a->stride;
    for (int pl = 0; pl < 2; ++pl) {
        const void* ab = pl ? a->a_lo : a->a_hi;
        if (ab) {
            uint64_t dims[4] = {(uint64_t)a->Ci, (uint64_t)a->Wi, (uint64_t)a->Hi, (uint64_t)a->B};
            uint64_t strides[3] = {(uint64_t)a->Ci * 2, (uint64_t)a->Wi * a->Ci * 2,
                                   (uint64_t)a->Hi * a->Wi * a->Ci * 2};
            uint32_t box[4] = {BLOCK_K, (uint32_t)(TILE_W * a_step), (uint32_t)((TILE_H << base.m2) * a_step), 1};
            uint32_t es[4] = {1, (uint32_t)a_step, (uint32_t)a_step, 1};
            rc = encode_tmap_16b(&base.tmA[pl], ab, 4, dims, strides, box, false, es);
            if (rc) return rc;
        } else {
            base.tmA[pl] = base.tmA[0];
        }
        base.tmA[2 + pl] = base.tmA[pl];
        const void* wb = pl ? a->w_lo : a->w_hi;
        if (wb) {
            const uint64_t Ktot = (uint64_t)T * cpad;
            uint64_t dims[2] = {Ktot, (uint64_t)a->n_total};
            uint64_t strides[1] = {Ktot * 2};
            uint32_t box[2] = {BLOCK_K, (uint32_t)base.b_rows};
            rc = encode_tmap_16b(&base.tmB[pl], wb, 2, dims, strides, box, false);
            if (rc) return rc;
        } else {
            base.tmB[pl] = base.tmB[0];
        }
    }
    const int classes = a->transposed ? a->stride : 1;
    for (int py = 0; py < classes; ++py)
        for (int px = 0; px < classes; ++px) {
            ConvParams p = base;
            int nt = 0;
            for (int ky = 0; ky < a->KH; ++ky)
                for (int kx = 0; kx < a->KW; ++kx) {
                    int dy, dx;
                    if (!a->transposed) {
                        dy = ky - a->pad;
                        dx = kx - a->pad;
                    } else {
                        // dX[y] = sum_ky dY[(y + pad - ky) / stride] * w[ky] over exact divisions
                        const int ty = py + a->pad - ky, tx = px + a->pad - kx;
                        if (ty % a->stride != 0 || tx % a->stride != 0) continue;
                        dy = ty / a->stride;
                        dx = tx / a->stride;
                    }
                    p.tap_dy[nt] = (int8_t)dy;
                    p.tap_dx[nt] = (int8_t)dx;
                    p.tap_k[nt] = (ky * a->KW + kx) * cpad;
                    ++nt;
                }
            p.ntaps = nt;
            p.a_step = a_step;
            if (a->transposed) {
                p.H = (a->Ho - py + a->stride - 1) / a->stride;
                p.W = (a->Wo - px + a->stride - 1) / a->stride;
                p.o_step = a->stride;
                p.o_offy = py;
                p.o_offx = px;
            } else {
                p.H = a->Ho;
                p.W = a->Wo;
                p.o_step = 1;
            }
            if (p.H <= 0 || p.W <= 0) continue;
            DSEE_CHECK_ARG(nt > 0, "a parity class of the transposed conv has no taps (kernel < stride)");
            p.tiles_w = (p.W + TILE_W - 1) / TILE_W;
            p.tiles_h = (p.H + (TILE_H << p.m2) - 1) / (TILE_H << p.m2);
            p.num_tiles = p.B * p.tiles_h * p.tiles_w * p.n_tiles;
            rc = launch<EPI_CONV>(p, (cudaStream_t)stream);
            if (rc) return rc;
        }
    return 0;
}

extern "C" int dsee_subpixel_dgrad(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale,
                                   const void* w_hi, const void* w_lo, const float* w_inv_scale, int B, int H,
                                   int W, int n_total, int C, int passes, float* out, float* amax_out,
                                   void* stream) {
    DSEE_CHECK_ARG(dy_hi && w_hi && w_inv_scale && out, "NULL pointer");
    DSEE_CHECK_ARG(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "bad geometry (H, W even)");
    DSEE_CHECK_ARG(n_total > 0 && n_total % BLOCK_K == 0 && C > 0 && C % 32 == 0, "bad channel counts");
    DSEE_CHECK_ARG(passes == 1 || (passes == 3 && dy_lo && w_lo), "passes must be 1, or 3 with lo planes");
    int rc = require_sm100();
    if (rc) return rc;
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.B = B;
    p.H = p.Hm = H / 2;   // output = the half-resolution gradient
    p.W = p.Wm = W / 2;
    p.o_step = 1;
    p.a_step = 2;         // A = dY planes at full resolution, element stride 2
    p.ntaps = 16;
    for (int k = 0; k < 4; ++k)          // parity class (py, px)
        for (int t = 0; t < 4; ++t) {    // its 2x2 tap (ty, tx) at half-resolution offset r = t + p - 1
            const int py = k >> 1, px = k & 1;
            const int ry = (t >> 1) + py - 1, rx = (t & 1) + px - 1;
            // dA[Y] += dY[2 (Y - ry) + py] * wc: box origin = 2 * Y0 + (py - 2 ry)
            p.tap_dy[k * 4 + t] = (int8_t)(py - 2 * ry);
            p.tap_dx[k * 4 + t] = (int8_t)(px - 2 * rx);
            p.tap_k[k * 4 + t] = (k * 4 + t) * n_total;
        }
    p.cb0 = p.cb_total = n_total / BLOCK_K;
    p.passes = passes;
    p.n_total = C;
    p.w_inv_scale = w_inv_scale;
    p.a_inv_scale = dy_inv_scale;
    p.noise_epoch = noise_epoch_ptr();
    p.b_rows = round_up(C < BLOCK_N ? C : BLOCK_N, 16);
    p.idesc = (1u << 4) | ((uint32_t)(p.b_rows >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
    p.n_tiles = (C + BLOCK_N - 1) / BLOCK_N;
    p.m2 = C <= 128 ? 1 : 0;
    p.tiles_w = (p.W + TILE_W - 1) / TILE_W;
    p.tiles_h = (p.H + (TILE_H << p.m2) - 1) / (TILE_H << p.m2);
    p.num_tiles = p.B * p.tiles_h * p.tiles_w * p.n_tiles;
    p.out = out;
    p.amax_out = amax_out;
    if (amax_out) DSEE_CUDA(cudaMemsetAsync(amax_out, 0, sizeof(float), (cudaStream_t)stream));
    for (int pl = 0; pl < 2; ++pl) {
        const void* ab = pl ? dy_lo : dy_hi;
        if (ab) {
            uint64_t dims[4] = {(uint64_t)n_total, (uint64_t)W, (uint64_t)H, (uint64_t)B};
            uint64_t strides[3] = {(uint64_t)n_total * 2, (uint64_t)W * n_total * 2, (uint64_t)H * W * n_total * 2};
            uint32_t box[4] = {BLOCK_K, (uint32_t)(TILE_W * 2), (uint32_t)((TILE_H << p.m2) * 2), 1};
            uint32_t es[4] = {1, 2, 2, 1};
            rc = encode_tmap_16b(&p.tmA[pl], ab, 4, dims, strides, box, false, es);
            if (rc) return rc;
        } else {
            p.tmA[pl] = p.tmA[0];
        }
        p.tmA[2 + pl] = p.tmA[pl];
        const void* wb = pl ? w_lo : w_hi;
        if (wb) {
            const uint64_t Ktot = (uint64_t)16 * n_total;
            uint64_t dims[2] = {Ktot, (uint64_t)C};
            uint64_t strides[1] = {Ktot * 2};
            uint32_t box[2] = {BLOCK_K, (uint32_t)p.b_rows};
            rc = encode_tmap_16b(&p.tmB[pl], wb, 2, dims, strides, box, false);
            if (rc) return rc;
        } else {
            p.tmB[pl] = p.tmB[0];
        }
    }
    return launch<EPI_CONV>(p, (cudaStream_t)stream);
}

extern "C" int dsee_spade_modulate_fwd(const dsee_conv_operands* ops, const dsee_modulate_args* mod,
                                       void* stream) {
    ConvParams p;
    memset(&p, 0, sizeof(p));
    int rc = fill_common(p, ops, false, true);
    if (rc) return rc;
    DSEE_CHECK_ARG(mod != nullptr, "modulate args are NULL");
    DSEE_CHECK_ARG(mod->C > 0 && mod->C % 128 == 0, "C must be a multiple of 128 (got %d)", mod->C);
    DSEE_CHECK_ARG(ops->n_total == 2 * mod->C, "n_total (%d) must be 2*C (%d)", ops->n_total,
                   2 * mod->C);
    DSEE_CHECK_ARG(mod->x && mod->bn_scale && mod->bn_shift && mod->gamma_bias && mod->beta_bias &&
                       mod->out_hi,
                   "NULL modulate pointer");
    DSEE_CHECK_ARG(mod->x_ups == 0 || mod->x_ups == 1, "x_ups must be 0 or 1");
    DSEE_CHECK_ARG(mod->x_ups == 0 || (ops->H % 2 == 0 && ops->W % 2 == 0),
                   "folded upsample needs even H, W");
    DSEE_CHECK_ARG((mod->noise != nullptr || mod->noise_seed != 0) == (mod->noise_w != nullptr),
                   "noise (tensor or seed) and noise_w must be given together");
    p.x = mod->x;
    p.x_ups = mod->x_ups;
    p.noise = mod->noise;
    p.noise_w = mod->noise_w;
    p.noise_seed = mod->noise_seed;
    p.bn_scale = mod->bn_scale;
    p.bn_shift = mod->bn_shift;
    p.gamma_bias = mod->gamma_bias;
    p.beta_bias = mod->beta_bias;
    p.out_hi = (__half*)mod->out_hi;
    p.out_lo = (__half*)mod->out_lo;
    p.g_hi = (__half*)mod->g_hi;
    p.g_lo = (__half*)mod->g_lo;
    DSEE_CHECK_ARG((mod->out8_lo == nullptr) == (mod->out8_hi == nullptr), "out8_lo / out8_hi go together");
    p.out8_lo = (uint8_t*)mod->out8_lo;
    p.out8_hi = (uint8_t*)mod->out8_hi;
    p.C = mod->C;
    return launch<EPI_MODULATE>(p, (cudaStream_t)stream);
}

extern "C" int dsee_spade_modulate_bwd(const dsee_conv_operands* ops, const dsee_modulate_bwd_args* a,
                                       void* stream) {
    ConvParams p;
    memset(&p, 0, sizeof(p));
    int rc = fill_common(p, ops);
    if (rc) return rc;
    DSEE_CHECK_ARG(a != nullptr, "modulate bwd args are NULL");
    DSEE_CHECK_ARG(a->C > 0 && a->C % 128 == 0 && ops->n_total == a->C,
                   "gamma-only weights expected: n_total (%d) must equal C (%d)", ops->n_total, a->C);
    DSEE_CHECK_ARG(a->x && a->dt && a->bn_scale && a->bn_shift && a->gamma_bias && a->dxhat &&
                       a->dgb_hi && a->partial && a->dt_amax && a->dgb_inv_scale,
                   "NULL modulate bwd pointer");
    DSEE_CHECK_ARG(a->x_ups == 0 || a->x_ups == 1, "x_ups must be 0 or 1");
    DSEE_CHECK_ARG((a->noise == nullptr) == (a->noise_w == nullptr), "noise/noise_w mismatch");
    p.x = a->x;
    p.x_ups = a->x_ups;
    p.noise = a->noise;
    p.noise_w = a->noise_w;
    p.bn_scale = a->bn_scale;
    p.bn_shift = a->bn_shift;
    p.gamma_bias = a->gamma_bias;
    p.C = a->C;
    p.dt = a->dt;
    p.dxhat = a->dxhat;
    p.dgb_hi = (__half*)a->dgb_hi;
    p.dgb_lo = (__half*)a->dgb_lo;
    p.dt_amax = a->dt_amax;
    p.dgb_inv_scale = a->dgb_inv_scale;
    p.bwd_partial = a->partial;
    return launch<EPI_MODULATE_BWD>(p, (cudaStream_t)stream);
}

extern "C" int dsee_dgrad_modulate_bwd(const dsee_conv_operands* ops, const dsee_dgrad_modbwd_args* a,
                                       void* stream) {
    ConvParams p;
    memset(&p, 0, sizeof(p));
    int rc = fill_common(p, ops);
    if (rc) return rc;
    DSEE_CHECK_ARG(a != nullptr, "dgrad+modulate bwd args are NULL");
    DSEE_CHECK_ARG(a->C > 0 && a->C % 128 == 0 && ops->n_total == a->C,
                   "transposed weights expected: n_total (%d) must equal C (%d)", ops->n_total, a->C);
    DSEE_CHECK_ARG(a->act_mask && a->g_hi && a->x && a->bn_scale && a->bn_shift && a->dy_amax && a->w_l1 &&
                       a->dxhat && a->dgb_hi && a->dgb_inv_scale && a->partial,
                   "NULL dgrad+modulate bwd pointer");
    DSEE_CHECK_ARG(a->x_ups == 0 || a->x_ups == 1, "x_ups must be 0 or 1");
    DSEE_CHECK_ARG(a->x_ups == 0 || (ops->H % 2 == 0 && ops->W % 2 == 0), "folded upsample needs even H, W");
    DSEE_CHECK_ARG((a->noise != nullptr || a->noise_seed != 0) == (a->noise_w != nullptr),
                   "noise (tensor or seed) and noise_w must be given together");
    p.act_mask = (const __half*)a->act_mask;
    p.gs_hi = (const __half*)a->g_hi;
    p.gs_lo = (const __half*)a->g_lo;
    p.x = a->x;
    p.x_ups = a->x_ups;
    p.noise = a->noise;
    p.noise_seed = a->noise_seed;
    p.noise_w = a->noise_w;
    p.bn_scale = a->bn_scale;
    p.bn_shift = a->bn_shift;
    p.dy_amax = a->dy_amax;
    p.w_l1 = a->w_l1;
    p.C = a->C;
    p.dxhat = a->dxhat;
    p.dgb_hi = (__half*)a->dgb_hi;
    p.dgb_lo = (__half*)a->dgb_lo;
    p.dgb_inv_scale = a->dgb_inv_scale;
    p.bwd_partial = a->partial;
    return launch<EPI_DGRAD_MODBWD>(p, (cudaStream_t)stream);
}
