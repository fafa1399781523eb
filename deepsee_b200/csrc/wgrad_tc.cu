// Weight gradient of the 3x3 convolutions on tcgen05 tensor cores.
//
//   dW[n][tap][c] = sum_{b,y,x} dY[b,y,x,n] * A[b,y+dy,x+dx,c]
//
// is, for every tap, a GEMM whose reduction dimension is the pixel index.  Both operands are
// stored channel-contiguous (NHWC), i.e. "MN-major" from the GEMM's point of view, which
// tcgen05.mma consumes directly (instruction-descriptor a_major = b_major = 1): a 4-D TMA box
// {64 ch, 16 w, 8 h, 1 b} lands in shared memory as 128 pixel rows of 128 B, which is exactly the
// canonical 128B-swizzled MN-major layout (8-row swizzle atoms 1024 B apart along K, 64-channel
// column blocks one box = 16 KB apart along M/N).  The shifted box of the activation supplies the
// tap offset and, through TMA out-of-bounds zero fill, the convolution's zero padding.
//
// Work unit = (128 or 2 x 128 dY-channels) x (tap) x (<=256 A-channels) x (pixel split); the K loop walks
// the split's pixel tiles.  Units are spread over persistent CTAs; each writes its fp32 accumulator
// tile(s) to a per-split partial buffer that a small kernel reduces in a fixed order (deterministic,
// no atomics).  Same warp roles / mbarrier pipeline as conv_tc.cu, with three TMA-issuing warps.
// The two-tile form (WgCfg<4, 3, 2>, used whenever n_total is a multiple of 256) exists because the
// one-tile form is bound by the L2 -> SM feed: see WgCfg.
//
// Reference semantics: autograd of nn.Conv2d at architecture.py:98,122 and
// normalization.py:116-117,198-201,283-284.
#include "common.cuh"
#include "../../include/deepsee_b200.h"
#include "launch_count.h"
#include <stdlib.h>
#include <string.h>

namespace dsee {

constexpr int WG_M = 128;        // dY channels per unit
constexpr int WG_NMAX = 256;     // activation channels per unit
constexpr int WG_TW = 16;
constexpr int WG_PRODUCERS = 3;   // warp 0 and warps 6, 7: one elected thread each issues a third of a stage's boxes
constexpr int WG_THREADS = 192 + 32 * (WG_PRODUCERS - 1);
// Two forms of the kernel:
//   <TH 8, 2 stages, MT 1>  K step = 8 x 16 pixels, one 128-channel dY tile per unit: 6 boxes of 16 KB per
//                           stage (96 KB), accumulators double-buffered in TMEM;
//   <TH 4, 3 stages, MT 2>  K step = 4 x 16 pixels, TWO dY tiles per unit sharing the activation boxes
//                           (two MMA groups per K step into TMEM columns [0,256) / [256,512)): 8 boxes of
//                           8 KB per stage (64 KB) for the same MMA time - 2/3 of the L2 -> SM bytes per
//                           FLOP, one more stage in flight.  The unit's accumulators fill TMEM, so its
//                           epilogue is not overlapped (a unit's main loop is hundreds of K steps).
template <int TH, int STAGES, int MT>
struct WgCfg {
    static constexpr int KPIX = WG_TW * TH;                      // pixels per K step
    static constexpr int BOX_BYTES = KPIX * 64 * 2;              // pixel rows x 64 ch
    static constexpr int NBOX_D = MT * WG_M / 64;                // dY boxes per stage
    static constexpr int STAGE_BYTES = (NBOX_D + WG_NMAX / 64) * BOX_BYTES;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
};

struct alignas(64) WgradParams {
    CUtensorMap tmD[2];  // dY planes (hi, lo)
    CUtensorMap tmA[2];  // activation planes (hi, lo)
    CUtensorMap tmA2[2]; // optional second activation source (channels c_split .. c_total)
    int c_split;         // channels taken from the first source (== c_total when there is one)
    int B, H, W;
    int tiles_w, tiles_h, ptiles;  // pixel tiles
    int n_total, c_total;
    int ntaps, a_step;           // filter taps; A box origin = dY tile origin * a_step + tap offset
    int d_step, d_offy, d_offx;  // dY box origin = tile origin * d_step + offset (sub-pixel classes: 2)
    int8_t tap_dy[16], tap_dx[16];
    int n_tiles, c_tiles, splits, num_units;
    int n_cols;  // activation channels per unit (<= 256)
    int passes;
    uint32_t idesc;
    float* partial;  // [splits][n_total][9][c_total]
    const float* inv_scale[2];  // device scalars undoing the operand pre-scaling (NULL = 1)
};

// MN-major, 128B-swizzled operand descriptor (see header comment).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t box_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((box_bytes >> 4) & 0x3FFF) << 16;  // LBO: next 64-channel block
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;          // SBO: next 8 pixel rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void wg_decode(const WgradParams& p, int unit, int& nt, int& tap, int& ct,
                                          int& split) {
    ct = unit % p.c_tiles;
    unit /= p.c_tiles;
    tap = unit % p.ntaps;
    unit /= p.ntaps;
    nt = unit % p.n_tiles;
    split = unit / p.n_tiles;
}

template <int TH, int STAGES, int MT>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
    using Cfg = WgCfg<TH, STAGES, MT>;
    constexpr int WG_STAGES = STAGES, WG_STAGE_BYTES = Cfg::STAGE_BYTES, WG_BOX_BYTES = Cfg::BOX_BYTES;
    constexpr int WG_KPIX = Cfg::KPIX, WG_TH = TH, NBOX_D = Cfg::NBOX_D;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + WG_STAGES;
    uint64_t* tfull_bar = bars + 2 * WG_STAGES;
    uint64_t* tempty_bar = bars + 2 * WG_STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmD[0]);
        tma_prefetch_desc(&p.tmA[0]);
        for (int s = 0; s < WG_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nboxA = p.n_cols / 64;
    const uint32_t stage_tx = (uint32_t)(NBOX_D + nboxA) * WG_BOX_BYTES;

    if (warp == 0 || warp >= 6) {
        // TMA producers.  A stage is 2 + nboxA boxes of 64 channels; one thread issuing all of them needs
        // ~1100 cycles of address arithmetic and uniform-register moves per stage - as long as the MMAs
        // of the stage (ncu source view: the MMA warp waited for data a third of the time).  Three
        // elected threads in three warps issue every third box each; producer 0 posts the byte count.
        const int prod = warp == 0 ? 0 : warp - 5;
        if (lane == 0) {
            uint32_t it = 0;
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
                int nt, tap, ct, split;
                wg_decode(p, unit, nt, tap, ct, split);
                const int dy = p.tap_dy[tap], dx = p.tap_dx[tap];
                const int pt0 = (int)((int64_t)p.ptiles * split / p.splits);
                const int pt1 = (int)((int64_t)p.ptiles * (split + 1) / p.splits);
                int tw = pt0 % p.tiles_w, th = (pt0 / p.tiles_w) % p.tiles_h, b = pt0 / (p.tiles_w * p.tiles_h);
                for (int pt = pt0; pt < pt1; ++pt) {
                    const int h0 = th * WG_TH, w0 = tw * WG_TW;
                    for (int pass = 0; pass < p.passes; ++pass, ++it) {
                        const int pd = (pass == 1) ? 1 : 0;  // dY plane
                        const int pa = (pass == 2) ? 1 : 0;  // activation plane
                        const int s = it % WG_STAGES;
                        const uint32_t ph = (it / WG_STAGES) & 1;
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        if (prod == 0) mbar_expect_tx(&full_bar[s], stage_tx);
                        uint8_t* sd = smem + s * WG_STAGE_BYTES;
                        uint8_t* sa = sd + NBOX_D * WG_BOX_BYTES;
#pragma unroll
                        for (int i = 0; i < NBOX_D; ++i) {
                            if (i % WG_PRODUCERS != prod) continue;
                            tma_load_4d(&p.tmD[pd], &full_bar[s], sd + i * WG_BOX_BYTES,
                                        nt * (WG_M * MT) + i * 64, w0 * p.d_step + p.d_offx,
                                        h0 * p.d_step + p.d_offy, b);
                        }
#pragma unroll
                        for (int i = 0; i < WG_NMAX / 64; ++i) {
                            if (i >= nboxA || (i + NBOX_D) % WG_PRODUCERS != prod) continue;
                            const int c = ct * WG_NMAX + i * 64;
                            const bool second = c >= p.c_split;
                            tma_load_4d(second ? &p.tmA2[pa] : &p.tmA[pa], &full_bar[s],
                                        sa + i * WG_BOX_BYTES, second ? c - p.c_split : c,
                                        w0 * p.a_step + dx, h0 * p.a_step + dy, b);
                        }
                    }
                    if (++tw == p.tiles_w) {
                        tw = 0;
                        if (++th == p.tiles_h) {
                            th = 0;
                            ++b;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0;
            int lu = 0;
            for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++lu) {
                int nt, tap, ct, split;
                wg_decode(p, unit, nt, tap, ct, split);
                const int pt0 = (int)((int64_t)p.ptiles * split / p.splits);
                const int pt1 = (int)((int64_t)p.ptiles * (split + 1) / p.splits);
                const int kiters = (pt1 - pt0) * p.passes;
                // MT == 1: two accumulator stages of 256 columns; MT == 2: the unit's two tiles fill TMEM
                const int as = MT == 2 ? 0 : (lu & 1);
                const uint32_t aph = MT == 2 ? (lu & 1) : ((lu >> 1) & 1);
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * WG_NMAX;
                for (int kit = 0; kit < kiters; ++kit, ++it) {
                    const int s = it % WG_STAGES;
                    const uint32_t ph = (it / WG_STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sd = smem_u32(smem + s * WG_STAGE_BYTES);
                    const uint32_t sa = sd + NBOX_D * WG_BOX_BYTES;
                    const uint64_t db = umma_desc_mn_sw128(sa, WG_BOX_BYTES);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint64_t da = umma_desc_mn_sw128(sd + mt * (WG_M / 64) * WG_BOX_BYTES, WG_BOX_BYTES);
#pragma unroll
                        for (int k = 0; k < WG_KPIX / 16; ++k) {
                            // advance 16 pixel rows = 2048 B = 128 (16 B units) along K
                            umma_f16(tmem_d + mt * WG_NMAX, da + 128 * k, db + 128 * k, p.idesc, (kit | k) != 0);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                if (kiters == 0) {
                    // empty split (fewer pixel tiles than splits): nothing accumulated; the
                    // epilogue writes zeros (it checks the same condition)
                }
                umma_commit(&tfull_bar[as]);
            }
        }
    } else if (warp < 6) {
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const float scale = (p.inv_scale[0] ? __ldg(p.inv_scale[0]) : 1.f) *
                            (p.inv_scale[1] ? __ldg(p.inv_scale[1]) : 1.f);
        int lu = 0;
        for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x, ++lu) {
            int nt, tap, ct, split;
            wg_decode(p, unit, nt, tap, ct, split);
            const int pt0 = (int)((int64_t)p.ptiles * split / p.splits);
            const int pt1 = (int)((int64_t)p.ptiles * (split + 1) / p.splits);
            const bool empty = pt1 <= pt0;
            const int as = MT == 2 ? 0 : (lu & 1);
            const uint32_t aph = MT == 2 ? (lu & 1) : ((lu >> 1) & 1);
            mbar_wait(&tfull_bar[as], aph);
            tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (as + mt) * WG_NMAX;
                const int n = (nt * MT + mt) * WG_M + m;
                float* orow = p.partial + (((size_t)split * p.n_total + n) * p.ntaps + tap) * p.c_total +
                              (size_t)ct * WG_NMAX;
#pragma unroll 1
                for (int ch = 0; ch < p.n_cols / 32; ++ch) {
                    uint32_t v[32];
                    tmem_ld32(taddr + ch * 32, v);
                    tmem_ld_wait();
                    if (n < p.n_total) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 o = make_float4(__uint_as_float(v[4 * j]) * scale,
                                                   __uint_as_float(v[4 * j + 1]) * scale,
                                                   __uint_as_float(v[4 * j + 2]) * scale,
                                                   __uint_as_float(v[4 * j + 3]) * scale);
                            if (empty) o = make_float4(0.f, 0.f, 0.f, 0.f);
                            reinterpret_cast<float4*>(orow + ch * 32)[j] = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// out = sum_s partial[s] (fixed order), optionally transposing [n][T][c] -> [n][c][T].
// block = (n, 32-channel slab): coalesced reads of T x 32 partial rows, smem transpose, coalesced
// write of the 32*T contiguous outputs.
// blockIdx.y = output group (per-image results: group g sums splits [g*splits, (g+1)*splits)).
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                    int splits, int n_total, int c_total, int to_nc9, int T) {
    __shared__ float tile[16][33];
    const int cblocks = c_total / 32;
    const int n = blockIdx.x / cblocks, c0 = (blockIdx.x % cblocks) * 32;
    const int64_t total = (int64_t)n_total * T * c_total;
    partial += (size_t)blockIdx.y * splits * total;
    out += (size_t)blockIdx.y * total;
    const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;  // 32 x 16 threads
    float a = 0.f;
    if (row < T) {
        const size_t i = ((size_t)n * T + row) * c_total + c0 + lane;
        for (int s = 0; s < splits; ++s) a += __ldg(partial + (size_t)s * total + i);
        if (!to_nc9) out[i] = a;
        tile[row][lane] = a;
    }
    if (!to_nc9) return;
    __syncthreads();
    // out[(n*c_total + c0 + cl) * T + tap], contiguous over (cl, tap)
    for (int e = threadIdx.x; e < 32 * T; e += blockDim.x) {
        const int cl = e / T, tap = e % T;
        out[((size_t)n * c_total + c0) * T + e] = tile[tap][cl];
    }
}

}  // namespace dsee

using namespace dsee;

// dY tiles per unit: two whenever the dY channels come in pairs of 128-channel tiles (see WgCfg)
static int g_wgrad_pairs = -1;
static int wgrad_mt(int n_total) {
    if (g_wgrad_pairs < 0) {
        const char* e = getenv("DSEE_WGRAD_PAIRS");
        g_wgrad_pairs = (e && e[0] == '0') ? 0 : 1;
    }
    return (g_wgrad_pairs && n_total % (2 * WG_M) == 0) ? 2 : 1;
}
static int wgrad_th(int mt) { return mt == 2 ? 4 : 8; }

static int wgrad_plan(int B, int H, int W, int n_total, int c_total, int* splits_out, int T = 9,
                      bool per_image = false) {
    const int mt = wgrad_mt(n_total), th_ = wgrad_th(mt);
    const int ptiles = B * ((H + th_ - 1) / th_) * ((W + WG_TW - 1) / WG_TW);
    const int n_tiles = (n_total + WG_M * mt - 1) / (WG_M * mt);
    const int c_tiles = (c_total + WG_NMAX - 1) / WG_NMAX;
    const int base = n_tiles * T * c_tiles;
    // Pixel splits: units = base * splits run as ceil(units / 148) waves of ceil(ptiles / splits)
    // K steps each.  Pick the split count with the shortest makespan (a wave that is only partly
    // full costs as much as a full one: base = 72 with 5 splits is 3 waves of 820 steps, with 2 or
    // 4 splits it is 2048 steps in total); ties go to fewer splits (less partial-buffer traffic).
    // Each unit also pays a fixed cost (accumulator drain + pipeline fill) of about 8 K steps.
    const int sms = 148;
    int splits = 1;
    long best = -1;
    // per_image: split boundaries must coincide with image boundaries (ptiles = B * tiles per image,
    // image-major), so the split count is a multiple of B
    // (few units per split - a narrow 1x1 layer like the image head, base = 2 - may split further, up
    // to one unit per SM)
    const int wide = sms / base > 16 ? sms / base : 16;
    const int step = per_image ? B : 1, smax = per_image ? (B > 16 ? B : 16 / B * B) : wide;
    for (int s = step; s <= smax && s <= ptiles; s += step) {
        const long waves = ((long)base * s + sms - 1) / sms;
        const long steps = (ptiles + s - 1) / s;
        const long cost = waves * (steps + 8);
        if (best < 0 || cost < best) {
            best = cost;
            splits = s;
        }
    }
    *splits_out = splits;
    return ptiles;
}

extern "C" int64_t dsee_conv3x3_wgrad_workspace_floats(int B, int H, int W, int n_total, int c_total) {
    int splits;
    wgrad_plan(B, H, W, n_total, c_total, &splits);
    return (int64_t)splits * n_total * 9 * c_total;
}

static int wgrad_impl(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale, const void* a_hi,
                      const void* a_lo, const float* a_inv_scale, int dtype, int B, int H, int W,
                      int Hi, int Wi, int n_total, int a_channels, int c_total, int KH, int KW,
                      int stride, int pad, int passes, float* workspace, float* dw, int layout_nc9,
                      void* stream, const void* a2_hi = nullptr, const void* a2_lo = nullptr,
                      int a2_channels = 0, bool per_image = false, int sub_py = -1, int sub_px = -1) {
    // sub_py / sub_px >= 0: one parity class of the sub-pixel form - H, W are the HALF-resolution
    // tile space, the dY planes are [B, 2H, 2W, n_total] read with element stride 2 from (sub_py,
    // sub_px), the 2x2 taps sit at rows ty + sub_py - 1 of the half-resolution activation
    // H, W: dY (= forward output) size; Hi, Wi: activation (= forward input) size;
    // a_channels: channels stored in the activation planes, c_total: dW columns (a multiple of 64)
    const int T = KH * KW;
    int rc = require_sm100();
    if (rc) return rc;
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.B = B;
    p.H = H;
    p.W = W;
    p.tiles_w = (W + WG_TW - 1) / WG_TW;
    const int mt = wgrad_mt(n_total), th_ = wgrad_th(mt);
    p.tiles_h = (H + th_ - 1) / th_;
    p.n_total = n_total;
    p.c_total = c_total;
    p.c_split = a2_channels > 0 ? a_channels : c_total;
    p.ntaps = T;
    p.a_step = stride;
    const bool sub = sub_py >= 0;
    p.d_step = sub ? 2 : 1;
    p.d_offy = sub ? sub_py : 0;
    p.d_offx = sub ? sub_px : 0;
    for (int t = 0; t < T; ++t) {
        p.tap_dy[t] = (int8_t)(t / KW - (sub ? 1 - sub_py : pad));
        p.tap_dx[t] = (int8_t)(t % KW - (sub ? 1 - sub_px : pad));
    }
    p.n_tiles = (n_total + WG_M * mt - 1) / (WG_M * mt);
    p.c_tiles = (c_total + WG_NMAX - 1) / WG_NMAX;
    p.n_cols = c_total < WG_NMAX ? c_total : WG_NMAX;
    p.ptiles = wgrad_plan(B, H, W, n_total, c_total, &p.splits, T, per_image);
    p.num_units = p.n_tiles * T * p.c_tiles * p.splits;
    p.passes = passes;
    p.partial = workspace;
    p.inv_scale[0] = dy_inv_scale;
    p.inv_scale[1] = a_inv_scale;
    // kind::f16, fp32 accumulate, both operands MN-major, M = 128, N = n_cols
    p.idesc = (1u << 4) | ((uint32_t)dtype << 7) | ((uint32_t)dtype << 10) | (1u << 15) | (1u << 16) |
              ((uint32_t)(p.n_cols >> 3) << 17) | ((uint32_t)(WG_M >> 4) << 24);
    uint32_t boxd[4] = {64, (uint32_t)(WG_TW * p.d_step), (uint32_t)(th_ * p.d_step), 1};
    uint32_t esd[4] = {1, (uint32_t)p.d_step, (uint32_t)p.d_step, 1};
    uint32_t boxa[4] = {64, (uint32_t)(WG_TW * stride), (uint32_t)(th_ * stride), 1};
    uint32_t esa[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    for (int pl = 0; pl < 2; ++pl) {
        const void* d = pl ? dy_lo : dy_hi;
        const void* a = pl ? a_lo : a_hi;
        if (d) {
            const uint64_t Wd = (uint64_t)W * p.d_step, Hd = (uint64_t)H * p.d_step;
            uint64_t dims[4] = {(uint64_t)n_total, Wd, Hd, (uint64_t)B};
            uint64_t st[3] = {(uint64_t)n_total * 2, Wd * n_total * 2, Hd * Wd * n_total * 2};
            rc = encode_tmap_16b(&p.tmD[pl], d, 4, dims, st, boxd, dtype == 1, esd);
            if (rc) return rc;
        } else {
            p.tmD[pl] = p.tmD[0];
        }
        if (a) {
            uint64_t dims[4] = {(uint64_t)a_channels, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)B};
            uint64_t st[3] = {(uint64_t)a_channels * 2, (uint64_t)Wi * a_channels * 2,
                              (uint64_t)Hi * Wi * a_channels * 2};
            rc = encode_tmap_16b(&p.tmA[pl], a, 4, dims, st, boxa, dtype == 1, esa);
            if (rc) return rc;
        } else {
            p.tmA[pl] = p.tmA[0];
        }
        const void* a2 = pl ? a2_lo : a2_hi;
        if (a2 && a2_channels > 0) {
            uint64_t dims[4] = {(uint64_t)a2_channels, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)B};
            uint64_t st[3] = {(uint64_t)a2_channels * 2, (uint64_t)Wi * a2_channels * 2,
                              (uint64_t)Hi * Wi * a2_channels * 2};
            rc = encode_tmap_16b(&p.tmA2[pl], a2, 4, dims, st, boxa, dtype == 1, esa);
            if (rc) return rc;
        } else {
            p.tmA2[pl] = p.tmA[pl];
        }
    }
    using Cfg1 = WgCfg<8, 2, 1>;
    using Cfg2 = WgCfg<4, 3, 2>;
    static bool configured[64] = {false};
    int dev = 0;
    DSEE_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        DSEE_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<8, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg1::SMEM));
        DSEE_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<4, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg2::SMEM));
        configured[dev] = true;
    }
    int sms = 0;
    DSEE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = p.num_units < sms ? p.num_units : sms;
    cudaStream_t st = (cudaStream_t)stream;
    if (mt == 2)
        wgrad_tc_kernel<4, 3, 2><<<grid, WG_THREADS, Cfg2::SMEM, st>>>(p);
    else
        wgrad_tc_kernel<8, 2, 1><<<grid, WG_THREADS, Cfg1::SMEM, st>>>(p);
    count_launch();
    DSEE_CUDA(cudaGetLastError());
    const int64_t total = (int64_t)n_total * T * c_total;
    (void)total;
    const int groups = per_image ? B : 1;
    wgrad_reduce_kernel<<<dim3(n_total * (c_total / 32), groups), 512, 0, st>>>(
        workspace, dw, p.splits / groups, n_total, c_total, layout_nc9, T);
    count_launch();
    DSEE_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int dsee_conv3x3_wgrad(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale,
                                  const void* a_hi, const void* a_lo, const float* a_inv_scale,
                                  int dtype, int B, int H, int W, int n_total, int c_total, int passes,
                                  float* workspace, float* dw, int layout_nc9, void* stream) {
    DSEE_CHECK_ARG(dy_hi && a_hi && workspace && dw, "NULL pointer");
    DSEE_CHECK_ARG(B > 0 && H > 0 && W > 0, "bad geometry");
    DSEE_CHECK_ARG(n_total % 128 == 0, "n_total must be a multiple of 128 (got %d)", n_total);
    DSEE_CHECK_ARG(c_total == 64 || c_total == 128 || c_total % 256 == 0,
                   "c_total must be 64, 128 or a multiple of 256 (got %d)", c_total);
    DSEE_CHECK_ARG(passes == 1 || (passes == 3 && dy_lo && a_lo), "passes must be 1, or 3 with lo planes");
    DSEE_CHECK_ARG(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
    return wgrad_impl(dy_hi, dy_lo, dy_inv_scale, a_hi, a_lo, a_inv_scale, dtype, B, H, W, H, W, n_total,
                      c_total, c_total, 3, 3, 1, 1, passes, workspace, dw, layout_nc9, stream);
}

extern "C" int dsee_conv3x3_wgrad2(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale,
                                   const void* const* a_hi, const void* const* a_lo,
                                   const int* a_channels, int dtype, int B, int H, int W, int n_total,
                                   int passes, float* workspace, float* dw, int layout_nc9, void* stream) {
    DSEE_CHECK_ARG(dy_hi && a_hi && a_channels && a_hi[0] && workspace && dw, "NULL pointer");
    DSEE_CHECK_ARG(B > 0 && H > 0 && W > 0, "bad geometry");
    DSEE_CHECK_ARG(n_total % 128 == 0, "n_total must be a multiple of 128 (got %d)", n_total);
    DSEE_CHECK_ARG(a_channels[0] > 0 && a_channels[0] % 64 == 0 && a_channels[1] >= 0 &&
                       a_channels[1] % 64 == 0 && (a_channels[1] == 0 || a_hi[1]),
                   "source channel counts must be multiples of 64");
    const int c_total = a_channels[0] + a_channels[1];
    DSEE_CHECK_ARG(c_total == 64 || c_total == 128 || c_total % 256 == 0,
                   "total channels must be 64, 128 or a multiple of 256 (got %d)", c_total);
    DSEE_CHECK_ARG(passes == 1 || (passes == 3 && dy_lo && a_lo && a_lo[0] && (a_channels[1] == 0 || a_lo[1])),
                   "passes must be 1, or 3 with lo planes");
    DSEE_CHECK_ARG(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
    return wgrad_impl(dy_hi, dy_lo, dy_inv_scale, a_hi[0], a_lo ? a_lo[0] : nullptr, nullptr, dtype, B, H,
                      W, H, W, n_total, a_channels[0], c_total, 3, 3, 1, 1, passes, workspace, dw,
                      layout_nc9, stream, a_hi[1], a_lo ? a_lo[1] : nullptr, a_channels[1]);
}

extern "C" int64_t dsee_conv3x3_wgrad_per_image_workspace_floats(int B, int H, int W, int n_total,
                                                                 int c_total) {
    int splits;
    wgrad_plan(B, H, W, n_total, c_total, &splits, 9, true);
    return (int64_t)splits * n_total * 9 * c_total;
}

extern "C" int dsee_conv3x3_wgrad2_per_image(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale,
                                             const void* const* a_hi, const void* const* a_lo,
                                             const int* a_channels, int dtype, int B, int H, int W,
                                             int n_total, int passes, float* workspace, float* dw,
                                             void* stream) {
    DSEE_CHECK_ARG(dy_hi && a_hi && a_channels && a_hi[0] && workspace && dw, "NULL pointer");
    DSEE_CHECK_ARG(B > 0 && H > 0 && W > 0, "bad geometry");
    DSEE_CHECK_ARG(n_total % 128 == 0, "n_total must be a multiple of 128 (got %d)", n_total);
    DSEE_CHECK_ARG(a_channels[0] > 0 && a_channels[0] % 64 == 0 && a_channels[1] >= 0 &&
                       a_channels[1] % 64 == 0 && (a_channels[1] == 0 || a_hi[1]),
                   "source channel counts must be multiples of 64");
    const int c_total = a_channels[0] + a_channels[1];
    DSEE_CHECK_ARG(c_total <= 256 || c_total % 256 == 0,
                   "total channels must be at most 256 or a multiple of 256 (got %d)", c_total);
    DSEE_CHECK_ARG(passes == 1 || (passes == 3 && dy_lo && a_lo && a_lo[0] && (a_channels[1] == 0 || a_lo[1])),
                   "passes must be 1, or 3 with lo planes");
    DSEE_CHECK_ARG(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
    return wgrad_impl(dy_hi, dy_lo, dy_inv_scale, a_hi[0], a_lo ? a_lo[0] : nullptr, nullptr, dtype, B, H,
                      W, H, W, n_total, a_channels[0], c_total, 3, 3, 1, 1, passes, workspace, dw, 1, stream,
                      a_hi[1], a_lo ? a_lo[1] : nullptr, a_channels[1], true);
}

extern "C" int64_t dsee_subpixel_wgrad_workspace_floats(int B, int H, int W, int n_total, int c_total) {
    int splits;
    wgrad_plan(B, H / 2, W / 2, n_total, c_total, &splits, 4);
    return (int64_t)splits * n_total * 4 * c_total;
}

extern "C" int dsee_subpixel_wgrad(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale,
                                   const void* const* a_hi, const void* const* a_lo, const int* a_channels,
                                   int B, int H, int W, int n_total, int sub_py, int sub_px, int passes,
                                   float* workspace, float* dwc, void* stream) {
    DSEE_CHECK_ARG(dy_hi && a_hi && a_channels && a_hi[0] && workspace && dwc, "NULL pointer");
    DSEE_CHECK_ARG(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "bad geometry (H, W even)");
    DSEE_CHECK_ARG((sub_py | 1) == 1 && (sub_px | 1) == 1, "parity class must be in {0,1}^2");
    DSEE_CHECK_ARG(n_total % 128 == 0, "n_total must be a multiple of 128 (got %d)", n_total);
    DSEE_CHECK_ARG(a_channels[0] > 0 && a_channels[0] % 64 == 0 && a_channels[1] >= 0 &&
                       a_channels[1] % 64 == 0 && (a_channels[1] == 0 || a_hi[1]),
                   "source channel counts must be multiples of 64");
    const int c_total = a_channels[0] + a_channels[1];
    DSEE_CHECK_ARG(c_total <= 256 || c_total % 256 == 0,
                   "total channels must be at most 256 or a multiple of 256 (got %d)", c_total);
    DSEE_CHECK_ARG(passes == 1 || (passes == 3 && dy_lo && a_lo && a_lo[0] && (a_channels[1] == 0 || a_lo[1])),
                   "passes must be 1, or 3 with lo planes");
    return wgrad_impl(dy_hi, dy_lo, dy_inv_scale, a_hi[0], a_lo ? a_lo[0] : nullptr, nullptr, 0, B, H / 2,
                      W / 2, H / 2, W / 2, n_total, a_channels[0], c_total, 2, 2, 1, 0, passes, workspace, dwc,
                      1, stream, a_hi[1], a_lo ? a_lo[1] : nullptr, a_channels[1], false, sub_py, sub_px);
}

extern "C" int64_t dsee_conv2d_tc_wgrad_workspace_floats(int B, int Ho, int Wo, int n_total, int Ci,
                                                         int KH, int KW) {
    int splits;
    const int cpad = (Ci + 63) / 64 * 64;
    wgrad_plan(B, Ho, Wo, n_total, cpad, &splits, KH * KW);
    return (int64_t)splits * n_total * KH * KW * cpad;
}

extern "C" int dsee_conv2d_tc_wgrad(const void* dy_hi, const void* dy_lo, const float* dy_inv_scale,
                                    const void* a_hi, const void* a_lo, const float* a_inv_scale, int B,
                                    int Ho, int Wo, int Hi, int Wi, int n_total, int Ci, int KH, int KW,
                                    int stride, int pad, int passes, float* workspace, float* dw,
                                    void* stream) {
    DSEE_CHECK_ARG(dy_hi && a_hi && workspace && dw, "NULL pointer");
    DSEE_CHECK_ARG(B > 0 && Ho > 0 && Wo > 0 && Hi > 0 && Wi > 0, "bad geometry");
    DSEE_CHECK_ARG(n_total % 8 == 0 && Ci % 8 == 0, "channel counts must be multiples of 8");
    DSEE_CHECK_ARG(KH * KW <= 16 && (stride == 1 || stride == 2), "at most 16 taps, stride 1 or 2");
    DSEE_CHECK_ARG(passes == 1 || (passes == 3 && dy_lo && a_lo), "passes must be 1, or 3 with lo planes");
    const int cpad = (Ci + 63) / 64 * 64;
    DSEE_CHECK_ARG(cpad == 64 || cpad == 128 || cpad % 256 == 0,
                   "input channels (padded to 64) must be 64, 128 or a multiple of 256");
    return wgrad_impl(dy_hi, dy_lo, dy_inv_scale, a_hi, a_lo, a_inv_scale, 0, B, Ho, Wo, Hi, Wi, n_total,
                      Ci, cpad, KH, KW, stride, pad, passes, workspace, dw, 1, stream);
}
