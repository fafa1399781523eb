// Shared device/host helpers for the deepsee_b200 CUDA library (sm_100a only).
//
// PTX wrappers for mbarrier / TMA (cp.async.bulk.tensor) / tcgen05 (UMMA + TMEM),
// the thread-local error string behind dsee_last_error(), and the host-side
// cuTensorMapEncodeTiled loader (resolved through cudaGetDriverEntryPoint so the
// library links against libcudart only and can be built on a box with no driver).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace dsee {

// ----------------------------------------------------------------------------------------------
// host: error convention: 0 = ok, <0 = invalid argument / unsupported device,
// >0 = CUDA error code. No exceptions, no exit().
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define DSEE_CHECK_ARG(cond, ...)                   \
    do {                                            \
        if (!(cond)) {                              \
            ::dsee::set_error(__VA_ARGS__);         \
            return -1;                              \
        }                                           \
    } while (0)

#define DSEE_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::dsee::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                              __FILE__, __LINE__);                                        \
            return (int)_e;                                                               \
        }                                                                                 \
    } while (0)

// Checks the current device is sm_100 (B200). Returns 0 or a negative error.
int require_sm100();

// Device address of the current device's noise epoch: a 64-bit counter every noise-regenerating
// kernel folds into its seed (eff_noise_seed below).  dsee_noise_epoch_advance bumps it with a
// one-thread kernel, so a CUDA graph that replays launches with baked-in seeds still draws fresh
// noise on every replay, consistently across the forward and backward kernels of one step.
const unsigned long long* noise_epoch_ptr();

// Encodes a tiled tensor map (rank <= 5) for a 16-bit element tensor with 128B swizzle.
// dims/strides innermost-first; strides[0] is implied (= 2 bytes).
// elem_strides (optional, per dimension): traversal stride; the box then lands
// ceil(box[i] / elem_strides[i]) elements of dimension i in shared memory.
int encode_tmap_16b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, bool bf16,
                    const uint32_t* elem_strides = nullptr);
// Same for an 8-bit element tensor (fp8 operands of tcgen05 kind::f8f6f4 are plain bytes to TMA).
int encode_tmap_8b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);

#ifdef __CUDACC__
// ----------------------------------------------------------------------------------------------
// device: PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Spin on the phase parity. A watchdog turns a dead pipeline into a trap (the launch then fails
// with an error the host reports) rather than a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    uint32_t spins = 0;
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > (1u << 26)) {
            printf("dsee: mbarrier watchdog block %d thread %d bar %u parity %u\n", blockIdx.x,
                   threadIdx.x, addr, parity);
            __trap();
        }
    }
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0,
                                            int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0,
                                            int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- tcgen05 / TMEM ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; kind::f16 covers fp16 and bf16 operands with fp32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with 8-bit float operands (e4m3 / e5m2 chosen per operand in idesc): K = 32 per instruction,
// i.e. the same 32 bytes of a 128B-swizzled K-major row as a kind::f16 step, at twice the FLOP rate.
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                        uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// ---- CTA pair (cta_group::2): two SMs of one TPC execute one M = 256 MMA --------------------------
// Each CTA of the 2-CTA cluster stages its own 128 A rows and its own half of the B rows; the leader
// (cluster rank 0) issues the MMA, which reads both CTAs' shared memory and writes each CTA's half of
// the accumulator into that CTA's TMEM.  Shared::cluster addresses of the two CTAs differ in bit 24:
// clearing it turns "my barrier" into "the leader's barrier at the same offset".
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx_leader(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(smem_u32(bar) & PEER_BIT_MASK),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                                 int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(m), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    const uint32_t z = 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
__device__ __forceinline__ void umma_f8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    const uint32_t z = 0u;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
                     "r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

// Arrives on the mbarrier when all tcgen05.mma issued so far by this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                     "r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31},"
        "[%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled shared-memory operand descriptor (rows of 64 16-bit elements = 128 B,
// 8-row swizzle atoms 1024 B apart). Field layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);       // start address
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for SW128 K-major)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}

// Two fp32 -> one packed fp16x2 word (a0 in the low half), round-to-nearest with saturation to
// +-65504: a single full-rate F2FP.SATFINITE.F16.F32.PACK_AB instead of two clamps and two quarter-rate F2F.
__device__ __forceinline__ uint32_t pack_half2_sat(float a0, float a1) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a1), "f"(a0));
    return r;
}
__device__ __forceinline__ float2 unpack_half2(uint32_t h) {
    return __half22float2(*reinterpret_cast<const __half2*>(&h));
}
// hi = fp16(a), lo = fp16(a - hi) for two values
__device__ __forceinline__ void split_half2_sat(float a0, float a1, uint32_t& hi, uint32_t& lo) {
    hi = pack_half2_sat(a0, a1);
    const float2 f = unpack_half2(hi);
    lo = pack_half2_sat(a0 - f.x, a1 - f.y);
}
// two fp32 -> packed e5m2x2 (a0 in the low byte), saturating
__device__ __forceinline__ uint32_t pack_e5m2x2_sat(float a0, float a1) {
    uint16_t r;
    asm("cvt.rn.satfinite.e5m2x2.f32 %0, %1, %2;" : "=h"(r) : "f"(a1), "f"(a0));
    return (uint32_t)r;
}

// 32 x 32 fp32 transpose tile of one warp in shared memory (4 KB): the 16-byte granule g of row r sits at
// granule g ^ (r & 7).  A row write (lane = row, 8 x 128 bit) and the transposed read (8 lanes share a
// row, one granule each) are both conflict-free 128-bit accesses.
__device__ __forceinline__ void tile_put_row(float* T, int row, const uint32_t (&v)[32]) {
#pragma unroll
    for (int g = 0; g < 8; ++g)
        *reinterpret_cast<uint4*>(T + row * 32 + ((g ^ (row & 7)) << 2)) =
            make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
}
__device__ __forceinline__ float4 tile_get4(const float* T, int row, int g) {
    return *reinterpret_cast<const float4*>(T + row * 32 + ((g ^ (row & 7)) << 2));
}

// 2^e with amax * 2^e in [2^(top-1), 2^top): the power-of-two operand scale that keeps the hi plane
// of an fp16 split well inside the normal range (and the lo plane out of the subnormals).
__device__ __forceinline__ float pow2_scale_for(float amax, int top) {
    if (!(amax > 0.f) || isinf(amax)) return 1.f;
    int ex;
    frexpf(amax, &ex);  // amax = f * 2^ex, f in [0.5, 1)
    return ldexpf(1.f, top - ex);
}
// max over non-negative floats through their (order-preserving) bit patterns
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
    atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// ----------------------------------------------------------------------------------------------
// Counter-based N(0,1) noise for NoiseInjection (normalization.py:299-304).  The value of element
// e (linear NHWC index) of a noise tensor is a pure function of (seed, e): Philox4x32-7 on the
// counter e/4 gives the 4 values of channels e..e+3 through two Box-Muller pairs.  Every kernel
// that needs the tensor (statistics, K1, K2's epilogue, the backward passes) regenerates it, so it
// is never written to or read from HBM.  dsee_noise_fill materialises the same stream for tests.
// ----------------------------------------------------------------------------------------------
// Rounds: 7 is the smallest round count of Philox4x32 that is Crush-resistant (Salmon et al.,
// "Parallel random numbers: as easy as 1, 2, 3", SC'11, table 2); cuRAND's default of 10 adds safety
// margin the noise injection does not need, and the generator sits in compute-bound epilogues.
constexpr int NOISE_PHILOX_ROUNDS = 7;
__device__ __forceinline__ unsigned long long eff_noise_seed(unsigned long long seed,
                                                             const unsigned long long* epoch) {
    return seed ? seed + __ldg(epoch) * 0x9E3779B97F4A7C15ull : 0ull;
}
__device__ __forceinline__ float4 noise_normal4(unsigned long long seed, unsigned long long idx4) {
    uint32_t c0 = (uint32_t)idx4, c1 = (uint32_t)(idx4 >> 32), c2 = 0u, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < NOISE_PHILOX_ROUNDS; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    const float k = 2.3283064365386963e-10f;  // 2^-32
    const float u0 = ((float)c0 + 0.5f) * k, u1 = (float)c1 * k;
    const float u2 = ((float)c2 + 0.5f) * k, u3 = (float)c3 * k;
    const float r0 = sqrtf(-2.f * __logf(fminf(u0, 1.f))), r1 = sqrtf(-2.f * __logf(fminf(u2, 1.f)));
    float s0, q0, s1, q1;
    __sincosf(6.2831853071795865f * u1 - 3.1415926535897932f, &s0, &q0);
    __sincosf(6.2831853071795865f * u3 - 3.1415926535897932f, &s1, &q1);
    return make_float4(r0 * q0, r0 * s0, r1 * q1, r1 * s1);
}
// 4 consecutive channels of a noise tensor starting at element e (a multiple of 4): from HBM when an
// explicit tensor is given, regenerated from the seed otherwise.
__device__ __forceinline__ float4 load_noise4(const float* noise, unsigned long long seed, size_t e) {
    if (noise) return __ldg(reinterpret_cast<const float4*>(noise + e));
    return noise_normal4(seed, (unsigned long long)(e >> 2));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif  // __CUDACC__

}  // namespace dsee
