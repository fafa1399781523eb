// Latency-bound exchange of small statistics vectors between the GPUs of one node over NVLink peer
// memory: the all-reduce(sum) of the per-channel batch-norm sums of Sync-BN (one [2, C] vector per
// conditional-norm layer forward, one per backward; sync_batchnorm/batchnorm.py:80-93 does it with
// Python-thread pipes, NCCL needs ~15-25 us per call for 4 KB).
//
// Every rank owns one symmetric allocation (torch.distributed._symmetric_memory), mapped into all
// peers.  Layout per rank:  data  [ring][world][maxn] float   - slot s, sender r: r's vector
//                           flags [ring][world]       u64     - sequence number of the last vector
//                           seq                       u64     - exchanges completed by this rank
// One kernel, one block:
//   1. PUSH: store my vector into slot[s % ring][my rank] of EVERY peer (remote NVLink stores);
//   2. fence.sys, then release-store the sequence number s into each peer's flag for me;
//   3. acquire-spin on my own flags until every peer's s has arrived;
//   4. sum the world's vectors from my OWN memory in rank order (bit-identical on every rank).
// A rank can run at most one exchange ahead of a peer (step 3 of exchange s+1 needs the peer's flag,
// which it only sends after finishing exchange s), so a ring of >= 2 slots is never overwritten
// while it is read.  All values live in device memory, so the kernel replays inside a CUDA graph.
#include "common.cuh"
#include "launch_count.h"
#include "../../include/deepsee_b200.h"

namespace dsee {

constexpr int PEER_MAX_WORLD = 8;

struct PeerArgs {
    unsigned char* buf[PEER_MAX_WORLD];  // every rank's allocation as mapped in this process
    int world, rank, ring, maxn;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(1024) peer_allreduce_small_kernel(const PeerArgs a, const float* in,
                                                                    float* out, int n) {
    const size_t data_bytes = (size_t)a.ring * a.world * a.maxn * sizeof(float);
    const size_t flag_bytes = (size_t)a.ring * a.world * sizeof(unsigned long long);
    unsigned long long* my_seq = reinterpret_cast<unsigned long long*>(a.buf[a.rank] + data_bytes + flag_bytes);
    const unsigned long long s = *my_seq + 1;
    const int slot = (int)(s % (unsigned long long)a.ring);
    // 1. push
    for (int r = 0; r < a.world; ++r) {
        float* dst = reinterpret_cast<float*>(a.buf[r]) + ((size_t)slot * a.world + a.rank) * a.maxn;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = in[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. signal, 3. wait
    if (threadIdx.x < a.world) {
        unsigned long long* pf = reinterpret_cast<unsigned long long*>(a.buf[threadIdx.x] + data_bytes);
        st_release_sys(pf + (size_t)slot * a.world + a.rank, s);
        const unsigned long long* mf = reinterpret_cast<const unsigned long long*>(a.buf[a.rank] + data_bytes) +
                                       (size_t)slot * a.world + threadIdx.x;
        unsigned long long spins = 0;
        while (ld_acquire_sys(mf) < s) {
            if (++spins > (1ull << 31)) {
                printf("dsee: peer exchange watchdog rank %d waiting for rank %d (seq %llu)\n", a.rank,
                       (int)threadIdx.x, s);
                __trap();
            }
        }
    }
    __syncthreads();
    // 4. reduce from local memory (written by the peers: bypass L1)
    const float* mine = reinterpret_cast<const float*>(a.buf[a.rank]) + (size_t)slot * a.world * a.maxn;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < a.world; ++r) acc += __ldcg(mine + (size_t)r * a.maxn + i);
        out[i] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) *my_seq = s;
}

}  // namespace dsee

using namespace dsee;

extern "C" int64_t dsee_peer_exchange_bytes(int world, int ring, int maxn) {
    return (int64_t)ring * world * maxn * 4 + (int64_t)ring * world * 8 + 64;
}

extern "C" int dsee_peer_allreduce_small(const void* const* peer_bufs, int world, int rank, int ring, int maxn,
                                         const float* in, float* out, int n, void* stream) {
    DSEE_CHECK_ARG(peer_bufs && in && out, "NULL pointer");
    DSEE_CHECK_ARG(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "bad world / rank");
    DSEE_CHECK_ARG(ring >= 2 && maxn > 0 && n > 0 && n <= maxn, "bad ring / vector length (n <= maxn)");
    int rc = require_sm100();
    if (rc) return rc;
    PeerArgs a;
    for (int r = 0; r < PEER_MAX_WORLD; ++r) a.buf[r] = r < world ? (unsigned char*)peer_bufs[r] : nullptr;
    for (int r = 0; r < world; ++r) DSEE_CHECK_ARG(a.buf[r] != nullptr, "peer buffer %d is NULL", r);
    a.world = world;
    a.rank = rank;
    a.ring = ring;
    a.maxn = maxn;
    peer_allreduce_small_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(a, in, out, n);
    count_launch();
    DSEE_CUDA(cudaGetLastError());
    return 0;
}
