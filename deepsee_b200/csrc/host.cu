// Host-side plumbing of the C ABI: error string, device check, tensor-map encoding, launch count.
#include "common.cuh"
#include "launch_count.h"
#include "../../include/deepsee_b200.h"

#include <atomic>
#include <mutex>

namespace dsee {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

int require_sm100() {
    static std::mutex mu;
    static int cached[64];
    static bool have[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("cudaGetDevice failed: %s (no CUDA device; this library has no CPU fallback)",
                  cudaGetErrorString(e));
        return -2;
    }
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 64 && have[dev]) {
        if (cached[dev] != 10) {
            set_error("device %d is sm_%d0, this library is built for sm_100a only", dev,
                      cached[dev]);
            return -2;
        }
        return 0;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
        return -2;
    }
    if (dev < 64) {
        cached[dev] = major;
        have[dev] = true;
    }
    if (major != 10) {
        set_error("device %d is sm_%d0, this library is built for sm_100a only", dev, major);
        return -2;
    }
    return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    });
    return fn;
}

static int encode_tmap_any(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank,
                           const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                           const uint32_t* elem_strides);

int encode_tmap_16b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, bool bf16,
                    const uint32_t* elem_strides) {
    return encode_tmap_any(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                           base, rank, dims, strides_bytes, box, elem_strides);
}

int encode_tmap_8b(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box) {
    return encode_tmap_any(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, base, rank, dims, strides_bytes, box,
                           nullptr);
}

static int encode_tmap_any(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank,
                           const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                           const uint32_t* elem_strides) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return -3;
    }
    if (((uintptr_t)base & 15) != 0) {
        set_error("TMA base pointer %p is not 16-byte aligned", base);
        return -1;
    }
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        es[i] = elem_strides ? elem_strides[i] : 1;
    }
    for (int i = 0; i < rank - 1; ++i) {
        s[i] = strides_bytes[i];
        if (s[i] % 16 != 0) {
            set_error("TMA stride %llu is not a multiple of 16 bytes", (unsigned long long)s[i]);
            return -1;
        }
    }
    CUresult r = enc(out, dtype,
                     (cuuint32_t)rank, const_cast<void*>(base), d, s, b, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)",
                  (int)r, rank, (unsigned long long)d[0], (unsigned long long)d[1], b[0], b[1]);
        return -3;
    }
    return 0;
}

__device__ unsigned long long g_noise_epoch = 0ull;
__global__ void noise_epoch_kernel(unsigned long long* e, unsigned long long add, int set) {
    *e = set ? add : *e + add;
}

const unsigned long long* noise_epoch_ptr() {
    static std::mutex mu;
    static unsigned long long* cached[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 64 && cached[dev]) return cached[dev];
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_noise_epoch) != cudaSuccess) return nullptr;
    if (dev < 64) cached[dev] = (unsigned long long*)p;
    return (unsigned long long*)p;
}

}  // namespace dsee

extern "C" int dsee_noise_epoch_advance(void* stream) {
    int rc = dsee::require_sm100();
    if (rc) return rc;
    unsigned long long* e = const_cast<unsigned long long*>(dsee::noise_epoch_ptr());
    DSEE_CHECK_ARG(e != nullptr, "noise epoch symbol not found");
    dsee::noise_epoch_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(e, 1ull, 0);
    dsee::count_launch();
    DSEE_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int dsee_noise_epoch_set(unsigned long long value, void* stream) {
    int rc = dsee::require_sm100();
    if (rc) return rc;
    unsigned long long* e = const_cast<unsigned long long*>(dsee::noise_epoch_ptr());
    DSEE_CHECK_ARG(e != nullptr, "noise epoch symbol not found");
    dsee::noise_epoch_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(e, value, 1);
    dsee::count_launch();
    DSEE_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int dsee_version(void) { return DSEE_ABI_VERSION; }
extern "C" const char* dsee_last_error(void) { return dsee::get_error(); }
extern "C" int64_t dsee_launch_count(void) { return dsee::launch_count(); }
