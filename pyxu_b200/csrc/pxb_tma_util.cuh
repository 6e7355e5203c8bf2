// pxb_tma_util.cuh -- the few PTX wrappers the TMA kernels share (mbarrier + cp.async.bulk.tensor) and the host-side
// tensor-map encoder (cuTensorMapEncodeTiled, resolved through the runtime: no link against libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>

static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

static __device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Spin on the phase parity; a bounded number of polls, then trap: a wrong transaction count must fail the launch
// loudly instead of hanging the device.
static __device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
#ifdef PXB_MBAR_BACKOFF
        __nanosleep(PXB_MBAR_BACKOFF);
#endif
    }
    __trap();
}
static __device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

static __device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
static __device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
static __device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- tensor maps ---------------------------------------------------------------------------------------
typedef CUresult (*PxbEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PxbEncodeTiledFn pxb_encode_fn() {
    static PxbEncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) p = nullptr;
        return (PxbEncodeTiledFn)p;
    }();
    return fn;
}

// rank-`rank` tiled map of a T array: dims / strides in ELEMENTS (stride[0] == 1), box in elements; out-of-bounds
// elements read as zero
template <class T>
static inline bool pxb_tma_encode(int rank, const void* base, const uint64_t* dim, const uint64_t* stride, const uint32_t* box, CUtensorMap* out) {
    PxbEncodeTiledFn fn = pxb_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[5], strides[4];
    cuuint32_t bx[5], estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) { dims[i] = dim[i]; bx[i] = box[i]; }
    for (int i = 1; i < rank; ++i) strides[i - 1] = stride[i] * sizeof(T);
    const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    return fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- launch-side caches ---------------------------------------------------------------------------------
// A solver iterates on the same few (pointer, geometry) pairs -- the two halves of a ping-pong, the boundary / interior
// launches of a slab: their tensor maps are encoded once.  The key holds everything that goes into the encoding, so a
// recycled pointer with another geometry can never hit a stale entry.  Per host thread, no lock.
struct PxbMapKey {
    const void* base;
    uint64_t dim[5], stride[5];
    uint32_t box[5];
    int32_t rank, esize;
};
template <class T>
static inline bool pxb_tma_encode_cached(int rank, const void* base, const uint64_t* dim, const uint64_t* stride, const uint32_t* box, CUtensorMap* out) {
    constexpr int N = 32;
    struct Entry { PxbMapKey key; CUtensorMap map; bool used; };
    static thread_local Entry cache[N];
    static thread_local int next = 0;
    PxbMapKey k;
    memset(&k, 0, sizeof(k));
    k.base = base; k.rank = rank; k.esize = (int32_t)sizeof(T);
    for (int i = 0; i < rank; ++i) { k.dim[i] = dim[i]; k.stride[i] = stride[i]; k.box[i] = box[i]; }
    for (int i = 0; i < N; ++i)
        if (cache[i].used && memcmp(&cache[i].key, &k, sizeof(k)) == 0) { *out = cache[i].map; return true; }
    if (!pxb_tma_encode<T>(rank, base, dim, stride, box, out)) return false;
    Entry& e = cache[next];
    next = (next + 1) % N;
    e.key = k; e.map = *out; e.used = true;
    return true;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the function: raise it when a launch needs more than any
// earlier launch of that kernel from this process did (some kernels size their shared memory at run time), not on every launch
static inline cudaError_t pxb_smem_attr_once(const void* kern, int bytes) {
    constexpr int N = 256;
    static const void* seen[N];
    static int granted[N];
    static int nseen = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    int slot = -1;
    for (int i = 0; i < nseen; ++i)
        if (seen[i] == kern) { slot = i; break; }
    if (slot >= 0 && granted[slot] >= bytes) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    if (slot < 0 && nseen < N) { slot = nseen++; seen[slot] = kern; }
    if (slot >= 0) granted[slot] = bytes;
    return cudaSuccess;
}
