// fl_rt.h -- thin runtime layer under the solver.
//
// Normal build (nvcc, sm_100a): CUDA runtime, real kernels, CUB.
// FL_EMU build (g++ -DFL_EMU): the SAME kernel bodies and the SAME host orchestration run serially on
// the host, one "thread" after another.  The emu build exists only so that the CPU-only test tier
// (pytest -m "not gpu") can check the solver's logic against the oracle before GPU time is spent;
// it is built into tests/_emu/, never shipped, and the product loader (fastlem_b200/_native.py)
// never looks for it.  Kernels that need block-level cooperation are not available under FL_EMU.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifdef FL_EMU
// ------------------------------------------------------------------------------------------------
#include <algorithm>
#include <chrono>
#include <cmath>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define FL_DEVICE_BUILD 0

struct fl_dim3 {
    unsigned x, y, z;
    fl_dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
typedef fl_dim3 dim3;
extern thread_local fl_dim3 threadIdx, blockIdx, blockDim, gridDim;

struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v; v.x = x; v.y = y; return v; }
typedef int cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

struct fl_event { std::chrono::steady_clock::time_point t; };
typedef fl_event* cudaEvent_t;

template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline void __threadfence() {}
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }

template <class F> inline void fl_emu_launch(unsigned grid, unsigned block, F&& f) {
    gridDim = fl_dim3(grid);
    blockDim = fl_dim3(block);
    for (unsigned b = 0; b < grid; ++b) {
        blockIdx = fl_dim3(b);
        for (unsigned t = 0; t < block; ++t) {
            threadIdx = fl_dim3(t);
            f();
        }
    }
}
#define FL_LAUNCH(kernel, grid, block, stream, ...) \
    do { (void)(stream); fl_emu_launch((grid), (block), [&] { kernel(__VA_ARGS__); }); } while (0)
#define FL_RANGE(name) ((void)0)

inline cudaError_t fl_set_device(int) { return 0; }
inline cudaError_t fl_stream_create(cudaStream_t* s) { *s = 0; return 0; }
inline cudaError_t fl_stream_destroy(cudaStream_t) { return 0; }
inline cudaError_t fl_stream_sync(cudaStream_t) { return 0; }
inline cudaError_t fl_malloc(void** p, size_t bytes) { *p = std::malloc(bytes ? bytes : 1); return *p ? 0 : 2; }
inline cudaError_t fl_free(void* p, cudaStream_t = 0) { std::free(p); return 0; }
inline cudaError_t fl_trim_pool() { return 0; }
inline cudaError_t fl_malloc_host(void** p, size_t bytes) { *p = std::malloc(bytes ? bytes : 1); return *p ? 0 : 2; }
inline cudaError_t fl_free_host(void* p) { std::free(p); return 0; }
inline cudaError_t fl_h2d(void* d, const void* h, size_t bytes, cudaStream_t) { std::memcpy(d, h, bytes); return 0; }
inline cudaError_t fl_d2h(void* h, const void* d, size_t bytes, cudaStream_t) { std::memcpy(h, d, bytes); return 0; }
inline cudaError_t fl_d2d(void* d, const void* s, size_t bytes, cudaStream_t) { std::memcpy(d, s, bytes); return 0; }
inline cudaError_t fl_memset(void* d, int v, size_t bytes, cudaStream_t) { std::memset(d, v, bytes); return 0; }
inline cudaError_t fl_event_create(cudaEvent_t* e) { *e = new fl_event; return 0; }
inline cudaError_t fl_event_destroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t fl_event_record(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return 0; }
inline cudaError_t fl_event_sync(cudaEvent_t) { return 0; }
inline cudaError_t fl_event_elapsed(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return 0;
}
inline cudaError_t fl_last_error() { return 0; }
inline int fl_sm_count() { return 4; }

// stable LSD sort of (key,value) pairs on the low `end_bit` bits, like cub::DeviceRadixSort::SortPairs
inline cudaError_t fl_sort_pairs(void*, size_t& temp_bytes, const uint32_t* kin, uint32_t* kout, const uint32_t* vin,
                                 uint32_t* vout, uint32_t n, int end_bit, cudaStream_t, bool query) {
    if (query) { temp_bytes = 1; return 0; }
    std::vector<uint32_t> idx(n);
    for (uint32_t i = 0; i < n; ++i) idx[i] = i;
    uint32_t mask = end_bit >= 32 ? 0xFFFFFFFFu : ((1u << end_bit) - 1u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return (kin[a] & mask) < (kin[b] & mask); });
    for (uint32_t i = 0; i < n; ++i) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
    return 0;
}

inline cudaError_t fl_exclusive_sum(void*, size_t& temp_bytes, const uint32_t* in, uint32_t* out, uint32_t n,
                                    cudaStream_t, bool query) {
    if (query) { temp_bytes = 1; return 0; }
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n; ++i) { uint32_t v = in[i]; out[i] = acc; acc += v; }
    return 0;
}

inline cudaError_t fl_inclusive_max(void*, size_t& temp_bytes, const uint32_t* in, uint32_t* out, uint32_t n,
                                    cudaStream_t, bool query) {
    if (query) { temp_bytes = 1; return 0; }
    uint32_t acc = 0;
    for (uint32_t i = 0; i < n; ++i) { acc = in[i] > acc ? in[i] : acc; out[i] = acc; }
    return 0;
}

// 64-bit keys (flood order on the device, fl_floodgpu.cuh)
inline cudaError_t fl_sort_pairs64(void*, size_t& temp_bytes, const unsigned long long* kin, unsigned long long* kout,
                                   const uint32_t* vin, uint32_t* vout, uint32_t n, int end_bit, cudaStream_t,
                                   bool query) {
    if (query) { temp_bytes = 1; return 0; }
    std::vector<uint32_t> idx(n);
    for (uint32_t i = 0; i < n; ++i) idx[i] = i;
    unsigned long long mask = end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1ull);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return (kin[a] & mask) < (kin[b] & mask); });
    for (uint32_t i = 0; i < n; ++i) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
    return 0;
}
inline cudaError_t fl_sort_keys64(void*, size_t& temp_bytes, const unsigned long long* kin, unsigned long long* kout,
                                  uint32_t n, cudaStream_t, bool query) {
    if (query) { temp_bytes = 1; return 0; }
    std::vector<unsigned long long> v(kin, kin + n);
    std::sort(v.begin(), v.end());
    for (uint32_t i = 0; i < n; ++i) kout[i] = v[i];
    return 0;
}
inline cudaError_t fl_exclusive_sum64(void*, size_t& temp_bytes, const unsigned long long* in, unsigned long long* out,
                                      uint32_t n, cudaStream_t, bool query) {
    if (query) { temp_bytes = 1; return 0; }
    unsigned long long acc = 0;
    for (uint32_t i = 0; i < n; ++i) { unsigned long long v = in[i]; out[i] = acc; acc += v; }
    return 0;
}

#else
// ------------------------------------------------------------------------------------------------
#include <cuda_runtime.h>
#include <mutex>
#include <nvtx3/nvToolsExt.h>  // header-only; a no-op unless a profiler (Nsight Systems / Compute) is attached
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#define FL_DEVICE_BUILD 1

#define FL_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)

// NVTX range over a scope: the API calls and the stages of an iteration show up named on a profiler's timeline
struct FlRange {
    explicit FlRange(const char* name) { nvtxRangePushA(name); }
    ~FlRange() { nvtxRangePop(); }
    FlRange(const FlRange&) = delete;
    FlRange& operator=(const FlRange&) = delete;
};
#define FL_RANGE(name) FlRange fl_range_##__LINE__(name)

inline cudaError_t fl_set_device(int d) { return cudaSetDevice(d); }
inline cudaError_t fl_stream_create(cudaStream_t* s) { return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking); }
inline cudaError_t fl_stream_destroy(cudaStream_t s) { return cudaStreamDestroy(s); }
inline cudaError_t fl_stream_sync(cudaStream_t s) { return cudaStreamSynchronize(s); }
// Device memory comes from a stream-ordered pool the library keeps per device (cudaMemPool, release threshold = never):
// a context's buffers go back to the pool when it is destroyed and the next context (the next `generate()` of a caller
// that creates one per call, as the Rust shim does) gets them back without a trip to the driver.  cudaMalloc / cudaFree
// of the ~0.5 GB slab of a 1M-site model were measured at up to 0.3 s EACH on some boxes (profiles/r2s_bench_1M.json:
// destroy 0.32 s) against 0.62 s for the whole solve.  fastlem_trim_memory() returns the cached memory to the driver;
// FASTLEM_NO_POOL=1 goes back to cudaMalloc / cudaFree.  A buffer is freed in the order of its owner's stream
// (cudaFreeAsync), so it is never handed out again while a kernel still uses it.
struct FlPool {
    cudaMemPool_t pool = nullptr;
    cudaStream_t stream = nullptr;
    int state = 0;  // 0 = not tried, 1 = in use, -1 = unavailable
};
inline FlPool* fl_pool() {
    static std::mutex mu;
    static FlPool pools[64];
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    FlPool& P = pools[d];
    if (P.state == 0) {
        P.state = -1;
        const char* off = std::getenv("FASTLEM_NO_POOL");
        if (!(off && off[0] == '1')) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = d;
            unsigned long long keep = ~0ull;
            if (cudaMemPoolCreate(&P.pool, &props) == cudaSuccess &&
                cudaMemPoolSetAttribute(P.pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess &&
                cudaStreamCreateWithFlags(&P.stream, cudaStreamNonBlocking) == cudaSuccess)
                P.state = 1;
            else
                (void)cudaGetLastError();
        }
    }
    return P.state == 1 ? &P : nullptr;
}
inline cudaError_t fl_malloc(void** p, size_t bytes) {
    if (FlPool* P = fl_pool()) {
        cudaError_t e = cudaMallocFromPoolAsync(p, bytes ? bytes : 1, P->pool, P->stream);
        if (e != cudaSuccess) return e;
        return cudaStreamSynchronize(P->stream);
    }
    return cudaMalloc(p, bytes ? bytes : 1);
}
// `owner`: the stream whose work last used the buffer -- the free is ordered behind that work
inline cudaError_t fl_free(void* p, cudaStream_t owner) {
    if (fl_pool()) return cudaFreeAsync(p, owner);
    return cudaFree(p);
}
inline cudaError_t fl_trim_pool() {
    if (FlPool* P = fl_pool()) {
        cudaError_t e = cudaStreamSynchronize(P->stream);
        return e != cudaSuccess ? e : cudaMemPoolTrimTo(P->pool, 0);
    }
    return cudaSuccess;
}
inline cudaError_t fl_malloc_host(void** p, size_t bytes) { return cudaMallocHost(p, bytes ? bytes : 1); }
inline cudaError_t fl_free_host(void* p) { return cudaFreeHost(p); }
inline cudaError_t fl_h2d(void* d, const void* h, size_t bytes, cudaStream_t s) {
    return cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s);
}
inline cudaError_t fl_d2h(void* h, const void* d, size_t bytes, cudaStream_t s) {
    return cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s);
}
inline cudaError_t fl_d2d(void* d, const void* s_, size_t bytes, cudaStream_t s) {
    return cudaMemcpyAsync(d, s_, bytes, cudaMemcpyDeviceToDevice, s);
}
inline cudaError_t fl_memset(void* d, int v, size_t bytes, cudaStream_t s) { return cudaMemsetAsync(d, v, bytes, s); }
inline cudaError_t fl_event_create(cudaEvent_t* e) { return cudaEventCreate(e); }
inline cudaError_t fl_event_destroy(cudaEvent_t e) { return cudaEventDestroy(e); }
inline cudaError_t fl_event_record(cudaEvent_t e, cudaStream_t s) { return cudaEventRecord(e, s); }
inline cudaError_t fl_event_sync(cudaEvent_t e) { return cudaEventSynchronize(e); }
inline cudaError_t fl_event_elapsed(float* ms, cudaEvent_t a, cudaEvent_t b) { return cudaEventElapsedTime(ms, a, b); }
inline cudaError_t fl_last_error() { return cudaGetLastError(); }
inline int fl_sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

inline cudaError_t fl_sort_pairs(void* temp, size_t& temp_bytes, const uint32_t* kin, uint32_t* kout,
                                 const uint32_t* vin, uint32_t* vout, uint32_t n, int end_bit, cudaStream_t s,
                                 bool query) {
    return cub::DeviceRadixSort::SortPairs(query ? nullptr : temp, temp_bytes, kin, kout, vin, vout, (int)n, 0, end_bit,
                                           s);
}

inline cudaError_t fl_exclusive_sum(void* temp, size_t& temp_bytes, const uint32_t* in, uint32_t* out, uint32_t n,
                                    cudaStream_t s, bool query) {
    return cub::DeviceScan::ExclusiveSum(query ? nullptr : temp, temp_bytes, in, out, (int)n, s);
}
struct FlMaxOp {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};
inline cudaError_t fl_inclusive_max(void* temp, size_t& temp_bytes, const uint32_t* in, uint32_t* out, uint32_t n,
                                    cudaStream_t s, bool query) {
    return cub::DeviceScan::InclusiveScan(query ? nullptr : temp, temp_bytes, in, out, FlMaxOp(), (int)n, s);
}
inline cudaError_t fl_sort_pairs64(void* temp, size_t& temp_bytes, const unsigned long long* kin,
                                   unsigned long long* kout, const uint32_t* vin, uint32_t* vout, uint32_t n,
                                   int end_bit, cudaStream_t s, bool query) {
    return cub::DeviceRadixSort::SortPairs(query ? nullptr : temp, temp_bytes, kin, kout, vin, vout, (int)n, 0, end_bit,
                                           s);
}
inline cudaError_t fl_sort_keys64(void* temp, size_t& temp_bytes, const unsigned long long* kin, unsigned long long* kout,
                                  uint32_t n, cudaStream_t s, bool query) {
    return cub::DeviceRadixSort::SortKeys(query ? nullptr : temp, temp_bytes, kin, kout, (int)n, 0, 64, s);
}
inline cudaError_t fl_exclusive_sum64(void* temp, size_t& temp_bytes, const unsigned long long* in,
                                      unsigned long long* out, uint32_t n, cudaStream_t s, bool query) {
    return cub::DeviceScan::ExclusiveSum(query ? nullptr : temp, temp_bytes, in, out, (int)n, s);
}
#endif
