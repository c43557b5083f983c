// Latency calibration for the dataflow sweeps (debug tool): dependent DADD, LDS+DADD groups, pointer chase in L2 / DRAM,
// store+fence, atomicAdd round trip.   nvcc -arch=sm_100a -O3 -fmad=false tools/microbench.cu -o tools/_dbg/microbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_dadd(double* out, double a, int n, long long* cyc) {
    double r = out[0];
    long long t0 = clock64();
    for (int i = 0; i < n; i += 8) {
        r = r + a; r = r + a; r = r + a; r = r + a; r = r + a; r = r + a; r = r + a; r = r + a;
    }
    long long t1 = clock64();
    out[threadIdx.x] = r; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds_group(double* out, int n, long long* cyc) {
    __shared__ double t[512];
    for (int i = threadIdx.x; i < 512; i += 32) t[i] = 1.0 + i;
    __syncwarp();
    double r = out[0];
    long long t0 = clock64();
    for (int t0i = 0; t0i < n; t0i += 8) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = t[(t0i + j) & 511];
#pragma unroll
        for (int j = 0; j < 8; ++j) { r = r + v[j]; v[j] = r; }
        if (threadIdx.x == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) t[(t0i + j) & 511] = v[j];
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = r; if (threadIdx.x == 0) cyc[1] = t1 - t0;
}
__global__ void k_chase(const uint32_t* next, int n, uint32_t* out, long long* cyc, int slot) {
    uint32_t p = 0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) p = __ldcg(&next[p]);
    long long t1 = clock64();
    out[0] = p; cyc[slot] = t1 - t0;
}
__global__ void k_fence(double* buf, unsigned* ctr, int n, long long* cyc) {
    long long t0 = clock64();
    unsigned acc = 0;
    for (int i = 0; i < n; ++i) {
        buf[(i * 4099) & 0xFFFFF] = (double)i;
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        acc += atomicAdd(&ctr[(i * 977) & 0xFFFF], 1u);
    }
    long long t1 = clock64();
    ctr[0x10000] = acc; cyc[4] = t1 - t0;
}
__global__ void k_fence_sc(double* buf, unsigned* ctr, int n, long long* cyc) {
    long long t0 = clock64();
    unsigned acc = 0;
    for (int i = 0; i < n; ++i) {
        buf[(i * 4099) & 0xFFFFF] = (double)i;
        __threadfence();
        acc += atomicAdd(&ctr[(i * 977) & 0xFFFF], 1u);
    }
    long long t1 = clock64();
    ctr[0x10000] = acc; cyc[5] = t1 - t0;
}
__global__ void k_atomic(unsigned* ctr, int n, long long* cyc) {
    long long t0 = clock64();
    unsigned acc = 0;
    for (int i = 0; i < n; ++i) acc = atomicAdd(&ctr[(acc + i * 977) & 0xFFFF], 1u);
    long long t1 = clock64();
    ctr[0x10000] = acc; cyc[6] = t1 - t0;
}
int main() {
    long long* cyc; cudaMallocManaged(&cyc, 64 * 8);
    double* d; cudaMalloc(&d, 8 << 20); cudaMemset(d, 0, 8 << 20);
    unsigned* ctr; cudaMalloc(&ctr, 4 * 0x10004); cudaMemset(ctr, 0, 4 * 0x10004);
    const int N = 4096;
    for (int rep = 0; rep < 2; ++rep) { k_dadd<<<1, 32>>>(d, 1.5, N, cyc); k_lds_group<<<1, 32>>>(d, N, cyc); }
    cudaDeviceSynchronize();
    printf("dependent DADD: %.1f cycles each\n", (double)cyc[0] / N);
    printf("LDS x8 + DADD x8 group chain: %.1f cycles per term\n", (double)cyc[1] / N);
    // pointer chase: random cycle over m entries
    for (int pass = 0; pass < 2; ++pass) {
        size_t m = pass == 0 ? (4u << 20) : (128u << 20);  // 16 MB (L2) / 512 MB (DRAM)
        uint32_t* h = (uint32_t*)malloc(m * 4);
        // simple LCG permutation cycle: next[i] = (i * a + c) mod m with m power of two, a = 4k+1, c odd -> full cycle
        for (size_t i = 0; i < m; ++i) h[i] = (uint32_t)((i * 1664525ull + 1013904223ull) & (m - 1));
        uint32_t* dn; cudaMalloc(&dn, m * 4); cudaMemcpy(dn, h, m * 4, cudaMemcpyHostToDevice);
        uint32_t* o; cudaMalloc(&o, 4);
        k_chase<<<1, 1>>>(dn, 2000, o, cyc, 2 + pass); cudaDeviceSynchronize();
        k_chase<<<1, 1>>>(dn, 2000, o, cyc, 2 + pass); cudaDeviceSynchronize();
        printf("pointer chase over %zu MB: %.1f cycles per load\n", m * 4 >> 20, (double)cyc[2 + pass] / 2000);
        cudaFree(dn); free(h);
    }
    k_fence<<<1, 1>>>(d, ctr, 1000, cyc); k_fence_sc<<<1, 1>>>(d, ctr, 1000, cyc); k_atomic<<<1, 1>>>(ctr, 1000, cyc);
    cudaDeviceSynchronize();
    printf("store + fence.acq_rel.gpu + atomicAdd: %.1f cycles\n", (double)cyc[4] / 1000);
    printf("store + __threadfence + atomicAdd:     %.1f cycles\n", (double)cyc[5] / 1000);
    printf("dependent atomicAdd round trip:        %.1f cycles\n", (double)cyc[6] / 1000);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock rate attr %d kHz\n", clk);
    return 0;
}
