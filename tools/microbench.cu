// Micro-benchmarks that decide kernel design: FP64/FP32 atomic (RED) throughput on a small L2-resident
// array, L2 load latency, MUFU and F2F conversion rates.  Build: nvcc -arch=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void k_red64(double *a, int n, int iters) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        atomicAdd(&a[(s >> 8) % n], 1.0);
    }
}
__global__ void k_red32(float *a, int n, int iters) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        atomicAdd(&a[(s >> 8) % n], 1.0f);
    }
}
// neighbouring lanes hit neighbouring addresses (3 doubles per atom, 9 consecutive per water)
__global__ void k_red64_coalesced(double *a, int n, int iters) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) / 8 * 2654435761u + 12345u;
    int l = threadIdx.x & 7;
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        atomicAdd(&a[((s >> 8) % (n / 8)) * 8 + l], 1.0);
    }
}
__global__ void k_chase(const int *p, int *out, int iters) {
    int j = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) j = p[j];
    long long t1 = clock64();
    out[0] = j;
    out[1] = (int)((t1 - t0) / iters);
}
__global__ void k_cvt(double *out, int iters) {
    double v = threadIdx.x * 1e-3 + 1.0;
    float acc = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) { acc += (float)(v + k); v += 1e-9; }
    }
    if (acc == 12345.f) out[0] = acc;
}
__global__ void k_rsq(float *out, int iters) {
    float v[8];
    for (int k = 0; k < 8; k++) v[k] = threadIdx.x + k + 1.0f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = rsqrtf(v[k]) + 1.0f;
    }
    float s = 0;
    for (int k = 0; k < 8; k++) s += v[k];
    if (s == 12345.f) out[0] = s;
}

template <typename F>
float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    const int n = 36864;   // ~ 3*natom of the 12k-atom system
    double *a; float *b;
    cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 4);
    cudaMemset(a, 0, n * 8); cudaMemset(b, 0, n * 4);
    const int blocks = 148 * 8, threads = 256, iters = 256;
    const double nops = (double)blocks * threads * iters;
    float ms = timeit([&] { k_red64<<<blocks, threads>>>(a, n, iters); });
    printf("RED.F64 random over %d doubles: %.1f G atomics/s (%.3f ms)\n", n, nops / ms / 1e6, ms);
    ms = timeit([&] { k_red64_coalesced<<<blocks, threads>>>(a, n, iters); });
    printf("RED.F64 8-consecutive groups:   %.1f G atomics/s (%.3f ms)\n", nops / ms / 1e6, ms);
    ms = timeit([&] { k_red32<<<blocks, threads>>>(b, n, iters); });
    printf("RED.F32 random over %d floats:  %.1f G atomics/s (%.3f ms)\n", n, nops / ms / 1e6, ms);
    for (int big : {1 << 22}) {
        double *c; cudaMalloc(&c, (size_t)big * 8); cudaMemset(c, 0, (size_t)big * 8);
        ms = timeit([&] { k_red64<<<blocks, threads>>>(c, big, iters); });
        printf("RED.F64 random over %d doubles: %.1f G atomics/s\n", big, nops / ms / 1e6);
        cudaFree(c);
    }
    // pointer chase, L2-resident (4 MB) and L1-resident (16 KB)
    for (int sz : {4096, 1 << 20}) {
        std::vector<int> h(sz);
        for (int i = 0; i < sz; i++) h[i] = (int)(((long long)i * 40503 + 7919) % sz);
        int *p, *o; cudaMalloc(&p, sz * 4); cudaMalloc(&o, 8);
        cudaMemcpy(p, h.data(), sz * 4, cudaMemcpyHostToDevice);
        k_chase<<<1, 1>>>(p, o, 2000); k_chase<<<1, 1>>>(p, o, 2000);
        int r[2]; cudaMemcpy(r, o, 8, cudaMemcpyDeviceToHost);
        printf("dependent load latency, %d KB working set: %d cycles\n", sz * 4 / 1024, r[1]);
        cudaFree(p); cudaFree(o);
    }
    ms = timeit([&] { k_cvt<<<blocks, threads>>>(a, 1024); });
    printf("F2F f64->f32 (+DADD): %.1f G conv/s per SM-clock: %.2f conv/clk/SM\n", (double)blocks * threads * 1024 * 8 / ms / 1e6,
           (double)blocks * threads * 1024 * 8 / (ms * 1e-3) / 148 / 1.965e9);
    ms = timeit([&] { k_rsq<<<blocks, threads>>>(b, 1024); });
    printf("MUFU.RSQ: %.2f /clk/SM\n", (double)blocks * threads * 1024 * 8 / (ms * 1e-3) / 148 / 1.965e9);
    return 0;
}
