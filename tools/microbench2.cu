// Issue-rate micro-benchmarks behind the row-kernel design (round 2): packed FP32 (FFMA2), FP64, MUFU, F2F and
// their mixes, per SM clock.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench2 microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

typedef unsigned long long u64;
#define ILP 8

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

// mode bits: 1 FFMA, 2 FFMA2, 4 DFMA, 8 MUFU.RSQ, 16 F2F f32->f64, 32 LDS.128, 64 FADD (alu?), 128 SHFL
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_mix(float *out, long long *cyc, int iters) {
    __shared__ float4 sm[512];
    float f[ILP], g[ILP];
    u64 p[ILP];
    double d[ILP];
    float m[ILP];
    double c[ILP];
    float4 l[2];
    float s[ILP];
    sm[threadIdx.x] = make_float4(1.f, 2.f, 3.f, 4.f);
    l[0] = l[1] = make_float4(0, 0, 0, 0);
    for (int k = 0; k < ILP; k++) {
        f[k] = threadIdx.x * 1e-3f + k; g[k] = 1.0f + k; p[k] = pk(f[k], f[k] + 1.f); d[k] = f[k]; m[k] = 1.0f + f[k]; c[k] = 0; s[k] = f[k];
    }
    const float a = 0.999f, b = 1e-3f;
    const u64 pa = pk(a, a), pb = pk(b, b);
    const double da = 0.999, db = 1e-3;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) {
            if (MODE & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(a), "f"(b));
            if (MODE & 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[k]) : "l"(pa), "l"(pb));
            if (MODE & 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(da), "d"(db));
            if (MODE & 8) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[k]));
            if (MODE & 16) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(c[k]) : "f"(g[k]));
            if (MODE & 64) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(g[k]) : "f"(b));
            if (MODE & 128) s[k] = __shfl_xor_sync(0xffffffffu, s[k], 1);
        }
        if (MODE & 32) {
            float4 v0 = sm[(threadIdx.x + i) & 511], v1 = sm[(threadIdx.x + 2 * i + 7) & 511];
            l[0].x += v0.x; l[1].x += v1.y;
        }
    }
    long long t1 = clock64();
    float acc = l[0].x + l[1].x;
    for (int k = 0; k < ILP; k++) {
        float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[k]));
        acc += f[k] + lo + hi + (float)d[k] + m[k] + (float)c[k] + g[k] + s[k];
    }
    if (acc == 12345.678f) out[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int nops_per_iter_lane /* counted unit ops per k */) {
    float *out; long long *cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    k_mix<MODE><<<148, 512>>>(out, cyc, iters);
    k_mix<MODE><<<148, 512>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double c = (double)h[74];
    // warp-instructions per SM clock (16 warps per SM)
    printf("%-34s %8.0f cyc  %6.3f warp-inst/clk/SM  (%5.2f clk per warp-inst per SMSP)\n", name, c,
           16.0 * iters * ILP * nops_per_iter_lane / c, c / (4.0 * iters * ILP * nops_per_iter_lane));
    cudaFree(out); cudaFree(cyc);
}

// coalesced RED: every warp adds 32 consecutive doubles / floats; rows rotate over an L2-resident array
__global__ void k_red_rows(double *a, int nrows, int iters, long long *cyc) {
    const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned s = w * 2654435761u + 12345u;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        atomicAdd(&a[(size_t)((s >> 8) % nrows) * 32 + lane], 1.0);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_red_rows32(float *a, int nrows, int iters, long long *cyc) {
    const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned s = w * 2654435761u + 12345u;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        atomicAdd(&a[(size_t)((s >> 8) % nrows) * 32 + lane], 1.0f);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// plain coalesced store/load-add of per-pair partial forces (the alternative to RED): STG.64 rows
__global__ void k_stg_rows(double *a, int nrows, int iters, long long *cyc) {
    const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned s = w * 2654435761u + 12345u;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        s = s * 1664525u + 1013904223u;
        a[(size_t)((s >> 8) % nrows) * 32 + lane] = (double)i;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    run<1>("FFMA", 1);
    run<2>("FFMA2", 1);
    run<64>("FADD", 1);
    run<4>("DFMA", 1);
    run<8>("MUFU.RSQ", 1);
    run<16>("F2F.F64.F32", 1);
    run<128>("SHFL", 1);
    run<1 | 64>("FFMA + FADD", 2);
    run<2 | 64>("FFMA2 + FADD", 2);
    run<1 | 4>("FFMA + DFMA", 2);
    run<2 | 4>("FFMA2 + DFMA", 2);
    run<1 | 8>("FFMA + MUFU", 2);
    run<2 | 4 | 8>("FFMA2 + DFMA + MUFU", 3);
    run<2 | 4 | 8 | 16>("FFMA2 + DFMA + MUFU + F2F", 4);
    run<1 | 4 | 8 | 16>("FFMA + DFMA + MUFU + F2F", 4);
    run<2 | 32>("FFMA2 + 2 LDS.128 per 8", 1);
    run<1 | 128>("FFMA + SHFL", 2);

    long long *cyc; cudaMalloc(&cyc, 148 * 8 * 8);
    for (int nrows : {1152, 36864}) {
        double *a; cudaMalloc(&a, (size_t)nrows * 32 * 8); cudaMemset(a, 0, (size_t)nrows * 32 * 8);
        const int iters = 512, blocks = 148 * 4, threads = 256;
        std::vector<long long> h(blocks);
        k_red_rows<<<blocks, threads>>>(a, nrows, iters, cyc); k_red_rows<<<blocks, threads>>>(a, nrows, iters, cyc);
        cudaMemcpy(h.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost); std::sort(h.begin(), h.end());
        printf("RED.F64 warp-coalesced rows (%d rows): %.2f clk per warp-RED per SM (32 warps/SM resident)\n", nrows,
               (double)h[blocks / 2] / (iters * 32.0));
        k_red_rows32<<<blocks, threads>>>((float *)a, nrows, iters, cyc); k_red_rows32<<<blocks, threads>>>((float *)a, nrows, iters, cyc);
        cudaMemcpy(h.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost); std::sort(h.begin(), h.end());
        printf("RED.F32 warp-coalesced rows (%d rows): %.2f clk per warp-RED per SM\n", nrows, (double)h[blocks / 2] / (iters * 32.0));
        k_stg_rows<<<blocks, threads>>>(a, nrows, iters, cyc); k_stg_rows<<<blocks, threads>>>(a, nrows, iters, cyc);
        cudaMemcpy(h.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost); std::sort(h.begin(), h.end());
        printf("STG.64 warp-coalesced rows (%d rows): %.2f clk per warp-STG per SM\n", nrows, (double)h[blocks / 2] / (iters * 32.0));
        cudaFree(a);
    }
    return 0;
}
