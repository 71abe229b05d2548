// micro-benchmark: issue throughput of sm_100 FFMA2 (fma.rn.f32x2) vs FFMA, alone and mixed with ALU-pipe work.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_issue benchmarks/ffma2_issue.cu && ./ffma2_issue   (results: profiles/r01_ffma2_issue.txt)
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a[8]; u64 p[8]; unsigned x[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; float2 f = make_float2(a[i], a[i] + 1); p[i] = *reinterpret_cast<u64*>(&f); x[i] = threadIdx.x + i; }
    float2 sf = make_float2(s, s); u64 s2 = *reinterpret_cast<u64*>(&sf);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = ffma1(a[i], s, s);                 // 1 FFMA
                if (MODE == 1) p[i] = ffma2(p[i], s2, s2);               // 1 FFMA2
                if (MODE == 2) { a[i] = ffma1(a[i], s, s); x[i] = lop(x[i], 0x5a5a5a5au + i); }   // FFMA + LOP
                if (MODE == 3) { p[i] = ffma2(p[i], s2, s2); x[i] = lop(x[i], 0x5a5a5a5au + i); } // FFMA2 + LOP
                if (MODE == 4) { p[i] = ffma2(p[i], s2, s2); a[i] = ffma1(a[i], s, s); }          // FFMA2 + FFMA
                if (MODE == 5) x[i] = lop(x[i], 0x5a5a5a5au + i);       // LOP only
            }
        }
    }
    float r = 0; for (int i = 0; i < 8; ++i) { float2 f = *reinterpret_cast<float2*>(&p[i]); r += a[i] + f.x + f.y + (float)x[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* out, int ops_per) {
    int iters = 4096; dim3 g(148 * 4), b(512);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<g, b>>>(out, 16, 1.0001f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<g, b>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)g.x * (b.x / 32) * iters * 32.0 * ops_per;  // warp instructions
    printf("%-14s %.3f ms  %.1f G warp-inst/s  (%.2f per SMSP-cycle @1.965GHz)\n", name, ms, winst / ms / 1e6, winst / (ms * 1e-3) / (148 * 4 * 1.965e9));
}
int main() {
    float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
    run<0>("FFMA", out, 1); run<1>("FFMA2", out, 1); run<2>("FFMA+LOP", out, 2); run<3>("FFMA2+LOP", out, 2); run<4>("FFMA2+FFMA", out, 2); run<5>("LOP", out, 1);
    return 0;
}
