// ubench2.cu -- second round of instruction-throughput probes for the fused forward kernel (B200, sm_100a):
// packed fp32x2 (FFMA2 / FADD2, Blackwell-only), directed-rounding adds used as exact floor()s,
// FSET, 3-input integer min/max, and a clean shared-memory LUT gather (LCG index, 2 ALU ops per gather).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu
// Output: thread-operations per clock per SM, measured with clock64() of block 0 (single wave).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8
typedef unsigned long long u64;

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2_rm(u64 a, u64 b) { u64 r; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int OP>
__global__ void __launch_bounds__(256) probe(float *out, const float *in, int n, long long *cycles)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float f[ILP];
    u64 p[ILP];
    int k[ILP];
    const float c1 = in[1], c2 = in[2];
    const u64 pc1 = ((u64)__float_as_uint(c1) << 32) | __float_as_uint(c2), pc2 = ((u64)__float_as_uint(c2) << 32) | __float_as_uint(c1);
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        f[i] = in[(t + i) % n];
        k[i] = (int)(f[i] * 1000.f) + t;
        p[i] = ((u64)__float_as_uint(f[i]) << 32) | (unsigned)k[i];
    }
    long long c0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) f[i] = __fmaf_rn(f[i], c1, c2);                                  // FFMA
            if (OP == 1) p[i] = ffma2(p[i], pc1, pc2);                                    // FFMA2 (2 flops x2 per lane-op)
            if (OP == 2) p[i] = fadd2_rm(p[i], pc1);                                      // FADD2.RM
            if (OP == 3) f[i] = __fadd_rd(f[i], c1);                                      // FADD.RM
            if (OP == 4) f[i] = f[i] < c1 ? __fadd_rn(f[i], 1.0f) : f[i];                 // FSETP+FSEL/FADD pattern
            if (OP == 5) k[i] = min(max(k[i], 64), 940 + it);                             // int clamp
            if (OP == 6) k[i] = (k[i] >> 3) + (int)((unsigned)k[i] >> 31) + it;           // SHF + SHF + IADD3
            if (OP == 7) k[i] = __byte_perm(k[i], it, 0x5410) ^ (k[i] >> 16);             // PRMT + SHF + LOP
            if (OP == 8) p[i] = fmul2(p[i], pc1);                                         // FMUL2
            if (OP == 9) f[i] = fminf(fmaxf(f[i], c1), c2 + (float)it);                   // FMNMX x2 (+ I2F hoisted?)
            if (OP == 10) { f[i] = (float)k[i]; k[i] += __float_as_int(f[i]) & 3; }        // I2F.S32 + LOP + IADD
            if (OP == 11) { k[i] = __float2int_rz(f[i]); f[i] = __int_as_float(__float_as_int(f[i]) + (k[i] & 1)); }  // F2I + LOP + IADD
            if (OP == 12) p[i] = fadd2(p[i], pc1);                                        // FADD2
        }
    }
    long long c1c = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc += f[i] + (float)k[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[t] = acc;
    if (t == 0) *cycles = c1c - c0;
}

// clean LDS gather: index = 15-bit LCG state (IMAD + LOP per gather), table 32768 floats (128 KB)
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) probe_lds(float *out, long long *cycles, int seed)
{
    extern __shared__ float lut[];
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) lut[i] = (float)(i & 7);
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned a[ILP];
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = (t * 2654435761u + i * 40503u + seed) & 32767u; acc[i] = 0; }
    long long c0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            acc[i] += lut[a[i]];
            a[i] = (a[i] * 20077u + 12345u) & 32767u;
        }
    }
    long long c1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i] + (float)a[i];
    out[t] = s;
    if (t == 0) *cycles = c1 - c0;
}

int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs\n", pr.name, sms);
    const int threads = 256, bps = 4, blocks = sms * bps, n = 1 << 16;
    float *in, *out;
    long long *cyc, hc;
    cudaMalloc(&in, n * 4); cudaMalloc(&out, (size_t)blocks * 1024 * 4); cudaMalloc(&cyc, 8);
    float *hin = (float *)malloc(n * 4);
    for (int i = 0; i < n; i++) hin[i] = 1.0f + (rand() % 1000) * 0.37f;
    hin[1] = 1.0000001f; hin[2] = 0.5f;
    cudaMemcpy(in, hin, n * 4, cudaMemcpyHostToDevice);
    const char *names[] = {"FFMA", "FFMA2 (lane-ops)", "FADD2.RM (lane-ops)", "FADD.RM", "FSETP+pred FADD", "int clamp min(max())",
                           "SHF+SHF+IADD3 (x3)", "PRMT+SHF+LOP (x3)", "FMUL2 (lane-ops)", "float clamp (x2)", "I2F+LOP+IADD",
                           "F2I+LOP+IADD", "FADD2 (lane-ops)"};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(OP)                                                                                     \
    {                                                                                               \
        probe<OP><<<blocks, threads>>>(out, in, n, cyc);                                            \
        cudaEventRecord(e0);                                                                        \
        probe<OP><<<blocks, threads>>>(out, in, n, cyc);                                            \
        cudaEventRecord(e1);                                                                        \
        cudaEventSynchronize(e1);                                                                   \
        float ms;                                                                                   \
        cudaEventElapsedTime(&ms, e0, e1);                                                          \
        cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);                                            \
        double ops = (double)blocks * threads * ITERS * ILP;                                        \
        printf("%-24s %8.3f ms  %9.1f Gop/s  %7.1f thread-ops/clk/SM\n", names[OP], ms, ops / ms * 1e-6, \
               (double)bps * threads * ITERS * ILP / (double)hc);                                   \
    }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
    cudaFuncSetAttribute(probe_lds<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    cudaFuncSetAttribute(probe_lds<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    for (int w : {16, 32}) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (w == 16) probe_lds<16><<<sms, 512, 131072>>>(out, cyc, rep); else probe_lds<32><<<sms, 1024, 131072>>>(out, cyc, rep);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        double ops = (double)sms * w * 32 * ITERS * ILP;
        printf("LDS random gather, %2d warps/SM: %8.3f ms  %8.1f Ggather/s  %6.2f gathers/clk/SM\n", w, ms, ops / ms * 1e-6,
               (double)w * 32 * ITERS * ILP / (double)hc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
