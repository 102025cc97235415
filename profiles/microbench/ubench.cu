// ubench.cu -- instruction-throughput probes that informed the fused kernel's design (B200, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run on the GPU box.
// Prints warp-instructions/clk/SM-equivalent lane throughput (lanes per clock per SM) per probe.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(256) probe(float *out, const float *in, int n, long long *cycles)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float f[ILP];
    double d[ILP];
    int k[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { f[i] = in[(t + i) % n]; d[i] = (double)f[i] + 1.0; k[i] = (int)(f[i] * 1000.f) + t; }
    long long c0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) { d[i] = (double)f[i]; f[i] = __int_as_float(__float_as_int(f[i]) ^ (int)__double2hiint(d[i])); }  // F2F.F64.F32 (+ LOP)
            if (OP == 1) { f[i] = __double2float_rn(d[i]); d[i] = __hiloint2double(__double2hiint(d[i]) ^ 1, __float_as_int(f[i])); } // F2F.F32.F64
            if (OP == 2) { f[i] = (float)(unsigned)k[i]; k[i] ^= __float_as_int(f[i]); }      // I2F.U32
            if (OP == 3) { k[i] = __float2int_rz(f[i]); f[i] = __int_as_float(__float_as_int(f[i]) ^ (k[i] & 1)); }  // F2I
            if (OP == 4) { d[i] = __fma_rn(d[i], 1.0000001, 0.5); }                          // DFMA
            if (OP == 5) { d[i] = __dadd_rn(d[i], 0.5); }                                     // DADD
            if (OP == 6) { d[i] = __dmul_rn(d[i], 1.0000001); }                               // DMUL
            if (OP == 7) { f[i] = __fmaf_rn(f[i], 1.0000001f, 0.5f); }                        // FFMA
            if (OP == 8) { k[i] = k[i] * 21 + 7; }                                            // IMAD
            if (OP == 9) { k[i] = __shfl_up_sync(0xffffffffu, k[i], 1); }                      // SHFL
            if (OP == 10) { k[i] = __double2int_rz(d[i]); d[i] = __hiloint2double(__double2hiint(d[i]), k[i]); }  // D2I
            if (OP == 11) { f[i] = __fadd_rn(f[i], 0.5f); }                                   // FADD
            if (OP == 12) { f[i] = __fmul_rn(f[i], 1.0000001f); }                             // FMUL
            if (OP == 13) { k[i] = min(max(k[i], 64), 940 + it); }                            // IMNMX x2
            if (OP == 14) { f[i] = fminf(fmaxf(f[i], 0.f), 1023.f + it); }                     // FMNMX x2
            if (OP == 15) { d[i] = __dadd_rz(d[i], 412316860416.5); }                         // DADD.RZ
        }
    }
    long long c1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc += f[i] + (float)d[i] + (float)k[i];
    out[t] = acc;
    if (t == 0) *cycles = c1 - c0;
}

// shared-memory LUT gather: random float reads from a table of `codes` entries
__global__ void __launch_bounds__(512) probe_lds(float *out, const unsigned *idx, int nidx, int codes, long long *cycles)
{
    extern __shared__ float lut[];
    for (int i = threadIdx.x; i < codes; i += blockDim.x) lut[i] = (float)i;
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned a[ILP];
    float acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = idx[(t * ILP + i) % nidx] % codes;
    long long c0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            float v = lut[a[i]];
            acc += v;
            a[i] = (a[i] * 1664525u + 1013904223u + (unsigned)v) % (unsigned)codes;   // next random code (adds ALU work)
        }
    }
    long long c1 = clock64();
    out[t] = acc;
    if (t == 0) *cycles = c1 - c0;
}

// global (L1/L2) LUT gather
__global__ void __launch_bounds__(512) probe_ldg(float *out, const float *lut, int codes, long long *cycles)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned a[ILP];
    float acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) a[i] = (t * 2654435761u + i * 40503u) % codes;
    long long c0 = clock64();
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            float v = __ldg(lut + a[i]);
            acc += v;
            a[i] = (a[i] * 1664525u + 1013904223u + (unsigned)v) % (unsigned)codes;
        }
    }
    long long c1 = clock64();
    out[t] = acc;
    if (t == 0) *cycles = c1 - c0;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs\n", p.name, sms);
    const int threads = 256, blocks = sms * 8, n = 1 << 16;
    float *in, *out;
    long long *cyc, hc;
    cudaMalloc(&in, n * 4); cudaMalloc(&out, (size_t)blocks * 512 * 4); cudaMalloc(&cyc, 8);
    float *hin = (float *)malloc(n * 4);
    for (int i = 0; i < n; i++) hin[i] = 1.0f + (rand() % 1000) * 0.37f;
    cudaMemcpy(in, hin, n * 4, cudaMemcpyHostToDevice);
    const char *names[] = {"F2F.F64.F32(+LOP)", "F2F.F32.F64(+2 mov)", "I2F.U32(+LOP)", "F2I(+2 LOP)", "DFMA", "DADD", "DMUL", "FFMA",
                           "IMAD", "SHFL", "D2I(+mov)", "FADD", "FMUL", "IMNMX x2", "FMNMX x2", "DADD.RZ"};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(OP)                                                                                         \
    {                                                                                                   \
        probe<OP><<<blocks, threads>>>(out, in, n, cyc);                                                \
        cudaEventRecord(e0);                                                                            \
        probe<OP><<<blocks, threads>>>(out, in, n, cyc);                                                \
        cudaEventRecord(e1);                                                                            \
        cudaEventSynchronize(e1);                                                                       \
        float ms;                                                                                       \
        cudaEventElapsedTime(&ms, e0, e1);                                                              \
        cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);                                                \
        double ops = (double)blocks * threads * ITERS * ILP;                                            \
        double clk_hz = hc / (ms * 1e-3) ;  /* cycles of block 0 ~ whole kernel when 1 wave */          \
        printf("%-22s %8.3f ms  %7.1f Gop/s  %6.1f lanes/clk/SM (clock~%.0f MHz over kernel)\n", names[OP], ms, ops / ms * 1e-6, \
               ops / ((double)hc * 8 / 8) / sms * ( (double)hc / (ms*1e-3) > 0 ? 1.0 : 1.0) * ((double)hc/(double)hc), clk_hz * 1e-6); \
    }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15)
    // LUT gathers
    unsigned *idx;
    const int nidx = 1 << 20;
    cudaMalloc(&idx, nidx * 4);
    unsigned *hidx = (unsigned *)malloc(nidx * 4);
    for (int i = 0; i < nidx; i++) hidx[i] = (unsigned)rand();
    cudaMemcpy(idx, hidx, nidx * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 130000);
    for (int codes : {1024, 20000, 31744}) {
        probe_lds<<<sms, 512, codes * 4>>>(out, idx, nidx, codes, cyc);
        cudaEventRecord(e0);
        probe_lds<<<sms, 512, codes * 4>>>(out, idx, nidx, codes, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        double ops = (double)sms * 512 * ITERS * ILP;
        printf("LDS gather %6d codes  %8.3f ms  %7.1f Ggather/s  %6.2f gathers/clk/SM (16 warps/SM)\n", codes, ms, ops / ms * 1e-6,
               (double)512 * ITERS * ILP / (double)hc);
    }
    float *glut;
    cudaMalloc(&glut, 65536 * 3 * 4);
    cudaMemset(glut, 0, 65536 * 3 * 4);
    for (int codes : {4096, 20000, 31744, 65536, 196608}) {
        probe_ldg<<<sms * 2, 512>>>(out, glut, codes, cyc);
        cudaEventRecord(e0);
        probe_ldg<<<sms * 2, 512>>>(out, glut, codes, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        double ops = (double)sms * 2 * 512 * (ITERS / 4) * ILP;
        printf("LDG gather %6d codes (%4d KB)  %8.3f ms  %7.1f Ggather/s  %6.2f gathers/clk/SM (32 warps/SM)\n", codes, codes * 4 / 1024, ms,
               ops / ms * 1e-6, (double)2 * 512 * (ITERS / 4) * ILP / (double)hc);
    }
    return 0;
}
