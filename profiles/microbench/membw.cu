// membw.cu -- read-bandwidth probes for the statistics pass (B200): how many bytes must be in flight per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membw membw.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldg_na(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <int UNROLL, bool NA>
__global__ void __launch_bounds__(256) k_read(const uint4 *__restrict__ p, long n, unsigned *out)
{
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long)gridDim.x * blockDim.x;
    unsigned acc = 0;
    long i = tid;
    for (; i + (UNROLL - 1) * nthr < n; i += UNROLL * nthr) {
        uint4 v[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) v[j] = NA ? ldg_na(p + i + j * nthr) : __ldg(p + i + j * nthr);
#pragma unroll
        for (int j = 0; j < UNROLL; j++) acc = max(acc, max(max(v[j].x, v[j].y), max(v[j].z, v[j].w)));
    }
    for (; i < n; i += nthr) { uint4 v = __ldg(p + i); acc = max(acc, max(max(v.x, v.y), max(v.z, v.w))); }
    if (acc == 0xdeadbeef) out[0] = acc;
}

// TMA bulk (cp.async.bulk global->shared, mbarrier completion): one thread issues, everyone reduces from smem
template <int STAGES, int CHUNK>
__global__ void __launch_bounds__(256) k_read_tma(const uint8_t *__restrict__ p, long nbytes, unsigned *out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar[STAGES];
    const long nchunks = nbytes / CHUNK;
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES; s++) {
            unsigned a = (unsigned)__cvta_generic_to_shared(&bar[s]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
        }
    __syncthreads();
    unsigned acc = 0;
    const long my = blockIdx.x;
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES; s++) {
            long c = my + (long)s * gridDim.x;
            if (c < nchunks) {
                unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]), d = (unsigned)__cvta_generic_to_shared(smem + s * CHUNK);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CHUNK));
                asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(p + c * CHUNK), "r"(CHUNK), "r"(b) : "memory");
            }
        }
    int it = 0;
    for (long c = my; c < nchunks; c += gridDim.x, it++) {
        const int s = it % STAGES;
        const unsigned parity = (it / STAGES) & 1;
        unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b), "r"(parity) : "memory");
        const uint4 *sp = reinterpret_cast<const uint4 *>(smem + s * CHUNK);
#pragma unroll 4
        for (int i = threadIdx.x; i < CHUNK / 16; i += 256) { uint4 v = sp[i]; acc = max(acc, max(max(v.x, v.y), max(v.z, v.w))); }
        __syncthreads();
        if (threadIdx.x == 0) {
            long nc = c + (long)STAGES * gridDim.x;
            if (nc < nchunks) {
                unsigned d = (unsigned)__cvta_generic_to_shared(smem + s * CHUNK);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(CHUNK));
                asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(p + nc * CHUNK), "r"(CHUNK), "r"(b) : "memory");
            }
        }
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

int main()
{
    const long nbytes = 3L << 30;
    uint8_t *buf; unsigned *out;
    cudaMalloc(&buf, nbytes); cudaMalloc(&out, 64);
    cudaMemset(buf, 1, nbytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long n = nbytes / 16;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
#define T(name, launch) { launch; cudaEventRecord(e0); launch; launch; cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); \
    printf("%-44s %8.1f GB/s  (%s)\n", name, 2.0 * nbytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError())); }
    T("ldg      unroll1  8 CTA/SM", (k_read<1, false><<<sms * 8, 256>>>((const uint4 *)buf, n, out)));
    T("ldg      unroll4  8 CTA/SM", (k_read<4, false><<<sms * 8, 256>>>((const uint4 *)buf, n, out)));
    T("ldg      unroll8  8 CTA/SM", (k_read<8, false><<<sms * 8, 256>>>((const uint4 *)buf, n, out)));
    T("ldg.na   unroll4  8 CTA/SM", (k_read<4, true><<<sms * 8, 256>>>((const uint4 *)buf, n, out)));
    T("ldg.na   unroll8  8 CTA/SM", (k_read<8, true><<<sms * 8, 256>>>((const uint4 *)buf, n, out)));
    T("ldg.na   unroll8  4 CTA/SM", (k_read<8, true><<<sms * 4, 256>>>((const uint4 *)buf, n, out)));
    T("ldg.na   unroll8 16 CTA/SM (2 waves)", (k_read<8, true><<<sms * 16, 256>>>((const uint4 *)buf, n, out)));
    T("ldg      unroll4 64 CTA/SM (8 waves)", (k_read<4, false><<<sms * 64, 256>>>((const uint4 *)buf, n, out)));
    cudaFuncSetAttribute(k_read_tma<4, 16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384);
    cudaFuncSetAttribute(k_read_tma<6, 32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768);
    cudaFuncSetAttribute(k_read_tma<3, 16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16384);
    T("TMA bulk 4 x 16 KB, 1 CTA/SM", (k_read_tma<4, 16384><<<sms, 256, 4 * 16384>>>(buf, nbytes, out)));
    T("TMA bulk 6 x 32 KB, 1 CTA/SM", (k_read_tma<6, 32768><<<sms, 256, 6 * 32768>>>(buf, nbytes, out)));
    T("TMA bulk 3 x 16 KB, 4 CTA/SM", (k_read_tma<3, 16384><<<sms * 4, 256, 3 * 16384>>>(buf, nbytes, out)));
    return 0;
}
