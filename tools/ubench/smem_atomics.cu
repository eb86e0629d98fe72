// Microbenchmark: shared-memory atomic throughput on B200 (lane-ops per cycle per SM), the quantity that bounds an
// open-addressing count table in shared memory.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef uint32_t u32;

__device__ __forceinline__ u32 rng(u32 &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

// MODE 0: CAS.64 on random distinct-ish slots   1: atomicAdd u32 no return   2: atomicAdd u32 with return
// MODE 3: CAS.64 + atomicAdd (one insert)       4: plain LDS.64 + STS.64     5: CAS.32   6: atomicAdd u64 no return
// MODE 7: match_any on 64-bit keys (no smem)
template <int MODE, int SLOTS>
__global__ void k_bench(u64 *out, int iters)
{
    extern __shared__ __align__(16) unsigned char raw[];
    u64 *key = reinterpret_cast<u64*>(raw);
    u32 *cnt = reinterpret_cast<u32*>(key + SLOTS);
    for (int i = threadIdx.x; i < SLOTS; i += blockDim.x) { key[i] = ~0ull; cnt[i] = 0; }
    __syncthreads();
    u32 s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    u64 acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
        u32 r0 = rng(s), r1 = rng(s);
        u32 a = r0 & (SLOTS - 1), b = r1 & (SLOTS - 1);
        u64 k0 = ((u64)r0 << 32) | r1, k1 = ((u64)r1 << 32) | r0;
        if (MODE == 0) { acc += atomicCAS(&key[a], ~0ull, k0); acc += atomicCAS(&key[b], ~0ull, k1); }
        if (MODE == 1) { atomicAdd(&cnt[a], 1u); atomicAdd(&cnt[b], 1u); }
        if (MODE == 2) { acc += atomicAdd(&cnt[a], 1u); acc += atomicAdd(&cnt[b], 1u); }
        if (MODE == 3) { acc += atomicCAS(&key[a], ~0ull, k0); acc += atomicCAS(&key[b], ~0ull, k1); atomicAdd(&cnt[a], 1u); atomicAdd(&cnt[b], 1u); }
        if (MODE == 4) { acc += key[a]; acc += key[b]; key[a ^ 1] = k0; key[b ^ 1] = k1; }
        if (MODE == 5) { acc += atomicCAS(&cnt[a], 0u, r0); acc += atomicCAS(&cnt[b], 0u, r1); }
        if (MODE == 6) { atomicAdd(&key[a], 1ull); atomicAdd(&key[b], 1ull); }
        if (MODE == 7) { acc += __match_any_sync(0xffffffffu, k0 & 0xFFull); acc += __match_any_sync(0xffffffffu, k1 & 0xFFull); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x * 2] = (u64)(t1 - t0);
    if (acc == 0x123456789ull) out[blockIdx.x * 2 + 1] = acc;
}

template <int MODE, int SLOTS>
void run(const char *name, int threads, int ctas_per_sm, u64 *d_out, int sms)
{
    const int iters = 4096;
    size_t smem = (size_t)SLOTS * 12;
    cudaFuncSetAttribute(k_bench<MODE, SLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = sms * ctas_per_sm;
    k_bench<MODE, SLOTS><<<grid, threads, smem>>>(d_out, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_bench<MODE, SLOTS><<<grid, threads, smem>>>(d_out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    u64 h[2]; cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    int nat = (MODE == 3) ? 4 : 2;
    double laneops = (double)iters * nat * threads * ctas_per_sm;          // per SM
    printf("%-28s slots %5d thr %4d x%d/SM : %8.0f cyc  %6.3f lane-ops/cyc/SM  (%.3f ms, %s)\n", name, SLOTS, threads, ctas_per_sm,
           (double)h[0], laneops / (double)h[0], ms, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    u64 *d; cudaMalloc(&d, 16 * 4096);
    for (int thr : {256, 512, 1024})
    {
        int c = 1024 / thr;
        run<0, 2048>("CAS.64", thr, c, d, sms);
        run<0, 8192>("CAS.64", thr, c, d, sms);
        run<1, 2048>("ADD.32 (no return)", thr, c, d, sms);
        run<2, 2048>("ADD.32 (return)", thr, c, d, sms);
        run<3, 2048>("CAS.64 + ADD.32 (insert)", thr, c, d, sms);
        run<4, 2048>("LDS.64 + STS.64", thr, c, d, sms);
        run<5, 2048>("CAS.32", thr, c, d, sms);
        run<6, 2048>("ADD.64 (no return)", thr, c, d, sms);
        run<7, 2048>("match_any (8-bit keys)", thr, c, d, sms);
    }
    run<3, 2048>("CAS.64 + ADD.32 (insert)", 256, 8, d, sms);
    run<0, 2048>("CAS.64", 256, 8, d, sms);
    run<0, 2048>("CAS.64", 128, 1, d, sms);
    run<0, 2048>("CAS.64", 32, 1, d, sms);
    return 0;
}
