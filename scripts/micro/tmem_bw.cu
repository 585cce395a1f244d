// tmem_bw.cu — micro-benchmark: tensor-memory READ bandwidth per SM (tcgen05.ld.32x32b.x32), the roofline of the
// fused-MLP epilogues (every fp32 accumulator element has to cross TMEM -> registers once per layer).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
// Prints bytes/clk/SM for 4, 8 and 16 reading warps per CTA (1 CTA per SM, all SMs busy), with 1 and 4 loads in flight
// per warp between waits.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t r[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}

template <int kInFlight>
__global__ void tmem_read_kernel(uint32_t* sink, long long* cycles, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[kInFlight][32];
#pragma unroll
        for (int q = 0; q < kInFlight; ++q) ld32(base + (uint32_t)(((i * kInFlight + q + (warp >> 2) * 3) & 15) * 32), r[q]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < kInFlight; ++q)
#pragma unroll
            for (int c = 0; c < 32; ++c) acc ^= r[q][c];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int kInFlight>
void run(int warps, int sms) {
    uint32_t* sink;
    long long* cyc;
    cudaMalloc(&sink, 4);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    const int iters = 4096;
    tmem_read_kernel<kInFlight><<<sms, warps * 32>>>(sink, cyc, 64);
    tmem_read_kernel<kInFlight><<<sms, warps * 32>>>(sink, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("error: %s\n", cudaGetErrorString(e));
        return;
    }
    long long* h = new long long[sms];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)h[i];
    mean /= sms;
    const double bytes = (double)iters * kInFlight * warps * 4096.0;
    printf("warps/CTA %2d  loads in flight %d : %7.1f bytes/clk/SM  (%.0f cycles)\n", warps, kInFlight, bytes / mean, mean);
    delete[] h;
    cudaFree(sink);
    cudaFree(cyc);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("tcgen05.ld.32x32b.x32 read bandwidth, %d SMs, 1 CTA/SM\n", sms);
    for (int w : {4, 8, 16}) {
        run<1>(w, sms);
        run<4>(w, sms);
    }
    return 0;
}
