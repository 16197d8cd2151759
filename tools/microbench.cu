// Device micro-benchmarks that calibrate the roofline denominators on the box (development aid,
// not part of the product): POPC issue rate, tcgen05.ld (TMEM read) bandwidth, and the issue rate
// of tcgen05.mma kind::i8 with operands already in shared memory.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I slam_experiments_b200/csrc \
//        -I include -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "hm_tcgen05.cuh"

using namespace hm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// ---------------------------------------------------------------- POPC
template <bool kXor>
__global__ void popc_kernel(unsigned* out, int iters, unsigned seed)
{
    unsigned a[8], acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            unsigned v = kXor ? (a[i] ^ (unsigned)it) : a[i];
            acc += __popc(v);
            if (!kXor) a[i] += acc;   // keep the popc inputs changing without an extra ALU op per popc
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---------------------------------------------------------------- TMEM read
__global__ void __launch_bounds__(256, 1) ldtm_kernel(unsigned* out, int iters, long long* cycles)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t r[32];
    unsigned acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            ptx::tmem_ld_32x32(base + ((c * 32 + (warp >> 2) * 256) & 511), r);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" : "+r"(r[0]), "+r"(r[31]) :: "memory");
            acc += r[0] ^ r[31];
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(slot, 512); }
}

// ---------------------------------------------------------------- MMA issue rate
template <int N>
__global__ void __launch_bounds__(128, 1) mma_kernel(int iters, long long* cycles)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01FF01FFu;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
    if (threadIdx.x == 32) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 1) {
        const uint32_t a = ptx::smem_u32(smem), b = a + 32768;
        constexpr uint32_t idesc = ptx::make_i8_idesc(128, N);
        long long t0 = clock64();
        if (ptx::elect_one()) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::mma_i8_ss(slot + (it & 1) * 256, ptx::make_kmajor_sw128_desc(a + s * 16384 + k * 32),
                                       ptx::make_kmajor_sw128_desc(b + s * 32768 + k * 32), idesc, (s | k) != 0);
            }
            ptx::tc_commit(&bar);
        }
        __syncwarp();
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 32) cycles[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(slot, 512); }
}

int main()
{
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    printf("{\"sm_count\": %d, \"clock_khz\": %d", sms, khz);
    unsigned* out;
    long long* cyc;
    CK(cudaMalloc(&out, sizeof(unsigned) * sms * 8 * 1024));
    CK(cudaMalloc(&cyc, sizeof(long long) * sms * 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    // POPC: 8 CTAs of 256 threads per SM
    for (int x = 0; x < 2; ++x) {
        const int iters = 20000;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            if (x) popc_kernel<true><<<sms * 8, 256>>>(out, iters, 12345u);
            else   popc_kernel<false><<<sms * 8, 256>>>(out, iters, 12345u);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
        }
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double popc = (double)sms * 8 * 256 * iters * 8;
        printf(", \"popc%s_per_s\": %.4g, \"popc%s_per_clk_per_sm_at_max_clock\": %.3f", x ? "_xor" : "", popc / (ms * 1e-3),
               x ? "_xor" : "", popc / (ms * 1e-3) / sms / (khz * 1e3));
    }
    // LDTM: 4 or 8 warps per SM
    for (int warps = 4; warps <= 8; warps += 4) {
        const int iters = 4000;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            ldtm_kernel<<<sms, warps * 32>>>(out, iters, cyc);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
        }
        CK(cudaEventElapsedTime(&ms, e0, e1));
        long long h;
        CK(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double bytes = (double)warps * iters * 8 * 4096;
        printf(", \"ldtm_%dwarps_bytes_per_clk_per_sm\": %.2f", warps, bytes / (double)h);
    }
    // MMA issue loop
    {
        const int iters = 4000;
        CK(cudaFuncSetAttribute(mma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 2048));
        CK(cudaFuncSetAttribute(mma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 2048));
        for (int n = 256; n >= 128; n -= 128) {
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                if (n == 256) mma_kernel<256><<<sms, 128, 100 * 1024 + 2048>>>(iters, cyc);
                else          mma_kernel<128><<<sms, 128, 100 * 1024 + 2048>>>(iters, cyc);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
            }
            CK(cudaEventElapsedTime(&ms, e0, e1));
            long long h;
            CK(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
            double macs = (double)iters * 8 * 128 * n * 32;
            printf(", \"mma_i8_n%d_cycles_per_mma\": %.1f, \"mma_i8_n%d_mac_per_clk_per_sm\": %.0f, \"mma_i8_n%d_chip_tops\": %.1f",
                   n, (double)h / (iters * 8), n, macs / (double)h, n, 2.0 * macs * sms / (ms * 1e-3) / 1e12);
        }
    }
    printf("}\n");
    return 0;
}
