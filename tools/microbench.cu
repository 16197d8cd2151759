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

// ---------------------------------------------------------------- MMA issue patterns
// issuers: 1 or 2 warps, each issuing groups of 8 MMAs (N=128) into its own accumulators;
// commit_every: 0 = one commit at the end, 8 = a tcgen05.commit (to a scratch mbarrier) after every group
__global__ void __launch_bounds__(128, 1) mma_pattern_kernel(int iters, int issuers, int commit_every, long long* cycles)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar[2], scratch[2];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01FF01FFu;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
    if (threadIdx.x == 32) {
        ptx::mbar_init(&bar[0], 1); ptx::mbar_init(&bar[1], 1);
        ptx::mbar_init(&scratch[0], 1 << 20); ptx::mbar_init(&scratch[1], 1 << 20);
        ptx::fence_barrier_init();
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 1 || (warp == 2 && issuers == 2)) {
        const int m = warp - 1;
        const uint32_t a = ptx::smem_u32(smem) + m * 32768, b = ptx::smem_u32(smem) + 65536;
        constexpr uint32_t idesc = ptx::make_i8_idesc(128, 128);
        long long t0 = clock64();
        if (ptx::elect_one()) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::mma_i8_ss(slot + ((it & 1) * 2 + m) * 128, ptx::make_kmajor_sw128_desc(a + s * 16384 + k * 32),
                                       ptx::make_kmajor_sw128_desc(b + s * 16384 + k * 32), idesc, (s | k) != 0);
                if (commit_every) { ptx::tc_commit(&scratch[m]); ptx::tc_commit(&scratch[m]); }
            }
            ptx::tc_commit(&bar[m]);
        }
        __syncwarp();
        ptx::mbar_wait(&bar[m], 0);
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 2 + m] = t1 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(slot, 512); }
}

// ---------------------------------------------------------------- MMA with concurrent shared-memory writes / TMEM reads
// mode bit 0: a second warp streams bulk copies (UBLKCP) into a separate smem region while the MMAs run
// mode bit 1: four more warps loop on tcgen05.ld of the other accumulator half
__global__ void __launch_bounds__(256, 1) mma_interference_kernel(int iters, int mode, const uint8_t* gsrc, long long* cycles,
                                                                  unsigned* sink, long long* copied)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar, cbar[2];
    __shared__ volatile int done;
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < (32768 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01FF01FFu;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
    if (threadIdx.x == 32) { ptx::mbar_init(&bar, 1); ptx::mbar_init(&cbar[0], 1); ptx::mbar_init(&cbar[1], 1); ptx::fence_barrier_init(); done = 0; }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 1) {
        const uint32_t a = ptx::smem_u32(smem), b = a + 32768;
        constexpr uint32_t idesc = ptx::make_i8_idesc(128, 128);
        long long t0 = clock64();
        if (ptx::elect_one()) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::mma_i8_ss(slot + (it & 1) * 128, ptx::make_kmajor_sw128_desc(a + s * 16384 + k * 32),
                                       ptx::make_kmajor_sw128_desc(b + s * 16384 + k * 32), idesc, (s | k) != 0);
            }
            ptx::tc_commit(&bar);
        }
        __syncwarp();
        ptx::mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 32) { cycles[blockIdx.x] = t1 - t0; done = 1; }
    } else if (warp == 2 && (mode & 1)) {
        if (threadIdx.x == 64) {
            uint8_t* dst = smem + 65536;
            long long n = 0;
            const uint8_t* src = gsrc + (size_t)blockIdx.x * (1 << 20);
            for (int it = 0; !done; ++it) {
                const int sgn = it & 1;
                ptx::mbar_arrive_expect_tx(&cbar[sgn], 32768);
                ptx::bulk_g2s(dst + sgn * 32768, src + (size_t)(it & 31) * 32768, 32768, &cbar[sgn]);
                if (it > 0) ptx::mbar_wait(&cbar[sgn ^ 1], ((it - 1) >> 1) & 1);
                n += 32768;
            }
            copied[blockIdx.x] = n;
        }
    } else if (warp >= 4 && (mode & 2)) {
        uint32_t r[32];
        unsigned acc = 0;
        const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + 256;
        while (!done) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                ptx::tmem_ld_32x32(base + c * 32, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;\n" : "+r"(r[0]), "+r"(r[31]) :: "memory");
                acc += r[0] ^ r[31];
            }
        }
        sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
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
    // issue patterns
    {
        const int iters = 4000;
        const int smem_bytes = 65536 + 32768 + 2048;
        CK(cudaFuncSetAttribute(mma_pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        for (int issuers = 1; issuers <= 2; ++issuers)
            for (int ce = 0; ce <= 8; ce += 8) {
                for (int rep = 0; rep < 2; ++rep) {
                    mma_pattern_kernel<<<sms, 128, smem_bytes>>>(iters, issuers, ce, cyc);
                    CK(cudaDeviceSynchronize());
                }
                long long h[2];
                CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
                printf(", \"mma_pattern_issuers%d_commit%d_cycles_per_mma\": %.1f", issuers, ce,
                       (double)h[0] / (iters * 8 * issuers));
            }
    }
    // MMA N=128 with concurrent smem writes / TMEM reads
    {
        const int iters = 4000;
        const int smem_bytes = 65536 + 65536 + 2048;
        uint8_t* gsrc;
        long long* copied;
        CK(cudaMalloc(&gsrc, (size_t)sms << 20));
        CK(cudaMemset(gsrc, 1, (size_t)sms << 20));
        CK(cudaMalloc(&copied, sizeof(long long) * sms));
        CK(cudaFuncSetAttribute(mma_interference_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        for (int mode = 0; mode < 4; ++mode) {
            CK(cudaMemset(copied, 0, sizeof(long long) * sms));
            for (int rep = 0; rep < 2; ++rep) {
                mma_interference_kernel<<<sms, 256, smem_bytes>>>(iters, mode, gsrc, cyc, out, copied);
                CK(cudaDeviceSynchronize());
            }
            long long h, n;
            CK(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&n, copied, sizeof(n), cudaMemcpyDeviceToHost));
            printf(", \"mma_n128_mode%d_cycles_per_mma\": %.1f, \"mma_n128_mode%d_copy_bytes_per_clk\": %.1f", mode,
                   (double)h / (iters * 8), mode, (double)n / (double)h);
        }
    }
    printf("}\n");
    return 0;
}
