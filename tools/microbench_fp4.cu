// Feasibility probe for a third distance core (development aid, not part of the product):
// tcgen05.mma kind::mxf4 (block-scaled e2m1, K = 64 per instruction) on a +/-1 expansion of the
// descriptor bits into 4-bit floats (bit 1 -> +1.0 = 0x2, bit 0 -> -1.0 = 0xA), every block scale
// = 1.0 (ue8m0 0x7F).  dot = 256 - 2*H is an exact small integer in the fp32 accumulator.
//   1. exactness: one 128 x 128 x 256 tile against a host popcount
//   2. issue rate of the instruction (N = 128 / 256) next to kind::i8
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I slam_experiments_b200/csrc \
//        -I include -o tools/microbench_fp4 tools/microbench_fp4.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "hm_tcgen05.cuh"

using namespace hm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// instruction descriptor, block-scaled kinds: [4,6) B sf id  [7,10) A format (mxf4: 1 = e2m1)  [10,13) B format
// [17,23) N >> 3   bit 23 scale format (1 = ue8m0)   [24,29) M >> 4   [29,31) A sf id   bit 31 K size (0 = K64)
__host__ __device__ constexpr uint32_t make_mxf4_idesc(int m, int n)
{
    return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_mxf4_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}

__device__ __forceinline__ void tmem_st_32x32_x1(uint32_t taddr, uint32_t v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};\n" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

constexpr int kSfCol = 384;   // scale factors: columns [384, 512), every byte 0x7F (= 1.0)

__device__ void fill_scales(uint32_t tmem_base, int warp)
{
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int c = kSfCol; c < 512; ++c) tmem_st_32x32_x1(tmem_base + lane_base + c, 0x7F7F7F7Fu);
    tmem_st_wait();
}

// ---------------------------------------------------------------- exactness
__global__ void __launch_bounds__(128, 1) fp4_tile_kernel(const uint8_t* a_img, const uint8_t* b_img, float* out)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 16384 / 16; i += blockDim.x) {
        reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(a_img)[i];
        reinterpret_cast<uint4*>(smem + 16384)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
    if (threadIdx.x == 32) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = slot;
    fill_scales(tm, warp);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 1) {
        if (ptx::elect_one()) {
            const uint32_t a = ptx::smem_u32(smem), b = a + 16384;
            constexpr uint32_t idesc = make_mxf4_idesc(128, 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                mma_mxf4_ss(tm, ptx::make_kmajor_sw128_desc(a + k * 32), ptx::make_kmajor_sw128_desc(b + k * 32), idesc,
                            tm + kSfCol, tm + kSfCol + 64, k != 0);
            ptx::tc_commit(&bar);
        }
        __syncwarp();
    }
    ptx::mbar_wait(&bar, 0);
    ptx::tc_fence_after();
    const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr + c * 32, r);
        ptx::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(size_t)threadIdx.x * 128 + c * 32 + j] = __uint_as_float(r[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(slot, 512); }
}

// ---------------------------------------------------------------- issue rate
template <int N, bool kFp4>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* cycles)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = kFp4 ? 0x2A2A2A2Au : 0x01FF01FFu;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
    if (threadIdx.x == 32) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tm = slot;
    if (kFp4) fill_scales(tm, warp);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 1) {
        const uint32_t a = ptx::smem_u32(smem), b = a + 32768;
        long long t0 = clock64();
        if (ptx::elect_one()) {
            if (kFp4) {
                constexpr uint32_t idesc = make_mxf4_idesc(128, N);
                // accumulators alternate between columns [0, N) and (N = 128 only) [128, 256)
                for (int it = 0; it < iters; ++it)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_mxf4_ss(tm + (N == 128 ? (it & 1) * 128 : 0), ptx::make_kmajor_sw128_desc(a + k * 32),
                                    ptx::make_kmajor_sw128_desc(b + k * 32), idesc, tm + kSfCol, tm + kSfCol + 64, k != 0);
            } else {
                constexpr uint32_t idesc = ptx::make_i8_idesc(128, N);
                for (int it = 0; it < iters; ++it)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::mma_i8_ss(tm + (N == 128 ? (it & 1) * 128 : 0), ptx::make_kmajor_sw128_desc(a + k * 32),
                                       ptx::make_kmajor_sw128_desc(b + k * 32), idesc, k != 0);
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

// host: packed bits -> e2m1 nibbles in the K-major SWIZZLE_128B image of one 128-row block (16 KB)
static void expand_fp4(const uint8_t* bits, uint8_t* img)
{
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 8; ++c) {               // logical 16-byte chunk = 32 descriptor bits
            uint8_t* dst = img + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) * 16);
            for (int j = 0; j < 16; ++j) {
                const int bit0 = c * 32 + j * 2;
                const int b0 = (bits[r * 32 + bit0 / 8] >> (bit0 & 7)) & 1;
                const int b1 = (bits[r * 32 + (bit0 + 1) / 8] >> ((bit0 + 1) & 7)) & 1;
                dst[j] = (uint8_t)((b0 ? 0x2 : 0xA) | ((b1 ? 0x2 : 0xA) << 4));
            }
        }
}

template <int N, bool kFp4>
static void run_rate(const char* name, int sms, long long* cyc)
{
    const int iters = 8000;
    const int smem_bytes = 100 * 1024 + 2048;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(rate_kernel<N, kFp4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        rate_kernel<N, kFp4><<<sms, 128, smem_bytes>>>(iters, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
    }
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h;
    CK(cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    const double macs = (double)iters * 4 * 128 * N * (kFp4 ? 64 : 32);
    printf(", \"%s_cycles_per_mma\": %.1f, \"%s_mac_per_clk_per_sm\": %.0f, \"%s_chip_tops\": %.1f", name,
           (double)h / (iters * 4), name, macs / (double)h, name, 2.0 * macs * sms / (ms * 1e-3) / 1e12);
}

int main()
{
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    printf("{\"sm_count\": %d, \"clock_khz\": %d", sms, khz);

    // 1. exactness of one tile
    static uint8_t q[128 * 32], t[128 * 32], qi[16384], ti[16384];
    static float out[128 * 128];
    srand(7);
    for (int i = 0; i < 128 * 32; ++i) { q[i] = rand() & 255; t[i] = rand() & 255; }
    memcpy(t, q, 32);                                  // one exact duplicate: distance 0 -> dot 256
    for (int i = 0; i < 32; ++i) t[32 + i] = ~q[32 + i];   // one complement: distance 256 -> dot -256
    expand_fp4(q, qi); expand_fp4(t, ti);
    uint8_t *da, *db; float* dout;
    CK(cudaMalloc(&da, 16384)); CK(cudaMalloc(&db, 16384)); CK(cudaMalloc(&dout, sizeof(out)));
    CK(cudaMemcpy(da, qi, 16384, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, ti, 16384, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(fp4_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
    fp4_tile_kernel<<<1, 128, 40 * 1024>>>(da, db, dout);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, dout, sizeof(out), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 128; ++j) {
            int h = 0;
            for (int b = 0; b < 32; ++b) h += __builtin_popcount(q[i * 32 + b] ^ t[j * 32 + b]);
            if (out[i * 128 + j] != (float)(256 - 2 * h)) {
                if (bad < 5) fprintf(stderr, "mismatch [%d][%d]: got %g want %d\n", i, j, out[i * 128 + j], 256 - 2 * h);
                ++bad;
            }
        }
    printf(", \"fp4_tile_mismatches\": %d, \"fp4_dot_dup\": %g, \"fp4_dot_complement\": %g", bad, out[0], out[128 + 1]);

    // 2. issue rates
    long long* cyc;
    CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    run_rate<128, false>("mma_i8_n128", sms, cyc);
    run_rate<128, true>("mma_mxf4_n128", sms, cyc);
    run_rate<256, true>("mma_mxf4_n256", sms, cyc);
    printf("}\n");
    return 0;
}
