// Epilogues over packed top-2 keys: shard/split merge, Lowe ratio test, mutual
// cross-check, the reference's distance filter, ordered stream compaction.
//
//  - merge: top-k of a union = top-k of the per-set top-k's;
//    unsigned min over (dist << 32 | idx) is cv2's tie order.
//  - ratio: integer LUT form of `m.distance < ratio * n.distance` (float64 on the
//    host, see oracle/hamming_oracle.py:ratio_lut) so bit-exactness never depends
//    on GPU floating point.
//  - mutual: cv2.BFMatcher(crossCheck=True).match == strict mutual NN with
//    lowest-index ties both ways; the reverse pass is a k-NN with
//    roles swapped.
//  - distance filter: /root/reference/feature_matchers.py:41-43,
//    keep d < max(2 * min_d, dist_threshold) (strict).
#include "hm_common.cuh"

namespace hm {

namespace {

__global__ void __launch_bounds__(256) hm_merge_top2_kernel(const unsigned long long* __restrict__ keys,
                                                            int groups, long long rows,
                                                            unsigned long long* __restrict__ out)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    unsigned long long k1 = kNoMatch, k2 = kNoMatch;
    fold_partials(keys, groups, rows * 2, r, k1, k2);
    *reinterpret_cast<ulonglong2*>(out + r * 2) = make_ulonglong2(k1, k2);
}

// ---- fused cross-GPU exchange + merge ------------------------------------------------------------
// Symmetric buffer layout (identical on every rank):
//   [2 parities][world slots][max_rows][2] u64 keys | [world][max_ctas] u32 epoch flags
struct ExchangeParams {
    const unsigned long long* local;   // [local_groups][rows][2]
    int local_groups;
    unsigned long long* out;
    long long rows, max_rows;
    int world, rank;
    unsigned epoch;
    unsigned char* peer[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__host__ __device__ inline long long exchange_max_ctas(long long max_rows) { return (max_rows + kExchangeThreads - 1) / kExchangeThreads; }
__host__ __device__ inline size_t exchange_keys_bytes(long long max_rows, int world)
{
    return (size_t)2 * world * max_rows * 2 * sizeof(unsigned long long);
}

__global__ void __launch_bounds__(kExchangeThreads) hm_exchange_merge_kernel(const ExchangeParams P)
{
    const long long r = (long long)blockIdx.x * kExchangeThreads + threadIdx.x;
    const int parity = P.epoch & 1;
    const size_t slot_keys = (size_t)P.max_rows * 2;
    const size_t keys_bytes = exchange_keys_bytes(P.max_rows, P.world);
    const long long max_ctas = exchange_max_ctas(P.max_rows);

    // 1. push this rank's candidates into slot `rank` of every rank's buffer (peer stores over NVLink)
    ulonglong2 mine = make_ulonglong2(kNoMatch, kNoMatch);
    if (r < P.rows) {
        fold_partials(P.local, P.local_groups, P.rows * 2, r, mine.x, mine.y);   // fold the k-NN kernel's train splits
        for (int p = 0; p < P.world; ++p) {
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(P.peer[p]) +
                                      ((size_t)parity * P.world + P.rank) * slot_keys + r * 2;
            *reinterpret_cast<ulonglong2*>(dst) = mine;
        }
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: flag[rank][cta] = epoch on every peer; 3. wait for every peer's flag for this CTA
    if (threadIdx.x < P.world) {
        const int p = threadIdx.x;
        unsigned* flag = reinterpret_cast<unsigned*>(P.peer[p] + keys_bytes) + (size_t)P.rank * max_ctas + blockIdx.x;
        st_release_sys(flag, P.epoch);
        const unsigned* want = reinterpret_cast<const unsigned*>(P.peer[P.rank] + keys_bytes) + (size_t)p * max_ctas + blockIdx.x;
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(want) - P.epoch) < 0) {
            if (++spins > (1u << 27)) __trap();          // a missing peer must not hang the GPU
        }
    }
    __syncthreads();
    // 4. merge the `world` slots of this rank's own buffer
    if (r < P.rows) {
        unsigned long long k1 = kNoMatch, k2 = kNoMatch;
        const unsigned long long* base = reinterpret_cast<const unsigned long long*>(P.peer[P.rank]) +
                                         (size_t)parity * P.world * slot_keys + r * 2;
        for (int g = 0; g < P.world; ++g) {
            ulonglong2 k = mine;
            if (g != P.rank) {   // written by a peer GPU: bypass L1
                asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];\n"
                             : "=l"(k.x), "=l"(k.y)
                             : "l"(base + (size_t)g * slot_keys)
                             : "memory");
            }
            top2_insert(k1, k2, k.x);
            top2_insert(k1, k2, k.y);
        }
        *reinterpret_cast<ulonglong2*>(P.out + r * 2) = make_ulonglong2(k1, k2);
    }
}

constexpr int kFilterThreads = 1024;

struct FilterParams {
    const unsigned long long* fwd;
    const unsigned long long* bwd;
    long long nq, nt;
    unsigned flags;
    int thr_ceil;                    // ceil(dist_threshold), exact integer form of the float compare
    int* out_q;
    int* out_t;
    int* out_d;
    int* out_count;
    RatioLut lut;
};

// One CTA per problem: rows are visited in order so the output stays sorted by queryIdx.
__global__ void __launch_bounds__(kFilterThreads) hm_filter_kernel(const FilterParams P)
{
    __shared__ int warp_sums[32];
    __shared__ int s_base;
    __shared__ unsigned s_min;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.x;
    const unsigned long long* fwd = P.fwd + (long long)b * P.nq * 2;
    const unsigned long long* bwd = P.bwd ? P.bwd + (long long)b * P.nt * 2 : nullptr;
    int* oq = P.out_q + (long long)b * P.nq;
    int* ot = P.out_t + (long long)b * P.nq;
    int* od = P.out_d + (long long)b * P.nq;

    int limit = 0x7FFFFFFF;
    if (P.flags & HM_FLAG_DIST_THRESHOLD) {
        if (tid == 0) s_min = 0xFFFFFFFFu;
        __syncthreads();
        unsigned m = 0xFFFFFFFFu;
        for (long long r = tid; r < P.nq; r += kFilterThreads) {
            const unsigned long long k1 = fwd[r * 2];
            if (k1 != kNoMatch) m = min(m, (unsigned)(k1 >> 32));
        }
        for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        if (lane == 0) atomicMin(&s_min, m);
        __syncthreads();
        const unsigned mn = s_min;
        limit = mn == 0xFFFFFFFFu ? 0 : max((int)(2 * mn), P.thr_ceil);
    }
    if (tid == 0) s_base = 0;
    __syncthreads();

    for (long long r0 = 0; r0 < P.nq; r0 += kFilterThreads) {
        const long long r = r0 + tid;
        int keep = 0, t1 = 0, d1 = 0;
        if (r < P.nq) {
            const ulonglong2 k = *reinterpret_cast<const ulonglong2*>(fwd + r * 2);
            keep = k.x != kNoMatch;
            t1 = (int)(unsigned)(k.x & 0xFFFFFFFFull);
            d1 = (int)(k.x >> 32);
            if (keep && (P.flags & HM_FLAG_RATIO)) {
                keep = (k.y != kNoMatch) && d1 < (int)P.lut.v[min((unsigned)(k.y >> 32), 256u)];
            }
            if (keep && (P.flags & HM_FLAG_MUTUAL)) {
                keep = (long long)(bwd[(long long)t1 * 2] & 0xFFFFFFFFull) == r;
            }
            if (keep && (P.flags & HM_FLAG_DIST_THRESHOLD)) keep = d1 < limit;
        }
        // ordered block scan of the keep flags
        const unsigned ballot = __ballot_sync(0xFFFFFFFFu, keep);
        const int prefix = __popc(ballot & ((1u << lane) - 1));
        if (lane == 0) warp_sums[wid] = __popc(ballot);
        __syncthreads();
        int wbase = 0, total = 0;
        {
            int v = lane < (kFilterThreads / 32) ? warp_sums[lane] : 0;
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                int n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += n;
            }
            wbase = __shfl_sync(0xFFFFFFFFu, incl - v, wid);
            total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        const int base = s_base;
        if (keep) {
            const int pos = base + wbase + prefix;
            oq[pos] = (int)r;
            ot[pos] = t1;
            od[pos] = d1;
        }
        __syncthreads();
        if (tid == 0) s_base = base + total;
        __syncthreads();
    }
    if (tid == 0) P.out_count[b] = s_base;
}

}  // namespace

int launch_merge_top2(const unsigned long long* keys, int groups, long long rows, unsigned long long* out,
                      cudaStream_t stream)
{
    if (rows <= 0) return HM_OK;
    const int threads = 64;      // few rows: spread them over many SMs
    const long long blocks = ceil_div(rows, threads);
    hm_merge_top2_kernel<<<(unsigned)blocks, threads, 0, stream>>>(keys, groups, rows, out);
    HM_CUDA_CHECK(cudaGetLastError());
    return HM_OK;
}

size_t exchange_bytes(long long max_rows, int world)
{
    return exchange_keys_bytes(max_rows, world) + (size_t)world * exchange_max_ctas(max_rows) * sizeof(unsigned) + 256;
}

int launch_exchange_merge(const unsigned long long* local_keys, int local_groups, long long rows, int world, int rank,
                          void* const* peers, long long max_rows, unsigned epoch, unsigned long long* out,
                          cudaStream_t stream)
{
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || rows > max_rows || rows <= 0 || epoch == 0 ||
        local_groups < 1) {
        set_error("hm_exchange_merge: bad arguments (world %d rank %d rows %lld max_rows %lld epoch %u)", world, rank,
                  rows, max_rows, epoch);
        return HM_ERR_INVALID_ARGUMENT;
    }
    ExchangeParams P{};
    P.local = local_keys; P.local_groups = local_groups; P.out = out; P.rows = rows; P.max_rows = max_rows; P.world = world; P.rank = rank; P.epoch = epoch;
    for (int i = 0; i < world; ++i) {
        if (!peers[i]) {
            set_error("hm_exchange_merge: null peer buffer %d", i);
            return HM_ERR_INVALID_ARGUMENT;
        }
        P.peer[i] = static_cast<unsigned char*>(peers[i]);
    }
    // the grid must cover max_rows (not just rows) so that every rank runs the same CTAs and flags
    const long long ctas = exchange_max_ctas(max_rows);
    hm_exchange_merge_kernel<<<(unsigned)ctas, kExchangeThreads, 0, stream>>>(P);
    HM_CUDA_CHECK(cudaGetLastError());
    return HM_OK;
}

int launch_filter(const unsigned long long* fwd, long long nq, const unsigned long long* bwd, long long nt,
                  int batch, unsigned flags, const RatioLut& lut, int thr_ceil, int* out_q, int* out_t,
                  int* out_d, int* out_count, cudaStream_t stream)
{
    if (batch <= 0) return HM_OK;
    FilterParams P{};
    P.fwd = fwd; P.bwd = bwd; P.nq = nq; P.nt = nt; P.flags = flags;
    P.out_q = out_q; P.out_t = out_t; P.out_d = out_d; P.out_count = out_count;
    P.lut = lut;
    P.thr_ceil = thr_ceil;
    hm_filter_kernel<<<batch, kFilterThreads, 0, stream>>>(P);
    HM_CUDA_CHECK(cudaGetLastError());
    return HM_OK;
}

}  // namespace hm
