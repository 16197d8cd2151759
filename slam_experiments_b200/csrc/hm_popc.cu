// Variant (a): LOP3 XOR + POPC distance core with a register top-2.
//
// Replaces the O(Nq*Nt) normHamming loop that cv::batchDistance runs for
// cv2.BFMatcher.knnMatch (reached from /root/reference/feature_matchers.py:39).
//
// Mapping: one thread owns QPT query descriptors in registers (8 x 32-bit words
// each); the CTA streams its slice of the train set through shared memory with
// cp.async double buffering (16-byte requests, fully coalesced), and every lane
// reads the same train row (two LDS.128 broadcasts, conflict free).  Per pair:
// 8 LOP3 (xor) + 8 POPC + 4 IADD3, then a 3-instruction top-2 update on a packed
// 32-bit key (dist << 23 | local row), so unsigned min() is cv2's tie rule
// (lowest trainIdx).  The train dimension is split over blockIdx.y so small
// problems still fill 148 SMs; split results are 64-bit keys merged by
// hm_merge_top2 (same kernel as the multi-GPU merge).
//
// Roofline: POPC issue rate (XU pipe), 8 POPC per pair -- see DESIGN.md.
#include "hm_common.cuh"

namespace hm {

namespace {

constexpr int kThreads = 128;
constexpr int kTileRows = 256;      // train rows per smem stage (8 KB)
constexpr int kLocalBits = 23;      // local row index bits inside the 32-bit key
constexpr long long kMaxChunk = 1ll << kLocalBits;

struct PopcParams {
    KnnProblem p;
    long long chunk;                 // train rows per split
    int splits;
    unsigned long long* out;         // [split][batch][nq][2] (split stride 0 when splits == 1)
    long long out_split_stride;      // in keys
    unsigned long long* final_out;   // [batch][nq][2], written by the last CTA of each query tile (splits > 1)
    unsigned* counters;              // [batch][query tiles] arrival counters, zeroed by the launcher
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ unsigned dist256(const unsigned (&q)[8], const uint4& a, const uint4& b)
{
    return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
           __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
}

template <int QPT>
__global__ void __launch_bounds__(kThreads) hm_popc_knn2_kernel(const PopcParams P)
{
    __shared__ __align__(16) uint4 tile[2][kTileRows * 2];

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const long long row0 = (long long)blockIdx.y * P.chunk;           // first train row of this split
    const long long rows = min(P.chunk, P.p.nt - row0);               // > 0 by construction
    const uint8_t* tbase = P.p.t + (long long)b * P.p.t_batch_stride + row0 * P.p.t_stride;
    const uint8_t* qbase = P.p.q + (long long)b * P.p.q_batch_stride;

    // ---- this thread's queries -> registers (two 128-bit loads each) -------------------
    unsigned q[QPT][8];
    long long qrow[QPT];
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
        qrow[u] = ((long long)blockIdx.x * QPT + u) * kThreads + tid;
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (qrow[u] < P.p.nq) {
            const uint4* src = reinterpret_cast<const uint4*>(qbase + qrow[u] * P.p.q_stride);
            lo = __ldg(src);
            hi = __ldg(src + 1);
        }
        q[u][0] = lo.x; q[u][1] = lo.y; q[u][2] = lo.z; q[u][3] = lo.w;
        q[u][4] = hi.x; q[u][5] = hi.y; q[u][6] = hi.z; q[u][7] = hi.w;
    }
    unsigned m1[QPT], m2[QPT];
#pragma unroll
    for (int u = 0; u < QPT; ++u) m1[u] = m2[u] = 0xFFFFFFFFu;

    const int ntiles = (int)ceil_div(rows, kTileRows);
    auto issue = [&](int tix, int buf) {
        const long long r0 = (long long)tix * kTileRows;
        const int valid = (int)min((long long)kTileRows, rows - r0);
        // 2 x 16-byte pieces per row; consecutive threads take consecutive pieces
        for (int c = tid; c < valid * 2; c += kThreads) {
            const int r = c >> 1, h = c & 1;
            cp_async16(&tile[buf][c], tbase + (r0 + r) * P.p.t_stride + h * 16);
        }
        cp_async_commit();
    };

    issue(0, 0);
    for (int tix = 0; tix < ntiles; ++tix) {
        const int buf = tix & 1;
        if (tix + 1 < ntiles) {
            issue(tix + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int valid = (int)min((long long)kTileRows, rows - (long long)tix * kTileRows);
        const unsigned jbase = (unsigned)tix * kTileRows;
        const uint4* s = tile[buf];
        int j = 0;
#pragma unroll 1
        for (; j + 4 <= valid; j += 4) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const uint4 a = s[(j + jj) * 2], c = s[(j + jj) * 2 + 1];
#pragma unroll
                for (int u = 0; u < QPT; ++u) {
                    const unsigned key = (dist256(q[u], a, c) << kLocalBits) + (jbase + j + jj);
                    m2[u] = min(m2[u], max(m1[u], key));
                    m1[u] = min(m1[u], key);
                }
            }
        }
        for (; j < valid; ++j) {
            const uint4 a = s[j * 2], c = s[j * 2 + 1];
#pragma unroll
            for (int u = 0; u < QPT; ++u) {
                const unsigned key = (dist256(q[u], a, c) << kLocalBits) + (jbase + j);
                m2[u] = min(m2[u], max(m1[u], key));
                m1[u] = min(m1[u], key);
            }
        }
        __syncthreads();   // everyone done with `buf` before it is refilled
    }

    // ---- widen to the 64-bit global key and store ---------------------------------------
    unsigned long long* out = P.out + (long long)blockIdx.y * P.out_split_stride + (long long)b * P.p.nq * 2;
    const unsigned long long gbase = P.p.train_base + (unsigned long long)row0;
#pragma unroll
    for (int u = 0; u < QPT; ++u) {
        if (qrow[u] < P.p.nq) {
            ulonglong2 k;
            k.x = m1[u] == 0xFFFFFFFFu ? kNoMatch
                : ((unsigned long long)(m1[u] >> kLocalBits) << 32) | (gbase + (m1[u] & (kMaxChunk - 1)));
            k.y = m2[u] == 0xFFFFFFFFu ? kNoMatch
                : ((unsigned long long)(m2[u] >> kLocalBits) << 32) | (gbase + (m2[u] & (kMaxChunk - 1)));
            *reinterpret_cast<ulonglong2*>(out + qrow[u] * 2) = k;
        }
    }
    // ---- in-kernel merge of the train splits: the last CTA of this query tile folds all partials ----
    if (P.counters) {
        __shared__ int last_flag;
        if (last_cta_arrives(&P.counters[(long long)b * gridDim.x + blockIdx.x], (unsigned)P.splits, &last_flag)) {
#pragma unroll
            for (int u = 0; u < QPT; ++u) {
                if (qrow[u] < P.p.nq) {
                    unsigned long long k1 = kNoMatch, k2 = kNoMatch;
                    fold_partials(P.out, P.splits, P.out_split_stride, (long long)b * P.p.nq + qrow[u], k1, k2);
                    *reinterpret_cast<ulonglong2*>(P.final_out + ((long long)b * P.p.nq + qrow[u]) * 2) = make_ulonglong2(k1, k2);
                }
            }
        }
    }
}

int queries_per_thread(long long nq, int batch, int sm_count)
{
    // two queries per thread halve the shared-memory reads per pair once the query side
    // alone fills the machine
    return (nq * batch >= (long long)sm_count * kThreads * 4) ? 2 : 1;
}

void plan(const KnnProblem& p, int sm_count, int* qpt, int* splits, long long* chunk)
{
    *qpt = queries_per_thread(p.nq, p.batch, sm_count);
    const long long qtiles = ceil_div(p.nq, (long long)kThreads * *qpt) * p.batch;
    const long long target = (long long)sm_count * 4;            // CTAs wanted in flight
    long long s = ceil_div(target, qtiles);
    s = min(s, ceil_div(p.nt, 64));                             // at least 64 rows per split
    s = max(s, ceil_div(p.nt, kMaxChunk));
    s = max(s, 1ll);
    s = min(s, 65535ll);
    long long c = ceil_div(p.nt, s);
    *chunk = c;
    *splits = (int)ceil_div(p.nt, c);
}

}  // namespace

int popc_splits(const KnnProblem& p, int sm_count)
{
    int qpt, splits;
    long long chunk;
    plan(p, sm_count, &qpt, &splits, &chunk);
    return splits;
}

size_t popc_workspace_bytes(long long nq, long long nt, int batch, int sm_count)
{
    KnnProblem p{};
    p.nq = nq; p.nt = nt; p.batch = batch;
    int qpt, s;
    long long chunk;
    plan(p, sm_count, &qpt, &s, &chunk);
    if (s <= 1) return 0;
    const long long qtiles = ceil_div(nq, (long long)kThreads * qpt);
    return counters_bytes(qtiles * batch) + (size_t)s * batch * nq * 2 * sizeof(unsigned long long);
}

int launch_popc_knn2(const KnnProblem& p, unsigned long long* out, void* ws, size_t ws_bytes, int sm_count,
                     cudaStream_t stream)
{
    int qpt, splits;
    long long chunk;
    plan(p, sm_count, &qpt, &splits, &chunk);
    if (ceil_div(p.nt, chunk) * chunk > (1ll << 32)) {
        set_error("train set too large for 32-bit trainIdx");
        return HM_ERR_UNSUPPORTED;
    }
    PopcParams P{};
    P.p = p;
    P.chunk = chunk;
    P.splits = splits;
    const long long rows = p.nq * p.batch;
    const long long qtiles = ceil_div(p.nq, (long long)kThreads * qpt);
    if (splits > 1) {
        const size_t cbytes = counters_bytes(qtiles * p.batch);
        const size_t need = cbytes + (size_t)splits * rows * 2 * sizeof(unsigned long long);
        if (!ws || ws_bytes < need) {
            set_error("workspace too small: need %zu bytes, got %zu", need, ws_bytes);
            return HM_ERR_WORKSPACE;
        }
        P.counters = static_cast<unsigned*>(ws);
        P.final_out = out;
        P.out = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ws) + cbytes);
        P.out_split_stride = rows * 2;
        HM_CUDA_CHECK(cudaMemsetAsync(P.counters, 0, cbytes, stream));
    } else {
        P.out = out;
        P.out_split_stride = 0;
    }
    if (qtiles > 0x7FFFFFFFll || p.batch > 65535) {
        set_error("grid too large");
        return HM_ERR_UNSUPPORTED;
    }
    dim3 grid((unsigned)qtiles, (unsigned)splits, (unsigned)p.batch);
    profile_mark(true, stream);
    if (qpt == 2)
        hm_popc_knn2_kernel<2><<<grid, kThreads, 0, stream>>>(P);
    else
        hm_popc_knn2_kernel<1><<<grid, kThreads, 0, stream>>>(P);
    profile_mark(false, stream);
    HM_CUDA_CHECK(cudaGetLastError());
    return HM_OK;
}

}  // namespace hm
