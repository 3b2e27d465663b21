// K3 — warp-per-row CSR SpMM with 128-bit embedding gathers and a fused layer-mean epilogue.
//
// Replaces torch.sparse.mm(norm_adj, E) (model/LightGCN.py:90 and the other call sites listed in
// include/coldrec_b200.h) plus the stack/mean of LightGCN.py:92-93.  HBM-bound: per nonzero one
// 4-byte column id, one 4-byte value and one d*4-byte row gather; per row one write of y and an
// optional read-modify-write of the running layer sum.
//
// Layout: a row of d floats is covered by LPR lanes holding NV float4 each (d = 4*LPR*NV); the
// 32/LPR lane groups of a warp walk different nonzeros of the same row, U nonzeros each in flight,
// so a warp keeps U*(32/LPR) row gathers (2 KB at d=64) outstanding.  Rows longer than kLongRow are
// cut into kChunk-nonzero chunks (one warp each, partial sums in the plan buffer) and reduced in
// chunk order by a third kernel, so results do not depend on scheduling.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kLongRow = 512;       // graphs of >= kSmallNnz nonzeros: rows longer than this take the split path ...
constexpr int kChunk = 512;         // ... in chunks of this many nonzeros
constexpr int kSmallNnz = 1 << 23;  // dataset-sized graphs (MovieLens / CiteULike / XING: 10^5 .. 10^6.5 nonzeros) are latency
constexpr int kSmallLongRow = 64;   // bound: a 500-nonzero row walked by one 8-lane group is the whole kernel's critical path
constexpr int kSmallChunk = 64;     // there, so they split at 64 nonzeros and spread the chunks over the idle SMs
constexpr int kThreads = 256;
// d = 64 geometry of the grouped kernel (A/B knobs, tools/build_variant.sh): lanes per row, gathers in flight per lane group,
// CTAs per SM.  Measured on one box at C4 (3-layer propagation, ms): 8 lanes x 2 float4, U=4, 3 CTAs (80 regs) 31.6;
// same with 4 CTAs (64 regs) 32.6; U=8 with 2 CTAs 33.0; 16 lanes x 1 float4: U=4 / 5 CTAs (48 regs) 41.9, U=8 / 5 CTAs 42.5,
// U=8 / 3 CTAs 30.3, U=16 / 2 CTAs 34.9, **U=8 / 4 CTAs (64 regs) 29.4**; 4 lanes x 4 float4: 36.4 - 40.8.
#ifndef CR_SPMM_GLPR
#define CR_SPMM_GLPR 16
#endif
#ifndef CR_SPMM_U
#define CR_SPMM_U 8
#endif
#ifndef CR_SPMM_MINB
#define CR_SPMM_MINB 4
#endif
#ifndef CR_SPMM_LONG_U
#define CR_SPMM_LONG_U 4      // gathers in flight per lane group in the warp-per-row / long-row-chunk kernels
#endif
// Work-balanced row ranges (graphs with a plan, bandwidth-bound geometry): warp w owns the consecutive rows whose cumulated
// weight  rowptr[r] + kRowCost * r  falls in [w * kWarpWork, (w + 1) * kWarpWork) — about 2048 nonzero-equivalents, whether
// that is 20 user rows of 100 nonzeros or 200 item rows of 2.  With a fixed 128 rows per warp a row block of a partitioned
// bipartite graph (8 GPUs: 125k user rows + 1.26M item rows each) put half of its nonzeros into 977 warps of 12,800
// nonzeros each — a 3.3 MB dependent gather stream per warp that set the duration of the whole launch (r02 probe: 2.4 ms
// per layer and GPU against 1.2 ms of HBM time).  The table lives in the plan (built once per graph).
constexpr int kRowCost = 8;
constexpr int kWarpWork = 2048;
inline int64_t balanced_warps(int64_t n_rows, int64_t nnz) { return (nnz + (int64_t)kRowCost * n_rows + kWarpWork - 1) / kWarpWork; }

inline int long_row_of(int64_t nnz) { return nnz >= kSmallNnz ? kLongRow : kSmallLongRow; }
inline int chunk_of(int64_t nnz) { return nnz >= kSmallNnz ? kChunk : kSmallChunk; }

struct PlanHeader {
    int n_chunks;
    int n_long;
    int overflow;
    int pad;
};
struct LongRow {
    int row;
    int chunk_base;
    int n_chunks;
    int pad;
};
struct Chunk {
    int64_t start;
    int len;
    int row;
};

struct PlanLayout {
    int64_t max_long, max_chunks;
    size_t off_long, off_chunks, off_partial, off_warp_rows, total;     // total = up to the warp table; + warp_table_bytes()
};
inline size_t warp_table_bytes(int64_t n_rows, int64_t nnz) { return cr::align_up((size_t)(balanced_warps(n_rows, nnz) + 2) * sizeof(int32_t), 256); }

PlanLayout plan_layout(int64_t nnz, int d) {
    PlanLayout L;
    L.max_long = nnz / (long_row_of(nnz) + 1) + 1;
    L.max_chunks = nnz / chunk_of(nnz) + L.max_long + 1;
    L.off_long = 256;
    L.off_chunks = cr::align_up(L.off_long + (size_t)L.max_long * sizeof(LongRow), 256);
    L.off_partial = cr::align_up(L.off_chunks + (size_t)L.max_chunks * sizeof(Chunk), 256);
    L.off_warp_rows = cr::align_up(L.off_partial + (size_t)L.max_chunks * d * sizeof(float), 256);
    L.total = L.off_warp_rows;      // offsets depend on (nnz, d) only: a plan built for n rows serves calls on its first m <= n rows
    return L;
}

__device__ __forceinline__ void fma4(float4& a, float w, const float4& x) {
    a.x = fmaf(w, x.x, a.x);
    a.y = fmaf(w, x.y, a.y);
    a.z = fmaf(w, x.z, a.z);
    a.w = fmaf(w, x.w, a.w);
}

// Accumulate nonzeros [s, e) of one row into per-lane-group partial sums a[].
template <int LPR, int NV, bool HAS_VAL, bool BOUNDS>
__device__ __forceinline__ void accumulate_range(const int32_t* __restrict__ col, const float* __restrict__ val,
                                                 int64_t s, int64_t e, const float4* __restrict__ X4, int d4, int lane,
                                                 float4 (&a)[NV]) {
    constexpr int G = 32 / LPR;
    constexpr int U = CR_SPMM_LONG_U;
    const int grp = lane / LPR, sub = lane % LPR;
    for (int64_t base = s; base < e; base += 32) {
        const int64_t j = base + lane;
        int c = 0;
        float v = 0.f;
        if (j < e) {
            c = __ldg(col + j);
            v = HAS_VAL ? __ldg(val + j) : 1.f;
        }
        const int cnt = (int)min((int64_t)32, e - base);
        for (int t = 0; t < cnt; t += G * U) {
            float4 x[U][NV];
            float w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int src = t + u * G + grp;
                const int cc = __shfl_sync(CR_FULL_MASK, c, src & 31);
                const float ww = __shfl_sync(CR_FULL_MASK, v, src & 31);
                const bool ok = src < cnt;
                w[u] = ok ? ww : 0.f;
                const float4* p = X4 + (int64_t)cc * d4 + sub;
#pragma unroll
                for (int nv = 0; nv < NV; ++nv) {
                    const bool okc = ok && (!BOUNDS || sub + nv * LPR < d4);
                    x[u][nv] = okc ? __ldg(p + nv * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int nv = 0; nv < NV; ++nv) fma4(a[nv], w[u], x[u][nv]);
        }
    }
}

template <int LPR, int NV>
__device__ __forceinline__ void reduce_groups(float4 (&a)[NV]) {
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
        for (int nv = 0; nv < NV; ++nv) {
            a[nv].x += __shfl_xor_sync(CR_FULL_MASK, a[nv].x, off);
            a[nv].y += __shfl_xor_sync(CR_FULL_MASK, a[nv].y, off);
            a[nv].z += __shfl_xor_sync(CR_FULL_MASK, a[nv].z, off);
            a[nv].w += __shfl_xor_sync(CR_FULL_MASK, a[nv].w, off);
        }
}

// Where a finished row goes: Y and/or the running layer sum, and (multi-GPU) the same row of every peer's
// gather table — the all-gather of the next layer's input fused into the SpMM epilogue as peer-memory stores.
struct Epi {
    float4* Y4; const float4* acc_in4; float4* acc4; float beta, div;
    float4* const* peers; int n_peers;
    // peers[p][peer_off4 + idx] = y (or the acc result) for idx < peer_split4, peers[p][peer_off_hi4 + idx] above: local rows
    // [0, split) and [split, n) may land in two different row ranges of the destination (a rank owns one range of user rows
    // and one of item rows: the last layer is scattered straight into the original node numbering)
    int64_t peer_off4, peer_split4, peer_off_hi4;
    int peers_get_acc;
    // optional: bit p of peer_need[row] says whether GPU p gathers this row at all (its row block has a nonzero in that
    // column).  Most item rows of a bipartite graph have a handful of nonzeros, i.e. a handful of readers: the fused
    // all-gather becomes a sparse one (C4, 8 GPUs: 2.9 remote copies per row instead of 7).
    const uint8_t* peer_need;
    // optional NVLS multicast alias of the destination table: ONE multimem.st replicated by the switch replaces the n_peers
    // unicast stores of a row that every GPU wants (user rows in every layer, all rows of a fully replicated result)
    float4* mc4; unsigned all_mask;
};

__device__ __forceinline__ void multimem_st4(float4* p, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void store_peers(const Epi& ep, unsigned need, int64_t di, const float4& v) {
    if (ep.mc4 && (need & ep.all_mask) == ep.all_mask) {
        multimem_st4(ep.mc4 + di, v);
        return;
    }
    for (int p = 0; p < ep.n_peers; ++p)
        if ((need >> p) & 1u) ep.peers[p][di] = v;
}

__device__ __forceinline__ int64_t peer_index(const Epi& ep, int64_t idx) {
    return idx + (idx < ep.peer_split4 ? ep.peer_off4 : ep.peer_off_hi4);
}

// Local part of the epilogue (Y / running layer sum) for one float4; returns the value the peers receive.
__device__ __forceinline__ float4 epilogue_local(const Epi& ep, int64_t idx, float4 y) {
    if (ep.Y4) ep.Y4[idx] = y;
    if (!ep.acc4) return y;
    float4 o = y;
    if (ep.beta != 0.f) {
        const float4 q = ep.acc_in4[idx];
        o.x = fmaf(ep.beta, q.x, y.x);
        o.y = fmaf(ep.beta, q.y, y.y);
        o.z = fmaf(ep.beta, q.z, y.z);
        o.w = fmaf(ep.beta, q.w, y.w);
    }
    if (ep.div != 1.f) {
        o.x = __fdiv_rn(o.x, ep.div);
        o.y = __fdiv_rn(o.y, ep.div);
        o.z = __fdiv_rn(o.z, ep.div);
        o.w = __fdiv_rn(o.w, ep.div);
    }
    ep.acc4[idx] = o;
    return ep.peers_get_acc ? o : y;
}

// y -> Y and/or acc = (beta*acc_in + y) / div for one float4 of one row (acc_in may alias acc), then the SM-issued peer stores.
__device__ __forceinline__ void store_epilogue(const Epi& ep, int64_t row, int64_t idx, float4 y) {
    const float4 v = epilogue_local(ep, idx, y);
    if (ep.peers) {
        const unsigned need = ep.peer_need ? (unsigned)__ldg(ep.peer_need + row) : 0xffffffffu;
        store_peers(ep, need, peer_index(ep, idx), v);
    }
}

// ---- staged peer stores (multi-GPU): finished rows are parked in shared memory, 8 consecutive rows per half-buffer, and
// pushed to the peers' gather tables by the TMA engine (cp.async.bulk shared -> peer global over NVLink) in runs of
// consecutive rows per destination: 256 B for a row only some GPUs read, up to 2 KB where 8 neighbours go to the same
// GPU (user rows in every layer, every row of the replicated last layer).  The SM-issued variant above (one 16-byte
// st.global per lane and destination) held the whole SpMM back once 4+ GPUs exchange rows (r01: 8 GPUs slower than 4):
// the stores sit in the LSU queues in front of the row gathers.  Bulk copies are asynchronous, leave the load path
// alone and are tracked per issuing thread (bulk groups), so a warp only waits when both of its half-buffers are in flight.
constexpr int kStageRows = 8;        // rows per half-buffer (two half-buffers per warp)

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One half-buffer (rows [row0, row0 + 8) of this warp, d4 float4 each) -> every GPU that reads them.  Lane p serves
// destination p: its bit mask over the 8 rows is cut into runs of consecutive rows (and at the boundary between the two
// destination ranges), one bulk copy per run.  skip8: rows not produced by this warp (split-path rows, rows past the end).
__device__ __forceinline__ void flush_stage(const Epi& ep, const float4* half, int64_t row0, int64_t n_rows, unsigned skip8,
                                            int d4, int lane) {
    fence_proxy_async_smem();        // this lane's st.shared -> visible to the async proxy (TMA) ...
    __syncwarp();                    // ... and every lane's before any lane issues a copy
    if (lane < ep.n_peers) {
        unsigned m = 0;
#pragma unroll
        for (int j = 0; j < kStageRows; ++j) {
            const int64_t row = row0 + j;
            if (row < n_rows && !((skip8 >> j) & 1u)) {
                const unsigned nd = ep.peer_need ? (unsigned)__ldg(ep.peer_need + row) : 0xffffffffu;
                m |= ((nd >> lane) & 1u) << j;
            }
        }
        float4* dst = ep.peers[lane];
        while (m) {
            const int j0 = __ffs(m) - 1;
            int len = __ffs(~(m >> j0)) - 1;                        // consecutive rows wanted by this GPU
            const int64_t idx0 = (row0 + j0) * d4;
            if (idx0 < ep.peer_split4 && idx0 + (int64_t)len * d4 > ep.peer_split4) len = (int)((ep.peer_split4 - idx0) / d4);
            bulk_store_s2g(dst + peer_index(ep, idx0), half + j0 * d4, (uint32_t)(len * d4 * sizeof(float4)));
            m &= ~(((1u << len) - 1u) << j0);
        }
        bulk_commit();               // exactly one group per flush and lane (empty ones included): wait_group counts flushes
    }
}

template <int LPR, int NV, bool HAS_VAL, bool BOUNDS>
__global__ void __launch_bounds__(kThreads)
spmm_rows_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
                 int64_t n_rows, const float4* __restrict__ X4, int d4, const Epi ep, int long_row) {
    const int64_t row = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const int64_t s = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
    if (e - s > long_row) return;   // split path owns this row
    float4 a[NV];
#pragma unroll
    for (int nv = 0; nv < NV; ++nv) a[nv] = make_float4(0.f, 0.f, 0.f, 0.f);
    accumulate_range<LPR, NV, HAS_VAL, BOUNDS>(col, val, s, e, X4, d4, lane, a);
    reduce_groups<LPR, NV>(a);
    if (lane < LPR) {
#pragma unroll
        for (int nv = 0; nv < NV; ++nv)
            if (!BOUNDS || lane + nv * LPR < d4) store_epilogue(ep, row, row * d4 + lane + nv * LPR, a[nv]);
    }
}

// One warp per chunk of a long row (grid-stride over the plan's chunk list), partial sums to the plan buffer.  Out of line:
// it must not take registers from the row loop of the kernel that hosts it.
template <int LPR, int NV, bool HAS_VAL>
__device__ __noinline__ void walk_chunks(const PlanHeader* __restrict__ hdr, const Chunk* __restrict__ chunks,
                                         const int32_t* __restrict__ col, const float* __restrict__ val,
                                         const float4* __restrict__ X4, float4* __restrict__ partial4, int chunk_ctas) {
    constexpr int d4 = LPR * NV;
    const int lane = threadIdx.x & 31;
    const int n_chunks = hdr->n_chunks;
    const int warps = (chunk_ctas * kThreads) >> 5;
    for (int c = (blockIdx.x * kThreads + threadIdx.x) >> 5; c < n_chunks; c += warps) {
        const Chunk ch = chunks[c];
        float4 a[NV];
#pragma unroll
        for (int nv = 0; nv < NV; ++nv) a[nv] = make_float4(0.f, 0.f, 0.f, 0.f);
        accumulate_range<LPR, NV, HAS_VAL, false>(col, val, ch.start, ch.start + ch.len, X4, d4, lane, a);
        reduce_groups<LPR, NV>(a);
        if (lane < LPR) {
#pragma unroll
            for (int nv = 0; nv < NV; ++nv) partial4[(int64_t)c * d4 + lane + nv * LPR] = a[nv];
        }
    }
}

// Grouped variant (d = 4*LPR*NV with LPR < 32): the 32/LPR lane groups of a warp walk DIFFERENT rows.
// ncu (profiles/r01) showed the warp-per-row kernel latency bound, not bandwidth bound (DRAM 31 % busy, 20
// resident warps, long-scoreboard stalls): 80 % of the rows of a bipartite item-side graph have <= 8 nonzeros,
// and each such row paid the full rowptr -> col -> gather dependency chain for a few hundred bytes.  Here a warp
// owns 128 consecutive rows: row pointers are fetched 32 rows at a time, column ids / values of the next batch
// of rows and of the next chunk of the same row are prefetched while the current gathers are in flight, and
// 32/LPR rows are gathered concurrently, so the chain is paid once per 32 rows instead of once per row.
template <int LPR, int NV, bool HAS_VAL, int U, int MINB, bool STAGED = false>
__global__ void __launch_bounds__(kThreads, MINB)
spmm_rows_grouped_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
                         int64_t n_rows, const float4* __restrict__ X4, const Epi ep, int long_row, int kRowsPerWarp,
                         const PlanHeader* __restrict__ hdr, const Chunk* __restrict__ chunks, float4* __restrict__ partial4,
                         int chunk_ctas, const int32_t* __restrict__ warp_rows, int64_t n_warps) {
    constexpr int RPW = 32 / LPR;     // rows in flight per warp; U = gathers issued back to back per group
    constexpr int d4 = LPR * NV;
    static_assert(LPR % U == 0, "a group's LPR column ids are consumed U at a time");
    const int lane = threadIdx.x & 31, grp = lane / LPR, sub = lane % LPR;
    // STAGED (multi-GPU): two half-buffers of kStageRows rows per warp for the TMA peer stores (see flush_stage)
    constexpr int kBatchesPerHalf = kStageRows / RPW;
    static_assert(kStageRows % RPW == 0 && (32 / kStageRows) % 2 == 0, "half-buffers alternate 0,1 inside every 32-row block");
    __shared__ __align__(128) float4 s_stage[STAGED ? (kThreads / 32) * 2 * kStageRows * d4 : 1];
    float4* const stage = s_stage + (STAGED ? (threadIdx.x >> 5) * 2 * kStageRows * d4 : 0);
    if ((int)blockIdx.x < chunk_ctas) {
        // The first CTAs of the grid walk the chunks of the long rows (one warp per chunk, partial sums to the plan buffer):
        // the longest work items start first and share the launch — and its tail — with the ordinary rows instead of
        // following them in a kernel of their own.
        walk_chunks<LPR, NV, HAS_VAL>(hdr, chunks, col, val, X4, partial4, chunk_ctas);
        return;
    }
    const int64_t wid = ((int64_t)(blockIdx.x - chunk_ctas) * kThreads + threadIdx.x) >> 5;
    int64_t row0 = wid * kRowsPerWarp, row_end = row0 + kRowsPerWarp;
    if (warp_rows) {                  // work-balanced ranges from the plan
        if (wid >= n_warps) return;
        row0 = __ldg(warp_rows + wid);
        row_end = __ldg(warp_rows + wid + 1);
    }
    row_end = min(n_rows, row_end);
    if (row0 >= row_end) return;
    for (int64_t r32 = row0; r32 < row_end; r32 += 32) {
        const int64_t rr = r32 + lane;
        int64_t lo = __ldg(rowptr + min(rr, n_rows));
        int64_t hi = __ldg(rowptr + min(rr + 1, n_rows));
        const bool skip_row = (rr >= row_end) || (hi - lo > long_row);  // long rows belong to the split path, later rows to the next warp
        if (skip_row) hi = lo;
        [[maybe_unused]] const unsigned skip32 = STAGED ? __ballot_sync(CR_FULL_MASK, skip_row) : 0u;
        int c_nb = 0;
        float v_nb = 0.f;
#pragma unroll 1
        for (int b = 0; b < 32 / RPW; ++b) {
            const int src = b * RPW + grp;
            const int64_t s = __shfl_sync(CR_FULL_MASK, lo, src), e = __shfl_sync(CR_FULL_MASK, hi, src);
            const bool skip = __shfl_sync(CR_FULL_MASK, (int)skip_row, src) != 0;
            const int64_t row = r32 + src;
            int c = c_nb;
            float v = v_nb;
            if (b == 0) {          // first batch of a 32-row block: nothing was prefetched
                c = 0; v = 0.f;
                if (s + sub < e) { c = __ldg(col + s + sub); v = HAS_VAL ? __ldg(val + s + sub) : 1.f; }
            }
            if (b + 1 < 32 / RPW) {   // first chunk of the next batch's rows
                const int64_t s2 = __shfl_sync(CR_FULL_MASK, lo, src + RPW), e2 = __shfl_sync(CR_FULL_MASK, hi, src + RPW);
                c_nb = 0; v_nb = 0.f;
                if (s2 + sub < e2) { c_nb = __ldg(col + s2 + sub); v_nb = HAS_VAL ? __ldg(val + s2 + sub) : 1.f; }
            }
            const int iters = (int)((e - s + LPR - 1) / LPR);
            const int max_it = __reduce_max_sync(CR_FULL_MASK, iters);
            float4 a[NV];
#pragma unroll
            for (int nv = 0; nv < NV; ++nv) a[nv] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int it = 0; it < max_it; ++it) {
                const int64_t base = s + (int64_t)it * LPR;
                const int64_t nj = base + LPR + sub;      // next chunk of the same row
                int c_nc = 0;
                float v_nc = 0.f;
                if (nj < e) { c_nc = __ldg(col + nj); v_nc = HAS_VAL ? __ldg(val + nj) : 1.f; }
                const int cnt = (int)max((int64_t)0, min((int64_t)LPR, e - base));
#pragma unroll
                for (int t = 0; t < LPR; t += U) {
                    float4 x[U][NV];
                    float w[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int cc = __shfl_sync(CR_FULL_MASK, c, grp * LPR + t + u);
                        const float ww = __shfl_sync(CR_FULL_MASK, v, grp * LPR + t + u);
                        const bool ok = t + u < cnt;
                        w[u] = ok ? ww : 0.f;
                        const float4* px = X4 + (int64_t)cc * d4 + sub;
#pragma unroll
                        for (int nv = 0; nv < NV; ++nv) x[u][nv] = ok ? __ldg(px + nv * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int nv = 0; nv < NV; ++nv) fma4(a[nv], w[u], x[u][nv]);
                }
                c = c_nc;
                v = v_nc;
            }
            if constexpr (STAGED) {
                const int half = (b / kBatchesPerHalf) & 1, slot = (b % kBatchesPerHalf) * RPW + grp;
                float4* const hb = stage + half * kStageRows * d4;
                if (b % kBatchesPerHalf == 0) {      // about to refill this half: its previous copies must have read it
                    if (lane < ep.n_peers) bulk_wait_read<1>();
                    __syncwarp();
                }
                if (!skip && row < n_rows) {
#pragma unroll
                    for (int nv = 0; nv < NV; ++nv)
                        hb[slot * d4 + sub + nv * LPR] = epilogue_local(ep, row * d4 + sub + nv * LPR, a[nv]);
                }
                if (b % kBatchesPerHalf == kBatchesPerHalf - 1) {
                    const int r8 = (b / kBatchesPerHalf) * kStageRows;
                    flush_stage(ep, hb, r32 + r8, n_rows, (skip32 >> r8) & 0xffu, d4, lane);
                }
            } else if (!skip && row < n_rows) {
#pragma unroll
                for (int nv = 0; nv < NV; ++nv) store_epilogue(ep, row, row * d4 + sub + nv * LPR, a[nv]);
            }
        }
    }
    if constexpr (STAGED) {
        if (lane < ep.n_peers) bulk_wait_all();      // the copies must have landed before the CTA (and its smem) goes away
    }
}

__global__ void spmm_plan_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows, PlanHeader* hdr, LongRow* long_rows,
                                 Chunk* chunks, int max_long, int max_chunks, int kLongRow, int kChunk) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int64_t s = rowptr[row], len = rowptr[row + 1] - s;
    if (len <= kLongRow) return;
    const int nch = (int)((len + kChunk - 1) / kChunk);
    const int base = atomicAdd(&hdr->n_chunks, nch);
    const int li = atomicAdd(&hdr->n_long, 1);
    if (base + nch > max_chunks || li >= max_long) {   // unreachable with plan_layout()'s bounds
        hdr->overflow = 1;
        return;
    }
    long_rows[li] = LongRow{(int)row, base, nch, 0};
    for (int c = 0; c < nch; ++c) {
        const int64_t off = (int64_t)c * kChunk;
        chunks[base + c] = Chunk{s + off, (int)min((int64_t)kChunk, len - off), (int)row};
    }
}

// warp_rows[w] = first row r with rowptr[r] + kRowCost * r >= w * kWarpWork (w = 0 .. n_warps; the last entry is n_rows)
__global__ void spmm_warp_rows_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows, int64_t n_warps, int32_t* __restrict__ warp_rows) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w > n_warps) return;
    const int64_t target = w * kWarpWork;
    int64_t lo = 0, hi = n_rows;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (rowptr[mid] + (int64_t)kRowCost * mid < target) lo = mid + 1; else hi = mid;
    }
    warp_rows[w] = (int32_t)(w == n_warps ? n_rows : lo);
}

template <int LPR, int NV, bool HAS_VAL, bool BOUNDS>
__global__ void __launch_bounds__(kThreads)
spmm_long_chunks_kernel(const PlanHeader* __restrict__ hdr, const Chunk* __restrict__ chunks,
                        const int32_t* __restrict__ col, const float* __restrict__ val, const float4* __restrict__ X4,
                        int d4, float4* __restrict__ partial4) {
    const int lane = threadIdx.x & 31;
    const int n_chunks = hdr->n_chunks;
    const int warps = (gridDim.x * kThreads) >> 5;
    for (int c = (blockIdx.x * kThreads + threadIdx.x) >> 5; c < n_chunks; c += warps) {
        const Chunk ch = chunks[c];
        float4 a[NV];
#pragma unroll
        for (int nv = 0; nv < NV; ++nv) a[nv] = make_float4(0.f, 0.f, 0.f, 0.f);
        accumulate_range<LPR, NV, HAS_VAL, BOUNDS>(col, val, ch.start, ch.start + ch.len, X4, d4, lane, a);
        reduce_groups<LPR, NV>(a);
        if (lane < LPR) {
#pragma unroll
            for (int nv = 0; nv < NV; ++nv)
                if (!BOUNDS || lane + nv * LPR < d4) partial4[(int64_t)c * d4 + lane + nv * LPR] = a[nv];
        }
    }
}

// Partial sums of a long row -> the row, in a fixed order (deterministic, no atomics).  A warp per row; for d <= 64 the
// 32/d4 lane groups take every G-th chunk each and every group keeps four independent running sums, so the head rows
// (hundreds of chunks) cost ~n_chunks/(4G) dependent L2 round trips instead of n_chunks (0.37 ms per layer at C4 before).
__global__ void __launch_bounds__(kThreads)
spmm_long_reduce_kernel(const PlanHeader* __restrict__ hdr, const LongRow* __restrict__ long_rows,
                        const float4* __restrict__ partial4, int d4, const Epi ep) {
    const int lane = threadIdx.x & 31;
    const int n_long = hdr->n_long;
    const int warps = (gridDim.x * kThreads) >> 5;
    const int G = d4 <= 16 ? 32 / d4 : 1;               // d4 in {8, 16} (d = 32, 64): 4 or 2 lane groups; otherwise one
    const auto add = [](float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; };
    for (int li = (blockIdx.x * kThreads + threadIdx.x) >> 5; li < n_long; li += warps) {
        const LongRow lr = long_rows[li];
        if (G > 1 && 32 % d4 == 0) {
            const int g = lane / d4, i = lane % d4;
            const float4* base = partial4 + (int64_t)lr.chunk_base * d4 + i;
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
            int c = g;
            for (; c + 3 * G < lr.n_chunks; c += 4 * G) {
                const float4 p0 = base[(int64_t)c * d4], p1 = base[(int64_t)(c + G) * d4];
                const float4 p2 = base[(int64_t)(c + 2 * G) * d4], p3 = base[(int64_t)(c + 3 * G) * d4];
                add(s0, p0); add(s1, p1); add(s2, p2); add(s3, p3);
            }
            for (; c < lr.n_chunks; c += G) add(s0, base[(int64_t)c * d4]);
            add(s0, s1); add(s2, s3); add(s0, s2);
            for (int off = d4; off < 32; off <<= 1) {
                s0.x += __shfl_down_sync(CR_FULL_MASK, s0.x, off);
                s0.y += __shfl_down_sync(CR_FULL_MASK, s0.y, off);
                s0.z += __shfl_down_sync(CR_FULL_MASK, s0.z, off);
                s0.w += __shfl_down_sync(CR_FULL_MASK, s0.w, off);
            }
            if (lane < d4) store_epilogue(ep, lr.row, (int64_t)lr.row * d4 + i, s0);
        } else {
            for (int i = lane; i < d4; i += 32) {
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int c = 0; c < lr.n_chunks; ++c) add(sum, partial4[(int64_t)(lr.chunk_base + c) * d4 + i]);
                store_epilogue(ep, lr.row, (int64_t)lr.row * d4 + i, sum);
            }
        }
    }
}

struct SpmmArgs {
    const int64_t* rowptr; const int32_t* col; const float* val; int64_t n_rows; int64_t nnz; const float4* X4; int d4;
    Epi ep;
    PlanHeader* hdr; LongRow* long_rows; Chunk* chunks; float4* partial4;
    cudaStream_t stream;
    int long_row;
    const int32_t* warp_rows;      // work-balanced row ranges of the plan (nullptr: fixed rows per warp)
};

template <int LPR, int NV, bool BOUNDS, int GLPR = 0, int GNV = 0, int GU = 4, int GMINB = 3, int SLPR = GLPR, int SNV = GNV>
int launch_spmm(const SpmmArgs& a) {
    const int long_row = a.hdr ? a.long_row : 0x7fffffff;
    bool chunks_fused = false;
    const int64_t blocks = (a.n_rows * 32 + kThreads - 1) / kThreads;
    if (blocks > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    if constexpr (GLPR > 0) {
        if (a.n_rows > 0) {
            // 128 rows per warp amortise the row-pointer fetches on big graphs; dataset-sized graphs (10^4 .. 10^5 rows) get 32
            // rows per warp so that the grid still covers the SMs (CiteULike-shaped, 22.5k rows: 22 -> 88 CTAs)
            const bool big = a.n_rows >= (int64_t)148 * 24 * 128 || getenv("CR_SPMM_FORCE_BIG");   // (test knob: wide geometry on small graphs)
            const int rows_per_warp = big ? 128 : 32;
            // Work-balanced ranges pay where a fixed 128 rows per warp leaves few, uneven warps: the row block of an 8-way
            // partition (2.3 waves of CTAs; r02, 8 GPUs: 11.3 -> 8.8 ms per propagation).  With 4.5 waves (4-way partition) the
            // fixed mapping is ahead again (10.4 vs 10.9 ms, same box, `profiles/r02_prop_probe_n4.jsonl`), and on the whole
            // 11M-row graph (18 waves) by ~1 % (29.9 vs 30.2 ms): balanced below 4 waves of fixed warps only.
            const bool few_warps = a.n_rows < (int64_t)148 * 4 * 8 * 4 * 128;
            const char* force = getenv("CR_SPMM_FIXED_ROWS");          // (A/B knob: 1 = always fixed, 0 = always balanced)
            const bool balanced = big && (force ? force[0] == '0' : few_warps);
            const int32_t* warp_rows = balanced ? a.warp_rows : nullptr;
            const int64_t n_bal = balanced_warps(a.n_rows, a.nnz);
            const int64_t warps = warp_rows ? n_bal : (a.n_rows + rows_per_warp - 1) / rows_per_warp;
            // long-row chunks ride in the first CTAs of the same launch (the plan lives on the device: size by its upper bound)
            const int64_t max_chunks = a.hdr ? a.nnz / chunk_of(a.nnz) + a.nnz / (long_row_of(a.nnz) + 1) + 2 : 0;
            const int chunk_ctas = (int)min((int64_t)148 * GMINB, (max_chunks * 32 + kThreads - 1) / kThreads);
            chunks_fused = a.hdr != nullptr;
            const unsigned gblocks = (unsigned)((warps * 32 + kThreads - 1) / kThreads) + (unsigned)chunk_ctas;
            cr::prof_start(cr::PROF_SPMM_ROWS, a.stream);
            // Bandwidth-bound graphs take the (GLPR, GU, GMINB) geometry; dataset-sized ones are latency bound and keep more rows
            // in flight per warp with the narrower groups (SLPR lanes x SNV float4, 4 gathers, 3 CTAs/SM): CiteULike-shaped
            // training step 0.37 ms vs 0.43 ms with the wide geometry.
            // multi-GPU: rows leave through shared-memory staging + TMA bulk stores (d <= 64: 32 KB of staging per CTA); the
            // SM-issued stores remain for the NVLS multicast variant, wider rows and as an A/B knob (CR_SPMM_PEER_ST=1)
            const bool staged = a.ep.peers && !a.ep.mc4 && GLPR * GNV <= 16 && !getenv("CR_SPMM_PEER_ST");
#define CR_GROUPED_(L_, N_, V_, U_, B_, S_)                                                                                     \
    spmm_rows_grouped_kernel<L_, N_, V_, U_, B_, S_><<<gblocks, kThreads, 0, a.stream>>>(                                       \
        a.rowptr, a.col, a.val, a.n_rows, a.X4, a.ep, long_row, rows_per_warp, a.hdr, a.chunks, a.partial4, chunk_ctas,     \
        warp_rows, n_bal)
#define CR_GROUPED(L_, N_, V_, U_, B_)                                                                                          \
    do {                                                                                                                        \
        if constexpr ((L_) * (N_) <= 16) { if (staged) CR_GROUPED_(L_, N_, V_, U_, B_, true); else CR_GROUPED_(L_, N_, V_, U_, B_, false); } \
        else CR_GROUPED_(L_, N_, V_, U_, B_, false);                                                                            \
    } while (0)
            if (big) { if (a.val) CR_GROUPED(GLPR, GNV, true, GU, GMINB); else CR_GROUPED(GLPR, GNV, false, GU, GMINB); }
            else { if (a.val) CR_GROUPED(SLPR, SNV, true, 4, 3); else CR_GROUPED(SLPR, SNV, false, 4, 3); }
#undef CR_GROUPED
#undef CR_GROUPED_
            CR_LAUNCH_CHECK("spmm_rows_grouped_kernel");
            cr::prof_stop(cr::PROF_SPMM_ROWS, a.stream);
        }
    } else if (a.n_rows > 0) {
        cr::prof_start(cr::PROF_SPMM_ROWS, a.stream);
        if (a.val)
            spmm_rows_kernel<LPR, NV, true, BOUNDS><<<(unsigned)blocks, kThreads, 0, a.stream>>>(
                a.rowptr, a.col, a.val, a.n_rows, a.X4, a.d4, a.ep, long_row);
        else
            spmm_rows_kernel<LPR, NV, false, BOUNDS><<<(unsigned)blocks, kThreads, 0, a.stream>>>(
                a.rowptr, a.col, a.val, a.n_rows, a.X4, a.d4, a.ep, long_row);
        CR_LAUNCH_CHECK("spmm_rows_kernel");
        cr::prof_stop(cr::PROF_SPMM_ROWS, a.stream);
    }
    if (a.hdr) {
        const int grid = 148 * 8;
        if (!chunks_fused) {
            if (a.val)
                spmm_long_chunks_kernel<LPR, NV, true, BOUNDS><<<grid, kThreads, 0, a.stream>>>(a.hdr, a.chunks, a.col, a.val,
                                                                                                a.X4, a.d4, a.partial4);
            else
                spmm_long_chunks_kernel<LPR, NV, false, BOUNDS><<<grid, kThreads, 0, a.stream>>>(a.hdr, a.chunks, a.col, a.val,
                                                                                                 a.X4, a.d4, a.partial4);
            CR_LAUNCH_CHECK("spmm_long_chunks_kernel");
        }
        spmm_long_reduce_kernel<<<148, kThreads, 0, a.stream>>>(a.hdr, a.long_rows, a.partial4, a.d4, a.ep);
        CR_LAUNCH_CHECK("spmm_long_reduce_kernel");
    }
    return CR_OK;
}

}  // namespace

extern "C" {

size_t cr_spmm_plan_bytes(int64_t n_rows, int64_t nnz, int d) {
    if (nnz < 0 || d <= 0 || n_rows < 0) return 0;
    return plan_layout(nnz, d).total + warp_table_bytes(n_rows, nnz);
}

int cr_spmm_plan(const int64_t* rowptr, int64_t n_rows, int64_t nnz, int d, void* plan, size_t plan_bytes, void* stream) {
    if (!rowptr || !plan || n_rows < 0 || nnz < 0 || d <= 0) return CR_ERR_ARG;
    if (n_rows > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    const PlanLayout L = plan_layout(nnz, d);
    if (plan_bytes < L.total + warp_table_bytes(n_rows, nnz)) return CR_ERR_WORKSPACE;
    if (L.max_chunks > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)plan;
    CR_CUDA_TRY(cudaMemsetAsync(base, 0, 256, st));
    if (n_rows > 0) {
        const unsigned blocks = (unsigned)((n_rows + 255) / 256);
        spmm_plan_kernel<<<blocks, 256, 0, st>>>(rowptr, n_rows, (PlanHeader*)base, (LongRow*)(base + L.off_long),
                                                 (Chunk*)(base + L.off_chunks), (int)L.max_long, (int)L.max_chunks, long_row_of(nnz),
                                                 chunk_of(nnz));
        CR_LAUNCH_CHECK("spmm_plan_kernel");
        const int64_t n_warps = balanced_warps(n_rows, nnz);
        spmm_warp_rows_kernel<<<(unsigned)((n_warps + 1 + 255) / 256), 256, 0, st>>>(rowptr, n_rows, n_warps, (int32_t*)(base + L.off_warp_rows));
        CR_LAUNCH_CHECK("spmm_warp_rows_kernel");
    }
    return CR_OK;
}

static int spmm_entry(const int64_t* rowptr, const int32_t* col, const float* val, int64_t n_rows, int64_t nnz, const float* X,
                      int d, float* Y, const float* acc_in, float* acc, float acc_beta, float acc_div, void* plan,
                      size_t plan_bytes, float* const* peers, int n_peers, int64_t peer_row_offset, int64_t peer_row_split,
                      int64_t peer_row_offset_hi, int bcast_acc, const uint8_t* peer_need, float* mc_table, void* stream) {
    if (!rowptr || (!col && nnz > 0) || !X || n_rows < 0 || nnz < 0 || (!Y && !acc && !peers) || acc_div == 0.f) return CR_ERR_ARG;
    if (d <= 0 || d % 4 != 0 || d > 512) return CR_ERR_UNSUPPORTED;
    if (!cr::aligned16(X) || !cr::aligned16(Y) || !cr::aligned16(acc) || !cr::aligned16(acc_in)) return CR_ERR_ALIGN;
    if (acc_in && !acc) return CR_ERR_ARG;
    if ((peer_need && n_peers > 8) || (peers && n_peers > 32)) return CR_ERR_UNSUPPORTED;
    if (!cr::aligned16(mc_table)) return CR_ERR_ALIGN;
    if (peers && (n_peers < 1 || peer_row_offset < 0 || peer_row_split < 0 || peer_row_split + peer_row_offset_hi < 0)) return CR_ERR_ARG;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    SpmmArgs a{rowptr, col, val, n_rows, nnz, (const float4*)X, d / 4,
               Epi{(float4*)Y, (const float4*)(acc_in ? acc_in : acc), (float4*)acc, acc_beta, acc_div, (float4* const*)peers,
                   peers ? n_peers : 0, peer_row_offset * (d / 4), peer_row_split * (d / 4), peer_row_offset_hi * (d / 4), bcast_acc, peers ? peer_need : nullptr,
                   peers ? (float4*)mc_table : nullptr, (peers && n_peers < 32) ? (1u << n_peers) - 1u : 0xffffffffu},
               nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream, long_row_of(nnz), nullptr};
    if (plan) {
        const PlanLayout L = plan_layout(nnz, d);
        // the plan may have been built for more rows than this call covers (empty padding rows at the end of a partition)
        if (plan_bytes < L.total + warp_table_bytes(n_rows, nnz)) return CR_ERR_WORKSPACE;
        a.warp_rows = (const int32_t*)((char*)plan + L.off_warp_rows);
        char* base = (char*)plan;
        a.hdr = (PlanHeader*)base;
        a.long_rows = (LongRow*)(base + L.off_long);
        a.chunks = (Chunk*)(base + L.off_chunks);
        a.partial4 = (float4*)(base + L.off_partial);
    }
    switch (d) {
        case 32: return launch_spmm<8, 1, false, 8, 1>(a);
        case 64: return launch_spmm<16, 1, false, CR_SPMM_GLPR, 16 / CR_SPMM_GLPR, CR_SPMM_U, CR_SPMM_MINB, 8, 2>(a);
        case 128: return launch_spmm<32, 1, false, 16, 2>(a);
        case 256: return launch_spmm<32, 2, false>(a);
        default:
            if (d <= 128) return launch_spmm<32, 1, true>(a);
            if (d <= 256) return launch_spmm<32, 2, true>(a);
            return launch_spmm<32, 4, true>(a);
    }
}

int cr_spmm_csr_f32(const int64_t* rowptr, const int32_t* col, const float* val, int64_t n_rows, int64_t nnz,
                    const float* X, int d, float* Y, const float* acc_in, float* acc, float acc_beta, float acc_div,
                    void* plan, size_t plan_bytes, void* stream) {
    return spmm_entry(rowptr, col, val, n_rows, nnz, X, d, Y, acc_in, acc, acc_beta, acc_div, plan, plan_bytes, nullptr, 0, 0, 0, 0, 0, nullptr, nullptr, stream);
}

int cr_spmm_csr_bcast_f32(const int64_t* rowptr, const int32_t* col, const float* val, int64_t n_rows, int64_t nnz,
                          const float* X, int d, float* const* peer_tables, int n_peers, int64_t peer_row_offset,
                          int64_t peer_row_split, int64_t peer_row_offset_hi, int bcast_acc, const uint8_t* peer_need,
                          float* mc_table, const float* acc_in, float* acc, float acc_beta, float acc_div, void* plan,
                          size_t plan_bytes, void* stream) {
    if (!peer_tables || (bcast_acc && !acc)) return CR_ERR_ARG;
    return spmm_entry(rowptr, col, val, n_rows, nnz, X, d, nullptr, acc_in, acc, acc_beta, acc_div, plan, plan_bytes, peer_tables,
                      n_peers, peer_row_offset, peer_row_split, peer_row_offset_hi, bcast_acc, peer_need, mc_table, stream);
}

}  // extern "C"
