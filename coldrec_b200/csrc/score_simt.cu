// K1 (exact path) — fused fp32 scorer: U.I^T tile -> train/flag mask -> threshold top-K, plus the
// candidate-list merge and the row gather.  No score matrix is ever written to HBM.
//
// Replaces MF.batch_predict (model/MF.py:58-63) + the mask writes and torch.topk of
// BaseColdStartTrainer._evaluate (model/BaseRecommender.py:170-182).  This is the CR_SCORE_EXACT_F32
// precision: plain FFMA, k = 0..d-1 in order, used on its own and as the refinement path of the
// tcgen05 scorer (score_tc.cu).
//
// One CTA owns kTU=32 queries and sweeps an item range in tiles of kTI=128 items.  Each thread
// computes a 4-query x 4-item micro-tile.  Selection is a threshold filter: a score is looked at
// again only if it beats the query's current K-th best (kept in shared memory); survivors are mask-
// checked (binary search in the query's sorted train row, item flag byte) and appended to a per-query
// shared-memory buffer that a warp compacts by ranking when it could overflow.  Items are visited in
// ascending global-id order, so "strictly greater than the K-th" implements (score desc, id asc).
#include "common.cuh"

namespace {

constexpr int kTU = 32;              // queries per CTA
constexpr int kTI = 128;             // items per tile
constexpr int kKC = 32;              // k-chunk (floats) staged per step
constexpr int kIS = kKC + 4;         // padded item-tile row stride (floats): conflict-free float4 reads
constexpr int kCap = CR_MAX_K + kTI; // candidate buffer entries per query
constexpr int kThreads = 256;
constexpr int kMaxSplits = 32;
constexpr int kMaxD = 256;            // widest table whose query vectors stay resident in shared memory; wider ones are
                                     // streamed k-chunk by k-chunk next to the item tile (content kNN: d = 300 .. 2740)

struct Cand {
    float s;
    int id;
};

struct ScoreParams {
    const float* user_tab; const int32_t* user_ids; int64_t n_q;
    const float* item_tab; const int32_t* item_gids; int64_t item_id_base; int64_t n_items; int d;
    const int64_t* mask_rowptr; const int32_t* mask_col; const uint8_t* item_flags; uint8_t flag_exclude;
    const int32_t* q_list;        // optional: sweep only these query indices (refinement path)
    const int32_t* q_count;       // device count for q_list (nullable -> n_q)
    int64_t gate_lo, gate_hi;     // with q_count: run only if gate_lo < count <= gate_hi
    int K; int n_splits; int64_t items_per_split;
    int64_t part_stride; int compact;   // partial lists indexed by list position instead of query
    float* out_score; int32_t* out_id;      // [n_splits, n_q, K] (or the final output when n_splits == 1)
};

// Rank-compact one query's buffer to its best K entries, sorted.  Called by one full warp.
__device__ void compact_query(Cand* buf, int* cnt, float* thr, Cand* tmp, int K, int lane) {
    const int n = *cnt;
    __syncwarp();
    for (int e = lane; e < n; e += 32) {
        const Cand c = buf[e];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const Cand o = buf[j];
            rank += cr::better(o.s, o.id, c.s, c.id) ? 1 : 0;
        }
        if (rank < K) tmp[rank] = c;
    }
    __syncwarp();
    const int m = min(n, K);
    for (int e = lane; e < m; e += 32) buf[e] = tmp[e];
    if (lane == 0) {
        *cnt = m;
        if (m == K) *thr = tmp[K - 1].s;
    }
    __syncwarp();
}

template <bool RESIDENT>
__global__ void __launch_bounds__(kThreads) score_topk_exact_kernel(const ScoreParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = p.d;
    const int ustride = RESIDENT ? d : kKC;                          // floats per query row held in shared memory
    float* s_user = reinterpret_cast<float*>(smem_raw);              // [kTU][d] (resident) or [kTU][kKC] (streamed)
    float* s_item = s_user + kTU * ustride;                          // [kTI][kIS]
    Cand* s_buf = reinterpret_cast<Cand*>(s_item + kTI * kIS);       // [kTU][kCap]
    Cand* s_tmp = s_buf + kTU * kCap;                                // [8 warps][CR_MAX_K]
    float* s_thr = reinterpret_cast<float*>(s_tmp + 8 * CR_MAX_K);   // [kTU]
    int* s_cnt = reinterpret_cast<int*>(s_thr + kTU);                // [kTU]
    int* s_q = s_cnt + kTU;                                          // [kTU] query index or -1

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_active = p.q_count ? (int64_t)*p.q_count : p.n_q;
    const int64_t q0 = (int64_t)blockIdx.x * kTU;
    if (q0 >= n_active) return;
    if (p.q_count && (n_active <= p.gate_lo || n_active > p.gate_hi)) return;
    const int split = blockIdx.y;
    const int64_t it_begin = (int64_t)split * p.items_per_split;
    const int64_t it_end = min(p.n_items, it_begin + p.items_per_split);
    const int K = p.K;

    if (tid < kTU) {
        const int64_t qi = q0 + tid;
        int q = -1;
        if (qi < n_active) q = p.q_list ? p.q_list[qi] : (int)qi;
        s_q[tid] = q;
        s_thr[tid] = -CUDART_INF_F;
        s_cnt[tid] = 0;
    }
    __syncthreads();
    // stage the query vectors (gathered through user_ids)
    for (int i = tid; RESIDENT && i < kTU * (d / 4); i += kThreads) {
        const int u = i / (d / 4), c = i % (d / 4);
        const int q = s_q[u];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q >= 0) {
            const int64_t row = p.user_ids ? (int64_t)p.user_ids[q] : (int64_t)q;
            v = __ldg(reinterpret_cast<const float4*>(p.user_tab + row * d) + c);
        }
        reinterpret_cast<float4*>(s_user + u * d)[c] = v;
    }

    const int ug = warp * 4;   // this thread's 4 queries: ug..ug+3 ; its items: lane + 32*i
    for (int64_t tile = it_begin; tile < it_end; tile += kTI) {
        float acc[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[u][i] = 0.f;

        for (int k0 = 0; k0 < d; k0 += kKC) {
            __syncthreads();   // previous chunk consumed / candidate phase finished
            const int kc4 = min(kKC, d - k0) / 4;
            for (int i = tid; i < kTI * (kKC / 4); i += kThreads) {
                const int r = i / (kKC / 4), c = i % (kKC / 4);
                const int64_t it = tile + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (it < it_end && c < kc4) v = __ldg(reinterpret_cast<const float4*>(p.item_tab + it * d + k0) + c);
                reinterpret_cast<float4*>(s_item + r * kIS)[c] = v;
            }
            if (!RESIDENT) {        // this k-chunk of the 32 query vectors (re-read per item tile: L2 resident, 4 KB)
                for (int i = tid; i < kTU * (kKC / 4); i += kThreads) {
                    const int u = i / (kKC / 4), c = i % (kKC / 4);
                    const int q = s_q[u];
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (q >= 0 && c < kc4) {
                        const int64_t row = p.user_ids ? (int64_t)p.user_ids[q] : (int64_t)q;
                        v = __ldg(reinterpret_cast<const float4*>(p.user_tab + row * d + k0) + c);
                    }
                    reinterpret_cast<float4*>(s_user + u * kKC)[c] = v;
                }
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < kKC; kk += 4) {
                if (k0 + kk < d) {
                    float4 uu[4], ii[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) uu[u] = *reinterpret_cast<const float4*>(s_user + (ug + u) * ustride + (RESIDENT ? k0 : 0) + kk);
#pragma unroll
                    for (int i = 0; i < 4; ++i) ii[i] = *reinterpret_cast<const float4*>(s_item + (lane + 32 * i) * kIS + kk);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            acc[u][i] = fmaf(uu[u].x, ii[i].x, acc[u][i]);
                            acc[u][i] = fmaf(uu[u].y, ii[i].y, acc[u][i]);
                            acc[u][i] = fmaf(uu[u].z, ii[i].z, acc[u][i]);
                            acc[u][i] = fmaf(uu[u].w, ii[i].w, acc[u][i]);
                        }
                }
            }
        }

        // threshold filter + mask check + append
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = s_q[ug + u];
            if (q < 0) continue;
            const float thr = s_thr[ug + u];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t it = tile + lane + 32 * i;
                float s = acc[u][i];
                if (it < it_end && s > thr) {
                    const int gid = p.item_gids ? __ldg(p.item_gids + it) : (int)(p.item_id_base + it);
                    bool masked = p.item_flags && (__ldg(p.item_flags + gid) & p.flag_exclude);
                    if (!masked && p.mask_rowptr)
                        masked = cr::csr_row_contains(p.mask_col, __ldg(p.mask_rowptr + q), __ldg(p.mask_rowptr + q + 1), gid);
                    if (masked) s = CR_MASK_SCORE;
                    if (s > thr) {
                        const int slot = atomicAdd(&s_cnt[ug + u], 1);
                        s_buf[(ug + u) * kCap + slot] = Cand{s, gid};
                    }
                }
            }
        }
        __syncthreads();
        // a buffer that could overflow on the next tile is compacted by this query group's warp
#pragma unroll 1
        for (int u = 0; u < 4; ++u)
            if (s_cnt[ug + u] > kCap - kTI)
                compact_query(s_buf + (ug + u) * kCap, &s_cnt[ug + u], &s_thr[ug + u], s_tmp + warp * CR_MAX_K, K, lane);
    }
    __syncthreads();
#pragma unroll 1
    for (int u = 0; u < 4; ++u) {
        const int q = s_q[ug + u];
        if (q < 0) continue;
        compact_query(s_buf + (ug + u) * kCap, &s_cnt[ug + u], &s_thr[ug + u], s_tmp + warp * CR_MAX_K, K, lane);
        const int m = s_cnt[ug + u];
        const int64_t o = ((int64_t)split * p.part_stride + (p.compact ? q0 + ug + u : (int64_t)q)) * K;
        for (int k = lane; k < K; k += 32) {
            const Cand c = (k < m) ? s_buf[(ug + u) * kCap + k] : Cand{-CUDART_INF_F, -1};
            p.out_score[o + k] = c.s;
            p.out_id[o + k] = c.id;
        }
    }
}

size_t exact_smem_bytes(int d) {
    return (size_t)kTU * (d <= kMaxD ? d : kKC) * 4 + (size_t)kTI * kIS * 4 + (size_t)kTU * kCap * sizeof(Cand) + 8 * CR_MAX_K * sizeof(Cand) +
           kTU * 4 * 3;
}

// One CTA per query: rank every valid entry of the G sorted lists by binary search in the others.
__global__ void __launch_bounds__(128) topk_merge_kernel(const float* __restrict__ in_score, const int32_t* __restrict__ in_id,
                                                         int G, int64_t n_q, int K, const int32_t* q_list,
                                                         const int32_t* q_count, int64_t part_stride, int compact,
                                                         int64_t gate_lo, int64_t gate_hi, float* __restrict__ out_score,
                                                         int32_t* __restrict__ out_id) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cand* s_in = reinterpret_cast<Cand*>(smem_raw);   // [G][K]
    Cand* s_out = s_in + G * K;                       // [K]
    const int64_t n_active = q_count ? (int64_t)*q_count : n_q;
    if (q_count && (n_active <= gate_lo || n_active > gate_hi)) return;
    for (int64_t qi = blockIdx.x; qi < n_active; qi += gridDim.x) {
        const int64_t q = q_list ? (int64_t)q_list[qi] : qi;
        const int n = G * K;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int g = e / K, k = e % K;
            const int64_t src = ((int64_t)g * part_stride + (compact ? qi : q)) * K + k;
            s_in[e] = Cand{in_score[src], in_id[src]};
        }
        for (int k = threadIdx.x; k < K; k += blockDim.x) s_out[k] = Cand{-CUDART_INF_F, -1};
        __syncthreads();
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const Cand c = s_in[e];
            if (c.id < 0) continue;
            int rank = 0;
            for (int g = 0; g < G; ++g) {
                const Cand* L = s_in + g * K;
                int lo = 0, hi = K;   // number of valid entries of list g that beat c
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const Cand o = L[mid];
                    if (o.id >= 0 && cr::better(o.s, o.id, c.s, c.id)) lo = mid + 1; else hi = mid;
                }
                rank += lo;
            }
            if (rank < K) s_out[rank] = c;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            out_score[q * K + k] = s_out[k].s;
            out_id[q * K + k] = s_out[k].id;
        }
        __syncthreads();
    }
}

__global__ void gather_rows_kernel(const float4* __restrict__ src, const int32_t* __restrict__ ids, int64_t n, int d4,
                                   float4* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * d4) return;
    const int64_t r = i / d4;
    const int c = (int)(i % d4);
    dst[i] = __ldg(src + (int64_t)__ldg(ids + r) * d4 + c);
}

// dst[dst_ids ? dst_ids[j] : j, :] = src[src_ids ? src_ids[j] : j, :]
__global__ void copy_rows_kernel(const float4* __restrict__ src, const int32_t* __restrict__ src_ids,
                                 const int32_t* __restrict__ dst_ids, int64_t n, int d4, float4* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * d4) return;
    const int64_t r = i / d4;
    const int c = (int)(i % d4);
    const int64_t sr = src_ids ? (int64_t)__ldg(src_ids + r) : r, dr = dst_ids ? (int64_t)__ldg(dst_ids + r) : r;
    dst[dr * d4 + c] = __ldg(src + sr * d4 + c);
}

// Queries that found fewer than K candidates in the swept (flag-compacted) tables: the reference would
// list masked ids at CR_MASK_SCORE there (any of them: they tie), so fill with the smallest masked ids.
__global__ void fill_masked_kernel(float* __restrict__ out_score, int32_t* __restrict__ out_id, int64_t n_q, int K,
                                   int64_t n_items_total, const uint8_t* __restrict__ item_flags, uint8_t flag_exclude,
                                   const int64_t* __restrict__ mask_rowptr, const int32_t* __restrict__ mask_col) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_q) return;
    float* os = out_score + q * K;
    int32_t* oi = out_id + q * K;
    int m = K;
    while (m > 0 && oi[m - 1] < 0) --m;
    if (m == K) return;
    const int64_t lo = mask_rowptr ? mask_rowptr[q] : 0, hi = mask_rowptr ? mask_rowptr[q + 1] : 0;
    for (int64_t gid = 0; gid < n_items_total && m < K; ++gid) {
        bool masked = item_flags && (item_flags[gid] & flag_exclude);
        if (!masked && mask_rowptr) masked = cr::csr_row_contains(mask_col, lo, hi, (int32_t)gid);
        if (!masked) continue;
        bool present = false;
        for (int k = 0; k < m; ++k) present |= (oi[k] == (int32_t)gid);
        if (present) continue;
        os[m] = CR_MASK_SCORE;
        oi[m] = (int32_t)gid;
        ++m;
    }
}

}  // namespace

namespace cr {

int exact_splits(int64_t n_q, int64_t n_items) {
    const int64_t ctas = (n_q + kTU - 1) / kTU;
    int64_t s = (2 * 148 + ctas - 1) / (ctas > 0 ? ctas : 1);
    const int64_t max_by_items = (n_items + 4 * kTI - 1) / (4 * kTI);
    if (s > max_by_items) s = max_by_items;
    if (s > kMaxSplits) s = kMaxSplits;
    if (s < 1) s = 1;
    return (int)s;
}

size_t exact_workspace_bytes(int64_t n_q, int64_t n_items, int K) {
    const int S = exact_splits(n_q, n_items);
    if (S == 1) return 0;
    return align_up((size_t)S * n_q * K * 4, 256) * 2;
}

int launch_merge(const float* in_score, const int32_t* in_id, int G, int64_t n_q, int K, float* out_score, int32_t* out_id,
                 cudaStream_t st, const RefineList* rl, int64_t part_stride, int compact, int64_t gate_lo, int64_t gate_hi) {
    if (n_q == 0) return CR_OK;
    const size_t smem = (size_t)(G + 1) * K * sizeof(Cand);
    if (smem > 200 * 1024) return CR_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        CR_CUDA_TRY(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n_grid = rl ? rl->cap : n_q;
    const unsigned grid = (unsigned)(n_grid < 148 * 64 ? n_grid : 148 * 64);
    topk_merge_kernel<<<grid, 128, smem, st>>>(in_score, in_id, G, n_q, K, rl ? rl->list : nullptr, rl ? rl->count : nullptr,
                                               part_stride > 0 ? part_stride : n_q, compact, gate_lo, gate_hi, out_score, out_id);
    CR_LAUNCH_CHECK("topk_merge_kernel");
    return CR_OK;
}

static int launch_exact_impl(const ExactJob& j, const RefineList* rl, int S, int64_t n_sweep, int compact, int64_t gate_lo,
                             int64_t gate_hi, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (n_sweep == 0) return CR_OK;
    ScoreParams p{j.user_tab, j.user_ids, j.n_q, j.item_tab, j.item_gids, j.item_id_base, j.n_items, j.d, j.mask_rowptr,
                  j.mask_col, j.item_flags, j.flag_exclude, rl ? rl->list : nullptr, rl ? rl->count : nullptr, gate_lo, gate_hi,
                  j.K, S, 0, j.n_q, 0, j.out_score, j.out_id};
    p.items_per_split = ((j.n_items + S - 1) / S + kTI - 1) / kTI * kTI;
    if (p.items_per_split == 0) p.items_per_split = kTI;
    float* part_s = nullptr;
    int32_t* part_i = nullptr;
    if (S > 1) {
        p.part_stride = compact ? n_sweep : j.n_q;
        p.compact = compact;
        const size_t half = align_up((size_t)S * p.part_stride * j.K * 4, 256);
        if (!ws || ws_bytes < 2 * half) return CR_ERR_WORKSPACE;
        part_s = (float*)ws;
        part_i = (int32_t*)((char*)ws + half);
        p.out_score = part_s;
        p.out_id = part_i;
    }
    const size_t smem = exact_smem_bytes(j.d);
    const int64_t gx = (n_sweep + kTU - 1) / kTU;
    if (gx > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    if (j.d <= kMaxD) {
        CR_CUDA_TRY(cudaFuncSetAttribute(score_topk_exact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        score_topk_exact_kernel<true><<<dim3((unsigned)gx, (unsigned)S), kThreads, smem, st>>>(p);
    } else {
        CR_CUDA_TRY(cudaFuncSetAttribute(score_topk_exact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        score_topk_exact_kernel<false><<<dim3((unsigned)gx, (unsigned)S), kThreads, smem, st>>>(p);
    }
    CR_LAUNCH_CHECK("score_topk_exact_kernel");
    if (S > 1) {
        RefineList sub{rl ? rl->list : nullptr, rl ? rl->count : nullptr, n_sweep};
        return launch_merge(part_s, part_i, S, j.n_q, j.K, j.out_score, j.out_id, st, rl ? &sub : nullptr, p.part_stride, compact,
                            gate_lo, gate_hi);
    }
    return CR_OK;
}

int launch_exact_scorer(const ExactJob& j, void* ws, size_t ws_bytes, cudaStream_t st) {
    return launch_exact_impl(j, nullptr, exact_splits(j.n_q, j.n_items), j.n_q, 0, -1, INT64_MAX, ws, ws_bytes, st);
}

// Refinement of the TF32 path: re-run the queries listed on the device (count unknown to the host).
// Two gated launches cover both regimes without a host sync: a few queries -> 16-way item split with
// compact partial lists (a lone CTA sweeping 10M items for one query would be a 40 ms tail); many queries -> one sweep per 32 queries writing the output rows directly.
constexpr int64_t kRefineSmall = 2048;
constexpr int kRefineSplits = 128;

size_t refine_workspace_bytes(int K) { return align_up((size_t)kRefineSplits * kRefineSmall * K * 4, 256) * 2; }

int launch_exact_refine(const ExactJob& j, const RefineList& rl, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (rl.cap == 0) return CR_OK;
    int S = kRefineSplits;
    const int64_t max_by_items = (j.n_items + 4 * kTI - 1) / (4 * kTI);
    if (S > max_by_items) S = (int)max_by_items;
    const int64_t small_cap = rl.cap < kRefineSmall ? rl.cap : kRefineSmall;
    int rc = launch_exact_impl(j, &rl, S, small_cap, 1, 0, kRefineSmall, ws, ws_bytes, st);
    if (rc != CR_OK || rl.cap <= kRefineSmall) return rc;
    return launch_exact_impl(j, &rl, 1, rl.cap, 0, kRefineSmall, INT64_MAX, nullptr, 0, st);
}

}  // namespace cr

extern "C" {

int cr_topk_merge(const float* in_score, const int32_t* in_id, int G, int64_t n_q, int K, float* out_score,
                  int32_t* out_id, void* stream) {
    if (!in_score || !in_id || !out_score || !out_id || G < 1 || n_q < 0) return CR_ERR_ARG;
    if (K < 1 || K > CR_MAX_K) return CR_ERR_UNSUPPORTED;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    return cr::launch_merge(in_score, in_id, G, n_q, K, out_score, out_id, (cudaStream_t)stream, nullptr, 0, 0, -1, INT64_MAX);
}

int cr_fill_masked(float* out_score, int32_t* out_id, int64_t n_q, int K, int64_t n_items_total, const uint8_t* item_flags,
                   uint8_t flag_exclude, const int64_t* mask_rowptr, const int32_t* mask_col, void* stream) {
    if (!out_score || !out_id || n_q < 0 || n_items_total < 0 || (mask_rowptr && !mask_col)) return CR_ERR_ARG;
    if (K < 1 || K > CR_MAX_K) return CR_ERR_UNSUPPORTED;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    if (n_q == 0) return CR_OK;
    fill_masked_kernel<<<(unsigned)((n_q + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        out_score, out_id, n_q, K, n_items_total, item_flags, flag_exclude, mask_rowptr, mask_col);
    CR_LAUNCH_CHECK("fill_masked_kernel");
    return CR_OK;
}

int cr_gather_rows_f32(const float* src, const int32_t* ids, int64_t n, int d, float* dst, void* stream) {
    if (!src || !ids || !dst || n < 0) return CR_ERR_ARG;
    if (d <= 0 || d % 4 != 0) return CR_ERR_UNSUPPORTED;
    if (!cr::aligned16(src) || !cr::aligned16(dst)) return CR_ERR_ALIGN;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    if (n == 0) return CR_OK;
    const int64_t total = n * (d / 4);
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    gather_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)src, ids, n, d / 4, (float4*)dst);
    CR_LAUNCH_CHECK("gather_rows_kernel");
    return CR_OK;
}

int cr_copy_rows_f32(const float* src, const int32_t* src_ids, const int32_t* dst_ids, int64_t n, int d, float* dst, void* stream) {
    if (!src || !dst || n < 0) return CR_ERR_ARG;
    if (d <= 0 || d % 4 != 0) return CR_ERR_UNSUPPORTED;
    if (!cr::aligned16(src) || !cr::aligned16(dst)) return CR_ERR_ALIGN;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    if (n == 0) return CR_OK;
    const int64_t total = n * (d / 4);
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    copy_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)src, src_ids, dst_ids, n, d / 4, (float4*)dst);
    CR_LAUNCH_CHECK("copy_rows_kernel");
    return CR_OK;
}

}  // extern "C"
