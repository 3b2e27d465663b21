// Row-wise top-K of a dense score block — the selection half of the content-kNN path for wide tables.
//
// Replaces the `index.search(query, k)` of model/KNN.py:63-77 (faiss IndexFlatIP) and the argpartition of FSGNN's chunked
// cosine kNN (model/FSGNN.py:106-152) when the content width (300 / 2,738) does not fit the fused sweep's TMEM-resident
// query tiles: the inner products of a block of queries against all values are produced by the tensor-core layer kernel
// (cr_linear_act_tc_f32 with the value table as the weight matrix: S = Q . V^T at fp32 accuracy, tower_tc.cu) and this kernel
// keeps the K best of every row.  One warp per row: the row is streamed 128 columns at a time, a value is looked at again only
// if it beats the row's current K-th best, survivors go to a shared-memory buffer that the warp rank-compacts when it could
// overflow.  Order: score descending, then column id ascending (the order of every top-K structure in this library).
#include "common.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kCapExtra = 128;      // one step can add at most 128 candidates (4 per lane)

struct Cand {
    float s;
    int id;
};

__global__ void __launch_bounds__(kWarps * 32)
topk_rows_kernel(const float* __restrict__ S, int64_t n_rows, int n_cols, int64_t ld, int K, const int32_t* __restrict__ exclude_col,
                 int col_id_base, float* __restrict__ out_score, int32_t* __restrict__ out_id) {
    extern __shared__ unsigned char smem_raw[];
    const int cap = K + kCapExtra;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Cand* buf = reinterpret_cast<Cand*>(smem_raw) + (size_t)w * 2 * cap;      // [cap] candidates + [cap] scratch for the compaction
    Cand* tmp = buf + cap;
    const int64_t row = (int64_t)blockIdx.x * kWarps + w;
    if (row >= n_rows) return;
    const float* s = S + row * ld;
    const int excl = exclude_col ? __ldg(exclude_col + row) : -1;
    int cnt = 0;
    float thr = -CUDART_INF_F;
    auto compact = [&]() {
        __syncwarp();
        for (int e = lane; e < cnt; e += 32) {
            const Cand c = buf[e];
            int rank = 0;
            for (int j = 0; j < cnt; ++j) rank += cr::better(buf[j].s, buf[j].id, c.s, c.id) ? 1 : 0;
            if (rank < K) tmp[rank] = c;
        }
        __syncwarp();
        const int m = min(cnt, K);
        for (int e = lane; e < m; e += 32) buf[e] = tmp[e];
        __syncwarp();
        cnt = m;
        if (m == K) thr = buf[K - 1].s;
    };
    const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(S) & 15u) == 0);
    for (int c0 = 0; c0 < n_cols; c0 += 128) {
        float v[4];
        const int c = c0 + lane * 4;
        if (vec && c + 4 <= n_cols) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(s + c));
            v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        } else {
#pragma unroll
            for (int y = 0; y < 4; ++y) v[y] = (c + y < n_cols) ? __ldg(s + c + y) : -CUDART_INF_F;
        }
        if (cnt + 128 > cap) compact();
        // columns are visited in ascending id order, so "strictly greater than the K-th" implements (score desc, id asc)
        int mine = 0;
#pragma unroll
        for (int y = 0; y < 4; ++y) mine += (c + y < n_cols && c + y != excl && v[y] > thr) ? 1 : 0;
        int pre = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(CR_FULL_MASK, pre, off);
            if (lane >= off) pre += t;
        }
        const int total = __shfl_sync(CR_FULL_MASK, pre, 31);
        int slot = cnt + pre - mine;
#pragma unroll
        for (int y = 0; y < 4; ++y)
            if (c + y < n_cols && c + y != excl && v[y] > thr) buf[slot++] = Cand{v[y], col_id_base + c + y};
        cnt += total;
        __syncwarp();
    }
    compact();
    for (int e = lane; e < K; e += 32) {
        out_score[row * K + e] = e < cnt ? buf[e].s : -CUDART_INF_F;
        out_id[row * K + e] = e < cnt ? buf[e].id : -1;
    }
}

}  // namespace

extern "C" int cr_topk_rows_f32(const float* S, int64_t n_rows, int n_cols, int64_t ld, int K, const int32_t* exclude_col, int col_id_base,
                                float* out_score, int32_t* out_id, void* stream) {
    if (!S || !out_score || !out_id || n_rows < 0 || n_cols <= 0 || ld < n_cols || K <= 0) return CR_ERR_ARG;
    if (K > 1024) return CR_ERR_UNSUPPORTED;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    if (n_rows == 0) return CR_OK;
    const size_t smem = (size_t)kWarps * 2 * (K + kCapExtra) * sizeof(Cand);
    if (smem > 48 * 1024) CR_CUDA_TRY(cudaFuncSetAttribute(topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = (n_rows + kWarps - 1) / kWarps;
    if (blocks > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    topk_rows_kernel<<<(unsigned)blocks, kWarps * 32, smem, (cudaStream_t)stream>>>(S, n_rows, n_cols, ld, K, exclude_col, col_id_base,
                                                                                  out_score, out_id);
    CR_LAUNCH_CHECK("topk_rows_kernel");
    return CR_OK;
}
