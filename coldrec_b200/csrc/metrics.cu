// K2 — Hit-ratio / Precision / Recall / NDCG partials reduced on device.
//
// Replaces ranking_evaluation + Metric.hits/hit_ratio/precision/recall/NDCG
// (util/evaluator.py:9-32, 47-63, 95-115, 153-187).  One thread per query walks its sorted top-K
// list, looks every id up in the query's ascending ground-truth row (binary search) and
// accumulates hits and DCG *in list order with the host-provided 1/log(n+2,2) table*, exactly the
// sequence of fp64 additions of evaluator.py:104-109, so per-query values are bit-equal to the
// reference.  Cross-query sums are reduced in a fixed order (block tree, then one ordered pass over
// the block partials), so results do not depend on scheduling.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxN = 8;
constexpr int kStats = 6;   // sum_hits, sum_gt, sum_recall, n_recall, sum_ndcg, n_ndcg

struct MetricParams {
    const int32_t* topk_id; int64_t n_q; int K;
    const int64_t* gt_rowptr; const int32_t* gt_col;
    int Ns[kMaxN]; int nN;
    const double* inv_log2; const double* idcg_prefix;
    int32_t* hits; double* dcg;
    double* block_partials;   // [gridDim.x][nN][kStats]
};

__global__ void __launch_bounds__(kThreads) rank_metrics_kernel(const MetricParams p) {
    __shared__ double s_red[kThreads];
    const int64_t q = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    double st[kMaxN][kStats];
#pragma unroll
    for (int a = 0; a < kMaxN; ++a)
#pragma unroll
        for (int b = 0; b < kStats; ++b) st[a][b] = 0.0;

    if (q < p.n_q) {
        const int64_t lo = p.gt_rowptr[q], hi = p.gt_rowptr[q + 1];
        const int n_gt = (int)(hi - lo);
        const int32_t* row = p.topk_id + q * p.K;
        for (int a = 0; a < p.nN; ++a) {
            // restart per cut-off: keeps the fp64 addition order identical to the reference for every N
            int hit_run = 0;
            double dcg_run = 0.0;
            const int N = p.Ns[a];
            for (int k = 0; k < N; ++k) {
                const int id = row[k];
                if (id >= 0 && cr::csr_row_contains(p.gt_col, lo, hi, id)) {
                    hit_run += 1;
                    dcg_run += p.inv_log2[k];
                }
            }
            if (p.hits) p.hits[(int64_t)a * p.n_q + q] = hit_run;
            if (p.dcg) p.dcg[(int64_t)a * p.n_q + q] = dcg_run;
            st[a][0] = (double)hit_run;
            st[a][1] = (double)n_gt;
            if (n_gt > 0) {
                st[a][2] = (double)hit_run / (double)n_gt;
                st[a][3] = 1.0;
            }
            const double idcg = p.idcg_prefix[min(n_gt, N)];
            if (idcg != 0.0) {
                st[a][4] = dcg_run / idcg;
                st[a][5] = 1.0;
            }
        }
    }
    for (int a = 0; a < p.nN; ++a)
        for (int b = 0; b < kStats; ++b) {
            s_red[threadIdx.x] = st[a][b];
            __syncthreads();
            for (int off = kThreads / 2; off > 0; off >>= 1) {
                if (threadIdx.x < off) s_red[threadIdx.x] += s_red[threadIdx.x + off];
                __syncthreads();
            }
            if (threadIdx.x == 0) p.block_partials[((int64_t)blockIdx.x * p.nN + a) * kStats + b] = s_red[0];
            __syncthreads();
        }
}

__global__ void rank_metrics_final_kernel(const double* __restrict__ block_partials, int n_blocks, int n_vals,
                                          double* __restrict__ sums) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vals) return;
    double s = 0.0;
    for (int b = 0; b < n_blocks; ++b) s += block_partials[(int64_t)b * n_vals + v];
    sums[v] = s;
}

}  // namespace

extern "C" {

size_t cr_rank_metrics_workspace_bytes(int64_t n_q, int nN) {
    const int64_t blocks = (n_q + kThreads - 1) / kThreads;
    return cr::align_up((size_t)(blocks > 0 ? blocks : 1) * (size_t)(nN > 0 ? nN : 1) * kStats * sizeof(double), 256);
}

int cr_rank_metrics(const int32_t* topk_id, int64_t n_q, int K, const int64_t* gt_rowptr, const int32_t* gt_col,
                    const int32_t* Ns_host, int nN, const double* inv_log2, const double* idcg_prefix, int32_t* hits,
                    double* dcg, double* sums, void* workspace, size_t ws_bytes, void* stream) {
    if (!topk_id || !gt_rowptr || !Ns_host || !inv_log2 || !idcg_prefix || !sums || n_q < 0) return CR_ERR_ARG;
    if (nN < 1 || nN > kMaxN || K < 1 || K > CR_MAX_K) return CR_ERR_UNSUPPORTED;
    for (int a = 0; a < nN; ++a)
        if (Ns_host[a] < 1 || Ns_host[a] > K) return CR_ERR_ARG;
    if (!workspace || ws_bytes < cr_rank_metrics_workspace_bytes(n_q, nN)) return CR_ERR_WORKSPACE;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    MetricParams p{};
    p.topk_id = topk_id; p.n_q = n_q; p.K = K; p.gt_rowptr = gt_rowptr; p.gt_col = gt_col; p.nN = nN;
    for (int a = 0; a < nN; ++a) p.Ns[a] = Ns_host[a];
    p.inv_log2 = inv_log2; p.idcg_prefix = idcg_prefix; p.hits = hits; p.dcg = dcg;
    p.block_partials = (double*)workspace;
    const int64_t blocks = (n_q + kThreads - 1) / kThreads;
    if (blocks > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    if (blocks == 0) {
        CR_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * nN * kStats, st));
        return CR_OK;
    }
    rank_metrics_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(p);
    CR_LAUNCH_CHECK("rank_metrics_kernel");
    rank_metrics_final_kernel<<<1, 64, 0, st>>>(p.block_partials, (int)blocks, nN * kStats, sums);
    CR_LAUNCH_CHECK("rank_metrics_final_kernel");
    return CR_OK;
}

}  // extern "C"
