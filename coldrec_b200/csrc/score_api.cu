// cr_score_topk_f32 — argument validation and dispatch between the exact fp32 scorer (score_simt.cu)
// and the tcgen05 TF32-checked scorer (score_tc.cu).
#include "common.cuh"

extern "C" {

size_t cr_score_topk_workspace_bytes(int64_t n_q, int64_t n_items, int d, int K, int precision) {
    if (n_q < 0 || n_items < 0 || d <= 0 || K < 1 || K > CR_MAX_K) return 0;
    if (precision == CR_SCORE_TF32_CHECKED) return cr::tc_workspace_bytes(n_q, n_items, d, K);
    return cr::exact_workspace_bytes(n_q, n_items, K);
}

int cr_score_topk_f32(const float* user_tab, const int32_t* user_ids, int64_t n_q, const float* item_tab,
                      const int32_t* item_gids, int64_t item_id_base, int64_t n_items, int d, const int64_t* mask_rowptr,
                      const int32_t* mask_col, const uint8_t* item_flags, uint8_t flag_exclude, int K, float* out_score,
                      int32_t* out_id, int32_t* n_refined, int precision, void* workspace, size_t ws_bytes, void* stream) {
    if (!user_tab || !item_tab || !out_score || !out_id || n_q < 0 || n_items < 0 || item_id_base < 0) return CR_ERR_ARG;
    if (mask_rowptr && !mask_col) return CR_ERR_ARG;
    if (K < 1 || K > CR_MAX_K || d <= 0 || d % 4 != 0) return CR_ERR_UNSUPPORTED;
    if (item_id_base + n_items > 0x7fffffffLL || n_q > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    if (precision != CR_SCORE_EXACT_F32 && precision != CR_SCORE_TF32_CHECKED) return CR_ERR_UNSUPPORTED;
    if (!cr::aligned16(user_tab) || !cr::aligned16(item_tab)) return CR_ERR_ALIGN;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const cr::ExactJob job{user_tab, user_ids, n_q, item_tab, item_gids, item_id_base, n_items, d, mask_rowptr, mask_col,
                           flag_exclude ? item_flags : nullptr, flag_exclude, K, out_score, out_id};
    if (flag_exclude && !item_flags) return CR_ERR_ARG;
    if (precision == CR_SCORE_TF32_CHECKED) return cr::launch_tc_scorer(job, n_refined, workspace, ws_bytes, st, nullptr);
    if (n_refined) CR_CUDA_TRY(cudaMemsetAsync(n_refined, 0, sizeof(int32_t), st));
    return cr::launch_exact_scorer(job, workspace, ws_bytes, st);
}

// Probe: raw TF32 scores of the first 256 queries x 96 items as the tensor-core sweep sees them
// (dbg [256*96] floats).  Test/diagnostic entry point; runs a full cr_score_topk_f32 underneath.
int cr_debug_tc_tile(const float* user_tab, int64_t n_q, const float* item_tab, int64_t n_items, int K, float* out_score,
                     int32_t* out_id, float* dbg, void* workspace, size_t ws_bytes, void* stream) {
    if (!user_tab || !item_tab || !out_score || !out_id || !dbg) return CR_ERR_ARG;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    const cr::ExactJob job{user_tab, nullptr, n_q, item_tab, nullptr, 0, n_items, 64, nullptr, nullptr, nullptr, 0, K, out_score, out_id};
    return cr::launch_tc_scorer(job, nullptr, workspace, ws_bytes, (cudaStream_t)stream, dbg);
}

int cr_debug_tc_timeline(unsigned long long* host_out, int n_units) {
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    return cr::read_tc_timeline(host_out, n_units);
}

}  // extern "C"
