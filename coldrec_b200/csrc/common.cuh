// Shared helpers for the coldrec_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include <math_constants.h>

#include "../../include/coldrec_b200.h"

#define CR_FULL_MASK 0xffffffffu

namespace cr {

// Record a CUDA error for cr_last_cuda_error() and return CR_ERR_CUDA.
int note_cuda_error(cudaError_t e, const char* where);
int require_device();   // CR_OK iff current device is sm_100; cached per device
void count_launch();    // every kernel this library launches is counted (cr_launch_count)
// Optional CUDA-event brackets around the dominant kernels (cr_profile_*): tag 0 = scoring sweep, 1 = SpMM rows.
enum { PROF_SCORE_SWEEP = 0, PROF_SPMM_ROWS = 1, PROF_TAGS = 2 };
void prof_start(int tag, cudaStream_t st);
void prof_stop(int tag, cudaStream_t st);

#define CR_CUDA_TRY(expr)                                          \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return cr::note_cuda_error(_e, #expr); \
    } while (0)

#define CR_LAUNCH_CHECK(name)                                      \
    do {                                                           \
        cudaError_t _e = cudaGetLastError();                       \
        if (_e != cudaSuccess) return cr::note_cuda_error(_e, name); \
        cr::count_launch();                                        \
    } while (0)

// One exact-scoring job (score_simt.cu); shared with the TF32 path, which uses it to refine.
struct ExactJob {
    const float* user_tab; const int32_t* user_ids; int64_t n_q;
    const float* item_tab; const int32_t* item_gids; int64_t item_id_base; int64_t n_items; int d;
    const int64_t* mask_rowptr; const int32_t* mask_col; const uint8_t* item_flags; uint8_t flag_exclude;
    int K; float* out_score; int32_t* out_id;
};
struct RefineList {
    const int32_t* list; const int32_t* count; int64_t cap;   // device list of query indices, device count, capacity
};
int exact_splits(int64_t n_q, int64_t n_items);
size_t exact_workspace_bytes(int64_t n_q, int64_t n_items, int K);
size_t refine_workspace_bytes(int K);
int launch_exact_scorer(const ExactJob& job, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_exact_refine(const ExactJob& job, const RefineList& rl, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_merge(const float* in_score, const int32_t* in_id, int G, int64_t n_q, int K, float* out_score, int32_t* out_id,
                 cudaStream_t st, const RefineList* rl, int64_t part_stride, int compact, int64_t gate_lo, int64_t gate_hi);
size_t tc_workspace_bytes(int64_t n_q, int64_t n_items, int d, int K);
int launch_tc_scorer(const ExactJob& job, int32_t* n_refined, void* ws, size_t ws_bytes, cudaStream_t st, float* dbg_scores);
int read_tc_timeline(unsigned long long* host_out, int n_units);

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Total order used by every top-K structure: higher score first, then smaller id.
__device__ __forceinline__ bool better(float sa, int ia, float sb, int ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// membership of `key` in the ascending int32 range [lo, hi) of `col`
__device__ __forceinline__ bool csr_row_contains(const int32_t* __restrict__ col, int64_t lo, int64_t hi, int32_t key) {
    int64_t end = hi;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        int32_t v = __ldg(col + mid);
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(col + lo) == key;
}

}  // namespace cr
