/*
 * coldrec_b200.h — C ABI of the B200-native embedding-generation-and-scoring path of ColdRec.
 *
 * The reference (YuanchenBei/ColdRec) is pure Python; the "FFI" a maintainer binds is therefore a
 * ctypes stub.  Every entry point below replaces a torch call site of the
 * reference, cited as file:line relative to the ColdRec tree.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.  All pointers are DEVICE pointers on the
 *     current CUDA device unless a parameter says "host".
 *   - The caller owns every buffer including workspaces; functions never allocate, free or retain
 *     pointers, and are asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - Return value: CR_OK (0) or a negative CR_ERR_* code; no C++ exception crosses the ABI.
 *   - Tables are fp32 row-major, rows 16-byte aligned (d % 4 == 0); ids are int32; CSR row pointers int64.
 *   - There is NO CPU fallback: without an sm_100 device every compute call returns CR_ERR_NO_DEVICE.
 */
#ifndef COLDREC_B200_H
#define COLDREC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CR_API __attribute__((visibility("default")))
#else
#define CR_API
#endif

#define CR_OK 0
#define CR_ERR_ARG (-1)         /* null pointer / negative size / inconsistent shape            */
#define CR_ERR_ALIGN (-2)       /* pointer or leading dimension not 16-byte aligned             */
#define CR_ERR_UNSUPPORTED (-3) /* d, K or another parameter outside the supported set           */
#define CR_ERR_WORKSPACE (-4)   /* workspace smaller than cr_*_workspace_bytes() says            */
#define CR_ERR_CUDA (-5)        /* a CUDA runtime/driver call failed (see cr_last_cuda_error)    */
#define CR_ERR_NO_DEVICE (-6)   /* no CUDA device, or the current device is not sm_100           */

#define CR_MAX_K 64             /* largest top-K per call (reference default max(topN)=20)       */
#define CR_MASK_SCORE (-1.0e9f) /* model/BaseRecommender.py:177,180  (`-10e8`)                   */

#define CR_SCORE_EXACT_F32 0    /* fp32 FFMA scoring, no approximation anywhere                  */
#define CR_SCORE_TF32_CHECKED 1 /* tcgen05 TF32 selection + exact fp32 rescoring + margin check  */

#define CR_ACT_NONE 0
#define CR_ACT_TANH 1
#define CR_ACT_LEAKY_RELU 2     /* slope 0.01 (torch F.leaky_relu default), model/NGCF.py:99     */

CR_API const char *cr_strerror(int code);
CR_API const char *cr_last_cuda_error(void); /* text of the last CUDA error seen by this library (thread-local) */
CR_API int cr_version(void);
CR_API int cr_device_check(void);            /* CR_OK iff the current device is compute capability 10.x */

/* Measurement hooks (bench.py): number of kernels this library has launched so far, and CUDA-event
 * brackets around the dominant kernels (tag 0 = tcgen05 scoring sweep, tag 1 = SpMM row kernel) recorded
 * on the launching stream while enabled; cr_profile_read synchronises on the recorded events. */
CR_API unsigned long long cr_launch_count(void);
CR_API int cr_profile_enable(int on);
CR_API int cr_profile_read(int tag, double *total_ms, int *launches);

/* ------------------------------------------------------------------------------------------------
 * K3 — CSR SpMM with fused layer accumulation.
 * Replaces `torch.sparse.mm(self.sparse_norm_adj, ego_embeddings)` (model/LightGCN.py:90,
 * model/NGCF.py:95, model/SimGCL.py:105, model/XSimGCL.py:111, model/NCL.py:190, model/CGRC.py:70,89,
 * model/FSGNN.py:363,404,438) and the `torch.stack` + `torch.mean` that follows (LightGCN.py:92-93).
 *
 *   y[r,:]   = sum_j val[j] * X[col[j],:]   for j in [rowptr[r], rowptr[r+1])      r in [0, n_rows)
 *   Y[r,:]   = y[r,:]                                  if Y   != NULL
 *   acc[r,:] = (acc_beta*acc_in[r,:] + y[r,:]) / acc_div   if acc != NULL   (acc_in == NULL means in place;
 *                                                           acc_in is not read when acc_beta == 0)
 *
 * rowptr/Y/acc describe the caller's LOCAL row block (row-partitioned multi-GPU passes its slice);
 * X is the full gather source.  val == NULL means all-ones.  `plan` (from cr_spmm_plan) enables the
 * split path for rows longer than 512 nonzeros; plan == NULL processes every row with one warp.
 * A plan belongs to one (rowptr, nnz, d) triple and is reused across layers and epochs.  It also holds the partial-sum
 * scratch of the split path and (round 2) the table of work-balanced warp row ranges: calls that share a plan must be
 * ordered on one stream (one plan per stream otherwise).
 * ------------------------------------------------------------------------------------------------ */
CR_API size_t cr_spmm_plan_bytes(int64_t n_rows, int64_t nnz, int d);
CR_API int cr_spmm_plan(const int64_t *rowptr, int64_t n_rows, int64_t nnz, int d, void *plan, size_t plan_bytes,
                 void *stream);
CR_API int cr_spmm_csr_f32(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n_rows, int64_t nnz,
                    const float *X, int d, float *Y, const float *acc_in, float *acc, float acc_beta, float acc_div,
                    void *plan, size_t plan_bytes, void *stream);

/* Multi-GPU form: the all-gather that rebuilds the next layer's gather source is fused into the SpMM
 * epilogue.  peer_tables (DEVICE array of n_peers pointers, e.g. torch symmetric-memory buffer_ptrs_dev)
 * names the full (padded) embedding table of every GPU of the node, this one included; finished row r is
 * stored to peer_tables[p][(peer_row_offset + r), :] (r < peer_row_split) or peer_tables[p][(peer_row_offset_hi + r), :]
 * (r >= peer_row_split; pass peer_row_split = n_rows for a single destination range) for every p with plain peer-memory
 * stores over NVLink, overlapped with the gathers of the rows still in flight.  Two ranges, because a rank of the
 * bipartite row partition owns one block of user rows and one of item rows: the last layer lands directly in the
 * reference's node numbering (users then items, model/LightGCN.py:87,94-95).  peer_need (nullable, n_peers <= 8):
 * one byte per local row, bit p set iff GPU p reads that row in the next layer (or wants it in the result) — rows are
 * only stored where they are needed (a sparse all-gather); NULL = every row to every GPU.  mc_table (nullable): NVLS
 * multicast address of the same destination table (torch symmetric memory `multicast_ptr`); a row wanted by every GPU is
 * then written with ONE multimem.st that the NVSwitch replicates, instead of n_peers unicast stores.  The caller synchronises the GPUs between layers
 * (symmetric-memory barrier).  acc as in cr_spmm_csr_f32 (local rows); with bcast_acc != 0 the peers receive
 * the acc result (the finished layer mean) instead of y. */
CR_API int cr_spmm_csr_bcast_f32(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n_rows, int64_t nnz,
                                 const float *X, int d, float *const *peer_tables, int n_peers, int64_t peer_row_offset,
                                 int64_t peer_row_split, int64_t peer_row_offset_hi, int bcast_acc, const uint8_t *peer_need,
                                 float *mc_table, const float *acc_in,
                                 float *acc, float acc_beta, float acc_div, void *plan, size_t plan_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * K1 — fused full-catalogue scorer: score -> train mask -> flag mask -> top-K, no score matrix in HBM.
 * Replaces `batch_predict` (model/MF.py:58-63 and its 21 copies) + the mask writes + `torch.topk`
 * inside `BaseColdStartTrainer._evaluate` (model/BaseRecommender.py:170-182).
 *
 *   query j   = user_tab[user_ids ? user_ids[j] : j, :]                          j in [0, n_q)
 *   item  p   = item_tab[p, :], global id gid(p) = item_gids ? item_gids[p] : item_id_base + p
 *   s(j,p)    = CR_MASK_SCORE  if gid(p) is in row j of the CSR (mask_rowptr, mask_col)   [train mask, :175-177]
 *             = CR_MASK_SCORE  if item_flags[gid(p)] & flag_exclude                      [column mask, :179-180]
 *             = <query j, item p>  (fp32)                                                [MF.py:62]
 *   out       = the K best (s desc, gid asc) per query: out_score[j,:], out_id[j,:] (global ids).
 *
 * item_gids, when given, must be strictly increasing; mask_col rows must be sorted ascending.
 * If fewer than K items exist the tail is padded with (-inf, -1).  With CR_SCORE_TF32_CHECKED the
 * selection runs on tcgen05 TF32 and every returned score is re-computed in fp32; queries whose
 * TF32 margin could not be proven are re-run exactly inside the same call, and their number is
 * written to *n_refined (device int32, nullable).  The tensor-core path serves K <= 52 and every width d <= 128 with
 * d % 4 == 0: d = 64 and d = 128 natively (d = 128 = VBPR / AMR's concatenated tables, model/VBPR.py:68-75), other widths
 * zero-padded to the next of the two inside the workspace; wider tables and K > 52 take the exact fp32 kernel.
 * ------------------------------------------------------------------------------------------------ */
CR_API size_t cr_score_topk_workspace_bytes(int64_t n_q, int64_t n_items, int d, int K, int precision);
CR_API int cr_score_topk_f32(const float *user_tab, const int32_t *user_ids, int64_t n_q, const float *item_tab,
                      const int32_t *item_gids, int64_t item_id_base, int64_t n_items, int d,
                      const int64_t *mask_rowptr, const int32_t *mask_col, const uint8_t *item_flags,
                      uint8_t flag_exclude, int K, float *out_score, int32_t *out_id, int32_t *n_refined,
                      int precision, void *workspace, size_t ws_bytes, void *stream);

/* Diagnostic probe of the tcgen05 path: like cr_score_topk_f32 (d = 64, TF32-checked, no masks) and
 * additionally dumps the raw TF32 scores of the first 256 queries x first 96 items into dbg[256*96]. */
CR_API int cr_debug_tc_tile(const float *user_tab, int64_t n_q, const float *item_tab, int64_t n_items, int K,
                            float *out_score, int32_t *out_id, float *dbg, void *workspace, size_t ws_bytes, void *stream);

/* Diagnostic: with the environment variable CR_TC_DEBUG_MODE=8 the tcgen05 sweep stamps %globaltimer at fixed points
 * of every unit (16 uint64 slots per unit: entry, setup, queries in TMEM, after tile 0/15/127/1023/4095/8191/last, exit);
 * this copies the stamps of the first n_units (<= 8192) units of the last sweep to HOST memory. */
CR_API int cr_debug_tc_timeline(unsigned long long *host_out, int n_units);

/* Merge G candidate lists per query (item shards / ALDI's two item groups, model/ALDI.py:149-160 /
 * per-GPU candidates after the NCCL allgather) into one top-K by (score desc, id asc).
 * in_score/in_id: [G, n_q, K] (list g of query j at ((g*n_q)+j)*K); ids < 0 are padding. */
CR_API int cr_topk_merge(const float *in_score, const int32_t *in_id, int G, int64_t n_q, int K, float *out_score,
                  int32_t *out_id, void *stream);

/* Lists that came up short (trailing ids < 0) because flagged items were compacted away before the sweep:
 * fill the tail with masked ids at CR_MASK_SCORE, which is what the reference's top-K shows when a query
 * has fewer than K unmasked items (model/BaseRecommender.py:177-182; which masked ids is unspecified). */
CR_API int cr_fill_masked(float *out_score, int32_t *out_id, int64_t n_q, int K, int64_t n_items_total,
                          const uint8_t *item_flags, uint8_t flag_exclude, const int64_t *mask_rowptr,
                          const int32_t *mask_col, void *stream);

/* dst[j,:] = src[ids[j],:]  — `self.user_emb[users]` (model/MF.py:62) and flag-compaction of item tables. */
CR_API int cr_gather_rows_f32(const float *src, const int32_t *ids, int64_t n, int d, float *dst, void *stream);

/* dst[dst_ids ? dst_ids[j] : j, :] = src[src_ids ? src_ids[j] : j, :]  — row overwrite between tables:
 * `h[cold_rows] = item_x[cold_item_idx]` after every convolution of CGRC's frozen-cold propagation
 * (model/CGRC.py:90-91), `item_emb[cold] = generated` (model/GAR.py:44-46).  dst_ids must not repeat. */
CR_API int cr_copy_rows_f32(const float *src, const int32_t *src_ids, const int32_t *dst_ids, int64_t n, int d, float *dst,
                            void *stream);

/* ------------------------------------------------------------------------------------------------
 * K2 — ranking metrics reduced on device.  Replaces `ranking_evaluation` / `Metric.*`
 * (util/evaluator.py:9-32, 47-63, 95-115, 153-187).
 *   topk_id [n_q, K]   sorted top-K global ids per query (ids < 0 ignored)
 *   gt CSR             row j = ground-truth item ids of query j, sorted ascending
 *   Ns (host) [nN]     cut-offs, each <= K
 *   inv_log2 [K]       1/log(n+2, 2); idcg_prefix [K+1]: prefix sums of inv_log2 (both computed by the
 *                      host exactly as util/evaluator.py:106-109 does, so per-user DCG/IDCG are bit-equal)
 *   hits [nN, n_q] int32 and dcg [nN, n_q] double per-query outputs (nullable)
 *   sums [nN, 6] double: sum_hits, sum_gt, sum(hits/|gt|) over |gt|>0, #{|gt|>0}, sum(dcg/idcg) over idcg>0, #{idcg>0}
 * ------------------------------------------------------------------------------------------------ */
CR_API size_t cr_rank_metrics_workspace_bytes(int64_t n_q, int nN);
CR_API int cr_rank_metrics(const int32_t *topk_id, int64_t n_q, int K, const int64_t *gt_rowptr, const int32_t *gt_col,
                    const int32_t *Ns_host, int nN, const double *inv_log2, const double *idcg_prefix, int32_t *hits,
                    double *dcg, double *sums, void *workspace, size_t ws_bytes, void *stream);

/* ------------------------------------------------------------------------------------------------
 * K4 — content->embedding towers.  One fused layer:
 *   Y[yrow ? yrow[r] : r, :] = act( ([X1[xr,:] | X2[xr,:]] . W^T + bias) * scale + shift ),  xr = xrow ? xrow[r] : r
 * W is torch nn.Linear layout [n_out, d1+d2]; scale/shift (nullable) carry an eval-mode BatchNorm1d.
 * Replaces nn.Linear -> BatchNorm1d(eval) -> tanh chains: model/DropoutNet.py:204-212,222-236,
 * model/Heater.py:143-167,218-222, model/GAR.py:102-107, model/ALDI.py:204-208; the two-segment
 * input is the `torch.cat((Vin, Vcontent), 1)` of DropoutNet.py:199-202; xrow/yrow are the
 * `content[cold_idx]` gather and `item_emb[cold_idx] = ...` scatter of GAR.py:44-46 / ALDI.py:96.
 * ------------------------------------------------------------------------------------------------ */
CR_API int cr_linear_act_f32(const float *X1, int64_t ld1, int d1, const float *X2, int64_t ld2, int d2, const int32_t *xrow,
                      int64_t n_rows, const float *W, const float *bias, const float *scale, const float *shift,
                      int n_out, int act, float *Y, int64_t ldy, const int32_t *yrow, void *stream);
/* The same layer on the tensor cores (tcgen05, kind::tf32) at fp32 accuracy ("3xTF32"): every operand comes split as
 * x = hi + lo with hi exactly representable in TF32 (cr_split_tf32), and hi.hi + lo.hi + hi.lo is accumulated in fp32.
 * Rows are contiguous (no xrow gather); X1/X2/W tables need 16-byte aligned bases and row strides (ld % 4 == 0);
 * n_out <= 256.  Outputs: Y (nullable, scattered through yrow if given) and/or the split (Yhi, Ylo) of the same values
 * with row stride ldh >= n_out (columns [n_out, ldh) are zeroed) — the next layer's input without a split pass.
 * X1lo == NULL (and X2lo == NULL) = "raw" mode: X1hi / X2hi are plain fp32 rows, split hi + lo inside the kernel (one HBM
 * read, bit-identical results; measured slower on wide-K layers, see tower_tc.cu).
 * Same replaced reference lines as cr_linear_act_f32. */
CR_API int cr_linear_act_tc_f32(const float *X1hi, const float *X1lo, int64_t ld1, int d1, const float *X2hi, const float *X2lo,
                         int64_t ld2, int d2, int64_t n_rows, const float *Whi, const float *Wlo, int64_t ldw,
                         const float *bias, const float *scale, const float *shift, int n_out, int act, float *Y,
                         int64_t ldy, const int32_t *yrow, float *Yhi, float *Ylo, int64_t ldh, void *stream);
/* hi = round-to-nearest-TF32(src), lo = src - hi, for a rows x cols table (row stride ld_src); destination row stride
 * ld_dst >= cols, columns [cols, ld_dst) zeroed (pads a 2,738-wide content table to a TMA-legal stride). */
CR_API int cr_split_tf32(const float *src, int64_t ld_src, int64_t rows, int cols, float *hi, float *lo, int64_t ld_dst,
                  void *stream);
/* Row-wise top-K of a dense score block S [n_rows, n_cols] (row stride ld): out ids = col_id_base + column, ordered by
 * (score desc, id asc); exclude_col (nullable, one column per row) is skipped — the "not myself" of a kNN graph.  With
 * cr_linear_act_tc_f32 producing S = Q . V^T this is the brute-force inner-product search of model/KNN.py:63-77
 * (faiss.IndexFlatIP) and model/FSGNN.py:106-152 for content tables too wide for the fused sweep (d = 300 / 2,738). */
CR_API int cr_topk_rows_f32(const float *S, int64_t n_rows, int n_cols, int64_t ld, int K, const int32_t *exclude_col,
                     int col_id_base, float *out_score, int32_t *out_id, void *stream);
/* scale = gamma / sqrt(var + eps); shift = beta - mean * scale   (eval BatchNorm1d, DropoutNet.py:226-230) */
CR_API int cr_bn_fold_f32(const float *gamma, const float *beta, const float *mean, const float *var, float eps, int n,
                   float *scale, float *shift, void *stream);
/* out = Vin*keep + tanh(sum_g gate[:,g] * expert) * one_minus_keep   (model/Heater.py:189-198) */
CR_API int cr_heater_blend_f32(const float *gate, int n_expert, const float *expert, const float *Vin, float keep,
                        float one_minus_keep, int64_t n_rows, int d, float *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * K5 — fused BPR training step for the propagated tables.
 * Replaces, per mini-batch of model/LightGCN.py:21-28 (same loop in MF / SimGCL / NGCF / KNN ...):
 *     u_e, p_e, n_e = rec_user_emb[u_idx], rec_item_emb[i_idx], rec_item_emb[j_idx]             (:24)
 *     loss = bpr_loss(u_e, p_e, n_e) + l2_reg_loss(reg, u_e, p_e, n_e)        (util/utils.py:25-29, 43-47)
 *     loss.backward()                       (the scatter-add gradients of the three row gathers)
 *   bpr  = mean_b -log(1e-5 + sigmoid(<u,p> - <u,n>)),  regl = reg * (|U_b|_F + |P_b|_F + |N_b|_F) / B
 *   loss[0..2] = bpr + regl, bpr, regl (device fp32[4]);
 *   grad_user[u,:] / grad_item[i,:] += d loss / d row  (dense (n_users,d) / (n_items,d) tables the caller zeroed;
 *   rows that occur several times in the batch accumulate; fp32 vector reductions, order unspecified).
 * The backward of the propagation is cr_spmm_csr_f32 applied to these gradient tables: the normalised
 * adjacency is symmetric (util/databuilder.py:236-248), so d loss / d E0 = mean_k A^k . grad.
 * ------------------------------------------------------------------------------------------------ */
CR_API size_t cr_bpr_workspace_bytes(int64_t batch);
CR_API int cr_bpr_fwd_bwd_f32(const float *user_emb, const float *item_emb, int d, const int32_t *u_idx, const int32_t *i_idx,
                              const int32_t *j_idx, int64_t batch, float reg, float *loss, float *grad_user, float *grad_item,
                              void *workspace, size_t ws_bytes, void *stream);

/* torch.optim.Adam(lr, betas, eps) single-tensor update (model/LightGCN.py:16 + optimizer.step() :28), no weight
 * decay / amsgrad; `step` is the 1-based step count; the gradient is read as grad * grad_scale.
 *   m += (g - m)(1 - beta1);  v = v beta2 + (1 - beta2) g g;  p -= lr/(1 - beta1^t) * m / (sqrt(v)/sqrt(1 - beta2^t) + eps)
 * dev_scalars (device float[2], nullable): when given, the two step-dependent factors lr/(1 - beta1^t) and sqrt(1 - beta2^t)
 * are read from device memory instead of being computed from `step` — a captured CUDA graph of the training step is then
 * replayed for every step with only those 8 bytes refreshed (cr_adam_scalars fills the host copy). */
CR_API int cr_adam_step_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, double lr,
                            double beta1, double beta2, double eps, int64_t step, float grad_scale, const float *dev_scalars,
                            void *stream);
CR_API int cr_adam_scalars(double lr, double beta1, double beta2, int64_t step, float *host_out2);

/* ------------------------------------------------------------------------------------------------
 * K6 — pairwise (user, positive, negative) sampler.
 * Replaces next_batch_pairwise (util/utils.py:123-157): the batch [begin, begin+count) of one epoch's shuffled
 * training pairs, each with one negative item drawn uniformly from [0, n_items) and re-drawn while it is one of
 * the user's training items (train CSR: train_rowptr int64 [n_users+1], train_col int32 ascending per row —
 * the scorer's mask CSR over all users).  The epoch permutation is a keyed bijection of [0, n_pairs) evaluated
 * per index (6-round Feistel network, cycle-walked), the draws are Philox4x32-10 words keyed by (seed, epoch)
 * with counter (position, attempt): batches are reproducible and independent of launch geometry.
 * n_exhausted (device int32, nullable, caller-zeroed) counts samples that still collided after 4096 draws
 * (a user who interacted with nearly every item; the reference loops forever there).
 * ------------------------------------------------------------------------------------------------ */
CR_API int cr_sample_pairwise(const int32_t *pair_user, const int32_t *pair_item, int64_t n_pairs,
                              const int64_t *train_rowptr, const int32_t *train_col, int32_t n_items, uint64_t seed,
                              uint64_t epoch, int64_t begin, int64_t count, int32_t *out_user, int32_t *out_pos,
                              int32_t *out_neg, int32_t *n_exhausted, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* COLDREC_B200_H */
