// K5 / K6 — the training-side callers of the propagation path.
//
// K5  fused BPR step.  Replaces, for one mini-batch of the LightGCN-family training loop
//     (model/LightGCN.py:21-28, identical in MF.py / SimGCL.py / NGCF.py / KNN.py ...):
//         user_emb, pos, neg = rec_user_emb[u], rec_item_emb[i], rec_item_emb[j]          (:24)
//         loss = bpr_loss(...) + l2_reg_loss(reg, user_emb, pos, neg)                     (:25, util/utils.py:25-29, 43-47)
//         loss.backward()        -> the index_put/scatter-add gradients of the three gathers
//     by two kernels: bpr_forward (scores, loss, Frobenius norms, per-sample coefficient) and bpr_backward
//     (row gradients accumulated into dense (N, d) gradient tables with 128-bit vector reductions).  The
//     backward of the propagation itself is the SAME SpMM kernel applied to those gradient tables (the
//     normalised adjacency is symmetric, util/databuilder.py:236-248), see coldrec_b200/training.py.
//     adam_step is torch.optim.Adam's single-tensor update (lr, betas, eps; no weight decay / amsgrad), :16.
//
// K6  pairwise sampler.  Replaces next_batch_pairwise (util/utils.py:123-157): an epoch permutation of the
//     training pairs and, per pair, one negative item drawn uniformly from the item table and re-drawn while
//     it is one of the user's training items.  The permutation is a keyed Feistel bijection evaluated per
//     index (no shuffle pass, no O(E) permutation array), the draws come from Philox4x32-10 keyed by
//     (seed, epoch) with counter (sample index, attempt), and membership is a binary search in the same sorted
//     train CSR the scorer uses as its mask.
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------ BPR
struct BprParams {
    const float* user_emb; const float* item_emb; int d;
    const int32_t* u; const int32_t* i; const int32_t* j; int64_t B;
    float reg;
    float* coef;              // [B] dL/d(pos - neg) per sample, already divided by B
    double* partials;         // [n_blocks][4]: loss sum, |U_b|^2, |I_b|^2, |J_b|^2
    int n_blocks;             // blocks of the forward kernel
    double* totals;           // [4] the partials reduced by bpr_reduce_kernel
    float* loss;              // [4]: total, bpr, reg, unused
    float* grad_user; float* grad_item;
};

// LPR lanes own one sample; each lane carries NV float4 of every row.
template <int LPR, int NV>
__global__ void __launch_bounds__(kThreads) bpr_forward_kernel(const BprParams p) {
    __shared__ double s_red[kThreads / 32][4];
    constexpr int SPW = 32 / LPR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / LPR, sub = lane % LPR;
    const int64_t s = ((int64_t)blockIdx.x * (kThreads / 32) + warp) * SPW + grp;
    float pos = 0.f, neg = 0.f, su = 0.f, si = 0.f, sj = 0.f;
    const bool live = s < p.B;
    if (live) {
        const float4* ur = reinterpret_cast<const float4*>(p.user_emb + (int64_t)__ldg(p.u + s) * p.d);
        const float4* ir = reinterpret_cast<const float4*>(p.item_emb + (int64_t)__ldg(p.i + s) * p.d);
        const float4* jr = reinterpret_cast<const float4*>(p.item_emb + (int64_t)__ldg(p.j + s) * p.d);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = sub + v * LPR;
            if (c * 4 < p.d) {
                const float4 a = __ldg(ur + c), b = __ldg(ir + c), n = __ldg(jr + c);
                pos += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
                neg += a.x * n.x + a.y * n.y + a.z * n.z + a.w * n.w;
                su += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
                si += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
                sj += n.x * n.x + n.y * n.y + n.z * n.z + n.w * n.w;
            }
        }
    }
#pragma unroll
    for (int off = LPR / 2; off > 0; off >>= 1) {
        pos += __shfl_xor_sync(CR_FULL_MASK, pos, off);
        neg += __shfl_xor_sync(CR_FULL_MASK, neg, off);
        su += __shfl_xor_sync(CR_FULL_MASK, su, off);
        si += __shfl_xor_sync(CR_FULL_MASK, si, off);
        sj += __shfl_xor_sync(CR_FULL_MASK, sj, off);
    }
    double st[4] = {0.0, 0.0, 0.0, 0.0};
    if (live && sub == 0) {
        // util/utils.py:25-29: -log(10e-6 + sigmoid(pos - neg)), fp32 like the reference
        const float x = pos - neg;
        const float sg = 1.0f / (1.0f + expf(-x));
        st[0] = (double)(-logf(1e-5f + sg));
        st[1] = su; st[2] = si; st[3] = sj;
        // d/dx [-log(eps + s(x))] = -s(1-s)/(eps + s); the mean over the batch folds in 1/B
        p.coef[s] = -(sg * (1.0f - sg)) / (1e-5f + sg) / (float)p.B;
    }
    // fixed-order reduction: lanes -> warp -> block (deterministic)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        for (int off = 16; off > 0; off >>= 1) st[k] += __shfl_down_sync(CR_FULL_MASK, st[k], off);
        if (lane == 0) s_red[warp][k] = st[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += s_red[w][threadIdx.x];
        p.partials[(int64_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}

__device__ __forceinline__ void red_add4(float* dst, float4 v) {
    // one 128-bit reduction instead of four scalar atomics (sm_90+)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// One block: the forward partials summed in a fixed order (threads stride the blocks, then a fixed tree).
__global__ void __launch_bounds__(kThreads) bpr_reduce_kernel(const double* __restrict__ partials, int n_blocks, double* __restrict__ totals) {
    __shared__ double s_red[kThreads][4];
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < n_blocks; b += kThreads)
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] += partials[(int64_t)b * 4 + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) s_red[threadIdx.x][k] = t[k];
    __syncthreads();
    for (int off = kThreads / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off)
#pragma unroll
            for (int k = 0; k < 4; ++k) s_red[threadIdx.x][k] += s_red[threadIdx.x + off][k];
        __syncthreads();
    }
    if (threadIdx.x < 4) totals[threadIdx.x] = s_red[0][threadIdx.x];
}

template <int LPR, int NV>
__global__ void __launch_bounds__(kThreads) bpr_backward_kernel(const BprParams p) {
    const double* s_tot = p.totals;
    const double B = (double)p.B;
    // l2_reg_loss (util/utils.py:43-47): reg * sum_t ||T||_F / B ; d/dT = reg * T / (B ||T||_F)
    const float nu = (float)sqrt(s_tot[1]), ni = (float)sqrt(s_tot[2]), nj = (float)sqrt(s_tot[3]);
    const float cu = nu > 0.f ? p.reg / ((float)p.B * nu) : 0.f;
    const float ci = ni > 0.f ? p.reg / ((float)p.B * ni) : 0.f;
    const float cj = nj > 0.f ? p.reg / ((float)p.B * nj) : 0.f;
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.loss) {
        const float bpr = (float)(s_tot[0] / B);
        const float regl = p.reg * (nu / (float)p.B + ni / (float)p.B + nj / (float)p.B);
        p.loss[0] = bpr + regl; p.loss[1] = bpr; p.loss[2] = regl; p.loss[3] = 0.f;
    }
    constexpr int SPW = 32 / LPR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / LPR, sub = lane % LPR;
    const int64_t s = ((int64_t)blockIdx.x * (kThreads / 32) + warp) * SPW + grp;
    if (s >= p.B) return;
    const int64_t u = __ldg(p.u + s), i = __ldg(p.i + s), j = __ldg(p.j + s);
    const float g = __ldg(p.coef + s);
    const float4* ur = reinterpret_cast<const float4*>(p.user_emb + u * p.d);
    const float4* ir = reinterpret_cast<const float4*>(p.item_emb + i * p.d);
    const float4* jr = reinterpret_cast<const float4*>(p.item_emb + j * p.d);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int c = sub + v * LPR;
        if (c * 4 < p.d) {
            const float4 a = __ldg(ur + c), b = __ldg(ir + c), n = __ldg(jr + c);
            float4 gu, gi, gj;
            gu.x = g * (b.x - n.x) + cu * a.x; gu.y = g * (b.y - n.y) + cu * a.y;
            gu.z = g * (b.z - n.z) + cu * a.z; gu.w = g * (b.w - n.w) + cu * a.w;
            gi.x = g * a.x + ci * b.x; gi.y = g * a.y + ci * b.y; gi.z = g * a.z + ci * b.z; gi.w = g * a.w + ci * b.w;
            gj.x = cj * n.x - g * a.x; gj.y = cj * n.y - g * a.y; gj.z = cj * n.z - g * a.z; gj.w = cj * n.w - g * a.w;
            red_add4(p.grad_user + u * p.d + c * 4, gu);
            red_add4(p.grad_item + i * p.d + c * 4, gi);
            red_add4(p.grad_item + j * p.d + c * 4, gj);
        }
    }
}

// ------------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam, single-tensor path (torch/optim/adam.py::_single_tensor_adam), per element:
//   m += (g - m) * (1 - beta1);  v = v * beta2 + (1 - beta2) * g * g;
//   p += -step_size * (m / (sqrt(v) / bc2_sqrt + eps));   step_size = lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t)
// 28 B/element of HBM traffic (read p, g, m, v; write p, m, v): one pass, float4 accesses.
__global__ void __launch_bounds__(kThreads) adam_step_kernel(float4* __restrict__ p4, const float4* __restrict__ g4,
                                                             float4* __restrict__ m4, float4* __restrict__ v4, int64_t n4,
                                                             float* __restrict__ p, const float* __restrict__ g,
                                                             float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                             float w1, float beta2, float w2, float step_size, float bc2_sqrt,
                                                             float eps, float grad_scale, const float* __restrict__ dev_scalars) {
    if (dev_scalars) {      // CUDA-graph replay: the step-dependent bias corrections come from device memory
        step_size = __ldg(dev_scalars);
        bc2_sqrt = __ldg(dev_scalars + 1);
    }
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= grad_scale;
        mm = mm + (gg - mm) * w1;
        vv = vv * beta2 + (w2 * gg) * gg;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        pp = pp + (-step_size) * (mm / denom);
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += stride) {
        float4 pp = p4[e], mm = m4[e], vv = v4[e];
        const float4 gg = __ldg(g4 + e);
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        p4[e] = pp; m4[e] = mm; v4[e] = vv;
    }
    for (int64_t e = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) upd(p[e], g[e], m[e], v[e]);
}

// ------------------------------------------------------------------------------------------------ sampler
__device__ __forceinline__ uint32_t philox_word0(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
// Bijection of [0, n): 6-round balanced Feistel network on 2*half_bits bits, cycle-walked into range.
__device__ __forceinline__ uint64_t feistel_perm(uint64_t x, uint64_t n, int half_bits, uint32_t k0, uint32_t k1) {
    const uint32_t mask = (half_bits >= 32) ? 0xffffffffu : ((1u << half_bits) - 1u);
    do {
        uint32_t L = (uint32_t)(x >> half_bits) & mask, R = (uint32_t)x & mask;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const uint32_t f = fmix32(R ^ (k0 + 0x9E3779B9u * (uint32_t)r)) ^ fmix32((R + k1) ^ (0x85EBCA6Bu * (uint32_t)(r + 1)));
            const uint32_t nl = R;
            R = (L ^ f) & mask;
            L = nl;
        }
        x = ((uint64_t)L << half_bits) | R;
    } while (x >= n);
    return x;
}

struct SampleParams {
    const int32_t* pair_user; const int32_t* pair_item; int64_t n_pairs;
    const int64_t* rowptr; const int32_t* col; int32_t n_items;
    uint32_t k0, k1, e0, e1; int half_bits;
    int64_t begin, count; int max_attempts;
    int32_t* out_u; int32_t* out_i; int32_t* out_j; int32_t* n_exhausted;
};

__global__ void __launch_bounds__(kThreads) sample_pairwise_kernel(const SampleParams p) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.count) return;
    const uint64_t pos = (uint64_t)(p.begin + t);
    const uint64_t src = feistel_perm(pos, (uint64_t)p.n_pairs, p.half_bits, p.k0 ^ p.e0, p.k1 ^ p.e1);
    const int32_t u = __ldg(p.pair_user + src), i = __ldg(p.pair_item + src);
    const int64_t lo = __ldg(p.rowptr + u), hi = __ldg(p.rowptr + u + 1);
    int32_t j = 0;
    int a = 0;
    for (;; ++a) {
        const uint32_t r = philox_word0((uint32_t)pos, (uint32_t)(pos >> 32), (uint32_t)a, p.e0, p.k0, p.k1 ^ p.e1);
        j = (int32_t)__umulhi(r, (uint32_t)p.n_items);          // uniform over [0, n_items)
        if (!cr::csr_row_contains(p.col, lo, hi, j)) break;
        if (a + 1 >= p.max_attempts) {                          // user interacted with (almost) every item: the reference
            if (p.n_exhausted) atomicAdd(p.n_exhausted, 1);     // would spin forever here (util/utils.py:140-152)
            break;
        }
    }
    p.out_u[t] = u; p.out_i[t] = i; p.out_j[t] = j;
}

template <typename F>
int dispatch_lpr(int d, F&& f) {
    // LPR lanes x NV float4 per row, LPR*NV*4 >= d
    if (d <= 32) return f(std::integral_constant<int, 8>{}, std::integral_constant<int, 1>{});
    if (d <= 64) return f(std::integral_constant<int, 16>{}, std::integral_constant<int, 1>{});
    if (d <= 128) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 1>{});
    if (d <= 256) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 2>{});
    return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 4>{});
}

}  // namespace

extern "C" {

size_t cr_bpr_workspace_bytes(int64_t batch) {
    if (batch < 0) return 0;
    const int64_t blocks = (batch + 7) / 8 + 1;      // >= forward blocks for every d (>= 8 samples per block)
    return cr::align_up((size_t)batch * 4, 256) + 256 + (size_t)blocks * 4 * sizeof(double) + 256;
}

int cr_bpr_fwd_bwd_f32(const float* user_emb, const float* item_emb, int d, const int32_t* u_idx, const int32_t* i_idx,
                       const int32_t* j_idx, int64_t batch, float reg, float* loss, float* grad_user, float* grad_item,
                       void* workspace, size_t ws_bytes, void* stream) {
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    if (batch < 0 || !loss) return CR_ERR_ARG;
    if (batch == 0) return CR_OK;
    if (!user_emb || !item_emb || !u_idx || !i_idx || !j_idx || !grad_user || !grad_item) return CR_ERR_ARG;
    if (d <= 0 || d % 4 != 0 || d > 512) return CR_ERR_UNSUPPORTED;
    if (!cr::aligned16(user_emb) || !cr::aligned16(item_emb) || !cr::aligned16(grad_user) || !cr::aligned16(grad_item))
        return CR_ERR_ALIGN;
    if (!workspace || ws_bytes < cr_bpr_workspace_bytes(batch)) return CR_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    BprParams p{};
    p.user_emb = user_emb; p.item_emb = item_emb; p.d = d; p.u = u_idx; p.i = i_idx; p.j = j_idx; p.B = batch; p.reg = reg;
    p.coef = (float*)workspace;
    p.totals = (double*)((char*)workspace + cr::align_up((size_t)batch * 4, 256));
    p.partials = p.totals + 32;
    p.loss = loss; p.grad_user = grad_user; p.grad_item = grad_item;
    return dispatch_lpr(d, [&](auto lpr, auto nv) -> int {
        constexpr int LPR = decltype(lpr)::value, NV = decltype(nv)::value;
        const int64_t per_block = (kThreads / 32) * (32 / LPR);
        const unsigned blocks = (unsigned)((batch + per_block - 1) / per_block);
        p.n_blocks = (int)blocks;
        bpr_forward_kernel<LPR, NV><<<blocks, kThreads, 0, st>>>(p);
        CR_LAUNCH_CHECK("bpr_forward_kernel");
        bpr_reduce_kernel<<<1, kThreads, 0, st>>>(p.partials, p.n_blocks, p.totals);
        CR_LAUNCH_CHECK("bpr_reduce_kernel");
        bpr_backward_kernel<LPR, NV><<<blocks, kThreads, 0, st>>>(p);
        CR_LAUNCH_CHECK("bpr_backward_kernel");
        return CR_OK;
    });
}

int cr_adam_scalars(double lr, double beta1, double beta2, int64_t step, float* host_out2) {
    if (!host_out2 || step < 1) return CR_ERR_ARG;
    host_out2[0] = (float)(lr / (1.0 - pow(beta1, (double)step)));
    host_out2[1] = (float)sqrt(1.0 - pow(beta2, (double)step));
    return CR_OK;
}

int cr_adam_step_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                     double beta2, double eps, int64_t step, float grad_scale, const float* dev_scalars, void* stream) {
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    if (n < 0 || step < 1) return CR_ERR_ARG;
    if (n == 0) return CR_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq) return CR_ERR_ARG;
    if (!cr::aligned16(param) || !cr::aligned16(grad) || !cr::aligned16(exp_avg) || !cr::aligned16(exp_avg_sq)) return CR_ERR_ALIGN;
    // scalar prefactors in double, as the Python floats of torch/optim/adam.py
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const float step_size = (float)(lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    const int64_t n4 = n / 4;
    const int64_t want = (n4 + kThreads - 1) / kThreads;
    const unsigned blocks = (unsigned)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
    adam_step_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>((float4*)param, (const float4*)grad, (float4*)exp_avg,
                                                                   (float4*)exp_avg_sq, n4, param, grad, exp_avg, exp_avg_sq, n,
                                                                   (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), step_size,
                                                                   bc2_sqrt, (float)eps, grad_scale, dev_scalars);
    CR_LAUNCH_CHECK("adam_step_kernel");
    return CR_OK;
}

int cr_sample_pairwise(const int32_t* pair_user, const int32_t* pair_item, int64_t n_pairs, const int64_t* train_rowptr,
                       const int32_t* train_col, int32_t n_items, uint64_t seed, uint64_t epoch, int64_t begin, int64_t count,
                       int32_t* out_user, int32_t* out_pos, int32_t* out_neg, int32_t* n_exhausted, void* stream) {
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    if (n_pairs < 0 || count < 0 || begin < 0 || begin + count > n_pairs || n_items <= 0) return CR_ERR_ARG;
    if (count == 0) return CR_OK;
    if (!pair_user || !pair_item || !train_rowptr || !train_col || !out_user || !out_pos || !out_neg) return CR_ERR_ARG;
    SampleParams p{};
    p.pair_user = pair_user; p.pair_item = pair_item; p.n_pairs = n_pairs; p.rowptr = train_rowptr; p.col = train_col;
    p.n_items = n_items; p.k0 = (uint32_t)seed; p.k1 = (uint32_t)(seed >> 32); p.e0 = (uint32_t)epoch; p.e1 = (uint32_t)(epoch >> 32);
    int bits = 1;
    while (bits < 64 && ((uint64_t)1 << bits) < (uint64_t)n_pairs) ++bits;
    p.half_bits = (bits + 1) / 2;
    p.begin = begin; p.count = count; p.max_attempts = 4096;
    p.out_u = out_user; p.out_i = out_pos; p.out_j = out_neg; p.n_exhausted = n_exhausted;
    sample_pairwise_kernel<<<(unsigned)((count + kThreads - 1) / kThreads), kThreads, 0, (cudaStream_t)stream>>>(p);
    CR_LAUNCH_CHECK("sample_pairwise_kernel");
    return CR_OK;
}

}  // extern "C"
