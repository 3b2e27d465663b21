// K4 (tensor-core path) — one fused tower layer  Y = act((X . W^T + bias) * scale + shift)  on tcgen05, fp32-accurate.
//
// Replaces the nn.Linear -> BatchNorm1d(eval) -> tanh chains of model/DropoutNet.py:204-212,222-236,
// model/Heater.py:143-167,218-222, model/GAR.py:102-107 and model/ALDI.py:204-208 for layers whose rows are contiguous
// (the SIMT kernel of towers.cu keeps the gathered / odd-shaped cases).
//
// Precision.  The parity bar for generated embeddings is 1e-5 norm-wise; one TF32 pass (10 mantissa bits) misses it by two
// orders of magnitude, so every operand is split  x = hi + lo  with hi = round-to-nearest TF32 (exactly representable:
// the tensor core's own fp32 -> tf32 conversion leaves it untouched) and lo = x - hi (exact in fp32), and each K = 8 slice
// issues THREE MMAs into the same fp32 TMEM accumulator:  hi.hi + lo.hi + hi.lo.  The dropped lo.lo term and the TF32
// rounding of lo are ~2^-21 relative — the "3xTF32" scheme; measured against the reference towers it stays at a few 1e-7.
// The split of a constant table (item content: 20,519 x 2,738 at XING shape) is done once (cr_split_tf32) and reused;
// a layer's epilogue can emit its output already split, so a tower chain never runs a separate split pass in between.
// "Raw" mode (X1lo == NULL) takes plain fp32 rows instead: TMA stages the fp32 tile and the four epilogue warps rewrite it in
// shared memory as hi (in place) + lo (second buffer) between accumulator drains, fence it towards the async proxy and hand
// the stage to the MMA warp — one HBM read of X instead of two.  Bit-identical results, but MEASURED SLOWER where it was meant
// to help (XING content layer 0.31 vs 0.26 ms, kNN 4.2 vs 3.0 ms: 48 KB of extra shared-memory traffic per chunk sits on the
// MMA's critical path); it wins only on short-K layers with wide outputs, where it saves two output tables.  towers.py uses
// pre-split operands; the mode stays for callers that cannot keep a split copy.
//
// Accumulation.  The tensor core adds each MMA's result to the TMEM accumulator with truncation, not round-to-nearest: over the
// 1056 MMAs of a K = 2,802 layer that bias reached 1.7e-5 (r02, first version: one accumulator for the whole K loop) — more
// than the whole parity budget.  So the K loop is cut into blocks of kBlockChunks chunks (24 MMAs): each block accumulates
// from zero into one of two TMEM accumulator stages, and the epilogue warps drain a finished block into fp32 running sums
// held in REGISTERS (round-to-nearest adds) while the next block is being issued.  Error now: a few 1e-7.
//
// One CTA = 128 rows x NB (64 or 128) outputs (grid.y walks wider layers); K is walked in 32-float chunks (one SWIZZLE_128B
// TMA box per operand and half: A.hi, A.lo 16 KB each, W.hi, W.lo NB x 128 B each) through a ring of shared-memory stages.  The two-segment
// K loop is the torch.cat of DropoutNet.py:199-202 ([V | content]): chunks of X1 first, then chunks of X2 against the W
// columns starting at d1; TMA zero-fills past the end of a segment / of W's rows, so no width needs padding to 32.
//   warp 0      TMA producer
//   warp 1      TMEM allocation + MMA issuer (one elected lane): 4 slices x 3 MMAs (M128 N=NB K8, kind::tf32) per chunk
//   warps 2-5   epilogue: thread = one row = one TMEM lane; per K block tcgen05.ld 32 columns at a time into the running sums;
//               after the last block bias / folded BN / activation,
//               row-major stores of Y and (optionally) of its hi / lo split; the row map yrow is the item_emb[cold_idx]
//               scatter of GAR.py:44-46.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kBM = 128;
constexpr int kBK = 32;
constexpr int kThreads = 192;
constexpr int kABytes = kBM * 128;          // one A box: 128 rows x 32 fp32
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 200 * 1024;
constexpr uint32_t kSpinLimit = 1u << 26;
constexpr int kBlockChunks = 2;             // K chunks (x 12 MMAs) accumulated in TMEM before the block is drained into registers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {      // bounded: a protocol bug traps instead of hanging
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32 (both operands K-major SWIZZLE_128B)
__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
#define CR_T_R32(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
        "%26, %27, %28, %29, %30, %31}, [%32];"
        : CR_T_R32(r, 0), CR_T_R32(r, 8), CR_T_R32(r, 16), CR_T_R32(r, 24)
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: 8-row groups 1024 B apart (SBO), version 1, layout type 2.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == CR_ACT_TANH) return tanhf(v);
    if (act == CR_ACT_LEAKY_RELU) return v > 0.f ? v : 0.01f * v;
    return v;
}

struct TowerTcParams {
    int64_t n_rows;
    int n_out;                     // outputs of the layer (grid.y x NB covers them)
    int chunks1, chunks2, d1;      // 32-float K chunks of the two segments; W column where segment 2 starts
    const float* bias; const float* scale; const float* shift; int act;
    float* Y; int64_t ldy; const int32_t* yrow;
    float* Yhi; float* Ylo; int64_t ldh;
    uint32_t idesc; int stages;
    int row_blocks, col_blocks, col_fastest;   // 1-D grid of row_blocks x col_blocks CTAs; which index runs fastest (see the launch)
};

template <int NB, bool RAW>
__global__ void __launch_bounds__(kThreads, 1)
tower_layer_tc_kernel(const __grid_constant__ CUtensorMap mapA1h, const __grid_constant__ CUtensorMap mapA1l,
                      const __grid_constant__ CUtensorMap mapA2h, const __grid_constant__ CUtensorMap mapA2l,
                      const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWl, const TowerTcParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int b_bytes = NB * 128;
    constexpr int stage_bytes = 2 * kABytes + 2 * b_bytes;
    constexpr int kTmemCols = 2 * NB;          // two accumulator stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* full = bars;                     // [stages] TMA -> MMA
    uint64_t* empty = bars + kMaxStages;       // [stages] MMA -> TMA
    uint64_t* tfull = bars + 2 * kMaxStages;   // [2] MMA -> epilogue: a K block is complete in accumulator stage a
    uint64_t* tempty = tfull + 2;              // [2] epilogue -> MMA: stage a has been drained
    uint64_t* conv = tempty + 2;               // [stages] RAW: epilogue warps -> MMA: the stage's A tile has been split into hi / lo
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(conv + kMaxStages);
    const int warp = __shfl_sync(CR_FULL_MASK, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int row_blk = p.col_fastest ? (int)(blockIdx.x / p.col_blocks) : (int)(blockIdx.x % p.row_blocks);
    const int col_blk = p.col_fastest ? (int)(blockIdx.x % p.col_blocks) : (int)(blockIdx.x / p.row_blocks);
    const int64_t m0 = (int64_t)row_blk * kBM;
    const int n0 = col_blk * NB;               // first output column of this CTA
    const int n_chunks = p.chunks1 + p.chunks2;
    const int n_blocks = (n_chunks + kBlockChunks - 1) / kBlockChunks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&conv[s], 4); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int c = 0; c < n_chunks; ++c) {
                const int s = c % p.stages;
                if (c >= p.stages) mbar_wait(&empty[s], ((c / p.stages) - 1) & 1);
                unsigned char* st = smem + (size_t)s * stage_bytes;
                mbar_expect_tx(&full[s], (uint32_t)(RAW ? stage_bytes - kABytes : stage_bytes));
                const bool seg2 = c >= p.chunks1;
                const int kk = (seg2 ? c - p.chunks1 : c) * kBK;
                tma_load_2d(st, seg2 ? &mapA2h : &mapA1h, &full[s], kk, (int)m0);          // RAW: the fp32 tile itself
                if constexpr (!RAW) tma_load_2d(st + kABytes, seg2 ? &mapA2l : &mapA1l, &full[s], kk, (int)m0);
                const int wcol = seg2 ? p.d1 + kk : kk;
                tma_load_2d(st + 2 * kABytes, &mapWh, &full[s], wcol, n0);
                tma_load_2d(st + 2 * kABytes + b_bytes, &mapWl, &full[s], wcol, n0);
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c % p.stages, blk = c / kBlockChunks, a = blk & 1;
            const bool first = c % kBlockChunks == 0, last = (c % kBlockChunks == kBlockChunks - 1) || c == n_chunks - 1;
            if (first && blk >= 2) mbar_wait(&tempty[a], ((blk >> 1) - 1) & 1);
            mbar_wait(RAW ? &conv[s] : &full[s], (c / p.stages) & 1);      // RAW: the converters waited for the TMA bytes of the stage
            tc_fence_after();
            if (leader) {
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t a_hi = smem_desc_sw128(sa), a_lo = smem_desc_sw128(sa + kABytes);
                const uint64_t b_hi = smem_desc_sw128(sa + 2 * kABytes), b_lo = smem_desc_sw128(sa + 2 * kABytes + b_bytes);
                const uint32_t d = tmem_base + a * NB;
#pragma unroll
                for (int k = 0; k < kBK / 8; ++k) {          // a K = 8 slice is 32 bytes further along the swizzled row: +2 in 16-byte units
                    umma_tf32_ss(d, a_hi + 2 * k, b_hi + 2 * k, p.idesc, (!first || k > 0) ? 1u : 0u);
                    umma_tf32_ss(d, a_lo + 2 * k, b_hi + 2 * k, p.idesc, 1u);
                    umma_tf32_ss(d, a_hi + 2 * k, b_lo + 2 * k, p.idesc, 1u);
                }
                umma_commit(&empty[s]);
                if (last) umma_commit(&tfull[a]);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: thread = one row = one TMEM lane; running sums of the K blocks in registers =====
        const int quad = warp & 3;
        const int64_t row = m0 + quad * 32 + lane;
        const bool valid = row < p.n_rows;
        float run[NB];
#pragma unroll
        for (int x = 0; x < NB; ++x) run[x] = 0.f;
        [[maybe_unused]] int cconv = 0;            // RAW: next chunk whose A tile these warps split
        [[maybe_unused]] const int et = (quad * 32 + lane);
        for (int blk = 0; blk < n_blocks; ++blk) {
            if constexpr (RAW) {
                // stay one K block ahead of the drain: the MMA warp works on block blk + 1 while block blk is drained
                const int upto = min(n_chunks, (blk + 2) * kBlockChunks);
                for (; cconv < upto; ++cconv) {
                    const int s = cconv % p.stages;
                    mbar_wait(&full[s], (cconv / p.stages) & 1);
                    float4* hi4 = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
                    float4* lo4 = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + kABytes);
#pragma unroll
                    for (int j = 0; j < kABytes / 16 / 128; ++j) {       // elementwise: the swizzled layout is the same for hi and lo
                        const float4 x = hi4[j * 128 + et];
                        float4 h;
                        h.x = tf32_rna(x.x); h.y = tf32_rna(x.y); h.z = tf32_rna(x.z); h.w = tf32_rna(x.w);
                        hi4[j * 128 + et] = h;
                        lo4[j * 128 + et] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to tcgen05.mma
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&conv[s]);
                }
            }
            const int a = blk & 1;
            mbar_wait(&tfull[a], (blk >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < NB; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + a * NB + c0, r);
#pragma unroll
                for (int x = 0; x < 32; ++x) run[c0 + x] += __uint_as_float(r[x]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[a]);
        }
        if (valid) {
            const int64_t orow = p.yrow ? (int64_t)__ldg(p.yrow + row) : row;
            float* yp = p.Y ? p.Y + orow * p.ldy : nullptr;
            float* hp = p.Yhi ? p.Yhi + row * p.ldh : nullptr;
            float* lp = p.Ylo ? p.Ylo + row * p.ldh : nullptr;
            const bool vec = (p.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.Y) & 15u) == 0);
            const bool vech = (p.ldh % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.Yhi) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(p.Ylo) & 15u) == 0);
#pragma unroll
            for (int x = 0; x < NB; x += 4) {
                const int o0 = n0 + x;
                if (o0 < p.n_out) {
                    float v[4], h[4], l[4];
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const int o = o0 + y;
                        float t = 0.f;
                        if (o < p.n_out) {
                            t = run[x + y] + (p.bias ? __ldg(p.bias + o) : 0.f);
                            if (p.scale) t = t * __ldg(p.scale + o) + __ldg(p.shift + o);
                            t = apply_act(t, p.act);
                        }
                        v[y] = t;
                        h[y] = tf32_rna(t);
                        l[y] = t - h[y];
                    }
                    if (o0 + 4 <= p.n_out) {
                        if (yp) {
                            if (vec) *reinterpret_cast<float4*>(yp + o0) = make_float4(v[0], v[1], v[2], v[3]);
                            else { yp[o0] = v[0]; yp[o0 + 1] = v[1]; yp[o0 + 2] = v[2]; yp[o0 + 3] = v[3]; }
                        }
                        if (hp) {
                            if (vech) {
                                *reinterpret_cast<float4*>(hp + o0) = make_float4(h[0], h[1], h[2], h[3]);
                                *reinterpret_cast<float4*>(lp + o0) = make_float4(l[0], l[1], l[2], l[3]);
                            } else {
#pragma unroll
                                for (int y = 0; y < 4; ++y) { hp[o0 + y] = h[y]; lp[o0 + y] = l[y]; }
                            }
                        }
                    } else {
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                            if (o0 + y < p.n_out) {
                                if (yp) yp[o0 + y] = v[y];
                                if (hp) { hp[o0 + y] = h[y]; lp[o0 + y] = l[y]; }
                            }
                        }
                    }
                }
            }
            // the split tables are padded to ldh columns: the padding must read as zero in the next layer's K loop
            if (hp && n0 + NB >= p.n_out) {
                for (int o = p.n_out; o < p.ldh; ++o) { hp[o] = 0.f; lp[o] = 0.f; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

// x -> (hi, lo): hi = round-to-nearest TF32, lo = x - hi; columns [cols, ld_dst) of the destination rows are zeroed.
__global__ void split_tf32_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows, int cols, float* __restrict__ hi,
                                  float* __restrict__ lo, int64_t ld_dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ld_dst) return;
    const int64_t r = i / ld_dst;
    const int c = (int)(i - r * ld_dst);
    float h = 0.f, l = 0.f;
    if (c < cols) {
        const float x = __ldg(src + r * ld_src + c);
        h = tf32_rna(x);
        l = x - h;
    }
    hi[i] = h;
    lo[i] = l;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode(EncodeTiledFn* out) {
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return cr::note_cuda_error(e, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
        if (qres != cudaDriverEntryPointSuccess || !fn) return cr::note_cuda_error(cudaErrorUnknown, "cuTensorMapEncodeTiled lookup");
        cached = (EncodeTiledFn)fn;
    }
    *out = cached;
    return CR_OK;
}

// rows x cols fp32, row stride ld floats; box = 32 floats x box_rows rows, SWIZZLE_128B, out-of-range elements read as zero
int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn enc;
    int rc = get_encode(&enc);
    if (rc != CR_OK) return rc;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cr::note_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (tower)");
    return CR_OK;
}

inline bool tma_ok(const void* p, int64_t ld) { return p && cr::aligned16(p) && ld > 0 && (ld * 4) % 16 == 0; }

}  // namespace

extern "C" {

int cr_split_tf32(const float* src, int64_t ld_src, int64_t rows, int cols, float* hi, float* lo, int64_t ld_dst, void* stream) {
    if (!src || !hi || !lo || rows < 0 || cols <= 0 || ld_src < cols || ld_dst < cols) return CR_ERR_ARG;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    if (rows == 0) return CR_OK;
    const int64_t total = rows * ld_dst;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, rows, cols, hi, lo, ld_dst);
    CR_LAUNCH_CHECK("split_tf32_kernel");
    return CR_OK;
}

int cr_linear_act_tc_f32(const float* X1hi, const float* X1lo, int64_t ld1, int d1, const float* X2hi, const float* X2lo, int64_t ld2,
                         int d2, int64_t n_rows, const float* Whi, const float* Wlo, int64_t ldw, const float* bias, const float* scale,
                         const float* shift, int n_out, int act, float* Y, int64_t ldy, const int32_t* yrow, float* Yhi, float* Ylo,
                         int64_t ldh, void* stream) {
    const bool raw = X1lo == nullptr;          // plain fp32 rows, split inside the kernel (both segments alike)
    if (!X1hi || !Whi || !Wlo || n_rows < 0 || d1 <= 0 || d2 < 0 || (d2 > 0 && !X2hi) || n_out <= 0) return CR_ERR_ARG;
    if (d2 > 0 && ((X2lo == nullptr) != raw)) return CR_ERR_ARG;
    if ((!Y && !Yhi) || ((Yhi == nullptr) != (Ylo == nullptr)) || (Y && ldy < n_out) || (Yhi && ldh < n_out)) return CR_ERR_ARG;
    if (((scale == nullptr) != (shift == nullptr)) || ldw < d1 + d2 || ld1 < d1 || (d2 > 0 && ld2 < d2)) return CR_ERR_ARG;
    if (act < CR_ACT_NONE || act > CR_ACT_LEAKY_RELU || n_rows > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    if (yrow && Yhi) return CR_ERR_UNSUPPORTED;       // split outputs are written in input row order
    if (!tma_ok(X1hi, ld1) || (!raw && !tma_ok(X1lo, ld1)) || !tma_ok(Whi, ldw) || !tma_ok(Wlo, ldw)) return CR_ERR_ALIGN;
    if (d2 > 0 && (!tma_ok(X2hi, ld2) || (!raw && !tma_ok(X2lo, ld2)))) return CR_ERR_ALIGN;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    if (n_rows == 0) return CR_OK;
    TowerTcParams p{};
    p.n_rows = n_rows; p.n_out = n_out;
    p.chunks1 = (d1 + kBK - 1) / kBK; p.chunks2 = (d2 + kBK - 1) / kBK; p.d1 = d1;
    p.bias = bias; p.scale = scale; p.shift = shift; p.act = act;
    p.Y = Y; p.ldy = ldy; p.yrow = yrow; p.Yhi = Yhi; p.Ylo = Ylo; p.ldh = ldh;
    const int NB = n_out <= 64 ? 64 : 128;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    const int stage_bytes = 2 * kABytes + 2 * NB * 128;
    p.stages = (kSmemBudget - 1024) / stage_bytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
    const int smem = p.stages * stage_bytes + 1024;
    CUtensorMap a1h, a1l, a2h, a2l, wh, wl;
    if ((rc = make_map(&a1h, X1hi, n_rows, d1, ld1, kBM)) != CR_OK) return rc;
    if (raw) a1l = a1h;
    else if ((rc = make_map(&a1l, X1lo, n_rows, d1, ld1, kBM)) != CR_OK) return rc;
    if (d2 > 0) {
        if ((rc = make_map(&a2h, X2hi, n_rows, d2, ld2, kBM)) != CR_OK) return rc;
        if (raw) a2l = a2h;
        else if ((rc = make_map(&a2l, X2lo, n_rows, d2, ld2, kBM)) != CR_OK) return rc;
    } else {
        a2h = a1h; a2l = a1l;
    }
    if ((rc = make_map(&wh, Whi, n_out, d1 + d2, ldw, NB)) != CR_OK) return rc;
    if ((rc = make_map(&wl, Wlo, n_out, d1 + d2, ldw, NB)) != CR_OK) return rc;
    // CTA order.  The CTAs of one row block (one per NB output columns) read the same A tiles; the CTAs of one column block the same
    // W tiles.  Whatever is re-read later must come from L2: a big A (the 450 MB split content table: ncu r02 showed 937 MB of
    // DRAM reads with the row blocks running fastest — every A tile fetched twice) wants its column blocks side by side; a small A
    // (the query block of the kNN path, which fits L2) against a big W wants the row blocks side by side.
    p.row_blocks = (int)((n_rows + kBM - 1) / kBM);
    p.col_blocks = (n_out + NB - 1) / NB;
    const double a_bytes = (raw ? 1.0 : 2.0) * (double)n_rows * (d1 + d2) * 4.0;
    p.col_fastest = a_bytes > 96.0 * 1024 * 1024 ? 1 : 0;
    if ((int64_t)p.row_blocks * p.col_blocks > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)((int64_t)p.row_blocks * p.col_blocks);
#define CR_TOWER(NB_, RAW_)                                                                                                     \
    do {                                                                                                                        \
        CR_CUDA_TRY(cudaFuncSetAttribute(tower_layer_tc_kernel<NB_, RAW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget)); \
        tower_layer_tc_kernel<NB_, RAW_><<<grid, kThreads, smem, (cudaStream_t)stream>>>(a1h, a1l, a2h, a2l, wh, wl, p);        \
    } while (0)
    if (NB == 64) { if (raw) CR_TOWER(64, true); else CR_TOWER(64, false); }
    else { if (raw) CR_TOWER(128, true); else CR_TOWER(128, false); }
#undef CR_TOWER
    CR_LAUNCH_CHECK("tower_layer_tc_kernel");
    return CR_OK;
}

}  // extern "C"
