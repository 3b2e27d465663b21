// K1 (tensor-core path) — fused TF32 scorer on tcgen05: query tiles resident in TMEM, TMA-staged item
// tiles, TMEM accumulators, an in-epilogue threshold filter, exact fp32 rescoring and a proven margin.
//
// Replaces MF.batch_predict (model/MF.py:58-63) + mask writes + torch.topk of _evaluate
// (model/BaseRecommender.py:170-182) for d = 64.  The (B, I) score matrix never exists: scores live
// only in TMEM and registers.
//
// One CTA = one "unit" = 256 queries x one contiguous range of 96-item tiles.  Warp roles (384 threads):
//   warp 0      TMA producer: item tiles (96 items x 64 fp32 = two SWIZZLE_128B boxes) into a 6-stage ring
//   warp 1      MMA issuer: per tile 2 x 8 tcgen05.mma kind::tf32 (M128 N96 K8) with the A operand (queries)
//               read from TENSOR MEMORY — re-reading A from shared memory for every K=8 slice made the
//               first version shared-memory-bandwidth bound (ncu r01: tensor pipe 41 % active)
//   warp 2      mask producer: per tile a 96-bit "do not take" bitmap per query (train items via a monotone
//               cursor in the sorted CSR row, flagged items, items past the end) in shared memory
//   warp 3      owns the TMEM allocation
//   warps 4-11  epilogue: thread = one query = one TMEM lane.  At start each thread stores its own query
//               vector into TMEM (tcgen05.st) — that is the A operand.  Per 32-column chunk: tcgen05.ld
//               (software pipelined one chunk ahead), 3-input max tree, one compare against the query's
//               running threshold (the KSEL-th best approximate score).  Only when some lane beats its
//               threshold does the warp enter the cooperative slow path: the lane's 32 values are
//               transposed through shared memory, each lane tests one column against threshold and bitmap,
//               survivors are appended to the query's candidate buffer (global memory, L2 resident); a
//               full buffer is rank-compacted by the warp.
// TMEM columns: [0,128) queries (2 tiles x 64), then 2 accumulator stages x 2 query tiles x 96 columns.
//
// After the sweep: rescore_kernel re-scores every candidate in exact fp32 (k = 0..63 in order) and keeps the
// top-K by (score desc, id asc) per (query, item split); the splits are merged; verify_kernel proves the
// result: every rejected item has TF32 score <= thr (the largest final threshold over the splits), hence exact
// score <= thr + eps with eps = 2^-8.9 |q| max|x| (TF32 truncation of both operands), so the list is exact
// if its K-th exact score exceeds thr + eps.  Queries that fail the proof are re-run by the exact fp32
// scorer (score_simt.cu) in the same call.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kBM = 256;                // queries per unit (2 x 128-row MMA tiles)
// Tile geometry knobs of the d = 64 instantiation (tools/build_variant.sh): 512 TMEM columns = 2 x 64 (queries) + kAcc x 2 x kBN.
// r02 A/B on one box (TFLOP/s at 10M / 2.5M / 1.25M items per shard): 96 x 2 accumulator stages 674 / 620 / 551;
// 64 x 3 stages 567 / 544 / 510 (the N=64 MMA is less efficient than N=96 and the third stage does not pay it back);
// 32 x 6 stages 347 / 337 / 327.  96 x 2 stays.
#ifndef CR_TC_BN
#define CR_TC_BN 96
#endif
#ifndef CR_TC_ACC
#define CR_TC_ACC 2
#endif
#ifndef CR_TC_STAGES
#define CR_TC_STAGES 6
#endif
#ifndef CR_MASK_STAGES
#define CR_MASK_STAGES 4
#endif
constexpr int kMaskStages = CR_MASK_STAGES;          // mask-bitmap ring (decoupled from the accumulators: ncu r01b showed the
                                        // epilogue waiting 40 % of its time on a bitmap tied to the 2 TMEM stages)
constexpr int kThreads = 384;
constexpr int kEpiWarp0 = 4;            // first epilogue warp
constexpr float kEpsFactor = 2.1e-3f;   // > 2^-9 (1 + 2^-10) + fp32 accumulation slack
constexpr uint32_t kSpinLimit = 1u << 26;

// Geometry of one instantiation.  D = 64: the MF / LightGCN / generator tables (kBN = 96, two accumulator stages, 6-stage
// item ring).  D = 128: VBPR / AMR's concatenated tables (model/VBPR.py:68-75): the two query tiles take 2 x 128 TMEM
// columns, which leaves 2 stages x 2 tiles x 64 accumulator columns; a 64-item tile is 32 KB (four SWIZZLE_128B boxes),
// five of them in flight; 2 x 16 MMAs (M128 N64 K8) per tile.
template <int D>
struct Geo {
    static_assert(D == 64 || D == 128, "instantiated widths");
    static constexpr int kD = D;
    static constexpr int kBN = D == 64 ? CR_TC_BN : 64;          // items per tile
    static constexpr int kAcc = D == 64 ? CR_TC_ACC : 2;         // TMEM accumulator stages
    static constexpr int kStages = D == 64 ? CR_TC_STAGES : 5;   // item smem ring
    static constexpr int kChunks = kBN / 32;                     // 32-column epilogue chunks per tile
    static constexpr int kBoxes = D / 32;                        // SWIZZLE_128B boxes (32 fp32 wide) per item tile
    static constexpr int kChunkBytes = kBN * 128;                // one box: kBN rows x 32 fp32
    static constexpr int kTileBytes = kBoxes * kChunkBytes;
    static constexpr int kTmemA = 0;                             // query tiles: columns [0, 2 D)
    static constexpr int kTmemAcc = 2 * D;                       // accumulators: 2 D + a * 2 kBN + t * kBN
    static_assert(kBN % 32 == 0 && kBN >= 32 && 2 * D + kAcc * 2 * kBN <= 512, "TMEM: 2 D query columns + kAcc x 2 x kBN accumulator columns");
    static_assert(8 * kChunks <= 32, "the mask producer's dirty-word bitmap holds 8 queries x kChunks bits per lane");
    // kind::tf32 instruction descriptor: D=F32, A=B=TF32, both K-major, N=kBN, M=128.
    static constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    struct Smem {
        static constexpr int kB = 0;                                       // kStages x kTileBytes
        static constexpr int kMask = kB + kStages * kTileBytes;            // [kMaskStages][kChunks][256 queries] u32
        static constexpr int kCommon = kMask + kMaskStages * kChunks * kBM * 4;   // [kMaskStages][kChunks] u32 (padded to 64 B)
        static constexpr int kDirty = kCommon + ((kMaskStages * kChunks * 4 + 63) / 64) * 64;                        // [kMaskStages][32 lanes] u32: words a lane must clear
        static constexpr int kScratch = kDirty + kMaskStages * 32 * 4;     // 8 warps x 32 floats
        static constexpr int kBars = kScratch + 8 * 128;
        static constexpr int kTotal = kBars + 512;
        static_assert(kTotal <= 227 * 1024, "shared memory per CTA");
    };
};

struct Cand {
    float s;
    int p;   // position in the local item table
};

// ---------------------------------------------------------------------------------------------- PTX
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
// Bounded wait: a protocol bug becomes a trap (launch failure) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]^T, kind::tf32
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#define CR_R32(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
#define CR_I32(r, o) "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7])
#define CR_LIST32                                                                                                           \
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, " \
    "%26, %27, %28, %29, %30, %31}"
// issue only; the caller waits with tmem_ld_wait() before touching r[]
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " CR_LIST32 ", [%32];"
                 : CR_R32(r, 0), CR_R32(r, 8), CR_R32(r, 16), CR_R32(r, 24)
                 : "r"(taddr)
                 : "memory");
}
#define CR_RW32(r, o) "+r"(r[o + 0]), "+r"(r[o + 1]), "+r"(r[o + 2]), "+r"(r[o + 3]), "+r"(r[o + 4]), "+r"(r[o + 5]), "+r"(r[o + 6]), "+r"(r[o + 7])
// wait for the outstanding tcgen05.ld; r[] is an in/out operand so no use of it can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : CR_RW32(r, 0), CR_RW32(r, 8), CR_RW32(r, 16), CR_RW32(r, 24)::"memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], " CR_LIST32 ";" ::CR_I32(r, 0), CR_I32(r, 8), CR_I32(r, 16),
                 CR_I32(r, 24), "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// max of 32 registers as a depth-4 tree of 3-input max (17 ALU ops, no long dependency chain)
__device__ __forceinline__ float max32(const uint32_t (&r)[32]) {
    float l1[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) l1[i] = max3(__uint_as_float(r[3 * i]), __uint_as_float(r[3 * i + 1]), __uint_as_float(r[3 * i + 2]));
    l1[10] = fmaxf(__uint_as_float(r[30]), __uint_as_float(r[31]));
    const float a = max3(l1[0], l1[1], l1[2]), b = max3(l1[3], l1[4], l1[5]), c = max3(l1[6], l1[7], l1[8]);
    const float d = fmaxf(l1[9], l1[10]);
    return fmaxf(max3(a, b, c), d);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 8-row groups are
// 1024 B apart (SBO), LBO unused for swizzled K-major, version 1, layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// ---------------------------------------------------------------------------------------------- sweep
// Diagnostic timeline (CR_TC_DEBUG_MODE & 8): %globaltimer stamps of epilogue warp 0 of every unit, read back by
// cr_debug_tc_timeline.  Slots: 0 entry, 1 setup done, 2 queries in TMEM, 3.. after tile 0, 15, 127, 1023, 4095, 8191, last, 10 exit (ns);
// 11-15 blocked SM cycles: epilogue warp 0 on tfull / on mfull, MMA issuer on full (TMA) / on tempty (epilogue), mask producer on mempty.
constexpr int kTlSlots = 16, kTlUnits = 8192;
__device__ unsigned long long g_tc_timeline[kTlUnits * kTlSlots];
__device__ __forceinline__ unsigned long long globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define CR_TL(slot)                                                                                         \
    do {                                                                                                    \
        if constexpr (DBG) {                                                                                \
            if ((p.dbg_mode & 8) && threadIdx.x == kEpiWarp0 * 32 && blockIdx.x < kTlUnits)                 \
                g_tc_timeline[blockIdx.x * kTlSlots + (slot)] = globaltimer();                              \
        }                                                                                                   \
    } while (0)

// DBG builds account the cycles a role spends blocked on each barrier (slots 11..15 of the unit's timeline row)
#define CR_WAIT(acc, bar, parity)                                     \
    do {                                                              \
        if constexpr (DBG) {                                          \
            const long long _t0 = clock64();                          \
            mbar_wait(bar, parity);                                   \
            acc += clock64() - _t0;                                   \
        } else {                                                      \
            mbar_wait(bar, parity);                                   \
        }                                                             \
    } while (0)
#define CR_TL_PUT(slot, value)                                                                              \
    do {                                                                                                    \
        if constexpr (DBG) {                                                                                \
            if ((p.dbg_mode & 8) && lane == 0 && blockIdx.x < kTlUnits)                                     \
                g_tc_timeline[blockIdx.x * kTlSlots + (slot)] = (unsigned long long)(value);                \
        }                                                                                                   \
    } while (0)

struct SweepParams {
    const float* Q;          // [n_q, 64] gathered query vectors
    int64_t n_q;             // valid queries
    int n_q_pad;             // n_utiles * 256
    int n_utiles;
    int64_t n_items;
    int tiles_per_split;
    int n_tiles;             // ceil(n_items / kBN)
    const int32_t* item_gids;
    int64_t item_id_base;
    const int64_t* mask_rowptr;
    const int32_t* mask_col;
    const uint8_t* item_flags;
    uint8_t flag_exclude;
    const uint32_t* bad_bits; // [n_tiles][kChunks] "never take" bit per item (flagged, or past the end of the table), precomputed
                              // once per call when item flags are tested (item_bad_bits_kernel); nullptr: only the table tail
    Cand* buf;               // [S][n_q_pad][CAP]
    int* cnt;                // [S][n_q_pad]
    float* thr;              // [S][n_q_pad]
    float* dbg_scores;       // optional: raw TF32 scores of the first 256 x 96 block (probe)
    int seed_tiles;          // threshold seed phase: the first seed_tiles tiles are swept twice (see the kernel)
    int dbg_mode;            // timing experiments only (env CR_TC_DEBUG_MODE): 1 = skip TMEM loads, 2 = skip MMAs, 4 = interleave
};

// Rank-compact one query's candidate buffer to its best KSEL entries, sorted; returns the KSEL-th score.
template <int EPL>
__device__ float warp_shrink(Cand* buf, int cnt, int ksel, int lane) {
    Cand e[EPL];
    int rank[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        const int idx = lane + 32 * i;
        e[i] = Cand{-CUDART_INF_F, 0x7fffffff};
        if (idx < cnt) {
            const float2 raw = __ldcg(reinterpret_cast<const float2*>(buf) + idx);
            e[i] = Cand{raw.x, __float_as_int(raw.y)};
        }
        rank[i] = 0;
    }
    for (int j = 0; j < 32; ++j) {
#pragma unroll
        for (int i2 = 0; i2 < EPL; ++i2) {
            const float s = __shfl_sync(CR_FULL_MASK, e[i2].s, j);
            const int p = __shfl_sync(CR_FULL_MASK, e[i2].p, j);
#pragma unroll
            for (int i = 0; i < EPL; ++i) rank[i] += cr::better(s, p, e[i].s, e[i].p) ? 1 : 0;
        }
    }
    float thr = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        const bool keep = (lane + 32 * i < cnt) && rank[i] < ksel;
        if (keep) reinterpret_cast<float2*>(buf)[rank[i]] = make_float2(e[i].s, __int_as_float(e[i].p));
        const unsigned who = __ballot_sync(CR_FULL_MASK, keep && rank[i] == ksel - 1);
        if (who) thr = __shfl_sync(CR_FULL_MASK, e[i].s, __ffs(who) - 1);
    }
    __syncwarp();
    return thr;
}

// End of the seed phase: every lane's query gets the KSEL-th largest of its T0 tile maxima (ties broken by tile index), one
// ulp lower so that the item that set the bound still passes "score > thr" (at least KSEL items do).  Each lane passes the
// base of ITS query's buffer; lane L's buffer is ranked by the whole warp.  Once per unit: kept out of line.
template <int KSEL>
__device__ __noinline__ float seed_threshold(const float* mine, int cap, int T0, int lane, bool valid) {
    float thr = CUDART_INF_F;                    // padding lanes never take the slow path
    __syncwarp();
    for (int L = 0; L < 32; ++L) {
        const float* sb = mine + (int64_t)(L - lane) * cap * 2;      // cap Cand entries = 2*cap floats per query
        float ev4[4];
        int rk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = lane + 32 * k;
            ev4[k] = idx < T0 ? __ldcg(sb + idx) : -CUDART_INF_F;
            rk[k] = 0;
        }
        for (int j = 0; j < 32; ++j) {
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
                const float sv = __shfl_sync(CR_FULL_MASK, ev4[k2], j);
#pragma unroll
                for (int k = 0; k < 4; ++k) rk[k] += cr::better(sv, j + 32 * k2, ev4[k], lane + 32 * k) ? 1 : 0;
            }
        }
        float t0v = -CUDART_INF_F;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned who = __ballot_sync(CR_FULL_MASK, rk[k] == KSEL - 1);
            if (who) t0v = __shfl_sync(CR_FULL_MASK, ev4[k], __ffs(who) - 1);
        }
        if (lane == L && valid) thr = t0v > -CUDART_INF_F ? nextafterf(t0v, -CUDART_INF_F) : -CUDART_INF_F;
    }
    __syncwarp();
    return thr;
}

// DBG = true compiles the probe / timing hooks (dbg_scores, dbg_mode, timeline) in; the production instantiation has none
// of them: the hot loops of the four warp roles must stay inside the 32 KB L1.5 instruction cache (B300_MICROARCH.md) —
// an earlier seed-phase variant that grew the kernel from 43 KB to 57 KB of SASS ran 14 % slower with identical hot loops.
template <int D, int KSEL, bool DBG>
__global__ void __launch_bounds__(kThreads, 1) score_sweep_tc_kernel(const __grid_constant__ CUtensorMap map_i, const SweepParams p) {
    using G = Geo<D>;
    using SmemLayout = typename G::Smem;
    constexpr int kD = G::kD, kBN = G::kBN, kAcc = G::kAcc, kStages = G::kStages, kChunks = G::kChunks, kChunkBytes = G::kChunkBytes,
                  kTileBytes = G::kTileBytes, kTmemA = G::kTmemA, kTmemAcc = G::kTmemAcc;
    constexpr uint32_t kIdesc = G::kIdesc;
    constexpr int CAP = KSEL + 32;
    constexpr int EPL = CAP / 32;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem + SmemLayout::kB;
    uint32_t* sMask = reinterpret_cast<uint32_t*>(smem + SmemLayout::kMask);
    uint32_t* sCommon = reinterpret_cast<uint32_t*>(smem + SmemLayout::kCommon);
    float* sScratch = reinterpret_cast<float*>(smem + SmemLayout::kScratch);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SmemLayout::kBars);
    uint64_t* full = bars;                      // [kStages]  TMA -> MMA
    uint64_t* empty = bars + kStages;           // [kStages]  MMA -> TMA
    uint64_t* tfull = bars + 2 * kStages;       // [kAcc]     MMA -> epilogue
    uint64_t* tempty = tfull + kAcc;            // [kAcc]     epilogue -> MMA
    uint64_t* mfull = tempty + kAcc;            // [kMaskStages] mask producer -> epilogue
    uint64_t* mempty = mfull + kMaskStages;     // [kMaskStages] epilogue -> mask producer
    uint64_t* aready = mempty + kMaskStages;    // [1]        query tiles stored in TMEM (8 epilogue warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aready + 1);

    const int warp = __shfl_sync(CR_FULL_MASK, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const int utile = blockIdx.x % p.n_utiles, split = blockIdx.x / p.n_utiles;
    const int tile_begin = split * p.tiles_per_split;
    const int tile_end = min(p.n_tiles, tile_begin + p.tiles_per_split);
    const int n_local = tile_end - tile_begin;
    // Threshold seed phase.  A query's threshold starts at -inf, so the first ~10^5 items of every sweep flood the slow
    // path: the per-unit timeline (CR_TC_DEBUG_MODE=8, profiles/r01_sweep_timeline.txt) shows 31 us per tile over the
    // first 16 tiles, 7 us up to tile 127, 1.7 us up to tile 1023 against 0.66 us in steady state — 2.1 ms lost per
    // unit, whatever the length of the sweep.  So the first T0 tiles are first swept in "seed mode": the epilogue only
    // records each query's best unmasked score per tile; T0 distinct items reach the KSEL-th largest of those T0 tile
    // maxima, so it is a valid lower bound of the query's KSEL-th best score, and the real sweep — which starts over at
    // tile 0 — begins with a threshold that is already tight.  Virtual tile v = tile v (seed, v < T0) or v - T0 (sweep).
    const int T0 = (p.seed_tiles > 0 && n_local >= 8 * p.seed_tiles) ? p.seed_tiles : 0;
    // The seed tiles are spread evenly over the unit's item range (every seed_stride-th tile), not its first T0 tiles: on a
    // table ordered by popularity or norm the head of the range says nothing about the scores to come (r02: a table sorted by
    // ascending norm ran at 0.66 of the unsorted one with head seeding, the thresholds kept rising to the last tile).
#ifdef CR_TC_SEED_HEAD                  // (A/B knob: seed from the first T0 tiles, as r01 did)
    const int seed_stride = 1;
#else
    const int seed_stride = T0 > 0 ? n_local / T0 : 1;
#endif
    const int n_virtual = n_local + T0;
    CR_TL(0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < kAcc; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        for (int m = 0; m < kMaskStages; ++m) { mbar_init(&mfull[m], 1); mbar_init(&mempty[m], 8); }
        mbar_init(aready, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 3) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    CR_TL(1);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int i = 0; i < n_virtual; ++i) {
                const int s = i % kStages;
                if (i >= kStages) mbar_wait(&empty[s], ((i / kStages) - 1) & 1);
                mbar_expect_tx(&full[s], kTileBytes);
                const int row = (tile_begin + (i < T0 ? i * seed_stride : i - T0)) * kBN;
#pragma unroll
                for (int bx = 0; bx < G::kBoxes; ++bx) tma_load_2d(sB + s * kTileBytes + bx * kChunkBytes, &map_i, &full[s], bx * 32, row);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop (uniform control flow), one elected lane issues =====
        mbar_wait(aready, 0);
        tc_fence_after();
        const bool leader = elect_one();
        [[maybe_unused]] long long w_tempty = 0, w_full = 0;
        // B descriptor = constant high word | (start address >> 4): per MMA only the low word moves
        constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
        for (int i = 0; i < n_virtual; ++i) {
            const int s = i % kStages, a = i % kAcc;
            if (i >= kAcc) CR_WAIT(w_tempty, &tempty[a], ((i / kAcc) - 1) & 1);
            CR_WAIT(w_full, &full[s], (i / kStages) & 1);
            tc_fence_after();
            if (leader) {
                const uint32_t dlo = ((smem_u32(sB + s * kTileBytes) >> 4) & 0x3FFF) | (1u << 16);
                const uint32_t d0 = tmem_base + kTmemAcc + a * (2 * kBN);
                if (!DBG || !(p.dbg_mode & 2)) {
#pragma unroll
                    for (int m = 0; m < 2 * (kD / 8); ++m) {
                        const int t = m / (kD / 8), k = m % (kD / 8);
                        const uint32_t off16 = ((k >> 2) * kChunkBytes + (k & 3) * 32) >> 4;
                        const uint64_t bdesc = ((uint64_t)kDescHi << 32) | (uint64_t)(dlo + off16);
                        umma_tf32_ts(d0 + t * kBN, tmem_base + kTmemA + t * kD + k * 8, bdesc, kIdesc, k > 0 ? 1u : 0u);
                    }
                }
                umma_commit(&empty[s]);
                umma_commit(&tfull[a]);
            }
            __syncwarp();
        }
        CR_TL_PUT(13, w_full);
        CR_TL_PUT(14, w_tempty);
    } else if (warp == 2) {
        // ===== mask producer: lane owns queries lane + 32*j, j = 0..7 =====
        // nxt[j] = next train item of the query, nx2[j] = the one after it, loaded one consumption EARLY: advancing a
        // cursor must not wait for a global load.  (With the load issued at the advance, every tile holding a train item
        // stalled this warp for a DRAM round trip per query; at 100 train items x 256 queries per unit that was a fixed
        // ~3 ms per unit however long the sweep — 4 % of a 10M-item sweep, 25 % of a 1.25M-item shard.)
        int cur[8], endp[8], nxt[8], nx2[8];
        int64_t rlo[8];
        uint32_t* sDirty = reinterpret_cast<uint32_t*>(smem + SmemLayout::kDirty);
        for (int w = lane; w < kMaskStages * kChunks * kBM; w += 32) sMask[w] = 0;     // bitmaps start clean and are
        for (int m = 0; m < kMaskStages; ++m) sDirty[m * 32 + lane] = 0;               // cleaned lazily afterwards
        __syncwarp();
        // Cursor setup: first train item at or after the first global id of this split.  The eight binary searches of a lane
        // advance together (eight independent loads per step instead of 8 x 7 dependent DRAM round trips, ~50 us per unit).
        int cur0[8];
        {
            int64_t lo8[8], hi8[8];
            const int64_t pos_first = (int64_t)tile_begin * kBN;
            const int first_gid = (p.item_gids && n_local > 0) ? __ldg(p.item_gids + pos_first) : (int)(p.item_id_base + pos_first);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t q = (int64_t)utile * kBM + lane + 32 * j;
                lo8[j] = hi8[j] = 0; rlo[j] = 0; endp[j] = 0;
                if (p.mask_rowptr && q < p.n_q && n_local > 0) {
                    lo8[j] = p.mask_rowptr[q]; hi8[j] = p.mask_rowptr[q + 1];
                    rlo[j] = lo8[j]; endp[j] = (int)(hi8[j] - lo8[j]);
                }
            }
            bool more = true;
            while (more) {
                more = false;
                int v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = lo8[j] < hi8[j] ? __ldg(p.mask_col + ((lo8[j] + hi8[j]) >> 1)) : 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (lo8[j] < hi8[j]) {
                        const int64_t mid = (lo8[j] + hi8[j]) >> 1;
                        if (v[j] < first_gid) lo8[j] = mid + 1; else hi8[j] = mid;
                        more |= lo8[j] < hi8[j];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) cur0[j] = (int)(lo8[j] - rlo[j]);
        }
        [[maybe_unused]] long long w_mempty = 0;
        const bool plain = !p.item_gids && !p.bad_bits;
        uint32_t bits_next = (p.bad_bits && lane < kChunks && n_local > 0) ? __ldg(p.bad_bits + (int64_t)tile_begin * kChunks + lane) : 0u;
        // two passes over the tiles when the seed phase is on: [0, T0) in seed mode, then the whole sweep from tile 0
        for (int pass = (T0 > 0 ? 0 : 1), i = 0; pass < 2; ++pass) {
            const int n_pass = pass == 0 ? T0 : n_local;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                cur[j] = cur0[j];
                nxt[j] = (cur[j] < endp[j]) ? __ldg(p.mask_col + rlo[j] + cur[j]) : 0x7fffffff;
                nx2[j] = (cur[j] + 1 < endp[j]) ? __ldg(p.mask_col + rlo[j] + cur[j] + 1) : 0x7fffffff;
            }
            for (int k = 0; k < n_pass; ++k, ++i) {
                const int a = i % kMaskStages;
                if (i >= kMaskStages) CR_WAIT(w_mempty, &mempty[a], ((i / kMaskStages) - 1) & 1);
                uint32_t* mk = sMask + a * kChunks * kBM;
                {   // clear only the words this lane set the last time the stage was used (bit b -> chunk b%3, query lane+32*(b/3))
                    uint32_t dm = sDirty[a * 32 + lane];
                    while (dm) {
                        const int b = __ffs(dm) - 1;
                        dm &= dm - 1;
                        mk[(b % kChunks) * kBM + lane + 32 * (b / kChunks)] = 0;
                    }
                }
                uint32_t dirty = 0;
                const int tstep = pass == 0 ? seed_stride : 1;       // seed pass: every seed_stride-th tile (cursors skip what lies between)
                const int64_t pos0 = (int64_t)(tile_begin + k * tstep) * kBN;
                int gid_lo = 0, gid_hi = 0;   // global id range covered by this tile: [gid_lo, gid_hi]
                // Items nobody may take (flagged by the warm/cold setting, or past the end of the table): one word per
                // 32-item chunk, the same for every query.  With item flags the words come from the per-call table
                // (item_bad_bits_kernel) and are fetched ONE TILE AHEAD: r02 measured the former in-loop version — a flag byte
                // per item loaded, balloted and shuffled by this warp inside every tile — at 0.54 of the unflagged sweep.
                if (plain && pos0 + kBN <= p.n_items) {      // common case: contiguous ids, no flags, full tile
                    gid_lo = (int)(p.item_id_base + pos0);
                    gid_hi = gid_lo + kBN - 1;
                    if (lane < kChunks) sCommon[a * kChunks + lane] = 0;
                } else {
                    if (p.bad_bits) {
                        if (lane < kChunks) sCommon[a * kChunks + lane] = bits_next;
                        const int kn = (k + 1 < n_pass) ? (k + 1) * tstep : 0;   // (the sweep pass restarts at tile 0 after the seed pass)
                        if (lane < kChunks) bits_next = __ldg(p.bad_bits + (int64_t)(tile_begin + kn) * kChunks + lane);
                    } else if (lane < kChunks) {
                        const int64_t left = p.n_items - (pos0 + lane * 32);
                        sCommon[a * kChunks + lane] = left >= 32 ? 0u : (left <= 0 ? 0xffffffffu : ~((1u << (int)left) - 1u));
                    }
                    const int n_valid = (int)min((int64_t)kBN, p.n_items - pos0);
                    if (!p.item_gids) {           // contiguous ids
                        gid_lo = (int)(p.item_id_base + pos0);
                        gid_hi = gid_lo + n_valid - 1;
                    } else {                      // compacted table: the ids of its first and last row in this tile
                        gid_lo = __ldg(p.item_gids + pos0);
                        gid_hi = __ldg(p.item_gids + pos0 + n_valid - 1);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    while (nxt[j] <= gid_hi) {
                        int pos = -1;
                        if (!p.item_gids) {
                            pos = nxt[j] - gid_lo;
                        } else if (nxt[j] >= gid_lo) {   // locate the id inside this tile of the compacted table
                            int lo2 = 0, hi2 = (int)min((int64_t)kBN, p.n_items - pos0);
                            while (lo2 < hi2) {
                                const int mid = (lo2 + hi2) >> 1;
                                if (__ldg(p.item_gids + pos0 + mid) < nxt[j]) lo2 = mid + 1; else hi2 = mid;
                            }
                            if (pos0 + lo2 < p.n_items && __ldg(p.item_gids + pos0 + lo2) == nxt[j]) pos = lo2;
                        }
                        if (pos >= 0 && pos < kBN) {
                            mk[(pos >> 5) * kBM + lane + 32 * j] |= 1u << (pos & 31);
                            dirty |= 1u << (j * kChunks + (pos >> 5));
                        }
                        ++cur[j];
                        nxt[j] = nx2[j];
                        nx2[j] = (cur[j] + 1 < endp[j]) ? __ldg(p.mask_col + rlo[j] + cur[j] + 1) : 0x7fffffff;
                    }
                }
                sDirty[a * 32 + lane] = dirty;
                __syncwarp();
                if (lane == 0) mbar_arrive(&mfull[a]);
            }
        }
        CR_TL_PUT(15, w_mempty);
    } else if (warp >= kEpiWarp0) {
        // ===== epilogue: thread = one query = one TMEM lane =====
        const int e = warp - kEpiWarp0;
        const int t = e >> 2, quad = warp & 3;
        const int ulocal = t * 128 + quad * 32 + lane;               // query index inside the unit
        const int64_t q = (int64_t)utile * kBM + ulocal;
        const bool valid = q < p.n_q;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        {   // A operand: this thread's query vector -> TMEM lane (quad*32 + lane), columns [t*64, t*64+64)
            uint32_t r[32];
#pragma unroll
            for (int h = 0; h < kD / 32; ++h) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid) v = __ldg(reinterpret_cast<const float4*>(p.Q + q * kD + h * 32) + k);
                    r[4 * k] = __float_as_uint(v.x); r[4 * k + 1] = __float_as_uint(v.y);
                    r[4 * k + 2] = __float_as_uint(v.z); r[4 * k + 3] = __float_as_uint(v.w);
                }
                tmem_st32(tmem_base + lane_addr + kTmemA + t * kD + h * 32, r);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(aready);
        }
        CR_TL(2);
        // seed phase: thr = +inf keeps every lane off the slow path while the tile maxima are collected
        float thr = (valid && T0 == 0) ? -CUDART_INF_F : CUDART_INF_F;
        float tmax = -CUDART_INF_F;
        [[maybe_unused]] long long w_tfull = 0, w_mfull = 0;
        int cnt = 0;
        Cand* mybuf = p.buf + ((int64_t)split * p.n_q_pad + utile * kBM + ulocal) * CAP;
        float* scratch = sScratch + e * 32;

        for (int i = 0; i < n_virtual; ++i) {
            const int a = i % kAcc;
            CR_WAIT(w_tfull, &tfull[a], (i / kAcc) & 1);
            const int ms = i % kMaskStages;
            CR_WAIT(w_mfull, &mfull[ms], (i / kMaskStages) & 1);
            tc_fence_after();
            const bool seeding = i < T0;
            const int64_t pos0 = (int64_t)(tile_begin + (seeding ? i * seed_stride : i - T0)) * kBN;
            const uint32_t acc_addr = tmem_base + lane_addr + kTmemAcc + a * (2 * kBN) + t * kBN;
            uint32_t rbuf[2][32];
            if constexpr (DBG) {
                if (p.dbg_mode & 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(&tempty[a]); mbar_arrive(&mempty[ms]); }
                    continue;
                }
            }
            tmem_ld32_issue(acc_addr, rbuf[0]);
#pragma unroll
            for (int c = 0; c < kChunks; ++c) {
                uint32_t (&r)[32] = rbuf[c & 1];
                tmem_ld_wait(r);
                if (c + 1 < kChunks) tmem_ld32_issue(acc_addr + (c + 1) * 32, rbuf[(c + 1) & 1]);   // next chunk in flight
                if constexpr (DBG) {
                    if (p.dbg_scores && blockIdx.x == 0 && i == T0) {
#pragma unroll
                        for (int x = 0; x < 32; ++x) p.dbg_scores[ulocal * kBN + c * 32 + x] = __uint_as_float(r[x]);
                    }
                }
                float m = max32(r);
                if (seeding) {      // (warp-uniform) seed mode: this query's best UNMASKED score of the tile
                    const uint32_t bad = sMask[(ms * kChunks + c) * kBM + ulocal] | sCommon[ms * kChunks + c];
#ifndef CR_TC_SEED_MASKED_MAX
                    if (bad) m = -CUDART_INF_F;      // a chunk holding a masked item is skipped: the bound stays valid, barely weaker
#else                                            // (A/B knob) seed from the unmasked columns of such a chunk instead.  r02, one box:
                    if (bad) {                   // +1-4 % on the in-kernel warm / cold settings, -1 % on the unflagged headline case
                        m = -CUDART_INF_F;       // (136 more instructions next to the hot loop) — not worth it, off by default
#pragma unroll
                        for (int x = 0; x < 32; ++x) m = ((bad >> x) & 1u) ? m : fmaxf(m, __uint_as_float(r[x]));
                    }
#endif
                    tmax = fmaxf(tmax, m);
                }
                unsigned ev = __ballot_sync(CR_FULL_MASK, m > thr);
                while (ev) {
                    // slow path (rare): one winning lane per iteration; its 32 values are transposed through shared
                    // memory so that every lane tests one column.  (A lane-local scan of the 32 registers was tried and
                    // was 17-35 % slower: 32 predicated compare/append steps cost more issue slots than this.)
                    const int L = __ffs(ev) - 1;
                    ev &= ev - 1;
                    if (lane == L) {
                        uint4* dst = reinterpret_cast<uint4*>(scratch);
#pragma unroll
                        for (int x = 0; x < 8; ++x) dst[x] = make_uint4(r[4 * x], r[4 * x + 1], r[4 * x + 2], r[4 * x + 3]);
                    }
                    __syncwarp();
                    const float x = scratch[lane];
                    float thrL = __shfl_sync(CR_FULL_MASK, thr, L);
                    int cntL = __shfl_sync(CR_FULL_MASK, cnt, L);
                    const int uL = t * 128 + quad * 32 + L;
                    const uint32_t bad = sMask[(ms * kChunks + c) * kBM + uL] | sCommon[ms * kChunks + c];
                    const bool ok = !((bad >> lane) & 1u);
                    bool pass = ok && x > thrL;
                    unsigned pm = __ballot_sync(CR_FULL_MASK, pass);
                    Cand* bufL = mybuf + (int64_t)(L - lane) * CAP;
                    if (cntL + __popc(pm) > CAP) {
                        thrL = warp_shrink<EPL>(bufL, cntL, KSEL, lane);
                        cntL = KSEL;
                        pass = ok && x > thrL;
                        pm = __ballot_sync(CR_FULL_MASK, pass);
                    }
                    if (pass) {
                        const int slot = cntL + __popc(pm & ((1u << lane) - 1u));
                        reinterpret_cast<float2*>(bufL)[slot] = make_float2(x, __int_as_float((int)(pos0 + c * 32 + lane)));
                    }
                    if (lane == L) {
                        cnt = cntL + __popc(pm);
                        thr = thrL;
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&tempty[a]); mbar_arrive(&mempty[ms]); }
            if (seeding) {
                // the tile maxima live in the query's (still empty) candidate buffer: 2*CAP floats >= T0
                reinterpret_cast<float*>(mybuf)[i] = valid ? tmax : -CUDART_INF_F;
                tmax = -CUDART_INF_F;
                if (i == T0 - 1) thr = seed_threshold<KSEL>(reinterpret_cast<const float*>(mybuf), CAP, T0, lane, valid);
            }
            if constexpr (DBG) {
                if (p.dbg_mode & 8) {
                    const int v = i - T0;
                    const int slot = v == 0 ? 3 : v == 15 ? 4 : v == 127 ? 5 : v == 1023 ? 6 : v == 4095 ? 7 : v == 8191 ? 8 : v == n_local - 1 ? 9 : -1;
                    if (slot >= 0) CR_TL(slot);
                }
            }
        }
        if (e == 0) { CR_TL_PUT(11, w_tfull); CR_TL_PUT(12, w_mfull); }
        const int64_t o = (int64_t)split * p.n_q_pad + utile * kBM + ulocal;
        p.cnt[o] = valid ? cnt : 0;
        p.thr[o] = thr;
    }
    tc_fence_before();
    __syncthreads();
    CR_TL(10);
    if (warp == 3) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------- rescore / verify
struct RescoreParams {
    const float* Q;              // [n_q, 64] gathered queries
    const float* item_tab;
    const int32_t* item_gids; int64_t item_id_base;
    int64_t n_q; int n_q_pad; int n_splits; int K; int cap;
    const Cand* buf; const int* cnt; const float* thr;
    const float* item_norm2_max;  // device scalar: max_i |x_i|^2
    float* part_score; int32_t* part_id;   // [S][n_q][K]
    const float* out_score;      // merged [n_q][K] (verify)
    int32_t* refine_list; int32_t* refine_count;
};

// One warp per (query, split): exact fp32 scores of the candidates, top-K by (score desc, gid asc).
// Candidates are never masked items (the sweep's bitmap dropped those).
template <int EPL, int kD>
__global__ void __launch_bounds__(256) rescore_kernel(const RescoreParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= p.n_q * p.n_splits) return;
    const int split = (int)(w / p.n_q);
    const int64_t q = w % p.n_q;
    const int64_t o = (int64_t)split * p.n_q_pad + q;
    const int cnt = min(p.cnt[o], p.cap);
    const Cand* buf = p.buf + o * p.cap;
    const float4* qv = reinterpret_cast<const float4*>(p.Q + q * kD);

    float es[EPL];
    int eg[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        const int idx = lane + 32 * i;
        es[i] = -CUDART_INF_F;
        eg[i] = 0x7fffffff;
        if (idx < cnt) {
            const int pos = buf[idx].p;
            const float4* xv = reinterpret_cast<const float4*>(p.item_tab + (int64_t)pos * kD);
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < kD / 4; ++k) {
                const float4 a = __ldg(qv + k), b = __ldg(xv + k);
                s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
            }
            es[i] = s;
            eg[i] = p.item_gids ? __ldg(p.item_gids + pos) : (int)(p.item_id_base + pos);
        }
    }
    int rank[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) rank[i] = 0;
    for (int j = 0; j < 32; ++j) {
#pragma unroll
        for (int i2 = 0; i2 < EPL; ++i2) {
            const float s = __shfl_sync(CR_FULL_MASK, es[i2], j);
            const int g = __shfl_sync(CR_FULL_MASK, eg[i2], j);
#pragma unroll
            for (int i = 0; i < EPL; ++i) rank[i] += cr::better(s, g, es[i], eg[i]) ? 1 : 0;
        }
    }
    const int K = p.K;
    float* os = p.part_score + ((int64_t)split * p.n_q + q) * K;
    int32_t* oi = p.part_id + ((int64_t)split * p.n_q + q) * K;
    for (int k = cnt + lane; k < K; k += 32) { os[k] = -CUDART_INF_F; oi[k] = -1; }
#pragma unroll
    for (int i = 0; i < EPL; ++i)
        if (lane + 32 * i < cnt && rank[i] < K) { os[rank[i]] = es[i]; oi[rank[i]] = eg[i]; }
}

// One thread per query, after the merge: prove the list or queue the query for the exact re-run.
template <int kD>
__global__ void verify_kernel(const RescoreParams p) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= p.n_q) return;
    float thr = -CUDART_INF_F;
    for (int s = 0; s < p.n_splits; ++s) thr = fmaxf(thr, p.thr[(int64_t)s * p.n_q_pad + q]);
    if (thr == -CUDART_INF_F) return;          // nothing was ever rejected by a threshold
    const float4* qv = reinterpret_cast<const float4*>(p.Q + q * kD);
    float qn2 = 0.f;
#pragma unroll
    for (int k = 0; k < kD / 4; ++k) {
        const float4 a = __ldg(qv + k);
        qn2 = fmaf(a.x, a.x, qn2); qn2 = fmaf(a.y, a.y, qn2); qn2 = fmaf(a.z, a.z, qn2); qn2 = fmaf(a.w, a.w, qn2);
    }
    const float eps = kEpsFactor * sqrtf(qn2) * sqrtf(*p.item_norm2_max);
    const float kth = p.out_score[q * p.K + p.K - 1];      // -inf if the merged list is short
    if (!(kth > thr + eps)) p.refine_list[atomicAdd(p.refine_count, 1)] = (int32_t)q;
}

template <int kD>
__global__ void item_norm_max_kernel(const float4* __restrict__ item4, int64_t n_items, float* out) {
    float m = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_items; r += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kD / 4; ++k) {
            const float4 a = __ldg(item4 + r * (kD / 4) + k);
            s = fmaf(a.x, a.x, s); s = fmaf(a.y, a.y, s); s = fmaf(a.z, a.z, s); s = fmaf(a.w, a.w, s);
        }
        m = fmaxf(m, s);
    }
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(CR_FULL_MASK, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // m >= 0: int order == float order
}

// One warp per (tile, 32-item chunk): bit l = item (tile * kBN + chunk * 32 + l) must never be taken — it carries the
// excluded flag, or lies past the end of the table.
__global__ void item_bad_bits_kernel(const uint8_t* __restrict__ item_flags, uint8_t flag_exclude, const int32_t* __restrict__ item_gids,
                                     int64_t item_id_base, int64_t n_items, int64_t n_words, uint32_t* __restrict__ bits) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // word index = tile * kChunks + chunk = pos / 32
    if (w >= n_words) return;
    const int64_t pos = w * 32 + (threadIdx.x & 31);
    bool bad = pos >= n_items;
    if (!bad) {
        const int64_t gid = item_gids ? (int64_t)__ldg(item_gids + pos) : item_id_base + pos;
        bad = (__ldg(item_flags + gid) & flag_exclude) != 0;
    }
    const unsigned b = __ballot_sync(CR_FULL_MASK, bad);
    if ((threadIdx.x & 31) == 0) bits[w] = b;
}

// dst[r, :] = [src[ids[r], 0:src_d] | 0 ...]: rows of width src_d (<= kD) gathered into kD-wide rows (zero columns do not change
// an inner product: a 32-, 48- or 96-wide table is swept by the 64 / 128 instantiation)
template <int kD>
__global__ void gather_q_kernel(const float4* __restrict__ src, const int32_t* __restrict__ ids, int64_t n, int src_d4, float4* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * (kD / 4)) return;
    const int64_t r = i / (kD / 4);
    const int c = (int)(i % (kD / 4));
    const int64_t row = ids ? (int64_t)__ldg(ids + r) : r;
    dst[i] = c < src_d4 ? __ldg(src + row * src_d4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// Lists that came up short because masked items never become candidates here: append the first masked
// items of the local table at CR_MASK_SCORE (the reference's top-K shows masked ids in that case).
__global__ void fill_masked_local_kernel(float* __restrict__ out_score, int32_t* __restrict__ out_id, int64_t n_q, int K,
                                         int64_t n_items, const int32_t* __restrict__ item_gids, int64_t item_id_base,
                                         const uint8_t* __restrict__ item_flags, uint8_t flag_exclude,
                                         const int64_t* __restrict__ mask_rowptr, const int32_t* __restrict__ mask_col) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_q) return;
    float* os = out_score + q * K;
    int32_t* oi = out_id + q * K;
    int m = K;
    while (m > 0 && oi[m - 1] < 0) --m;
    if (m == K) return;
    const int64_t lo = mask_rowptr ? mask_rowptr[q] : 0, hi = mask_rowptr ? mask_rowptr[q + 1] : 0;
    for (int64_t pos = 0; pos < n_items && m < K; ++pos) {
        const int gid = item_gids ? item_gids[pos] : (int)(item_id_base + pos);
        bool masked = item_flags && (item_flags[gid] & flag_exclude);
        if (!masked && mask_rowptr) masked = cr::csr_row_contains(mask_col, lo, hi, gid);
        if (!masked) continue;
        os[m] = CR_MASK_SCORE;
        oi[m] = gid;
        ++m;
    }
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
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

// rows x kD fp32, row-major; box = 32 floats x kBN rows, SWIZZLE_128B; out-of-range rows read as zero
int make_map(CUtensorMap* map, const float* base, int64_t rows, int kD, int kBN) {
    EncodeTiledFn enc;
    int rc = get_encode_fn(&enc);
    if (rc != CR_OK) return rc;
    cuuint64_t dims[2] = {(cuuint64_t)kD, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kD * sizeof(float)};
    cuuint32_t box[2] = {32, (cuuint32_t)kBN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cr::note_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
    return CR_OK;
}

struct TcPlan {
    int ksel, cap, n_utiles, n_q_pad, n_tiles, S, tiles_per_split;
    size_t off_q, off_buf, off_cnt, off_thr, off_norm, off_list, off_count, off_ps, off_pi, off_bits, off_pad, off_exact, exact_bytes, total;
};

TcPlan tc_plan(int64_t n_q, int64_t n_items, int K, int kD, int src_d = 0) {
    const int kBN = kD == 64 ? Geo<64>::kBN : Geo<128>::kBN;
    TcPlan P{};
    P.ksel = (K <= 24) ? 32 : 64;
    P.cap = P.ksel + 32;
    P.n_utiles = (int)((n_q + kBM - 1) / kBM);
    P.n_q_pad = P.n_utiles * kBM;
    P.n_tiles = (int)((n_items + kBN - 1) / kBN);
    // Item-range splits trade wave balance (units vs 148 SMs) against extra selection work: every split re-learns
    // its thresholds, so the slow-path count grows ~linearly with S.  Pick the S that minimises rounds x tiles/S x penalty.
    int best = 1;
    double best_cost = 1e300;
    const int max_by_tiles = P.n_tiles / 64 > 0 ? P.n_tiles / 64 : 1;
    for (int S = 1; S <= 32 && S <= max_by_tiles; ++S) {
        const double rounds = (double)(((int64_t)P.n_utiles * S + 147) / 148);
        const double cost = rounds * ((P.n_tiles + S - 1) / S) * (1.0 + 0.04 * (S - 1));
        if (cost < best_cost * 0.999) { best_cost = cost; best = S; }
    }
    P.tiles_per_split = (P.n_tiles + best - 1) / best;
    if (P.tiles_per_split < 1) P.tiles_per_split = 1;
    P.S = (P.n_tiles + P.tiles_per_split - 1) / P.tiles_per_split;
    if (P.S < 1) P.S = 1;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = cr::align_up(off + bytes, 256); return o; };
    P.off_q = take((size_t)(n_q > 0 ? n_q : 1) * kD * 4);
    P.off_buf = take((size_t)P.S * P.n_q_pad * P.cap * sizeof(Cand));
    P.off_cnt = take((size_t)P.S * P.n_q_pad * 4);
    P.off_thr = take((size_t)P.S * P.n_q_pad * 4);
    P.off_norm = take(256);
    P.off_list = take((size_t)(n_q > 0 ? n_q : 1) * 4);
    P.off_count = take(256);
    P.off_ps = take((size_t)P.S * n_q * K * 4);
    P.off_pi = take((size_t)P.S * n_q * K * 4);
    P.off_bits = take((size_t)(P.n_tiles + 1) * (kBN / 32) * 4);      // "never take" words of the flag mask (tile-chunk granularity)
    P.off_pad = take(src_d > 0 && src_d != kD ? (size_t)n_items * kD * 4 : 0);    // zero-padded copy of a narrower item table
    P.exact_bytes = cr::refine_workspace_bytes(K);
    P.off_exact = take(P.exact_bytes);
    P.total = off;
    return P;
}

}  // namespace

namespace cr {

// widths served by the tensor-core sweep: 64 and 128 natively; any other multiple of 4 up to 128 zero-padded to the next of the two
// (costs a padded copy of the item table in the workspace; d = 32 -> 64 doubles the FLOPs and still beats the FFMA kernel 25x)
static inline int tc_width(int d) { return (d <= 0 || d > 128 || d % 4 != 0) ? 0 : (d <= 64 ? 64 : 128); }

size_t tc_workspace_bytes(int64_t n_q, int64_t n_items, int d, int K) {
    if (!tc_width(d) || K > 52) return exact_workspace_bytes(n_q, n_items, K);
    return tc_plan(n_q, n_items, K, tc_width(d), d).total;
}

template <int kD>
static int launch_tc_scorer_d(const ExactJob& j, int32_t* n_refined, void* ws, size_t ws_bytes, cudaStream_t st, float* dbg_scores) {
    using G = Geo<kD>;
    const int64_t n_q = j.n_q, n_items = j.n_items;
    const int K = j.K;
    if (n_q == 0) return CR_OK;
    const TcPlan P = tc_plan(n_q, n_items, K, kD, j.d);
    if (!ws || ws_bytes < P.total) return CR_ERR_WORKSPACE;
    if (!aligned16(ws)) return CR_ERR_ALIGN;
    char* base = (char*)ws;
    float* Q = (float*)(base + P.off_q);
    Cand* buf = (Cand*)(base + P.off_buf);
    int* cnt = (int*)(base + P.off_cnt);
    float* thr = (float*)(base + P.off_thr);
    float* norm = (float*)(base + P.off_norm);
    int32_t* rlist = (int32_t*)(base + P.off_list);
    int32_t* rcount = (int32_t*)(base + P.off_count);
    float* part_s = (P.S > 1) ? (float*)(base + P.off_ps) : j.out_score;
    int32_t* part_i = (P.S > 1) ? (int32_t*)(base + P.off_pi) : j.out_id;
    const uint8_t* flags = j.flag_exclude ? j.item_flags : nullptr;
    const float* item_tab = j.item_tab;

    CR_CUDA_TRY(cudaMemsetAsync(norm, 0, 256, st));
    CR_CUDA_TRY(cudaMemsetAsync(rcount, 0, 256, st));
    {
        const int64_t total = n_q * (kD / 4);
        gather_q_kernel<kD><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float4*)j.user_tab, j.user_ids, n_q, j.d / 4, (float4*)Q);
        CR_LAUNCH_CHECK("gather_q_kernel");
        if (j.d != kD) {       // narrower table: zero-padded copy (the sweep, the rescoring and the norm bound read the copy)
            const int64_t tot_i = n_items * (kD / 4);
            gather_q_kernel<kD><<<(unsigned)((tot_i + 255) / 256), 256, 0, st>>>((const float4*)j.item_tab, nullptr, n_items, j.d / 4,
                                                                                (float4*)(base + P.off_pad));
            CR_LAUNCH_CHECK("gather_q_kernel");
            item_tab = (const float*)(base + P.off_pad);
        }
        item_norm_max_kernel<kD><<<148 * 4, 256, 0, st>>>((const float4*)item_tab, n_items, norm);
        CR_LAUNCH_CHECK("item_norm_max_kernel");
    }
    CUtensorMap map_i;
    int rc = make_map(&map_i, item_tab, n_items, kD, G::kBN);
    if (rc != CR_OK) return rc;

    SweepParams sp{};
    sp.Q = Q; sp.n_q = n_q; sp.n_q_pad = P.n_q_pad; sp.n_utiles = P.n_utiles; sp.n_items = n_items;
    sp.tiles_per_split = P.tiles_per_split; sp.n_tiles = P.n_tiles; sp.item_gids = j.item_gids; sp.item_id_base = j.item_id_base;
    sp.mask_rowptr = j.mask_rowptr; sp.mask_col = j.mask_col; sp.item_flags = flags;
    sp.flag_exclude = j.flag_exclude; sp.buf = buf; sp.cnt = cnt; sp.thr = thr; sp.dbg_scores = dbg_scores;
    sp.bad_bits = nullptr;
    if (flags) {
        uint32_t* bits = (uint32_t*)(base + P.off_bits);
        const int64_t n_words = (int64_t)P.n_tiles * G::kChunks;
        item_bad_bits_kernel<<<(unsigned)((n_words * 32 + 255) / 256), 256, 0, st>>>(flags, j.flag_exclude, j.item_gids, j.item_id_base, n_items,
                                                                                   n_words, bits);
        CR_LAUNCH_CHECK("item_bad_bits_kernel");
        sp.bad_bits = bits;
    }
    sp.seed_tiles = 128;     // <= 2*CAP: the tile maxima live in the query's (still empty) candidate buffer
    if (const char* e = getenv("CR_TC_SEED_TILES")) sp.seed_tiles = min(128, max(0, atoi(e)));   // A/B knob (0 disables)
    {
        const char* e = getenv("CR_TC_DEBUG_MODE");
        sp.dbg_mode = e ? atoi(e) : 0;
    }
    const unsigned grid = (unsigned)(P.n_utiles * P.S);
    prof_start(PROF_SCORE_SWEEP, st);
    const bool dbg = sp.dbg_mode != 0 || dbg_scores != nullptr;
    constexpr int kSmem = G::Smem::kTotal;
#define CR_SWEEP(KS, DB)                                                                                                        \
    do {                                                                                                                        \
        CR_CUDA_TRY(cudaFuncSetAttribute(score_sweep_tc_kernel<kD, KS, DB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); \
        score_sweep_tc_kernel<kD, KS, DB><<<grid, kThreads, kSmem, st>>>(map_i, sp);                                            \
    } while (0)
    if (P.ksel == 32) { if (dbg) CR_SWEEP(32, true); else CR_SWEEP(32, false); }
    else { if (dbg) CR_SWEEP(64, true); else CR_SWEEP(64, false); }
#undef CR_SWEEP
    CR_LAUNCH_CHECK("score_sweep_tc_kernel");
    prof_stop(PROF_SCORE_SWEEP, st);

    RescoreParams rp{Q, item_tab, j.item_gids, j.item_id_base, n_q, P.n_q_pad, P.S, K, P.cap, buf, cnt, thr, norm,
                     part_s, part_i, j.out_score, rlist, rcount};
    const int64_t warps = n_q * P.S;
    const unsigned rgrid = (unsigned)((warps * 32 + 255) / 256);
    if (P.cap == 64) rescore_kernel<2, kD><<<rgrid, 256, 0, st>>>(rp); else rescore_kernel<3, kD><<<rgrid, 256, 0, st>>>(rp);
    CR_LAUNCH_CHECK("rescore_kernel");
    if (P.S > 1) {
        rc = launch_merge(part_s, part_i, P.S, n_q, K, j.out_score, j.out_id, st, nullptr, 0, 0, -1, INT64_MAX);
        if (rc != CR_OK) return rc;
    }
    verify_kernel<kD><<<(unsigned)((n_q + 255) / 256), 256, 0, st>>>(rp);
    CR_LAUNCH_CHECK("verify_kernel");
    // lists shorter than K (fewer than K unmasked items): show masked ids at -1e9 like the reference's top-K
    if (j.mask_rowptr || flags) {
        fill_masked_local_kernel<<<(unsigned)((n_q + 127) / 128), 128, 0, st>>>(j.out_score, j.out_id, n_q, K, n_items, j.item_gids,
                                                                               j.item_id_base, flags, j.flag_exclude,
                                                                               j.mask_rowptr, j.mask_col);
        CR_LAUNCH_CHECK("fill_masked_local_kernel");
    }
    // queries whose margin could not be proven: exact fp32 re-run, rows overwritten in place
    ExactJob ej = j;
    ej.item_flags = flags;
    rc = launch_exact_refine(ej, RefineList{rlist, rcount, n_q}, base + P.off_exact, P.exact_bytes, st);
    if (rc != CR_OK) return rc;
    if (n_refined) CR_CUDA_TRY(cudaMemcpyAsync(n_refined, rcount, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return CR_OK;
}

int launch_tc_scorer(const ExactJob& j, int32_t* n_refined, void* ws, size_t ws_bytes, cudaStream_t st, float* dbg_scores) {
    if (n_refined) CR_CUDA_TRY(cudaMemsetAsync(n_refined, 0, sizeof(int32_t), st));
    if (!tc_width(j.d) || j.K > 52 || j.n_items == 0) {
        // exact fp32 everywhere: a stricter result than TF32-checked asks for, never a weaker one
        return launch_exact_scorer(j, ws, ws_bytes, st);
    }
    return tc_width(j.d) == 64 ? launch_tc_scorer_d<64>(j, n_refined, ws, ws_bytes, st, dbg_scores)
                               : launch_tc_scorer_d<128>(j, n_refined, ws, ws_bytes, st, dbg_scores);
}

int read_tc_timeline(unsigned long long* host_out, int n_units) {
    if (!host_out || n_units < 0 || n_units > kTlUnits) return CR_ERR_ARG;
    CR_CUDA_TRY(cudaMemcpyFromSymbol(host_out, g_tc_timeline, (size_t)n_units * kTlSlots * sizeof(unsigned long long)));
    return CR_OK;
}

}  // namespace cr
