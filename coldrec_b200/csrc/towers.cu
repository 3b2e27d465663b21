// K4 — content->embedding towers: one fused layer  Y = act((X . W^T + bias) * scale + shift).
//
// Replaces the nn.Linear -> BatchNorm1d(eval) -> tanh chains of model/DropoutNet.py:204-212,222-236,
// model/Heater.py:143-167,218-222, model/GAR.py:102-107 and model/ALDI.py:204-208.  The concat of
// DropoutNet.py:199-202 is folded in as a two-segment K loop, the content[cold_idx] gather and the
// item_emb[cold_idx] scatter (GAR.py:44-46, ALDI.py:96) as row maps.  fp32 FFMA: the parity bar for
// generated embeddings is 1e-5 norm-wise, which a single TF32 pass cannot meet.
#include "common.cuh"

namespace {

constexpr int kBM = 64, kBN = 64, kBK = 32;
constexpr int kLd = kBM + 4;   // smem leading dim (floats): keeps float4 reads aligned
constexpr int kThreads = 256;

struct LinearParams {
    const float* X1; int64_t ld1; int d1;
    const float* X2; int64_t ld2; int d2;
    const int32_t* xrow; int64_t n_rows;
    const float* W; const float* bias; const float* scale; const float* shift;
    int n_out; int act;
    float* Y; int64_t ldy; const int32_t* yrow;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == CR_ACT_TANH) return tanhf(v);
    if (act == CR_ACT_LEAKY_RELU) return v > 0.f ? v : 0.01f * v;
    return v;
}

__global__ void __launch_bounds__(kThreads) linear_act_kernel(const LinearParams p) {
    __shared__ __align__(16) float As[kBK][kLd];
    __shared__ __align__(16) float Bs[kBK][kLd];
    __shared__ int64_t s_row[kBM];
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int64_t m0 = (int64_t)blockIdx.x * kBM;
    const int n0 = blockIdx.y * kBN;
    const int kdim = p.d1 + p.d2;
    if (tid < kBM) {
        const int64_t r = m0 + tid;
        s_row[tid] = (r < p.n_rows) ? (p.xrow ? (int64_t)p.xrow[r] : r) : -1;
    }
    __syncthreads();
    const int ty = tid / 16, tx = tid % 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < kdim; k0 += kBK) {
        const int k = k0 + lane;
#pragma unroll
        for (int i = 0; i < kBM / 8; ++i) {
            const int r = wrp + 8 * i;
            const int64_t row = s_row[r];
            float v = 0.f;
            if (row >= 0 && k < kdim) v = (k < p.d1) ? __ldg(p.X1 + row * p.ld1 + k) : __ldg(p.X2 + row * p.ld2 + (k - p.d1));
            As[lane][r] = v;
            const int o = n0 + r;
            float w = 0.f;
            if (o < p.n_out && k < kdim) w = __ldg(p.W + (int64_t)o * kdim + k);
            Bs[lane][r] = w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = m0 + ty * 4 + i;
        if (r >= p.n_rows) continue;
        const int64_t orow = p.yrow ? (int64_t)p.yrow[r] : r;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = n0 + tx * 4 + j;
            if (o >= p.n_out) continue;
            float v = acc[i][j] + (p.bias ? __ldg(p.bias + o) : 0.f);
            if (p.scale) v = v * __ldg(p.scale + o) + __ldg(p.shift + o);
            p.Y[orow * p.ldy + o] = apply_act(v, p.act);
        }
    }
}

__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int n,
                               float* scale, float* shift) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float invstd = 1.f / sqrtf(var[i] + eps);
    const float sc = (gamma ? gamma[i] : 1.f) * invstd;
    scale[i] = sc;
    shift[i] = (beta ? beta[i] : 0.f) - mean[i] * sc;
}

__global__ void heater_blend_kernel(const float* __restrict__ gate, int n_expert, const float* __restrict__ expert,
                                    const float* __restrict__ Vin, float keep, float drop, int64_t n_rows, int d,
                                    float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * d) return;
    const int64_t r = i / d;
    const float e = expert[i];
    float s = 0.f;
    for (int g = 0; g < n_expert; ++g) s = fmaf(gate[r * n_expert + g], e, s);
    out[i] = Vin[i] * keep + tanhf(s) * drop;
}

}  // namespace

extern "C" {

int cr_linear_act_f32(const float* X1, int64_t ld1, int d1, const float* X2, int64_t ld2, int d2, const int32_t* xrow,
                      int64_t n_rows, const float* W, const float* bias, const float* scale, const float* shift, int n_out,
                      int act, float* Y, int64_t ldy, const int32_t* yrow, void* stream) {
    if (!X1 || !W || !Y || n_rows < 0 || d1 <= 0 || d2 < 0 || (d2 > 0 && !X2) || n_out <= 0) return CR_ERR_ARG;
    if (ld1 < d1 || (d2 > 0 && ld2 < d2) || ldy < n_out || ((scale == nullptr) != (shift == nullptr))) return CR_ERR_ARG;
    if (act < CR_ACT_NONE || act > CR_ACT_LEAKY_RELU) return CR_ERR_UNSUPPORTED;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    if (n_rows == 0) return CR_OK;
    LinearParams p{X1, ld1, d1, X2, ld2, d2, xrow, n_rows, W, bias, scale, shift, n_out, act, Y, ldy, yrow};
    const int64_t gx = (n_rows + kBM - 1) / kBM;
    const int gy = (n_out + kBN - 1) / kBN;
    if (gx > 0x7fffffffLL || gy > 65535) return CR_ERR_UNSUPPORTED;
    linear_act_kernel<<<dim3((unsigned)gx, (unsigned)gy), kThreads, 0, (cudaStream_t)stream>>>(p);
    CR_LAUNCH_CHECK("linear_act_kernel");
    return CR_OK;
}

int cr_bn_fold_f32(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int n, float* scale,
                   float* shift, void* stream) {
    if (!mean || !var || !scale || !shift || n <= 0) return CR_ERR_ARG;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    bn_fold_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, eps, n, scale, shift);
    CR_LAUNCH_CHECK("bn_fold_kernel");
    return CR_OK;
}

int cr_heater_blend_f32(const float* gate, int n_expert, const float* expert, const float* Vin, float keep,
                        float one_minus_keep, int64_t n_rows, int d, float* out, void* stream) {
    if (!gate || !expert || !Vin || !out || n_expert < 1 || n_rows < 0 || d <= 0) return CR_ERR_ARG;
    int rc = cr::require_device();
    if (rc != CR_OK) return rc;
    if (n_rows == 0) return CR_OK;
    const int64_t blocks = (n_rows * d + 255) / 256;
    if (blocks > 0x7fffffffLL) return CR_ERR_UNSUPPORTED;
    heater_blend_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(gate, n_expert, expert, Vin, keep, one_minus_keep,
                                                                            n_rows, d, out);
    CR_LAUNCH_CHECK("heater_blend_kernel");
    return CR_OK;
}

}  // extern "C"
