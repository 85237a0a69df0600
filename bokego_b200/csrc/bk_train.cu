// bk_train.cu -- the REINFORCE step of bin/selfplay.py:59-122: PolicyNet.forward in train() mode,
// the policy-gradient loss, its backward pass through the conv stack, and torch.optim.AdamW's update.
//
// What the reference computes (oracle/train.py restates it, tests/golden/reinforce.npz pins it):
//   * `pi` is in train() mode and is called with ONE position per call (nnet.py:265-275), so every BatchNorm2d
//     normalises a position with that position's own statistics (81 squares per channel, biased variance, eps 1e-5):
//     bn_mode 0.  bn_mode 1 = running statistics (eval mode), for training with frozen BatchNorm.
//   * loss = sum_p coef[p] * -log(clamp(softmax(logits_p)[move_p]))        (coef = reward / bs, selfplay.py:98-117)
//   * AdamW: decoupled weight decay, bias-corrected moments (selfplay.py:138).
//
// Layout: activations are channel-last float32 [P][81][C]; a conv is an implicit GEMM over rows m = (position, square)
// with K = (tap, ci) and the 3x3 / 5x5 window gathered while the A tile is staged (zero-filled cp.async for taps that
// leave the board).  The three GEMMs per layer -- forward, data gradient (the same kernel on dZ with mirrored taps and
// the transposed weights) and weight gradient (reduction over all rows, split over CTAs, fixed-order second pass) -- run
// on the tensor cores with TF32 operands and fp32 accumulation (mma.sync.m16n8k8; prec 1 = 3xTF32 split for
// fp32-grade results, prec 2 = plain FFMA over the same tiles, the validation path).  Parameters, gradients and Adam
// moments are flat float32 buffers in the GEMM layout (offsets in include/bokego_b200.h); everything is deterministic
// (no atomics).  This first version of the row keeps the training GEMMs on the warp-level MMA path; moving them onto
// tcgen05 like the inference kernel is the next step.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <mutex>

#include "../../include/bokego_b200.h"
#include "bk_layout.h"
#include "bk_train_args.h"

namespace {

constexpr int C = 128;              // trunk width
constexpr int NSQ = 81;
constexpr int C0 = 32;              // layer-0 input channels, 27 padded to 32
constexpr int BM = 128, BK = 32;    // CTA tile: 128 GEMM rows x 128 columns, 32 deep
constexpr int A_LD = BK + 4;        // 36 floats: conflict-free fragment reads (4*g + t)
constexpr int B_LD = C + 8;         // 136 floats: conflict-free fragment reads (8*t + g)
constexpr int NT = 256;
constexpr int CONV_SMEM = 2 * (BM * A_LD + BK * B_LD) * 4;   // 71,680
constexpr int WGRAD_SMEM = 2 * 2 * BK * B_LD * 4;            // 69,632
constexpr int SPLIT_MAX_P = 64;     // up to this many positions the tcgen05 3x3 forward convs run one CTA per channel group (4 partial results)
constexpr float BN_EPS = 1e-5f;
constexpr float PROB_EPS = 1.1920928955078125e-07f;          // torch.finfo(float32).eps (Categorical's clamp)

__device__ __forceinline__ void cp_async16(void *smem, const void *g, bool valid)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;                            // 0 source bytes = 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(g), "r"(sz));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ uint32_t f2tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// One warp's 64 x 32 tile over a BK-deep slab.  A element (i, k) = a[i * A_SI + k * A_SK], B element (k, j) = b[k * B_LD + j].
// Accumulator fragment (mma.m16n8k8): c0,c1 = (row g, cols 2t, 2t+1), c2,c3 = (row g+8, same cols); g = lane / 4, t = lane % 4.
template <int PREC, int A_SI, int A_SK>
__device__ __forceinline__ void warp_tile(float (&acc)[4][4][4], const float *a, const float *b, int lane)
{
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int k8 = 0; k8 < BK; k8 += 8) {
        if (PREC == 2) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                float bv[4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bv[j][0] = b[(k8 + kk) * B_LD + 8 * j + 2 * t];
                    bv[j][1] = b[(k8 + kk) * B_LD + 8 * j + 2 * t + 1];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float a0 = a[(16 * i + g) * A_SI + (k8 + kk) * A_SK];
                    const float a1 = a[(16 * i + g + 8) * A_SI + (k8 + kk) * A_SK];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        acc[i][j][0] = fmaf(a0, bv[j][0], acc[i][j][0]);
                        acc[i][j][1] = fmaf(a0, bv[j][1], acc[i][j][1]);
                        acc[i][j][2] = fmaf(a1, bv[j][0], acc[i][j][2]);
                        acc[i][j][3] = fmaf(a1, bv[j][1], acc[i][j][3]);
                    }
                }
            }
        } else {
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x0 = b[(k8 + t) * B_LD + 8 * j + g];          // b0 = (k = t, n = g)
                const float x1 = b[(k8 + t + 4) * B_LD + 8 * j + g];      // b1 = (k = t + 4, n = g)
                bh[j][0] = f2tf32(x0);
                bh[j][1] = f2tf32(x1);
                if (PREC == 1) {
                    bl[j][0] = f2tf32(x0 - __uint_as_float(bh[j][0]));
                    bl[j][1] = f2tf32(x1 - __uint_as_float(bh[j][1]));
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float y[4];
                y[0] = a[(16 * i + g) * A_SI + (k8 + t) * A_SK];          // a0 = (row g,     k = t)
                y[1] = a[(16 * i + g + 8) * A_SI + (k8 + t) * A_SK];      // a1 = (row g + 8, k = t)
                y[2] = a[(16 * i + g) * A_SI + (k8 + t + 4) * A_SK];      // a2 = (row g,     k = t + 4)
                y[3] = a[(16 * i + g + 8) * A_SI + (k8 + t + 4) * A_SK];  // a3 = (row g + 8, k = t + 4)
                uint32_t ah[4], al[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    ah[r] = f2tf32(y[r]);
                    if (PREC == 1) al[r] = f2tf32(y[r] - __uint_as_float(ah[r]));
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (PREC == 1) {
                        // the tensor core's fp32 accumulation truncates; keep its chains three MMAs short and add the
                        // partial sums with IEEE fp32 adds, otherwise the split buys nothing over a long reduction
                        // (interleaving the chains of the four column blocks was tried: slower, it spills)
                        float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                        mma_tf32(c, al, bh[j]);
                        mma_tf32(c, ah, bl[j]);
                        mma_tf32(c, ah, bh[j]);
#pragma unroll
                        for (int r = 0; r < 4; ++r) acc[i][j][r] += c[r];
                    } else {
                        mma_tf32(acc[i][j], ah, bh[j]);
                    }
                }
            }
        }
    }
}

// ---- implicit-GEMM convolution: out[m][co] = sum_{tap,ci} in[m shifted by sign*tap][ci] * w[(tap,ci)][co] (+ bias[co]) ----
// sign = +1: forward conv (cross-correlation, zero padding R/2).  sign = -1 with w = the per-tap transposed weights: the
// data gradient.  in: [M/81][81][Cin], w: [R*R*Cin][128], out: [M][128].
using ConvArgs = BkConvArgs;

template <int PREC>
__global__ void __launch_bounds__(NT, 2) bk_train_conv_kernel(const ConvArgs a)
{
    extern __shared__ __align__(16) float smem[];
    float *As = smem;                        // [2][BM][A_LD]
    float *Bs = smem + 2 * BM * A_LD;        // [2][BK][B_LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int m0 = blockIdx.x * BM;
    const int ch = tid & 7;                  // 16-byte chunk of a 32-float A row
    int rbase[4], rx[4], ry[4];
    bool rvalid[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + (tid >> 3) + 32 * i;
        rvalid[i] = m < a.M;
        const int p = m / NSQ, sq = m - p * NSQ;
        rx[i] = sq / 9;
        ry[i] = sq - 9 * rx[i];
        rbase[i] = p * NSQ;
    }
    const int half = a.R >> 1;
    const int KT = a.R * a.R * a.Cin / BK;
    auto load = [&](int kt, int buf) {
        const int k0 = kt * BK;
        const int tap = k0 / a.Cin, c0 = k0 - tap * a.Cin;
        const int ti = tap / a.R;
        const int dx = a.sign * (ti - half), dy = a.sign * (tap - ti * a.R - half);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = rx[i] + dx, y = ry[i] + dy;
            const bool ok = rvalid[i] && (unsigned)x < 9u && (unsigned)y < 9u;
            const float *src = ok ? a.in + ((size_t)(rbase[i] + 9 * x + y) * a.Cin + c0 + 4 * ch) : a.in;
            cp_async16(As + (buf * BM + (tid >> 3) + 32 * i) * A_LD + 4 * ch, src, ok);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = (tid >> 5) + 8 * i;
            cp_async16(Bs + (buf * BK + r) * B_LD + 4 * lane, a.w + (size_t)(k0 + r) * C + 4 * lane, true);
        }
    };
    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.0f;
    load(0, 0);
    cp_commit();
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) {
            load(kt + 1, buf ^ 1);
            cp_commit();
            cp_wait<1>();
        } else {
            cp_wait<0>();
        }
        __syncthreads();
        warp_tile<PREC, A_LD, 1>(acc, As + (buf * BM + wm * 64) * A_LD, Bs + buf * BK * B_LD + wn * 32, lane);
        __syncthreads();
    }
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = wn * 32 + 8 * j + 2 * t;
            const float b0 = a.bias ? a.bias[col] : 0.0f, b1 = a.bias ? a.bias[col + 1] : 0.0f;
            const int r0 = m0 + wm * 64 + 16 * i + g;
            if (r0 < a.M) *reinterpret_cast<float2 *>(a.out + (size_t)r0 * C + col) = make_float2(acc[i][j][0] + b0, acc[i][j][1] + b1);
            if (r0 + 8 < a.M)
                *reinterpret_cast<float2 *>(a.out + (size_t)(r0 + 8) * C + col) = make_float2(acc[i][j][2] + b0, acc[i][j][3] + b1);
        }
    }
}

// ---- weight gradient: part[split][k][co] = sum over the split's rows m of act[m shifted by tap(k)][ci(k)] * dz[m][co] ----
// grid = (k tiles of 128, splits).  For Cin = 128 a k tile is one tap; for Cin = 32 it is four taps.
using WgradArgs = BkWgradArgs;

template <int PREC>
__global__ void __launch_bounds__(NT, 2) bk_train_wgrad_kernel(const WgradArgs a)
{
    extern __shared__ __align__(16) float smem[];
    float *Am = smem;                        // [2][BK rows m][B_LD]  (k along the row)
    float *Bm = smem + 2 * BK * B_LD;        // [2][BK rows m][B_LD]  (co along the row)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int kb = blockIdx.x, split = blockIdx.y;
    const int r_lo = split * a.rows_per_split;
    const int r_hi = min(a.M, r_lo + a.rows_per_split);
    const int kglob = kb * 128 + 4 * lane;   // this thread's 16-byte chunk of the k tile: fixed tap and channel
    const int tap = kglob / a.Cin, ci = kglob - tap * a.Cin;
    const bool tap_ok = tap < a.R * a.R;
    const int half = a.R >> 1, ti = tap / a.R;
    const int dx = ti - half, dy = tap - ti * a.R - half;
    auto load = [&](int r0, int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = (tid >> 5) + 8 * i;
            const int m = r0 + row;
            const bool in_rows = m < r_hi;
            const int p = m / NSQ, sq = m - p * NSQ;
            const int x = sq / 9 + dx, y = sq - 9 * (sq / 9) + dy;
            const bool ok = in_rows && tap_ok && (unsigned)x < 9u && (unsigned)y < 9u;
            const float *src = ok ? a.act + ((size_t)(p * NSQ + 9 * x + y) * a.Cin + ci) : a.act;
            cp_async16(Am + (buf * BK + row) * B_LD + 4 * lane, src, ok);
            const float *srcb = in_rows ? a.dz + (size_t)m * C + 4 * lane : a.dz;
            cp_async16(Bm + (buf * BK + row) * B_LD + 4 * lane, srcb, in_rows);
        }
    };
    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.0f;
    const int n_it = r_hi > r_lo ? (r_hi - r_lo + BK - 1) / BK : 0;
    if (n_it > 0) {
        load(r_lo, 0);
        cp_commit();
    }
    for (int it = 0; it < n_it; ++it) {
        const int buf = it & 1;
        if (it + 1 < n_it) {
            load(r_lo + (it + 1) * BK, buf ^ 1);
            cp_commit();
            cp_wait<1>();
        } else {
            cp_wait<0>();
        }
        __syncthreads();
        // GEMM rows = k (A element (i, m) = Am[m][i]), GEMM columns = co, depth = the 32 staged rows m
        warp_tile<PREC, 1, B_LD>(acc, Am + buf * BK * B_LD + wm * 64, Bm + buf * BK * B_LD + wn * 32, lane);
        __syncthreads();
    }
    const int g = lane >> 2, t = lane & 3;
    float *out = a.part + (size_t)split * a.K * C;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = wn * 32 + 8 * j + 2 * t;
            const int k0 = kb * 128 + wm * 64 + 16 * i + g;
            if (k0 < a.K) *reinterpret_cast<float2 *>(out + (size_t)k0 * C + col) = make_float2(acc[i][j][0], acc[i][j][1]);
            if (k0 + 8 < a.K) *reinterpret_cast<float2 *>(out + (size_t)(k0 + 8) * C + col) = make_float2(acc[i][j][2], acc[i][j][3]);
        }
    }
}

// grad[i] (+)= sum_s part[s][i], splits added in order
__global__ void bk_train_reduce_kernel(const float *part, float *grad, int n, int S, int accumulate)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = accumulate ? grad[i] : 0.0f;
    for (int k = 0; k < S; ++k) s += part[(size_t)k * n + i];
    grad[i] = s;
}

// out[j] (+)= sum_p in[p][j]; rows added in a fixed order (8 interleaved partial sums, then those in order)
__global__ void bk_train_colsum_kernel(const float *in, float *out, int P, int width, int accumulate)
{
    __shared__ float s[8][32];
    const int j = blockIdx.x * 32 + threadIdx.x;
    float v = 0.0f;
    if (j < width)
        for (int p = threadIdx.y; p < P; p += 8) v += in[(size_t)p * width + j];
    s[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.y == 0 && j < width) {
        float t = accumulate ? out[j] : 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += s[k][threadIdx.x];
        out[j] = t;
    }
}

// uint8 planes [P][27][81] (bk_encode's planes_u8 == nnet.features values) -> float32 [P][81][32], channels 27..31 zero
__global__ void bk_train_pack_kernel(const uint8_t *planes, float *x0, int P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * NSQ * C0) return;
    const int c = i & 31, r = i >> 5;
    const int p = r / NSQ, sq = r - p * NSQ;
    x0[i] = c < 27 ? (float)planes[((size_t)p * 27 + c) * NSQ + sq] : 0.0f;
}

// wd[(tap, co)][ci] = w[(tap, ci)][co] for the six 3x3 layers (operand of the data-gradient GEMM)
__global__ void bk_train_transpose_kernel(const float *w, float *wd, int n_mat)
{
    __shared__ float tile[32][33];
    const int mat = blockIdx.z;
    if (mat >= n_mat) return;
    const float *src = w + (size_t)mat * C * C;
    float *dst = wd + (size_t)mat * C * C;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) tile[r][threadIdx.x] = src[(size_t)(by + r) * C + bx + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) dst[(size_t)(bx + r) * C + by + threadIdx.x] = tile[threadIdx.x][r];
}

// The K-major B operand of the tcgen05 kernels (bk_train_tc.cu), split for 3xTF32 once per step instead of once per tile:
// hi[k / 4][co][k % 4] = rna_tf32(w[k][co]), lo[...] = rna_tf32(w - hi); n floats in total
// bh / bl: the same two parts rounded to bf16 and packed [k / 8][co][k % 8] (K-major operands of kind::f16: the 3x3 kernel's
// BK_R3_LO_BF16 build computes the low-order products with them)
__global__ void bk_train_pack_w_kernel(const float *w, float *hi, float *lo, __nv_bfloat16 *bh, __nv_bfloat16 *bl, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = i / C, co = i - k * C;
    const size_t o = (size_t)(k >> 2) * (4 * C) + co * 4 + (k & 3);
    const float x = w[i], h = __uint_as_float(f2tf32(x));
    hi[o] = h;
    lo[o] = __uint_as_float(f2tf32(x - h));
    if (bh) {                                     // only the BK_R3_LO_BF16 measurement build of the 3x3 kernel reads these
        const size_t o8 = (size_t)(k >> 3) * (8 * C) + co * 8 + (k & 7);
        bh[o8] = __float2bfloat16_rn(x);
        bl[o8] = __float2bfloat16_rn(x - h);
    }
}

// z[i] = z[i] + z[n + i] + z[2n + i] + z[3n + i] (in that order): the four channel-group partial results of a small-batch conv
__global__ void bk_train_sum4_kernel(float4 *z, size_t n4)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 v = z[i];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        const float4 u = z[k * n4 + i];
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    z[i] = v;
}

// ---- BatchNorm (+ ReLU) forward for one position per CTA, thread = channel ----
// bn_mode 0: statistics of this position (train-mode batch of one); 1: running statistics.
__global__ void __launch_bounds__(C) bk_train_bn_fwd_kernel(const float *z, const float *gamma, const float *beta,
                                                            const float *run_mean, const float *run_var, float *act,
                                                            float *mean_out, float *rstd_out, float *stats_out, int layer,
                                                            int bn_mode)
{
    const int p = blockIdx.x, c = threadIdx.x;
    const float *zp = z + (size_t)p * NSQ * C + c;
    float v[NSQ];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NSQ; ++i) {
        v[i] = zp[i * C];
        s += v[i];
    }
    const float mu_p = s * (1.0f / NSQ);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NSQ; ++i) q = fmaf(v[i] - mu_p, v[i] - mu_p, q);
    float mu = mu_p, rstd = 1.0f / sqrtf(q * (1.0f / NSQ) + BN_EPS);
    if (bn_mode == 1) {
        mu = run_mean[c];
        rstd = 1.0f / sqrtf(run_var[c] + BN_EPS);
    }
    mean_out[p * C + c] = mu;
    rstd_out[p * C + c] = rstd;
    if (stats_out) {   // what a train-mode call feeds into the running averages: mean and UNBIASED variance
        stats_out[((size_t)p * 7 + layer) * 2 * C + c] = mu_p;
        stats_out[((size_t)p * 7 + layer) * 2 * C + C + c] = q * (1.0f / (NSQ - 1));
    }
    const float ga = gamma[c], be = beta[c];
    float *ap = act + (size_t)p * NSQ * C + c;
#pragma unroll
    for (int i = 0; i < NSQ; ++i) ap[i * C] = fmaxf(fmaf((v[i] - mu) * rstd, ga, be), 0.0f);
}

// ---- BatchNorm (+ ReLU) backward for one position per CTA; part[p][0..2][c] = d conv-bias, d gamma, d beta contributions (the flat layout's order) ----
__global__ void __launch_bounds__(C) bk_train_bn_bwd_kernel(const float *dact, const float *act, const float *z,
                                                            const float *gamma, const float *mean, const float *rstd, float *dz,
                                                            float *part, int bn_mode)
{
    const int p = blockIdx.x, c = threadIdx.x;
    const size_t off = (size_t)p * NSQ * C + c;
    const float mu = mean[p * C + c], rs = rstd[p * C + c], ga = gamma[c];
    float g[NSQ], xh[NSQ];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < NSQ; ++i) {
        g[i] = act[off + i * C] > 0.0f ? dact[off + i * C] : 0.0f;
        xh[i] = (z[off + i * C] - mu) * rs;
        s1 += g[i];
        s2 = fmaf(g[i], xh[i], s2);
    }
    const float m1 = bn_mode == 0 ? s1 * (1.0f / NSQ) : 0.0f, m2 = bn_mode == 0 ? s2 * (1.0f / NSQ) : 0.0f;
    const float k = ga * rs;
    float sb = 0.0f;
#pragma unroll
    for (int i = 0; i < NSQ; ++i) {
        const float d = k * (g[i] - m1 - xh[i] * m2);
        dz[off + i * C] = d;
        sb += d;
    }
    part[((size_t)p * 3 + 0) * C + c] = sb;
    part[((size_t)p * 3 + 1) * C + c] = s2;
    part[((size_t)p * 3 + 2) * C + c] = s1;
}

// ---- head: Conv2dUntiedBias 1x1 (nnet.py:138-180) -> logits [P][81] and SOFT (nnet.py:16) probabilities ----
__global__ void __launch_bounds__(128) bk_train_head_fwd_kernel(const float *a6, const float *hw, const float *hb, float *logits,
                                                                 float *probs)
{
    __shared__ float lg[NSQ];
    __shared__ float red[2];
    const int p = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 w = *reinterpret_cast<const float4 *>(hw + 4 * lane);
    for (int sq = warp; sq < NSQ; sq += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(a6 + ((size_t)p * NSQ + sq) * C + 4 * lane);
        float s = v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) lg[sq] = s + hb[sq];
    }
    __syncthreads();
    if (threadIdx.x < NSQ) logits[p * NSQ + threadIdx.x] = lg[threadIdx.x];
    if (!probs) return;
    if (warp == 0) {
        float mx = -INFINITY;
        for (int i = lane; i < NSQ; i += 32) mx = fmaxf(mx, lg[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.0f;
        for (int i = lane; i < NSQ; i += 32) s += expf(lg[i] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            red[0] = mx;
            red[1] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x < NSQ) probs[p * NSQ + threadIdx.x] = expf(lg[threadIdx.x] - red[0]) / red[1];
}

// ---- loss and head backward: nlp[p] = -log_prob(move), dlogit = coef * (softmax - onehot), d a6, d head_w contributions ----
__global__ void __launch_bounds__(128) bk_train_head_bwd_kernel(const float *logits, const float *a6, const float *hw,
                                                                 const int16_t *moves, const float *coef, float *nlp,
                                                                 float *dlogit, float *da6, float *dw_part)
{
    __shared__ float pr[NSQ];
    __shared__ float dl[NSQ];
    __shared__ float red[2];
    const int p = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < NSQ) pr[threadIdx.x] = logits[p * NSQ + threadIdx.x];
    __syncthreads();
    if (warp == 0) {
        float mx = -INFINITY;
        for (int i = lane; i < NSQ; i += 32) mx = fmaxf(mx, pr[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.0f;
        for (int i = lane; i < NSQ; i += 32) s += expf(pr[i] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            red[0] = mx;
            red[1] = s;
        }
    }
    __syncthreads();
    const int mv = moves[p];
    const float cf = coef[p];
    if (threadIdx.x < NSQ) {
        const float q = expf(pr[threadIdx.x] - red[0]) / red[1];
        pr[threadIdx.x] = q;
    }
    __syncthreads();
    const bool mv_ok = mv >= 0 && mv < NSQ;
    const float pm = mv_ok ? pr[mv] : 1.0f;
    const bool clamped = pm < PROB_EPS || pm > 1.0f - PROB_EPS;   // Categorical clamps before the log: no gradient there
    if (threadIdx.x == 0) nlp[p] = mv_ok ? -logf(fminf(fmaxf(pm, PROB_EPS), 1.0f - PROB_EPS)) : 0.0f;
    if (threadIdx.x < NSQ) {
        const float d = (!mv_ok || clamped) ? 0.0f : cf * (pr[threadIdx.x] - (threadIdx.x == mv ? 1.0f : 0.0f));
        dl[threadIdx.x] = d;
        dlogit[p * NSQ + threadIdx.x] = d;
    }
    __syncthreads();
    const int c = threadIdx.x;
    const float w = hw[c];
    float s = 0.0f;
    const size_t off = (size_t)p * NSQ * C + c;
#pragma unroll 9
    for (int i = 0; i < NSQ; ++i) {
        s = fmaf(dl[i], a6[off + i * C], s);
        da6[off + i * C] = w * dl[i];
    }
    dw_part[p * C + c] = s;
}

// running_mean / running_var after train-mode calls on positions seq[0], seq[1], ... in that order (momentum filter)
__global__ void bk_train_running_kernel(float *running, const float *stats, const int32_t *seq, int S, float momentum)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (layer, which, channel)
    if (i >= 7 * 2 * C) return;
    const int layer = i / (2 * C), r = i - layer * 2 * C;  // r = which * C + c
    const int which = r / C, c = r - which * C;
    float v = running[(which * 7 + layer) * C + c];
    const float keep = 1.0f - momentum;
    for (int k = 0; k < S; ++k) {
        const int p = seq ? seq[k] : k;
        v = keep * v + momentum * stats[((size_t)p * 7 + layer) * 2 * C + r];
    }
    running[(which * 7 + layer) * C + c] = v;
}

__global__ void bk_adamw_kernel(float *p, const float *g, float *m, float *v, size_t n, float decay, float w1, float b2, float w2,
                                float bc2_sqrt, float eps, float step_size)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    float pi = p[i] * decay;
    const float mi = m[i] + w1 * (gi - m[i]);
    const float vi = v[i] * b2 + (w2 * gi) * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
}

// ---- workspace (floats) ----
struct Ws {
    size_t x0, z[7], a[7], mean[7], rstd[7], da, dz, logits, dlogit, part, dwpart, wd, wpack, wpack_lo, wdpack, wdpack_lo, wpart, tail, wb16, wdb16, total;
    int splits, rows_per_split;
};

Ws ws_layout(int P)
{
    Ws w;
    size_t o = 0;
    auto take = [&](size_t n) {
        const size_t r = o;
        o += (n + 31) / 32 * 32;
        return r;
    };
    const size_t act = (size_t)P * NSQ * C;
    w.x0 = take((size_t)P * NSQ * C0);
    for (int l = 0; l < 7; ++l) {
        w.z[l] = take(P <= SPLIT_MAX_P ? 4 * act : act);           // small batches: four channel-group partial results, summed into the first
        w.a[l] = take(act);
        w.mean[l] = take((size_t)P * C);
        w.rstd[l] = take((size_t)P * C);
    }
    w.da = take(act);
    w.dz = take(act);
    w.logits = take((size_t)P * NSQ);
    w.dlogit = take((size_t)P * NSQ);
    w.part = take((size_t)P * 3 * C);
    w.dwpart = take((size_t)P * C);
    w.wd = take((size_t)6 * 9 * C * C);
    w.wpack = take((size_t)BK_TP_VEC);              // all conv weights, K-major packed TF32 high parts (tcgen05 path)
    w.wpack_lo = take((size_t)BK_TP_VEC);           // ... and low parts
    w.wdpack = take((size_t)6 * 9 * C * C);         // the transposed weights of the data gradient, packed the same way
    w.wdpack_lo = take((size_t)6 * 9 * C * C);
    const int M = P * NSQ;
    int splits = (M + 255) / 256;                   // >= 256 rows per split, at most 32 splits (small batches: more CTAs, shorter loops)
    splits = splits < 1 ? 1 : (splits > 32 ? 32 : splits);
    int rps = (M + splits - 1) / splits;
    rps = (rps + BK - 1) / BK * BK;
    w.splits = splits;
    w.rows_per_split = rps;
    w.wpart = take((size_t)splits * 9 * C * C);
    w.tail = take((size_t)BK_CONV3_TAIL_FLOATS);    // partial results of the split tail of the staged-once 3x3 kernel
    w.wb16 = take((size_t)BK_TP_VEC);               // all conv weights as bf16 high parts, then bf16 low parts (BK_TP_VEC elements each)
    w.wdb16 = take((size_t)6 * 9 * C * C);          // ... of the data gradient's transposed weights
    w.total = o;
    return w;
}

// kernel attributes are per device: one flag per device, guarded by a mutex (several host threads / devices per process)
std::mutex attrs_mutex;
bool attrs_set[BK_MAX_DEVICES] = {};
int set_attrs()
{
    const int slot = bk_current_device_slot();
    if (slot < 0) return -2;
    std::lock_guard<std::mutex> lock(attrs_mutex);
    if (attrs_set[slot]) return 0;
    cudaError_t e = cudaSuccess;
#define BK_SET(k, n) if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, n)
    BK_SET(bk_train_conv_kernel<0>, CONV_SMEM);
    BK_SET(bk_train_conv_kernel<1>, CONV_SMEM);
    BK_SET(bk_train_conv_kernel<2>, CONV_SMEM);
    BK_SET(bk_train_wgrad_kernel<0>, WGRAD_SMEM);
    BK_SET(bk_train_wgrad_kernel<1>, WGRAD_SMEM);
    BK_SET(bk_train_wgrad_kernel<2>, WGRAD_SMEM);
#undef BK_SET
    if (e != cudaSuccess || bk_tc_set_attrs() != 0) return -3;
    attrs_set[slot] = true;
    return 0;
}

void launch_conv(const ConvArgs &a, int prec, cudaStream_t st)
{
    const int grid = (a.M + BM - 1) / BM;
    if (prec >= 4) bk_tc_launch_conv(a, prec == 5, st);
    else if (prec == 1) bk_train_conv_kernel<1><<<grid, NT, CONV_SMEM, st>>>(a);
    else if (prec == 2) bk_train_conv_kernel<2><<<grid, NT, CONV_SMEM, st>>>(a);
    else bk_train_conv_kernel<0><<<grid, NT, CONV_SMEM, st>>>(a);
}

void launch_wgrad(const WgradArgs &a, int splits, int prec, cudaStream_t st)
{
    const dim3 grid((a.K + 127) / 128, splits);
    if (prec >= 4) bk_tc_launch_wgrad(a, splits, prec == 5, st);
    else if (prec == 1) bk_train_wgrad_kernel<1><<<grid, NT, WGRAD_SMEM, st>>>(a);
    else if (prec == 2) bk_train_wgrad_kernel<2><<<grid, NT, WGRAD_SMEM, st>>>(a);
    else bk_train_wgrad_kernel<0><<<grid, NT, WGRAD_SMEM, st>>>(a);
}

inline size_t w_off(int l) { return l == 0 ? (size_t)BK_TP_W0 : (size_t)BK_TP_W1 + (size_t)(l - 1) * 9 * C * C; }
inline size_t vec_off(int l, int which) { return (size_t)BK_TP_VEC + ((size_t)l * 3 + which) * C; }

}   // namespace

extern "C" size_t bk_train_param_count(void) { return BK_TP_COUNT; }

extern "C" size_t bk_train_workspace_bytes(int P) { return P <= 0 ? 0 : ws_layout(P).total * sizeof(float); }

extern "C" int bk_train_launches(int which, int P, int prec)
{
    const int tc = prec >= 4 ? 1 : 0;                               // the tcgen05 path repacks the weights once per call
    const int tail = tc && P > SPLIT_MAX_P && !getenv("BK_TC_OLD_CONV") && bk_tc_conv3_tail(P) > 0 ? 6 : 0;   // one more launch per 3x3 conv: its split tail
    if (which == 0) return 1 + tc + 7 * 2 + 1 + (tc && P <= SPLIT_MAX_P ? 6 : 0) + tail;
    return 1 + tc + 1 + 2 + 7 * 4 + 6 + (tc && !getenv("BK_TC_OLD_CONV") && bk_tc_conv3_tail(P) > 0 ? 6 : 0);
}

extern "C" int bk_train_forward(const float *params, const float *running, const uint8_t *planes_u8, int P, int bn_mode, int prec,
                                void *workspace, float *logits, float *probs, float *stats_out, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P <= 0) return 0;
    if (!params || !running || !planes_u8 || !workspace || bn_mode < 0 || bn_mode > 1 || prec < 0 || prec > 5 || prec == 3) return -1;
    if (set_attrs() != 0) return -3;
    const Ws w = ws_layout(P);
    float *ws = static_cast<float *>(workspace);
    const int n0 = P * NSQ * C0;
    bk_train_pack_kernel<<<(n0 + 255) / 256, 256, 0, st>>>(planes_u8, ws + w.x0, P);
    __nv_bfloat16 *wb = bk_tc_lo_bf16() ? reinterpret_cast<__nv_bfloat16 *>(ws + w.wb16) : nullptr;
    if (prec >= 4) bk_train_pack_w_kernel<<<(BK_TP_VEC + 255) / 256, 256, 0, st>>>(params, ws + w.wpack, ws + w.wpack_lo, wb, wb ? wb + BK_TP_VEC : nullptr, BK_TP_VEC);
    for (int l = 0; l < 7; ++l) {
        ConvArgs a;
        a.in = l == 0 ? ws + w.x0 : ws + w.a[l - 1];
        a.w = prec >= 4 ? ws + w.wpack + w_off(l) : params + w_off(l);
        a.w_lo = ws + w.wpack_lo + w_off(l);
        if (wb) {
            a.w_bh = reinterpret_cast<const unsigned short *>(wb + w_off(l));
            a.w_bl = reinterpret_cast<const unsigned short *>(wb + BK_TP_VEC + w_off(l));
        }
        a.bias = params + vec_off(l, 0);
        a.out = ws + w.z[l];
        a.M = P * NSQ;
        a.Cin = l == 0 ? C0 : C;
        a.R = l == 0 ? 5 : 3;
        a.sign = 1;
        a.ksplit = (prec >= 4 && l > 0 && P <= SPLIT_MAX_P) ? 4 : 1;
        a.tail_part = ws + w.tail;
        launch_conv(a, prec, st);
        if (a.ksplit > 1) {
            const size_t n4 = (size_t)P * NSQ * C / 4;
            bk_train_sum4_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<float4 *>(ws + w.z[l]), n4);
        }
        bk_train_bn_fwd_kernel<<<P, C, 0, st>>>(ws + w.z[l], params + vec_off(l, 1), params + vec_off(l, 2), running + l * C,
                                                running + (7 + l) * C, ws + w.a[l], ws + w.mean[l], ws + w.rstd[l], stats_out, l,
                                                bn_mode);
    }
    bk_train_head_fwd_kernel<<<P, 128, 0, st>>>(ws + w.a[6], params + BK_TP_HEADW, params + BK_TP_HEADB, ws + w.logits, probs);
    if (logits && cudaMemcpyAsync(logits, ws + w.logits, (size_t)P * NSQ * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return -3;
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int bk_train_backward(const float *params, const int16_t *moves, const float *coef, int P, int bn_mode, int prec,
                                 void *workspace, float *grads, int accumulate, float *nlp_out, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (P <= 0) return 0;
    if (!params || !moves || !coef || !workspace || !grads || !nlp_out || bn_mode < 0 || bn_mode > 1 || prec < 0 || prec > 5 || prec == 3) return -1;
    if (set_attrs() != 0) return -3;
    const Ws w = ws_layout(P);
    float *ws = static_cast<float *>(workspace);
    const int M = P * NSQ;
    bk_train_transpose_kernel<<<dim3(4, 4, 54), dim3(32, 8), 0, st>>>(params + BK_TP_W1, ws + w.wd, 54);
    __nv_bfloat16 *wdb = bk_tc_lo_bf16() ? reinterpret_cast<__nv_bfloat16 *>(ws + w.wdb16) : nullptr;
    if (prec >= 4) bk_train_pack_w_kernel<<<(6 * 9 * C * C + 255) / 256, 256, 0, st>>>(ws + w.wd, ws + w.wdpack, ws + w.wdpack_lo, wdb, wdb ? wdb + 6 * 9 * C * C : nullptr, 6 * 9 * C * C);
    bk_train_head_bwd_kernel<<<P, 128, 0, st>>>(ws + w.logits, ws + w.a[6], params + BK_TP_HEADW, moves, coef, nlp_out,
                                                ws + w.dlogit, ws + w.da, ws + w.dwpart);
    bk_train_colsum_kernel<<<(C + 31) / 32, dim3(32, 8), 0, st>>>(ws + w.dwpart, grads + BK_TP_HEADW, P, C, accumulate);
    bk_train_colsum_kernel<<<(NSQ + 31) / 32, dim3(32, 8), 0, st>>>(ws + w.dlogit, grads + BK_TP_HEADB, P, NSQ, accumulate);
    for (int l = 6; l >= 0; --l) {
        // d a_l (in ws.da) -> d z_l (ws.dz) and the per-position contributions to d gamma / d beta / d conv bias
        bk_train_bn_bwd_kernel<<<P, C, 0, st>>>(ws + w.da, ws + w.a[l], ws + w.z[l], params + vec_off(l, 1), ws + w.mean[l],
                                                ws + w.rstd[l], ws + w.dz, ws + w.part, bn_mode);
        bk_train_colsum_kernel<<<(3 * C + 31) / 32, dim3(32, 8), 0, st>>>(ws + w.part, grads + vec_off(l, 0), P, 3 * C, accumulate);
        WgradArgs g;
        g.act = l == 0 ? ws + w.x0 : ws + w.a[l - 1];
        g.dz = ws + w.dz;
        g.part = ws + w.wpart;
        g.M = M;
        g.Cin = l == 0 ? C0 : C;
        g.R = l == 0 ? 5 : 3;
        g.K = g.R * g.R * g.Cin;
        g.rows_per_split = w.rows_per_split;
        launch_wgrad(g, w.splits, prec, st);
        const int nw = g.K * C;
        bk_train_reduce_kernel<<<(nw + 255) / 256, 256, 0, st>>>(ws + w.wpart, grads + w_off(l), nw, w.splits, accumulate);
        if (l > 0) {
            ConvArgs a;
            a.in = ws + w.dz;
            a.w = (prec >= 4 ? ws + w.wdpack : ws + w.wd) + (size_t)(l - 1) * 9 * C * C;
            a.w_lo = ws + w.wdpack_lo + (size_t)(l - 1) * 9 * C * C;
            if (wdb) {
                a.w_bh = reinterpret_cast<const unsigned short *>(wdb + (size_t)(l - 1) * 9 * C * C);
                a.w_bl = reinterpret_cast<const unsigned short *>(wdb + (size_t)(6 + l - 1) * 9 * C * C);
            }
            a.bias = nullptr;
            a.out = ws + w.da;
            a.M = M;
            a.Cin = C;
            a.R = 3;
            a.sign = -1;
            a.ksplit = 1;
            a.tail_part = ws + w.tail;
            launch_conv(a, prec, st);
        }
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int bk_train_running_stats(float *running, const float *stats, const int32_t *seq, int S, float momentum, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (S <= 0) return 0;
    if (!running || !stats) return -1;
    bk_train_running_kernel<<<(7 * 2 * C + 127) / 128, 128, 0, st>>>(running, stats, seq, S, momentum);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int bk_adamw_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, size_t n, double lr, double beta1,
                             double beta2, double eps, double weight_decay, int step, void *stream)
{
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) return 0;
    if (!params || !grads || !exp_avg || !exp_avg_sq || step < 1) return -1;
    double p1 = 1.0, p2 = 1.0;
    for (int i = 0; i < step; ++i) {
        p1 *= beta1;
        p2 *= beta2;
    }
    const double bc1 = 1.0 - p1, bc2 = 1.0 - p2;
    const double r = sqrt(bc2);
    bk_adamw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, (float)(1.0 - lr * weight_decay),
                                                                 (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)r,
                                                                 (float)eps, (float)(lr / bc1));
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
