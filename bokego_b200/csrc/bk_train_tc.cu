// bk_train_tc.cu -- the three GEMMs of the REINFORCE step (forward conv, data gradient, weight gradient; see bk_train.cu) on the
// 5th-generation tensor cores: tcgen05.mma kind::tf32 issued by one thread, operands staged in shared memory, accumulators in TMEM.
//
// One CTA computes a 128 x 128 output tile.  Warp roles (288 threads; measured alternatives that were slower: 8 + 8 producer / result warps
// at 96 registers, 8 result warps on column halves, 16-deep slabs in six stages for 3xTF32, a separate raw landing ring with three
// slabs of gathers in flight and two operand stages for 3xTF32):
//   warps 4-7  producers: gather the operands of a 32-deep slab with 16-byte cp.async (zero fill for taps that leave the board),
//              and -- in 3xTF32 mode -- split every value into a TF32 high part and a TF32 low part (hi = x cut to 10 mantissa
//              bits, lo = the exact remainder cut the same way: two ANDs, where cvt.rna costs a sequence of instructions);
//   warp  8    issues the MMAs of a slab (4 K steps of 8; 3xTF32: lo*hi, hi*lo, hi*hi per step) and commits the slab's buffers back;
//   warps 0-3  own the result: thread = output row.  The tensor core's fp32 accumulation truncates every time it adds to the
//              accumulator (measured: chains of 48 MMAs leave errors of 1e-3 in the data gradient, whose sums cancel heavily), so in
//              3xTF32 mode a chain is one slab (twelve MMAs; 3, 6 and 12 measured equally accurate): the MMA warp rotates through four TMEM accumulators, and these
//              warps add each finished chain into fp32 registers with IEEE adds while the next chains run -- the same reason
//              bk_train.cu keeps its mma.sync chains three MMAs short.  At the end they add the bias and store the rows.
// Operand layout in shared memory: every operand is K-major without swizzle, [chunk of 4 along the reduction][row][4 floats]
// (8 x 16-byte core matrices, SBO 128 B between 8-row groups, LBO 2048 B between the two reduction chunks of one K = 8 MMA).
// MN-major TF32 operands without swizzle are not accepted by the tensor core (tools/probes/tc_probe.cu: the MMA returns zeros),
// so whatever arrives with its GEMM-M / GEMM-N index contiguous is turned on the way in:
//   conv:  A = gathered activation rows [m][ci] -- K-major as they come, 16-byte cp.async;
//          B = weights, repacked and split into TF32 high / low parts once per step by bk_train_pack_w_kernel into
//              [k / 4][co][k % 4]: a slab is 16 KiB contiguous and already in operand layout -- one cp.async.bulk per part;
//   wgrad: A = activation rows [m][k index] and B = dZ rows [m][co], reduction over m: the producers load 4 rows x 4 columns into
//          registers, transpose the 4 x 4 block and store it (the 3xTF32 split happens in the same registers).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>

#include "bk_train_args.h"

namespace {

constexpr int C = 128;
constexpr int NSQ = 81;
constexpr int SLAB = 32;                 // reduction depth of one stage
constexpr int MAX_STAGES = 6;
// measurement only (results are wrong): 1 = no 3xTF32 split in the conv producers, 2 = no gathers / loads of the A and dZ operands,
// 4 = no MMAs (commits only), 8 = no TMEM read-out / adds in the result warps, 16 = no weight bulk copies, 32 = 3x3 kernel: operand
// windows on 8-row boundaries
#ifndef BK_TC_DIAG
#define BK_TC_DIAG 0
#endif
#ifndef BK_R3_LO_BF16
#define BK_R3_LO_BF16 0                            // measurement build, rejected: the low-order products (A_hi x B_lo, A_lo x B_hi: 2^-11 of the
#endif                                             // result) on kind::f16 with bf16 copies of the four parts, one K = 16 MMA where two TF32 MMAs
                                                   // run.  3.5 % off the step, but 2^-20 per product is not enough under the cancellation of the
                                                   // data gradient: nine in ten entries of conv.1.weight's gradient within 2.5e-4 instead of
                                                   // 1e-4 of the largest (test_reference_iteration fails; profiles/r02zz_lo_bf16.txt)
#ifndef BK_TC_CHAIN
#define BK_TC_CHAIN 4    // K steps per 3xTF32 accumulation chain = one slab (12 MMAs).  Measured on the recorded reference iteration: 1, 2 and 4
#endif                   // are equally accurate (1e-6 against FFMA) and 4 is 5 % faster; 16 (four slabs, 48 MMAs) leaves 1e-3 in the early layers
constexpr int NBUF = 4;                  // TMEM accumulators of 128 columns (all 512 columns)
constexpr int N_THREADS = 288;
constexpr int WARP_MMA = 8;
enum { BAR_FULL = 0, BAR_EMPTY = MAX_STAGES, BAR_ACCF = 2 * MAX_STAGES, BAR_ACCE = 2 * MAX_STAGES + NBUF, N_BARS = 2 * MAX_STAGES + 2 * NBUF };

template <int MODE, int PREC>
struct Sizes {
    // Byte distance between consecutive 8-row core matrices.  The conv operands are written one core matrix per quarter
    // warp (dense, 128 B); the weight-gradient operands are written transposed, one row of every second core matrix per
    // lane, and a pitch of 144 B spreads those 16-byte stores over all banks (ncu: 75 % of the wavefronts of the dense
    // version were bank conflicts).
    static constexpr int SBO = MODE == 0 ? 128 : 144;
    static constexpr int LBO = 16 * SBO;                             // between the K chunks (of 4) of a 128-row operand
    static constexpr int OP_BYTES = (SLAB / 4) * LBO;                // one operand tile of a stage: 16 / 18 KiB
    static constexpr int STAGE = (PREC ? 4 : 2) * OP_BYTES;          // A_hi, B_hi (, A_lo, B_lo)
    static constexpr int STAGES = PREC ? 3 : 6;
    static constexpr int LOOKAHEAD = PREC ? 1 : 3;                   // slabs of cp.async a producer keeps in flight; the other stages are slack for the MMAs
    static constexpr int CHAIN = PREC ? BK_TC_CHAIN : 4;                       // K steps (of 8) accumulated in the tensor core before the fp32 add
    static constexpr int SMEM = STAGES * STAGE + 256;
    static constexpr uint32_t DESC_HI = ((uint32_t)SBO >> 4) | (1u << 14);   // SBO, descriptor version 1
};

// instruction descriptor: D = f32 (bit 4), A = B = TF32 (2 at bits 7, 10), majors at bits 15 / 16 (1 = MN-major), N >> 3 at 17, M >> 4 at 24
constexpr uint32_t IDESC_BASE = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
// kind::f16 with bf16 operands (format 1 at bits 7, 10), D = f32, K-major, N = 128, M = 128: one MMA covers K = 16
constexpr uint32_t IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_N256 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);   // N = 256

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must end in a trap (launch error), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000LL) asm volatile("trap;");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *g, bool valid)
{
    const int sz = valid ? 16 : 0;                            // 0 source bytes = 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// contiguous global -> shared bulk copy (async proxy); its bytes complete on the barrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// completion of all MMAs issued so far by this thread -> one arrival on the barrier
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
#if BK_R3_LO_BF16
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
#endif
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// the producers' split: high part = the value cut to TF32 (what the tensor core would do with the raw word), low part = the exact
// remainder cut to TF32.  One AND each instead of cvt.rna's multi-instruction sequence; what is dropped is below 2^-21 of the value.
__device__ __forceinline__ float tf32_cut(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t hi)
{
    return ((uint64_t)hi << 32) | (((saddr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16));
}

#ifndef BK_R3_PROF
#define BK_R3_PROF 0                               // measurement build: clock64 stamps of CTA 0 (tools/prof_train_conv3.py)
#endif
#if BK_R3_PROF
__device__ long long g_r3_prof[5 * 1024];          // [role: MMA issue / weight loader / result of the 3x3 kernel, MMA issue / producer warp 4 of a 3x3 weight gradient][use < 256][4 stamps]
#define R3_STAMP(role, idx, k) do { if (blockIdx.x == 0 && (idx) < 256) g_r3_prof[(role) * 1024 + (idx) * 4 + (k)] = clock64(); } while (0)
#else
#define R3_STAMP(role, idx, k) do {} while (0)
#endif

// The result rows of a 128 x 128 tile leave through a per-warp transposition buffer (32 x 33 floats).  A thread owns a ROW of 128
// floats, and stored from there a warp instruction touches 32 different lines, 16 bytes each: 12 k cycles per tile
// (profiles/r02u_train_conv3.md).  Instead, 32 columns at a time: thread = row writes its 32 values (pitch 33: conflict-free), then
// lane = column reads the 32 rows back -- ALL loads before the first store; a loop of load / store pairs took as long as the direct
// stores, 110 cycles per row -- and every store is one whole 128-byte line of one row: 5 k cycles per tile.
// at = this thread's output row (in rows of 128 floats from obase), -1 = none; bias = one value per column or null.
__device__ __forceinline__ void store_rows_transposed(const float (&acc)[128], float *sst, float *obase, int at, const float *bias, int lane)
{
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const float bv = bias ? bias[ch * 32 + lane] : 0.0f;
#pragma unroll
        for (int i = 0; i < 32; ++i) sst[lane * 33 + i] = acc[ch * 32 + i];
        __syncwarp();
        float v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) v[r] = sst[r * 33 + lane] + bv;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int at_r = __shfl_sync(0xffffffffu, at, r);
            if (at_r >= 0) obase[(size_t)at_r * 128 + ch * 32 + lane] = v[r];
        }
        __syncwarp();
    }
}

// MODE 0: convolution / data gradient (BkConvArgs).  MODE 1: weight gradient (BkWgradArgs), grid = (k tiles, splits).
template <int MODE, int PREC, typename Args>
__global__ void __launch_bounds__(N_THREADS, 1) bk_train_gemm_tc_kernel(const Args a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    using SZ = Sizes<MODE, PREC>;
    constexpr int STAGE = SZ::STAGE, STAGES = SZ::STAGES, LOOKAHEAD = SZ::LOOKAHEAD, CHAIN = SZ::CHAIN;
    constexpr int OP_BYTES = SZ::OP_BYTES, SBO = SZ::SBO, LBO = SZ::LBO;
    constexpr uint32_t DESC_HI_K = SZ::DESC_HI;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_bar = s_base + STAGES * STAGE;
    const uint32_t s_tmem = s_bar + 8 * N_BARS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- the reduction range of this CTA ----
    int KT;                       // slabs
    int m0 = 0, r_lo = 0, r_hi = 0, kb = 0;
    if constexpr (MODE == 0) {
        m0 = blockIdx.x * 128;
        KT = a.R * a.R * a.Cin / SLAB;
    } else {
        kb = blockIdx.x;
        r_lo = blockIdx.y * a.rows_per_split;
        r_hi = min(a.M, r_lo + a.rows_per_split);
        KT = r_hi > r_lo ? (r_hi - r_lo + SLAB - 1) / SLAB : 0;
    }
    const int n_chains = KT * (SLAB / 8) / CHAIN;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(s_bar + 8 * (BAR_FULL + s), MODE == 0 ? 129 : 128);    // every producer thread (+ the weight slab's expect_tx)
            mbar_init(s_bar + 8 * (BAR_EMPTY + s), 1);     // tcgen05.commit
        }
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(s_bar + 8 * (BAR_ACCF + b), 1);      // tcgen05.commit
            mbar_init(s_bar + 8 * (BAR_ACCE + b), 128);    // every result thread
        }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(s_tmem, 128 * NBUF);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + STAGES * STAGE + 8 * N_BARS);

    if (warp >= 4 && warp < 8) {
        // =============================== producers ===============================
        const int pw = warp - 4, ptid = tid - 128;
        const int half = a.R >> 1;
        if constexpr (MODE == 0) {
            // ---- conv: both operands by 16-byte cp.async; chunk i of this thread = (k chunk (i >> 2) * 4 + q4, row (pw + 4 (i & 3)) * 8 + r8)
            const int r8 = lane & 7, q4 = lane >> 3;
            uint32_t off[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) off[i] = (uint32_t)(((i >> 2) * 4 + q4) * LBO + ((pw + 4 * (i & 3)) * 8 + r8) * 16);
            int rbase[4], rx[4], ry[4];
            bool rvalid[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int m = m0 + (pw + 4 * q) * 8 + r8;
                rvalid[q] = m < a.M;
                const int p = m / NSQ, sq = m - p * NSQ;
                rx[q] = sq / 9;
                ry[q] = sq - 9 * rx[q];
                rbase[q] = p * NSQ;
            }
            auto issue = [&](int kt) {
                const uint32_t st = s_base + (uint32_t)((kt % STAGES) * STAGE);
                const int k0 = kt * SLAB;
                const int tap = k0 / a.Cin, c0 = k0 - tap * a.Cin;
                const int ti = tap / a.R;
                const int dx = a.sign * (ti - half), dy = a.sign * (tap - ti * a.R - half);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int q = i & 3, kc = (i >> 2) * 4 + q4;
                    const int x = rx[q] + dx, y = ry[q] + dy;
                    const bool ok = rvalid[q] && (unsigned)x < 9u && (unsigned)y < 9u;
                    const float *src = ok ? a.in + ((size_t)(rbase[q] + 9 * x + y) * a.Cin + c0 + 4 * kc) : a.in;
                    if (!(BK_TC_DIAG & 2)) cp_async16(st + off[i], src, ok);
                }
                if (ptid == 0) {
                    // the weight slab: 8 K chunks x 128 co x 16 bytes, contiguous in the packed (and pre-split) weights and
                    // already in the K-major operand layout: one bulk copy per part, no thread touches it
                    const uint32_t bar = s_bar + 8 * (BAR_FULL + kt % STAGES);
                    if (BK_TC_DIAG & 16) {
                        mbar_arrive(bar);
                    } else {
                        mbar_arrive_expect_tx(bar, PREC ? 2u * OP_BYTES : (uint32_t)OP_BYTES);
                        bulk_g2s(st + OP_BYTES, a.w + (size_t)(k0 / 4) * (4 * C), OP_BYTES, bar);
                        if (PREC) bulk_g2s(st + 3 * OP_BYTES, a.w_lo + (size_t)(k0 / 4) * (4 * C), OP_BYTES, bar);
                    }
                }
            };
            auto publish = [&](int kt) {
                const int s = kt % STAGES;
                if constexpr (PREC != 0 && !(BK_TC_DIAG & 1)) {
                    uint8_t *st = smem + s * STAGE;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {                 // the activation chunks this thread gathered
                        const uint32_t o = off[i];
                        const float4 v = *reinterpret_cast<const float4 *>(st + o);   // stays as the high part: the tensor core cuts it
                        float4 l;
                        l.x = v.x - tf32_cut(v.x); l.y = v.y - tf32_cut(v.y); l.z = v.z - tf32_cut(v.z); l.w = v.w - tf32_cut(v.w);
                        *reinterpret_cast<float4 *>(st + 2 * OP_BYTES + o) = l;
                    }
                }
                fence_proxy_async();                          // the tensor core reads shared memory through the async proxy
                mbar_arrive(s_bar + 8 * (BAR_FULL + s));
            };
            for (int kt = 0; kt < KT + LOOKAHEAD; ++kt) {
                if (kt < KT) {
                    if (kt >= STAGES) mbar_wait(s_bar + 8 * (BAR_EMPTY + kt % STAGES), ((kt / STAGES) & 1) ^ 1);
                    issue(kt);
                }
                cp_commit();
                if (kt >= LOOKAHEAD) {
                    cp_wait<LOOKAHEAD>();
                    publish(kt - LOOKAHEAD);
                }
            }
        } else {
            // ---- wgrad: 4 x 4 blocks (4 rows m x 4 consecutive columns) through registers, transposed on the way.
            // Block (j, mc): columns 4j .. 4j+3, rows 4 mc .. 4 mc + 3 of the slab.  This thread: j = lane, mc = pw and pw + 4,
            // for A (activation rows, column = k index) and for B (dZ rows, column = co).
            const int j = ptid & 31;
            const int kglob = kb * 128 + 4 * j;
            const int tap = kglob / a.Cin, ci = kglob - tap * a.Cin;
            const bool tap_ok = tap < a.R * a.R;
            const int ti = tap / a.R;
            const int dx = ti - half, dy = tap - ti * a.R - half;
            float4 va[2][4], vb[2][4], na[2][4], nb[2][4];
            auto fetch = [&](int kt, float4 (&A)[2][4], float4 (&B)[2][4]) {
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    // four consecutive rows: one division for the first, the others by stepping (x, y); the shifted row of a valid tap
                    // is row m + 9 dx + dy of the dense layout (the producers are instruction-bound: every row used to divide twice)
                    const int mf = r_lo + kt * SLAB + 4 * (pw + 4 * u);
                    const int sq0 = mf - (mf / NSQ) * NSQ;
                    int x = sq0 / 9, y = sq0 - 9 * x;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int m = mf + r;
                        const bool in_rows = m < r_hi;
                        const bool ok = in_rows && tap_ok && (unsigned)(x + dx) < 9u && (unsigned)(y + dy) < 9u;
                        A[u][r] = (ok && !(BK_TC_DIAG & 2)) ? __ldg(reinterpret_cast<const float4 *>(a.act + ((size_t)(m + 9 * dx + dy) * a.Cin + ci))) : zero;
                        B[u][r] = (in_rows && !(BK_TC_DIAG & 2)) ? __ldg(reinterpret_cast<const float4 *>(a.dz + (size_t)m * C + 4 * j)) : zero;
                        if (++y == 9) {
                            y = 0;
                            if (++x == 9) x = 0;
                        }
                    }
                }
            };
            // (starting each lane at a different column of its block makes the 16-byte stores below conflict-free on paper; measured,
            // the kernel was 25 % slower with it, so every lane stores its columns in order)
            constexpr int rot = 0;
            if (KT > 0) fetch(0, va, vb);
            // Clock stamps of a slab (profiles/r02z_wgrad_clocks.txt; 1,820 cycles against 1,250 of tensor time): issuing the 16 loads of
            // the next slab 530 (32 KiB per slab and SM is what the L2 port delivers in that time; the warp waits in the issue),
            // transposing and storing 860, the register copies at the end 290.  Tried and slower: half of the loads between the stores
            // (2,470 per slab: the later loads are then waited for), two register sets that swap roles instead of the copies
            // (+0.5 % on the step).
            for (int kt = 0; kt < KT; ++kt) {
                const int s = kt % STAGES;
                const bool stamp = PREC != 0 && a.Cin == C && blockIdx.y == 0 && ptid == 0;
                if (stamp) R3_STAMP(4, kt, 0);
                if (kt + 1 < KT) fetch(kt + 1, na, nb);           // the next slab's loads are in flight while this one is stored
                if (stamp) R3_STAMP(4, kt, 1);
                if (kt >= STAGES) mbar_wait(s_bar + 8 * (BAR_EMPTY + s), ((kt / STAGES) & 1) ^ 1);
                if (stamp) R3_STAMP(4, kt, 2);
                uint8_t *st = smem + s * STAGE;
                auto put = [&](const float4 (&v)[4], uint32_t base) {
                    // column c of the block = the four rows' c-th components: one 16-byte K chunk of row (4 j + c)
                    const float4 c0 = make_float4(v[0].x, v[1].x, v[2].x, v[3].x), c1 = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
                    const float4 c2 = make_float4(v[0].z, v[1].z, v[2].z, v[3].z), c3 = make_float4(v[0].w, v[1].w, v[2].w, v[3].w);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int c = (t + rot) & 3;
                        float4 h = c == 0 ? c0 : (c == 1 ? c1 : (c == 2 ? c2 : c3));
                        if constexpr (PREC != 0) {
                            // the tensor core cuts the words it reads to TF32 itself: the raw value serves as the high part and the
                            // exact remainder as the low part, bit for bit what explicit cuts of both would give
                            float4 l;
                            l.x = h.x - tf32_cut(h.x); l.y = h.y - tf32_cut(h.y); l.z = h.z - tf32_cut(h.z); l.w = h.w - tf32_cut(h.w);
                            *reinterpret_cast<float4 *>(st + 2 * OP_BYTES + base + ((4 * j + c) >> 3) * SBO + ((4 * j + c) & 7) * 16) = l;
                        }
                        *reinterpret_cast<float4 *>(st + base + ((4 * j + c) >> 3) * SBO + ((4 * j + c) & 7) * 16) = h;
                    }
                };
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    put(va[u], (uint32_t)((pw + 4 * u) * LBO));
                    put(vb[u], (uint32_t)(OP_BYTES + (pw + 4 * u) * LBO));
                }
                fence_proxy_async();
                mbar_arrive(s_bar + 8 * (BAR_FULL + s));
                if (stamp) R3_STAMP(4, kt, 3);
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        va[u][r] = na[u][r];
                        vb[u][r] = nb[u][r];
                    }
            }
        }
    } else if (warp == WARP_MMA) {
        // =============================== MMA issue ===============================
        int c = 0;                                     // chain counter: chain c accumulates in TMEM buffer c % NBUF
        for (int kt = 0; kt < KT; ++kt) {
            const int s = kt % STAGES;
            bool stamp = false;                       // (blockIdx.x == 0 inside the macro)
            if constexpr (MODE == 1 && PREC != 0) stamp = a.Cin == C && blockIdx.y == 0 && lane == 0;
            if (stamp) R3_STAMP(3, kt, 0);
            mbar_wait(s_bar + 8 * (BAR_FULL + s), (kt / STAGES) & 1);
            if (stamp) R3_STAMP(3, kt, 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t st = s_base + (uint32_t)(s * STAGE);
                int cc = c;
#pragma unroll
                for (int g = 0; g < SLAB / 8; ++g) {
                    const int b = cc % NBUF;
                    const bool first = g % CHAIN == 0;
                    if (first) {
                        mbar_wait(s_bar + 8 * (BAR_ACCE + b), ((cc / NBUF) & 1) ^ 1);   // the result warps have drained this accumulator
                        tc_fence_after();
                    }
                    const uint32_t d = tmem + (uint32_t)(b * 128);
                    const uint64_t ah = make_desc(st + g * 2 * LBO, LBO, DESC_HI_K);
                    const uint64_t bh = make_desc(st + OP_BYTES + g * 2 * LBO, LBO, DESC_HI_K);
                    const uint64_t al = make_desc(st + 2 * OP_BYTES + g * 2 * LBO, LBO, DESC_HI_K);      // 3xTF32 only
                    const uint64_t bl = make_desc(st + 3 * OP_BYTES + g * 2 * LBO, LBO, DESC_HI_K);
                    constexpr uint32_t idesc = IDESC_BASE;
                    if constexpr ((BK_TC_DIAG & 4) != 0) {
                    } else if constexpr (PREC != 0) {
                        umma_tf32(d, al, bh, idesc, first ? 0u : 1u);
                        umma_tf32(d, ah, bl, idesc, 1u);
                        umma_tf32(d, ah, bh, idesc, 1u);
                    } else {
                        umma_tf32(d, ah, bh, idesc, first ? 0u : 1u);
                    }
                    if (g % CHAIN == CHAIN - 1) {
                        umma_commit(s_bar + 8 * (BAR_ACCF + b));          // this chain is complete
                        ++cc;
                    }
                }
                umma_commit(s_bar + 8 * (BAR_EMPTY + s));                 // slab consumed -> the producers may refill it
            }
            c += (SLAB / 8) / CHAIN;
            __syncwarp();
            if (stamp) R3_STAMP(3, kt, 2);
        }
    } else {
        // =============================== result warps (thread = output row) ===============================
        float acc[C];
#pragma unroll
        for (int i = 0; i < C; ++i) acc[i] = 0.0f;
        const uint32_t t_lane = tmem + ((uint32_t)(32 * warp) << 16);
        for (int c = 0; c < n_chains; ++c) {
            const int b = c % NBUF;
            mbar_wait(s_bar + 8 * (BAR_ACCF + b), (c / NBUF) & 1);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                uint32_t v[32];
                if (BK_TC_DIAG & 8) continue;
                tmem_ld32(t_lane + (uint32_t)(b * 128 + h * 32), v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[h * 32 + i] += __uint_as_float(v[i]);
            }
            tc_fence_before();
            mbar_arrive(s_bar + 8 * (BAR_ACCE + b));
        }
        // (thread = row stores; through the transposition buffer of the 3x3 kernel this epilogue, which nothing overlaps, was 2 % slower)
        const int row = 32 * warp + lane;
        if constexpr (MODE == 0) {
            const int m = m0 + row;
            if (m < a.M) {
                float4 *o = reinterpret_cast<float4 *>(a.out + (size_t)m * C);
#pragma unroll
                for (int i = 0; i < C / 4; ++i) {
                    float4 v = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                    if (a.bias) {
                        const float4 bb = *reinterpret_cast<const float4 *>(a.bias + 4 * i);
                        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                    }
                    o[i] = v;
                }
            }
        } else {
            const int k = kb * 128 + row;
            if (k < a.K) {
                float4 *o = reinterpret_cast<float4 *>(a.part + ((size_t)blockIdx.y * a.K + k) * C);
#pragma unroll
                for (int i = 0; i < C / 4; ++i) o[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem, 128 * NBUF);
    }
}


// ------------------------------------------------------------------------------------------------------------------------------
// 3x3 convolution / data gradient with the activation tile staged ONCE (the structural fix the stage-by-stage breakdown of the
// kernel above asks for; profiles/r01h_tc_stage_breakdown.md).  GEMM rows are the rows of a padded raster: position p, square (x, y)
// sits at row 100 p + 10 + 10 x + y -- ten zero rows above every board, one zero column to its right -- so a tap (dx, dy) is the row
// shift 10 dx + dy of ONE staged operand, exactly as in the inference kernel, and zero padding comes from the zero rows / column.
// A tile is 128 consecutive raster rows: rows R0 - 11 .. R0 + 138 (all 128 channels) are gathered once -- 4,800 16-byte chunks
// instead of the 36,864 of nine separate window gathers -- in four channel groups of 32 that the MMAs follow group by group; beside
// that only the pre-split weight slabs stream (one bulk copy each).  81 of every 100 raster rows are real squares; the result
// threads write those to the dense [P][81][128] output.
// Round 2 (profiles/r02u_train_conv3.md):
//   * PERSISTENT CTAs, one per SM, each working through its tiles as ONE pipeline: with one tile per CTA (round 1) the prologue
//     (barriers, TMEM allocation, the first group's gather, the first weight slabs) and the epilogue (last read-out, stores) of
//     every tile were exposed -- 57 k cycles per tile against 28 k of tensor time, and switching the MMAs off saved only a third.
//     Now all roles run loops over the CTA's work items with ring positions that carry over from item to item: the staging warps
//     gather the first groups of the next tile into free group buffers (three rotating buffers) while the MMAs of the current one
//     run, the weight loader never stops, and the result warps store tile t while the tensor pipe is in the first chains of t + 1;
//   * a kind::tf32 MMA of 128 x 128 x 8 (cta_group::1, both operands in shared memory) takes ~104 cycles here, not 64 (clock stamps:
//     the issuing thread waits on no barrier; windows on 8-row boundaries / no weight copies / other accumulators change nothing).
//     The two products that share the operand A_hi are ONE MMA with N = 256, A_hi x [B_hi | B_lo] -> the [hi | lo] halves of a
//     256-column accumulator, and A_lo x B_hi goes into the lo half: two MMAs per K step instead of three.  The N = 256 form takes
//     ~190 cycles -- the cost follows the tensor work, not the operand bytes -- so this bought 5 % per slab, and what it really gave is
//     the next point;
//   * the accumulation chains are split by magnitude.  What the tensor core's truncating fp32 accumulation costs is one ulp of the
//     ACCUMULATOR per MMA, so the large hi * hi products are kept apart from the lo * hi and hi * lo products (2^-11 of them);
//     a chain is BK_R3_CHAIN_SLABS slabs = 12 accumulations per half -- as many as the 12-MMA chains measured equivalent to
//     one-K-step chains -- and the result warps add both halves of a finished chain into fp32 registers with IEEE adds: 12
//     read-outs per tile instead of 36;
//   * work items: whole tiles; for small batches (ksplit) one channel group of a tile each; and the tiles left over after the last
//     full round of the grid as four channel-group items each (see BkConvArgs::tail_part).
// Warps: 0-3 result (thread = raster row), 4-7 staging, 8 MMA issue, 9 weight loader (one thread).
// Barrier rule (profiles/r02_handover_experiments.md): every waiter consumes every phase of the barriers it waits on -- ring
// position and phase advance together, once per use, in every role.
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int R3_ROWS = 152;                       // staged rows: 128 + 2 * 11 halo, rounded up to 8
constexpr int R3_LBO = R3_ROWS * 16;               // bytes between K chunks of the staged tile
constexpr int R3_G_BYTES = 8 * R3_LBO;             // one part (hi or lo) of one channel group of the tile: 19,456
constexpr int R3_THREADS = 320;
constexpr int R3_NA = 3;                           // rotating group buffers
constexpr int R3_NROT = 2;                         // two accumulators of 256 columns ([hi | lo] halves), the chains take them in turn
#ifndef BK_R3_CHAIN_SLABS
#define BK_R3_CHAIN_SLABS 3                        // slabs per hi * hi chain (4 MMAs each)
#endif
template <int PREC>
struct R3 {
    static constexpr int NW = PREC ? 3 : 6;                          // weight stages
    static constexpr int W_STAGE = (PREC ? 2 : 1) * 16384;           // B_hi (, B_lo) of a 32-deep slab
    static constexpr int ABUF = (PREC ? 2 : 1) * R3_G_BYTES;         // one group buffer: A_hi (, A_lo)
    static constexpr int A_BYTES = R3_NA * ABUF;
    static constexpr int SMEM = A_BYTES + NW * W_STAGE + 256 + 4 * 32 * 33 * 4;      // + the result warps' transposition buffers
    static constexpr int CHAIN = BK_R3_CHAIN_SLABS;
    static_assert(SMEM <= 227 * 1024, "shared memory");
};
enum { R3_AFULL = 0, R3_AEMPTY = R3_NA, R3_WFULL = 2 * R3_NA, R3_WEMPTY = R3_WFULL + MAX_STAGES, R3_ACCF = R3_WEMPTY + MAX_STAGES,
       R3_ACCE = R3_ACCF + R3_NROT, R3_NBARS = R3_ACCE + R3_NROT };
static_assert(8 * R3_NBARS + 4 <= 256, "barrier area");

#ifndef BK_R3_STORE
#define BK_R3_STORE 2                              // measurement builds: 0 = no output stores, 1 = thread = row stores
#endif
struct R3Ring {                                    // position in a ring of n buffers and the phase of its current use
    int i;
    uint32_t ph;
    __device__ __forceinline__ void next(int n)
    {
        if (++i == n) {
            i = 0;
            ph ^= 1u;
        }
    }
};
struct R3Item { int tile, g_lo, ng, kind; };       // kind: 0 = whole tile, 1 = one group of a small batch (ksplit), 2 = one group of a tail tile
__device__ __forceinline__ R3Item r3_item(const BkConvArgs &a, int it)
{
    R3Item w;
    if (a.ksplit > 1) {
        w.tile = it >> 2; w.g_lo = it & 3; w.ng = 1; w.kind = 1;
    } else if (a.tail_full > 0 && it >= a.tail_full) {
        const int t = it - a.tail_full;
        w.tile = a.tail_full + (t >> 2); w.g_lo = t & 3; w.ng = 1; w.kind = 2;
    } else {
        w.tile = it; w.g_lo = 0; w.ng = 4; w.kind = 0;
    }
    return w;
}

template <int PREC>
__global__ void __launch_bounds__(R3_THREADS, 1) bk_train_conv3_tc_kernel(const BkConvArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    using Z = R3<PREC>;
    constexpr int NW = Z::NW, W_STAGE = Z::W_STAGE, CHAIN = Z::CHAIN;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_w = s_base + Z::A_BYTES;
    const uint32_t s_bar = s_w + NW * W_STAGE;
    const uint32_t s_tmem = s_bar + 8 * R3_NBARS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P = a.M / NSQ;

    if (tid == 0) {
        for (int i = 0; i < R3_NA; ++i) {
            mbar_init(s_bar + 8 * (R3_AFULL + i), 128);        // every staging thread
            mbar_init(s_bar + 8 * (R3_AEMPTY + i), 1);         // tcgen05.commit
        }
        for (int s = 0; s < NW; ++s) {
            mbar_init(s_bar + 8 * (R3_WFULL + s), 1);          // the loader's expect_tx
            mbar_init(s_bar + 8 * (R3_WEMPTY + s), 1);         // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(s_bar + 8 * (R3_ACCF + b), 1);           // tcgen05.commit
            mbar_init(s_bar + 8 * (R3_ACCE + b), 128);         // every result thread
        }
        fence_barrier_init();
    }
    if (warp == 4) tmem_alloc(s_tmem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(smem + Z::A_BYTES + NW * W_STAGE + 8 * R3_NBARS);

    if (warp >= 4 && warp < 8) {
        // =============================== staging of the activation tiles, group by group ===============================
        const int pw = warp - 4, r8 = lane & 7, q4 = lane >> 3;
        R3Ring ra = {0, 0};
        for (int it = blockIdx.x; it < a.n_items; it += gridDim.x) {
            const R3Item w = r3_item(a, it);
            const int R0 = w.tile * 128;                                       // first raster row of the tile
            for (int g = w.g_lo; g < w.g_lo + w.ng; ++g) {
                mbar_wait(s_bar + 8 * (R3_AEMPTY + ra.i), ra.ph ^ 1u);         // the MMAs of the buffer's previous group have finished
                const int abuf = ra.i * Z::ABUF;
                // a block = 8 rows x 4 K chunks; 19 row groups x 2 chunk halves per channel group of 8 chunks
                for (int b = pw; b < 38; b += 4) {
                    const int row = (b >> 1) * 8 + r8, kl = (b & 1) * 4 + q4, kc = g * 8 + kl;
                    const int rr = R0 - 11 + row;                              // raster row
                    const int p = rr >= 0 ? rr / 100 : -1;
                    const int o = rr - 100 * p - 10;                           // 10 x + y, negative in the zero rows above the board
                    const int x = o / 10, y = o - 10 * x;
                    const bool ok = rr >= 0 && p < P && o >= 0 && y < 9 && row < 150;
                    const float *src = ok ? a.in + ((size_t)(p * NSQ + 9 * x + y) * C + 4 * kc) : a.in;
                    if (!(BK_TC_DIAG & 2)) cp_async16(s_base + (uint32_t)(abuf + kl * R3_LBO + row * 16), src, ok);
                }
                cp_commit();
                cp_wait<0>();
                if constexpr (PREC != 0) {
                    for (int b = pw; b < 38; b += 4) {
                        const int row = (b >> 1) * 8 + r8, kl = (b & 1) * 4 + q4;
                        // the raw words stay where they landed as the high part (the tensor core cuts what it reads to TF32); the low
                        // part is the exact remainder, cut by the tensor core as well
                        const float4 v = *reinterpret_cast<const float4 *>(smem + abuf + kl * R3_LBO + row * 16);
                        float4 l;
                        l.x = v.x - tf32_cut(v.x); l.y = v.y - tf32_cut(v.y); l.z = v.z - tf32_cut(v.z); l.w = v.w - tf32_cut(v.w);
#if BK_R3_LO_BF16
                        // bf16 copies of the value and of its low part: four K chunks of 8 channels, [chunk][row][8 x bf16]; this
                        // thread's four channels are one half of a chunk's row
                        const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
                        const __nv_bfloat162 l01 = __floats2bfloat162_rn(l.x, l.y), l23 = __floats2bfloat162_rn(l.z, l.w);
                        uint8_t *bp = smem + abuf + R3_G_BYTES + (kl >> 1) * R3_LBO + row * 16 + (kl & 1) * 8;
                        *reinterpret_cast<uint2 *>(bp) = make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
                        *reinterpret_cast<uint2 *>(bp + R3_G_BYTES / 2) = make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
#else
                        *reinterpret_cast<float4 *>(smem + abuf + R3_G_BYTES + kl * R3_LBO + row * 16) = l;
#endif
                    }
                }
                fence_proxy_async();
                mbar_arrive(s_bar + 8 * (R3_AFULL + ra.i));
                ra.next(R3_NA);
            }
        }
    } else if (warp == 9) {
        // =============================== weight slabs: one bulk copy per part ===============================
        if (lane == 0) {
            R3Ring rw = {0, 0};
            int n_slab = 0;
            for (int it = blockIdx.x; it < a.n_items; it += gridDim.x) {
                const R3Item w = r3_item(a, it);
                for (int s = 0; s < 9 * w.ng; ++s, ++n_slab) {
                    R3_STAMP(1, n_slab, 0);
                    mbar_wait(s_bar + 8 * (R3_WEMPTY + rw.i), rw.ph ^ 1u);     // the MMAs of the stage's previous slab have finished
                    R3_STAMP(1, n_slab, 1);
                    const int g = w.g_lo + s / 9, tap = s % 9;
                    const int k0 = tap * C + 32 * g;
                    const uint32_t bar = s_bar + 8 * (R3_WFULL + rw.i), dst = s_w + (uint32_t)(rw.i * W_STAGE);
                    if (BK_TC_DIAG & 16) {
                        mbar_arrive(bar);
                    } else {
                        mbar_arrive_expect_tx(bar, (uint32_t)W_STAGE);
                        if (PREC && BK_R3_LO_BF16) {        // [B_hi tf32: 16 KiB][B_hi bf16: 8 KiB][B_lo bf16: 8 KiB]
                            bulk_g2s(dst, a.w + (size_t)(k0 / 4) * (4 * C), 16384, bar);
                            bulk_g2s(dst + 16384, a.w_bh + (size_t)(k0 / 8) * (8 * C), 8192, bar);
                            bulk_g2s(dst + 24576, a.w_bl + (size_t)(k0 / 8) * (8 * C), 8192, bar);
                        } else if (PREC) {         // K chunk by K chunk [B_hi: 128 rows | B_lo: 128 rows]: one 256-row operand
                            for (int c = 0; c < 8; ++c) {
                                bulk_g2s(dst + c * 4096, a.w + (size_t)(k0 / 4 + c) * (4 * C), 2048, bar);
                                bulk_g2s(dst + c * 4096 + 2048, a.w_lo + (size_t)(k0 / 4 + c) * (4 * C), 2048, bar);
                            }
                        } else {
                            bulk_g2s(dst, a.w + (size_t)(k0 / 4) * (4 * C), 16384, bar);
                        }
                    }
                    rw.next(NW);
                }
            }
        }
    } else if (warp == 8) {
        // =============================== MMA issue ===============================
        R3Ring ra = {0, 0}, rw = {0, 0}, rc = {0, 0};
        int n_slab = 0;
        for (int it = blockIdx.x; it < a.n_items; it += gridDim.x) {
            const R3Item w = r3_item(a, it);
            const int KT = 9 * w.ng;
            for (int s = 0; s < KT; ++s, ++n_slab) {
                const int gi = s / 9, tap = s - 9 * gi;
                const bool first = s % CHAIN == 0, last = s % CHAIN == CHAIN - 1 || s == KT - 1;
                if (lane == 0) R3_STAMP(0, n_slab, 0);
                if (tap == 0) mbar_wait(s_bar + 8 * (R3_AFULL + ra.i), ra.ph);
                if (lane == 0) R3_STAMP(0, n_slab, 1);
                mbar_wait(s_bar + 8 * (R3_WFULL + rw.i), rw.ph);
                if (lane == 0) R3_STAMP(0, n_slab, 2);
                if (first) mbar_wait(s_bar + 8 * (R3_ACCE + rc.i), rc.ph ^ 1u);                // the result warps have read the chain two back
                if (lane == 0) R3_STAMP(0, n_slab, 3);
                tc_fence_after();
                if (elect_one()) {
                    const int ti = tap / 3;
                    const int shift = (BK_TC_DIAG & 32) ? -11 + 8 * (tap % 3) : a.sign * (10 * (ti - 1) + (tap - 3 * ti - 1));   // 32: windows on 8-row boundaries
                    const uint32_t a0 = s_base + (uint32_t)(ra.i * Z::ABUF + (11 + shift) * 16), w0 = s_w + (uint32_t)(rw.i * W_STAGE);
                    const uint32_t d = tmem + (uint32_t)(rc.i * 256);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t akc = (uint32_t)(2 * ks * R3_LBO);
                        const uint32_t acc_flag = (first && ks == 0) ? 0u : 1u;
                        const uint64_t ah = make_desc(a0 + akc, R3_LBO, (128u >> 4) | (1u << 14));
                        if constexpr ((BK_TC_DIAG & 4) != 0) {
                        } else if constexpr (PREC != 0 && BK_R3_LO_BF16 != 0) {
                            // hi half += A_hi x B_hi on kind::tf32 (the raw words); lo half += A_hi x B_lo + A_lo x B_hi on kind::f16, K = 16
                            // per MMA: K steps 0 and 2 of the slab issue the two bf16 MMAs of their half of the slab
                            umma_tf32(d, ah, make_desc(w0 + ks * 4096, 2048, (128u >> 4) | (1u << 14)), IDESC_BASE, acc_flag);
                            if ((ks & 1) == 0) {
                                const uint32_t ab = a0 + R3_G_BYTES + (uint32_t)(ks * R3_LBO);         // bf16 chunks ks, ks + 1 of the group's four
                                const uint64_t a_hb = make_desc(ab, R3_LBO, (128u >> 4) | (1u << 14));
                                const uint64_t a_lb = make_desc(ab + R3_G_BYTES / 2, R3_LBO, (128u >> 4) | (1u << 14));
                                const uint64_t b_hb = make_desc(w0 + 16384 + ks * 2048, 2048, (128u >> 4) | (1u << 14));
                                const uint64_t b_lb = make_desc(w0 + 24576 + ks * 2048, 2048, (128u >> 4) | (1u << 14));
                                umma_bf16(d + 128u, a_hb, b_lb, IDESC_BF16, acc_flag);
                                umma_bf16(d + 128u, a_lb, b_hb, IDESC_BF16, 1u);
                            }
                        } else if constexpr (PREC != 0) {
                            const uint64_t al = make_desc(a0 + R3_G_BYTES + akc, R3_LBO, (128u >> 4) | (1u << 14));
                            const uint64_t bw = make_desc(w0 + ks * 8192, 4096, (128u >> 4) | (1u << 14));     // rows 0..127 = B_hi, 128..255 = B_lo
                            umma_tf32(d, ah, bw, IDESC_N256, acc_flag);              // [hi | lo] (+)= A_hi x [B_hi | B_lo]
                            umma_tf32(d + 128u, al, bw, IDESC_BASE, 1u);             // lo += A_lo x B_hi
                        } else {
                            umma_tf32(d, ah, make_desc(w0 + ks * 4096, 2048, (128u >> 4) | (1u << 14)), IDESC_BASE, acc_flag);
                        }
                    }
                    umma_commit(s_bar + 8 * (R3_WEMPTY + rw.i));                            // slab consumed -> the loader may refill the stage
                    if (tap == 8) umma_commit(s_bar + 8 * (R3_AEMPTY + ra.i));               // group consumed -> its buffer may be restaged
                    if (last) umma_commit(s_bar + 8 * (R3_ACCF + rc.i));                     // this chain is complete
                }
                __syncwarp();
                rw.next(NW);
                if (tap == 8) ra.next(R3_NA);
                if (last) rc.next(R3_NROT);
            }
        }
    } else {
        // =============================== result warps (thread = raster row) ===============================
        const uint32_t t_lane = tmem + ((uint32_t)(32 * warp) << 16);
        const int n_tiles = (P * 100 + 127) / 128;
        R3Ring rc = {0, 0};
        uint32_t n_item = 0;
        for (int it = blockIdx.x; it < a.n_items; it += gridDim.x, ++n_item) {
            const R3Item w = r3_item(a, it);
            const int n_chains = (9 * w.ng + CHAIN - 1) / CHAIN;
            float acc[C];
#pragma unroll
            for (int i = 0; i < C; ++i) acc[i] = 0.0f;
            for (int c = 0; c < n_chains; ++c) {
                if (tid == 0) R3_STAMP(2, (int)n_item * 14 + c, 0);
                mbar_wait(s_bar + 8 * (R3_ACCF + rc.i), rc.ph);
                if (tid == 0) R3_STAMP(2, (int)n_item * 14 + c, 1);
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < (PREC ? 8 : 4); ++h) {                 // the hi half, then the lo half
                    uint32_t v[32];
                    if (BK_TC_DIAG & 8) continue;
                    tmem_ld32(t_lane + (uint32_t)(rc.i * 256 + h * 32), v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[(h & 3) * 32 + i] += __uint_as_float(v[i]);
                }
                tc_fence_before();
                if (tid == 0) R3_STAMP(2, (int)n_item * 14 + c, 2);
                mbar_arrive(s_bar + 8 * (R3_ACCE + rc.i));
                rc.next(R3_NROT);
            }
            if (tid == 0) R3_STAMP(2, (int)n_item * 14 + 13, 3);       // acc complete
            // ---- the rows go out through the warp's transposition buffer (store_rows_transposed): 5 k cycles per tile, under the first
            // two chains of the next one; stored straight from the row-owning threads they took 12 k, and the MMA issuer waited for them
            const int row = 32 * warp + lane;
            const int rr = w.tile * 128 + row;
            const int p = rr / 100, o = rr - 100 * p - 10;
            const int x = o / 10, y = o - 10 * x;
            // whole tile: the dense output; small batch: one dense copy per group (bk_train_sum4_kernel adds them); tail tile:
            // partial result [group][tail tile][row][128] (bk_train_conv3_tail_kernel adds them)
            const int at = !(p < P && o >= 0 && y < 9) ? -1
                           : w.kind == 2 ? (w.g_lo * (n_tiles - a.tail_full) + w.tile - a.tail_full) * 128 + row
                                         : (w.kind == 1 ? w.g_lo : 0) * a.M + p * NSQ + 9 * x + y;
            float *obase = w.kind == 2 ? a.tail_part : a.out;
            float *sst = reinterpret_cast<float *>(smem + Z::A_BYTES + NW * W_STAGE + 256) + warp * (32 * 33);
            const bool with_bias = a.bias && w.g_lo == 0;
#if BK_R3_STORE == 0
            if (at == -2) obase[0] = acc[0];       // measurement: no stores
#elif BK_R3_STORE == 1
            if (at >= 0) {                          // measurement: thread = row, 32 16-byte stores per thread
                float4 *out = reinterpret_cast<float4 *>(obase + (size_t)at * C);
#pragma unroll
                for (int i = 0; i < C / 4; ++i) {
                    float4 v = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
                    if (with_bias) {
                        const float4 bb = *reinterpret_cast<const float4 *>(a.bias + 4 * i);
                        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
                    }
                    out[i] = v;
                }
            }
#else
            store_rows_transposed(acc, sst, obase, at, with_bias ? a.bias : nullptr, lane);
#endif
            if (tid == 0) R3_STAMP(2, (int)n_item * 14 + 13, 2);       // stores issued
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// the split tail of bk_train_conv3_tc_kernel: out[real square of raster row row0 + r] = part[0][r] + part[1][r] + part[2][r] + part[3][r]
__global__ void bk_train_conv3_tail_kernel(const float4 *__restrict__ part, float *__restrict__ out, int row0, int n_rows, int P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = i >> 5, q = i & 31;
    if (r >= n_rows) return;
    const int rr = row0 + r;
    const int p = rr / 100, o = rr - 100 * p - 10;
    const int x = o / 10, y = o - 10 * x;
    if (p >= P || o < 0 || y >= 9) return;
    float4 v = part[(size_t)r * 32 + q];
#pragma unroll
    for (int g = 1; g < 4; ++g) {
        const float4 u = part[((size_t)g * n_rows + r) * 32 + q];
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    reinterpret_cast<float4 *>(out + (size_t)(p * NSQ + 9 * x + y) * C)[q] = v;
}

}   // namespace

#if BK_R3_PROF
extern "C" int bk_r3_prof_read(long long *host) { return (int)cudaMemcpyFromSymbol(host, g_r3_prof, sizeof(g_r3_prof)); }
#endif

int bk_tc_lo_bf16(void) { return BK_R3_LO_BF16; }

// SM count of the current device (per device, queried once)
static int bk_tc_sm_count()
{
    static std::mutex mu;
    static int cached[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    std::lock_guard<std::mutex> lock(mu);
    if (!cached[dev] && cudaDeviceGetAttribute(&cached[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) cached[dev] = 0;
    return cached[dev];
}

// The tiles left over after the last complete round of the grid run as four quarter-length work items each when those fit beside
// each other: the number of tiles that run whole (a multiple of the SM count), 0 = no split tail.  Pure arithmetic: bk_train_conv3_schedule
// exposes it to the tests.
static int conv3_tail_full(int P, int n_sm)
{
    const int n_tiles = (P * 100 + 127) / 128;
    const int full = n_sm > 0 ? n_tiles / n_sm * n_sm : 0, rest = n_tiles - full;
    return (full > 0 && rest > 0 && rest <= BK_CONV3_TAIL_MAX && 4 * rest <= n_sm) ? full : 0;
}

// ... for P positions on the current device
int bk_tc_conv3_tail(int P)
{
    const bool off = getenv("BK_TC_NO_TAIL") != nullptr;             // measurement / tests: read per call, so one process can compare both
    return off ? 0 : conv3_tail_full(P, bk_tc_sm_count());
}

// Work items of the persistent 3x3 training kernel for P positions on a device of n_sm SMs (ksplit = 4: the small-batch forward, one
// channel group per item): out[0] = tiles of 128 raster rows, out[1] = tiles that run whole when the tail is split (0 = no split),
// out[2] = work items, out[3] = CTAs launched.  No device needed.
extern "C" int bk_train_conv3_schedule(int P, int n_sm, int ksplit, int *out)
{
    if (P <= 0 || n_sm <= 0 || !out || (ksplit != 1 && ksplit != 4)) return -1;
    const int n_tiles = (P * 100 + 127) / 128;
    const int full = ksplit == 1 ? conv3_tail_full(P, n_sm) : 0;
    const int n_items = ksplit > 1 ? 4 * n_tiles : (full > 0 ? full + 4 * (n_tiles - full) : n_tiles);
    out[0] = n_tiles;
    out[1] = full;
    out[2] = n_items;
    out[3] = n_sm < n_items ? n_sm : n_items;
    return 0;
}

static bool getenv_old_conv()
{
    static const bool v = getenv("BK_TC_OLD_CONV") != nullptr;    // measurement: the window-gather kernel for the 3x3 layers as well
    return v;
}

int bk_tc_set_attrs(void)
{
    cudaError_t e = cudaSuccess;
#define BK_SET(k, n) if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, n)
    BK_SET((bk_train_gemm_tc_kernel<0, 0, BkConvArgs>), (Sizes<0, 0>::SMEM));
    BK_SET((bk_train_gemm_tc_kernel<0, 1, BkConvArgs>), (Sizes<0, 1>::SMEM));
    BK_SET((bk_train_gemm_tc_kernel<1, 0, BkWgradArgs>), (Sizes<1, 0>::SMEM));
    BK_SET((bk_train_gemm_tc_kernel<1, 1, BkWgradArgs>), (Sizes<1, 1>::SMEM));
    BK_SET((bk_train_conv3_tc_kernel<0>), (R3<0>::SMEM));
    BK_SET((bk_train_conv3_tc_kernel<1>), (R3<1>::SMEM));
#undef BK_SET
    return e == cudaSuccess ? 0 : -3;
}

void bk_tc_launch_conv(const BkConvArgs &a, int three_x, cudaStream_t st)
{
    if (a.R == 3 && a.Cin == C && !getenv_old_conv()) {          // the 3x3 layers and their data gradients: tile staged once
        const int n_tiles = (a.M / NSQ * 100 + 127) / 128;
        BkConvArgs b = a;
        b.tail_full = a.ksplit == 1 && a.tail_part ? bk_tc_conv3_tail(a.M / NSQ) : 0;
        const int rest = n_tiles - b.tail_full;
        // work items of the persistent CTAs (one per SM): whole tiles, or one channel group of a tile (small batches, split tail)
        b.n_items = a.ksplit > 1 ? 4 * n_tiles : (b.tail_full > 0 ? b.tail_full + 4 * rest : n_tiles);
        const int n_sm = bk_tc_sm_count();
        const int tiles = n_sm > 0 && n_sm < b.n_items ? n_sm : b.n_items;
        if (three_x) bk_train_conv3_tc_kernel<1><<<tiles, R3_THREADS, R3<1>::SMEM, st>>>(b);
        else bk_train_conv3_tc_kernel<0><<<tiles, R3_THREADS, R3<0>::SMEM, st>>>(b);
        if (b.tail_full > 0)
            bk_train_conv3_tail_kernel<<<(rest * 128 * 32 + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float4 *>(a.tail_part), a.out,
                                                                                       b.tail_full * 128, rest * 128, a.M / NSQ);
        return;
    }
    const int grid = (a.M + 127) / 128;
    if (three_x) bk_train_gemm_tc_kernel<0, 1, BkConvArgs><<<grid, N_THREADS, Sizes<0, 1>::SMEM, st>>>(a);
    else bk_train_gemm_tc_kernel<0, 0, BkConvArgs><<<grid, N_THREADS, Sizes<0, 0>::SMEM, st>>>(a);
}

void bk_tc_launch_wgrad(const BkWgradArgs &a, int splits, int three_x, cudaStream_t st)
{
    const dim3 grid((a.K + 127) / 128, splits);
    if (three_x) bk_train_gemm_tc_kernel<1, 1, BkWgradArgs><<<grid, N_THREADS, Sizes<1, 1>::SMEM, st>>>(a);
    else bk_train_gemm_tc_kernel<1, 0, BkWgradArgs><<<grid, N_THREADS, Sizes<1, 0>::SMEM, st>>>(a);
}
