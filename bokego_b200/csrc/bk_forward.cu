// bk_forward.cu -- kernel (b): PolicyNet / ValueNet forward for batches of 9x9 positions on sm_100a.
//
// Replaces PolicyNet.forward (/root/reference/bokego/nnet.py:19-57), ValueNet.forward (nnet.py:59-113),
// Conv2dUntiedBias.forward (nnet.py:175-180) and SOFT (nnet.py:16) in eval mode, BatchNorm folded.
//
// Persistent CTA PAIRS (clusters of two, one CTA per SM).  A work item of one CTA is one trunk (policy or value)
// over up to 5 boards of one group; the two CTAs of a pair work on two items of the SAME net in lock step and
// share the weights through tcgen05 cta_group::2:
//   * activations of the item (<= 5 x 100 padded rows x 128 ch, fp16) stay in shared memory for all 7 conv
//     layers; each layer is an implicit GEMM  D[rows, 128 co] += A[rows + tap shift, ci] * W[tap]  issued by the
//     pair's leader as tcgen05.mma.cta_group::2.kind::f16, M=256 (128 rows of each CTA), N=128, K=16, with fp32
//     accumulators for up to four 128-row tiles per CTA resident in TMEM (4 x 128 = 512 columns);
//     a single-CTA M=128 x N=128 MMA reads 8 KiB of operands per 64 tensor cycles, i.e. it is bound by the
//     128 B/clk shared-memory port (measured 76 cycles); in the pair each CTA reads its own A rows and only
//     HALF of B (the weights of 64 output channels), 6 KiB per MMA, and the MMA runs at its 64-cycle floor;
//   * weights stream L2 -> shared memory in stages of four K steps (8 KiB per CTA, tensor-map TMA + mbarrier,
//     a 6-slot ring); every stage is used by all tiles of the pass before its slot is recycled; the last three stages
//     of a pass (and the bias K step) are issued tile by tile with one completion barrier per tile, so the read-out of
//     tile t runs under the MMAs of the tiles behind it and only the last tile's read-out is exposed;
//   * the folded bias enters through the tensor core as well: two extra K rows (bias split into fp16 hi + lo)
//     multiplied by a constant all-ones operand, so the accumulators leave TMEM ready for ReLU;
//   * 16 epilogue warps per CTA (one thread per GEMM row: TMEM lane quarter = warp % 4, tile = warp / 4) pull the
//     accumulators out of TMEM (tcgen05.ld), apply ReLU, round to fp16 and write the next layer's operand in
//     place; after the last layer they compute the 1x1 head with the untied bias and finish with the 81-way
//     softmax (policy) or the small dense tail + tanh (value).
// Warp roles: warps 0-15 epilogue, warp 16 = bulk-copy producer (+ TMEM allocation), warp 17 = MMA issuer in the
// leader CTA / forwarder in the peer CTA (tells the leader when the peer's feature planes have landed).
// Items that do not fill a whole round of the grid are split into smaller board ranges (fewer M tiles each)
// so that the last round, and small batches, spread over more SMs.
//
// PLAYOUT instantiation (bk_playout_run): the same kernel keeps its <= 5 boards for a WHOLE playout.  After the policy head of
// move k the probabilities stay in shared memory; three epilogue warps per board sample the move, play it, refresh the liberty
// cache and encode the new position (bk_step_core.cuh, the code of the stepping kernel) straight into the shared-memory
// feature operand of move k + 1 -- no launch, no grid-wide dependency and no HBM round trip between the moves of a game
// (/root/reference/bokego/mcts.py:195-206, bin/selfplay.py:18-33).  Boards, ko / last / turn, the liberty cache and the
// move record stay in global memory (L2); the two policy nets of a self-play game alternate by move parity.
//
// A plain CUDA-core kernel over the same packed operands (BK_FWD_SIMT) exists to validate the packing and
// the tensor-core path against each other on the GPU; it is not a fallback and is never selected implicitly.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>

#include "bk_layout.h"
#include "bk_step_core.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------
// shared-memory plan of the tcgen05 kernel (identical in both CTAs of a pair)
// ------------------------------------------------------------------------------------------------------
constexpr int A_MARGIN = 12;                       // zero rows in front of GEMM row 0 (|tap shift| <= 11)
constexpr int A_ROWS = A_MARGIN + 512;             // 524: the rows behind a chunk are the next chunk's margin
constexpr int A_LBO = A_ROWS * 16;                 // bytes between consecutive 8-channel chunks
constexpr int A_BYTES = 16 * A_LBO;                // 134,144
constexpr int F_MARGIN = 24;                       // |5x5 tap shift| <= 24
constexpr int F_ROWS = F_MARGIN + BK_F_ROWS_G;     // 629
constexpr int F_LBO = F_ROWS * 16;
constexpr int F_BYTES = (BK_F_CHUNKS * F_LBO + F_MARGIN * 16 + 127) / 128 * 128;   // 40,704 (+ zero rows behind chunk 3)
constexpr int CTA_STAGE_BYTES = BK_STAGE_BYTES / 2;   // this CTA's N half of a stage: [4 k-steps][2 k-chunks][64 co][8 k]
constexpr int CTA_KSTEP_BYTES = BK_KSTEP_BYTES / 2;
constexpr int CTA_BIAS_BYTES = BK_BIAS_BYTES / 2;  // this CTA's half of a layer's bias rows
constexpr int N_STAGES = 6;                        // 48 KiB weight ring
#ifndef BK_HANDOVER
#define BK_HANDOVER 0                              // 1, 2 = measurement builds (see the epilogue)
#endif
#ifndef BK_HEAD_PER_TILE
#define BK_HEAD_PER_TILE 1                         // measurement builds: 0 = every pass waits for the whole hand-over of the previous one, 2 = per tile between 3x3 layers only
#endif
#ifndef BK_HEAD
#define BK_HEAD 2                                  // stages at the start of a pass that are issued tile by tile
#endif
#ifndef BK_TAIL
#define BK_TAIL 3                                  // stages at the end of a pass that are issued tile by tile
#endif
#ifndef BK_TAIL_PLAYOUT
#define BK_TAIL_PLAYOUT 2                          // the same in the persistent playout kernel (measured: 2 is 0.6 % faster per move than 3)
#endif
constexpr int ONES_BYTES = 4096;                   // [2 k-chunks][128 rows][8]: 1.0 in k = 0, 1
constexpr int OFF_A = 0;
constexpr int OFF_F = OFF_A + A_BYTES;             // F's front margin doubles as the rows behind A's last chunk
constexpr int OFF_W = OFF_F + F_BYTES;
constexpr int OFF_BIASW = OFF_W + N_STAGES * CTA_STAGE_BYTES;
constexpr int OFF_ONES = OFF_BIASW + CTA_BIAS_BYTES;
constexpr int OFF_BAR = OFF_ONES + ONES_BYTES;     // barriers, 8 B each
constexpr int BAR_BYTES = 208;
constexpr int OFF_TMEM = OFF_BAR + BAR_BYTES;
constexpr int OFF_LOGIT = OFF_TMEM + 16;           // float[5][81]: head output per square
constexpr int OFF_RES = OFF_LOGIT + 1664;          // PLAYOUT: resident positions of the item's boards (BkResident[5])
constexpr int SMEM_BYTES = OFF_RES + 240 + 128;    // 232,384 (ENCODE: logit area + resident area + tail = one group table, see below)
static_assert(SMEM_BYTES <= 232448, "shared memory plan exceeds 227 KiB");
// PLAYOUT: between the last layer of move k and layer 0 of move k + 1 the activation buffer is dead; the step phase keeps its
// per-board scratch (group table, arg-max slots) at its start and zeroes it again (padding rows must read as zero)
constexpr int STEP_SCRATCH = (int)((sizeof(BkStepScratch) + 15) / 16 * 16);
static_assert(BK_GROUP * STEP_SCRATCH <= A_BYTES, "step scratch");
constexpr int HALF_HEAD_OFF = 16384;               // PLAYOUT: 64 partial head sums of a half tile (dead activation buffer, re-zeroed)
static_assert(BK_GROUP * STEP_SCRATCH <= HALF_HEAD_OFF && HALF_HEAD_OFF + 256 <= A_BYTES, "half-tile head scratch");
static_assert(BK_GROUP * sizeof(BkResident) <= 240, "resident positions");
// ENCODE: while the tensor pipe runs layers 1..5 of an item, three epilogue warps encode the boards of the NEXT item one at a time;
// their group table lives in the logit area (written only by the last layer's epilogue) and the bytes behind it
static_assert(sizeof(BKGroups) <= SMEM_BYTES - OFF_LOGIT, "encode scratch");
static_assert(OFF_W % 128 == 0 && OFF_ONES % 128 == 0 && OFF_BIASW % 128 == 0, "operand alignment");

// WFULL (leader): both CTAs' halves of a stage have landed -- each CTA's tensor-map copy (cta_group::2) reports its
// bytes to the LEADER's barrier.  WEMPTY: the pair's MMAs that read the stage are complete.  BFULL / BEMPTY: the same for
// the layer's bias rows (single slot).  ACC[t]: accumulator tile t of the pass complete (the last stages of a pass are issued
// tile by tile, so the read-out of the first tiles runs under the MMAs of the last).  ACT[t] (leader): the warp groups of tile t
// of both CTAs have read their accumulators out and rewritten their rows.
// FFULL / PFFULL (leader) / FEMPTY: the same for the feature planes of an item (the peer forwards its FFULL to the
// leader's PFFULL, once per item).
enum { BAR_WFULL = 0, BAR_WEMPTY = N_STAGES, BAR_BFULL = 2 * N_STAGES, BAR_BEMPTY, BAR_ACC, BAR_ACT = BAR_ACC + 4,
       BAR_FFULL = BAR_ACT + 4, BAR_PFFULL, BAR_FEMPTY, N_BARS };
static_assert(N_BARS * 8 <= BAR_BYTES, "barrier area");

// the conv weights of a blob seen as a 2-D tensor of 512-byte rows (256 fp16): one CTA's half of a stage = 16 rows,
// its half of the bias rows = 4 rows (a second tensor map with a smaller box)
constexpr int TM_ROW_BYTES = 512;
constexpr int TM_BOX_ROWS = CTA_STAGE_BYTES / TM_ROW_BYTES;     // 16
constexpr int TM_BIAS_BOX_ROWS = CTA_BIAS_BYTES / TM_ROW_BYTES; // 4
constexpr int TM_ROWS = BK_W_BIAS_OFF / TM_ROW_BYTES;           // 3928
static_assert(BK_W_BIAS_OFF % TM_ROW_BYTES == 0 && BK_L0_BYTES % TM_ROW_BYTES == 0 && BK_L_BYTES % TM_ROW_BYTES == 0, "row grid");

constexpr int N_EPI_WARPS = 16;
constexpr int WARP_PRODUCER = 16;
constexpr int WARP_MMA = 17;
constexpr int N_THREADS = 576;

// instruction descriptor: D=f32 (bit 4), A=B=f16 (0), both K-major (0), N=128 (>>3 at bit 17), M=256 (>>4 at bit 24)
constexpr uint32_t IDESC = (1u << 4) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
// high word of every shared-memory matrix descriptor used here: stride between 8-row groups = 128 B, version 1
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);

// ------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on a barrier of another (or this) CTA of the cluster; `cbar` is a shared::cluster address from mapa()
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cbar)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cbar) : "memory");
}
// The epilogue's hand-over to the leader.  What the leader's MMAs read next is this CTA's OWN shared memory, through this
// SM's tensor core; the stores were made visible to that (async) proxy by fence.proxy.async before this arrive, so only the
// signal crosses the SM boundary and a cluster-scope release (which costs ~1,000 cycles here) is not needed.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cbar)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must end in a trap (launch error), never in a hung GPU.
__device__ unsigned int *g_dbg = nullptr;
#ifdef BK_TRACE   // measurement build (tools/trace_handover.py): one hash per epilogue thread and pass of what it read out of TMEM
__device__ unsigned int *g_trace = nullptr;
constexpr int TRACE_PASSES = 24;
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t tag = 0)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            if (g_dbg) {
                g_dbg[0] = 0xDEAD0000u | (threadIdx.x & 0xFFFFu); g_dbg[1] = blockIdx.x; g_dbg[2] = bar; g_dbg[3] = parity;
                g_dbg[4] = tag;
                __threadfence_system();
            }
            asm volatile("trap;");
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// 8 rows x 512 B of the weight tensor -> this CTA's shared memory; the byte count is reported to `cbar`, a barrier
// that may live in the other CTA of the pair (shared::cluster address)
__device__ __forceinline__ void tma_rows_g2s(uint32_t dst, const CUtensorMap *tm, int row, uint32_t cbar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(0), "r"(row), "r"(cbar)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMEM of the pair: the same warp of both CTAs allocates / frees, both get the same column range
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem] over the CTA pair: rows 0..127 from the leader's A, 128..255 from the peer's A at the
// same shared-memory offset; output channels 0..63 from the leader's B half, 64..127 from the peer's
// instruction descriptor of the HALF tile: M = 128 over the pair = 64 rows of each CTA, at half the tensor time (32 cycles).
// Its accumulator layout (tools/probes/tc_probe_m128x2.cu, profiles/r02_probe_m128x2.txt): row r (0..63) x channel n sits at TMEM
// lane r + 64 * (n / 64), column n % 64 -- 64 columns used; the other 64 columns of the 128-column slot are clobbered.
constexpr uint32_t IDESC_HALF = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t accum, uint32_t idesc = IDESC);
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t accum, uint32_t idesc)
{
    const uint64_t adesc = ((uint64_t)DESC_HI << 32) | a_lo, bdesc = ((uint64_t)DESC_HI << 32) | b_lo;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// completion of all MMAs issued so far -> one arrival on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
// 32 lanes x 32 consecutive columns of 32-bit accumulators -> 32 registers per thread
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
// one elected lane of a converged warp (the compiler then knows a single thread issues what follows)
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
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Low word of a K-major, no-swizzle shared-memory matrix descriptor: start address and the byte distance
// between the two 8-element K chunks of one MMA.  Rows are 16 B apart, so adding n to the word moves the
// operand window n rows down (all shared-memory addresses are below 2^18).
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo)
{
    return ((saddr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16);
}

// ------------------------------------------------------------------------------------------------------
// work items, passes, row bookkeeping
// ------------------------------------------------------------------------------------------------------
struct FwdArgs {
    const uint8_t *feats;      // [G][BK_F_GROUP_BYTES]
    const uint8_t *blob[2];    // policy, value
    float *logits, *probs, *value;
    int B, G, n_nets, first_net;
    // schedule: per net the sub-items are [whole groups 0..g_whole-1] then [groups g_whole..G-1 split into `split`
    // board ranges each]; a PAIR takes sub-items 2j (leader) and 2j+1 (peer) of one net; pair v = j * n_nets + net
    int g_whole, split, n_sub, n_pairs;
    float *dump;               // diagnostic: raw accumulators [640][128] of pass `dump_pass` (first item of CTA 0)
    int dump_pass;
    int diag;                  // diagnostic bits: 1 = producer signals stages without copying, 2 = aligned A windows
    long long *prof;           // diagnostic: clock64 stamps of CTA 0, 4 per pass
    unsigned int *dbg;         // host-mapped words written before a bounded wait traps
    // PLAYOUT only: whole games inside the kernel.  blob[0] / blob[1] are then the policy nets of the moves made at even / odd
    // (first_turn + k); moves_out is [n_steps][B]
    int8_t *boards; int16_t *ko, *last, *turn; uint8_t *libs, *done; int16_t *moves_out;
    int n_steps, mode, max_turn, first_turn;
    // ENCODE only: positions in (boards / ko / last / turn above, libs = carried liberty cache or null for fresh Games), optional
    // outputs of the encoder
    uint8_t *legal_out, *libs_out;
    int play_group;            // boards per item (1..5): the smallest that puts every board on an SM in one round
    int fresh_libs;            // the starting positions have no liberty cache (exact liberties, like fresh Games)
    unsigned long long seed;
    uint32_t game0;
};

struct Item { int g, net, lo, nb; };

// sub-item u of a net -> (group, first board, board count); nb = 0 when out of range or the range is empty
__device__ __forceinline__ void decode_sub(const FwdArgs &a, int u, Item &it)
{
    it.g = 0; it.lo = 0; it.nb = 0;
    if (u >= a.n_sub) return;
    int lo = 0, hi = BK_GROUP;
    if (u < a.g_whole) {
        it.g = u;
    } else {
        const int t = u - a.g_whole;
        it.g = a.g_whole + t / a.split;
        const int piece = t - (t / a.split) * a.split;
        lo = piece * BK_GROUP / a.split;
        hi = (piece + 1) * BK_GROUP / a.split;
    }
    const int in_group = min(BK_GROUP, a.B - BK_GROUP * it.g);
    it.lo = min(lo, in_group);
    it.nb = min(hi, in_group) - it.lo;
}
// PLAYOUT: sub-item u = boards [u * play_group, (u + 1) * play_group) of the batch (it.g = first board, it.lo = 0)
__device__ __forceinline__ void decode_play(const FwdArgs &a, int u, Item &it)
{
    it.net = 0; it.lo = 0;
    it.g = u * a.play_group;
    it.nb = u < a.n_sub ? min(a.play_group, a.B - it.g) : 0;
    if (it.nb < 0) it.nb = 0;
}
template <bool PLAYOUT>
__device__ __forceinline__ bool decode_pair_t(const FwdArgs &a, int v, int rank, Item &mine, int &pair_nb);
// pair v -> this CTA's item and the board count that fixes the pass structure of the pair; false = nothing to do
__device__ __forceinline__ bool decode_pair(const FwdArgs &a, int v, int rank, Item &mine, int &pair_nb)
{
    const int j = v / a.n_nets;
    Item other;
    decode_sub(a, 2 * j + rank, mine);
    decode_sub(a, 2 * j + (rank ^ 1), other);
    mine.net = a.first_net + (v - j * a.n_nets);
    pair_nb = max(mine.nb, other.nb);
    return pair_nb > 0;
}

template <bool PLAYOUT>
__device__ __forceinline__ bool decode_pair_t(const FwdArgs &a, int v, int rank, Item &mine, int &pair_nb)
{
    if (!PLAYOUT) return decode_pair(a, v, rank, mine, pair_nb);
    Item other;
    decode_play(a, 2 * v + rank, mine);
    decode_play(a, 2 * v + (rank ^ 1), other);
    pair_nb = max(mine.nb, other.nb);
    return pair_nb > 0;
}

// A full item takes 8 passes (layer 0 needs five 128-row tiles but TMEM holds four), a smaller one 7.
struct Pass { int layer, tile0, n_tiles; bool l0_last; };
__device__ __forceinline__ int n_passes(int nb) { return nb == BK_GROUP ? 8 : 7; }
__device__ __forceinline__ Pass pass_info(int nb, int ps)
{
    Pass p;
    const int nt = (100 * nb - 1 + 127) >> 7;           // tiles of the stride-10 raster (layers 1..6)
    if (nb == BK_GROUP) {                               // layer 0: tiles 0..2, then tiles 3..4 (3 + 2 balances the two passes)
        p.layer = ps < 2 ? 0 : ps - 1;
        p.tile0 = ps == 1 ? 3 : 0;
        p.n_tiles = ps == 0 ? 3 : (ps == 1 ? 2 : nt);
        p.l0_last = ps == 1;
    } else {
        p.layer = ps;
        p.tile0 = 0;
        p.n_tiles = ps == 0 ? (121 * nb - 2 + 127) >> 7 : nt;   // tiles of the stride-11 raster (layer 0)
        p.l0_last = ps == 0;
    }
    return p;
}
// PLAYOUT: index of the pass's last tile when it may run as a HALF tile (64 rows per CTA, M = 128 MMAs at half the tensor
// time), else -1.  Layers 1..6 of a 4-board item end at GEMM row 398 (tile 3 holds 15 real rows), of a 3-board item at row 298
// (tile 2: 43 rows); 1, 2 and 5 boards fill more than half of their last tile.  Only the persistent playout kernel uses it:
// there an item lasts as long as its passes, and the half tile takes 1/8 off the layers of a 4-board item.
template <bool PLAYOUT>
__device__ __forceinline__ int half_last_tile(int pair_nb, const Pass &pi)
{
    return (PLAYOUT && pi.layer >= 1 && (pair_nb == 3 || pair_nb == 4)) ? pi.n_tiles - 1 : -1;
}
__device__ __forceinline__ int n_stages_of(int layer) { return layer == 0 ? BK_L0_STAGES : BK_L_STAGES; }

// layers 1..6: GEMM row r (0..511) -> is it a real square of boards 0..nb-1?  (in place: destination row = r)
__device__ __forceinline__ bool act_row_valid(int r, int nb, int &board, int &sq)
{
    board = r / 100;
    const int rem = r - 100 * board;
    const int q = rem - 10;
    const int x = q / 10, y = q - 10 * x;
    sq = 9 * x + y;
    return board < nb && rem >= 10 && y < 9;
}
// layer 0: GEMM row r0 (0..639, stride-11 raster) -> destination activation row, or -1
__device__ __forceinline__ int l0_dest_row(int r0, int nb)
{
    const int board = r0 / BK_F_ROWS_B;
    const int rem = r0 - BK_F_ROWS_B * board;
    if (board >= nb || rem < 22) return -1;
    const int q = rem - 22;
    const int x = q / 11, y = q - 11 * x;
    if (y >= 9) return -1;
    return 100 * board + 10 + 10 * x + y;
}

__device__ __forceinline__ uint32_t relu_pack(uint32_t a, uint32_t b)
{
    const __half2 h = __hmax2(__floats2half2_rn(__uint_as_float(a), __uint_as_float(b)), __float2half2_rn(0.0f));
    return *reinterpret_cast<const uint32_t *>(&h);
}
// ReLU + fp16 rounding of 32 consecutive output channels of one row -> 4 chunks of the activation operand
__device__ __forceinline__ void store_act32(uint8_t *smem, const uint32_t (&v)[32], int chunk0, int dest)
{
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        uint4 o;
        o.x = relu_pack(v[c8 * 8 + 0], v[c8 * 8 + 1]);
        o.y = relu_pack(v[c8 * 8 + 2], v[c8 * 8 + 3]);
        o.z = relu_pack(v[c8 * 8 + 4], v[c8 * 8 + 5]);
        o.w = relu_pack(v[c8 * 8 + 6], v[c8 * 8 + 7]);
        *reinterpret_cast<uint4 *>(smem + OFF_A + (chunk0 + c8) * A_LBO + (A_MARGIN + dest) * 16) = o;
    }
}
// 1x1 head conv: sum over 32 channels of relu(acc) * w, continuing the chain `acc` (a row's head output is ONE chain of 128 fused
// multiply-adds in channel order in every instantiation; four interleaved chains were measured 1.5 % slower per playout move)
__device__ __forceinline__ float head_dot32(const uint32_t (&v)[32], const float4 *hw4, float acc)
{
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
        const float4 w = __ldg(hw4 + c4);
        acc = fmaf(fmaxf(__uint_as_float(v[c4 * 4 + 0]), 0.0f), w.x, acc);
        acc = fmaf(fmaxf(__uint_as_float(v[c4 * 4 + 1]), 0.0f), w.y, acc);
        acc = fmaf(fmaxf(__uint_as_float(v[c4 * 4 + 2]), 0.0f), w.z, acc);
        acc = fmaf(fmaxf(__uint_as_float(v[c4 * 4 + 3]), 0.0f), w.w, acc);
    }
    return acc;
}

// Final stage shared by both kernels: `logit` holds the 81 head outputs (1x1 conv + untied bias) of one
// board; one warp turns them into probs (policy) or the scalar value (value net).
__device__ __forceinline__ void finish_board(const float *logit, int net, const uint8_t *blob, float *logits_out,
                                             float *probs_out, float *value_out, int lane)
{
    if (net == 0) {
        float v[3], mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int p = lane + 32 * k;
            v[k] = p < 81 ? logit[p] : -INFINITY;
            mx = fmaxf(mx, v[k]);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float e[3], sum = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            e[k] = (lane + 32 * k) < 81 ? expf(v[k] - mx) : 0.0f;
            sum += e[k];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int p = lane + 32 * k;
            if (p < 81) {
                if (logits_out) logits_out[p] = v[k];
                if (probs_out) probs_out[p] = e[k] / sum;
            }
        }
    } else {
        const float *vt = reinterpret_cast<const float *>(blob + BK_W_VT_OFF);
        const float *w1t = reinterpret_cast<const float *>(blob + BK_W_VT_W1T_OFF);
        const float *b1 = reinterpret_cast<const float *>(blob + BK_W_VT_B1_OFF);
        const float *w2 = reinterpret_cast<const float *>(blob + BK_W_VT_W2_OFF);
        const float s = __ldg(vt + 0), t = __ldg(vt + 1), b2 = __ldg(vt + 2);
        float h0 = __ldg(b1 + lane), h1 = __ldg(b1 + lane + 32);
        for (int p = 0; p < 81; ++p) {
            const float a = fmaxf(fmaf(logit[p], s, t), 0.0f);
            h0 = fmaf(__ldg(w1t + p * 64 + lane), a, h0);
            h1 = fmaf(__ldg(w1t + p * 64 + lane + 32), a, h1);
        }
        float acc = fmaxf(h0, 0.0f) * __ldg(w2 + lane) + fmaxf(h1, 0.0f) * __ldg(w2 + lane + 32);
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0 && value_out) *value_out = tanhf(acc + b2);
    }
}

// ------------------------------------------------------------------------------------------------------
// the tcgen05 kernel
// ------------------------------------------------------------------------------------------------------
// MODE 0: planes from global memory (bk_encode's operand, bulk copies); 1: PLAYOUT (whole games, positions resident on chip);
// 2: ENCODE (planes from the positions: the epilogue warps run nnet.features for the NEXT item while the tensor pipe works on
// the current one -- one launch from positions to probabilities and values)
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(N_THREADS, 1)
bk_forward_tc_kernel(const __grid_constant__ FwdArgs args, const __grid_constant__ CUtensorMap tm_policy,
                     const __grid_constant__ CUtensorMap tm_value, const __grid_constant__ CUtensorMap tb_policy,
                     const __grid_constant__ CUtensorMap tb_value)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp-uniform role index
    const uint32_t s_base = smem_u32(smem);
    const uint32_t sA = s_base + OFF_A, sF = s_base + OFF_F, sW = s_base + OFF_W, sBar = s_base + OFF_BAR;
    float *logit = reinterpret_cast<float *>(smem + OFF_LOGIT);
    const int rank = (int)cluster_ctarank();           // 0 = leader (issues the pair's MMAs), 1 = peer
    const int pair0 = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    constexpr bool PLAYOUT = MODE == 1, ENCODE = MODE == 2;
    const int n_steps = PLAYOUT ? args.n_steps : 1;     // moves every item is kept for (1 = plain forward)
    if (threadIdx.x == 0 && args.dbg) g_dbg = args.dbg;

    // ---- one-time setup: zero the operand buffers (pad rows must read as 0), the ones operand, barriers, TMEM
    {
        uint4 *z = reinterpret_cast<uint4 *>(smem);
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < (A_BYTES + F_BYTES) / 16; i += N_THREADS) z[i] = zero;
        uint4 *o = reinterpret_cast<uint4 *>(smem + OFF_ONES);
        for (int i = threadIdx.x; i < ONES_BYTES / 16; i += N_THREADS)
            o[i] = i < 128 ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : zero;      // k = 0, 1 -> 1.0 (fp16)
        fence_proxy_async();
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < N_STAGES; ++s) {
            mbar_init(sBar + 8 * (BAR_WFULL + s), 1);
            mbar_init(sBar + 8 * (BAR_WEMPTY + s), 1);
        }
        mbar_init(sBar + 8 * BAR_BFULL, 1);
        mbar_init(sBar + 8 * BAR_BEMPTY, 1);
        for (int t = 0; t < 4; ++t) mbar_init(sBar + 8 * (BAR_ACC + t), 1);
        for (int t = 0; t < 4; ++t) mbar_init(sBar + 8 * (BAR_ACT + t), 2 * (N_EPI_WARPS / 4));
        mbar_init(sBar + 8 * BAR_FFULL, 1);
        mbar_init(sBar + 8 * BAR_PFFULL, 1);
        mbar_init(sBar + 8 * BAR_FEMPTY, 1);
        fence_barrier_init();
    }
    if (warp == WARP_PRODUCER) tmem_alloc(s_base + OFF_TMEM, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                 // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t *>(smem + OFF_TMEM), 0);

    if (warp == WARP_PRODUCER) {
        // =========================== bulk-copy producer (both CTAs: own planes, own N half of the weights) =========
        if (lane == 0) {
            uint32_t wit = 0;   // weight stage counter over the whole kernel
            auto load_feats = [&](const Item &it) {
                // rows of boards lo..lo+nb-1; a partial range is followed by 22 zero rows (board 0's top padding)
                const uint8_t *src = args.feats + (size_t)it.g * BK_F_GROUP_BYTES;
                const uint32_t n = (uint32_t)(BK_F_ROWS_B * it.nb) * 16u;
                const uint32_t tail = (it.nb > 0 && it.nb < BK_GROUP) ? 22u * 16u : 0u;
                mbar_arrive_expect_tx(sBar + 8 * BAR_FFULL, BK_F_CHUNKS * (n + tail));
                if (n == 0) return;                      // the partner has boards, this CTA does not
#pragma unroll
                for (int c = 0; c < BK_F_CHUNKS; ++c) {
                    const uint32_t dst = sF + c * F_LBO + F_MARGIN * 16;
                    bulk_g2s(dst, src + ((size_t)c * BK_F_ROWS_G + BK_F_ROWS_B * it.lo) * 16, n, sBar + 8 * BAR_FFULL);
                    if (tail) bulk_g2s(dst + n, src + (size_t)c * BK_F_ROWS_G * 16, tail, sBar + 8 * BAR_FFULL);
                }
            };
            const uint32_t leader_wfull = mapa(sBar + 8 * BAR_WFULL, 0), leader_bfull = mapa(sBar + 8 * BAR_BFULL, 0);
            uint32_t n_bias = 0;    // bias loads so far (one per pass)
            // this CTA's half of every stage of one pass over a layer, and of the layer's bias rows.  The bias slot is
            // refilled a few stages into the pass: by then the previous pass has handed it back, so the wait is free and
            // the first stages of the pass are prefetched while the previous pass is still running.
            auto stream = [&](const CUtensorMap *tm, const CUtensorMap *tb, int layer_off, int n_stages) {
                for (int s = 0; s < n_stages; ++s, ++wit) {
                    const uint32_t st = wit % N_STAGES, ph = (wit / N_STAGES) & 1u;
                    const int row = layer_off / TM_ROW_BYTES + s * (BK_STAGE_BYTES / TM_ROW_BYTES) + rank * TM_BOX_ROWS;
                    mbar_wait(sBar + 8 * (BAR_WEMPTY + st), ph ^ 1u, 0x100u + wit);
                    if ((args.diag & 1) && wit >= N_STAGES) {          // measurement only: skip the copies (wrong results)
                        if (rank == 0) mbar_arrive(sBar + 8 * (BAR_WFULL + st));
                    } else {
                        if (rank == 0) mbar_arrive_expect_tx(sBar + 8 * (BAR_WFULL + st), 2 * CTA_STAGE_BYTES);   // both halves
                        tma_rows_g2s(sW + st * CTA_STAGE_BYTES, tm, row, leader_wfull + 8 * st);
                    }
                    if (s == N_STAGES - 1) {
                        mbar_wait(sBar + 8 * BAR_BEMPTY, (n_bias & 1u) ^ 1u, 0x180u);
                        if (rank == 0) mbar_arrive_expect_tx(sBar + 8 * BAR_BFULL, 2 * CTA_BIAS_BYTES);
                        tma_rows_g2s(s_base + OFF_BIASW, tb, (layer_off + n_stages * BK_STAGE_BYTES) / TM_ROW_BYTES + rank * TM_BIAS_BOX_ROWS,
                                     leader_bfull);
                        ++n_bias;
                    }
                }
            };
            Item it, nx;
            int pair_nb, nx_nb;
            uint32_t n_femp = 0;    // FEMPTY phases consumed (one per item and move)
            if (pair0 < args.n_pairs && decode_pair_t<PLAYOUT>(args, pair0, rank, it, pair_nb) && MODE == 0) load_feats(it);
            for (int v = pair0; v < args.n_pairs; v += n_clusters) {
                if (!decode_pair_t<PLAYOUT>(args, v, rank, it, pair_nb)) continue;
                for (int k = 0; k < n_steps; ++k) {
                    const int net = PLAYOUT ? ((args.first_turn + k) & 1) : it.net;
                    const CUtensorMap *tm = net == 0 ? &tm_policy : &tm_value;
                    const CUtensorMap *tb = net == 0 ? &tb_policy : &tb_value;
                    stream(tm, tb, BK_W_L0_OFF, n_stages_of(0));
                    if (pair_nb == BK_GROUP) stream(tm, tb, BK_W_L0_OFF, n_stages_of(0));   // layer 0, second pass (tiles 3, 4)
                    stream(tm, tb, BK_W_L_OFF(1), n_stages_of(1));
                    // layer 0 of this move has read the feature planes; the planes of the NEXT item may be fetched once the
                    // item's last move is past that point (in between, the epilogue warps write the next move's planes)
                    mbar_wait(sBar + 8 * BAR_FEMPTY, n_femp & 1u, 0x200u);
                    ++n_femp;
                    // (PLAYOUT: the planes never come from global memory -- the epilogue warps encode every position in place)
                    if (MODE == 0 && v + n_clusters < args.n_pairs && decode_pair(args, v + n_clusters, rank, nx, nx_nb)) load_feats(nx);
                    for (int l = 2; l <= 6; ++l) stream(tm, tb, BK_W_L_OFF(l), n_stages_of(l));
                }
            }
        }
    } else if (warp == WARP_MMA && rank == 1) {
        // =========================== peer: tell the leader when this CTA's feature planes have landed ===============
        // (PLAYOUT: the planes are written by the epilogue warps, which report to the leader themselves -- n_pairs loop skipped)
        uint32_t n_done = 0;
        const uint32_t leader_bar = mapa(sBar, 0);
        Item it;
        int pair_nb;
        for (int v = pair0; v < args.n_pairs && !PLAYOUT; v += n_clusters) {
            if (!decode_pair_t<PLAYOUT>(args, v, rank, it, pair_nb)) continue;
            for (int k = 0; k < n_steps; ++k) {
                mbar_wait(sBar + 8 * BAR_FFULL, n_done & 1u, 0x300u);
                ++n_done;
                if (lane == 0) mbar_arrive_cluster(leader_bar + 8 * BAR_PFFULL);
            }
        }
    } else if (warp == WARP_MMA) {
        // =========================== leader: MMA issuer (whole warp runs the loop, one elected lane issues) ===========
        // The tensor pipe queues only a few MMAs, so the scalar work between two stages has to stay small:
        // descriptors are formed with adds of compile-time offsets, and one elected lane issues all MMAs of a stage
        // plus the commit that hands the stage back to both producers.
        uint32_t st = 0, ph = 0, pass = 0, n_done = 0;   // weight ring slot / phase
        const uint32_t a_lo0 = desc_lo(sA + A_MARGIN * 16, A_LBO);
        const uint32_t f_lo0 = desc_lo(sF + F_MARGIN * 16, F_LBO);
        const uint32_t one_lo = desc_lo(s_base + OFF_ONES, 2048);
        const uint32_t w_lo0 = desc_lo(sW, 1024);          // B half: [2 k-chunks][64 co][8 k] per K step
        const uint32_t b_lo0 = desc_lo(s_base + OFF_BIASW, 1024);
        int n_tiles = 0;
        int half_tile = -1;        // PLAYOUT, layers 1..6 of 3- / 4-board items: the last tile holds <= 64 real rows -> M = 128 MMAs
        long long tw = 0, tp = 0, ti = 0;                  // diagnostic: cycles waiting for weights / bias rows, issuing
        const bool profiling = !PLAYOUT && args.prof != nullptr && blockIdx.x == 0;
        constexpr int TAIL = PLAYOUT ? BK_TAIL_PLAYOUT : BK_TAIL;   // stages at the end of a pass that are issued tile by tile
        constexpr int HEAD = BK_HEAD;                      // stages at the start of a pass that are issued tile by tile (tap 0)
        static_assert(HEAD >= 1 && HEAD <= 8, "the head may only hold taps with negative row shifts (taps 0..3)");
        static_assert(HEAD + TAIL < N_STAGES, "head and tail slots are held across all tiles: leave ring slots to prefetch into");
        static_assert(TAIL >= 1 && TAIL <= 5, "the tail may only hold taps with non-negative row shifts, and must leave ring slots free");
        // A windows of stage s of a pass (descriptor words of the pass's first tile): K steps 0, 1 at w.x, w.x + w.z and K steps
        // 2, 3 at w.y, w.y + w.z; nk = how many of the four exist
        auto windows = [&](const Pass &pi, int s, int &nk) -> uint3 {
            uint3 w;
            if (pi.layer == 0) {       // stage s = taps 2s, 2s+1; a tap is two K steps (channel chunks 0,1 / 2,3); the last stage holds tap 24 only
                const uint32_t fb = f_lo0 + (uint32_t)(128 * pi.tile0);
                const int t0 = 2 * s, t1 = 2 * s + 1;
                w.x = fb + (uint32_t)((t0 / 5 - 2) * 11 + (t0 % 5 - 2));
                w.y = fb + (uint32_t)((t1 / 5 - 2) * 11 + (t1 % 5 - 2));
                w.z = 2u * (F_LBO >> 4);
                nk = t1 < 25 ? 4 : 2;
            } else {                   // stage s = half a tap: K steps with channel chunks (8*part + 2j, +1), j = 0..3
                const int tap = s >> 1, part = s & 1;
                const int ti = tap / 3, tj = tap - 3 * ti;
                w.x = a_lo0 + (uint32_t)((ti - 1) * 10 + (tj - 1)) + (uint32_t)(8 * part) * (A_LBO >> 4);
                if (args.diag & 2) w.x = a_lo0 - 4u + (uint32_t)(8 * part) * (A_LBO >> 4);   // measurement only: 128 B aligned windows
                w.y = w.x + 4u * (A_LBO >> 4);
                w.z = 2u * (A_LBO >> 4);
                nk = 4;
            }
            return w;
        };
        // the K steps of one stage for tiles [t_lo, t_hi): B = the 2 KiB K steps of ring slot `slot` in both CTAs
        auto issue = [&](const uint3 w, int nk, uint32_t slot, int t_lo, int t_hi, uint32_t accum0) {
            const uint32_t w_lo = w_lo0 + slot * (CTA_STAGE_BYTES >> 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nk) {
                    const uint32_t aj = (j < 2 ? w.x : w.y) + (uint32_t)(j & 1) * w.z;
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (t >= t_lo && t < t_hi)
                            umma_f16(tmem + (uint32_t)(t * 128), aj + (uint32_t)t * 128u, w_lo + (uint32_t)j * (CTA_KSTEP_BYTES >> 4),
                                     j == 0 ? accum0 : 1u, t == half_tile ? IDESC_HALF : IDESC);
                }
            }
        };
        Item it;
        int pair_nb;
        for (int v = pair0; v < args.n_pairs; v += n_clusters) {
            if (!decode_pair_t<PLAYOUT>(args, v, rank, it, pair_nb)) continue;
            const int np = n_passes(pair_nb);
            for (int k = 0; k < n_steps; ++k)
            for (int ps = 0; ps < np; ++ps, ++pass) {
                if (ps == 0) {        // the feature planes of this move are in place in both CTAs (bulk copy, or the step phase)
                    mbar_wait(sBar + 8 * BAR_FFULL, n_done & 1u, 0x300u);
                    mbar_wait(sBar + 8 * BAR_PFFULL, n_done & 1u, 0x380u);
                    ++n_done;
                    if (PLAYOUT && args.prof && blockIdx.x == 0 && v == pair0 && k < 32 && lane == 0) args.prof[16 * k + 11] = clock64();
                }
                const Pass pi = pass_info(pair_nb, ps);
                n_tiles = pi.n_tiles;
                half_tile = half_last_tile<PLAYOUT>(pair_nb, pi);
                const int S = n_stages_of(pi.layer);
                // ---- the first HEAD stages (tap 0, row shift -11: tile t reads rows 128 t - 11 .. 128 t + 116): tile by tile.
                //      Between two 3x3 layers tile t starts as soon as the groups of tiles <= t have handed over (ACT[t]) -- the
                //      read-out of the previous pass's last tile then runs under the head of tiles 0..2.  Layer 0 reads the
                //      feature planes, so its tile t only needs accumulator slot t back: a new item starts under the head
                //      computation of the previous item's last layer, the second pass of layer 0 under the read-out of the first
                //      (groups whose slot this pass does not use are waited for after the middle stages).  Layer 1 reads what
                //      layer 0 wrote through another raster and waits for all four.
                //      The issuing thread consumes every phase of every ACT[t]: all four are waited for in every pass.
                const bool per_tile = BK_HEAD_PER_TILE && pass > 0 && (BK_HEAD_PER_TILE == 2 ? pi.layer >= 2 : pi.layer != 1);
                {
                    if (pass > 0 && !per_tile) {
#pragma unroll
                        for (int t = 0; t < 4; ++t) mbar_wait(sBar + 8 * (BAR_ACT + t), (pass - 1) & 1u, 0x400u + pass);
                    }
                    uint32_t slot[HEAD];
                    uint3 w[HEAD];
                    int nk[HEAD];
#pragma unroll
                    for (int r = 0; r < HEAD; ++r) {
                        slot[r] = st;
                        mbar_wait(sBar + 8 * (BAR_WFULL + st), ph, 0x520u + st);
                        if (++st == N_STAGES) { st = 0; ph ^= 1u; }
                        w[r] = windows(pi, r, nk[r]);
                    }
                    tc_fence_after();
                    if (!PLAYOUT && args.prof && blockIdx.x == 0 && pass < 64 && lane == 0) args.prof[pass * 4 + 0] = clock64();
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (per_tile && t < n_tiles) {
                            mbar_wait(sBar + 8 * (BAR_ACT + t), (pass - 1) & 1u, 0x440u + pass);
                            tc_fence_after();
                        }
                        if (t < n_tiles && elect_one()) {
#pragma unroll
                            for (int r = 0; r < HEAD; ++r) issue(w[r], nk[r], slot[r], t, t + 1, r != 0);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) {
#pragma unroll
                        for (int r = 0; r < HEAD; ++r) umma_commit_pair(sBar + 8 * (BAR_WEMPTY + slot[r]));
                    }
                    __syncwarp();
                }
                // ---- the middle stages: stage by stage, every tile uses the stage before its slot is handed back
                for (int s2 = HEAD; s2 < S - TAIL; ++s2) {
                    long long c0 = 0, c1 = 0;
                    if (profiling) c0 = clock64();
                    mbar_wait(sBar + 8 * (BAR_WFULL + st), ph, 0x500u + st);
                    if (profiling) { c1 = clock64(); tw += c1 - c0; }
                    tc_fence_after();
                    int nk;
                    const uint3 w = windows(pi, s2, nk);
                    if (elect_one()) {
                        issue(w, nk, st, 0, n_tiles, 1u);
                        umma_commit_pair(sBar + 8 * (BAR_WEMPTY + st));   // stage consumed -> both producers may refill
                    }
                    __syncwarp();
                    if (profiling) ti += clock64() - c1;
                    if (++st == N_STAGES) { st = 0; ph ^= 1u; }
                }
                if (per_tile) {       // the groups of the previous pass whose accumulator slot this pass does not use
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if (t >= n_tiles) mbar_wait(sBar + 8 * (BAR_ACT + t), (pass - 1) & 1u, 0x480u + pass);
                    tc_fence_after();
                }
                // ---- the last TAIL stages and the bias rows (x the all-ones operand: one K step, every row of the ones operand is
                //      the same): tile by tile, so that tile t is complete (ACC[t]) while the tiles behind it are still running and
                //      the TMEM read-out and the operand stores of its epilogue run under their MMAs
                {
                    long long c0 = 0, c1 = 0;
                    if (profiling) c0 = clock64();
                    uint32_t slot[TAIL];
                    uint3 w[TAIL];
                    int nk[TAIL];
#pragma unroll
                    for (int r = 0; r < TAIL; ++r) {
                        slot[r] = st;
                        mbar_wait(sBar + 8 * (BAR_WFULL + st), ph, 0x540u + st);
                        if (++st == N_STAGES) { st = 0; ph ^= 1u; }
                        w[r] = windows(pi, S - TAIL + r, nk[r]);
                    }
                    mbar_wait(sBar + 8 * BAR_BFULL, pass & 1u, 0x5C0u);
                    if (profiling) { c1 = clock64(); tp += c1 - c0; }
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            if (t < n_tiles) {
#pragma unroll
                                for (int r = 0; r < TAIL; ++r) issue(w[r], nk[r], slot[r], t, t + 1, 1u);
                                umma_f16(tmem + (uint32_t)(t * 128), one_lo, b_lo0, 1u, t == half_tile ? IDESC_HALF : IDESC);
                            }
                            umma_commit_pair(sBar + 8 * (BAR_ACC + t));   // (tiles the pass does not have complete with the last real one)
                        }
#pragma unroll
                        for (int r = 0; r < TAIL; ++r) umma_commit_pair(sBar + 8 * (BAR_WEMPTY + slot[r]));
                        umma_commit_pair(sBar + 8 * BAR_BEMPTY);
                        if (pi.l0_last) umma_commit_pair(sBar + 8 * BAR_FEMPTY);   // feature planes no longer needed
                    }
                    __syncwarp();
                    if (profiling) ti += clock64() - c1;
                }
                if (!PLAYOUT && args.prof && blockIdx.x == 0 && pass < 64 && lane == 0) {
                    args.prof[pass * 4 + 1] = clock64();
                    args.prof[256 + pass * 4 + 0] = tw; args.prof[256 + pass * 4 + 1] = tp; args.prof[256 + pass * 4 + 2] = ti;
                    tw = tp = ti = 0;
                }
                __syncwarp();
            }
        }
    } else {
        // =========================== epilogue warps (both CTAs, each on its own rows) ===========================
        const int quad = warp & 3, wq = warp >> 2;
        const uint32_t t_lane = tmem + ((uint32_t)(32 * quad) << 16);
        const uint32_t leader_act = mapa(sBar + 8 * (BAR_ACT + wq), 0);
        // PLAYOUT: "the planes of the next move are in place in this CTA" -- the leader's own FFULL, or, from the peer, straight
        // onto the leader's PFFULL (a relaxed arrive behind the proxy fences and the barrier of the 16 warps, like the pass
        // hand-over; no tensor work is in flight here; the forwarding warp's release.cluster round trip cost 1.4 k cycles a move)
        const uint32_t leader_pffull = mapa(sBar + 8 * BAR_PFFULL, 0);
        auto planes_ready = [&]() {
            if (rank == 0) mbar_arrive(sBar + 8 * BAR_FFULL);
            else { __threadfence_block(); mbar_arrive_cluster_relaxed(leader_pffull); }
        };
        uint32_t pass = 0;
        bool first = true;
        Item it;
        int pair_nb;
        // ENCODE: nnet.features of board j of item `e` into block j of the feature operand (and, for the first net's items, the
        // encoder's other outputs), by the three warps `sub` = 0..2 on barrier `sy`, group table in `grp`
        auto encode_board = [&](const Item &e, int j, BKGroups &grp, const BkSyncNamed &sy, int sub) {
            const int b = e.g * BK_GROUP + e.lo + j;
            const bool outs = e.net == args.first_net;
            bk_encode_board(sy, grp, 32 * sub + lane, b, args.boards + (size_t)b * BK_NSQ, args.ko, args.last, args.turn,
                            args.libs ? args.libs + (size_t)b * BK_NSQ : nullptr,
                            reinterpret_cast<uint4 *>(smem + OFF_F) + F_MARGIN + BK_F_ROWS_B * j, F_ROWS,
                            (outs && args.legal_out) ? args.legal_out + (size_t)b * BK_NSQ : nullptr,
                            (outs && args.libs_out) ? args.libs_out + (size_t)b * BK_NSQ : nullptr);
        };
        if (ENCODE) {
            // the planes of the CTA's first item: every board by its own three warps, tables in the (still unused) activation buffer
            Item f;
            int fnb;
            if (pair0 < args.n_pairs && decode_pair(args, pair0, rank, f, fnb)) {
                const int bi = warp / 3, sub = warp - 3 * bi;
                if (bi < f.nb) {
                    BkStepScratch &sc = *reinterpret_cast<BkStepScratch *>(smem + OFF_A + bi * STEP_SCRATCH);
                    encode_board(f, bi, sc.grp, BkSyncNamed{2 + bi}, sub);
                    uint4 *z = reinterpret_cast<uint4 *>(&sc);
                    for (int i = 32 * sub + lane; i < STEP_SCRATCH / 16; i += 96) z[i] = make_uint4(0u, 0u, 0u, 0u);
                }
                fence_proxy_async();
                named_bar_sync(1, N_EPI_WARPS * 32);
                if (threadIdx.x == 0) mbar_arrive(sBar + 8 * BAR_FFULL);
            }
        }
        for (int v = pair0; v < args.n_pairs; v += n_clusters) {
            if (!decode_pair_t<PLAYOUT>(args, v, rank, it, pair_nb)) continue;
            const int np = n_passes(pair_nb);
            // ENCODE: the item this CTA takes next; its boards are encoded by warps 0..2 at the top of the passes of layers 1..5
            // (one board per pass): layer 0 has read the current planes by then, and the tensor pipe is busy for ~19 k cycles
            Item nx;
            int nx_nb = 0;
            const bool have_next = ENCODE && v + n_clusters < args.n_pairs && decode_pair(args, v + n_clusters, rank, nx, nx_nb);
            const int first_enc = pair_nb == BK_GROUP ? 2 : 1;
            // PLAYOUT: three warps per board (warp 15 idles) own one board of the item for the whole playout: its position is
            // resident in shared memory (BkResident) plus one register per thread (the liberty-cache entry of the thread's square)
            const int bi = warp / 3, sub = warp - 3 * bi;
            const bool has_board = PLAYOUT && bi < it.nb;
            const int gb = it.g + bi;                         // PLAYOUT: global index of the trio's board
            BkStepScratch &sc = *reinterpret_cast<BkStepScratch *>(smem + OFF_A + (bi < BK_GROUP ? bi : 0) * STEP_SCRATCH);
            BkResident &res = reinterpret_cast<BkResident *>(smem + OFF_RES)[bi < BK_GROUP ? bi : 0];
            uint4 *const planes = reinterpret_cast<uint4 *>(smem + OFF_F) + F_MARGIN + BK_F_ROWS_B * (bi < BK_GROUP ? bi : 0);
            const BkSyncNamed sy{2 + bi};
            int my_lib = 0;
            if (PLAYOUT) {
                // ---- the starting positions: load, refresh the liberty cache, write the planes of move 0 (nnet.features)
                if (has_board) {
                    bk_resident_load(sy, sc, res, 32 * sub + lane, my_lib, gb, args.boards + (size_t)gb * BK_NSQ, args.ko, args.last,
                                     args.turn, args.libs + (size_t)gb * BK_NSQ, args.done, args.fresh_libs != 0, planes, F_ROWS);
                    uint4 *z = reinterpret_cast<uint4 *>(&sc);       // the scratch lives in the activation buffer: leave zeros
                    for (int i = 32 * sub + lane; i < STEP_SCRATCH / 16; i += 96) z[i] = make_uint4(0u, 0u, 0u, 0u);
                }
                fence_proxy_async();
                named_bar_sync(1, N_EPI_WARPS * 32);
                if (threadIdx.x == 0) planes_ready();
            }
            for (int k = 0; k < n_steps; ++k)
            for (int ps = 0; ps < np; ++ps, ++pass) {
                const int net = PLAYOUT ? ((args.first_turn + k) & 1) : it.net;
                const uint8_t *blob = args.blob[net];
                const Pass pi = pass_info(pair_nb, ps);
                // (warps 2, 3 and 6: schedulers 2 and 3 -- not the scheduler of the MMA-issuing warp 17 nor of the producer warp 16,
                // whose instruction issue paces the tensor pipe)
                const int enc_sub = warp == 2 ? 0 : (warp == 3 ? 1 : (warp == 6 ? 2 : -1));
                if (ENCODE && have_next && enc_sub >= 0) {
                    const int j = ps - first_enc, nj = nx.nb > 0 ? nx.nb : 1;
                    if (j >= 0 && j < nj) {
                        const BkSyncNamed sy{2};
                        if (j < nx.nb) encode_board(nx, j, *reinterpret_cast<BKGroups *>(smem + OFF_LOGIT), sy, enc_sub);
                        if (j == nj - 1) {                 // the next item's planes are complete (a CTA without boards just reports)
                            fence_proxy_async();
                            sy.sync();
                            if (enc_sub == 0 && lane == 0) mbar_arrive(sBar + 8 * BAR_FFULL);
                        }
                    }
                }
                // Tile wq is read out, and its rows rewritten in place, as soon as IT is complete (ACC[wq]): the MMAs still in
                // flight then are the last TAIL stages -- taps 7 and 8, row shifts +10 / +11 -- of the tiles behind it, which read
                // rows >= 128 (wq + 1) + 10 and write other TMEM columns (layer 0 reads the feature planes, not this buffer).
                // Only the read-out of the pass's last tile is left exposed.
                // EVERY group waits on its OWN barrier in EVERY pass (a tile the pass does not have completes with the last real
                // one): a parity wait cannot tell "phase k complete" from "phase k - 2 complete", so a warp may only wait on a
                // barrier whose every phase it consumes.  Round 1's form -- idle groups waiting on ACC[3] -- let a warp of group 2
                // run a whole pass ahead of the issuing thread (see the hand-over below).
                const int half_tile = half_last_tile<PLAYOUT>(pair_nb, pi);
                const bool own_tile = wq < pi.n_tiles;
#if BK_HANDOVER == 1   // measurement build: round 1's barrier choice (fails ~1 cold-L2 launch in 150 ... 12,000)
                mbar_wait(sBar + 8 * (BAR_ACC + (own_tile ? wq : 3)), pass & 1u, 0x600u + pass);
#else
                mbar_wait(sBar + 8 * (BAR_ACC + wq), pass & 1u, 0x600u + pass);
#endif
                tc_fence_after();
                const bool prof = !PLAYOUT && args.prof && blockIdx.x == 0 && pass < 64 && threadIdx.x == 32 * 12;   // last group
                if (prof) args.prof[pass * 4 + 2] = clock64();
                const bool dump = args.dump && blockIdx.x == 0 && first && ps == args.dump_pass;
                if (PLAYOUT && args.prof && blockIdx.x == 0 && first && k < 32 && threadIdx.x == 0) {
                    if (ps == 0) args.prof[16 * k + 10] = clock64();               // accumulators of the move's first pass ready
                    if (pi.layer == 6) args.prof[16 * k + 9] = clock64();          // accumulators of the last layer ready
                }
                if (PLAYOUT && wq == half_tile) {
                    // HALF tile (M = 128 MMAs): 64 rows; lanes 0..63 hold channels 0..63 of row = lane, lanes 64..127 channels
                    // 64..127 of row = lane - 64, in the first 64 columns of the slot -- two threads per row, 64 channels each
                    const int hc = quad >> 1;                                   // which half of the channels
                    const int r = 128 * (pi.tile0 + wq) + 32 * (quad & 1) + lane;
                    int board = 0, sq = 0;
                    const int dest = act_row_valid(r, it.nb, board, sq) ? r : -1;
                    uint32_t v0[32], v1[32];
                    tmem_ld32(t_lane + (uint32_t)(wq * 128), v0);
                    tmem_ld32(t_lane + (uint32_t)(wq * 128 + 32), v1);
                    tc_wait_ld();
                    if (pi.layer < 6) {
                        if (dest >= 0) {
                            store_act32(smem, v0, hc * 8, dest);
                            store_act32(smem, v1, hc * 8 + 4, dest);
                        }
                    } else {
                        // the 1x1 head is ONE chain of 128 fused multiply-adds per row (channel order), as in the full tiles: the
                        // thread with channels 0..63 starts it, hands the partial sum over, the other thread finishes it
                        const float4 *hw4 = reinterpret_cast<const float4 *>(blob + BK_W_HEADW_OFF);
                        float *hpart = reinterpret_cast<float *>(smem + OFF_A + HALF_HEAD_OFF);
                        const int slot = 32 * (quad & 1) + lane;
                        if (hc == 0) {
                            float hs = head_dot32(v0, hw4, 0.0f);
                            hpart[slot] = head_dot32(v1, hw4 + 8, hs);
                        }
                        named_bar_sync(8, 128);
                        if (hc == 1) {
                            float hsum = head_dot32(v0, hw4 + 16, hpart[slot]);
                            hsum = head_dot32(v1, hw4 + 24, hsum);
                            hpart[slot] = 0.0f;                                 // the activation buffer's padding must read as zero
                            if (dest >= 0) logit[board * 81 + sq] = hsum + __ldg(reinterpret_cast<const float *>(blob + BK_W_HEADB_OFF) + sq);
                        }
                    }
                } else if (own_tile) {
                    // one thread per GEMM row: all 128 output channels of row r (accumulator slot wq holds tile tile0 + wq)
                    const int r = 128 * (pi.tile0 + wq) + 32 * quad + lane;
                    int dest, board = 0, sq = 0;
                    if (pi.layer == 0) dest = l0_dest_row(r, it.nb);
                    else dest = act_row_valid(r, it.nb, board, sq) ? r : -1;
                    float hsum = 0.0f;
                    const float4 *hw4 = reinterpret_cast<const float4 *>(blob + BK_W_HEADW_OFF);
#ifdef BK_TRACE
                    unsigned int trace_h = 17u;
#endif
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t v0[32], v1[32];
                        tmem_ld32(t_lane + (uint32_t)(wq * 128 + h * 64), v0);
                        tmem_ld32(t_lane + (uint32_t)(wq * 128 + h * 64 + 32), v1);
                        tc_wait_ld();
#ifdef BK_TRACE
                        trace_h = (trace_h * 31u + v0[5]) * 31u + v1[17];      // (a light trace: the full hash moved the failure away)
#endif
                        if (dump) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                args.dump[(size_t)r * 128 + h * 64 + i] = __uint_as_float(v0[i]);
                                args.dump[(size_t)r * 128 + h * 64 + 32 + i] = __uint_as_float(v1[i]);
                            }
                        }
                        if (pi.layer < 6) {
                            if (dest >= 0) {
                                store_act32(smem, v0, h * 8, dest);
                                store_act32(smem, v1, h * 8 + 4, dest);
                            }
                        } else {
                            hsum = head_dot32(v0, hw4 + h * 16, hsum);
                            hsum = head_dot32(v1, hw4 + h * 16 + 8, hsum);
                        }
                    }
                    if (pi.layer == 6 && dest >= 0)   // 1x1 conv 128->1 plus the per-square bias
                        logit[board * 81 + sq] = hsum + __ldg(reinterpret_cast<const float *>(blob + BK_W_HEADB_OFF) + sq);
#ifdef BK_TRACE
                    if (g_trace && pass < TRACE_PASSES)   // (rows that are no square read past the operand buffers: not deterministic, never used)
                        g_trace[((size_t)blockIdx.x * TRACE_PASSES + pass) * 512 + threadIdx.x] = dest >= 0 ? trace_h : 0u;
#endif
                }
                // Hand-over: this group's TMEM reads are complete and its operand rows are visible to the tensor core; the group
                // arrives as soon as its own tile is done -- the leader starts the next pass when all 32 warps of the pair have.
                // Root cause of round 1's corrupted launches (1 cold-L2 launch in 150 ... 12,000, always 32-row blocks of tile 2
                // in a CTA's first item; tools/trace_handover.py, profiles/r02_handover_experiments.md): there, a group without a
                // tile waited on ACC[3] instead of its own barrier.  In the first item layer 0 runs as passes of 3 and 2 tiles, so
                // group 2 waited ACC[2] (parity 0), ACC[3] (parity 1), ACC[2] (parity 0).  With a cold L2 the issuing thread can
                // stall for a microsecond on an instruction fetch BETWEEN the commits of ACC[2] and ACC[3] of pass 0; a warp of
                // group 2 that had finished its read-out in the meantime found ACC[3] still in phase 0, for which a parity-1 wait
                // succeeds at once, "finished" pass 1, found ACC[2] in phase 1, for which a parity-0 wait succeeds at once, and
                // ran pass 2's read-out -- stale TMEM into 32 rows of the operand, two extra arrives -- before pass 1 had started.
                // Waiting for ACC[3] before the arrive (round 2's first fix, BK_HANDOVER=2) hid it because every warp then
                // consumed every phase of ACC[3]; waiting on the own barrier only removes it.
#if BK_HANDOVER == 2   // measurement build: the hand-over waits for the whole pass
                if (wq != 3) mbar_wait(sBar + 8 * (BAR_ACC + 3), pass & 1u, 0x680u + pass);
#endif
                tc_fence_before();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();     // the warp's stores are performed before the (relaxed) arrive can be observed
                    mbar_arrive_cluster_relaxed(leader_act);
                }
                if (prof) args.prof[pass * 4 + 3] = clock64();
                if (pi.layer == 6) {
                    // logit[] holds nb boards x 81 head outputs; it is next written seven passes from now, and
                    // every pass in between needs all 16 warps to have arrived, so no trailing barrier
                    named_bar_sync(1, N_EPI_WARPS * 32);
                    if (!PLAYOUT) {
                        if (warp < it.nb) {
                            const int b = it.g * BK_GROUP + it.lo + warp;
                            finish_board(logit + warp * 81, it.net, blob, args.logits ? args.logits + (size_t)b * 81 : nullptr,
                                         args.probs ? args.probs + (size_t)b * 81 : nullptr,
                                         args.value ? args.value + b : nullptr, lane);
                        }
                        // ENCODE: the logit area is the encode scratch of the next item's passes -- nobody may still be reading it
                        if (ENCODE) named_bar_sync(1, N_EPI_WARPS * 32);
                    } else {
                        // ---- step phase: the board's three warps sample the move from the policy's probabilities, play it
                        // and write the planes of the new position into the feature operand of the next move.  The activation
                        // buffer is dead here and lends its first bytes as scratch (zeroed again afterwards).
                        const bool last_move = k == n_steps - 1;
                        // diagnostic (bk_playout_run_debug): clock64 stamps per move of CTA 0's first item, taken by thread 0
                        long long *stamps = (args.prof && blockIdx.x == 0 && first && k < 32) ? args.prof + 16 * k : nullptr;
                        if (stamps && threadIdx.x == 0) stamps[6] = clock64();
                        if (has_board) {
                            float *pr = logit + bi * 81;
                            if (sub == 0) finish_board(pr, 0, blob, nullptr, pr, nullptr, lane);       // softmax in place
                            sy.sync();
                            if (stamps && threadIdx.x == 0) stamps[0] = clock64();
                            bk_resident_move(sy, sc, res, 32 * sub + lane, my_lib, pr, args.seed, args.game0 + (uint32_t)gb, args.mode,
                                             args.max_turn, args.moves_out + (size_t)k * args.B + gb, last_move ? nullptr : planes, F_ROWS,
                                             bi == 0 ? stamps : nullptr);
                            if (stamps && threadIdx.x == 0) stamps[7] = clock64();
                            sy.sync();                    // the trio is done with the scratch: zero it again
                            uint4 *z = reinterpret_cast<uint4 *>(&sc);
                            for (int i = 32 * sub + lane; i < STEP_SCRATCH / 16; i += 96) z[i] = make_uint4(0u, 0u, 0u, 0u);
                            if (last_move)                // the playout ends here: the position goes back to global memory
                                bk_resident_store(res, 32 * sub + lane, my_lib, gb, args.boards + (size_t)gb * BK_NSQ, args.ko, args.last,
                                                  args.turn, args.libs + (size_t)gb * BK_NSQ, args.done);
                        }
                        fence_proxy_async();              // planes and zeroed scratch are visible to the tensor core
                        named_bar_sync(1, N_EPI_WARPS * 32);
                        if (!last_move && threadIdx.x == 0) planes_ready();                        // planes of move k + 1 are in place
                        if (stamps && threadIdx.x == 0) stamps[8] = clock64();
                    }
                }
            }
            first = false;
        }
    }

    // ---- teardown: nobody leaves while the partner may still signal its barriers or read its shared memory ------
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == WARP_PRODUCER) { __syncwarp(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------------
// validation kernel: the same packed operands on CUDA cores, one CTA per (board, net), thread = out channel
// ------------------------------------------------------------------------------------------------------
constexpr int SIMT_FIN_BYTES = 169 * 32 * 2;          // 13x13 padded input, 32 ch
constexpr int SIMT_ACT_BYTES = 121 * 128 * 2;         // 11x11 padded activations, 128 ch
constexpr int SIMT_SMEM = SIMT_FIN_BYTES + 2 * SIMT_ACT_BYTES + 4 * 81 * 4 + 96 * 4;

__global__ void __launch_bounds__(128) bk_forward_simt_kernel(const FwdArgs args)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __half *fin = reinterpret_cast<__half *>(smem);
    __half *act0 = reinterpret_cast<__half *>(smem + SIMT_FIN_BYTES);
    __half *act1 = reinterpret_cast<__half *>(smem + SIMT_FIN_BYTES + SIMT_ACT_BYTES);
    float *red = reinterpret_cast<float *>(smem + SIMT_FIN_BYTES + 2 * SIMT_ACT_BYTES);   // [4][81]
    float *logit = red + 4 * 81;
    const int item = blockIdx.x;
    const int b = item / args.n_nets, net = args.first_net + item % args.n_nets;
    const int co = threadIdx.x, lane = co & 31, warp = co >> 5;
    const uint8_t *blob = args.blob[net];
    const int g = b / BK_GROUP, bi = b - g * BK_GROUP;

    for (int i = co; i < (SIMT_FIN_BYTES + 2 * SIMT_ACT_BYTES) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    __syncthreads();
    // unpack this board's planes from the conv layout into a 13x13 zero-padded raster
    const uint4 *src = reinterpret_cast<const uint4 *>(args.feats + (size_t)g * BK_F_GROUP_BYTES);
    for (int i = co; i < 81 * BK_F_CHUNKS; i += 128) {
        const int p = i / BK_F_CHUNKS, c = i - p * BK_F_CHUNKS;
        const int x = p / 9, y = p - 9 * x;
        const uint4 v = src[c * BK_F_ROWS_G + bi * BK_F_ROWS_B + 22 + 11 * x + y];
        *reinterpret_cast<uint4 *>(fin + ((x + 2) * 13 + (y + 2)) * 32 + c * 8) = v;
    }
    __syncthreads();

    float acc[81];
    const float *bias = reinterpret_cast<const float *>(blob + BK_W_BIAS_OFF);
    __half *in = act0, *out = act1;
    for (int layer = 0; layer < 7; ++layer) {
#pragma unroll
        for (int p = 0; p < 81; ++p) acc[p] = 0.0f;
        const int ntap = layer == 0 ? 25 : 9, nchunk = layer == 0 ? 4 : 16, kw = layer == 0 ? 5 : 3;
        const uint8_t *wl = blob + (layer == 0 ? BK_W_L0_OFF : BK_W_L_OFF(layer));
        for (int tap = 0; tap < ntap; ++tap) {
            const int ti = tap / kw, tj = tap - kw * ti;
            for (int c = 0; c < nchunk; ++c) {
                const int k = tap * nchunk * 8 + c * 8;
                const uint4 wv = *reinterpret_cast<const uint4 *>(wl + BK_W_OFF(k, co));
                const __half2 *wh = reinterpret_cast<const __half2 *>(&wv);
                const float2 w01 = __half22float2(wh[0]), w23 = __half22float2(wh[1]), w45 = __half22float2(wh[2]),
                             w67 = __half22float2(wh[3]);
#pragma unroll
                for (int p = 0; p < 81; ++p) {
                    const int x = p / 9, y = p - 9 * (p / 9);
                    const uint4 av = layer == 0
                                         ? *reinterpret_cast<const uint4 *>(fin + ((x + ti) * 13 + (y + tj)) * 32 + c * 8)
                                         : *reinterpret_cast<const uint4 *>(in + ((x + ti) * 11 + (y + tj)) * 128 + c * 8);
                    const __half2 *ah = reinterpret_cast<const __half2 *>(&av);
                    const float2 a01 = __half22float2(ah[0]), a23 = __half22float2(ah[1]), a45 = __half22float2(ah[2]),
                                 a67 = __half22float2(ah[3]);
                    float s = acc[p];
                    s = fmaf(a01.x, w01.x, s); s = fmaf(a01.y, w01.y, s);
                    s = fmaf(a23.x, w23.x, s); s = fmaf(a23.y, w23.y, s);
                    s = fmaf(a45.x, w45.x, s); s = fmaf(a45.y, w45.y, s);
                    s = fmaf(a67.x, w67.x, s); s = fmaf(a67.y, w67.y, s);
                    acc[p] = s;
                }
            }
        }
        const float bv = bias[layer * 128 + co];
        if (layer < 6) {
#pragma unroll
            for (int p = 0; p < 81; ++p) {
                const int x = p / 9, y = p - 9 * (p / 9);
                out[((x + 1) * 11 + (y + 1)) * 128 + co] = __float2half_rn(fmaxf(acc[p] + bv, 0.0f));
            }
            __syncthreads();
            __half *tmp = in; in = out; out = tmp;
            if (layer == 0) { in = act1; out = act0; }
        } else {
            const float hw = reinterpret_cast<const float *>(blob + BK_W_HEADW_OFF)[co];
#pragma unroll
            for (int p = 0; p < 81; ++p) {
                float v = fmaxf(acc[p] + bv, 0.0f) * hw;
#pragma unroll
                for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[warp * 81 + p] = v;
            }
            __syncthreads();
            if (co < 81)
                logit[co] = red[co] + red[81 + co] + red[162 + co] + red[243 + co] +
                            reinterpret_cast<const float *>(blob + BK_W_HEADB_OFF)[co];
            __syncthreads();
            if (warp == 0)
                finish_board(logit, net, blob, args.logits ? args.logits + (size_t)b * 81 : nullptr,
                             args.probs ? args.probs + (size_t)b * 81 : nullptr, args.value ? args.value + b : nullptr, lane);
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static inline uint16_t f2h(float f)
{
    const __half h = __float2half_rn(f);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
}
static inline float h2f(uint16_t u)
{
    __half h;
    memcpy(&h, &u, 2);
    return __half2float(h);
}

// Packs BatchNorm-folded fp32 parameters into the blob described in bk_layout.h (host memory -> host memory).
//   w0 [128][27][5][5], w16 [6][128][128][3][3], bias [7][128], head_w [128], head_b [81],
//   vtail (nullable) = {bn_scale, bn_shift, lin2_b, W1'[64][81], b1'[64], w2[64]}
extern "C" int bk_weights_pack(const float *w0, const float *w16, const float *bias, const float *head_w,
                               const float *head_b, const float *vtail, void *blob_out)
{
    if (!w0 || !w16 || !bias || !head_w || !head_b || !blob_out) return -1;
    uint8_t *blob = static_cast<uint8_t *>(blob_out);
    memset(blob, 0, BK_W_BLOB_BYTES);
    uint8_t *l0 = blob + BK_W_L0_OFF;
    auto put = [](uint8_t *at, uint16_t v) { memcpy(at, &v, 2); };
    for (int co = 0; co < 128; ++co)
        for (int ci = 0; ci < 27; ++ci)
            for (int tap = 0; tap < 25; ++tap) put(l0 + BK_W_OFF(tap * 32 + ci, co), f2h(w0[(co * 27 + ci) * 25 + tap]));
    // folded bias as two extra K rows (fp16 hi + lo) that meet the all-ones operand, behind the stages of each layer
    for (int co = 0; co < 128; ++co) {
        const uint16_t hi = f2h(bias[co]);
        put(l0 + BK_W_BIAS_ROW_OFF(BK_L0_STAGES, 0, co), hi);
        put(l0 + BK_W_BIAS_ROW_OFF(BK_L0_STAGES, 1, co), f2h(bias[co] - h2f(hi)));
    }
    for (int l = 1; l <= 6; ++l) {
        uint8_t *ll = blob + BK_W_L_OFF(l);
        const float *wl = w16 + (size_t)(l - 1) * 128 * 128 * 9;
        for (int co = 0; co < 128; ++co) {
            const float b = bias[l * 128 + co];
            const uint16_t hi = f2h(b);
            put(ll + BK_W_BIAS_ROW_OFF(BK_L_STAGES, 0, co), hi);
            put(ll + BK_W_BIAS_ROW_OFF(BK_L_STAGES, 1, co), f2h(b - h2f(hi)));
            for (int ci = 0; ci < 128; ++ci)
                for (int tap = 0; tap < 9; ++tap)
                    put(ll + BK_W_OFF(tap * 128 + ci, co), f2h(wl[((size_t)co * 128 + ci) * 9 + tap]));
        }
    }
    memcpy(blob + BK_W_BIAS_OFF, bias, 7 * 128 * 4);
    memcpy(blob + BK_W_HEADW_OFF, head_w, 128 * 4);
    memcpy(blob + BK_W_HEADB_OFF, head_b, 81 * 4);
    if (vtail) {
        float *vt = reinterpret_cast<float *>(blob + BK_W_VT_OFF);
        vt[0] = vtail[0]; vt[1] = vtail[1]; vt[2] = vtail[2]; vt[3] = 0.0f;
        float *w1t = reinterpret_cast<float *>(blob + BK_W_VT_W1T_OFF);
        const float *w1 = vtail + 3;
        for (int j = 0; j < 64; ++j)
            for (int p = 0; p < 81; ++p) w1t[p * 64 + j] = w1[j * 81 + p];
        memcpy(blob + BK_W_VT_B1_OFF, vtail + 3 + 64 * 81, 64 * 4);
        memcpy(blob + BK_W_VT_W2_OFF, vtail + 3 + 64 * 81 + 64, 64 * 4);
    }
    return 0;
}

// Host-side launch state.  Everything that depends on the device (kernel attributes, SM count, tensor maps over a blob) is
// kept PER DEVICE and guarded by one mutex, so the entry points may be called from several host threads and for several
// devices of one process (ctypes releases the GIL during a call).
static std::mutex g_fwd_mutex;
static unsigned int *g_dbg_host = nullptr;   // pinned, portable, device-visible; survives a trapped kernel
struct FwdMapEntry { const void *blob; CUtensorMap stage, bias; unsigned long long used; };
struct FwdDeviceState {
    int n_sm = 0;                    // 0 = tcgen05 kernel attributes not set yet on this device
    bool simt_attr = false;
    FwdMapEntry maps[8] = {};        // tensor maps depend on the blob address only; a handful of nets are live at a time
    unsigned long long tick = 0;
};
static FwdDeviceState g_fwd_dev[BK_MAX_DEVICES];

// Tensor map over the conv weights of one blob: rows of 512 bytes, boxes of 8 rows (one CTA's half of a stage).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_weight_map(const void *blob, int box_rows, CUtensorMap *out)
{
    static EncodeTiledFn encode = nullptr;   // process-wide driver entry point; callers hold g_fwd_mutex
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return -3;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    memset(out, 0, sizeof(*out));
    if (!blob) return 0;
    const cuuint64_t dims[2] = {TM_ROW_BYTES / 2, (cuuint64_t)TM_ROWS};
    const cuuint64_t strides[1] = {TM_ROW_BYTES};
    const cuuint32_t box[2] = {TM_ROW_BYTES / 2, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(blob), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

// tensor maps of `blob` on the current device (callers hold g_fwd_mutex); zero maps for a null blob
static int weight_maps(FwdDeviceState &ds, const void *blob, CUtensorMap *stage, CUtensorMap *bias)
{
    if (!blob) {
        memset(stage, 0, sizeof(*stage));
        memset(bias, 0, sizeof(*bias));
        return 0;
    }
    FwdMapEntry *victim = &ds.maps[0];
    for (FwdMapEntry &e : ds.maps) {
        if (e.blob == blob) {
            e.used = ++ds.tick;
            *stage = e.stage; *bias = e.bias;
            return 0;
        }
        if (e.used < victim->used) victim = &e;
    }
    int rc = make_weight_map(blob, TM_BOX_ROWS, &victim->stage);
    if (rc == 0) rc = make_weight_map(blob, TM_BIAS_BOX_ROWS, &victim->bias);
    if (rc != 0) { victim->blob = nullptr; victim->used = 0; return rc; }
    victim->blob = blob;
    victim->used = ++ds.tick;
    *stage = victim->stage; *bias = victim->bias;
    return 0;
}

#ifdef BK_TRACE
// measurement build: device buffer of n_blocks x 24 passes x 512 words (or null to switch the trace off)
extern "C" int bk_debug_trace(unsigned int *dev_buf)
{
    return cudaMemcpyToSymbol(g_trace, &dev_buf, sizeof(dev_buf)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int bk_debug_words(unsigned int *out8)
{
    for (int i = 0; i < 8; ++i) out8[i] = g_dbg_host ? g_dbg_host[i] : 0u;
    return g_dbg_host ? 0 : -1;
}

// positions for the ENCODE mode (planes computed inside the kernel); boards == nullptr: planes come from feats_conv
struct FwdPositions {
    const int8_t *boards; const int16_t *ko, *last, *turn; const uint8_t *libs_in; uint8_t *legal_out, *libs_out;
};
static int forward_impl(const void *feats_conv, const void *blob_policy, const void *blob_value, float *logits,
                        float *probs, float *value, int B, int flags, cudaStream_t stream, float *dump, int dump_pass,
                        long long *prof, const FwdPositions *pos = nullptr);

// diagnostics: dump = raw accumulators [640][128] of pass `dump_pass` of CTA 0's first item; prof = clock64 stamps of
// CTA 0, four per pass {MMA issue start, MMA issue end, accumulators ready, epilogue done}, room for 64 passes
extern "C" int bk_forward_debug(const void *feats_conv, const void *blob_policy, const void *blob_value, float *logits,
                                float *probs, float *value, int B, int flags, cudaStream_t stream, float *dump, int dump_pass,
                                long long *prof)
{
    return forward_impl(feats_conv, blob_policy, blob_value, logits, probs, value, B, flags, stream, dump, dump_pass, prof);
}

extern "C" int bk_forward(const void *feats_conv, const void *blob_policy, const void *blob_value, float *logits,
                          float *probs, float *value, int B, int flags, cudaStream_t stream)
{
    return forward_impl(feats_conv, blob_policy, blob_value, logits, probs, value, B, flags, stream, nullptr, -1, nullptr);
}

// bk_forward_positions: nnet.features + both nets in ONE launch -- the ENCODE instantiation of the conv kernel computes the
// planes of every item from the positions while the tensor pipe works on the item before (see MODE 2 at the kernel)
extern "C" int bk_forward_positions(const int8_t *boards, const int16_t *ko, const int16_t *last, const int16_t *turn,
                                    const uint8_t *libs_in, const void *blob_policy, const void *blob_value, float *logits,
                                    float *probs, float *value, uint8_t *legal_out, uint8_t *libs_out, int B, int flags,
                                    cudaStream_t stream)
{
    if (B <= 0) return 0;
    if (!boards || !ko || !last || !turn || (flags & BK_FWD_SIMT)) return -1;
    // with both nets every board is encoded twice (once per net, by different CTAs): an in-place cache update would let one of
    // them read a half-updated cache
    if (libs_out && libs_out == libs_in && (flags & BK_FWD_POLICY) && (flags & BK_FWD_VALUE)) return -1;
    const FwdPositions pos = {boards, ko, last, turn, libs_in, legal_out, libs_out};
    return forward_impl(nullptr, blob_policy, blob_value, logits, probs, value, B, flags, stream, nullptr, -1, nullptr, &pos);
}

static int forward_impl(const void *feats_conv, const void *blob_policy, const void *blob_value, float *logits,
                        float *probs, float *value, int B, int flags, cudaStream_t stream, float *dump, int dump_pass,
                        long long *prof, const FwdPositions *pos)
{
    if (B <= 0) return 0;
    const bool do_p = flags & BK_FWD_POLICY, do_v = flags & BK_FWD_VALUE;
    if (!do_p && !do_v) return -1;
    if ((do_p && !blob_policy) || (do_v && (!blob_value || !value)) || (!feats_conv && !pos)) return -1;
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    if (pos) {
        a.boards = const_cast<int8_t *>(pos->boards); a.ko = const_cast<int16_t *>(pos->ko); a.last = const_cast<int16_t *>(pos->last);
        a.turn = const_cast<int16_t *>(pos->turn); a.libs = const_cast<uint8_t *>(pos->libs_in);
        a.legal_out = pos->legal_out; a.libs_out = pos->libs_out;
    }
    a.feats = static_cast<const uint8_t *>(feats_conv);
    a.blob[0] = static_cast<const uint8_t *>(blob_policy);
    a.blob[1] = static_cast<const uint8_t *>(blob_value);
    a.logits = logits; a.probs = probs; a.value = value;
    a.B = B; a.G = (B + BK_GROUP - 1) / BK_GROUP;
    a.n_nets = (do_p ? 1 : 0) + (do_v ? 1 : 0);
    a.first_net = do_p ? 0 : 1;
    a.g_whole = a.G; a.split = 1; a.n_sub = a.G; a.n_pairs = 0;
    a.dump = dump; a.dump_pass = dump_pass; a.prof = prof;
    a.diag = prof ? ((flags & 0x200) ? 1 : 0) | ((flags & 0x400) ? 2 : 0) : 0;
    const int slot = bk_current_device_slot();
    if (slot < 0) return -2;
    std::lock_guard<std::mutex> lock(g_fwd_mutex);     // launches of one process are serialised here; the kernels are not
    FwdDeviceState &ds = g_fwd_dev[slot];
    if (!g_dbg_host) {
        if (cudaHostAlloc((void **)&g_dbg_host, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) g_dbg_host = nullptr;
        else memset(g_dbg_host, 0, 64);
    }
    a.dbg = nullptr;
    if (g_dbg_host) cudaHostGetDevicePointer((void **)&a.dbg, g_dbg_host, 0);
    cudaError_t e;
    if (flags & BK_FWD_SIMT) {
        if (!ds.simt_attr) {
            e = cudaFuncSetAttribute(bk_forward_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SIMT_SMEM);
            if (e != cudaSuccess) return -3;
            ds.simt_attr = true;
        }
        bk_forward_simt_kernel<<<B * a.n_nets, 128, SIMT_SMEM, stream>>>(a);
    } else {
        if (!ds.n_sm) {
            int n_sm = 0;
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, slot);
            e = cudaFuncSetAttribute(bk_forward_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(bk_forward_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(bk_forward_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
            if (e != cudaSuccess || n_sm < 2) return -3;
            ds.n_sm = n_sm;
        }
        const int n_sm = ds.n_sm;
        // Schedule (see FwdArgs): CTA pairs take two groups of one net at a time.  The pairs of whole groups fill
        // complete rounds of the n_sm / 2 clusters; the groups left over for the last, partial round are split into
        // board ranges so that it spreads over the idle SMs (fewer M tiles per CTA).
        const int n_clusters = n_sm / 2;
        const int whole_pairs = a.n_nets * (a.G / 2);
        const int rounds = whole_pairs / n_clusters;
        a.g_whole = (flags & BK_FWD_NOSPLIT) ? a.G : 2 * (rounds * n_clusters / a.n_nets);
        const int rest = (a.G - a.g_whole) * a.n_nets;          // whole-group items left for the last round
        if (rest > 0) {
            a.split = n_sm / rest < BK_GROUP ? n_sm / rest : BK_GROUP;
            if (a.split < 1) a.split = 1;
        }
        a.n_sub = a.g_whole + (a.G - a.g_whole) * a.split;
        a.n_pairs = a.n_nets * ((a.n_sub + 1) / 2);
        const int grid = 2 * (a.n_pairs < n_clusters ? a.n_pairs : n_clusters);
        CUtensorMap map[2], bias_map[2];
        for (int i = 0; i < 2; ++i) {
            const int rc = weight_maps(ds, a.blob[i], &map[i], &bias_map[i]);
            if (rc != 0) return rc;
        }
        if (pos) bk_forward_tc_kernel<2><<<grid, N_THREADS, SMEM_BYTES, stream>>>(a, map[0], map[1], bias_map[0], bias_map[1]);
        else bk_forward_tc_kernel<0><<<grid, N_THREADS, SMEM_BYTES, stream>>>(a, map[0], map[1], bias_map[0], bias_map[1]);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

static int playout_impl(int8_t *boards, int16_t *ko, int16_t *last, int16_t *turn, uint8_t *libs, uint8_t *done,
                        const void *blob_even, const void *blob_odd, uint64_t seed, uint32_t game0, int mode, int max_turn,
                        int first_turn, int n_steps, int fresh_libs, int16_t *moves_out, int B, cudaStream_t stream, long long *prof);

// Whole playouts in ONE launch (see the PLAYOUT note at the top): B boards, n_steps moves each.
extern "C" int bk_playout_run(int8_t *boards, int16_t *ko, int16_t *last, int16_t *turn, uint8_t *libs, uint8_t *done,
                              const void *blob_even, const void *blob_odd, uint64_t seed, uint32_t game0, int mode, int max_turn,
                              int first_turn, int n_steps, int fresh_libs, int16_t *moves_out, int B, cudaStream_t stream)
{
    return playout_impl(boards, ko, last, turn, libs, done, blob_even, blob_odd, seed, game0, mode, max_turn, first_turn, n_steps,
                        fresh_libs, moves_out, B, stream, nullptr);
}

// diagnostics: prof = room for 32 moves x 16 clock64 stamps of CTA 0's first item (tools/prof_playout.py)
extern "C" int bk_playout_run_debug(int8_t *boards, int16_t *ko, int16_t *last, int16_t *turn, uint8_t *libs, uint8_t *done,
                                    const void *blob_even, const void *blob_odd, uint64_t seed, uint32_t game0, int mode,
                                    int max_turn, int first_turn, int n_steps, int fresh_libs, int16_t *moves_out, int B,
                                    cudaStream_t stream, long long *prof)
{
    return playout_impl(boards, ko, last, turn, libs, done, blob_even, blob_odd, seed, game0, mode, max_turn, first_turn, n_steps,
                        fresh_libs, moves_out, B, stream, prof);
}

static int playout_impl(int8_t *boards, int16_t *ko, int16_t *last, int16_t *turn, uint8_t *libs, uint8_t *done,
                        const void *blob_even, const void *blob_odd, uint64_t seed, uint32_t game0, int mode, int max_turn,
                        int first_turn, int n_steps, int fresh_libs, int16_t *moves_out, int B, cudaStream_t stream, long long *prof)
{
    if (B <= 0 || n_steps <= 0) return 0;
    if (!boards || !ko || !last || !turn || !libs || !done || !blob_even || !moves_out) return -1;
    if (mode != 0 && mode != 1) return -1;
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    a.blob[0] = static_cast<const uint8_t *>(blob_even);
    a.blob[1] = static_cast<const uint8_t *>(blob_odd ? blob_odd : blob_even);
    a.B = B; a.G = (B + BK_GROUP - 1) / BK_GROUP;
    a.n_nets = 1; a.first_net = 0;
    a.dump_pass = -1;
    a.prof = prof;
    a.boards = boards; a.ko = ko; a.last = last; a.turn = turn; a.libs = libs; a.done = done; a.moves_out = moves_out;
    a.n_steps = n_steps; a.mode = mode; a.max_turn = max_turn; a.first_turn = first_turn; a.seed = seed; a.game0 = game0;
    a.fresh_libs = fresh_libs;
    const int slot = bk_current_device_slot();
    if (slot < 0) return -2;
    std::lock_guard<std::mutex> lock(g_fwd_mutex);
    FwdDeviceState &ds = g_fwd_dev[slot];
    if (!g_dbg_host) {
        if (cudaHostAlloc((void **)&g_dbg_host, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) g_dbg_host = nullptr;
        else memset(g_dbg_host, 0, 64);
    }
    if (g_dbg_host) cudaHostGetDevicePointer((void **)&a.dbg, g_dbg_host, 0);
    if (!ds.n_sm) {
        int n_sm = 0;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, slot);
        cudaError_t e = cudaFuncSetAttribute(bk_forward_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(bk_forward_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(bk_forward_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess || n_sm < 2) return -3;
        ds.n_sm = n_sm;
    }
    // Boards per item.  An item keeps its boards for the whole playout, so the run lasts (rounds of the grid) x (time of one
    // item's move), and an item's move grows with its M tiles: measured on the B200 (profiles/r02k_playout_per_move.jsonl)
    // 48 / 66 / 76 / 85 / 100 us for 1 .. 5 boards (3 and 4 boards run their last tile as a half tile, 5 boards need two layer-0
    // passes).  Up to two rounds of full groups take the size that minimises rounds x move time (512 boards -> 4 per item in one
    // round, 1,024 -> 4 per item in two rounds); beyond that full groups of 5, the best throughput per board.
    const int n_clusters = ds.n_sm / 2;
    static const double move_us[BK_GROUP + 1] = {0.0, 48.0, 66.0, 76.0, 85.0, 100.0};
    int grp = BK_GROUP;
    double best = 1e30;
    for (int g = 1; g <= BK_GROUP && B <= 2 * BK_GROUP * ds.n_sm; ++g) {
        const int pairs = ((B + g - 1) / g + 1) / 2;
        const double t = (double)((pairs + n_clusters - 1) / n_clusters) * move_us[g];
        if (t < best - 1e-9) { best = t; grp = g; }
    }
    a.play_group = grp;
    a.n_sub = (B + grp - 1) / grp;
    a.g_whole = a.n_sub; a.split = 1;
    a.n_pairs = (a.n_sub + 1) / 2;
    const int grid = 2 * (a.n_pairs < n_clusters ? a.n_pairs : n_clusters);
    CUtensorMap map[2], bias_map[2];
    for (int i = 0; i < 2; ++i) {
        const int rc = weight_maps(ds, a.blob[i], &map[i], &bias_map[i]);
        if (rc != 0) return rc;
    }
    bk_forward_tc_kernel<1><<<grid, N_THREADS, SMEM_BYTES, stream>>>(a, map[0], map[1], bias_map[0], bias_map[1]);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
