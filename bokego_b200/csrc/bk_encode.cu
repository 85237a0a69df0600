// bk_encode.cu -- kernel (a): batched board -> 27 feature planes, three warps (one thread per square) per 9x9 board.
//
// Replaces nnet.features (/root/reference/bokego/nnet.py:182-262) and the go.py routines it calls
// (get_legal_moves go.py:245-260, get_liberties go.py:220-243, get_caps go.py:404-418).
// Integer only; bit-exact against the reference in both liberty-cache modes.
//
// Work split: one block of three warps per board; the board's two colour sets are built with warp ballots into
// 81-bit bit-boards that every thread keeps in registers; thread t owns square t.  A stone floods its own group
// once (iterated dilation) and the group's lowest square publishes its stones and liberties in a small
// shared-memory group table; an empty square evaluates the move "stone on t" from its <= 4 neighbouring groups by
// look-up (captures with the reference's double counting, liberties of the merged group) and popcounts.
// No atomics.  Outputs are optional (null pointer = skip):
//   feats_conv  fp16 operand layout of the conv kernel: [group of 5 boards][4 channel chunks]
//               [605 rows][8 channels]; row = 121*board_in_group + 22 + 11*x + y, i.e. a stride-11
//               raster with two zero columns and two zero rows between boards, so that every tap of
//               the 5x5 first layer is a pure row shift.  Pad rows/columns and channels 27..31 are 0.
//   feats_f32   float32 [B][27][9][9], what nnet.features returns
//   planes_u8   uint8   [B][27][81], the same values as bytes
//   legal_out   uint8   [B][81]   (plane 5)
//   libs_out    uint8   [B][81]   Game._libs after the call
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "bk_bitboard.cuh"
#include "bk_layout.h"

namespace {

__device__ __forceinline__ uint32_t half_bits(int n)   // fp16 bit pattern of a small non-negative integer
{
    return (uint32_t)__half_as_ushort(__int2half_rn(n));
}

__device__ __forceinline__ uint32_t pack2(int lo, int hi) { return half_bits(lo) | (half_bits(hi) << 16); }

__global__ void __launch_bounds__(96)
bk_encode_kernel(const int8_t *__restrict__ boards, const int16_t *__restrict__ ko_arr,
                 const int16_t *__restrict__ last_arr, const int16_t *__restrict__ turn_arr,
                 const uint8_t *__restrict__ libs_in, uint4 *__restrict__ feats_conv,
                 float *__restrict__ feats_f32, uint8_t *__restrict__ planes_u8,
                 uint8_t *__restrict__ legal_out, uint8_t *__restrict__ libs_out, int B, int slots)
{
    // one block of three warps per board: warp w owns squares 32w .. 32w+31 (one square per thread)
    const int slot = (int)blockIdx.x;
    const int lane = threadIdx.x & 31, wsq = threadIdx.x >> 5;
    if (slot >= slots) return;

    uint4 *conv_base = nullptr;
    if (feats_conv) {
        const int g = slot / BK_GROUP, bi = slot - g * BK_GROUP;
        conv_base = feats_conv + (size_t)g * (BK_F_CHUNKS * BK_F_ROWS_G) + bi * BK_F_ROWS_B;
        // zero rows/columns of this board's block (and the whole block for a slot past the batch)
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int r = threadIdx.x; r < BK_F_ROWS_B; r += 96) {
            const bool pad = r < 22 || ((r - 22) % 11) >= 9 || slot >= B;
            if (pad) {
#pragma unroll
                for (int c = 0; c < BK_F_CHUNKS; ++c) conv_base[c * BK_F_ROWS_G + r] = z;
            }
        }
    }
    if (slot >= B) return;
    const int b = slot;

    // ---- load the position, build the bit-boards ------------------------------------------------
    const int ko = ko_arr[b], last = last_arr[b], turn = turn_arr[b];
    uint32_t bl[3], wh[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int p = lane + 32 * k;
        const int v = p < BK_NSQ ? (int)boards[(size_t)b * BK_NSQ + p] : 0;
        bl[k] = __ballot_sync(0xffffffffu, v == 1);
        wh[k] = __ballot_sync(0xffffffffu, v == -1);
    }
    BB black, white;
    black.w[0] = bl[0] & BK_M27;
    black.w[1] = ((bl[0] >> 27) | (bl[1] << 5)) & BK_M27;
    black.w[2] = ((bl[1] >> 22) | (bl[2] << 10)) & BK_M27;
    white.w[0] = wh[0] & BK_M27;
    white.w[1] = ((wh[0] >> 27) | (wh[1] << 5)) & BK_M27;
    white.w[2] = ((wh[1] >> 22) | (wh[2] << 10)) & BK_M27;
    const bool blk = (turn & 1) == 0;
    const BB own = blk ? black : white, opp = blk ? white : black;
    const bool carried = libs_in != nullptr;
    const bool stale = carried && last >= 0 && libs_in[(size_t)b * BK_NSQ + last] == 0;

    // ---- group table of the board (shared by the block): every stone floods its own group once; the group's lowest
    //      square is its id and publishes the group's stones and liberties.  Empty squares then evaluate "what if a stone
    //      is put here" from their <= 4 neighbouring groups by look-up instead of flooding up to five groups each. --------
    __shared__ BKGroups grp;
    const int p = lane + 32 * wsq;
    const bool active = p < BK_NSQ;
    const bool mine = active && bb_test(own, p), theirs = active && bb_test(opp, p);
    bk_groups_build(grp, black, white, p);
    __syncthreads();   // the table is complete; also: every thread has read libs_in[last], so libs_out may alias libs_in
    if (!active) return;

    // ---- per-square evaluation --------------------------------------------------------------------
    {
        int lib = 0;
        if (mine || theirs) {
            if (!carried) lib = bb_count(bk_group_libs(grp, p));            // fresh Game: exact liberties
            else lib = bk_groups_lazy_lib(grp, black, white, last, stale, p, (int)libs_in[(size_t)b * BK_NSQ + p]);
        } else if (carried) {
            lib = (int)libs_in[(size_t)b * BK_NSQ + p];                      // stale values persist on empty squares (go.py:220-243)
        }
        int la = 0, cp = 0;
        bool lg = false;
        if (!mine && !theirs) {
            // the move "own stone on p" (nnet.py:241-247, go.py:404-418): a dead opponent group is counted once per
            // neighbour of p that belongs to it; liberties of the merged own group after the removal
            const Cand c = bk_groups_candidate(grp, own, opp, p, nullptr);
            lg = bb_listed_legal(own, opp, ko, p, c);
            if (lg) { la = c.libs_after; cp = c.caps; }
        }
        if (libs_out) libs_out[(size_t)b * BK_NSQ + p] = (uint8_t)lib;
        if (legal_out) legal_out[(size_t)b * BK_NSQ + p] = (uint8_t)lg;

        // plane values (nnet.py:249-262): planes 6..12 / 13..19 / 20..26 hold min(v,7) in slot min(v,7)-1
        const int l7 = lib > 6 ? 7 : lib, a7 = la > 6 ? 7 : la, c7 = cp > 6 ? 7 : cp;
        int v[32];
        v[0] = mine; v[1] = theirs; v[2] = (!mine && !theirs); v[3] = blk; v[4] = (p == last); v[5] = lg;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            v[6 + i] = (l7 == i + 1) ? l7 : 0;
            v[13 + i] = (a7 == i + 1) ? a7 : 0;
            v[20 + i] = (c7 == i + 1) ? c7 : 0;
        }
#pragma unroll
        for (int i = 27; i < 32; ++i) v[i] = 0;

        if (planes_u8) {
#pragma unroll
            for (int c = 0; c < 27; ++c) planes_u8[((size_t)b * 27 + c) * BK_NSQ + p] = (uint8_t)v[c];
        }
        if (feats_f32) {
#pragma unroll
            for (int c = 0; c < 27; ++c) feats_f32[((size_t)b * 27 + c) * BK_NSQ + p] = (float)v[c];
        }
        if (conv_base) {
            const int x = p / 9, y = p - 9 * x;
            const int r = 22 + 11 * x + y;
#pragma unroll
            for (int c = 0; c < BK_F_CHUNKS; ++c) {
                uint4 o;
                o.x = pack2(v[8 * c + 0], v[8 * c + 1]);
                o.y = pack2(v[8 * c + 2], v[8 * c + 3]);
                o.z = pack2(v[8 * c + 4], v[8 * c + 5]);
                o.w = pack2(v[8 * c + 6], v[8 * c + 7]);
                conv_base[c * BK_F_ROWS_G + r] = o;
            }
        }
    }
}

// float32 planes [B][27][81] (what nnet.features returns) -> the conv kernel's fp16 operand layout.
// Used by the drop-in PolicyNet/ValueNet.forward, whose argument is the reference's float tensor.
__global__ void __launch_bounds__(128)
bk_repack_kernel(const float *__restrict__ feats_f32, uint4 *__restrict__ feats_conv, int B, int slots)
{
    const int slot = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (slot >= slots) return;
    const int g = slot / BK_GROUP, bi = slot - g * BK_GROUP;
    uint4 *base = feats_conv + (size_t)g * (BK_F_CHUNKS * BK_F_ROWS_G) + bi * BK_F_ROWS_B;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int r = lane; r < BK_F_ROWS_B; r += 32) {
        const int q = r - 22;
        const int x = q / 11, y = q - 11 * x;
        const bool pad = r < 22 || y >= 9 || slot >= B;
        if (pad) {
#pragma unroll
            for (int c = 0; c < BK_F_CHUNKS; ++c) base[c * BK_F_ROWS_G + r] = z;
        } else {
            const float *src = feats_f32 + (size_t)slot * 27 * BK_NSQ + 9 * x + y;
#pragma unroll
            for (int c = 0; c < BK_F_CHUNKS; ++c) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = (8 * c + e) < 27 ? src[(8 * c + e) * BK_NSQ] : 0.0f;
                uint4 o;
                const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
                const __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
                o.x = *reinterpret_cast<const uint32_t *>(&h0); o.y = *reinterpret_cast<const uint32_t *>(&h1);
                o.z = *reinterpret_cast<const uint32_t *>(&h2); o.w = *reinterpret_cast<const uint32_t *>(&h3);
                base[c * BK_F_ROWS_G + r] = o;
            }
        }
    }
}

}  // namespace

extern "C" int bk_repack_f32(const float *feats_f32, void *feats_conv, int B, cudaStream_t stream)
{
    if (B <= 0) return 0;
    if (!feats_f32 || !feats_conv) return -1;
    const int slots = ((B + BK_GROUP - 1) / BK_GROUP) * BK_GROUP;
    bk_repack_kernel<<<(slots + 3) / 4, 128, 0, stream>>>(feats_f32, (uint4 *)feats_conv, B, slots);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int bk_encode(const int8_t *boards, const int16_t *ko, const int16_t *last, const int16_t *turn,
                                const uint8_t *libs_in, void *feats_conv, float *feats_f32, uint8_t *planes_u8,
                                uint8_t *legal_out, uint8_t *libs_out, int B, cudaStream_t stream)
{
    if (B <= 0) return 0;
    if (!boards || !ko || !last || !turn) return -1;
    const int slots = feats_conv ? ((B + BK_GROUP - 1) / BK_GROUP) * BK_GROUP : B;
    bk_encode_kernel<<<slots, 96, 0, stream>>>(boards, ko, last, turn, libs_in, (uint4 *)feats_conv,
                                                                 feats_f32, planes_u8, legal_out, libs_out, B, slots);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
