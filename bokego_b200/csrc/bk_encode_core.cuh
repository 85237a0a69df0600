// bk_encode_core.cuh -- the per-square part of nnet.features (/root/reference/bokego/nnet.py:213-262) shared by the encoder
// kernel (bk_encode.cu) and the fused "play a move, then encode the new position" kernel (bk_step.cu).  Device only.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "bk_bitboard.cuh"
#include "bk_layout.h"

__device__ __forceinline__ uint32_t bk_half_bits(int n)   // fp16 bit pattern of a small non-negative integer
{
    return (uint32_t)__half_as_ushort(__int2half_rn(n));
}
__device__ __forceinline__ uint32_t bk_pack2(int lo, int hi) { return bk_half_bits(lo) | (bk_half_bits(hi) << 16); }

// first uint4 of board `slot` inside the conv operand (bk_layout.h): [group][4 chunks][605 rows], row 121 * board + ...
__device__ __forceinline__ uint4 *bk_conv_base(uint4 *feats_conv, int slot)
{
    const int g = slot / BK_GROUP, bi = slot - g * BK_GROUP;
    return feats_conv + (size_t)g * (BK_F_CHUNKS * BK_F_ROWS_G) + bi * BK_F_ROWS_B;
}

// Planes of square p of board b from the group table `grp` of the position (black, white, ko, last, turn parity `blk`).
// carried = the position has a liberty cache (Game._libs): lib_carried is its entry for p, `stale` = last >= 0 and the
// cache entry of `last` is 0 (go.py:226); otherwise exact liberties (fresh Game).  Every output pointer may be null.
// Returns the new liberty-cache entry of p (what libs_out receives).
// conv_base = first uint4 of this board's rows in the conv operand, chunk_stride = distance between channel chunks in uint4
// (BK_F_ROWS_G in global memory; the row count of the shared-memory copy inside the conv kernel).
__device__ __forceinline__ int bk_encode_square(const BKGroups &grp, BB black, BB white, bool blk, int ko, int last, bool carried,
                                                 bool stale, int lib_carried, int p, size_t b, uint4 *conv_base, int chunk_stride,
                                                 float *feats_f32, uint8_t *planes_u8, uint8_t *legal_out, uint8_t *libs_out)
{
    const BB own = blk ? black : white, opp = blk ? white : black;
    const bool mine = bb_test(own, p), theirs = bb_test(opp, p);
    int lib = 0;
    if (mine || theirs) {
        if (!carried) lib = bb_count(bk_group_libs(grp, p));            // fresh Game: exact liberties
        else lib = bk_groups_lazy_lib(grp, black, white, last, stale, p, lib_carried);
    } else if (carried) {
        lib = lib_carried;                                               // stale values persist on empty squares (go.py:220-243)
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
    if (libs_out) libs_out[b * BK_NSQ + p] = (uint8_t)lib;
    if (legal_out) legal_out[b * BK_NSQ + p] = (uint8_t)lg;

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
        for (int c = 0; c < 27; ++c) planes_u8[(b * 27 + c) * BK_NSQ + p] = (uint8_t)v[c];
    }
    if (feats_f32) {
#pragma unroll
        for (int c = 0; c < 27; ++c) feats_f32[(b * 27 + c) * BK_NSQ + p] = (float)v[c];
    }
    if (conv_base) {
        const int x = p / 9, y = p - 9 * x;
        const int r = 22 + 11 * x + y;
#pragma unroll
        for (int c = 0; c < BK_F_CHUNKS; ++c) {
            uint4 o;
            o.x = bk_pack2(v[8 * c + 0], v[8 * c + 1]);
            o.y = bk_pack2(v[8 * c + 2], v[8 * c + 3]);
            o.z = bk_pack2(v[8 * c + 4], v[8 * c + 5]);
            o.w = bk_pack2(v[8 * c + 6], v[8 * c + 7]);
            conv_base[c * chunk_stride + r] = o;
        }
    }
    return lib;
}
