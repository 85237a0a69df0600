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
#include "bk_encode_core.cuh"
#include "bk_layout.h"

namespace {

__global__ void __launch_bounds__(96)
bk_encode_kernel(const int8_t *__restrict__ boards, const int16_t *__restrict__ ko_arr,
                 const int16_t *__restrict__ last_arr, const int16_t *__restrict__ turn_arr,
                 const uint8_t *libs_in, uint4 *__restrict__ feats_conv,
                 float *__restrict__ feats_f32, uint8_t *__restrict__ planes_u8,
                 uint8_t *__restrict__ legal_out, uint8_t *libs_out, int B, int slots)   // libs_out may alias libs_in
{
    // one block of three warps per board: warp w owns squares 32w .. 32w+31 (one square per thread)
    const int slot = (int)blockIdx.x;
    const int lane = threadIdx.x & 31, wsq = threadIdx.x >> 5;
    if (slot >= slots) return;

    uint4 *conv_base = nullptr;
    if (feats_conv) {
        conv_base = bk_conv_base(feats_conv, slot);
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
    const bool carried = libs_in != nullptr;
    const bool stale = carried && last >= 0 && libs_in[(size_t)b * BK_NSQ + last] == 0;

    // ---- group table of the board (shared by the block): every stone floods its own group once; the group's lowest
    //      square is its id and publishes the group's stones and liberties.  Empty squares then evaluate "what if a stone
    //      is put here" from their <= 4 neighbouring groups by look-up instead of flooding up to five groups each. --------
    __shared__ BKGroups grp;
    const int p = lane + 32 * wsq;
    const bool active = p < BK_NSQ;
    bk_groups_build(grp, black, white, p);
    __syncthreads();   // the table is complete; also: every thread has read libs_in[last], so libs_out may alias libs_in
    if (!active) return;

    // ---- per-square evaluation (bk_encode_core.cuh) ---------------------------------------------------
    bk_encode_square(grp, black, white, blk, ko, last, carried, stale, carried ? (int)libs_in[(size_t)b * BK_NSQ + p] : 0, p,
                     (size_t)b, conv_base, BK_F_ROWS_G, feats_f32, planes_u8, legal_out, libs_out);
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
