// bk_step.cu -- kernel (c): batched playout stepping, scoring, and the shared random stream.
//
//   bk_step_kernel   one warp per board: pick a move from the policy's probabilities and play it.
//       mode 0 (mcts flavour)      Go_MCTS.get_move + make_move + is_game_over
//                                  (/root/reference/bokego/mcts.py:340-364): exponential-race sample,
//                                  reject illegal moves and own-eye fills by zeroing their probability and
//                                  drawing again, PASS when no mass is left (SURVEY F7 shim), terminal when
//                                  turn > max_turn or the move was PASS.
//       mode 1 (self-play flavour) legal_sample + playout loop (/root/reference/bin/selfplay.py:18-47): one
//                                  unmasked draw; if illegal the legal move of highest probability (ties:
//                                  lowest index); none => the game stops; stop when turn > max_turn + 1.
//       The sample is torch.multinomial's single-draw algorithm, argmax_i p_i / q_i with q ~ Exp(1), IEEE
//       fp32 division, first maximum.  q comes either from a caller-supplied buffer (parity tests) or from
//       the counter-based stream bk_exp_draw(seed, game, turn, try, square) -- identical bits on host and
//       device, independent of how games are sharded over GPUs.
//       The move is played with Game.play_move semantics (go.py:123-182) including the lazy liberty cache
//       refresh that precedes the board update (go.py:160).
//   bk_score_kernel  Game.score() (go.py:202-218) and the +-1 reward of Go_MCTS.reward (mcts.py:330-338).
#include <cuda_runtime.h>
#include <stdint.h>

#include "bk_bitboard.cuh"

namespace {

__device__ __forceinline__ void load_boards(const int8_t *bd, int lane, BB &black, BB &white)
{
    uint32_t bl[3], wh[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int p = lane + 32 * k;
        const int v = p < BK_NSQ ? (int)bd[p] : 0;
        bl[k] = __ballot_sync(0xffffffffu, v == 1);
        wh[k] = __ballot_sync(0xffffffffu, v == -1);
    }
    black.w[0] = bl[0] & BK_M27;
    black.w[1] = ((bl[0] >> 27) | (bl[1] << 5)) & BK_M27;
    black.w[2] = ((bl[1] >> 22) | (bl[2] << 10)) & BK_M27;
    white.w[0] = wh[0] & BK_M27;
    white.w[1] = ((wh[0] >> 27) | (wh[1] << 5)) & BK_M27;
    white.w[2] = ((wh[1] >> 22) | (wh[2] << 10)) & BK_M27;
}

// warp-wide (max value, lowest index) -- the "first maximum" of a sequential argmax
__device__ __forceinline__ void warp_argmax(float &v, int &i)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

__global__ void __launch_bounds__(128)
bk_step_kernel(int8_t *__restrict__ boards, int16_t *__restrict__ ko_arr, int16_t *__restrict__ last_arr,
               int16_t *__restrict__ turn_arr, uint8_t *__restrict__ libs, uint8_t *__restrict__ done,
               const float *__restrict__ probs, const float *__restrict__ q_inj, int q_vecs, uint64_t seed,
               uint32_t game0, int mode, int max_turn, int16_t *__restrict__ moves_out, int B)
{
    const int b = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    if (done[b]) {
        if (lane == 0 && moves_out) moves_out[b] = -3;
        return;
    }
    int8_t *bd = boards + (size_t)b * BK_NSQ;
    int ko = ko_arr[b], last = last_arr[b], turn = turn_arr[b];
    BB black, white;
    load_boards(bd, lane, black, white);
    const bool blk = (turn & 1) == 0;
    const BB own = blk ? black : white, opp = blk ? white : black;
    const int me = blk ? 1 : -1;

    // per-square probability and accept mask (bit k = square lane + 32k)
    float pr[3];
    uint32_t ok_mask = 0u;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int p = lane + 32 * k;
        pr[k] = 0.0f;
        if (p < BK_NSQ) {
            pr[k] = probs[(size_t)b * BK_NSQ + p];
            bool ok = bb_is_legal_quirk(own, opp, ko, p);
            if (mode == 0 && ok) ok = bb_possible_eye(black, white, p) != me;
            if (ok) ok_mask |= 1u << k;
        }
    }

    int mv = BK_NONE;
    int t = 0;
    for (;;) {
        if (t > 0) {
            const bool mass = pr[0] > 0.0f || pr[1] > 0.0f || pr[2] > 0.0f;
            if (!__any_sync(0xffffffffu, mass)) { mv = BK_PASS; break; }
            if (q_inj && t >= q_vecs) { mv = -4; break; }
        }
        float bv = -1.0f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int p = lane + 32 * k;
            if (p < BK_NSQ) {
                const float q = q_inj ? q_inj[((size_t)b * q_vecs + t) * BK_NSQ + p]
                                      : bk_exp_draw(seed, game0 + (uint32_t)b, (uint32_t)turn, (uint32_t)t, p);
                const float v = __fdiv_rn(pr[k], q);
                if (v > bv) { bv = v; bi = p; }
            }
        }
        warp_argmax(bv, bi);
        ++t;
        const int owner = bi & 31, slot = bi >> 5;
        const bool accept = (__shfl_sync(0xffffffffu, ok_mask, owner) >> slot) & 1u;
        if (mode == 1) {
            if (accept) { mv = bi; break; }
            // highest-probability legal move, lowest index on ties
            float fv = -1.0f;
            int fi = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int p = lane + 32 * k;
                if (p < BK_NSQ && ((ok_mask >> k) & 1u) && pr[k] > fv) { fv = pr[k]; fi = p; }
            }
            warp_argmax(fv, fi);
            mv = fi == 0x7fffffff ? BK_NONE : fi;
            break;
        }
        if (accept) { mv = bi; break; }
        if (t - 1 >= BK_NSQ) { mv = BK_PASS; break; }   // tries >= 81 (mcts.py:354)
        if (lane == owner) {
            if (slot == 0) pr[0] = 0.0f; else if (slot == 1) pr[1] = 0.0f; else pr[2] = 0.0f;
        }
    }

    if (lane == 0 && moves_out) moves_out[b] = (int16_t)mv;
    if (mv == BK_NONE || mv == -4) {
        if (lane == 0) done[b] = 1;
        return;
    }

    // lazy liberty cache on the position BEFORE the move (go.py:160); a PASS does not touch it
    if (libs && mv >= 0) {
        uint8_t *lb = libs + (size_t)b * BK_NSQ;
        const bool stale = last >= 0 && lb[last] == 0;
        __syncwarp();
        int nl[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int p = lane + 32 * k;
            nl[k] = p < BK_NSQ ? bb_lazy_lib_of(black, white, last, stale, p, (int)lb[p]) : 0;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int p = lane + 32 * k;
            if (p < BK_NSQ) lb[p] = (uint8_t)nl[k];
        }
    }

    const int st = bb_play(black, white, ko, last, turn, mv);
    if (st != 0) {   // cannot happen for a position reached by legal play; flag instead of corrupting state
        if (lane == 0) { done[b] = 1; if (moves_out) moves_out[b] = (int16_t)(-10 - st); }
        return;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int p = lane + 32 * k;
        if (p < BK_NSQ) bd[p] = bb_test(black, p) ? 1 : (bb_test(white, p) ? -1 : 0);
    }
    if (lane == 0) {
        ko_arr[b] = (int16_t)ko; last_arr[b] = (int16_t)last; turn_arr[b] = (int16_t)turn;
        const bool over = mode == 0 ? (turn > max_turn || last == BK_PASS) : (turn > max_turn + 1);
        if (over) done[b] = 1;
    }
}

__global__ void __launch_bounds__(128)
bk_score_kernel(const int8_t *__restrict__ boards, float komi, float *__restrict__ score_out,
                int8_t *__restrict__ reward_out, int B)
{
    const int b = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    BB black, white;
    load_boards(boards + (size_t)b * BK_NSQ, lane, black, white);
    if (lane == 0) {
        const float sc = (float)bb_score_diff(black, white) - komi;
        if (score_out) score_out[b] = sc;
        if (reward_out) reward_out[b] = sc > 0.0f ? 1 : -1;
    }
}

// Go_MCTS.make_move (/root/reference/bokego/mcts.py:340-346) for a batch of children: child c is a copy of parent
// parent_idx[c] with moves[c] played by Game.play_move (go.py:123-182; -1 = play_pass go.py:109-121).  One warp per child.
// The carried liberty cache is refreshed on the PARENT position before the move, as go.py:160 does, and handed to the child
// (deepcopy carries `_libs`, mcts.py:281-307); a parent without a cache (libs == nullptr) hands down exact liberties of the
// parent position, which is what the first get_liberties() call of a fresh Game computes.
// status: 0 ok, 1 ko, 2 not_empty, 3 suicide -- the reference raises IllegalMove there; the child is then the unchanged parent.
__global__ void __launch_bounds__(128)
bk_make_moves_kernel(const int8_t *__restrict__ boards, const int16_t *__restrict__ ko_arr, const int16_t *__restrict__ last_arr,
                     const int16_t *__restrict__ turn_arr, const uint8_t *__restrict__ libs, const int32_t *__restrict__ parent_idx,
                     const int16_t *__restrict__ moves, int8_t *__restrict__ boards_out, int16_t *__restrict__ ko_out,
                     int16_t *__restrict__ last_out, int16_t *__restrict__ turn_out, uint8_t *__restrict__ libs_out,
                     uint8_t *__restrict__ status_out, int C)
{
    const int c = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    const int par = parent_idx[c];
    const int mv = moves[c];
    int ko = ko_arr[par], last = last_arr[par], turn = turn_arr[par];
    BB black, white;
    load_boards(boards + (size_t)par * BK_NSQ, lane, black, white);
    if (libs_out) {
        const uint8_t *lb = libs ? libs + (size_t)par * BK_NSQ : nullptr;
        const bool stale = lb && mv >= 0 && last >= 0 && lb[last] == 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int p = lane + 32 * k;
            if (p < BK_NSQ) {
                int v;
                if (!lb) v = bb_exact_lib_of(black, white, p);
                else v = mv >= 0 ? bb_lazy_lib_of(black, white, last, stale, p, (int)lb[p]) : (int)lb[p];
                libs_out[(size_t)c * BK_NSQ + p] = (uint8_t)v;
            }
        }
    }
    const int st = bb_play(black, white, ko, last, turn, mv);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int p = lane + 32 * k;
        if (p < BK_NSQ) boards_out[(size_t)c * BK_NSQ + p] = bb_test(black, p) ? 1 : (bb_test(white, p) ? -1 : 0);
    }
    if (lane == 0) {
        ko_out[c] = (int16_t)ko; last_out[c] = (int16_t)last; turn_out[c] = (int16_t)turn;
        if (status_out) status_out[c] = (uint8_t)st;
    }
}

__global__ void bk_exp_draws_kernel(uint64_t seed, uint32_t game0, uint32_t move, uint32_t tr, float *__restrict__ q, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * BK_NSQ) return;
    const int b = i / BK_NSQ, p = i - b * BK_NSQ;
    q[i] = bk_exp_draw(seed, game0 + (uint32_t)b, move, tr, p);
}

}  // namespace

extern "C" int bk_playout_step(int8_t *boards, int16_t *ko, int16_t *last, int16_t *turn, uint8_t *libs, uint8_t *done,
                              const float *probs, const float *q_inj, int q_vecs, uint64_t seed, uint32_t game0, int mode,
                              int max_turn, int16_t *moves_out, int B, cudaStream_t stream)
{
    if (B <= 0) return 0;
    bk_step_kernel<<<(B + 3) / 4, 128, 0, stream>>>(boards, ko, last, turn, libs, done, probs, q_inj, q_vecs, seed, game0,
                                                    mode, max_turn, moves_out, B);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int bk_make_moves(const int8_t *boards, const int16_t *ko, const int16_t *last, const int16_t *turn, const uint8_t *libs,
                             const int32_t *parent_idx, const int16_t *moves, int8_t *boards_out, int16_t *ko_out,
                             int16_t *last_out, int16_t *turn_out, uint8_t *libs_out, uint8_t *status_out, int C,
                             cudaStream_t stream)
{
    if (C <= 0) return 0;
    if (!boards || !ko || !last || !turn || !parent_idx || !moves || !boards_out || !ko_out || !last_out || !turn_out) return -1;
    bk_make_moves_kernel<<<(C + 3) / 4, 128, 0, stream>>>(boards, ko, last, turn, libs, parent_idx, moves, boards_out, ko_out,
                                                          last_out, turn_out, libs_out, status_out, C);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int bk_score(const int8_t *boards, float komi, float *score_out, int8_t *reward_out, int B,
                               cudaStream_t stream)
{
    if (B <= 0) return 0;
    bk_score_kernel<<<(B + 3) / 4, 128, 0, stream>>>(boards, komi, score_out, reward_out, B);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int bk_exp_draws(uint64_t seed, uint32_t game0, uint32_t move, uint32_t tr, float *q, int B,
                                   cudaStream_t stream)
{
    if (B <= 0) return 0;
    const int n = B * BK_NSQ;
    bk_exp_draws_kernel<<<(n + 255) / 256, 256, 0, stream>>>(seed, game0, move, tr, q, B);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
