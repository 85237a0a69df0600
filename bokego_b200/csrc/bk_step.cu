// bk_step.cu -- kernel (c): batched playout stepping, scoring, and the shared random stream.
//
//   bk_step_kernel   three warps per board (thread = square): pick a move from the policy's probabilities and play it.
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
//       With `feats_conv` the kernel goes on to encode the position AFTER the move (nnet.features with the carried liberty
//       cache, bk_encode_core.cuh) into the conv operand of the next policy evaluation: "sample, play, capture, re-encode"
//       is one launch, and a playout move is two launches (policy forward, this kernel) instead of three.
//   bk_score_kernel  Game.score() (go.py:202-218) and the +-1 reward of Go_MCTS.reward (mcts.py:330-338).
#include <cuda_runtime.h>
#include <stdint.h>

#include "bk_bitboard.cuh"
#include "bk_encode_core.cuh"
#include "bk_step_core.cuh"

namespace {

// One block of three warps per board, thread t = square t (as in bk_encode_kernel); the move itself is bk_step_board
// (bk_step_core.cuh), shared with the persistent playout kernel.
__global__ void __launch_bounds__(96)
bk_step_kernel(int8_t *__restrict__ boards, int16_t *__restrict__ ko_arr, int16_t *__restrict__ last_arr,
               int16_t *__restrict__ turn_arr, uint8_t *__restrict__ libs, uint8_t *__restrict__ done,
               const float *__restrict__ probs, const float *__restrict__ q_inj, int q_vecs, uint64_t seed,
               uint32_t game0, int mode, int max_turn, int16_t *__restrict__ moves_out, uint4 *__restrict__ feats_conv, int B)
{
    __shared__ BkStepScratch sc;
    const int b = (int)blockIdx.x;
    if (b >= B) return;
    bk_step_board(BkSyncBlock(), sc, (int)threadIdx.x, b, boards + (size_t)b * BK_NSQ, ko_arr, last_arr, turn_arr,
                  libs ? libs + (size_t)b * BK_NSQ : nullptr, done, probs + (size_t)b * BK_NSQ,
                  q_inj ? q_inj + (size_t)b * q_vecs * BK_NSQ : nullptr, q_vecs, seed, game0 + (uint32_t)b, mode, max_turn,
                  moves_out ? moves_out + b : nullptr, feats_conv ? bk_conv_base(feats_conv, b) : nullptr, BK_F_ROWS_G);
}

__global__ void __launch_bounds__(128)
bk_score_kernel(const int8_t *__restrict__ boards, float komi, float *__restrict__ score_out,
                int8_t *__restrict__ reward_out, int B)
{
    const int b = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    BB black, white;
    bk_load_boards(boards + (size_t)b * BK_NSQ, lane, black, white);
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
    bk_load_boards(boards + (size_t)par * BK_NSQ, lane, black, white);
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

// per-game records of a finished batch of playouts: rec int16 [B][T + 3] = { turn reached, reward, 2 * score, move 0 .. T-1 }
// from the move log int16 [T][B] of the playout loop (bokego_b200/playout.py: the rows that cross ranks in the result gather)
__global__ void bk_pack_records_kernel(const int16_t *__restrict__ moves, const int16_t *__restrict__ turn,
                                       const float *__restrict__ score, const int8_t *__restrict__ reward,
                                       int16_t *__restrict__ rec, int T, int B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int W = T + 3;
    if (i >= B * W) return;
    const int b = i / W, c = i - b * W;
    int16_t v;
    if (c == 0) v = turn[b];
    else if (c == 1) v = (int16_t)reward[b];
    else if (c == 2) v = (int16_t)__float2int_rn(score[b] * 2.0f);
    else v = moves[(size_t)(c - 3) * B + b];
    rec[i] = v;
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
    if (!boards || !ko || !last || !turn || !done || !probs) return -1;
    bk_step_kernel<<<B, 96, 0, stream>>>(boards, ko, last, turn, libs, done, probs, q_inj, q_vecs, seed, game0,
                                                    mode, max_turn, moves_out, nullptr, B);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int bk_playout_step_encode(int8_t *boards, int16_t *ko, int16_t *last, int16_t *turn, uint8_t *libs, uint8_t *done,
                                     const float *probs, const float *q_inj, int q_vecs, uint64_t seed, uint32_t game0,
                                     int mode, int max_turn, int16_t *moves_out, void *feats_conv, int B, cudaStream_t stream)
{
    if (B <= 0) return 0;
    if (!boards || !ko || !last || !turn || !done || !probs || !libs || !feats_conv) return -1;
    bk_step_kernel<<<B, 96, 0, stream>>>(boards, ko, last, turn, libs, done, probs, q_inj, q_vecs, seed, game0,
                                                    mode, max_turn, moves_out, (uint4 *)feats_conv, B);
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

extern "C" int bk_pack_records(const int16_t *moves, const int16_t *turn, const float *score, const int8_t *reward,
                               int16_t *rec, int T, int B, cudaStream_t stream)
{
    if (B <= 0) return 0;
    if (!moves || !turn || !score || !reward || !rec || T < 0) return -1;
    const int n = B * (T + 3);
    bk_pack_records_kernel<<<(n + 255) / 256, 256, 0, stream>>>(moves, turn, score, reward, rec, T, B);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
