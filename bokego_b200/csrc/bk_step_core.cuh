// bk_step_core.cuh -- one playout move of ONE board by 96 threads (thread t = square t): sample, play, capture, refresh the
// liberty cache and (optionally) encode the new position.  Shared by the stepping kernel (bk_step.cu: one block per board)
// and by the persistent playout kernel (bk_forward.cu: three epilogue warps per board, on a named barrier).  Device only.
//   mode 0  Go_MCTS.get_move + make_move + is_game_over   /root/reference/bokego/mcts.py:340-364
//   mode 1  legal_sample + the playout loop                /root/reference/bin/selfplay.py:18-47
//   Game.play_move with the lazy liberty refresh before the board update   /root/reference/bokego/go.py:123-182, 160
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bk_bitboard.cuh"
#include "bk_encode_core.cuh"

// shared-memory scratch of one board
struct BkStepScratch {
    BKGroups grp;
    float s_v[3];
    int s_i[3];
    uint8_t s_ok[96];
};

// how the 96 threads of a board synchronise: the whole block ...
struct BkSyncBlock {
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ bool sync_or(bool p) const { return __syncthreads_or(p) != 0; }
};
// ... or three warps of a larger block on the named barrier `id`
struct BkSyncNamed {
    int id;
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, 96;" ::"r"(id) : "memory"); }
    __device__ __forceinline__ bool sync_or(bool p) const
    {
        uint32_t r;
        asm volatile(
            "{\n\t.reg .pred pi, po;\n\t"
            "setp.ne.u32 pi, %1, 0;\n\t"
            "bar.red.or.pred po, %2, 96, pi;\n\t"
            "selp.u32 %0, 1, 0, po;\n\t}"
            : "=r"(r)
            : "r"((uint32_t)p), "r"(id)
            : "memory");
        return r != 0;
    }
};

__device__ __forceinline__ void bk_load_boards(const int8_t *bd, int lane, BB &black, BB &white)
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
__device__ __forceinline__ void bk_warp_argmax(float &v, int &i)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

// (max value, lowest index) over the 96 threads: warp shuffles, then the three warps' candidates through shared memory
template <class Sync>
__device__ __forceinline__ void bk_board_argmax(const Sync &sy, float &v, int &i, float *sv, int *si, int wsq, int lane)
{
    bk_warp_argmax(v, i);
    sy.sync();                             // the previous round's readers are done with sv / si
    if (lane == 0) { sv[wsq] = v; si[wsq] = i; }
    sy.sync();
    v = sv[0]; i = si[0];
#pragma unroll
    for (int w = 1; w < 3; ++w)
        if (sv[w] > v || (sv[w] == v && si[w] < i)) { v = sv[w]; i = si[w]; }
}

// One move of board b.  tid = 0..95 (thread = square, 81..95 idle); every branch that leads to a barrier is uniform over the
// 96 threads.  bd / libs_row: this board's 81 bytes; probs_row: its 81 probabilities (global or shared memory); q_row: injected
// draws [q_vecs][81] or null (then the counter-based stream of game `game_id`); move_out: where the move code goes (or null);
// conv_base / chunk_stride (in uint4): where the planes of the position AFTER the move are written (bk_encode_core.cuh), null =
// no re-encode.  Boards that are done are skipped (move code -3).
template <class Sync>
__device__ __forceinline__ void bk_step_board(const Sync &sy, BkStepScratch &sc, int tid, int b, int8_t *bd, int16_t *ko_arr,
                                              int16_t *last_arr, int16_t *turn_arr, uint8_t *libs_row, uint8_t *done,
                                              const float *probs_row, const float *q_row, int q_vecs, uint64_t seed,
                                              uint32_t game_id, int mode, int max_turn, int16_t *move_out, uint4 *conv_base,
                                              int chunk_stride)
{
    const int lane = tid & 31, wsq = tid >> 5;
    const int p = tid;                               // this thread's square (81..95 idle)
    const bool active = p < BK_NSQ;
    if (done[b]) {
        if (tid == 0 && move_out) *move_out = -3;
        return;
    }
    int ko = ko_arr[b], last = last_arr[b], turn = turn_arr[b];
    BB black, white;
    bk_load_boards(bd, lane, black, white);
    const bool blk = (turn & 1) == 0;
    const BB own = blk ? black : white, opp = blk ? white : black;
    const int me = blk ? 1 : -1;
    bk_groups_build(sc.grp, black, white, p);
    const bool stale = libs_row && last >= 0 && libs_row[last] == 0;
    sy.sync();

    // probability and accept flag of this thread's square: Game.is_legal (go.py:184-200, early exit kept) and, in the
    // mcts flavour, not an own eye (mcts.py:354)
    float pr = 0.0f;
    bool ok = false;
    if (active) {
        pr = probs_row[p];
        if (!bb_test(own, p) && !bb_test(opp, p)) {
            int nb[4];
            const int n = bb_nbr_list(p, nb);
            const BB occ = bb_or(own, opp);
            int empties = 0;
            bool early = false;
            for (int k = 0; k < n; ++k) {
                if (empties > 1) { early = true; break; }
                if (!bb_test(occ, nb[k])) ++empties;
            }
            ok = early || (p != ko && bk_groups_candidate(sc.grp, own, opp, p, nullptr).libs_after > 0);
            if (mode == 0 && ok) ok = bb_possible_eye(black, white, p) != me;
        }
        sc.s_ok[p] = ok;
    }
    sy.sync();

    int mv = BK_NONE;
    int t = 0;
    for (;;) {
        if (t > 0) {
            if (!sy.sync_or(pr > 0.0f)) { mv = BK_PASS; break; }
            if (q_row && t >= q_vecs) { mv = -4; break; }
        }
        float bv = -1.0f;
        int bi = 0x7fffffff;
        if (active) {
            const float q = q_row ? q_row[(size_t)t * BK_NSQ + p] : bk_exp_draw(seed, game_id, (uint32_t)turn, (uint32_t)t, p);
            bv = __fdiv_rn(pr, q);
            bi = p;
        }
        bk_board_argmax(sy, bv, bi, sc.s_v, sc.s_i, wsq, lane);
        ++t;
        const bool accept = sc.s_ok[bi] != 0;
        if (mode == 1) {
            if (accept) { mv = bi; break; }
            // highest-probability legal move, lowest index on ties
            float fv = (active && ok) ? pr : -1.0f;
            int fi = (active && ok) ? p : 0x7fffffff;
            bk_board_argmax(sy, fv, fi, sc.s_v, sc.s_i, wsq, lane);
            mv = fi == 0x7fffffff ? BK_NONE : fi;
            break;
        }
        if (accept) { mv = bi; break; }
        if (t - 1 >= BK_NSQ) { mv = BK_PASS; break; }   // tries >= 81 (mcts.py:354)
        if (p == bi) pr = 0.0f;
    }

    if (tid == 0 && move_out) *move_out = (int16_t)mv;
    if (mv == BK_NONE || mv == -4) {
        if (tid == 0) done[b] = 1;
        return;
    }

    // lazy liberty cache on the position BEFORE the move (go.py:160); a PASS does not touch it
    if (libs_row && mv >= 0 && active)
        libs_row[p] = (uint8_t)bk_groups_lazy_lib(sc.grp, black, white, last, stale, p, (int)libs_row[p]);   // `stale` was read before the barrier

    // Game.play_move (go.py:123-182) from the table: every thread derives the same outcome
    int st = 0;
    if (mv == BK_PASS) {
        turn += 1; ko = -1; last = BK_PASS;
    } else if (mv == ko) {
        st = 1;
    } else if (bb_test(black, mv) || bb_test(white, mv)) {
        st = 2;
    } else {
        const int pk = bb_possible_ko(black, white, mv);
        BB dead;
        const Cand c = bk_groups_candidate(sc.grp, own, opp, mv, &dead);
        if (c.libs_after == 0) {
            st = 3;
        } else {
            ko = (c.caps == 1 && pk == (blk ? -1 : 1)) ? c.single_cap : -1;
            const BB own2 = bb_or(own, bb_bit(mv)), opp2 = bb_andn(opp, dead);
            black = blk ? own2 : opp2;
            white = blk ? opp2 : own2;
            last = mv;
            turn += 1;
        }
    }
    if (st != 0) {   // cannot happen for a position reached by legal play; flag instead of corrupting state
        if (tid == 0) { done[b] = 1; if (move_out) *move_out = (int16_t)(-10 - st); }
        return;
    }
    if (active) bd[p] = bb_test(black, p) ? 1 : (bb_test(white, p) ? -1 : 0);
    const bool over = mode == 0 ? (turn > max_turn || last == BK_PASS) : (turn > max_turn + 1);
    if (tid == 0) {
        ko_arr[b] = (int16_t)ko; last_arr[b] = (int16_t)last; turn_arr[b] = (int16_t)turn;
        if (over) done[b] = 1;
    }
    if (!conv_base || over) return;          // uniform over the 96 threads

    // ---- re-encode: nnet.features of the new position with the carried cache (what the next policy call sees) -------
    sy.sync();                               // every thread is done with the old table; the cache refresh above is visible
    bk_groups_build(sc.grp, black, white, p);
    const bool stale2 = last >= 0 && libs_row[last] == 0;
    sy.sync();                               // table complete; libs_row[last] has been read by everyone before it is rewritten
    if (active)
        bk_encode_square(sc.grp, black, white, (turn & 1) == 0, ko, last, true, stale2, (int)libs_row[p], p, 0, conv_base,
                         chunk_stride, nullptr, nullptr, nullptr, libs_row);
}
