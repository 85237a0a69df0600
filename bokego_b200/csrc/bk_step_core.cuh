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
    int s_x;                 // one word handed from one thread to the others across a barrier
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
                                              int chunk_stride, long long *stamps = nullptr)
{
    // diagnostic: clock64 of thread 0 after {table, flags, sampling, play + state, new table, planes}
#define BK_STAMP(i) do { if (stamps && tid == 0) stamps[i] = clock64(); } while (0)
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
    BK_STAMP(0);

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
    BK_STAMP(1);

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

    BK_STAMP(2);
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
    BK_STAMP(3);
    if (!conv_base || over) return;          // uniform over the 96 threads

    // ---- re-encode: nnet.features of the new position with the carried cache (what the next policy call sees) -------
    sy.sync();                               // every thread is done with the old table; the cache refresh above is visible
    bk_groups_build(sc.grp, black, white, p);
    const bool stale2 = last >= 0 && libs_row[last] == 0;
    sy.sync();                               // table complete; libs_row[last] has been read by everyone before it is rewritten
    BK_STAMP(4);
    if (active)
        bk_encode_square(sc.grp, black, white, (turn & 1) == 0, ko, last, true, stale2, (int)libs_row[p], p, 0, conv_base,
                         chunk_stride, nullptr, nullptr, nullptr, libs_row);
    BK_STAMP(5);
#undef BK_STAMP
}

// nnet.features of ONE board by 96 threads (the body of bk_encode_kernel): group table into `grp`, then planes / legal moves /
// refreshed liberty cache.  libs_in_row == null: a fresh Game (exact liberties).  Ends with a barrier: `grp` may be reused.
template <class Sync>
__device__ __forceinline__ void bk_encode_board(const Sync &sy, BKGroups &grp, int tid, int b, const int8_t *bd, const int16_t *ko_arr,
                                                const int16_t *last_arr, const int16_t *turn_arr, const uint8_t *libs_in_row,
                                                uint4 *planes, int chunk_stride, uint8_t *legal_row, uint8_t *libs_out_row)
{
    const int lane = tid & 31, p = tid;
    const bool active = p < BK_NSQ;
    const int ko = ko_arr[b], last = last_arr[b], turn = turn_arr[b];
    BB black, white;
    bk_load_boards(bd, lane, black, white);
    const bool carried = libs_in_row != nullptr;
    const bool stale = carried && last >= 0 && libs_in_row[last] == 0;
    const int lib_c = (carried && active) ? (int)libs_in_row[p] : 0;
    bk_groups_build(grp, black, white, p);
    sy.sync();                               // table complete; every thread has read libs_in_row[last] (libs_out may alias libs_in)
    if (active)
        bk_encode_square(grp, black, white, (turn & 1) == 0, ko, last, carried, stale, lib_c, p, 0, planes, chunk_stride, nullptr,
                         nullptr, legal_row, libs_out_row);
    sy.sync();
}

// ---------------------------------------------------------------------------------------------------------------------
// Resident form, for the persistent playout kernel: the position of a board lives in shared memory (bit-boards, ko / last /
// turn / done, the liberty-cache entry of `last`) and in one register per thread (the liberty-cache entry of the thread's own
// square) for all moves of a playout; global memory is read once (bk_resident_load) and written once (bk_resident_store).
// Same results as bk_step_board move by move.  What is off the critical path compared with it: no global round trips, and
// legality is evaluated for the sampled square only (by every thread, from the group table) instead of for all 81 squares.
// ---------------------------------------------------------------------------------------------------------------------
struct BkResident {
    uint32_t black[3], white[3];
    int ko, last, turn, done;
    int lib_last;            // liberty-cache entry of square `last` (the stale test of go.py:226)
};

// Game.is_legal with its early exit (go.py:184-200) and, in the mcts flavour, "does not fill an own eye" (mcts.py:354) for
// ONE square s; uniform when every thread passes the same s
__device__ __forceinline__ bool bk_accept_move(const BKGroups &g, BB black, BB white, bool blk, int ko, int s, int mode)
{
    const BB own = blk ? black : white, opp = blk ? white : black;
    if (bb_test(own, s) || bb_test(opp, s)) return false;
    int nb[4];
    const int n = bb_nbr_list(s, nb);
    const BB occ = bb_or(own, opp);
    int empties = 0;
    bool early = false;
    for (int k = 0; k < n; ++k) {
        if (empties > 1) { early = true; break; }
        if (!bb_test(occ, nb[k])) ++empties;
    }
    bool ok = early || (s != ko && bk_groups_candidate(g, own, opp, s, nullptr).libs_after > 0);
    if (mode == 0 && ok) ok = bb_possible_eye(black, white, s) != (blk ? 1 : -1);
    return ok;
}

// position of board b from global memory; fresh = the caller has no liberty cache (exact liberties, like a fresh Game).  Builds
// the group table, applies the lazy refresh (or takes exact liberties) and writes the planes of the position; my_lib receives the
// cache entry of this thread's square.  All 96 threads.
template <class Sync>
__device__ __forceinline__ void bk_resident_load(const Sync &sy, BkStepScratch &sc, BkResident &st, int tid, int &my_lib, int b,
                                                 const int8_t *bd, const int16_t *ko_arr, const int16_t *last_arr,
                                                 const int16_t *turn_arr, const uint8_t *libs_row, const uint8_t *done, bool fresh,
                                                 uint4 *planes, int chunk_stride)
{
    const int lane = tid & 31, p = tid;
    const bool active = p < BK_NSQ;
    const int ko = ko_arr[b], last = last_arr[b], turn = turn_arr[b], dn = done[b];
    BB black, white;
    bk_load_boards(bd, lane, black, white);
    my_lib = (!fresh && active) ? (int)libs_row[p] : 0;
    const int lib_last = (!fresh && last >= 0) ? (int)libs_row[last] : 0;
    if (!dn) bk_groups_build(sc.grp, black, white, p);
    sy.sync();
    if (!dn && active)
        my_lib = bk_encode_square(sc.grp, black, white, (turn & 1) == 0, ko, last, !fresh, !fresh && last >= 0 && lib_last == 0, my_lib, p,
                                  0, planes, chunk_stride, nullptr, nullptr, nullptr, nullptr);
    if (!dn && last >= 0 && p == last) st.lib_last = my_lib;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { st.black[k] = black.w[k]; st.white[k] = white.w[k]; }
        st.ko = ko; st.last = last; st.turn = turn; st.done = dn;
        if (dn || last < 0) st.lib_last = lib_last;
    }
    sy.sync();                               // the state is published; the scratch may be reused
}

// One move: sample from probs_row, play, refresh the cache, publish the new state and (planes != null) write the planes of the
// new position; with planes == null the cache refresh of the encoder still happens (last move of a run that stops before the
// end of the game), so the cache ends as it does after the stepping kernel.  All 96 threads; uniform control flow.
template <class Sync>
__device__ __forceinline__ void bk_resident_move(const Sync &sy, BkStepScratch &sc, BkResident &st, int tid, int &my_lib,
                                                 const float *probs_row, uint64_t seed, uint32_t game_id, int mode, int max_turn,
                                                 int16_t *move_out, uint4 *planes, int chunk_stride, long long *stamps = nullptr)
{
    // diagnostic: clock64 of thread 0 after {table, sampling, play + publish, new table, planes}
#define BK_STAMP(i) do { if (stamps && tid == 0) stamps[i] = clock64(); } while (0)
    const int lane = tid & 31, wsq = tid >> 5, p = tid;
    const bool active = p < BK_NSQ;
    if (st.done) {
        if (tid == 0 && move_out) *move_out = -3;
        return;
    }
    BB black, white;
#pragma unroll
    for (int k = 0; k < 3; ++k) { black.w[k] = st.black[k]; white.w[k] = st.white[k]; }
    int ko = st.ko, last = st.last, turn = st.turn;
    const bool stale = last >= 0 && st.lib_last == 0;
    const bool blk = (turn & 1) == 0;
    const BB own = blk ? black : white, opp = blk ? white : black;
    bk_groups_build(sc.grp, black, white, p);
    float pr = active ? probs_row[p] : 0.0f;
    sy.sync();
    BK_STAMP(1);

    int mv = BK_NONE;
    int t = 0;
    for (;;) {
        if (t > 0 && !sy.sync_or(pr > 0.0f)) { mv = BK_PASS; break; }
        float bv = -1.0f;
        int bi = 0x7fffffff;
        if (active) {
            bv = __fdiv_rn(pr, bk_exp_draw(seed, game_id, (uint32_t)turn, (uint32_t)t, p));
            bi = p;
        }
        bk_board_argmax(sy, bv, bi, sc.s_v, sc.s_i, wsq, lane);
        ++t;
        if (bk_accept_move(sc.grp, black, white, blk, ko, bi, mode)) { mv = bi; break; }
        if (mode == 1) {
            // highest-probability legal move, lowest index on ties (selfplay.py:38-47): now every square needs its flag
            const bool ok = active && bk_accept_move(sc.grp, black, white, blk, ko, p, mode);
            float fv = ok ? pr : -1.0f;
            int fi = ok ? p : 0x7fffffff;
            bk_board_argmax(sy, fv, fi, sc.s_v, sc.s_i, wsq, lane);
            mv = fi == 0x7fffffff ? BK_NONE : fi;
            break;
        }
        if (t - 1 >= BK_NSQ) { mv = BK_PASS; break; }   // tries >= 81 (mcts.py:354)
        if (p == bi) pr = 0.0f;
    }
    BK_STAMP(2);
    if (tid == 0 && move_out) *move_out = (int16_t)mv;
    if (mv == BK_NONE) {
        if (tid == 0) st.done = 1;
        return;
    }
    // lazy liberty cache on the position BEFORE the move (go.py:160); a PASS does not touch it
    if (mv >= 0 && active) my_lib = bk_groups_lazy_lib(sc.grp, black, white, last, stale, p, my_lib);
    int err = 0;
    if (mv == BK_PASS) {
        turn += 1; ko = -1; last = BK_PASS;
    } else {
        const int pk = bb_possible_ko(black, white, mv);
        BB dead;
        const Cand c = bk_groups_candidate(sc.grp, own, opp, mv, &dead);
        if (mv == ko || bb_test(black, mv) || bb_test(white, mv) || c.libs_after == 0) {
            err = 1;         // cannot happen for an accepted move; flag instead of corrupting state
        } else {
            ko = (c.caps == 1 && pk == (blk ? -1 : 1)) ? c.single_cap : -1;
            const BB own2 = bb_or(own, bb_bit(mv)), opp2 = bb_andn(opp, dead);
            black = blk ? own2 : opp2;
            white = blk ? opp2 : own2;
            last = mv;
            turn += 1;
        }
    }
    if (err) {
        if (tid == 0) { st.done = 1; if (move_out) *move_out = (int16_t)-13; }
        return;
    }
    const bool over = mode == 0 ? (turn > max_turn || last == BK_PASS) : (turn > max_turn + 1);
    if (last >= 0 && p == last) sc.s_x = my_lib;             // cache entry of the square just played (an empty square keeps its stale value)
    sy.sync();                               // every thread is done with the old table and the old state
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { st.black[k] = black.w[k]; st.white[k] = white.w[k]; }
        st.ko = ko; st.last = last; st.turn = turn; st.done = over ? 1 : 0;
    }
    BK_STAMP(3);
    if (over) {
        if (last >= 0 && p == last) st.lib_last = my_lib;
        return;
    }
    // ---- nnet.features of the new position with the carried cache (what the next policy call sees) ----------------------
    const bool stale2 = last >= 0 && sc.s_x == 0;
    bk_groups_build(sc.grp, black, white, p);
    sy.sync();
    BK_STAMP(4);
    if (active)
        my_lib = bk_encode_square(sc.grp, black, white, (turn & 1) == 0, ko, last, true, stale2, my_lib, p, 0, planes, chunk_stride,
                                  nullptr, nullptr, nullptr, nullptr);
    if (last >= 0 && p == last) st.lib_last = my_lib;
    BK_STAMP(5);
#undef BK_STAMP
}

// the position back to global memory, in the encoding of the stepping kernel
__device__ __forceinline__ void bk_resident_store(const BkResident &st, int tid, int my_lib, int b, int8_t *bd, int16_t *ko_arr,
                                                  int16_t *last_arr, int16_t *turn_arr, uint8_t *libs_row, uint8_t *done)
{
    BB black, white;
#pragma unroll
    for (int k = 0; k < 3; ++k) { black.w[k] = st.black[k]; white.w[k] = st.white[k]; }
    if (tid < BK_NSQ) {
        bd[tid] = bb_test(black, tid) ? 1 : (bb_test(white, tid) ? -1 : 0);
        libs_row[tid] = (uint8_t)my_lib;
    }
    if (tid == 0) {
        ko_arr[b] = (int16_t)st.ko; last_arr[b] = (int16_t)st.last; turn_arr[b] = (int16_t)st.turn; done[b] = (uint8_t)(st.done != 0);
    }
}

