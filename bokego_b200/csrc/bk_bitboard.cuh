// bk_bitboard.cuh -- 9x9 Go rules on 81-bit bit-boards (three 27-bit words, one word = three rows).
//
// Everything here is __host__ __device__ so the same code is unit-tested on the CPU
// (tests/test_bitboard_host.py builds it with g++) before it ever runs in a kernel.
// Semantics follow the reference, quirks included; each routine cites what it replaces:
//   bb_flood / bb_libs        go.py:375-402  flood_fill, get_stone_lib
//   bb_candidate              nnet.py:241-247 + go.py:404-418 (capture list counts a group once per
//                             touching neighbour) + suicide test of go.py:154-157
//   bb_is_legal_quirk         go.py:184-200 (early exit before the ko test)
//   bb_possible_ko / _eye     go.py:461-485 with the DIAGONALS table of go.py:372-373 (typo kept, F6)
//   bb_play                   go.py:123-182 / 109-121
//   bb_score                  go.py:202-218 (ascending regions, border stones repainted, F6)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BK_HD __host__ __device__ __forceinline__
#else
#define BK_HD inline
#endif

#define BK_NSQ 81
#define BK_PASS (-1)
#define BK_NONE (-2)
#define BK_M27 0x07FFFFFFu
#define BK_COL0 0x00040201u   // bits 0, 9, 18  (y == 0 of each row in a word)
#define BK_COL8 0x04020100u   // bits 8, 17, 26 (y == 8)

struct BB { uint32_t w[3]; };

BK_HD int bk_popc(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
BK_HD int bk_ffs(uint32_t v)   // index of lowest set bit, v != 0
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

BK_HD BB bb_zero() { BB r; r.w[0] = r.w[1] = r.w[2] = 0u; return r; }
BK_HD BB bb_and(BB a, BB b) { BB r; r.w[0] = a.w[0] & b.w[0]; r.w[1] = a.w[1] & b.w[1]; r.w[2] = a.w[2] & b.w[2]; return r; }
BK_HD BB bb_or(BB a, BB b) { BB r; r.w[0] = a.w[0] | b.w[0]; r.w[1] = a.w[1] | b.w[1]; r.w[2] = a.w[2] | b.w[2]; return r; }
BK_HD BB bb_andn(BB a, BB b) { BB r; r.w[0] = a.w[0] & ~b.w[0]; r.w[1] = a.w[1] & ~b.w[1]; r.w[2] = a.w[2] & ~b.w[2]; return r; }
BK_HD bool bb_any(BB a) { return (a.w[0] | a.w[1] | a.w[2]) != 0u; }
BK_HD bool bb_eq(BB a, BB b) { return a.w[0] == b.w[0] && a.w[1] == b.w[1] && a.w[2] == b.w[2]; }
BK_HD int bb_count(BB a) { return bk_popc(a.w[0]) + bk_popc(a.w[1]) + bk_popc(a.w[2]); }
BK_HD bool bb_test(BB a, int s)
{
    // s / 27 without a division: s < 81
    int wi = (s >= 54) ? 2 : (s >= 27 ? 1 : 0);
    uint32_t word = wi == 0 ? a.w[0] : (wi == 1 ? a.w[1] : a.w[2]);
    return (word >> (s - 27 * wi)) & 1u;
}
BK_HD BB bb_bit(int s)
{
    BB r = bb_zero();
    int wi = (s >= 54) ? 2 : (s >= 27 ? 1 : 0);
    uint32_t m = 1u << (s - 27 * wi);
    if (wi == 0) r.w[0] = m; else if (wi == 1) r.w[1] = m; else r.w[2] = m;
    return r;
}
BK_HD int bb_first(BB a)   // lowest square index, a non-empty
{
    if (a.w[0]) return bk_ffs(a.w[0]);
    if (a.w[1]) return 27 + bk_ffs(a.w[1]);
    return 54 + bk_ffs(a.w[2]);
}

// the four orthogonal neighbours of every set square (self not included)
BK_HD BB bb_neighbours(BB a)
{
    BB r;
    r.w[0] = ((a.w[0] << 1) & ~BK_COL0 & BK_M27) | ((a.w[0] >> 1) & ~BK_COL8) | ((a.w[0] << 9) & BK_M27) |
             (a.w[0] >> 9) | ((a.w[1] << 18) & BK_M27);
    r.w[1] = ((a.w[1] << 1) & ~BK_COL0 & BK_M27) | ((a.w[1] >> 1) & ~BK_COL8) | ((a.w[1] << 9) & BK_M27) |
             (a.w[1] >> 9) | (a.w[0] >> 18) | ((a.w[2] << 18) & BK_M27);
    r.w[2] = ((a.w[2] << 1) & ~BK_COL0 & BK_M27) | ((a.w[2] >> 1) & ~BK_COL8) | ((a.w[2] << 9) & BK_M27) |
             (a.w[2] >> 9) | (a.w[1] >> 18);
    return r;
}

// connected component of `seed` inside `mask` (seed must lie inside mask)
BK_HD BB bb_flood(BB seed, BB mask)
{
    BB g = seed;
    for (;;) {
        BB n = bb_or(g, bb_and(bb_neighbours(g), mask));
        if (bb_eq(n, g)) return g;
        g = n;
    }
}

// liberties of a group = empty squares next to it
BK_HD BB bb_libs(BB grp, BB empty) { return bb_and(bb_neighbours(grp), empty); }

struct Cand { int libs_after; int caps; int single_cap; };

// Put a stone of the side owning `own` on the empty square s.  caps = length of the reference's
// capture list (a dead group is listed once per neighbour of s that belongs to it); libs_after =
// liberties of the new stone's group once the captured stones are gone (0 => suicide);
// single_cap = the captured square when the list has exactly one entry (ko candidate), else -1.
// `after_own/after_opp` (optional) receive the position after the move.
BK_HD Cand bb_candidate(BB own, BB opp, int s, BB *after_own, BB *after_opp)
{
    BB sb = bb_bit(s);
    BB all; all.w[0] = BK_M27; all.w[1] = BK_M27; all.w[2] = BK_M27;
    BB empty_after = bb_andn(bb_andn(bb_andn(all, own), opp), sb);
    BB nb = bb_neighbours(sb);
    BB opp_nb = bb_and(nb, opp);
    BB dead = bb_zero();
    Cand c; c.caps = 0; c.single_cap = -1;
    while (bb_any(opp_nb)) {
        int v = bb_first(opp_nb);
        BB vb = bb_bit(v);
        opp_nb = bb_andn(opp_nb, vb);
        BB g = bb_flood(vb, opp);
        if (!bb_any(bb_libs(g, empty_after))) {
            int n = bb_count(g);
            if (c.caps == 0 && n == 1) c.single_cap = v;
            c.caps += n;
            dead = bb_or(dead, g);
        }
    }
    if (c.caps != 1) c.single_cap = -1;
    BB own2 = bb_or(own, sb);
    BB mine = bb_flood(sb, own2);
    c.libs_after = bb_count(bb_libs(mine, bb_or(empty_after, dead)));
    if (after_own) *after_own = own2;
    if (after_opp) *after_opp = bb_andn(opp, dead);
    return c;
}

// neighbours of s in the reference's NEIGHBORS order (x+1, x-1, y+1, y-1), off-board skipped
BK_HD int bb_nbr_list(int s, int out[4])
{
    int x = s / 9, y = s - 9 * x, n = 0;
    if (x < 8) out[n++] = s + 9;
    if (x > 0) out[n++] = s - 9;
    if (y < 8) out[n++] = s + 1;
    if (y > 0) out[n++] = s - 1;
    return n;
}

// go.py:461-468: colour surrounding the empty square s (+1 black, -1 white), 0 if none/mixed
BK_HD int bb_possible_ko(BB black, BB white, int s)
{
    if (bb_test(black, s) || bb_test(white, s)) return 0;
    BB nb = bb_neighbours(bb_bit(s));
    if (bb_eq(bb_and(nb, black), nb)) return 1;
    if (bb_eq(bb_and(nb, white), nb)) return -1;
    return 0;
}

// go.py:470-485 with DIAGONALS = [(x+1,y+1),(x+1,y-1),(x-1,y-1),(x-1,y-1)] filtered on-board
BK_HD int bb_possible_eye(BB black, BB white, int s)
{
    int c = bb_possible_ko(black, white, s);
    if (c == 0) return 0;
    BB other = c == 1 ? white : black;
    int x = s / 9, y = s - 9 * x;
    int n_diag = 0, faults = 0;
    if (x < 8 && y < 8) { ++n_diag; if (bb_test(other, s + 10)) ++faults; }
    if (x < 8 && y > 0) { ++n_diag; if (bb_test(other, s + 8)) ++faults; }
    if (x > 0 && y > 0) { n_diag += 2; if (bb_test(other, s - 10)) faults += 2; }
    if (n_diag < 4) ++faults;
    return faults > 1 ? 0 : c;
}

// go.py:184-200: `own` = stones of the side to move.  The neighbour scan returns True as soon as
// two empty neighbours have been seen BEFORE the current one -- without looking at ko.
BK_HD bool bb_is_legal_quirk(BB own, BB opp, int ko, int s)
{
    if (bb_test(own, s) || bb_test(opp, s)) return false;
    int nb[4];
    int n = bb_nbr_list(s, nb), empties = 0;
    BB occ = bb_or(own, opp);
    for (int k = 0; k < n; ++k) {
        if (empties > 1) return true;
        if (!bb_test(occ, nb[k])) ++empties;
    }
    if (s == ko) return false;
    return bb_candidate(own, opp, s, 0, 0).libs_after > 0;
}

// go.py:245-260 + 184-200 as reached from get_legal_moves: an empty square with an empty neighbour
// is always listed; an isolated one goes through is_legal (ko and suicide tests).
BK_HD bool bb_listed_legal(BB own, BB opp, int ko, int s, const Cand &c)
{
    BB all; all.w[0] = BK_M27; all.w[1] = BK_M27; all.w[2] = BK_M27;
    BB empty = bb_andn(bb_andn(all, own), opp);
    if (bb_any(bb_and(bb_neighbours(bb_bit(s)), empty))) return true;
    return s != ko && c.libs_after > 0;
}

// Game.play_move / play_pass on (black, white, ko, last, turn).  Returns 0 ok, 1 ko, 2 not_empty,
// 3 suicide; the state is untouched on error.  (The liberty cache is handled by the caller.)
BK_HD int bb_play(BB &black, BB &white, int &ko, int &last, int &turn, int mv)
{
    if (mv == BK_PASS) { turn += 1; ko = -1; last = BK_PASS; return 0; }
    if (mv == ko) return 1;
    if (bb_test(black, mv) || bb_test(white, mv)) return 2;
    bool blk = (turn & 1) == 0;
    int pk = bb_possible_ko(black, white, mv);
    BB own2, opp2;
    Cand c = bb_candidate(blk ? black : white, blk ? white : black, mv, &own2, &opp2);
    if (c.libs_after == 0) return 3;
    ko = (c.caps == 1 && pk == (blk ? -1 : 1)) ? c.single_cap : -1;
    black = blk ? own2 : opp2;
    white = blk ? opp2 : own2;
    last = mv;
    turn += 1;
    return 0;
}

// Game.score(): returns #X - #O after the reference's region painting; caller subtracts komi.
BK_HD int bb_score_diff(BB black, BB white)
{
    BB all; all.w[0] = BK_M27; all.w[1] = BK_M27; all.w[2] = BK_M27;
    BB empty = bb_andn(bb_andn(all, black), white);
    while (bb_any(empty)) {
        BB region = bb_flood(bb_bit(bb_first(empty)), empty);
        BB border = bb_andn(bb_neighbours(region), region);
        bool hx = bb_any(bb_and(border, black)), ho = bb_any(bb_and(border, white));
        BB paint = bb_or(region, border);
        black = bb_andn(black, paint);
        white = bb_andn(white, paint);
        if (hx && !ho) black = bb_or(black, paint);
        else if (ho && !hx) white = bb_or(white, paint);
        empty = bb_andn(empty, region);
    }
    return bb_count(black) - bb_count(white);
}

// Lazy liberty cache (go.py:220-243) for ONE square p given the carried value: returns the new
// value of libs[p].  `last_stale` = (last >= 0 && libs_in[last] == 0).
BK_HD int bb_lazy_lib_of(BB black, BB white, int last, bool last_stale, int p, int carried)
{
    if (!last_stale) return carried;
    bool pb = bb_test(black, p), pw = bb_test(white, p);
    if (!pb && !pw) return carried;
    BB all; all.w[0] = BK_M27; all.w[1] = BK_M27; all.w[2] = BK_M27;
    BB empty = bb_andn(bb_andn(all, black), white);
    BB g = bb_flood(bb_bit(p), pb ? black : white);
    BB seeds = bb_or(bb_neighbours(bb_bit(last)), bb_bit(last));
    if (!bb_any(bb_and(g, seeds))) return carried;
    return bb_count(bb_libs(g, empty));
}

// exact liberties of the stone at p (fresh Game), 0 on empty squares
BK_HD int bb_exact_lib_of(BB black, BB white, int p)
{
    bool pb = bb_test(black, p), pw = bb_test(white, p);
    if (!pb && !pw) return 0;
    BB all; all.w[0] = BK_M27; all.w[1] = BK_M27; all.w[2] = BK_M27;
    BB empty = bb_andn(bb_andn(all, black), white);
    return bb_count(bb_libs(bb_flood(bb_bit(p), pb ? black : white), empty));
}

// ---- counter-based Exp(1) stream (bit-identical to oracle/bk_oracle.c: bko_exp_draws) ------------
BK_HD uint32_t bk_mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

BK_HD void bk_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4])
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = bk_mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = bk_mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

BK_HD float bk_fma(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}
BK_HD float bk_mul(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}

// q = -log(u), u = (k + 0.5) * 2^-23 from the top 23 bits; only correctly-rounded single operations
// in a fixed order, so host and device agree bit for bit.
BK_HD float bk_exp_from_bits(uint32_t bits)
{
    float u = bk_mul((float)(bits >> 9) + 0.5f, 1.1920928955078125e-07f);
    union { float f; uint32_t i; } cv;
    cv.f = u;
    int e = (int)(cv.i >> 23) - 127;
    cv.i = (cv.i & 0x007FFFFFu) | 0x3F800000u;
    float m = cv.f;
    if (m > 1.41421356f) { m = bk_mul(m, 0.5f); e += 1; }
    float t = m - 1.0f;
    float p = -0.0833333333f;
    p = bk_fma(p, t, 0.0909090909f);
    p = bk_fma(p, t, -0.1f);
    p = bk_fma(p, t, 0.1111111111f);
    p = bk_fma(p, t, -0.125f);
    p = bk_fma(p, t, 0.1428571429f);
    p = bk_fma(p, t, -0.1666666667f);
    p = bk_fma(p, t, 0.2f);
    p = bk_fma(p, t, -0.25f);
    p = bk_fma(p, t, 0.3333333333f);
    p = bk_fma(p, t, -0.5f);
    p = bk_fma(p, t, 1.0f);
    float lg = bk_fma((float)e, 0.693147180559945f, bk_mul(p, t));
    return -lg;
}

// Exp(1) variate for square i of draw (seed, game, move, try)
BK_HD float bk_exp_draw(uint64_t seed, uint32_t game, uint32_t move, uint32_t tr, int i)
{
    uint32_t r[4];
    bk_philox(seed, game, move, tr, (uint32_t)(i >> 2), r);
    return bk_exp_from_bits(r[i & 3]);
}

// ---- group table of one board (shared memory of a block of 96 threads, thread t = square t) -----------------------
// Every stone floods its own group once; the group's lowest square is its id and publishes the group's stones and
// liberties.  "What if the side to move puts a stone on the empty square s" (bb_candidate) then needs no flood at all:
// the <= 4 neighbouring groups are looked up.  Device only.
#if defined(__CUDACC__)
struct BKGroups {
    uint32_t stones[BK_NSQ][3];
    uint32_t libs[BK_NSQ][3];
    uint8_t root[BK_NSQ];
};

// call with all threads of the block (p = square of this thread, p >= 81 for the idle ones), then __syncthreads()
__device__ __forceinline__ void bk_groups_build(BKGroups &g, BB black, BB white, int p)
{
    if (p >= BK_NSQ) return;
    const bool pb = bb_test(black, p), pw = bb_test(white, p);
    if (!pb && !pw) return;
    BB all; all.w[0] = BK_M27; all.w[1] = BK_M27; all.w[2] = BK_M27;
    const BB empty = bb_andn(bb_andn(all, black), white);
    const BB grp = bb_flood(bb_bit(p), pb ? black : white);
    const int r = bb_first(grp);
    g.root[p] = (uint8_t)r;
    if (r == p) {
        const BB l = bb_libs(grp, empty);
#pragma unroll
        for (int k = 0; k < 3; ++k) { g.stones[p][k] = grp.w[k]; g.libs[p][k] = l.w[k]; }
    }
}
__device__ __forceinline__ BB bk_group_stones(const BKGroups &g, int q)
{
    const int r = g.root[q];
    BB b; b.w[0] = g.stones[r][0]; b.w[1] = g.stones[r][1]; b.w[2] = g.stones[r][2];
    return b;
}
__device__ __forceinline__ BB bk_group_libs(const BKGroups &g, int q)
{
    const int r = g.root[q];
    BB b; b.w[0] = g.libs[r][0]; b.w[1] = g.libs[r][1]; b.w[2] = g.libs[r][2];
    return b;
}
// bb_candidate(own, opp, s, ...) from the table; `dead` receives the captured stones
__device__ __forceinline__ Cand bk_groups_candidate(const BKGroups &g, BB own, BB opp, int s, BB *dead_out)
{
    BB all; all.w[0] = BK_M27; all.w[1] = BK_M27; all.w[2] = BK_M27;
    const BB sb = bb_bit(s);
    const BB empty_after = bb_andn(bb_andn(bb_andn(all, own), opp), sb);
    BB dead = bb_zero(), merged = sb;
    Cand c; c.caps = 0; c.single_cap = -1;
    int nb[4];
    const int nn = bb_nbr_list(s, nb);
    for (int k = 0; k < nn; ++k) {
        const int q = nb[k];
        const bool qo = bb_test(opp, q), qm = bb_test(own, q);
        if (!qo && !qm) continue;
        const BB grp = bk_group_stones(g, q);
        if (qm) { merged = bb_or(merged, grp); continue; }
        if (!bb_any(bb_and(bk_group_libs(g, q), empty_after))) {          // its only liberty was s
            const int n = bb_count(grp);
            if (c.caps == 0 && n == 1) c.single_cap = q;
            c.caps += n;
            dead = bb_or(dead, grp);
        }
    }
    if (c.caps != 1) c.single_cap = -1;
    c.libs_after = bb_count(bb_libs(merged, bb_or(empty_after, dead)));
    if (dead_out) *dead_out = dead;
    return c;
}
// the lazy liberty cache for the stone / empty square p (bb_lazy_lib_of) from the table
__device__ __forceinline__ int bk_groups_lazy_lib(const BKGroups &g, BB black, BB white, int last, bool last_stale, int p, int carried)
{
    if (!last_stale) return carried;
    if (!bb_test(black, p) && !bb_test(white, p)) return carried;
    const BB seeds = bb_or(bb_neighbours(bb_bit(last)), bb_bit(last));
    if (!bb_any(bb_and(bk_group_stones(g, p), seeds))) return carried;
    return bb_count(bk_group_libs(g, p));
}
#endif
