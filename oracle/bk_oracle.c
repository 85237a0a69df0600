/*
 * bk_oracle.c -- CPU restatement of BokeGo's integer hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the *checker* for the CUDA kernels in bokego_b200/csrc.  Only tests/, the
 * smoke() entry and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * path never does.  It restates, in plain C with simple array walks (deliberately NOT the
 * bit-board formulation the kernels use, so the two are independent), the algorithms of the
 * reference at /root/reference:
 *
 *   go.py:375-390  flood_fill            -> grp_fill()
 *   go.py:392-402  get_stone_lib         -> grp_libs()
 *   go.py:404-418  get_caps              -> captures_of() (duplicate counting)
 *   go.py:123-182  Game.play_move        -> bko_play()
 *   go.py:109-121  Game.play_pass        -> bko_play(move = -1)
 *   go.py:184-200  Game.is_legal         -> bko_is_legal() (early-exit quirk kept)
 *   go.py:245-260  Game.get_legal_moves  -> legal part of bko_features()
 *   go.py:220-243  Game.get_liberties    -> lazy_libs() (stale cache semantics)
 *   go.py:202-218  Game.score            -> bko_score() (order dependent, border overwrite, F6)
 *   go.py:461-485  possible_ko/eye       -> bko_possible_ko()/bko_possible_eye() (DIAGONALS typo, F6)
 *   nnet.py:182-262 features             -> bko_features()
 *   mcts.py:348-364 get_move/make_move/is_game_over -> bko_step_mcts() (+ F7 shim: no mass => PASS)
 *   bin/selfplay.py:35-47 legal_sample   -> bko_step_selfplay()
 *   torch.multinomial (1 sample)         -> race_argmax(): argmax_i p_i / q_i, first max wins
 *
 * Parity pin: tests/test_oracle_golden.py checks every function here against vectors produced by
 * running the unmodified reference in the build container (tests/golden/make_golden.py).
 *
 * Encoding: board int8 [81], index 9*x+y; +1 = BLACK 'X', -1 = WHITE 'O', 0 = EMPTY.
 *           ko: -1 = None.  last: -1 = PASS, -2 = None.  turn even => BLACK to move.
 */
#include <stdint.h>
#include <string.h>
#include <math.h>

#define NSQ 81
#define BK_NONE (-2)
#define BK_PASS (-1)

/* ---- static geometry, same order as go.py:370-373 ------------------------------------------- */
static int g_ready = 0;
static int8_t NB[NSQ][4];  static int NBN[NSQ];
static int8_t DG[NSQ][4];  static int DGN[NSQ];

static void geom_init(void)
{
    if (g_ready) return;
    static const int nd[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
    /* go.py:372-373: the last diagonal offset repeats (-1,-1); (-1,+1) never appears. */
    static const int dd[4][2] = {{1, 1}, {1, -1}, {-1, -1}, {-1, -1}};
    for (int x = 0; x < 9; ++x)
        for (int y = 0; y < 9; ++y) {
            int s = 9 * x + y, n = 0, d = 0;
            for (int k = 0; k < 4; ++k) {
                int a = x + nd[k][0], b = y + nd[k][1];
                if (a >= 0 && a < 9 && b >= 0 && b < 9) NB[s][n++] = (int8_t)(9 * a + b);
            }
            NBN[s] = n;
            for (int k = 0; k < 4; ++k) {
                int a = x + dd[k][0], b = y + dd[k][1];
                if (a >= 0 && a < 9 && b >= 0 && b < 9) DG[s][d++] = (int8_t)(9 * a + b);
            }
            DGN[s] = d;
        }
    g_ready = 1;
}

/* ---- group primitives ------------------------------------------------------------------------ */
/* Fill the 4-connected component of `s` (same colour code as board[s]); returns its size, writes
 * members to grp[] and marks them in in_grp[].  border[] gets 1 on every non-member neighbour. */
static int grp_fill(const int8_t *bd, int s, uint8_t *in_grp, int8_t *grp, uint8_t *border)
{
    int8_t col = bd[s];
    int8_t stack[NSQ];
    int sp = 0, n = 0;
    memset(in_grp, 0, NSQ);
    if (border) memset(border, 0, NSQ);
    stack[sp++] = (int8_t)s; in_grp[s] = 1;
    while (sp) {
        int c = stack[--sp];
        grp[n++] = (int8_t)c;
        for (int k = 0; k < NBN[c]; ++k) {
            int v = NB[c][k];
            if (bd[v] == col) { if (!in_grp[v]) { in_grp[v] = 1; stack[sp++] = (int8_t)v; } }
            else if (border) border[v] = 1;
        }
    }
    return n;
}

/* liberties (number of EMPTY border squares) of the group at s; optionally the group */
static int grp_libs(const int8_t *bd, int s, uint8_t *in_grp, int8_t *grp, int *n_grp)
{
    uint8_t border[NSQ];
    int n = grp_fill(bd, s, in_grp, grp, border), libs = 0;
    for (int i = 0; i < NSQ; ++i) if (border[i] && bd[i] == 0) ++libs;
    if (n_grp) *n_grp = n;
    return libs;
}

/* go.py:404-418 on a board that already has `col` on s.  For EACH opponent-coloured neighbour of s
 * (NEIGHBORS order) whose group has no liberty, the group's stones are appended to the capture
 * list -- so a group touching s through k neighbours is listed k times.  Returns the list length,
 * removes the stones from `bd`, reports the first listed stone. */
static int captures_of(int8_t *bd, int s, int col, int *first)
{
    int8_t work[NSQ];
    uint8_t in_grp[NSQ], kill[NSQ];
    int8_t grp[NSQ];
    int total = 0, n;
    memcpy(work, bd, NSQ);
    memset(kill, 0, NSQ);
    if (first) *first = -1;
    for (int k = 0; k < NBN[s]; ++k) {
        int v = NB[s][k];
        if (work[v] != -col) continue;
        if (grp_libs(work, v, in_grp, grp, &n) == 0) {
            if (total == 0 && first) *first = v;   /* only consulted when the list has one stone */
            for (int i = 0; i < n; ++i) kill[grp[i]] = 1;
            total += n;
        }
    }
    for (int i = 0; i < NSQ; ++i) if (kill[i]) bd[i] = 0;
    return total;
}

int bko_possible_ko(const int8_t *bd, int s)
{
    geom_init();
    if (bd[s] != 0) return 0;
    int c = bd[NB[s][0]];
    if (c == 0) return 0;
    for (int k = 1; k < NBN[s]; ++k) if (bd[NB[s][k]] != c) return 0;
    return c;
}

int bko_possible_eye(const int8_t *bd, int s)
{
    int c = bko_possible_ko(bd, s);
    if (!c) return 0;
    int faults = DGN[s] < 4 ? 1 : 0;
    for (int k = 0; k < DGN[s]; ++k) { int v = bd[DG[s][k]]; if (v != c && v != 0) ++faults; }
    return faults > 1 ? 0 : c;
}

/* go.py:220-243.  libs==NULL is not allowed here; `fresh` means Game._libs was None. */
static void lazy_libs(const int8_t *bd, int last, uint8_t *libs, int fresh)
{
    uint8_t in_grp[NSQ], seen[NSQ];
    int8_t grp[NSQ];
    int n;
    if (fresh) {
        memset(libs, 0, NSQ);
        memset(seen, 0, NSQ);
        for (int s = 0; s < NSQ; ++s) {
            if (bd[s] == 0 || seen[s]) continue;
            int l = grp_libs(bd, s, in_grp, grp, &n);
            for (int i = 0; i < n; ++i) { libs[grp[i]] = (uint8_t)l; seen[grp[i]] = 1; }
        }
        return;
    }
    if (last >= 0 && libs[last] == 0) {
        memset(seen, 0, NSQ);
        for (int k = 0; k <= NBN[last]; ++k) {
            int v = k < NBN[last] ? NB[last][k] : last;
            if (bd[v] == 0 || seen[v]) continue;
            int l = grp_libs(bd, v, in_grp, grp, &n);
            for (int i = 0; i < n; ++i) { libs[grp[i]] = (uint8_t)l; seen[grp[i]] = 1; }
        }
    }
}

/* Place `col` on empty s, remove captures; returns own-group liberties afterwards, *caps = length
 * of the (duplicated) capture list, *first_cap = its first element. */
static int try_place(const int8_t *bd, int s, int col, int8_t *out, int *caps, int *first_cap)
{
    int8_t nb[NSQ];
    uint8_t in_grp[NSQ];
    int8_t grp[NSQ];
    memcpy(nb, bd, NSQ);
    nb[s] = (int8_t)col;
    *caps = captures_of(nb, s, col, first_cap);
    int l = grp_libs(nb, s, in_grp, grp, 0);
    if (out) memcpy(out, nb, NSQ);
    return l;
}

/* go.py:184-200 including the early exit that skips the ko test. */
int bko_is_legal(const int8_t *bd, int ko, int turn, int s)
{
    geom_init();
    if (s == BK_PASS) return 1;
    if (bd[s] != 0) return 0;
    int empties = 0;
    for (int k = 0; k < NBN[s]; ++k) {
        if (empties > 1) return 1;
        if (bd[NB[s][k]] == 0) ++empties;
    }
    if (s == ko) return 0;
    int col = (turn & 1) ? -1 : 1, caps, fc;
    return try_place(bd, s, col, 0, &caps, &fc) > 0;
}

/* play_move / play_pass.  Returns 0 ok, 1 ko, 2 not_empty, 3 suicide (state untouched on error).
 * libs may be NULL (liberty cache not tracked); libs_fresh!=0 means Game._libs is None on entry. */
int bko_play(int8_t *bd, int16_t *ko, int16_t *last, int16_t *turn, uint8_t *libs, int libs_fresh, int mv)
{
    geom_init();
    if (mv == BK_PASS) { *turn += 1; *ko = -1; *last = BK_PASS; return 0; }
    if (mv == *ko) return 1;
    if (bd[mv] != 0) return 2;
    int col = (*turn & 1) ? -1 : 1;
    int pk = bko_possible_ko(bd, mv);
    int8_t nb[NSQ];
    int caps, fc;
    int l = try_place(bd, mv, col, nb, &caps, &fc);
    if (l == 0) return 3;
    int new_ko = (caps == 1 && pk == -col) ? fc : -1;
    if (libs) lazy_libs(bd, *last, libs, libs_fresh);   /* go.py:160, BEFORE the board changes */
    memcpy(bd, nb, NSQ);
    *last = (int16_t)mv; *ko = (int16_t)new_ko; *turn += 1;
    return 0;
}

/* nnet.py:182-262.  feats: uint8 [27][81] holding the plane VALUES (0..7).  libs_in==NULL => fresh
 * Game (exact liberties); otherwise the carried cache, updated lazily.  legal/libs_out may be NULL. */
void bko_features(const int8_t *bd, int ko, int last, int turn, const uint8_t *libs_in,
                  uint8_t *feats, uint8_t *legal_out, uint8_t *libs_out)
{
    geom_init();
    int col = (turn & 1) ? -1 : 1;
    uint8_t libs[NSQ], legal[NSQ];
    int la[NSQ], cp[NSQ];
    memset(feats, 0, 27 * NSQ);
    if (libs_in) memcpy(libs, libs_in, NSQ);
    /* legal moves first (go.py:245-260), then liberties (nnet.py:233-237) */
    for (int s = 0; s < NSQ; ++s) {
        legal[s] = 0; la[s] = 0; cp[s] = 0;
        if (bd[s] != 0) continue;
        int has_empty_nb = 0;
        for (int k = 0; k < NBN[s]; ++k) if (bd[NB[s][k]] == 0) has_empty_nb = 1;
        int caps, fc;
        int l = try_place(bd, s, col, 0, &caps, &fc);
        int ok = has_empty_nb ? 1 : (s != ko && l > 0);
        if (ok) { legal[s] = 1; la[s] = l; cp[s] = caps; }
    }
    lazy_libs(bd, last, libs, libs_in == 0);
    for (int s = 0; s < NSQ; ++s) {
        if (bd[s] == col) feats[0 * NSQ + s] = 1;
        if (bd[s] == -col) feats[1 * NSQ + s] = 1;
        if (bd[s] == 0) feats[2 * NSQ + s] = 1;
        if (col == 1) feats[3 * NSQ + s] = 1;
        if (legal[s]) feats[5 * NSQ + s] = 1;
        int v = libs[s];
        if (v >= 1) feats[(6 + (v > 6 ? 6 : v - 1)) * NSQ + s] = (uint8_t)(v > 6 ? 7 : v);
        v = la[s];
        if (v >= 1) feats[(13 + (v > 6 ? 6 : v - 1)) * NSQ + s] = (uint8_t)(v > 6 ? 7 : v);
        v = cp[s];
        if (v >= 1) feats[(20 + (v > 6 ? 6 : v - 1)) * NSQ + s] = (uint8_t)(v > 6 ? 7 : v);
    }
    if (last >= 0) feats[4 * NSQ + last] = 1;
    if (legal_out) memcpy(legal_out, legal, NSQ);
    if (libs_out) memcpy(libs_out, libs, NSQ);
}

/* go.py:202-218: regions in ascending first-empty order; the region AND its border stones are
 * repainted with the border colour ('?' = 2 when mixed/none), which later regions then see. */
double bko_score(const int8_t *bd_in, double komi)
{
    geom_init();
    int8_t bd[NSQ], grp[NSQ];
    uint8_t in_grp[NSQ], border[NSQ];
    memcpy(bd, bd_in, NSQ);
    for (;;) {
        int e = -1;
        for (int s = 0; s < NSQ; ++s) if (bd[s] == 0) { e = s; break; }
        if (e < 0) break;
        int n = grp_fill(bd, e, in_grp, grp, border);
        int hasx = 0, haso = 0;
        for (int s = 0; s < NSQ; ++s) if (border[s]) { if (bd[s] == 1) hasx = 1; else if (bd[s] == -1) haso = 1; }
        int8_t c = (hasx && !haso) ? 1 : (haso && !hasx) ? -1 : 2;
        for (int s = 0; s < NSQ; ++s) if (border[s]) bd[s] = c;
        for (int i = 0; i < n; ++i) bd[grp[i]] = c;
    }
    int nx = 0, no = 0;
    for (int s = 0; s < NSQ; ++s) { if (bd[s] == 1) ++nx; else if (bd[s] == -1) ++no; }
    return nx - (no + komi);
}

/* One draw of torch.multinomial(probs, 1): exponential race, IEEE fp32 divide, first max. */
static int race_argmax(const float *p, const float *q)
{
    int best = 0;
    float bv = p[0] / q[0];
    for (int i = 1; i < NSQ; ++i) { float v = p[i] / q[i]; if (v > bv) { bv = v; best = i; } }
    return best;
}

/* forward declaration: counter-based Exp(1) draws, defined below */
void bko_exp_draws(uint64_t seed, uint32_t game, uint32_t move, uint32_t tr, float *q);

/* mcts.py:348-360 get_move, given the node's (already normalised) dist.probs and one Exp(1) vector
 * per draw: q_inj[t*81+i] when injected (q_vecs of them), else bko_exp_draws(seed, game, turn, t).
 * probs is modified in place like the cached dist.  Returns the move or PASS; *n_draws = number of
 * vectors consumed.  Shim: no probability mass left => PASS instead of the crash. */
int bko_get_move_mcts(const int8_t *bd, int ko, int turn, float *probs, const float *q_inj, int q_vecs,
                      uint64_t seed, uint32_t game, int *n_draws)
{
    geom_init();
    int col = (turn & 1) ? -1 : 1, t = 0, tries = 0;
    float qbuf[NSQ];
    const float *q;
    if (q_inj) q = q_inj; else { bko_exp_draws(seed, game, (uint32_t)turn, 0, qbuf); q = qbuf; }
    int mv = race_argmax(probs, q); ++t;
    while (!bko_is_legal(bd, ko, turn, mv) || bko_possible_eye(bd, mv) == col) {
        if (tries >= NSQ) { mv = BK_PASS; break; }
        probs[mv] = 0.0f;
        int any = 0;
        for (int i = 0; i < NSQ; ++i) if (probs[i] > 0.0f) any = 1;
        if (!any) { mv = BK_PASS; break; }
        if (q_inj) { if (t >= q_vecs) { mv = -4; break; } q = q_inj + NSQ * t; }   /* -4: ran out of injected draws */
        else { bko_exp_draws(seed, game, (uint32_t)turn, (uint32_t)t, qbuf); q = qbuf; }
        mv = race_argmax(probs, q); ++t;
        ++tries;
    }
    if (n_draws) *n_draws = t;
    return mv;
}

/* bin/selfplay.py:35-47 legal_sample: one unmasked draw; if illegal, the legal move of highest
 * probability (ties: lowest index -- torch.topk leaves tie order unspecified).  -2 => None. */
int bko_get_move_selfplay(const int8_t *bd, int ko, int turn, const float *probs, const float *q)
{
    geom_init();
    int mv = race_argmax(probs, q);
    if (bko_is_legal(bd, ko, turn, mv)) return mv;
    int best = BK_NONE; float bv = -1.0f;
    for (int i = 0; i < NSQ; ++i)
        if (probs[i] > bv && bko_is_legal(bd, ko, turn, i)) { bv = probs[i]; best = i; }
    return best;
}

/* ---- counter-based Exp(1) stream shared with the CUDA kernels -------------------------------- *
 * Philox4x32-10 keyed by (seed), counter (game, move, try, lane/4); 24-bit uniforms mapped to
 * q = -log(u) by a fixed sequence of correctly-rounded fp32 operations (fmaf only), so that host
 * and device produce the same bits.  Not part of the reference: it replaces torch's global RNG so
 * results are independent of how games are sharded over GPUs. */
static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void bko_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4])
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c[4] = {c0, c1, c2, c3};
    for (int r = 0; r < 10; ++r) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    memcpy(out, c, sizeof(c));
}

float bko_exp_from_bits(uint32_t bits)
{
    /* u = (k + 0.5) * 2^-23, k = top 23 bits: exactly representable, never 0 or 1 */
    float u = ((float)(bits >> 9) + 0.5f) * 1.1920928955078125e-07f;
    uint32_t ib; memcpy(&ib, &u, 4);
    int e = (int)(ib >> 23) - 127;
    uint32_t mb = (ib & 0x007FFFFFu) | 0x3F800000u;
    float m; memcpy(&m, &mb, 4);
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }          /* exact */
    float t = m - 1.0f;                                       /* exact (Sterbenz) */
    /* log1p(t) on [-0.2929, 0.4142]: fixed-order Horner evaluation of the Maclaurin series */
    float p = -0.0833333333f;
    p = fmaf(p, t, 0.0909090909f);
    p = fmaf(p, t, -0.1f);
    p = fmaf(p, t, 0.1111111111f);
    p = fmaf(p, t, -0.125f);
    p = fmaf(p, t, 0.1428571429f);
    p = fmaf(p, t, -0.1666666667f);
    p = fmaf(p, t, 0.2f);
    p = fmaf(p, t, -0.25f);
    p = fmaf(p, t, 0.3333333333f);
    p = fmaf(p, t, -0.5f);
    p = fmaf(p, t, 1.0f);
    float lg = fmaf((float)e, 0.693147180559945f, p * t);
    return -lg;
}

/* q[81] for (seed, game, move, try) */
void bko_exp_draws(uint64_t seed, uint32_t game, uint32_t move, uint32_t tr, float *q)
{
    for (uint32_t blk = 0; blk < 21; ++blk) {
        uint32_t r[4];
        bko_philox(seed, game, move, tr, blk, r);
        for (int j = 0; j < 4; ++j) { uint32_t i = 4 * blk + j; if (i < NSQ) q[i] = bko_exp_from_bits(r[j]); }
    }
}

/* ---- batch drivers (OpenMP) used by the tests and by the CPU-baseline legs of bench.py -------- */
void bko_features_batch(const int8_t *bd, const int16_t *ko, const int16_t *last, const int16_t *turn,
                        const uint8_t *libs_in, uint8_t *feats, uint8_t *legal, uint8_t *libs_out, int B)
{
    geom_init();
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        bko_features(bd + (size_t)b * NSQ, ko[b], last[b], turn[b], libs_in ? libs_in + (size_t)b * NSQ : 0,
                     feats + (size_t)b * 27 * NSQ, legal ? legal + (size_t)b * NSQ : 0,
                     libs_out ? libs_out + (size_t)b * NSQ : 0);
}

void bko_score_batch(const int8_t *bd, double komi, double *out, int B)
{
    geom_init();
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) out[b] = bko_score(bd + (size_t)b * NSQ, komi);
}

/* One playout step for B boards.  mode 0 = mcts flavour, 1 = self-play flavour.  done[b]!=0 boards
 * are left alone.  If q_inj!=NULL it holds q_vecs Exp(1) vectors per board, else draws come from
 * bko_exp_draws(seed, game0+b, turn, try).  probs [B][81] (used as given, no renormalisation).
 * moves_out[b]: chosen move (-1 PASS, -2 none/stopped, -3 board was already done). */
void bko_step_batch(int8_t *bd, int16_t *ko, int16_t *last, int16_t *turn, uint8_t *libs, uint8_t *done,
                    const float *probs, const float *q_inj, int q_vecs, uint64_t seed, uint32_t game0,
                    int mode, int max_turn, int16_t *moves_out, int B)
{
    geom_init();
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        int8_t *bb = bd + (size_t)b * NSQ;
        if (done[b]) { moves_out[b] = -3; continue; }
        float p[NSQ], qbuf[NSQ];
        memcpy(p, probs + (size_t)b * NSQ, sizeof(p));
        const float *q = q_inj ? q_inj + (size_t)b * q_vecs * NSQ : 0;
        int mv;
        if (mode == 0) mv = bko_get_move_mcts(bb, ko[b], turn[b], p, q, q_vecs, seed, game0 + (uint32_t)b, 0);
        else {
            if (!q) { bko_exp_draws(seed, game0 + (uint32_t)b, (uint32_t)turn[b], 0, qbuf); q = qbuf; }
            mv = bko_get_move_selfplay(bb, ko[b], turn[b], p, q);
        }
        moves_out[b] = (int16_t)mv;
        if (mv == BK_NONE || mv == -4) { done[b] = 1; continue; }
        bko_play(bb, ko + b, last + b, turn + b, libs ? libs + (size_t)b * NSQ : 0, 0, mv);
        if (mode == 0) { if (turn[b] > max_turn || last[b] == BK_PASS) done[b] = 1; }   /* mcts.py:362-364 */
        else { if (turn[b] > max_turn + 1) done[b] = 1; }   /* selfplay.py:21-33: pair-wise check, even start */
    }
}
