// Host build of the kernels' bit-board rules (bokego_b200/csrc/bk_bitboard.cuh) so that they can be
// checked against the oracle and the golden vectors without a GPU.  Test-only.
#include "../bokego_b200/csrc/bk_bitboard.cuh"
#include <cstring>

static void load(const int8_t *bd, BB &black, BB &white)
{
    black = bb_zero(); white = bb_zero();
    for (int s = 0; s < 81; ++s) {
        if (bd[s] == 1) black = bb_or(black, bb_bit(s));
        else if (bd[s] == -1) white = bb_or(white, bb_bit(s));
    }
}
static void store(BB black, BB white, int8_t *bd)
{
    for (int s = 0; s < 81; ++s) bd[s] = bb_test(black, s) ? 1 : (bb_test(white, s) ? -1 : 0);
}

extern "C" {

// mirrors the per-lane work of bk_encode_kernel, serially over the 81 squares
void hb_features(const int8_t *bd, int ko, int last, int turn, const uint8_t *libs_in, uint8_t *feats,
                 uint8_t *legal, uint8_t *libs_out)
{
    BB black, white;
    load(bd, black, white);
    bool blk = (turn & 1) == 0;
    BB own = blk ? black : white, opp = blk ? white : black;
    memset(feats, 0, 27 * 81);
    bool stale = libs_in && last >= 0 && libs_in[last] == 0;
    for (int p = 0; p < 81; ++p) {
        bool mine = bb_test(own, p), theirs = bb_test(opp, p);
        int lib = libs_in ? bb_lazy_lib_of(black, white, last, stale, p, libs_in[p]) : bb_exact_lib_of(black, white, p);
        libs_out[p] = (uint8_t)lib;
        int la = 0, cp = 0; bool lg = false;
        if (!mine && !theirs) {
            Cand c = bb_candidate(own, opp, p, 0, 0);
            lg = bb_listed_legal(own, opp, ko, p, c);
            if (lg) { la = c.libs_after; cp = c.caps; }
        }
        legal[p] = lg;
        feats[0 * 81 + p] = mine; feats[1 * 81 + p] = theirs; feats[2 * 81 + p] = !mine && !theirs;
        feats[3 * 81 + p] = blk; feats[4 * 81 + p] = (p == last); feats[5 * 81 + p] = lg;
        if (lib) feats[(6 + (lib > 6 ? 6 : lib - 1)) * 81 + p] = lib > 6 ? 7 : lib;
        if (la) feats[(13 + (la > 6 ? 6 : la - 1)) * 81 + p] = la > 6 ? 7 : la;
        if (cp) feats[(20 + (cp > 6 ? 6 : cp - 1)) * 81 + p] = cp > 6 ? 7 : cp;
    }
}

int hb_play(int8_t *bd, int *ko, int *last, int *turn, int mv)
{
    BB black, white;
    load(bd, black, white);
    int st = bb_play(black, white, *ko, *last, *turn, mv);
    store(black, white, bd);
    return st;
}

int hb_is_legal(const int8_t *bd, int ko, int turn, int s)
{
    BB black, white;
    load(bd, black, white);
    bool blk = (turn & 1) == 0;
    return bb_is_legal_quirk(blk ? black : white, blk ? white : black, ko, s);
}

int hb_eye(const int8_t *bd, int s)
{
    BB black, white;
    load(bd, black, white);
    return bb_possible_eye(black, white, s);
}

int hb_score_diff(const int8_t *bd)
{
    BB black, white;
    load(bd, black, white);
    return bb_score_diff(black, white);
}

void hb_exp_draws(uint64_t seed, uint32_t game, uint32_t move, uint32_t tr, float *q)
{
    for (int i = 0; i < 81; ++i) q[i] = bk_exp_draw(seed, game, move, tr, i);
}
}
