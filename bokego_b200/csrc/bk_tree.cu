// bk_tree.cu -- host-side core of the batched tree search (bokego_b200/mcts.py): PUCT descents, virtual loss and back-up on
// the flat arrays of the tree.  Plain C++ (no device code); lives in the shared library so that a search of 1600 playouts is
// not bound by interpreter overhead.  Search rule = the reference's MCTS, in no_sim mode (value net only) and in --simulate
// mode (every rollout also plays the leaf out; Q = playout rewards, mixed with the value sums V by value_net_weight):
//   _descend / _puct_select  /root/reference/bokego/mcts.py:172-183, 219-234
//   rollout / _simulate       mcts.py:133-151, 195-206 (the playouts themselves run on the device, bokego_b200/playout.py)
//   _backpropagate            mcts.py:208-217
// All arithmetic is IEEE double in the same order as the Python expressions of the reference, so that visit counts agree
// count for count (tests/golden/mcts.npz).
#include <math.h>
#include <stdint.h>

namespace {

struct Tree {
    int64_t *N; double *V; const int32_t *child0, *nchild; const int16_t *move; const float *prior; const double *val;
    double *Q; double w;     // --simulate: playout reward sums and value_net_weight (Q == nullptr: no_sim, w = 1)
};

// child of i with the highest PUCT score; ties go to the lowest index (= lowest move)
inline int select_child(const Tree &t, int i, double c_puct)
{
    const int lo = t.child0[i], c = t.nchild[i];
    int64_t total = 0;
    for (int k = 0; k < c; ++k) total += t.N[lo + k];
    if (total < 1) total = 1;
    const double sq = sqrt((double)total);
    int best = lo;
    double best_s = -INFINITY;
    for (int k = 0; k < c; ++k) {
        const int ch = lo + k;
        const double n = (double)t.N[ch];
        // mcts.py:228-230: ((1 - w) * Q + w * V) / N, the same expression order (w = 1 without simulations => V / N)
        const double avg = t.N[ch] > 0 ? (t.Q ? ((1.0 - t.w) * t.Q[ch] + t.w * t.V[ch]) / n : t.V[ch] / n) : 0.0;
        const double p = (double)t.prior[(size_t)i * 81 + t.move[ch]];
        const double s = -avg + c_puct * p * sq / (1.0 + n);
        if (s > best_s) { best_s = s; best = ch; }
    }
    return best;
}

// mcts.py:208-217.  reward = playout result seen by the player to move at the leaf (0 = no simulation: Q untouched, which is
// also what the reference's `if reward:` does); use_val = a value net exists (mcts.py:215)
inline void backup(const Tree &t, const int32_t *path, int len, double reward, bool use_val)
{
    double v = use_val ? t.val[path[len - 1]] : 0.0;
    for (int d = len - 1; d >= 0; --d) {
        t.N[path[d]] += 1;
        if (t.Q && reward != 0.0) { t.Q[path[d]] += reward; reward = -reward; }
        if (use_val) { t.V[path[d]] += v; v = -v; }
    }
}

}  // namespace

// Runs rollouts from `root` until `n_rollouts` are complete or `leaf_batch` descents are waiting for device work.
// A descent whose leaf already has a value and needs no expansion is backed up at once.  Any other descent is parked:
// its path goes to pend_nodes[j * max_depth ...] (length pend_len[j]), the node that has to be expanded (visited more than
// expand_thresh times, not expanded yet, already evaluated) to pend_expand[j] (-1 = none), and, when leaf_batch > 1, a virtual
// loss (N += 1, V += 1) is left on its path.  Returns the number of completed rollouts; *n_pending receives the parked ones.
// --simulate (Q != nullptr): EVERY descent is parked, because each rollout plays its leaf out on the device (mcts.py:147-148);
// the virtual loss then also covers Q.  have_value = 0: there is no value net (then val is never waited for).
extern "C" int bk_tree_run(int64_t *N, double *V, const int32_t *child0, const int32_t *nchild, const int16_t *move,
                           const float *prior, const double *val, int root, int n_rollouts, int leaf_batch, int expand_thresh,
                           double c_puct, int32_t *pend_nodes, int32_t *pend_len, int32_t *pend_expand, int max_depth,
                           int *n_pending, double *Q, double value_weight, int have_value)
{
    const Tree t = {N, V, child0, nchild, move, prior, val, Q, value_weight};
    const bool sim = Q != nullptr;
    int done = 0, pend = 0;
    while (done + pend < n_rollouts && pend < leaf_batch) {
        int32_t *path = pend_nodes + (size_t)pend * max_depth;
        int len = 0, i = root, want_expand = -1;
        path[len++] = i;
        for (;;) {
            if (nchild[i] <= 0) {
                if (nchild[i] < 0 && N[i] > expand_thresh && (!have_value || !isnan(val[i]))) want_expand = i;
                break;
            }
            if (len >= max_depth) break;
            i = select_child(t, i, c_puct);
            path[len++] = i;
        }
        if (!sim && !isnan(val[i]) && want_expand < 0) {
            backup(t, path, len, 0.0, true);
            ++done;
            continue;
        }
        pend_len[pend] = len;
        pend_expand[pend] = want_expand;
        if (leaf_batch > 1)
            for (int d = 0; d < len; ++d) {
                N[path[d]] += 1; V[path[d]] += 1.0;
                if (sim) Q[path[d]] += 1.0;
            }
        ++pend;
    }
    *n_pending = pend;
    return done;
}

// Takes the virtual losses of the parked descents back (when leaf_batch > 1) and backs their leaf values up.
// --simulate: reward[j] = result of the playout from the leaf of parked descent j as the player to move AT THE LEAF sees it
// (+1 / -1: Go_MCTS.reward is Black's view and _simulate inverts it for a leaf with odd turn, mcts.py:199-204).
extern "C" int bk_tree_finish(int64_t *N, double *V, const double *val, const int32_t *pend_nodes, const int32_t *pend_len,
                              int n_pending, int max_depth, int leaf_batch, double *Q, const double *reward, int have_value)
{
    const Tree t = {N, V, nullptr, nullptr, nullptr, nullptr, val, Q, 1.0};
    if (leaf_batch > 1)
        for (int j = 0; j < n_pending; ++j)
            for (int d = 0; d < pend_len[j]; ++d) {
                const int node = pend_nodes[(size_t)j * max_depth + d];
                N[node] -= 1; V[node] -= 1.0;
                if (Q) Q[node] -= 1.0;
            }
    for (int j = 0; j < n_pending; ++j)
        backup(t, pend_nodes + (size_t)j * max_depth, pend_len[j], (Q && reward) ? reward[j] : 0.0, have_value != 0);
    return n_pending;
}
