/* bokego_b200.h -- C ABI of libbokego_b200.so: the B200 (sm_100a) implementation of BokeGo's batched
 * leaf-evaluation / playout hot path.
 *
 * The reference (meiji163/bokego) is pure Python and has no FFI; the interface each entry point
 * replaces is therefore a Python function of the reference, cited per function below
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add to bokego/nnet.py / bokego/mcts.py to bind them.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name says host; buffers are caller-allocated,
 *     contiguous, and owned by the caller; the library allocates no device memory.  Its only state is launch state kept PER
 *     DEVICE behind a mutex (kernel attributes, SM count, tensor maps over the weight blobs it has seen), so the entry
 *     points may be called from several host threads and for several devices of one process; the current device
 *     (cudaSetDevice) selects where the work runs;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value: 0 = ok, BK_ERR_ARG = bad argument, BK_ERR_DEVICE = not an sm_100 device,
 *     BK_ERR_LAUNCH = CUDA launch error (see cudaGetLastError);
 *   - batched calls never fail per board: per-board conditions come back in masks / move codes.
 *
 * Position encoding (mirrors go.Game, go.py:51-66): board int8[81], index 9*x+y, +1 = 'X' (black),
 * -1 = 'O' (white), 0 = '.'; ko int16 (-1 = None); last int16 (-1 = PASS, -2 = None); turn int16
 * (even = black to move); libs uint8[81] = Game._libs (lazy liberty cache, go.py:220-243).
 */
#ifndef BOKEGO_B200_H
#define BOKEGO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BK_OK 0
#define BK_ERR_ARG (-1)
#define BK_ERR_DEVICE (-2)
#define BK_ERR_LAUNCH (-3)

/* bk_forward flags */
#define BK_FLAG_POLICY 1   /* policy trunk: logits + softmax probabilities */
#define BK_FLAG_VALUE 2    /* value trunk: tanh value */
#define BK_FLAG_SIMT 4     /* validation only: CUDA-core kernel over the same packed operands */

/* move codes written by bk_playout_step */
#define BK_MOVE_PASS (-1)
#define BK_MOVE_NONE (-2)      /* self-play flavour: no legal move, game stopped (selfplay.py:44-45) */
#define BK_MOVE_FINISHED (-3)  /* board was already done */
#define BK_MOVE_NO_DRAWS (-4)  /* injected draw buffer exhausted */

int bk_version(void);
const char *bk_strerror(int code);
/* 0 when the current device is sm_100 (B200); BK_ERR_DEVICE otherwise.  There is no other path. */
int bk_device_check(void);

/* ---- (a) feature planes: nnet.features (bokego/nnet.py:182-262) for B positions ----------------
 * libs_in == NULL  -> fresh Game objects (exact liberties);  otherwise the carried Game._libs.
 * Outputs (each may be NULL): feats_conv = fp16 operand of bk_forward, bk_feats_conv_bytes(B) bytes;
 * feats_f32 = float32 [B][27][9][9] exactly as nnet.features returns; planes_u8 = same values as
 * uint8 [B][27][81]; legal_out uint8 [B][81] (Game.get_legal_moves, go.py:245-260);
 * libs_out uint8 [B][81] (Game._libs after the call; may alias libs_in for an in-place update). */
size_t bk_feats_conv_bytes(int B);
int bk_encode(const int8_t *boards, const int16_t *ko, const int16_t *last, const int16_t *turn,
              const uint8_t *libs_in, void *feats_conv, float *feats_f32, uint8_t *planes_u8,
              uint8_t *legal_out, uint8_t *libs_out, int B, void *stream);

/* float32 planes [B][27][9][9] (the tensor nnet.features returns and PolicyNet.forward receives,
 * nnet.py:54-57) -> the fp16 operand of bk_forward.  Values are rounded to fp16 (exact for feature planes). */
int bk_repack_f32(const float *feats_f32, void *feats_conv, int B, void *stream);

/* ---- (b) nets: PolicyNet.forward (nnet.py:19-57), ValueNet.forward (nnet.py:59-113), SOFT (nnet.py:16)
 * Weights: fold BatchNorm on the host, then bk_weights_pack (HOST pointers in, HOST blob out, blob is
 * bk_weights_blob_bytes() long) and copy the blob to the device once per net.
 *   w0 [128][27][5][5], w16 [6][128][128][3][3], bias [7][128], head_w [128] (conv.21.weight),
 *   head_b [81] (conv.21.bias), vtail NULL for a PolicyNet, else
 *   {bn_scale, bn_shift, lin2_bias, W1[64][81], b1[64], w2[64]} with BatchNorm folded.
 * bk_forward: logits/probs float32 [B][81] (may be NULL), value float32 [B]. */
size_t bk_weights_blob_bytes(void);
int bk_weights_pack(const float *w0, const float *w16, const float *bias, const float *head_w,
                    const float *head_b, const float *vtail, void *blob_host_out);
int bk_forward(const void *feats_conv, const void *blob_policy, const void *blob_value, float *logits,
               float *probs, float *value, int B, int flags, void *stream);

/* bk_encode + bk_forward in ONE launch: positions in, probabilities / values out.  The conv kernel computes the planes of every
 * item itself -- three of its epilogue warps run nnet.features (nnet.py:182-262) for the boards of the NEXT item, straight into
 * the shared-memory operand of layer 0, while the tensor pipe works on the current item -- so the planes never exist in global
 * memory and no encoder launch precedes the nets.  libs_in == NULL: fresh Games (exact liberties), else the carried Game._libs.
 * legal_out / libs_out (may be NULL) as in bk_encode, except that libs_out may alias libs_in only when ONE net is evaluated
 * (with both, every board is encoded once per net); the other arguments as in bk_forward (BK_FLAG_SIMT is not available here).  Results are bit-identical to bk_encode followed by bk_forward. */
int bk_forward_positions(const int8_t *boards, const int16_t *ko, const int16_t *last, const int16_t *turn,
                         const uint8_t *libs_in, const void *blob_policy, const void *blob_value, float *logits, float *probs,
                         float *value, uint8_t *legal_out, uint8_t *libs_out, int B, int flags, void *stream);

/* ---- (c) playout stepping ------------------------------------------------------------------------
 * mode 0: Go_MCTS.get_move + make_move + is_game_over (bokego/mcts.py:340-364)
 * mode 1: legal_sample + playout loop (bin/selfplay.py:18-47)
 * State is updated in place; done[b] != 0 boards are skipped.  probs float32 [B][81] are used as given.
 * Random draws: q_inj != NULL -> float32 [B][q_vecs][81] Exp(1) variates (what torch.multinomial
 * would draw); else the counter-based stream keyed (seed, game0 + b, turn, try). */
int bk_playout_step(int8_t *boards, int16_t *ko, int16_t *last, int16_t *turn, uint8_t *libs,
                    uint8_t *done, const float *probs, const float *q_inj, int q_vecs, uint64_t seed,
                    uint32_t game0, int mode, int max_turn, int16_t *moves_out, int B, void *stream);
/* bk_playout_step followed, in the same launch, by bk_encode of the position after the move (carried liberty cache, in place)
 * into feats_conv -- the "sample, play, capture, re-encode" step of a playout loop (mcts.py:195-206, selfplay.py:18-33), so that
 * a playout move is two launches: bk_forward (policy), bk_playout_step_encode.  libs must not be NULL; feats_conv must be the
 * buffer a previous bk_encode of the same B boards wrote (its padding rows stay untouched); boards that are finished after the
 * move keep their old planes. */
int bk_playout_step_encode(int8_t *boards, int16_t *ko, int16_t *last, int16_t *turn, uint8_t *libs, uint8_t *done,
                           const float *probs, const float *q_inj, int q_vecs, uint64_t seed, uint32_t game0, int mode,
                           int max_turn, int16_t *moves_out, void *feats_conv, int B, void *stream);
/* Whole playouts in ONE launch: the conv kernel keeps every item of <= 5 boards on its SM for n_steps moves.  Three warps per
 * board hold the position in shared memory and registers: they encode the starting position (nnet.features; fresh_libs != 0:
 * exact liberties like a fresh Game, else the carried cache in `libs`), and after every policy forward (blob_even for the moves
 * made at even first_turn + k, blob_odd for the odd ones; blob_odd NULL = one net) they sample, play, capture, refresh the cache
 * and write the planes of the new position straight into the shared-memory operand of the next move -- the loop of
 * MCTS._simulate (mcts.py:195-206) and of selfplay.playout (bin/selfplay.py:18-33) without a launch, a grid-wide dependency or
 * an HBM round trip between the moves.  boards / ko / last / turn / libs / done are read once and written once, and end exactly as
 * after n_steps calls of bk_playout_step_encode; moves_out int16 [n_steps][B].  Random stream: (seed, game0 + b, turn, try).
 * The item size adapts to B: the smallest that gives every board an SM in one round (a playout lasts as long as one item). */
int bk_playout_run(int8_t *boards, int16_t *ko, int16_t *last, int16_t *turn, uint8_t *libs, uint8_t *done,
                   const void *blob_even, const void *blob_odd, uint64_t seed, uint32_t game0, int mode, int max_turn,
                   int first_turn, int n_steps, int fresh_libs, int16_t *moves_out, int B, void *stream);
/* Go_MCTS.make_move (bokego/mcts.py:340-346) for C children: child c = copy of parent parent_idx[c] (an index into the
 * parent arrays) with moves[c] played by Game.play_move (go.py:123-182; -1 = play_pass go.py:109-121).  The lazy liberty
 * cache is refreshed on the parent position before the move (go.py:160) and handed to the child; libs == NULL means the
 * parents are fresh Games (exact liberties).  libs_out / status_out may be NULL.  status: 0 ok, 1 ko, 2 not_empty,
 * 3 suicide (the reference raises IllegalMove; the child is then the unchanged parent). */
int bk_make_moves(const int8_t *boards, const int16_t *ko, const int16_t *last, const int16_t *turn, const uint8_t *libs,
                  const int32_t *parent_idx, const int16_t *moves, int8_t *boards_out, int16_t *ko_out, int16_t *last_out,
                  int16_t *turn_out, uint8_t *libs_out, uint8_t *status_out, int C, void *stream);
/* Game.score (go.py:202-218) minus komi, and Go_MCTS.reward's +-1 (mcts.py:330-338) */
int bk_score(const int8_t *boards, float komi, float *score_out, int8_t *reward_out, int B, void *stream);
/* fixed-size result records of B finished playouts, the rows that cross ranks in the one gather of a multi-GPU run: rec int16
 * [B][T + 3] = { turn reached, reward, 2 * score, move 0 .. T-1 } from the move log moves int16 [T][B] of the playout loop */
int bk_pack_records(const int16_t *moves, const int16_t *turn, const float *score, const int8_t *reward, int16_t *rec, int T,
                    int B, void *stream);
/* the counter-based Exp(1) stream itself: q float32 [B][81] for (seed, game0 + b, move, try) */
int bk_exp_draws(uint64_t seed, uint32_t game0, uint32_t move, uint32_t tr, float *q, int B, void *stream);

/* ---- host-side core of the batched tree search (bokego/mcts.py:172-234: _descend, _puct_select, _backpropagate) ----------
 * HOST pointers to the flat arrays of one tree: N visits, V value sums, child0 / nchild (children are contiguous, -1 = not
 * expanded), move (the move that led to the node), prior float [nodes][81] (policy probabilities of the node's position), val
 * (value-net output of the node, NaN = not evaluated yet).  bk_tree_run makes descents until n_rollouts are complete or
 * leaf_batch of them wait for device work (an unevaluated leaf, or a leaf visited more than expand_thresh times that has to be
 * expanded); those are parked in pend_nodes [leaf_batch][max_depth] / pend_len / pend_expand under a virtual loss.  After the
 * caller has evaluated / expanded, bk_tree_finish removes the virtual losses and backs the values up. */
int bk_tree_run(int64_t *N, double *V, const int32_t *child0, const int32_t *nchild, const int16_t *move, const float *prior,
                const double *val, int root, int n_rollouts, int leaf_batch, int expand_thresh, double c_puct,
                int32_t *pend_nodes, int32_t *pend_len, int32_t *pend_expand, int max_depth, int *n_pending,
                double *Q, double value_weight, int have_value);
int bk_tree_finish(int64_t *N, double *V, const double *val, const int32_t *pend_nodes, const int32_t *pend_len,
                   int n_pending, int max_depth, int leaf_batch, double *Q, const double *reward, int have_value);
/* --simulate mode (MCTS(no_sim=False), mcts.py:133-151, 195-217): pass Q (playout reward sums, one per node) and
 * value_weight (value_net_weight of mcts.py:66-71: selection uses ((1 - w) Q + w V) / N); every descent is then parked, the
 * caller plays every parked leaf out on the device (bk_forward + bk_playout_step_encode until the boards are done, bk_score)
 * and hands bk_tree_finish reward[j] = +-1 as seen by the player to move at leaf j.  Q == NULL is the no_sim search.
 * have_value = 0: no value net (val is ignored, V stays 0). */

/* ---- REINFORCE step (bin/selfplay.py:59-122): train-mode PolicyNet forward, policy-gradient loss, backward, AdamW ---------
 * All pointers are device pointers.  Parameters, gradients and the Adam moments are flat float32 buffers of
 * bk_train_param_count() elements in the layout of the training GEMMs (k = tap * Cin + ci, tap = kh * R + kw):
 *   BK_TP_W0     conv.0.weight   as [25][32][128]  (tap, ci padded 27 -> 32 with zeros, co)
 *   BK_TP_W1     conv.3..18.weight, six layers of [9][128][128]  (tap, ci, co)
 *   BK_TP_VEC    seven layers of { conv.N.bias[128], BatchNorm weight[128], BatchNorm bias[128] }
 *   BK_TP_HEADW  conv.21.weight[128];  BK_TP_HEADB  conv.21.bias[81]
 * running = BatchNorm running statistics, float32 [2][7][128] (means of the 7 layers, then variances).
 * bn_mode 0: every position is normalised with its own statistics -- PolicyNet in train() mode called with one position
 *            per call, which is how policy_dist / policy_sample call it (nnet.py:265-297); 1: running statistics (eval()).
 * prec 5: 3xTF32 on tcgen05 (operands split into two TF32 values, TMEM accumulation chains kept short and summed in fp32
 *         registers: fp32-grade, the default of the host mirror); 4: TF32 on tcgen05; 1 / 0: the same two on warp-level
 *         mma.sync; 2: FFMA fp32 (validation path).
 * bk_train_forward: planes_u8 uint8 [P][27][81] (bk_encode's planes_u8 == nnet.features values) -> logits float32 [P][81]
 *   (PolicyNet.forward, nnet.py:54-57), probs = SOFT(logits) (nnet.py:16; may be NULL), stats_out (may be NULL) float32
 *   [P][7][2][128] = per-position channel mean and unbiased variance, i.e. what each train-mode call feeds into the running
 *   averages.  Activations stay in `workspace` (bk_train_workspace_bytes(P) bytes) for bk_train_backward.
 * bk_train_backward: loss = sum_p coef[p] * nlp[p], nlp[p] = -Categorical(probs_p).log_prob(moves[p]) (selfplay.py:98-100;
 *   coef[p] = reward / bs, selfplay.py:115-117); writes nlp_out float32 [P] and d loss / d params into `grads`
 *   (accumulate != 0: adds to it).  Must follow bk_train_forward with the same P, bn_mode and workspace.
 * bk_train_running_stats: the BatchNorm momentum filter (momentum 0.1 in the reference) over train-mode calls on positions
 *   seq[0..S) of stats (seq == NULL: 0..S-1), in call order.
 * bk_adamw_step: torch.optim.AdamW's update (selfplay.py:138) of n elements; step counts from 1. */
#define BK_TP_W0 0
#define BK_TP_W1 102400
#define BK_TP_VEC 987136
#define BK_TP_HEADW 989824
#define BK_TP_HEADB 989952
#define BK_TP_COUNT 990033
size_t bk_train_param_count(void);
size_t bk_train_workspace_bytes(int P);
int bk_train_launches(int which, int P, int prec); /* kernels launched by bk_train_forward (0) / bk_train_backward (1) for P positions on the
                                                      current device (the tcgen05 3x3 kernel adds a launch per layer when the tiles left over
                                                      after the last full round of the SMs are split, which depends on the SM count) */
int bk_train_conv3_schedule(int P, int n_sm, int ksplit, int *out4); /* host arithmetic only: the work items of the persistent tcgen05 3x3
                                                      kernel for P positions on n_sm SMs -- out4 = tiles of 128 raster rows, tiles that run
                                                      whole when the left-over tiles are split into single-channel-group items (0 = no split),
                                                      work items, CTAs; ksplit = 4 is the small-batch forward.  -1 on bad arguments */
int bk_train_forward(const float *params, const float *running, const uint8_t *planes_u8, int P, int bn_mode, int prec,
                     void *workspace, float *logits, float *probs, float *stats_out, void *stream);
int bk_train_backward(const float *params, const int16_t *moves, const float *coef, int P, int bn_mode, int prec,
                      void *workspace, float *grads, int accumulate, float *nlp_out, void *stream);
int bk_train_running_stats(float *running, const float *stats, const int32_t *seq, int S, float momentum, void *stream);
int bk_adamw_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, size_t n, double lr, double beta1,
                  double beta2, double eps, double weight_decay, int step, void *stream);

#ifdef __cplusplus
}
#endif
#endif
