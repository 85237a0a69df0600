/* bk_train_args.h -- argument blocks shared by the two implementations of the training GEMMs
 * (bk_train.cu: warp-level mma.sync / FFMA; bk_train_tc.cu: tcgen05 with TMEM accumulators). */
#ifndef BK_TRAIN_ARGS_H
#define BK_TRAIN_ARGS_H

#include <cuda_runtime.h>

/* implicit-GEMM convolution: out[m][co] = sum_{tap,ci} in[m shifted by sign*tap][ci] * w[(tap,ci)][co] (+ bias[co])
 * in: [M/81][81][Cin], w: [R*R*Cin][128], out: [M][128].  sign = -1 with per-tap transposed weights = data gradient. */
struct BkConvArgs {
    const float *in, *w, *bias;
    float *out;
    int M, Cin, R, sign;
    const float *w_lo;   /* tcgen05 path only: w = packed TF32 high parts [k / 4][co][k % 4], w_lo = the low parts (3xTF32) */
    int ksplit;          /* staged-once 3x3 kernel only: 1, or 4 = one CTA per channel group of 32 (grid.y), group g writes its partial
                            result to out + g * M * 128 and bk_train_sum4_kernel adds them in order.  For small batches: a 16-position
                            forward has 13 tiles, this puts 52 CTAs to work on a quarter of the K loop each. */
    float *tail_part = nullptr;   /* staged-once 3x3 kernel, ksplit == 1: scratch of BK_CONV3_TAIL_FLOATS floats.  When the tiles left over after
                            the last complete wave of the grid are few, they run as four channel-group CTAs each (a quarter of the K
                            loop), write their partial results here and bk_train_conv3_tail_kernel adds them in group order --
                            576 positions: 450 tiles on 148 SMs = three waves and a quarter instead of four. */
    int tail_full = 0;   /* set by bk_tc_launch_conv: tiles below this index run whole, 0 = no split tail */
    int n_items = 0;     /* set by bk_tc_launch_conv: work items of the persistent CTAs (whole tiles + single-group items) */
    const unsigned short *w_bh = nullptr, *w_bl = nullptr;   /* staged-once 3x3 kernel built with BK_R3_LO_BF16: the weights' high and low
                            parts as bf16, packed [k / 8][co][k % 8] (the low-order products of 3xTF32 on kind::f16) */
};
#define BK_CONV3_TAIL_MAX 24                                  /* most tiles a split tail may have */
#define BK_CONV3_TAIL_FLOATS (4 * BK_CONV3_TAIL_MAX * 128 * 128)

/* weight gradient: part[split][k][co] = sum over the split's rows m of act[m shifted by tap(k)][ci(k)] * dz[m][co] */
struct BkWgradArgs {
    const float *act, *dz;
    float *part;
    int M, Cin, R, K, rows_per_split;
};

/* tcgen05 versions (bk_train_tc.cu); three_x != 0: 3xTF32 split operands.  Return a cudaError_t as int. */
int bk_tc_set_attrs(void);
int bk_tc_lo_bf16(void);       /* 1 in the BK_R3_LO_BF16 measurement build of the 3x3 kernel (it reads w_bh / w_bl) */
int bk_tc_conv3_tail(int P);   /* tiles of a P-position 3x3 conv that run whole on this device when its tail is split (see tail_part), else 0 */
void bk_tc_launch_conv(const BkConvArgs &a, int three_x, cudaStream_t st);
void bk_tc_launch_wgrad(const BkWgradArgs &a, int splits, int three_x, cudaStream_t st);

#endif
