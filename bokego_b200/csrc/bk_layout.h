/* bk_layout.h -- data layouts shared by the kernels, the weight packer and the host code.
 *
 * Conv work item = one trunk (policy or value) over a GROUP of 5 boards.
 *
 * Feature operand (written by bk_encode, read by layer 0), per group, fp16:
 *     [4 channel chunks][605 rows][8 channels]           (38,720 bytes, contiguous)
 *     row = 121*board + 22 + 11*x + y : stride-11 raster, columns 9,10 and the first two rows of
 *     every board are zero, so a 5x5 tap (i,j) is the row shift (i-2)*11 + (j-2).
 *
 * Activation operand (on chip only), fp16, 128 channels:
 *     [16 channel chunks][524 rows][8 channels], row = 12 + 100*board + 10 + 10*x + y :
 *     stride-10 raster with one zero column, one zero row above every board; a 3x3 tap (i,j) is the
 *     row shift (i-1)*10 + (j-1).  GEMM rows 0..511 = 4 M-tiles of 128; 405 of them are real squares.
 *
 * Both are "K-major, no swizzle" UMMA operands whose 8-row core matrices are contiguous in M
 * (stride-byte-offset 128) -- so a row-shifted window is just a different start address.
 *
 * Weight blob (per net), bytes.  The conv kernel runs on CTA pairs (tcgen05 cta_group::2): the two CTAs share
 * the B operand, each holding the weights of 64 of the 128 output channels ("N half" h = co / 64).
 *     stage   = four K steps (64 K values) = 16 KiB: [2 N-halves][4 k-steps][2 k-chunks][64 co][8 k] fp16,
 *               so the 8 KiB a CTA needs of a stage are contiguous
 *     layer 0 : 13 stages, K = tap*32 + ci (25 taps, ci padded 27->32; the last stage is half empty), then 4 KiB of bias rows
 *     layer l : 18 stages, K = tap*128 + ci (l = 1..6),                                               then 4 KiB of bias rows
 *     bias rows = [2 N-halves][2 k-chunks][64 co][8 k]: k = 0, 1 hold the folded bias as fp16 hi + lo; they are
 *               multiplied by an all-ones A operand, so the bias is added by the tensor core
 *     then fp32: bias[7][128] (BatchNorm folded; read by the validation kernel), head_w[128], head_b[96] (81 used),
 *     value tail: {bn_scale, bn_shift, lin2_b, 0}, W1T[81][64] (BN1d folded, transposed), b1[64], w2[64]
 */
#ifndef BK_LAYOUT_H
#define BK_LAYOUT_H

#define BK_GROUP 5
#define BK_F_CHUNKS 4
#define BK_F_ROWS_B 121
#define BK_F_ROWS_G 605
#define BK_F_GROUP_BYTES (BK_F_CHUNKS * BK_F_ROWS_G * 16)

#define BK_KSTEP_BYTES 4096
#define BK_STAGE_BYTES 16384
#define BK_BIAS_BYTES 4096
#define BK_L0_STAGES 13
#define BK_L_STAGES 18
#define BK_L0_BYTES (BK_L0_STAGES * BK_STAGE_BYTES + BK_BIAS_BYTES) /* 217,088 */
#define BK_L_BYTES (BK_L_STAGES * BK_STAGE_BYTES + BK_BIAS_BYTES)   /* 299,008 */
#define BK_W_L0_OFF 0
#define BK_W_L_OFF(l) (BK_L0_BYTES + ((l) - 1) * BK_L_BYTES)
#define BK_W_BIAS_OFF (BK_L0_BYTES + 6 * BK_L_BYTES) /* 2,011,136 */
/* byte offset inside a layer of weight (K index k, output channel co) */
#define BK_W_OFF(k, co) ((size_t)((k) >> 6) * BK_STAGE_BYTES + ((co) >> 6) * 8192 + (((k) >> 4) & 3) * 2048 + \
                         (((k) >> 3) & 1) * 1024 + ((co) & 63) * 16 + ((k) & 7) * 2)
/* byte offset inside a layer of bias row k (0 = hi, 1 = lo) of output channel co; n_stages = stages of the layer */
#define BK_W_BIAS_ROW_OFF(n_stages, k, co) ((size_t)(n_stages) * BK_STAGE_BYTES + ((co) >> 6) * 2048 + ((co) & 63) * 16 + (k) * 2)
#define BK_W_HEADW_OFF (BK_W_BIAS_OFF + 7 * 128 * 4)
#define BK_W_HEADB_OFF (BK_W_HEADW_OFF + 128 * 4)
#define BK_W_VT_OFF (BK_W_HEADB_OFF + 96 * 4)
#define BK_W_VT_W1T_OFF (BK_W_VT_OFF + 16)
#define BK_W_VT_B1_OFF (BK_W_VT_W1T_OFF + 81 * 64 * 4)
#define BK_W_VT_W2_OFF (BK_W_VT_B1_OFF + 64 * 4)
#define BK_W_BLOB_BYTES (((BK_W_VT_W2_OFF + 64 * 4) + 127) / 128 * 128)

/* bk_forward flags */
#define BK_FWD_POLICY 1      /* run the policy trunk: logits + probs */
#define BK_FWD_VALUE 2       /* run the value trunk: value */
#define BK_FWD_SIMT 4        /* validation path: plain CUDA-core kernel instead of tcgen05 */
#define BK_FWD_NOSPLIT 8     /* measurement only: do not split the items of the last partial round */


/* per-device launch state (kernel attributes, SM count, tensor maps) is kept in arrays of this size, guarded by a mutex */
#define BK_MAX_DEVICES 64
#ifdef __cplusplus
int bk_current_device_slot(void);   /* bk_api.cu: index of the current CUDA device, -1 if none / out of range */
#endif

#endif
