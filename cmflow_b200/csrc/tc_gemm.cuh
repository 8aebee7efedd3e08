// tc_gemm.cuh -- launch interface of the tcgen05 (5th-gen tensor core) GEMM used by the fast pipeline.
#pragma once
#include "cmf_common.cuh"

enum { TC_PROD_PLAIN = 0, TC_PROD_FC_H1 = 1, TC_PROD_SC2_Y1 = 2, TC_PROD_TILED = 3 };
enum { TC_EPI_STORE = 0, TC_EPI_MAXK = 1, TC_EPI_WSUM = 2 };

// Out[c][m] = epi( sum_k W[m][k] * B[c][k] )   computed as 3xTF32 (W_hi*B_hi + W_hi*B_lo + W_lo*B_hi), fp32 accumulate in TMEM.
struct TcArgs {
    // A operand: weights pre-tiled by cmf_tc_tile_weights(): [m_block][k_block]{hi tile, lo tile}, tile = 128 x 32 fp32, 128B-swizzled
    const float *Wt;
    int m_blocks, k_blocks;      // 128-row blocks, 32-col blocks
    int M;                       // true number of output channels (rows beyond M are zero padding)
    long long cols;              // number of B rows (columns of the output)
    // B operand producer
    int prod;
    const float *X; int ldx;                                     // PLAIN: row c = X + c*ldx
    const float *U1, *U2, *Wsmall;                               // FC_H1: leaky(U1[i]+U2[j]+Wd.dir); SC2_Y1: relu(P[j]+Wx.rel) with U2=P, Wsmall=Wx/Wd (C x 4)
    const float *xyz_q, *xyz_c; const int *nbr;                  // planar (B,3,N) clouds of the query / candidate points, neighbour table
    int n_pts, ksamp, nbr_ld, nbr_off, ld_u2, off_u2;            // points per (query) cloud, neighbours per point, table row stride/offset, gathered-row stride/offset
    int n_cand;                                                  // points per candidate cloud (xyz_c, U2 rows) when it differs from n_pts; 0 = same
    const float *Xt;                                             // TILED: activations already split + swizzled by a previous tc GEMM (out_tiled)
    // epilogue
    int epi;
    int out_tiled;                                               // STORE only: write [col_tile][k_block]{hi,lo} 256x32 swizzled tiles instead of rows
    float *Out; int ldo;
    const float *bias, *pbias; int pb_ld, cols_per_pair, act;
    // WSUM epilogue (flow embedding, radarflow_util.py:215-225): Out[point][m] = sum_k WeightNet(dir_ik)[m] * act(acc + bias)[column (point,k)]
    // WeightNet = 3 -> 8 -> 8 -> C, ReLU after every layer; dir = xyz_c[nbr] - xyz_q[point]; ksamp neighbours per point (pair kernel, TILED producer only)
    const float *wnA1, *wna1, *wnA2, *wna2, *wnA3, *wna3;
    // ---- operand format ----------------------------------------------------------------------------------------------------------
    // fmt 0: 3xTF32 (kind::tf32, K = 16 floats per 64-byte stage row).  fmt 1: 3xFP16 (kind::f16, K = 32 halfs per 64-byte stage row):
    // same 22-bit split (11 + 11 significand bits) at twice the tensor rate and half the operand bytes; fp16's 5-bit exponent is
    // handled by exact power-of-two scaling: weight row m is stored times 2^e(m) (a_inv[m] = 2^-e(m)), the activation rows of frame
    // pair p times b_scale(p), and the epilogue multiplies the accumulator by a_inv[m] / b_scale(p) before bias / activation.
    int fmt;
    const float *a_inv;                                          // fmt 1: per-output-channel 1/scale of the tiled weights (NULL = 1)
    // b_scale(p): bs_mode 0 -> 1; 1 -> pow2_scale(bound(p)), bound(p) = bs_const + sum_i bs_coef[i] * bs_src[i][p] (a rigorous
    // per-pair bound on |activation| built from measured per-pair maxima of the upstream tensors); 2 -> bs_src[0][p] holds the scale
    int bs_mode;
    const float *bs_src[3]; float bs_coef[3]; float bs_const;
    // tiled output (fmt 1): scale(p) = pow2_scale(out_mul * bound(p) + out_add) (out_mul = max_m |W[m]|_1, out_add = max_m |bias[m]|),
    // also written to out_scale_store[p] for the consuming GEMM (bs_mode 2)
    float out_mul, out_add; float *out_scale_store;
    // row-major STORE epilogue: atomicMax of |out| per frame pair into amax_out[(m / amax_group) * amax_ld + p] (uint bits; NULL = off)
    unsigned int *amax_out; int amax_group, amax_ld;
    // optional wait-time instrumentation (pair kernel): long long[gridDim.x][8] cycles = {total, mma:tempty, mma:full, mma:peer_full,
    // loader:empty, producer(warp 8):empty, epilogue(warp 4):tfull, tiles}; NULL in production
    long long *dbg;
};

size_t cmf_tc_tiled_floats(int M, int K);                                      // floats needed for the pre-tiled copy of an M x K matrix (either format)
size_t cmf_tc_act_tiled_floats(long long cols, int C);                          // floats of a tiled activation buffer (cols x C channels; sized for fmt 0, fmt 1 uses half)
int cmf_tc_tile_weights(const float *W, int ldw, int M, int K, float *Wt, cudaStream_t st);
// fmt 1: a_inv (M floats, rounded up to 128) receives 2^-e(m); Wt receives the fp16 hi/lo tiles of W[m][:] * 2^e(m)
int cmf_tc_tile_weights_f16(const float *W, int ldw, int M, int K, float *Wt, float *a_inv, cudaStream_t st);
int cmf_launch_tc_gemm(const TcArgs &a, cudaStream_t st);      // one CTA per 128 x 256 tile
int cmf_launch_tc_gemm2(const TcArgs &a, cudaStream_t st);     // CTA pair (cta_group::2) per 256 x 256 tile; needs M % 256 == 0
int cmf_tc_pair_enabled();
// set-conv #2 layers 2 + 3 + max over the K neighbours in one kernel (tc_sc2.cu; 3xFP16 only).  l2 = the layer-2 arguments exactly as for
// cmf_launch_tc_gemm2 with the SC2_Y1 producer (its Out / out_tiled / out_scale_store are ignored); Wt3 = tiled 64 x 256 layer-3 weights with
// per-channel un-scale a_inv3 and bias3; out[point][0..63] (row stride ldo floats) = max_k relu(W3 relu(W2 x_k + b2) + b3)
int cmf_launch_sc2_fused(const TcArgs &l2, const float *Wt3, const float *a_inv3, const float *bias3, float *out, int ldo, cudaStream_t st);
// 2-D tensor map (CUtensorMap *, passed as void * to keep <cuda.h> out of this header) over a row-major fp32 matrix, box = box_cols x box_rows
int cmf_make_row_map(void *tensor_map, const float *base, long long rows, int ld, int box_cols, int box_rows);
int cmf_launch_tc_auto(const TcArgs &a, cudaStream_t st);
// watchdog record (tc_dev.cuh): one setter per translation unit with mbarrier waits
int cmf_wd_set_tc_gemm(unsigned long long *dev_ptr);
int cmf_wd_set_tc_gemm2(unsigned long long *dev_ptr);
int cmf_wd_set_tc_sc2(unsigned long long *dev_ptr);
int cmf_wd_set_tc_chain(unsigned long long *dev_ptr);      // pair kernel when M % 256 == 0 (unless CMF_TC2=0), else the one-CTA kernel

// ---- narrow MLP chains with the activations as the A operand (tc_chain.cu; 3xFP16 only) ---------------------------------------------
// Weight tiles come from cmf_tc_tile_weights_f16 (one 128-row block, K/32 blocks of {hi, lo}); ainv = its per-row un-scale.
struct TcChainSc1W { const float *W1, *b1, *b2, *b3, *W2t, *ainv2, *W3t, *ainv3; };     // one scale of mse_layer: 32x8 fp32, biases, 32x32 / 64x32 tiles
struct TcChainMlp2W { const float *Vt[3], *ainv[3], *c[3]; };                           // one scale of mlp2: three 64x64 tiles, biases
// set-conv #1 up to the max over neighbours: out (bc*n, 256) = [scale0 64 | scale1 64 | scale2 64 | scale3 64]
int cmf_launch_setconv1_tc(int bc, int n, const float *xyz_planar, const float *ft_planar, const int *idx60, const TcChainSc1W *w4, float *out,
                           cudaStream_t st);
// out[row][s*64 + o] = mlp2_s(in[row][s*64 .. +63]) for the four scales
// amax_out (optional): per-pair atomicMax of the outputs (uint bit patterns, caller zeroes), rows_per_pair rows per frame pair
// gmax (optional): (pairs, 256) per-pair per-channel maximum over the pair's rows = the global max-pooled feature; the caller zeroes it
int cmf_launch_mlp2_tc(long long rows, const float *in, int ld_in, float *out, int ld_out, const TcChainMlp2W *w4, unsigned int *amax_out,
                       int rows_per_pair, cudaStream_t st, float *gmax = nullptr);
