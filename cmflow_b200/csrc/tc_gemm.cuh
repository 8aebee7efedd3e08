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
    int n_pts, ksamp, nbr_ld, nbr_off, ld_u2, off_u2;            // points per cloud, neighbours per point, table row stride/offset, gathered-row stride/offset
    const float *Xt;                                             // TILED: activations already split + swizzled by a previous tc GEMM (out_tiled)
    // epilogue
    int epi;
    int out_tiled;                                               // STORE only: write [col_tile][k_block]{hi,lo} 256x32 swizzled tiles instead of rows
    float *Out; int ldo;
    const float *bias, *pbias; int pb_ld, cols_per_pair, act;
    // WSUM epilogue (flow embedding, radarflow_util.py:215-225): Out[point][m] = sum_k WeightNet(dir_ik)[m] * act(acc + bias)[column (point,k)]
    // WeightNet = 3 -> 8 -> 8 -> C, ReLU after every layer; dir = xyz_c[nbr] - xyz_q[point]; ksamp neighbours per point (pair kernel, TILED producer only)
    const float *wnA1, *wna1, *wnA2, *wna2, *wnA3, *wna3;
    // optional wait-time instrumentation (pair kernel): long long[gridDim.x][8] cycles = {total, mma:tempty, mma:full, mma:peer_full,
    // loader:empty, producer(warp 8):empty, epilogue(warp 4):tfull, tiles}; NULL in production
    long long *dbg;
};

size_t cmf_tc_tiled_floats(int M, int K);
size_t cmf_tc_act_tiled_floats(long long cols, int C);                          // floats of a tiled activation buffer (cols x C channels)                                        // floats needed for the pre-tiled copy of an M x K matrix
int cmf_tc_tile_weights(const float *W, int ldw, int M, int K, float *Wt, cudaStream_t st);
int cmf_launch_tc_gemm(const TcArgs &a, cudaStream_t st);      // one CTA per 128 x 256 tile
int cmf_launch_tc_gemm2(const TcArgs &a, cudaStream_t st);     // CTA pair (cta_group::2) per 256 x 256 tile; needs M % 256 == 0
int cmf_tc_pair_enabled();
int cmf_launch_tc_auto(const TcArgs &a, cudaStream_t st);      // pair kernel when M % 256 == 0 (unless CMF_TC2=0), else the one-CTA kernel
