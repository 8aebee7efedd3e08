// model_kernels.cuh -- launch prototypes of the model-path kernels (model_kernels.cu) used by engine.cu.
#pragma once
#include "cmf_common.cuh"

enum { CMF_ACT_NONE = 0, CMF_ACT_RELU = 1, CMF_ACT_LEAKY = 2 };

// Out[c][m] = act( sum_k W[m][k] * X[c][k] + bias[m] + pbias[c / cols_per_pair][m] ),  fp32 FMA.
// W (M x K) row-major with leading dimension ldw; X rows of K contiguous floats, leading dim ldx;
// Out rows of M floats, leading dim ldo.  K, ldw, ldx, ldo, M must be multiples of 4 and all base
// pointers 16-byte aligned.  bias / pbias may be NULL.
struct GemmArgs {
    const float *W; const float *X; float *Out; const float *bias; const float *pbias;
    int ldw, ldx, ldo, pb_ld, cols_per_pair;
    int M, K, cols, act;
};
struct GemmBatch { GemmArgs g[4]; int count; };
int cmf_launch_gemm(const GemmBatch &gb, cudaStream_t st);
int cmf_launch_gemm1(const GemmArgs &g, cudaStream_t st);
// the same for problems with few columns (one per frame pair): 64-output x 16-column CTAs; K % 64 == 0, no activation / per-pair bias
int cmf_launch_pair_gemv(const GemmBatch &gb, cudaStream_t st);

int cmf_launch_ball_query_ms(int b, int n, const float *xyz_planar, int *idx60, cudaStream_t st);
int cmf_launch_knn_point8(int b, int n_cand, int n_query, const float *cand_aos, const float *query_aos, int *idx, cudaStream_t st);
// both clouds' ball queries + point-major coordinate copies + (optional, cloud 1) E[:, off .. off+2] = ft, zero pad, per-pair |max|: one launch
struct SearchPrologueArgs {
    int n[2]; const float *xyz[2]; int *idx60[2]; float *aos[2];
    const float *ft; float *E; int lde, off, pad; unsigned int *amax_ft;
};
int cmf_launch_search_prologue(int b, const SearchPrologueArgs &a, cudaStream_t st);
// thread-per-query forms of the two search launches (cmf_search_small_ok: at least 128 blocks of 128 queries, clouds of at most 65535
// points); the k-NN reads the
// planar coordinates, so `aos` is not written.  Bit-identical results to the warp-cooperative kernels.
int cmf_search_small_ok(int b, int n, int n2);
int cmf_launch_search_prologue_small(int b, const SearchPrologueArgs &a, cudaStream_t st);
int cmf_launch_knn_point8_dual_small(int b, int n_query, const float *xyzq_planar, int n_cand0, const float *xyzc0_planar, int *idx0,
                                     int n_cand1, const float *xyzc1_planar, int *idx1, unsigned int *dirmax0, cudaStream_t st);
// the engine's two 8-NN searches of cloud-1 queries (against cloud 2 and against cloud 1) in one launch; dirmax (optional): per-pair
// atomicMax of |candidate - query| components over the FIRST search's neighbours (uint bit patterns; caller zeroes)
int cmf_launch_knn_point8_dual(int b, int n_query, const float *query_aos, int n_cand0, const float *cand0_aos, int *idx0,
                               int n_cand1, const float *cand1_aos, int *idx1, unsigned int *dirmax0, cudaStream_t st);

// mse_layer input rows: X0[scale s][(b*N+i)*K_s + kk][0..7] = [xyz_j - xyz_i, ft_j, 0, 0]
int cmf_launch_build_x0(int b, int n, const float *xyz_planar, const float *ft_planar, const int *idx60,
                        float *x0 /* 4 scale segments, see engine */, cudaStream_t st);
// fused set-conv #1 up to the max over neighbours: out (B*N, 256) = [scale0 64 | scale1 64 | scale2 64 | scale3 64]
int cmf_launch_setconv1_fused(int b, int n, const float *xyz_planar, const float *ft_planar, const int *idx60,
                              const float *const *seg12x4, float *out, cudaStream_t st);
// out[(b*N+i)*ldo + c] = max_kk Y[((b*N+i)*K + kk)*ldy + c], c < C (C % 4 == 0)
int cmf_launch_maxk(long long points, int K, int C, const float *Y, int ldy, float *out, int ldo, cudaStream_t st);
// G[b][c] = max_i F[(b*N+i)*ldf + c]
int cmf_launch_globalmax(int b, int n, int C, const float *F, int ldf, float *G, cudaStream_t st);
// amax_out (optional, in fc_reduce): per-pair atomicMax of |values written| (uint bit patterns; caller zeroes)

// flow embedding (FeatureCorrelator) pieces
int cmf_launch_fc_build_h1(int b, int n, int n2, const float *xyz1_planar, const float *xyz2_planar, const int *knn12,
                           const float *U1, const float *U2, const float *Wd /*512x4*/, float *H1, cudaStream_t st);
struct WeightNetP { const float *A1, *a1, *A2, *a2, *A3, *a3; };   // 8x4, 8, 8x8, 8, 512x8, 512
// n = points of the query cloud (rows of out), n_cand = points of the candidate cloud (xyzc, and src when gather = 1)
int cmf_launch_fc_reduce(int b, int n, int n_cand, const float *xyzq_planar, const float *xyzc_planar, const int *knn,
                         WeightNetP wn, const float *src, int gather /*0: src rows (b*N+i)*8+k ; 1: src rows b*N+j*/,
                         float *out, int ldo, cudaStream_t st, unsigned int *amax_out = nullptr);
// set-conv #2 first layer after hoisting: Y1[((b*N+i)*K + kk)][c] = relu(P[(b*N+j)*ldp + poff + c] + Wx[c][0..2] . rel)
int cmf_launch_mse2_build_y1(int b, int n, int K, int koff, const float *xyz_planar, const int *idx60,
                             const float *P, int ldp, int poff, const float *Wx /*512x4 rows for this scale*/,
                             float *Y1, cudaStream_t st);
// heads' last layer: flow (B,3,N) = W4f . h[:, 0:64] ; cls (B,N) = sigmoid(W4m . h[:, 64:128])
int cmf_launch_head_final(int b, int n, const float *H3, int ldh, const float *W4f, const float *W4m,
                          float *flow_planar, float *cls, cudaStream_t st);
// GRU gates (cmflow_t.py:101): gi, gh (B,768) pre-activations incl. biases; h_prev (B,256) or NULL -> h_new (B,256)
int cmf_launch_gru_gates(int b, const float *gi, const float *gh, const float *h_prev, float *h_new, cudaStream_t st);
// weighted Kabsch (+ optional refine). mode 0: W given (normalised); mode 1: score -> (s+eps)/sum
int cmf_launch_kabsch(int b, int n, const float *pc1, const float *pc_or_flow, int second_is_flow, const float *w,
                      int normalise, float eps, float stat_thres, float *trans, float *sf_agg, uint8_t *mask,
                      cudaStream_t st);

// RaFlow SFR module (models/raflow.py:79-117): flow (B,3,N) raw scene flow -> sf_agg, pre_trans (B,4,4), mask_s (B,N)
int cmf_launch_raflow_sfr(int b, int n, const float *pc1, const float *ft1, const float *flow, const float *interval,
                          float rigid_thres, float rigid_pcs, float *sf_agg, float *trans, uint8_t *mask, cudaStream_t st);

// fp16x3 mode: out[b] = max(out[b], bits(max |X[(b*N+i)*ld + c]|, c < width))  (uint bit patterns; caller zeroes)
int cmf_launch_pair_absmax(int b, int n, const float *X, int ld, int width, unsigned int *out, cudaStream_t st);
// out[b] = max over points i and their k neighbours j of max(|xc_j - xq_i| per component)
