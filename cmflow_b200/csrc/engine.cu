// engine.cu -- Part 3 of the C ABI: the whole-forward engine behind CMFlow.forward (models/cmflow.py:171-197)
// and CMFlow_T.forward (models/cmflow_t.py:185-211).
//
// The engine owns (a) the packed weights on the device, (b) one workspace arena sized for a CHUNK of frame
// pairs, and (c) the launch sequence.  Pairs are independent in eval mode (BatchNorm uses running statistics),
// so a batch is processed as consecutive chunks of pairs: the arena stays a few GB regardless of B and N,
// and a chunk's activations are mostly L2-resident between the kernels that produce and consume them.
//
// Algebra (exact in real arithmetic, differs from the reference only in fp32 summation order):
//  * BatchNorm (eval) is folded into the preceding 1x1 conv on the host in fp64 (cmflow_b200/weights.py).
//  * conv(cat[...]) is split by column blocks.  Blocks that multiply a per-cloud constant (the broadcast
//    global max feature, cmflow.py:76-81) become a per-pair bias; blocks that multiply a gathered neighbour
//    feature are applied ONCE PER POINT before the gather (the 1x1 conv commutes with the gather):
//       set-conv #2 layer 1:   8.1 GMAC/pair -> 0.41 GMAC/pair   (SURVEY.md section 7 "first-layer hoisting")
//       flow-embedding layer 1: 1.08 GMAC/pair -> 0.07 GMAC/pair
//  * mse_layer and mse_layer2 query the same cloud with the same radii (cmflow.py:21-24,35-39), so the
//    reference's 12 ball queries are 8, and each cloud's four radii are answered in one pass.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "model_kernels.cuh"
#include "tc_gemm.cuh"

namespace {

constexpr int HDR = 512;                 // header floats (int32 view): magic, nseg, temporal, then (off, rows, cols) per segment
constexpr int MAGIC = 0x434D4642;        // "CMFB"
constexpr int KS[4] = {4, 8, 16, 32};
constexpr int KOFF[4] = {0, 4, 12, 28};
constexpr int E_LD = 800;                // embedding row: [f1 256 | cor 512 | ft 3 | zero pad 29] (K padded to a multiple of 32 for the tensor-core path)

// ---- segment table (order is the contract with cmflow_b200/weights.py) ---------------------------
enum {
    M1_BASE = 0,          // 4 scales x 12: W1[32x8] B1 W2[32x32] B2 W3[64x32] B3 V1[64x64] C1 V2 C2 V3 C3
    FC_WC = 48, FC_WCG, FC_WN, FC_WNG, FC_WD, FC_B1, FC_W2, FC_B2, FC_W3, FC_B3,
    WN1_BASE = 58,        // A1[8x4] a1[8] A2[8x8] a2[8] A3[512x8] a3[512]
    WN2_BASE = 64,
    M2_WP = 70, M2_WG, M2_T1, M2_WX,
    M2_BASE = 74,         // 4 scales x 10: W2[256x512] T2 W3[64x256] T3 V1 C1 V2 C2 V3 C3
    HD_W1 = 114, HD_W1G, HD_T1, HD_W2F, HD_T2F, HD_W2M, HD_T2M, HD_W3F, HD_T3F, HD_W3M, HD_T3M, HD_W4,
    GRU_WIH = 126, GRU_WHH, GRU_BIH, GRU_BHH,
    NSEG_STATIC = 126, NSEG_TEMPORAL = 130
};

struct SegShape { int rows, cols; };

std::vector<SegShape> expected_shapes(int temporal) {
    std::vector<SegShape> v;
    for (int s = 0; s < 4; ++s) {
        v.push_back({32, 8}); v.push_back({1, 32}); v.push_back({32, 32}); v.push_back({1, 32});
        v.push_back({64, 32}); v.push_back({1, 64});
        for (int l = 0; l < 3; ++l) { v.push_back({64, 64}); v.push_back({1, 64}); }
    }
    v.push_back({512, 256}); v.push_back({512, 256}); v.push_back({512, 256}); v.push_back({512, 256});
    v.push_back({512, 4}); v.push_back({1, 512}); v.push_back({512, 512}); v.push_back({1, 512});
    v.push_back({512, 512}); v.push_back({1, 512});
    for (int w = 0; w < 2; ++w) {
        v.push_back({8, 4}); v.push_back({1, 8}); v.push_back({8, 8}); v.push_back({1, 8});
        v.push_back({512, 8}); v.push_back({1, 512});
    }
    v.push_back({2048, E_LD}); v.push_back({2048, 256}); v.push_back({1, 2048}); v.push_back({2048, 4});
    for (int s = 0; s < 4; ++s) {
        v.push_back({256, 512}); v.push_back({1, 256}); v.push_back({64, 256}); v.push_back({1, 64});
        for (int l = 0; l < 3; ++l) { v.push_back({64, 64}); v.push_back({1, 64}); }
    }
    v.push_back({512, 256}); v.push_back({512, 256}); v.push_back({1, 512});
    v.push_back({128, 256}); v.push_back({1, 128}); v.push_back({128, 256}); v.push_back({1, 128});
    v.push_back({64, 128}); v.push_back({1, 64}); v.push_back({64, 128}); v.push_back({1, 64});
    v.push_back({4, 64});
    if (temporal) { v.push_back({768, 256}); v.push_back({768, 256}); v.push_back({1, 768}); v.push_back({1, 768}); }
    return v;
}

size_t pad4(size_t x) { return (x + 3) & ~(size_t)3; }

struct Arena {
    char *base = nullptr;
    size_t off = 0;
    template <typename T> T *take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct Work {
    float *X1T, *X2T; int *BQ1, *BQ2, *KNN12, *KNN11;
    float *E, *F2, *G1, *G2;
    float *X0, *T32a, *T32b, *T64, *M64, *Q1, *Q2;
    float *PB1, *PB2, *U1, *U2, *H1, *H2, *COST1;
    float *PBM, *P, *Y1, *Y2, *Y3, *PROP, *GP, *GI, *GH, *GNEW, *ZERO;
    float *PBH, *HD1, *HD2, *HD3, *FLOW;
    unsigned int *AMAX;      // fp16x3 mode: per-pair max |x| of the tensors feeding a tensor-core GEMM (uint bit patterns of floats), AM_COUNT x bc
    float *SCL;              // fp16x3 mode: per-pair scales of the tiled intermediates (H2, Y2 x 4 scales), SC_COUNT x bc
};
enum { AM_F1 = 0, AM_F2, AM_E, AM_PROP, AM_U1, AM_U2, AM_DIR, AM_HD1, AM_HD2, AM_COR, AM_FT, AM_P0, AM_COUNT = AM_P0 + 4 };
enum { SC_H2 = 0, SC_Y2, SC_COUNT = SC_Y2 + 4 };

// What a mode does not touch is not allocated: the strict-fp32 pipeline materialises the gathered layer-1 tensors (Y1: 4.3 GB at 256 pairs of
// 256 points) that the tensor-core pipelines build inside their GEMM producers, and the unfused set-conv #1 scratch exists only for its A/B switch.
struct CarveOpts { int mode; bool unfused_sc1, need_h1, need_y2; };
void carve(Arena &a, Work &w, int bc, int n, const CarveOpts &o) {
    const size_t bn = (size_t)bc * n;
    const bool tc = o.mode != 0;
    w = Work{};
    w.X1T = a.take<float>(bn * 3); w.X2T = a.take<float>(bn * 3);
    w.BQ1 = a.take<int>(bn * 60); w.BQ2 = a.take<int>(bn * 60);
    w.KNN12 = a.take<int>(bn * 8); w.KNN11 = a.take<int>(bn * 8);
    w.E = a.take<float>(bn * E_LD); w.F2 = a.take<float>(bn * 256);
    if (o.unfused_sc1) {
        w.X0 = a.take<float>(bn * 60 * 8); w.T32a = a.take<float>(bn * 60 * 32); w.T32b = a.take<float>(bn * 60 * 32);
        w.T64 = a.take<float>(bn * 60 * 64);
    }
    w.M64 = a.take<float>(bn * 256); w.Q1 = a.take<float>(bn * 256); w.Q2 = a.take<float>(bn * 256);
    w.PB1 = a.take<float>((size_t)bc * 512); w.PB2 = a.take<float>((size_t)bc * 512);
    w.U1 = a.take<float>(bn * 512); w.U2 = a.take<float>(bn * 512);
    if (o.need_h1) w.H1 = a.take<float>(bn * 8 * 512);
    // H2 / Y2: row-major fp32 (strict mode) or tiled hi/lo: 3xTF32 tiles hold two floats per element, 3xFP16 tiles two halfs
    const size_t h2_tiled = cmf_tc_act_tiled_floats((long long)bn * 8, 512), y2_tiled = cmf_tc_act_tiled_floats((long long)bn * 32, 256);
    w.H2 = a.take<float>(o.mode == 1 ? h2_tiled : (o.mode == 2 ? h2_tiled / 2 : bn * 8 * 512));
    w.COST1 = a.take<float>(bn * 512);
    w.PBM = a.take<float>((size_t)bc * 2048); w.P = a.take<float>(bn * 2048);
    if (!tc) { w.Y1 = a.take<float>(bn * 32 * 512); w.Y3 = a.take<float>(bn * 32 * 64); }
    if (o.need_y2) w.Y2 = a.take<float>(o.mode == 1 ? y2_tiled : (o.mode == 2 ? y2_tiled / 2 : bn * 32 * 256));
    w.PROP = a.take<float>(bn * 256);
    w.GI = a.take<float>((size_t)bc * 768); w.GH = a.take<float>((size_t)bc * 768);
    w.GNEW = a.take<float>((size_t)bc * 256); w.ZERO = a.take<float>((size_t)bc * 256);
    w.PBH = a.take<float>((size_t)bc * 512);
    w.HD1 = a.take<float>(bn * 512); w.HD2 = a.take<float>(bn * 256); w.HD3 = a.take<float>(bn * 128);
    w.FLOW = a.take<float>(bn * 3);
    // cleared together at the top of a forward (one memset): the scale maxima and the three global max-pooled features (atomicMax targets)
    w.AMAX = a.take<unsigned int>((size_t)AM_COUNT * bc);
    w.G1 = a.take<float>((size_t)bc * 256); w.G2 = a.take<float>((size_t)bc * 256); w.GP = a.take<float>((size_t)bc * 256);
    w.SCL = a.take<float>((size_t)SC_COUNT * bc);
}

}  // namespace

static unsigned long long *g_wd_host[64];          // per device: host-mapped watchdog record (tc_dev.cuh), installed by the first engine created there

struct cmf_model {
    int device = 0;              // ordinal the engine was created on: every entry point switches to it (and back) for its own duration
    int temporal = 0;
    float stat_thres = 0.5f;
    float *d_blob = nullptr;
    std::vector<const float *> seg;
    std::vector<SegShape> shape;
    char *ws = nullptr; size_t ws_bytes = 0; int cap_bc = 0, cap_n = 0, ws_mode = -1;    // arena carved for cap_bc pairs x cap_n points under carve signature ws_mode (mode + fusion switches)
    Work w{};
    // staging for the host entry points: two slots so that copies of neighbouring calls overlap the kernels (cmf_model_submit_host)
    struct HostSlot { float *d_in = nullptr; char *d_out = nullptr; size_t in_floats = 0, out_bytes = 0;
                      cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr; bool busy = false; } slots[2];
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    int launches = 0;
    int last_b = 0, last_n = 0;
    // host entry point: the kernel sequence of one (b, n, mode) forward between the staging buffers, captured once as a CUDA graph
    // (39 launches -> 1 cudaGraphLaunch: the host path synchronises every call, so launch latency is not hidden behind the GPU there)
    struct HostGraph { int slot, b, n, n2, mode, has_g; cudaGraphExec_t exec; bool failed; };
    std::vector<HostGraph> graphs;
    // tensor-core (tcgen05, 3xTF32) mode: pre-tiled hi/lo copies of the big weight matrices
    int tc = 0;
    int fused_sc1 = 1;           // CMF_FUSED_SC1=0 falls back to the unfused (GEMM-per-layer) set-conv #1 for A/B testing
    // pre-tiled weights per operand format (index 0: 3xTF32, 1: 3xFP16 with per-row scales a_inv)
    struct TcW { const float *wt = nullptr, *ainv = nullptr; };
    struct TcSet { float *buf = nullptr; TcW fc_wc, fc_wn, fc_w2, fc_w3, m2_wp, m2_w2[4], m2_w3[4], hd_w1;
                   TcW m1_w2[4], m1_w3[4], m1_v[4][3], m2_v[4][3];
                   TcW hd_w2; const float *hd_t2 = nullptr;
                   TcW hd_w3; const float *hd_t3 = nullptr; } tcw[2];      // heads layer 3: [W3F 0; 0 W3M] (128 x 256) and its stacked bias      // heads layer 2: [W2F 0; 0 W2M] (256 x 512) and its stacked bias      // narrow chains (tc_chain.cu, fmt 1 only)
    // RaFlow (models/raflow.py): same backbone, no motion head (its weights are zero in the blob), SFR module instead of the Kabsch head
    int raflow = 0; float rigid_thres = 0.15f, rigid_pcs = 0.25f;
    int chain = 1;               // CMF_CHAIN=0: keep the fp32 FMA kernels for set-conv #1 / mlp2 in fp16x3 mode (A/B testing)
    // host-side norms for the fp16x3 scale bounds: max row L1 of the rel-xyz / direction columns, max row L1 and max |bias| of the layers
    // whose outputs are written pre-split (flow-embedding conv1, set-conv #2 layer 2)
    float wd_l1 = 0.f, wx_l1[4] = {0.f, 0.f, 0.f, 0.f}, fc_w2_l1 = 0.f, fc_b2_max = 0.f, m2_w2_l1[4] = {0.f, 0.f, 0.f, 0.f}, m2_t2_max[4] = {0.f, 0.f, 0.f, 0.f};
    // profiling
    struct ProfRec { int cat; cudaEvent_t e0, e1; };
    int profiling = 0;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> pool; size_t pool_used = 0;
    double work[16] = {0}; int nlaunch[16] = {0}; float ms[16] = {0};
};

// Every allocation, weight-tiling kernel and launch of an engine belongs to the device it was created on.  A caller that holds the
// model on cuda:1 while cuda:0 is current (single-process multi-GPU) must not have them land on the wrong GPU.
struct DeviceScope {
    int prev = -1; bool switched = false;
    explicit DeviceScope(const cmf_model *m) {
        if (m && cudaGetDevice(&prev) == cudaSuccess && prev != m->device) switched = cudaSetDevice(m->device) == cudaSuccess;
    }
    ~DeviceScope() { if (switched) cudaSetDevice(prev); }
};

static void drop_graphs(cmf_model *m) {
    for (auto &g : m->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    m->graphs.clear();
}

static void prof_events(cmf_model *m, cudaEvent_t *a, cudaEvent_t *b) {
    while (m->pool.size() < m->pool_used + 2) { cudaEvent_t e; cudaEventCreate(&e); m->pool.push_back(e); }
    *a = m->pool[m->pool_used++]; *b = m->pool[m->pool_used++];
}

static GemmArgs mk(const float *W, int ldw, const float *X, int ldx, float *Out, int ldo, const float *bias,
                   int M, int K, long long cols, int act, const float *pbias = nullptr, int pb_ld = 0, int cpp = 1) {
    GemmArgs g;
    g.W = W; g.X = X; g.Out = Out; g.bias = bias; g.pbias = pbias;
    g.ldw = ldw; g.ldx = ldx; g.ldo = ldo; g.pb_ld = pb_ld; g.cols_per_pair = cpp;
    g.M = M; g.K = K; g.cols = (int)cols; g.act = act;
    return g;
}

// One launch: category (for the per-category device timers of cmf_model_read_profile) and its algorithmic
// work (FLOPs for the GEMM categories, bytes for the others).
#define RUN(cat_, work_, call)                                      \
    do {                                                            \
        cudaEvent_t e0_ = nullptr, e1_ = nullptr;                   \
        if (m->profiling) { prof_events(m, &e0_, &e1_); cudaEventRecord(e0_, st); } \
        int rc_ = (call);                                           \
        if (rc_ != CMF_OK) return rc_;                              \
        if (m->profiling) { cudaEventRecord(e1_, st); m->prof.push_back({(cat_), e0_, e1_}); } \
        m->work[(cat_)] += (double)(work_);                         \
        ++m->nlaunch[(cat_)];                                       \
        ++m->launches;                                              \
    } while (0)

enum { C_SEARCH = 0, C_GEMM_SC1, C_GEMM_FC_HOIST, C_GEMM_FC_MLP, C_GEMM_SC2_HOIST, C_GEMM_SC2_L2, C_GEMM_SC2_L3, C_GEMM_SC2_L2L3,
       C_GEMM_POINTWISE, C_GATHER, C_REDUCE, C_HEAD_KABSCH, C_COUNT };
static const char *const kCatNames[C_COUNT] = {"search", "gemm_setconv1", "gemm_flowembed_hoist", "gemm_flowembed_mlp",
    "gemm_setconv2_hoist", "gemm_setconv2_l2", "gemm_setconv2_l3", "gemm_setconv2_l2l3", "gemm_pointwise", "gather_build", "reduce", "head_kabsch"};
static double gflops(const GemmArgs &g) { return 2.0 * g.M * (double)g.K * g.cols; }
static double gflops(const GemmBatch &gb) { double f = 0; for (int i = 0; i < gb.count; ++i) f += gflops(gb.g[i]); return f; }

static TcArgs tc_plain(const cmf_model::TcW &W, int fmt, int M, int K, const float *X, int ldx, float *Out, int ldo, const float *bias,
                       long long cols, int act, const float *pbias, int pb_ld, long long cpp) {
    TcArgs a{};
    a.fmt = fmt; a.a_inv = fmt ? W.ainv : nullptr; a.amax_group = 1 << 30;
    a.Wt = W.wt; a.m_blocks = cmf_divup(M, 128); a.k_blocks = cmf_divup(K, 32); a.M = M; a.cols = cols;
    a.prod = TC_PROD_PLAIN; a.X = X; a.ldx = ldx;
    a.epi = TC_EPI_STORE; a.Out = Out; a.ldo = ldo; a.bias = bias; a.pbias = pbias; a.pb_ld = pb_ld; a.cols_per_pair = (int)cpp; a.act = act;
    a.ksamp = 1;
    return a;
}
// fp16x3: B-operand scale from a bound  const + sum coef_i * amax_i[pair]
static void tc_bound(TcArgs &a, float c0, const unsigned int *s0, float k0, const unsigned int *s1 = nullptr, float k1 = 0.f,
                     const unsigned int *s2 = nullptr, float k2 = 0.f) {
    if (!a.fmt) return;
    a.bs_mode = 1; a.bs_const = c0;
    a.bs_src[0] = reinterpret_cast<const float *>(s0); a.bs_coef[0] = k0;
    a.bs_src[1] = reinterpret_cast<const float *>(s1); a.bs_coef[1] = k1;
    a.bs_src[2] = reinterpret_cast<const float *>(s2); a.bs_coef[2] = k2;
}
static void tc_scaled(TcArgs &a, const float *scale) {          // fp16x3: B operand pre-split by the previous GEMM with per-pair scale[]
    if (!a.fmt) return;
    a.bs_mode = 2; a.bs_src[0] = scale;
}
static double tflops(const TcArgs &a, int K) { return 2.0 * a.M * (double)K * (double)a.cols; }

static int ensure_tc_weights(cmf_model *m, int fmt) {
    cmf_model::TcSet &S = m->tcw[fmt];
    if (S.buf) return CMF_OK;
    struct Item { cmf_model::TcW *dst; int seg, M, K, ldw; };
    std::vector<Item> items = {
        {&S.fc_wc, FC_WC, 512, 256, 256}, {&S.fc_wn, FC_WN, 512, 256, 256}, {&S.fc_w2, FC_W2, 512, 512, 512},
        {&S.fc_w3, FC_W3, 512, 512, 512}, {&S.m2_wp, M2_WP, 2048, E_LD, E_LD}, {&S.hd_w1, HD_W1, 512, 256, 256}};
    for (int s = 0; s < 4; ++s) {
        items.push_back({&S.m2_w2[s], M2_BASE + s * 10, 256, 512, 512});
        items.push_back({&S.m2_w3[s], M2_BASE + s * 10 + 2, 64, 256, 256});
    }
    if (fmt == 1)
        for (int s = 0; s < 4; ++s) {
            items.push_back({&S.m1_w2[s], M1_BASE + s * 12 + 2, 32, 32, 32});
            items.push_back({&S.m1_w3[s], M1_BASE + s * 12 + 4, 64, 32, 32});
            for (int l = 0; l < 3; ++l) {
                items.push_back({&S.m1_v[s][l], M1_BASE + s * 12 + 6 + l * 2, 64, 64, 64});
                items.push_back({&S.m2_v[s][l], M2_BASE + s * 10 + 4 + l * 2, 64, 64, 64});
            }
        }
    auto ainv_floats = [](int M) { return (size_t)cmf_divup(M, 128) * 128; };
    size_t tot = cmf_tc_tiled_floats(256, 512) + ainv_floats(256) + 256            // + block-diagonal heads layer 2
               + cmf_tc_tiled_floats(128, 256) + ainv_floats(128) + 128;           // + block-diagonal heads layer 3
    for (auto &it : items) tot += cmf_tc_tiled_floats(it.M, it.K) + ainv_floats(it.M);
    cudaError_t e = cudaMalloc(&S.buf, tot * sizeof(float));
    if (e != cudaSuccess) { S.buf = nullptr; cmf_set_error("tc weights cudaMalloc failed: %s", cudaGetErrorString(e)); return CMF_ERR_NOMEM; }
    size_t off = 0;
    for (auto &it : items) {
        float *wt = S.buf + off, *ainv = wt + cmf_tc_tiled_floats(it.M, it.K);
        int rc = fmt ? cmf_tc_tile_weights_f16(m->seg[it.seg], it.ldw, it.M, it.K, wt, ainv, 0)
                     : cmf_tc_tile_weights(m->seg[it.seg], it.ldw, it.M, it.K, wt, 0);
        if (rc) return rc;
        it.dst->wt = wt; it.dst->ainv = fmt ? ainv : nullptr;
        off += cmf_tc_tiled_floats(it.M, it.K) + ainv_floats(it.M);
    }
    {   // heads layer 2 (radarflow_util.py:246,274): the flow and motion branches read disjoint halves of the stacked layer-1 output, so
        // both are one GEMM with the block-diagonal matrix [W2F 0; 0 W2M] (M = 256 suits the CTA-pair kernel; half of its MMAs multiply zeros)
        float *tmp = nullptr;
        CMF_CUDA(cudaMalloc(&tmp, (size_t)256 * 512 * sizeof(float)));
        CMF_CUDA(cudaMemset(tmp, 0, (size_t)256 * 512 * sizeof(float)));
        CMF_CUDA(cudaMemcpy2D(tmp, 512 * sizeof(float), m->seg[HD_W2F], 256 * sizeof(float), 256 * sizeof(float), 128, cudaMemcpyDeviceToDevice));
        CMF_CUDA(cudaMemcpy2D(tmp + (size_t)128 * 512 + 256, 512 * sizeof(float), m->seg[HD_W2M], 256 * sizeof(float), 256 * sizeof(float), 128, cudaMemcpyDeviceToDevice));
        float *wt = S.buf + off, *ainv = wt + cmf_tc_tiled_floats(256, 512), *bias = ainv + ainv_floats(256);
        int rc = fmt ? cmf_tc_tile_weights_f16(tmp, 512, 256, 512, wt, ainv, 0) : cmf_tc_tile_weights(tmp, 512, 256, 512, wt, 0);
        if (rc) { cudaFree(tmp); return rc; }
        CMF_CUDA(cudaMemcpy(bias, m->seg[HD_T2F], 128 * sizeof(float), cudaMemcpyDeviceToDevice));
        CMF_CUDA(cudaMemcpy(bias + 128, m->seg[HD_T2M], 128 * sizeof(float), cudaMemcpyDeviceToDevice));
        CMF_CUDA(cudaDeviceSynchronize());
        cudaFree(tmp);
        S.hd_w2.wt = wt; S.hd_w2.ainv = fmt ? ainv : nullptr; S.hd_t2 = bias;
        off += cmf_tc_tiled_floats(256, 512) + ainv_floats(256) + 256;
    }
    {   // heads layer 3 (radarflow_util.py:247,275), the same way: [W3F 0; 0 W3M] (128 x 256) on the one-CTA tensor-core kernel
        float *tmp = nullptr;
        CMF_CUDA(cudaMalloc(&tmp, (size_t)128 * 256 * sizeof(float)));
        CMF_CUDA(cudaMemset(tmp, 0, (size_t)128 * 256 * sizeof(float)));
        CMF_CUDA(cudaMemcpy2D(tmp, 256 * sizeof(float), m->seg[HD_W3F], 128 * sizeof(float), 128 * sizeof(float), 64, cudaMemcpyDeviceToDevice));
        CMF_CUDA(cudaMemcpy2D(tmp + (size_t)64 * 256 + 128, 256 * sizeof(float), m->seg[HD_W3M], 128 * sizeof(float), 128 * sizeof(float), 64, cudaMemcpyDeviceToDevice));
        float *wt = S.buf + off, *ainv = wt + cmf_tc_tiled_floats(128, 256), *bias = ainv + ainv_floats(128);
        int rc = fmt ? cmf_tc_tile_weights_f16(tmp, 256, 128, 256, wt, ainv, 0) : cmf_tc_tile_weights(tmp, 256, 128, 256, wt, 0);
        if (rc) { cudaFree(tmp); return rc; }
        CMF_CUDA(cudaMemcpy(bias, m->seg[HD_T3F], 64 * sizeof(float), cudaMemcpyDeviceToDevice));
        CMF_CUDA(cudaMemcpy(bias + 64, m->seg[HD_T3M], 64 * sizeof(float), cudaMemcpyDeviceToDevice));
        CMF_CUDA(cudaDeviceSynchronize());
        cudaFree(tmp);
        S.hd_w3.wt = wt; S.hd_w3.ainv = fmt ? ainv : nullptr; S.hd_t3 = bias;
    }
    CMF_CUDA(cudaDeviceSynchronize());
    return CMF_OK;
}

static bool sc2_fused(const cmf_model *m) {              // CMF_SC2_FUSED=0: the two-kernel set-conv #2 of round 1 (A/B testing)
    const char *e = getenv("CMF_SC2_FUSED");
    return m->tc == 2 && cmf_tc_pair_enabled() && !(e && e[0] == '0');
}
static bool wsum_fused(const cmf_model *m) { return m->tc && cmf_tc_pair_enabled() && !getenv("CMF_NO_WSUM"); }
static CarveOpts carve_opts(const cmf_model *m) {
    const bool chain = m->tc == 2 && m->chain;
    return CarveOpts{m->tc, !chain && !m->fused_sc1, !wsum_fused(m), !sc2_fused(m)};
}
static int carve_sig(const cmf_model *m) {
    const CarveOpts o = carve_opts(m);
    return o.mode | (o.unfused_sc1 ? 8 : 0) | (o.need_h1 ? 16 : 0) | (o.need_y2 ? 32 : 0);
}
static size_t chunk_bytes(const cmf_model *m, int bc, int n) {
    Arena a; Work w;
    carve(a, w, bc, n, carve_opts(m));
    return a.off + 256;
}

static int ensure_workspace(cmf_model *m, int b, int n) {
    // chunk size: as many pairs as fit the workspace arena (default 64 GiB of the 180 GB, at most half of what is free; CMF_WS_GB / CMF_CHUNK_PAIRS override).
    // Few large chunks amortise the per-launch prologue of the persistent tensor-core kernels (cluster launch, TMEM allocation,
    // pipeline fill: ~40 us each, ~14 such launches per chunk).
    // `n` = the larger of the two clouds.  The arena is carved for (cap_bc pairs, cap_n points) and serves every call that needs no more:
    // the reference's evaluation loop feeds one pair at a time with a different point count per frame (dataset/vod.py:92-93,
    // main.py:203), which must not cost a cudaFree + cudaMalloc per frame.
    const bool env_chunk = getenv("CMF_CHUNK_PAIRS") != nullptr;
    if (m->ws && m->ws_mode != carve_sig(m)) { cudaFree(m->ws); m->ws = nullptr; m->ws_bytes = 0; m->cap_bc = m->cap_n = 0; drop_graphs(m); }   // carved for another mode
    if (m->ws && n <= m->cap_n && b <= m->cap_bc && !env_chunk) return CMF_OK;       // (also the path taken under stream capture)
    int n_cap = n;
    if (m->ws && n > m->cap_n && n < 2048) n_cap = (n + 127) & ~127;                  // growing point counts: leave headroom instead of growing again next frame
    if (n_cap < m->cap_n) n_cap = m->cap_n;
    size_t budget = (size_t)64 << 30;                 // of the 180 GB: one chunk for B=256 at N=256 and for B=64 at N=4096
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            free_b += m->ws_bytes;                    // what we hold is ours to re-use
            if (free_b / 2 < budget) budget = free_b / 2 > ((size_t)1 << 30) ? free_b / 2 : ((size_t)1 << 30);
        }
    }
    const char *wsenv = getenv("CMF_WS_GB");
    if (wsenv && atof(wsenv) > 0) budget = (size_t)(atof(wsenv) * (double)((size_t)1 << 30));
    int bc = b;
    const char *env = getenv("CMF_CHUNK_PAIRS");
    if (env && atoi(env) > 0) bc = atoi(env) < b ? atoi(env) : b;
    else {
        const size_t per = chunk_bytes(m, 1, n_cap);
        size_t fit = budget / per;
        if (fit < 1) fit = 1;
        if ((size_t)bc > fit) bc = (int)fit;
    }
    if (m->ws && n_cap <= m->cap_n && m->cap_bc >= bc) return CMF_OK;
    if (bc < m->cap_bc && n_cap == m->cap_n) bc = m->cap_bc;                           // never shrink the pair capacity at the same point count
    if (m->ws) { cudaFree(m->ws); m->ws = nullptr; m->ws_bytes = 0; }
    drop_graphs(m);                                   // captured graphs hold workspace pointers
    const size_t bytes = chunk_bytes(m, bc, n_cap);
    cudaError_t e = cudaMalloc(&m->ws, bytes);
    if (e != cudaSuccess) {
        cmf_set_error("cmf_model: workspace cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        m->cap_bc = m->cap_n = 0; m->ws_bytes = 0; m->ws = nullptr;
        return CMF_ERR_NOMEM;
    }
    m->ws_bytes = bytes; m->cap_bc = bc; m->cap_n = n_cap; m->ws_mode = carve_sig(m);
    Arena a; a.base = m->ws;
    carve(a, m->w, bc, n_cap, carve_opts(m));
    CMF_CUDA(cudaMemset(m->w.ZERO, 0, (size_t)bc * 256 * sizeof(float)));
    return CMF_OK;
}

// set-conv over a small-channel cloud (mse_layer, C=3): PointLocalFeature x 4 scales (radarflow_util.py:144-162)
static int run_mse_layer(cmf_model *m, int bc, int n, const float *pc, const float *ft, const int *bq,
                         float *dest, int ldd, float *G, cudaStream_t st, unsigned int *amax_dest = nullptr) {
    Work &w = m->w;
    const long long bn = (long long)bc * n;
    GemmBatch gb;
    gb.count = 4;
    if (m->tc == 2 && m->chain) {
        // tensor-core chains (tc_chain.cu): gather + 6->32->32->64 + max over K in one launch, the three 64->64 layers of mlp2 in another
        const cmf_model::TcSet &T = m->tcw[1];
        TcChainSc1W cw[4]; TcChainMlp2W mw[4];
        for (int s = 0; s < 4; ++s) {
            const int sb = M1_BASE + s * 12;
            cw[s] = TcChainSc1W{m->seg[sb], m->seg[sb + 1], m->seg[sb + 3], m->seg[sb + 5], T.m1_w2[s].wt, T.m1_w2[s].ainv, T.m1_w3[s].wt, T.m1_w3[s].ainv};
            for (int l = 0; l < 3; ++l) { mw[s].Vt[l] = T.m1_v[s][l].wt; mw[s].ainv[l] = T.m1_v[s][l].ainv; mw[s].c[l] = m->seg[sb + 7 + l * 2]; }
        }
        RUN(C_GEMM_SC1, 2.0 * 3264.0 * 60.0 * (double)bn, cmf_launch_setconv1_tc(bc, n, pc, ft, bq, cw, w.M64, st));
        // (the global max over the points -- G, zeroed with the scale maxima at the top of the forward -- is taken in mlp2's epilogue)
        RUN(C_GEMM_POINTWISE, 2.0 * 3 * 4 * 64 * 64 * (double)bn, cmf_launch_mlp2_tc(bn, w.M64, 256, dest, ldd, mw, amax_dest, n, st, G));
        return CMF_OK;
    }
    if (m->fused_sc1) {
        // gather + 3-layer MLP + max over neighbours in one kernel (all four scales)
        const float *segs[24];
        for (int s = 0; s < 4; ++s)
            for (int k = 0; k < 6; ++k) segs[s * 6 + k] = m->seg[M1_BASE + s * 12 + k];
        RUN(C_GEMM_SC1, 2.0 * 3264.0 * 60.0 * (double)bn, cmf_launch_setconv1_fused(bc, n, pc, ft, bq, segs, w.M64, st));
    } else {
    RUN(C_GATHER, 0, cmf_launch_build_x0(bc, n, pc, ft, bq, w.X0, st));
    for (int s = 0; s < 4; ++s) {
        const int sb = M1_BASE + s * 12;
        gb.g[s] = mk(m->seg[sb + 0], 8, w.X0 + (size_t)bn * KOFF[s] * 8, 8, w.T32a + (size_t)bn * KOFF[s] * 32, 32,
                     m->seg[sb + 1], 32, 8, bn * KS[s], CMF_ACT_RELU);
    }
    RUN(C_GEMM_SC1, gflops(gb), cmf_launch_gemm(gb, st));
    for (int s = 0; s < 4; ++s) {
        const int sb = M1_BASE + s * 12;
        gb.g[s] = mk(m->seg[sb + 2], 32, w.T32a + (size_t)bn * KOFF[s] * 32, 32, w.T32b + (size_t)bn * KOFF[s] * 32, 32,
                     m->seg[sb + 3], 32, 32, bn * KS[s], CMF_ACT_RELU);
    }
    RUN(C_GEMM_SC1, gflops(gb), cmf_launch_gemm(gb, st));
    for (int s = 0; s < 4; ++s) {
        const int sb = M1_BASE + s * 12;
        gb.g[s] = mk(m->seg[sb + 4], 32, w.T32b + (size_t)bn * KOFF[s] * 32, 32, w.T64 + (size_t)bn * KOFF[s] * 64, 64,
                     m->seg[sb + 5], 64, 32, bn * KS[s], CMF_ACT_RELU);
    }
    RUN(C_GEMM_SC1, gflops(gb), cmf_launch_gemm(gb, st));
    for (int s = 0; s < 4; ++s)
        RUN(C_REDUCE, 0, cmf_launch_maxk(bn, KS[s], 64, w.T64 + (size_t)bn * KOFF[s] * 64, 64, w.M64 + s * 64, 256, st));
    }
    const float *src[3] = {w.M64, w.Q1, w.Q2};
    float *dst[3] = {w.Q1, w.Q2, dest};
    const int lds[3] = {256, 256, 256}, ldo[3] = {256, 256, ldd};
    for (int l = 0; l < 3; ++l) {
        for (int s = 0; s < 4; ++s) {
            const int sb = M1_BASE + s * 12 + 6 + l * 2;
            gb.g[s] = mk(m->seg[sb], 64, src[l] + s * 64, lds[l], dst[l] + s * 64, ldo[l], m->seg[sb + 1], 64, 64, bn, CMF_ACT_RELU);
        }
        RUN(C_GEMM_POINTWISE, gflops(gb), cmf_launch_gemm(gb, st));
    }
    RUN(C_REDUCE, 0, cmf_launch_globalmax(bc, n, 256, dest, ldd, G, st));
    return CMF_OK;
}

// n = points of cloud 1 (the queries; every output is per point of cloud 1), n2 = points of cloud 2.  The reference's evaluation loop
// feeds clouds of different sizes (dataset/vod.py:92-93 resamples to num_points only when training; main.py:203 evaluates one pair at a time).
static int forward_chunk(cmf_model *m, int bc, int n, int n2, const float *pc1, const float *pc2, const float *ft1,
                         const float *ft2, const float *gprev, float *sf_agg, float *stat_cls, float *pre_trans,
                         uint8_t *mask, float *gfeat_out, cudaStream_t st, const float *interval = nullptr, float *raw_flow = nullptr,
                         const float *label_m = nullptr) {
    Work &w = m->w;
    const long long bn = (long long)bc * n, bn2 = (long long)bc * n2;
    auto S = [&](int i) { return m->seg[i]; };
    const int F = m->tc == 2 ? 1 : 0;                                   // operand format of the tensor-core GEMMs (0: 3xTF32, 1: 3xFP16)
    const cmf_model::TcSet &T = m->tcw[F];
    auto AM = [&](int slot) { return w.AMAX + (size_t)slot * m->cap_bc; };
    auto SC = [&](int slot) { return w.SCL + (size_t)slot * m->cap_bc; };
    static const float RADII[4] = {2.f, 4.f, 8.f, 16.f};               // models/cmflow.py:21,35
    if (F) CMF_CUDA(cudaMemsetAsync(w.AMAX, 0, (size_t)(reinterpret_cast<char *>(w.GP + (size_t)m->cap_bc * 256) - reinterpret_cast<char *>(w.AMAX)), st));   // AMAX, G1, G2, GP

    // fp16x3 + chain kernels: the per-pair maxima behind the consumer GEMMs' fp16 scales are taken by the producing kernels' epilogues
    const bool fused_amax = F && m->chain;
    // neighbour search: two launches.  (1) both clouds' four-radius ball queries, the point-major coordinate copies the k-NN reads and the
    // radar-feature columns of E; (2) the cross-frame and the self 8-NN of cloud 1's points, with the per-pair direction maximum (fp16 scale bound)
    {
        SearchPrologueArgs sp;
        sp.n[0] = n; sp.n[1] = n2; sp.xyz[0] = pc1; sp.xyz[1] = pc2; sp.idx60[0] = w.BQ1; sp.idx60[1] = w.BQ2; sp.aos[0] = w.X1T; sp.aos[1] = w.X2T;
        sp.ft = ft1; sp.E = w.E; sp.lde = E_LD; sp.off = 768; sp.pad = E_LD - 771; sp.amax_ft = fused_amax ? AM(AM_FT) : nullptr;
        // batches of radar-sized clouds: one thread per query (the candidates of a pair fit in shared memory); dense clouds and calls with a
        // handful of pairs: warp-cooperative kernels.  CMF_SEARCH_WARP / CMF_SEARCH_THREAD force one or the other (tests)
        const bool small = (cmf_search_small_ok(bc, n, n2) || getenv("CMF_SEARCH_THREAD")) && n <= 65535 && n2 <= 65535 && !getenv("CMF_SEARCH_WARP");
        RUN(C_SEARCH, 0, small ? cmf_launch_search_prologue_small(bc, sp, st) : cmf_launch_search_prologue(bc, sp, st));
        if (small) RUN(C_SEARCH, 0, cmf_launch_knn_point8_dual_small(bc, n, pc1, n2, pc2, w.KNN12, n, pc1, w.KNN11, F ? AM(AM_DIR) : nullptr, st));
        else RUN(C_SEARCH, 0, cmf_launch_knn_point8_dual(bc, n, w.X1T, n2, w.X2T, w.KNN12, n, w.X1T, w.KNN11, F ? AM(AM_DIR) : nullptr, st));
    }

    // multi-scale encoders (cmflow.py:72-77); cloud 1 writes straight into the embedding rows E[:, 0:256]
    { int rc = run_mse_layer(m, bc, n, pc1, ft1, w.BQ1, w.E, E_LD, w.G1, st, fused_amax ? AM(AM_F1) : nullptr); if (rc) return rc; }
    { int rc = run_mse_layer(m, bc, n2, pc2, ft2, w.BQ2, w.F2, 256, w.G2, st, fused_amax ? AM(AM_F2) : nullptr); if (rc) return rc; }

    // flow embedding (FeatureCorrelator, radarflow_util.py:185-237)
    bool fused_wsum = false;
    {   // per-pair bias vectors (the broadcast global-max features times their column blocks): flow embedding x2 and set-conv #2, one batched launch
        GemmBatch gb; gb.count = 3;
        gb.g[0] = mk(S(FC_WCG), 256, w.G1, 256, w.PB1, 512, S(FC_B1), 512, 256, bc, CMF_ACT_NONE);
        gb.g[1] = mk(S(FC_WNG), 256, w.G2, 256, w.PB2, 512, nullptr, 512, 256, bc, CMF_ACT_NONE);
        gb.g[2] = mk(S(M2_WG), 256, w.G1, 256, w.PBM, 2048, S(M2_T1), 2048, 256, bc, CMF_ACT_NONE);
        RUN(C_GEMM_FC_HOIST, gflops(gb), m->tc ? cmf_launch_pair_gemv(gb, st) : cmf_launch_gemm(gb, st));
    }
    if (m->tc) {
        if (F) {    // per-pair maxima of the GEMM inputs (fp16 scales): encoder features, kNN direction components
            if (!fused_amax) {
                RUN(C_REDUCE, 0, cmf_launch_pair_absmax(bc, n, w.E, E_LD, 256, AM(AM_F1), st));
                RUN(C_REDUCE, 0, cmf_launch_pair_absmax(bc, n2, w.F2, 256, 256, AM(AM_F2), st));
            }
        }
        {
            TcArgs ta_ = tc_plain(T.fc_wc, F, 512, 256, w.E, E_LD, w.U1, 512, nullptr, bn, CMF_ACT_NONE, w.PB1, 512, n);
            tc_bound(ta_, 0.f, AM(AM_F1), 1.f); if (F) { ta_.amax_out = AM(AM_U1); }
            RUN(C_GEMM_FC_HOIST, tflops(ta_, 256), cmf_launch_tc_auto(ta_, st));
        }
        {
            TcArgs ta_ = tc_plain(T.fc_wn, F, 512, 256, w.F2, 256, w.U2, 512, nullptr, bn2, CMF_ACT_NONE, w.PB2, 512, n2);
            tc_bound(ta_, 0.f, AM(AM_F2), 1.f); if (F) { ta_.amax_out = AM(AM_U2); }
            RUN(C_GEMM_FC_HOIST, tflops(ta_, 256), cmf_launch_tc_auto(ta_, st));
        }
        {   // conv1 with the gather + hoisted conv0 epilogue fused into the B-operand producer (no H1 round trip)
            TcArgs ta_ = tc_plain(T.fc_w2, F, 512, 512, nullptr, 0, w.H2, 512, S(FC_B2), bn * 8, CMF_ACT_LEAKY, nullptr, 0, (long long)n * 8);
            ta_.prod = TC_PROD_FC_H1; ta_.U1 = w.U1; ta_.U2 = w.U2; ta_.ld_u2 = 512; ta_.off_u2 = 0; ta_.Wsmall = S(FC_WD);
            ta_.xyz_q = pc1; ta_.xyz_c = pc2; ta_.nbr = w.KNN12; ta_.nbr_ld = 8; ta_.nbr_off = 0; ta_.ksamp = 8; ta_.n_pts = n; ta_.n_cand = n2;
            ta_.out_tiled = 1;                 // conv2's B operand is written split + swizzled, ready for a bulk copy
            // |leaky(U1[i] + U2[j] + Wd.dir)| <= max|U1| + max|U2| + max_c |Wd[c]|_1 * max|dir component|
            tc_bound(ta_, 0.f, AM(AM_U1), 1.f, AM(AM_U2), 1.f, AM(AM_DIR), m->wd_l1);
            ta_.out_mul = m->fc_w2_l1; ta_.out_add = m->fc_b2_max; ta_.out_scale_store = F ? SC(SC_H2) : nullptr;
            static long long *fc_dbg = nullptr;                      // CMF_FC_DBG=1: per-role wait cycles of this launch on stderr (diagnostic)
            const bool dbg = getenv("CMF_FC_DBG") != nullptr;
            if (dbg) { if (!fc_dbg) cudaMalloc(&fc_dbg, 512 * 8 * sizeof(long long)); cudaMemsetAsync(fc_dbg, 0, 512 * 8 * sizeof(long long), st); ta_.dbg = fc_dbg; }
            RUN(C_GEMM_FC_MLP, tflops(ta_, 512), cmf_launch_tc_auto(ta_, st));
            if (dbg) {
                std::vector<long long> h(512 * 8);
                cudaStreamSynchronize(st);
                cudaMemcpy(h.data(), fc_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
                double av[8] = {0}; int nb = 0;
                for (int b = 0; b < 512; b += 2) if (h[b * 8]) { ++nb; for (int k = 0; k < 8; ++k) av[k] += (double)h[b * 8 + k]; }
                if (nb) fprintf(stderr, "fc conv1 leaders=%d cycles: total %.0f | issuer tempty %.0f full %.0f peer_full %.0f | loader empty %.0f | producer empty %.0f | epilogue tfull %.0f | tiles %.0f\n",
                                nb, av[0] / nb, av[1] / nb, av[2] / nb, av[3] / nb, av[4] / nb, av[5] / nb, av[6] / nb, av[7] / nb);
            }
        }
        fused_wsum = wsum_fused(m);
        {
            TcArgs ta_ = tc_plain(T.fc_w3, F, 512, 512, nullptr, 0, fused_wsum ? w.COST1 : w.H1, 512, S(FC_B3), bn * 8, CMF_ACT_LEAKY, nullptr, 0, (long long)n * 8);
            ta_.prod = TC_PROD_TILED; ta_.Xt = w.H2;
            tc_scaled(ta_, SC(SC_H2));
            if (fused_wsum) {        // conv2 + LeakyReLU + WeightNet1 weighting + sum over the 8 neighbours in the TMEM epilogue
                ta_.epi = TC_EPI_WSUM; ta_.ksamp = 8; ta_.n_pts = n; ta_.n_cand = n2; ta_.xyz_q = pc1; ta_.xyz_c = pc2; ta_.nbr = w.KNN12; ta_.nbr_ld = 8; ta_.nbr_off = 0;
                ta_.wnA1 = S(WN1_BASE); ta_.wna1 = S(WN1_BASE + 1); ta_.wnA2 = S(WN1_BASE + 2); ta_.wna2 = S(WN1_BASE + 3);
                ta_.wnA3 = S(WN1_BASE + 4); ta_.wna3 = S(WN1_BASE + 5);
            }
            RUN(C_GEMM_FC_MLP, tflops(ta_, 512), cmf_launch_tc_auto(ta_, st));
        }
    } else {
    { const GemmArgs ga_ = mk(S(FC_WC), 256, w.E, E_LD, w.U1, 512, nullptr, 512, 256, bn, CMF_ACT_NONE, w.PB1, 512, n); RUN(C_GEMM_FC_HOIST, gflops(ga_), cmf_launch_gemm1(ga_, st)); }
    { const GemmArgs ga_ = mk(S(FC_WN), 256, w.F2, 256, w.U2, 512, nullptr, 512, 256, bn2, CMF_ACT_NONE, w.PB2, 512, n2); RUN(C_GEMM_FC_HOIST, gflops(ga_), cmf_launch_gemm1(ga_, st)); }
    RUN(C_GATHER, 0, cmf_launch_fc_build_h1(bc, n, n2, pc1, pc2, w.KNN12, w.U1, w.U2, S(FC_WD), w.H1, st));
    { const GemmArgs ga_ = mk(S(FC_W2), 512, w.H1, 512, w.H2, 512, S(FC_B2), 512, 512, bn * 8, CMF_ACT_LEAKY); RUN(C_GEMM_FC_MLP, gflops(ga_), cmf_launch_gemm1(ga_, st)); }
    { const GemmArgs ga_ = mk(S(FC_W3), 512, w.H2, 512, w.H1, 512, S(FC_B3), 512, 512, bn * 8, CMF_ACT_LEAKY); RUN(C_GEMM_FC_MLP, gflops(ga_), cmf_launch_gemm1(ga_, st)); }
    }
    WeightNetP wn1{S(WN1_BASE), S(WN1_BASE + 1), S(WN1_BASE + 2), S(WN1_BASE + 3), S(WN1_BASE + 4), S(WN1_BASE + 5)};
    WeightNetP wn2{S(WN2_BASE), S(WN2_BASE + 1), S(WN2_BASE + 2), S(WN2_BASE + 3), S(WN2_BASE + 4), S(WN2_BASE + 5)};
    if (!fused_wsum) RUN(C_REDUCE, 0, cmf_launch_fc_reduce(bc, n, n2, pc1, pc2, w.KNN12, wn1, w.H1, 0, w.COST1, 512, st));
    RUN(C_REDUCE, 0, cmf_launch_fc_reduce(bc, n, n, pc1, pc1, w.KNN11, wn2, w.COST1, 1, w.E + 256, E_LD, st, fused_amax ? AM(AM_COR) : nullptr));

    // set-conv #2 (mse_layer2, cmflow.py:87-89)
    if (m->tc) {
        if (F && !fused_amax) RUN(C_REDUCE, 0, cmf_launch_pair_absmax(bc, n, w.E, E_LD, E_LD, AM(AM_E), st));
        TcArgs ta_ = tc_plain(T.m2_wp, F, 2048, E_LD, w.E, E_LD, w.P, 2048, nullptr, bn, CMF_ACT_NONE, w.PBM, 2048, n);
        // E = [f1 | cor | ft]: max|E| <= max|f1| + max|cor| + max|ft| (a bound up to 3x loose costs < 2 of fp16's 5 exponent bits of headroom)
        if (fused_amax) tc_bound(ta_, 0.f, AM(AM_F1), 1.f, AM(AM_COR), 1.f, AM(AM_FT), 1.f);
        else tc_bound(ta_, 0.f, AM(AM_E), 1.f);
        if (F) { ta_.amax_out = AM(AM_P0); ta_.amax_group = 512; ta_.amax_ld = m->cap_bc; }     // one maximum per scale's 512 channels
        RUN(C_GEMM_SC2_HOIST, tflops(ta_, 771), cmf_launch_tc_auto(ta_, st));
    } else {
        const GemmArgs ga_ = mk(S(M2_WP), E_LD, w.E, E_LD, w.P, 2048, nullptr, 2048, E_LD, bn, CMF_ACT_NONE, w.PBM, 2048, n);
        RUN(C_GEMM_SC2_HOIST, 2.0 * 2048 * 771.0 * bn, cmf_launch_gemm1(ga_, st));
    }
    for (int s = 0; s < 4; ++s) {
        const int sb = M2_BASE + s * 10;
        if (m->tc) {
            {   // layer 2 (512->256): neighbour gather of the hoisted layer-1 rows + rel-xyz term + ReLU fused into the B producer
                TcArgs ta_ = tc_plain(T.m2_w2[s], F, 256, 512, nullptr, 0, w.Y2, 256, S(sb + 1), bn * KS[s], CMF_ACT_RELU, nullptr, 0, (long long)n * KS[s]);
                ta_.prod = TC_PROD_SC2_Y1; ta_.U1 = nullptr; ta_.U2 = w.P; ta_.ld_u2 = 2048; ta_.off_u2 = s * 512;
                ta_.Wsmall = S(M2_WX) + (size_t)s * 512 * 4; ta_.xyz_q = pc1; ta_.xyz_c = pc1; ta_.nbr = w.BQ1; ta_.nbr_ld = 60;
                ta_.nbr_off = KOFF[s]; ta_.ksamp = KS[s]; ta_.n_pts = n;
                ta_.out_tiled = 1;
                // |relu(P[j] + Wx.rel)| <= max|P (this scale)| + max_c |Wx[c]|_1 * radius   (every rel component is below the ball radius)
                tc_bound(ta_, m->wx_l1[s] * RADII[s], AM(AM_P0 + s), 1.f);
                ta_.out_mul = m->m2_w2_l1[s]; ta_.out_add = m->m2_t2_max[s]; ta_.out_scale_store = F ? SC(SC_Y2 + s) : nullptr;
                if (sc2_fused(m)) {      // layers 2 + 3 + max over K in one kernel: the 256-channel layer-2 output stays in tensor memory
                    static long long *dbg_buf = nullptr;                 // CMF_SC2_DBG=1: per-role wait cycles of the fused kernel on stderr (diagnostic)
                    const bool dbg = getenv("CMF_SC2_DBG") != nullptr;
                    if (dbg) { if (!dbg_buf) cudaMalloc(&dbg_buf, 512 * 16 * sizeof(long long)); cudaMemsetAsync(dbg_buf, 0, 512 * 16 * sizeof(long long), st); ta_.dbg = dbg_buf; }
                    RUN(C_GEMM_SC2_L2L3, tflops(ta_, 512) + 2.0 * 64 * 256.0 * (double)ta_.cols,
                        cmf_launch_sc2_fused(ta_, T.m2_w3[s].wt, T.m2_w3[s].ainv, S(sb + 3), w.M64 + s * 64, 256, st));
                    if (dbg) {
                        std::vector<long long> h(512 * 16);
                        cudaStreamSynchronize(st);
                        cudaMemcpy(h.data(), dbg_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
                        double av[16] = {0}; int nb = 0;
                        for (int b = 0; b < 512; b += 2) if (h[b * 16]) { ++nb; for (int k = 0; k < 16; ++k) av[k] += (double)h[b * 16 + k]; }
                        if (nb) {
                            fprintf(stderr, "sc2 fused K=%d leaders=%d cycles: total %.0f | issuer tempty %.0f operands %.0f | producer gather %.0f | epilogue tfull %.0f g1done %.0f d3full %.0f | producer empty %.0f\n",
                                    KS[s], nb, av[0] / nb, av[1] / nb, av[2] / nb, av[3] / nb, av[4] / nb, av[5] / nb, av[6] / nb, av[7] / nb);
                            fprintf(stderr, "    issuer: of the operand wait, %.0f cycles were for the peer's half after its own was complete\n", av[13] / nb);
                            if (av[9] > 0) fprintf(stderr, "    producer sections (SC2_PROD_PROFILE build): gather issue %.0f | loads + arithmetic %.0f | stores %.0f | proxy fence %.0f | arrive + bookkeeping %.0f\n",
                                                   av[8] / nb, av[9] / nb, av[10] / nb, av[11] / nb, av[12] / nb);
                        }
                    }
                    continue;
                }
                RUN(C_GEMM_SC2_L2, tflops(ta_, 512), cmf_launch_tc_auto(ta_, st));
            }
            {   // layer 3 (256->64) with ReLU + max over the K neighbours fused into the TMEM epilogue; B operand bulk-copied
                TcArgs ta_ = tc_plain(T.m2_w3[s], F, 64, 256, nullptr, 0, w.M64 + s * 64, 256, S(sb + 3), bn * KS[s], CMF_ACT_RELU, nullptr, 0, (long long)n * KS[s]);
                ta_.prod = TC_PROD_TILED; ta_.Xt = w.Y2;
                tc_scaled(ta_, SC(SC_Y2 + s));
                ta_.epi = TC_EPI_MAXK; ta_.ksamp = KS[s];
                RUN(C_GEMM_SC2_L3, tflops(ta_, 256), cmf_launch_tc_auto(ta_, st));
            }
        } else {
        RUN(C_GATHER, 0, cmf_launch_mse2_build_y1(bc, n, KS[s], KOFF[s], pc1, w.BQ1, w.P, 2048, s * 512, S(M2_WX) + (size_t)s * 512 * 4, w.Y1, st));
        { const GemmArgs ga_ = mk(S(sb), 512, w.Y1, 512, w.Y2, 256, S(sb + 1), 256, 512, bn * KS[s], CMF_ACT_RELU); RUN(C_GEMM_SC2_L2, gflops(ga_), cmf_launch_gemm1(ga_, st)); }
        { const GemmArgs ga_ = mk(S(sb + 2), 256, w.Y2, 256, w.Y3, 64, S(sb + 3), 64, 256, bn * KS[s], CMF_ACT_RELU); RUN(C_GEMM_SC2_L3, gflops(ga_), cmf_launch_gemm1(ga_, st)); }
        RUN(C_REDUCE, 0, cmf_launch_maxk(bn, KS[s], 64, w.Y3, 64, w.M64 + s * 64, 256, st));
        }
    }
    if (m->tc == 2 && m->chain) {
        TcChainMlp2W mw[4];
        for (int s = 0; s < 4; ++s)
            for (int l = 0; l < 3; ++l) { mw[s].Vt[l] = T.m2_v[s][l].wt; mw[s].ainv[l] = T.m2_v[s][l].ainv; mw[s].c[l] = S(M2_BASE + s * 10 + 5 + l * 2); }
        RUN(C_GEMM_POINTWISE, 2.0 * 3 * 4 * 64 * 64 * (double)bn, cmf_launch_mlp2_tc(bn, w.M64, 256, w.PROP, 256, mw, F ? AM(AM_PROP) : nullptr, n, st, w.GP));
    } else {
        GemmBatch gb; gb.count = 4;
        const float *src[3] = {w.M64, w.Q1, w.Q2};
        float *dst[3] = {w.Q1, w.Q2, w.PROP};
        for (int l = 0; l < 3; ++l) {
            for (int s = 0; s < 4; ++s) {
                const int sb = M2_BASE + s * 10 + 4 + l * 2;
                gb.g[s] = mk(S(sb), 64, src[l] + s * 64, 256, dst[l] + s * 64, 256, S(sb + 1), 64, 64, bn, CMF_ACT_RELU);
            }
            RUN(C_GEMM_POINTWISE, gflops(gb), cmf_launch_gemm(gb, st));
        }
    }
    if (!(m->tc == 2 && m->chain)) RUN(C_REDUCE, 0, cmf_launch_globalmax(bc, n, 256, w.PROP, 256, w.GP, st));
    const float *gvec = w.GP;
    if (m->temporal) {            // CMFlow-T: GRU over the global feature (cmflow_t.py:94-105)
        GemmBatch gg; gg.count = 2;
        gg.g[0] = mk(S(GRU_WIH), 256, w.GP, 256, w.GI, 768, S(GRU_BIH), 768, 256, bc, CMF_ACT_NONE);
        gg.g[1] = mk(S(GRU_WHH), 256, gprev ? gprev : w.ZERO, 256, w.GH, 768, S(GRU_BHH), 768, 256, bc, CMF_ACT_NONE);
        RUN(C_GEMM_POINTWISE, gflops(gg), m->tc ? cmf_launch_pair_gemv(gg, st) : cmf_launch_gemm(gg, st));
        float *gnew = gfeat_out ? gfeat_out : w.GNEW;
        RUN(C_HEAD_KABSCH, 0, cmf_launch_gru_gates(bc, w.GI, w.GH, gprev, gnew, st));
        gvec = gnew;
    }

    // heads (FlowHead / MotionHead, radarflow_util.py:240-285), first layer stacked [fp ; mp]
    { GemmBatch gh; gh.count = 1; gh.g[0] = mk(S(HD_W1G), 256, gvec, 256, w.PBH, 512, S(HD_T1), 512, 256, bc, CMF_ACT_NONE); RUN(C_GEMM_POINTWISE, gflops(gh), m->tc ? cmf_launch_pair_gemv(gh, st) : cmf_launch_gemm(gh, st)); }
    if (m->tc) {
        if (F && !fused_amax) RUN(C_REDUCE, 0, cmf_launch_pair_absmax(bc, n, w.PROP, 256, 256, AM(AM_PROP), st));
        TcArgs ta_ = tc_plain(T.hd_w1, F, 512, 256, w.PROP, 256, w.HD1, 512, nullptr, bn, CMF_ACT_RELU, w.PBH, 512, n);
        tc_bound(ta_, 0.f, AM(AM_PROP), 1.f); if (F) { ta_.amax_out = AM(AM_HD1); }
        RUN(C_GEMM_POINTWISE, tflops(ta_, 256), cmf_launch_tc_auto(ta_, st));
        TcArgs tb_ = tc_plain(T.hd_w2, F, 256, 512, w.HD1, 512, w.HD2, 256, T.hd_t2, bn, CMF_ACT_RELU, nullptr, 0, n);
        tc_bound(tb_, 0.f, AM(AM_HD1), 1.f); if (F) { tb_.amax_out = AM(AM_HD2); }
        RUN(C_GEMM_POINTWISE, tflops(tb_, 256), cmf_launch_tc_auto(tb_, st));
        TcArgs tc_ = tc_plain(T.hd_w3, F, 128, 256, w.HD2, 256, w.HD3, 128, T.hd_t3, bn, CMF_ACT_RELU, nullptr, 0, n);
        tc_bound(tc_, 0.f, AM(AM_HD2), 1.f);
        RUN(C_GEMM_POINTWISE, tflops(tc_, 128), cmf_launch_tc_auto(tc_, st));
    } else {
        const GemmArgs ga_ = mk(S(HD_W1), 256, w.PROP, 256, w.HD1, 512, nullptr, 512, 256, bn, CMF_ACT_RELU, w.PBH, 512, n);
        RUN(C_GEMM_POINTWISE, gflops(ga_), cmf_launch_gemm1(ga_, st));
    }
    {
        GemmBatch gb; gb.count = 2;
        if (!m->tc) {
            gb.g[0] = mk(S(HD_W2F), 256, w.HD1, 512, w.HD2, 256, S(HD_T2F), 128, 256, bn, CMF_ACT_RELU);
            gb.g[1] = mk(S(HD_W2M), 256, w.HD1 + 256, 512, w.HD2 + 128, 256, S(HD_T2M), 128, 256, bn, CMF_ACT_RELU);
            RUN(C_GEMM_POINTWISE, gflops(gb), cmf_launch_gemm(gb, st));
        }
        if (!m->tc) {
            gb.g[0] = mk(S(HD_W3F), 128, w.HD2, 256, w.HD3, 128, S(HD_T3F), 64, 128, bn, CMF_ACT_RELU);
            gb.g[1] = mk(S(HD_W3M), 128, w.HD2 + 128, 256, w.HD3 + 64, 128, S(HD_T3M), 64, 128, bn, CMF_ACT_RELU);
            RUN(C_GEMM_POINTWISE, gflops(gb), cmf_launch_gemm(gb, st));
        }
    }
    if (m->raflow) {       // FlowPredictor read-out into the caller's `output`, then the SFR module (raflow.py:79-117); w.FLOW takes the unused scores
        RUN(C_HEAD_KABSCH, 0, cmf_launch_head_final(bc, n, w.HD3, 128, S(HD_W4), S(HD_W4) + 192, raw_flow, w.FLOW, st));
        RUN(C_HEAD_KABSCH, 0, cmf_launch_raflow_sfr(bc, n, pc1, ft1, raw_flow, interval, m->rigid_thres, m->rigid_pcs, sf_agg, pre_trans, mask, st));
        return CMF_OK;
    }
    RUN(C_HEAD_KABSCH, 0, cmf_launch_head_final(bc, n, w.HD3, 128, S(HD_W4), S(HD_W4) + 192, w.FLOW, stat_cls, st));
    // ego-motion head + refinement (cmflow.py:96-125); CMFlow-T omits the +1e-4 (cmflow_t.py:119)
    // mode='train' with pseudo labels (cmflow.py:181-182): the labels take the scores' place in the weights AND in the refinement mask
    RUN(C_HEAD_KABSCH, 0, cmf_launch_kabsch(bc, n, pc1, w.FLOW, 1, label_m ? label_m : stat_cls, 1, m->temporal ? 0.f : 1e-4f, m->stat_thres, pre_trans, sf_agg, mask, st));
    return CMF_OK;
}

extern "C" size_t cmf_model_blob_floats(int temporal) {
    size_t tot = HDR;
    for (const SegShape &s : expected_shapes(temporal)) tot += pad4((size_t)s.rows * s.cols);
    return tot;
}

extern "C" int cmf_model_create(cmf_model **out, const float *blob, size_t blob_floats, int temporal, float stat_thres) {
    CMF_REQUIRE(out && blob, "null pointer");
    *out = nullptr;
    if (cmf_device_check() != CMF_OK) return CMF_ERR_STATE;
    const int *hdr = reinterpret_cast<const int *>(blob);
    const std::vector<SegShape> exp = expected_shapes(temporal);
    CMF_REQUIRE(blob_floats == cmf_model_blob_floats(temporal), "blob size does not match cmf_model_blob_floats()");
    CMF_REQUIRE(hdr[0] == MAGIC, "bad blob magic");
    CMF_REQUIRE(hdr[1] == (int)exp.size() && hdr[2] == (temporal ? 1 : 0), "blob segment count / model kind mismatch");
    cmf_model *m = new cmf_model();
    CMF_CUDA(cudaGetDevice(&m->device));
    m->temporal = temporal ? 1 : 0;
    m->stat_thres = stat_thres;
    cudaError_t e = cudaMalloc(&m->d_blob, blob_floats * sizeof(float));
    if (e != cudaSuccess) { delete m; cmf_set_error("cmf_model_create: cudaMalloc failed: %s", cudaGetErrorString(e)); return CMF_ERR_NOMEM; }
    e = cudaMemcpy(m->d_blob, blob, blob_floats * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(m->d_blob); delete m; cmf_set_error("cmf_model_create: cudaMemcpy failed: %s", cudaGetErrorString(e)); return CMF_ERR_CUDA; }
    size_t off = HDR;
    for (size_t i = 0; i < exp.size(); ++i) {
        const int so = hdr[3 + 3 * i], sr = hdr[4 + 3 * i], sc = hdr[5 + 3 * i];
        if ((size_t)so != off || sr != exp[i].rows || sc != exp[i].cols) {
            cmf_set_error("cmf_model_create: segment %zu is (off %d, %dx%d), expected (off %zu, %dx%d)", i, so, sr, sc, off, exp[i].rows, exp[i].cols);
            cudaFree(m->d_blob); delete m;
            return CMF_ERR_INVALID;
        }
        m->seg.push_back(m->d_blob + off);
        off += pad4((size_t)sr * sc);
    }
    m->shape = exp;
    {   // watchdog record of this device's tensor-core kernels: host-mapped, so that it can be read after a trap has poisoned the context
        int dev = m->device;
        if (dev >= 0 && dev < 64 && !g_wd_host[dev]) {
            unsigned long long *h = nullptr, *d = nullptr;
            if (cudaHostAlloc(&h, 4 * sizeof(unsigned long long), cudaHostAllocMapped) == cudaSuccess &&
                cudaHostGetDevicePointer(reinterpret_cast<void **>(&d), h, 0) == cudaSuccess) {
                h[0] = h[1] = h[2] = h[3] = 0ull;
                if (cmf_wd_set_tc_gemm(d) == CMF_OK && cmf_wd_set_tc_gemm2(d) == CMF_OK && cmf_wd_set_tc_sc2(d) == CMF_OK && cmf_wd_set_tc_chain(d) == CMF_OK)
                    g_wd_host[dev] = h;
            }
            cudaGetLastError();          // the record is a diagnostic: never fail the create over it
        }
    }
    {   // norms behind the fp16x3 scale bounds (host copy of the blob is still at hand)
        auto seg_host = [&](int i) { return blob + hdr[3 + 3 * i]; };
        auto max_row_l1 = [&](int i, int r0, int r1) {
            const float *W = seg_host(i); const int cols = exp[i].cols;
            double best = 0;
            for (int r = r0; r < r1; ++r) { double a = 0; for (int c = 0; c < cols; ++c) a += fabs((double)W[(size_t)r * cols + c]); if (a > best) best = a; }
            return (float)(best * 1.0001);
        };
        auto max_abs = [&](int i) { const float *W = seg_host(i); float b = 0.f; for (int k = 0; k < exp[i].rows * exp[i].cols; ++k) b = fmaxf(b, fabsf(W[k])); return b; };
        m->wd_l1 = max_row_l1(FC_WD, 0, 512);
        m->fc_w2_l1 = max_row_l1(FC_W2, 0, 512); m->fc_b2_max = max_abs(FC_B2);
        for (int s = 0; s < 4; ++s) {
            m->wx_l1[s] = max_row_l1(M2_WX, s * 512, (s + 1) * 512);
            m->m2_w2_l1[s] = max_row_l1(M2_BASE + s * 10, 0, 256); m->m2_t2_max[s] = max_abs(M2_BASE + s * 10 + 1);
        }
    }
    *out = m;
    { const char *ce = getenv("CMF_CHAIN"); if (ce && ce[0] == '0') m->chain = 0; }
    const char *fe = getenv("CMF_FUSED_SC1");
    if (fe && fe[0] == '0') m->fused_sc1 = 0;
    const char *env = getenv("CMF_MODE");            // "fp32" (default) | "tf32x3"
    if (env && (!strcmp(env, "tf32x3") || !strcmp(env, "fp16x3"))) {
        int rc = cmf_model_set_mode(m, !strcmp(env, "fp16x3") ? 2 : 1);
        if (rc) { cmf_model_destroy(m); *out = nullptr; return rc; }
    }
    return CMF_OK;
}

extern "C" void cmf_model_destroy(cmf_model *m) {
    if (!m) return;
    DeviceScope dev_(m);
    drop_graphs(m);
    if (m->ws) cudaFree(m->ws);
    if (m->d_blob) cudaFree(m->d_blob);
    for (auto &hs : m->slots) {
        if (hs.busy && hs.ev_out) cudaEventSynchronize(hs.ev_out);
        if (hs.d_in) cudaFree(hs.d_in);
        if (hs.d_out) cudaFree(hs.d_out);
        if (hs.ev_in) { cudaEventDestroy(hs.ev_in); cudaEventDestroy(hs.ev_done); cudaEventDestroy(hs.ev_out); }
    }
    if (m->s_h2d) { cudaStreamDestroy(m->s_h2d); cudaStreamDestroy(m->s_d2h); }
    for (int f = 0; f < 2; ++f) if (m->tcw[f].buf) cudaFree(m->tcw[f].buf);
    for (cudaEvent_t e : m->pool) cudaEventDestroy(e);
    delete m;
}

extern "C" size_t cmf_model_workspace_bytes(const cmf_model *m) { return m ? m->ws_bytes : 0; }
extern "C" int cmf_model_host_graphs(const cmf_model *m) {
    int n = 0;
    if (m) for (const auto &g : m->graphs) n += g.exec ? 1 : 0;
    return n;
}
extern "C" int cmf_model_launches_per_forward(const cmf_model *m) { return m ? m->launches : 0; }

// {0 = no watchdog fired on any device of this process | translation unit (1 tc_gemm, 2 tc_gemm2, 3 tc_sc2, 4 tc_chain), block << 32 | thread,
//  grid << 32 | block size, nanoseconds waited}: readable after the launch failure the trap causes
extern "C" int cmf_watchdog_read(unsigned long long *out4) {
    CMF_REQUIRE(out4, "null pointer");
    out4[0] = out4[1] = out4[2] = out4[3] = 0ull;
    for (int d = 0; d < 64; ++d)
        if (g_wd_host[d] && g_wd_host[d][0]) { for (int k = 0; k < 4; ++k) out4[k] = g_wd_host[d][k]; break; }
    return CMF_OK;
}

// Shared body of the device entry points.  n = points of cloud 1, n2 = points of cloud 2 (0 = same as n).
struct FwdArgs {
    const float *pc1, *pc2, *ft1, *ft2, *gfeat_prev, *label_m, *interval;
    float *sf_agg, *stat_cls, *pre_trans; uint8_t *mask; float *gfeat_out, *raw_flow;
};
static int forward_impl(cmf_model *m, int b, int n, int n2, const FwdArgs &f, void *stream) {
    CMF_REQUIRE(m, "null model");
    DeviceScope dev_(m);
    if (n2 == 0) n2 = n;
    CMF_REQUIRE(b >= 0 && n >= 0 && n2 >= 0, "negative size");
    if (b == 0) return CMF_OK;
    CMF_REQUIRE(n >= 8 && n2 >= 8, "need at least 8 points per cloud (knn_point(8, ...): torch.topk raises below that)");
    const int nmax = n > n2 ? n : n2;
    CMF_REQUIRE((long long)nmax * 32 * 512 < 2147483647LL / 2, "N too large for 32-bit column indices");
    CMF_REQUIRE(f.pc1 && f.pc2 && f.ft1 && f.ft2 && f.sf_agg && f.pre_trans && f.mask, "null pointer");
    if (m->raflow) CMF_REQUIRE(f.interval && f.raw_flow, "RaFlow engine: call cmf_model_forward_raflow (interval, output)");
    else {
        CMF_REQUIRE(!f.interval, "this engine is not a RaFlow engine: call cmf_model_set_raflow first");
        CMF_REQUIRE(f.stat_cls, "null pointer");
        CMF_REQUIRE(!m->temporal || f.gfeat_out, "CMFlow-T needs gfeat_out");
    }
    int rc = ensure_workspace(m, b, nmax);
    if (rc != CMF_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    m->launches = 0;
    for (int i = 0; i < 16; ++i) { m->work[i] = 0; m->nlaunch[i] = 0; }
    m->prof.clear(); m->pool_used = 0;
    m->last_b = b < m->cap_bc ? b : m->cap_bc; m->last_n = n;
    const size_t pn = (size_t)3 * n, pn2 = (size_t)3 * n2;
    for (int b0 = 0; b0 < b; b0 += m->cap_bc) {
        const int bc = (b - b0) < m->cap_bc ? (b - b0) : m->cap_bc;
        rc = forward_chunk(m, bc, n, n2, f.pc1 + b0 * pn, f.pc2 + b0 * pn2, f.ft1 + b0 * pn, f.ft2 + b0 * pn2,
                           f.gfeat_prev ? f.gfeat_prev + (size_t)b0 * 256 : nullptr,
                           f.sf_agg + b0 * pn, f.stat_cls ? f.stat_cls + (size_t)b0 * n : nullptr, f.pre_trans + (size_t)b0 * 16, f.mask + (size_t)b0 * n,
                           f.gfeat_out ? f.gfeat_out + (size_t)b0 * 256 : nullptr, st, f.interval ? f.interval + b0 : nullptr,
                           f.raw_flow ? f.raw_flow + b0 * pn : nullptr, f.label_m ? f.label_m + (size_t)b0 * n : nullptr);
        if (rc != CMF_OK) return rc;
    }
    return CMF_OK;
}

extern "C" int cmf_model_forward(cmf_model *m, int b, int n, const float *pc1, const float *pc2, const float *ft1,
                                 const float *ft2, const float *gfeat_prev, float *sf_agg, float *stat_cls,
                                 float *pre_trans, uint8_t *mask, float *gfeat_out, void *stream) {
    return forward_impl(m, b, n, n, FwdArgs{pc1, pc2, ft1, ft2, gfeat_prev, nullptr, nullptr, sf_agg, stat_cls, pre_trans, mask, gfeat_out, nullptr}, stream);
}

extern "C" int cmf_model_forward2(cmf_model *m, int b, int n1, int n2, const float *pc1, const float *pc2, const float *ft1,
                                  const float *ft2, const float *gfeat_prev, const float *label_m, float *sf_agg, float *stat_cls,
                                  float *pre_trans, uint8_t *mask, float *gfeat_out, void *stream) {
    return forward_impl(m, b, n1, n2, FwdArgs{pc1, pc2, ft1, ft2, gfeat_prev, label_m, nullptr, sf_agg, stat_cls, pre_trans, mask, gfeat_out, nullptr}, stream);
}

extern "C" int cmf_model_set_raflow(cmf_model *m, float rigid_thres, float rigid_pcs) {
    CMF_REQUIRE(m, "null model");
    CMF_REQUIRE(!m->temporal, "RaFlow has no temporal variant");
    m->raflow = 1; m->rigid_thres = rigid_thres; m->rigid_pcs = rigid_pcs;
    return CMF_OK;
}

extern "C" int cmf_model_forward_raflow(cmf_model *m, int b, int n, const float *pc1, const float *pc2, const float *ft1, const float *ft2,
                                        const float *interval, float *output, float *sf_agg, float *pre_trans, uint8_t *mask_s, void *stream) {
    CMF_REQUIRE(m && m->raflow, "call cmf_model_set_raflow first");
    CMF_REQUIRE(interval && output, "null pointer");
    return forward_impl(m, b, n, n, FwdArgs{pc1, pc2, ft1, ft2, nullptr, nullptr, interval, sf_agg, nullptr, pre_trans, mask_s, nullptr, output}, stream);
}

extern "C" int cmf_model_forward_raflow2(cmf_model *m, int b, int n1, int n2, const float *pc1, const float *pc2, const float *ft1, const float *ft2,
                                         const float *interval, float *output, float *sf_agg, float *pre_trans, uint8_t *mask_s, void *stream) {
    CMF_REQUIRE(m && m->raflow, "call cmf_model_set_raflow first");
    CMF_REQUIRE(interval && output, "null pointer");
    return forward_impl(m, b, n1, n2, FwdArgs{pc1, pc2, ft1, ft2, nullptr, nullptr, interval, sf_agg, nullptr, pre_trans, mask_s, nullptr, output}, stream);
}

// ---- host entry point -----------------------------------------------------------------------------------------------------------------
// Two staging slots (device input / output buffers, one copy stream each way, events): the H2D copy of call i+1 and the D2H copy of
// call i-1 overlap the kernels of call i when the caller uses the split submit / wait form (cmf_model_submit_host / cmf_model_wait_host);
// cmf_model_forward_host is submit + wait of one call.
static int host_slot_reserve(cmf_model *m, cmf_model::HostSlot &hs, size_t in_floats, size_t out_bytes) {
    if (hs.in_floats < in_floats || hs.out_bytes < out_bytes) drop_graphs(m);         // captured graphs hold staging pointers
    if (hs.in_floats < in_floats) {
        if (hs.d_in) cudaFree(hs.d_in);
        hs.d_in = nullptr; hs.in_floats = 0;
        CMF_CUDA(cudaMalloc(&hs.d_in, in_floats * sizeof(float)));
        hs.in_floats = in_floats;
    }
    if (hs.out_bytes < out_bytes) {
        if (hs.d_out) cudaFree(hs.d_out);
        hs.d_out = nullptr; hs.out_bytes = 0;
        CMF_CUDA(cudaMalloc(&hs.d_out, out_bytes));
        hs.out_bytes = out_bytes;
    }
    if (!hs.ev_in) {
        CMF_CUDA(cudaEventCreateWithFlags(&hs.ev_in, cudaEventDisableTiming));
        CMF_CUDA(cudaEventCreateWithFlags(&hs.ev_done, cudaEventDisableTiming));
        CMF_CUDA(cudaEventCreateWithFlags(&hs.ev_out, cudaEventDisableTiming));
    }
    return CMF_OK;
}

// The kernel sequence between the staging buffers of one slot: eager, or (CMF_HOST_GRAPH=1) captured once per (slot, shape, mode) and replayed.
static int host_run(cmf_model *m, int slot, int b, int n, int n2, const FwdArgs &f, cudaStream_t st) {
    const char *ge = getenv("CMF_HOST_GRAPH");
    const bool use_graph = ge && ge[0] == '1' && !m->profiling && !m->raflow;
    if (!use_graph) return forward_impl(m, b, n, n2, f, st);
    const int has_g = f.gfeat_prev ? 1 : 0;
    cmf_model::HostGraph *hg = nullptr;
    for (auto &g : m->graphs) if (g.slot == slot && g.b == b && g.n == n && g.n2 == n2 && g.mode == m->tc && g.has_g == has_g) hg = &g;
    if (!hg) {
        // first call of this shape: eager (one-time attribute settings, weight tiling and allocations happen here, outside any capture)
        int rc = forward_impl(m, b, n, n2, f, st);
        if (rc == CMF_OK) {
            if (m->graphs.size() >= 32) {              // bounded cache: forget the oldest shape
                if (m->graphs.front().exec) cudaGraphExecDestroy(m->graphs.front().exec);
                m->graphs.erase(m->graphs.begin());
            }
            m->graphs.push_back({slot, b, n, n2, m->tc, has_g, nullptr, false});
        }
        return rc;
    }
    if (!hg->exec && !hg->failed) {
        cudaGraph_t graph = nullptr;
        const int launches = m->launches;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            int rc = forward_impl(m, b, n, n2, f, st);
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc == CMF_OK && ce == cudaSuccess && graph) {
                if (cudaGraphInstantiate(&hg->exec, graph, 0) != cudaSuccess) hg->exec = nullptr;
            }
            if (graph) cudaGraphDestroy(graph);
            if (rc != CMF_OK) return rc;
        }
        if (!hg->exec) { cudaGetLastError(); hg->failed = true; m->launches = launches; }  // capture unavailable: the entry keeps its key and routes this shape to the eager path from now on
    }
    if (hg->exec) { CMF_CUDA(cudaGraphLaunch(hg->exec, st)); return CMF_OK; }
    return forward_impl(m, b, n, n2, f, st);
}

struct HostIo {
    const float *pc1, *pc2, *ft1, *ft2, *gfeat_prev;
    float *sf_agg, *stat_cls, *pre_trans; uint8_t *mask; float *gfeat_out;
};

static int host_submit(cmf_model *m, int slot, int b, int n, int n2, const HostIo &io, cudaStream_t st, bool pipelined) {
    cmf_model::HostSlot &hs = m->slots[slot];
    const size_t p1 = (size_t)b * 3 * n, p2 = (size_t)b * 3 * n2;
    const size_t in_floats = 2 * p1 + 2 * p2 + (size_t)b * 256;
    const size_t out_bytes = (p1 + (size_t)b * n + (size_t)b * 16 + (size_t)b * 256) * sizeof(float) + (size_t)b * n;
    int rc = host_slot_reserve(m, hs, in_floats, out_bytes);
    if (rc != CMF_OK) return rc;
    float *d_pc1 = hs.d_in, *d_pc2 = d_pc1 + p1, *d_ft1 = d_pc2 + p2, *d_ft2 = d_ft1 + p1, *d_g = d_ft2 + p2;
    float *d_sf = reinterpret_cast<float *>(hs.d_out), *d_cls = d_sf + p1, *d_tr = d_cls + (size_t)b * n, *d_go = d_tr + (size_t)b * 16;
    uint8_t *d_mask = reinterpret_cast<uint8_t *>(d_go + (size_t)b * 256);
    cudaStream_t s_in = st, s_out = st;
    if (pipelined) {
        if (!m->s_h2d) { CMF_CUDA(cudaStreamCreateWithFlags(&m->s_h2d, cudaStreamNonBlocking)); CMF_CUDA(cudaStreamCreateWithFlags(&m->s_d2h, cudaStreamNonBlocking)); }
        s_in = m->s_h2d; s_out = m->s_d2h;
        // the slot's previous outputs must have left before its buffers are overwritten: the caller waited on that slot (cmf_model_wait_host)
    }
    CMF_CUDA(cudaMemcpyAsync(d_pc1, io.pc1, p1 * sizeof(float), cudaMemcpyHostToDevice, s_in));
    CMF_CUDA(cudaMemcpyAsync(d_pc2, io.pc2, p2 * sizeof(float), cudaMemcpyHostToDevice, s_in));
    CMF_CUDA(cudaMemcpyAsync(d_ft1, io.ft1, p1 * sizeof(float), cudaMemcpyHostToDevice, s_in));
    CMF_CUDA(cudaMemcpyAsync(d_ft2, io.ft2, p2 * sizeof(float), cudaMemcpyHostToDevice, s_in));
    if (io.gfeat_prev) CMF_CUDA(cudaMemcpyAsync(d_g, io.gfeat_prev, (size_t)b * 256 * sizeof(float), cudaMemcpyHostToDevice, s_in));
    if (pipelined) { CMF_CUDA(cudaEventRecord(hs.ev_in, s_in)); CMF_CUDA(cudaStreamWaitEvent(st, hs.ev_in, 0)); }
    rc = host_run(m, slot, b, n, n2, FwdArgs{d_pc1, d_pc2, d_ft1, d_ft2, io.gfeat_prev ? d_g : nullptr, nullptr, nullptr, d_sf, d_cls, d_tr, d_mask, d_go, nullptr}, st);
    if (rc != CMF_OK) return rc;
    if (pipelined) { CMF_CUDA(cudaEventRecord(hs.ev_done, st)); CMF_CUDA(cudaStreamWaitEvent(s_out, hs.ev_done, 0)); }
    CMF_CUDA(cudaMemcpyAsync(io.sf_agg, d_sf, p1 * sizeof(float), cudaMemcpyDeviceToHost, s_out));
    CMF_CUDA(cudaMemcpyAsync(io.stat_cls, d_cls, (size_t)b * n * sizeof(float), cudaMemcpyDeviceToHost, s_out));
    CMF_CUDA(cudaMemcpyAsync(io.pre_trans, d_tr, (size_t)b * 16 * sizeof(float), cudaMemcpyDeviceToHost, s_out));
    CMF_CUDA(cudaMemcpyAsync(io.mask, d_mask, (size_t)b * n, cudaMemcpyDeviceToHost, s_out));
    if (m->temporal && io.gfeat_out) CMF_CUDA(cudaMemcpyAsync(io.gfeat_out, d_go, (size_t)b * 256 * sizeof(float), cudaMemcpyDeviceToHost, s_out));
    CMF_CUDA(cudaEventRecord(hs.ev_out, s_out));
    hs.busy = true;
    return CMF_OK;
}

static int host_check(cmf_model *m, int b, int n, int n2, const HostIo &io) {
    CMF_REQUIRE(m, "null model");
    CMF_REQUIRE(!m->raflow, "the host entry points serve CMFlow / CMFlow-T engines");
    CMF_REQUIRE(b >= 0 && n >= 0 && n2 >= 0, "negative size");
    CMF_REQUIRE(b == 0 || (io.pc1 && io.pc2 && io.ft1 && io.ft2 && io.sf_agg && io.stat_cls && io.pre_trans && io.mask), "null pointer");
    return CMF_OK;
}

extern "C" int cmf_model_forward_host(cmf_model *m, int b, int n, const float *pc1, const float *pc2, const float *ft1,
                                      const float *ft2, const float *gfeat_prev, float *sf_agg, float *stat_cls,
                                      float *pre_trans, uint8_t *mask, float *gfeat_out, void *stream) {
    const HostIo io{pc1, pc2, ft1, ft2, gfeat_prev, sf_agg, stat_cls, pre_trans, mask, gfeat_out};
    int rc = host_check(m, b, n, n, io);
    if (rc != CMF_OK || b == 0) return rc;
    DeviceScope dev_(m);
    if (m->slots[0].busy) { CMF_CUDA(cudaEventSynchronize(m->slots[0].ev_out)); m->slots[0].busy = false; }
    rc = host_submit(m, 0, b, n, n, io, (cudaStream_t)stream, false);
    if (rc != CMF_OK) return rc;
    CMF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    m->slots[0].busy = false;
    return CMF_OK;
}

extern "C" int cmf_model_forward_host2(cmf_model *m, int b, int n1, int n2, const float *pc1, const float *pc2, const float *ft1,
                                       const float *ft2, const float *gfeat_prev, float *sf_agg, float *stat_cls,
                                       float *pre_trans, uint8_t *mask, float *gfeat_out, void *stream) {
    const HostIo io{pc1, pc2, ft1, ft2, gfeat_prev, sf_agg, stat_cls, pre_trans, mask, gfeat_out};
    if (n2 == 0) n2 = n1;
    int rc = host_check(m, b, n1, n2, io);
    if (rc != CMF_OK || b == 0) return rc;
    DeviceScope dev_(m);
    if (m->slots[0].busy) { CMF_CUDA(cudaEventSynchronize(m->slots[0].ev_out)); m->slots[0].busy = false; }
    rc = host_submit(m, 0, b, n1, n2, io, (cudaStream_t)stream, false);
    if (rc != CMF_OK) return rc;
    CMF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    m->slots[0].busy = false;
    return CMF_OK;
}

extern "C" int cmf_model_submit_host(cmf_model *m, int slot, int b, int n1, int n2, const float *pc1, const float *pc2, const float *ft1,
                                     const float *ft2, const float *gfeat_prev, float *sf_agg, float *stat_cls,
                                     float *pre_trans, uint8_t *mask, float *gfeat_out, void *stream) {
    const HostIo io{pc1, pc2, ft1, ft2, gfeat_prev, sf_agg, stat_cls, pre_trans, mask, gfeat_out};
    if (n2 == 0) n2 = n1;
    int rc = host_check(m, b, n1, n2, io);
    if (rc != CMF_OK || b == 0) return rc;
    CMF_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
    CMF_REQUIRE(!m->slots[slot].busy, "slot still in flight: call cmf_model_wait_host(slot) first");
    DeviceScope dev_(m);
    return host_submit(m, slot, b, n1, n2, io, (cudaStream_t)stream, true);
}

extern "C" int cmf_model_wait_host(cmf_model *m, int slot) {
    CMF_REQUIRE(m, "null model");
    CMF_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
    DeviceScope dev_(m);
    cmf_model::HostSlot &hs = m->slots[slot];
    if (!hs.busy) return CMF_OK;
    CMF_CUDA(cudaEventSynchronize(hs.ev_out));
    hs.busy = false;
    return CMF_OK;
}

extern "C" int cmf_model_set_mode(cmf_model *m, int mode) {
    CMF_REQUIRE(m, "null model");
    DeviceScope dev_(m);
    CMF_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (strict fp32 FMA), 1 (tcgen05 3xTF32) or 2 (tcgen05 3xFP16)");
    if (mode) { int rc = ensure_tc_weights(m, mode - 1); if (rc) return rc; }
    m->tc = mode;
    return CMF_OK;
}
extern "C" int cmf_model_get_mode(const cmf_model *m) { return m ? m->tc : -1; }

extern "C" int cmf_model_set_profiling(cmf_model *m, int enable) {
    CMF_REQUIRE(m, "null model");
    m->profiling = enable ? 1 : 0;
    return CMF_OK;
}
extern "C" int cmf_model_profile_categories(void) { return C_COUNT; }
extern "C" const char *cmf_model_profile_name(int cat) { return (cat >= 0 && cat < C_COUNT) ? kCatNames[cat] : ""; }

extern "C" int cmf_model_read_profile(cmf_model *m, float *ms, int *launches, double *work) {
    CMF_REQUIRE(m && ms && launches && work, "null pointer");
    DeviceScope dev_(m);
    for (int i = 0; i < C_COUNT; ++i) { ms[i] = 0.f; launches[i] = m->nlaunch[i]; work[i] = m->work[i]; }
    for (const auto &r : m->prof) {
        CMF_CUDA(cudaEventSynchronize(r.e1));
        float t = 0.f;
        CMF_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms[r.cat] += t;
    }
    return CMF_OK;
}

extern "C" const void *cmf_model_tap(const cmf_model *m, const char *name) {
    if (!m || !m->ws || !name) return nullptr;
    const Work &w = m->w;
    struct { const char *n; const void *p; } tab[] = {
        {"E", w.E}, {"f2", w.F2}, {"g1", w.G1}, {"g2", w.G2}, {"prop", w.PROP}, {"flow", w.FLOW},
        {"bq1", w.BQ1}, {"bq2", w.BQ2}, {"knn12", w.KNN12}, {"knn11", w.KNN11}, {"gp", w.GP}, {"P", w.P},
        {"cost1", w.COST1}, {"u1", w.U1}, {"u2", w.U2}, {"hd3", w.HD3}, {"m64", w.M64}};
    for (auto &t : tab)
        if (!strcmp(t.n, name)) return t.p;
    return nullptr;
}
