// model_kernels.cu -- model-path kernels of the strict-fp32 pipeline (see DESIGN.md "Pipeline").
//
// Everything here is exact-fp32 (FMA) arithmetic; this is the parity build of the hot path.  The big
// 1x1-conv stacks go through one generic NT GEMM (activations are kept POINT-MAJOR, i.e. one row of C
// contiguous channels per point, so a neighbour gather is one coalesced 2 KB row read and GEMM outputs
// chain without transposes).
#include <float.h>
#include <limits.h>
#include <math.h>

#include "model_kernels.cuh"

// =================================================================================================
// Generic fp32 NT GEMM: Out[c][m] = act(sum_k W[m][k] X[c][k] + bias[m] + pbias[pair(c)][m])
// 256 threads, BM x 128 tile, BK = 16, 2-stage smem ring with register prefetch; thread (tm,tc) owns
// rows {tm*4..+3, 64+tm*4..+3} x cols {tc*4..+3, 64+tc*4..+3} (split-4 layout: conflict-free LDS.128,
// 256-byte coalesced stores along m).
// =================================================================================================
constexpr int G_BN = 128;
constexpr int G_BK = 16;

template <int BM>
__global__ void __launch_bounds__(256, 2)
gemm_nt_kernel(const GemmBatch gb) {
    const GemmArgs &g = gb.g[blockIdx.z];
    const int m0 = blockIdx.y * BM;
    const int c0 = blockIdx.x * G_BN;
    if (m0 >= g.M || c0 >= g.cols) return;
    constexpr int TMG = BM / 64;                  // groups of 4 rows per thread (2 for BM=128, 1 for BM=64)
    constexpr int TM = TMG * 4;
    __shared__ __align__(16) float As[2][G_BK][BM + 4];
    __shared__ __align__(16) float Bs[2][G_BK][G_BN + 4];

    const int tid = threadIdx.x;
    const int tm = tid & 15, tc = tid >> 4;
    const int lrow = tid >> 2, lkq = (tid & 3) * 4;     // loader mapping: 64 rows x 4 k-quads per pass

    // accumulators as float2 pairs of adjacent columns: one FFMA2 (Blackwell packed fp32 FMA, scalar operand broadcast) does two of the
    // tile's FMAs -- each half is an IEEE fma, so results are bit-identical with the scalar form at twice the FMA-pipe throughput
    float2 acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

    const int ktiles = (g.K + G_BK - 1) / G_BK;
    float4 ra[TMG], rb[2];

    auto gload = [&](int kt) {
        const int k = kt * G_BK + lkq;
        const bool kok = k < g.K;
#pragma unroll
        for (int r = 0; r < TMG; ++r) {
            const int m = m0 + lrow + 64 * r;
            ra[r] = (kok && m < g.M) ? __ldg(reinterpret_cast<const float4 *>(g.W + (size_t)m * g.ldw + k))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int c = c0 + lrow + 64 * r;
            rb[r] = (kok && c < g.cols) ? __ldg(reinterpret_cast<const float4 *>(g.X + (size_t)c * g.ldx + k))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int r = 0; r < TMG; ++r) {
            As[buf][lkq + 0][lrow + 64 * r] = ra[r].x; As[buf][lkq + 1][lrow + 64 * r] = ra[r].y;
            As[buf][lkq + 2][lrow + 64 * r] = ra[r].z; As[buf][lkq + 3][lrow + 64 * r] = ra[r].w;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            Bs[buf][lkq + 0][lrow + 64 * r] = rb[r].x; Bs[buf][lkq + 1][lrow + 64 * r] = rb[r].y;
            Bs[buf][lkq + 2][lrow + 64 * r] = rb[r].z; Bs[buf][lkq + 3][lrow + 64 * r] = rb[r].w;
        }
    };

    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < ktiles; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < ktiles) gload(kt + 1);
#pragma unroll
        for (int k = 0; k < G_BK; ++k) {
            float a[TM];
            float2 b[4];
#pragma unroll
            for (int r = 0; r < TMG; ++r) {
                const float4 v = *reinterpret_cast<const float4 *>(&As[buf][k][tm * 4 + 64 * r]);
                a[r * 4 + 0] = v.x; a[r * 4 + 1] = v.y; a[r * 4 + 2] = v.z; a[r * 4 + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float4 v = *reinterpret_cast<const float4 *>(&Bs[buf][k][tc * 4 + 64 * r]);
                b[r * 2 + 0] = make_float2(v.x, v.y); b[r * 2 + 1] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), b[j], acc[i][j]);
        }
        if (kt + 1 < ktiles) sstore(buf ^ 1);
        __syncthreads();
    }

    // epilogue
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + (j < 4 ? tc * 4 + j : 64 + tc * 4 + (j - 4));
        if (c >= g.cols) continue;
        const float *pb = g.pbias ? g.pbias + (size_t)(c / g.cols_per_pair) * g.pb_ld : nullptr;
#pragma unroll
        for (int r = 0; r < TMG; ++r) {
            const int m = m0 + 64 * r + tm * 4;
            if (m >= g.M) continue;
            float4 v = (j & 1) ? make_float4(acc[r * 4 + 0][j >> 1].y, acc[r * 4 + 1][j >> 1].y, acc[r * 4 + 2][j >> 1].y, acc[r * 4 + 3][j >> 1].y)
                               : make_float4(acc[r * 4 + 0][j >> 1].x, acc[r * 4 + 1][j >> 1].x, acc[r * 4 + 2][j >> 1].x, acc[r * 4 + 3][j >> 1].x);
            if (g.bias) {
                const float4 bb = __ldg(reinterpret_cast<const float4 *>(g.bias + m));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            }
            if (pb) {
                const float4 bb = __ldg(reinterpret_cast<const float4 *>(pb + m));
                v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            }
            if (g.act == CMF_ACT_RELU) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            } else if (g.act == CMF_ACT_LEAKY) {
                v.x = v.x > 0.f ? v.x : 0.1f * v.x; v.y = v.y > 0.f ? v.y : 0.1f * v.y;
                v.z = v.z > 0.f ? v.z : 0.1f * v.z; v.w = v.w > 0.f ? v.w : 0.1f * v.w;
            }
            *reinterpret_cast<float4 *>(g.Out + (size_t)c * g.ldo + m) = v;
        }
    }
}

int cmf_launch_gemm(const GemmBatch &gb, cudaStream_t st) {
    int maxM = 0, maxC = 0;
    for (int i = 0; i < gb.count; ++i) {
        const GemmArgs &g = gb.g[i];
        if ((g.K & 3) || (g.ldw & 3) || (g.ldx & 3) || (g.ldo & 3) || (g.M & 3)) {
            cmf_set_error("gemm: K/ld/M must be multiples of 4 (K=%d ldw=%d ldx=%d ldo=%d M=%d)", g.K, g.ldw, g.ldx, g.ldo, g.M);
            return CMF_ERR_INVALID;
        }
        if (g.M > maxM) maxM = g.M;
        if (g.cols > maxC) maxC = g.cols;
    }
    if (maxM == 0 || maxC == 0) return CMF_OK;
    if (maxM <= 64) {
        dim3 grid(cmf_divup(maxC, G_BN), cmf_divup(maxM, 64), gb.count);
        gemm_nt_kernel<64><<<grid, 256, 0, st>>>(gb);
    } else {
        dim3 grid(cmf_divup(maxC, G_BN), cmf_divup(maxM, 128), gb.count);
        gemm_nt_kernel<128><<<grid, 256, 0, st>>>(gb);
    }
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// ---- few-column GEMMs (one column per frame pair: the per-pair bias vectors, the GRU gate pre-activations) -------------------------------
// Out[c][m] = bias[m] + sum_k W[m][k] * X[c][k] for up to four problems in one launch.  With 256 columns the 128 x 128 tiles of gemm_nt_kernel
// give a few dozen CTAs (36 us for 0.4 GFLOP); here a CTA takes 64 outputs x 32 columns, K in chunks of 32 through shared memory (both
// chunks transposed; 4 x 4 register tiles; the next chunk's loads in flight under the arithmetic), so that a 3072-output batch is 384 CTAs.
// fp32 FMA, k ascending: deterministic.
constexpr int PG_O = 64, PG_C = 32, PG_K = 32, PG_T = 128;
__global__ void __launch_bounds__(PG_T)
pair_gemv_kernel(const GemmBatch gb, int tiles0, int tiles1, int tiles2) {
    __shared__ __align__(16) float sW[PG_K][PG_O + 4];                     // W chunk transposed: sW[k][o], rows of 68 floats (16-byte aligned float4 reads)
    __shared__ __align__(16) float sX[PG_K][PG_C + 4];                     // X chunk transposed: sX[k][c]
    int ot = blockIdx.x, seg = 0;
    if (ot >= tiles0) { ot -= tiles0; seg = 1; if (ot >= tiles1) { ot -= tiles1; seg = 2; if (ot >= tiles2) { ot -= tiles2; seg = 3; } } }
    const GemmArgs &g = gb.g[seg];
    const int o0 = ot * PG_O, c0 = blockIdx.y * PG_C;
    if (c0 >= g.cols) return;
    // this thread: outputs o0 + 4 * (tid & 15) .. + 3, columns c0 + 4 * (tid >> 4) .. + 3  (4 x 4 accumulators, two 128-bit shared loads per 16 FMAs)
    const int to = (threadIdx.x & 15) * 4, tc = (threadIdx.x >> 4) * 4;
    // loaders: W chunk = 64 rows x 32 k = 2048 floats = 16 per thread (row = idx >> 5, k = idx & 31: a warp reads one 128-byte row slice);
    //          X chunk = 32 cols x 32 k = 1024 floats = 8 per thread
    float wreg[16], xreg[8];
    auto fetch = [&](int kc) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int idx = threadIdx.x + i * PG_T, k = idx & (PG_K - 1), r = idx >> 5;
            wreg[i] = (o0 + r < g.M) ? __ldg(g.W + (size_t)(o0 + r) * g.ldw + kc + k) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = threadIdx.x + i * PG_T, k = idx & (PG_K - 1), c = idx >> 5;
            xreg[i] = (c0 + c < g.cols) ? __ldg(g.X + (size_t)(c0 + c) * g.ldx + kc + k) : 0.f;
        }
    };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    fetch(0);
    for (int kc = 0; kc < g.K; kc += PG_K) {
        __syncthreads();                                                   // the previous chunk has been consumed
#pragma unroll
        for (int i = 0; i < 16; ++i) { const int idx = threadIdx.x + i * PG_T; sW[idx & (PG_K - 1)][idx >> 5] = wreg[i]; }
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int idx = threadIdx.x + i * PG_T; sX[idx & (PG_K - 1)][idx >> 5] = xreg[i]; }
        __syncthreads();
        if (kc + PG_K < g.K) fetch(kc + PG_K);                             // the next chunk's global loads fly under this chunk's arithmetic
#pragma unroll 8
        for (int k = 0; k < PG_K; ++k) {
            const float4 w = *reinterpret_cast<const float4 *>(&sW[k][to]);
            const float4 x = *reinterpret_cast<const float4 *>(&sX[k][tc]);
            const float wv[4] = {w.x, w.y, w.z, w.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = c0 + tc + j;
        if (c >= g.cols) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int o = o0 + to + i;
            if (o < g.M) g.Out[(size_t)c * g.ldo + o] = acc[i][j] + (g.bias ? __ldg(g.bias + o) : 0.f);
        }
    }
}
int cmf_launch_pair_gemv(const GemmBatch &gb, cudaStream_t st) {
    int tiles[4] = {0, 0, 0, 0}, total = 0, maxC = 0;
    for (int i = 0; i < gb.count; ++i) {
        const GemmArgs &g = gb.g[i];
        if ((g.K % PG_K) || g.act != CMF_ACT_NONE || g.pbias) { cmf_set_error("pair_gemv: K %% 32 == 0, no activation, no per-pair bias (K=%d)", g.K); return CMF_ERR_INVALID; }
        tiles[i] = cmf_divup(g.M, PG_O); total += tiles[i];
        if (g.cols > maxC) maxC = g.cols;
    }
    if (total == 0 || maxC == 0) return CMF_OK;
    pair_gemv_kernel<<<dim3(total, cmf_divup(maxC, PG_C)), PG_T, 0, st>>>(gb, tiles[0], tiles[1], tiles[2]);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

int cmf_launch_gemm1(const GemmArgs &g, cudaStream_t st) {
    GemmBatch gb;
    gb.g[0] = g;
    gb.count = 1;
    return cmf_launch_gemm(gb, st);
}

// =================================================================================================
// small layout kernels
// =================================================================================================
// ---- multi-radius ball query: one pass over the candidates for all four CMFlow scales -------------
// (models/cmflow.py:21-22; semantics per scale = lib/src/ball_query_gpu.cu:9-45)
constexpr int MS_CHUNK = 2048;
constexpr int MS_QPW = 2;
__constant__ float c_ms_r2[4] = {4.0f, 16.0f, 64.0f, 256.0f};     // radius*radius for 2,4,8,16 (exact in fp32)

__device__ __forceinline__ void ball_query_ms_body(int b, int n, const float *__restrict__ xyz, int *__restrict__ idx60) {
    __shared__ float sx[MS_CHUNK], sy[MS_CHUNK], sz[MS_CHUNK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const float *px = xyz + (size_t)b * 3 * n, *py = px + n, *pz = py + n;
    const int q0 = (blockIdx.x * 8 + warp) * MS_QPW;
    constexpr int KS[4] = {4, 8, 16, 32};
    constexpr int OFF[4] = {0, 4, 12, 28};
    float qx[MS_QPW], qy[MS_QPW], qz[MS_QPW];
    int cnt[MS_QPW][4], first[MS_QPW][4];
    float rmax[MS_QPW];                     // r^2 of the largest radius whose row is still open (0 = all four rows full): the radii nest,
                                            // so a candidate outside it hits nothing and one ballot dismisses the whole batch of 32
    bool all_done = true;
#pragma unroll
    for (int t = 0; t < MS_QPW; ++t) {
        const int q = q0 + t;
        const bool valid = q < n;
        qx[t] = __ldg(px + (valid ? q : 0)); qy[t] = __ldg(py + (valid ? q : 0)); qz[t] = __ldg(pz + (valid ? q : 0));
#pragma unroll
        for (int s = 0; s < 4; ++s) { cnt[t][s] = valid ? 0 : KS[s]; first[t][s] = -1; }
        rmax[t] = valid ? c_ms_r2[3] : 0.f;
        all_done = all_done && !valid;
    }
    for (int base = 0; base < n; base += MS_CHUNK) {
        if (__syncthreads_and(all_done)) break;
        const int cn = min(MS_CHUNK, n - base);
        for (int i = threadIdx.x; i < cn; i += blockDim.x) {
            sx[i] = __ldg(px + base + i); sy[i] = __ldg(py + base + i); sz[i] = __ldg(pz + base + i);
        }
        __syncthreads();
        if (all_done) continue;
        for (int j = 0; j < cn; j += 32) {
            const int k = j + lane;
            const bool in = k < cn;
            const float cx = in ? sx[k] : 0.f, cy = in ? sy[k] : 0.f, cz = in ? sz[k] : 0.f;
#pragma unroll
            for (int t = 0; t < MS_QPW; ++t) {
                const float d2 = cmf_sqdist_ref(qx[t], qy[t], qz[t], cx, cy, cz);
                if (!__any_sync(0xffffffffu, in && d2 < rmax[t])) continue;      // nothing in this batch for any open row
                float open_r2 = 0.f;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    if (cnt[t][s] >= KS[s]) continue;               // warp-uniform
                    const bool hit = in && (d2 < c_ms_r2[s]);
                    const unsigned mask = __ballot_sync(0xffffffffu, hit);
                    if (mask) {
                        if (first[t][s] < 0) first[t][s] = base + j + __ffs(mask) - 1;
                        const int pos = cnt[t][s] + __popc(mask & lt);
                        if (hit && pos < KS[s]) idx60[((size_t)b * n + q0 + t) * 60 + OFF[s] + pos] = base + k;
                        cnt[t][s] += __popc(mask);
                    }
                    if (cnt[t][s] < KS[s]) open_r2 = c_ms_r2[s];    // ascending s: ends as the largest open radius
                }
                rmax[t] = open_r2;
            }
            bool open = false;
#pragma unroll
            for (int t = 0; t < MS_QPW; ++t) open = open || rmax[t] > 0.f;
            if (!open) { all_done = true; break; }
        }
    }
#pragma unroll
    for (int t = 0; t < MS_QPW; ++t) {
        if (q0 + t >= n) continue;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            int *row = idx60 + ((size_t)b * n + q0 + t) * 60 + OFF[s];
            // a cloud queried against itself always hits itself, but keep the reference's "no hit -> 0" rule
            const int fill = first[t][s] >= 0 ? first[t][s] : 0;
            const int from = first[t][s] >= 0 ? cnt[t][s] : 0;
            for (int l = from + lane; l < KS[s]; l += 32) row[l] = fill;
        }
    }
}
__global__ void __launch_bounds__(256)
ball_query_ms_kernel(int n, const float *__restrict__ xyz, int *__restrict__ idx60) { ball_query_ms_body(blockIdx.y, n, xyz, idx60); }

// The engine's search prologue in ONE launch (blockIdx.z = cloud): the multi-radius ball query of both clouds, plus -- for the 16 points a
// block queries -- the point-major copy of the coordinates the k-NN kernel reads and (cloud 1) the radar-feature columns of the embedding
// matrix E with their per-pair |max| (what transpose3 x2 / ball_query_ms x2 / scatter_ft did in five launches).
__global__ void __launch_bounds__(256)
search_prologue_kernel(const SearchPrologueArgs a) {
    const int z = blockIdx.z, b = blockIdx.y, n = z ? a.n[1] : a.n[0];         // (ternaries: no runtime indexing of the parameter struct)
    const int p0 = blockIdx.x * 8 * MS_QPW;
    if (p0 >= n) return;                                                      // block-uniform
    const float *xyz_z = z ? a.xyz[1] : a.xyz[0];
    const float *xyz = xyz_z + (size_t)b * 3 * n;
    float *aos = z ? a.aos[1] : a.aos[0];
    if (threadIdx.x < 3 * 8 * MS_QPW) {
        const int q = p0 + threadIdx.x / 3, comp = threadIdx.x % 3;
        if (q < n) aos[((size_t)b * n + q) * 3 + comp] = __ldg(xyz + (size_t)comp * n + q);
    } else if (z == 0 && a.E && threadIdx.x >= 64 && threadIdx.x < 64 + 8 * MS_QPW) {      // lanes 0..15 of warp 2
        const int i = p0 + (int)threadIdx.x - 64;
        float m = 0.f;
        if (i < n) {
            const float *p = a.ft + (size_t)b * 3 * n;
            float *o = a.E + ((size_t)b * n + i) * a.lde + a.off;
            const float f0 = __ldg(p + i), f1 = __ldg(p + n + i), f2 = __ldg(p + 2 * n + i);
            o[0] = f0; o[1] = f1; o[2] = f2;
            for (int d = 0; d < a.pad; ++d) o[3 + d] = 0.f;
            m = fmaxf(fabsf(f0), fmaxf(fabsf(f1), fabsf(f2)));
        }
        if (a.amax_ft) {
#pragma unroll
            for (int sft = 8; sft > 0; sft >>= 1) m = fmaxf(m, __shfl_xor_sync(0x0000ffffu, m, sft));
            if (threadIdx.x == 64 && m > 0.f) atomicMax(a.amax_ft + b, __float_as_uint(m));
        }
    }
    ball_query_ms_body(b, n, xyz_z, z ? a.idx60[1] : a.idx60[0]);
}
int cmf_launch_search_prologue(int b, const SearchPrologueArgs &a, cudaStream_t st) {
    const int nmax = a.n[0] > a.n[1] ? a.n[0] : a.n[1];
    if (b <= 0 || nmax <= 0) return CMF_OK;
    search_prologue_kernel<<<dim3(cmf_divup(nmax, 8 * MS_QPW), b, 2), 256, 0, st>>>(a);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// ---- thread-per-query searches (batches with enough queries to fill the chip; clouds of up to 65535 points) --------------------------
// With 256 candidates a warp-cooperative search spends ~700 warp instructions per query on ballots, shuffles and ordered insertion; one
// THREAD per query -- the reference's own decomposition (lib/src/ball_query_gpu.cu:9-45, radarflow_util.py:88-99) -- walks the candidates
// from shared memory (broadcast reads: every lane looks at the same candidate) at ~10-30 warp instructions per candidate for 32 queries.
// Same distance functions and the same ordering rules as the warp kernels (first hits in index order; ascending by (d, index) with the
// strict `<` of an insertion sort), so the results are bit-identical to them -- tests/test_gpu_pointops.py compares the two.
constexpr int ST_THREADS = 128, ST_CHUNK = 1024, ST_MAXN = 65535;      // candidates go through shared memory ST_CHUNK at a time; 16-bit row buffer

__global__ void __launch_bounds__(ST_THREADS)
search_prologue_thread_kernel(const SearchPrologueArgs a) {
    // dynamic shared memory: one chunk of candidates {x, y, z, -} (ST_CHUNK points or the whole cloud if smaller) and a 16-bit row buffer (indices
    // < 65536): a query's 60 neighbour slots, rows padded to 61 halfwords.  ~20 KB per block at 256 points: every block of a 256-pair batch is resident.
    extern __shared__ float4 st_smem[];
    const int z = blockIdx.z, b = blockIdx.y, n = z ? a.n[1] : a.n[0];
    const int nmax = a.n[0] > a.n[1] ? a.n[0] : a.n[1];
    const int chunk = nmax < ST_CHUNK ? nmax : ST_CHUNK;
    float4 *sc = st_smem;
    unsigned short *srow = reinterpret_cast<unsigned short *>(st_smem + chunk);
    const int q0 = blockIdx.x * ST_THREADS;
    if (q0 >= n) return;                                                        // block-uniform
    const float *px = (z ? a.xyz[1] : a.xyz[0]) + (size_t)b * 3 * n, *py = px + n, *pz = py + n;
    int *idx60 = z ? a.idx60[1] : a.idx60[0];
    const int q = q0 + threadIdx.x;
    const bool valid = q < n;
    if (z == 0 && a.E) {                                                        // cloud 1: the radar-feature columns of E and their per-pair |max|
        float m = 0.f;
        if (valid) {
            const float *p = a.ft + (size_t)b * 3 * n;
            float *o = a.E + ((size_t)b * n + q) * a.lde + a.off;
            const float f0 = __ldg(p + q), f1 = __ldg(p + n + q), f2 = __ldg(p + 2 * n + q);
            o[0] = f0; o[1] = f1; o[2] = f2;
            for (int d = 0; d < a.pad; ++d) o[3 + d] = 0.f;
            m = fmaxf(fabsf(f0), fmaxf(fabsf(f1), fabsf(f2)));
        }
        if (a.amax_ft) {
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sft));
            if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(a.amax_ft + b, __float_as_uint(m));
        }
    }
    constexpr int KS[4] = {4, 8, 16, 32};
    constexpr int OFF[4] = {0, 4, 12, 28};
    const float qx = valid ? __ldg(px + q) : 0.f, qy = valid ? __ldg(py + q) : 0.f, qz = valid ? __ldg(pz + q) : 0.f;
    int cnt[4], first[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) { cnt[s] = valid ? 0 : KS[s]; first[s] = -1; }
    unsigned short *row = srow + threadIdx.x * 61;
    bool done = !valid;
    for (int base = 0; base < n; base += chunk) {
        if (__syncthreads_and(done)) break;                                     // (also: everybody is past the previous chunk)
        const int cn = min(chunk, n - base);
        for (int i = threadIdx.x; i < cn; i += ST_THREADS) sc[i] = make_float4(__ldg(px + base + i), __ldg(py + base + i), __ldg(pz + base + i), 0.f);
        __syncthreads();
        for (int k = 0; k < cn; ++k) {
            const float4 c = sc[k];                                             // broadcast read
            const float d2 = cmf_sqdist_ref(qx, qy, qz, c.x, c.y, c.z);
            if (d2 < c_ms_r2[3]) {                                              // the radii nest: outside the largest one nothing hits
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    if (d2 < c_ms_r2[s] && cnt[s] < KS[s]) {
                        row[OFF[s] + cnt[s]] = (unsigned short)(base + k);
                        if (cnt[s] == 0) first[s] = base + k;
                        ++cnt[s];
                    }
            }
            if ((k & 31) == 31) {
                done = cnt[0] >= KS[0] && cnt[1] >= KS[1] && cnt[2] >= KS[2] && cnt[3] >= KS[3];
                if (__all_sync(0xffffffffu, done)) break;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        // a cloud queried against itself always hits itself, but keep the reference's "no hit -> 0" rule
        const int fill = first[s] >= 0 ? first[s] : 0;
        for (int l = first[s] >= 0 ? cnt[s] : 0; l < KS[s]; ++l) row[OFF[s] + l] = (unsigned short)fill;
    }
    __syncthreads();
    const int rows = min(ST_THREADS, n - q0);
    int *dst = idx60 + ((size_t)b * n + q0) * 60;
    for (int i = threadIdx.x; i < rows * 60; i += ST_THREADS) dst[i] = (int)srow[(i / 60) * 61 + i % 60];      // coalesced
}

// the cross-frame and the self 8-NN of cloud 1's points (blockIdx.z), one thread per query; coordinates in the reference's planar layout
__global__ void __launch_bounds__(ST_THREADS)
knn_point8_thread_kernel(int nq, const float *__restrict__ xyzq, int mc0, const float *__restrict__ xyzc0, int *__restrict__ idx0,
                         int mc1, const float *__restrict__ xyzc1, int *__restrict__ idx1, unsigned int *__restrict__ dirmax0) {
    extern __shared__ float4 st_smem[];                                         // one chunk of candidates {x, y, z, |x|^2}, sized by the launch
    float4 *sc = st_smem;
    const int z = blockIdx.z, b = blockIdx.y, mc = z ? mc1 : mc0;
    const int mcmax = mc0 > mc1 ? mc0 : mc1, chunk = mcmax < ST_CHUNK ? mcmax : ST_CHUNK;
    const float *pc = (z ? xyzc1 : xyzc0) + (size_t)b * 3 * mc;
    const int q = blockIdx.x * ST_THREADS + threadIdx.x;
    const bool valid = q < nq;
    const float *pq = xyzq + (size_t)b * 3 * nq;
    const float qx = __ldg(pq + (valid ? q : 0)), qy = __ldg(pq + nq + (valid ? q : 0)), qz = __ldg(pq + 2 * nq + (valid ? q : 0));
    const float nqn = cmf_sqnorm3(qx, qy, qz);
    float bd[8]; int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { bd[j] = INFINITY; bi[j] = INT_MAX; }
    for (int base = 0; base < mc; base += chunk) {
        __syncthreads();                                                        // everybody is past the previous chunk
        const int cn = min(chunk, mc - base);
        for (int i = threadIdx.x; i < cn; i += ST_THREADS) {
            const float x = __ldg(pc + base + i), y = __ldg(pc + mc + base + i), zz = __ldg(pc + 2 * mc + base + i);
            sc[i] = make_float4(x, y, zz, cmf_sqnorm3(x, y, zz));
        }
        __syncthreads();
        for (int i = 0; i < cn; ++i) {
            const float4 c = sc[i];
            float cd = cmf_sqdist_expanded(qx, qy, qz, nqn, c.x, c.y, c.z, c.w);
            if (cd < bd[7]) {                           // also rejects inf / NaN, like the reference's topk over finite values
                int ci = base + i;
                bool placed = false;                    // ordered insertion: behind every element with d' <= d (earlier index wins ties), then shift
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    placed = placed || (cd < bd[j]);
                    if (placed) { const float td = bd[j]; const int ti = bi[j]; bd[j] = cd; bi[j] = ci; cd = td; ci = ti; }
                }
            }
        }
    }
    if (valid) {
        int *o = (z ? idx1 : idx0) + ((size_t)b * nq + q) * 8;
        int r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = bi[j] == INT_MAX ? 0 : bi[j];        // unfilled slot: (1e40 -> inf, 0) in the reference
        reinterpret_cast<int4 *>(o)[0] = make_int4(r[0], r[1], r[2], r[3]);
        reinterpret_cast<int4 *>(o)[1] = make_int4(r[4], r[5], r[6], r[7]);
    }
    if (z == 0 && dirmax0) {                            // per-pair max |candidate - query| component over the neighbours found (fp16 scale bound)
        float mx = 0.f;
        if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int cj = bi[j] == INT_MAX ? 0 : bi[j];
                mx = fmaxf(mx, fmaxf(fabsf(__fsub_rn(__ldg(pc + cj), qx)), fmaxf(fabsf(__fsub_rn(__ldg(pc + mc + cj), qy)), fabsf(__fsub_rn(__ldg(pc + 2 * mc + cj), qz)))));
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(dirmax0 + b, __float_as_uint(mx));
    }
}
// Worth it when there are enough queries to fill the chip with 128-query blocks: one thread walks ALL candidates of its query, so with a
// handful of pairs (the reference's one-pair evaluation calls) the warp-cooperative kernels have the shorter critical path.
int cmf_search_small_ok(int b, int n, int n2) {
    const int nmax = n > n2 ? n : n2;
    return n <= ST_MAXN && n2 <= ST_MAXN && (long long)b * cmf_divup(nmax, ST_THREADS) >= 128;
}
int cmf_launch_search_prologue_small(int b, const SearchPrologueArgs &a, cudaStream_t st) {
    const int nmax = a.n[0] > a.n[1] ? a.n[0] : a.n[1];
    if (b <= 0 || nmax <= 0) return CMF_OK;
    if (nmax > ST_MAXN) { cmf_set_error("search prologue (thread per query): more than %d points", ST_MAXN); return CMF_ERR_INVALID; }
    const size_t smem = (size_t)(nmax < ST_CHUNK ? nmax : ST_CHUNK) * sizeof(float4) + (size_t)ST_THREADS * 61 * sizeof(unsigned short) + 16;
    search_prologue_thread_kernel<<<dim3(cmf_divup(nmax, ST_THREADS), b, 2), ST_THREADS, smem, st>>>(a);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}
int cmf_launch_knn_point8_dual_small(int b, int n_query, const float *xyzq_planar, int n_cand0, const float *xyzc0_planar, int *idx0,
                                     int n_cand1, const float *xyzc1_planar, int *idx1, unsigned int *dirmax0, cudaStream_t st) {
    if (b <= 0 || n_query <= 0) return CMF_OK;
    if (n_cand0 < 8 || n_cand1 < 8) { cmf_set_error("knn_point8_dual: fewer than 8 candidates (torch.topk raises too)"); return CMF_ERR_INVALID; }
    if (n_cand0 > ST_MAXN || n_cand1 > ST_MAXN) { cmf_set_error("knn (thread per query): more than %d candidates", ST_MAXN); return CMF_ERR_INVALID; }
    const int mcmax = n_cand0 > n_cand1 ? n_cand0 : n_cand1;
    const size_t smem = (size_t)(mcmax < ST_CHUNK ? mcmax : ST_CHUNK) * sizeof(float4);
    knn_point8_thread_kernel<<<dim3(cmf_divup(n_query, ST_THREADS), b, 2), ST_THREADS, smem, st>>>(n_query, xyzq_planar, n_cand0, xyzc0_planar, idx0,
                                                                                               n_cand1, xyzc1_planar, idx1, dirmax0);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

int cmf_launch_ball_query_ms(int b, int n, const float *xyz_planar, int *idx60, cudaStream_t st) {
    ball_query_ms_kernel<<<dim3(cmf_divup(n, 8 * MS_QPW), b), 256, 0, st>>>(n, xyz_planar, idx60);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}
extern "C" int cmf_ball_query_ms(int b, int n, const float *xyz_planar, int *idx60, void *stream) {
    CMF_REQUIRE(b >= 0 && n >= 0, "negative size");
    if (b == 0 || n == 0) return CMF_OK;
    CMF_REQUIRE(xyz_planar && idx60, "null pointer");
    CMF_REQUIRE(b <= 65535, "batch > 65535");
    return cmf_launch_ball_query_ms(b, n, xyz_planar, idx60, (cudaStream_t)stream);
}

int cmf_launch_knn_point8(int b, int n_cand, int n_query, const float *cand_aos, const float *query_aos, int *idx, cudaStream_t st) {
    return cmf_knn_point(b, n_cand, n_query, 8, cand_aos, query_aos, idx, nullptr, (void *)st);
}

// ---- mse_layer (C=3) input rows -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
build_x0_kernel(int bn_total, int n, const float *__restrict__ xyz, const float *__restrict__ ft,
                const int *__restrict__ idx60, float *__restrict__ x0) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)bn_total * 60) return;
    const int slot = (int)(t % 60);
    const int bi = (int)(t / 60);
    const int b = bi / n, i = bi - b * n;
    const int s = slot < 4 ? 0 : (slot < 12 ? 1 : (slot < 28 ? 2 : 3));
    const int off = s == 0 ? 0 : (s == 1 ? 4 : (s == 2 ? 12 : 28));
    const int K = 4 << s;
    const int kk = slot - off;
    const int j = __ldg(idx60 + (size_t)bi * 60 + slot);
    const float *px = xyz + (size_t)b * 3 * n, *pf = ft + (size_t)b * 3 * n;
    float4 v0, v1;
    v0.x = __fsub_rn(__ldg(px + j), __ldg(px + i));
    v0.y = __fsub_rn(__ldg(px + n + j), __ldg(px + n + i));
    v0.z = __fsub_rn(__ldg(px + 2 * n + j), __ldg(px + 2 * n + i));
    v0.w = __ldg(pf + j);
    v1.x = __ldg(pf + n + j); v1.y = __ldg(pf + 2 * n + j); v1.z = 0.f; v1.w = 0.f;
    float *row = x0 + ((size_t)bn_total * off + (size_t)bi * K + kk) * 8;
    reinterpret_cast<float4 *>(row)[0] = v0;
    reinterpret_cast<float4 *>(row)[1] = v1;
}
int cmf_launch_build_x0(int b, int n, const float *xyz_planar, const float *ft_planar, const int *idx60,
                        float *x0, cudaStream_t st) {
    const long long tot = (long long)b * n * 60;
    build_x0_kernel<<<cmf_divup(tot, 256), 256, 0, st>>>(b * n, n, xyz_planar, ft_planar, idx60, x0);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// ---- fused set-conv #1 (mse_layer, C=3): gather -> 6->32->32->64 (BN folded, ReLU) -> max over the K neighbours ----
// PointLocalFeature up to the max (radarflow_util.py:147-155) for ONE scale per blockIdx.y.  One thread owns two neighbour
// columns (c, c+128 of the CTA's 256), so every weight fetched from shared memory (128-bit broadcast) feeds 8 FMAs; the
// three layers stay in registers; the max over a point's K consecutive lanes is an xor-shuffle butterfly.  Replaces
// build_x0 + three GEMM launches + maxk and their 4.3 GB/cloud/step of activation round trips.
struct SetConv1W { const float *W1, *b1, *W2, *b2, *W3, *b3; };      // per scale: 32x8, 32, 32x32, 32, 64x32, 64
struct SetConv1Args { SetConv1W w[4]; };

__global__ void __launch_bounds__(128)
setconv1_fused_kernel(int n, const float *__restrict__ xyz, const float *__restrict__ ft, const int *__restrict__ idx60,
                      const SetConv1Args args, float *__restrict__ out /* (B*N, 256) */) {
    __shared__ __align__(16) float sW1[32 * 8], sW2[32 * 32], sW3[64 * 32], sb1[32], sb2[32], sb3[64];
    const int s = blockIdx.y, b = blockIdx.z;
    const int K = 4 << s, koff = (s == 0) ? 0 : (s == 1 ? 4 : (s == 2 ? 12 : 28));
    const int total_cols = n * K;
    if ((int)blockIdx.x * 256 >= total_cols) return;
    const SetConv1W &w = args.w[s];
    for (int i = threadIdx.x; i < 32 * 8; i += 128) sW1[i] = __ldg(w.W1 + i);
    for (int i = threadIdx.x; i < 32 * 32; i += 128) sW2[i] = __ldg(w.W2 + i);
    for (int i = threadIdx.x; i < 64 * 32; i += 128) sW3[i] = __ldg(w.W3 + i);
    if (threadIdx.x < 32) { sb1[threadIdx.x] = __ldg(w.b1 + threadIdx.x); sb2[threadIdx.x] = __ldg(w.b2 + threadIdx.x); }
    if (threadIdx.x < 64) sb3[threadIdx.x] = __ldg(w.b3 + threadIdx.x);
    __syncthreads();

    const float *px = xyz + (size_t)b * 3 * n, *pf = ft + (size_t)b * 3 * n;
    float x0[2][6];
    int pi[2]; bool ok[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = blockIdx.x * 256 + h * 128 + threadIdx.x;
        ok[h] = c < total_cols;
        const int i = ok[h] ? c / K : 0, kk = ok[h] ? c - i * K : 0;
        pi[h] = i;
        const int j = __ldg(idx60 + ((size_t)b * n + i) * 60 + koff + kk);
        x0[h][0] = __fsub_rn(__ldg(px + j), __ldg(px + i));
        x0[h][1] = __fsub_rn(__ldg(px + n + j), __ldg(px + n + i));
        x0[h][2] = __fsub_rn(__ldg(px + 2 * n + j), __ldg(px + 2 * n + i));
        x0[h][3] = __ldg(pf + j); x0[h][4] = __ldg(pf + n + j); x0[h][5] = __ldg(pf + 2 * n + j);
    }
    // the thread's two columns ride in the two halves of a float2: every weight (scalar, broadcast) feeds one FFMA2 = two FMAs
    float2 h1[32], h2[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        const float4 wa = *reinterpret_cast<const float4 *>(&sW1[o * 8]), wb = *reinterpret_cast<const float4 *>(&sW1[o * 8 + 4]);
        float2 a = make_float2(sb1[o], sb1[o]);
        a = __ffma2_rn(make_float2(wa.x, wa.x), make_float2(x0[0][0], x0[1][0]), a);
        a = __ffma2_rn(make_float2(wa.y, wa.y), make_float2(x0[0][1], x0[1][1]), a);
        a = __ffma2_rn(make_float2(wa.z, wa.z), make_float2(x0[0][2], x0[1][2]), a);
        a = __ffma2_rn(make_float2(wa.w, wa.w), make_float2(x0[0][3], x0[1][3]), a);
        a = __ffma2_rn(make_float2(wb.x, wb.x), make_float2(x0[0][4], x0[1][4]), a);
        a = __ffma2_rn(make_float2(wb.y, wb.y), make_float2(x0[0][5], x0[1][5]), a);
        h1[o] = make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
    }
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float2 a = make_float2(sb2[o], sb2[o]);
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
            const float4 wv = *reinterpret_cast<const float4 *>(&sW2[o * 32 + k]);
            a = __ffma2_rn(make_float2(wv.x, wv.x), h1[k], a); a = __ffma2_rn(make_float2(wv.y, wv.y), h1[k + 1], a);
            a = __ffma2_rn(make_float2(wv.z, wv.z), h1[k + 2], a); a = __ffma2_rn(make_float2(wv.w, wv.w), h1[k + 3], a);
        }
        h2[o] = make_float2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f));
    }
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int oc = 0; oc < 64; oc += 8) {
        float r[2][8];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            float2 a = make_float2(sb3[oc + o], sb3[oc + o]);
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
                const float4 wv = *reinterpret_cast<const float4 *>(&sW3[(oc + o) * 32 + k]);
                a = __ffma2_rn(make_float2(wv.x, wv.x), h2[k], a); a = __ffma2_rn(make_float2(wv.y, wv.y), h2[k + 1], a);
                a = __ffma2_rn(make_float2(wv.z, wv.z), h2[k + 2], a); a = __ffma2_rn(make_float2(wv.w, wv.w), h2[k + 3], a);
            }
            r[0][o] = fmaxf(a.x, 0.f); r[1][o] = fmaxf(a.y, 0.f);
        }
        // max over the K consecutive lanes of a point (K | 32, groups are lane-aligned because 128 and 256 are multiples of K)
        for (int off = 1; off < K; off <<= 1)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int o = 0; o < 8; ++o) r[h][o] = fmaxf(r[h][o], __shfl_xor_sync(0xffffffffu, r[h][o], off));
        if ((lane & (K - 1)) == 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (ok[h]) {
                    float *dst = out + ((size_t)b * n + pi[h]) * 256 + s * 64 + oc;
                    *reinterpret_cast<float4 *>(dst) = make_float4(r[h][0], r[h][1], r[h][2], r[h][3]);
                    *reinterpret_cast<float4 *>(dst + 4) = make_float4(r[h][4], r[h][5], r[h][6], r[h][7]);
                }
        }
    }
}

int cmf_launch_setconv1_fused(int b, int n, const float *xyz_planar, const float *ft_planar, const int *idx60,
                              const float *const *seg12x4 /* 4 scales x {W1,b1,W2,b2,W3,b3} */, float *out, cudaStream_t st) {
    SetConv1Args a;
    for (int s = 0; s < 4; ++s) {
        a.w[s].W1 = seg12x4[s * 6 + 0]; a.w[s].b1 = seg12x4[s * 6 + 1]; a.w[s].W2 = seg12x4[s * 6 + 2];
        a.w[s].b2 = seg12x4[s * 6 + 3]; a.w[s].W3 = seg12x4[s * 6 + 4]; a.w[s].b3 = seg12x4[s * 6 + 5];
    }
    dim3 grid(cmf_divup((long long)n * 32, 256), 4, b);
    setconv1_fused_kernel<<<grid, 128, 0, st>>>(n, xyz_planar, ft_planar, idx60, a, out);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

__global__ void __launch_bounds__(256)
maxk_kernel(long long points, int K, int C4, const float *__restrict__ Y, int ldy, float *__restrict__ out, int ldo) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= points * C4) return;
    const int c4 = (int)(t % C4);
    const long long p = t / C4;
    const float *src = Y + (size_t)p * K * ldy + c4 * 4;
    float4 m = __ldg(reinterpret_cast<const float4 *>(src));
    for (int kk = 1; kk < K; ++kk) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(src + (size_t)kk * ldy));
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
    *reinterpret_cast<float4 *>(out + (size_t)p * ldo + c4 * 4) = m;
}
int cmf_launch_maxk(long long points, int K, int C, const float *Y, int ldy, float *out, int ldo, cudaStream_t st) {
    maxk_kernel<<<cmf_divup(points * (C / 4), 256), 256, 0, st>>>(points, K, C / 4, Y, ldy, out, ldo);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

__global__ void __launch_bounds__(256)
globalmax_kernel(int n, int C, const float *__restrict__ F, int ldf, float *__restrict__ G, int rows_per_slice) {
    __shared__ float red[4][64];
    const int b = blockIdx.x, c = blockIdx.y * 64 + (threadIdx.x & 63), rg = threadIdx.x >> 6;
    const int r0 = blockIdx.z * rows_per_slice, r1 = min(n, r0 + rows_per_slice);
    float m = -FLT_MAX;
    if (c < C)
        for (int i = r0 + rg; i < r1; i += 4) m = fmaxf(m, __ldg(F + ((size_t)b * n + i) * ldf + c));
    red[rg][threadIdx.x & 63] = m;
    __syncthreads();
    if (rg == 0 && c < C) {
        m = fmaxf(fmaxf(red[0][threadIdx.x], red[1][threadIdx.x]), fmaxf(red[2][threadIdx.x], red[3][threadIdx.x]));
        float *g = G + (size_t)b * C + c;
        if (gridDim.z == 1) *g = m;
        // row slices combine by a sign-aware atomic max on the bit pattern; G was preset to 0xFFFFFFFF, which is below every float in both views
        else if (m >= 0.f) atomicMax(reinterpret_cast<int *>(g), __float_as_int(m));
        else atomicMin(reinterpret_cast<unsigned int *>(g), __float_as_uint(m));
    }
}
int cmf_launch_globalmax(int b, int n, int C, const float *F, int ldf, float *G, cudaStream_t st) {
    // few pairs with many points (the N=4096 configuration): split the rows over gridDim.z so that the grid still fills the chip
    int slices = 1;
    if ((long long)b * cmf_divup(C, 64) < 592 && n > 512) slices = cmf_divup(n, 256);
    if (slices > 1) CMF_CUDA(cudaMemsetAsync(G, 0xFF, (size_t)b * C * sizeof(float), st));
    globalmax_kernel<<<dim3(b, cmf_divup(C, 64), slices), 256, 0, st>>>(n, C, F, ldf, G, slices > 1 ? 256 : n);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// ---- per-pair maxima for the fp16x3 operand scales ------------------------------------------------
__global__ void __launch_bounds__(256)
pair_absmax_kernel(int n, const float *__restrict__ X, int ld, int width4, unsigned int *__restrict__ out) {
    const int b = blockIdx.x;
    const long long total = (long long)n * width4;
    float mx = 0.f;
    for (long long t = (long long)blockIdx.y * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.y * blockDim.x) {
        const int i = (int)(t / width4), c4 = (int)(t - (long long)i * width4);
        const float4 v = __ldg(reinterpret_cast<const float4 *>(X + ((size_t)b * n + i) * ld) + c4);
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(out + b, __float_as_uint(mx));
}
int cmf_launch_pair_absmax(int b, int n, const float *X, int ld, int width, unsigned int *out, cudaStream_t st) {
    if ((width & 3) || (ld & 3)) { cmf_set_error("pair_absmax: width / ld must be multiples of 4"); return CMF_ERR_INVALID; }
    const int split = cmf_divup((long long)n * (width / 4), 256 * 8) < 1 ? 1 : cmf_divup((long long)n * (width / 4), 256 * 8);
    pair_absmax_kernel<<<dim3(b, split), 256, 0, st>>>(n, X, ld, width / 4, out);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// =================================================================================================
// FeatureCorrelator (radarflow_util.py:185-237) after hoisting conv0 over the concat:
//   conv0([f1_i ; g1 ; f2_j ; g2 ; dir]) = U1[i] + U2[j] + Wd.dir      (U1,U2 carry the per-cloud constants and the bias)
// =================================================================================================
__device__ __forceinline__ float4 ld4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float leaky01(float v) { return v > 0.f ? v : 0.1f * v; }

__global__ void __launch_bounds__(128)
fc_build_h1_kernel(int n, int n2, const float *__restrict__ xyz1, const float *__restrict__ xyz2, const int *__restrict__ knn12,
                   const float *__restrict__ U1, const float *__restrict__ U2, const float *__restrict__ Wd,
                   float *__restrict__ H1) {
    const int bi = blockIdx.x, b = bi / n, i = bi - b * n, t = threadIdx.x;
    const float *p1 = xyz1 + (size_t)b * 3 * n, *p2 = xyz2 + (size_t)b * 3 * n2;
    const float qx = __ldg(p1 + i), qy = __ldg(p1 + n + i), qz = __ldg(p1 + 2 * n + i);
    const float4 u1 = ld4(U1 + (size_t)bi * 512 + t * 4);
    const float4 w0 = ld4(Wd + (t * 4 + 0) * 4), w1 = ld4(Wd + (t * 4 + 1) * 4), w2 = ld4(Wd + (t * 4 + 2) * 4), w3 = ld4(Wd + (t * 4 + 3) * 4);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int j = __ldg(knn12 + (size_t)bi * 8 + k);
        const float dx = __fsub_rn(__ldg(p2 + j), qx), dy = __fsub_rn(__ldg(p2 + n2 + j), qy), dz = __fsub_rn(__ldg(p2 + 2 * n2 + j), qz);
        const float4 u2 = ld4(U2 + ((size_t)b * n2 + j) * 512 + t * 4);
        float4 v;
        v.x = leaky01(u1.x + u2.x + fmaf(w0.z, dz, fmaf(w0.y, dy, w0.x * dx)));
        v.y = leaky01(u1.y + u2.y + fmaf(w1.z, dz, fmaf(w1.y, dy, w1.x * dx)));
        v.z = leaky01(u1.z + u2.z + fmaf(w2.z, dz, fmaf(w2.y, dy, w2.x * dx)));
        v.w = leaky01(u1.w + u2.w + fmaf(w3.z, dz, fmaf(w3.y, dy, w3.x * dx)));
        *reinterpret_cast<float4 *>(H1 + ((size_t)bi * 8 + k) * 512 + t * 4) = v;
    }
}
int cmf_launch_fc_build_h1(int b, int n, int n2, const float *xyz1_planar, const float *xyz2_planar, const int *knn12,
                           const float *U1, const float *U2, const float *Wd, float *H1, cudaStream_t st) {
    fc_build_h1_kernel<<<b * n, 128, 0, st>>>(n, n2, xyz1_planar, xyz2_planar, knn12, U1, U2, Wd, H1);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// out[i][c] = sum_k WeightNet(dir_ik)[c] * src_row(i,k)[c]     (radarflow_util.py:223-225 and 233-235)
// One CTA handles FCR_PTS consecutive points: thread (p, k) first runs the two hidden WeightNet layers (3 -> 8 -> 8) of its (point,
// neighbour) pair in registers; then thread t owns channels 4t..4t+3, keeps its four rows of the last WeightNet layer (8 -> 512) in
// registers for all FCR_PTS points (they used to be re-read per point: as many bytes as the gathered rows), and software-pipelines the
// eight 2 KB row gathers of point p+1 under the arithmetic of point p.
constexpr int FCR_PTS = 16;
__global__ void __launch_bounds__(128)
fc_reduce_kernel(long long points, int n, int nc, const float *__restrict__ xyzq, const float *__restrict__ xyzc, const int *__restrict__ knn,
                 WeightNetP wn, const float *__restrict__ src, int gather, float *__restrict__ out, int ldo, unsigned int *__restrict__ amax_out) {
    __shared__ __align__(16) float sh2[FCR_PTS][8][8];
    __shared__ float samx[4];
    __shared__ long long srow[FCR_PTS][8];
    const int t = threadIdx.x;
    const long long base = (long long)blockIdx.x * FCR_PTS;
    {
        const int p = t >> 3, k = t & 7;
        const long long bi = base + p;
        float h2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        long long row = 0;
        if (bi < points) {
            const long long b = bi / n; const int i = (int)(bi - b * n);
            const float *pq = xyzq + (size_t)b * 3 * n, *pc = xyzc + (size_t)b * 3 * nc;
            const int j = __ldg(knn + (size_t)bi * 8 + k);
            row = gather ? b * nc + j : bi * 8 + k;
            const float dx = __fsub_rn(__ldg(pc + j), __ldg(pq + i)), dy = __fsub_rn(__ldg(pc + nc + j), __ldg(pq + n + i)),
                        dz = __fsub_rn(__ldg(pc + 2 * nc + j), __ldg(pq + 2 * n + i));
            float h1[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float4 a = ld4(wn.A1 + u * 4);
                h1[u] = fmaxf(fmaf(a.z, dz, fmaf(a.y, dy, fmaf(a.x, dx, __ldg(wn.a1 + u)))), 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float sacc = __ldg(wn.a2 + u);
#pragma unroll
                for (int v = 0; v < 8; ++v) sacc = fmaf(__ldg(wn.A2 + u * 8 + v), h1[v], sacc);
                h2[u] = fmaxf(sacc, 0.f);
            }
        }
        *reinterpret_cast<float4 *>(&sh2[p][k][0]) = make_float4(h2[0], h2[1], h2[2], h2[3]);
        *reinterpret_cast<float4 *>(&sh2[p][k][4]) = make_float4(h2[4], h2[5], h2[6], h2[7]);
        srow[p][k] = row;
    }
    // last WeightNet layer of this thread's four channels, paired for packed fp32: P[h][j] = (A3[2h][j], A3[2h+1][j])
    float2 P[2][8], a3p[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float4 r0a = ld4(wn.A3 + (t * 4 + 2 * h) * 8), r0b = ld4(wn.A3 + (t * 4 + 2 * h) * 8 + 4);
        const float4 r1a = ld4(wn.A3 + (t * 4 + 2 * h + 1) * 8), r1b = ld4(wn.A3 + (t * 4 + 2 * h + 1) * 8 + 4);
        P[h][0] = make_float2(r0a.x, r1a.x); P[h][1] = make_float2(r0a.y, r1a.y); P[h][2] = make_float2(r0a.z, r1a.z); P[h][3] = make_float2(r0a.w, r1a.w);
        P[h][4] = make_float2(r0b.x, r1b.x); P[h][5] = make_float2(r0b.y, r1b.y); P[h][6] = make_float2(r0b.z, r1b.z); P[h][7] = make_float2(r0b.w, r1b.w);
        a3p[h] = make_float2(__ldg(wn.a3 + t * 4 + 2 * h), __ldg(wn.a3 + t * 4 + 2 * h + 1));
    }
    __syncthreads();
    const int np = (points - base) < FCR_PTS ? (int)(points - base) : FCR_PTS;
    float4 cur[8], nxt[8];
    float amx = 0.f;                                   // |out| maximum of this thread (per-pair fp16 scale of the consumer GEMM)
    const long long pair_first = base / n, pair_last = (base + np - 1) / n;
#pragma unroll
    for (int k = 0; k < 8; ++k) cur[k] = ld4(src + (size_t)srow[0][k] * 512 + t * 4);
    for (int p = 0; p < np; ++p) {
        if (p + 1 < np) {
#pragma unroll
            for (int k = 0; k < 8; ++k) nxt[k] = ld4(src + (size_t)srow[p + 1][k] * 512 + t * 4);
        }
        float2 acc2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 ha = *reinterpret_cast<const float4 *>(&sh2[p][k][0]), hb = *reinterpret_cast<const float4 *>(&sh2[p][k][4]);
            const float hv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
            const float2 sv[2] = {make_float2(cur[k].x, cur[k].y), make_float2(cur[k].z, cur[k].w)};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float2 w = a3p[h];                      // same fma order per channel as the scalar form: bit-identical
#pragma unroll
                for (int j = 0; j < 8; ++j) w = __ffma2_rn(P[h][j], make_float2(hv[j], hv[j]), w);
                acc2[h] = __ffma2_rn(make_float2(fmaxf(w.x, 0.f), fmaxf(w.y, 0.f)), sv[h], acc2[h]);
            }
        }
        const float acc[4] = {acc2[0].x, acc2[0].y, acc2[1].x, acc2[1].y};
        *reinterpret_cast<float4 *>(out + (size_t)(base + p) * ldo + t * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (amax_out) {
            const float m4 = fmaxf(fmaxf(fabsf(acc[0]), fabsf(acc[1])), fmaxf(fabsf(acc[2]), fabsf(acc[3])));
            if (pair_first == pair_last) amx = fmaxf(amx, m4);
            else if (m4 > 0.f) atomicMax(amax_out + (base + p) / n, __float_as_uint(m4));      // CTA straddles two pairs (N % 16 != 0): per point
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) cur[k] = nxt[k];
    }
    if (amax_out && pair_first == pair_last) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, off));
        if ((t & 31) == 0) samx[t >> 5] = amx;
        __syncthreads();
        if (t == 0) {
            const float m = fmaxf(fmaxf(samx[0], samx[1]), fmaxf(samx[2], samx[3]));
            if (m > 0.f) atomicMax(amax_out + pair_first, __float_as_uint(m));
        }
    }
}
int cmf_launch_fc_reduce(int b, int n, int n_cand, const float *xyzq_planar, const float *xyzc_planar, const int *knn,
                         WeightNetP wn, const float *src, int gather, float *out, int ldo, cudaStream_t st, unsigned int *amax_out) {
    const long long points = (long long)b * n;
    if (points <= 0) return CMF_OK;
    fc_reduce_kernel<<<cmf_divup(points, FCR_PTS), 128, 0, st>>>(points, n, n_cand, xyzq_planar, xyzc_planar, knn, wn, src, gather, out, ldo, amax_out);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// =================================================================================================
// set-conv #2 (mse_layer2) first layer after hoisting the 1027-channel conv over the gather
// =================================================================================================
__global__ void __launch_bounds__(128)
mse2_build_y1_kernel(int n, int K, int koff, const float *__restrict__ xyz, const int *__restrict__ idx60,
                     const float *__restrict__ P, int ldp, int poff, const float *__restrict__ Wx, float *__restrict__ Y1) {
    const int bi = blockIdx.x, b = bi / n, i = bi - b * n, t = threadIdx.x;
    const float *px = xyz + (size_t)b * 3 * n;
    const float qx = __ldg(px + i), qy = __ldg(px + n + i), qz = __ldg(px + 2 * n + i);
    const float4 w0 = ld4(Wx + (t * 4 + 0) * 4), w1 = ld4(Wx + (t * 4 + 1) * 4), w2 = ld4(Wx + (t * 4 + 2) * 4), w3 = ld4(Wx + (t * 4 + 3) * 4);
    for (int kk = 0; kk < K; ++kk) {
        const int j = __ldg(idx60 + (size_t)bi * 60 + koff + kk);
        const float dx = __fsub_rn(__ldg(px + j), qx), dy = __fsub_rn(__ldg(px + n + j), qy), dz = __fsub_rn(__ldg(px + 2 * n + j), qz);
        const float4 p = ld4(P + ((size_t)b * n + j) * ldp + poff + t * 4);
        float4 v;
        v.x = fmaxf(p.x + fmaf(w0.z, dz, fmaf(w0.y, dy, w0.x * dx)), 0.f);
        v.y = fmaxf(p.y + fmaf(w1.z, dz, fmaf(w1.y, dy, w1.x * dx)), 0.f);
        v.z = fmaxf(p.z + fmaf(w2.z, dz, fmaf(w2.y, dy, w2.x * dx)), 0.f);
        v.w = fmaxf(p.w + fmaf(w3.z, dz, fmaf(w3.y, dy, w3.x * dx)), 0.f);
        *reinterpret_cast<float4 *>(Y1 + ((size_t)bi * K + kk) * 512 + t * 4) = v;
    }
}
int cmf_launch_mse2_build_y1(int b, int n, int K, int koff, const float *xyz_planar, const int *idx60,
                             const float *P, int ldp, int poff, const float *Wx, float *Y1, cudaStream_t st) {
    mse2_build_y1_kernel<<<b * n, 128, 0, st>>>(n, K, koff, xyz_planar, idx60, P, ldp, poff, Wx, Y1);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// =================================================================================================
// heads' last 1x1 conv (radarflow_util.py:248,259 / 276,283): 64 -> 3 and 64 -> 1 + sigmoid
// =================================================================================================
__global__ void __launch_bounds__(128)
head_final_kernel(int n, const float *__restrict__ H3, int ldh, const float *__restrict__ W4f, const float *__restrict__ W4m,
                  float *__restrict__ flow, float *__restrict__ cls) {
    __shared__ float w[4][64];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) w[t >> 6][t & 63] = (t < 192) ? __ldg(W4f + t) : __ldg(W4m + (t - 192));
    __syncthreads();
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *h = H3 + ((size_t)b * n + i) * ldh;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int c = 0; c < 64; c += 4) {
        const float4 f = ld4(h + c), m = ld4(h + 64 + c);
        a0 = fmaf(w[0][c], f.x, a0); a0 = fmaf(w[0][c + 1], f.y, a0); a0 = fmaf(w[0][c + 2], f.z, a0); a0 = fmaf(w[0][c + 3], f.w, a0);
        a1 = fmaf(w[1][c], f.x, a1); a1 = fmaf(w[1][c + 1], f.y, a1); a1 = fmaf(w[1][c + 2], f.z, a1); a1 = fmaf(w[1][c + 3], f.w, a1);
        a2 = fmaf(w[2][c], f.x, a2); a2 = fmaf(w[2][c + 1], f.y, a2); a2 = fmaf(w[2][c + 2], f.z, a2); a2 = fmaf(w[2][c + 3], f.w, a2);
        a3 = fmaf(w[3][c], m.x, a3); a3 = fmaf(w[3][c + 1], m.y, a3); a3 = fmaf(w[3][c + 2], m.z, a3); a3 = fmaf(w[3][c + 3], m.w, a3);
    }
    float *fo = flow + (size_t)b * 3 * n;
    fo[i] = a0; fo[n + i] = a1; fo[2 * n + i] = a2;
    cls[(size_t)b * n + i] = 1.0f / (1.0f + expf(-a3));
}
int cmf_launch_head_final(int b, int n, const float *H3, int ldh, const float *W4f, const float *W4m,
                          float *flow_planar, float *cls, cudaStream_t st) {
    head_final_kernel<<<dim3(cmf_divup(n, 128), b), 128, 0, st>>>(n, H3, ldh, W4f, W4m, flow_planar, cls);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// =================================================================================================
// GRU gates (nn.GRU one step, gate order r,z,n; models/cmflow_t.py:46,101)
// =================================================================================================
__global__ void gru_gates_kernel(int total, const float *__restrict__ gi, const float *__restrict__ gh,
                                 const float *__restrict__ h_prev, float *__restrict__ h_new) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int b = t >> 8, u = t & 255;
    const float *i_ = gi + (size_t)b * 768, *h_ = gh + (size_t)b * 768;
    const float r = 1.0f / (1.0f + expf(-(i_[u] + h_[u])));
    const float z = 1.0f / (1.0f + expf(-(i_[256 + u] + h_[256 + u])));
    const float nn = tanhf(i_[512 + u] + r * h_[512 + u]);
    const float hp = h_prev ? h_prev[t] : 0.f;
    h_new[t] = (1.0f - z) * nn + z * hp;
}
int cmf_launch_gru_gates(int b, const float *gi, const float *gh, const float *h_prev, float *h_new, cudaStream_t st) {
    gru_gates_kernel<<<cmf_divup(b * 256, 256), 256, 0, st>>>(b * 256, gi, gh, h_prev, h_new);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// =================================================================================================
// Weighted Kabsch + refine (models/cmflow.py:96-169, 112-125): one CTA per frame pair.
// Moments are accumulated in fp64 in a single pass; the 3x3 SVD is a one-sided Jacobi in fp64, so the
// rotation is the polar factor V U^T to ~1e-15 -- the reference's fp32 cuSOLVER/LAPACK result agrees with
// it to fp32 rounding.  The reflection rule reproduces the reference's quirk: ROW 2 of V is negated
// (cmflow.py:162), i.e. R = diag(1,1,-1) V U^T when det(V U^T) < 0.
// =================================================================================================
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__device__ void polar_rotation_3x3(const double H[3][3], double R[3][3]) {
    double A[3][3], V[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { A[r][c] = H[r][c]; V[r][c] = (r == c) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double offn = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double al = 0, be = 0, ga = 0;
                for (int r = 0; r < 3; ++r) { al += A[r][p] * A[r][p]; be += A[r][q] * A[r][q]; ga += A[r][p] * A[r][q]; }
                if (fabs(ga) <= 1e-18 * sqrt(al * be) || ga == 0.0) continue;
                offn += fabs(ga);
                const double zeta = (be - al) / (2.0 * ga);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
                for (int r = 0; r < 3; ++r) {
                    double x = A[r][p], y = A[r][q];
                    A[r][p] = cs * x - sn * y; A[r][q] = sn * x + cs * y;
                    x = V[r][p]; y = V[r][q];
                    V[r][p] = cs * x - sn * y; V[r][q] = sn * x + cs * y;
                }
            }
        if (offn == 0.0) break;
    }
    // A = U S (columns); normalise to U, repairing a (near-)null column by a cross product
    double nrm[3], U[3][3];
    double nmax = 0;
    for (int c = 0; c < 3; ++c) { nrm[c] = sqrt(A[0][c] * A[0][c] + A[1][c] * A[1][c] + A[2][c] * A[2][c]); nmax = fmax(nmax, nrm[c]); }
    int bad = -1, nbad = 0;
    for (int c = 0; c < 3; ++c) {
        if (nrm[c] > 1e-13 * nmax && nrm[c] > 0) { for (int r = 0; r < 3; ++r) U[r][c] = A[r][c] / nrm[c]; }
        else { bad = c; ++nbad; }
    }
    if (nbad == 1) {
        const int a = (bad + 1) % 3, b2 = (bad + 2) % 3;
        double cx = U[1][a] * U[2][b2] - U[2][a] * U[1][b2], cy = U[2][a] * U[0][b2] - U[0][a] * U[2][b2], cz = U[0][a] * U[1][b2] - U[1][a] * U[0][b2];
        // choose the sign that makes det(U) = det(V) (so that V U^T is a proper rotation)
        double detV = V[0][0] * (V[1][1] * V[2][2] - V[1][2] * V[2][1]) - V[0][1] * (V[1][0] * V[2][2] - V[1][2] * V[2][0]) + V[0][2] * (V[1][0] * V[2][1] - V[1][1] * V[2][0]);
        U[0][bad] = cx; U[1][bad] = cy; U[2][bad] = cz;      // det(U) = +1 with (a,b,bad) cyclic
        if (detV < 0) { U[0][bad] = -cx; U[1][bad] = -cy; U[2][bad] = -cz; }
    } else if (nbad >= 2) {
        // rank <= 1: V U^T is not unique (the reference's answer depends on LAPACK's choice of the null-space bases).  Any rotation that takes
        // the one defined left singular vector u to its right partner v is optimal; return the smallest such rotation (identity for H = 0).
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r][c] = (r == c) ? 1.0 : 0.0;
        if (nbad == 3) return;
        int gc = 0;
        for (int c = 0; c < 3; ++c) if (nrm[c] > 1e-13 * nmax && nrm[c] > 0) gc = c;
        const double u[3] = {U[0][gc], U[1][gc], U[2][gc]}, v[3] = {V[0][gc], V[1][gc], V[2][gc]};
        const double cth = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
        if (cth > -1.0 + 1e-12) {                            // Rodrigues: R = I + [k]x + [k]x^2 / (1 + cos), k = u x v
            const double k[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
            const double Kx[3][3] = {{0, -k[2], k[1]}, {k[2], 0, -k[0]}, {-k[1], k[0], 0}};
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    double k2 = 0; for (int j = 0; j < 3; ++j) k2 += Kx[r][j] * Kx[j][c];
                    R[r][c] += Kx[r][c] + k2 / (1.0 + cth);
                }
        } else {                                             // u = -v: half turn about any axis n perpendicular to u, R = 2 n n^T - I
            const int ax = fabs(u[0]) < fabs(u[1]) ? (fabs(u[0]) < fabs(u[2]) ? 0 : 2) : (fabs(u[1]) < fabs(u[2]) ? 1 : 2);
            double e[3] = {0, 0, 0}; e[ax] = 1.0;
            double n[3] = {u[1] * e[2] - u[2] * e[1], u[2] * e[0] - u[0] * e[2], u[0] * e[1] - u[1] * e[0]};
            const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r][c] = 2.0 * n[r] * n[c] / (nn * nn) - (r == c ? 1.0 : 0.0);
        }
        return;
    }
    double Z[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Z[r][c] = V[r][0] * U[c][0] + V[r][1] * U[c][1] + V[r][2] * U[c][2];
    const double det = Z[0][0] * (Z[1][1] * Z[2][2] - Z[1][2] * Z[2][1]) - Z[0][1] * (Z[1][0] * Z[2][2] - Z[1][2] * Z[2][0]) + Z[0][2] * (Z[1][0] * Z[2][1] - Z[1][1] * Z[2][0]);
    const double s2 = det < 0 ? -1.0 : 1.0;                  // cmflow.py:157-162: negate ROW 2 of V
    for (int c = 0; c < 3; ++c) { R[0][c] = Z[0][c]; R[1][c] = Z[1][c]; R[2][c] = s2 * Z[2][c]; }
}

__global__ void __launch_bounds__(256)
kabsch_kernel(int n, const float *__restrict__ pc1, const float *__restrict__ second, int second_is_flow,
              const float *__restrict__ w, int normalise, float eps, float stat_thres,
              float *__restrict__ trans, float *__restrict__ sf_agg, uint8_t *__restrict__ mask) {
    __shared__ double red[8][16];
    __shared__ float sT[12];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *pa = pc1 + (size_t)b * 3 * n, *ps = second + (size_t)b * 3 * n, *pw = w + (size_t)b * n;
    double m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
        const float ax = pa[i], ay = pa[n + i], az = pa[2 * n + i];
        float bx = ps[i], by = ps[n + i], bz = ps[2 * n + i];
        if (second_is_flow) { bx = __fadd_rn(ax, bx); by = __fadd_rn(ay, by); bz = __fadd_rn(az, bz); }   // pc1_warp = pc1 + flow (cmflow.py:102)
        const double wi = normalise ? (double)__fadd_rn(pw[i], eps) : (double)pw[i];
        const double A[3] = {ax, ay, az}, Bv[3] = {bx, by, bz};
        m[0] += wi;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            m[1 + r] += wi * A[r];
            m[4 + r] += wi * Bv[r];
#pragma unroll
            for (int c = 0; c < 3; ++c) m[7 + r * 3 + c] += wi * A[r] * Bv[c];
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = warp_sum_d(m[i]);
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < 16; ++i) red[warp][i] = m[i];
    __syncthreads();
    if (tid == 0) {
        double t[16];
        for (int i = 0; i < 16; ++i) { t[i] = 0; for (int wv = 0; wv < 8; ++wv) t[i] += red[wv][i]; }
        double sw = t[0];
        if (normalise) { for (int i = 1; i < 16; ++i) t[i] /= sw; sw = 1.0; }
        const double cA[3] = {t[1], t[2], t[3]}, cB[3] = {t[4], t[5], t[6]};
        double H[3][3], R[3][3];
        // H = sum w (a-cA)(b-cB)^T with cA = sum w a, cB = sum w b (no division by sum w: cmflow.py:138-151)
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[r][c] = t[7 + r * 3 + c] - (2.0 - sw) * cA[r] * cB[c];
        polar_rotation_3x3(H, R);
        for (int r = 0; r < 3; ++r) {
            const double tr = -(R[r][0] * cA[0] + R[r][1] * cA[1] + R[r][2] * cA[2]) + cB[r];
            sT[r * 4 + 0] = (float)R[r][0]; sT[r * 4 + 1] = (float)R[r][1]; sT[r * 4 + 2] = (float)R[r][2]; sT[r * 4 + 3] = (float)tr;
        }
    }
    __syncthreads();
    if (tid < 16) trans[(size_t)b * 16 + tid] = tid < 12 ? sT[tid] : (tid == 15 ? 1.f : 0.f);
    if (!sf_agg) return;
    // refine_with_transform (cmflow.py:112-125): static points take the rigid flow T.[p;1] - p
    for (int i = tid; i < n; i += blockDim.x) {
        const float x = pa[i], y = pa[n + i], z = pa[2 * n + i];
        const bool st = pw[i] > stat_thres;                      // mask = scores > stat_thres (cmflow.py:188)
        float o[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float rg = __fsub_rn(fmaf(sT[r * 4 + 2], z, fmaf(sT[r * 4 + 1], y, fmaf(sT[r * 4 + 0], x, sT[r * 4 + 3]))), r == 0 ? x : (r == 1 ? y : z));
            o[r] = st ? rg : ps[r * n + i];
        }
        float *so = sf_agg + (size_t)b * 3 * n;
        so[i] = o[0]; so[n + i] = o[1]; so[2 * n + i] = o[2];
        if (mask) mask[(size_t)b * n + i] = st ? 1 : 0;
    }
}

int cmf_launch_kabsch(int b, int n, const float *pc1, const float *pc_or_flow, int second_is_flow, const float *w,
                      int normalise, float eps, float stat_thres, float *trans, float *sf_agg, uint8_t *mask,
                      cudaStream_t st) {
    kabsch_kernel<<<b, 256, 0, st>>>(n, pc1, pc_or_flow, second_is_flow, w, normalise, eps, stat_thres, trans, sf_agg, mask);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// =================================================================================================
// RaFlow's scene-flow refinement (SFR_module + rigid_transform_torch, models/raflow.py:79-156): one CTA per frame pair.
//   T0 = rigid fit of (pc1, pc1 + output) over all points; rigid flow sf_rg = T0.[p;1] - p;
//   mask_s = |(vel*interval - <sf_rg, p>/|p|) / vel| < rigid_thres  (vel = feature1 channel 0);
//   if more than rigid_pcs of the points are inliers: T1 = fit over the inliers, inliers take T1's rigid flow; else T0 and the raw flow.
// Reference quirks kept: centroids are divided by N even for the masked fit (raflow.py:129-130), all points are centred but only masked
// columns enter H (:137-140), ROW 2 of V is negated on reflection (:151).  Moments in fp64, one pass per fit.
// =================================================================================================
__device__ __forceinline__ void raflow_fit(const double t[16], int n, float *sT /* 12 */) {
    const double inv_n = 1.0 / (double)n;
    const double cA[3] = {t[1] * inv_n, t[2] * inv_n, t[3] * inv_n}, cB[3] = {t[4] * inv_n, t[5] * inv_n, t[6] * inv_n};
    double H[3][3], R[3][3];
    // sum_i w_i (a_i - cA)(b_i - cB)^T = M_ab - cA m_b^T - m_a cB^T + m_0 cA cB^T
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) H[r][c] = t[7 + r * 3 + c] - cA[r] * t[4 + c] - t[1 + r] * cB[c] + t[0] * cA[r] * cB[c];
    polar_rotation_3x3(H, R);
    for (int r = 0; r < 3; ++r) {
        const double tr = -(R[r][0] * cA[0] + R[r][1] * cA[1] + R[r][2] * cA[2]) + cB[r];
        sT[r * 4 + 0] = (float)R[r][0]; sT[r * 4 + 1] = (float)R[r][1]; sT[r * 4 + 2] = (float)R[r][2]; sT[r * 4 + 3] = (float)tr;
    }
}

__global__ void __launch_bounds__(256)
raflow_sfr_kernel(int n, const float *__restrict__ pc1, const float *__restrict__ ft1, const float *__restrict__ flow,
                  const float *__restrict__ interval, float rigid_thres, float rigid_pcs,
                  float *__restrict__ sf_agg, float *__restrict__ trans, uint8_t *__restrict__ mask) {
    __shared__ double red[8][17];
    __shared__ float sT0[12], sT1[12];
    __shared__ int s_use1;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *pa = pc1 + (size_t)b * 3 * n, *pf = flow + (size_t)b * 3 * n, *pv = ft1 + (size_t)b * 3 * n;
    uint8_t *pm = mask + (size_t)b * n;
    const float dt = interval[b];
    double m[17];
    auto accumulate = [&](bool masked) {
#pragma unroll
        for (int i = 0; i < 17; ++i) m[i] = 0.0;
        for (int i = tid; i < n; i += blockDim.x) {
            const float ax = pa[i], ay = pa[n + i], az = pa[2 * n + i];
            const float bx = __fadd_rn(ax, pf[i]), by = __fadd_rn(ay, pf[n + i]), bz = __fadd_rn(az, pf[2 * n + i]);   // pc1_warp (raflow.py:85)
            double wi = 1.0;
            if (masked) {
                // rigid flow of T0 and its radial projection (raflow.py:92-98)
                const float rx = __fsub_rn(fmaf(sT0[2], az, fmaf(sT0[1], ay, fmaf(sT0[0], ax, sT0[3]))), ax);
                const float ry = __fsub_rn(fmaf(sT0[6], az, fmaf(sT0[5], ay, fmaf(sT0[4], ax, sT0[7]))), ay);
                const float rz = __fsub_rn(fmaf(sT0[10], az, fmaf(sT0[9], ay, fmaf(sT0[8], ax, sT0[11]))), az);
                const float proj = __fdiv_rn(fmaf(rz, az, fmaf(ry, ay, __fmul_rn(rx, ax))), sqrtf(fmaf(az, az, fmaf(ay, ay, __fmul_rn(ax, ax)))));
                const float vel = pv[i];
                const float ratio = fabsf(__fdiv_rn(__fsub_rn(__fmul_rn(vel, dt), proj), vel));
                const bool in = ratio < rigid_thres;                      // NaN / inf (vel == 0) -> false, as in torch
                pm[i] = in ? 1 : 0;
                wi = in ? 1.0 : 0.0;
                m[16] += wi;
            }
            const double A[3] = {ax, ay, az}, Bv[3] = {bx, by, bz};
            m[0] += wi;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                m[1 + r] += wi * A[r];
                m[4 + r] += wi * Bv[r];
#pragma unroll
                for (int c = 0; c < 3; ++c) m[7 + r * 3 + c] += wi * A[r] * Bv[c];
            }
        }
#pragma unroll
        for (int i = 0; i < 17; ++i) m[i] = warp_sum_d(m[i]);
        if (lane == 0)
#pragma unroll
            for (int i = 0; i < 17; ++i) red[warp][i] = m[i];
        __syncthreads();
    };
    accumulate(false);
    if (tid == 0) {
        double t[16];
        for (int i = 0; i < 16; ++i) { t[i] = 0; for (int wv = 0; wv < 8; ++wv) t[i] += red[wv][i]; }
        raflow_fit(t, n, sT0);
    }
    __syncthreads();
    accumulate(true);
    if (tid == 0) {
        double t[17];
        for (int i = 0; i < 17; ++i) { t[i] = 0; for (int wv = 0; wv < 8; ++wv) t[i] += red[wv][i]; }
        // (mask_s[b].sum() / N) > rigid_pcs in float32 (raflow.py:104)
        s_use1 = __fdiv_rn((float)t[16], (float)n) > rigid_pcs ? 1 : 0;
        if (s_use1) raflow_fit(t, n, sT1);
    }
    __syncthreads();
    const float *T = s_use1 ? sT1 : sT0;
    if (tid < 16) trans[(size_t)b * 16 + tid] = tid < 12 ? T[tid] : (tid == 15 ? 1.f : 0.f);
    float *so = sf_agg + (size_t)b * 3 * n;
    for (int i = tid; i < n; i += blockDim.x) {
        const float x = pa[i], y = pa[n + i], z = pa[2 * n + i];
        const bool rigid = s_use1 && pm[i];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float rg = __fsub_rn(fmaf(T[r * 4 + 2], z, fmaf(T[r * 4 + 1], y, fmaf(T[r * 4 + 0], x, T[r * 4 + 3]))), r == 0 ? x : (r == 1 ? y : z));
            so[r * n + i] = rigid ? rg : pf[r * n + i];
        }
    }
}

int cmf_launch_raflow_sfr(int b, int n, const float *pc1, const float *ft1, const float *flow, const float *interval,
                          float rigid_thres, float rigid_pcs, float *sf_agg, float *trans, uint8_t *mask, cudaStream_t st) {
    raflow_sfr_kernel<<<b, 256, 0, st>>>(n, pc1, ft1, flow, interval, rigid_thres, rigid_pcs, sf_agg, trans, mask);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

extern "C" int cmf_raflow_refine(int b, int n, const float *pc1, const float *ft1, const float *flow, const float *interval,
                                 float rigid_thres, float rigid_pcs, float *sf_agg, float *trans, uint8_t *mask, void *stream) {
    CMF_REQUIRE(b >= 0 && n >= 1, "need b >= 0, n >= 1");
    if (b == 0) return CMF_OK;
    CMF_REQUIRE(pc1 && ft1 && flow && interval && sf_agg && trans && mask, "null pointer");
    return cmf_launch_raflow_sfr(b, n, pc1, ft1, flow, interval, rigid_thres, rigid_pcs, sf_agg, trans, mask, (cudaStream_t)stream);
}

extern "C" int cmf_kabsch_refine(int b, int n, const float *pc1, const float *flow, const float *score,
                                 float eps, float stat_thres, float *trans, float *sf_agg, uint8_t *mask, void *stream) {
    CMF_REQUIRE(b >= 0 && n >= 1, "need b >= 0, n >= 1");
    if (b == 0) return CMF_OK;
    CMF_REQUIRE(pc1 && flow && score && trans && sf_agg, "null pointer");
    return cmf_launch_kabsch(b, n, pc1, flow, 1, score, 1, eps, stat_thres, trans, sf_agg, mask, (cudaStream_t)stream);
}

extern "C" int cmf_weighted_kabsch(int b, int n, const float *A, const float *Bp, const float *W, float *trans, void *stream) {
    CMF_REQUIRE(b >= 0 && n >= 1, "need b >= 0, n >= 1");
    if (b == 0) return CMF_OK;
    CMF_REQUIRE(A && Bp && W && trans, "null pointer");
    return cmf_launch_kabsch(b, n, A, Bp, 0, W, 0, 0.f, 0.f, trans, nullptr, nullptr, (cudaStream_t)stream);
}
