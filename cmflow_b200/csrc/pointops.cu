// pointops.cu -- Part 1 of the C ABI: the ten launchers behind the reference's `pointnet2_cuda`
// module (lib/src/pointnet2_api.cpp:11-24), rewritten for sm_100a, plus the model-path kNN.
//
// Design notes (B200): these are integer/byte kernels bounded by HBM/L2 traffic and latency, not math.
//  * neighbour searches are warp-per-query: candidates are staged once per CTA in shared memory with
//    128-bit coalesced loads, 32 lanes test 32 candidates per step, __ballot_sync/__popc place hits in
//    index order (ball query) and a shuffle arg-min merges per-lane sorted lists (kNN);
//  * gathers keep the index in a register and loop over channels so idx is read once, not C times,
//    and every store is coalesced;
//  * all distance arithmetic uses the reference's exact fma contraction (cmf_common.cuh) so indices
//    are bit-identical to the reference's own CUDA build.
#include <limits.h>
#include <math.h>

#include "cmf_common.cuh"

// ------------------------------------------------------------------------------------------------
// ball query   (reference: lib/src/ball_query_gpu.cu:9-45)
// ------------------------------------------------------------------------------------------------
constexpr int BQ_THREADS = 256;
constexpr int BQ_QPW = 4;          // queries per warp
constexpr int BQ_CHUNK = 2048;     // candidates staged per pass (24 KB)

__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(int n, int m, float radius2, int nsample,
                  const float *__restrict__ new_xyz, const float *__restrict__ xyz, int *__restrict__ idx) {
    __shared__ __align__(16) float s[BQ_CHUNK * 3];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int q0 = (blockIdx.x * (BQ_THREADS / 32) + warp) * BQ_QPW;
    xyz += (size_t)b * n * 3;

    float qx[BQ_QPW], qy[BQ_QPW], qz[BQ_QPW];
    int cnt[BQ_QPW], first[BQ_QPW];
    int *out[BQ_QPW];
    bool all_done = true;
#pragma unroll
    for (int t = 0; t < BQ_QPW; ++t) {
        int q = q0 + t;
        bool valid = q < m;
        const float *p = new_xyz + ((size_t)b * m + (valid ? q : 0)) * 3;
        qx[t] = __ldg(p); qy[t] = __ldg(p + 1); qz[t] = __ldg(p + 2);
        cnt[t] = valid ? 0 : nsample;          // invalid queries are "full" from the start
        first[t] = -1;
        out[t] = idx + ((size_t)b * m + (valid ? q : 0)) * nsample;
        all_done = all_done && !valid;
    }
    if (nsample <= 0) all_done = true;

    for (int base = 0; base < n; base += BQ_CHUNK) {
        if (__syncthreads_and(all_done)) break;           // also fences reuse of s[]
        const int cn = min(BQ_CHUNK, n - base);
        cmf_stage_floats(s, xyz + (size_t)base * 3, cn * 3);
        __syncthreads();
        if (all_done) continue;
        for (int j = 0; j < cn; j += 32) {
            const int k = j + lane;
            float cx = 0.f, cy = 0.f, cz = 0.f;
            if (k < cn) { cx = s[3 * k]; cy = s[3 * k + 1]; cz = s[3 * k + 2]; }
            bool any_open = false;
#pragma unroll
            for (int t = 0; t < BQ_QPW; ++t) {
                if (cnt[t] >= nsample) continue;           // warp-uniform
                bool hit = (k < cn) && (cmf_sqdist_ref(qx[t], qy[t], qz[t], cx, cy, cz) < radius2);
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                if (mask) {
                    if (first[t] < 0) first[t] = base + j + __ffs(mask) - 1;
                    int pos = cnt[t] + __popc(mask & lt);
                    if (hit && pos < nsample) out[t][pos] = base + k;
                    cnt[t] += __popc(mask);
                }
                any_open = any_open || (cnt[t] < nsample);
            }
            if (!any_open) { all_done = true; break; }
        }
    }
    // pad the remainder of each row with the first hit (ball_query_gpu.cu:36-40); rows without hits stay untouched
#pragma unroll
    for (int t = 0; t < BQ_QPW; ++t) {
        if (q0 + t < m && first[t] >= 0)
            for (int l = cnt[t] + lane; l < nsample; l += 32) out[t][l] = first[t];
    }
}

extern "C" int cmf_ball_query(int b, int n, int m, float radius, int nsample,
                              const float *new_xyz, const float *xyz, int *idx, void *stream) {
    CMF_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 0, "negative size");
    if (b == 0 || m == 0 || nsample == 0 || n == 0) return CMF_OK;
    CMF_REQUIRE(new_xyz && xyz && idx, "null pointer");
    CMF_REQUIRE(b <= 65535, "batch > 65535");
    dim3 grid(cmf_divup(m, (BQ_THREADS / 32) * BQ_QPW), b);
    ball_query_kernel<<<grid, BQ_THREADS, 0, (cudaStream_t)stream>>>(n, m, radius * radius, nsample, new_xyz, xyz, idx);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// ------------------------------------------------------------------------------------------------
// kNN, k <= 32.  MODE 0: lib/src/interpolate_gpu.cu:9-57 (direct-form distance, ordered);
//                MODE 1: radarflow_util.py:88-99 knn_point (expanded-form distance).
// A warp answers KNN_QPW queries.  The k best of a query live ACROSS the warp, sorted: lane r holds the r-th best (d, index), and
// tau = the k-th best distance is warp-uniform.  Candidates are taken 32 at a time (one per lane, staged in shared memory, |x|^2
// precomputed once per CTA for the expanded form); a ballot of d < tau finds the few that matter -- about k ln(N/k) of N per query --
// and each of those is inserted by one ballot (its rank) and one shuffle-up.  Order rule as the reference: ascending by (d, index),
// because candidates arrive in increasing index order, hits of a batch are taken from the lowest lane up, and a new element goes
// behind every element with d' <= d (the reference's strict '<').  No final merge: lane r writes result r.
// (The earlier version kept k best PER LANE: with 32 private thresholds nearly every batch triggered an 8-step insertion in some
// lane -- 74 instructions per batch at N=4096; this one spends ~13.)
// ------------------------------------------------------------------------------------------------
constexpr int KNN_THREADS = 256;
constexpr int KNN_CHUNK = 2048;
constexpr int KNN_QPW = 2;

template <int MODE>
__device__ __forceinline__ void knn_warp_body(int b, int nq, int mc, int k, const float *__restrict__ query, const float *__restrict__ cand,
                                              float *__restrict__ dist_out, int *__restrict__ idx_out, unsigned int *__restrict__ dirmax) {
    __shared__ __align__(16) float s[KNN_CHUNK * 3];
    __shared__ float sn[MODE == 1 ? KNN_CHUNK : 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = (blockIdx.x * (KNN_THREADS / 32) + warp) * KNN_QPW;
    cand += (size_t)b * mc * 3;
    float qx[KNN_QPW], qy[KNN_QPW], qz[KNN_QPW], nqn[KNN_QPW], bd[KNN_QPW], tau[KNN_QPW];
    int bi[KNN_QPW];
#pragma unroll
    for (int t = 0; t < KNN_QPW; ++t) {
        const float *qp = query + ((size_t)b * nq + (q0 + t < nq ? q0 + t : 0)) * 3;
        qx[t] = __ldg(qp); qy[t] = __ldg(qp + 1); qz[t] = __ldg(qp + 2);
        nqn[t] = cmf_sqnorm3(qx[t], qy[t], qz[t]);
        bd[t] = INFINITY; bi[t] = INT_MAX; tau[t] = INFINITY;
    }
    const bool any_valid = q0 < nq;

    for (int base = 0; base < mc; base += KNN_CHUNK) {
        __syncthreads();
        const int cn = min(KNN_CHUNK, mc - base);
        cmf_stage_floats(s, cand + (size_t)base * 3, cn * 3);
        __syncthreads();
        if (MODE == 1) {
            for (int i = threadIdx.x; i < cn; i += KNN_THREADS) sn[i] = cmf_sqnorm3(s[3 * i], s[3 * i + 1], s[3 * i + 2]);
            __syncthreads();
        }
        if (!any_valid) continue;
        for (int j = 0; j < cn; j += 32) {
            const int kk = j + lane;
            const bool in = kk < cn;
            const int ks = in ? kk : 0;
            const float x = s[3 * ks], y = s[3 * ks + 1], z = s[3 * ks + 2];
            const float xn = MODE == 1 ? sn[ks] : 0.f;
#pragma unroll
            for (int t = 0; t < KNN_QPW; ++t) {
                float d;
                if (MODE == 0) d = cmf_sqdist_ref(qx[t], qy[t], qz[t], x, y, z);
                else d = cmf_sqdist_expanded(qx[t], qy[t], qz[t], nqn[t], x, y, z, xn);
                if (base == 0 && j == 0) {
                    // first batch, empty list: every lane would hit and be inserted one by one -- sort the 32 candidates instead (bitonic
                    // network over the lanes, ascending by (d, index)); inf / NaN / out-of-range lanes sort to the end as (inf, INT_MAX)
                    float sd = (in && d < INFINITY) ? d : INFINITY;
                    int si = (in && d < INFINITY) ? kk : INT_MAX;
#pragma unroll
                    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
                        for (int stride = size >> 1; stride > 0; stride >>= 1) {
                            const float od = __shfl_xor_sync(0xffffffffu, sd, stride);
                            const int oi = __shfl_xor_sync(0xffffffffu, si, stride);
                            const bool other_first = od < sd || (od == sd && oi < si);
                            const bool want_min = ((lane & stride) == 0) == ((lane & size) == 0);
                            if (want_min == other_first && (od != sd || oi != si)) { sd = od; si = oi; }
                        }
                    bd[t] = lane < k ? sd : INFINITY; bi[t] = lane < k ? si : INT_MAX;
                    tau[t] = __shfl_sync(0xffffffffu, bd[t], k - 1);
                    continue;
                }
                unsigned mask = __ballot_sync(0xffffffffu, in && d < tau[t]);      // also rejects inf / NaN like the reference's `d < best`
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float cd = __shfl_sync(0xffffffffu, d, src);
                    if (!(cd < tau[t])) continue;                                  // tau fell while this batch was being inserted
                    const int ci = base + j + src;
                    const int pos = __popc(__ballot_sync(0xffffffffu, bd[t] <= cd));   // elements that stay in front (earlier index wins ties)
                    const float pd = __shfl_up_sync(0xffffffffu, bd[t], 1);
                    const int pi = __shfl_up_sync(0xffffffffu, bi[t], 1);
                    if (lane > pos) { bd[t] = pd; bi[t] = pi; }
                    else if (lane == pos) { bd[t] = cd; bi[t] = ci; }
                    if (lane >= k) { bd[t] = INFINITY; bi[t] = INT_MAX; }
                    tau[t] = __shfl_sync(0xffffffffu, bd[t], k - 1);
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < KNN_QPW; ++t) {
        const int q = q0 + t;
        if (q >= nq || lane >= k) continue;
        idx_out[((size_t)b * nq + q) * k + lane] = (bi[t] == INT_MAX) ? 0 : bi[t];     // unfilled slot: (1e40 -> inf, 0) in the reference
        if (dist_out) dist_out[((size_t)b * nq + q) * k + lane] = bd[t];
    }
    if (dirmax) {          // engine only: per-pair max |candidate - query| component over the neighbours found (bound behind an fp16 operand scale)
        float mx = 0.f;
#pragma unroll
        for (int t = 0; t < KNN_QPW; ++t) {
            if (q0 + t >= nq || lane >= k) continue;
            const float *cp = cand + (size_t)((bi[t] == INT_MAX) ? 0 : bi[t]) * 3;
            mx = fmaxf(mx, fmaxf(fabsf(__fsub_rn(__ldg(cp), qx[t])), fmaxf(fabsf(__fsub_rn(__ldg(cp + 1), qy[t])), fabsf(__fsub_rn(__ldg(cp + 2), qz[t])))));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        if (lane == 0 && mx > 0.f) atomicMax(dirmax + b, __float_as_uint(mx));
    }
}
template <int MODE>
__global__ void __launch_bounds__(KNN_THREADS)
knn_warp_kernel(int nq, int mc, int k, const float *__restrict__ query, const float *__restrict__ cand,
                float *__restrict__ dist_out, int *__restrict__ idx_out) {
    knn_warp_body<MODE>(blockIdx.y, nq, mc, k, query, cand, dist_out, idx_out, nullptr);
}
// two searches of the same queries against two candidate clouds (blockIdx.z) in one launch: the engine's cross-frame and self 8-NN
__global__ void __launch_bounds__(KNN_THREADS)
knn_point_dual_kernel(int nq, int k, const float *__restrict__ query, int mc0, const float *__restrict__ cand0, int *__restrict__ idx0,
                      int mc1, const float *__restrict__ cand1, int *__restrict__ idx1, unsigned int *__restrict__ dirmax0) {
    if (blockIdx.z == 0) knn_warp_body<1>(blockIdx.y, nq, mc0, k, query, cand0, nullptr, idx0, dirmax0);
    else knn_warp_body<1>(blockIdx.y, nq, mc1, k, query, cand1, nullptr, idx1, nullptr);
}
int cmf_launch_knn_point8_dual(int b, int n_query, const float *query_aos, int n_cand0, const float *cand0_aos, int *idx0,
                               int n_cand1, const float *cand1_aos, int *idx1, unsigned int *dirmax0, cudaStream_t st) {
    if (b <= 0 || n_query <= 0) return CMF_OK;
    if (n_cand0 < 8 || n_cand1 < 8) { cmf_set_error("knn_point8_dual: fewer than 8 candidates (torch.topk raises too)"); return CMF_ERR_INVALID; }
    if (b > 65535) { cmf_set_error("knn_point8_dual: batch > 65535"); return CMF_ERR_INVALID; }
    dim3 grid(cmf_divup(n_query, (KNN_THREADS / 32) * KNN_QPW), b, 2);
    knn_point_dual_kernel<<<grid, KNN_THREADS, 0, st>>>(n_query, 8, query_aos, n_cand0, cand0_aos, idx0, n_cand1, cand1_aos, idx1, dirmax0);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// Fallback for 32 < k <= 200: one thread per query with local arrays, as the reference does.
__global__ void __launch_bounds__(128)
knn_thread_kernel(int nq, int mc, int k, const float *__restrict__ query, const float *__restrict__ cand,
                  float *__restrict__ dist_out, int *__restrict__ idx_out) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    cand += (size_t)b * mc * 3;
    const float *qp = query + ((size_t)b * nq + q) * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2];
    float bd[200]; int bi[200];
    for (int j = 0; j < k; ++j) { bd[j] = INFINITY; bi[j] = INT_MAX; }
    for (int i = 0; i < mc; ++i) {
        float d = cmf_sqdist_ref(qx, qy, qz, __ldg(cand + 3 * i), __ldg(cand + 3 * i + 1), __ldg(cand + 3 * i + 2));
        if (d < bd[k - 1]) {
            int j = k - 1;
            while (j > 0 && d < bd[j - 1]) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; --j; }
            bd[j] = d; bi[j] = i;
        }
    }
    for (int j = 0; j < k; ++j) {
        idx_out[((size_t)b * nq + q) * k + j] = bi[j] == INT_MAX ? 0 : bi[j];
        dist_out[((size_t)b * nq + q) * k + j] = bd[j];
    }
}

template <int MODE>
static int launch_knn(int b, int nq, int mc, int k, const float *query, const float *cand,
                      float *dist, int *idx, cudaStream_t st) {
    dim3 grid(cmf_divup(nq, (KNN_THREADS / 32) * KNN_QPW), b);
    knn_warp_kernel<MODE><<<grid, KNN_THREADS, 0, st>>>(nq, mc, k, query, cand, dist, idx);
    return 0;
}

extern "C" int cmf_knn(int b, int n, int m, int k, const float *unknown, const float *known,
                       float *dist2, int *idx, void *stream) {
    CMF_REQUIRE(b >= 0 && n >= 0 && m >= 0, "negative size");
    CMF_REQUIRE(k >= 1 && k <= 200, "k must be in [1,200] (reference: fixed best[200], interpolate_gpu.cu:30)");
    if (b == 0 || n == 0) return CMF_OK;
    CMF_REQUIRE(unknown && known && dist2 && idx, "null pointer");
    CMF_REQUIRE(b <= 65535, "batch > 65535");
    if (k <= 32) launch_knn<0>(b, n, m, k, unknown, known, dist2, idx, (cudaStream_t)stream);
    else knn_thread_kernel<<<dim3(cmf_divup(n, 128), b), 128, 0, (cudaStream_t)stream>>>(n, m, k, unknown, known, dist2, idx);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

extern "C" int cmf_three_nn(int b, int n, int m, const float *unknown, const float *known,
                            float *dist2, int *idx, void *stream) {
    // three_nn_kernel_fast (interpolate_gpu.cu:81-124) is the k=3 case of the same ordered search
    CMF_REQUIRE(b >= 0 && n >= 0 && m >= 0, "negative size");
    if (b == 0 || n == 0) return CMF_OK;
    CMF_REQUIRE(unknown && known && dist2 && idx, "null pointer");
    CMF_REQUIRE(b <= 65535, "batch > 65535");
    launch_knn<0>(b, n, m, 3, unknown, known, dist2, idx, (cudaStream_t)stream);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

extern "C" int cmf_knn_point(int b, int n, int s, int k, const float *xyz, const float *new_xyz,
                             int *idx, float *dist, void *stream) {
    CMF_REQUIRE(b >= 0 && n >= 0 && s >= 0, "negative size");
    CMF_REQUIRE(k >= 1 && k <= 32, "k must be in [1,32]");
    CMF_REQUIRE(k <= n || b == 0 || s == 0, "k > number of candidates (torch.topk raises too)");
    if (b == 0 || s == 0) return CMF_OK;
    CMF_REQUIRE(xyz && new_xyz && idx, "null pointer");
    CMF_REQUIRE(b <= 65535, "batch > 65535");
    launch_knn<1>(b, s, n, k, new_xyz, xyz, dist, idx, (cudaStream_t)stream);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// ------------------------------------------------------------------------------------------------
// grouping / gathering   (reference: lib/src/group_points_gpu.cu, sampling_gpu.cu:8-63)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
group_points_kernel(int c, int n, long long ps, const float *__restrict__ points,
                    const int *__restrict__ idx, float *__restrict__ out) {
    const int b = blockIdx.z;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ps) return;
    const int id = __ldg(idx + (size_t)b * ps + j);
    const float *src = points + (size_t)b * c * n + id;
    float *dst = out + (size_t)b * c * ps + j;
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y) dst[(size_t)ci * ps] = __ldg(src + (size_t)ci * n);
}

__global__ void __launch_bounds__(256)
group_points_grad_kernel(int c, int n, long long ps, const float *__restrict__ grad_out,
                         const int *__restrict__ idx, float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ps) return;
    const int id = __ldg(idx + (size_t)b * ps + j);
    const float *src = grad_out + (size_t)b * c * ps + j;
    float *dst = grad_points + (size_t)b * c * n + id;
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y) atomicAdd(dst + (size_t)ci * n, __ldg(src + (size_t)ci * ps));
}

static dim3 gather_grid(int b, int c, long long ps) {
    int bx = cmf_divup(ps, 256);
    long long want = 148LL * 16;                          // ~16 CTAs per SM worth of blocks
    long long per = (long long)bx * b;
    int cy = (int)((want + per - 1) / per);
    if (cy < 1) cy = 1;
    if (cy > c) cy = c;
    if (cy > 65535) cy = 65535;
    return dim3(bx, cy, b);
}

extern "C" int cmf_group_points(int b, int c, int n, int npoints, int nsample,
                                const float *points, const int *idx, float *out, void *stream) {
    CMF_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, "negative size");
    long long ps = (long long)npoints * nsample;
    if (b == 0 || c == 0 || ps == 0) return CMF_OK;
    CMF_REQUIRE(points && idx && out, "null pointer");
    CMF_REQUIRE(b <= 65535, "batch > 65535");
    group_points_kernel<<<gather_grid(b, c, ps), 256, 0, (cudaStream_t)stream>>>(c, n, ps, points, idx, out);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

extern "C" int cmf_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                     const float *grad_out, const int *idx, float *grad_points, void *stream) {
    CMF_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, "negative size");
    long long ps = (long long)npoints * nsample;
    if (b == 0 || c == 0 || ps == 0) return CMF_OK;
    CMF_REQUIRE(grad_out && idx && grad_points, "null pointer");
    CMF_REQUIRE(b <= 65535, "batch > 65535");
    group_points_grad_kernel<<<gather_grid(b, c, ps), 256, 0, (cudaStream_t)stream>>>(c, n, ps, grad_out, idx, grad_points);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// gather_points is group_points with nsample = 1 (out (B,C,M) = points[b,c,idx[b,m]])
extern "C" int cmf_gather_points(int b, int c, int n, int npoints,
                                 const float *points, const int *idx, float *out, void *stream) {
    return cmf_group_points(b, c, n, npoints, 1, points, idx, out, stream);
}
extern "C" int cmf_gather_points_grad(int b, int c, int n, int npoints,
                                      const float *grad_out, const int *idx, float *grad_points, void *stream) {
    return cmf_group_points_grad(b, c, n, npoints, 1, grad_out, idx, grad_points, stream);
}

// ------------------------------------------------------------------------------------------------
// three_interpolate   (reference: lib/src/interpolate_gpu.cu:149-214)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
three_interpolate_kernel(int c, int m, int n, const float *__restrict__ points, const int *__restrict__ idx,
                         const float *__restrict__ weight, float *__restrict__ out) {
    const int b = blockIdx.z;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float *w = weight + ((size_t)b * n + p) * 3;
    const int *ix = idx + ((size_t)b * n + p) * 3;
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const int i0 = __ldg(ix), i1 = __ldg(ix + 1), i2 = __ldg(ix + 2);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y) {
        const float *pt = points + ((size_t)b * c + ci) * m;
        // nvcc's contraction of w0*p0 + w1*p1 + w2*p2 in the reference: fma(w2,p2, fma(w0,p0, w1*p1))
        out[((size_t)b * c + ci) * n + p] = __fmaf_rn(w2, __ldg(pt + i2), __fmaf_rn(w0, __ldg(pt + i0), __fmul_rn(w1, __ldg(pt + i1))));
    }
}

__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out, const int *__restrict__ idx,
                              const float *__restrict__ weight, float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float *w = weight + ((size_t)b * n + p) * 3;
    const int *ix = idx + ((size_t)b * n + p) * 3;
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const int i0 = __ldg(ix), i1 = __ldg(ix + 1), i2 = __ldg(ix + 2);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y) {
        const float g = __ldg(grad_out + ((size_t)b * c + ci) * n + p);
        float *gp = grad_points + ((size_t)b * c + ci) * m;
        atomicAdd(gp + i0, __fmul_rn(g, w0));
        atomicAdd(gp + i1, __fmul_rn(g, w1));
        atomicAdd(gp + i2, __fmul_rn(g, w2));
    }
}

extern "C" int cmf_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                     const float *weight, float *out, void *stream) {
    CMF_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0, "negative size");
    if (b == 0 || c == 0 || n == 0) return CMF_OK;
    CMF_REQUIRE(points && idx && weight && out, "null pointer");
    CMF_REQUIRE(b <= 65535, "batch > 65535");
    three_interpolate_kernel<<<gather_grid(b, c, n), 256, 0, (cudaStream_t)stream>>>(c, m, n, points, idx, weight, out);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

extern "C" int cmf_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                          const float *weight, float *grad_points, void *stream) {
    CMF_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0, "negative size");
    if (b == 0 || c == 0 || n == 0) return CMF_OK;
    CMF_REQUIRE(grad_out && idx && weight && grad_points, "null pointer");
    CMF_REQUIRE(b <= 65535, "batch > 65535");
    three_interpolate_grad_kernel<<<gather_grid(b, c, n), 256, 0, (cudaStream_t)stream>>>(c, n, m, grad_out, idx, weight, grad_points);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// ------------------------------------------------------------------------------------------------
// furthest point sampling   (reference: lib/src/sampling_gpu.cu:93-209)
// One CTA per cloud with the reference's thread count bs = 2^floor(log2 n) (cuda_utils.h:9-13) so the
// per-thread scan order is the reference's.  The reference's smem tree keeps the LOWER slot on ties
// (__update, sampling_gpu.cu:86-91) at every level (half = bs/2, bs/4, ..., 1): the last level pits even
// against odd thread ids, the one before it (tid mod 4) 0 against 2, ... so among equal maxima the winner is
// the thread with the smallest BIT-REVERSED id.  (max value, then min bitrev(tid)) is a total order, hence
// associative: a shuffle butterfly + one cross-warp step yields the identical index.
// ------------------------------------------------------------------------------------------------
static int fps_threads(int work_size) {
    const int pow_2 = (int)(std::log(static_cast<double>(work_size)) / std::log(2.0));   // same expression as the reference
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

__global__ void __launch_bounds__(1024)
fps_kernel(int n, int m, const float *__restrict__ dataset, float *__restrict__ temp, int *__restrict__ idxs) {
    if (m <= 0) return;
    __shared__ float sv[32];
    __shared__ unsigned st[32];
    __shared__ int si[32];
    __shared__ int s_old;
    const int b = blockIdx.x, tid = threadIdx.x, bs = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = (bs + 31) >> 5;
    dataset += (size_t)b * n * 3;
    temp += (size_t)b * n;
    idxs += (size_t)b * m;
    int old = 0;
    if (tid == 0) idxs[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = dataset[old * 3], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
        float best = -1.f; int besti = 0;
        for (int k = tid; k < n; k += bs) {
            float d = cmf_sqdist_ref(dataset[k * 3], dataset[k * 3 + 1], dataset[k * 3 + 2], x1, y1, z1);
            float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            if (d2 > best) { best = d2; besti = k; }
        }
        // reduce (best desc, tid asc) carrying besti
        float v = best; unsigned t = __brev((unsigned)tid); int bi = besti;
        const unsigned full = (bs >= 32) ? 0xffffffffu : ((1u << bs) - 1u);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            float ov = __shfl_xor_sync(full, v, off);
            unsigned ot = __shfl_xor_sync(full, t, off);
            int oi = __shfl_xor_sync(full, bi, off);
            bool valid = (lane ^ off) < bs;                        // bs < 32: partner may not exist
            if (valid && (ov > v || (ov == v && ot < t))) { v = ov; t = ot; bi = oi; }
        }
        if (lane == 0) { sv[warp] = v; st[warp] = t; si[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            float v2 = lane < nwarp ? sv[lane] : -2.f;
            unsigned t2 = lane < nwarp ? st[lane] : 0xffffffffu; int i2 = lane < nwarp ? si[lane] : 0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                float ov = __shfl_xor_sync(full, v2, off);
                unsigned ot = __shfl_xor_sync(full, t2, off);
                int oi = __shfl_xor_sync(full, i2, off);
                bool valid = (lane ^ off) < bs;                    // bs < 32: only bs lanes exist
                if (valid && (ov > v2 || (ov == v2 && ot < t2))) { v2 = ov; t2 = ot; i2 = oi; }
            }
            if (lane == 0) { s_old = i2; idxs[j] = i2; }
        }
        __syncthreads();
        old = s_old;
    }
}

extern "C" int cmf_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp, int *idxs, void *stream) {
    CMF_REQUIRE(b >= 0 && n >= 1 && m >= 0, "need n >= 1, b,m >= 0");
    if (b == 0 || m == 0) return CMF_OK;
    CMF_REQUIRE(dataset && temp && idxs, "null pointer");
    fps_kernel<<<b, fps_threads(n), 0, (cudaStream_t)stream>>>(n, m, dataset, temp, idxs);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// ------------------------------------------------------------------------------------------------
// Kernel-density estimate of the training losses (utils/util.py:172-182 compute_density_loss; caller losses/radar_loss.py:39-40):
//   density[b][i] = mean_j exp(-d2(xyz1_i, xyz2_j) / (2 bw^2)) / (2.5 bw),  d2 = the clamped expanded form of utils/util.py:148-170.
// One thread per query, candidates staged in shared memory; the reference materialises the (B,N,M) distance matrix.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
kde_density_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2, float inv2bw2, float norm, float *__restrict__ out) {
    __shared__ float4 sc[512];
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    const float *q = xyz1 + ((size_t)b * n + (i < n ? i : 0)) * 3;
    const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
    const float nq = cmf_sqnorm3(qx, qy, qz);
    float acc = 0.f;
    for (int base = 0; base < m; base += 512) {
        const int cn = min(512, m - base);
        __syncthreads();
        for (int j = threadIdx.x; j < cn; j += blockDim.x) {
            const float *c = xyz2 + ((size_t)b * m + base + j) * 3;
            const float x = __ldg(c), y = __ldg(c + 1), z = __ldg(c + 2);
            sc[j] = make_float4(x, y, z, cmf_sqnorm3(x, y, z));
        }
        __syncthreads();
        for (int j = 0; j < cn; ++j) {
            const float4 c = sc[j];
            acc += __expf(-cmf_sqdist_expanded(qx, qy, qz, nq, c.x, c.y, c.z, c.w) * inv2bw2) * norm;
        }
    }
    if (i < n) out[(size_t)b * n + i] = acc / (float)m;
}

extern "C" int cmf_kde_density(int b, int n, int m, const float *xyz1, const float *xyz2, float bandwidth, float *density, void *stream) {
    CMF_REQUIRE(b >= 0 && n >= 0 && m >= 0, "negative size");
    if (b == 0 || n == 0) return CMF_OK;
    CMF_REQUIRE(m >= 1 && bandwidth > 0.f, "need at least one candidate and a positive bandwidth");
    CMF_REQUIRE(xyz1 && xyz2 && density, "null pointer");
    CMF_REQUIRE(b <= 65535, "batch > 65535");
    kde_density_kernel<<<dim3(cmf_divup(n, 128), b), 128, 0, (cudaStream_t)stream>>>(n, m, xyz1, xyz2, 1.f / (2.f * bandwidth * bandwidth), 1.f / (2.5f * bandwidth), density);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}
