// eval_kernels.cu -- the evaluation metrics of the reference's eval loop (utils/eval_util.py:42-117, main_util.py:175-193) as device
// reductions: each call adds one batch's SUMS into a small double array that stays on the GPU, so an evaluation run reads back (or
// all-reduces across ranks) a few dozen doubles once, instead of four .cpu().numpy() round trips and a host loop per batch.
// The Python mirror (cmflow_b200/eval_util.py) turns sums into the reference's dictionaries.
#include <math.h>

#include "cmf_common.cuh"

namespace {

__device__ __forceinline__ double block_sum(double v, double *red /* [32] */) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    }
    return t;        // valid in thread 0
}

// x/y/z measurement resolution of a point for a sensor with (range, elevation, azimuth) resolution `res` (eval_util.py:4-40):
// float32 spherical coordinates and gradients as numpy computes them, products with the float64 resolution vector in double.
__device__ __forceinline__ double cartesian_res_sum(float x, float y, float z, double r_res, double th_res, double ph_res) {
    const float r = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    const float theta = asinf(__fdiv_rn(z, r));
    const float phi = atan2f(y, x);
    const float ct = cosf(theta), st = sinf(theta), cp = cosf(phi), sp = sinf(phi);
    const float gx0 = __fmul_rn(cp, ct), gx1 = __fmul_rn(__fmul_rn(-r, st), cp), gx2 = __fmul_rn(__fmul_rn(-r, ct), sp);
    const float gy0 = __fmul_rn(sp, ct), gy1 = __fmul_rn(__fmul_rn(-r, sp), st), gy2 = __fmul_rn(__fmul_rn(r, ct), cp);
    const float gz0 = st, gz1 = __fmul_rn(r, ct);
    const double xr = fabs((double)gx0) * r_res + fabs((double)gx1) * th_res + fabs((double)gx2) * ph_res;
    const double yr = fabs((double)gy0) * r_res + fabs((double)gy1) * th_res + fabs((double)gy2) * ph_res;
    const double zr = fabs((double)gz0) * r_res + fabs((double)gz1) * th_res;
    return xr + yr + zr;
}

// sums[0..10]: points, sum error, #accs, #accr, sum re_error, sum re_error[mask==0], #(mask==0), sum re_error[mask==1], #(mask==1), #sas, #ras
__global__ void __launch_bounds__(256)
eval_scene_flow_kernel(long long total, int n, const float *__restrict__ pc /* (B,3,N) */, const float *__restrict__ pred /* (B,N,3) */,
                       const float *__restrict__ labels /* (B,N,3) */, const float *__restrict__ mask /* (B,N) */,
                       double r_res, double th_res, double ph_res, double *__restrict__ sums) {
    __shared__ double red[32];
    double acc[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) acc[i] = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / n; const int i = (int)(t - b * n);
        const float *p = pc + (size_t)b * 3 * n;
        const float x = p[i], y = p[n + i], z = p[2 * n + i];
        const float *pr = pred + (size_t)t * 3, *lb = labels + (size_t)t * 3;
        const float d0 = __fsub_rn(pr[0], lb[0]), d1 = __fsub_rn(pr[1], lb[1]), d2 = __fsub_rn(pr[2], lb[2]);
        // np.sqrt(np.sum((pred - labels)**2, 2) + 1e-20) in float32 (eval_util.py:48,52)
        const float err = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), 1e-20f));
        const float len = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(lb[0], lb[0]), __fmul_rn(lb[1], lb[1])), __fmul_rn(lb[2], lb[2])), 1e-20f));
        const float rel = __fdiv_rn(err, len);
        const double res_r = sqrt(cartesian_res_sum(x, y, z, r_res, th_res, ph_res) + 1e-20);                       // :62-63
        const double res_l = sqrt(cartesian_res_sum(x, y, z, 0.04, 0.4 * M_PI / 180.0, 0.08 * M_PI / 180.0) + 1e-20);  // :13-15, 64-65
        const double re = (double)err / (res_r / res_l);                                                              // :68
        const float mk = mask[t];
        acc[0] += 1.0;
        acc[1] += (double)err;
        acc[2] += (err <= 0.05f || rel <= 0.05f) ? 1.0 : 0.0;                                                         // :58
        acc[3] += (err <= 0.10f || rel <= 0.10f) ? 1.0 : 0.0;                                                         // :59
        acc[4] += re;
        if (mk == 0.f) { acc[5] += re; acc[6] += 1.0; }                                                               // :70
        if (mk == 1.f) { acc[7] += re; acc[8] += 1.0; }                                                               // :71
        const double rre = re / (double)len;
        acc[9] += (re <= 0.10 || rre <= 0.10) ? 1.0 : 0.0;                                                            // :75
        acc[10] += (re <= 0.20 || rre <= 0.20) ? 1.0 : 0.0;                                                           // :76
    }
#pragma unroll
    for (int i = 0; i < 11; ++i) {
        const double s = block_sum(acc[i], red);
        if (threadIdx.x == 0 && s != 0.0) atomicAdd(sums + i, s);
    }
}

// counts[0..3] += tp, tn, fp, fn (eval_util.py:103-106)
__global__ void __launch_bounds__(256)
eval_motion_seg_kernel(long long total, const float *__restrict__ pre, const float *__restrict__ gt, double *__restrict__ counts) {
    __shared__ double red[32];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const float p = pre[t], g = gt[t];
        acc[0] += (p == 1.f && g == 1.f); acc[1] += (p == 0.f && g == 0.f);
        acc[2] += (p == 1.f && g == 0.f); acc[3] += (p == 0.f && g == 1.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double s = block_sum(acc[i], red);
        if (threadIdx.x == 0 && s != 0.0) atomicAdd(counts + i, s);
    }
}

// sums[0..2] += pairs, sum |t(E)|, sum rotation angle of E in degrees; E = gt^-1 pred (odometry_util.py:62-117, 133-138)
__global__ void eval_rpe_kernel(int b, const float *__restrict__ gt, const float *__restrict__ pred, double *__restrict__ sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    double G[3][4], P[3][4];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) { G[r][c] = gt[(size_t)i * 16 + r * 4 + c]; P[r][c] = pred[(size_t)i * 16 + r * 4 + c]; }
    // se3_inverse(G) = [G_R^T, -G_R^T G_t];  E = inverse(G) . P
    double E[3][3], et[3];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) E[r][c] = G[0][r] * P[0][c] + G[1][r] * P[1][c] + G[2][r] * P[2][c];
        et[r] = G[0][r] * (P[0][3] - G[0][3]) + G[1][r] * (P[1][3] - G[1][3]) + G[2][r] * (P[2][3] - G[2][3]);
    }
    const double tn = sqrt(et[0] * et[0] + et[1] * et[1] + et[2] * et[2]);
    // rotation angle: |rotvec| = atan2(|axis part| , cos) -- well conditioned near the identity, where acos((tr-1)/2) is not
    const double vx = 0.5 * (E[2][1] - E[1][2]), vy = 0.5 * (E[0][2] - E[2][0]), vz = 0.5 * (E[1][0] - E[0][1]);
    const double ang = atan2(sqrt(vx * vx + vy * vy + vz * vz), 0.5 * (E[0][0] + E[1][1] + E[2][2] - 1.0)) * 180.0 / M_PI;
    atomicAdd(sums + 0, 1.0); atomicAdd(sums + 1, tn); atomicAdd(sums + 2, ang);
}

}  // namespace

extern "C" int cmf_eval_scene_flow_sums(int b, int n, const float *pc, const float *pred, const float *labels, const float *mask,
                                        double r_res, double theta_res, double phi_res, double *sums11, void *stream) {
    CMF_REQUIRE(b >= 0 && n >= 0, "negative size");
    if (b == 0 || n == 0) return CMF_OK;
    CMF_REQUIRE(pc && pred && labels && mask && sums11, "null pointer");
    const long long total = (long long)b * n;
    const int grid = (int)(cmf_divup(total, 256) < 1184 ? cmf_divup(total, 256) : 1184);
    eval_scene_flow_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(total, n, pc, pred, labels, mask, r_res, theta_res, phi_res, sums11);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

extern "C" int cmf_eval_motion_seg_counts(long long total, const float *pre, const float *gt, double *counts4, void *stream) {
    CMF_REQUIRE(total >= 0, "negative size");
    if (total == 0) return CMF_OK;
    CMF_REQUIRE(pre && gt && counts4, "null pointer");
    const int grid = (int)(cmf_divup(total, 256) < 1184 ? cmf_divup(total, 256) : 1184);
    eval_motion_seg_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(total, pre, gt, counts4);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

extern "C" int cmf_eval_rpe_sums(int b, const float *gt_trans, const float *pred_trans, double *sums3, void *stream) {
    CMF_REQUIRE(b >= 0, "negative size");
    if (b == 0) return CMF_OK;
    CMF_REQUIRE(gt_trans && pred_trans && sums3, "null pointer");
    eval_rpe_kernel<<<cmf_divup(b, 128), 128, 0, (cudaStream_t)stream>>>(b, gt_trans, pred_trans, sums3);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}
