// tc_gemm.cu -- tcgen05 / TMEM GEMM for the dense 1x1-conv stacks (the part of the hot path that really is a GEMM).
//
//   Out[c][m] = epi( sum_k W[m][k] * B[c][k] ),   W: M x K weights,  B: cols x K activations (point-major rows)
//
// * 3xTF32 split precision: every fp32 operand x is split as hi = rna_tf32(x), lo = x - hi and the product is
//   accumulated as W_hi*B_hi + W_hi*B_lo + W_lo*B_hi in fp32 (TMEM) -- ~2^-22 relative per product, which keeps the
//   forward inside the 1e-4 parity bar where single-pass TF32 (2^-11) would not.
// * CTA tile = 128 output channels (one UMMA M) x 256 activation rows (UMMA N), K consumed in stages of 16 floats
//   (64-byte rows, 64B swizzle).  A (weights) is pre-tiled in global memory in exactly the shared-memory image
//   (hi tile then lo tile), so one cp.async.bulk (TMA engine, UBLKCP) per stage brings it in.
//   B is either PRODUCED by 8 warps, one thread per activation row (coalesced 128-byte row-slice loads, or the fused
//   neighbour gather + first-layer epilogue of the set-conv / flow-embedding; hi/lo split; swizzled st.shared), or --
//   when the previous tc GEMM's epilogue already wrote it split + swizzled ("tiled") -- bulk-copied like A.
// * Warp roles (512 threads): w0 bulk-copy issuer, w1 MMA issuer (one elected lane), w2 TMEM allocator,
//   w4-7 epilogue (TMEM lane quarter = warp%4), w8-15 B producers.  4-stage smem ring (4 x 48 KB), 2 TMEM accumulator
//   stages (2 x 256 columns) so the epilogue of tile t overlaps the MMAs of tile t+1.  Persistent over tiles.
#include <stdlib.h>

#define CMF_WD_TU 1
#include "tc_dev.cuh"

using namespace tcdev;

namespace {

// Geometry.  One pipeline STAGE holds 64-byte operand rows (64B swizzle): K = 16 floats (fmt 0, 3xTF32) or K = 32 halfs (fmt 1, 3xFP16):
// A 128 rows hi+lo (16 KB) + B 256 rows hi+lo (32 KB) = 48 KB, four stages in flight (192 KB).
// TcArgs.k_blocks counts 32-element blocks (what a producer thread handles per iteration = two stages in fmt 0, one in fmt 1).
constexpr int TILE_A_FLOATS = BM * SK;              // 2048 floats =  8 KB
constexpr int TILE_B_FLOATS = BN * SK;              // 4096 floats = 16 KB
constexpr int STAGE_BYTES = (2 * TILE_A_FLOATS + 2 * TILE_B_FLOATS) * 4;   // 48 KB
constexpr int NSTAGE = 4;
constexpr int NTHREADS = 512;
constexpr int SMALL_BYTES = 512 * 16;               // rel-xyz / direction weights (C x 4 floats) of the gather producers
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + SMALL_BYTES;
constexpr uint32_t IDESC_TF32 = make_idesc(BM, BN, 0), IDESC_F16 = make_idesc(BM, BN, 1);

// ISSUE DISCIPLINE (see tc_sc2.cu): the issuer warp runs the issue code in uniform control flow with warp-uniform operands, every tcgen05
// instruction predicated on `el`, the flag of its elected lane (under `if (lane == 0)` each MMA sat in an ELECT / R2UR.BROADCAST / branch loop).
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t el;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(el));
    return el;
}
__device__ __forceinline__ void tc_commit(uint32_t el, uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(el) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t el, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(el) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t el, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(el) : "memory");
}

template <int PROD, int F16>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_gemm_kernel(const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t bar0 = base + NSTAGE * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8 * s; };
    auto empty_bar = [&](int s) { return bar0 + 32 + 8 * s; };
    auto tfull_bar = [&](int s) { return bar0 + 64 + 8 * s; };
    auto tempty_bar = [&](int s) { return bar0 + 80 + 8 * s; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + NSTAGE * STAGE_BYTES + 128);
    float4 *sW = reinterpret_cast<float4 *>(smem + NSTAGE * STAGE_BYTES + 256);
    if (PROD == TC_PROD_FC_H1 || PROD == TC_PROD_SC2_Y1)
        for (int i = threadIdx.x; i < a.k_blocks * PK; i += NTHREADS) sW[i] = __ldg(reinterpret_cast<const float4 *>(a.Wsmall) + i);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;      // broadcast: warp-uniform role branches
    const long long col_tiles = (a.cols + BN - 1) / BN;
    const long long ntiles = col_tiles * a.m_blocks;
    constexpr int SPB = F16 ? 1 : 2;                              // pipeline stages per 32-element K block
    const int nks = a.k_blocks * SPB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(full_bar(s), PROD == TC_PROD_TILED ? 1 : 1 + 8);   // bulk-copy issuer (expect_tx) [+ one arrive per producer warp]
            mbar_init(empty_bar(s), 1);           // tcgen05.commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);           // tcgen05.commit
            mbar_init(tempty_bar(s), 4);          // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== bulk-copy issuer: A (pre-tiled weights) every stage; B too when the activations arrive pre-tiled =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int mb = (int)(t % a.m_blocks);
                const long long ct = t / a.m_blocks;
                for (int ks = 0; ks < nks; ++ks) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const float *src = a.Wt + ((size_t)mb * nks + ks) * (2 * TILE_A_FLOATS);
                    const uint32_t dst = base + stage * STAGE_BYTES;
                    if (PROD == TC_PROD_TILED) {
                        const float *bsrc = a.Xt + ((size_t)ct * nks + ks) * (2 * TILE_B_FLOATS);
                        mbar_arrive_expect_tx(full_bar(stage), (2 * TILE_A_FLOATS + 2 * TILE_B_FLOATS) * 4);
                        bulk_g2s(dst, src, 2 * TILE_A_FLOATS * 4, full_bar(stage));
                        bulk_g2s(dst + 2 * TILE_A_FLOATS * 4, bsrc, 2 * TILE_B_FLOATS * 4, full_bar(stage));
                    } else {
                        mbar_arrive_expect_tx(full_bar(stage), 2 * TILE_A_FLOATS * 4);
                        bulk_g2s(dst, src, 2 * TILE_A_FLOATS * 4, full_bar(stage));
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t el = elect_one();
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int ks = 0; ks < nks; ++ks) {
                mbar_wait(full_bar(stage), phase);
                __syncwarp();                                                 // converged warp: one elect-predicated MMA per pass of the warp
                tc_fence_after();
                {
                    const uint32_t sa = base + stage * STAGE_BYTES;
                    const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + TILE_A_FLOATS * 4);
                    const uint64_t b_hi = make_desc(sa + 2 * TILE_A_FLOATS * 4), b_lo = make_desc(sa + 2 * TILE_A_FLOATS * 4 + TILE_B_FLOATS * 4);
#pragma unroll
                    for (int k8 = 0; k8 < SK / 8; ++k8) {
                        const uint64_t adv = (uint64_t)(k8 * 32 >> 4);      // 32 bytes per K=8 step inside the swizzle atom
                        if (F16) {
                            tc_mma_f16(el, d_tmem, a_lo + adv, b_hi + adv, IDESC_F16, (ks | k8) ? 1u : 0u);
                            tc_mma_f16(el, d_tmem, a_hi + adv, b_lo + adv, IDESC_F16, 1u);
                            tc_mma_f16(el, d_tmem, a_hi + adv, b_hi + adv, IDESC_F16, 1u);
                        } else {
                            tc_mma_tf32(el, d_tmem, a_lo + adv, b_hi + adv, IDESC_TF32, (ks | k8) ? 1u : 0u);
                            tc_mma_tf32(el, d_tmem, a_hi + adv, b_lo + adv, IDESC_TF32, 1u);
                            tc_mma_tf32(el, d_tmem, a_hi + adv, b_hi + adv, IDESC_TF32, 1u);
                        }
                    }
                    tc_commit(el, empty_bar(stage));                          // frees the smem stage when these MMAs retire
                    if (ks == nks - 1) tc_commit(el, tfull_bar(acc));         // accumulator complete
                }
                __syncwarp();
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===== epilogue: TMEM -> registers -> global; this warp owns TMEM lanes [32*(warp%4), +32) =====
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int mb = (int)(t % a.m_blocks);
            const long long ct = t / a.m_blocks;
            const long long c0 = ct * BN;
            const int m = mb * BM + q * 32 + lane;
            const bool m_ok = m < a.M;
            const float bias = (a.bias && m_ok) ? __ldg(a.bias + m) : 0.f;
            EpiState es = epi_begin(a, c0, m, m_ok);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < BN; cc += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cc, r);
                epilogue_chunk(a, r, ct, c0, cc, m, m_ok, bias, es, TILE_B_FLOATS);
            }
            epi_end(a, es, m);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 8 && PROD != TC_PROD_TILED) {
        // ===== B operand producers: thread `row` builds activation row c0+row; one iteration = 32 floats = two stages =====
        const int row = threadIdx.x - 256;
        int stage = 0; uint32_t phase = 0;
        long long t = blockIdx.x;
        if (t < ntiles) {
            RowCtx rc = make_row(a, (t / a.m_blocks) * BN + row);
            float4 v[8], vn[8];
            float4 u = make_float4(0.f, 0.f, 0.f, 0.f), un = u;
            load_row<PROD>(rc, 0, lane & 7, v, u);
            while (true) {
                RowCtx rcn = rc;
                const long long tn = t + gridDim.x;
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    // prefetch the next 32-block's row slice (or the NEXT TILE's row context + first slice) while we wait for the stages
                    if (kb + 1 < a.k_blocks) load_row<PROD>(rc, kb + 1, lane & 7, vn, un);
                    else if (tn < ntiles) { rcn = make_row(a, (tn / a.m_blocks) * BN + row); load_row<PROD>(rcn, 0, lane & 7, vn, un); }
#pragma unroll
                    for (int half = 0; half < SPB; ++half) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        uint8_t *Bb = smem + stage * STAGE_BYTES + 2 * TILE_A_FLOATS * 4;
                        if (F16) store_row_f16<PROD>(sW, rc, kb, row, lane, v, u, Bb, Bb + TILE_B_FLOATS * 4);
                        else store_half<PROD>(sW, rc, kb, row, lane, half, v, u, reinterpret_cast<float *>(Bb), reinterpret_cast<float *>(Bb) + TILE_B_FLOATS);
                        fence_async_smem();                                   // generic-proxy writes -> visible to the tensor core (async proxy)
                        __syncwarp();
                        if (lane == 0) mbar_arrive(full_bar(stage));
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) v[qq] = vn[qq];
                    u = un;
                }
                if (tn >= ntiles) break;
                t = tn; rc = rcn;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// W (M x K, row-major, ld) -> tiles [m_block][16-block]{hi, lo}; tile element (r, kk) at sw_off(r, kk)
__global__ void tile_weights_kernel(const float *__restrict__ W, int ldw, int M, int K, int m_blocks, int nks, float *__restrict__ Wt) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)m_blocks * nks * TILE_A_FLOATS;
    if (t >= total) return;
    const int e = (int)(t % TILE_A_FLOATS);
    const long long tile = t / TILE_A_FLOATS;
    const int ks = (int)(tile % nks), mb = (int)(tile / nks);
    const int r = e / SK, kk = e % SK;
    const int m = mb * BM + r, k = ks * SK + kk;
    const float x = (m < M && k < K) ? W[(size_t)m * ldw + k] : 0.f;
    float hi, lo;
    split_tf32(x, hi, lo);
    const int off = sw_off(r, kk);
    float *dst = Wt + (size_t)tile * (2 * TILE_A_FLOATS);
    dst[off] = hi;
    dst[TILE_A_FLOATS + off] = lo;
}

// fmt 1: per-row power-of-two scale (a_inv[m] = 1/scale; padded rows get 1)
__global__ void weight_row_scale_kernel(const float *__restrict__ W, int ldw, int M, int K, int m_rows_padded, float *__restrict__ a_inv) {
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= m_rows_padded) return;
    float mx = 0.f;
    if (m < M)
        for (int k = lane; k < K; k += 32) mx = fmaxf(mx, fabsf(W[(size_t)m * ldw + k]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if (lane == 0) a_inv[m] = 1.f / pow2_scale(mx);
}
// W (M x K) -> tiles [m_block][32-block]{hi, lo}; tile = 128 rows x 32 halfs (64-byte rows), half element (r, kk) at byte sw_off_h(r, kk)
__global__ void tile_weights_f16_kernel(const float *__restrict__ W, int ldw, int M, int K, int m_blocks, int nks, const float *__restrict__ a_inv,
                                        float *__restrict__ Wt) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)m_blocks * nks * (BM * PK);
    if (t >= total) return;
    const int e = (int)(t % (BM * PK));
    const long long tile = t / (BM * PK);
    const int ks = (int)(tile % nks), mb = (int)(tile / nks);
    const int r = e / PK, kk = e % PK;
    const int m = mb * BM + r, k = ks * PK + kk;
    const float x = (m < M && k < K) ? W[(size_t)m * ldw + k] * (1.f / a_inv[m]) : 0.f;      // exact: power-of-two scale
    unsigned short hi, lo;
    split_f16(x, hi, lo);
    uint8_t *dst = reinterpret_cast<uint8_t *>(Wt + (size_t)tile * (2 * TILE_A_FLOATS));
    const int off = sw_off_h(r, kk);
    *reinterpret_cast<unsigned short *>(dst + off) = hi;
    *reinterpret_cast<unsigned short *>(dst + TILE_A_FLOATS * 4 + off) = lo;
}

template <int PROD>
cudaError_t set_smem1() {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<PROD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc_gemm_kernel<PROD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
}
template <int PROD>
void launch1(const TcArgs &a, int grid, cudaStream_t st) {
    if (a.fmt == 1) tc_gemm_kernel<PROD, 1><<<grid, NTHREADS, SMEM_BYTES, st>>>(a);
    else tc_gemm_kernel<PROD, 0><<<grid, NTHREADS, SMEM_BYTES, st>>>(a);
}

}  // namespace

size_t cmf_tc_act_tiled_floats(long long cols, int C) {
    return (size_t)((cols + BN - 1) / BN) * (size_t)(cmf_divup(C, PK) * 2) * 2 * TILE_B_FLOATS;
}

size_t cmf_tc_tiled_floats(int M, int K) {
    return (size_t)cmf_divup(M, BM) * (cmf_divup(K, PK) * 2) * 2 * TILE_A_FLOATS;
}

int cmf_tc_tile_weights(const float *W, int ldw, int M, int K, float *Wt, cudaStream_t st) {
    const int mb = cmf_divup(M, BM), nks = cmf_divup(K, PK) * 2;
    const long long total = (long long)mb * nks * TILE_A_FLOATS;
    tile_weights_kernel<<<cmf_divup(total, 256), 256, 0, st>>>(W, ldw, M, K, mb, nks, Wt);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

int cmf_tc_tile_weights_f16(const float *W, int ldw, int M, int K, float *Wt, float *a_inv, cudaStream_t st) {
    const int mb = cmf_divup(M, BM), nks = cmf_divup(K, PK);
    weight_row_scale_kernel<<<cmf_divup(mb * BM, 8), 256, 0, st>>>(W, ldw, M, K, mb * BM, a_inv);
    CMF_LAUNCH_CHECK();
    const long long total = (long long)mb * nks * (BM * PK);
    tile_weights_f16_kernel<<<cmf_divup(total, 256), 256, 0, st>>>(W, ldw, M, K, mb, nks, a_inv, Wt);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

int cmf_launch_tc_gemm(const TcArgs &a, cudaStream_t st) {
    // function attributes are per device: a process that drives several GPUs (the reference's nn.DataParallel mode) sets them on each
    static int num_sms_of[64];
    static bool attr_set_of[64];
    int dev = 0;
    CMF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { cmf_set_error("tc gemm: device ordinal %d out of range", dev); return CMF_ERR_STATE; }
    if (!attr_set_of[dev]) {
        CMF_CUDA(set_smem1<TC_PROD_PLAIN>());
        CMF_CUDA(set_smem1<TC_PROD_FC_H1>());
        CMF_CUDA(set_smem1<TC_PROD_SC2_Y1>());
        CMF_CUDA(set_smem1<TC_PROD_TILED>());
        CMF_CUDA(cudaDeviceGetAttribute(&num_sms_of[dev], cudaDevAttrMultiProcessorCount, dev));
        attr_set_of[dev] = true;
    }
    const int num_sms = num_sms_of[dev];
    if (a.cols <= 0 || a.m_blocks <= 0) return CMF_OK;
    if (a.out_tiled && ((a.M & 127) || a.epi != TC_EPI_STORE)) { cmf_set_error("tc_gemm: tiled output needs M % 128 == 0 and the STORE epilogue"); return CMF_ERR_INVALID; }
    if (a.epi == TC_EPI_MAXK && a.ksamp != 4 && a.ksamp != 8 && a.ksamp != 16 && a.ksamp != 32) { cmf_set_error("tc_gemm: MAXK needs ksamp in {4,8,16,32}"); return CMF_ERR_INVALID; }
    if (a.prod == TC_PROD_FC_H1 && a.ksamp != 8) { cmf_set_error("tc_gemm: the flow-embedding producer assumes 8 neighbours per point"); return CMF_ERR_INVALID; }
    if ((a.pbias || a.bs_mode || a.amax_out) && a.cols_per_pair <= 0) { cmf_set_error("tc_gemm: cols_per_pair must be set"); return CMF_ERR_INVALID; }
    if (a.amax_out && a.amax_group <= 0) { cmf_set_error("tc_gemm: amax_group must be positive"); return CMF_ERR_INVALID; }
    const long long ntiles = ((a.cols + BN - 1) / BN) * a.m_blocks;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    if (a.prod == TC_PROD_PLAIN) launch1<TC_PROD_PLAIN>(a, grid, st);
    else if (a.prod == TC_PROD_FC_H1) launch1<TC_PROD_FC_H1>(a, grid, st);
    else if (a.prod == TC_PROD_TILED) launch1<TC_PROD_TILED>(a, grid, st);
    else launch1<TC_PROD_SC2_Y1>(a, grid, st);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

int cmf_tc_pair_enabled() {
    static int use2 = -1;
    if (use2 < 0) { const char *e = getenv("CMF_TC2"); use2 = (e && e[0] == '0') ? 0 : 1; }
    return use2;
}

int cmf_launch_tc_auto(const TcArgs &a, cudaStream_t st) {
    if (cmf_tc_pair_enabled() && (a.M & 255) == 0 && (a.m_blocks & 1) == 0) return cmf_launch_tc_gemm2(a, st);
    return cmf_launch_tc_gemm(a, st);
}

// ---- test doorway: plain split-precision GEMM through the C ABI (tests/test_gpu_tc_gemm.py) --------------------------------
static long long *g_test_dbg = nullptr;
extern "C" void cmf_test_tc_set_dbg(long long *dbg) { g_test_dbg = dbg; }      // device buffer long long[grid][8] or NULL

extern "C" int cmf_test_tc_gemm_fmt(int fmt, int M, int K, long long cols, const float *W, int ldw, const float *X, int ldx,
                                    const float *bias, int act, float *Out, int ldo, float *scratch_tiles,
                                    int cols_per_pair, const float *amax_in, unsigned int *amax_out, void *stream) {
    CMF_REQUIRE(W && X && Out && scratch_tiles, "null pointer");
    CMF_REQUIRE(fmt == 0 || fmt == 1, "fmt must be 0 (3xTF32) or 1 (3xFP16)");
    CMF_REQUIRE((ldx & 3) == 0 && ldx >= cmf_divup(K, PK) * PK, "ldx must be a multiple of 4 and cover K padded to 32");
    cudaStream_t st = (cudaStream_t)stream;
    float *a_inv = scratch_tiles + cmf_tc_tiled_floats(M, K);
    int rc = fmt == 1 ? cmf_tc_tile_weights_f16(W, ldw, M, K, scratch_tiles, a_inv, st) : cmf_tc_tile_weights(W, ldw, M, K, scratch_tiles, st);
    if (rc) return rc;
    TcArgs a{};
    a.Wt = scratch_tiles; a.m_blocks = cmf_divup(M, BM); a.k_blocks = cmf_divup(K, PK); a.M = M; a.cols = cols;
    a.prod = TC_PROD_PLAIN; a.X = X; a.ldx = ldx;
    a.epi = TC_EPI_STORE; a.Out = Out; a.ldo = ldo; a.bias = bias; a.pbias = nullptr; a.act = act;
    a.cols_per_pair = cols_per_pair > 0 ? cols_per_pair : 1;
    a.fmt = fmt; a.a_inv = fmt == 1 ? a_inv : nullptr;
    if (fmt == 1 && amax_in) { a.bs_mode = 1; a.bs_src[0] = amax_in; a.bs_coef[0] = 1.f; }
    a.amax_out = amax_out; a.amax_group = 1 << 30; a.amax_ld = 0;
    a.dbg = g_test_dbg;
    return cmf_launch_tc_auto(a, st);
}
extern "C" int cmf_test_tc_gemm(int M, int K, long long cols, const float *W, int ldw, const float *X, int ldx,
                                const float *bias, int act, float *Out, int ldo, float *scratch_tiles, void *stream) {
    return cmf_test_tc_gemm_fmt(0, M, K, cols, W, ldw, X, ldx, bias, act, Out, ldo, scratch_tiles, 1, nullptr, nullptr, stream);
}
extern "C" size_t cmf_test_tc_tiled_floats(int M, int K) { return cmf_tc_tiled_floats(M, K) + (size_t)cmf_divup(M, BM) * BM; }

// installs the host-mapped watchdog record of this translation unit's kernels (tc_dev.cuh) on the current device
int cmf_wd_set_tc_gemm(unsigned long long *dev_ptr) {
    CMF_CUDA(cudaMemcpyToSymbol(tcdev::g_cmf_wd_record, &dev_ptr, sizeof(dev_ptr)));
    return CMF_OK;
}
