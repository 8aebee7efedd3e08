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

#include "tc_gemm.cuh"

namespace {

// Geometry.  One pipeline STAGE holds K = 16 floats (64-byte rows, 64B swizzle): A 128x16 hi+lo (16 KB) + B 256x16 hi+lo (32 KB)
// = 48 KB, four stages in flight (192 KB).  Finer stages than the swizzle-128 variant buy latency tolerance: the ring
// holds the same bytes but a slot is recycled every 6 MMAs (768 clk) instead of every 12.
// TcArgs.k_blocks counts 32-float blocks (what a producer thread handles per iteration = two stages).
constexpr int BM = 128, BN = 256, SK = 16, PK = 32;
constexpr int TILE_A_FLOATS = BM * SK;              // 2048 floats =  8 KB
constexpr int TILE_B_FLOATS = BN * SK;              // 4096 floats = 16 KB
constexpr int STAGE_BYTES = (2 * TILE_A_FLOATS + 2 * TILE_B_FLOATS) * 4;   // 48 KB
constexpr int NSTAGE = 4;
constexpr int NTHREADS = 512;
constexpr int SMALL_BYTES = 512 * 16;               // rel-xyz / direction weights (C x 4 floats) of the gather producers
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + SMALL_BYTES;
constexpr unsigned SPIN_LIMIT = 20000u;         // x 1 ms suspend hint = 20 s before a stuck wait traps

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (!done && ++spins > SPIN_LIMIT) __trap();        // never hang the GPU: fail the launch instead
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// hi = round-to-nearest TF32 (low 13 mantissa bits zero), lo = x - hi (exact)
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    lo = x - hi;
}

// Float offset of element (row, kk) inside a K-major [rows x 16] tile with the 64-byte swizzle:
// rows are 64 bytes, the 16-byte chunk index (2 bits) is XORed with address bits [7,9) = (row >> 1) & 3.
__host__ __device__ __forceinline__ int sw_off(int row, int kk) { return row * SK + ((((kk >> 2) ^ ((row >> 1) & 3))) << 2) + (kk & 3); }

// K-major, 64B-swizzled operand tile (tile base 1024-aligned): 8-row groups are 512 bytes apart.
//   bits [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=32) | [46,48) version=1 | [61,64) layout=4 (SWIZZLE_64B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M=128, N=256:
//   c_format[4,6)=1 (F32) | a_format[7,10)=2 (TF32) | b_format[10,13)=2 | n_dim[17,23)=N>>3 | m_dim[24,29)=M>>4
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : 0.1f * v;
    return v;
}

template <int KSAMP>
__device__ __forceinline__ void maxk_groups(const uint32_t (&r)[32], float bias, long long cbase, int m, bool m_ok, const TcArgs &a) {
#pragma unroll
    for (int g0 = 0; g0 < 32; g0 += KSAMP) {
        float mx = 0.f;                                            // relu output >= 0
#pragma unroll
        for (int e = 0; e < KSAMP; ++e) mx = fmaxf(mx, __uint_as_float(r[g0 + e]) + bias);
        const long long c = cbase + g0;
        if (c < a.cols && m_ok) a.Out[(size_t)(c / KSAMP) * a.ldo + m] = mx;
    }
}

struct RowCtx {          // per-producer-thread description of its activation row for the current tile
    bool valid;
    const float *src0, *src1;      // PLAIN: src0 ; FC_H1: src0 = U1 row (centre point), src1 = U2 row (neighbour) ; SC2_Y1: src1 = P row
    float dx, dy, dz;
};

__device__ __forceinline__ RowCtx make_row(const TcArgs &a, long long c) {
    RowCtx r;
    r.valid = c < a.cols;
    r.src0 = r.src1 = nullptr; r.dx = r.dy = r.dz = 0.f;
    if (!r.valid) return r;
    if (a.prod == TC_PROD_PLAIN) { r.src0 = a.X + (size_t)c * a.ldx; return r; }
    const long long bi = c / a.ksamp;
    const int kk = (int)(c - bi * a.ksamp);
    const int b = (int)(bi / a.n_pts), i = (int)(bi - (long long)b * a.n_pts);
    const int j = __ldg(a.nbr + (size_t)bi * a.nbr_ld + a.nbr_off + kk);
    const float *pq = a.xyz_q + (size_t)b * 3 * a.n_pts, *pc = a.xyz_c + (size_t)b * 3 * a.n_pts;
    r.dx = __fsub_rn(__ldg(pc + j), __ldg(pq + i));
    r.dy = __fsub_rn(__ldg(pc + a.n_pts + j), __ldg(pq + a.n_pts + i));
    r.dz = __fsub_rn(__ldg(pc + 2 * a.n_pts + j), __ldg(pq + 2 * a.n_pts + i));
    r.src0 = a.U1 ? a.U1 + (size_t)bi * 512 : nullptr;
    r.src1 = a.U2 + ((size_t)b * a.n_pts + j) * a.ld_u2 + a.off_u2;
    return r;
}

// One 32-float block of this thread's (gathered) row, BEFORE the tf32 split: 128 contiguous bytes.  For FC_H1 the
// centre-point row is shared by the 8 consecutive rows of a point: each of those 8 lanes fetches one 16-byte chunk of it
// (`u`) and the chunks are exchanged by shuffle at store time.
template <int PROD>
__device__ __forceinline__ void load_row(const RowCtx &r, int kb, int sub, float4 (&v)[8], float4 &u) {
    if (!r.valid) return;
    const float *src = (PROD == TC_PROD_PLAIN ? r.src0 : r.src1) + kb * PK;
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __ldg(reinterpret_cast<const float4 *>(src) + q);
    if (PROD == TC_PROD_FC_H1) u = __ldg(reinterpret_cast<const float4 *>(r.src0 + kb * PK) + sub);
}

__device__ __forceinline__ float small_term(const float4 *sW, int ch, const RowCtx &r) {
    const float4 w = sW[ch];                                   // shared-memory broadcast (all 32 lanes read the same channel)
    return fmaf(w.z, r.dz, fmaf(w.y, r.dy, w.x * r.dx));
}

// transform + split + swizzled store of half a 32-block (chunks q0..q0+3) into one stage's B tiles
template <int PROD>
__device__ __forceinline__ void store_half(const float4 *sW, const RowCtx &r, int kb, int row, int lane, int half,
                                           const float4 (&v)[8], const float4 &u, float *Bhi, float *Blo) {
    const int k0 = kb * PK + half * SK;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
        const int q = half * 4 + qq;
        float x[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
        if (PROD == TC_PROD_FC_H1) {
            const int srcl = (lane & ~7) + q;                     // the lane of this point's group that holds chunk q of the centre row
            const float uu[4] = {__shfl_sync(0xffffffffu, u.x, srcl), __shfl_sync(0xffffffffu, u.y, srcl),
                                 __shfl_sync(0xffffffffu, u.z, srcl), __shfl_sync(0xffffffffu, u.w, srcl)};
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = act_apply(uu[e] + x[e] + small_term(sW, k0 + qq * 4 + e, r), 2);
        } else if (PROD == TC_PROD_SC2_Y1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e] + small_term(sW, k0 + qq * 4 + e, r), 0.f);
        }
        if (!r.valid) { x[0] = x[1] = x[2] = x[3] = 0.f; }
        float4 h, l;
        split_tf32(x[0], h.x, l.x); split_tf32(x[1], h.y, l.y); split_tf32(x[2], h.z, l.z); split_tf32(x[3], h.w, l.w);
        const int off = sw_off(row, qq * 4);
        *reinterpret_cast<float4 *>(Bhi + off) = h;
        *reinterpret_cast<float4 *>(Blo + off) = l;
    }
}

template <int PROD>
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long col_tiles = (a.cols + BN - 1) / BN;
    const long long ntiles = col_tiles * a.m_blocks;
    const int nks = a.k_blocks * 2;                               // 16-float stages per tile

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
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int ks = 0; ks < nks; ++ks) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = base + stage * STAGE_BYTES;
                    const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + TILE_A_FLOATS * 4);
                    const uint64_t b_hi = make_desc(sa + 2 * TILE_A_FLOATS * 4), b_lo = make_desc(sa + 2 * TILE_A_FLOATS * 4 + TILE_B_FLOATS * 4);
#pragma unroll
                    for (int k8 = 0; k8 < SK / 8; ++k8) {
                        const uint64_t adv = (uint64_t)(k8 * 32 >> 4);      // 32 bytes per K=8 step inside the swizzle atom
                        tc_mma_tf32(d_tmem, a_lo + adv, b_hi + adv, IDESC, (ks | k8) ? 1u : 0u);
                        tc_mma_tf32(d_tmem, a_hi + adv, b_lo + adv, IDESC, 1u);
                        tc_mma_tf32(d_tmem, a_hi + adv, b_hi + adv, IDESC, 1u);
                    }
                    tc_commit(empty_bar(stage));                              // frees the smem stage when these MMAs retire
                    if (ks == nks - 1) tc_commit(tfull_bar(acc));             // accumulator complete
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
            // per-pair bias: track the pair of the current column incrementally (no division per element)
            long long pair = 0, pair_end = 0x7fffffffffffffffLL;
            float pb = 0.f;
            if (a.pbias) {
                pair = c0 / a.cols_per_pair; pair_end = (pair + 1) * (long long)a.cols_per_pair;
                if (m_ok) pb = __ldg(a.pbias + (size_t)pair * a.pb_ld + m);
            }
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < BN; cc += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cc, r);
                if (a.epi == TC_EPI_STORE && a.out_tiled) {
                    // next GEMM's B operand, already TF32-split and swizzled: tile (col_tile, 16-block = m/16), row = column in tile
                    float *tb = a.Out + ((size_t)ct * (a.M >> 4) + (m >> 4)) * (2 * TILE_B_FLOATS);
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const long long c = c0 + cc + e;
                        if (c >= pair_end) {
                            ++pair; pair_end += a.cols_per_pair;
                            if (c < a.cols) pb = __ldg(a.pbias + (size_t)pair * a.pb_ld + m);
                        }
                        const float v = c < a.cols ? act_apply(__uint_as_float(r[e]) + bias + pb, a.act) : 0.f;
                        float hi, lo;
                        split_tf32(v, hi, lo);
                        const int off = sw_off(cc + e, m & 15);
                        tb[off] = hi;
                        tb[TILE_B_FLOATS + off] = lo;
                    }
                } else if (a.epi == TC_EPI_STORE) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const long long c = c0 + cc + e;
                        if (c >= pair_end) {                                  // warp-uniform: crossed into the next frame pair
                            ++pair; pair_end += a.cols_per_pair;
                            if (m_ok && c < a.cols) pb = __ldg(a.pbias + (size_t)pair * a.pb_ld + m);
                        }
                        if (c < a.cols && m_ok) a.Out[(size_t)c * a.ldo + m] = act_apply(__uint_as_float(r[e]) + bias + pb, a.act);
                    }
                } else if (a.epi == TC_EPI_MAXK) {
                    // relu(acc + bias) then max over each group of `ksamp` consecutive rows (one point's neighbours); ksamp | 32
                    if (a.ksamp == 4) maxk_groups<4>(r, bias, c0 + cc, m, m_ok, a);
                    else if (a.ksamp == 8) maxk_groups<8>(r, bias, c0 + cc, m, m_ok, a);
                    else if (a.ksamp == 16) maxk_groups<16>(r, bias, c0 + cc, m, m_ok, a);
                    else maxk_groups<32>(r, bias, c0 + cc, m, m_ok, a);
                }
            }
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
                    for (int half = 0; half < 2; ++half) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        float *Bhi = reinterpret_cast<float *>(smem + stage * STAGE_BYTES + 2 * TILE_A_FLOATS * 4);
                        store_half<PROD>(sW, rc, kb, row, lane, half, v, u, Bhi, Bhi + TILE_B_FLOATS);
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

int cmf_launch_tc_gemm(const TcArgs &a, cudaStream_t st) {
    static int num_sms = 0;
    static bool attr_set = false;
    if (!attr_set) {
        CMF_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TC_PROD_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CMF_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TC_PROD_FC_H1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CMF_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TC_PROD_SC2_Y1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CMF_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TC_PROD_TILED>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        int dev = 0;
        CMF_CUDA(cudaGetDevice(&dev));
        CMF_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr_set = true;
    }
    if (a.cols <= 0 || a.m_blocks <= 0) return CMF_OK;
    if (a.out_tiled && ((a.M & 127) || a.epi != TC_EPI_STORE)) { cmf_set_error("tc_gemm: tiled output needs M % 128 == 0 and the STORE epilogue"); return CMF_ERR_INVALID; }
    if (a.epi == TC_EPI_MAXK && a.ksamp != 4 && a.ksamp != 8 && a.ksamp != 16 && a.ksamp != 32) { cmf_set_error("tc_gemm: MAXK needs ksamp in {4,8,16,32}"); return CMF_ERR_INVALID; }
    if (a.prod == TC_PROD_FC_H1 && a.ksamp != 8) { cmf_set_error("tc_gemm: the flow-embedding producer assumes 8 neighbours per point"); return CMF_ERR_INVALID; }
    const long long ntiles = ((a.cols + BN - 1) / BN) * a.m_blocks;
    const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
    if (a.prod == TC_PROD_PLAIN) tc_gemm_kernel<TC_PROD_PLAIN><<<grid, NTHREADS, SMEM_BYTES, st>>>(a);
    else if (a.prod == TC_PROD_FC_H1) tc_gemm_kernel<TC_PROD_FC_H1><<<grid, NTHREADS, SMEM_BYTES, st>>>(a);
    else if (a.prod == TC_PROD_TILED) tc_gemm_kernel<TC_PROD_TILED><<<grid, NTHREADS, SMEM_BYTES, st>>>(a);
    else tc_gemm_kernel<TC_PROD_SC2_Y1><<<grid, NTHREADS, SMEM_BYTES, st>>>(a);
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

// ---- test doorway: plain 3xTF32 GEMM through the C ABI (tests/test_gpu_tc_gemm.py) --------------------------------
static long long *g_test_dbg = nullptr;
extern "C" void cmf_test_tc_set_dbg(long long *dbg) { g_test_dbg = dbg; }      // device buffer long long[grid][8] or NULL

extern "C" int cmf_test_tc_gemm(int M, int K, long long cols, const float *W, int ldw, const float *X, int ldx,
                                const float *bias, int act, float *Out, int ldo, float *scratch_tiles, void *stream) {
    CMF_REQUIRE(W && X && Out && scratch_tiles, "null pointer");
    CMF_REQUIRE((ldx & 3) == 0 && ldx >= cmf_divup(K, PK) * PK, "ldx must be a multiple of 4 and cover K padded to 32");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cmf_tc_tile_weights(W, ldw, M, K, scratch_tiles, st);
    if (rc) return rc;
    TcArgs a{};
    a.Wt = scratch_tiles; a.m_blocks = cmf_divup(M, BM); a.k_blocks = cmf_divup(K, PK); a.M = M; a.cols = cols;
    a.prod = TC_PROD_PLAIN; a.X = X; a.ldx = ldx;
    a.epi = TC_EPI_STORE; a.Out = Out; a.ldo = ldo; a.bias = bias; a.pbias = nullptr; a.act = act; a.cols_per_pair = 1;
    a.dbg = g_test_dbg;
    return cmf_launch_tc_auto(a, st);
}
extern "C" size_t cmf_test_tc_tiled_floats(int M, int K) { return cmf_tc_tiled_floats(M, K); }
