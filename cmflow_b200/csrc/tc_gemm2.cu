// tc_gemm2.cu -- CTA-PAIR (cta_group::2) variant of the tcgen05 3xTF32 GEMM for M % 256 == 0.
//
// Why: in the one-CTA kernel (tc_gemm.cu) every tcgen05.mma re-reads its A (4 KB) and B (8 KB) operand slices from shared
// memory, the three split-precision MMAs of a K=8 step triple that, and the producers' swizzled stores, the bulk copies and
// the gathers share the same 128 B/clk L1/shared-memory data path: ~2500 clk of that path per 32-deep K block against
// 1536 clk of tensor-pipe time -- the kernel is shared-memory-bandwidth bound.  A CTA pair computes a 256(M) x 256(N) tile with
// ONE instruction stream (tcgen05.mma.cta_group::2, M=256): each CTA holds its own 128 weight rows and only HALF of the
// activation rows (128), the hardware feeds both tensor cores from both halves.  Per CTA that is 4+4 KB instead of 4+8 KB per MMA,
// half the producer stores / gathers per MAC, and -- for M = 256 layers -- every activation row is produced exactly once.
//
// Cluster of 2 CTAs, 512 threads each: w0 bulk-copy issuer, w1 MMA issuer (leader CTA only), w2 TMEM allocator,
// w1 / w3 forwarders (peer CTA only: relay "my half of stage s is full" to the leader, alternating stages), w4-7 epilogue (own 128 accumulator rows),
// w8-15 producers (128 rows per CTA; lane = (row, 16-byte chunk) so that eight lanes read one 128-byte row slice: coalesced gathers).
// 6-stage ring of 32 KB stages (K = 16 per stage).
// Protocol (barriers at identical offsets in both CTAs):
//   full_local[s]  : local  -- bulk copy expect_tx + 8 producer warps                      (count 1 + 8, or 1 when B is bulk-copied)
//   peer_full[s]   : leader -- remote arrive by the peer's forwarder                        (count 1)
//   empty[s]       : both   -- tcgen05.commit.cta_group::2 multicast from the leader        (count 1)
//   tfull[a]       : both   -- commit multicast by the MMA issuer after the last K stage of a tile (count 1)
//   tempty[a]      : leader -- 4 local + 4 remote epilogue warps                            (count 8)
#define CMF_WD_TU 2
#include "tc_dev.cuh"

using namespace tcdev;

namespace {

constexpr int HALF_N = 128;                          // activation rows held by each CTA
constexpr int TILE_A_FLOATS = BM * SK;               // 2048 floats = 8 KB
constexpr int TILE_BH_FLOATS = HALF_N * SK;          // 2048 floats = 8 KB (local half of B)
constexpr int TILE_B_FLOATS = BN * SK;               // 4096 floats: the 256-row tile of the tiled activation FORMAT in global memory
constexpr int STAGE_BYTES = (2 * TILE_A_FLOATS + 2 * TILE_BH_FLOATS) * 4;   // 32 KB
constexpr int NTHREADS = 512;
// Shared-memory plan (bytes, after 1024-alignment):
//   bulk-copied activations (TILED): 6 stages x 32 KB                          | barriers 256 | aux 16 KB = WSUM WeightNet hidden vectors
//   produced activations:            4 stages x 32 KB + 72 KB cp.async staging | barriers 256 | aux 20 KB = rel-xyz weights 8 KB + row contexts 12 KB
// The staging ring decouples the producers from global-memory latency: every producer thread keeps up to PF-1 K-blocks of its own
// 16-byte row chunks in flight with cp.async (it reads back only what it copied itself, so cp.async.wait_group is the only
// synchronisation), instead of a one-block-ahead register prefetch that left the loop latency-bound (~2100 clk per block measured
// against 768 clk of tensor time).
constexpr int NSTAGE_TILED = 6, NSTAGE_PROD = 4;
constexpr int STG_BYTES = 72 * 1024;
constexpr int RING_BYTES = NSTAGE_TILED * STAGE_BYTES + 8192;       // = NSTAGE_PROD * STAGE_BYTES + STG_BYTES = 204800
static_assert(NSTAGE_PROD * STAGE_BYTES + STG_BYTES == RING_BYTES, "smem plan");
constexpr int AUX_BYTES = 20 * 1024;                               // rel-xyz weights 8 KB + 3 tiles of row contexts 12 KB (or 16 KB of WSUM vectors)
constexpr int SMEM_BYTES = RING_BYTES + 1024 + 256 + AUX_BYTES;
constexpr uint32_t IDESC2_TF32 = make_idesc(256, BN, 0), IDESC2_F16 = make_idesc(256, BN, 1);

// ISSUE DISCIPLINE (see tc_sc2.cu): the whole issuer warp runs the issue code in uniform control flow and every tcgen05 instruction is
// predicated on `el`, the flag of the lane elected at role start.  Under `if (lane == 0)` the compiler wraps each MMA in an
// ELECT / R2UR.BROADCAST / branch loop (~100 clk of issue per MMA).
// CAUTION: ptxas turns `@el tcgen05.mma` into ONE unguarded UTCHMMA per pass of the warp through the code (operands broadcast from the elected
// lane).  The issuer warp must therefore be CONVERGED there: if its lanes drift apart (per-lane polling results), every divergent group issues
// the MMA again -- with stale uniform registers (observed: launch failure).  Poll with warp-uniform votes / __syncwarp() before issuing.
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t el;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(el));
    return el;
}
__device__ __forceinline__ void tc_commit2_mc(uint32_t el, uint32_t bar) {      // arrive on `bar` in BOTH CTAs when all prior MMAs of the elected thread retire
    const uint16_t mask = 3;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar), "h"(mask), "r"(el) : "memory");
}
__device__ __forceinline__ void tc_mma2_tf32(uint32_t el, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(el) : "memory");
}

__device__ __forceinline__ void tc_mma2_f16(uint32_t el, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(el) : "memory");
}
// The hi x lo and hi x hi products use the same A tile back to back: the first keeps it in the tensor core's A collector, the second reads it
// from there instead of from shared memory (4 KB less on the SM's shared-memory port per K=16 step, see profiles/r01b_tc_kernels_ncu_full.md).
__device__ __forceinline__ void tc_mma2_f16_keep(uint32_t el, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, 1, 0;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(el) : "memory");
}
__device__ __forceinline__ void tc_mma2_f16_reuse(uint32_t el, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, 1, 0;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(el) : "memory");
}

// F16 = 0: 3xTF32, one pipeline stage = 16 floats of K;  F16 = 1: 3xFP16, one stage = 32 halfs of K (same bytes, see tc_dev.cuh)
template <int PROD, int F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
tc_gemm2_kernel(const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    constexpr int NSTAGE = PROD == TC_PROD_TILED ? NSTAGE_TILED : NSTAGE_PROD;
    const uint32_t bar0 = base + RING_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8 * s; };
    auto pfull_bar = [&](int s) { return bar0 + 48 + 8 * s; };
    auto empty_bar = [&](int s) { return bar0 + 96 + 8 * s; };
    auto tfull_bar = [&](int s) { return bar0 + 144 + 8 * s; };
    auto tempty_bar = [&](int s) { return bar0 + 160 + 8 * s; };
    auto h2full_bar = [&](int s) { return bar0 + 176 + 8 * s; };
    auto h2empty_bar = [&](int s) { return bar0 + 208 + 8 * s; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + RING_BYTES + 240);
    float *h2s = reinterpret_cast<float *>(smem + RING_BYTES + 256);                 // TILED + WSUM only
    float4 *sW = reinterpret_cast<float4 *>(smem + RING_BYTES + 256);                // gather producers only
    // rel-xyz / direction weights, transposed for the producers: sW[(kb*3 + comp)*8 + q] = {W[c][comp], c = kb*32 + q*4 .. +3}.  A quarter
    // warp (q = 0..7) reads 8 consecutive float4 = one 128-byte wavefront, broadcast to the four row groups (the channel-major float4
    // layout put the 8 chunks 64 bytes apart: 4-way bank conflicts, 16 wavefronts per load -- 55 % of the kernel's shared-memory traffic).
    if (PROD == TC_PROD_FC_H1 || PROD == TC_PROD_SC2_Y1)
        for (int i = threadIdx.x; i < a.k_blocks * 24; i += NTHREADS) {
            const int kb_ = i / 24, comp = (i >> 3) % 3, q_ = i & 7;
            const float *w = a.Wsmall + (size_t)(kb_ * PK + q_ * 4) * 4 + comp;
            sW[i] = make_float4(__ldg(w), __ldg(w + 4), __ldg(w + 8), __ldg(w + 12));
        }

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;      // broadcast: warp-uniform role branches for the compiler
    const long long t_start = a.dbg ? clock64() : 0;
    long long dw0 = 0, dw1 = 0, dw2 = 0;                             // per-role accumulated wait cycles (instrumented runs only)
#define TIMED(acc_, stmt_) do { if (a.dbg) { const long long c0_ = clock64(); stmt_; acc_ += clock64() - c0_; } else { stmt_; } } while (0)
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const long long col_tiles = (a.cols + BN - 1) / BN;
    const int m_pairs = a.m_blocks >> 1;
    const long long ntiles = col_tiles * m_pairs;
    const long long cl_id = blockIdx.x >> 1, n_cl = gridDim.x >> 1;
    constexpr int SPB = F16 ? 1 : 2;                              // pipeline stages per 32-element K block
    const int nks = a.k_blocks * SPB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(full_bar(s), PROD == TC_PROD_TILED ? 1 : 1 + 8);
            mbar_init(pfull_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 8); mbar_init(h2full_bar(s), 8); mbar_init(h2empty_bar(s), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();                         // both CTAs: barriers initialised, TMEM allocated, sW visible
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Row contexts of 3 tiles (current, next, next but one) in shared memory: gathered-row pointer (NULL = row out of range),
    // centre-row pointer (flow embedding only) and {dx, dy, dz, scale}.  Written by warp 2 two tiles ahead of the producers.
    const float **cs1 = reinterpret_cast<const float **>(smem + RING_BYTES + 256 + 8192);          // [3][128]
    const float **cs0 = cs1 + 3 * HALF_N;                                                           // [3][128]
    float4 *cgeo = reinterpret_cast<float4 *>(cs0 + 3 * HALF_N);                                    // [3][128]

    if (warp == 0) {
        // ===== bulk-copy issuer: this CTA's 128 weight rows (and, when pre-tiled, its half of the activation rows) =====
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long t = cl_id; t < ntiles; t += n_cl) {
                const int mb = (int)(t % m_pairs) * 2 + (int)rank;
                const long long ct = t / m_pairs;
                for (int ks = 0; ks < nks; ++ks) {
                    TIMED(dw0, mbar_wait(empty_bar(stage), phase ^ 1));
                    const float *src = a.Wt + ((size_t)mb * nks + ks) * (2 * TILE_A_FLOATS);
                    const uint32_t dst = base + stage * STAGE_BYTES;
                    if (PROD == TC_PROD_TILED) {
                        const float *bsrc = a.Xt + ((size_t)ct * nks + ks) * (2 * TILE_B_FLOATS) + rank * TILE_BH_FLOATS;
                        mbar_arrive_expect_tx(full_bar(stage), (2 * TILE_A_FLOATS + 2 * TILE_BH_FLOATS) * 4);
                        bulk_g2s(dst, src, 2 * TILE_A_FLOATS * 4, full_bar(stage));
                        bulk_g2s(dst + 2 * TILE_A_FLOATS * 4, bsrc, TILE_BH_FLOATS * 4, full_bar(stage));                                   // hi half
                        bulk_g2s(dst + 2 * TILE_A_FLOATS * 4 + TILE_BH_FLOATS * 4, bsrc + TILE_B_FLOATS, TILE_BH_FLOATS * 4, full_bar(stage));  // lo half
                    } else {
                        mbar_arrive_expect_tx(full_bar(stage), 2 * TILE_A_FLOATS * 4);
                        bulk_g2s(dst, src, 2 * TILE_A_FLOATS * 4, full_bar(stage));
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
            if (a.dbg) a.dbg[(size_t)blockIdx.x * 8 + 4] = dw0;
        }
    } else if (leader && warp == 1) {
        // ===== MMA issuer (leader CTA only): ONE warp drives both SMs' tensor cores =====
        // Round 1 alternated two issuer warps because the issue itself was slow (per-MMA ELECT / R2UR loops under `if (lane == 0)`, ~100 clk
        // each).  With the elect-predicated, warp-uniform issue the six MMAs and two commits of a stage take a few tens of cycles, so one
        // warp keeps the pipe fed -- and a single issuing thread is what makes the accumulation order (hence the result bits) reproducible:
        // MMAs of DIFFERENT warps reach the tensor pipe through their own sub-partition queues and may interleave differently from run to
        // run even when a shared-memory turn counter orders their issue (measured: run-to-run differences of a few ulp).
        const uint32_t el = elect_one();                            // the one lane of this warp that issues (and commits) every MMA
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long t = cl_id; t < ntiles; t += n_cl) {
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int ks = 0; ks < nks; ++ks) {
                if (ks == 0) { TIMED(dw0, mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1)); }     // the tile's first MMA overwrites the accumulator
                TIMED(dw1, mbar_wait(full_bar(stage), phase));                 // my half
                TIMED(dw2, mbar_wait_cluster(pfull_bar(stage), phase));        // the peer's half (relayed)
                __syncwarp();                                                  // converged warp: ptxas issues an elect-predicated MMA ONCE PER WARP PASS
                tc_fence_after();
                const uint32_t sa = base + stage * STAGE_BYTES;
                const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + TILE_A_FLOATS * 4);
                const uint64_t b_hi = make_desc(sa + 2 * TILE_A_FLOATS * 4), b_lo = make_desc(sa + 2 * TILE_A_FLOATS * 4 + TILE_BH_FLOATS * 4);
#pragma unroll
                for (int k8 = 0; k8 < SK / 8; ++k8) {
                    const uint64_t adv = (uint64_t)(k8 * 32 >> 4);
                    if (F16) {
                        tc_mma2_f16(el, d_tmem, a_lo + adv, b_hi + adv, IDESC2_F16, (ks | k8) ? 1u : 0u);
                        tc_mma2_f16_keep(el, d_tmem, a_hi + adv, b_lo + adv, IDESC2_F16);
                        tc_mma2_f16_reuse(el, d_tmem, a_hi + adv, b_hi + adv, IDESC2_F16);
                    } else {
                        tc_mma2_tf32(el, d_tmem, a_lo + adv, b_hi + adv, IDESC2_TF32, (ks | k8) ? 1u : 0u);
                        tc_mma2_tf32(el, d_tmem, a_hi + adv, b_lo + adv, IDESC2_TF32, 1u);
                        tc_mma2_tf32(el, d_tmem, a_hi + adv, b_hi + adv, IDESC2_TF32, 1u);
                    }
                }
                tc_commit2_mc(el, empty_bar(stage));
                if (ks == nks - 1) tc_commit2_mc(el, tfull_bar(acc));          // the accumulator is complete when the tile's last MMAs retire
                __syncwarp();
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (a.dbg && lane == 0) { a.dbg[(size_t)blockIdx.x * 8 + 1] = dw0; a.dbg[(size_t)blockIdx.x * 8 + 2] = dw1; a.dbg[(size_t)blockIdx.x * 8 + 3] = dw2; }
    } else if (warp == 2 && PROD != TC_PROD_TILED) {
        // ===== context filler (idle after the TMEM allocation): neighbour index -> row pointers, rel-xyz, fp16 scale for the tile two
        // ahead of the producers, so that those dependent global loads never sit on the producers' critical path =====
        auto fill_ctx = [&](long long tt, int buf) {
            RowCtx rc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) rc[k] = make_row(a, (tt / m_pairs) * BN + rank * HALF_N + lane + 32 * k);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = buf * HALF_N + lane + 32 * k;
                cs1[r] = rc[k].valid ? ((PROD == TC_PROD_PLAIN) ? rc[k].src0 : rc[k].src1) : nullptr;
                if (PROD == TC_PROD_FC_H1) cs0[r] = rc[k].src0;
                cgeo[r] = make_float4(rc[k].dx, rc[k].dy, rc[k].dz, rc[k].valid ? rc[k].scale : 0.f);
            }
        };
        long long t = cl_id;
        if (t < ntiles) {
            int buf = 0;
            fill_ctx(t, 0);
            if (t + n_cl < ntiles) fill_ctx(t + n_cl, 1);
            asm volatile("bar.sync 1, 288;" ::: "memory");
            while (true) {
                const long long tn = t + n_cl;
                if (tn + n_cl < ntiles) fill_ctx(tn + n_cl, buf == 0 ? 2 : buf - 1);    // the slot the previous tile has vacated
                if (tn >= ntiles) break;
                asm volatile("bar.sync 1, 288;" ::: "memory");                           // ends the producers' tile t; publishes tile t + 2
                t = tn; buf = buf == 2 ? 0 : buf + 1;
            }
        }
    } else if (!leader && (warp == 1 || warp == 3)) {
        // ===== forwarders (peer CTA: its warps 1 and 3, alternating stages): tell the leader when this CTA's half of a stage is complete =====
        // (the relay is a serial wait -> remote arrive per stage; a release.cluster arrive on it was the pair kernel's critical path)
        const int me = warp == 3 ? 1 : 0;
        int stage = 0; uint32_t phase = 0;
        uint32_t g = 0;
        for (long long t = cl_id; t < ntiles; t += n_cl)
            for (int ks = 0; ks < nks; ++ks, ++g) {
                if ((int)(g & 1u) == me) {
                    mbar_wait(full_bar(stage), phase);
                    if (lane == 0) mbar_arrive_remote_relaxed(pfull_bar(stage), 0);
                    __syncwarp();
                }
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
    } else if (warp >= 4 && warp < 8) {
        // ===== epilogue: this CTA's 128 accumulator rows x 256 columns =====
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long t = cl_id; t < ntiles; t += n_cl) {
            const int mb = (int)(t % m_pairs) * 2 + (int)rank;
            const long long ct = t / m_pairs;
            const long long c0 = ct * BN;
            const int m = mb * BM + q * 32 + lane;
            const bool m_ok = m < a.M;
            const float bias = (a.bias && m_ok) ? __ldg(a.bias + m) : 0.f;
            if (a.dbg && warp == 4 && lane == 0) dw1 += 1;                     // tiles processed
            EpiState es = epi_begin(a, c0, m, m_ok);
            float4 w3a = make_float4(0.f, 0.f, 0.f, 0.f), w3b = w3a; float w3c = 0.f;
            if (a.epi == TC_EPI_WSUM && m_ok) {                      // last WeightNet layer row of this thread's channel
                w3a = __ldg(reinterpret_cast<const float4 *>(a.wnA3 + (size_t)m * 8)); w3b = __ldg(reinterpret_cast<const float4 *>(a.wnA3 + (size_t)m * 8 + 4));
                w3c = __ldg(a.wna3 + m);
            }
            TIMED(dw0, mbar_wait_cluster(tfull_bar(acc), acc_phase));
            if (a.epi == TC_EPI_WSUM) mbar_wait(h2full_bar(acc), acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < BN; cc += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cc, r);
                if (a.epi == TC_EPI_WSUM) {
                    // h2s holds the WeightNet hidden vectors of column PAIRS interleaved ([pair][j][2]): one 128-bit broadcast read brings
                    // (h[e][j], h[e+1][j], h[e][j+1], h[e+1][j+1]), and the last WeightNet layer of two columns is 8 FFMA2 instead of 16 FFMA
                    const float4 *hv = reinterpret_cast<const float4 *>(h2s + ((size_t)acc * BN + cc) * 8);
                    if (es.track && c0 + cc >= es.pair_end) epi_advance(a, es, c0 + cc, m, m_ok);
                    const bool one_pair = (c0 + cc + 32 <= a.cols) && (c0 + cc + 32 <= es.pair_end);
                    const float slope = act_slope(a.act);
#pragma unroll
                    for (int g0 = 0; g0 < 32; g0 += 8) {             // one point = 8 consecutive columns (always inside one frame pair)
                        const long long c = c0 + cc + g0;
                        float inv = es.inv;
                        if (F16 && !one_pair && c < a.cols) inv = es.ainv * __frcp_rn(b_scale_of(a, div_i(c, a.cols_per_pair)));
                        const float2 inv2 = make_float2(inv, inv), bias2 = make_float2(bias, bias), slope2 = make_float2(slope, slope);
                        float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
                        for (int e = 0; e < 8; e += 2) {
                            const float4 *hp = hv + ((g0 + e) >> 1) * 4;                    // broadcast reads
                            const float4 q0 = hp[0], q1 = hp[1], q2 = hp[2], q3 = hp[3];
                            float2 w = make_float2(w3c, w3c);                               // per column the same fma chain as the scalar form
                            w = __ffma2_rn(make_float2(q0.x, q0.y), make_float2(w3a.x, w3a.x), w); w = __ffma2_rn(make_float2(q0.z, q0.w), make_float2(w3a.y, w3a.y), w);
                            w = __ffma2_rn(make_float2(q1.x, q1.y), make_float2(w3a.z, w3a.z), w); w = __ffma2_rn(make_float2(q1.z, q1.w), make_float2(w3a.w, w3a.w), w);
                            w = __ffma2_rn(make_float2(q2.x, q2.y), make_float2(w3b.x, w3b.x), w); w = __ffma2_rn(make_float2(q2.z, q2.w), make_float2(w3b.y, w3b.y), w);
                            w = __ffma2_rn(make_float2(q3.x, q3.y), make_float2(w3b.z, w3b.z), w); w = __ffma2_rn(make_float2(q3.z, q3.w), make_float2(w3b.w, w3b.w), w);
                            float2 v = __ffma2_rn(make_float2(__uint_as_float(r[g0 + e]), __uint_as_float(r[g0 + e + 1])), inv2, bias2);
                            const float2 vs = __fmul2_rn(v, slope2);
                            v = make_float2(fmaxf(v.x, vs.x), fmaxf(v.y, vs.y));           // act as max(v, slope * v)
                            sum2 = __ffma2_rn(make_float2(fmaxf(w.x, 0.f), fmaxf(w.y, 0.f)), v, sum2);
                        }
                        if (c < a.cols && m_ok) a.Out[(size_t)(c >> 3) * a.ldo + m] = sum2.x + sum2.y;
                    }
                } else {
                    epilogue_chunk(a, r, ct, c0, cc, m, m_ok, bias, es, TILE_B_FLOATS);
                }
            }
            tc_fence_before();
            __syncwarp();
            epi_end(a, es, m);
            if (lane == 0) {
                if (leader) mbar_arrive(tempty_bar(acc)); else mbar_arrive_remote(tempty_bar(acc), 0);
                if (a.epi == TC_EPI_WSUM) mbar_arrive(h2empty_bar(acc));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (a.dbg && warp == 4 && lane == 0) { a.dbg[(size_t)blockIdx.x * 8 + 6] = dw0; a.dbg[(size_t)blockIdx.x * 8 + 7] = dw1; }
    } else if (warp >= 8 && PROD != TC_PROD_TILED) {
        // ===== producers (256 threads): build this CTA's 128 activation rows, one 32-element K block per iteration. =====
        // Mapping: lane = (row-in-group-of-4, chunk): lane l handles 16-byte chunk q = l & 7 of rows  w*16 + 4*i + (l >> 3), i = 0..3.
        // Eight lanes read one 128-byte row slice -> fully coalesced 16-byte copies, and a thread's four channels are the same for all of
        // its rows, so their rel-xyz weights are fetched once per iteration.  Global -> shared staging by cp.async, PF-1 blocks ahead.
        constexpr int NSL = PROD == TC_PROD_FC_H1 ? 6 : 4;          // 16-byte slots per thread per block: 4 rows (+ 2 centre-point rows)
        constexpr int PF = STG_BYTES / (NSL * 256 * 16);            // staging ring depth: 4 blocks (3 for the flow-embedding producer)
        static_assert(PF >= 3, "staging ring too shallow");
        const int p = threadIdx.x - 256;
        const int pw = p >> 5;                                     // producer warp 0..7 -> rows pw*16 .. pw*16+15
        const int q = lane & 7, rsub = lane >> 3;
        const int row0 = pw * 16 + rsub;                            // this thread's rows: row0 + 4*i
        const uint32_t stg0 = base + NSTAGE * STAGE_BYTES + p * 16; // slot (ring r, i) of this thread at + (r*NSL + i) * 4096
        const int pf = a.k_blocks + 1 < PF ? a.k_blocks + 1 : PF;   // look-ahead never reaches beyond the next tile
        int stage = 0; uint32_t phase = 0;
        // staging slots start as zeros: rows beyond the last column are never copied, and their (stale but finite) slot contents are
        // multiplied by a zero scale below instead of being selected away
        for (int r = 0; r < PF * NSL; ++r)
            asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(stg0 + r * 4096), "f"(0.f));
        // NOTE: the asm statements below carry no "memory" clobber on purpose -- they are volatile, so they keep their order among
        // themselves (copy -> commit -> wait_group -> ld.shared), while the compiler stays free to hoist the plain shared-memory loads
        // of contexts / weights above them and to interleave the four rows' arithmetic and stores.
        auto cp16 = [&](uint32_t dst, const float *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src)); };
        // gathered-row (and centre-row) pointers of the LOOK-AHEAD tile live in registers: they change once per tile, not per K block
        const float *ls1[4] = {nullptr, nullptr, nullptr, nullptr}, *ls0[2] = {nullptr, nullptr};
        auto load_ptrs = [&](int buf) {
#pragma unroll
            for (int i = 0; i < 4; ++i) ls1[i] = cs1[buf * HALF_N + row0 + 4 * i];
            if (PROD == TC_PROD_FC_H1) { ls0[0] = cs0[buf * HALF_N + row0]; ls0[1] = cs0[buf * HALF_N + row0 + 8]; }
        };
        auto issue = [&](int kb, int ring) {                       // copies of K block kb of the look-ahead tile
            const int koff = kb * PK + q * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (ls1[i]) {
                    cp16(stg0 + (ring * NSL + i) * 4096, ls1[i] + koff);
                    // the centre-point row is shared by the 8 neighbour rows of a point: rows i = 0,1 and i = 2,3 of this thread
                    if (PROD == TC_PROD_FC_H1 && (i & 1) == 0) cp16(stg0 + (ring * NSL + 4 + (i >> 1)) * 4096, ls0[i >> 1] + koff);
                }
            }
            asm volatile("cp.async.commit_group;");
        };
        auto lds16 = [&](uint32_t addr) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
            return v;
        };
        long long t = cl_id;
        if (t < ntiles) {
            int buf = 0, ring = 0;                                  // buf = context slot of the current tile, ring = staging slot being consumed
            asm volatile("bar.sync 1, 288;" ::: "memory");          // contexts of the first two tiles are in place (warp 2)
            // prologue: blocks 0 .. pf-2 (pf - 1 <= k_blocks: all inside the first tile)
            load_ptrs(0);
            for (int g = 0; g < pf - 1; ++g) issue(g, g);
            // look-ahead cursor: block (current + pf - 1)
            int la_buf = 0, la_kb = pf - 1; long long la_t = t;
            if (la_kb >= a.k_blocks) { la_kb -= a.k_blocks; la_buf = 1; la_t += n_cl; if (la_t < ntiles) load_ptrs(1); }
            while (true) {
                const long long tn = t + n_cl;
                // {s*dx, s*dy, s*dz, s} of this thread's four rows, fixed for the whole tile (s = power-of-two fp16 scale of the row's
                // frame pair, 1 in TF32 mode, 0 for rows beyond the last column).  s > 0 commutes exactly with the affine term, ReLU /
                // LeakyReLU and every rounding, so scaling the inputs gives bit-identical results to scaling the activated value.
                float4 geo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 g = cgeo[buf * HALF_N + row0 + 4 * i];
                    geo[i] = make_float4(g.x * g.w, g.y * g.w, g.z * g.w, g.w);
                }
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    {   // keep pf-1 blocks in flight
                        int lring = ring + pf - 1; if (lring >= pf) lring -= pf;
                        if (la_t < ntiles) issue(la_kb, lring); else asm volatile("cp.async.commit_group;");
                        if (++la_kb == a.k_blocks) {                // the cursor moves on to the next tile: fetch its row pointers
                            la_kb = 0; la_buf = la_buf == 2 ? 0 : la_buf + 1; la_t += n_cl;
                            if (la_t < ntiles) load_ptrs(la_buf);
                        }
                    }
                    // this thread's four channels of the 32-block and their rel-xyz weights
                    float4 wx = make_float4(0.f, 0.f, 0.f, 0.f), wy = wx, wz = wx;
                    if (PROD == TC_PROD_FC_H1 || PROD == TC_PROD_SC2_Y1) {
                        const float4 *wp = sW + kb * 24 + q;
                        wx = wp[0]; wy = wp[8]; wz = wp[16];
                    }
                    if (pf == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
                    else if (pf == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
                    else asm volatile("cp.async.wait_group 1;" ::: "memory");
                    float4 v[4], uc[2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = lds16(stg0 + (ring * NSL + i) * 4096);
                    if (PROD == TC_PROD_FC_H1) { uc[0] = lds16(stg0 + (ring * NSL + 4) * 4096); uc[1] = lds16(stg0 + (ring * NSL + 5) * 4096); }
                    // arithmetic first (results in registers), THEN wait for the stage to be free: after the tensor core releases a stage
                    // only the stores, the proxy fence and the arrive remain on the critical path of the ring
                    uint4 hh[4], ll[4];                             // fmt 0: four hi / lo floats; fmt 1: .x,.y = four hi halfs, ll .x,.y = four lo halfs
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // packed fp32 (FFMA2 / FMUL2 / FADD2): two channels per instruction
                        const float4 g = geo[i];
                        const float2 gx = make_float2(g.x, g.x), gy = make_float2(g.y, g.y), gz = make_float2(g.z, g.z), gs = make_float2(g.w, g.w);
                        float2 xa = make_float2(v[i].x, v[i].y), xb = make_float2(v[i].z, v[i].w);
                        if (PROD == TC_PROD_FC_H1) {
                            xa = __fadd2_rn(make_float2(uc[i >> 1].x, uc[i >> 1].y), xa);
                            xb = __fadd2_rn(make_float2(uc[i >> 1].z, uc[i >> 1].w), xb);
                        }
                        if (PROD == TC_PROD_FC_H1 || PROD == TC_PROD_SC2_Y1) {
                            float2 ta = __fmul2_rn(make_float2(wx.x, wx.y), gx), tb = __fmul2_rn(make_float2(wx.z, wx.w), gx);
                            ta = __ffma2_rn(make_float2(wy.x, wy.y), gy, ta); tb = __ffma2_rn(make_float2(wy.z, wy.w), gy, tb);
                            ta = __ffma2_rn(make_float2(wz.x, wz.y), gz, ta); tb = __ffma2_rn(make_float2(wz.z, wz.w), gz, tb);
                            xa = __ffma2_rn(xa, gs, ta); xb = __ffma2_rn(xb, gs, tb);
                            if (PROD == TC_PROD_FC_H1) {            // LeakyReLU(0.1) = max(v, 0.1 v)
                                const float2 la = __fmul2_rn(xa, make_float2(0.1f, 0.1f)), lb = __fmul2_rn(xb, make_float2(0.1f, 0.1f));
                                xa = make_float2(fmaxf(xa.x, la.x), fmaxf(xa.y, la.y)); xb = make_float2(fmaxf(xb.x, lb.x), fmaxf(xb.y, lb.y));
                            } else {
                                xa = make_float2(fmaxf(xa.x, 0.f), fmaxf(xa.y, 0.f)); xb = make_float2(fmaxf(xb.x, 0.f), fmaxf(xb.y, 0.f));
                            }
                        } else {
                            xa = __fmul2_rn(xa, gs); xb = __fmul2_rn(xb, gs);
                        }
                        if (F16) {
                            split_f16x2(xa, hh[i].x, ll[i].x); split_f16x2(xb, hh[i].y, ll[i].y);
                        } else {
                            float h[4], l[4];
                            split_tf32(xa.x, h[0], l[0]); split_tf32(xa.y, h[1], l[1]); split_tf32(xb.x, h[2], l[2]); split_tf32(xb.y, h[3], l[3]);
                            hh[i] = make_uint4(__float_as_uint(h[0]), __float_as_uint(h[1]), __float_as_uint(h[2]), __float_as_uint(h[3]));
                            ll[i] = make_uint4(__float_as_uint(l[0]), __float_as_uint(l[1]), __float_as_uint(l[2]), __float_as_uint(l[3]));
                        }
                    }
                    const int st0 = stage, st1 = stage + SPB - 1;   // fmt 0: NSTAGE is even, a 32-block never wraps between its two stages
                    TIMED(dw0, mbar_wait(empty_bar(st0), phase ^ 1));
                    if (!F16) TIMED(dw0, mbar_wait(empty_bar(st1), phase ^ 1));
                    uint8_t *Bhi = smem + ((F16 || q < 4) ? st0 : st1) * STAGE_BYTES + 2 * TILE_A_FLOATS * 4;
                    uint8_t *Blo = Bhi + TILE_BH_FLOATS * 4;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = row0 + 4 * i;
                        if (F16) {                    // this thread's four halfs: bytes [8*(q&1), +8) of 16-byte chunk q>>1 of the 64-byte stage row
                            const int off = sw_off_h(row, q * 4);
                            *reinterpret_cast<uint2 *>(Bhi + off) = make_uint2(hh[i].x, hh[i].y);
                            *reinterpret_cast<uint2 *>(Blo + off) = make_uint2(ll[i].x, ll[i].y);
                        } else {
                            const int off = sw_off(row, (q & 3) * 4) * 4;
                            *reinterpret_cast<uint4 *>(Bhi + off) = hh[i];
                            *reinterpret_cast<uint4 *>(Blo + off) = ll[i];
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(full_bar(st0)); if (!F16) mbar_arrive(full_bar(st1)); }
                    stage += SPB;
                    if (stage == NSTAGE) { stage = 0; phase ^= 1; }
                    if (++ring == pf) ring = 0;
                }
                if (tn >= ntiles) break;
                asm volatile("bar.sync 1, 288;" ::: "memory");     // producers are done with this tile's contexts; warp 2 has published tile t + 2
                t = tn; buf = buf == 2 ? 0 : buf + 1;
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (a.dbg && warp == 8 && lane == 0) a.dbg[(size_t)blockIdx.x * 8 + 5] = dw0;
    }
    else if (warp >= 8 && PROD == TC_PROD_TILED && a.epi == TC_EPI_WSUM) {
        // ===== WeightNet hidden layers for the WSUM epilogue: thread p owns column ct*256 + p of the tile (all 256 columns, both CTAs) =====
        const int p = threadIdx.x - 256;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long t = cl_id; t < ntiles; t += n_cl) {
            const long long c = (t / m_pairs) * BN + p;
            float h2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (c < a.cols) {
                const long long bi = c >> 3;
                const int b = (int)div_i(bi, a.n_pts), i = (int)(bi - (long long)b * a.n_pts);
                const int j = __ldg(a.nbr + (size_t)bi * a.nbr_ld + a.nbr_off + (int)(c & 7));
                const int nc = a.n_cand ? a.n_cand : a.n_pts;
                const float *pq = a.xyz_q + (size_t)b * 3 * a.n_pts, *pc = a.xyz_c + (size_t)b * 3 * nc;
                const float dx = __fsub_rn(__ldg(pc + j), __ldg(pq + i)), dy = __fsub_rn(__ldg(pc + nc + j), __ldg(pq + a.n_pts + i)),
                            dz = __fsub_rn(__ldg(pc + 2 * nc + j), __ldg(pq + 2 * a.n_pts + i));
                float h1[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 w = __ldg(reinterpret_cast<const float4 *>(a.wnA1) + u);
                    h1[u] = fmaxf(fmaf(w.z, dz, fmaf(w.y, dy, fmaf(w.x, dx, __ldg(a.wna1 + u)))), 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float sacc = __ldg(a.wna2 + u);
#pragma unroll
                    for (int v = 0; v < 8; ++v) sacc = fmaf(__ldg(a.wnA2 + u * 8 + v), h1[v], sacc);
                    h2[u] = fmaxf(sacc, 0.f);
                }
            }
            mbar_wait(h2empty_bar(acc), acc_phase ^ 1);
            float *dst = h2s + ((((size_t)acc * BN + p) >> 1) * 16) + (p & 1);            // [column pair][j][2]
#pragma unroll
            for (int u = 0; u < 8; ++u) dst[u * 2] = h2[u];
            __syncwarp();
            if (lane == 0) mbar_arrive(h2full_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    if (a.dbg && threadIdx.x == 0) a.dbg[(size_t)blockIdx.x * 8 + 0] = clock64() - t_start;
    tc_fence_before();
    cluster_sync_all();                         // nobody frees TMEM / exits while the pair still uses it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

template <int PROD>
int launch2(const TcArgs &a, int n_clusters, cudaStream_t st) {
    if (a.fmt == 1) tc_gemm2_kernel<PROD, 1><<<2 * n_clusters, NTHREADS, SMEM_BYTES, st>>>(a);      // cluster shape comes from __cluster_dims__(2,1,1)
    else tc_gemm2_kernel<PROD, 0><<<2 * n_clusters, NTHREADS, SMEM_BYTES, st>>>(a);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}
template <int PROD>
cudaError_t set_smem2() {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm2_kernel<PROD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc_gemm2_kernel<PROD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
}

}  // namespace

int cmf_launch_tc_gemm2(const TcArgs &a, cudaStream_t st) {
    // function attributes are per device: a process that drives several GPUs (the reference's nn.DataParallel mode) sets them on each
    static int num_sms_of[64];
    static bool attr_set_of[64];
    int dev = 0;
    CMF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { cmf_set_error("tc gemm: device ordinal %d out of range", dev); return CMF_ERR_STATE; }
    if (!attr_set_of[dev]) {
        CMF_CUDA(set_smem2<TC_PROD_PLAIN>());
        CMF_CUDA(set_smem2<TC_PROD_FC_H1>());
        CMF_CUDA(set_smem2<TC_PROD_SC2_Y1>());
        CMF_CUDA(set_smem2<TC_PROD_TILED>());
        CMF_CUDA(cudaDeviceGetAttribute(&num_sms_of[dev], cudaDevAttrMultiProcessorCount, dev));
        attr_set_of[dev] = true;
    }
    const int num_sms = num_sms_of[dev];
    if (a.cols <= 0 || a.m_blocks <= 0) return CMF_OK;
    if ((a.m_blocks & 1) || (a.M & 255)) { cmf_set_error("tc_gemm2: needs M %% 256 == 0"); return CMF_ERR_INVALID; }
    if (a.out_tiled && a.epi != TC_EPI_STORE) { cmf_set_error("tc_gemm2: tiled output needs the STORE epilogue"); return CMF_ERR_INVALID; }
    if (a.epi == TC_EPI_MAXK && a.ksamp != 4 && a.ksamp != 8 && a.ksamp != 16 && a.ksamp != 32) { cmf_set_error("tc_gemm2: MAXK needs ksamp in {4,8,16,32}"); return CMF_ERR_INVALID; }
    if (a.prod == TC_PROD_FC_H1 && a.ksamp != 8) { cmf_set_error("tc_gemm2: the flow-embedding producer assumes 8 neighbours per point"); return CMF_ERR_INVALID; }
    if (a.epi == TC_EPI_WSUM && (a.prod != TC_PROD_TILED || a.ksamp != 8)) { cmf_set_error("tc_gemm2: WSUM needs the TILED producer and 8 neighbours per point"); return CMF_ERR_INVALID; }
    if ((a.pbias || a.bs_mode || a.amax_out) && a.cols_per_pair <= 0) { cmf_set_error("tc_gemm2: cols_per_pair must be set"); return CMF_ERR_INVALID; }
    if (a.amax_out && a.amax_group <= 0) { cmf_set_error("tc_gemm2: amax_group must be positive"); return CMF_ERR_INVALID; }
    const long long ntiles = ((a.cols + BN - 1) / BN) * (a.m_blocks >> 1);
    const int max_cl = num_sms / 2;
    const int n_cl = (int)(ntiles < max_cl ? ntiles : max_cl);
    if (a.prod == TC_PROD_PLAIN) return launch2<TC_PROD_PLAIN>(a, n_cl, st);
    if (a.prod == TC_PROD_FC_H1) return launch2<TC_PROD_FC_H1>(a, n_cl, st);
    if (a.prod == TC_PROD_TILED) return launch2<TC_PROD_TILED>(a, n_cl, st);
    return launch2<TC_PROD_SC2_Y1>(a, n_cl, st);
}

// installs the host-mapped watchdog record of this translation unit's kernels (tc_dev.cuh) on the current device
int cmf_wd_set_tc_gemm2(unsigned long long *dev_ptr) {
    CMF_CUDA(cudaMemcpyToSymbol(tcdev::g_cmf_wd_record, &dev_ptr, sizeof(dev_ptr)));
    return CMF_OK;
}
