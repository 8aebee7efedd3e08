// tc_sc2.cu -- set-conv #2 (mse_layer2, radarflow_util.py:144-155) after the hoisted first layer: neighbour gather + layer 2 (512 -> 256)
// + layer 3 (256 -> 64) + max over the K neighbours in ONE kernel (3xFP16 split precision, fp32 accumulation in TMEM).
//
// Round 1 ran this stage as two kernels per scale: layer 2 wrote its 256-channel output pre-split to HBM (4 GB per step at 256 pairs)
// and a second, HBM-bound kernel read it back for layer 3 + max.  Here layer 2's output never leaves the SM pair:
//
//   * ORIENTATION.  The ACTIVATIONS are the A operand (M = 256 neighbour columns per CTA pair, 128 per CTA), the weights the B operand
//     (N = 256 output channels, 128 rows staged by each CTA: cta_group::2 splits B across the pair).  The shared-memory stage of a
//     CTA holds the same bytes as in tc_gemm2.cu -- 128 weight rows + 128 activation rows, hi and lo -- only the two descriptors swap
//     places in the MMA.  The accumulator D2[128 lanes x 256 columns] of a CTA then has one LANE per neighbour column and the 256
//     channels of that column along the TMEM columns.
//   * LAYER 3 READS ITS A OPERAND FROM TMEM.  An epilogue thread owns one lane = one neighbour column: it loads 16 channels at a time,
//     applies un-scale / bias / ReLU, splits into fp16 hi + lo and writes the 8 + 8 packed 32-bit words back over the 16 fp32 columns it
//     just read (tcgen05.st): D2 turns IN PLACE into the K-major fp16 A operand of layer 3, one K=16 step per 16 columns (hi at +0,
//     lo at +8).  tcgen05.mma with A in TMEM and W3 (64 x 256, resident in shared memory, 32 rows per CTA) as B accumulates
//     D3[128 x 64]; max over a point's K consecutive lanes by warp shuffles; one 256-byte row per point goes to HBM.
//   * TMEM BUDGET.  Two 256-column D2 buffers use all 512 columns, so D3 has no columns of its own: it lives in columns 0..63 of the SAME
//     buffer.  The epilogue first takes channels 0..63 into registers (already converted: 64 packed words), which frees those columns
//     for D3; channels 64..255 are converted in place and multiplied first; when the K steps of channels 64..127 have retired, their
//     columns take the held channels 0..63 for the last four K steps.
//   * Layer-3 MMAs are issued by the (single) MMA-issuer warp between its main-loop stages; all of its waits poll for pending layer-3 work
//     (the main loop's next-but-one tile needs the buffer back, so a blocking wait there would deadlock).
//
//   * THE NEIGHBOUR GATHER IS TMA.  The hoisted layer-1 matrix P (one 512-float row slice per point and scale) is described by a 2-D tensor
//     map (box = 64 floats x 1 row); cp.async.bulk.tensor.2d ... tile::gather4 fetches four arbitrary rows x 256 bytes per instruction -- two
//     K blocks of each row -- into a staging ring, completing on an mbarrier by byte count.  Each producer warp issues the four gathers of
//     its own 16 rows (one elected lane, one group of two K blocks ahead) and waits only on its own barrier, so the 256 producer threads
//     just wait, transform and store: no per-thread address arithmetic, no cp.async groups, nothing of their own in flight at the proxy
//     fence.  A gather instruction costs its warp ~150 clk whatever the box width: with one K block per gather the issue alone was 46 % of
//     the producers' busy time and the kernel ran at the producers' pace, not the tensor pipe's.
//
// Cluster of 2 CTAs, 512 threads each -- roles as in tc_gemm2.cu: w0 bulk-copy issuer (weights), w1 MMA issuer (leader) / w1 + w3
// forwarders (peer), w2 row-context filler, w4-7 epilogue, w8.. producers (PW = 8 warps; gather + rel-xyz term + ReLU + fp16 split).
#include <cuda.h>
#include <stdlib.h>

#define CMF_WD_TU 3
#include "tc_dev.cuh"

using namespace tcdev;

namespace {

constexpr int HALF_ROWS = 128;                       // activation rows (neighbour columns) per CTA
constexpr int TILE_BYTES = 128 * 64;                 // 128 rows x 32 halfs: 8 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;          // W hi, W lo, X hi, X lo: 32 KB
#ifndef SC2_NSTAGE
#define SC2_NSTAGE 3
#endif
#ifndef SC2_POLL_SLEEP
#define SC2_POLL_SLEEP 0
#endif
#ifndef SC2_PF
#define SC2_PF 2
#endif
#ifndef SC2_GK
#define SC2_GK 2
#endif
#ifndef SC2_PW
#define SC2_PW 8
#endif
constexpr int NSTAGE = SC2_NSTAGE;
// Producer warps.  The producers' loop is a chain of waits, shared-memory loads, ~11 packed instructions per channel pair, stores, a proxy
// fence and an arrive; with 8 producer warps (two per scheduler) those latencies set the kernel's pace, not the tensor pipe.  16 producer
// warps (768 threads) halve the rows per thread and double the warps the schedulers can switch between; the register file is redistributed
// by warpgroup with setmaxnreg: utility warps 48, epilogue 144, producers 72.  The pool is what the CTA was LAUNCHED with (768 x 80 registers
// = 61,440 = 128 x (48 + 144 + 4 x 72)): an increase beyond what the other groups release never completes.
constexpr int PW = SC2_PW;                           // producer warps: 8 or 16
constexpr int RPW = HALF_ROWS / PW;                  // rows per producer warp: 16 or 8
constexpr int RPT = RPW / 4;                         // rows per producer thread: 4 or 2
constexpr int NTHREADS = 256 + 32 * PW;
static_assert(PW == 8 || PW == 16, "8 or 16 producer warps");
// Gather granularity.  One tile::gather4 instruction costs the issuing warp ~150 clk whatever its box width (measured with the section
// timers below: four 32-float boxes per K block were 46 % of a producer warp's busy time), so a gather fetches GK K blocks of a row at
// once (box = 32 * GK floats) and the staging ring holds PF such groups: 4 instructions per warp per GK K blocks.
constexpr int GK = SC2_GK;                           // K blocks per gather
constexpr int PF = SC2_PF;                           // staging ring depth in gather groups: PF - 1 groups in flight per producer warp
constexpr int ROW_BYTES = 128 * GK;                  // staged bytes per row and group
constexpr int SLOT_BYTES = HALF_ROWS * ROW_BYTES;    // 32 KB: [128 rows][32 * GK floats], rows in tile order
constexpr int STG_BYTES = PF * SLOT_BYTES;           // 64 KB
constexpr int W3_KB = 8;                             // 256 channels = 8 K blocks of 32
constexpr int W3_BYTES = W3_KB * 2 * 2048;           // this CTA's 32 rows of W3: per K block {hi 2 KB, lo 2 KB}
constexpr int OFF_STG = NSTAGE * STAGE_BYTES;        // 98304
constexpr int OFF_W3 = OFF_STG + STG_BYTES;          // 163840
constexpr int OFF_BAR = OFF_W3 + W3_BYTES;           // 196608
constexpr int OFF_SW = OFF_BAR + 256;                // rel-xyz weights: 16 K blocks x 24 float4 = 6 KB
constexpr int OFF_CS1 = OFF_SW + 6144;               // gathered-row indices (rows of P) [3][128] int
constexpr int OFF_GEO = OFF_CS1 + 3 * HALF_ROWS * 4; // {dx, dy, dz, scale} [3][128]
constexpr int OFF_AB2 = OFF_GEO + 3 * HALF_ROWS * 16;// {a_inv, bias} of the 256 layer-2 channels (float2)
constexpr int OFF_AB3 = OFF_AB2 + 256 * 8;           // {a_inv, bias} of the 64 layer-3 channels
constexpr int OFF_SBAR = OFF_AB3 + 64 * 8;           // staging barriers [PF][PW producer warps]
constexpr int OFF_CBAR = OFF_SBAR + PF * PW * 8;     // row-context barriers: ready[3] (filled, count 1), free[3] (the PW producer warps are done with it)
constexpr int SMEM_BYTES = OFF_CBAR + 6 * 8 + 1024;  // + alignment slack
static_assert(SMEM_BYTES <= 232448, "shared-memory plan exceeds 227 KB");
constexpr uint32_t IDESC_L2 = make_idesc(256, 256, 1), IDESC_L3 = make_idesc(256, 64, 1);

struct Sc2Args {
    TcArgs g;                        // layer 2: tiled W2 (Wt, a_inv, bias), SC2_Y1 producer fields, scales (bs_*, out_mul / out_add), cols, cols_per_pair, ksamp
    const float *Wt3, *a_inv3, *bias3;   // layer 3: tiled W3 (one 128-row block, 8 K blocks), per-channel un-scale, bias
    float *out; int ldo;             // out[point][0..63]
    int expt;                        // timing experiments (results invalid): 1 = no layer 3 at all, 3 = conversions but no layer-3 MMAs
};

// ISSUE DISCIPLINE.  The whole issuer warp executes these in uniform control flow and the instruction is predicated on `el`, the flag of the one
// lane elected at role start (elect.sync).  Under an `if (lane == 0)` the compiler must assume per-thread operands: it wraps EVERY tcgen05.mma
// in an ELECT / 5 x R2UR.BROADCAST / branch loop, ~100 clk of issue per MMA -- which is what made "MMA issue block at the pipe's rate" in round 1
// and made a 32-clk layer-3 MMA cost 115.  With uniform operands the descriptors live in uniform registers and the MMAs issue back to back.
// CAUTION: ptxas turns `@el tcgen05.mma` into ONE unguarded UTCHMMA per pass of the warp through the code (operands broadcast from the elected
// lane), so the issuer warp must be CONVERGED wherever it issues: it polls its barriers with warp-uniform votes (mbar_test_warp), never per lane
// (lanes that drifted apart re-issued MMAs with stale uniform registers: launch failure).
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t el;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(el));
    return el;
}
__device__ __forceinline__ void commit2_mc(uint32_t el, uint32_t bar) {       // arrive on `bar` in BOTH CTAs when all prior MMAs of the elected thread retire
    const uint16_t mask = 3;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar), "h"(mask), "r"(el) : "memory");
}
__device__ __forceinline__ void mma2_ss(uint32_t el, uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(el) : "memory");
}
__device__ __forceinline__ void mma2_ss_keep(uint32_t el, uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, 1, 0;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(el) : "memory");
}
__device__ __forceinline__ void mma2_ss_reuse(uint32_t el, uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, 1, 0;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(el) : "memory");
}
// A operand in tensor memory (K-major, two fp16 per 32-bit column, one lane per row), B from shared memory
__device__ __forceinline__ void mma2_ts(uint32_t el, uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(el) : "memory");
}
// TMEM load without the wait: tmem_ld16_done() is the wait plus a register dependency, so that one load can be in flight under
// the arithmetic on the previous one
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_done(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// truly non-blocking test of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
// the same as a warp-uniform decision (true only when EVERY lane has observed -- and acquired -- the phase): the issuer warp polls with this, so
// that its lanes never drift apart and the compiler sees uniform control flow around the MMA issue
__device__ __forceinline__ bool mbar_test_warp(uint32_t bar, uint32_t parity) { return __all_sync(0xffffffffu, mbar_test(bar, parity)) != 0; }

// max over the K consecutive lanes of a point for 64 channels, as a halving butterfly: at the stage with lane offset `off` a lane keeps one
// half of its channels and trades the other half with its partner, so the stages cost 32 + 16 + ... shuffles instead of 64 each, and the K
// lanes of a point end up with 64 / K channels apiece -- together one contiguous 256-byte row.
template <int K>
__device__ __forceinline__ void maxk_store(float (&v)[64], int lane, bool valid, float *orow) {
    int ch0 = 0;
#pragma unroll
    for (int off = 1, n = 64; off < K; off <<= 1, n >>= 1) {
        const bool up = (lane & off) != 0;                   // this lane keeps the upper half of its n channels
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
        }
        ch0 += up ? n / 2 : 0;
    }
    constexpr int CNT = 64 / K;                              // 16, 8, 4 or 2 channels left in this lane
    if (valid) {
        if (CNT >= 4) {
#pragma unroll
            for (int i = 0; i < CNT; i += 4) *reinterpret_cast<float4 *>(orow + ch0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
            *reinterpret_cast<float2 *>(orow + ch0) = make_float2(v[0], v[1]);
        }
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
sc2_fused_kernel(const Sc2Args s, const __grid_constant__ CUtensorMap tmapP) {
    const TcArgs &a = s.g;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    const uint32_t bar0 = base + OFF_BAR;
    auto full_bar = [&](int i) { return bar0 + 8 * i; };                 // [NSTAGE <= 4] local: bulk copy (expect_tx) + PW producer warps
    auto pfull_bar = [&](int i) { return bar0 + 32 + 8 * i; };           // [NSTAGE <= 4] leader: the peer's half of the stage is complete
    auto empty_bar = [&](int i) { return bar0 + 64 + 8 * i; };           // [NSTAGE <= 4] both: stage consumed (MMA commit, multicast)
    auto tfull_bar = [&](int i) { return bar0 + 96 + 8 * i; };           // [2] both: layer-2 accumulator complete
    auto tempty_bar = [&](int i) { return bar0 + 112 + 8 * i; };         // [2] leader: 4 + 4 epilogue warps have drained the buffer
    auto a3r_bar = [&](int acc, int g) { return bar0 + 128 + 8 * (acc * 4 + g); };   // [2][4] leader: layer-3 operand group written (4 + 4 warps)
    auto g1done_bar = [&](int i) { return bar0 + 192 + 8 * i; };         // [2] both: the K steps of channels 64..127 have retired
    auto d3full_bar = [&](int i) { return bar0 + 208 + 8 * i; };         // [2] both: layer-3 accumulator complete
    const uint32_t w3_bar = bar0 + 224;                                  // local: resident W3 slice has landed
    auto ctx_ready_bar = [&](int i) { return base + OFF_CBAR + 8 * i; };
    auto ctx_free_bar = [&](int i) { return base + OFF_CBAR + 24 + 8 * i; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 240);
    float4 *sW = reinterpret_cast<float4 *>(smem + OFF_SW);
    int *cidx = reinterpret_cast<int *>(smem + OFF_CS1);
    float4 *cgeo = reinterpret_cast<float4 *>(smem + OFF_GEO);
    float2 *sAB2 = reinterpret_cast<float2 *>(smem + OFF_AB2);
    float2 *sAB3 = reinterpret_cast<float2 *>(smem + OFF_AB3);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;      // broadcast: the compiler then knows the role branches are warp-uniform
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    // optional wait-time instrumentation (a.dbg = long long[gridDim.x][16]): {total, issuer: tempty, full + peer | producer warp 8: gathered rows | epilogue warp 4: tfull,
    // g1done, d3full | producer warp 8: empty}
    const long long t_start = a.dbg ? clock64() : 0;
    long long dw0 = 0, dw1 = 0, dw2 = 0;
#define TIMED(acc_, stmt_) do { if (a.dbg) { const long long c0_ = clock64(); stmt_; acc_ += clock64() - c0_; } else { stmt_; } } while (0)
    const long long ntiles = (a.cols + 255) / 256;                       // M = 256 output channels: one 256-row tile covers them all
    const long long cl_id = blockIdx.x >> 1, n_cl = gridDim.x >> 1;
    const int nks = a.k_blocks;                                          // K blocks of 32 (512 / 32 = 16)

    // rel-xyz weights transposed for the producers: sW[(kb*3 + comp)*8 + q] = {Wx[c][comp], c = kb*32 + q*4 .. +3} (see tc_gemm2.cu)
    for (int i = threadIdx.x; i < nks * 24; i += NTHREADS) {
        const int kb_ = i / 24, comp = (i >> 3) % 3, q_ = i & 7;
        const float *w = a.Wsmall + (size_t)(kb_ * PK + q_ * 4) * 4 + comp;
        sW[i] = make_float4(__ldg(w), __ldg(w + 4), __ldg(w + 8), __ldg(w + 12));
    }
    for (int i = threadIdx.x; i < 256; i += NTHREADS) sAB2[i] = make_float2(__ldg(a.a_inv + i), a.bias ? __ldg(a.bias + i) : 0.f);
    if (threadIdx.x < 64) sAB3[threadIdx.x] = make_float2(__ldg(s.a_inv3 + threadIdx.x), s.bias3 ? __ldg(s.bias3 + threadIdx.x) : 0.f);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(full_bar(i), 1 + PW); mbar_init(pfull_bar(i), 1); mbar_init(empty_bar(i), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 8); mbar_init(g1done_bar(i), 1); mbar_init(d3full_bar(i), 1);
            for (int g = 0; g < 4; ++g) mbar_init(a3r_bar(i, g), 8);
        }
        mbar_init(w3_bar, 1);
        for (int i = 0; i < PF * PW; ++i) mbar_init(base + OFF_SBAR + 8 * i, 1);
        for (int i = 0; i < 3; ++i) { mbar_init(base + OFF_CBAR + 8 * i, 1); mbar_init(base + OFF_CBAR + 24 + 8 * i, PW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    cluster_sync_all();                         // both CTAs: barriers initialised, TMEM allocated, tables visible
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Roles by warpgroup (warps 0-3 utility, 4-7 epilogue, 8.. producers): with 16 producer warps each group first takes its share of the
    // register file (setmaxnreg must dominate the code it governs, and every warp of a group executes it with the same count).
    if (warp < 4) {
    if (PW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
        // ===== bulk-copy issuer: the resident W3 slice once, then this CTA's 128 rows of W2 per stage =====
        if (lane == 0) {
            mbar_arrive_expect_tx(w3_bar, W3_BYTES);
            for (int kb = 0; kb < W3_KB; ++kb)
                for (int hl = 0; hl < 2; ++hl)
                    bulk_g2s(base + OFF_W3 + (kb * 2 + hl) * 2048,
                             reinterpret_cast<const uint8_t *>(s.Wt3) + (size_t)(kb * 2 + hl) * TILE_BYTES + rank * 2048, 2048, w3_bar);
            int stage = 0; uint32_t phase = 0;
            for (long long t = cl_id; t < ntiles; t += n_cl)
                for (int ks = 0; ks < nks; ++ks) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint8_t *src = reinterpret_cast<const uint8_t *>(a.Wt) + ((size_t)rank * nks + ks) * (2 * TILE_BYTES);
                    mbar_arrive_expect_tx(full_bar(stage), 2 * TILE_BYTES);
                    bulk_g2s(base + stage * STAGE_BYTES, src, 2 * TILE_BYTES, full_bar(stage));
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
        }
    } else if (leader && warp == 1) {
        // ===== MMA issuer (leader CTA): one warp issues the main loop AND layer 3 =====
        // A single issuing thread fixes the order in which the MMAs enter the in-order tensor pipe, hence the accumulation order and the result
        // bits (two alternating issuer warps -- round 1 -- are reproducible only while the issue itself is slow: MMAs of different warps may
        // interleave differently from run to run).  Layer 3 goes out in whole operand groups of four K steps (twelve MMAs) between main-loop
        // stages and whenever the issuer would otherwise wait.  Measured alternative: a second thread issuing layer 3 interleaves its small
        // MMAs one-to-one with the big ones, every one of them then waits a full main MMA (~200 clk) and the 48 of a tile stretch the tile's
        // layer-3 chain beyond the next tile's main loop.
        const uint32_t el = elect_one();                   // the one lane of this warp that issues (and commits) every MMA
        const long long my_tiles = cl_id < ntiles ? (ntiles - cl_id + n_cl - 1) / n_cl : 0;
        const long long l3_tiles = (s.expt == 0 || s.expt == 4) ? my_tiles : 0;
        long long l3_next = 0;                              // next layer-3 group: local tile * 4 + group
        bool w3_ready = false;
        // groups: 0 = K steps 4..7 (channels 64..127, in place), 1 = 8..11, 2 = 12..15, 3 = K steps 0..3 (channels 0..63, parked in the columns
        // of group 0 once that has retired).  Non-blocking.
        auto serve_l3 = [&]() {
            while (true) {
                const long long l3_tile = l3_next >> 2; const int l3_grp = (int)(l3_next & 3);
                if (l3_tile >= l3_tiles) return;
                const int acc3 = (int)(l3_tile & 1);
                if (!mbar_test_warp(a3r_bar(acc3, l3_grp), (uint32_t)((l3_tile >> 1) & 1))) return;
                if (!w3_ready) { mbar_wait(w3_bar, 0); w3_ready = true; __syncwarp(); }      // (per-lane wait loop: converge before issuing)
                tc_fence_after();
                const uint32_t d3 = tmem_base + acc3 * 256;
                const int ks0 = l3_grp == 3 ? 0 : 4 + 4 * l3_grp;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ks = ks0 + j;
                    const uint32_t a_hi = tmem_base + acc3 * 256 + (l3_grp == 3 ? 64 : 0) + 16 * ks, a_lo = a_hi + 8;
                    const uint32_t wb = base + OFF_W3 + (ks >> 1) * 4096;
                    const uint64_t b_hi = make_desc(wb) + (uint64_t)((ks & 1) * 2), b_lo = make_desc(wb + 2048) + (uint64_t)((ks & 1) * 2);
                    if (s.expt != 4) {
                        mma2_ts(el, d3, a_lo, b_hi, IDESC_L3, (l3_grp == 0 && j == 0) ? 0u : 1u);
                        mma2_ts(el, d3, a_hi, b_lo, IDESC_L3, 1u);
                        mma2_ts(el, d3, a_hi, b_hi, IDESC_L3, 1u);
                    } else {
                        mma2_ts(el, d3, a_hi, b_hi, IDESC_L3, (l3_grp == 0 && j == 0) ? 0u : 1u);      // timing experiment: one MMA per K step
                    }
                }
                if (l3_grp == 0) commit2_mc(el, g1done_bar(acc3));
                if (l3_grp == 3) commit2_mc(el, d3full_bar(acc3));
                ++l3_next;                                  // (tile, 3) + 1 = (tile + 1, 0)
            }
        };
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (long long t = cl_id; t < ntiles; t += n_cl) {
            const uint32_t d2 = tmem_base + acc * 256;
            for (int ks = 0; ks < nks; ++ks) {
                {   // the stage's operands (both halves); pending layer-3 groups go out while I wait
                    unsigned spins = 0; unsigned long long t0 = 0ull;
                    const long long c0_ = a.dbg ? clock64() : 0;
                    bool have_full = false, have_peer = false;
                    long long cf_ = 0;
                    while (true) {
                        if (!have_full) { have_full = mbar_test_warp(full_bar(stage), phase); if (have_full && a.dbg) cf_ = clock64(); }
                        if (!have_peer) { have_peer = mbar_test_warp(pfull_bar(stage), phase); if (have_peer && a.dbg && !have_full) cf_ = -1; }
                        if (have_full && have_peer) { if (a.dbg && cf_ > 0) dw2 += clock64() - cf_; break; }      // dw2: waiting for the PEER's half after my own was complete
                        serve_l3();
                        if (SC2_POLL_SLEEP) __nanosleep(SC2_POLL_SLEEP);                 // an always-eligible polling warp takes issue slots from the producers of its scheduler
                        watchdog(spins, t0);
                    }
                    if (a.dbg) dw1 += clock64() - c0_;
                }
                if (ks == 0) {
                    // the accumulator buffer comes back only when ITS layer 3 is done, and layer 3 is issued here
                    unsigned spins = 0; unsigned long long t0 = 0ull;
                    const long long c0_ = a.dbg ? clock64() : 0;
                    while (!mbar_test_warp(tempty_bar(acc), acc_phase ^ 1)) { serve_l3(); watchdog(spins, t0); }
                    if (a.dbg) dw0 += clock64() - c0_;
                }
                __syncwarp();
                tc_fence_after();
                {
                    const uint32_t sa = base + stage * STAGE_BYTES;
                    const uint64_t w_hi = make_desc(sa), w_lo = make_desc(sa + TILE_BYTES);
                    const uint64_t x_hi = make_desc(sa + 2 * TILE_BYTES), x_lo = make_desc(sa + 3 * TILE_BYTES);
#pragma unroll
                    for (int k16 = 0; k16 < 2; ++k16) {
                        const uint64_t adv = (uint64_t)(k16 * 2);             // 32 bytes = 16 halfs, in 16-byte descriptor units
                        mma2_ss(el, d2, x_lo + adv, w_hi + adv, IDESC_L2, (ks | k16) ? 1u : 0u);
                        mma2_ss_keep(el, d2, x_hi + adv, w_lo + adv, IDESC_L2);
                        mma2_ss_reuse(el, d2, x_hi + adv, w_hi + adv, IDESC_L2);
                    }
                    commit2_mc(el, empty_bar(stage));
                    if (ks == nks - 1) commit2_mc(el, tfull_bar(acc));
                }
                serve_l3();                                                          // whole groups between main-loop stages
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        {                                                   // drain: layer 3 of the last tile(s)
            unsigned spins = 0; unsigned long long t0 = 0ull;
            while ((l3_next >> 2) < l3_tiles) { serve_l3(); watchdog(spins, t0); }
            if (a.dbg && lane == 0) { a.dbg[(size_t)blockIdx.x * 16 + 1] = dw0; a.dbg[(size_t)blockIdx.x * 16 + 2] = dw1; a.dbg[(size_t)blockIdx.x * 16 + 13] = dw2; }
        }
    } else if (warp == 2) {
        // ===== row contexts: neighbour index -> row of P, rel-xyz, fp16 scale, up to three tiles ahead of the producers =====
        auto fill_ctx = [&](long long tt, int buf) {
            RowCtx rc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) rc[k] = make_row(a, tt * 256 + rank * HALF_ROWS + lane + 32 * k);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = buf * HALF_ROWS + lane + 32 * k;
                // row of P this column gathers; columns beyond the last one read row 0 (any valid row) and are zeroed by their scale
                cidx[r] = rc[k].valid ? (int)((rc[k].src1 - (a.U2 + a.off_u2)) / a.ld_u2) : 0;
                cgeo[r] = make_float4(rc[k].dx, rc[k].dy, rc[k].dz, rc[k].valid ? rc[k].scale : 0.f);
            }
        };
        const long long my_tiles = cl_id < ntiles ? (ntiles - cl_id + n_cl - 1) / n_cl : 0;
        long long fi = 0;                                                   // next local tile whose contexts are to be filled
        while (fi < my_tiles) {                                             // up to three tiles ahead of the producers
            const int slot = (int)(fi % 3);
            if (fi >= 3) mbar_wait(ctx_free_bar(slot), (uint32_t)(((fi / 3) - 1) & 1));
            fill_ctx(cl_id + fi * n_cl, slot);
            __syncwarp();
            if (lane == 0) mbar_arrive(ctx_ready_bar(slot));
            ++fi;
        }
    } else if (!leader && (warp == 1 || warp == 3)) {
        // ===== forwarders (peer CTA): relay "my half of stage s is complete" to the leader =====
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
    }
    } else if (warp < 8) {
        if (PW == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        // ===== epilogue: this thread owns lane q*32 + lane of the CTA's accumulator = one neighbour column =====
        const int q = warp & 3;
        const int K = a.ksamp;
        int acc = 0; uint32_t acc_phase = 0;
        auto arrive_leader = [&](uint32_t bar) { if (lane == 0) { if (leader) mbar_arrive(bar); else mbar_arrive_remote(bar, 0); } };
        for (long long t = cl_id; t < ntiles; t += n_cl) {
            const long long c = t * 256 + rank * HALF_ROWS + q * 32 + lane;
            const bool valid = c < a.cols;
            const long long pair = div_i(valid ? c : a.cols - 1, a.cols_per_pair);
            const float osc = out_scale_of(a, pair);                               // power of two: layer-3 operand scale of this pair
            const float k2 = __frcp_rn(b_scale_of(a, pair)) * osc;                  // accumulator un-scale (times a_inv[ch]) with the output scale folded in
            const float2 k2v = make_float2(k2, k2), oscv = make_float2(osc, osc);
            const uint32_t tcol = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
            // 16 accumulator columns (channels ch0 .. ch0+15) -> relu(acc * a_inv * k2 + bias * osc) -> 8 packed hi + 8 packed lo words
            auto convert16 = [&](const uint32_t (&r)[16], int ch0, uint32_t *hi, uint32_t *lo) {
                const float4 *ab = reinterpret_cast<const float4 *>(sAB2 + ch0);      // {a_inv[c], bias[c], a_inv[c+1], bias[c+1]}: broadcast reads
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 w = ab[i];
                    float2 v = __ffma2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])),
                                          __fmul2_rn(make_float2(w.x, w.z), k2v), __fmul2_rn(make_float2(w.y, w.w), oscv));
                    v = make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
                    split_f16x2(v, hi[i], lo[i]);
                }
            };
            TIMED(dw0, mbar_wait_cluster(tfull_bar(acc), acc_phase));
            tc_fence_after();
            if (s.expt == 1) {
                tc_fence_before(); __syncwarp(); arrive_leader(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            // channels 0..63 leave their columns to D3 at once: raw into registers first, converted while layer 3 already runs on the others
            uint32_t H0[16], H1[16], H2[16], H3[16];
            tmem_ld16_issue(tcol, H0); tmem_ld16_issue(tcol + 16, H1); tmem_ld16_issue(tcol + 32, H2); tmem_ld16_issue(tcol + 48, H3);
            uint32_t ra[16], rb[16];
            tmem_ld16_issue(tcol + 64, ra);
            tmem_ld16_done(H0); tmem_ld16_done(H1); tmem_ld16_done(H2); tmem_ld16_done(H3); tmem_ld16_done(ra);
            // channels 64..255 in place, four K steps per group; the load of K step ks+1 is in flight under the arithmetic of K step ks
#pragma unroll 1
            for (int grp = 0; grp < 3; ++grp) {
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    const int ks = 4 + 4 * grp + j;
                    uint32_t hi[8], lo[8];
                    tmem_ld16_issue(tcol + 16 * (ks + 1), rb);
                    convert16(ra, 16 * ks, hi, lo);
                    tmem_st8(tcol + 16 * ks, hi);                    // in place: the 16 fp32 columns become K step ks of layer 3's A operand
                    tmem_st8(tcol + 16 * ks + 8, lo);
                    tmem_ld16_done(rb);
                    if (ks + 2 < 16) tmem_ld16_issue(tcol + 16 * (ks + 2), ra);
                    convert16(rb, 16 * (ks + 1), hi, lo);
                    tmem_st8(tcol + 16 * (ks + 1), hi);
                    tmem_st8(tcol + 16 * (ks + 1) + 8, lo);
                    if (ks + 2 < 16) tmem_ld16_done(ra);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                arrive_leader(a3r_bar(acc, grp));
            }
            if (s.expt == 3) {
                tc_fence_before(); __syncwarp(); arrive_leader(tempty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            uint32_t Hh[32], Hl[32];
            convert16(H0, 0, Hh, Hl); convert16(H1, 16, Hh + 8, Hl + 8); convert16(H2, 32, Hh + 16, Hl + 16); convert16(H3, 48, Hh + 24, Hl + 24);
            TIMED(dw1, mbar_wait_cluster(g1done_bar(acc), acc_phase));   // the K steps that read columns 64..127 have retired
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                tmem_st8(tcol + 64 + 16 * ks, Hh + 8 * ks);
                tmem_st8(tcol + 64 + 16 * ks + 8, Hl + 8 * ks);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            arrive_leader(a3r_bar(acc, 3));
            TIMED(dw2, mbar_wait_cluster(d3full_bar(acc), acc_phase));
            tc_fence_after();
            // layer-3 epilogue: un-scale, bias, ReLU, max over the point's K consecutive lanes (halving butterfly), one 256-byte row per point
            {
                const float inv3 = __frcp_rn(osc);
                float v[64];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t r[32];
                    tmem_ld32(tcol + 32 * half, r);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float2 ab = sAB3[32 * half + i];
                        v[32 * half + i] = fmaxf(fmaf(__uint_as_float(r[i]), ab.x * inv3, ab.y), 0.f);
                    }
                }
                float *orow = s.out + (size_t)div_i(c, K) * s.ldo;
                if (K == 4) maxk_store<4>(v, lane, valid, orow);
                else if (K == 8) maxk_store<8>(v, lane, valid, orow);
                else if (K == 16) maxk_store<16>(v, lane, valid, orow);
                else maxk_store<32>(v, lane, valid, orow);
            }
            tc_fence_before();
            __syncwarp();
            arrive_leader(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (a.dbg && warp == 4 && lane == 0) { a.dbg[(size_t)blockIdx.x * 16 + 4] = dw0; a.dbg[(size_t)blockIdx.x * 16 + 5] = dw1; a.dbg[(size_t)blockIdx.x * 16 + 6] = dw2; }
    } else {
        if (PW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        // ===== producers (PW warps): this CTA's 128 activation rows, one 32-channel K block per iteration (as tc_gemm2.cu, SC2_Y1) =====
        const int pw = warp - 8;                                    // warp-uniform (warp is a broadcast value)
        const uint32_t el = elect_one();                            // the lane that issues this warp's gathers
        const int q = lane & 7, rsub = lane >> 3;
        const int row0 = pw * RPW + rsub;                           // this thread's rows: row0 + 4*i, i < RPT
        const uint32_t stg_w = base + OFF_STG + pw * RPW * ROW_BYTES; // this warp's RPW rows of ring slot r at + r * SLOT_BYTES
        const uint32_t stg_t = stg_w + rsub * ROW_BYTES + q * 16;   // this thread's 16-byte chunk of row 4*i + rsub at + i * 4 * ROW_BYTES, K block h of the group at + h * 128
        const uint32_t sbar_w = base + OFF_SBAR + pw * 8;           // this warp's barrier of ring slot r at + r * PW * 8
        int stage = 0; uint32_t phase = 0;
#ifdef SC2_PROD_PROFILE
        long long ps0 = 0, ps1 = 0, ps2 = 0, ps3 = 0, ps4 = 0;      // producer sections: gather issue | loads + arithmetic | stores | proxy fence | arrive
#endif
        // four rows x (128 * GK) bytes per instruction, rows picked by index: the warp's 16 rows of gather group kg (K blocks kg * GK ...) of the
        // tile whose contexts sit in `lbuf`.  (Measured alternative: holding the 16 indices in registers for the whole tile instead of re-reading
        // them from shared memory per gather made the kernel 5 % SLOWER -- 16 more live registers in a 128-register kernel.)
        // Issued like the MMAs (see ISSUE DISCIPLINE): the whole warp runs this in uniform control flow with warp-uniform operands and the
        // instructions are predicated on the elected lane.  Under `if (lane == 0)` every UTMALDG sat in an ELECT / 7 x R2UR.BROADCAST /
        // branch loop: ~180 clk per gather, 46 % of a producer warp's busy time (section timers, SC2_PROD_PROFILE).
        auto issue = [&](int kg, int ring, int lbuf) {
            const int4 *ip = reinterpret_cast<const int4 *>(cidx + lbuf * HALF_ROWS + pw * RPW);     // warp-uniform address: broadcast loads
            const uint32_t dst = stg_w + ring * SLOT_BYTES, bar = sbar_w + ring * (PW * 8);
            const int col = a.off_u2 + kg * (PK * GK);
            __syncwarp();
            asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
                         ::"r"(bar), "r"(RPW * ROW_BYTES), "r"(el) : "memory");
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int4 id = ip[i];
                asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %8, 0;\n\t"
                             "@q cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n\t}"
                             ::"r"(dst + i * 4 * ROW_BYTES), "l"(&tmapP), "r"(col), "r"(id.x), "r"(id.y), "r"(id.z), "r"(id.w), "r"(bar), "r"(el) : "memory");
            }
        };
        auto lds16 = [&](uint32_t addr) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
            return v;
        };
        long long t = cl_id;
        if (t < ntiles) {
            int buf = 0, ring = 0; uint32_t rphase = 0;             // ring slot being consumed and the parity of its barrier
            long long ti = 0, la_i = 0;                             // local tile numbers of the current tile and of the look-ahead cursor
            mbar_wait(ctx_ready_bar(0), 0);                         // contexts of the first tile are in place (warp 2)
            const int n_groups = a.k_blocks / GK;                   // 8 gather groups per tile
            for (int g = 0; g < PF - 1; ++g) issue(g, g, 0);        // n_groups >= PF - 1: all inside the first tile
            int la_buf = 0, la_kg = PF - 1; long long la_t = t;     // look-ahead cursor: gather group (current + PF - 1)
            while (true) {
                const long long tn = t + n_cl;
                float4 geo[RPT];
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const float4 g = cgeo[buf * HALF_ROWS + row0 + 4 * i];
                    geo[i] = make_float4(g.x * g.w, g.y * g.w, g.z * g.w, g.w);
                }
                for (int kb = 0; kb < a.k_blocks; ++kb) {
#ifdef SC2_PROD_PROFILE
                    const long long pc0 = clock64();
#endif
                    const int hk = kb % GK;                           // K block within its gather group
                    if (hk == 0) {
                        // keep PF - 1 groups in flight: refill the slot consumed during the previous group (its reads are done: their
                        // results were stored below, and the proxy fence there orders them before this asynchronous write)
                        int lring = ring + PF - 1; if (lring >= PF) lring -= PF;
                        if (la_t < ntiles) issue(la_kg, lring, la_buf);
                        if (++la_kg == n_groups) {                  // the cursor moves on to the next tile: its contexts must have been filled
                            la_kg = 0; la_buf = la_buf == 2 ? 0 : la_buf + 1; la_t += n_cl; ++la_i;
                            if (la_t < ntiles) mbar_wait(ctx_ready_bar(la_buf), (uint32_t)((la_i / 3) & 1));
                        }
                    }
                    const float4 *wp = sW + kb * 24 + q;
                    const float4 wx = wp[0], wy = wp[8], wz = wp[16];
#ifdef SC2_PROD_PROFILE
                    const long long pc1 = clock64(); ps0 += pc1 - pc0;
#endif
                    if (hk == 0) { TIMED(dw1, mbar_wait(sbar_w + ring * (PW * 8), rphase)); }      // the whole group lands on one barrier
#ifdef SC2_PROD_PROFILE
                    const long long pc2 = clock64();
#endif
                    float4 v[RPT];
#pragma unroll
                    for (int i = 0; i < RPT; ++i) v[i] = lds16(stg_t + ring * SLOT_BYTES + i * 4 * ROW_BYTES + hk * 128);
                    uint2 hh[RPT], ll[RPT];
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        const float4 g = geo[i];
                        const float2 gx = make_float2(g.x, g.x), gy = make_float2(g.y, g.y), gz = make_float2(g.z, g.z), gs = make_float2(g.w, g.w);
                        float2 xa = make_float2(v[i].x, v[i].y), xb = make_float2(v[i].z, v[i].w);
                        float2 ta = __fmul2_rn(make_float2(wx.x, wx.y), gx), tb = __fmul2_rn(make_float2(wx.z, wx.w), gx);
                        ta = __ffma2_rn(make_float2(wy.x, wy.y), gy, ta); tb = __ffma2_rn(make_float2(wy.z, wy.w), gy, tb);
                        ta = __ffma2_rn(make_float2(wz.x, wz.y), gz, ta); tb = __ffma2_rn(make_float2(wz.z, wz.w), gz, tb);
                        xa = __ffma2_rn(xa, gs, ta); xb = __ffma2_rn(xb, gs, tb);
                        xa = make_float2(fmaxf(xa.x, 0.f), fmaxf(xa.y, 0.f)); xb = make_float2(fmaxf(xb.x, 0.f), fmaxf(xb.y, 0.f));
                        split_f16x2(xa, hh[i].x, ll[i].x); split_f16x2(xb, hh[i].y, ll[i].y);
                    }
#ifdef SC2_PROD_PROFILE
                    const long long pc3 = clock64(); ps1 += pc3 - pc2;
#endif
                    TIMED(dw0, mbar_wait(empty_bar(stage), phase ^ 1));
#ifdef SC2_PROD_PROFILE
                    const long long pc4 = clock64();
#endif
                    uint8_t *Xhi = smem + stage * STAGE_BYTES + 2 * TILE_BYTES;
                    uint8_t *Xlo = Xhi + TILE_BYTES;
#pragma unroll
                    for (int i = 0; i < RPT; ++i) {
                        const int off = sw_off_h(row0 + 4 * i, q * 4);
                        *reinterpret_cast<uint2 *>(Xhi + off) = hh[i];
                        *reinterpret_cast<uint2 *>(Xlo + off) = ll[i];
                    }
#ifdef SC2_PROD_PROFILE
                    const long long pc5 = clock64(); ps2 += pc5 - pc4;
#endif
                    fence_async_smem();
#ifdef SC2_PROD_PROFILE
                    const long long pc6 = clock64(); ps3 += pc6 - pc5;
#endif
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full_bar(stage));
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    if (hk == GK - 1 && ++ring == PF) { ring = 0; rphase ^= 1; }
#ifdef SC2_PROD_PROFILE
                    ps4 += clock64() - pc6;
#endif
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(ctx_free_bar(buf));      // this warp is done with the tile's contexts
                if (tn >= ntiles) break;
                t = tn; buf = buf == 2 ? 0 : buf + 1; ++ti;
                mbar_wait(ctx_ready_bar(buf), (uint32_t)((ti / 3) & 1));
            }
        }
        if (a.dbg && warp == 8 && lane == 0) { a.dbg[(size_t)blockIdx.x * 16 + 7] = dw0; a.dbg[(size_t)blockIdx.x * 16 + 3] = dw1; }
#ifdef SC2_PROD_PROFILE
        if (a.dbg && warp == 8 && lane == 0) {
            long long *d = a.dbg + (size_t)blockIdx.x * 16 + 8;
            d[0] = ps0; d[1] = ps1; d[2] = ps2; d[3] = ps3; d[4] = ps4;
        }
#endif
    }
    if (a.dbg && threadIdx.x == 0) a.dbg[(size_t)blockIdx.x * 16 + 0] = clock64() - t_start;
#undef TIMED
    tc_fence_before();
    cluster_sync_all();                         // nobody frees TMEM / exits while the pair still uses it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

}  // namespace

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links no libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int cmf_make_row_map(void *tm_, const float *base, long long rows, int ld, int box_cols, int box_rows) {
    CUtensorMap *tm = reinterpret_cast<CUtensorMap *>(tm_);
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
            cmf_set_error("tensor map: cuTensorMapEncodeTiled is not available from this driver"); return CMF_ERR_STATE;
        }
        enc = (EncodeTiledFn)fn;
    }
    // rows x ld floats, row pitch ld * 4 bytes; box = box_cols floats x box_rows rows (tile::gather4 fetches four one-row boxes at four row indices);
    // rows beyond the matrix read as zeros
    const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows}, gstride[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    const CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { cmf_set_error("tensor map: cuTensorMapEncodeTiled failed (%d)", (int)rc); return CMF_ERR_CUDA; }
    return CMF_OK;
}

int cmf_launch_sc2_fused(const TcArgs &l2, const float *Wt3, const float *a_inv3, const float *bias3, float *out, int ldo, cudaStream_t st) {
    static int num_sms_of[64];
    static bool attr_set_of[64];
    int dev = 0;
    CMF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { cmf_set_error("sc2 fused: device ordinal %d out of range", dev); return CMF_ERR_STATE; }
    if (!attr_set_of[dev]) {
        CMF_CUDA(cudaFuncSetAttribute(sc2_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CMF_CUDA(cudaDeviceGetAttribute(&num_sms_of[dev], cudaDevAttrMultiProcessorCount, dev));
        attr_set_of[dev] = true;
    }
    if (l2.cols <= 0) return CMF_OK;
    if (l2.fmt != 1 || l2.prod != TC_PROD_SC2_Y1 || l2.M != 256 || l2.m_blocks != 2 || l2.k_blocks != 16 || l2.bs_mode != 1) {
        cmf_set_error("sc2 fused: needs the 3xFP16 format, the set-conv #2 producer, a 512 -> 256 layer and bound-based scales"); return CMF_ERR_INVALID;
    }
    if (l2.ksamp != 4 && l2.ksamp != 8 && l2.ksamp != 16 && l2.ksamp != 32) { cmf_set_error("sc2 fused: ksamp must be 4, 8, 16 or 32"); return CMF_ERR_INVALID; }
    if (l2.cols % l2.ksamp != 0 || l2.cols_per_pair <= 0 || (ldo & 3)) { cmf_set_error("sc2 fused: cols must be points x ksamp, ldo a multiple of 4"); return CMF_ERR_INVALID; }
    if ((l2.ld_u2 & 3) || (l2.off_u2 & 31) || (reinterpret_cast<uintptr_t>(l2.U2) & 15)) { cmf_set_error("sc2 fused: gathered matrix must be 16-byte aligned, ld a multiple of 4, offset a multiple of 32"); return CMF_ERR_INVALID; }
    CUtensorMap tmapP;
    {   // the gathered matrix: one row per point of the (query = candidate) cloud, l2.cols / ksamp of them
        int rc = cmf_make_row_map(&tmapP, l2.U2, l2.cols / l2.ksamp, l2.ld_u2, 32 * GK, 1);
        if (rc != CMF_OK) return rc;
    }
    Sc2Args s;
    s.g = l2; s.Wt3 = Wt3; s.a_inv3 = a_inv3; s.bias3 = bias3; s.out = out; s.ldo = ldo;
    { const char *e = getenv("CMF_SC2_EXPT"); s.expt = e ? atoi(e) : 0; }
    const long long ntiles = (l2.cols + 255) / 256;
    const int max_cl = num_sms_of[dev] / 2;
    const int n_cl = (int)(ntiles < max_cl ? ntiles : max_cl);
    sc2_fused_kernel<<<2 * n_cl, NTHREADS, SMEM_BYTES, st>>>(s, tmapP);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// installs the host-mapped watchdog record of this translation unit's kernels (tc_dev.cuh) on the current device
int cmf_wd_set_tc_sc2(unsigned long long *dev_ptr) {
    CMF_CUDA(cudaMemcpyToSymbol(tcdev::g_cmf_wd_record, &dev_ptr, sizeof(dev_ptr)));
    return CMF_OK;
}
