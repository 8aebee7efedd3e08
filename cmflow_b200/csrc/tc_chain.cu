// tc_chain.cu -- tcgen05 kernels for the NARROW per-row MLP chains of the encoders (3xFP16 split precision, fp32 accumulate in TMEM):
//
//   * set-conv #1 (mse_layer, C = 3; radarflow_util.py:144-157): gather -> 6->32 (fp32 FMA, in registers) -> 32->32 -> 32->64 -> max over K
//   * the per-point tail of both set-convs (mlp2: 64->64->64->64 per scale; radarflow_util.py:158-161)
//
// These layers are 32-64 channels wide: as the N x 256-column "weights are A" tiles of tc_gemm.cu they would waste 50-75 % of a
// 128-row MMA and leave every epilogue thread with 2-byte scattered stores.  Here the ACTIVATIONS are the A operand: a CTA tile is 128
// rows (neighbour columns / points), one thread per row; the weights (N = 32 / 64 rows) are the B operand.  The accumulator row of a
// thread then is ITS row's output channels, so the epilogue of layer l (bias, ReLU, fp16 hi/lo split) writes layer l+1's A row with
// 16-byte shared-memory stores and nothing leaves the SM between layers.  Because a thread sees its whole row, the power-of-two fp16
// scale is the row's own exact maximum -- no bounds, no absmax pre-pass.
// The last layer of set-conv #1 is flipped (weights as A, duplicated into both 64-row halves; activations as B) so that the max over a
// point's K neighbour columns runs along a thread's TMEM columns instead of across lanes.
//
// Small tiles, no K pipeline: a CTA runs gather / FMA / epilogue phases and MMA phases back to back; 4 (set-conv) or 2 (mlp2) CTAs per
// SM overlap one another's phases.
#define CMF_WD_TU 4
#include "tc_dev.cuh"

using namespace tcdev;

namespace {

constexpr int CH_THREADS = 128;
constexpr int TILE_BYTES = 128 * 64;                 // one 128-row x 32-half operand tile (64-byte rows, 64B swizzle)

// ISSUE DISCIPLINE (see tc_sc2.cu): warp 0 runs the issue code in uniform control flow with warp-uniform operands and every tcgen05
// instruction is predicated on `el`, the flag of its elected lane.  Under `if (tid == 0)` each MMA sat in an ELECT / R2UR.BROADCAST / branch
// loop (~100 clk per MMA, with the CTA's other 127 threads waiting for the commit).  The warp must be converged where it issues.
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t el;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(el));
    return el;
}
__device__ __forceinline__ void tc_commit1(uint32_t el, uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(el) : "memory");
}
__device__ __forceinline__ void tc_mma1_f16(uint32_t el, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(el) : "memory");
}
// D = A * B^T over nkb 32-half K blocks, as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi.  Tile kb of an operand sits kb*stride bytes on.
__device__ __forceinline__ void issue_split_mma(uint32_t el, uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_stride, uint32_t b_hi, uint32_t b_lo,
                                                uint32_t b_stride, int nkb, uint32_t idesc) {
#pragma unroll
    for (int kb = 0; kb < nkb; ++kb) {
        const uint64_t dah = make_desc(a_hi + kb * a_stride), dal = make_desc(a_lo + kb * a_stride);
        const uint64_t dbh = make_desc(b_hi + kb * b_stride), dbl = make_desc(b_lo + kb * b_stride);
#pragma unroll
        for (int k16 = 0; k16 < 2; ++k16) {
            const uint64_t adv = (uint64_t)(k16 * 2);            // 32 bytes = 16 halfs, in 16-byte descriptor units
            tc_mma1_f16(el, d_tmem, dal + adv, dbh + adv, idesc, (kb | k16) ? 1u : 0u);
            tc_mma1_f16(el, d_tmem, dah + adv, dbl + adv, idesc, 1u);
            tc_mma1_f16(el, d_tmem, dah + adv, dbh + adv, idesc, 1u);
        }
    }
}
__device__ __forceinline__ void tmem_alloc1(uint32_t slot_smem, int cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc1(uint32_t taddr, int cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
// byte offset of 16-byte chunk ch (8 halfs) of row r inside a tile
__device__ __forceinline__ uint32_t chunk_off(int r, int ch) { return (uint32_t)(r * 64 + ((ch ^ ((r >> 1) & 3)) << 4)); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// this thread's row: 8 scaled values (4 float2) -> one 16-byte chunk of the hi tile and of the lo tile
__device__ __forceinline__ void store_chunk(uint32_t hi_tile, uint32_t lo_tile, int row, int ch, const float2 (&x)[4]) {
    uint4 h, l;
    split_f16x2(x[0], h.x, l.x); split_f16x2(x[1], h.y, l.y); split_f16x2(x[2], h.z, l.z); split_f16x2(x[3], h.w, l.w);
    const uint32_t off = chunk_off(row, ch);
    sts128(hi_tile + off, h);
    sts128(lo_tile + off, l);
}
// copy `bytes` (multiple of 16) global -> shared with all CH_THREADS threads
__device__ __forceinline__ void copy_g2s(uint8_t *dst, const void *src, int bytes) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (int i = threadIdx.x; i < bytes / 16; i += CH_THREADS) d[i] = __ldg(s + i);
}

// =====================================================================================================================================
// set-conv #1
// =====================================================================================================================================
struct Sc1Scale { const float *W1, *b1, *b2, *b3, *W2t, *ainv2, *W3t, *ainv3; };
struct Sc1Args {
    Sc1Scale w[4];
    int bc, n;
    const float *xyz, *ft; const int *idx60;
    float *out;                 // (bc*n, 256): [scale 0: 64 | scale 1: 64 | ...]
};

// shared-memory plan (bytes from the 1024-aligned base)
constexpr int SC1_A2 = 0;                       // layer-2 A operand: hi tile, lo tile
constexpr int SC1_B3 = 2 * TILE_BYTES;          // layer-3 B operand (activation rows): hi, lo
constexpr int SC1_W3 = 4 * TILE_BYTES;          // layer-3 A operand: W3 rows 0..63 twice (rows 64..127 = rows 0..63): hi, lo
constexpr int SC1_W2 = 6 * TILE_BYTES;          // layer-2 B operand: 32 rows: hi 2 KB, lo 2 KB (each 1024-aligned)
constexpr int SC1_F = SC1_W2 + 4096;            // floats: W1t[6][32] b1[32] b2[32] ainv2[32] b3[64] ainv3[64] sinv[128]
constexpr int SC1_F_FLOATS = 192 + 32 + 32 + 32 + 64 + 64 + 128;
constexpr int SC1_BAR = SC1_F + SC1_F_FLOATS * 4;
constexpr int SC1_SMEM = SC1_BAR + 16 + 1024;
constexpr uint32_t IDESC_128x32 = make_idesc(128, 32, 1), IDESC_128x128 = make_idesc(128, 128, 1), IDESC_128x64 = make_idesc(128, 64, 1);

// orow = the output row of the tile's first point (+ scale and channel offset); ncols = valid columns of the tile: 32-bit addressing inside the tile
template <int K>
__device__ __forceinline__ void sc1_maxk(const uint32_t (&r)[32], int col0, int ncols, float ainv, float bias, const float *sinv, float *orow) {
#pragma unroll
    for (int g0 = 0; g0 < 32; g0 += K) {
        float mx = __uint_as_float(r[g0]);
#pragma unroll
        for (int e = 1; e < K; ++e) mx = fmaxf(mx, __uint_as_float(r[g0 + e]));
        // the K columns of a point share one scale (phase 3), so the max of the scaled accumulators is the scaled max
        if (col0 + g0 < ncols) orow[((col0 + g0) / K) * 256] = fmaxf(fmaf(mx, ainv * sinv[col0 + g0], bias), 0.f);
    }
}

__global__ void __launch_bounds__(CH_THREADS, 4)
setconv1_tc_kernel(const Sc1Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    float *sf = reinterpret_cast<float *>(smem + SC1_F);
    float *sW1t = sf, *sb1 = sf + 192, *sb2 = sb1 + 32, *sainv2 = sb2 + 32, *sb3 = sainv2 + 32, *sainv3 = sb3 + 64, *sinv = sainv3 + 64;
    const uint32_t bar = base + SC1_BAR;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SC1_BAR + 8);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;      // warp: broadcast, so that `warp == 0` is provably uniform
    const uint32_t el = elect_one();

    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc1(smem_u32(tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t parity = 0;

    // flattened tile list over the four scales, split into contiguous per-CTA ranges
    const long long pts = (long long)a.bc * a.n;
    long long tstart[5];
    tstart[0] = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) tstart[s + 1] = tstart[s] + (pts * (4 << s) + 127) / 128;
    const long long t0 = tstart[4] * blockIdx.x / gridDim.x, t1 = tstart[4] * (blockIdx.x + 1) / gridDim.x;

    // per-thread gather cursor of a tile: neighbour index (stage A), then the six inputs (stage B)
    auto scale_of = [&](long long t) { return t >= tstart[3] ? 3 : (t >= tstart[2] ? 2 : (t >= tstart[1] ? 1 : 0)); };
    auto load_j = [&](long long t, int &j, long long &gp) {
        j = -1; gp = 0;
        if (t >= t1) return;
        const int s = scale_of(t), K = 4 << s, koff = K - 4;                        // KOFF = {0, 4, 12, 28} = K - 4
        const long long c = (t - tstart[s]) * 128 + tid;
        if (c >= pts * K) return;
        gp = c >> (2 + s);
        j = __ldg(a.idx60 + (size_t)gp * 60 + koff + (int)(c & (K - 1)));
    };
    auto load_x = [&](int j, long long gp, float (&x)[6]) {
        if (j < 0) { x[0] = x[1] = x[2] = x[3] = x[4] = x[5] = 0.f; return; }
        const unsigned b = (unsigned)gp / (unsigned)a.n; const int i = (int)((unsigned)gp - b * (unsigned)a.n);     // 32-bit: bc * n < 2^31 (launcher)
        const float *px = a.xyz + (size_t)b * 3 * a.n, *pf = a.ft + (size_t)b * 3 * a.n;
        x[0] = __fsub_rn(__ldg(px + j), __ldg(px + i));
        x[1] = __fsub_rn(__ldg(px + a.n + j), __ldg(px + a.n + i));
        x[2] = __fsub_rn(__ldg(px + 2 * a.n + j), __ldg(px + 2 * a.n + i));
        x[3] = __ldg(pf + j); x[4] = __ldg(pf + a.n + j); x[5] = __ldg(pf + 2 * a.n + j);
    };
    int jA, jB; long long gpA, gpB;
    float xB[6];
    load_j(t0, jB, gpB);
    load_x(jB, gpB, xB);
    load_j(t0 + 1, jA, gpA);

    int cur_s = -1;
    for (long long t = t0; t < t1; ++t) {
        const int s = scale_of(t);
        if (s != cur_s) {               // (re)load this scale's weights
            __syncthreads();            // nobody still reads the previous scale's tables
            const Sc1Scale &w = a.w[s];
            for (int i = tid; i < 192; i += CH_THREADS) { const int c = i >> 5, o = i & 31; sW1t[i] = __ldg(w.W1 + o * 8 + c); }
            if (tid < 32) { sb1[tid] = __ldg(w.b1 + tid); sb2[tid] = __ldg(w.b2 + tid); sainv2[tid] = __ldg(w.ainv2 + tid); }
            if (tid < 64) { sb3[tid] = __ldg(w.b3 + tid); sainv3[tid] = __ldg(w.ainv3 + tid); }
            const uint8_t *w2 = reinterpret_cast<const uint8_t *>(w.W2t), *w3 = reinterpret_cast<const uint8_t *>(w.W3t);
            copy_g2s(smem + SC1_W2, w2, 2048); copy_g2s(smem + SC1_W2 + 2048, w2 + TILE_BYTES, 2048);
            copy_g2s(smem + SC1_W3, w3, 4096); copy_g2s(smem + SC1_W3 + 4096, w3, 4096);
            copy_g2s(smem + SC1_W3 + TILE_BYTES, w3 + TILE_BYTES, 4096); copy_g2s(smem + SC1_W3 + TILE_BYTES + 4096, w3 + TILE_BYTES, 4096);
            cur_s = s;
            __syncthreads();
        }
        const int K = 4 << s;
        const long long tile_col0 = (t - tstart[s]) * 128, total_cols = pts * K;

        // rotate the gather pipeline: this tile's inputs are in xB; start the next tile's loads now
        float x0[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) x0[c] = xB[c];
        const bool valid = jB >= 0;
        jB = jA; gpB = gpA;
        load_x(jB, gpB, xB);
        load_j(t + 2, jA, gpA);

        // ---- phase 1: layer 1 (6 -> 32) in fp32, ReLU, row maximum, scale, split -> A2 row ----
        float2 h[16];
        float s1;
        {
            float mx = 0.f;
#pragma unroll
            for (int o = 0; o < 32; o += 4) {
                const float4 bb = *reinterpret_cast<const float4 *>(sb1 + o);
                float2 p0 = make_float2(bb.x, bb.y), p1 = make_float2(bb.z, bb.w);
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const float4 wv = *reinterpret_cast<const float4 *>(sW1t + c * 32 + o);
                    const float2 xc = make_float2(x0[c], x0[c]);
                    p0 = __ffma2_rn(make_float2(wv.x, wv.y), xc, p0); p1 = __ffma2_rn(make_float2(wv.z, wv.w), xc, p1);
                }
                p0 = make_float2(fmaxf(p0.x, 0.f), fmaxf(p0.y, 0.f)); p1 = make_float2(fmaxf(p1.x, 0.f), fmaxf(p1.y, 0.f));
                mx = fmaxf(mx, fmaxf(fmaxf(p0.x, p0.y), fmaxf(p1.x, p1.y)));
                h[o >> 1] = p0; h[(o >> 1) + 1] = p1;
            }
            s1 = valid ? pow2_scale(mx) : 0.f;
            const float2 s1v = make_float2(s1, s1);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float2 x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = __fmul2_rn(h[ch * 4 + e], s1v);
                store_chunk(base + SC1_A2, base + SC1_A2 + TILE_BYTES, tid, ch, x);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- phase 2: D2[128 x 32] = A2 * W2^T ----
        if (warp == 0) {
            tc_fence_after();
            issue_split_mma(el, tmem_base, base + SC1_A2, base + SC1_A2 + TILE_BYTES, 0, base + SC1_W2, base + SC1_W2 + 2048, 0, 1, IDESC_128x32);
            tc_commit1(el, bar);
        }
        mbar_wait(bar, parity); parity ^= 1;
        tc_fence_after();
        // ---- phase 3: layer-2 epilogue -> B3 row (layer 3 runs flipped: this row is a COLUMN of its output) ----
        {
            uint32_t r[32];
            tmem_ld32(tmem_row, r);
            const float inv1 = s1 > 0.f ? __frcp_rn(s1) : 0.f;
            const float2 inv1v = make_float2(inv1, inv1);
            float mx = 0.f;
#pragma unroll
            for (int o = 0; o < 32; o += 4) {
                const float4 ai = *reinterpret_cast<const float4 *>(sainv2 + o), bb = *reinterpret_cast<const float4 *>(sb2 + o);
                float2 p0 = __ffma2_rn(make_float2(__uint_as_float(r[o]), __uint_as_float(r[o + 1])), __fmul2_rn(make_float2(ai.x, ai.y), inv1v), make_float2(bb.x, bb.y));
                float2 p1 = __ffma2_rn(make_float2(__uint_as_float(r[o + 2]), __uint_as_float(r[o + 3])), __fmul2_rn(make_float2(ai.z, ai.w), inv1v), make_float2(bb.z, bb.w));
                p0 = make_float2(fmaxf(p0.x, 0.f), fmaxf(p0.y, 0.f)); p1 = make_float2(fmaxf(p1.x, 0.f), fmaxf(p1.y, 0.f));
                mx = fmaxf(mx, fmaxf(fmaxf(p0.x, p0.y), fmaxf(p1.x, p1.y)));
                h[o >> 1] = p0; h[(o >> 1) + 1] = p1;
            }
            // one scale per POINT: the maximum over the point's K rows (K consecutive lanes), so that the max over K of the scaled layer-3
            // accumulators is the max of the true values
            for (int off = 1; off < K; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float s2 = valid ? pow2_scale(mx) : 0.f;
            const float2 s2v = make_float2(s2, s2);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                float2 x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = __fmul2_rn(h[ch * 4 + e], s2v);
                store_chunk(base + SC1_B3, base + SC1_B3 + TILE_BYTES, tid, ch, x);
            }
            sinv[tid] = s2 > 0.f ? __frcp_rn(s2) : 0.f;
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---- phase 4: D3[128 (2 x 64 channels) x 128 columns] = W3dup * B3^T ----
        if (warp == 0) {
            tc_fence_after();
            issue_split_mma(el, tmem_base, base + SC1_W3, base + SC1_W3 + TILE_BYTES, 0, base + SC1_B3, base + SC1_B3 + TILE_BYTES, 0, 1, IDESC_128x128);
            tc_commit1(el, bar);
        }
        mbar_wait(bar, parity); parity ^= 1;
        tc_fence_after();
        // ---- phase 5: thread = channel (tid & 63); lanes 0..63 take columns 0..63, lanes 64..127 (duplicate rows) columns 64..127 ----
        {
            const int ch = tid & 63, half = tid >> 6;
            const float ainv = sainv3[ch], bias = sb3[ch];
            const int ncols = (int)(total_cols - tile_col0 < 128 ? total_cols - tile_col0 : 128);      // tile_col0 is a multiple of 128, hence of K
            float *orow = a.out + (size_t)(tile_col0 >> (2 + s)) * 256 + s * 64 + ch;                  // K = 4 << s
#pragma unroll 1
            for (int cc = 0; cc < 64; cc += 32) {
                uint32_t r[32];
                const int col0 = half * 64 + cc;
                tmem_ld32(tmem_row + col0, r);
                if (s == 0) sc1_maxk<4>(r, col0, ncols, ainv, bias, sinv, orow);
                else if (s == 1) sc1_maxk<8>(r, col0, ncols, ainv, bias, sinv, orow);
                else if (s == 2) sc1_maxk<16>(r, col0, ncols, ainv, bias, sinv, orow);
                else sc1_maxk<32>(r, col0, ncols, ainv, bias, sinv, orow);
            }
        }
        tc_fence_before();          // the next tile's MMA (after its phase-1 barrier) overwrites these TMEM columns
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc1(tmem_base, 128); }
}

// =====================================================================================================================================
// mlp2: Out[row][s*64 + o] = relu(V3 relu(V2 relu(V1 x + c1) + c2) + c3), x = In[row][s*64 .. +63], per scale s
// =====================================================================================================================================
struct Mlp2Args {
    const float *Vt[4][3], *ainv[4][3], *c[4][3];
    const float *in; int ld_in;
    float *out; int ld_out;
    long long rows;
    unsigned int *amax_out; int rows_per_pair;        // optional: atomicMax of the (non-negative) outputs per frame pair, uint bit patterns
    unsigned int *gmax;                               // optional: per-pair, per-channel maximum over the pair's rows, (pairs, 256) uint bit patterns (caller zeroes)
};
constexpr int ML_A = 0;                              // A operand: [kb 0: hi, lo][kb 1: hi, lo] = 4 tiles
// B operands: a ring of TWO layers' weights, [slot][kb]{hi 4 KB, lo 4 KB} (64 rows each).  Holding all three layers (48 KB) made the CTA 83 KB: two
// CTAs = 8 warps per SM for a kernel that is a chain of dependent phases per tile.  With the ring the next layer's 16 KB arrive by bulk copy
// (L2 -> shared, mbarrier byte count) while the current layer runs, the CTA is 67 KB and three fit.
constexpr int ML_W = 4 * TILE_BYTES;
constexpr int ML_F = ML_W + 2 * 2 * 8192;            // floats: ainv[3][64], c[3][64]
constexpr int ML_BAR = ML_F + 6 * 64 * 4;            // mma barrier, weight barriers [2], TMEM slot
constexpr int ML_SMEM = ML_BAR + 32 + 1024;

// 64 values of this thread's row (as 32 float2, all >= 0 or raw input) -> scale by the row's own power of two, split, store both K blocks
__device__ __forceinline__ float mlp2_store_row(uint32_t a_base, int row, const float2 (&h)[32], float mx, bool valid) {
    const float sc = valid ? pow2_scale(mx) : 0.f;
    const float2 scv = make_float2(sc, sc);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float2 x[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __fmul2_rn(h[kb * 16 + ch * 4 + e], scv);
            store_chunk(a_base + kb * 2 * TILE_BYTES, a_base + kb * 2 * TILE_BYTES + TILE_BYTES, row, ch, x);
        }
    return sc;
}

__global__ void __launch_bounds__(CH_THREADS, 3)
mlp2_tc_kernel(const Mlp2Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw);
    float *sainv = reinterpret_cast<float *>(smem + ML_F), *sc = sainv + 192;
    const uint32_t bar = base + ML_BAR, wbar = base + ML_BAR + 8;               // wbar + 8 * slot
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + ML_BAR + 24);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t el = elect_one();

    if (tid == 0) { mbar_init(bar, 1); mbar_init(wbar, 1); mbar_init(wbar + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc1(smem_u32(tmem_slot), 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t parity = 0;

    const long long ntile = (a.rows + 127) / 128, nitem = ntile * 4;           // item = scale * ntile + tile
    const long long i0 = nitem * blockIdx.x / gridDim.x, i1 = nitem * (blockIdx.x + 1) / gridDim.x;
    // weight ring: element j = (item, layer) in processing order uses slot j & 1; its 16 KB are requested one element ahead
    auto load_w = [&](long long item, int l, unsigned j) {                    // thread 0 only
        const uint8_t *wt = reinterpret_cast<const uint8_t *>(a.Vt[(int)(item / ntile)][l]);
        const uint32_t dst = base + ML_W + (j & 1u) * 16384, wb = wbar + 8 * (j & 1u);
        mbar_arrive_expect_tx(wb, 16384);
        for (int kb = 0; kb < 2; ++kb) {
            bulk_g2s(dst + kb * 8192, wt + (size_t)kb * 2 * TILE_BYTES, 4096, wb);
            bulk_g2s(dst + kb * 8192 + 4096, wt + (size_t)kb * 2 * TILE_BYTES + TILE_BYTES, 4096, wb);
        }
    };
    unsigned wj = 0;
    if (tid == 0 && i0 < i1) load_w(i0, 0, 0);
    int cur_s = -1;
    for (long long it = i0; it < i1; ++it) {
        const int s = (int)(it / ntile);
        const long long row = (it - (long long)s * ntile) * 128 + tid;
        const bool valid = row < a.rows;
        // this row's 64 inputs (issued before the weight reload so that the loads overlap it)
        float2 h[32];
        float mx = 0.f;
        {
            const float4 *src = reinterpret_cast<const float4 *>(a.in + (size_t)(valid ? row : 0) * a.ld_in + s * 64);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float4 v = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                h[2 * q] = make_float2(v.x, v.y); h[2 * q + 1] = make_float2(v.z, v.w);
                mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            }
        }
        if (s != cur_s) {
            __syncthreads();            // nobody still reads the previous scale's tables
            for (int l = 0; l < 3; ++l)
                if (tid < 64) { sainv[l * 64 + tid] = __ldg(a.ainv[s][l] + tid); sc[l * 64 + tid] = __ldg(a.c[s][l] + tid); }
            cur_s = s;
            // visibility: the barrier before the first MMA below
        }
        float rs = mlp2_store_row(base + ML_A, tid, h, mx, valid);
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (warp == 0) {
                // request the NEXT element's weights into the other slot: its last reader (the element before this one) has retired
                if (tid == 0) {
                    if (l < 2) load_w(it, l + 1, wj + 1);
                    else if (it + 1 < i1) load_w(it + 1, 0, wj + 1);
                }
                __syncwarp();
                mbar_wait(wbar + 8 * (wj & 1u), (wj >> 1) & 1u);                  // this element's weights have landed
                __syncwarp();                                                     // converged warp for the elect-predicated issue
                tc_fence_after();
                const uint32_t wb = base + ML_W + (wj & 1u) * 16384;
                issue_split_mma(el, tmem_base, base + ML_A, base + ML_A + TILE_BYTES, 2 * TILE_BYTES, wb, wb + 4096, 8192, 2, IDESC_128x64);
                tc_commit1(el, bar);
            }
            ++wj;
            mbar_wait(bar, parity); parity ^= 1;
            tc_fence_after();
            const float inv = rs > 0.f ? __frcp_rn(rs) : 0.f;
            const float2 invv = make_float2(inv, inv);
            mx = 0.f;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t r[32];
                tmem_ld32(tmem_row + half * 32, r);
#pragma unroll
                for (int o = 0; o < 32; o += 4) {
                    const float4 ai = *reinterpret_cast<const float4 *>(sainv + l * 64 + half * 32 + o), bb = *reinterpret_cast<const float4 *>(sc + l * 64 + half * 32 + o);
                    float2 p0 = __ffma2_rn(make_float2(__uint_as_float(r[o]), __uint_as_float(r[o + 1])), __fmul2_rn(make_float2(ai.x, ai.y), invv), make_float2(bb.x, bb.y));
                    float2 p1 = __ffma2_rn(make_float2(__uint_as_float(r[o + 2]), __uint_as_float(r[o + 3])), __fmul2_rn(make_float2(ai.z, ai.w), invv), make_float2(bb.z, bb.w));
                    p0 = make_float2(fmaxf(p0.x, 0.f), fmaxf(p0.y, 0.f)); p1 = make_float2(fmaxf(p1.x, 0.f), fmaxf(p1.y, 0.f));
                    mx = fmaxf(mx, fmaxf(fmaxf(p0.x, p0.y), fmaxf(p1.x, p1.y)));
                    h[half * 16 + (o >> 1)] = p0; h[half * 16 + (o >> 1) + 1] = p1;
                }
            }
            if (l < 2) {
                rs = mlp2_store_row(base + ML_A, tid, h, mx, valid);       // all MMAs of layer l have retired: A is free
            } else {
                if (valid) {
                    float4 *dst = reinterpret_cast<float4 *>(a.out + (size_t)row * a.ld_out + s * 64);
#pragma unroll
                    for (int q = 0; q < 16; ++q) dst[q] = make_float4(h[2 * q].x, h[2 * q].y, h[2 * q + 1].x, h[2 * q + 1].y);
                }
                if (a.gmax) {
                    // The global max-pooled feature (cmflow.py:76, 89: max over the pair's points) is taken here instead of by a pass over the
                    // output: the tile's 128 x 64 outputs go through shared memory (the A-operand tiles are free -- every MMA of the item has
                    // retired), thread (c = tid & 63, half = tid >> 6) takes the maximum of channel c over 64 rows, one atomicMax per thread.
                    // Outputs are ReLU values, so their uint bit patterns order like the floats and 0 is the identity.
                    const long long row0 = row - tid, last = (row0 + 127 < a.rows ? row0 + 127 : a.rows - 1);
                    const long long pfirst = div_i(row0, a.rows_per_pair);
                    if (pfirst == div_i(last, a.rows_per_pair)) {                       // (block-uniform) the whole tile belongs to one frame pair
                        float *sT = reinterpret_cast<float *>(smem + ML_A);       // [128 rows][64], 16-byte chunk q of row r at chunk q ^ (r & 15)
                        __syncthreads();
#pragma unroll
                        for (int q = 0; q < 16; ++q)
                            *reinterpret_cast<float4 *>(sT + tid * 64 + ((q ^ (tid & 15)) << 2)) =
                                valid ? make_float4(h[2 * q].x, h[2 * q].y, h[2 * q + 1].x, h[2 * q + 1].y) : make_float4(0.f, 0.f, 0.f, 0.f);
                        __syncthreads();
                        const int c = tid & 63, r0 = (tid >> 6) * 64;
                        float cm = 0.f;
#pragma unroll 16
                        for (int i = 0; i < 64; ++i) { const int r = r0 + i; cm = fmaxf(cm, sT[r * 64 + ((((c >> 2) ^ (r & 15)) << 2) | (c & 3))]); }
                        if (cm > 0.f) atomicMax(a.gmax + (size_t)pfirst * 256 + s * 64 + c, __float_as_uint(cm));
                        __syncthreads();                                          // the next item's A rows overwrite sT
                    } else if (valid) {                                           // a tile across a pair boundary (points per pair not a multiple of 128)
                        unsigned int *g = a.gmax + (size_t)div_i(row, a.rows_per_pair) * 256 + s * 64;
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            if (h[q].x > 0.f) atomicMax(g + 2 * q, __float_as_uint(h[q].x));
                            if (h[q].y > 0.f) atomicMax(g + 2 * q + 1, __float_as_uint(h[q].y));
                        }
                    }
                }
                if (a.amax_out) {          // the consumer GEMM's per-pair fp16 scale comes from here instead of a separate pass over the output
                    const long long pair = valid ? div_i(row, a.rows_per_pair) : -1;
                    const long long p0 = __shfl_sync(0xffffffffu, pair, 0);
                    const float v = valid ? mx : 0.f;
                    if (__all_sync(0xffffffffu, pair == p0 || !valid) && p0 >= 0) {
                        float wm = v;
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, off));
                        if ((tid & 31) == 0 && wm > 0.f) atomicMax(a.amax_out + p0, __float_as_uint(wm));
                    } else if (valid && v > 0.f) {
                        atomicMax(a.amax_out + pair, __float_as_uint(v));
                    }
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc1(tmem_base, 64); }
}

// function attributes are per device: set them on every device this process uses; returns the SM count of the current device in `sms`
int init_device(int &sms) {
    static int num_sms_of[64];
    static bool done_of[64];
    int dev = 0;
    CMF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { cmf_set_error("tc chain: device ordinal %d out of range", dev); return CMF_ERR_STATE; }
    if (!done_of[dev]) {
        CMF_CUDA(cudaFuncSetAttribute(setconv1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SC1_SMEM));
        CMF_CUDA(cudaFuncSetAttribute(mlp2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ML_SMEM));
        CMF_CUDA(cudaDeviceGetAttribute(&num_sms_of[dev], cudaDevAttrMultiProcessorCount, dev));
        done_of[dev] = true;
    }
    sms = num_sms_of[dev];
    return CMF_OK;
}

}  // namespace

int cmf_launch_setconv1_tc(int bc, int n, const float *xyz_planar, const float *ft_planar, const int *idx60, const TcChainSc1W *w4, float *out,
                           cudaStream_t st) {
    int g_num_sms = 0;
    int rc = init_device(g_num_sms);
    if (rc) return rc;
    if (bc <= 0 || n <= 0) return CMF_OK;
    if ((long long)bc * n >= (1LL << 31)) { cmf_set_error("setconv1_tc: more than 2^31 points in one chunk"); return CMF_ERR_INVALID; }
    Sc1Args a;
    for (int s = 0; s < 4; ++s)
        a.w[s] = Sc1Scale{w4[s].W1, w4[s].b1, w4[s].b2, w4[s].b3, w4[s].W2t, w4[s].ainv2, w4[s].W3t, w4[s].ainv3};
    a.bc = bc; a.n = n; a.xyz = xyz_planar; a.ft = ft_planar; a.idx60 = idx60; a.out = out;
    const long long tiles = ((long long)bc * n * 60 + 127) / 128 + 4;
    const int grid = (int)(tiles < 4LL * g_num_sms ? tiles : 4LL * g_num_sms);
    setconv1_tc_kernel<<<grid, CH_THREADS, SC1_SMEM, st>>>(a);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

int cmf_launch_mlp2_tc(long long rows, const float *in, int ld_in, float *out, int ld_out, const TcChainMlp2W *w4, unsigned int *amax_out,
                       int rows_per_pair, cudaStream_t st, float *gmax) {
    int g_num_sms = 0;
    int rc = init_device(g_num_sms);
    if (rc) return rc;
    if (rows <= 0) return CMF_OK;
    if ((ld_in & 3) || (ld_out & 3)) { cmf_set_error("mlp2_tc: leading dimensions must be multiples of 4"); return CMF_ERR_INVALID; }
    Mlp2Args a;
    for (int s = 0; s < 4; ++s)
        for (int l = 0; l < 3; ++l) { a.Vt[s][l] = w4[s].Vt[l]; a.ainv[s][l] = w4[s].ainv[l]; a.c[s][l] = w4[s].c[l]; }
    a.in = in; a.ld_in = ld_in; a.out = out; a.ld_out = ld_out; a.rows = rows;
    a.amax_out = amax_out; a.rows_per_pair = rows_per_pair > 0 ? rows_per_pair : 1;
    a.gmax = reinterpret_cast<unsigned int *>(gmax);
    const long long items = ((rows + 127) / 128) * 4;
    const int grid = (int)(items < 3LL * g_num_sms ? items : 3LL * g_num_sms);
    mlp2_tc_kernel<<<grid, CH_THREADS, ML_SMEM, st>>>(a);
    CMF_LAUNCH_CHECK();
    return CMF_OK;
}

// installs the host-mapped watchdog record of this translation unit's kernels (tc_dev.cuh) on the current device
int cmf_wd_set_tc_chain(unsigned long long *dev_ptr) {
    CMF_CUDA(cudaMemcpyToSymbol(tcdev::g_cmf_wd_record, &dev_ptr, sizeof(dev_ptr)));
    return CMF_OK;
}
