// tc_dev.cuh -- device helpers shared by the tcgen05 GEMM kernels (tc_gemm.cu: one CTA per tile; tc_gemm2.cu: CTA pairs).
//
// Two operand formats (TcArgs.fmt), both a 22-bit hi/lo split of the fp32 operands with fp32 accumulation in TMEM:
//   fmt 0  3xTF32: kind::tf32, stage row = 16 floats (64 bytes).
//   fmt 1  3xFP16: kind::f16,  stage row = 32 halfs  (64 bytes); operands pre-scaled by exact powers of two (see tc_gemm.cuh).
// The shared-memory geometry in BYTES is identical (64-byte rows, 64B swizzle, 8 KB per 128-row tile), so descriptors, bulk
// copies and the MMA issue loop are the same code; only the element packing and the instruction kind differ.
#pragma once
#include <cuda_fp16.h>

#include "tc_gemm.cuh"

namespace tcdev {

constexpr int BM = 128, BN = 256, SK = 16, PK = 32;
// Watchdog of the mbarrier waits: a wait that has not completed WATCHDOG_NS of wall-clock (%globaltimer) after it began traps, so that a
// protocol bug fails the launch instead of hanging the GPU.  NOTE: a trap poisons the CUDA context (sticky error for every later call of
// the process), so the budget is generous -- time-slicing, MPS, compute-sanitizer or a debugger can stretch a legitimate wait a lot.
// Compile with -DCMF_NO_WATCHDOG to remove it.
#ifndef CMF_WATCHDOG_MS
#define CMF_WATCHDOG_MS 30000
#endif
constexpr unsigned long long WATCHDOG_NS = (unsigned long long)CMF_WATCHDOG_MS * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// Where a watchdog fired: a trap poisons the context, so the thread first leaves {translation unit + 1, block << 32 | thread, grid << 32 | block size,
// waited ns} in a host-mapped record (one per translation unit that includes this header: CMF_WD_TU names it; cmf_wd_set_* installs the pointer,
// cmf_watchdog_read() -- callable after the failure -- returns it).  threadIdx.x >> 5 is the role warp, blockDim.x tells the kernel families apart.
#ifndef CMF_WD_TU
#define CMF_WD_TU 0
#endif
static __device__ unsigned long long *g_cmf_wd_record = nullptr;
__device__ __forceinline__ void watchdog(unsigned &spins, unsigned long long &t0) {
#ifndef CMF_NO_WATCHDOG
    if ((++spins & 63u) == 0u) {                 // the suspend-time hint bounds a try_wait from above only: look at the clock every 64 wake-ups
        const unsigned long long now = global_ns();
        if (t0 == 0ull) t0 = now;
        else if (now - t0 > WATCHDOG_NS) {
            unsigned long long *r = g_cmf_wd_record;
            if (r) {
                r[1] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x; r[2] = ((unsigned long long)gridDim.x << 32) | blockDim.x; r[3] = now - t0;
                r[0] = CMF_WD_TU + 1;
                __threadfence_system();
            }
            __trap();
        }
    }
#endif
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// x / d for a non-negative 64-bit x and a positive int d: column / row counters fit 32 bits in practice, and a 64-bit division is ~100
// instructions per call on the role warps' issue slots (the 32-bit one ~25)
__device__ __forceinline__ long long div_i(long long x, int d) {
    return (unsigned long long)x <= 0xffffffffull ? (long long)((unsigned)x / (unsigned)d) : x / d;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(cta) : "memory");
}
// relaxed form: only the arrival count travels.  For the forwarder -- the data it announces was written and proxy-fenced by OTHER threads
// (producers: st.shared + fence.proxy.async; bulk copies: complete_tx) and is read by the tensor core, not by the waiting thread.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned spins = 0; unsigned long long t0 = 0ull;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (!done) watchdog(spins, t0);                     // never hang the GPU: fail the launch instead
    }
}
// same, acquiring at cluster scope (the arrive came from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    unsigned spins = 0; unsigned long long t0 = 0ull;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (!done) watchdog(spins, t0);
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

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

// ---- operand splits ---------------------------------------------------------------------------------------------------------------
// hi = round-to-nearest TF32 (low 13 mantissa bits zero), lo = x - hi (exact)
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    lo = x - hi;
}
// two (already scaled) floats -> packed fp16 hi pair and fp16 lo pair: hi = rn_f16(x), lo = rn_f16(x - hi)
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
// packed form: the residual of both halves in one FFMA2 (x - hi = fma(hi, -1, x), exact)
__device__ __forceinline__ void split_f16x2(float2 x, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(x.x, x.y);
    const float2 d = __ffma2_rn(__half22float2(h), make_float2(-1.f, -1.f), x);
    const __half2 l = __floats2half2_rn(d.x, d.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ void split_f16(float x, unsigned short &hi, unsigned short &lo) {
    const __half h = __float2half_rn(x);
    const __half l = __float2half_rn(x - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}

// Power of two s with bound * s in [2^13, 2^14) (fp16 overflows at 65504: two binades of slack); 1 for zero / non-finite bounds.
__host__ __device__ __forceinline__ float pow2_scale(float bound) {
#ifdef __CUDA_ARCH__
    const unsigned u = __float_as_uint(bound);
#else
    union { float f; unsigned u; } cv; cv.f = bound; const unsigned u = cv.u;
#endif
    const int e = (int)((u >> 23) & 0xffu);                 // bound in [2^(e-127), 2^(e-126))
    if (e == 0 || e == 0xff) return 1.f;
    int se = 267 - e;                                       // biased exponent of 2^(14 - (e - 126))
    se = se < 1 ? 1 : (se > 254 ? 254 : se);
#ifdef __CUDA_ARCH__
    return __uint_as_float((unsigned)se << 23);
#else
    cv.u = (unsigned)se << 23; return cv.f;
#endif
}

// rigorous per-pair bound on |B operand element| (bs_mode 1), with a little slack for the fp32 rounding of the bound itself
__device__ __forceinline__ float pair_bound(const TcArgs &a, long long pair) {
    float bnd = a.bs_const;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (a.bs_src[i]) bnd = fmaf(a.bs_coef[i], __ldg(a.bs_src[i] + pair), bnd);
    return bnd * 1.001f;
}
__device__ __forceinline__ float b_scale_of(const TcArgs &a, long long pair) {
    if (a.bs_mode == 0) return 1.f;
    if (a.bs_mode == 2) return __ldg(a.bs_src[0] + pair);
    return pow2_scale(pair_bound(a, pair));
}
__device__ __forceinline__ float out_scale_of(const TcArgs &a, long long pair) {
    if (a.bs_mode != 1) return 1.f;
    return pow2_scale(fmaf(a.out_mul, pair_bound(a, pair), a.out_add) * 1.001f);
}

// Float offset of element (row, kk) inside a K-major [rows x 16 floats] tile with the 64-byte swizzle:
// rows are 64 bytes, the 16-byte chunk index (2 bits) is XORed with address bits [7,9) = (row >> 1) & 3.
__host__ __device__ __forceinline__ int sw_off(int row, int kk) { return row * SK + ((((kk >> 2) ^ ((row >> 1) & 3))) << 2) + (kk & 3); }
// Same tile geometry holding 32 halfs per row: BYTE offset of half element (row, kk), kk in [0,32)
__host__ __device__ __forceinline__ int sw_off_h(int row, int kk) { return row * 64 + ((((kk >> 3) ^ ((row >> 1) & 3))) << 4) + (kk & 7) * 2; }

// K-major, 64B-swizzled operand tile (tile base 1024-aligned): 8-row groups are 512 bytes apart.
//   bits [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=32) | [46,48) version=1 | [61,64) layout=4 (SWIZZLE_64B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// fp32 accumulate, A and B K-major:
//   c_format[4,6)=1 (F32) | a_format[7,10) | b_format[10,13) (kind::tf32: 2 = TF32; kind::f16: 0 = F16) | n_dim[17,23)=N>>3 | m_dim[24,29)=M>>4
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int f16 = 0) {
    return (1u << 4) | ((f16 ? 0u : 2u) << 7) | ((f16 ? 0u : 2u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return v > 0.f ? v : 0.1f * v;
    return v;
}

struct RowCtx {          // per-producer-thread description of its activation row for the current tile
    bool valid;
    const float *src0, *src1;      // PLAIN: src0 ; FC_H1: src0 = U1 row (centre point), src1 = U2 row (neighbour) ; SC2_Y1: src1 = P row
    float dx, dy, dz;
    float scale;                   // fmt 1: power-of-two scale of this row's frame pair (1 otherwise)
};

__device__ __forceinline__ RowCtx make_row(const TcArgs &a, long long c) {
    RowCtx r;
    r.valid = c < a.cols;
    r.src0 = r.src1 = nullptr; r.dx = r.dy = r.dz = 0.f; r.scale = 1.f;
    if (!r.valid) return r;
    if (a.prod == TC_PROD_PLAIN) {
        r.src0 = a.X + (size_t)c * a.ldx;
        if (a.bs_mode) r.scale = b_scale_of(a, div_i(c, a.cols_per_pair));
        return r;
    }
    const long long bi = div_i(c, a.ksamp);
    const int kk = (int)(c - bi * a.ksamp);
    const int b = (int)div_i(bi, a.n_pts), i = (int)(bi - (long long)b * a.n_pts);
    const int j = __ldg(a.nbr + (size_t)bi * a.nbr_ld + a.nbr_off + kk);
    const int nc = a.n_cand ? a.n_cand : a.n_pts;                      // candidate cloud may hold a different number of points (N1 != N2)
    const float *pq = a.xyz_q + (size_t)b * 3 * a.n_pts, *pc = a.xyz_c + (size_t)b * 3 * nc;
    r.dx = __fsub_rn(__ldg(pc + j), __ldg(pq + i));
    r.dy = __fsub_rn(__ldg(pc + nc + j), __ldg(pq + a.n_pts + i));
    r.dz = __fsub_rn(__ldg(pc + 2 * nc + j), __ldg(pq + 2 * a.n_pts + i));
    r.src0 = a.U1 ? a.U1 + (size_t)bi * 512 : nullptr;
    r.src1 = a.U2 + ((size_t)b * nc + j) * a.ld_u2 + a.off_u2;
    if (a.bs_mode) r.scale = b_scale_of(a, b);
    return r;
}

// One 32-float block of this thread's (gathered) row, BEFORE the split: 128 contiguous bytes.  For FC_H1 the
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

// the 4 transformed (pre-split) values of 16-byte chunk q of the thread's 32-block (one-CTA kernel: thread = row)
template <int PROD>
__device__ __forceinline__ void row_chunk(const float4 *sW, const RowCtx &r, int kb, int lane, int q, const float4 (&v)[8], const float4 &u, float (&x)[4]) {
    x[0] = v[q].x; x[1] = v[q].y; x[2] = v[q].z; x[3] = v[q].w;
    const int k0 = kb * PK + q * 4;
    if (PROD == TC_PROD_FC_H1) {
        const int srcl = (lane & ~7) + q;                     // the lane of this point's group that holds chunk q of the centre row
        const float uu[4] = {__shfl_sync(0xffffffffu, u.x, srcl), __shfl_sync(0xffffffffu, u.y, srcl),
                             __shfl_sync(0xffffffffu, u.z, srcl), __shfl_sync(0xffffffffu, u.w, srcl)};
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = act_apply(uu[e] + x[e] + small_term(sW, k0 + e, r), 2);
    } else if (PROD == TC_PROD_SC2_Y1) {
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e] + small_term(sW, k0 + e, r), 0.f);
    }
    if (!r.valid) { x[0] = x[1] = x[2] = x[3] = 0.f; }
}

// fmt 0: transform + split + swizzled store of half a 32-block (chunks 4*half .. 4*half+3) into one 16-float stage's B tiles
template <int PROD>
__device__ __forceinline__ void store_half(const float4 *sW, const RowCtx &r, int kb, int row, int lane, int half,
                                           const float4 (&v)[8], const float4 &u, float *Bhi, float *Blo) {
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
        float x[4];
        row_chunk<PROD>(sW, r, kb, lane, half * 4 + qq, v, u, x);
        float4 h, l;
        split_tf32(x[0], h.x, l.x); split_tf32(x[1], h.y, l.y); split_tf32(x[2], h.z, l.z); split_tf32(x[3], h.w, l.w);
        const int off = sw_off(row, qq * 4);
        *reinterpret_cast<float4 *>(Bhi + off) = h;
        *reinterpret_cast<float4 *>(Blo + off) = l;
    }
}
// fmt 1: the whole 32-block -> one 32-half stage row (four 16-byte chunks of hi and of lo)
template <int PROD>
__device__ __forceinline__ void store_row_f16(const float4 *sW, const RowCtx &r, int kb, int row, int lane,
                                              const float4 (&v)[8], const float4 &u, uint8_t *Bhi, uint8_t *Blo) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        float x0[4], x1[4];
        row_chunk<PROD>(sW, r, kb, lane, 2 * ch, v, u, x0);
        row_chunk<PROD>(sW, r, kb, lane, 2 * ch + 1, v, u, x1);
        uint4 h, l;
        split_f16x2(x0[0] * r.scale, x0[1] * r.scale, h.x, l.x); split_f16x2(x0[2] * r.scale, x0[3] * r.scale, h.y, l.y);
        split_f16x2(x1[0] * r.scale, x1[1] * r.scale, h.z, l.z); split_f16x2(x1[2] * r.scale, x1[3] * r.scale, h.w, l.w);
        const int off = sw_off_h(row, ch * 8);
        *reinterpret_cast<uint4 *>(Bhi + off) = h;
        *reinterpret_cast<uint4 *>(Blo + off) = l;
    }
}

// ---- epilogue ---------------------------------------------------------------------------------------------------------------------
// One pass over 32 accumulator columns held in r[] (this thread = output channel m): un-scale, bias / per-pair bias, activation and
// either a row-major store (+ per-pair |max| for the consumer's fp16 scale), a tiled (split, swizzled) store for the next GEMM, or
// the max over each point's ksamp neighbours.  The common case -- the 32 columns are all valid and belong to one frame pair -- takes a
// branch-free fast path; the generic path (bounds / pair-boundary checks per element) is kept for edge tiles.
struct EpiState {
    long long pair, pair_end;
    float pb;            // per-pair bias of channel m
    float ainv;          // a_inv[m]
    float inv;           // accumulator un-scale = a_inv[m] / b_scale(pair)
    float osc;           // tiled output scale of the pair (fmt 1)
    float amx;           // running max |out| of this thread in the current pair
    bool track;          // pair cursor in use
};

// act as max(v, v*slope): slope 1 = identity, 0 = ReLU, 0.1 = LeakyReLU(0.1)
__device__ __forceinline__ float act_slope(int act) { return act == 1 ? 0.f : (act == 2 ? 0.1f : 1.f); }

__device__ __forceinline__ void epi_flush_amax(const TcArgs &a, EpiState &es, int m) {      // warp-uniform call sites only
    if (!a.amax_out) return;
    float v = es.amx;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    if ((threadIdx.x & 31) == 0 && v > 0.f)
        atomicMax(a.amax_out + (size_t)(m / a.amax_group) * a.amax_ld + es.pair, __float_as_uint(v));
    es.amx = 0.f;
}
__device__ __forceinline__ void epi_load_pair(const TcArgs &a, EpiState &es, int m, bool m_ok, bool in_range) {
    if (a.pbias && m_ok && in_range) es.pb = __ldg(a.pbias + (size_t)es.pair * a.pb_ld + m);
    if (a.fmt == 1 && in_range) {
        es.inv = es.ainv * __frcp_rn(b_scale_of(a, es.pair));
        if (a.out_tiled) {
            es.osc = out_scale_of(a, es.pair);
            if (a.out_scale_store && m == 0) a.out_scale_store[es.pair] = es.osc;
        }
    }
}
__device__ __forceinline__ EpiState epi_begin(const TcArgs &a, long long c0, int m, bool m_ok) {
    EpiState es;
    es.pair = 0; es.pair_end = 0x7fffffffffffffffLL; es.pb = 0.f; es.inv = 1.f; es.osc = 1.f; es.amx = 0.f;
    es.ainv = (a.fmt == 1 && a.a_inv && m_ok) ? __ldg(a.a_inv + m) : 1.f;
    es.inv = es.ainv;
    es.track = a.pbias || a.bs_mode || a.amax_out;
    if (es.track) {
        es.pair = div_i(c0, a.cols_per_pair); es.pair_end = (es.pair + 1) * (long long)a.cols_per_pair;
        epi_load_pair(a, es, m, m_ok, c0 < a.cols);
    }
    return es;
}
__device__ __forceinline__ void epi_advance(const TcArgs &a, EpiState &es, long long c, int m, bool m_ok) {   // c >= es.pair_end (warp-uniform)
    epi_flush_amax(a, es, m);
    while (c >= es.pair_end) { ++es.pair; es.pair_end += a.cols_per_pair; }
    epi_load_pair(a, es, m, m_ok, c < a.cols);
}
__device__ __forceinline__ void epi_end(const TcArgs &a, EpiState &es, int m) { epi_flush_amax(a, es, m); }

template <int KSAMP>
__device__ __forceinline__ void maxk_groups(const uint32_t (&r)[32], float bias, long long cbase, int m, bool m_ok, const TcArgs &a,
                                            EpiState &es, bool uniform) {
#pragma unroll
    for (int g0 = 0; g0 < 32; g0 += KSAMP) {
        const long long c = cbase + g0;
        float inv = es.inv;
        if (!uniform && a.fmt == 1 && c < a.cols) inv = es.ainv * __frcp_rn(b_scale_of(a, div_i(c, a.cols_per_pair)));   // a point's columns share a pair
        float mx = 0.f;                                            // relu output >= 0
#pragma unroll
        for (int e = 0; e < KSAMP; ++e) mx = fmaxf(mx, fmaf(__uint_as_float(r[g0 + e]), inv, bias));
        if (c < a.cols && m_ok) a.Out[(size_t)(c / KSAMP) * a.ldo + m] = mx;
    }
}

__device__ __forceinline__ void epilogue_chunk(const TcArgs &a, const uint32_t (&r)[32], long long ct, long long c0, int cc, int m, bool m_ok,
                                               float bias, EpiState &es, int tile_b_floats) {
    const long long cfirst = c0 + cc;
    if (es.track && cfirst >= es.pair_end) epi_advance(a, es, cfirst, m, m_ok);       // pair cursor -> the chunk's first column
    const bool one_pair = (cfirst + 32 <= a.cols) && (cfirst + 32 <= es.pair_end);
    if (a.epi == TC_EPI_MAXK) {
        if (a.ksamp == 4) maxk_groups<4>(r, bias, cfirst, m, m_ok, a, es, one_pair);
        else if (a.ksamp == 8) maxk_groups<8>(r, bias, cfirst, m, m_ok, a, es, one_pair);
        else if (a.ksamp == 16) maxk_groups<16>(r, bias, cfirst, m, m_ok, a, es, one_pair);
        else maxk_groups<32>(r, bias, cfirst, m, m_ok, a, es, one_pair);
        return;
    }
    const bool fast = one_pair && (m_ok || a.out_tiled);
    const float slope = act_slope(a.act);
    if (fast) {
        const float badd = bias + es.pb;
        if (a.out_tiled && a.fmt == 1) {
            // tile (col_tile, 32-block m/32) = {hi 256 rows x 64 B, lo}; half element (row = column in tile, kk = m%32) at sw_off_h(row, kk);
            // (row>>1)&3 == (e>>1)&3 since cc % 8 == 0
            uint8_t *tb = reinterpret_cast<uint8_t *>(a.Out + ((size_t)ct * (a.M >> 5) + (m >> 5)) * (2 * (size_t)tile_b_floats)) + cc * 64 + (m & 7) * 2;
            const int kq = (m & 31) >> 3;
            // the (power-of-two) output scale is folded into the un-scale and the bias: act(s v) = s act(v) for s > 0, every product exact.
            // Two adjacent columns per packed-fp32 instruction.
            const float2 inv2 = make_float2(es.inv * es.osc, es.inv * es.osc), badd2 = make_float2(badd * es.osc, badd * es.osc);
            const float2 slope2 = make_float2(slope, slope);
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
                float2 v = __ffma2_rn(make_float2(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), inv2, badd2);
                const float2 vs = __fmul2_rn(v, slope2);
                v = make_float2(fmaxf(v.x, vs.x), fmaxf(v.y, vs.y));
                uint32_t hi, lo;
                split_f16x2(v, hi, lo);
                const int off = e * 64 + ((kq ^ ((e >> 1) & 3)) << 4);        // e even: columns e and e + 1 share the swizzle term
                *reinterpret_cast<unsigned short *>(tb + off) = (unsigned short)(hi & 0xffffu);
                *reinterpret_cast<unsigned short *>(tb + off + 64) = (unsigned short)(hi >> 16);
                *reinterpret_cast<unsigned short *>(tb + tile_b_floats * 4 + off) = (unsigned short)(lo & 0xffffu);
                *reinterpret_cast<unsigned short *>(tb + tile_b_floats * 4 + off + 64) = (unsigned short)(lo >> 16);
            }
        } else if (a.out_tiled) {
            // tile (col_tile, 16-block m/16); element (row = column in tile, kk = m%16) at sw_off(row, kk)
            float *tb = a.Out + ((size_t)ct * (a.M >> 4) + (m >> 4)) * (2 * (size_t)tile_b_floats) + cc * SK + (m & 3);
            const int kq = (m & 15) >> 2;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                float v = fmaf(__uint_as_float(r[e]), es.inv, badd);
                v = fmaxf(v, v * slope);
                float hi, lo;
                split_tf32(v, hi, lo);
                const int off = e * SK + ((kq ^ ((e >> 1) & 3)) << 2);
                tb[off] = hi;
                tb[tile_b_floats + off] = lo;
            }
        } else {
            float *o = a.Out + (size_t)cfirst * a.ldo + m;
            float amx = es.amx;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                float v = fmaf(__uint_as_float(r[e]), es.inv, badd);
                v = fmaxf(v, v * slope);
                amx = fmaxf(amx, fabsf(v));
                o[(size_t)e * a.ldo] = v;
            }
            es.amx = amx;
        }
        return;
    }
    // generic path (warp-uniform control flow: all lanes of a warp share the columns); fully unrolled so that r[] stays in registers
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        const long long c = cfirst + e;
        if (es.track && c >= es.pair_end) epi_advance(a, es, c, m, m_ok);
        const bool ok = c < a.cols;
        float v = fmaf(__uint_as_float(r[e]), es.inv, bias + es.pb);
        v = ok ? fmaxf(v, v * slope) : 0.f;
        if (a.out_tiled && a.fmt == 1) {
            uint8_t *tb = reinterpret_cast<uint8_t *>(a.Out + ((size_t)ct * (a.M >> 5) + (m >> 5)) * (2 * (size_t)tile_b_floats));
            unsigned short hi, lo;
            split_f16(v * es.osc, hi, lo);
            const int off = sw_off_h(cc + e, m & 31);
            *reinterpret_cast<unsigned short *>(tb + off) = hi;
            *reinterpret_cast<unsigned short *>(tb + tile_b_floats * 4 + off) = lo;
        } else if (a.out_tiled) {
            float *tb = a.Out + ((size_t)ct * (a.M >> 4) + (m >> 4)) * (2 * (size_t)tile_b_floats);
            float hi, lo;
            split_tf32(v, hi, lo);
            const int off = sw_off(cc + e, m & 15);
            tb[off] = hi;
            tb[tile_b_floats + off] = lo;
        } else if (ok && m_ok) {
            es.amx = fmaxf(es.amx, fabsf(v));
            a.Out[(size_t)c * a.ldo + m] = v;
        }
    }
}

}  // namespace tcdev
