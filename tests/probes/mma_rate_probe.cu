// Probe (GPU box): issue rate of tcgen05.mma.cta_group::2 kind::f16, M=256, as a function of the operand layout / N / operand source.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tests/probes/_bin/mma_probe tests/probes/mma_rate_probe.cu && tests/probes/_bin/mma_probe
// One cluster of two CTAs; the leader's thread 0 issues `iters` MMAs back to back on fixed (zeroed) shared-memory operands, then commits and
// waits.  Prints clocks per MMA.  (No data dependence: this is the pipe's own rate for that instruction shape, the ceiling of the GEMM kernels.)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw64(uint32_t a) { return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) { return (uint64_t)((a >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61); }
__host__ __device__ constexpr uint32_t idesc(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// modes 5 / 6: the 3-product operand pattern over 4 rotating stages without / with the collector hints;
// mode 0: SS, SW64, one accumulator; 1: SS, SW64, keep/reuse pattern of the 3-product split; 2: SS, SW128; 3: TS (A in TMEM), SW64 B; 4: SS SW64 rotating over 4 stages
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(int mode, int N, int iters, long long *out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t *smem = raw + (base - smem_u32(raw));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_slot;
    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t id = idesc(256, N);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t st = (mode >= 4) ? (uint32_t)(i / 6 % 4) * 32768u : 0u;
            const uint32_t koff = (uint32_t)(i & 1) * 2;                     // alternate the two K16 halves of a 32-half row
            const uint32_t d = tm + (uint32_t)((i / 96) & 1) * 256;           // switch accumulator every 96 MMAs (a "tile")
            if (mode == 2) {
                const uint64_t a = desc_sw128(base + st) + koff, b = desc_sw128(base + st + 16384) + koff;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
            } else if (mode == 3) {
                const uint64_t b = desc_sw64(base + 16384) + koff;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm + 256), "r"(tm + 64 + (uint32_t)(i & 7) * 8), "l"(b), "r"(id) : "memory");
            } else if (mode == 5) {
                const uint64_t a = desc_sw64(base + st + (i % 3 == 0 ? 8192 : 0)) + (uint32_t)((i / 3) & 1) * 2, b = desc_sw64(base + st + 16384 + (i % 3 == 1 ? 8192 : 0)) + (uint32_t)((i / 3) & 1) * 2;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
            } else if (mode == 1 || mode == 6) {
                const uint32_t k3 = mode == 6 ? (uint32_t)((i / 3) & 1) * 2 : koff;
                const uint64_t a = desc_sw64(base + st + (i % 3 == 0 ? 8192 : 0)) + k3, b = desc_sw64(base + st + 16384 + (i % 3 == 1 ? 8192 : 0)) + k3;
                if (i % 3 == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
                else if (i % 3 == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
                else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
            } else {
                const uint64_t a = desc_sw64(base + st) + koff, b = desc_sw64(base + st + 16384) + koff;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
            }
        }
        const long long t1 = clock64();
        const uint16_t mask = 3;
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"(mask) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    } else if (threadIdx.x == 0) {
        uint32_t done = 0;      // the peer waits for the multicast commit too, so that it does not free TMEM early
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
    }
}

int main() {
    long long *out; cudaMalloc(&out, 16);
    const int SM = 170 * 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
    const char *names[7] = {"SS SW64 plain", "SS SW64 fill/lastuse pattern (3-product split)", "SS SW128", "TS (A in TMEM), B SW64", "SS SW64 rotating 4 stages",
                            "3-product pattern, 4 rotating stages, NO collector hints", "3-product pattern, 4 rotating stages, fill/lastuse hints"};
    const int iters = 3000;
    for (int mode = 0; mode < 7; ++mode)
        for (int N = 256; N >= 64; N >>= 1) {
            if (mode == 3 && N == 256) continue;          // D + A would not fit next to each other the way the probe places them
            for (int rep = 0; rep < 2; ++rep) {
                probe<<<2, 128, SM>>>(mode, N, iters, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e)); return 1; }
            }
            long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("%-48s M=256 N=%3d K=16: issue %.1f clk/MMA, retired %.1f clk/MMA (ideal %d)\n", names[mode], N, (double)h[0] / iters, (double)h[1] / iters, N / 2);
        }
    return 0;
}
