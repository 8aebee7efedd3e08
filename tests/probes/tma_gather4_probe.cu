// Probe (run on the GPU box): semantics of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe tests/probes/tma_gather4_probe.cu && /tmp/tma_probe
// Question answered: which boxDim the tensor map needs (cols x 1 or cols x 4) and how the four rows land in shared memory.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int col, int r0, int r1, int r2, int r3, float *out) {
    __shared__ __align__(128) float dst[4 * 32];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 128; i += blockDim.x) dst[i] = -1.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(512) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(d), "l"(&tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(b) : "memory");
    }
    uint32_t done = 0; int spins = 0;
    while (!done && ++spins < 100000)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(b), "r"(0) : "memory");
    __syncthreads();
    for (int i = threadIdx.x; i < 128; i += blockDim.x) out[i] = dst[i];
    if (threadIdx.x == 0) out[128] = done ? 1.f : 0.f;
}

int main() {
    const int rows = 1000, cols = 2048;
    std::vector<float> h((size_t)rows * cols);
    for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) h[(size_t)r * cols + c] = (float)(r * 10000 + c);
    float *g, *out;
    cudaMalloc(&g, h.size() * 4); cudaMemcpy(g, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 129 * 4);
    EncodeTiled enc = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &q);
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    for (int boxrows = 1; boxrows <= 4; boxrows += 3) {
        CUtensorMap tm;
        cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, gstride[1] = {(cuuint64_t)cols * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)boxrows}, estr[2] = {1, 1};
        CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("boxDim {32,%d}: encode rc=%d\n", boxrows, (int)rc);
        if (rc != CUDA_SUCCESS) continue;
        const int col = 64, r[4] = {5, 900, 17, 3};
        cudaMemset(out, 0, 129 * 4);
        probe<<<1, 128>>>(tm, col, r[0], r[1], r[2], r[3], out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  launch: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 2;
        float ho[129]; cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
        printf("  barrier completed: %g\n", ho[128]);
        int ok = 1;
        for (int i = 0; i < 4; ++i) {
            printf("  smem row %d: %.0f %.0f ... %.0f (want %d ... %d)\n", i, ho[i * 32], ho[i * 32 + 1], ho[i * 32 + 31], r[i] * 10000 + col, r[i] * 10000 + col + 31);
            for (int j = 0; j < 32; ++j) ok &= ho[i * 32 + j] == (float)(r[i] * 10000 + col + j);
        }
        printf("  => %s\n", ok ? "MATCH: four rows land contiguously, 128 bytes each, in the order given" : "MISMATCH");
    }
    return 0;
}
