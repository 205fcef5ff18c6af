// Probe: 2-D tensor TMA row loads (BOXW x 1 float boxes, SWIZZLE_NONE) exactly as stream.cuh issues them:
// negative / overhanging x coordinates, mbarrier complete_tx, one elected thread.  Bounded waits (test_wait).
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int BOXW>
__global__ void k(const __grid_constant__ CUtensorMap map, int x0, int y0, int rows, float* out, int* status) {
    extern __shared__ __align__(128) float smem[];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + rows * BOXW);
    const unsigned bar0 = smem_u32(bars);
    if (threadIdx.x == 0) {
        for (int r = 0; r < rows; ++r) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u * r));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int r = 0; r < rows; ++r) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * r), "r"(BOXW * 4) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                ::"r"(smem_u32(smem + r * BOXW)), "l"(&map), "r"(x0), "r"(y0 + r), "r"(bar0 + 8u * r) : "memory");
        }
    }
    int ok = 1;
    for (int r = 0; r < rows; ++r) {
        unsigned done = 0;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar0 + 8u * r), "r"(0u) : "memory");
        if (!done) ok = 0;
    }
    if (threadIdx.x == 0) status[0] = ok;
    __syncthreads();
    for (int i = threadIdx.x; i < rows * BOXW; i += blockDim.x) out[i] = smem[i];
}

int main() {
    const int w = 1000, h = 50, pitch = 1024, BOXW = 160, rows = 8;
    float* d; cudaMalloc(&d, (size_t)pitch * h * 4);
    float* hbuf = new float[pitch * h];
    for (int y = 0; y < h; ++y) for (int x = 0; x < pitch; ++x) hbuf[y * pitch + x] = y * 10000.f + x;
    cudaMemcpy(d, hbuf, (size_t)pitch * h * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    CUtensorMap map; memset(&map, 0, sizeof map);
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BOXW, 1u};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    float* out; int* st; cudaMalloc(&out, rows * BOXW * 4); cudaMalloc(&st, 4);
    const size_t smem = rows * BOXW * 4 + rows * 8;
    cudaFuncSetAttribute(k<BOXW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int x0 : {0, 40, -20, 900, 960}) {
        cudaMemset(st, 0xff, 4);
        k<BOXW><<<1, 96, smem>>>(map, x0, 3, rows, out, st);
        cudaError_t e = cudaDeviceSynchronize();
        int hs = -1; float ho[rows * BOXW];
        cudaMemcpy(&hs, st, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(ho, out, sizeof ho, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < rows; ++rr) for (int c = 0; c < BOXW; ++c) {
            const int x = x0 + c; const float want = (x >= 0 && x < w) ? (3 + rr) * 10000.f + x : 0.f;
            if (ho[rr * BOXW + c] != want) ++bad;
        }
        printf("x0 %4d: %s, all barriers completed %d, mismatches %d (first values %.0f %.0f)\n", x0, cudaGetErrorString(e), hs, bad, ho[0], ho[BOXW - 1]);
    }
    return 0;
}
