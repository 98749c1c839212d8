#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, float* out, int c0, int c1, int c2, int elems, int variant) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    float* s = (float*)smem_raw;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        if (variant == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(elems * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
            "r"(smem_u32(s)), "l"(&tmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    for (int i = threadIdx.x; i < elems; i += blockDim.x) out[i] = s[i];
}
int run(int w2, int rows, int planes, int b0, int b1, int b2, int c0, int c1, int c2, int variant) {
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn encode = (EncodeTiledFn)fn;
    size_t n = (size_t)w2 * rows * planes;
    std::vector<float> h(n); for (size_t i = 0; i < n; ++i) h[i] = (float)(i + 1);
    float *d, *o; cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    int elems = b0 * b1 * b2; cudaMalloc(&o, elems * 4); cudaMemset(o, 0xff, elems * 4);
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)w2, (cuuint64_t)rows, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)w2 * 4, (cuuint64_t)w2 * 4 * rows};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2}, es[3] = {1, 1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("dims %dx%dx%d box %dx%dx%d at (%d,%d,%d) variant %d: encode=%d ", w2, rows, planes, b0, b1, b2, c0, c1, c2, variant, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<<<1, 128, elems * 4 + 128>>>(m, o, c0, c1, c2, elems, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> res(elems); cudaMemcpy(res.data(), o, elems * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int z = 0; z < b2; ++z) for (int y = 0; y < b1; ++y) for (int x = 0; x < b0; ++x) {
            int gx = c0 + x, gy = c1 + y, gz = c2 + z;
            float want = (gx >= 0 && gx < w2 && gy >= 0 && gy < rows && gz >= 0 && gz < planes) ? h[((size_t)gz * rows + gy) * w2 + gx] : 0.0f;
            if (res[((size_t)z * b1 + y) * b0 + x] != want) ++bad;
        }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    return e != cudaSuccess;
}
int main(int argc, char** argv) {
    int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0) return run(576, 192, 6, 148, 42, 1, -10, -5, 2, 0);     // production shape
    if (which == 1) return run(48, 16, 3, 148, 42, 1, -10, -5, 1, 0);       // box larger than tensor
    if (which == 2) return run(48, 16, 3, 148, 42, 1, -10, -5, 1, 1);       // other fence
    if (which == 3) return run(576, 192, 18, 148, 20, 3, -4, -2, 3, 0);     // kernel A shape
    if (which == 4) return run(48, 16, 3, 48, 16, 1, 0, 0, 1, 0);           // box == tensor
    return 0;
}
