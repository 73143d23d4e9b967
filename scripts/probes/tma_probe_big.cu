// Stand-alone probe of the tensor-map chunk copies used by the row kernels
// (fft_core.cuh: tma_load_chunks / tma_store_chunks).  nvcc -arch=sm_100a tma_probe.cu -o tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>
typedef unsigned long long mbar_t;
__global__ void probe(const void* tmap, float* out, int y, int xb0, int img) {
    extern __shared__ __align__(128) unsigned char smem[];
    mbar_t* bar = (mbar_t*)(smem + 136 * 64);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(136 * 64) : "memory");
        const int c0 = y * 4 * 2;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(d), "l"(tmap), "r"(b), "r"(c0), "r"(xb0), "r"(img) : "memory");
    }
    unsigned done = 0; int spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(b), "r"(0u) : "memory");
    } while (!done && ++spins < 1000000);
    for (int i = threadIdx.x; i < 4 * 16; i += blockDim.x) out[i] = ((float*)smem)[i];
    if (threadIdx.x == 0) out[64] = (float)spins;
    // store test: write back +1000 to the same box of image 1 - img
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 16; i += blockDim.x) ((float*)smem)[i] += 1000.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const int c0 = y * 4 * 2;
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(tmap), "r"(d), "r"(c0), "r"(xb0), "r"(1 - img) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}
int main() {
    const int rows_e = 1184, C = 4, nxb = 271, nimg = 2;
    const size_t slab = (size_t)rows_e * C * 2, n = slab * nxb * nimg;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)i;
    float* d; cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    void* fn = 0; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap map;
    const cuuint64_t dims[3] = {slab, (cuuint64_t)nxb, (cuuint64_t)nimg};
    const cuuint64_t strides[2] = {slab * 4, slab * 4 * nxb};
    const cuuint32_t box[3] = {16, 136, 1}, es[3] = {1, 1, 1};
    CUresult r = ((enc_t)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d entry %p q %d\n", (int)r, fn, (int)q);
    void* dmap; cudaMalloc(&dmap, sizeof(map)); cudaMemcpy(dmap, &map, sizeof(map), cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 65 * 4);
    const int y = 1100, xb0 = 136, img = 1;
    probe<<<1, 64, 16384>>>(dmap, out, y, xb0, img);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    float res[65]; cudaMemcpy(res, out, 65 * 4, cudaMemcpyDeviceToHost);
    printf("spins %g\n", res[64]);
    for (int blk = 0; blk < 4; ++blk) {
        printf("box row %d (xb %d): got %g..%g, expect %g..%g%s\n", blk, xb0 + blk, res[blk * 16], res[blk * 16 + 15],
               (xb0 + blk) < nxb ? (float)(img * slab * nxb + (xb0 + blk) * slab + y * 8) : 0.f,
               (xb0 + blk) < nxb ? (float)(img * slab * nxb + (xb0 + blk) * slab + y * 8 + 15) : 0.f, (xb0 + blk) < nxb ? "" : " (OOB)");
    }
    cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost);
    const size_t o = (size_t)(1 - img) * slab * nxb + (size_t)xb0 * slab + y * 8;
    printf("store: image 1 block %d got %g (expect %g)\n", xb0, h[o], (float)(img * slab * nxb + xb0 * slab + y * 8) + 1000.f);
    return 0;
}
