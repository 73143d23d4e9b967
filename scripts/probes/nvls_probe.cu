// Probe: does this box give CUDA multicast objects (NVLS: in-switch reduction / multicast over
// NVSwitch) to a plain process?  One process, all visible GPUs: one multicast object over N
// devices, a buffer of each GPU bound to it, then on GPU 0
//   multimem.ld_reduce.add.v4.f32 (sum of the N replicas, reduced in the switch) and
//   multimem.st.v4.f32 (one store lands in every replica).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o nvls_probe nvls_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CU(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_ = 0; cuGetErrorString(r_, &s_); \
    printf("FAILED %s -> %d (%s) at line %d\n", #x, (int)r_, s_ ? s_ : "?", __LINE__); return 1; } } while (0)
#define RT(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("FAILED %s -> %s at line %d\n", #x, cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void fill(float* p, size_t n, float v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v + (float)(i % 7);
}
__global__ void reduce_and_broadcast(float* mc, size_t n4) {   // n4 = float4 elements
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float a, b, c, d;
        const float* p = mc + 4 * i;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(p) : "memory");
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
    }
}

int main() {
    CU(cuInit(0));
    int n = 0;
    RT(cudaGetDeviceCount(&n));
    printf("devices: %d\n", n);
    if (n < 2) { printf("need 2 GPUs\n"); return 0; }
    std::vector<CUdevice> dev(n);
    for (int i = 0; i < n; ++i) {
        CU(cuDeviceGet(&dev[i], i));
        int mc = 0;
        CU(cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev[i]));
        printf("device %d multicast supported: %d\n", i, mc);
        if (!mc) return 0;
        RT(cudaSetDevice(i));
        RT(cudaFree(0));
    }
    const size_t want = 32u << 20;
    CUmulticastObjectProp mp = {};
    mp.numDevices = n; mp.size = want; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR; mp.flags = 0;
    size_t gran = 0;
    CU(cuMulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    const size_t size = (want + gran - 1) / gran * gran;
    mp.size = size;
    printf("granularity %zu, size %zu\n", gran, size);
    CUmemGenericAllocationHandle mch;
    CU(cuMulticastCreate(&mch, &mp));
    int fd = -1;
    CU(cuMemExportToShareableHandle(&fd, mch, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    printf("multicast object created, exportable fd %d\n", fd);
    for (int i = 0; i < n; ++i) CU(cuMulticastAddDevice(mch, dev[i]));
    std::vector<CUdeviceptr> uc(n), mc(n);
    std::vector<CUmemGenericAllocationHandle> mem(n);
    for (int i = 0; i < n; ++i) {
        RT(cudaSetDevice(i));
        CUmemAllocationProp ap = {};
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ap.location.id = i;
        ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        CU(cuMemCreate(&mem[i], size, &ap, 0));
        CU(cuMulticastBindMem(mch, 0, mem[i], 0, size, 0));
        CUmemAccessDesc ad = {};
        ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = i; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        CU(cuMemAddressReserve(&uc[i], size, gran, 0, 0));
        CU(cuMemMap(uc[i], size, 0, mem[i], 0));
        CU(cuMemSetAccess(uc[i], size, &ad, 1));
        CU(cuMemAddressReserve(&mc[i], size, gran, 0, 0));
        CU(cuMemMap(mc[i], size, 0, mch, 0));
        CU(cuMemSetAccess(mc[i], size, &ad, 1));
        fill<<<256, 256>>>((float*)uc[i], size / 4, (float)(i + 1));
        RT(cudaDeviceSynchronize());
    }
    RT(cudaSetDevice(0));
    cudaEvent_t e0, e1;
    RT(cudaEventCreate(&e0)); RT(cudaEventCreate(&e1));
    for (int rep = 0; rep < 3; ++rep) {
        for (int i = 0; i < n; ++i) { RT(cudaSetDevice(i)); fill<<<256, 256>>>((float*)uc[i], size / 4, (float)(i + 1)); RT(cudaDeviceSynchronize()); }
        RT(cudaSetDevice(0));
        RT(cudaEventRecord(e0));
        reduce_and_broadcast<<<148 * 4, 256>>>((float*)mc[0], size / 16);
        RT(cudaEventRecord(e1));
        RT(cudaDeviceSynchronize());
        float ms = 0; RT(cudaEventElapsedTime(&ms, e0, e1));
        printf("one GPU reduces + broadcasts %zu MB over %d replicas: %.3f ms (%.1f GB/s of buffer)\n", size >> 20, n, ms, size / ms / 1e6);
    }
    // check: every replica holds sum_i (i+1 + k%7)
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        RT(cudaSetDevice(i));
        std::vector<float> h(1024);
        RT(cudaMemcpy(h.data(), (void*)uc[i], sizeof(float) * h.size(), cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < h.size(); ++k) {
            const float want_v = n * (n + 1) / 2.0f + n * (float)(k % 7);
            if (h[k] != want_v) { if (bad < 5) printf("replica %d [%zu] = %g, want %g\n", i, k, h[k], want_v); ++bad; }
        }
    }
    printf(bad ? "MISMATCH (%d)\n" : "multicast reduce + broadcast verified on every replica (%d mismatches)\n", bad);
    return 0;
}
