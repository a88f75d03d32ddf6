// SM-issued stores into page-locked host memory (the zero-copy route of gpr_step_host): GB/s by store shape.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/sm_store_bw tools/sm_store_bw.cu && /tmp/sm_store_bw
// Shapes: bytes per lane per store instruction and the stride between lanes (stride > width = partially written sectors).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <typename T>
__global__ void fill(T* dst, size_t n_elems, int stride_elems, T v) {
    // element i of the logical stream goes to dst[i * stride]; consecutive lanes write consecutive logical elements
    const size_t total = n_elems;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        dst[i * (size_t)stride_elems] = v;
}

template <typename T>
static double run(T* dst, size_t bytes_written, int stride_elems, int ctas, T v) {
    const size_t n = bytes_written / sizeof(T);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) fill<T><<<ctas, 128>>>(dst, n, stride_elems, v);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int reps = 10;
    for (int r = 0; r < reps; ++r) fill<T><<<ctas, 128>>>(dst, n, stride_elems, v);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return (double)reps * (double)bytes_written / (ms * 1e-3) / 1e9;
}

int main() {
    const size_t cap = 64u << 20;
    void* h = nullptr;
    cudaHostAlloc(&h, cap, cudaHostAllocMapped);
    void* d = nullptr;
    cudaHostGetDevicePointer(&d, h, 0);
    void* g = nullptr;
    cudaMalloc(&g, cap);
    const size_t mb8 = 8u << 20;
    printf("{");
    for (int ctas : {148, 148 * 6, 148 * 16}) {
        printf("\"ctas_%d\": {", ctas);
        printf("\"1B_contig\": %.1f, ", run<uint8_t>((uint8_t*)d, mb8 / 4, 1, ctas, 1));
        printf("\"4B_contig\": %.1f, ", run<uint32_t>((uint32_t*)d, mb8, 1, ctas, 1u));
        printf("\"8B_contig\": %.1f, ", run<uint2>((uint2*)d, mb8, 1, ctas, make_uint2(1, 2)));
        printf("\"16B_contig\": %.1f, ", run<uint4>((uint4*)d, mb8, 1, ctas, make_uint4(1, 2, 3, 4)));
        printf("\"8B_stride16\": %.1f, ", run<uint2>((uint2*)d, mb8, 2, ctas, make_uint2(1, 2)));
        printf("\"8B_stride32\": %.1f, ", run<uint2>((uint2*)d, mb8 / 2, 4, ctas, make_uint2(1, 2)));
        printf("\"4B_stride32\": %.1f, ", run<uint32_t>((uint32_t*)d, mb8 / 4, 8, ctas, 1u));
        printf("\"16B_contig_to_hbm\": %.1f}, ", run<uint4>((uint4*)g, mb8, 1, ctas, make_uint4(1, 2, 3, 4)));
    }
    // copy engine for comparison
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaMemcpyAsync(h, g, mb8, cudaMemcpyDeviceToHost);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) cudaMemcpyAsync(h, g, mb8, cudaMemcpyDeviceToHost);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("\"copy_engine_d2h_8MB\": %.1f, \"unit\": \"GB/s of bytes written, 8 MB streams into cudaHostAllocMapped memory\"}\n", 10.0 * mb8 / (ms * 1e-3) / 1e9);
    return 0;
}
