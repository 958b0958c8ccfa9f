// Micro-benchmark: achievable HBM throughput for the step kernel's read/write mix (reads R MB, writes W MB per launch).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mixbw tools/mixbw.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mix(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t nin, size_t nout, int ratio_num, int ratio_den) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    // every thread: for each block of work, read `ratio_den` vectors and write `ratio_num` vectors
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (size_t k = i; k * ratio_den < nin; k += stride) {
        for (int r = 0; r < ratio_den; ++r) {
            size_t j = k + (size_t)r * (nin / ratio_den);
            if (j < nin) { uint4 v = in[j]; acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w; }
        }
        for (int w = 0; w < ratio_num; ++w) {
            size_t j = k + (size_t)w * (nout / ratio_num);
            if (j < nout) __stcs(out + j, acc);
        }
    }
}
int main() {
    const size_t R = 169u << 20, W = 387u << 20;
    uint4 *in, *out;
    cudaMalloc(&in, R); cudaMalloc(&out, W);
    cudaMemset(in, 1, R); cudaMemset(out, 0, W);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks : {148 * 8, 148 * 16, 148 * 32}) {
        for (int it = 0; it < 5; ++it) mix<<<blocks, 256>>>(in, out, R / 16, W / 16, 16, 7);
        cudaEventRecord(e0);
        const int K = 200;
        for (int it = 0; it < K; ++it) mix<<<blocks, 256>>>(in, out, R / 16, W / 16, 16, 7);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("blocks=%d: %.1f us per launch, %.0f GB/s (read %zu MB + write %zu MB)\n", blocks, 1e3 * ms / K, (R + W) / (ms / K * 1e-3) / 1e9, R >> 20, W >> 20);
    }
    // plain copy for reference (read W/2.. use 278 MB each way)
    const size_t Cb = 278u << 20;
    uint4 *a, *b; cudaMalloc(&a, Cb); cudaMalloc(&b, Cb);
    for (int it = 0; it < 5; ++it) cudaMemcpyAsync(b, a, Cb, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    for (int it = 0; it < 200; ++it) cudaMemcpyAsync(b, a, Cb, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("memcpy d2d 278 MB: %.1f us, %.0f GB/s (read+write)\n", 1e3 * ms / 200, 2.0 * Cb / (ms / 200 * 1e-3) / 1e9);
    // pure write
    for (int it = 0; it < 5; ++it) cudaMemsetAsync(out, 0, W);
    cudaEventRecord(e0);
    for (int it = 0; it < 200; ++it) cudaMemsetAsync(out, 0, W);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("memset 387 MB: %.1f us, %.0f GB/s (write only)\n", 1e3 * ms / 200, (double)W / (ms / 200 * 1e-3) / 1e9);
    return 0;
}
