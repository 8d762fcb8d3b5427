// Throughput probes that decide the scatter design (run on the B200 via gpurun).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__global__ void k_red_scalar(float* buf, size_t n, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) { size_t j = (i + (size_t)it * stride) % n; atomicAdd(buf + j, 1.0f); }
}
__global__ void k_red_v4(float4* buf, size_t n4, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) { size_t j = (i + (size_t)it * stride) % n4; atomicAdd(buf + j, make_float4(1.f,1.f,1.f,1.f)); }
}
__global__ void k_red_v2(float2* buf, size_t n2, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) { size_t j = (i + (size_t)it * stride) % n2; atomicAdd(buf + j, make_float2(1.f,1.f)); }
}
__global__ void k_store(float4* buf, size_t n4, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) { size_t j = (i + (size_t)it * stride) % n4; buf[j] = make_float4(1.f,1.f,1.f,1.f); }
}
// shared-memory variants: mode 0 atomicAdd lane-consecutive, 1 atomicAdd pseudo-random, 2 plain RMW lane-consecutive
__global__ void k_smem(float* out, int iters, int mode) {
    __shared__ float acc[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    unsigned r = threadIdx.x * 2654435761u + blockIdx.x;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) { atomicAdd(&acc[(warp * 997 + it * 37 + lane) & 8191], 1.0f); }
        else if (mode == 1) { r = r * 1664525u + 1013904223u; atomicAdd(&acc[(r >> 8) & 8191], 1.0f); }
        else if (mode == 2) { int a = (warp * 1024 + ((it * 37 + lane) & 1023)); float v = acc[a]; acc[a] = v + 1.0f; __syncwarp(); }
        else if (mode == 3) { atomicAdd(reinterpret_cast<int*>(&acc[(warp * 997 + it * 37 + lane) & 8191]), 1); }
        else if (mode == 4) { r = r * 1664525u + 1013904223u; atomicAdd(reinterpret_cast<int*>(&acc[(r >> 8) & 8191]), 1); }
        else if (mode == 5) { atomicAdd(reinterpret_cast<unsigned long long*>(&acc[((warp * 997 + it * 37 + lane) * 2) & 8190]), 1ull); }
        else if (mode == 6) { atomicAdd(reinterpret_cast<int*>(&acc[(warp * 997 + it * 37 + lane - (lane >> 3)) & 8191]), 1); }   // 3 duplicate pairs per warp-op
        else if (mode >= 8) { int a = (warp * 997 + it * 37 + lane) & 8191; const int act = mode == 8 ? 16 : (mode == 9 ? 8 : 32);
               if (lane < act) { atomicAdd(reinterpret_cast<int*>(&acc[a]), 2); atomicAdd(reinterpret_cast<int*>(&acc[(a + 1) & 8191]), 3);
               atomicAdd(reinterpret_cast<int*>(&acc[(a + 96) & 8191]), 4); atomicAdd(reinterpret_cast<int*>(&acc[(a + 97) & 8191]), 5); } }
        else { int a = (warp * 997 + it * 37 + lane) & 8191; atomicAdd(reinterpret_cast<int*>(&acc[a]), 1); atomicAdd(reinterpret_cast<int*>(&acc[(a + 1) & 8191]), 2);
               atomicAdd(reinterpret_cast<int*>(&acc[(a + 96) & 8191]), 3); atomicAdd(reinterpret_cast<int*>(&acc[(a + 97) & 8191]), 4); }
    }
    __syncthreads();
    float s = 0; for (int i = threadIdx.x; i < 8192; i += blockDim.x) s += acc[i];
    if (s == -1.f) out[0] = s;
}
int main() {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int big = 0; big < 2; ++big) {
        size_t n = big ? (size_t)400 << 20 : (size_t)8 << 20;   // floats: 1.6 GB or 32 MB (L2 resident)
        float* buf; CK(cudaMalloc(&buf, n * 4)); CK(cudaMemset(buf, 0, n * 4));
        int grid = 148 * 8, block = 256, iters = 256;
        double ops = (double)grid * block * iters;
        k_red_scalar<<<grid, block>>>(buf, n, 8); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k_red_scalar<<<grid, block>>>(buf, n, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%s red.f32   coalesced: %.1f G lane-ops/s  = %.1f GB/s payload\n", big ? "1.6GB" : "32MB ", ops / ms / 1e6, ops * 4 / ms / 1e6);
        cudaEventRecord(e0); k_red_v2<<<grid, block>>>((float2*)buf, n / 2, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%s red.v2.f32 coalesced: %.1f G lane-ops/s = %.1f GB/s payload\n", big ? "1.6GB" : "32MB ", ops / ms / 1e6, ops * 8 / ms / 1e6);
        cudaEventRecord(e0); k_red_v4<<<grid, block>>>((float4*)buf, n / 4, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%s red.v4.f32 coalesced: %.1f G lane-ops/s = %.1f GB/s payload\n", big ? "1.6GB" : "32MB ", ops / ms / 1e6, ops * 16 / ms / 1e6);
        cudaEventRecord(e0); k_store<<<grid, block>>>((float4*)buf, n / 4, iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%s st.v4      coalesced: %.1f GB/s\n", big ? "1.6GB" : "32MB ", ops * 16 / ms / 1e6);
        cudaFree(buf);
    }
    float* o; cudaMalloc(&o, 4);
    for (int mode = 0; mode < 11; ++mode) {
        int grid = 148 * 4, block = 256, iters = 4096;
        k_smem<<<grid, block>>>(o, 16, mode); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k_smem<<<grid, block>>>(o, iters, mode); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)grid * block * iters;
        const char* nm[] = {"smem atomicAdd f32 consecutive", "smem atomicAdd f32 random", "smem plain RMW + syncwarp", "smem atomicAdd i32 consecutive", "smem atomicAdd i32 random", "smem atomicAdd u64 consecutive", "smem atomicAdd i32 consecutive, 3 duplicate pairs", "smem atomicAdd i32 4 taps per iteration (ops counted per iteration)", "4 taps, 16 of 32 lanes active", "4 taps, 8 of 32 lanes active", "4 taps, 32 lanes (ATOMS.ADD form)"};
        printf("%s: %.1f G lane-ops/s  (%.2f clk per warp-op per SM at 1.965 GHz)\n", nm[mode], ops / ms / 1e6,
               1.965e9 / (ops / 32 / 148 / (ms * 1e-3)));
    }
    return 0;
}
