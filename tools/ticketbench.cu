// How fast can CTAs draw tickets from ONE global counter? (decides the scheduling of persistent kernels)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/ticketbench tools/ticketbench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ticket(int* counter, int total, int work, float* sink)
{
    __shared__ int s_t;
    float acc = 0.f;
    for (;;) {
        if (threadIdx.x == 0) s_t = atomicAdd(counter, 1);
        __syncthreads();
        const int t = s_t;
        __syncthreads();
        if (t >= total) break;
        for (int i = 0; i < work; ++i) acc = acc * 1.0001f + t;      // stand-in for the tile
    }
    if (acc == -1.f) sink[0] = acc;
}
__global__ void k_static(int total, int work, float* sink)
{
    float acc = 0.f;
    for (int t = blockIdx.x; t < total; t += gridDim.x)
        for (int i = 0; i < work; ++i) acc = acc * 1.0001f + t;
    if (acc == -1.f) sink[0] = acc;
}
int main()
{
    int* c; float* s; cudaMalloc(&c, 4); cudaMalloc(&s, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas : {148, 296, 444, 1184})
        for (int work : {0, 2000, 20000}) {
            const int total = 130000;
            float ms, ms2;
            cudaMemset(c, 0, 4);
            k_ticket<<<ctas, 256>>>(c, 1000, work, s); cudaDeviceSynchronize();
            cudaMemset(c, 0, 4);
            cudaEventRecord(e0); k_ticket<<<ctas, 256>>>(c, total, work, s); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            cudaEventRecord(e0); k_static<<<ctas, 256>>>(total, work, s); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms2, e0, e1);
            printf("ctas %4d work %5d: dynamic tickets %.3f ms (%.1f ns/ticket)   static %.3f ms\n", ctas, work, ms,
                   ms * 1e6 / total, ms2);
        }
    return 0;
}
