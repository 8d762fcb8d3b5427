// Read-bandwidth probes for the streaming kernels (run on the B200 via gpurun): how do the number of concurrent
// streams, the per-CTA chunking and the loads in flight per thread change the achieved HBM read rate?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/readbench tools/readbench.cu && /tmp/readbench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// mode 0: ONE contiguous stream of 2*HW floats per item; mode 1: two plane streams (f+i, f+HW+i);
// mode 2: two planes + byte mask.  ITERS x float4 per thread, CTA-contiguous chunks, UNR loads in flight.
template <int MODE, int ITERS, int UNR>
__global__ void __launch_bounds__(256) k_read(const float* __restrict__ flow, const uint8_t* __restrict__ mask,
                                              unsigned* out, int HW)
{
    const int n = blockIdx.y;
    const float* f = flow + (size_t)n * 2 * HW;
    const uint8_t* m = mask + (size_t)n * HW;
    unsigned acc = 0;
    const int span = (MODE == 0) ? 2 * HW : HW;
    const int base = blockIdx.x * (ITERS * 1024) + threadIdx.x * 4;
#pragma unroll UNR
    for (int k = 0; k < ITERS; ++k) {
        const int i = base + k * 1024;
        if (i >= span) break;
        const float4 a = __ldg(reinterpret_cast<const float4*>(f + i));
        acc = max(acc, max(max(__float_as_uint(a.x), __float_as_uint(a.y)), max(__float_as_uint(a.z), __float_as_uint(a.w))));
        if (MODE >= 1) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(f + HW + i));
            acc = max(acc, max(max(__float_as_uint(b.x), __float_as_uint(b.y)), max(__float_as_uint(b.z), __float_as_uint(b.w))));
        }
        if (MODE >= 2) {
            const uchar4 mm = __ldg(reinterpret_cast<const uchar4*>(m + i));
            acc = max(acc, (unsigned)(mm.x | mm.y | mm.z | mm.w));
        }
    }
    acc = __reduce_max_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && acc == 0x7fffffffu) atomicMax(out + n, acc);
}

// persistent grid-stride over the whole batch as one flat stream (what torch's reduction does)
template <int UNR>
__global__ void __launch_bounds__(512) k_flat(const float4* __restrict__ p, size_t n4, unsigned* out)
{
    unsigned acc = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
#pragma unroll UNR
    for (; i < n4; i += stride) {
        const float4 a = __ldg(p + i);
        acc = max(acc, max(max(__float_as_uint(a.x), __float_as_uint(a.y)), max(__float_as_uint(a.z), __float_as_uint(a.w))));
    }
    acc = __reduce_max_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && acc == 0x7fffffffu) atomicMax(out, acc);
}

template <typename F> float timeit(F launch, int reps = 20)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main()
{
    const int N = 64, H = 1080, W = 1920, HW = H * W;
    float* flow; uint8_t* mask; unsigned* out;
    cudaMalloc(&flow, (size_t)N * 2 * HW * 4); cudaMalloc(&mask, (size_t)N * HW); cudaMalloc(&out, 4 * N);
    cudaMemset(flow, 0, (size_t)N * 2 * HW * 4); cudaMemset(mask, 1, (size_t)N * HW); cudaMemset(out, 0, 4 * N);
    const double gb_flow = (double)N * 2 * HW * 4 / 1e9, gb_mask = (double)N * HW / 1e9;
#define RUN(MODE, ITERS, UNR)                                                                                      \
    {                                                                                                              \
        const int span = (MODE == 0) ? 2 * HW : HW;                                                                \
        dim3 grid((span + ITERS * 1024 - 1) / (ITERS * 1024), N);                                                  \
        float ms = timeit([&] { k_read<MODE, ITERS, UNR><<<grid, 256>>>(flow, mask, out, HW); });                  \
        double gb = gb_flow + (MODE == 2 ? gb_mask : 0);                                                           \
        printf("mode %d iters %2d unroll %d grid %5d x %d: %.3f ms  %.0f GB/s\n", MODE, ITERS, UNR, grid.x, N, ms, \
               gb / ms * 1e3);                                                                                     \
    }
    RUN(0, 8, 4) RUN(1, 8, 4) RUN(2, 8, 4)
    RUN(0, 8, 8) RUN(1, 8, 8) RUN(2, 8, 8)
    RUN(0, 4, 4) RUN(1, 4, 4) RUN(2, 4, 4)
    RUN(0, 16, 4) RUN(1, 16, 4) RUN(2, 16, 4)
    RUN(0, 2, 2) RUN(2, 2, 2) RUN(0, 1, 1) RUN(2, 1, 1)
    const size_t n4 = (size_t)N * 2 * HW / 4;
    for (int g : {148, 148 * 2, 148 * 4, 148 * 8}) {
        float ms = timeit([&] { k_flat<4><<<g, 512>>>(reinterpret_cast<const float4*>(flow), n4, out); });
        printf("flat grid-stride grid %4d x 512 unroll 4: %.3f ms  %.0f GB/s\n", g, ms, gb_flow / ms * 1e3);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
