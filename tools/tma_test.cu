// Minimal TMA box-load probe (debugging aid).
#include <cstdio>
#include <vector>
#include "../oflibpytorch_b200/csrc/tma.cuh"
using namespace ofb;

template <int BW, int BH, int BC>
__global__ void k(const __grid_constant__ CUtensorMap map, int ox, int oy, int oz, float* out)
{
    extern __shared__ unsigned char raw[];
    unsigned char* sm = (unsigned char*)(((uintptr_t)raw + 127) & ~(uintptr_t)127);
    __shared__ unsigned long long bar;
    if (threadIdx.x == 0) mbar_init((uint64_t*)&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx((uint64_t*)&bar, BW * BH * BC * 4);
        tma_load_3d(sm, &map, ox, oy, oz, (uint64_t*)&bar);
    }
    __syncthreads();
    mbar_wait((uint64_t*)&bar, 0);
    float* t = (float*)sm;
    for (int i = threadIdx.x; i < BW * BH * BC; i += blockDim.x) out[i] = t[i];
}

template <int BW, int BH, int BC>
int run(int W, int H, int P, int ox, int oy)
{
    std::vector<float> h((size_t)W * H * P);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, BW * BH * BC * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap m;
    bool ok = make_map_3d(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, W, H, P, BW, BH, BC);
    printf("box %dx%dx%d encode %d\n", BW, BH, BC, (int)ok);
    if (!ok) return 1;
    int smem = BW * BH * BC * 4 + 128;
    cudaFuncSetAttribute(k<BW, BH, BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<BW, BH, BC><<<1, 128, smem>>>(m, ox, oy, 0, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<float> r(BW * BH * BC);
    cudaMemcpy(r.data(), o, r.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < BC; ++c) for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
        int gx = ox + x, gy = oy + y;
        float expv = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? (float)((size_t)c * W * H + (size_t)gy * W + gx) : 0.f;
        if (r[(c * BH + y) * BW + x] != expv) ++bad;
    }
    printf("  mismatches %d\n", bad);
    return 0;
}
int main(int argc, char** argv)
{
    int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0) return run<64, 8, 1>(128, 64, 3, -4, -2);
    if (which == 1) return run<80, 48, 1>(128, 64, 3, -4, 5);
    if (which == 2) return run<80, 48, 3>(128, 64, 3, 8, 30);
    if (which == 3) return run<80, 48, 3>(128, 64, 3, 6, 30);
    if (which == 4) return run<64, 8, 1>(128, 64, 3, 2, 1);
    return 0;
}
