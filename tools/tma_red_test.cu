// Minimal TMA tensor reduce-add probe (debugging aid): smem box -> global += box, clipped at the tensor edges.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../oflibpytorch_b200/csrc/tma.cuh"
using namespace ofb;

template <int BW, int BH, int BC>
__global__ void k(const __grid_constant__ CUtensorMap map, int ox, int oy, int oz, int smem_off)
{
    extern __shared__ __align__(1024) unsigned char raw[];
    float* t = (float*)(raw + smem_off);
    for (int i = threadIdx.x; i < BW * BH * BC; i += blockDim.x) t[i] = 1.0f + (float)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                     ::"l"(reinterpret_cast<uint64_t>(&map)), "r"(ox), "r"(oy), "r"(oz), "r"(smem_u32(t)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
}

template <int BW, int BH, int BC>
int run(int W, int H, int P, int ox, int oy, int oz, int smem_off)
{
    std::vector<float> h((size_t)W * H * P, 100.0f);
    float* d;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap m;
    bool ok = make_map_3d(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, W, H, P, BW, BH, BC);
    printf("box %dx%dx%d at (%d,%d,%d) smem_off %d: encode %d\n", BW, BH, BC, ox, oy, oz, smem_off, (int)ok);
    if (!ok) return 1;
    int smem = BW * BH * BC * 4 + smem_off;
    cudaFuncSetAttribute(k<BW, BH, BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<BW, BH, BC><<<1, 32, smem>>>(m, ox, oy, oz, smem_off);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<float> r(h.size());
    cudaMemcpy(r.data(), d, r.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < P; ++c) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
        float expv = 100.0f;
        int bx = x - ox, by = y - oy, bc = c - oz;
        if (bx >= 0 && bx < BW && by >= 0 && by < BH && bc >= 0 && bc < BC) expv += 1.0f + (float)((bc * BH + by) * BW + bx);
        if (r[((size_t)c * H + y) * W + x] != expv) ++bad;
    }
    printf("  mismatches %d\n", bad);
    cudaFree(d);
    return 0;
}
int main(int argc, char** argv)
{
    int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0) return run<64, 2, 3>(128, 64, 6, 8, 4, 3, 0);
    if (which == 1) return run<64, 2, 3>(128, 64, 6, -4, -1, 0, 0);
    if (which == 2) return run<64, 2, 3>(128, 64, 6, 100, 63, 4, 0);
    if (which == 3) return run<64, 2, 3>(128, 64, 6, 8, 4, 3, 1536);
    if (which == 4) return run<64, 2, 3>(128, 64, 6, 6, 4, 3, 0);
    if (which == 5) return run<32, 2, 3>(128, 64, 6, 8, 4, 3, 0);
    if (which == 6) return run<64, 2, 1>(128, 64, 6, 8, 4, 3, 0);
    if (which == 7) return run<64, 2, 3>(1920, 1080, 6, 8, 4, 3, 13824);
    return 0;
}
