// Row-band sharding of one oversized field: the halo step of the splat and of the W adjoint.
//
// Every rank accumulates contributions that land in rows owned by its neighbours into halo rows of its own buffer.
// The buffers are peer-mapped (torch symmetric memory: every GPU of the NVSwitch domain maps every other GPU's
// buffer), so the owner of the rows simply READS its neighbours' halo rows over NVLink and adds them to its own --
// no staging copy, no send/recv.  The order is fixed by the caller (own + upper neighbour, then + lower neighbour:
// two launches on one stream), so the sums do not depend on arrival order.
#include "common.cuh"

namespace ofb {

// dst[n, r, :] += src[n, r, :] for r < rows; rows are row_floats floats wide (a multiple of 4 when VEC)
template <bool VEC>
__global__ void __launch_bounds__(256) add_rows_kernel(float *__restrict__ dst, int64_t dst_item, const float *__restrict__ src,
                                                       int64_t src_item, int N, int64_t per_item)
{
    const int64_t total = (int64_t)N * per_item;
    const int64_t step = (int64_t)gridDim.x * blockDim.x * (VEC ? 4 : 1);
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * (VEC ? 4 : 1); i < total; i += step) {
        const int64_t n = i / per_item, k = i - n * per_item;
        float *d = dst + n * dst_item + k;
        const float *s = src + n * src_item + k;
        if (VEC) {
            const float4 a = *reinterpret_cast<const float4 *>(s);      // peer memory: a P2P load over NVLink
            float4 b = *reinterpret_cast<float4 *>(d);
            b.x += a.x; b.y += a.y; b.z += a.z; b.w += a.w;
            *reinterpret_cast<float4 *>(d) = b;
        } else {
            *d += *s;
        }
    }
}

}  // namespace ofb

using namespace ofb;

extern "C" int ofb_add_rows(float *dst, int64_t dst_item_stride, const float *src, int64_t src_item_stride,
                            int N, int64_t floats_per_item, void *stream)
{
    if (!dst || !src || N <= 0 || floats_per_item <= 0) return OFB_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = aligned16(dst) && aligned16(src) && (dst_item_stride % 4 == 0) && (src_item_stride % 4 == 0) &&
                     (floats_per_item % 4 == 0);
    const int64_t work = (int64_t)N * floats_per_item / (vec ? 4 : 1);
    int blocks = (int)((work + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    if (vec) add_rows_kernel<true><<<blocks, 256, 0, st>>>(dst, dst_item_stride, src, src_item_stride, N, floats_per_item);
    else add_rows_kernel<false><<<blocks, 256, 0, st>>>(dst, dst_item_stride, src, src_item_stride, N, floats_per_item);
    OFB_CHECK_LAUNCH();
    return 0;
}
