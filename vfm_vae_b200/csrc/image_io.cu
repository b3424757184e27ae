// Decode I/O path: the pixel conversion between the decoder's output and the PNG encoder of the reference's decode /
// reconstruct tools,
//     images = ((images + 1) / 2).clamp(0, 1)                      tools/decode/decode_latents_to_images.py:92
//     to_pil_image(tensor.clamp(0, 1))  ==  tensor.mul(255).byte() -> HWC     (safe_save, :20-24; torchvision's float path truncates)
// as ONE pass: NCHW fp32 / fp16 in [-1, 1]  ->  NHWC uint8, bit-identical to the reference's fp32 arithmetic (add, exact halving, clamp,
// one rounded multiply by 255, truncation).  HBM-bound byte work: a thread converts 4 consecutive pixels of a row -- one 16-byte (fp32) /
// 8-byte (fp16) load per channel plane, one contiguous 4*C-byte store -- so a warp reads C x 512 contiguous bytes and writes 128*C.
#include "common.cuh"

namespace vfm {
namespace {

__device__ __forceinline__ uint32_t to_u8(float x, float pre_add, float pre_div, float scale) {
    float t = __fdiv_rn(__fadd_rn(x, pre_add), pre_div);
    t = fminf(fmaxf(t, 0.f), 1.f);               // NaN -> 0 like the byte cast of clamp(NaN) is unspecified in the reference; pinned to 0 here
    if (!(t == t)) t = 0.f;
    return (uint32_t)(int)__fmul_rn(t, scale);   // truncation, as Tensor.byte()
}

template <class T, int C>
__global__ void __launch_bounds__(256) image_to_u8_kernel(const T* __restrict__ x, uint8_t* __restrict__ y, int64_t n_quads, int HW, int W,
                                                          float pre_add, float pre_div, float scale, int vec_ok) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = q * 4;                 // first of 4 consecutive pixels of one image (HW % 4 == 0 when vec_ok)
        const int64_t n = pix / HW;
        const int64_t p = pix - n * HW;
        float v[C][4];
#pragma unroll
        for (int c = 0; c < C; c++) {
            const T* src = x + (n * C + c) * HW + p;
            if (vec_ok) {
                if (sizeof(T) == 4) { const float4 t = *(const float4*)src; v[c][0] = t.x; v[c][1] = t.y; v[c][2] = t.z; v[c][3] = t.w; }
                else { const uint2 t = *(const uint2*)src; const float2 a = __half22float2(*(const __half2*)&t.x), b = __half22float2(*(const __half2*)&t.y);
                       v[c][0] = a.x; v[c][1] = a.y; v[c][2] = b.x; v[c][3] = b.y; }
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) v[c][i] = (p + i < HW) ? to_acc(src[i]) : 0.f;
            }
        }
        uint8_t out[4 * C];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int c = 0; c < C; c++) out[i * C + c] = (uint8_t)to_u8(v[c][i], pre_add, pre_div, scale);
        uint8_t* dst = y + (n * HW + p) * C;
        if (vec_ok) {
#pragma unroll
            for (int w = 0; w < C; w++) ((uint32_t*)dst)[w] = *(const uint32_t*)&out[4 * w];
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (p + i < HW) {
#pragma unroll
                    for (int c = 0; c < C; c++) dst[i * C + c] = out[i * C + c];
                }
        }
    }
}

template <class T, int C>
int launch_image_to_u8(const vfm_image_to_u8_params* p, cudaStream_t stream) {
    const int64_t HW = (int64_t)p->height * p->width;
    const int vec_ok = (HW % 4 == 0) && aligned16(p->x) && ((reinterpret_cast<uintptr_t>(p->y) & 3u) == 0);
    const int64_t quads = vec_ok ? (int64_t)p->batch * HW / 4 : (int64_t)p->batch * ceil_div64(HW, 4);
    if (!vec_ok && HW % 4 != 0 && p->batch > 1) {
        // ragged planes: quads must not straddle images -> one launch per image keeps the kernel simple (not a decoder shape)
        for (int n = 0; n < p->batch; n++) {
            const int64_t w1 = ceil_div64(ceil_div64(HW, 4), 256);
            image_to_u8_kernel<T, C><<<(unsigned)(w1 < (int64_t)kNumSMs * 8 ? w1 : (int64_t)kNumSMs * 8), 256, 0, stream>>>(
                (const T*)p->x + (int64_t)n * C * HW, p->y + (int64_t)n * HW * C, ceil_div64(HW, 4), (int)HW, p->width, (float)p->pre_add, (float)p->pre_div,
                (float)p->scale, 0);
            int st = launch_status("image_to_u8_kernel"); if (st) return st;
        }
        return VFM_OK;
    }
    const int64_t want = ceil_div64(quads, 256);
    const unsigned grid = (unsigned)(want < (int64_t)kNumSMs * 8 ? (want > 0 ? want : 1) : (int64_t)kNumSMs * 8);
    KernelTimer timer("image_to_u8", stream, 0.0, (double)p->batch * C * HW * (sizeof(T) + 1.0));
    image_to_u8_kernel<T, C><<<grid, 256, 0, stream>>>((const T*)p->x, p->y, quads, (int)HW, p->width, (float)p->pre_add, (float)p->pre_div, (float)p->scale, vec_ok);
    return launch_status("image_to_u8_kernel");
}

}  // namespace
}  // namespace vfm

extern "C" int vfm_image_to_u8(const vfm_image_to_u8_params* p, void* stream_) {
    using namespace vfm;
    cudaStream_t stream = (cudaStream_t)stream_;
    VFM_CHECK_ARG(p != nullptr, "image_to_u8: params is NULL");
    VFM_CHECK_ARG(p->x && p->y, "image_to_u8: x and y must be non-NULL");
    VFM_CHECK_ARG(p->batch >= 1 && p->height >= 1 && p->width >= 1, "image_to_u8: empty image");
    VFM_CHECK_ARG((int64_t)p->height * p->width <= 0x7fffffffLL, "image_to_u8: image too large");
    VFM_CHECK_ARG(p->dtype == VFM_F16 || p->dtype == VFM_F32, "image_to_u8: unsupported dtype %d", p->dtype);
    VFM_CHECK_ARG(p->channels == 1 || p->channels == 3 || p->channels == 4, "image_to_u8: channels must be 1, 3 or 4 (got %d)", p->channels);
    VFM_CHECK_ARG(p->pre_div != 0.0, "image_to_u8: pre_div must be non-zero");
#define VFM_IMG_DISPATCH(T)                                                       \
    switch (p->channels) {                                                        \
        case 1: return launch_image_to_u8<T, 1>(p, stream);                       \
        case 3: return launch_image_to_u8<T, 3>(p, stream);                       \
        default: return launch_image_to_u8<T, 4>(p, stream);                      \
    }
    if (p->dtype == VFM_F16) { VFM_IMG_DISPATCH(__half) }
    VFM_IMG_DISPATCH(float)
#undef VFM_IMG_DISPATCH
}
