// Modulated conv on the 5th-generation tensor cores: implicit GEMM with tcgen05.mma, accumulators in TMEM, operands
// staged by TMA.  sm_100a only.
//
//   M = 128 output pixels (a tw x th x tn patch of the NHWC activation tensor),  N = BN output channels,
//   K = taps x Cin, walked as (tap, 64-channel chunk).
//
// * A operand: the activation already multiplied by the per-sample styles and transposed to NHWC fp16 by a pre-pass
//   (nhwc_prepass_kernel).  One TMA box [64 ch, tw, th, tn] per (tap, chunk); the tap shift is a coordinate offset and
//   the zero padding halo is TMA out-of-bounds fill -- no im2col buffer, no per-sample weights.
// * B operand: the shared weights re-laid out as [tap][Cout][Cin] fp16 (K-major), one TMA box [64, BN, 1].
// * both land in shared memory in the canonical 128-byte-swizzled K-major layout that tcgen05 smem descriptors read.
// * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-5 = epilogue
//   (tcgen05.ld 32 lanes x 16 columns, demodulation d[n,o] (* pre-normalisation a[o]) and noise applied in fp32, NCHW
//   store).  3-stage full/empty mbarrier ring; tcgen05.commit releases smem stages and publishes the accumulator.
// * the data gradient is the same kernel (activation = d*dy, weights transposed/flipped) with an epilogue that scales by
//   the styles and reduces sum_p x*dxpre into dstyles with warp shuffles.
//
// Numerics: fp16 operands, fp32 accumulation, fp32 epilogue -- the fp16 pre-normalisation of the reference
// (networks/generator.py:66-68) is only needed for the activation operand here, because nothing is ever accumulated or
// rescaled in fp16.
#include "modconv_common.cuh"
#include <cuda.h>

namespace vfm {
namespace modconv {

namespace {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (an error the host sees), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], fp16 x fp16 -> fp32, single CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor for a K-major tile whose rows are 128 bytes (64 fp16) with the 128-byte swizzle:
// 8-row groups are 1024 bytes apart (SBO); LBO is unused for swizzled K-major layouts (set to 1 like CUTLASS does);
// bits 46-47 = descriptor version 1 (sm_100); bits 61-63 = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)1 << 16;                                // leading byte offset (ignored)
    d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset
    d |= (uint64_t)1 << 46;                                // version
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}

// Instruction descriptor, kind::f16: D = fp32 (bits 4-5 = 1), A = B = fp16 (0), both K-major (bits 15,16 = 0),
// N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int BM = 128, BK = 64, STAGES = 3;
constexpr int A_BYTES = BM * BK * 2;   // 16 KB

struct TcArgs {
    int ntaps, kchunks;
    int tap_dy[9], tap_dx[9], tap_b[9];   // A coordinate offsets and the tap's slice index in the B tensor map
    int tw, th, tn;                        // M tile = tw x th pixels of tn consecutive samples (tw*th*tn == 128)
    int tiles_w, tiles_h;                  // tiles per image
    int N, H, W, Nout;                     // output geometry: [N, Nout, H, W] NCHW
    void* out;
    const float* oscale;                   // [N, Nout]
    const float* add;                      // noise or NULL
    int64_t add_sn;
    const void* aux;                       // dgrad: x, same shape/dtype as out
    float* aux_sum;                        // dgrad: [N, Nout] += sum_p aux * acc
};

template <int BN, class TOut, bool DGRAD>
__global__ void __launch_bounds__(192) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcArgs p) {
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B tiles need 1024-byte alignment
    uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);
    float* s_scale = (float*)(tmem_slot + 2);    // [tn][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates
    int mt = blockIdx.x;
    const int tile_w = mt % p.tiles_w; mt /= p.tiles_w;
    const int tile_h = mt % p.tiles_h; mt /= p.tiles_h;
    const int n0 = mt * p.tn, h0 = tile_h * p.th, w0 = tile_w * p.tw;
    const int o0 = blockIdx.y * BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN);      // BN fp32 accumulator columns (power of two >= 32)
        tmem_relinquish();
    }
    for (int i = threadIdx.x; i < p.tn * BN; i += blockDim.x) {
        int nl = i / BN, c = i - nl * BN;
        int n = n0 + nl;
        s_scale[i] = (n < p.N) ? p.oscale[(size_t)n * p.Nout + o0 + c] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int iters = p.ntaps * p.kchunks;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int it = 0; it < iters; it++) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
                uint8_t* sa = smem + s * STAGE_BYTES;
                tma_load_4d(sa, &tmA, &full_bar[s], kc * BK, w0 + p.tap_dx[tap], h0 + p.tap_dy[tap], n0);
                tma_load_3d(sa + A_BYTES, &tmB, &full_bar[s], kc * BK, o0, p.tap_b[tap]);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(BM, BN);
            for (int it = 0; it < iters; it++) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t adesc = make_kmajor_sw128_desc(sa);
                const uint64_t bdesc = make_kmajor_sw128_desc(sa + A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; k++) {
                    // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle span: +2 in the (>>4) address field
                    umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);     // frees this smem stage once the MMAs above have read it
            }
            umma_commit(accum_bar);             // accumulator complete
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps (2..5)
        const int lg = warp & 3;                           // TMEM lane group this warp may access
        const int r = lg * 32 + lane;                      // accumulator row = pixel index inside the tile
        const int nl = r / (p.tw * p.th);
        const int hl = (r / p.tw) % p.th, wl = r % p.tw;
        const int n = n0 + nl, h = h0 + hl, w = w0 + wl;
        const bool valid = (n < p.N) && (h < p.H) && (w < p.W);
        const size_t HW = (size_t)p.H * p.W;
        const size_t pix = (size_t)h * p.W + w;
        float addv = 0.f;
        if (!DGRAD && p.add && valid) addv = p.add[(size_t)n * p.add_sn + pix];
        TOut* outp = (TOut*)p.out + ((size_t)n * p.Nout + o0) * HW + pix;
        const TOut* auxp = DGRAD ? (const TOut*)p.aux + ((size_t)n * p.Nout + o0) * HW + pix : nullptr;
        const float* sc = s_scale + nl * BN;

        mbar_wait(accum_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < BN / 16; j++) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(j * 16), v);
#pragma unroll
            for (int c = 0; c < 16; c++) {
                const int col = j * 16 + c;
                float acc = v[c];
                if (DGRAD) {
                    float part = 0.f;
                    if (valid && p.aux_sum) part = to_acc(auxp[(size_t)col * HW]) * acc;
                    if (p.aux_sum) {
                        part = warp_sum(part);     // all 32 lanes of a warp belong to one sample (tw*th >= 32)
                        if (lane == 0 && n < p.N) atomicAdd(&p.aux_sum[(size_t)n * p.Nout + o0 + col], part);
                    }
                }
                float val = acc * sc[col] + addv;
                if (valid) outp[(size_t)col * HW] = from_acc<TOut, float>(val);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

// ------------------------------------------------------------------------------------------------ pre-pass kernels
// x [N,C,H,W] (TIn) * scale[n,c]  ->  xt [N,H,W,C] fp16.  64 channels x 64 pixels per CTA through shared memory:
// coalesced reads along pixels, coalesced 16-byte writes along channels.
template <class TIn>
__global__ void __launch_bounds__(256) nhwc_prepass_kernel(const TIn* __restrict__ x, const float* __restrict__ scale, __half* __restrict__ xt,
                                                           int C, int HW) {
    __shared__ __half s[64][66];     // [pixel][channel], 33-word pitch: conflict-free transposed stores
    const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
    const int tid = threadIdx.x;
    {
        const int pl = tid & 63, cg = tid >> 6;      // 4 channel groups x 64 pixels
        const int pidx = p0 + pl;
#pragma unroll 4
        for (int i = 0; i < 16; i++) {
            const int cl = cg + i * 4;
            const int c = c0 + cl;
            float v = 0.f;
            if (pidx < HW && c < C) v = to_acc(x[((size_t)n * C + c) * HW + pidx]) * scale[(size_t)n * C + c];
            s[pl][cl] = __float2half_rn(v);
        }
    }
    __syncthreads();
    {
        const int cv = tid & 7, pl0 = tid >> 3;      // 8 threads x 8 channels (16 bytes) per pixel, 32 pixels per pass
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int pl = pl0 + i * 32;
            const int pidx = p0 + pl;
            if (pidx >= HW) continue;
            union { uint4 u; uint32_t w[4]; } pk;
            const uint32_t* src = (const uint32_t*)&s[pl][cv * 8];
#pragma unroll
            for (int k = 0; k < 4; k++) pk.w[k] = src[k];
            if (c0 + cv * 8 + 8 <= C) *(uint4*)(xt + ((size_t)n * HW + pidx) * C + c0 + cv * 8) = pk.u;
        }
    }
}

// weight [O,I,KK] fp32 -> wt[t][O][I] (transpose == 0) or wt[t][I][O] (transpose == 1) fp16, t = position in the tap list
__global__ void weight_prep_kernel(const float* __restrict__ w, __half* __restrict__ wt, int O, int I, int KK, int ntaps, const int* __restrict__ widx_dev,
                                   int transpose, int widx0, int widx1, int widx2, int widx3, int widx4, int widx5, int widx6, int widx7, int widx8) {
    const int widx[9] = {widx0, widx1, widx2, widx3, widx4, widx5, widx6, widx7, widx8};
    size_t total = (size_t)ntaps * O * I;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int t = (int)(idx / ((size_t)O * I));
        size_t r = idx - (size_t)t * O * I;
        int o, i;
        if (!transpose) { o = (int)(r / I); i = (int)(r - (size_t)o * I); }
        else { i = (int)(r / O); o = (int)(r - (size_t)i * O); }
        wt[idx] = __float2half_rn(w[((size_t)o * I + i) * KK + widx[t]]);
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

int encode_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint32_t* box) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VFM_ERR_CUDA; }
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bdim[5], estr[5];
    uint64_t stride = 2;   // fp16
    for (int i = 0; i < rank; i++) {
        gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1;
        stride *= dims[i];
        if (i < rank - 1) gstride[i] = stride;
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, base, gdim, gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with code %d", (int)r); return VFM_ERR_CUDA; }
    return VFM_OK;
}

bool pick_tile(int N, int H, int W, int& tw, int& th, int& tn) {
    if (W >= 32) { tw = 32; th = 4; tn = 1; }
    else if (W == 16) { tw = 16; th = 8; tn = 1; }
    else if (W == 8) { tw = 8; th = 8; tn = 2; }
    else return false;
    return (W % tw == 0) && (H % th == 0) && (N % tn == 0);
}

size_t smem_bytes(int BN) { return (size_t)STAGES * (A_BYTES + BN * BK * 2) + 1024 + 256 + (size_t)2 * BN * sizeof(float); }

template <int BN, class TOut, bool DGRAD>
int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, int m_tiles, int n_tiles, double flops, cudaStream_t stream) {
    auto kern = conv_tc_kernel<BN, TOut, DGRAD>;
    size_t smem = smem_bytes(BN);
    VFM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KernelTimer timer(DGRAD ? "modconv_tc_dgrad" : "modconv_tc_fwd", stream, flops, 0.0);
    kern<<<dim3(m_tiles, n_tiles), 192, smem, stream>>>(tmA, tmB, a);
    return launch_status("modconv conv_tc_kernel");
}

// One implicit-GEMM conv over an NHWC fp16 activation `act` [N,H,W,Cin] with weights `wt` [ntaps][Nout][Cin]
int run_tc_conv(int out_dtype, bool dgrad, __half* act, __half* wt, const TapTable& taps, int N, int H, int W, int Cin, int Nout,
                void* out, const float* oscale, const float* add, int64_t add_sn, const void* aux, float* aux_sum, cudaStream_t stream) {
    TcArgs a;
    a.ntaps = taps.ntaps; a.kchunks = Cin / BK;
    for (int t = 0; t < taps.ntaps; t++) { a.tap_dy[t] = taps.off_y[t]; a.tap_dx[t] = taps.off_x[t]; a.tap_b[t] = t; }
    if (!pick_tile(N, H, W, a.tw, a.th, a.tn)) { set_error("tcgen05 path: unsupported image size %dx%d", H, W); return VFM_ERR_NO_KERNEL; }
    a.tiles_w = W / a.tw; a.tiles_h = H / a.th;
    a.N = N; a.H = H; a.W = W; a.Nout = Nout;
    a.out = out; a.oscale = oscale; a.add = add; a.add_sn = add_sn; a.aux = aux; a.aux_sum = aux_sum;
    const int BN = 128;
    CUtensorMap tmA, tmB;
    uint64_t adims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint32_t abox[4] = {(uint32_t)BK, (uint32_t)a.tw, (uint32_t)a.th, (uint32_t)a.tn};
    int st = encode_map(&tmA, act, 4, adims, abox); if (st) return st;
    uint64_t bdims[3] = {(uint64_t)Cin, (uint64_t)Nout, (uint64_t)taps.ntaps};
    uint32_t bbox[3] = {(uint32_t)BK, (uint32_t)BN, 1u};
    st = encode_map(&tmB, wt, 3, bdims, bbox); if (st) return st;
    const int m_tiles = a.tiles_w * a.tiles_h * (N / a.tn), n_tiles = Nout / BN;
    const double flops = 2.0 * N * H * W * (double)Nout * Cin * taps.ntaps;
    if (out_dtype == VFM_F16) {
        return dgrad ? launch_tc<128, __half, true>(tmA, tmB, a, m_tiles, n_tiles, flops, stream)
                     : launch_tc<128, __half, false>(tmA, tmB, a, m_tiles, n_tiles, flops, stream);
    }
    return dgrad ? launch_tc<128, float, true>(tmA, tmB, a, m_tiles, n_tiles, flops, stream)
                 : launch_tc<128, float, false>(tmA, tmB, a, m_tiles, n_tiles, flops, stream);
}

template <class TIn>
int run_prepass(const void* x, const float* scale, __half* xt, int N, int C, int HW, cudaStream_t stream) {
    dim3 grid(ceil_div(HW, 64), ceil_div(C, 64), N);
    KernelTimer timer("modconv_nhwc_prepass", stream, 0.0, (double)N * C * HW * (sizeof(TIn) + 2));
    nhwc_prepass_kernel<TIn><<<grid, 256, 0, stream>>>((const TIn*)x, scale, xt, C, HW);
    return launch_status("modconv nhwc_prepass_kernel");
}

int run_weight_prep(const float* w, __half* wt, int O, int I, int KK, const TapTable& taps, int transpose, cudaStream_t stream) {
    int wi[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < taps.ntaps; t++) wi[t] = taps.widx[t];
    size_t total = (size_t)taps.ntaps * O * I;
    int blocks = (int)((total + 255) / 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    weight_prep_kernel<<<blocks, 256, 0, stream>>>(w, wt, O, I, KK, taps.ntaps, nullptr, transpose, wi[0], wi[1], wi[2], wi[3], wi[4], wi[5], wi[6], wi[7], wi[8]);
    return launch_status("modconv weight_prep_kernel");
}

void fwd_taps(const vfm_modconv_desc& d, TapTable& t) {
    t.ntaps = d.kh * d.kw;
    for (int ky = 0; ky < d.kh; ky++)
        for (int kx = 0; kx < d.kw; kx++) {
            int i = ky * d.kw + kx;
            t.off_y[i] = ky - d.padding; t.off_x[i] = kx - d.padding;
            t.widx[i] = d.flip_weight ? i : (d.kh - 1 - ky) * d.kw + (d.kw - 1 - kx);
        }
}

}  // namespace

bool tc_supported(const vfm_modconv_desc& d) {
    if (d.dtype != VFM_F16) return false;
    if (d.up != 1 || d.kh != d.kw || (d.kh != 3 && d.kh != 1) || d.padding != d.kh / 2) return false;
    if (d.in_channels % 128 != 0 || d.out_channels % 128 != 0) return false;   // both are an N dimension (fwd / dgrad) and a K dimension
    int tw, th, tn;
    if (!pick_tile(d.batch, d.in_h, d.in_w, tw, th, tn)) return false;
    return get_encode_fn() != nullptr;
}

size_t tc_workspace_bytes(const vfm_modconv_desc& d, int direction) {
    Carver cv(nullptr, ~(size_t)0);
    const size_t npix = (size_t)d.batch * d.in_h * d.in_w;
    const size_t wel = (size_t)d.kh * d.kw * d.out_channels * d.in_channels;
    if (direction == 0) {
        cv.take<__half>(npix * d.in_channels);
        cv.take<__half>(wel);
    } else {
        cv.take<__half>(npix * d.out_channels);   // d*dy, NHWC
        cv.take<__half>(wel);                     // transposed weights
    }
    return cv.off + 512;
}

int tc_forward(const vfm_modconv_fwd_params& p, const Coefs& k, void* ws, size_t ws_bytes, cudaStream_t stream) {
    const vfm_modconv_desc& d = p.d;
    Carver cv(ws, ws_bytes);
    const size_t npix = (size_t)d.batch * d.in_h * d.in_w;
    __half* xt = cv.take<__half>(npix * d.in_channels);
    __half* wt = cv.take<__half>((size_t)d.kh * d.kw * d.out_channels * d.in_channels);
    if (!cv.ok()) { set_error("modulated_conv2d: tcgen05 workspace too small"); return VFM_ERR_WORKSPACE; }
    TapTable taps; fwd_taps(d, taps);
    int st = run_prepass<__half>(p.x, k.iscale, xt, d.batch, d.in_channels, d.in_h * d.in_w, stream); if (st) return st;
    st = run_weight_prep(p.weight, wt, d.out_channels, d.in_channels, d.kh * d.kw, taps, 0, stream); if (st) return st;
    const int64_t noise_sn = (d.noise_mode == VFM_NOISE_N1HW) ? (int64_t)d.out_h * d.out_w : 0;
    return run_tc_conv(d.dtype, false, xt, wt, taps, d.batch, d.in_h, d.in_w, d.in_channels, d.out_channels, p.y, k.oscale, p.noise, noise_sn,
                       nullptr, nullptr, stream);
}

int run_wgrad(int dtype, WgradArgs a, cudaStream_t stream);   // generic SIMT wgrad (modconv_generic.cu) until the tcgen05 wgrad lands

int tc_backward(const vfm_modconv_bwd_params& p, const Coefs& k, float* g, float* dsum, void* ws, size_t ws_bytes, cudaStream_t stream) {
    const vfm_modconv_desc& d = p.d;
    const int N = d.batch, I = d.in_channels, O = d.out_channels, KK = d.kh * d.kw;
    Carver cv(ws, ws_bytes);
    const size_t npix = (size_t)N * d.in_h * d.in_w;
    __half* dyt = cv.take<__half>(npix * O);
    __half* wtT = cv.take<__half>((size_t)KK * O * I);
    if (!cv.ok()) { set_error("modulated_conv2d backward: tcgen05 workspace too small"); return VFM_ERR_WORKSPACE; }
    TapTable ftaps; fwd_taps(d, ftaps);
    int st;
    if (p.dx) {
        // dxpre[n,i,p] = sum_{o,t} W[o,i,widx(t)] * (d*a*dy)[n,o,p - off(t)]
        TapTable dt = ftaps;
        for (int t = 0; t < dt.ntaps; t++) { dt.off_y[t] = -ftaps.off_y[t]; dt.off_x[t] = -ftaps.off_x[t]; }
        st = run_prepass<__half>(p.dy, k.oscale, dyt, N, O, d.out_h * d.out_w, stream); if (st) return st;
        st = run_weight_prep(p.weight, wtT, O, I, KK, dt, 1, stream); if (st) return st;
        if (p.dstyles) VFM_CUDA_OK(cudaMemsetAsync(dsum, 0, sizeof(float) * (size_t)N * I, stream));
        st = run_tc_conv(d.dtype, true, dyt, wtT, dt, N, d.in_h, d.in_w, O, I, p.dx, k.iscale, nullptr, 0, p.dstyles ? p.x : nullptr,
                         p.dstyles ? dsum : nullptr, stream);
        if (st) return st;
    }
    if (p.dweight) {
        VFM_CUDA_OK(cudaMemsetAsync(p.dweight, 0, sizeof(float) * (size_t)O * I * KK, stream));
        WgradArgs w;
        w.dy = p.dy; w.x = p.x; w.oscale = k.oscale; w.iscale = k.iscale; w.dw = p.dweight;
        w.s_co = (int64_t)I * KK; w.s_ci = KK;
        w.N = N; w.Co = O; w.Ci = I; w.Hd = d.out_h; w.Wd = d.out_w; w.Hx = d.in_h; w.Wx = d.in_w;
        w.sn = 1; w.sd = 1; w.taps = ftaps; w.chunks = 0; w.chunk_pix = 0;
        st = run_wgrad(d.dtype, w, stream); if (st) return st;
    }
    (void)g;
    return VFM_OK;
}

}  // namespace modconv
}  // namespace vfm
