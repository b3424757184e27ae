// Modulated conv on the 5th-generation tensor cores: implicit GEMM with tcgen05.mma, accumulators in TMEM, operands
// staged by TMA.  sm_100a only.
//
// conv_tc_kernel (forward and data gradient), one persistent CTA of 320 threads per SM:
//   M = 128 output CHANNELS (UMMA A operand = a weight tile),  N = NPIX = 256 PIXELS (UMMA B operand = a tw x th x tn
//   patch of the NHWC activation; 128 pixels for the paired up=2 phases and the fp32 split),  K = taps x Cin, walked
//   as (tap, 64-channel chunk).  The accumulator therefore has one channel per TMEM lane and the pixels along the
//   columns: an epilogue thread owns one output channel and runs of 16 consecutive pixels.
// * A operand: the shared weights re-laid out as [tap][Cout][Cin] fp16 (K-major, weight_prep_kernel), one TMA box
//   [64, 128, 1] per (tap, chunk); the data gradient uses the transposed layout [tap][Cin][Cout].
// * B operand: the activation multiplied by the per-sample styles (and an optional GroupNorm affine map) and transposed
//   to NHWC fp16 by a pre-pass (nhwc_prepass_kernel).  One TMA box [64 ch, tw, th, tn] per (tap, chunk): the tap shift
//   is a coordinate offset, the zero-padding halo is TMA out-of-bounds fill -- no im2col buffer and no per-sample
//   [N,O,I,k,k] weights (the 604 MB tensor of networks/generator.py:73-99).  1x1 convs on fp16 NCHW rows of a multiple
//   of 64 pixels (MNP) skip the pre-pass: MN-major boxes [64 pixels, 64 channels] straight from the NCHW tensor and
//   per-sample weights that carry the modulation.
// * both operands land in shared memory in the canonical 128-byte-swizzled layout that tcgen05 smem descriptors read;
//   a ring of 192 KB / stage bytes stages (4 x 48 KB for fp16 at NPIX = 256) with full/empty mbarriers.
// * warp roles: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator (all 512 columns) + single-thread MMA
//   issuer, warps 2-9 = epilogue (two warps per TMEM lane group, one per half of the pixel columns).  The
//   accumulators are double-buffered in TMEM (tfull/tempty mbarriers): the epilogue of item i runs under the MMAs of
//   item i+1; tcgen05.commit releases smem stages and publishes accumulators.
// * epilogue: tcgen05.ld 32 lanes x 16 columns (next step's load in flight), demodulation d[n,o], noise (staged per
//   tile in shared memory), optional fused bias / lrelu|GELU / gain / clamp / layer-scaled residual (inference),
//   256-bit NCHW stores; side inputs (residual, or x for the dstyles reduction) prefetched in a rolling window.
// * fp16 3x3 stride-1 convs on images that are a multiple of 16 x 16 (ROW3): the three taps of one kernel column share ONE TMA box of (16 + 2)
//   rows x 16 pixels -- a tile row is two 1024-byte swizzle atoms, so the taps are UMMA descriptors 2 KB apart -- and the pixel operand is
//   fetched 3 instead of 9 times per K chunk (two rings: 3 pixel boxes of 36 KB, 5 weight tiles of 16 KB).
// * up=2 (transposed conv, stride 2; PAIR): 4 sub-pixel phases, each a stride-1 conv over the input grid with the taps
//   of matching parity.  An item computes the two horizontal phases of one row parity into two accumulators and the
//   epilogue interleaves them, so stores to the (2H+1) x (2W+1) intermediate are contiguous vectors.
// * the data gradient (DGRAD) is the same kernel (B operand = d*dz, weights transposed) with an epilogue that scales by
//   the styles and accumulates sum_p x*dxpre -> dstyles as a private register sum (one atomic per channel and sample);
//   for up=2 its B boxes are element-strided TMA boxes (stride 2).
// * fp32 tensors (SPLIT) use a 2-term fp16 split of both operands (x = hi + lo/2048): three MMAs per K step
//   (hi*hi -> acc0; hi*lo + lo*hi -> acc1), recombined in the epilogue as acc0 + acc1/2048.  That carries ~22 mantissa
//   bits per operand -- enough for the 1e-5 fp32 parity gate, which single-pass TF32 (10 bits) cannot meet -- at 3x the
//   fp16 cost instead of the 6x of 3xTF32.  Range: the styles are normalised per sample (max |s'| <= 1, a power-of-two
//   c2 undone in the epilogue) and the incoming GRADIENTS are brought into fp16 range by a device-computed power of two
//   (amax_kernel -> gscale); forward ACTIVATIONS are not rescaled (that would cost one more pass over x), so an fp32
//   forward needs |x * s'| < 65504 and loses mantissa below ~1e-4 of that -- true for the decoder's O(1) activations
//   (GroupNorm / lrelu-clamped layers), documented in INTEGRATION.md as the supported range.
//
// wgrad_tc_kernel (weight gradient): M = 128 output channels x N = 256 columns = two (tap, 128 input channel) blocks,
// K = pixels; both operands are MN-major TMA boxes [64 pixels, 64 channels] of the same NHWC tensors; split-K over
// pixel tiles with fp32 red.add into dweight.
//
// Numerics: fp16 operands, fp32 accumulation, fp32 epilogue.  The reference's fp16 pre-normalisation
// (networks/generator.py:66-68) maps to: B = x * s' (styles normalised per sample), A = W * a[o] (weights normalised per
// output channel), epilogue scale = d[n,o].
#include "modconv_common.cuh"
#include <cuda.h>
#include <type_traits>

namespace vfm {
namespace modconv {

int run_wgrad(int dtype, WgradArgs a, cudaStream_t stream);   // generic SIMT wgrad (modconv_generic.cu)

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

// MN-major operand tiles (rows of 64 MN-elements = 128 bytes, K = row index): used by the weight gradient (both operands) and by
// the 1x1 convs that read their pixel operand straight from the NCHW tensor
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;      // leading byte offset: next 64-channel block
    d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset: next group of 8 K rows
    d |= (uint64_t)1 << 46;                                // version
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_f16_mnmajor(int M, int N) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int BK = 64;
constexpr float kLoScale = 2048.f;     // the lo term of the fp16 split is stored scaled by 2^11

// ------------------------------------------------------------------------------------------------ conv / data gradient
// GEMM orientation: M = 128 *channels* (the weight tile is the UMMA A operand), N = NPIX *pixels* (the activation tile
// is the UMMA B operand), so the accumulator in TMEM has one channel per lane and the tile's pixels along the columns.
// An epilogue thread therefore owns one output channel and a run of consecutive pixels: per-channel coefficients
// (demodulation, bias, layer scale) are registers, NCHW stores / residual loads are 16-byte vectors along W, and the
// dstyles reduction of the data gradient is a private register sum -- no shuffles, no shared memory.
constexpr int CH = 128;                // channels per tile (UMMA M)
constexpr int W_BYTES = CH * BK * 2;   // weight tile per stage, 16 KB
constexpr int kConvThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kTmemCols = 512;         // the whole tensor memory: one persistent CTA per SM

struct TcPhase {
    int ntaps;
    int dy[9], dx[9], tb[9];   // pixel-operand coordinate offsets and the tap's slice index in the weight tensor map
    int oy, ox;                // output offset of this phase
    int Hg, Wg;                // extent of the phase grid
};

struct TcArgs {
    TcPhase ph[4];             // PAIR: index 2*py + px
    int kchunks;
    int tw, th, tn;            // pixel tile = tw x th pixels of tn consecutive samples (tw*th*tn == NPIX), all powers of two
    int tw_sh, th_sh;          // log2(tw), log2(th)
    int tiles_w, tiles_h;      // tile grid (covers the largest phase)
    int org_w, org_h;          // origin of the tile grid in phase-grid coordinates (edge strips of the up=2 phases start at the last column / row)
    int N, Nout;
    int out_H, out_W, out_s;   // output tensor [N, Nout, out_H, out_W]; output coordinate = g*out_s + o{y,x}
    int out_pitch;             // row pitch of the output tensor in elements (>= out_W; planes are out_H*out_pitch apart)
    int ngroups;               // phase groups an item belongs to: 1, or 2 (PAIR: py)
    int a_s;                   // pixel coordinate = g*a_s + d{y,x}  (2 for the strided gather of the transposed conv's backward)
    void* out;
    const float* oscale;       // [N, Nout]
    const float* gscale_inv;   // device scalar multiplied into the result, or NULL
    const float* add;          // noise or NULL
    int64_t add_sn;
    const void* aux;           // dgrad: x, same shape/dtype as out
    float* aux_sum;            // dgrad: [N, Nout] += sum_p aux * acc
    Epilogue ep;               // forward: optional fused bias/activation/residual
    int vec_out, vec_side, vec_add;   // 16-byte vector access allowed on out / (aux | residual) / noise
    const float* bias_nc;             // forward: optional per-(sample, channel) bias added before the activation (folded input shift)
    // ROW3 kernels: per phase group (PAIR: py; otherwise only [0]) the taps sorted into COLUMN groups -- taps of one pixel offset dx share ONE
    // pixel box that starts at row h0 + dy0 and is 2 rows taller than the tile; tap m of a group accumulates into phase sub[m], reads the box
    // from image row row[m] (0..2) on, and multiplies weight slice tb[m]
    struct ColGroup { int dx, nm; int sub[4], row[4], tb[4]; };
    int cg_n[2], cg_dy0[2];
    ColGroup cg[2][3];
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 256-bit global accesses (sm_100): one full 32-byte sector per lane
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
    asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// 16 consecutive elements: raw 32-bit words (8 for fp16, 16 for fp32) -> floats
template <class T> struct Raw16 { static constexpr int W = (sizeof(T) == 2) ? 8 : 16; };
template <class T> __device__ __forceinline__ void raw16_load(const T* p, uint32_t* r) {
    ldg256(p, r);
    if (sizeof(T) == 4) ldg256((const char*)p + 32, r + 8);
}
template <class T> __device__ __forceinline__ void raw16_unpack(const uint32_t* r, float* v);
template <> __device__ __forceinline__ void raw16_unpack<__half>(const uint32_t* r, float* v) {
#pragma unroll
    for (int i = 0; i < 8; i++) { const float2 f = __half22float2(*(const __half2*)&r[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void raw16_unpack<float>(const uint32_t* r, float* v) {
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
// n = number of valid elements; vec = all 16 valid and 32-byte aligned
template <class T> __device__ __forceinline__ void load16(const T* p, float* v, int n, bool vec) {
    if (vec && n == 16) {
        uint32_t r[Raw16<T>::W];
        raw16_load<T>(p, r);
        raw16_unpack<T>(r, v);
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = (i < n) ? to_acc(p[i]) : 0.f;
    }
}
template <class T> __device__ __forceinline__ void store16(T* p, const float* v, int n, bool vec);
template <> __device__ __forceinline__ void store16<__half>(__half* p, const float* v, int n, bool vec) {
    if (vec && n == 16) {
        uint32_t r[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]); r[i] = *(const uint32_t*)&h; }
        stg256(p, r);
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) if (i < n) p[i] = __float2half_rn(v[i]);
    }
}
template <> __device__ __forceinline__ void store16<float>(float* p, const float* v, int n, bool vec) {
    if (vec && n == 16) {
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; i++) r[i] = __float_as_uint(v[i]);
        stg256(p, r);
        stg256(p + 8, r + 8);
    } else {
#pragma unroll
        for (int i = 0; i < 16; i++) if (i < n) p[i] = v[i];
    }
}

// Persistent kernel: one CTA per SM walks over work items (phase group fastest, then channel tile, then pixel tile, so
// CTAs that share a pixel tile run at the same time).  The shared-memory ring (TMA -> MMA) keeps running across items;
// the accumulators are double-buffered in TMEM (tfull / tempty barriers), so the epilogue of item i overlaps the MMAs of
// item i+1 and the per-CTA setup (barrier init, TMEM allocation, descriptor prefetch) is paid once per SM.
//
// PAIR (transposed conv, stride 2): an item computes the two horizontal sub-pixel phases (px = 0, 1) of one row parity
// into two accumulators and the epilogue interleaves them, so the stores to the 2x-resolution output are contiguous.
// SPLIT (fp32 tensors): two accumulators per phase (hi*hi and the cross terms), recombined in the epilogue.
// MNP (1x1 convs, fp16): the pixel operand is read straight from the NCHW tensor as MN-major tiles (one TMA box [64 pixels of a
// row, 64 channels] per image row of the tile) and the weight operand is a per-sample weight [n][Cout][Cin] that carries the
// modulation (and a folded input affine map) -- no pre-pass over the activations at all.  Only possible without tap shifts: TMA
// needs the innermost box start 16-byte aligned, which a +-1 pixel shift of an NCHW row is not.
// ROW3 (fp16 3x3 stride-1 convs, 16 x 16-pixel tiles): the three taps of one kernel COLUMN read the same pixels shifted by whole image rows, and a
// row of a 16-pixel-wide tile is 16 x 128 B = two 1024-byte swizzle atoms -- so ONE TMA box of (16 + 2) rows x 16 pixels serves all three taps as
// UMMA descriptors 2 KB apart, and the pixel operand is fetched 3 times per K chunk instead of 9 (36 KB boxes: 108 KB instead of 288 KB; with the nine
// 16 KB weight tiles 252 KB instead of 432 KB per 64-channel chunk).  The kernel is bound by operand delivery from L2, which this cuts by
// 42 %.  Two rings instead of one: 3 pixel boxes (Bfull / Bempty) and 5 weight tiles (Afull / Aempty); a box is released after its last tap.
// The taps come as COLUMN-GROUP tables (TcArgs::cg: per pixel offset dx the taps with their box row, weight slice and target accumulator), built and
// checked on the host (run_tc_conv_one).
template <class TOut, bool DGRAD, bool SPLIT, bool PAIR, int NPIX, bool MNP, bool ROW3 = false>
__global__ void __launch_bounds__(kConvThreads, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmP,
                                                                  const __grid_constant__ CUtensorMap tmWlo, const __grid_constant__ CUtensorMap tmPlo, TcArgs p) {
    constexpr int P_BYTES = NPIX * BK * 2;
    constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (W_BYTES + P_BYTES);
    constexpr int NSTAGE = (192 * 1024) / STAGE_BYTES;
    constexpr int NSUB = PAIR ? 2 : 1;                          // phases per item
    constexpr int SUB_COLS = (SPLIT ? 2 : 1) * NPIX;            // TMEM columns of one phase
    constexpr int ITEM_COLS = NSUB * SUB_COLS;
    constexpr int NBUF = kTmemCols / ITEM_COLS;                 // accumulator buffers (1 = no overlap, only fp32 up-layers)
    static_assert(NBUF >= 1 && NSTAGE >= 2, "tile configuration does not fit");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B tiles need 1024-byte alignment
    uint64_t* full_bar = (uint64_t*)(smem + NSTAGE * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + NSTAGE;
    uint64_t* tfull_bar = empty_bar + NSTAGE;                   // [NBUF] accumulators ready
    uint64_t* tempty_bar = tfull_bar + 2;                       // [NBUF] accumulators drained
    uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
    float* s_noise = (float*)(tmem_slot + 4);                   // [2][NPIX] noise of the current / next item's pixels
    static_assert(!ROW3 || (!SPLIT && !MNP && NPIX == 256), "ROW3 is an fp16, 256-pixel-tile configuration");
    constexpr int R3_TW = 16, R3_TH = 16, R3_NB = 3, R3_NA = 5;
    constexpr int R3_PBOX = (R3_TH + 2) * R3_TW * BK * 2;       // 36 KB: 18 rows x 16 pixels x 64 channels
    static_assert(R3_NB * R3_PBOX + R3_NA * W_BYTES <= NSTAGE * STAGE_BYTES, "ROW3 rings do not fit");
    uint8_t* r3_b = smem;                                       // pixel-box ring
    uint8_t* r3_a = smem + R3_NB * R3_PBOX;                     // weight-tile ring
    uint64_t* r3_bfull = (uint64_t*)(s_noise + 2 * NPIX);       // [3], then bempty [3], afull [5], aempty [5]
    uint64_t* r3_bempty = r3_bfull + R3_NB;
    uint64_t* r3_afull = r3_bempty + R3_NB;
    uint64_t* r3_aempty = r3_afull + R3_NA;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmP);
        if (SPLIT) { tma_prefetch_desc(&tmWlo); tma_prefetch_desc(&tmPlo); }
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 8); }
        if (ROW3) {
            for (int i = 0; i < R3_NB; i++) { mbar_init(&r3_bfull[i], 1); mbar_init(&r3_bempty[i], 1); }
            for (int i = 0; i < R3_NA; i++) { mbar_init(&r3_afull[i], 1); mbar_init(&r3_aempty[i], 1); }
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int c_tiles = p.Nout / CH;
    const int n_tiles = (p.N + p.tn - 1) / p.tn;
    const int total_items = p.tiles_w * p.tiles_h * n_tiles * c_tiles * p.ngroups;

    // item index -> coordinates; false for items that fall outside their phase grid (skipped by every role alike)
    auto decode = [&](int t, int& grp, int& n0, int& h0, int& w0, int& c0) -> bool {
        grp = t % p.ngroups; t /= p.ngroups;
        c0 = (t % c_tiles) * CH; t /= c_tiles;
        const int tile_w = t % p.tiles_w; t /= p.tiles_w;
        const int tile_h = t % p.tiles_h; t /= p.tiles_h;
        n0 = t * p.tn; h0 = p.org_h + tile_h * p.th; w0 = p.org_w + tile_w * p.tw;
        const TcPhase& ph = p.ph[grp * NSUB];                  // PAIR: px = 0 has the larger (or equal) grid
        return h0 < ph.Hg && w0 < ph.Wg;
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (ROW3 && lane == 0) {
            uint32_t ib = 0, ia = 0;
            for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
                int grp, n0, h0, w0, c0;
                if (!decode(t, grp, n0, h0, w0, c0)) continue;
                for (int g = 0; g < p.cg_n[grp]; g++) {
                    const TcArgs::ColGroup& cg = p.cg[grp][g];
                    for (int kc = 0; kc < p.kchunks; kc++, ib++) {
                        const int bs = ib % R3_NB;
                        mbar_wait(&r3_bempty[bs], ((ib / R3_NB) & 1) ^ 1);
                        mbar_expect_tx(&r3_bfull[bs], R3_PBOX);
                        tma_load_4d(r3_b + bs * R3_PBOX, &tmP, &r3_bfull[bs], kc * BK, w0 + cg.dx, h0 + p.cg_dy0[grp], n0);
                        for (int m = 0; m < cg.nm; m++, ia++) {
                            const int as = ia % R3_NA;
                            mbar_wait(&r3_aempty[as], ((ia / R3_NA) & 1) ^ 1);
                            mbar_expect_tx(&r3_afull[as], W_BYTES);
                            tma_load_3d(r3_a + as * W_BYTES, &tmW, &r3_afull[as], kc * BK, c0, cg.tb[m]);
                        }
                    }
                }
            }
        } else if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
                int grp, n0, h0, w0, c0;
                if (!decode(t, grp, n0, h0, w0, c0)) continue;
#pragma unroll 1
                for (int sub = 0; sub < NSUB; sub++) {
                    const TcPhase& ph = p.ph[grp * NSUB + sub];
                    const int iters = ph.ntaps * p.kchunks;
                    for (int k = 0; k < iters; k++, it++) {
                        const int s = it % NSTAGE;
                        const uint32_t par = (it / NSTAGE) & 1;
                        mbar_wait(&empty_bar[s], par ^ 1);
                        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                        const int tap = k / p.kchunks, kc = k - tap * p.kchunks;
                        uint8_t* sw = smem + s * STAGE_BYTES;
                        const int cw = w0 * p.a_s + ph.dx[tap], chh = h0 * p.a_s + ph.dy[tap];
                        if (MNP) {
                            tma_load_4d(sw, &tmW, &full_bar[s], kc * BK, c0, ph.tb[tap], n0);
#pragma unroll
                            for (int r = 0; r < NPIX / 64; r++)      // 64-pixel box r of the tile (row-major over th x tw)
                                tma_load_4d(sw + W_BYTES + r * 8192, &tmP, &full_bar[s], cw + ((64 * r) & (p.tw - 1)), chh + ((64 * r) >> p.tw_sh), kc * BK, n0);
                        } else {
                            tma_load_3d(sw, &tmW, &full_bar[s], kc * BK, c0, ph.tb[tap]);
                            tma_load_4d(sw + W_BYTES, &tmP, &full_bar[s], kc * BK, cw, chh, n0);
                        }
                        if (SPLIT) {
                            tma_load_3d(sw + W_BYTES + P_BYTES, &tmWlo, &full_bar[s], kc * BK, c0, ph.tb[tap]);
                            tma_load_4d(sw + 2 * W_BYTES + P_BYTES, &tmPlo, &full_bar[s], kc * BK, cw, chh, n0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one thread)
        if (ROW3 && lane == 0) {
            constexpr uint32_t idesc3 = make_idesc_f16(CH, NPIX);
            uint32_t ib = 0, ia = 0, icount = 0;
            for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
                int grp, n0, h0, w0, c0;
                if (!decode(t, grp, n0, h0, w0, c0)) continue;
                const uint32_t buf = icount % NBUF;
                mbar_wait(&tempty_bar[buf], ((icount / NBUF) & 1) ^ 1);
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * ITEM_COLS;
                uint32_t accum[2] = {0u, 0u};             // per phase accumulator: the first MMA into it overwrites
                for (int g = 0; g < p.cg_n[grp]; g++) {
                    const TcArgs::ColGroup& cg = p.cg[grp][g];
                    for (int kc = 0; kc < p.kchunks; kc++, ib++) {
                        const int bs = ib % R3_NB;
                        mbar_wait(&r3_bfull[bs], (ib / R3_NB) & 1);
                        const uint32_t pbase = smem_u32(r3_b + bs * R3_PBOX);
                        for (int m = 0; m < cg.nm; m++, ia++) {
                            const int as = ia % R3_NA;
                            mbar_wait(&r3_afull[as], (ia / R3_NA) & 1);
                            tc_fence_after();
                            const int sub = cg.sub[m];
                            const uint64_t wdesc = make_kmajor_sw128_desc(smem_u32(r3_a + as * W_BYTES));
                            const uint64_t pdesc = make_kmajor_sw128_desc(pbase + cg.row[m] * (R3_TW * BK * 2));   // the box from image row row[m] on: 2 KB per row
#pragma unroll
                            for (int kk = 0; kk < BK / 16; kk++) {
                                umma_f16(acc + sub * SUB_COLS, wdesc + (uint64_t)(2 * kk), pdesc + (uint64_t)(2 * kk), idesc3, accum[sub]);
                                accum[sub] = 1u;
                            }
                            umma_commit(&r3_aempty[as]);
                        }
                        umma_commit(&r3_bempty[bs]);      // every tap of this box has been issued: free it once they have read it
                    }
                }
                umma_commit(&tfull_bar[buf]);
                icount++;
            }
        } else if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(CH, NPIX) | (MNP ? (1u << 16) : 0u);      // bit 16: B operand MN-major
            uint32_t it = 0, icount = 0;
            for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
                int grp, n0, h0, w0, c0;
                if (!decode(t, grp, n0, h0, w0, c0)) continue;
                const uint32_t buf = icount % NBUF;
                mbar_wait(&tempty_bar[buf], ((icount / NBUF) & 1) ^ 1);       // the epilogue has drained this buffer
                tc_fence_after();
#pragma unroll 1
                for (int sub = 0; sub < NSUB; sub++) {
                    const int iters = p.ph[grp * NSUB + sub].ntaps * p.kchunks;
                    const uint32_t acc = tmem_base + buf * ITEM_COLS + sub * SUB_COLS;
                    for (int k = 0; k < iters; k++, it++) {
                        const int s = it % NSTAGE;
                        const uint32_t par = (it / NSTAGE) & 1;
                        mbar_wait(&full_bar[s], par);
                        tc_fence_after();
                        const uint32_t sw = smem_u32(smem + s * STAGE_BYTES);
                        const uint64_t wdesc = make_kmajor_sw128_desc(sw);
                        const uint64_t pdesc = MNP ? make_mnmajor_sw128_desc(sw + W_BYTES, 8192) : make_kmajor_sw128_desc(sw + W_BYTES);
                        const uint64_t wdesc_lo = make_kmajor_sw128_desc(sw + W_BYTES + P_BYTES);
                        const uint64_t pdesc_lo = make_kmajor_sw128_desc(sw + 2 * W_BYTES + P_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / 16; kk++) {
                            // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle span: +2 in the (>>4) address field
                            const uint64_t ko = (uint64_t)(2 * kk);
                            const uint64_t pko = MNP ? (uint64_t)(kk * 128) : ko;                 // MN-major: 16 K rows of 128 bytes
                            const uint32_t accum = (k > 0 || kk > 0) ? 1u : 0u;
                            umma_f16(acc, wdesc + ko, pdesc + pko, idesc, accum);
                            if (SPLIT) {
                                umma_f16(acc + NPIX, wdesc + ko, pdesc_lo + ko, idesc, accum);
                                umma_f16(acc + NPIX, wdesc_lo + ko, pdesc + ko, idesc, 1u);
                            }
                        }
                        umma_commit(&empty_bar[s]);       // frees this smem stage once the MMAs above have read it
                    }
                }
                umma_commit(&tfull_bar[buf]);             // accumulators complete
                icount++;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps (2..9)
        const int lg = warp & 3;                           // TMEM lane group this warp may access (warp id % 4)
        const int chalf = (warp - 2) >> 2;                 // which half of the tile's pixel columns
        const int et = threadIdx.x - 64;                   // 0..255 among the epilogue threads
        const float gsv = p.gscale_inv ? *p.gscale_inv : 1.f;    // undoes the global power-of-two scale of the operand
        const size_t HW = (size_t)p.out_H * p.out_pitch;   // plane pitch of out (and of aux, same tensor shape)
        const size_t HWd = (size_t)p.out_H * p.out_W;      // dense plane (residual)
        const bool ep_on = !DGRAD && p.ep.enable;
        const bool has_res = ep_on && p.ep.residual != nullptr;
        const bool has_add = !DGRAD && p.add != nullptr;
        const bool do_ds = DGRAD && p.aux_sum != nullptr;
        const bool wide = p.tw >= 16;                      // every 16-column step lies inside one image row
        const TOut* side_base = DGRAD ? (do_ds ? (const TOut*)p.aux : nullptr) : (has_res ? (const TOut*)p.ep.residual : nullptr);
        const size_t side_plane = DGRAD ? HW : HWd;
        const int side_pitch = DGRAD ? p.out_pitch : p.out_W;
        const int tile_hw_sh = p.tw_sh + p.th_sh;
        constexpr int NIT = NPIX / 32;                     // 16-column steps (= 16-pixel chunks) of this warp
        constexpr int RW = Raw16<TOut>::W;                 // 32-bit words of 16 elements
        constexpr bool PRE = !PAIR;                        // side inputs exist only without pairing
        uint32_t icount = 0;

        // tile column -> image coordinates
        auto locate = [&](int j, int n0, int h0, int w0, int& n, int& gh, int& gw) {
            gw = w0 + (j & (p.tw - 1));
            gh = h0 + ((j >> p.tw_sh) & (p.th - 1));
            n = n0 + (j >> tile_hw_sh);
        };

        for (int t = blockIdx.x; t < total_items; t += gridDim.x) {
            int grp, n0, h0, w0, c0;
            if (!decode(t, grp, n0, h0, w0, c0)) continue;
            const TcPhase& ph0 = p.ph[grp * NSUB];
            const int ch = c0 + lg * 32 + lane;
            const uint32_t buf = icount % NBUF;
            float* nz_tile = s_noise + (icount & 1) * NPIX;

            // ---- everything that does not depend on the accumulators is fetched while the MMAs of this item still run ----
            if (has_add) {
                // noise of the tile's pixels: one pixel per thread -> shared memory (double-buffered by item parity)
                if (et < NPIX) {
                    int n, gh, gw;
                    locate(et, n0, h0, w0, n, gh, gw);
                    const int oy = gh * p.out_s + ph0.oy, ox = gw * p.out_s + ph0.ox;
                    float v = 0.f;
                    if (n < p.N && gh < ph0.Hg && gw < ph0.Wg && oy < p.out_H && ox < p.out_W) v = __ldg(p.add + (size_t)n * p.add_sn + (size_t)oy * p.out_W + ox);
                    nz_tile[et] = v;
                }
            }
            float bias = 0.f, gam = 1.f;
            if (ep_on) {
                if (p.ep.bias) bias = to_acc(((const TOut*)p.ep.bias)[ch]);
                if (has_res) gam = p.ep.gamma[ch] * p.ep.res_scale;
            }
            const float bias0 = bias;
            // can the whole item go through the vectorised fast path?  (complete 16-pixel chunks, 32-byte aligned tensors)
            const bool wide_item = wide && p.vec_out && (!side_base || p.vec_side) && (PAIR || p.out_s == 1) &&
                                   (PAIR ? (2 * (w0 + p.tw) <= p.out_W) : (w0 + p.tw <= ph0.Wg && w0 + p.tw + ph0.ox <= p.out_W));
            // side input (x of the dstyles reduction / residual of the fused layer): this thread's 16-pixel chunks, prefetched
            // (a rolling window of PF chunks: the first PF are requested before the accumulators are awaited, chunk c + PF when
            // chunk c has been unpacked -- holding all NIT chunks costs 64 registers and serialises the epilogue arithmetic)
            static_assert(NIT % 2 == 0, "the step loop is unrolled by two");
            constexpr int UNR = MNP ? 2 : NIT;                 // unroll factor of the step loop of the fast path (see there)
            constexpr int PF = (UNR < 4 || sizeof(TOut) == 4) ? 2 : 4;      // divides UNR (static register slots); fp32 chunks are 16 words
            uint32_t pre[PRE ? PF * RW : 1];
            auto side_fetch = [&](int c, uint32_t* dst) {
                int n, gh, gw;
                locate(chalf * (NPIX / 2) + c * 16, n0, h0, w0, n, gh, gw);
                if (n < p.N && gh < ph0.Hg && gh < p.out_H)
                    raw16_load<TOut>(side_base + ((size_t)n * p.Nout + ch) * side_plane + (size_t)gh * side_pitch + gw, dst);
            };
            if (PRE && side_base && wide_item) {
#pragma unroll
                for (int c = 0; c < PF; c++) side_fetch(c, &pre[c * RW]);
            }
            if (has_add) asm volatile("bar.sync 1, 256;" ::: "memory");

            int cur_n = -1;
            float scale = 0.f, ds_acc = 0.f;
            float res_a = p.ep.res_scale, res_b = 0.f;          // residual term = raw * res_a + res_b  (layer scale folded in)
            auto set_sample = [&](int n) {
                if (n == cur_n) return;
                if (do_ds && cur_n >= 0 && cur_n < p.N) atomicAdd(&p.aux_sum[(size_t)cur_n * p.Nout + ch], ds_acc * gsv);
                ds_acc = 0.f;
                cur_n = n;
                scale = (n >= 0 && n < p.N) ? p.oscale[(size_t)n * p.Nout + ch] * gsv : 0.f;
                if (!DGRAD && p.bias_nc && n >= 0 && n < p.N) bias = bias0 + p.bias_nc[(size_t)n * p.Nout + ch];
                if (!DGRAD && has_res && p.ep.res_a && n >= 0 && n < p.N) {
                    res_a = p.ep.res_a[(size_t)n * p.Nout + ch] * p.ep.res_scale;
                    res_b = p.ep.res_b[(size_t)n * p.Nout + ch] * p.ep.res_scale;
                }
            };
            set_sample(n0);

            mbar_wait(&tfull_bar[buf], (icount / NBUF) & 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + buf * ITEM_COLS + ((uint32_t)(lg * 32) << 16) + (uint32_t)(chalf * (NPIX / 2));
            constexpr int NLD = NSUB * (SPLIT ? 2 : 1);    // TMEM loads per 16-column step

            // forward epilogue of NV values in place: (+bias) + noise -> activation -> clamp -> layer-scaled residual
            auto finish_fwd = [&](float* o, const float* nz, const float* rs, auto nv_tag) {
                constexpr int NV = decltype(nv_tag)::value;
#pragma unroll
                for (int i = 0; i < NV; i++) o[i] = fmaf(o[i], scale, bias);
                if (has_add) {
#pragma unroll
                    for (int i = 0; i < NV; i++) o[i] += nz[i];
                }
                if (ep_on) {
                    if (p.ep.act == 3) {
                        const float gp = p.ep.gain, gn = p.ep.gain * p.ep.alpha;
#pragma unroll
                        for (int i = 0; i < NV; i++) o[i] *= (o[i] > 0.f) ? gp : gn;
                    } else if (p.ep.act == VFM_EP_ACT_GELU) {
                        const float hg = 0.5f * p.ep.gain;
                        if (sizeof(TOut) == 2) {
                            // fp16 output: erfc by Abramowitz-Stegun 7.1.28, erfc(z) = (1 + a1 z + .. + a6 z^6)^-16 (|error| < 3e-7), with
                            // z = |x| / sqrt2 folded into the coefficients and the SFU's approximate reciprocal: one SFU and 14 FP32
                            // instructions per element, branch-free (7.1.26 needs a second SFU op for exp(-z^2); the IEEE reciprocal
                            // and expf cost ~40) -- the 4C-wide GELU epilogue is bound by issue slots and the SFU, not by the MMAs.
                            //   gelu(x) g = h + |h| erf(z) = (h + |h|) - |h| erfc(z),  h = g x / 2,  g > 0  (no cancellation for x < 0)
                            // Written stage by stage over groups of 8 elements: the two epilogue warps per scheduler cannot hide the
                            // latencies of one element's dependent chain, eight interleaved chains can.
                            constexpr float r2 = 0.70710678118654752f;
                            constexpr float c1 = 0.0705230784f * r2, c2 = 0.0422820123f * r2 * r2, c3 = 0.0092705272f * r2 * r2 * r2,
                                            c4 = 0.0001520143f * r2 * r2 * r2 * r2, c5 = 0.0002765672f * r2 * r2 * r2 * r2 * r2,
                                            c6 = 0.0000430638f * r2 * r2 * r2 * r2 * r2 * r2;
                            constexpr int G = NV >= 8 ? 8 : NV;
#pragma unroll
                            for (int i0 = 0; i0 < NV; i0 += G) {
                                float a[G], q[G];
#pragma unroll
                                for (int i = 0; i < G; i++) { a[i] = fabsf(o[i0 + i]); q[i] = fmaf(c6, a[i], c5); }
#pragma unroll
                                for (int i = 0; i < G; i++) q[i] = fmaf(q[i], a[i], c4);
#pragma unroll
                                for (int i = 0; i < G; i++) q[i] = fmaf(q[i], a[i], c3);
#pragma unroll
                                for (int i = 0; i < G; i++) q[i] = fmaf(q[i], a[i], c2);
#pragma unroll
                                for (int i = 0; i < G; i++) q[i] = fmaf(q[i], a[i], c1);
#pragma unroll
                                for (int i = 0; i < G; i++) q[i] = fmaf(q[i], a[i], 1.f);
#pragma unroll
                                for (int i = 0; i < G; i++) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q[i]) : "f"(q[i]));
#pragma unroll
                                for (int i = 0; i < G; i++) q[i] *= q[i];
#pragma unroll
                                for (int i = 0; i < G; i++) q[i] *= q[i];
#pragma unroll
                                for (int i = 0; i < G; i++) q[i] *= q[i];
#pragma unroll
                                for (int i = 0; i < G; i++) q[i] *= q[i];               // erfc(|x| / sqrt2)
#pragma unroll
                                for (int i = 0; i < G; i++) {
                                    const float h = hg * o[i0 + i];
                                    o[i0 + i] = fmaf(-fabsf(h), q[i], h + fabsf(h));
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < NV; i++) o[i] = hg * o[i] * (1.f + erff(o[i] * 0.70710678118654752f));
                        }
                    } else if (p.ep.gain != 1.f) {
#pragma unroll
                        for (int i = 0; i < NV; i++) o[i] *= p.ep.gain;
                    }
                    if (p.ep.clamp >= 0.f) {
#pragma unroll
                        for (int i = 0; i < NV; i++) o[i] = fminf(fmaxf(o[i], -p.ep.clamp), p.ep.clamp);
                    }
                    if (has_res) {
#pragma unroll
                        for (int i = 0; i < NV; i++) o[i] = fmaf(rs[i], res_a, fmaf(o[i], gam, res_b));
                    }
                }
            };

            if (wide_item) {
                // fast path: every 16-pixel chunk of the item is complete and 32-byte aligned (all tiles of the decoder's layers);
                // ragged or unaligned items take the compact element-wise path below, which keeps this unrolled code small enough
                // for the instruction cache (a 1x1 conv's epilogue is not hidden behind MMAs)
                uint32_t raw[2][NLD][16];
                auto issue = [&](int step, int slot) {
#pragma unroll
                    for (int sub = 0; sub < NSUB; sub++) {
                        tmem_ld16_nowait(acc + sub * SUB_COLS + step * 16, raw[slot][sub * (SPLIT ? 2 : 1)]);
                        if (SPLIT) tmem_ld16_nowait(acc + sub * SUB_COLS + NPIX + step * 16, raw[slot][sub * 2 + 1]);
                    }
                };
                issue(0, 0);
                tmem_ld_wait();
                // 1x1 convs (MNP), whose epilogue is not hidden behind the MMAs: unrolled by two only (static register slots of the
                // TMEM / side-input double buffers) -- fully unrolled, the ~3000 straight-line instructions per item starve the two
                // warps per scheduler of instructions (30 % of the stall samples of the 1x1 GELU conv were no_instructions).  The
                // k x k convs keep the full unroll (compile-time addressing, deeper side-input window; their epilogue hides behind the MMAs).
#pragma unroll 1
                for (int step0 = 0; step0 < NIT; step0 += UNR) {
#pragma unroll
                  for (int u = 0; u < UNR; u++) {
                    const int step = step0 + u;
                    const int slot = u & 1;
                    if (step + 1 < NIT) issue(step + 1, slot ^ 1);     // in flight while this step is processed
                    do {
                        const int jw = step * 16;                      // column inside this warp's half
                        int n, gh, gw0;
                        locate(chalf * (NPIX / 2) + jw, n0, h0, w0, n, gh, gw0);
                        set_sample(n);
                        float sd[16];
                        if (PRE && side_base) {         // (before the range check: the window slot is recycled for every chunk)
                            raw16_unpack<TOut>(&pre[(u % PF) * RW], sd);
                            if (step + PF < NIT) side_fetch(step + PF, &pre[(u % PF) * RW]);
                        }
                        const int oy = gh * p.out_s + ph0.oy;
                        if (n >= p.N || gh >= ph0.Hg || oy >= p.out_H) break;
                        float v[NSUB][16];
#pragma unroll
                        for (int sub = 0; sub < NSUB; sub++)
#pragma unroll
                            for (int i = 0; i < 16; i++) {
                                float a = __uint_as_float(raw[slot][sub * (SPLIT ? 2 : 1)][i]);
                                if (SPLIT) a += __uint_as_float(raw[slot][sub * 2 + 1][i]) * (1.f / kLoScale);
                                v[sub][i] = a;
                            }
                        const size_t plane = (size_t)n * p.Nout + ch;
                        if (PAIR) {
                            // interleave the two horizontal phases: output pixels 2*gw0 .. 2*gw0+31
                            float o[32];
#pragma unroll
                            for (int i = 0; i < 16; i++) { o[2 * i] = v[0][i] * scale; o[2 * i + 1] = v[NSUB - 1][i] * scale; }
                            TOut* outp = (TOut*)p.out + plane * HW + (size_t)oy * p.out_pitch + gw0 * 2;
                            store16<TOut>(outp, o, 16, true);
                            store16<TOut>(outp + 16, o + 16, 16, true);
                            break;
                        }
                        const int ox0 = gw0 + ph0.ox;
                        const size_t off = plane * HW + (size_t)oy * p.out_pitch + ox0;
                        float* o = v[0];
                        if (DGRAD) {
                            if (do_ds) {
#pragma unroll
                                for (int i = 0; i < 16; i++) ds_acc = fmaf(sd[i], o[i], ds_acc);
                            }
#pragma unroll
                            for (int i = 0; i < 16; i++) o[i] *= scale;
                        } else {
                            float nz[16];
                            if (has_add) {
#pragma unroll
                                for (int q = 0; q < 4; q++) {
                                    const float4 a4 = *(const float4*)(nz_tile + chalf * (NPIX / 2) + jw + 4 * q);
                                    nz[4 * q] = a4.x; nz[4 * q + 1] = a4.y; nz[4 * q + 2] = a4.z; nz[4 * q + 3] = a4.w;
                                }
                            }
                            finish_fwd(o, nz, sd, std::integral_constant<int, 16>());
                        }
                        store16<TOut>((TOut*)p.out + off, o, 16, true);
                    } while (0);
                    if (step + 1 < NIT) tmem_ld_wait();
                  }
                }
            } else {
                // tiles of rows shorter than 16 pixels (images up to 8x8): one element at a time, compact code
#pragma unroll 1
                for (int step = 0; step < NIT; step++) {
                    uint32_t raw[NLD][16];
#pragma unroll
                    for (int sub = 0; sub < NSUB; sub++) {
                        tmem_ld16_nowait(acc + sub * SUB_COLS + step * 16, raw[sub * (SPLIT ? 2 : 1)]);
                        if (SPLIT) tmem_ld16_nowait(acc + sub * SUB_COLS + NPIX + step * 16, raw[sub * 2 + 1]);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int jw = step * 16 + i;
                        int n, gh, gw;
                        locate(chalf * (NPIX / 2) + jw, n0, h0, w0, n, gh, gw);
                        set_sample(n);
                        if (n >= p.N || gh >= ph0.Hg) continue;
                        const size_t plane = (size_t)n * p.Nout + ch;
#pragma unroll
                        for (int sub = 0; sub < NSUB; sub++) {
                            const TcPhase& ph = p.ph[grp * NSUB + sub];
                            const int oy = gh * p.out_s + ph.oy, ox = gw * p.out_s + ph.ox;
                            if (gw >= ph.Wg || oy >= p.out_H || ox >= p.out_W) continue;
                            float a = __uint_as_float(raw[sub * (SPLIT ? 2 : 1)][i]);
                            if (SPLIT) a += __uint_as_float(raw[sub * 2 + 1][i]) * (1.f / kLoScale);
                            const size_t off = plane * HW + (size_t)oy * p.out_pitch + ox;
                            float o[1], nz[1], sd[1];
                            o[0] = a;
                            if (DGRAD) {
                                if (do_ds) ds_acc = fmaf(to_acc(side_base[off]), a, ds_acc);
                                o[0] = a * scale;
                            } else {
                                nz[0] = has_add ? nz_tile[chalf * (NPIX / 2) + jw] : 0.f;
                                sd[0] = has_res ? to_acc(side_base[plane * HWd + (size_t)oy * p.out_W + ox]) : 0.f;
                                finish_fwd(o, nz, sd, std::integral_constant<int, 1>());
                            }
                            ((TOut*)p.out)[off] = from_acc<TOut, float>(o[0]);
                        }
                    }
                }
            }
            set_sample(-2);        // flush the dstyles partial of the last sample
            // all tcgen05.ld of this warp have completed: hand the accumulator buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
            icount++;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[o,i,t] = sum_{n, x-pixel q} (d*dz)[n, z(q,t), o] * (x*s')[n, q, i]      z(q,t) = q*sd - off_t
//
// GEMM with M = 128 output channels, N = 128 input channels, K = pixels.  Both operands come straight from the NHWC
// tensors the data-gradient already uses: a TMA box [64 ch, tw, th, tn] is 64 pixel rows of 128 bytes, i.e. an
// MN-major tile (channels contiguous, K = pixel rows) in the 128-byte-swizzled layout; two boxes side by side make the
// 128-wide M (resp. N) extent.  The tap shift (and the stride 2 of the transposed conv) is the coordinate / element
// stride of the dz box, zero padding is TMA out-of-bounds fill.  Split-K over pixel tiles, fp32 red.add into dweight.
constexpr int WK = 64;                      // pixels per K block
constexpr int WBOX_BYTES = WK * 128;        // one [64 px][64 ch] box = 8 KB
constexpr int WOP_BYTES = 2 * WBOX_BYTES;   // 128 channels

// 16-byte vector reduction (sm_90+): one L2 atomic request for four fp32 sums.  The split-K epilogue of the weight gradient is bound by
// the L2's atomic rate (~125 scalar atomics per ns chip-wide measured: 148 CTAs x 32 K atomics = 40 us per wave), so it accumulates into a
// [tap][O][I] scratch whose rows are contiguous along the input channels an epilogue thread holds, with a quarter of the requests.
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// dweight[o][i][k] = scratch[k][o][i]
__global__ void __launch_bounds__(256) dw_scatter_kernel(const float* __restrict__ scratch, float* __restrict__ dw, int OI, int KK) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= OI) return;
    for (int k = 0; k < KK; k++) dw[(size_t)idx * KK + k] = scratch[(size_t)k * OI + idx];
}

struct WgTcArgs {
    int tw, th, tn;              // pixel tile (tw*th*tn == 64)
    int tiles_w, tiles_h, tiles_n;
    int ksplit, ntaps;
    int tap_ay[9], tap_ax[9];    // dz coordinate = q*sd + tap_a
    int tap_widx[9];
    int sd;
    int O, I, KK;
    int ib;                      // I / 128
    int jobs_per_o;              // column-block pairs per 128 output channels
    float* dw;                   // [KK][O][I] fp32 scratch (workspace), accumulated with 16-byte vector reductions; dw_scatter_kernel re-lays it out
    const float* rowscale;       // a[o]
    const float* gscale;         // device scalar multiplied into the result, or NULL
};

// A CTA accumulates D[128 out-ch][up to 256 columns] over its share of the pixel tiles (split-K).  The 256 columns are two
// "column blocks" (tap, 128 input channels):
//   sd == 1: the tap shift is applied to the x operand (B), the dz tile (A) is the same for every tap, so the two blocks may
//            belong to different taps -- also 128-channel layers get N = 256 (shared-memory and L2 traffic per MMA drop by 25%);
//   sd == 2: the shift / stride 2 sits on the dz operand, both blocks share the tap and cover 256 input channels.
// MB = 2 (fp16, O % 256 == 0): the CTA owns TWO 128-channel row blocks (two A tiles, two accumulators = all 512 TMEM columns) that share the B
// tile: 64 KB of operands per 1024 tensor-pipe cycles instead of 48 KB per 512 -- the kernel is bound by operand delivery from L2 (ncu:
// tensor pipe 55-58 % with MB = 1, profiles/r02_ncu_full_bwd_*.md), so a third less traffic per MMA is the lever; 3 stages of 64 KB.
template <bool SPLIT, int MB>
__global__ void __launch_bounds__(192, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                          const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo, WgTcArgs p) {
    static_assert(MB == 1 || !SPLIT, "two row blocks need the TMEM columns the split accumulators use");
    constexpr int A_OP = MB * WOP_BYTES, B_OP = 2 * WOP_BYTES;               // 16 (32) KB + 32 KB
    constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_OP + B_OP);
    constexpr int NST = (192 * 1024) / STAGE_BYTES;                          // 4 (fp16) / 3 (fp16, MB = 2) / 2 (split)
    constexpr int TMEM_COLS = (SPLIT || MB == 2) ? 512 : 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = (uint64_t*)(smem + NST * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + NST;
    uint64_t* accum_bar = empty_bar + NST;
    uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // job -> output-channel block and column blocks
    const int o0 = (blockIdx.x / p.jobs_per_o) * (128 * MB);
    const int jc = blockIdx.x % p.jobs_per_o;
    const int ks = blockIdx.y;
    int btap[2], bi0[2], nb;
    if (p.sd == 1) {
        const int cb0 = 2 * jc, ncb = p.ntaps * p.ib;
        nb = (cb0 + 1 < ncb) ? 2 : 1;
        for (int j = 0; j < 2; j++) { const int cb = (cb0 + j < ncb) ? cb0 + j : cb0; btap[j] = cb / p.ib; bi0[j] = (cb % p.ib) * 128; }
    } else {
        const int pairs = (p.ib + 1) / 2;
        const int tap = jc / pairs, pr = jc % pairs;
        nb = (2 * pr + 1 < p.ib) ? 2 : 1;
        btap[0] = btap[1] = tap;
        bi0[0] = 2 * pr * 128; bi0[1] = (nb == 2) ? bi0[0] + 128 : bi0[0];
    }
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int t_begin = (int)((long long)total_tiles * ks / p.ksplit), t_end = (int)((long long)total_tiles * (ks + 1) / p.ksplit);
    const int iters = t_end - t_begin;
    if (iters <= 0) return;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < NST; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx_bytes = (SPLIT ? 2 : 1) * (A_OP + nb * WOP_BYTES);
            for (int it = 0; it < iters; it++) {
                const int s = it % NST;
                const uint32_t phs = (it / NST) & 1;
                mbar_wait(&empty_bar[s], phs ^ 1);
                mbar_expect_tx(&full_bar[s], tx_bytes);
                int t = t_begin + it;
                const int tile_w = t % p.tiles_w; t /= p.tiles_w;
                const int tile_h = t % p.tiles_h; t /= p.tiles_h;
                const int n0 = t * p.tn, h0 = tile_h * p.th, w0 = tile_w * p.tw;
                // sd == 1: tiles walk the dz grid, x is read at (z - tap_a);  sd == 2: tiles walk the x grid, dz at (q*2 + tap_a)
                const int aw = (p.sd == 1) ? w0 : w0 * p.sd + p.tap_ax[btap[0]], ah = (p.sd == 1) ? h0 : h0 * p.sd + p.tap_ay[btap[0]];
                uint8_t* sa = smem + s * STAGE_BYTES;
#pragma unroll
                for (int m = 0; m < MB; m++) {
                    tma_load_4d(sa + m * WOP_BYTES, &tmA, &full_bar[s], o0 + m * 128, aw, ah, n0);
                    tma_load_4d(sa + m * WOP_BYTES + WBOX_BYTES, &tmA, &full_bar[s], o0 + m * 128 + 64, aw, ah, n0);
                }
                if (SPLIT) {
                    tma_load_4d(sa + A_OP + B_OP, &tmAlo, &full_bar[s], o0, aw, ah, n0);
                    tma_load_4d(sa + A_OP + B_OP + WBOX_BYTES, &tmAlo, &full_bar[s], o0 + 64, aw, ah, n0);
                }
                for (int j = 0; j < nb; j++) {
                    const int bw = (p.sd == 1) ? w0 - p.tap_ax[btap[j]] : w0, bh = (p.sd == 1) ? h0 - p.tap_ay[btap[j]] : h0;
                    uint8_t* sb = sa + A_OP + j * WOP_BYTES;
                    tma_load_4d(sb, &tmB, &full_bar[s], bi0[j], bw, bh, n0);
                    tma_load_4d(sb + WBOX_BYTES, &tmB, &full_bar[s], bi0[j] + 64, bw, bh, n0);
                    if (SPLIT) {
                        tma_load_4d(sb + A_OP + B_OP, &tmBlo, &full_bar[s], bi0[j], bw, bh, n0);
                        tma_load_4d(sb + A_OP + B_OP + WBOX_BYTES, &tmBlo, &full_bar[s], bi0[j] + 64, bw, bh, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16_mnmajor(128, 128 * nb);
            for (int it = 0; it < iters; it++) {
                const int s = it % NST;
                const uint32_t phs = (it / NST) & 1;
                mbar_wait(&full_bar[s], phs);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t adesc = make_mnmajor_sw128_desc(sa, WBOX_BYTES);
                const uint64_t bdesc = make_mnmajor_sw128_desc(sa + A_OP, WBOX_BYTES);
                const uint64_t adesc_lo = make_mnmajor_sw128_desc(sa + A_OP + B_OP, WBOX_BYTES);
                const uint64_t bdesc_lo = make_mnmajor_sw128_desc(sa + 2 * A_OP + B_OP, WBOX_BYTES);
#pragma unroll
                for (int k = 0; k < WK / 16; k++) {
                    const uint64_t ko = (uint64_t)((k * 16 * 128) >> 4);     // 16 pixel rows of 128 bytes
                    const uint32_t first = (it > 0 || k > 0) ? 1u : 0u;
                    umma_f16(tmem_base, adesc + ko, bdesc + ko, idesc, first);
                    if (MB == 2) umma_f16(tmem_base + 256, make_mnmajor_sw128_desc(sa + WOP_BYTES, WBOX_BYTES) + ko, bdesc + ko, idesc, first);
                    if (SPLIT) {
                        umma_f16(tmem_base + 256, adesc + ko, bdesc_lo + ko, idesc, first);
                        umma_f16(tmem_base + 256, adesc_lo + ko, bdesc + ko, idesc, 1u);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(accum_bar);
        }
    } else {
        const int lg = warp & 3;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int mj = 0; mj < MB * nb; mj++) {
            const int m = mj / nb, j = mj - m * nb;
            const int o = o0 + m * 128 + lg * 32 + lane;
            const float rs = p.rowscale[o] * (p.gscale ? *p.gscale : 1.f);
            float* dst = p.dw + ((size_t)p.tap_widx[btap[j]] * p.O + o) * p.I + bi0[j];      // 128 consecutive input channels of row o
#pragma unroll 1
            for (int q = 0; q < 128 / 16; q++) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(m * 256 + j * 128 + q * 16), v);
                if (SPLIT) {
                    float v1[16];
                    tmem_ld16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(256 + j * 128 + q * 16), v1);
#pragma unroll
                    for (int c = 0; c < 16; c++) v[c] += v1[c] * (1.f / kLoScale);
                }
#pragma unroll
                for (int c = 0; c < 16; c += 4) red_add_v4(dst + q * 16 + c, v[c] * rs, v[c + 1] * rs, v[c + 2] * rs, v[c + 3] * rs);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// gs[2] = c2g (power of two keeping |x * s'| in fp16 range for every sample), gs[3] = 1 / (gk * c2g)
// One CTA (the result is a single scalar the next kernels read): 1024 threads x four independent 16-byte loads per round, so the
// N*I <= ~50 K styles take 3 rounds of memory latency -- the former 256-thread scalar loop was 128+ dependent rounds (77 us per call,
// 89 calls per training step: 3.4 % of the step in the ncu launch list).
__global__ void __launch_bounds__(1024) wgrad_scalars_kernel(const float* __restrict__ iscale, int n_el, float* gs, int has_gk) {
    __shared__ float red[32];
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
    const int nv = ((reinterpret_cast<uintptr_t>(iscale) & 15u) == 0) ? n_el >> 2 : 0;
    const float4* iv = (const float4*)iscale;
    int i = threadIdx.x;
    for (; i + 3 * (int)blockDim.x < nv; i += 4 * blockDim.x) {
        const float4 a = iv[i], b = iv[i + blockDim.x], c = iv[i + 2 * blockDim.x], d = iv[i + 3 * blockDim.x];
        m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
        m1 = fmaxf(m1, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
        m2 = fmaxf(m2, fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w))));
        m3 = fmaxf(m3, fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));
    }
    for (; i < nv; i += blockDim.x) { const float4 a = iv[i]; m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)))); }
    for (int j = 4 * nv + threadIdx.x; j < n_el; j += blockDim.x) m1 = fmaxf(m1, fabsf(iscale[j]));
    float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = 0.f;
        for (int k = 0; k < (int)(blockDim.x + 31) / 32; k++) mm = fmaxf(mm, red[k]);
        float c2g = (mm > 1.f && isfinite(mm)) ? exp2f(-ceilf(log2f(mm))) : 1.f;
        float gk = has_gk ? gs[0] : 1.f;
        gs[2] = c2g;
        gs[3] = 1.f / (gk * c2g);
    }
}

// ------------------------------------------------------------------------------------------------ pre-pass kernels
// x [N,C,H,W] (TIn) * scale[n,c] (* *gscale)  ->  xt [N,H,W,C] fp16 (hi) and, for the split path, the residual
// (v - hi) * 2048 as a second fp16 tensor (lo).  64 channels x 64 pixels per CTA through shared memory: coalesced reads
// along pixels, coalesced 16-byte writes along channels.
template <class TIn, bool SPLIT>
__global__ void __launch_bounds__(256) nhwc_prepass_kernel(const TIn* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                                           const float* __restrict__ gscale, __half* __restrict__ xt, __half* __restrict__ xt_lo, int C, int HW,
                                                           int W, int in_pitch) {
    __shared__ __half s[SPLIT ? 2 : 1][64][66];     // [pixel][channel], 33-word pitch: conflict-free transposed stores
    const int n = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
    const int tid = threadIdx.x;
    const float gs = gscale ? *gscale : 1.f;
    // fast path (fp16, dense 16-byte aligned planes, full tile): 16-byte loads of 8 pixels of one channel -- a warp reads four
    // 128-byte runs -- instead of 2-byte loads
    const bool vec_in = (sizeof(TIn) == 2) && !SPLIT && in_pitch == W && (HW % 8) == 0 && p0 + 64 <= HW && c0 + 64 <= C &&
                        ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
    if (vec_in) {
        const int pg = tid & 7, cl0 = tid >> 3;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int cl = cl0 + 32 * i, c = c0 + cl;
            const float sc = scale[(size_t)n * C + c] * gs;
            const float sh = shift ? shift[(size_t)n * C + c] * gs : 0.f;
            union { uint4 u; __half2 h[4]; } t;
            t.u = ldg_stream((const __half*)x + ((size_t)n * C + c) * HW + p0 + 8 * pg);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float2 f = __half22float2(t.h[k]);
                s[0][8 * pg + 2 * k][cl] = __float2half_rn(fmaf(f.x, sc, sh));
                s[0][8 * pg + 2 * k + 1][cl] = __float2half_rn(fmaf(f.y, sc, sh));
            }
        }
    } else {
        const int pl = tid & 63, cg = tid >> 6;      // 4 channel groups x 64 pixels
        const int pidx = p0 + pl;
        // input rows may be pitched (the blur backward writes 32-byte aligned rows): pixel -> offset inside the plane
        const size_t plane_sz = (in_pitch == W) ? (size_t)HW : (size_t)(HW / W) * in_pitch;
        const size_t poff = (in_pitch == W) ? (size_t)pidx : (size_t)(pidx / W) * in_pitch + (pidx % W);
#pragma unroll 4
        for (int i = 0; i < 16; i++) {
            const int cl = cg + i * 4;
            const int c = c0 + cl;
            float v = 0.f;
            if (pidx < HW && c < C) {
                v = to_acc(x[((size_t)n * C + c) * plane_sz + poff]) * (scale[(size_t)n * C + c] * gs);
                if (shift) v += shift[(size_t)n * C + c] * gs;
            }
            const __half hi = __float2half_rn(v);
            s[0][pl][cl] = hi;
            if (SPLIT) s[1][pl][cl] = __float2half_rn((v - __half2float(hi)) * kLoScale);
        }
    }
    __syncthreads();
    {
        const int cv = tid & 7, pl0 = tid >> 3;      // 8 threads x 8 channels (16 bytes) per pixel, 32 pixels per pass
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int pl = pl0 + i * 32;
            const int pidx = p0 + pl;
            if (pidx >= HW || c0 + cv * 8 + 8 > C) continue;
#pragma unroll
            for (int part = 0; part < (SPLIT ? 2 : 1); part++) {
                union { uint4 u; uint32_t w[4]; } pk;
                const uint32_t* src = (const uint32_t*)&s[part][pl][cv * 8];
#pragma unroll
                for (int k = 0; k < 4; k++) pk.w[k] = src[k];
                __half* dst = (part == 0 ? xt : xt_lo) + ((size_t)n * HW + pidx) * C + c0 + cv * 8;
                *(uint4*)dst = pk.u;
            }
        }
    }
}

struct WidxList { int v[9]; };

// weight [O,I,KK] fp32 (* a[o])  ->  wt[t][O][I] (transpose == 0) or wt[t][I][O] (transpose == 1) fp16 hi (+ lo)
__global__ void weight_prep_kernel(const float* __restrict__ w, const float* __restrict__ a, __half* __restrict__ wt, __half* __restrict__ wt_lo,
                                   int O, int I, int KK, int ntaps, int transpose, WidxList widx) {
    size_t total = (size_t)ntaps * O * I;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int t = (int)(idx / ((size_t)O * I));
        size_t r = idx - (size_t)t * O * I;
        int o, i;
        if (!transpose) { o = (int)(r / I); i = (int)(r - (size_t)o * I); }
        else { i = (int)(r / O); o = (int)(r - (size_t)i * O); }
        float v = w[((size_t)o * I + i) * KK + widx.v[t]] * a[o];
        __half hi = __float2half_rn(v);
        wt[idx] = hi;
        if (wt_lo) wt_lo[idx] = __float2half_rn((v - __half2float(hi)) * kLoScale);
    }
}

// 1x1 direct-NCHW path: per-sample weights  wt[n][o][i] = W[o,i] * a[o] * s'[n,i] * (xa[n,i] or 1)   (fp16)  and the folded
// input shift  pb[n,o] = d[n,o] * sum_i W[o,i] * a[o] * s'[n,i] * xb[n,i]   (fp32, added by the epilogue before the activation).
// One warp per (n, o): coalesced along i, the bias sum by warp shuffle.
__global__ void __launch_bounds__(256) weight_prep_mod1x1_kernel(const float* __restrict__ w, const float* __restrict__ a, const float* __restrict__ iscale,
                                                                 const float* __restrict__ xa, const float* __restrict__ xb, const float* __restrict__ d,
                                                                 __half* __restrict__ wt, float* __restrict__ pb, int N, int O, int I) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= N * O) return;
    const int n = warp / O, o = warp - n * O;
    const float ao = a[o];
    float acc = 0.f;
    for (int i = lane; i < I; i += 32) {
        const float base = w[(size_t)o * I + i] * ao * iscale[(size_t)n * I + i];
        wt[((size_t)n * O + o) * I + i] = __float2half_rn(xa ? base * xa[(size_t)n * I + i] : base);
        if (xb) acc = fmaf(base, xb[(size_t)n * I + i], acc);
    }
    if (pb) {
        acc = warp_sum(acc);
        if (lane == 0) pb[(size_t)n * O + o] = acc * d[(size_t)n * O + o];
    }
}

// per-sample power-of-two normalisation of the activation scale so that |x * scale| stays far from the fp16 limit:
//   a_scale[n,i] = in_scale[n,i] * c2[n],  o_scale[n,o] = out_scale[n,o] / c2[n],  c2 = 2^-ceil(log2(max_i |in_scale|)) if that max > 1
//   with an input affine map x -> x * xa + xb:  a_scale *= xa,  a_shift[n,i] = in_scale * c2 * xb
//   c2_global != NULL: ONE power of two for the whole batch (*c2_global, from wgrad_scalars_kernel) instead of the per-sample one, which
//   makes the scaled operand exactly what the weight gradient needs (training: the forward's operand is kept for the backward)
__global__ void scale_prep_kernel(const float* __restrict__ in_scale, const float* __restrict__ out_scale, float* a_scale, float* o_scale, int Cin, int Cout,
                                  const float* __restrict__ xa = nullptr, const float* __restrict__ xb = nullptr, float* a_shift = nullptr,
                                  const float* __restrict__ c2_global = nullptr) {
    __shared__ float red[32];
    __shared__ float s_c2;
    const int n = blockIdx.x;
    float m = 0.f;
    for (int i = threadIdx.x; i < Cin; i += blockDim.x) m = fmaxf(m, fabsf(in_scale[(size_t)n * Cin + i]));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = 0.f;
        for (int k = 0; k < (int)(blockDim.x + 31) / 32; k++) mm = fmaxf(mm, red[k]);
        s_c2 = c2_global ? *c2_global : ((mm > 1.f && isfinite(mm)) ? exp2f(-ceilf(log2f(mm))) : 1.f);
    }
    __syncthreads();
    const float c2 = s_c2, c2i = 1.f / s_c2;
    for (int i = threadIdx.x; i < Cin; i += blockDim.x) {
        const float sc = in_scale[(size_t)n * Cin + i] * c2;
        a_scale[(size_t)n * Cin + i] = xa ? sc * xa[(size_t)n * Cin + i] : sc;
        if (a_shift) a_shift[(size_t)n * Cin + i] = xb ? sc * xb[(size_t)n * Cin + i] : 0.f;
    }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) o_scale[(size_t)n * Cout + i] = out_scale[(size_t)n * Cout + i] * c2i;
}

// amax over |x * scale[n,c]| (bit pattern of a non-negative float is monotone -> atomicMax on uint)
template <class TIn>
__global__ void __launch_bounds__(256) amax_kernel(const TIn* __restrict__ x, const float* __restrict__ scale, int C, int HW, size_t total, unsigned int* amax_bits,
                                                   int W, int pitch) {
    float m = 0.f;
    if (pitch == W && (HW & 3) == 0 && total < (1ull << 31) && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && sizeof(TIn) == 4) {
        // dense fp32 planes: 16-byte loads, one 32-bit division per vector, two vectors in flight
        const unsigned nv = (unsigned)(total >> 2), step = gridDim.x * blockDim.x;
        const float4* xv = (const float4*)x;
        for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += 2 * step) {
            const float4 a = xv[v];
            const bool two = v + step < nv;
            const float4 b = two ? xv[v + step] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float sa = scale[(v * 4u) / (unsigned)HW], sb = two ? scale[((v + step) * 4u) / (unsigned)HW] : 0.f;
            m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))) * fabsf(sa));
            m = fmaxf(m, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))) * fabsf(sb));
        }
    } else
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t plane = i / HW;
        size_t off = i;
        if (pitch != W) { const int rem = (int)(i - plane * HW); off = plane * (size_t)(HW / W) * pitch + (size_t)(rem / W) * pitch + (rem % W); }   // pitched rows
        m = fmaxf(m, fabsf(to_acc(x[off]) * scale[plane]));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0 && m > 0.f && isfinite(m)) atomicMax(amax_bits, __float_as_uint(m));
}
// gs[0] = power of two bringing amax into [256, 512); gs[1] = 1 / gs[0]
__global__ void gscale_kernel(const unsigned int* amax_bits, float* gs) {
    float m = __uint_as_float(*amax_bits);
    float k = (m > 0.f && isfinite(m)) ? exp2f(8.f - floorf(log2f(m))) : 1.f;
    gs[0] = k; gs[1] = 1.f / k;
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

int encode_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint32_t* box, const uint32_t* estride) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return VFM_ERR_CUDA; }
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t bdim[5], estr[5];
    uint64_t stride = 2;   // fp16
    for (int i = 0; i < rank; i++) {
        gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = estride ? estride[i] : 1;
        stride *= dims[i];
        if (i < rank - 1) gstride[i] = stride;
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, base, gdim, gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with code %d", (int)r); return VFM_ERR_CUDA; }
    return VFM_OK;
}

// Pixel tile (tw x th pixels of tn samples, tw*th*tn == npix, powers of two) with the least padding waste for a Hg x Wg
// grid.  Wide tiles are preferred on ties (longer contiguous runs per epilogue thread); multi-sample tiles only when one
// image is smaller than the tile.  `a_s` is the element stride of the TMA box (box extents are limited to 256).
void pick_tile(int Hg, int Wg, int npix, int a_s, int& tw, int& th, int& tn) {
    double best = 1e30;
    for (int w = 256; w >= 4; w >>= 1) {
        for (int n = 1; n <= npix / w; n <<= 1) {
            const int h = npix / (w * n);
            if (h < 1 || w * h * n != npix) continue;
            if (w * a_s > 256 || h * a_s > 256 || n > 256) continue;
            if (w < 16 && Wg >= 16) continue;
            const double cover = (double)ceil_div(Wg, w) * w * ceil_div(Hg, h) * h;
            const double waste = cover / ((double)Wg * Hg) * (n > 1 && Hg * Wg > w * h ? 4.0 : 1.0);
            if (waste < best - 1e-9) { best = waste; tw = w; th = h; tn = n; }
        }
    }
}
int ilog2(int v) { int r = 0; while ((1 << r) < v) r++; return r; }

size_t smem_bytes() { return (size_t)192 * 1024 + 1024 + 512 + 2 * 256 * sizeof(float); }      // operand rings + alignment slack + barriers + noise tile

template <class TOut, bool DGRAD, bool SPLIT, bool PAIR, int NPIX, bool MNP = false, bool ROW3 = false>
int launch_tc(const CUtensorMap* maps, const TcArgs& a, dim3 grid, double flops, cudaStream_t stream) {
    auto kern = conv_tc_kernel<TOut, DGRAD, SPLIT, PAIR, NPIX, MNP, ROW3>;
    size_t smem = smem_bytes();
    VFM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KernelTimer timer(DGRAD ? (SPLIT ? "modconv_tc_dgrad_split" : "modconv_tc_dgrad") : (SPLIT ? "modconv_tc_fwd_split" : "modconv_tc_fwd"), stream, flops, 0.0,
                      "i%do%dh%d%s%s%s%s", a.kchunks * BK, a.Nout, a.out_H, MNP ? "m" : (ROW3 ? "3" : ""), PAIR ? "p" : "", (!DGRAD && a.ep.enable) ? "e" : "",
                      (DGRAD ? a.aux_sum != nullptr : (a.ep.enable && a.ep.residual)) ? "r" : "");
    kern<<<grid, kConvThreads, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], a);
    return launch_status("modconv conv_tc_kernel");
}

struct TcOperands {
    __half* act; __half* act_lo;     // NHWC [N, Ha, Wa, Cin]
    __half* wt; __half* wt_lo;       // [ntaps][Nout][Cin]
    int N, Ha, Wa, Cin, Nout, ntaps;
};

// One implicit-GEMM conv.  `args` must have ph[], out*, a_s, scales, add/aux already filled in; this sets the tiling.
// nphases == 4: the phases are the sub-pixel phases of a stride-2 transposed conv in the order 2*py + px (PAIR kernel).
// mnp: op.act is the NCHW fp16 tensor itself and op.wt the per-sample weights [N][ntaps][Nout][Cin] (1x1 convs only, see conv_tc_kernel).
// tile shape / origin of one launch when the caller fixes them (edge strips); flop_px = sum over phases of taps x pixels this launch computes
struct TileOverride { int tw, th, tn, org_w, org_h; double flop_px; };

int run_tc_conv_one(bool f32, bool dgrad, const TcOperands& op, TcArgs a, int nphases, cudaStream_t stream, bool mnp, const TileOverride* ov) {
    const bool pair = (nphases == 4);
    if (mnp && (f32 || pair || dgrad || a.a_s != 1 || op.ntaps != 1)) { set_error("tcgen05 conv: direct NCHW operands need an fp16 1x1 forward conv"); return VFM_ERR_INVALID; }
    if (nphases != 1 && nphases != 4) { set_error("tcgen05 conv: unsupported phase structure"); return VFM_ERR_INVALID; }
    if (pair && (dgrad || a.ep.enable || a.add)) { set_error("tcgen05 conv: the paired-phase kernel has no fused epilogue"); return VFM_ERR_INVALID; }
    // paired up=2 phases: 256-pixel tiles need all 512 TMEM columns for one item (no epilogue / MMA overlap) but a third less operand traffic per
    // MMA -- measured faster for the 65- and 129-wide phase grids (1.78 -> 1.58, 2.13 -> 1.98 ms per layer), much slower for the 33-wide one
    // (32-pixel-wide tiles waste half of the second tile column)
    int pair_w = 0;
    for (int i = 0; i < nphases; i++) pair_w = max(pair_w, a.ph[i].Wg);
    const bool pair256 = pair && !f32 && pair_w >= 64 && !ov;
    const int npix = (!f32 && (!pair || pair256)) ? 256 : 128;
    int Hg = 0, Wg = 0;
    double taps_px = 0;
    for (int i = 0; i < nphases; i++) { Hg = max(Hg, a.ph[i].Hg); Wg = max(Wg, a.ph[i].Wg); taps_px += (double)a.ph[i].ntaps * a.ph[i].Hg * a.ph[i].Wg; }
    pick_tile(Hg, Wg, npix, a.a_s, a.tw, a.th, a.tn);
    if (mnp) {                                                // 64-pixel TMA boxes; the widest rows that divide W keep the DRAM accesses long
        a.tw = (op.Wa % 256 == 0) ? 256 : (op.Wa % 128 == 0 ? 128 : 64);
        a.th = npix / a.tw; a.tn = 1;
    }
    // ROW3 (column groups): fp16, stride-1 pixel operand, image a multiple of the 16 x 16 tile; the taps of every phase group must sort into <= 3
    // pixel offsets dx with <= 4 taps each whose rows span <= 3 image rows: the 3x3 stride-1 conv / data gradient (3 groups x 3 taps).
    // The tables (and the kernel) also describe the paired up=2 phases -- py = 0: the box of dx = 0 serves both row taps of BOTH horizontal
    // phases, 2 boxes instead of 6 pixel tiles -- and that variant passes the whole GPU suite, but it measured SLOWER than the plain paired
    // kernel (512->256: 1.19 -> 1.26 ms, 256->128: 1.69 -> 1.74 ms per layer incl. pre-pass and blur), so up=2 is not routed here.
    bool row3 = false;
    if (!f32 && !mnp && !ov && !pair && a.a_s == 1 && Hg % 16 == 0 && Wg % 16 == 0 && nphases == 1 && a.ph[0].ntaps == 9) {
        row3 = true;
        const int ngrp = pair ? 2 : 1, nsub = pair ? 2 : 1;
        for (int grp = 0; grp < ngrp && row3; grp++) {
            int dy0 = 1 << 30, ntap = 0;
            for (int sub = 0; sub < nsub; sub++) for (int t = 0; t < a.ph[grp * nsub + sub].ntaps; t++) { dy0 = min(dy0, a.ph[grp * nsub + sub].dy[t]); ntap++; }
            a.cg_dy0[grp] = dy0;
            a.cg_n[grp] = 0;
            int placed = 0;
            for (int sub = 0; sub < nsub && row3; sub++) {
                const TcPhase& ph = a.ph[grp * nsub + sub];
                for (int t = 0; t < ph.ntaps && row3; t++) {
                    int g = 0;
                    while (g < a.cg_n[grp] && a.cg[grp][g].dx != ph.dx[t]) g++;
                    if (g == a.cg_n[grp]) { if (g == 3) { row3 = false; break; } a.cg[grp][g].dx = ph.dx[t]; a.cg[grp][g].nm = 0; a.cg_n[grp]++; }
                    TcArgs::ColGroup& cg = a.cg[grp][g];
                    const int row = ph.dy[t] - dy0;
                    if (cg.nm == 4 || row > 2) { row3 = false; break; }
                    cg.sub[cg.nm] = sub; cg.row[cg.nm] = row; cg.tb[cg.nm] = ph.tb[t]; cg.nm++;
                    placed++;
                }
            }
            if (placed != ntap) row3 = false;
        }
        if (row3) { a.tw = 16; a.th = 16; a.tn = 1; }
    }
    if (ov) {
        if (ov->tw * ov->th * ov->tn != npix) { set_error("tcgen05 conv: edge-strip tile does not match the kernel's pixel count"); return VFM_ERR_INVALID; }
        a.tw = ov->tw; a.th = ov->th; a.tn = ov->tn; a.org_w = ov->org_w; a.org_h = ov->org_h;
        taps_px = ov->flop_px;
    }
    a.tw_sh = ilog2(a.tw); a.th_sh = ilog2(a.th);
    a.tiles_w = ceil_div(Wg - a.org_w, a.tw); a.tiles_h = ceil_div(Hg - a.org_h, a.th);
    a.kchunks = op.Cin / BK;
    a.N = op.N; a.Nout = op.Nout;
    a.ngroups = pair ? 2 : 1;
    // 16-byte vector access: every 8-pixel chunk starts at a multiple of 8 pixels of a row, rows and planes must keep that alignment
    const int va = f32 ? 8 : 16;       // elements per 32 bytes
    auto aligned32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; };
    a.vec_out = aligned32(a.out) && a.out_pitch % va == 0 && a.tw >= 16;
    const void* side = dgrad ? a.aux : a.ep.residual;
    a.vec_side = side && aligned32(side) && (dgrad ? a.out_pitch : a.out_W) % va == 0 && a.tw >= 16;
    a.vec_add = 0;
    CUtensorMap maps[4];
    uint64_t pdims[4] = {(uint64_t)op.Cin, (uint64_t)op.Wa, (uint64_t)op.Ha, (uint64_t)op.N};
    uint32_t pbox[4] = {(uint32_t)BK, (uint32_t)(a.tw * a.a_s), (uint32_t)((a.th + (row3 ? 2 : 0)) * a.a_s), (uint32_t)a.tn};
    uint32_t pstr[4] = {1u, (uint32_t)a.a_s, (uint32_t)a.a_s, 1u};
    uint64_t wdims[3] = {(uint64_t)op.Cin, (uint64_t)op.Nout, (uint64_t)op.ntaps};
    uint32_t wbox[3] = {(uint32_t)BK, (uint32_t)CH, 1u};
    int st;
    if (mnp) {
        uint64_t wdims4[4] = {(uint64_t)op.Cin, (uint64_t)op.Nout, (uint64_t)op.ntaps, (uint64_t)op.N};
        uint32_t wbox4[4] = {(uint32_t)BK, (uint32_t)CH, 1u, 1u};
        uint64_t xdims[4] = {(uint64_t)op.Wa, (uint64_t)op.Ha, (uint64_t)op.Cin, (uint64_t)op.N};
        uint32_t xbox[4] = {64u, 1u, (uint32_t)BK, 1u};
        st = encode_map(&maps[0], op.wt, 4, wdims4, wbox4, nullptr); if (st) return st;
        st = encode_map(&maps[1], op.act, 4, xdims, xbox, nullptr); if (st) return st;
        maps[2] = maps[0]; maps[3] = maps[1];
        const long long items = (long long)a.tiles_w * a.tiles_h * op.N * (op.Nout / CH);
        dim3 g((unsigned)(items < kNumSMs ? items : kNumSMs), 1, 1);
        return launch_tc<__half, false, false, false, 256, true>(maps, a, g, 2.0 * op.N * taps_px * (double)op.Nout * op.Cin, stream);
    }
    st = encode_map(&maps[0], op.wt, 3, wdims, wbox, nullptr); if (st) return st;
    st = encode_map(&maps[1], op.act, 4, pdims, pbox, pstr); if (st) return st;
    st = encode_map(&maps[2], f32 ? op.wt_lo : op.wt, 3, wdims, wbox, nullptr); if (st) return st;
    st = encode_map(&maps[3], f32 ? op.act_lo : op.act, 4, pdims, pbox, pstr); if (st) return st;
    const long long total_items = (long long)a.tiles_w * a.tiles_h * ceil_div(op.N, a.tn) * a.ngroups * (op.Nout / CH);
    dim3 grid((unsigned)(total_items < kNumSMs ? total_items : kNumSMs), 1, 1);      // persistent: one CTA per SM
    const double flops = 2.0 * op.N * taps_px * (double)op.Nout * op.Cin;
    if (row3) {
        if (dgrad) return launch_tc<__half, true, false, false, 256, false, true>(maps, a, grid, flops, stream);
        return launch_tc<__half, false, false, false, 256, false, true>(maps, a, grid, flops, stream);
    }
    if (!f32) {
        if (dgrad) return launch_tc<__half, true, false, false, 256>(maps, a, grid, flops, stream);
        if (pair && pair256) return launch_tc<__half, false, false, true, 256>(maps, a, grid, flops, stream);
        if (pair) return launch_tc<__half, false, false, true, 128>(maps, a, grid, flops, stream);
        return launch_tc<__half, false, false, false, 256>(maps, a, grid, flops, stream);
    }
    if (dgrad) return launch_tc<float, true, true, false, 128>(maps, a, grid, flops, stream);
    if (pair) return launch_tc<float, false, true, true, 128>(maps, a, grid, flops, stream);
    return launch_tc<float, false, true, false, 128>(maps, a, grid, flops, stream);
}

// The four sub-pixel phase grids of the stride-2 transposed conv are (H+1) x (W+1), (H+1) x W, H x (W+1) and H x W for a power-of-two H x W
// input: the "+1" row and column of the (2H+1) x (2W+1) intermediate.  Tiled as one grid, the extra column and row cost a whole tile column and
// row of zero-filled MMA work -- 2.7x the useful work on the 17 x 17 grids of the 16x16 -> 32x32 layer, 1.5x at 65 x 65, 1.2x at 129 x 129.  So:
// one launch over the exact H x W part of every phase, one for the last column (a 1-pixel-wide strip, all rows incl. the corner) and one for the
// last row (a 1-pixel-high strip) with tiles shaped like the strips.
int run_tc_conv(bool f32, bool dgrad, const TcOperands& op, TcArgs a, int nphases, cudaStream_t stream, bool mnp = false) {
    a.org_w = a.org_h = 0;
    if (nphases != 4) return run_tc_conv_one(f32, dgrad, op, a, nphases, stream, mnp, nullptr);
    int Hmax = 0, Wmax = 0, Hmin = 1 << 30, Wmin = 1 << 30;
    for (int i = 0; i < 4; i++) { Hmax = max(Hmax, a.ph[i].Hg); Wmax = max(Wmax, a.ph[i].Wg); Hmin = min(Hmin, a.ph[i].Hg); Wmin = min(Wmin, a.ph[i].Wg); }
    const bool pow2 = (Hmin & (Hmin - 1)) == 0 && (Wmin & (Wmin - 1)) == 0;
    if (!(pow2 && Hmax == Hmin + 1 && Wmax == Wmin + 1 && Wmin >= 16 && Hmin >= 16)) return run_tc_conv_one(f32, dgrad, op, a, 4, stream, false, nullptr);
    TcArgs m = a;                                                  // the exact H x W part of every phase
    for (int i = 0; i < 4; i++) { m.ph[i].Hg = Hmin; m.ph[i].Wg = Wmin; }
    int st = run_tc_conv_one(f32, dgrad, op, m, 4, stream, false, nullptr); if (st) return st;
    {   // last column of the phases that have one (px = 0), every row
        TileOverride c;
        c.tw = 1; c.th = min(128, 1 << ilog2(Hmax)); c.tn = 128 / c.th; c.org_w = Wmin; c.org_h = 0; c.flop_px = 0;
        for (int i = 0; i < 4; i++) if (a.ph[i].Wg > Wmin) c.flop_px += (double)a.ph[i].ntaps * a.ph[i].Hg;
        st = run_tc_conv_one(f32, dgrad, op, a, 4, stream, false, &c); if (st) return st;
    }
    {   // last row of the phases that have one (py = 0), without the corner
        TcArgs r = a;
        for (int i = 0; i < 4; i++) r.ph[i].Wg = Wmin;
        TileOverride o;
        o.tw = min(128, Wmin); o.th = 1; o.tn = 128 / o.tw; o.org_w = 0; o.org_h = Hmin; o.flop_px = 0;
        for (int i = 0; i < 4; i++) if (a.ph[i].Hg > Hmin) o.flop_px += (double)a.ph[i].ntaps * Wmin;
        st = run_tc_conv_one(f32, dgrad, op, r, 4, stream, false, &o); if (st) return st;
    }
    return VFM_OK;
}

int run_prepass(int dtype, bool split, const void* x, const float* scale, const float* gscale, __half* xt, __half* xt_lo, int N, int C, int HW, cudaStream_t stream,
                const float* shift = nullptr, int W = 0, int in_pitch = 0) {
    if (W <= 0 || in_pitch <= 0) { W = HW; in_pitch = HW; }         // dense planes
    dim3 grid(ceil_div(HW, 64), ceil_div(C, 64), N);
    if (grid.z > 65535 || grid.y > 65535) { set_error("tcgen05 path: batch too large"); return VFM_ERR_INVALID; }
    KernelTimer timer("modconv_nhwc_prepass", stream, 0.0, (double)N * C * HW * ((dtype == VFM_F16 ? 2 : 4) + (split ? 4 : 2)), "c%dhw%d", C, HW);
    if (dtype == VFM_F16) nhwc_prepass_kernel<__half, false><<<grid, 256, 0, stream>>>((const __half*)x, scale, shift, gscale, xt, xt_lo, C, HW, W, in_pitch);
    else if (split) nhwc_prepass_kernel<float, true><<<grid, 256, 0, stream>>>((const float*)x, scale, shift, gscale, xt, xt_lo, C, HW, W, in_pitch);
    else nhwc_prepass_kernel<float, false><<<grid, 256, 0, stream>>>((const float*)x, scale, shift, gscale, xt, xt_lo, C, HW, W, in_pitch);
    return launch_status("modconv nhwc_prepass_kernel");
}

int run_weight_prep(const float* w, const float* a, __half* wt, __half* wt_lo, int O, int I, int KK, const TapTable& taps, int transpose, cudaStream_t stream) {
    WidxList wl;
    for (int t = 0; t < 9; t++) wl.v[t] = t < taps.ntaps ? taps.widx[t] : 0;
    size_t total = (size_t)taps.ntaps * O * I;
    int blocks = (int)((total + 255) / 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    weight_prep_kernel<<<blocks, 256, 0, stream>>>(w, a, wt, wt_lo, O, I, KK, taps.ntaps, transpose, wl);
    return launch_status("modconv weight_prep_kernel");
}

bool is_f32(const vfm_modconv_desc& d) { return d.dtype == VFM_F32; }

struct TcWorkspace {
    __half *act, *act_lo, *wt, *wt_lo;       // NHWC activation operand (x in forward, d*dz in backward), re-laid-out weights
    __half *xt, *xt_lo;                      // backward only: x*s' NHWC for the weight gradient
    __half* wt_mod; float* pb;               // forward, 1x1 direct-NCHW path: per-sample weights [N][O][I] and folded-shift bias [N][O]
    float *a_scale, *a_shift, *o_scale, *gs; // gs: [0] gk, [1] 1/gk, [2] c2g, [3] 1/(gk*c2g)
    float* dw_scratch;                       // backward only: [KK][O][I] accumulation buffer of the weight gradient
    unsigned int* amax;
};

// forward 1x1 convs that can read x straight from NCHW (conv_tc_kernel MNP): fp16, rows that split into 64-pixel boxes
bool mnp_candidate(const vfm_modconv_desc& d) {
    return d.dtype == VFM_F16 && d.kh == 1 && d.kw == 1 && d.up == 1 && d.in_w % 64 == 0 && d.in_channels % 64 == 0;
}

void carve_tc(Carver& cv, const vfm_modconv_desc& d, const Stage1& s, int direction, TcWorkspace& w) {
    const bool f32 = is_f32(d);
    const size_t x_el = (size_t)d.batch * d.in_h * d.in_w * d.in_channels;
    const size_t act_el = direction == 0 ? x_el : (size_t)d.batch * s.zh * s.zw * d.out_channels;
    const size_t wel = (size_t)d.kh * d.kw * d.out_channels * d.in_channels;
    w.act = cv.take<__half>(act_el);
    w.act_lo = f32 ? cv.take<__half>(act_el) : nullptr;
    w.wt = cv.take<__half>(wel);
    w.wt_lo = f32 ? cv.take<__half>(wel) : nullptr;
    w.xt = w.xt_lo = nullptr;
    w.wt_mod = nullptr; w.pb = nullptr;
    if (direction == 0 && mnp_candidate(d)) { w.wt_mod = cv.take<__half>(wel * d.batch); w.pb = cv.take<float>((size_t)d.batch * d.out_channels); }
    if (direction == 1) {
        w.xt = cv.take<__half>(x_el);
        w.xt_lo = f32 ? cv.take<__half>(x_el) : nullptr;
    }
    const size_t nin = (size_t)d.batch * (direction == 0 ? d.in_channels : d.out_channels);
    const size_t nout = (size_t)d.batch * (direction == 0 ? d.out_channels : d.in_channels);
    w.a_scale = cv.take<float>(nin);
    w.a_shift = cv.take<float>(nin);
    w.o_scale = cv.take<float>(nout);
    w.gs = cv.take<float>(4);
    w.amax = cv.take<unsigned int>(4);
    w.dw_scratch = direction == 1 ? cv.take<float>(wel) : nullptr;
}

// 64-pixel K tile with the least padding waste for an H x W grid
void pick_wtile(int H, int W, int& tw, int& th, int& tn) {
    const int cand[][3] = {{16, 4, 1}, {8, 8, 1}, {32, 2, 1}, {4, 16, 1}, {4, 4, 4}, {8, 4, 2}, {2, 2, 16}};
    double best = 1e30;
    for (auto& c : cand) {
        double cover = (double)ceil_div(W, c[0]) * c[0] * ceil_div(H, c[1]) * c[1];
        double waste = cover / ((double)W * H) * (c[2] > 1 && H * W > c[0] * c[1] ? 4.0 : 1.0);
        if (waste < best - 1e-9) { best = waste; tw = c[0]; th = c[1]; tn = c[2]; }
    }
}

// dweight += a[o] * gscale * sum_{n,q} dzt[n, q*sd - off_t, o] * xt[n, q, i]
int run_tc_wgrad(bool f32, const vfm_modconv_desc& d, const Stage1& s, const TcWorkspace& w, const float* rowscale, const float* gscale,
                 float* dweight, cudaStream_t stream) {
    const int N = d.batch, I = d.in_channels, O = d.out_channels;
    WgTcArgs a;
    a.ntaps = s.taps.ntaps; a.sd = s.sd;
    for (int t = 0; t < a.ntaps; t++) { a.tap_ay[t] = -s.taps.off_y[t]; a.tap_ax[t] = -s.taps.off_x[t]; a.tap_widx[t] = s.taps.widx[t]; }
    a.O = O; a.I = I; a.KK = d.kh * d.kw;
    a.dw = w.dw_scratch; a.rowscale = rowscale; a.gscale = gscale;
    VFM_CUDA_OK(cudaMemsetAsync(w.dw_scratch, 0, sizeof(float) * (size_t)O * I * a.KK, stream));
    // sd == 1: the pixel tiles walk the dz grid (== x grid for the padded stride-1 conv) and the shift is on x
    const int gw = (s.sd == 1) ? s.zw : d.in_w, gh = (s.sd == 1) ? s.zh : d.in_h;
    pick_wtile(gh, gw, a.tw, a.th, a.tn);
    a.tiles_w = ceil_div(gw, a.tw); a.tiles_h = ceil_div(gh, a.th); a.tiles_n = ceil_div(N, a.tn);
    const int tiles = a.tiles_w * a.tiles_h * a.tiles_n;
    a.ib = I / 128;
    a.jobs_per_o = (s.sd == 1) ? ceil_div(a.ntaps * a.ib, 2) : a.ntaps * ceil_div(a.ib, 2);
    const int mb = (!f32 && O % 256 == 0) ? 2 : 1;          // row blocks (of 128 output channels) per CTA
    const int jobs = (O / (128 * mb)) * a.jobs_per_o;
    // split-K factor: a CTA's time is (its share of the pixel tiles) x (time of one 64-pixel K block) + its epilogue, and the epilogue -- 32 K
    // (64 K with two row blocks) fp32 sums through L2 atomics while every other CTA of the wave does the same -- costs as much as ~35 K
    // blocks of an fp16 kernel.  Waves of one CTA per SM run back to back, so: minimise waves x (blocks per CTA x t_block + t_epilogue).
    {
        const double t_block = f32 ? 3.0 : (double)mb;             // in units of one 128 x 256 x 64 MMA block (512 tensor-pipe cycles)
        const double t_epi = 35.0 * mb;
        double best = 1e30;
        a.ksplit = 1;
        for (int ks = 1; ks <= tiles && (long long)jobs * ks <= 4LL * kNumSMs; ks++) {
            const double waves = (double)ceil_div(jobs * ks, kNumSMs);
            const double t = waves * (ceil_div(tiles, ks) * t_block + t_epi);
            if (t < best - 1e-9) { best = t; a.ksplit = ks; }
        }
    }
    CUtensorMap maps[4];
    uint64_t adims[4] = {(uint64_t)O, (uint64_t)s.zw, (uint64_t)s.zh, (uint64_t)N};
    uint32_t abox[4] = {64u, (uint32_t)(a.tw * s.sd), (uint32_t)(a.th * s.sd), (uint32_t)a.tn};
    uint32_t astr[4] = {1u, (uint32_t)s.sd, (uint32_t)s.sd, 1u};
    uint64_t bdims[4] = {(uint64_t)I, (uint64_t)d.in_w, (uint64_t)d.in_h, (uint64_t)N};
    uint32_t bbox[4] = {64u, (uint32_t)a.tw, (uint32_t)a.th, (uint32_t)a.tn};
    int st = encode_map(&maps[0], w.act, 4, adims, abox, astr); if (st) return st;
    st = encode_map(&maps[1], w.xt, 4, bdims, bbox, nullptr); if (st) return st;
    st = encode_map(&maps[2], f32 ? w.act_lo : w.act, 4, adims, abox, astr); if (st) return st;
    st = encode_map(&maps[3], f32 ? w.xt_lo : w.xt, 4, bdims, bbox, nullptr); if (st) return st;
    dim3 grid(jobs, a.ksplit, 1);
    if (grid.y > 65535) { set_error("tcgen05 wgrad: grid too large"); return VFM_ERR_INVALID; }
    const size_t smem = (size_t)192 * 1024 + 1024 + 256;
    const double flops = 2.0 * N * d.in_h * d.in_w * (double)O * I * a.ntaps;
    {
    KernelTimer timer(f32 ? "modconv_tc_wgrad_split" : "modconv_tc_wgrad", stream, flops, 0.0, "i%do%dh%ds%d", I, O, d.in_h, s.sd);
    if (f32) {
        VFM_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        wgrad_tc_kernel<true, 1><<<grid, 192, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], a);
    } else if (mb == 2) {
        VFM_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        wgrad_tc_kernel<false, 2><<<grid, 192, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], a);
    } else {
        VFM_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        wgrad_tc_kernel<false, 1><<<grid, 192, smem, stream>>>(maps[0], maps[1], maps[2], maps[3], a);
    }
    }
    st = launch_status("modconv wgrad_tc_kernel"); if (st) return st;
    dw_scatter_kernel<<<ceil_div(O * I, 256), 256, 0, stream>>>(w.dw_scratch, dweight, O * I, a.KK);
    return launch_status("modconv dw_scatter_kernel");
}

}  // namespace

bool tc_supported(const vfm_modconv_desc& d) {
    if (d.dtype != VFM_F16 && d.dtype != VFM_F32) return false;
    if (d.kh != d.kw || d.padding != d.kh / 2) return false;
    if (!(d.kh == 3 || (d.kh == 1 && d.up == 1))) return false;
    if (d.in_channels % 128 != 0 || d.out_channels % 128 != 0) return false;   // each is an N dimension (fwd / dgrad) and a K dimension
    if (d.batch > 65535) return false;
    return get_encode_fn() != nullptr;
}

size_t tc_workspace_bytes(const vfm_modconv_desc& d, int direction) {
    Carver cv(nullptr, ~(size_t)0);
    Stage1 s = make_stage1(d);
    TcWorkspace w;
    carve_tc(cv, d, s, direction, w);
    return cv.off + 512;
}

// where tc_stage1_forward(keep_operand) leaves x * s' * c2g (NHWC fp16 hi / lo) inside its workspace; NULL for the direct-NCHW 1x1 path
void tc_forward_operand(const vfm_modconv_desc& d, void* ws, size_t ws_bytes, const void** hi, const void** lo) {
    *hi = *lo = nullptr;
    if (mnp_candidate(d)) return;
    Carver cv(ws, ws_bytes);
    TcWorkspace w;
    carve_tc(cv, d, make_stage1(d), 0, w);
    if (!cv.ok()) return;
    *hi = w.act; *lo = w.act_lo;
}

int tc_stage1_forward(const vfm_modconv_desc& d, const Stage1& s, const void* x, const float* weight, const Coefs& k, void* z, int zpitch,
                      const float* noise, int64_t noise_sn, const Epilogue& ep, const float* x_scale, const float* x_shift,
                      void* ws, size_t ws_bytes, cudaStream_t stream, int keep_operand) {
    const bool f32 = is_f32(d);
    const int N = d.batch, I = d.in_channels, O = d.out_channels, KK = d.kh * d.kw;
    Carver cv(ws, ws_bytes);
    TcWorkspace w;
    carve_tc(cv, d, s, 0, w);
    if (!cv.ok()) { set_error("modulated_conv2d: tcgen05 workspace too small"); return VFM_ERR_WORKSPACE; }
    if (w.wt_mod && aligned16(x)) {
        // 1x1 direct path: A = x as it lies in memory (NCHW), B = per-sample W * a * s' (* input scale), epilogue scale = d,
        // the input shift becomes a per-(n,o) bias
        {
            KernelTimer timer("modconv_weight_mod1x1", stream, 0.0, (double)N * O * I * 2.0, "i%do%d", I, O);
            weight_prep_mod1x1_kernel<<<ceil_div(N * O * 32, 256), 256, 0, stream>>>(weight, k.a, k.iscale, x_scale, x_shift, k.d, w.wt_mod, x_shift ? w.pb : nullptr, N, O, I);
        }
        int st0 = launch_status("modconv weight_prep_mod1x1_kernel"); if (st0) return st0;
        TcArgs a;
        TcPhase& ph = a.ph[0];
        ph.ntaps = 1; ph.dy[0] = 0; ph.dx[0] = 0; ph.tb[0] = 0;
        ph.oy = ph.ox = 0; ph.Hg = s.zh; ph.Wg = s.zw;
        a.out_s = 1; a.out_H = s.zh; a.out_W = s.zw; a.out_pitch = zpitch; a.a_s = 1;
        a.out = z; a.oscale = k.d; a.gscale_inv = nullptr; a.add = noise; a.add_sn = noise_sn; a.aux = nullptr; a.aux_sum = nullptr; a.ep = ep;
        a.bias_nc = x_shift ? w.pb : nullptr;
        TcOperands op{(__half*)x, nullptr, w.wt_mod, nullptr, N, d.in_h, d.in_w, I, O, 1};
        return run_tc_conv(false, false, op, a, 1, stream, true);
    }
    // A = x * s' * c2, B = W * a, epilogue scale = d / c2
    const float* c2_global = nullptr;
    if (keep_operand && !x_scale) {
        // the operand survives for the weight gradient: one power of two for the whole batch (what run_tc_wgrad's own pre-pass would use)
        wgrad_scalars_kernel<<<1, 1024, 0, stream>>>(k.iscale, N * I, w.gs, 0);
        int st1 = launch_status("modconv wgrad_scalars_kernel"); if (st1) return st1;
        c2_global = w.gs + 2;
    }
    scale_prep_kernel<<<N, 256, 0, stream>>>(k.iscale, k.d, w.a_scale, w.o_scale, I, O, x_scale, x_shift, x_scale ? w.a_shift : nullptr, c2_global);
    int st = launch_status("modconv scale_prep_kernel"); if (st) return st;
    st = run_prepass(d.dtype, f32, x, w.a_scale, nullptr, w.act, w.act_lo, N, I, d.in_h * d.in_w, stream, x_scale ? w.a_shift : nullptr); if (st) return st;
    st = run_weight_prep(weight, k.a, w.wt, w.wt_lo, O, I, KK, s.taps, 0, stream); if (st) return st;

    TcArgs a;
    int nph = 0;
    if (s.sd == 1) {
        TcPhase& ph = a.ph[0];
        ph.ntaps = s.taps.ntaps;
        for (int t = 0; t < ph.ntaps; t++) { ph.dy[t] = s.taps.off_y[t]; ph.dx[t] = s.taps.off_x[t]; ph.tb[t] = t; }
        ph.oy = ph.ox = 0; ph.Hg = s.zh; ph.Wg = s.zw;
        nph = 1; a.out_s = 1;
    } else {
        for (int pa = 0; pa < 2; pa++)
            for (int pb = 0; pb < 2; pb++) {
                TcPhase& ph = a.ph[nph];            // index 2*py + px: what the paired-phase kernel expects
                ph.ntaps = 0;
                for (int t = 0; t < s.taps.ntaps; t++) {
                    int ny = pa + s.taps.off_y[t], nx = pb + s.taps.off_x[t];
                    if ((ny & 1) || (nx & 1)) continue;
                    ph.dy[ph.ntaps] = ny / 2; ph.dx[ph.ntaps] = nx / 2; ph.tb[ph.ntaps] = t;     // exact: even numerators
                    ph.ntaps++;
                }
                ph.oy = pa; ph.ox = pb; ph.Hg = (s.zh - pa + 1) / 2; ph.Wg = (s.zw - pb + 1) / 2;
                if (ph.ntaps == 0 || ph.Hg <= 0 || ph.Wg <= 0) { set_error("modulated_conv2d: degenerate sub-pixel phase"); return VFM_ERR_NO_KERNEL; }
                nph++;
            }
        a.out_s = 2;
    }
    a.out_H = s.zh; a.out_W = s.zw; a.out_pitch = zpitch; a.a_s = 1;
    a.out = z; a.oscale = w.o_scale; a.gscale_inv = nullptr; a.add = noise; a.add_sn = noise_sn; a.aux = nullptr; a.aux_sum = nullptr; a.ep = ep;
    a.bias_nc = nullptr;
    TcOperands op{w.act, w.act_lo, w.wt, w.wt_lo, N, d.in_h, d.in_w, I, O, s.taps.ntaps};
    return run_tc_conv(f32, false, op, a, nph, stream);
}

int tc_stage1_backward(const vfm_modconv_desc& d, const Stage1& s, const void* dz, int dz_pitch, const void* x, const float* weight, const Coefs& k,
                       void* dx, float* dsum, float* dweight, void* ws, size_t ws_bytes, cudaStream_t stream, const void* saved_xt, const void* saved_xt_lo) {
    const bool f32 = is_f32(d);
    const int N = d.batch, I = d.in_channels, O = d.out_channels, KK = d.kh * d.kw;
    Carver cv(ws, ws_bytes);
    TcWorkspace w;
    carve_tc(cv, d, s, 1, w);
    if (!cv.ok()) { set_error("modulated_conv2d backward: tcgen05 workspace too small"); return VFM_ERR_WORKSPACE; }
    int st;
    // A operand shared by the data and the weight gradient: (d * dz) NHWC (fp32: scaled by one power of two, gk)
    const size_t zel = (size_t)N * O * s.zh * s.zw;
    const float* gs = nullptr;
    if (f32) {
        // fp32 gradients can be arbitrarily small: bring the tensor into the fp16 sweet spot with one power-of-two scale
        VFM_CUDA_OK(cudaMemsetAsync(w.amax, 0, sizeof(unsigned int), stream));
        size_t want_blocks = (zel + 255) / 256;
        int blocks = (int)(want_blocks < (size_t)kNumSMs * 8 ? want_blocks : (size_t)kNumSMs * 8);
        amax_kernel<float><<<blocks, 256, 0, stream>>>((const float*)dz, k.d, O, s.zh * s.zw, zel, w.amax, s.zw, dz_pitch > 0 ? dz_pitch : s.zw);
        st = launch_status("modconv amax_kernel"); if (st) return st;
        gscale_kernel<<<1, 1, 0, stream>>>(w.amax, w.gs);
        st = launch_status("modconv gscale_kernel"); if (st) return st;
        gs = w.gs;
    }
    st = run_prepass(d.dtype, f32, dz, k.d, gs, w.act, w.act_lo, N, O, s.zh * s.zw, stream, nullptr, s.zw, dz_pitch > 0 ? dz_pitch : s.zw); if (st) return st;
    if (dx) {
        // dxpre[n,i,p] = sum_{o,t} (a*W)[o,i,widx(t)] * (d*dz)[n,o,z(p,t)];  dx = s' * dxpre;  dsum = sum_p x * dxpre
        TapTable dt; int sn, sd;
        dgrad_taps(s, dt, sn, sd);
        st = run_weight_prep(weight, k.a, w.wt, w.wt_lo, O, I, KK, dt, 1, stream); if (st) return st;
        TcArgs a;
        TcPhase& ph = a.ph[0];
        ph.ntaps = dt.ntaps;
        for (int t = 0; t < dt.ntaps; t++) { ph.dy[t] = dt.off_y[t]; ph.dx[t] = dt.off_x[t]; ph.tb[t] = t; }
        ph.oy = ph.ox = 0; ph.Hg = d.in_h; ph.Wg = d.in_w;
        a.out_s = 1; a.out_H = d.in_h; a.out_W = d.in_w; a.out_pitch = d.in_w; a.a_s = sn;
        a.out = dx; a.oscale = k.iscale; a.gscale_inv = gs ? gs + 1 : nullptr; a.add = nullptr; a.add_sn = 0;
        a.aux = dsum ? x : nullptr; a.aux_sum = dsum; a.ep = no_epilogue(); a.bias_nc = nullptr;
        TcOperands op{w.act, w.act_lo, w.wt, w.wt_lo, N, s.zh, s.zw, O, I, dt.ntaps};
        st = run_tc_conv(f32, true, op, a, 1, stream); if (st) return st;
    }
    if (dweight) {
        // B operand: x * s' * c2g NHWC; result scale a[o] / (gk * c2g)
        wgrad_scalars_kernel<<<1, 1024, 0, stream>>>(k.iscale, N * I, w.gs, gs ? 1 : 0);
        st = launch_status("modconv wgrad_scalars_kernel"); if (st) return st;
        if (saved_xt && (!f32 || saved_xt_lo)) {
            // the forward kept exactly this tensor (same iscale, same c2g): no second pass over x
            w.xt = (__half*)saved_xt; w.xt_lo = (__half*)saved_xt_lo;
        } else {
            st = run_prepass(d.dtype, f32, x, k.iscale, w.gs + 2, w.xt, w.xt_lo, N, I, d.in_h * d.in_w, stream); if (st) return st;
        }
        st = run_tc_wgrad(f32, d, s, w, k.a, w.gs + 3, dweight, stream); if (st) return st;
    }
    return VFM_OK;
}

}  // namespace modconv
}  // namespace vfm
