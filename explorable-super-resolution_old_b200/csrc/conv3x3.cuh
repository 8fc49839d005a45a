// Shared definitions of the 3x3 convolution kernels (tcgen05 and SIMT variants).
//
// Formulation (DESIGN.md "conv3x3 kernel"): for an output tile of 8 rows x 32
// columns the kernel computes, per 4-row band (128 pixels = the MMA M dimension),
//     D[p, (dx, co)] = sum_{dy, ci}  X[p + dy*row, ci] * W[co, ci, dy, dx]
// i.e. the three filter columns dx are stacked along the MMA N dimension
// (N = 3*cout_tile) and only the filter rows dy are walked in the K loop, as
// whole-row shifts of a halo tile that TMA loaded once.  The epilogue finishes
// the convolution with two warp shuffles per channel:
//     out[p] = D[p-1, (dx=-1)] + D[p, (dx=0)] + D[p+1, (dx=+1)]
// which is why each 32-column tile produces 30 output columns.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "esr_common.cuh"

namespace esr {

constexpr int kTileW = 32;           // tile columns loaded / MMA'd (one warp of lanes)
constexpr int kTileWOut = 30;        // columns produced per tile
constexpr int kBandRows = 4;         // rows per MMA (4 x 32 = 128 = M)
constexpr int kBands = 2;            // bands per tile
constexpr int kTileH = kBandRows * kBands;       // 8 output rows per tile
constexpr int kHaloRows = kTileH + 2;            // 10 input rows per tile
constexpr int kKB = ESR_KBLOCK_CH;               // 32 channels per K block
constexpr int kRowBytes = kKB * 2;               // 64-byte smem rows (SWIZZLE_64B)
constexpr int kABytes = kHaloRows * kTileW * kRowBytes;  // 20480

struct ConvLaunch {
    esr_conv_desc d;
    int tiles_x, tiles_y, spatial_tiles, total_tiles;
    uint32_t w_smem_bytes;   // resident weight region (w_tile_bytes rounded up to 1 KiB)
    int nstages;             // depth of the A-tile ring that fits beside it
    int pair_nb;             // 0: single-CTA kernel; else bands per CTA tile of the cta_group::2 kernel
    int a_stream;            // activation tiles of source 0 are dead after this launch (conv 4 of an RDB reads the whole
                             // dense block for the last time): TMA loads them with an L2 evict_first policy
    int reverse;             // walk the spatial tiles last-to-first: consecutive layers of a recorded sequence alternate
                             // direction, so a layer starts on the data the previous one touched last (still in L2)
    int debug;               // ESR_DEBUG_SKIP timing experiments (results invalid when non-zero)
    unsigned long long* prof; // optional [gridDim][16] per-role cycle counters (esr_debug_set_profile_buffer)
};

// Byte offset of element (row n, channel k) inside a [rows x 32ch] SWIZZLE_64B K-major
// slab: 16-byte chunk index is XORed with bits [7,9) of the row offset.
__host__ __device__ inline uint32_t sw64_offset(uint32_t n, uint32_t k) {
    return n * 64u + ((((k >> 3) ^ (n >> 1)) & 3u) << 4) + (k & 7u) * 2u;
}

// Same for a [rows x 16ch] SWIZZLE_32B slab (32-byte rows): the 16-byte chunk bit is XORed with address bit 7.
__host__ __device__ inline uint32_t sw32_offset(uint32_t n, uint32_t k) {
    return n * 32u + ((((k >> 3) ^ (n >> 2)) & 1u) << 4) + (k & 7u) * 2u;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&p);
}
// two floats -> packed fp16 pair (a in the low half), round to nearest, overflow saturates to +-65504
__device__ __forceinline__ uint32_t pack_f16x2_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// 16-bit operand storage: term 0 = bf16(v), 1 = bf16(v - bf16(v)), 2 = fp16(v)
__device__ __forceinline__ uint16_t operand_bits(float v, int term) {
    if (term == 2) return static_cast<uint16_t>(pack_f16x2_sat(v, 0.f) & 0xffffu);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 r = term == 0 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
    return *reinterpret_cast<const uint16_t*>(&r);
}
__device__ __forceinline__ float operand_value(uint16_t bits, bool f16) {
    return f16 ? __half2float(*reinterpret_cast<const __half*>(&bits)) : __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&bits));
}
__device__ __forceinline__ float bf16_lo_part(float a) {     // a - bf16(a)
    return a - __bfloat162float(__float2bfloat16_rn(a));
}

// 32-byte store: halves the number of store instructions (and L1 wavefronts) of the
// pixel-strided NHWC writes.  Needs a 32-byte aligned address.
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

__device__ __forceinline__ void st_global_v8f(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_v8f(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
// Streaming variants for the fp32 trunk: each value is written once and read once, five launches (~0.6 GB of other
// traffic) later, so it can never be an L2 hit; evict_first keeps it from displacing the bf16 dense-block lines
// that the next layer re-reads.
__device__ __forceinline__ void st_global_v8f_stream(float* p, const float* v) {
    asm volatile("st.global.L2::evict_first.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ld_global_v8f_stream(const float* p, float* v) {
    asm volatile("ld.global.L2::evict_first.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void ld_global_v8u(const void* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
// r[0..16) = hi + lo, two bf16 NHWC tensors (16 consecutive channels of one pixel each, 32-byte aligned)
__device__ __forceinline__ void load16_hilo(const esr_conv_desc& d, size_t pix, int co0, float* r) {
    uint32_t h[8], l[8];
    ld_global_v8u(reinterpret_cast<const __nv_bfloat16*>(d.res1_hi) + pix * d.res1_hi_stride + d.res1_hi_choff + co0, h);
    ld_global_v8u(reinterpret_cast<const __nv_bfloat16*>(d.res1_lo) + pix * d.res1_lo_stride + d.res1_lo_choff + co0, l);
#pragma unroll
    for (int i = 0; i < 8; ++i) {   // bf16 -> f32 is a 16-bit shift
        r[2 * i] = __uint_as_float(h[i] << 16) + __uint_as_float(l[i] << 16);
        r[2 * i + 1] = __uint_as_float(h[i] & 0xffff0000u) + __uint_as_float(l[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void store16_lo(const esr_conv_desc& d, size_t pix, int co0, const float* v) {
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(bf16_lo_part(v[2 * i]), bf16_lo_part(v[2 * i + 1]));
    st_global_v8(reinterpret_cast<__nv_bfloat16*>(d.out_lo) + pix * d.out_lo_stride + d.out_lo_choff + co0, pk);
}
// 16 consecutive floats, 32-byte vector accesses when `wide` (address 32-byte aligned)
__device__ __forceinline__ void load16f(const float* p, float* r, bool wide) {
    if (wide) { ld_global_v8f(p, r); ld_global_v8f(p + 8, r + 8); }
    else {
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(r + i) = *reinterpret_cast<const float4*>(p + i);
    }
}
__device__ __forceinline__ void store16f(float* p, const float* r, bool wide) {
    if (wide) { st_global_v8f(p, r); st_global_v8f(p + 8, r + 8); }
    else {
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(p + i) = *reinterpret_cast<const float4*>(r + i);
    }
}

constexpr int kEpiGeneric = 0;   // every flag / output decided at run time
constexpr int kEpiTrunk = 1;     // bias + LeakyReLU -> bf16 slice of the dense-block buffer (conv 0..3 of an RDB)
constexpr int kEpiRes = 2;       // conv 4 of an RDB: alpha*(acc+bias) + gamma*res1 [, beta*. + res2] -> blocked f32 trunk
                                 // + 16-bit slice (block.py:235, :270)
constexpr int kEpiAct = 3;       // bias [+ LeakyReLU] -> 16-bit NHWC (bf16 or fp16), optionally replicated 2x2 (upconv / HR conv)
constexpr int kEpiNchw = 4;      // bias -> f32 NCHW, first cout_real channels (last conv of the generator)
constexpr int kEpiMask = 5;      // dgrad of a trunk conv: (acc + bias) * LeakyReLU'(stored activation) -> bf16 slice
constexpr int kEpiDx0Mask = 7;   // kEpiDx0 whose 16-bit output is also scaled by LeakyReLU'(mask): the dgrads of the HR convs (f32 rows
                                 // for the latent + masked bf16 gradient of the conv below); round 1 ran them on the generic epilogue
constexpr int kEpiOuter = 8;     // pair kernel, cout tile 64: bias [, alpha*. + gamma*res1] -> [blocked f32 +] 16-bit NHWC [2x2 replicated]:
                                 // the first conv (f32 trunk + bf16 slice) and LR_conv (+ fea, nearest x2 for the upconv); they ran on
                                 // the generic epilogue (126 / 78 us per launch at config 2 for ~35 us of work)
constexpr int kEpiDx0 = 6;       // dgrad of a block input: per cout tile either v + out_f32 (accumulate, routed latent rows) or
                                 // alpha*v + gamma*res1 [, beta*. + res2] -> blocked f32 [+ scale*v as bf16]
// The generic epilogue costs ~5000 clk per 128-pixel x 64-channel tile (issue bound: two epilogue warps per
// scheduler walking run-time flags); the specialised ones are bound by the TMEM read of the three dx slabs (~1600).

// Address of 8 consecutive f32 channels [c, c+8) (c % 8 == 0) of pixel (n,y,x) in a trunk tensor with C = `stride`
// channels: NHWC, or blocked [B, C/8, H, W, 8] where the 32 pixels of a warp form one contiguous 1 KiB run.
__device__ __forceinline__ size_t f32_off(const esr_conv_desc& d, bool blocked, int stride, int n, int y, int x, int c) {
    if (blocked)
        return ((static_cast<size_t>(n) * (stride >> 3) + (c >> 3)) * d.H + y) * (static_cast<size_t>(d.W) * 8) +
               static_cast<size_t>(x) * 8;
    return ((static_cast<size_t>(n) * d.H + y) * d.W + x) * stride + c;
}
__device__ __forceinline__ void load16f_at(const esr_conv_desc& d, const float* base, int stride, int choff, int n,
                                           int y, int x, int co0, float* r) {
    const bool blocked = (d.flags & ESR_EPI_F32_BLOCKED) != 0;
    if (blocked) {
        ld_global_v8f(base + f32_off(d, true, stride, n, y, x, choff + co0), r);
        ld_global_v8f(base + f32_off(d, true, stride, n, y, x, choff + co0 + 8), r + 8);
    } else {
        load16f(base + f32_off(d, false, stride, n, y, x, choff + co0), r, (d.flags & ESR_EPI_WIDE_OK) != 0);
    }
}
__device__ __forceinline__ void store16f_at(const esr_conv_desc& d, float* base, int stride, int choff, int n, int y,
                                            int x, int co0, const float* r) {
    const bool blocked = (d.flags & ESR_EPI_F32_BLOCKED) != 0;
    if (blocked) {
        st_global_v8f(base + f32_off(d, true, stride, n, y, x, choff + co0), r);
        st_global_v8f(base + f32_off(d, true, stride, n, y, x, choff + co0 + 8), r + 8);
    } else {
        store16f(base + f32_off(d, false, stride, n, y, x, choff + co0), r, (d.flags & ESR_EPI_WIDE_OK) != 0);
    }
}

// Residual / accumulator / mask operands of one pixel's 16 channels.  They are fetched BEFORE the
// epilogue warp waits for its accumulator, so their L2/HBM latency hides behind the tile's MMAs.
struct EpiOperands {
    float r1[16];   // out_f32 (ACCUM) or res1
    float r2[16];   // res2
    uint4 m[2];     // LeakyReLU mask source (bf16 x 16)
};

// Output channel of column `col` of cout tile `ct` (dgrad launches re-route whole tiles).
__device__ __forceinline__ int tile_channel(const esr_conv_desc& d, int ct, int col) {
    const int base = d.tile_choff[ct];
    return (base >= 0 ? base : ct * d.cout_tile) + col;
}
__device__ __forceinline__ uint32_t tile_flags(const esr_conv_desc& d, int ct) {
    uint32_t f = d.flags;
    if ((d.no_accum_tiles >> ct) & 1) f &= ~static_cast<uint32_t>(ESR_EPI_ACCUM);
    if ((d.no_res_tiles >> ct) & 1) f &= ~static_cast<uint32_t>(ESR_EPI_RES1 | ESR_EPI_RES2);
    return f;
}

template <int MODE>
__device__ __forceinline__ void conv_epilogue_prefetch(const esr_conv_desc& d, int ct, int n, int y, int x, int co0,
                                                       EpiOperands& P) {
    if constexpr (MODE == kEpiTrunk || MODE == kEpiAct || MODE == kEpiNchw) return;
    if constexpr (MODE == kEpiDx0 || MODE == kEpiDx0Mask) {
        const uint32_t flags = tile_flags(d, ct);
        if constexpr (MODE == kEpiDx0Mask) {
            if (d.out_bf16 != nullptr && !((d.no_bf16_tiles >> ct) & 1)) {
                const uint4* m = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.mask) +
                                                                ((static_cast<size_t>(n) * d.H + y) * d.W + x) * d.mask_stride + d.mask_choff + co0);
                P.m[0] = __ldg(m);
                P.m[1] = __ldg(m + 1);
            }
        }
        if (flags & ESR_EPI_ACCUM) {
            ld_global_v8f(d.out_f32 + f32_off(d, true, d.out_f32_stride, n, y, x, d.out_f32_choff + co0), P.r1);
            ld_global_v8f(d.out_f32 + f32_off(d, true, d.out_f32_stride, n, y, x, d.out_f32_choff + co0 + 8), P.r1 + 8);
        } else if (flags & ESR_EPI_RES1) {
            ld_global_v8f(d.res1 + f32_off(d, true, d.res1_stride, n, y, x, d.res1_choff + co0), P.r1);
            ld_global_v8f(d.res1 + f32_off(d, true, d.res1_stride, n, y, x, d.res1_choff + co0 + 8), P.r1 + 8);
        }
        if (flags & ESR_EPI_RES2) {
            ld_global_v8f(d.res2 + f32_off(d, true, d.res2_stride, n, y, x, d.res2_choff + co0), P.r2);
            ld_global_v8f(d.res2 + f32_off(d, true, d.res2_stride, n, y, x, d.res2_choff + co0 + 8), P.r2 + 8);
        }
        return;
    }
    if constexpr (MODE == kEpiMask) {
        const size_t pix = (static_cast<size_t>(n) * d.H + y) * d.W + x;
        const uint4* m = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.mask) +
                                                        pix * d.mask_stride + d.mask_choff + co0);
        P.m[0] = __ldg(m);
        P.m[1] = __ldg(m + 1);
        return;
    }
    if constexpr (MODE == kEpiOuter) {
        if (d.flags & ESR_EPI_RES1) {
            ld_global_v8f(d.res1 + f32_off(d, true, d.res1_stride, n, y, x, d.res1_choff + co0), P.r1);
            ld_global_v8f(d.res1 + f32_off(d, true, d.res1_stride, n, y, x, d.res1_choff + co0 + 8), P.r1 + 8);
        }
        return;
    }
    if constexpr (MODE == kEpiRes) {
        if (d.flags & ESR_EPI_RES1_HILO) {
            load16_hilo(d, (static_cast<size_t>(n) * d.H + y) * d.W + x, co0, P.r1);
        } else {
            ld_global_v8f_stream(d.res1 + f32_off(d, true, d.res1_stride, n, y, x, d.res1_choff + co0), P.r1);
            ld_global_v8f_stream(d.res1 + f32_off(d, true, d.res1_stride, n, y, x, d.res1_choff + co0 + 8), P.r1 + 8);
        }
        if (d.flags & ESR_EPI_RES2) {
            ld_global_v8f_stream(d.res2 + f32_off(d, true, d.res2_stride, n, y, x, d.res2_choff + co0), P.r2);
            ld_global_v8f_stream(d.res2 + f32_off(d, true, d.res2_stride, n, y, x, d.res2_choff + co0 + 8), P.r2 + 8);
        }
        return;
    }
    const uint32_t flags = tile_flags(d, ct);
    const size_t pix = (static_cast<size_t>(n) * d.H + y) * d.W + x;
    if (flags & ESR_EPI_ACCUM) load16f_at(d, d.out_f32, d.out_f32_stride, d.out_f32_choff, n, y, x, co0, P.r1);
    else if ((flags & ESR_EPI_RES1) && (flags & ESR_EPI_RES1_HILO)) load16_hilo(d, pix, co0, P.r1);
    else if (flags & ESR_EPI_RES1) load16f_at(d, d.res1, d.res1_stride, d.res1_choff, n, y, x, co0, P.r1);
    if (flags & ESR_EPI_RES2) load16f_at(d, d.res2, d.res2_stride, d.res2_choff, n, y, x, co0, P.r2);
    // the mask only scales the 16-bit output: a cout tile without one (the routed latent rows of a dgrad, whose channel
    // index lies beyond the mask tensor's channels) must not touch it - for the tensor's last pixel that read ran past
    // the end of the allocation (found in round 2 when a different allocation order put it at the end of a segment)
    if ((flags & ESR_EPI_MASK) && d.out_bf16 != nullptr && !((d.no_bf16_tiles >> ct) & 1)) {
        const uint4* m = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.mask) +
                                                        pix * d.mask_stride + d.mask_choff + co0);
        P.m[0] = __ldg(m);
        P.m[1] = __ldg(m + 1);
    }
}

// Applies the fused epilogue to 16 consecutive output channels of one pixel.  `bias` points at the
// 16 biases of these channels (shared memory in the tcgen05 kernel, global in the SIMT check).
template <int MODE>
__device__ __forceinline__ void conv_epilogue16(const esr_conv_desc& d, const float* bias, int ct, int n, int y, int x,
                                                int co0, float (&v)[16], const EpiOperands& P) {
    const size_t pix = (static_cast<size_t>(n) * d.H + y) * d.W + x;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        const float4 b = *reinterpret_cast<const float4*>(bias + i);
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
    if constexpr (MODE == kEpiTrunk) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            pk[i] = pack_bf16x2(fmaxf(v[2 * i], d.slope * v[2 * i]), fmaxf(v[2 * i + 1], d.slope * v[2 * i + 1]));
        st_global_v8(reinterpret_cast<__nv_bfloat16*>(d.out_bf16) + pix * d.out_bf16_stride + d.out_bf16_choff + co0, pk);
        return;
    }
    if constexpr (MODE == kEpiRes) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = d.alpha * v[i] + d.gamma * P.r1[i];
        if (d.flags & ESR_EPI_RES2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = d.beta * v[i] + P.r2[i];
        }
        if (d.out_f32 != nullptr) {
            st_global_v8f_stream(d.out_f32 + f32_off(d, true, d.out_f32_stride, n, y, x, d.out_f32_choff + co0), v);
            st_global_v8f_stream(d.out_f32 + f32_off(d, true, d.out_f32_stride, n, y, x, d.out_f32_choff + co0 + 8), v + 8);
        }
        if (d.out_lo != nullptr) store16_lo(d, pix, co0, v);
        uint32_t pk[8];
        if (d.flags & ESR_EPI_OUT_F16) {
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_f16x2_sat(v[2 * i], v[2 * i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        }
        st_global_v8(reinterpret_cast<__nv_bfloat16*>(d.out_bf16) + pix * d.out_bf16_stride + d.out_bf16_choff + co0, pk);
        return;
    }
    if constexpr (MODE == kEpiAct || MODE == kEpiOuter) {
        if constexpr (MODE == kEpiOuter) {
            if (d.flags & ESR_EPI_RES1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = d.alpha * v[i] + d.gamma * P.r1[i];
            }
            if (d.out_f32 != nullptr) {
                st_global_v8f(d.out_f32 + f32_off(d, true, d.out_f32_stride, n, y, x, d.out_f32_choff + co0), v);
                st_global_v8f(d.out_f32 + f32_off(d, true, d.out_f32_stride, n, y, x, d.out_f32_choff + co0 + 8), v + 8);
            }
        } else if (d.flags & ESR_EPI_LRELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], d.slope * v[i]);
        }
        uint32_t pk[8];
        if (d.flags & ESR_EPI_OUT_F16) {
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_f16x2_sat(v[2 * i], v[2 * i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        }
        __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(d.out_bf16) + d.out_bf16_choff + co0;
        if (d.up == 1) {
            st_global_v8(ob + pix * d.out_bf16_stride, pk);
        } else {                                     // nearest x2: the next conv reads the upsampled tensor
            const size_t ow = static_cast<size_t>(d.W) * 2;
            const size_t o00 = (static_cast<size_t>(n) * d.H * 2 + static_cast<size_t>(y) * 2) * ow + static_cast<size_t>(x) * 2;
            st_global_v8(ob + o00 * d.out_bf16_stride, pk);
            st_global_v8(ob + (o00 + 1) * d.out_bf16_stride, pk);
            st_global_v8(ob + (o00 + ow) * d.out_bf16_stride, pk);
            st_global_v8(ob + (o00 + ow + 1) * d.out_bf16_stride, pk);
        }
        return;
    }
    if constexpr (MODE == kEpiDx0 || MODE == kEpiDx0Mask) {
        const uint32_t flags = tile_flags(d, ct);
        if (flags & ESR_EPI_ACCUM) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += P.r1[i];
        } else if (flags & ESR_EPI_RES1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = d.alpha * v[i] + d.gamma * P.r1[i];
        }
        if (flags & ESR_EPI_RES2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = d.beta * v[i] + P.r2[i];
        }
        st_global_v8f(d.out_f32 + f32_off(d, true, d.out_f32_stride, n, y, x, d.out_f32_choff + co0), v);
        st_global_v8f(d.out_f32 + f32_off(d, true, d.out_f32_stride, n, y, x, d.out_f32_choff + co0 + 8), v + 8);
        if (d.out_bf16 != nullptr && !((d.no_bf16_tiles >> ct) & 1)) {
            uint32_t pk[8];
            if constexpr (MODE == kEpiDx0Mask) {
                const uint16_t* mv = reinterpret_cast<const uint16_t*>(P.m);   // raw bits: bf16 or fp16 activations
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float s0 = ((mv[2 * i] & 0x8000u) == 0 && (mv[2 * i] & 0x7fffu) != 0) ? 1.f : d.slope;
                    const float s1 = ((mv[2 * i + 1] & 0x8000u) == 0 && (mv[2 * i + 1] & 0x7fffu) != 0) ? 1.f : d.slope;
                    pk[i] = pack_bf16x2(d.out_bf16_scale * v[2 * i] * s0, d.out_bf16_scale * v[2 * i + 1] * s1);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(d.out_bf16_scale * v[2 * i], d.out_bf16_scale * v[2 * i + 1]);
            }
            st_global_v8(reinterpret_cast<__nv_bfloat16*>(d.out_bf16) + pix * d.out_bf16_stride + d.out_bf16_choff + co0, pk);
        }
        return;
    }
    if constexpr (MODE == kEpiMask) {
        const uint16_t* mv = reinterpret_cast<const uint16_t*>(P.m);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float s0 = ((mv[2 * i] & 0x8000u) == 0 && (mv[2 * i] & 0x7fffu) != 0) ? 1.f : d.slope;
            const float s1 = ((mv[2 * i + 1] & 0x8000u) == 0 && (mv[2 * i + 1] & 0x7fffu) != 0) ? 1.f : d.slope;
            pk[i] = pack_bf16x2(v[2 * i] * s0, v[2 * i + 1] * s1);
        }
        st_global_v8(reinterpret_cast<__nv_bfloat16*>(d.out_bf16) + pix * d.out_bf16_stride + d.out_bf16_choff + co0, pk);
        return;
    }
    if constexpr (MODE == kEpiNchw) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int co = co0 + i;
            if (co < d.cout_real) d.out_nchw[((static_cast<size_t>(n) * d.cout_real + co) * d.H + y) * d.W + x] = v[i];
        }
        return;
    }
    const uint32_t flags = tile_flags(d, ct);
    if (flags & ESR_EPI_ACCUM) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += P.r1[i];
        if (flags & ESR_EPI_RES1) {
            float r[16];
            load16f_at(d, d.res1, d.res1_stride, d.res1_choff, n, y, x, co0, r);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = d.alpha * v[i] + d.gamma * r[i];
        }
    } else {
        if (flags & ESR_EPI_LRELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], d.slope * v[i]);
        }
        if (flags & ESR_EPI_RES1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = d.alpha * v[i] + d.gamma * P.r1[i];
        }
    }
    if (flags & ESR_EPI_RES2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = d.beta * v[i] + P.r2[i];
    }
    if (d.out_f32 != nullptr) store16f_at(d, d.out_f32, d.out_f32_stride, d.out_f32_choff, n, y, x, co0, v);
    if (d.out_lo != nullptr) store16_lo(d, pix, co0, v);
    if (d.out_nchw != nullptr) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int co = co0 + i;
            if (co < d.cout_real)
                d.out_nchw[((static_cast<size_t>(n) * d.cout_real + co) * d.H + y) * d.W + x] = v[i];
        }
    }
    if (d.out_bf16 != nullptr && !((d.no_bf16_tiles >> ct) & 1)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= d.out_bf16_scale;
        if (flags & ESR_EPI_MASK) {
            const uint16_t* mv = reinterpret_cast<const uint16_t*>(P.m);   // raw bits: bf16 or fp16 activations
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= ((mv[i] & 0x8000u) == 0 && (mv[i] & 0x7fffu) != 0) ? 1.f : d.slope;
        }
        uint32_t pk[8];
        if (flags & ESR_EPI_OUT_F16) {
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_f16x2_sat(v[2 * i], v[2 * i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        }
        const uint4 h0 = make_uint4(pk[0], pk[1], pk[2], pk[3]), h1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        const bool want_lo = d.out_bf16_lo_choff >= 0;
        uint4 l0 = h0, l1 = h1;
        if (want_lo) {
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(bf16_lo_part(v[2 * i]), bf16_lo_part(v[2 * i + 1]));
            l0 = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            l1 = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        const int up = d.up;
        const size_t ow = static_cast<size_t>(d.W) * up;
        __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(d.out_bf16);
        for (int a = 0; a < up; ++a) {
            for (int bb = 0; bb < up; ++bb) {
                const size_t opix = (static_cast<size_t>(n) * d.H * up + static_cast<size_t>(y) * up + a) * ow +
                                    static_cast<size_t>(x) * up + bb;
                uint4* o = reinterpret_cast<uint4*>(ob + opix * d.out_bf16_stride + d.out_bf16_choff + co0);
                o[0] = h0;
                o[1] = h1;
                if (want_lo) {
                    uint4* ol = reinterpret_cast<uint4*>(ob + opix * d.out_bf16_stride + d.out_bf16_lo_choff + co0);
                    ol[0] = l0;
                    ol[1] = l1;
                }
            }
        }
    }
}

// Picks the cheapest epilogue specialisation that implements the descriptor exactly.
inline int classify_epilogue(const esr_conv_desc& d) {
    const uint32_t f = d.flags & ~static_cast<uint32_t>(ESR_EPI_WIDE_OK | ESR_CONV_F16);
    const auto al32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
    bool routed = d.no_accum_tiles || d.no_bf16_tiles || d.no_res_tiles;
    for (int t = 0; t < d.cout_tiles; ++t) routed = routed || d.tile_choff[t] >= 0;
    // (a plain f32-only output - the dgrads of the upconvs - qualifies too: no other specialisation writes out_f32 alone)
    const bool f32_only = d.out_bf16 == nullptr && (f & ~static_cast<uint32_t>(ESR_EPI_F32_BLOCKED)) == 0;
    const bool with_mask = (f & ESR_EPI_MASK) != 0 && d.mask != nullptr && d.out_bf16 != nullptr && d.mask_stride % 8 == 0 &&
                           d.mask_choff % 8 == 0;
    if ((f & ESR_EPI_F32_BLOCKED) &&
        (f & ~static_cast<uint32_t>(ESR_EPI_RES1 | ESR_EPI_RES2 | ESR_EPI_ACCUM | ESR_EPI_F32_BLOCKED | (with_mask ? ESR_EPI_MASK : 0))) == 0 &&
        (routed || (f & ESR_EPI_ACCUM) || f32_only) && d.out_f32 != nullptr && d.out_nchw == nullptr && d.up == 1 && d.cout_tile == 32 &&
        d.out_lo == nullptr &&
        (d.out_bf16 == nullptr || (d.out_bf16_lo_choff < 0 && d.out_bf16_stride % 16 == 0 && d.out_bf16_choff % 16 == 0 &&
                                   al32(d.out_bf16)))) {
        bool ok = true;                    // a tile either accumulates or takes residuals, and routed tiles stay 16-aligned
        for (int t = 0; t < d.cout_tiles; ++t) {
            const bool acc = (f & ESR_EPI_ACCUM) && !((d.no_accum_tiles >> t) & 1);
            const bool res = (f & (ESR_EPI_RES1 | ESR_EPI_RES2)) && !((d.no_res_tiles >> t) & 1);
            ok = ok && !(acc && res) && (d.tile_choff[t] < 0 || d.tile_choff[t] % 8 == 0);
        }
        if (ok) return with_mask ? kEpiDx0Mask : kEpiDx0;
    }
    if (routed) return kEpiGeneric;
    const bool bf_ok = d.out_bf16 != nullptr && d.out_bf16_lo_choff < 0 && d.out_bf16_scale == 1.0f && d.out_bf16_stride % 16 == 0 &&
                       d.out_bf16_choff % 16 == 0 && al32(d.out_bf16);
    if ((f & ~static_cast<uint32_t>(ESR_EPI_F32_BLOCKED)) == ESR_EPI_LRELU && bf_ok && d.out_f32 == nullptr && d.out_nchw == nullptr &&
        d.up == 1 && d.cout_tile == 32)
        return kEpiTrunk;
    if ((f & ~static_cast<uint32_t>(ESR_EPI_RES2 | ESR_EPI_OUT_F16 | ESR_EPI_RES1_HILO)) == (ESR_EPI_RES1 | ESR_EPI_F32_BLOCKED) && bf_ok &&
        d.up == 1 && (d.out_f32 != nullptr || d.out_lo != nullptr) && d.out_nchw == nullptr &&
        ((f & ESR_EPI_RES1_HILO) || d.res1 != nullptr) && (!(f & ESR_EPI_RES2) || d.res2 != nullptr))
        return kEpiRes;      // blocked-layout / pair alignment was validated (validate_conv_desc)
    if ((f & ~static_cast<uint32_t>(ESR_EPI_LRELU | ESR_EPI_OUT_F16 | ESR_EPI_F32_BLOCKED)) == 0 && bf_ok && d.out_f32 == nullptr &&
        d.out_nchw == nullptr)
        return kEpiAct;
    if ((f & ~static_cast<uint32_t>(ESR_EPI_F32_BLOCKED)) == 0 && d.out_nchw != nullptr && d.out_bf16 == nullptr && d.out_f32 == nullptr)
        return kEpiNchw;
    if ((f & ~static_cast<uint32_t>(ESR_EPI_F32_BLOCKED)) == ESR_EPI_MASK && bf_ok && d.out_f32 == nullptr && d.out_nchw == nullptr &&
        d.up == 1 && d.cout_tile == 32 && d.mask != nullptr && d.mask_stride % 8 == 0 && d.mask_choff % 8 == 0)
        return kEpiMask;
    static const bool no_outer = []() { const char* v = getenv("ESR_NO_EPI_OUTER"); return v && atoi(v); }();   // A/B aid
    if (!no_outer && d.pair && d.cout_tile == 64 && (f & ESR_EPI_F32_BLOCKED) && (f & ~static_cast<uint32_t>(ESR_EPI_RES1 | ESR_EPI_OUT_F16 | ESR_EPI_F32_BLOCKED)) == 0 &&
        bf_ok && d.out_nchw == nullptr && d.out_lo == nullptr && (d.up == 1 || d.up == 2) && (!(f & ESR_EPI_RES1) || d.res1 != nullptr))
        return kEpiOuter;
    return kEpiGeneric;
}

}  // namespace esr
