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
    int tiles_x, tiles_y, total_tiles;
};

// Byte offset of element (row n, channel k) inside a [rows x 32ch] SWIZZLE_64B K-major
// slab: 16-byte chunk index is XORed with bits [7,9) of the row offset.
__host__ __device__ inline uint32_t sw64_offset(uint32_t n, uint32_t k) {
    return n * 64u + ((((k >> 3) ^ (n >> 1)) & 3u) << 4) + (k & 7u) * 2u;
}

// Applies the fused epilogue to 16 consecutive output channels of one pixel.
__device__ __forceinline__ void conv_epilogue16(const esr_conv_desc& d, int n, int y, int x, int co0, float (&v)[16]) {
    const size_t pix = (static_cast<size_t>(n) * d.H + y) * d.W + x;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(d.bias + co0 + i));
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
    if (d.flags & ESR_EPI_ACCUM) {
        const float* o = d.out_f32 + pix * d.out_f32_stride + d.out_f32_choff + co0;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const float4 r = *reinterpret_cast<const float4*>(o + i);
            v[i] += r.x; v[i + 1] += r.y; v[i + 2] += r.z; v[i + 3] += r.w;
        }
    }
    if (d.flags & ESR_EPI_LRELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : d.slope * v[i];
    }
    if (d.flags & ESR_EPI_RES1) {
        const float* r1 = d.res1 + pix * d.res1_stride + d.res1_choff + co0;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(r1 + i));
            v[i] = d.alpha * v[i] + r.x; v[i + 1] = d.alpha * v[i + 1] + r.y;
            v[i + 2] = d.alpha * v[i + 2] + r.z; v[i + 3] = d.alpha * v[i + 3] + r.w;
        }
    }
    if (d.flags & ESR_EPI_RES2) {
        const float* r2 = d.res2 + pix * d.res2_stride + d.res2_choff + co0;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(r2 + i));
            v[i] = d.beta * v[i] + r.x; v[i + 1] = d.beta * v[i + 1] + r.y;
            v[i + 2] = d.beta * v[i + 2] + r.z; v[i + 3] = d.beta * v[i + 3] + r.w;
        }
    }
    if (d.out_f32 != nullptr) {
        float* o = d.out_f32 + pix * d.out_f32_stride + d.out_f32_choff + co0;
#pragma unroll
        for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    if (d.out_nchw != nullptr) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int co = co0 + i;
            if (co < d.cout_real)
                d.out_nchw[((static_cast<size_t>(n) * d.cout_real + co) * d.H + y) * d.W + x] = v[i];
        }
    }
    if (d.out_bf16 != nullptr) {
        float w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = d.out_bf16_scale * v[i];
        if (d.flags & ESR_EPI_MASK) {
            const __nv_bfloat16* m = reinterpret_cast<const __nv_bfloat16*>(d.mask) + pix * d.mask_stride +
                                     d.mask_choff + co0;
            uint4 raw[2];
            raw[0] = __ldg(reinterpret_cast<const uint4*>(m));
            raw[1] = __ldg(reinterpret_cast<const uint4*>(m) + 1);
            const __nv_bfloat16* mv = reinterpret_cast<const __nv_bfloat16*>(raw);
#pragma unroll
            for (int i = 0; i < 16; ++i) w[i] *= (__bfloat162float(mv[i]) > 0.f ? 1.f : d.slope);
        }
        __align__(16) __nv_bfloat16 hi[16];
        __align__(16) __nv_bfloat16 lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            hi[i] = __float2bfloat16_rn(w[i]);
            lo[i] = __float2bfloat16_rn(w[i] - __bfloat162float(hi[i]));
        }
        const int up = d.up;
        const size_t ow = static_cast<size_t>(d.W) * up;
        __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(d.out_bf16);
        for (int a = 0; a < up; ++a) {
            for (int b = 0; b < up; ++b) {
                const size_t opix = (static_cast<size_t>(n) * d.H * up + static_cast<size_t>(y) * up + a) * ow +
                                    static_cast<size_t>(x) * up + b;
                uint4* o = reinterpret_cast<uint4*>(ob + opix * d.out_bf16_stride + d.out_bf16_choff + co0);
                o[0] = reinterpret_cast<const uint4*>(hi)[0];
                o[1] = reinterpret_cast<const uint4*>(hi)[1];
                if (d.out_bf16_lo_choff >= 0) {
                    uint4* ol = reinterpret_cast<uint4*>(ob + opix * d.out_bf16_stride + d.out_bf16_lo_choff + co0);
                    ol[0] = reinterpret_cast<const uint4*>(lo)[0];
                    ol[1] = reinterpret_cast<const uint4*>(lo)[1];
                }
            }
        }
    }
}

}  // namespace esr
