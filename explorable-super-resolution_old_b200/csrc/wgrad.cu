// Weight gradients of the generator's 3x3 convolutions (the GAN training step's generator half,
// codes/models/SRRaGAN_model.py:463-547: l_g_total.backward() with trainable G parameters).
//   dW[co, ci, ky, kx] = sum_{n,y,x} g[n, y, x, co] * X[n, y+ky-1, x+kx-1, ci]      (zero outside the image)
//   db[co]             = sum_{n,y,x} g[n, y, x, co]
// where X is the conv's input as the forward left it in HBM (16-bit NHWC slices of the dense-block / outer buffers,
// plus the small fp32 NCHW latent / LR planes) and g the gradient w.r.t. its pre-activation output as the data-gradient
// backward left it (bf16 NHWC).  Two kernels:
//   wgrad16_kernel    one CTA per (conv, block of 16 input channels): K = pixels, warp-level bf16 MMAs (mma.sync m16n8k16
//                     fed by ldmatrix.trans, fp32 accumulate) on shared-memory tiles; every CTA owns its slice of dW for all pixels (no atomics,
//                     deterministic) unless the item is a spatial chunk of a high-resolution conv (atomicAdd then).
//   wgrad_small_kernel  the <= 8 fp32 NCHW input channels (latent, LR image) and the bias: lanes = output channels.
// First version: legacy mma.sync tensor path (HMMA), not tcgen05 - DESIGN.md lists the tcgen05 form (MN-major
// operands straight from the NHWC tiles) as the next step.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "esr_common.cuh"

namespace esr {

constexpr int kWgTH = 8, kWgTW = 16;               // spatial tile: 8 rows x 16 pixels (one k-step per row)
constexpr int kWgThreads = 256;
constexpr int kWgXPitch = 16;                      // input channels per item
constexpr int kWgMaxCout = 64;

// ldmatrix.trans: both operands are stored pixel-major ([k = pixel][m or n = channel]) and the MMA wants K innermost, so
// every 8x8 block is transposed on the way into the fragment.  (The first version went through wmma::load_matrix_sync:
// generic LD.E loads + MOVM register transposes, 49 instructions per MMA; this form is 2 LDSM + 2 HMMA.)
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_row))));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t n = valid ? 16u : 0u;                            // src-size 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc), "r"(n)
                 : "memory");
}

constexpr int kWgXsElems = (kWgTH + 2) * (kWgTW + 2) * kWgXPitch;   // input halo tile (bf16 elements)
constexpr int kWgGsElems = kWgTH * kWgTW * (kWgMaxCout + 16);       // gradient tile
constexpr int kWgStageElems = kWgXsElems + kWgGsElems;
constexpr int kWgSmemBytes = 2 * kWgStageElems * 2;                 // two stages: the next tile lands while this one is multiplied

// Round-2b: the tiles come in through a two-stage cp.async pipeline.  The first version staged every tile with blocking
// loads between two __syncthreads(): ~100 MMA-side instructions per warp behind ~1 us of exposed load latency per tile,
// 13.5 ms of a 27 ms generator step.
__global__ void __launch_bounds__(kWgThreads, 3) wgrad16_kernel(const esr_wgrad_item* __restrict__ items) {
    const esr_wgrad_item it = items[blockIdx.x];
    extern __shared__ __align__(128) unsigned char wg_smem[];
    __nv_bfloat16* const stage0 = reinterpret_cast<__nv_bfloat16*>(wg_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncb = it.cout >> 4;                  // 16-channel blocks of the output
    const int nfrag = 9 * ncb;                     // (tap, co block) accumulators, round-robin over the 8 warps
    const int gp = it.cout + 16;                   // pixel pitch of the g tile (elements): rows stay 32-byte aligned
    constexpr int kMaxFr = (9 * (kWgMaxCout / 16) + 7) / 8;
    float acc[kMaxFr][2][4];                       // [fragment][ci half (n8)][mma accumulator]
#pragma unroll
    for (int f = 0; f < kMaxFr; ++f)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[f][q >> 2][q & 3] = 0.f;
    float bsum = 0.f;                              // bias gradient: thread (co = tid % 64, quarter = tid / 64) sums every 4th pixel
    const int b_co = threadIdx.x & 63, b_part = threadIdx.x >> 6;
    const bool do_bias = it.db != nullptr && b_co < it.cout;
    const int tiles_x = (it.W + kWgTW - 1) / kWgTW, tiles_y = (it.H + kWgTH - 1) / kWgTH;
    const int tiles_img = tiles_x * tiles_y;
    const int t_end = it.tile_end > 0 ? it.tile_end : it.B * tiles_img;
    const uint16_t* xg = static_cast<const uint16_t*>(it.x);
    const uint16_t* gg = static_cast<const uint16_t*>(it.g);
    const bool fix_x = it.x_f16 || it.n_ci < 16 || it.ci_lo > 0;    // the input chunks need a pass after they landed (CTA-uniform)
    const int g_vec = it.cout >> 3;                // uint4 per pixel of the gradient tile
    // ldmatrix row of this lane inside a 16 (k) x 16 (m or n) operand block: matrices 0..3 = lanes 0-7, 8-15, 16-23, 24-31
    const int lj = lane >> 3, lr = lane & 7;
    const int a_k = (lj >> 1) * 8 + lr, a_m = (lj & 1) * 8;        // A: (k0,m0) = (0,0), (0,8), (8,0), (8,8)  -> a0..a3
    const int b_k = (lj & 1) * 8 + lr, b_n = (lj >> 1) * 8;        // B: (k0,n0) = (0,0), (8,0), (0,8), (8,8)  -> b0,b1 | b0,b1

    // issues the copies of tile t into stage `buf` (zero outside the image = the conv's zero padding); one commit group
    auto prefetch = [&](int t, int buf) {
        __nv_bfloat16* Xs = stage0 + buf * kWgStageElems;
        __nv_bfloat16* Gs = Xs + kWgXsElems;
        const int n = t / tiles_img, r = t - n * tiles_img;
        const int y0 = (r / tiles_x) * kWgTH, x0 = (r % tiles_x) * kWgTW;
        for (int idx = threadIdx.x; idx < (kWgTH + 2) * (kWgTW + 2) * 2; idx += kWgThreads) {
            const int half = idx & 1, p = idx >> 1;
            const int yy = y0 - 1 + p / (kWgTW + 2), xx = x0 - 1 + p % (kWgTW + 2);
            const bool ok = yy >= 0 && yy < it.H && xx >= 0 && xx < it.W;
            const uint16_t* src = ok ? xg + (static_cast<size_t>(n) * it.H + yy) * it.W * it.x_stride + static_cast<size_t>(xx) * it.x_stride +
                                            it.x_c0 + half * 8
                                     : xg;
            cp_async16_zfill(Xs + p * kWgXPitch + half * 8, src, ok);
        }
        for (int idx = threadIdx.x; idx < kWgTH * kWgTW * g_vec; idx += kWgThreads) {
            const int p = idx / g_vec, q = idx - p * g_vec;
            const int yy = y0 + p / kWgTW, xx = x0 + p % kWgTW;
            const bool ok = yy < it.H && xx < it.W;
            const uint16_t* src = ok ? gg + (static_cast<size_t>(n) * it.H + yy) * it.W * it.g_stride + static_cast<size_t>(xx) * it.g_stride +
                                            it.g_c0 + q * 8
                                     : gg;
            cp_async16_zfill(Gs + p * gp + q * 8, src, ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (it.tile_begin < t_end) prefetch(it.tile_begin, 0);
    for (int t = it.tile_begin; t < t_end; ++t) {
        const int buf = (t - it.tile_begin) & 1;
        __nv_bfloat16* Xs = stage0 + buf * kWgStageElems;
        __nv_bfloat16* Gs = Xs + kWgXsElems;
        if (t + 1 < t_end) {
            prefetch(t + 1, buf ^ 1);              // the other stage was released by the barrier that ended tile t-1
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (fix_x) {
            // every thread post-processes the input chunks IT copied (its own cp.async results are visible to it after the wait)
            for (int idx = threadIdx.x; idx < (kWgTH + 2) * (kWgTW + 2) * 2; idx += kWgThreads) {
                const int half = idx & 1, p = idx >> 1;
                uint4 v = *reinterpret_cast<const uint4*>(Xs + p * kWgXPitch + half * 8);
                if (it.x_f16) {                    // fp16 activations of the outer convs -> bf16 operands
                    const __half2* h = reinterpret_cast<const __half2*>(&v);
                    __nv_bfloat162 o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) o[k] = __float22bfloat162_rn(__half22float2(h[k]));
                    v = *reinterpret_cast<const uint4*>(o);
                }
                if (it.n_ci < 16 || it.ci_lo > 0) { // channels outside [ci_lo, n_ci) belong to someone else: contribute nothing
                    uint16_t* e = reinterpret_cast<uint16_t*>(&v);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (half * 8 + k >= it.n_ci || half * 8 + k < it.ci_lo) e[k] = 0;
                }
                *reinterpret_cast<uint4*>(Xs + p * kWgXPitch + half * 8) = v;
            }
        }
        __syncthreads();
        // ---- K = the tile's 128 pixels, 16 per step (one tile row): D[co, ci] += sum_px g[px, co] * X[px + tap, ci]
#pragma unroll
        for (int f = 0; f < kMaxFr; ++f) {
            const int fr = warp + f * 8;
            if (fr < nfrag) {
                const int tap = fr / ncb, cb = fr - tap * ncb;
                const int ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
                for (int y = 0; y < kWgTH; ++y) {
                    uint32_t a[4], b[4];
                    ldsm_x4_trans(a, Gs + (y * kWgTW + a_k) * gp + cb * 16 + a_m);
                    ldsm_x4_trans(b, Xs + ((y + ky) * (kWgTW + 2) + kx + b_k) * kWgXPitch + b_n);
                    mma_bf16_16816(acc[f][0], a, b[0], b[1]);          // ci 0..7
                    mma_bf16_16816(acc[f][1], a, b[2], b[3]);          // ci 8..15
                }
            }
        }
        if (do_bias) {                             // pixels outside the image were zero-filled
            const uint16_t* gs16 = reinterpret_cast<const uint16_t*>(Gs) + b_co;
#pragma unroll 8
            for (int p = b_part; p < kWgTH * kWgTW; p += 4) bsum += __uint_as_float(static_cast<uint32_t>(gs16[p * gp]) << 16);
        }
        __syncthreads();
    }
    if (it.db != nullptr) {                        // CTA-uniform; the stages are free after the last barrier
        float* red = reinterpret_cast<float*>(wg_smem);
        red[threadIdx.x] = do_bias ? bsum : 0.f;
        __syncthreads();
        if (threadIdx.x < it.n_co) {
            const float v = (red[threadIdx.x] + red[64 + threadIdx.x]) + (red[128 + threadIdx.x] + red[192 + threadIdx.x]);
            if (it.tile_end > 0) atomicAdd(it.db + threadIdx.x, v);
            else it.db[threadIdx.x] = v;
        }
    }
    // ---- dW[co, ci0 + ci, ky, kx]: accumulator (row g / g+8, cols 2t, 2t+1) of each n8 half
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int f = 0; f < kMaxFr; ++f) {
        const int fr = warp + f * 8;
        if (fr < nfrag) {
            const int tap = fr / ncb, cb = fr - tap * ncb;
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int co = cb * 16 + gq + (e >> 1) * 8, ci = hlf * 8 + 2 * tq + (e & 1);
                    if (co < it.n_co && ci < it.n_ci && ci >= it.ci_lo) {
                        float* dst = it.dw + (static_cast<size_t>(co) * it.cin_total + it.ci0 + ci - it.ci_lo) * 9 + tap;
                        if (it.tile_end > 0) atomicAdd(dst, acc[f][hlf][e]);
                        else *dst = acc[f][hlf][e];
                    }
                }
        }
    }
}

// Round-2c: K split over the warps.  Warp w owns tile row y = w (one 16-pixel k-step) for ALL nine taps and both 16-channel
// output blocks: per tile it loads 2 gradient fragments and 9 input fragments (11 ldmatrix.x4) for 36 MMAs, where the
// fragment-per-warp mapping above loads 2 per 2 MMAs (the shared-memory pipe, not the tensor pipe, bounded it).  Items
// stage at most 32 output channels (the host cuts the 64-channel convs in two), so the 9 x 2 x 2 accumulator fragments are
// 144 registers; the eight warps' partial sums meet in shared memory once per item.  Four cp.async stages: a tile is now
// ~50 instructions per warp, less than the L2 latency.
constexpr int kWgRStages = 4;
constexpr int kWgRGsElems = kWgTH * kWgTW * (32 + 16);
constexpr int kWgRStageElems = kWgXsElems + kWgRGsElems;
constexpr int kWgRSmemBytes = kWgRStages * kWgRStageElems * 2;

__global__ void __launch_bounds__(kWgThreads, 1) wgrad16r_kernel(const esr_wgrad_item* __restrict__ items) {
    const esr_wgrad_item it = items[blockIdx.x];
    extern __shared__ __align__(128) unsigned char wg_smem[];
    __nv_bfloat16* const stage0 = reinterpret_cast<__nv_bfloat16*>(wg_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ncb = it.cout >> 4;                  // 1 or 2
    const int gp = it.cout + 16;
    float acc[9][2][2][4];                         // [tap][co block][ci half][mma accumulator]
#pragma unroll
    for (int t9 = 0; t9 < 9; ++t9)
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[t9][q >> 3][(q >> 2) & 1][q & 3] = 0.f;
    float bsum = 0.f;
    const int b_co = threadIdx.x & 63, b_part = threadIdx.x >> 6;
    const bool do_bias = it.db != nullptr && b_co < it.cout;
    const int tiles_x = (it.W + kWgTW - 1) / kWgTW, tiles_y = (it.H + kWgTH - 1) / kWgTH;
    const int tiles_img = tiles_x * tiles_y;
    const int t_end = it.tile_end > 0 ? it.tile_end : it.B * tiles_img;
    const uint16_t* xg = static_cast<const uint16_t*>(it.x);
    const uint16_t* gg = static_cast<const uint16_t*>(it.g);
    const bool fix_x = it.x_f16 || it.n_ci < 16 || it.ci_lo > 0;
    const int g_vec = it.cout >> 3;
    const int lj = lane >> 3, lr = lane & 7;
    const int a_k = (lj >> 1) * 8 + lr, a_m = (lj & 1) * 8;
    const int b_k = (lj & 1) * 8 + lr, b_n = (lj >> 1) * 8;

    auto prefetch = [&](int t, int buf) {
        __nv_bfloat16* Xs = stage0 + buf * kWgRStageElems;
        __nv_bfloat16* Gs = Xs + kWgXsElems;
        const int n = t / tiles_img, r = t - n * tiles_img;
        const int y0 = (r / tiles_x) * kWgTH, x0 = (r % tiles_x) * kWgTW;
        for (int idx = threadIdx.x; idx < (kWgTH + 2) * (kWgTW + 2) * 2; idx += kWgThreads) {
            const int half = idx & 1, p = idx >> 1;
            const int yy = y0 - 1 + p / (kWgTW + 2), xx = x0 - 1 + p % (kWgTW + 2);
            const bool ok = yy >= 0 && yy < it.H && xx >= 0 && xx < it.W;
            const uint16_t* src = ok ? xg + (static_cast<size_t>(n) * it.H + yy) * it.W * it.x_stride + static_cast<size_t>(xx) * it.x_stride +
                                            it.x_c0 + half * 8
                                     : xg;
            cp_async16_zfill(Xs + p * kWgXPitch + half * 8, src, ok);
        }
        for (int idx = threadIdx.x; idx < kWgTH * kWgTW * g_vec; idx += kWgThreads) {
            const int p = idx / g_vec, q = idx - p * g_vec;
            const int yy = y0 + p / kWgTW, xx = x0 + p % kWgTW;
            const bool ok = yy < it.H && xx < it.W;
            const uint16_t* src = ok ? gg + (static_cast<size_t>(n) * it.H + yy) * it.W * it.g_stride + static_cast<size_t>(xx) * it.g_stride +
                                            it.g_c0 + q * 8
                                     : gg;
            cp_async16_zfill(Gs + p * gp + q * 8, src, ok);
        }
    };
    // prologue: kWgRStages - 1 tiles in flight; one commit group per tile slot (empty groups keep the count uniform)
#pragma unroll
    for (int k = 0; k < kWgRStages - 1; ++k) {
        if (it.tile_begin + k < t_end) prefetch(it.tile_begin + k, k);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int t = it.tile_begin; t < t_end; ++t) {
        const int buf = (t - it.tile_begin) % kWgRStages;
        __nv_bfloat16* Xs = stage0 + buf * kWgRStageElems;
        __nv_bfloat16* Gs = Xs + kWgXsElems;
        // the stage of tile t + S - 1 was read during tile t - 1: released by the barrier that ended it
        if (t + kWgRStages - 1 < t_end) prefetch(t + kWgRStages - 1, (buf + kWgRStages - 1) % kWgRStages);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(kWgRStages - 1) : "memory");
        if (fix_x) {
            for (int idx = threadIdx.x; idx < (kWgTH + 2) * (kWgTW + 2) * 2; idx += kWgThreads) {
                const int half = idx & 1, p = idx >> 1;
                uint4 v = *reinterpret_cast<const uint4*>(Xs + p * kWgXPitch + half * 8);
                if (it.x_f16) {
                    const __half2* h = reinterpret_cast<const __half2*>(&v);
                    __nv_bfloat162 o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) o[k] = __float22bfloat162_rn(__half22float2(h[k]));
                    v = *reinterpret_cast<const uint4*>(o);
                }
                if (it.n_ci < 16 || it.ci_lo > 0) {
                    uint16_t* e = reinterpret_cast<uint16_t*>(&v);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (half * 8 + k >= it.n_ci || half * 8 + k < it.ci_lo) e[k] = 0;
                }
                *reinterpret_cast<uint4*>(Xs + p * kWgXPitch + half * 8) = v;
            }
        }
        __syncthreads();
        {
            const int y = warp;                    // this warp's k-step: the 16 pixels of tile row y
            const bool wide = it.n_ci > 8;         // the latent blocks use channels [3, 6) only: half the MMAs (CTA-uniform)
            uint32_t a[2][4];
            ldsm_x4_trans(a[0], Gs + (y * kWgTW + a_k) * gp + a_m);
            if (ncb > 1) ldsm_x4_trans(a[1], Gs + (y * kWgTW + a_k) * gp + 16 + a_m);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    uint32_t b[4];
                    ldsm_x4_trans(b, Xs + ((y + ky) * (kWgTW + 2) + kx + b_k) * kWgXPitch + b_n);
                    mma_bf16_16816(acc[ky * 3 + kx][0][0], a[0], b[0], b[1]);
                    if (wide) mma_bf16_16816(acc[ky * 3 + kx][0][1], a[0], b[2], b[3]);     // input channels 8..15
                    if (ncb > 1) {
                        mma_bf16_16816(acc[ky * 3 + kx][1][0], a[1], b[0], b[1]);
                        if (wide) mma_bf16_16816(acc[ky * 3 + kx][1][1], a[1], b[2], b[3]);
                    }
                }
        }
        if (do_bias) {
            const uint16_t* gs16 = reinterpret_cast<const uint16_t*>(Gs) + b_co;
#pragma unroll 8
            for (int p = b_part; p < kWgTH * kWgTW; p += 4) bsum += __uint_as_float(static_cast<uint32_t>(gs16[p * gp]) << 16);
        }
        __syncthreads();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float* red = reinterpret_cast<float*>(wg_smem);                    // [8 warps][32 lanes][16]: 16 KiB, one tap at a time
    if (it.db != nullptr) {
        red[threadIdx.x] = do_bias ? bsum : 0.f;
        __syncthreads();
        if (threadIdx.x < it.n_co) {
            const float v = (red[threadIdx.x] + red[64 + threadIdx.x]) + (red[128 + threadIdx.x] + red[192 + threadIdx.x]);
            if (it.tile_end > 0) atomicAdd(it.db + threadIdx.x, v);
            else it.db[threadIdx.x] = v;
        }
        __syncthreads();
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int q = 0; q < 16; ++q) red[(warp * 16 + q) * 32 + lane] = acc[tap][q >> 3][(q >> 2) & 1][q & 3];
        __syncthreads();
        // 512 outputs (q, lane) per tap, two per thread, summed over the warps in a fixed order
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int o = threadIdx.x + k * kWgThreads, q = o >> 5, ln = o & 31;
            float v = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) v += red[(w8 * 16 + q) * 32 + ln];
            const int cb = q >> 3, hlf = (q >> 2) & 1, e = q & 3;
            const int co = cb * 16 + (ln >> 2) + (e >> 1) * 8, ci = hlf * 8 + 2 * (ln & 3) + (e & 1);
            if (cb < ncb && co < it.n_co && ci < it.n_ci && ci >= it.ci_lo) {
                float* dst = it.dw + (static_cast<size_t>(co) * it.cin_total + it.ci0 + ci - it.ci_lo) * 9 + tap;
                if (it.tile_end > 0) atomicAdd(dst, v);
                else *dst = v;
            }
        }
        __syncthreads();
    }
}

// grid = (items, row chunks): a CTA sums over `rows_per_cta` image rows and adds its partial sums with atomicAdd (dW / db
// zeroed by the caller); one CTA per conv over all pixels took 71 ms on the 128x128 convs of config 5
__global__ void __launch_bounds__(256) wgrad_small_kernel(const esr_wgrad_small_item* __restrict__ items, int rows_per_cta) {
    const esr_wgrad_small_item it = items[blockIdx.x];
    const int row_lo = blockIdx.y * rows_per_cta, row_hi = min(row_lo + rows_per_cta, it.B * it.H);
    if (row_lo >= row_hi) return;
    constexpr int kMaxC = 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int co_blocks = (it.cout + 31) >> 5;
    __shared__ float red[8][32];
    const uint16_t* gg = static_cast<const uint16_t*>(it.g);
    const size_t plane = static_cast<size_t>(it.H) * it.W;
    for (int cb = 0; cb < co_blocks; ++cb) {
        const int co = cb * 32 + lane;
        const bool live = co < it.cout;
        float acc[kMaxC * 9 + 1];
#pragma unroll
        for (int k = 0; k < kMaxC * 9 + 1; ++k) acc[k] = 0.f;
        for (int rw = row_lo + warp; rw < row_hi; rw += 8) {       // a warp per image row
            const int n = rw / it.H, y = rw - n * it.H;
            for (int x = 0; x < it.W; ++x) {
                float gv = 0.f;
                if (live) {
                    const uint16_t bits = gg[(static_cast<size_t>(rw) * it.W + x) * it.g_stride + it.g_c0 + co];
                    gv = __uint_as_float(static_cast<uint32_t>(bits) << 16);
                }
                // bias gradient; the last conv's comes from the fp32 gradient planes (its true value is ~0 after the CEM
                // adjoint removed the DC component: a sum of bf16-rounded terms would be pure rounding noise)
                acc[kMaxC * 9] += (it.g32 != nullptr && live && co < it.n_co)
                                      ? it.g32[(static_cast<size_t>(n) * it.n_co + co) * plane + static_cast<size_t>(y) * it.W + x] : gv;
                if (it.n_c > 0) {
#pragma unroll
                    for (int c = 0; c < kMaxC; ++c) {
                        if (c >= it.n_c) break;
                        const float* sp = it.s + (static_cast<size_t>(n) * it.s_channels + it.s_c0 + c) * plane;
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            const int yy = y + ky - 1;
                            if (yy < 0 || yy >= it.H) continue;
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const int xx = x + kx - 1;
                                if (xx < 0 || xx >= it.W) continue;
                                acc[c * 9 + ky * 3 + kx] = fmaf(gv, __ldg(sp + static_cast<size_t>(yy) * it.W + xx), acc[c * 9 + ky * 3 + kx]);
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kMaxC * 9 + 1; ++k) {                  // cross-warp sums, one accumulator at a time (unrolled: acc stays in registers)
            if (k < kMaxC * 9 && k / 9 >= it.n_c) continue;        // block-uniform
            red[warp][lane] = acc[k];
            __syncthreads();
            if (warp == 0) {
                float v = 0.f;
#pragma unroll
                for (int w2 = 0; w2 < 8; ++w2) v += red[w2][lane];
                if (live && co < it.n_co) {
                    if (k == kMaxC * 9) { if (it.db != nullptr) atomicAdd(it.db + co, v); }
                    else atomicAdd(it.dw + (static_cast<size_t>(co) * it.cin_total + it.ci0 + k / 9) * 9 + k % 9, v);
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace esr

using namespace esr;

// Items with at most 32 staged output channels, K split over the warps (wgrad16r_kernel).
extern "C" int esr_wgrad16r(const esr_wgrad_item* items_device, int32_t n_items, void* stream) {
    ESR_CHECK_ARG(items_device != nullptr && n_items > 0, "esr_wgrad16r: bad arguments");
    ESR_ONCE_PER_DEVICE(ESR_CUDA(cudaFuncSetAttribute(wgrad16r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgRSmemBytes)););
    wgrad16r_kernel<<<n_items, kWgThreads, kWgRSmemBytes, static_cast<cudaStream_t>(stream)>>>(items_device);
    return check_launch("wgrad16r_kernel");
}

extern "C" int esr_wgrad16(const esr_wgrad_item* items_device, int32_t n_items, void* stream) {
    ESR_CHECK_ARG(items_device != nullptr && n_items > 0, "esr_wgrad16: bad arguments");
    ESR_ONCE_PER_DEVICE(ESR_CUDA(cudaFuncSetAttribute(wgrad16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes)););
    wgrad16_kernel<<<n_items, kWgThreads, kWgSmemBytes, static_cast<cudaStream_t>(stream)>>>(items_device);
    return check_launch("wgrad16_kernel");
}

extern "C" int esr_wgrad_small(const esr_wgrad_small_item* items_device, int32_t n_items, int32_t max_rows, void* stream) {
    ESR_CHECK_ARG(items_device != nullptr && n_items > 0 && max_rows > 0, "esr_wgrad_small: bad arguments");
    const int rows_per_cta = 32;
    wgrad_small_kernel<<<dim3(n_items, (max_rows + rows_per_cta - 1) / rows_per_cta), 256, 0, static_cast<cudaStream_t>(stream)>>>(items_device, rows_per_cta);
    return check_launch("wgrad_small_kernel");
}
