// CEM (Consistency Enforcing Module) fixed-filter operators, fp32, NCHW planes.
//
//   Down  (H)        : replicate-pad, correlate with the flipped ds kernel, keep phase `pre`
//                      of every sf x sf block                      (CEMnet.py:157-162)
//   InvHTH (K)       : replicate-pad, correlate with inv_hTh       (CEMnet.py:149-151)
//   Up    (~H^T)     : zero-stuff x sf at phase `pre`, replicate-pad, correlate with
//                      sf^2 * ds kernel (polyphase here)           (CEMnet.py:153-159)
//   project          : out = crop(y + Up(K * (x - Down(y))))       (CEMnet.py:183-190)
//
// All three filters of the default (bicubic) configuration are rank-1, so every operator
// runs as a horizontal pass into shared memory followed by a vertical pass.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "cem_tabs.cuh"
#include "esr_common.cuh"
#include "ptx_sm100.cuh"

namespace esr {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ------------------------------------------------------------------------ Down
// Block: DT_R x DT_C LR outputs of one (b,c) plane.
constexpr int DT_R = 8, DT_C = 32;

// out = (x ? x - Down(y) : Down(y)),  y: [P,H,W]  out/x: [P,H/sf,W/sf]   (P = B*C planes)
template <int SF>
__global__ void __launch_bounds__(256) cem_down_kernel(const __grid_constant__ esr_cem_filters f,
                                                       const float* __restrict__ y, const float* __restrict__ x,
                                                       float* __restrict__ out, int H, int W) {
    extern __shared__ float sm[];
    constexpr int sf = SF;
    const int nt = f.n_ds, pad = f.n_ds / 2;
    const int h = H / sf, w = W / sf;
    const int plane = blockIdx.z;
    const int i0 = blockIdx.y * DT_R, j0 = blockIdx.x * DT_C;
    const int rows_in = (DT_R - 1) * sf + nt;          // HR rows feeding the tile
    const int cols_in = (DT_C - 1) * sf + nt;
    float* tile = sm;                                   // [rows_in][cols_in]
    float* hbuf = sm + rows_in * cols_in;               // [rows_in][DT_C]
    const float* yp = y + static_cast<size_t>(plane) * H * W;
    const int r_base = i0 * sf + f.pre - pad, c_base = j0 * sf + f.pre - pad;
    for (int idx = threadIdx.x; idx < rows_in * cols_in; idx += blockDim.x) {
        const int r = idx / cols_in, c = idx - r * cols_in;
        tile[idx] = __ldg(yp + static_cast<size_t>(clampi(r_base + r, 0, H - 1)) * W + clampi(c_base + c, 0, W - 1));
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows_in * DT_C; idx += blockDim.x) {
        const int r = idx / DT_C, j = idx - r * DT_C;
        const float* t = tile + r * cols_in + j * sf;
        float acc = 0.f;
        for (int k = 0; k < nt; ++k) acc = fmaf(f.ds[nt - 1 - k], t[k], acc);
        hbuf[idx] = acc;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < DT_R * DT_C; idx += blockDim.x) {
        const int i = idx / DT_C, j = idx - i * DT_C;
        if (i0 + i >= h || j0 + j >= w) continue;
        float acc = 0.f;
        for (int k = 0; k < nt; ++k) acc = fmaf(f.ds[nt - 1 - k], hbuf[(i * sf + k) * DT_C + j], acc);
        const size_t o = (static_cast<size_t>(plane) * h + i0 + i) * w + j0 + j;
        out[o] = x != nullptr ? x[o] - acc : acc;
    }
}

// ---------------------------------------------------------------------- InvHTH
constexpr int IT_R = 16, IT_C = 32;
__global__ void __launch_bounds__(256) cem_inv_kernel(const __grid_constant__ esr_cem_filters f,
                                                      const float* __restrict__ x, float* __restrict__ out, int h,
                                                      int w) {
    extern __shared__ float sm[];
    const int nt = f.n_inv, pad = nt / 2;
    const int plane = blockIdx.z;
    const int i0 = blockIdx.y * IT_R, j0 = blockIdx.x * IT_C;
    const int rows_in = IT_R + nt - 1, cols_in = IT_C + nt - 1;
    float* tile = sm;
    float* hbuf = sm + rows_in * cols_in;               // [rows_in][IT_C]
    const float* xp = x + static_cast<size_t>(plane) * h * w;
    for (int idx = threadIdx.x; idx < rows_in * cols_in; idx += blockDim.x) {
        const int r = idx / cols_in, c = idx - r * cols_in;
        tile[idx] = __ldg(xp + static_cast<size_t>(clampi(i0 + r - pad, 0, h - 1)) * w + clampi(j0 + c - pad, 0, w - 1));
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows_in * IT_C; idx += blockDim.x) {
        const int r = idx / IT_C, j = idx - r * IT_C;
        const float* t = tile + r * cols_in + j;
        float acc = 0.f;
        for (int k = 0; k < nt; ++k) acc = fmaf(f.inv[k], t[k], acc);
        hbuf[idx] = acc;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < IT_R * IT_C; idx += blockDim.x) {
        const int i = idx / IT_C, j = idx - i * IT_C;
        if (i0 + i >= h || j0 + j >= w) continue;
        float acc = 0.f;
        for (int k = 0; k < nt; ++k) acc = fmaf(f.inv[k], hbuf[(i + k) * IT_C + j], acc);
        out[(static_cast<size_t>(plane) * h + i0 + i) * w + j0 + j] = acc;
    }
}

// -------------------------------------------------------------------------- Up
// out[Y-crop, X-crop] = (y ? y[Y,X] : 0) + sign * sum_{ty,tx} u[ty] u[tx] z[clamp(Y+ty-pad), clamp(X+tx-pad)],
// z = x zero-stuffed at phase pre.  Block: UT_R x UT_C LR cells -> (UT_R*sf) x (UT_C*sf) HR pixels.
// SF is a template parameter (divisions by it become multiplies); away from the border only every SF-th tap meets a
// sample, so the interior walks its phase's taps directly; within pad pixels of the border the replicate padding of
// the zero-stuffed image breaks the phase pattern and every tap is tested.
constexpr int UT_R = 8, UT_C = 32;
template <int SF>
__device__ __forceinline__ float up_taps_1d(const float* __restrict__ u, int nt, int pad, int pre, int P, int L,
                                            const float* __restrict__ src, int src_stride, int cell0) {
    // sum_t u[t] * z[clamp(P + t - pad, 0, L-1)] along one axis; src[(cell - cell0) * src_stride] = sample of LR cell
    float acc = 0.f;
    if (P >= pad && P + (nt - 1 - pad) <= L - 1) {
        const int t0 = ((pre + pad - P) % SF + SF) % SF;
        const float* s = src + ((P + t0 - pad - pre) / SF - cell0) * src_stride;
        for (int t = t0; t < nt; t += SF, s += src_stride) acc = fmaf(u[t] * SF, *s, acc);
    } else {
        for (int t = 0; t < nt; ++t) {
            const int q = clampi(P + t - pad, 0, L - 1) - pre;
            if (q >= 0 && q % SF == 0) acc = fmaf(u[t] * SF, src[(q / SF - cell0) * src_stride], acc);
        }
    }
    return acc;
}

template <int SF>
__global__ void __launch_bounds__(256) cem_up_kernel(const __grid_constant__ esr_cem_filters f,
                                                     const float* __restrict__ x, const float* __restrict__ y,
                                                     float* __restrict__ out, int h, int w, int crop, float sign) {
    extern __shared__ float sm[];
    const int nt = f.n_ds, pad = nt / 2, pre = f.pre;
    const int H = h * SF, W = w * SF;
    const int plane = blockIdx.z;
    const int I0 = blockIdx.y * UT_R, J0 = blockIdx.x * UT_C;
    const int ext = (nt - 1) / SF + 2;                  // LR halo cells needed on each side (conservative)
    const int rows_in = UT_R + 2 * ext, cols_in = UT_C + 2 * ext;
    float* tile = sm;                                   // x[I0-ext .., J0-ext ..] (zero outside)
    float* hbuf = sm + rows_in * cols_in;               // [rows_in][UT_C*SF] horizontally upsampled
    const float* xp = x + static_cast<size_t>(plane) * h * w;
    for (int idx = threadIdx.x; idx < rows_in * cols_in; idx += blockDim.x) {
        const int r = idx / cols_in, c = idx - r * cols_in;
        const int rr = I0 - ext + r, cc = J0 - ext + c;
        tile[idx] = (rr >= 0 && rr < h && cc >= 0 && cc < w) ? __ldg(xp + static_cast<size_t>(rr) * w + cc) : 0.f;
    }
    __syncthreads();
    constexpr int wc = UT_C * SF, hr = UT_R * SF;
    for (int idx = threadIdx.x; idx < rows_in * wc; idx += blockDim.x) {
        const int r = idx / wc, cx = idx - r * wc;
        const int X = J0 * SF + cx;
        hbuf[idx] = X < W ? up_taps_1d<SF>(f.ds, nt, pad, pre, X, W, tile + r * cols_in, 1, J0 - ext) : 0.f;
    }
    __syncthreads();
    const int Hout = H - 2 * crop, Wout = W - 2 * crop;
    for (int idx = threadIdx.x; idx < hr * wc; idx += blockDim.x) {
        const int ry = idx / wc, cx = idx - ry * wc;
        const int Y = I0 * SF + ry, X = J0 * SF + cx;
        if (Y < crop || Y >= H - crop || X < crop || X >= W - crop) continue;
        const float acc = up_taps_1d<SF>(f.ds, nt, pad, pre, Y, H, hbuf + cx, wc, I0 - ext);
        const size_t o = (static_cast<size_t>(plane) * Hout + (Y - crop)) * Wout + (X - crop);
        const float base = y != nullptr ? __ldg(y + (static_cast<size_t>(plane) * H + Y) * W + X) : 0.f;
        out[o] = base + sign * acc;
    }
}

// --------------------------------------------------------------- x4 streaming fast paths
// Register / warp-shuffle kernels for sf = 4 (the production scale).  A warp owns a strip of 28 LR
// columns (lane l holds LR cell j0-2+l, i.e. one float4 of 4 HR pixels per HR row) and a segment of
// kSegRows LR rows; the two halo cells on either side come from neighbouring lanes via shuffles, so
// every HR float4 is loaded / stored exactly once per strip, fully coalesced (448 B per warp row).
constexpr int kStripCells = 28;
constexpr int kSegRows = 8;          // LR rows per warp (host-chosen launch parameter of the kernels below)

// Vertical taps first: every HR row costs 20 FMAs into five float4 accumulators (LR rows I-2..I+2) and no
// shuffles; the horizontal taps run once per LR row on the finished accumulator (20 FMAs for the cell's five
// partial sums, which neighbours exchange with 4 shuffles).  Loads go through a per-warp shared-memory ring of
// kDownGroups 4-row groups filled with cp.async (16 bytes per lane per row, each lane only ever reads its own
// slots, so no barrier is needed): kDownGroups-1 groups = 12 HR rows (6 KiB per warp) are in flight without
// costing registers, which is what this latency-bound kernel needs (28 warps per SM instead of 16).
constexpr int kDownGroups = 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(128) cem_down4_kernel(const __grid_constant__ CemTab T, const float* __restrict__ y,
                                                        const float* __restrict__ x, float* __restrict__ out, int H,
                                                        int W, int seg) {
    __shared__ float4 ring[4][kDownGroups][4][32];                  // [warp][group slot][HR row of the group][lane]
    __shared__ float xring[4][kDownGroups][32];                     // x of the LR row that group completes (row I-2)
    const int h = H >> 2, w = W >> 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int plane = blockIdx.z;
    const int j = blockIdx.x * kStripCells - 2 + lane;             // LR cell of this lane
    const int i0 = (blockIdx.y * 4 + warp) * seg;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a PDL-launched K+Up may start its prologue early
    if (i0 >= h) return;
    const int i1 = min(i0 + seg, h);
    const float* yp = y + static_cast<size_t>(plane) * H * W;
    // cells outside the image replicate the edge pixel: fetch the edge cell, pick its outer element when read back
    const int edge = j < 0 ? -1 : (j >= w ? 1 : 0);
    const bool strip_has_edge = blockIdx.x == 0 || (blockIdx.x + 1) * kStripCells + 2 > w;   // warp-uniform
    const float* col = yp + 4 * min(max(j, 0), w - 1);
    const bool writer = lane >= 2 && lane < 2 + kStripCells && j < w;
    const int last = i1 + 1;                                        // LR row groups i0-2 .. i1+1 feed rows i0 .. i1-1
    const uint32_t ring_base = static_cast<uint32_t>(__cvta_generic_to_shared(&ring[warp][0][0][lane]));
    constexpr uint32_t kRowBytes = 32 * sizeof(float4), kGroupBytes = 4 * kRowBytes;
    const float* xcol = x != nullptr ? x + static_cast<size_t>(plane) * h * w + j : nullptr;
    const uint32_t xring_base = static_cast<uint32_t>(__cvta_generic_to_shared(&xring[warp][0][lane]));
    auto issue_group = [&](int I) {                                 // the 4 HR rows of LR row I (rows replicate padded)
        if (I <= last) {
            const uint32_t gslot = static_cast<uint32_t>((I - (i0 - 2)) % kDownGroups);
            const uint32_t dst = ring_base + gslot * kGroupBytes;
            if (xcol != nullptr && writer && I - 2 >= i0 && I - 2 < i1)   // same depth of prefetch as the HR rows
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(xring_base + gslot * 128u), "l"(xcol + static_cast<size_t>(I - 2) * w) : "memory");
            if (I >= 0 && I < h) {                                  // interior group: one row pointer, no clamps
                const float* src = col + static_cast<size_t>(4 * I) * W;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + q * kRowBytes), "l"(src + static_cast<size_t>(q) * W) : "memory");
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = min(max(4 * I + q, 0), H - 1);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + q * kRowBytes), "l"(col + static_cast<size_t>(r) * W) : "memory");
                }
            }
        }
        cp_async_commit();                                          // (possibly empty) group: keeps the wait counts uniform
    };
    float2 acc_lo[5], acc_hi[5];                                    // LR rows I-2 .. I+2 (4 HR columns of the cell each)
#pragma unroll
    for (int m = 0; m < 5; ++m) acc_lo[m] = acc_hi[m] = make_float2(0.f, 0.f);
#pragma unroll
    for (int g = 0; g < kDownGroups - 1; ++g) issue_group(i0 - 2 + g);
    for (int I = i0 - 2; I <= last; ++I) {
        cp_async_wait<kDownGroups - 2>();                           // group I has landed (this thread's own copies)
        const float4(*slot)[32] = ring[warp][(I - (i0 - 2)) % kDownGroups];
        const float x_cur = xring[warp][(I - (i0 - 2)) % kDownGroups][lane];   // only meaningful where it was fetched
        float4 c[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) c[q] = slot[q][lane];
        if (strip_has_edge) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (edge < 0) c[q] = make_float4(c[q].x, c[q].x, c[q].x, c[q].x);
                else if (edge > 0) c[q] = make_float4(c[q].w, c[q].w, c[q].w, c[q].w);
            }
        }
        issue_group(I + kDownGroups - 1);                           // refills the slot read one iteration ago
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int m = 0; m < 5; ++m) {                           // row 4I+q feeds LR row I-2+m with tap index 4-m
                const float2 wv = T.down_v2[q][4 - m];              // packed fp32: two FFMA2 instead of four FFMA
                acc_lo[m] = __ffma2_rn(wv, make_float2(c[q].x, c[q].y), acc_lo[m]);
                acc_hi[m] = __ffma2_rn(wv, make_float2(c[q].z, c[q].w), acc_hi[m]);
            }
        }
        // LR row I-2 is complete: this cell's contribution to the output columns (own cell) - k, k = -2..2
        const float4 a = make_float4(acc_lo[0].x, acc_lo[0].y, acc_hi[0].x, acc_hi[0].y);
        float hsum = 0.f;
#pragma unroll
        for (int k = -2; k <= 2; ++k) {
            float p = T.down_h[0][k + 2] * a.x;
            p = fmaf(T.down_h[1][k + 2], a.y, p);
            p = fmaf(T.down_h[2][k + 2], a.z, p);
            p = fmaf(T.down_h[3][k + 2], a.w, p);
            hsum += k == 0 ? p : __shfl_sync(0xffffffffu, p, lane + k);
        }
        const int i = I - 2;
        if (i >= i0 && i < i1 && writer) {
            const size_t o = (static_cast<size_t>(plane) * h + i) * w + j;
            out[o] = x != nullptr ? x_cur - hsum : hsum;
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) { acc_lo[m] = acc_lo[m + 1]; acc_hi[m] = acc_hi[m + 1]; }
        acc_lo[4] = acc_hi[4] = make_float2(0.f, 0.f);
    }
    cp_async_wait<0>();
}

// out = crop(y + Up(K * d)) with the 27x27 (HH^T)^-1 correlation fused in: the block builds the tile of
// e = K * d it needs (its LR rows +-2, 32 cells) in shared memory with two separable passes over a replicate-
// clamped tile of d, then every warp streams its LR rows: horizontal polyphase taps straight from the shared
// e tile (no shuffles), a rolling 5-row window for the vertical taps, y and out as coalesced float4 rows.
struct InvTaps {
    float t[ESR_CEM_MAX_TAPS];
    float2 t2[ESR_CEM_MAX_TAPS];   // (t, t): FFMA2 operand form
    int n;
};

// NT = compile-time length of the (HH^T)^-1 factor (27 for the bicubic x4 configuration; 0 = run-time loops).  With
// NT known the two separable passes are register tiled (4 outputs per thread share NT+3 loaded values) and fully
// unrolled: 30 shared loads + 108 FMAs per 4 outputs instead of 4 x 27 x (constant load, shared load, FMA, loop).
template <int NT, int MINB>
__global__ void __launch_bounds__(128, MINB) cem_invup4_kernel(const __grid_constant__ CemTab T, const __grid_constant__ InvTaps K,
                                                         const float* __restrict__ d, const float* __restrict__ y,
                                                         float* __restrict__ out, int h, int w, int crop, int seg) {
    extern __shared__ float sm[];
    const int H = h << 2, W = w << 2;
    const int Ho = H - 2 * crop, Wo = W - 2 * crop;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int plane = blockIdx.z;
    const int jb = blockIdx.x * kStripCells - 2;                   // LR cell of lane 0
    const int j = jb + lane;
    const int ib = blockIdx.y * 4 * seg;                           // first LR row of the block
    const int RB = 4 * seg, nt = NT > 0 ? NT : K.n, pad = nt >> 1;
    const int Re = RB + 4, Rd = Re + 2 * pad, Cd = 32 + 2 * pad, Cds = (Cd + 3) & ~3;
    float* dt = sm;                                                // [Rd][Cds]  d, replicate clamped
    float* hb = dt + Rd * Cds;                                     // [Rd][32]   horizontal pass
    float* et = hb + Rd * 32;                                      // [Re][32]   e (zero outside the image)
    const int i0 = ib + warp * seg, i1 = min(i0 + seg, h);
    const bool writer = lane >= 2 && lane < 2 + kStripCells && j < w && 4 * j >= crop && 4 * j + 3 < W - crop;
    // running pointers: LR row i's first HR row in y / out (one add per row instead of a 64-bit multiply per access)
    const int Wq = W >> 2;
    const float4* ybase = reinterpret_cast<const float4*>(y + static_cast<size_t>(plane) * H * W) + j;
    // y is fetched three LR rows (12 HR rows, 6 KiB per warp) ahead into four register buffers that take turns
    // (the streaming loop is unrolled by four): no register copies, so no instruction waits on the newest load.
    auto rows_kept = [&](int i) { return writer && i < i1 && 4 * i >= crop && 4 * i + 3 < H - crop; };   // crop % 4 == 0
    auto load_y = [&](int i, float4 (&b)[4]) {                     // the 4 HR rows of LR row i (base image)
        if (rows_kept(i)) {
            const float4* p = ybase + static_cast<size_t>(4 * i) * Wq;
#pragma unroll
            for (int psi = 0; psi < 4; ++psi) b[psi] = __ldg(p + psi * Wq);
        }
    };
    float4 yb0[4], yb1[4], yb2[4], yb3[4];
#pragma unroll
    for (int psi = 0; psi < 4; ++psi) yb0[psi] = yb1[psi] = yb2[psi] = yb3[psi] = make_float4(0.f, 0.f, 0.f, 0.f);
    load_y(i0, yb0);                                               // HBM loads fly while the e tile is built
    load_y(i0 + 1, yb1);
    load_y(i0 + 2, yb2);
    asm volatile("griddepcontrol.wait;" ::: "memory");             // d comes from the Down launch just before (PDL)
    const float* dp = d + static_cast<size_t>(plane) * h * w;
    for (int r = warp; r < Rd; r += 4) {                           // a warp per tile row: no div / mod, one clamp per row
        const float* row = dp + static_cast<size_t>(clampi(ib - 2 - pad + r, 0, h - 1)) * w;
        float* drow = NT > 0 ? dt + (r >> 1) * Cds * 2 + (r & 1) : dt + r * Cds;   // NT > 0: row pairs interleaved
        const int cstep = NT > 0 ? 2 : 1;
        for (int c = lane; c < Cd; c += 32)                        // cp.async: the whole tile is in flight at once
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(drow + c * cstep))),
                         "l"(row + clampi(jb - pad + c, 0, w - 1))
                         : "memory");
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if constexpr (NT > 0) {
        // horizontal pass: 2 rows x 2 columns per task.  The d tile is stored with row pairs interleaved
        // (dt[(r/2)*Cds + col][r&1]), so one 16-byte load brings two columns of both rows as two aligned register
        // pairs and every tap is one packed FMA per output column: 14 LDS.128 + 54 FFMA2 per 4 outputs.
        for (int task = threadIdx.x; task < (Rd >> 1) * 16; task += 128) {
            const int rp = task >> 4, c = (task & 15) * 2;
            const float4* src = reinterpret_cast<const float4*>(dt + (rp * Cds + c) * 2);
            float2 v[NT + 1];                                               // v[k] = (d[2rp][c+k], d[2rp+1][c+k])
#pragma unroll
            for (int q = 0; q < (NT + 1) / 2; ++q) {
                const float4 t4 = src[q];
                v[2 * q] = make_float2(t4.x, t4.y);
                v[2 * q + 1] = make_float2(t4.z, t4.w);
            }
            float2 o0 = make_float2(0.f, 0.f), o1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                const float2 tk = K.t2[k];
                o0 = __ffma2_rn(tk, v[k], o0);
                o1 = __ffma2_rn(tk, v[k + 1], o1);
            }
            *reinterpret_cast<float2*>(hb + (2 * rp) * 32 + c) = make_float2(o0.x, o1.x);
            *reinterpret_cast<float2*>(hb + (2 * rp + 1) * 32 + c) = make_float2(o0.y, o1.y);
        }
        __syncthreads();
        for (int task = threadIdx.x; task < (Re >> 2) * 16; task += 128) {  // 4 rows (Re % 4 == 0) x 2 columns per task
            const int r0 = (task >> 4) * 4, c = (task & 15) * 2;
            float2 v[NT + 3];                                               // column pairs: packed fp32 FMAs
#pragma unroll
            for (int k = 0; k < NT + 3; ++k) v[k] = *reinterpret_cast<const float2*>(hb + (r0 + k) * 32 + c);
            float2 o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                const float2 tk = K.t2[k];
#pragma unroll
                for (int q = 0; q < 4; ++q) o[q] = __ffma2_rn(tk, v[k + q], o[q]);
            }
            const int jj = jb + c;
            const bool in0 = jj >= 0 && jj < w, in1 = jj + 1 >= 0 && jj + 1 < w;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ii = ib - 2 + r0 + q;
                const bool rin = ii >= 0 && ii < h;
                *reinterpret_cast<float2*>(et + (r0 + q) * 32 + c) = make_float2(rin && in0 ? o[q].x : 0.f, rin && in1 ? o[q].y : 0.f);
            }
        }
    } else {
        for (int idx = threadIdx.x; idx < Rd * 32; idx += 128) {
            const int r = idx >> 5, c = idx & 31;
            const float* t = dt + r * Cds + c;
            float acc = 0.f;
            for (int k = 0; k < nt; ++k) acc = fmaf(K.t[k], t[k], acc);
            hb[idx] = acc;
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < Re * 32; idx += 128) {
            const int r = idx >> 5, c = idx & 31;
            const int ii = ib - 2 + r, jj = jb + c;
            float acc = 0.f;
            if (ii >= 0 && ii < h && jj >= 0 && jj < w) {
                const float* t = hb + r * 32 + c;
                for (int k = 0; k < nt; ++k) acc = fmaf(K.t[k], t[k * 32], acc);
            }
            et[idx] = acc;
        }
    }
    __syncthreads();
    if (i0 >= h) return;
    // horizontally upsampled e rows i-2 .. i+2 (4 HR phases each); et row index of LR row ii is ii - ib + 2
    float2 hu01[5], hu23[5];                                       // HR phases (0,1) and (2,3) of each row
    int lofs[5];                                                   // neighbour cells (edge lanes never write: clamped)
#pragma unroll
    for (int k = 0; k < 5; ++k) lofs[k] = min(max(lane + k - 2, 0), 31);
    auto hrow = [&](int ii, float2& o01, float2& o23) {
        const float* e = et + (ii - ib + 2) * 32;
        float2 p01 = make_float2(0.f, 0.f), p23 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float n = e[lofs[k]];
            const float2 n2 = make_float2(n, n);
            p01 = __ffma2_rn(T.up_h01[k], n2, p01);
            p23 = __ffma2_rn(T.up_h23[k], n2, p23);
        }
        o01 = p01; o23 = p23;
    };
#pragma unroll
    for (int kv = 0; kv < 4; ++kv) hrow(i0 - 2 + kv, hu01[kv], hu23[kv]);
    auto step = [&](int i, const float4 (&cur)[4], float4 (&fill)[4]) {   // emits LR row i from `cur`, refills `fill` with row i+3
        hrow(i + 2, hu01[4], hu23[4]);
        load_y(i + 3, fill);
        if (rows_kept(i)) {
            float4* op = reinterpret_cast<float4*>(out + (static_cast<size_t>(plane) * Ho + (4 * i - crop)) * Wo + (4 * j - crop));
            const int Woq = Wo >> 2;
#pragma unroll
            for (int psi = 0; psi < 4; ++psi) {
                float2 r01 = make_float2(cur[psi].x, cur[psi].y), r23 = make_float2(cur[psi].z, cur[psi].w);
#pragma unroll
                for (int kv = 0; kv < 5; ++kv) {
                    const float2 wv = T.up_v2[psi][kv];
                    r01 = __ffma2_rn(wv, hu01[kv], r01);
                    r23 = __ffma2_rn(wv, hu23[kv], r23);
                }
                op[psi * Woq] = make_float4(r01.x, r01.y, r23.x, r23.y);
            }
        }
#pragma unroll
        for (int kv = 0; kv < 4; ++kv) { hu01[kv] = hu01[kv + 1]; hu23[kv] = hu23[kv + 1]; }
    };
    for (int i = i0; i < i1; i += 4) {
        step(i, yb0, yb3);
        if (i + 1 >= i1) break;
        step(i + 1, yb1, yb0);
        if (i + 2 >= i1) break;
        step(i + 2, yb2, yb1);
        if (i + 3 >= i1) break;
        step(i + 3, yb3, yb2);
    }
}

__global__ void __launch_bounds__(128) cem_up4_kernel(const __grid_constant__ CemTab T, const float* __restrict__ e,
                                                      const float* __restrict__ y, float* __restrict__ out, int h,
                                                      int w, int crop, float sign) {
    const int H = h << 2, W = w << 2;
    const int Ho = H - 2 * crop, Wo = W - 2 * crop;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int plane = blockIdx.z;
    const int j = blockIdx.x * kStripCells - 2 + lane;
    const int i0 = (blockIdx.y * 4 + warp) * kSegRows;
    if (i0 >= h) return;
    const int i1 = min(i0 + kSegRows, h);
    const float* ep = e + static_cast<size_t>(plane) * h * w;
    const bool inside = j >= 0 && j < w;
    const bool writer = lane >= 2 && lane < 2 + kStripCells && j < w && 4 * j >= crop && 4 * j + 3 < W - crop;
    auto load_y = [&](int i, float4 (&b)[4]) {                     // the 4 HR rows of LR row i (base image)
#pragma unroll
        for (int psi = 0; psi < 4; ++psi) {
            const int Y = 4 * i + psi;
            b[psi] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (writer && y != nullptr && Y >= crop && Y < H - crop)
                b[psi] = __ldg(reinterpret_cast<const float4*>(y + (static_cast<size_t>(plane) * H + Y) * W) + j);
        }
    };
    float4 yc[4], yn[4];
    load_y(i0, yc);
    for (int i = i0; i < i1; ++i) {
        if (i + 1 < i1) load_y(i + 1, yn);                         // next row's HBM loads fly during this row's math
        float hu[5][4];                                            // horizontally upsampled LR rows i-2..i+2, 4 HR phases
#pragma unroll
        for (int kv = 0; kv < 5; ++kv) {
            const int ii = i + kv - 2;
            const float v = (inside && ii >= 0 && ii < h) ? __ldg(ep + static_cast<size_t>(ii) * w + j) : 0.f;
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
            for (int k = -2; k <= 2; ++k) {
                const float n = k == 0 ? v : __shfl_sync(0xffffffffu, v, lane + k);
                p0 = fmaf(T.up[0][k + 2], n, p0);
                p1 = fmaf(T.up[1][k + 2], n, p1);
                p2 = fmaf(T.up[2][k + 2], n, p2);
                p3 = fmaf(T.up[3][k + 2], n, p3);
            }
            hu[kv][0] = p0; hu[kv][1] = p1; hu[kv][2] = p2; hu[kv][3] = p3;
        }
        if (writer) {
#pragma unroll
            for (int psi = 0; psi < 4; ++psi) {
                const int Y = 4 * i + psi;
                if (Y < crop || Y >= H - crop) continue;
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int kv = 0; kv < 5; ++kv) {
                    const float wv = T.up[psi][kv];
                    r.x = fmaf(wv, hu[kv][0], r.x); r.y = fmaf(wv, hu[kv][1], r.y);
                    r.z = fmaf(wv, hu[kv][2], r.z); r.w = fmaf(wv, hu[kv][3], r.w);
                }
                float4 b = yc[psi];
                b.x = fmaf(sign, r.x, b.x); b.y = fmaf(sign, r.y, b.y); b.z = fmaf(sign, r.z, b.z); b.w = fmaf(sign, r.w, b.w);
                *reinterpret_cast<float4*>(out + (static_cast<size_t>(plane) * Ho + (Y - crop)) * Wo + (4 * j - crop)) = b;
            }
        }
#pragma unroll
        for (int psi = 0; psi < 4; ++psi) yc[psi] = yn[psi];
    }
}

// LR rows per warp (`seg`) of the streaming kernels.  A block covers 4*seg LR rows of a strip and pays a fixed
// halo on top (Down: 4 LR-row groups of loads + FMAs; K+Up: the 26 extra rows of its d / e tiles, worth ~1.2 rows of
// streaming), so long segments amortise the halo, but the last block row of a plane is padded up to 4*seg rows and
// small problems need >= 4 blocks (16 warps) per SM to hide latency.  Cost model fitted to a sweep on B200
// (tools/cem_seg_sweep.py; config 4: 3 planes of 512^2 cells, config 2: 48 planes of 148^2):
//   cost(seg) = ceil(h / 4seg) * (seg + halo) / min(1, blocks / (148 * 4)) * ceil(waves) / waves.
static int pick_seg(int planes, int h, int w, const char* env_name, int max_seg, float halo, int blocks_per_sm) {
    if (const char* v = getenv(env_name)) {                       // tuning aid
        const int forced = atoi(v);
        if (forced >= 2 && forced <= max_seg) return forced;
    }
    const long strips = ceil_div(w, kStripCells);
    int best = 8;
    float best_cost = 3.4e38f;
    for (int seg = 6; seg <= max_seg; ++seg) {
        const int block_rows = ceil_div(h, 4 * seg);
        const float blocks = static_cast<float>(strips * block_rows * planes);
        const float fill = blocks / (148.f * 4.f);
        float cost = block_rows * (seg + halo) / (fill < 1.f ? fill : 1.f);
        const float waves = blocks / (148.f * blocks_per_sm);     // a partly filled last wave costs a whole one (ncu: 38 %
        if (waves > 1.f) cost *= ceilf(waves) / waves;            // of K+Up's time had most SMs idle at 1.06 waves)
        if (cost < best_cost) { best_cost = cost; best = seg; }
    }
    return best;
}
static int pick_seg_down(int planes, int h, int w) { return pick_seg(planes, h, w, "ESR_CEM_SEG_DOWN", 16, 4.f, 6); }
static int pick_seg_up(int planes, int h, int w) { return pick_seg(planes, h, w, "ESR_CEM_SEG_UP", 16, 1.2f, 4); }

static bool fast4_ok(const esr_cem_filters& f, int H, int W, int crop, const void* a, const void* b, const void* c) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return f.sf == 4 && f.n_ds == 17 && H % 4 == 0 && W % 4 == 0 && crop % 4 == 0 && al(a) && al(b) && al(c);
}

CemTab make_tab(const esr_cem_filters& f) {
    CemTab T;
    const int nt = f.n_ds, pad = nt / 2, pre = f.pre;
    for (int p = 0; p < 4; ++p)
        for (int k = -2; k <= 2; ++k) {
            const int td = 4 * k + p + pad - pre;        // down: HR sample 4(j+k)+p = 4j + pre + t - pad
            const float wd = (td >= 0 && td < nt) ? f.ds[nt - 1 - td] : 0.f;
            T.down_h[p][k + 2] = wd;
            T.down_v[p][k + 2] = wd;                     // same relation with m = k (row 4I+q -> LR row I-m uses t = 4m+q+pad-pre)
            const int tu = 4 * k + pad + pre - p;        // up: 4j+p + t - pad - pre = 4(j+k)
            T.up[p][k + 2] = (tu >= 0 && tu < nt) ? f.ds[tu] * f.sf : 0.f;
        }
    for (int p = 0; p < 4; ++p)
        for (int c = 0; c < 5; ++c) {
            T.down_v2[p][c] = make_float2(T.down_v[p][c], T.down_v[p][c]);
            T.up_v2[p][c] = make_float2(T.up[p][c], T.up[p][c]);
        }
    for (int c = 0; c < 5; ++c) {
        T.up_h01[c] = make_float2(T.up[0][c], T.up[1][c]);
        T.up_h23[c] = make_float2(T.up[2][c], T.up[3][c]);
    }
    return T;
}

// ===================================================================== x4 projection, round 2: TMA-fed streaming kernels
// Same two-launch structure (Down, then K + Up + add as a programmatic dependent), same arithmetic as the kernels above,
// but (a) every HR row of y reaches a warp through its own ring of TMA tile loads (cp.async.bulk.tensor, one
// instruction per 4-row group issued by one lane, completion on an mbarrier) instead of 128 per-lane cp.async / LDG
// with address arithmetic, eight groups (32 HR rows, 16 KiB per warp) in flight without costing registers;
// (b) persistent CTAs, one per SM, that walk a list of equal work items sized by a cost model so that the last wave is
// full (ncu on the round-1 kernels: SMs active 67 % / 79 % of the time - ramp, 3.08 blocks per SM, tail);
// (c) structurally zero filter taps are skipped (16 of the 20 polyphase slots of the bicubic x4 kernel are non-zero)
// and the rolling accumulator windows rotate by renaming (the row loop is unrolled by five) instead of register copies.
// Border handling: TMA fills columns outside the image with zeros; the lanes holding such cells take the replicated
// edge value from the lane that owns the edge cell (one shuffle per row, only in strips that touch a border); row
// groups above / below the image are fetched as four single-row boxes at clamped coordinates.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();
int num_sms_cached();

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

template <int V> using IC = std::integral_constant<int, V>;

// A ring wait that never completes is a protocol bug.  Instead of trapping (which kills the context and says nothing)
// the streaming kernels record where it happened and carry on with whatever is in the slot: the launch ends, results
// are wrong, and esr_debug_cem_timeout() reports {code, block, warp, group}.
__device__ unsigned int g_cem_timeout[4];
__device__ __forceinline__ void ring_wait(uint64_t* bar, uint32_t parity, unsigned code, unsigned n) {
    for (uint32_t it = 0; !mbar_try_wait(bar, parity); ++it) {
        if (it > (1u << 20)) {
            if (atomicCAS(&g_cem_timeout[0], 0u, code) == 0u) {
                g_cem_timeout[1] = blockIdx.x; g_cem_timeout[2] = threadIdx.x; g_cem_timeout[3] = n;
            }
            return;
        }
    }
}

constexpr int kAW = 16;                         // Down: warps per CTA (each an independent streaming unit; with 8 the kernel was
                                                // latency bound: issue slots 41 % busy, mostly "wait" stalls, ncu r02)
constexpr int kAD = 4;                          // ring stages per warp, one 4-row group each (16 warps x 4 x 2 KiB in flight)
constexpr int kAStage = 4 * 512 + 256;          // 4 HR rows x 32 cells x 16 B, then 36 x values of the LR row it completes (cells j0-4 ..:
                                                // a TMA box must start on a 16-byte boundary, j0-2 does not)
constexpr uint32_t kAXBytes = 36 * 4;
constexpr uint32_t kAllTaps = 0xfffffu;

struct Down4Args {
    const float* x;
    float* out;
    int H, W, h, w;
    int strips, nseg, seg, items;
    int dbg;                                    // ESR_CEM_DBGA bisect bits: 1 no x TMA, 2 single-row boxes only, 4 no negative column start
};

// VMASK / HMASK: bit (p*5 + c) set <=> T.down_v[p][c] / T.down_h[p][c] may be non-zero (host checked)
template <uint32_t VMASK, uint32_t HMASK>
__global__ void __launch_bounds__(kAW * 32, 1)
cem_down4s_kernel(const __grid_constant__ CUtensorMap tmY4, const __grid_constant__ CUtensorMap tmY1,
                  const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CemTab T, const __grid_constant__ Down4Args A) {
    extern __shared__ uint8_t sm_raw[];
    uint8_t* sm = sm_raw + ((128u - (smem_u32(sm_raw) & 127u)) & 127u);     // 128-byte aligned, still a shared-space pointer
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform: the ring, item and
                                                                                       // TMA operands live in uniform registers
    uint8_t* ring = sm + warp * (kAD * kAStage);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + kAW * kAD * kAStage) + warp * kAD;
    if (lane == 0) {
        for (int s = 0; s < kAD; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    pdl_launch_dependents();                                       // the K + Up launch may take the SMs this grid leaves
    const int h = A.h, w = A.w, H = A.H;
    uint32_t cnt = 0;                                              // groups consumed so far: stage = cnt % kAD, parity = (cnt / kAD) & 1
    for (int item = blockIdx.x * kAW + warp; item < A.items; item += gridDim.x * kAW) {
        const int strip = item % A.strips;
        const int t = item / A.strips;
        const int sg = t % A.nseg, plane = t / A.nseg;
        const int j0 = strip * kStripCells, j = j0 - 2 + lane;
        const int i0 = sg * A.seg, i1 = min(i0 + A.seg, h);
        if (i0 >= h) continue;
        const int ngroups = i1 - i0 + 4;                           // LR row groups i0-2 .. i1+1 feed rows i0 .. i1-1
        const bool writer = lane >= 2 && lane < 2 + kStripCells && j < w;
        const bool edge_l = j0 == 0, edge_r = j0 + kStripCells + 2 > w;            // warp-uniform
        const int lane_last = w - 1 - (j0 - 2);                                    // lane of the last cell (edge_r strips)
        auto issue = [&](int n, uint32_t slot) {                                   // lane 0 only
            const int I = i0 - 2 + n;
            uint8_t* dst = ring + slot * kAStage;
            const bool want_x = A.x != nullptr && I - 2 >= i0 && I - 2 < i1 && !(A.dbg & 1);
            mbar_expect_tx(&bars[slot], 2048u + (want_x ? kAXBytes : 0u));
            const int c0 = (A.dbg & 4) ? max(4 * (j0 - 2), 0) : 4 * (j0 - 2);
            if (I >= 0 && I < h && !(A.dbg & 2)) {
                tma_load_3d(dst, &tmY4, &bars[slot], c0, 4 * I, plane);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    tma_load_3d(dst + q * 512, &tmY1, &bars[slot], c0, min(max(4 * I + q, 0), H - 1), plane);
            }
            if (want_x) tma_load_3d(dst + 2048, &tmX, &bars[slot], j0 - 4, I - 2, plane);
        };
        if (elect_one()) {
            const int npre = ngroups < kAD ? ngroups : kAD;
            for (int n = 0; n < npre; ++n) issue(n, (cnt + n) % kAD);
        }
        float2 acc_lo[5], acc_hi[5];
#pragma unroll
        for (int m = 0; m < 5; ++m) acc_lo[m] = acc_hi[m] = make_float2(0.f, 0.f);
        // group n of the item (LR row group I = i0-2+n); R = n % 5 is the rotation of the accumulator window:
        // LR row I-2+m lives in slot (m + R) % 5
        auto body = [&](auto Rc, int n) {
            constexpr int R = decltype(Rc)::value;
            const uint32_t g = cnt + n, slot = g % kAD;
            ring_wait(&bars[slot], (g / kAD) & 1u, 1u, static_cast<unsigned>(n));
            const uint8_t* st = ring + slot * kAStage;
            float4 c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) c[q] = reinterpret_cast<const float4*>(st + q * 512)[lane];
            const float x_cur = reinterpret_cast<const float*>(st + 2048)[lane + 2];
            if (edge_l || edge_r) {                                // replicate the edge pixel into the cells outside the image
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float first = __shfl_sync(0xffffffffu, c[q].x, 2);
                    const float last = __shfl_sync(0xffffffffu, c[q].w, lane_last & 31);
                    if (j < 0) c[q] = make_float4(first, first, first, first);
                    else if (j >= w) c[q] = make_float4(last, last, last, last);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int m = 0; m < 5; ++m) {                      // row 4I+q feeds LR row I-2+m with table entry [q][4-m]
                    if (!((VMASK >> (q * 5 + 4 - m)) & 1u)) continue;
                    const float2 wv = T.down_v2[q][4 - m];
                    acc_lo[(m + R) % 5] = __ffma2_rn(wv, make_float2(c[q].x, c[q].y), acc_lo[(m + R) % 5]);
                    acc_hi[(m + R) % 5] = __ffma2_rn(wv, make_float2(c[q].z, c[q].w), acc_hi[(m + R) % 5]);
                }
            }
            // the slot's values have been consumed by the FMAs above (data dependence, not just issue order): refill it
            __syncwarp();
            if (n + kAD < ngroups && elect_one()) issue(n + kAD, slot);
            // LR row I-2 is complete (if it belongs to the item: the first four groups only warm the window up):
            // this cell's contribution to the output columns (own cell) - k, k = -2..2
            const int i = i0 - 4 + n;                              // = I - 2
            if (i < i0) { acc_lo[R % 5] = acc_hi[R % 5] = make_float2(0.f, 0.f); return; }
            const float a0 = acc_lo[R % 5].x, a1 = acc_lo[R % 5].y, a2 = acc_hi[R % 5].x, a3 = acc_hi[R % 5].y;
            float hsum = 0.f;
#pragma unroll
            for (int k = -2; k <= 2; ++k) {
                float p = 0.f;
                if ((HMASK >> (0 * 5 + k + 2)) & 1u) p = T.down_h[0][k + 2] * a0;
                if ((HMASK >> (1 * 5 + k + 2)) & 1u) p = fmaf(T.down_h[1][k + 2], a1, p);
                if ((HMASK >> (2 * 5 + k + 2)) & 1u) p = fmaf(T.down_h[2][k + 2], a2, p);
                if ((HMASK >> (3 * 5 + k + 2)) & 1u) p = fmaf(T.down_h[3][k + 2], a3, p);
                hsum += k == 0 ? p : __shfl_sync(0xffffffffu, p, (lane + k) & 31);
            }
            if (i < i1 && writer) {
                const size_t o = (static_cast<size_t>(plane) * h + i) * w + j;
                A.out[o] = A.x != nullptr ? x_cur - hsum : hsum;
            }
            acc_lo[R % 5] = acc_hi[R % 5] = make_float2(0.f, 0.f);   // becomes LR row I+3 of the next group
        };
        for (int n = 0; n < ngroups; n += 5) {
            body(IC<0>{}, n);
            if (n + 1 < ngroups) body(IC<1>{}, n + 1);
            if (n + 2 < ngroups) body(IC<2>{}, n + 2);
            if (n + 3 < ngroups) body(IC<3>{}, n + 3);
            if (n + 4 < ngroups) body(IC<4>{}, n + 4);
        }
        cnt += ngroups;
    }
}

// ---- K + Up + add.  CTA = 4 * NS warps = NS strips (28 cells each) x 4 row quarters; the CTA builds e = K * d for its
// rows +-2 and 32 * NS columns in shared memory (same two register-tiled separable passes as cem_invup4_kernel), then
// every warp streams its rows with y arriving through a per-warp TMA ring (28 cells = 448 B per row, only what it writes).
// NS = 1 with several CTAs per SM balances best (one CTA of 8 warps per SM: 120 items on 148 SMs at config 4, latency
// bound with 2 warps per scheduler, ncu r02).
constexpr int kBD = 4;                          // ring stages per warp
constexpr int kBStage = 4 * 448;
constexpr int kBMaxRows = 128;                  // LR rows per CTA item

struct InvUp4Args {
    const float* d;
    float* out;
    int h, w, crop;
    int pairs, nbrow, rb, seg, items;           // strip pairs per plane, block rows per plane, rows per block, rows per warp
};

template <int NT, int NS>
__global__ void __launch_bounds__(NS * 128, NS == 1 ? 4 : 1)
cem_invup4s_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CemTab T, const __grid_constant__ InvTaps K,
                   const __grid_constant__ InvUp4Args A) {
    extern __shared__ uint8_t sm_raw[];
    uint8_t* smb = sm_raw + ((128u - (smem_u32(sm_raw) & 127u)) & 127u);    // 128-byte aligned, still a shared-space pointer
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform
    constexpr int kBW = 4 * NS, kBCE = 32 * NS;
    const int sx = warp % NS, sy = warp / NS;
    constexpr int pad = NT >> 1;
    const int h = A.h, w = A.w, crop = A.crop;
    const int H = h << 2, W = w << 2, Ho = H - 2 * crop, Wo = W - 2 * crop, Woq = Wo >> 2;
    const int RB = A.rb, Re = RB + 4, Rd = Re + 2 * pad;
    constexpr int Cd = kBCE + 2 * pad, Cds = (Cd + 3) & ~3;
    uint8_t* ring = smb + warp * (kBD * kBStage);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smb + kBW * kBD * kBStage) + warp * kBD;
    float* dt = reinterpret_cast<float*>(smb + kBW * kBD * kBStage + 256);     // [Rd][Cds] d, replicate clamped, row pairs interleaved
    float* hb = dt + ((Rd + 1) & ~1) * Cds;                                    // [Rd][64] horizontal pass
    float* et = dt;                                                            // [Re][64] e (zero outside the image); aliases dt
    if (lane == 0) {
        for (int s = 0; s < kBD; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    uint32_t cnt = 0;
    bool waited = false;
    int lofs[5];                                                   // neighbour cells in the e tile (edge lanes never write: clamped)
#pragma unroll
    for (int k = 0; k < 5; ++k) lofs[k] = min(max(lane + k - 2, 0), 31) + sx * kStripCells;
    for (int item = blockIdx.x; item < A.items; item += gridDim.x) {
        const int pr = item % A.pairs;
        const int t = item / A.pairs;
        const int br = t % A.nbrow, plane = t / A.nbrow;
        const int jb = pr * (NS * kStripCells) - 2;                 // LR cell of column 0 of the e tile
        const int ib = br * RB;                                    // first LR row of the block
        const int jw = jb + sx * kStripCells, j = jw + lane;       // this warp's strip: lane l holds cell jw + l
        const int i0 = ib + sy * A.seg, i1 = min(min(i0 + A.seg, ib + RB), h);
        const bool writer = lane >= 2 && lane < 2 + kStripCells && j < w && 4 * j >= crop && 4 * j + 3 < W - crop;
        // rows this warp really emits (crop % 4 == 0): [ia, ie)
        const int ia = max(i0, crop >> 2), ie = min(i1, (H - crop) >> 2);
        const int ngroups = ie > ia ? ie - ia : 0;
        const bool strip_live = jw + 2 < w;                        // the strip holds at least one image column
        auto issue = [&](int n, uint32_t slot) {                   // lane 0: the 4 HR rows of LR row ia + n, cells jw+2 .. jw+29
            mbar_expect_tx(&bars[slot], static_cast<uint32_t>(kBStage));
            tma_load_3d(ring + slot * kBStage, &tmY, &bars[slot], 4 * (jw + 2), 4 * (ia + n), plane);
        };
        const int live_groups = strip_live ? ngroups : 0;
        if (elect_one()) {
            const int npre = live_groups < kBD ? live_groups : kBD;
            for (int n = 0; n < npre; ++n) issue(n, (cnt + n) % kBD);
        }
        if (!waited) { pdl_wait(); waited = true; }                // d comes from the Down launch just before (PDL)
        // ---- d tile: rows ib-2-pad .. , columns jb-pad .. (replicate clamped), whole tile in flight at once
        const float* dp = A.d + static_cast<size_t>(plane) * h * w;
        for (int r = warp; r < Rd; r += kBW) {
            const float* row = dp + static_cast<size_t>(clampi(ib - 2 - pad + r, 0, h - 1)) * w;
            float* drow = dt + (r >> 1) * Cds * 2 + (r & 1);
            for (int c = lane; c < Cd; c += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(drow + c * 2)), "l"(row + clampi(jb - pad + c, 0, w - 1)) : "memory");
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        // ---- horizontal pass: 2 rows x 2 columns per task (one 16-byte load = two columns of both rows)
        for (int task = threadIdx.x; task < ((Rd + 1) >> 1) * (kBCE / 2); task += kBW * 32) {
            const int rp = task / (kBCE / 2), c = (task % (kBCE / 2)) * 2;
            const float4* src = reinterpret_cast<const float4*>(dt + (rp * Cds + c) * 2);
            float2 v[NT + 1];
#pragma unroll
            for (int q = 0; q < (NT + 1) / 2; ++q) {
                const float4 t4 = src[q];
                v[2 * q] = make_float2(t4.x, t4.y);
                v[2 * q + 1] = make_float2(t4.z, t4.w);
            }
            float2 o0 = make_float2(0.f, 0.f), o1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                const float2 tk = K.t2[k];
                o0 = __ffma2_rn(tk, v[k], o0);
                o1 = __ffma2_rn(tk, v[k + 1], o1);
            }
            *reinterpret_cast<float2*>(hb + (2 * rp) * kBCE + c) = make_float2(o0.x, o1.x);
            *reinterpret_cast<float2*>(hb + (2 * rp + 1) * kBCE + c) = make_float2(o0.y, o1.y);
        }
        __syncthreads();                                           // dt is dead from here on: et may overwrite it
        // ---- vertical pass: 4 rows x 2 columns per task
        for (int task = threadIdx.x; task < (Re >> 2) * (kBCE / 2); task += kBW * 32) {
            const int r0 = (task / (kBCE / 2)) * 4, c = (task % (kBCE / 2)) * 2;
            float2 v[NT + 3];
#pragma unroll
            for (int k = 0; k < NT + 3; ++k) v[k] = *reinterpret_cast<const float2*>(hb + (r0 + k) * kBCE + c);
            float2 o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                const float2 tk = K.t2[k];
#pragma unroll
                for (int q = 0; q < 4; ++q) o[q] = __ffma2_rn(tk, v[k + q], o[q]);
            }
            const int jj = jb + c;
            const bool in0 = jj >= 0 && jj < w, in1 = jj + 1 >= 0 && jj + 1 < w;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ii = ib - 2 + r0 + q;
                const bool rin = ii >= 0 && ii < h;
                *reinterpret_cast<float2*>(et + (r0 + q) * kBCE + c) = make_float2(rin && in0 ? o[q].x : 0.f, rin && in1 ? o[q].y : 0.f);
            }
        }
        __syncthreads();
        // ---- streaming: horizontally upsampled e rows i-2 .. i+2 in a rotating window, vertical taps, y +, crop
        if (live_groups > 0) {
            float2 hu01[5], hu23[5];
            auto hrow = [&](int ii, float2& o01, float2& o23) {   // et row index of LR row ii is ii - ib + 2
                const float* e = et + (ii - ib + 2) * kBCE;
                float2 p01 = make_float2(0.f, 0.f), p23 = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float nb = e[lofs[k]];
                    const float2 n2 = make_float2(nb, nb);
                    p01 = __ffma2_rn(T.up_h01[k], n2, p01);
                    p23 = __ffma2_rn(T.up_h23[k], n2, p23);
                }
                o01 = p01; o23 = p23;
            };
#pragma unroll
            for (int kv = 0; kv < 4; ++kv) hrow(ia - 2 + kv, hu01[kv], hu23[kv]);
            float4* obase = reinterpret_cast<float4*>(A.out + (static_cast<size_t>(plane) * Ho + (4 * ia - crop)) * Wo + (4 * j - crop));
            // row n of the warp (LR row ia + n); window slot of LR row i-2+kv is (kv + R) % 5, R = n % 5
            auto body = [&](auto Rc, int n) {
                constexpr int R = decltype(Rc)::value;
                const int i = ia + n;
                hrow(i + 2, hu01[(4 + R) % 5], hu23[(4 + R) % 5]);
                const uint32_t g = cnt + n, slot = g % kBD;
                ring_wait(&bars[slot], (g / kBD) & 1u, 2u, static_cast<unsigned>(n));
                float4 yv[4];
                const float4* sp = reinterpret_cast<const float4*>(ring + slot * kBStage) + min(max(lane - 2, 0), kStripCells - 1);
#pragma unroll
                for (int psi = 0; psi < 4; ++psi) yv[psi] = sp[psi * (kStripCells)];
                if (writer) {
                    float4* op = obase + static_cast<size_t>(4 * n) * Woq;
#pragma unroll
                    for (int psi = 0; psi < 4; ++psi) {
                        float2 r01 = make_float2(yv[psi].x, yv[psi].y), r23 = make_float2(yv[psi].z, yv[psi].w);
#pragma unroll
                        for (int kv = 0; kv < 5; ++kv) {
                            const float2 wv = T.up_v2[psi][kv];
                            r01 = __ffma2_rn(wv, hu01[(kv + R) % 5], r01);
                            r23 = __ffma2_rn(wv, hu23[(kv + R) % 5], r23);
                        }
                        op[psi * Woq] = make_float4(r01.x, r01.y, r23.x, r23.y);
                    }
                }
                // refill only after the values read from the slot have been USED (the FMAs above depend on them): a
                // __syncwarp() right after the loads only orders their issue, and with 12 warps per SM queueing on the
                // shared-memory pipe a refill from L2 was seen landing under loads still in flight (config-2 shape, 3 CTAs
                // per SM: sporadic wrong rows; never with one CTA per SM)
                __syncwarp();
                if (n + kBD < live_groups && elect_one()) issue(n + kBD, slot);
            };
            for (int n = 0; n < live_groups; n += 5) {
                body(IC<0>{}, n);
                if (n + 1 < live_groups) body(IC<1>{}, n + 1);
                if (n + 2 < live_groups) body(IC<2>{}, n + 2);
                if (n + 3 < live_groups) body(IC<3>{}, n + 3);
                if (n + 4 < live_groups) body(IC<4>{}, n + 4);
            }
            cnt += live_groups;
        }
        __syncthreads();                                           // the next item overwrites dt / hb / et
    }
    if (!waited) pdl_wait();
}

int make_plane_map(CUtensorMap* tm, const float* base, int planes, int rows, int cols, int box_cols, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return ESR_ERR_CUDA; }
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(planes)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(cols) * 4, static_cast<cuuint64_t>(rows) * cols * 4};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (CEM plane map) failed with CUresult %d", static_cast<int>(r)); return ESR_ERR_CUDA; }
    return ESR_OK;
}

// rows per work item minimising  ceil(items / units) * (rows + halo)  for `columns` independent columns of `h` rows on
// `units` parallel workers (items = columns * ceil(h / rows)); rows <= max_rows
static int pick_rows(long columns, int h, long units, float halo, int max_rows, int granule) {
    int best = granule;
    float best_cost = 3.4e38f;
    for (int rows = granule; rows <= max_rows; rows += granule) {
        const long items = columns * ceil_div(h, rows);
        const float cost = static_cast<float>((items + units - 1) / units) * (rows + halo);
        if (cost < best_cost - 1e-3f) { best_cost = cost; best = rows; }
        if (rows >= h) break;
    }
    return best;
}

static bool stream4_ok(const esr_cem_filters& f, int H, int W, int crop, const void* y, const void* x, const void* out, const void* d) {
    static const bool off = []() { const char* v = getenv("ESR_CEM_V1"); return v && atoi(v); }();     // A/B timing: round-1 kernels
    const int w = W / 4;
    return !off && f.n_inv == 27 && fast4_ok(f, H, W, crop, y, out, d) && (W - 2 * crop) % 4 == 0 && w % 4 == 0 && x != nullptr &&
           (reinterpret_cast<uintptr_t>(x) & 15) == 0 && H >= 8 && W >= 8;
}

// Down as the TMA streaming kernel; d = x - Down(y)
static int cem_down4s(const esr_cem_filters& f, const float* y, const float* x, int planes, int H, int W, float* out, cudaStream_t s) {
    const CemTab T = make_tab(f);
    uint32_t vmask = 0, hmask = 0;
    for (int p = 0; p < 4; ++p)
        for (int c = 0; c < 5; ++c) {
            if (T.down_v[p][c] != 0.f) vmask |= 1u << (p * 5 + c);
            if (T.down_h[p][c] != 0.f) hmask |= 1u << (p * 5 + c);
        }
    Down4Args A;
    A.x = x; A.out = out; A.H = H; A.W = W; A.h = H / 4; A.w = W / 4;
    A.dbg = []() { const char* v = getenv("ESR_CEM_DBGA"); return v ? atoi(v) : 0; }();
    A.strips = ceil_div(A.w, kStripCells);
    const int sms = num_sms_cached();
    A.seg = pick_rows(static_cast<long>(A.strips) * planes, A.h, static_cast<long>(sms) * kAW, 4.f, 1 << 20, 1);
    A.nseg = ceil_div(A.h, A.seg);
    A.items = planes * A.nseg * A.strips;
    CUtensorMap tmY4, tmY1, tmX;
    int rc;
    if ((rc = make_plane_map(&tmY4, y, planes, H, W, 128, 4))) return rc;
    if ((rc = make_plane_map(&tmY1, y, planes, H, W, 128, 1))) return rc;
    if ((rc = make_plane_map(&tmX, x, planes, A.h, A.w, 36, 1))) return rc;
    const int smem = 128 + kAW * kAD * kAStage + kAW * kAD * 8;
    const int grid = std::min(sms, ceil_div(A.items, kAW));   // one CTA of kAW warps per SM
    // the bicubic x4 table (sampling phase 1): taps [0][0], [0][1]; [2][4], [3][4] and one more are structurally zero
    constexpr uint32_t kBicV = 0xfffffu & ~((1u << (0 * 5 + 0)) | (1u << (1 * 5 + 0)) | (1u << (2 * 5 + 4)) | (1u << (3 * 5 + 4)));
    auto launch = [&](auto kern) -> int {
        ESR_ONCE_PER_DEVICE(ESR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)););
        kern<<<grid, kAW * 32, smem, s>>>(tmY4, tmY1, tmX, T, A);
        return check_launch("cem_down4s_kernel");
    };
    if ((vmask & ~kBicV) == 0 && (hmask & ~kBicV) == 0) return launch(cem_down4s_kernel<kBicV, kBicV>);
    return launch(cem_down4s_kernel<kAllTaps, kAllTaps>);
}

static int cem_invup4s(const esr_cem_filters& f, const float* d, const float* y, int planes, int h, int w, int crop, float* out,
                       cudaStream_t s) {
    InvTaps K;
    K.n = f.n_inv;
    for (int i = 0; i < f.n_inv; ++i) { K.t[i] = f.inv[i]; K.t2[i] = make_float2(f.inv[i], f.inv[i]); }
    InvUp4Args A;
    A.d = d; A.out = out; A.h = h; A.w = w; A.crop = crop;
    constexpr int NS = 1, kBW = 4 * NS, kBCE = 32 * NS;
    A.pairs = ceil_div(w, NS * kStripCells);
    const int sms = num_sms_cached();
    constexpr int NT = 27, pad = NT / 2, Cds = (kBCE + 2 * pad + 3) & ~3;
    auto smem_for = [&](int rb) {
        const int RdE = (rb + 4 + 2 * pad + 1) & ~1;
        return 128 + kBW * kBD * kBStage + 256 + 4 * (RdE * Cds + RdE * kBCE);
    };
    // rows per CTA item: a multiple of 16 (4 warps x 4-row tasks); the K pass of an item costs about 0.42 rows of
    // streaming per row plus a fixed ~7 rows (its 30 halo rows); CTAs per SM follow from the item's shared memory
    int best = 16;
    float best_cost = 3.4e38f;
    for (int rb = 16; rb <= kBMaxRows; rb += 16) {
        const int per_sm = std::max(1, std::min(NS == 1 ? 4 : 1, (225 * 1024) / (smem_for(rb) + 1024)));
        const long units = static_cast<long>(sms) * per_sm, items = static_cast<long>(A.pairs) * planes * ceil_div(h, rb);
        const float cost = static_cast<float>((items + units - 1) / units) * (rb + 6.f) * per_sm;   // time ~ rounds x item x CTAs sharing the SM
        if (cost < best_cost - 1e-3f) { best_cost = cost; best = rb; }
        if (rb >= h) break;
    }
    A.rb = best;
    A.seg = A.rb / 4;
    A.nbrow = ceil_div(h, A.rb);
    A.items = planes * A.nbrow * A.pairs;
    CUtensorMap tmY;
    int rc;
    if ((rc = make_plane_map(&tmY, y, planes, 4 * h, 4 * w, 4 * kStripCells, 4))) return rc;
    const int smem = smem_for(A.rb);
    const int per_sm = std::max(1, std::min(NS == 1 ? 4 : 1, (225 * 1024) / (smem + 1024)));
    cudaLaunchConfig_t cfg = {};
    static const int dbgb = []() { const char* v = getenv("ESR_CEM_DBGB"); return v ? atoi(v) : 0; }();   // bisect: 1 one item per CTA, 2 one CTA per SM
    cfg.gridDim = dim3((dbgb & 1) ? A.items : std::min(sms * ((dbgb & 2) ? 1 : per_sm), A.items));
    cfg.blockDim = dim3(kBW * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ESR_ONCE_PER_DEVICE(ESR_CUDA(cudaFuncSetAttribute(cem_invup4s_kernel<27, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)););
    const cudaError_t e = cudaLaunchKernelEx(&cfg, cem_invup4s_kernel<27, NS>, tmY, make_tab(f), K, A);
    if (e != cudaSuccess) { set_error("cem_invup4s_kernel launch failed: %s", cudaGetErrorString(e)); return ESR_ERR_CUDA; }
    return check_launch("cem_invup4s_kernel");
}

// ----------------------------------------------------------------- adjoint (backward)
// Every CEM operator is, per axis, F: out[a] = sum_t taps[t] * src[clamp(sa*a + off + t - pad, 0, Ls-1)]
// (src = the image for Down / K, the zero-stuffed image for Up).  The kernel below evaluates the
// exact adjoint of F, including the fold of the replicate padding onto the border samples,
//   gs[m] = sum_{a,t} [clamp(sa*a + off + t - pad) == m] taps[t] * g[a],
// at the positions m = so*j + po, j in [0,nout), along one axis of a [planes,R,C] array.
struct Adj1dArgs {
    float taps[ESR_CEM_MAX_TAPS];
    int nt, pad, sa, off, na, Ls, so, po, nout;
    int axis;            // 0: along columns (contiguous), 1: along rows
    int planes, other;   // `other` = extent of the untouched axis
    float scale;         // multiplies the result
    const float* g;      // [planes, (axis? na:other), (axis? other:na)]
    float* out;          // [planes, (axis? nout:other), (axis? other:nout)]
    const float* base;   // optional: out = base - result  (same shape as out)
};

// Round 2: taps in shared memory (the tap index differs from lane to lane: as kernel-parameter / constant-bank reads
// those were serialised up to 32x) and 32-bit index arithmetic (the 64-bit div / mod per output was most of the
// instruction count).  The six passes of esr_cem_project_bwd took 1.3 ms of a 10.7 ms Z-optimisation iteration.
__global__ void __launch_bounds__(256) cem_adj1d_kernel(const __grid_constant__ Adj1dArgs a) {
    __shared__ float taps[ESR_CEM_MAX_TAPS];
    for (int t = threadIdx.x; t < a.nt; t += blockDim.x) taps[t] = a.taps[t];
    __syncthreads();
    const uint32_t total = static_cast<uint32_t>(a.planes) * a.other * a.nout;      // host checks that it fits 32 bits
    const uint32_t nout = a.nout, other = a.other;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        uint32_t j, o, p;
        if (a.axis == 0) { const uint32_t q = idx / nout; j = idx - q * nout; p = q / other; o = q - p * other; }
        else { const uint32_t q = idx / other; o = idx - q * other; p = q / nout; j = q - p * nout; }
        const float* gl = a.axis == 0 ? a.g + (static_cast<size_t>(p) * other + o) * a.na : a.g + static_cast<size_t>(p) * a.na * other + o;
        const int gstride = a.axis == 0 ? 1 : static_cast<int>(other);
        const int m = a.so * static_cast<int>(j) + a.po;
        float acc = 0.f;
        // t = m + pad - off - sa*i (interior); borders collect every (i,t) that clamps onto them
        const int c0 = m + a.pad - a.off;
        if (m > 0 && m < a.Ls - 1) {
            int i_lo = (c0 - (a.nt - 1) + a.sa - 1);
            i_lo = i_lo <= 0 ? 0 : i_lo / a.sa;
            int i_hi = c0 < 0 ? -1 : c0 / a.sa;
            if (i_hi > a.na - 1) i_hi = a.na - 1;
            for (int i = i_lo; i <= i_hi; ++i) acc = fmaf(taps[c0 - a.sa * i], gl[static_cast<size_t>(i) * gstride], acc);
        } else {
            // only the samples whose window reaches past the border can clamp onto it (conservative bounds; the
            // exact test is inside): pos <= 0 needs sa*i <= pad - off, pos >= Ls-1 needs sa*i >= Ls-1-off+pad-(nt-1)
            int i_lo = 0, i_hi = a.na - 1;
            if (a.Ls > 1) {
                if (m == 0) { i_hi = (a.pad - a.off) / a.sa + 1; if (i_hi > a.na - 1) i_hi = a.na - 1; }
                else { i_lo = (a.Ls - 1 - a.off + a.pad - (a.nt - 1)) / a.sa - 1; if (i_lo < 0) i_lo = 0; }
            }
            for (int i = i_lo; i <= i_hi; ++i) {
                const int base_pos = a.sa * i + a.off - a.pad;   // position of tap 0
                float wsum = 0.f;
                for (int t = 0; t < a.nt; ++t) {
                    const int pos = base_pos + t;
                    const int cl = pos < 0 ? 0 : (pos > a.Ls - 1 ? a.Ls - 1 : pos);
                    if (cl == m) wsum += taps[t];
                }
                if (wsum != 0.f) acc = fmaf(wsum, gl[static_cast<size_t>(i) * gstride], acc);
            }
        }
        acc *= a.scale;
        a.out[idx] = a.base != nullptr ? a.base[idx] - acc : acc;
    }
}

// g_pad[Y,X] = (crop <= Y < H-crop && ...) ? g[Y-crop, X-crop] : 0   (adjoint of HR_unpadder)
__global__ void cem_pad_zero_kernel(const float* __restrict__ g, float* __restrict__ out, int planes, int H, int W,
                                    int crop) {
    const size_t total = static_cast<size_t>(planes) * H * W;
    const int Ho = H - 2 * crop, Wo = W - 2 * crop;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int X = static_cast<int>(idx % W);
        const int Y = static_cast<int>((idx / W) % H);
        const size_t p = idx / (static_cast<size_t>(W) * H);
        const bool in = Y >= crop && Y < H - crop && X >= crop && X < W - crop;
        out[idx] = in ? g[(p * Ho + (Y - crop)) * Wo + (X - crop)] : 0.f;
    }
}

int cem_pad_zero(const float* g, float* out, int planes, int H, int W, int crop, cudaStream_t s) {
    const size_t want = (static_cast<size_t>(planes) * H * W + 255) / 256;
    const int grid = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
    cem_pad_zero_kernel<<<grid, 256, 0, s>>>(g, out, planes, H, W, crop);
    return check_launch("cem_pad_zero_kernel");
}

static int launch_adj1d(Adj1dArgs& a, cudaStream_t s) {
    const size_t total = static_cast<size_t>(a.planes) * a.other * a.nout;
    ESR_CHECK_ARG(total < (1ull << 31), "CEM adjoint: %zu elements per pass exceed the 32-bit index range", total);
    const size_t want = (total + 255) / 256;
    const int grid = static_cast<int>(want < 148 * 16 ? (want ? want : 1) : 148 * 16);
    cem_adj1d_kernel<<<grid, 256, 0, s>>>(a);
    return check_launch("cem_adj1d_kernel");
}

static size_t down_smem(const esr_cem_filters& f) {
    const int rows_in = (DT_R - 1) * f.sf + f.n_ds, cols_in = (DT_C - 1) * f.sf + f.n_ds;
    return sizeof(float) * (static_cast<size_t>(rows_in) * cols_in + static_cast<size_t>(rows_in) * DT_C);
}
static size_t inv_smem(const esr_cem_filters& f) {
    const int rows_in = IT_R + f.n_inv - 1, cols_in = IT_C + f.n_inv - 1;
    return sizeof(float) * (static_cast<size_t>(rows_in) * cols_in + static_cast<size_t>(rows_in) * IT_C);
}
static size_t up_smem(const esr_cem_filters& f) {
    const int ext = (f.n_ds - 1) / f.sf + 2;
    const int rows_in = UT_R + 2 * ext, cols_in = UT_C + 2 * ext;
    return sizeof(float) * (static_cast<size_t>(rows_in) * cols_in + static_cast<size_t>(rows_in) * UT_C * f.sf);
}

static int check_filters(const esr_cem_filters* f) {
    ESR_CHECK_ARG(f != nullptr, "null CEM filters");
    ESR_CHECK_ARG(f->sf >= 2 && f->sf <= 4 && f->pre >= 0 && f->pre < f->sf, "unsupported CEM scale factor %d", f->sf);
    ESR_CHECK_ARG(f->n_ds > 0 && f->n_ds <= ESR_CEM_MAX_TAPS && (f->n_ds & 1), "bad ds kernel length");
    ESR_CHECK_ARG(f->n_inv > 0 && f->n_inv <= ESR_CEM_MAX_TAPS && (f->n_inv & 1), "bad inv_hTh length");
    return ESR_OK;
}

static int set_smem(const void* fn, size_t bytes) {
    if (bytes > 48 * 1024) ESR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    return ESR_OK;
}

int cem_down(const esr_cem_filters& f, const float* y, const float* x, int planes, int H, int W, float* out,
             cudaStream_t s) {
    ESR_CHECK_ARG(H % f.sf == 0 && W % f.sf == 0, "HR size %dx%d not divisible by %d", H, W, f.sf);
    if (fast4_ok(f, H, W, 0, y, nullptr, nullptr)) {
        const int seg = pick_seg_down(planes, H / 4, W / 4);
        dim3 grid(ceil_div(W / 4, kStripCells), ceil_div(ceil_div(H / 4, seg), 4), planes);
        cem_down4_kernel<<<grid, 128, 0, s>>>(make_tab(f), y, x, out, H, W, seg);
        return check_launch("cem_down4_kernel");
    }
    const size_t sm = down_smem(f);
    dim3 grid(ceil_div(W / f.sf, DT_C), ceil_div(H / f.sf, DT_R), planes);
    auto launch = [&](auto kern) -> int {
        int rc = set_smem(reinterpret_cast<const void*>(kern), sm);
        if (rc) return rc;
        kern<<<grid, 256, sm, s>>>(f, y, x, out, H, W);
        return check_launch("cem_down_kernel");
    };
    return f.sf == 2 ? launch(cem_down_kernel<2>) : f.sf == 3 ? launch(cem_down_kernel<3>) : launch(cem_down_kernel<4>);
}
int cem_inv(const esr_cem_filters& f, const float* x, int planes, int h, int w, float* out, cudaStream_t s) {
    const size_t sm = inv_smem(f);
    int rc = set_smem(reinterpret_cast<const void*>(cem_inv_kernel), sm);
    if (rc) return rc;
    dim3 grid(ceil_div(w, IT_C), ceil_div(h, IT_R), planes);
    cem_inv_kernel<<<grid, 256, sm, s>>>(f, x, out, h, w);
    return check_launch("cem_inv_kernel");
}
int cem_up(const esr_cem_filters& f, const float* x, const float* y, int planes, int h, int w, int crop, float sign,
           float* out, cudaStream_t s) {
    if (fast4_ok(f, h * 4, w * 4, crop, y, out, nullptr) && ((w * 4 - 2 * crop) % 4 == 0)) {
        dim3 grid(ceil_div(w, kStripCells), ceil_div(ceil_div(h, kSegRows), 4), planes);
        cem_up4_kernel<<<grid, 128, 0, s>>>(make_tab(f), x, y, out, h, w, crop, sign);
        return check_launch("cem_up4_kernel");
    }
    const size_t sm = up_smem(f);
    dim3 grid(ceil_div(w, UT_C), ceil_div(h, UT_R), planes);
    auto launch = [&](auto kern) -> int {
        int rc = set_smem(reinterpret_cast<const void*>(kern), sm);
        if (rc) return rc;
        kern<<<grid, 256, sm, s>>>(f, x, y, out, h, w, crop, sign);
        return check_launch("cem_up_kernel");
    };
    return f.sf == 2 ? launch(cem_up_kernel<2>) : f.sf == 3 ? launch(cem_up_kernel<3>) : launch(cem_up_kernel<4>);
}

// Fused (HH^T)^-1 + Up + residual add, x4 only; d = x - Down(y).
int cem_invup4(const esr_cem_filters& f, const float* d, const float* y, int planes, int h, int w, int crop, float* out,
               cudaStream_t s, bool after_down) {
    InvTaps K;
    K.n = f.n_inv;
    for (int i = 0; i < f.n_inv; ++i) { K.t[i] = f.inv[i]; K.t2[i] = make_float2(f.inv[i], f.inv[i]); }
    const int seg = pick_seg_up(planes, h, w), pad = f.n_inv / 2;   // <= 16: smem tile
    const int Re = 4 * seg + 4, Rd = Re + 2 * pad, Cd = 32 + 2 * pad, Cds = (Cd + 3) & ~3;
    const size_t sm = sizeof(float) * (static_cast<size_t>(Rd) * Cds + static_cast<size_t>(Rd) * 32 + static_cast<size_t>(Re) * 32);
    dim3 grid(ceil_div(w, kStripCells), ceil_div(h, 4 * seg), planes);
    // Programmatic dependent launch on the Down kernel that produced d: the blocks become resident while Down's last
    // blocks drain, issue their y prefetch, and only then wait for d (griddepcontrol.wait in the kernel).
    auto launch = [&](auto kern) -> int {
        int rc = set_smem(reinterpret_cast<const void*>(kern), sm);
        if (rc) return rc;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = sm;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = after_down ? 1 : 0;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, make_tab(f), K, d, y, out, h, w, crop, seg);
        if (e != cudaSuccess) { set_error("cem_invup4_kernel launch failed: %s", cudaGetErrorString(e)); return ESR_ERR_CUDA; }
        return ESR_OK;
    };
    int rc;
    if (f.n_inv == 27) rc = launch(cem_invup4_kernel<27, 4>);      // 5 blocks / SM (96 registers) measured 7-19 % slower
    else rc = launch(cem_invup4_kernel<0, 4>);
    if (rc) return rc;
    return check_launch("cem_invup4_kernel");
}

}  // namespace esr

using namespace esr;

extern "C" int esr_debug_cem_timeout(uint32_t* out4) {
    if (out4 == nullptr) { set_error("null pointer"); return ESR_ERR_INVALID; }
    unsigned int zero[4] = {0, 0, 0, 0};
    ESR_CUDA(cudaMemcpyFromSymbol(out4, g_cem_timeout, sizeof(zero)));
    ESR_CUDA(cudaMemcpyToSymbol(g_cem_timeout, zero, sizeof(zero)));
    return ESR_OK;
}

extern "C" int esr_cem_downscale(const esr_cem_filters* f, const float* y, int32_t B, int32_t C, int32_t H, int32_t W,
                                 float* out, void* stream) {
    int rc = check_filters(f);
    if (rc) return rc;
    ESR_CHECK_ARG(y && out && B > 0 && C > 0, "esr_cem_downscale: bad arguments");
    return cem_down(*f, y, nullptr, B * C, H, W, out, static_cast<cudaStream_t>(stream));
}

extern "C" int esr_cem_inv_hth(const esr_cem_filters* f, const float* x, int32_t B, int32_t C, int32_t h, int32_t w,
                               float* out, void* stream) {
    int rc = check_filters(f);
    if (rc) return rc;
    ESR_CHECK_ARG(x && out && B > 0 && C > 0 && h > 0 && w > 0, "esr_cem_inv_hth: bad arguments");
    return cem_inv(*f, x, B * C, h, w, out, static_cast<cudaStream_t>(stream));
}

extern "C" int esr_cem_upscale(const esr_cem_filters* f, const float* x, int32_t B, int32_t C, int32_t h, int32_t w,
                               float* out, void* stream) {
    int rc = check_filters(f);
    if (rc) return rc;
    ESR_CHECK_ARG(x && out && B > 0 && C > 0 && h > 0 && w > 0, "esr_cem_upscale: bad arguments");
    return cem_up(*f, x, nullptr, B * C, h, w, 0, 1.f, out, static_cast<cudaStream_t>(stream));
}

extern "C" int esr_cem_project(const esr_cem_filters* f, const float* y, const float* x, int32_t B, int32_t C,
                               int32_t H, int32_t W, int32_t crop, float* out, float* workspace, void* stream) {
    int rc = check_filters(f);
    if (rc) return rc;
    ESR_CHECK_ARG(y && x && out && workspace && B > 0 && C > 0, "esr_cem_project: bad arguments");
    ESR_CHECK_ARG(crop >= 0 && 2 * crop < H && 2 * crop < W, "esr_cem_project: crop %d too large", crop);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int h = H / f->sf, w = W / f->sf, planes = B * C;
    float* d = workspace;
    float* e = workspace + static_cast<size_t>(planes) * h * w;
    if (cem_fused_enabled() && (rc = cem_project4f(*f, y, x, planes, H, W, crop, out, workspace, s)) <= 0) return rc;   // opt-in: one launch, y read once
    if (stream4_ok(*f, H, W, crop, y, x, out, d)) {                         // round-2 TMA streaming kernels
        static const int dbg = []() { const char* v = getenv("ESR_CEM_DEBUG"); return v ? atoi(v) : 0; }();   // 1: new Down only, 2: new K+Up only
        if ((rc = dbg == 2 ? cem_down(*f, y, x, planes, H, W, d, s) : cem_down4s(*f, y, x, planes, H, W, d, s))) return rc;
        if (dbg == 1) return cem_invup4(*f, d, y, planes, h, w, crop, out, s, false);
        return cem_invup4s(*f, d, y, planes, h, w, crop, out, s);
    }
    if ((rc = cem_down(*f, y, x, planes, H, W, d, s))) return rc;
    if (fast4_ok(*f, H, W, crop, y, out, d) && ((W - 2 * crop) % 4 == 0))
        return cem_invup4(*f, d, y, planes, h, w, crop, out, s, true);    // two launches: Down, then K + Up + add
    if ((rc = cem_inv(*f, d, planes, h, w, e, s))) return rc;
    return cem_up(*f, e, y, planes, h, w, crop, 1.f, out, s);
}

extern "C" int esr_cem_project_bwd(const esr_cem_filters* f, const float* g_out, int32_t B, int32_t C, int32_t H,
                                   int32_t W, int32_t crop, float* g_y, float* workspace, void* stream) {
    int rc = check_filters(f);
    if (rc) return rc;
    ESR_CHECK_ARG(g_out && g_y && workspace && B > 0 && C > 0, "esr_cem_project_bwd: bad arguments");
    ESR_CHECK_ARG(H % f->sf == 0 && W % f->sf == 0 && crop >= 0 && 2 * crop < H && 2 * crop < W,
                  "esr_cem_project_bwd: bad geometry");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int sf = f->sf, h = H / sf, w = W / sf, planes = B * C;
    const size_t nHR = static_cast<size_t>(planes) * H * W;
    float* Gp = workspace;                                             // [planes,H,W]  zero-padded g_out
    float* tA = Gp + nHR;                                              // [planes,H,w]
    float* tB = tA + static_cast<size_t>(planes) * H * w;              // [planes,h,w]
    float* tC = tB + static_cast<size_t>(planes) * h * w;              // [planes,h,w]
    if ((rc = cem_pad_zero(g_out, Gp, planes, H, W, crop, s))) return rc;
    Adj1dArgs a;
    // Up^T: Up = correlate(u = sf*ds) over the replicate-padded zero-stuffed image; sample at sf*i+pre.
    for (int t = 0; t < f->n_ds; ++t) a.taps[t] = f->ds[t] * sf;
    a.nt = f->n_ds; a.pad = f->n_ds / 2; a.sa = 1; a.off = 0; a.so = sf; a.po = f->pre; a.scale = 1.f; a.base = nullptr;
    a.planes = planes;
    a.axis = 0; a.other = H; a.na = W; a.Ls = W; a.nout = w; a.g = Gp; a.out = tA;
    if ((rc = launch_adj1d(a, s))) return rc;
    a.axis = 1; a.other = w; a.na = H; a.Ls = H; a.nout = h; a.g = tA; a.out = tB;
    if ((rc = launch_adj1d(a, s))) return rc;
    // K^T
    for (int t = 0; t < f->n_inv; ++t) a.taps[t] = f->inv[t];
    a.nt = f->n_inv; a.pad = f->n_inv / 2; a.sa = 1; a.off = 0; a.so = 1; a.po = 0;
    a.axis = 0; a.other = h; a.na = w; a.Ls = w; a.nout = w; a.g = tB; a.out = tC;
    if ((rc = launch_adj1d(a, s))) return rc;
    a.axis = 1; a.other = w; a.na = h; a.Ls = h; a.nout = h; a.g = tC; a.out = tB;
    if ((rc = launch_adj1d(a, s))) return rc;
    // Down^T: Down = correlate(flipped ds) over the replicate-padded image, sampled at sf*i+pre.
    for (int t = 0; t < f->n_ds; ++t) a.taps[t] = f->ds[f->n_ds - 1 - t];
    a.nt = f->n_ds; a.pad = f->n_ds / 2; a.sa = sf; a.off = f->pre; a.so = 1; a.po = 0;
    a.axis = 1; a.other = w; a.na = h; a.Ls = H; a.nout = H; a.g = tB; a.out = tA;
    if ((rc = launch_adj1d(a, s))) return rc;
    a.axis = 0; a.other = H; a.na = w; a.Ls = W; a.nout = W; a.g = tA; a.out = g_y; a.base = Gp;
    return launch_adj1d(a, s);
}
