// CEM operators for general (non-separable) filters: direct 2-D stencils, fp32, NCHW planes.
//
// The default bicubic configuration never gets here (cem.cu runs it as rank-1 passes).  Non-default
// kernels -- an estimated / user-supplied downscaling kernel (imresize_CEM.py:22-32) or
// blurry_cubic_<sigma> with a strong blur (:37-41), for which the magnitude clamp of
// compute_inv_hTh (CEMnet.py:112) bites -- give a ds_kernel and / or an inv_hTh that is not an outer
// product, and the three operators
//   Down  : out[i,j] = sum rot180(ds)[tr,tc] * y[clamp(sf*i+pre+tr-p), clamp(sf*j+pre+tc-p)]   (CEMnet.py:157-162)
//   K     : out[i,j] = sum inv[tr,tc]       * x[clamp(i+tr-q), clamp(j+tc-q)]                  (CEMnet.py:149-151)
//   Up    : out[Y,X] = sum sf^2 ds[tr,tc]   * z[clamp(Y+tr-p), clamp(X+tc-p)],  z = zero-stuffed x  (CEMnet.py:153-159)
// run from a shared-memory tile with the taps staged in shared memory too.  The adjoint used by the
// data-gradient backward is one generic kernel (every operator is the same clamped strided
// correlation per axis, see cem.cu's 1-D adjoint).
#include "esr_common.cuh"

namespace esr {

int cem_pad_zero(const float* g, float* out, int planes, int H, int W, int crop, cudaStream_t s);   // cem.cu

namespace {

__device__ __forceinline__ int clamp2(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__host__ __device__ __forceinline__ int floordiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// ------------------------------------------------------------------------ Down
// Block: 8 x 32 LR outputs of one plane.  The HR tile is stored de-interleaved by column phase
// ([row][col % sf][col / sf]) so that the 32 lanes of a warp (consecutive LR columns, HR stride sf)
// read consecutive shared-memory words.
constexpr int D2_R = 8, D2_C = 32;

__global__ void __launch_bounds__(256) cem2d_down_kernel(const float* __restrict__ ds, int n, int sf, int pre,
                                                         const float* __restrict__ y, const float* __restrict__ x,
                                                         float* __restrict__ out, int H, int W) {
    extern __shared__ float sm[];
    const int pad = n / 2, h = H / sf, w = W / sf;
    const int plane = blockIdx.z, i0 = blockIdx.y * D2_R, j0 = blockIdx.x * D2_C;
    const int rows_in = (D2_R - 1) * sf + n, cols_in = (D2_C - 1) * sf + n;
    const int cph = (cols_in + sf - 1) / sf;
    float* taps = sm;                    // rot180(ds): the correlation taps of Down
    float* tile = sm + n * n;            // [rows_in][sf][cph]
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) taps[idx] = __ldg(ds + n * n - 1 - idx);
    const float* yp = y + static_cast<size_t>(plane) * H * W;
    const int r_base = i0 * sf + pre - pad, c_base = j0 * sf + pre - pad;
    for (int idx = threadIdx.x; idx < rows_in * cols_in; idx += blockDim.x) {
        const int r = idx / cols_in, c = idx - r * cols_in;
        tile[(r * sf + c % sf) * cph + c / sf] =
            __ldg(yp + static_cast<size_t>(clamp2(r_base + r, 0, H - 1)) * W + clamp2(c_base + c, 0, W - 1));
    }
    __syncthreads();
    const int i = threadIdx.x / D2_C, j = threadIdx.x % D2_C;
    if (i0 + i >= h || j0 + j >= w) return;
    float acc = 0.f;
    for (int tr = 0; tr < n; ++tr) {
        const float* trow = tile + (i * sf + tr) * sf * cph + j;
        const float* tp = taps + tr * n;
        for (int ph = 0; ph < sf; ++ph)
            for (int q = 0, tc = ph; tc < n; ++q, tc += sf) acc = fmaf(tp[tc], trow[ph * cph + q], acc);
    }
    const size_t o = (static_cast<size_t>(plane) * h + i0 + i) * w + j0 + j;
    out[o] = x != nullptr ? x[o] - acc : acc;
}

// --------------------------------------------------------------------------- K
// Block: 32 x 32 outputs, 256 threads, four rows (i, i+8, i+16, i+24) per thread sharing each tap.
constexpr int I2_R = 32, I2_C = 32;

__global__ void __launch_bounds__(256) cem2d_inv_kernel(const float* __restrict__ inv, int n, const float* __restrict__ x,
                                                        float* __restrict__ out, int h, int w) {
    extern __shared__ float sm[];
    const int pad = n / 2;
    const int plane = blockIdx.z, i0 = blockIdx.y * I2_R, j0 = blockIdx.x * I2_C;
    const int rows_in = I2_R + n - 1, cols_in = I2_C + n - 1;
    float* taps = sm;
    float* tile = sm + n * n;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) taps[idx] = __ldg(inv + idx);
    const float* xp = x + static_cast<size_t>(plane) * h * w;
    for (int idx = threadIdx.x; idx < rows_in * cols_in; idx += blockDim.x) {
        const int r = idx / cols_in, c = idx - r * cols_in;
        tile[idx] = __ldg(xp + static_cast<size_t>(clamp2(i0 + r - pad, 0, h - 1)) * w + clamp2(j0 + c - pad, 0, w - 1));
    }
    __syncthreads();
    const int i = threadIdx.x / I2_C, j = threadIdx.x % I2_C;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int tr = 0; tr < n; ++tr) {
        const float* t0 = tile + (i + tr) * cols_in + j;
        const float* tp = taps + tr * n;
        for (int tc = 0; tc < n; ++tc) {
            const float t = tp[tc];
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = fmaf(t, t0[k * 8 * cols_in + tc], acc[k]);
        }
    }
    if (j0 + j >= w) return;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (i0 + i + 8 * k < h) out[(static_cast<size_t>(plane) * h + i0 + i + 8 * k) * w + j0 + j] = acc[k];
}

// -------------------------------------------------------------------------- Up
// out[Y-crop, X-crop] = (y ? y[Y,X] : 0) + sign * Up(x)[Y,X].  Block: 8 x 32 LR cells -> (8 sf) x (32 sf) HR pixels.
// Away from the border only every sf-th tap of a row / column meets a sample (polyphase); within
// n/2 pixels of the border the replicate padding of the zero-stuffed image breaks that pattern and
// every tap is tested.
constexpr int U2_R = 8, U2_C = 32;

__global__ void __launch_bounds__(256) cem2d_up_kernel(const float* __restrict__ ds, int n, int sf, int pre,
                                                       const float* __restrict__ x, const float* __restrict__ y,
                                                       float* __restrict__ out, int h, int w, int crop, float sign) {
    extern __shared__ float sm[];
    const int pad = n / 2, H = h * sf, W = w * sf;
    const int plane = blockIdx.z, I0 = blockIdx.y * U2_R, J0 = blockIdx.x * U2_C;
    const int ext = (n - 1) / sf + 2;
    const int rows_in = U2_R + 2 * ext, cols_in = U2_C + 2 * ext;
    float* taps = sm;                    // sf^2 * ds
    float* tile = sm + n * n;            // x[I0-ext.., J0-ext..], zero outside the image
    const float gain = static_cast<float>(sf * sf);
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) taps[idx] = gain * __ldg(ds + idx);
    const float* xp = x + static_cast<size_t>(plane) * h * w;
    for (int idx = threadIdx.x; idx < rows_in * cols_in; idx += blockDim.x) {
        const int r = idx / cols_in, c = idx - r * cols_in;
        const int rr = I0 - ext + r, cc = J0 - ext + c;
        tile[idx] = (rr >= 0 && rr < h && cc >= 0 && cc < w) ? __ldg(xp + static_cast<size_t>(rr) * w + cc) : 0.f;
    }
    __syncthreads();
    const int Hout = H - 2 * crop, Wout = W - 2 * crop;
    const int hr = U2_R * sf, wc = U2_C * sf;
    for (int idx = threadIdx.x; idx < hr * wc; idx += blockDim.x) {
        const int ry = idx / wc, cx = idx - ry * wc;
        const int Y = I0 * sf + ry, X = J0 * sf + cx;
        if (Y < crop || Y >= H - crop || X < crop || X >= W - crop) continue;
        float acc = 0.f;
        const bool interior = Y >= pad && Y + (n - 1 - pad) <= H - 1 && X >= pad && X + (n - 1 - pad) <= W - 1;
        if (interior) {
            const int tr0 = ((pre + pad - Y) % sf + sf) % sf, tc0 = ((pre + pad - X) % sf + sf) % sf;
            const int qc0 = (X + tc0 - pad - pre) / sf - (J0 - ext);
            for (int tr = tr0; tr < n; tr += sf) {
                const float* trow = tile + ((Y + tr - pad - pre) / sf - (I0 - ext)) * cols_in + qc0;
                const float* tp = taps + tr * n;
                for (int tc = tc0, q = 0; tc < n; tc += sf, ++q) acc = fmaf(tp[tc], trow[q], acc);
            }
        } else {
            for (int tr = 0; tr < n; ++tr) {
                const int qr = clamp2(Y + tr - pad, 0, H - 1) - pre;
                if (qr < 0 || qr % sf != 0) continue;
                const float* trow = tile + (qr / sf - (I0 - ext)) * cols_in - (J0 - ext);
                const float* tp = taps + tr * n;
                for (int tc = 0; tc < n; ++tc) {
                    const int qc = clamp2(X + tc - pad, 0, W - 1) - pre;
                    if (qc >= 0 && qc % sf == 0) acc = fmaf(tp[tc], trow[qc / sf], acc);
                }
            }
        }
        const float base = y != nullptr ? __ldg(y + (static_cast<size_t>(plane) * H + Y) * W + X) : 0.f;
        out[(static_cast<size_t>(plane) * Hout + (Y - crop)) * Wout + (X - crop)] = base + sign * acc;
    }
}

// --------------------------------------------------------------------- adjoint
// Forward operator, per axis: F: out[a] = sum_t taps[t] * src[clamp(sa*a + off + t - pad, 0, Ls-1)].
// Adjoint evaluated at the source positions m = so*j + po:
//   gs[mr,mc] = sum_{(ar,tr) in S(mr)} sum_{(ac,tc) in S(mc)} taps[tr,tc] * g[ar,ac],
//   S(m) = {(a,t): clamp(sa*a + off + t - pad) == m}.
// For an interior m every a contributes exactly one t; on the two border positions the whole run of
// taps hanging over the edge folds onto m.
struct Adj2dArgs {
    const float* taps;   // [n*n], the forward correlation taps before `flip` / `gain`
    int n, pad, flip;    // flip: use taps[n*n-1-idx] (Down correlates with rot180(ds))
    float gain;
    int sa, off, so, po;
    int na_r, na_c;      // extent of g
    int Ls_r, Ls_c;      // extent of the forward operator's source
    int nout_r, nout_c;  // extent of out
    int planes;
    const float* g;
    float* out;
    const float* base;   // optional: out = base - result
};

struct AxisRange { int a_lo, a_hi; };

__device__ __forceinline__ AxisRange adj_axis(int m, int Ls, int n, int pad, int sa, int off, int na) {
    const int c0 = m + pad - off;
    AxisRange r;
    r.a_lo = m == 0 ? 0 : floordiv(c0 - (n - 1) + sa - 1, sa);
    r.a_hi = m == Ls - 1 ? na - 1 : floordiv(c0, sa);
    if (r.a_lo < 0) r.a_lo = 0;
    if (r.a_hi > na - 1) r.a_hi = na - 1;
    return r;
}
// taps of output sample `a` that land on m: [t_lo, t_hi] (possibly empty)
__device__ __forceinline__ void adj_taps(int m, int Ls, int n, int bp, int& t_lo, int& t_hi) {
    t_lo = m == 0 ? 0 : m - bp;
    t_hi = m == Ls - 1 ? n - 1 : m - bp;
    if (t_lo < 0) t_lo = 0;
    if (t_hi > n - 1) t_hi = n - 1;
}

__global__ void __launch_bounds__(256) cem2d_adj_kernel(const __grid_constant__ Adj2dArgs a) {
    extern __shared__ float taps[];
    const int nn = a.n * a.n;
    for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) taps[idx] = a.gain * __ldg(a.taps + (a.flip ? nn - 1 - idx : idx));
    __syncthreads();
    const size_t per_plane = static_cast<size_t>(a.nout_r) * a.nout_c;
    const size_t total = per_plane * a.planes;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int jc = static_cast<int>(idx % a.nout_c);
        const int jr = static_cast<int>((idx / a.nout_c) % a.nout_r);
        const size_t p = idx / per_plane;
        const float* gp = a.g + p * static_cast<size_t>(a.na_r) * a.na_c;
        const int mr = a.so * jr + a.po, mc = a.so * jc + a.po;
        const AxisRange R = adj_axis(mr, a.Ls_r, a.n, a.pad, a.sa, a.off, a.na_r);
        const AxisRange Cc = adj_axis(mc, a.Ls_c, a.n, a.pad, a.sa, a.off, a.na_c);
        const bool int_r = mr > 0 && mr < a.Ls_r - 1, int_c = mc > 0 && mc < a.Ls_c - 1;
        float acc = 0.f;
        if (int_r && int_c) {
            const int tr0 = mr - (a.sa * R.a_lo + a.off - a.pad), tc0 = mc - (a.sa * Cc.a_lo + a.off - a.pad);
            for (int ar = R.a_lo, tr = tr0; ar <= R.a_hi; ++ar, tr -= a.sa) {
                const float* gl = gp + static_cast<size_t>(ar) * a.na_c;
                const float* tp = taps + tr * a.n;
                for (int ac = Cc.a_lo, tc = tc0; ac <= Cc.a_hi; ++ac, tc -= a.sa) acc = fmaf(tp[tc], __ldg(gl + ac), acc);
            }
        } else {
            for (int ar = R.a_lo; ar <= R.a_hi; ++ar) {
                int tr_lo, tr_hi;
                adj_taps(mr, a.Ls_r, a.n, a.sa * ar + a.off - a.pad, tr_lo, tr_hi);
                if (tr_lo > tr_hi) continue;
                const float* gl = gp + static_cast<size_t>(ar) * a.na_c;
                for (int ac = Cc.a_lo; ac <= Cc.a_hi; ++ac) {
                    int tc_lo, tc_hi;
                    adj_taps(mc, a.Ls_c, a.n, a.sa * ac + a.off - a.pad, tc_lo, tc_hi);
                    float wsum = 0.f;
                    for (int tr = tr_lo; tr <= tr_hi; ++tr)
                        for (int tc = tc_lo; tc <= tc_hi; ++tc) wsum += taps[tr * a.n + tc];
                    acc = fmaf(wsum, __ldg(gl + ac), acc);
                }
            }
        }
        a.out[idx] = a.base != nullptr ? a.base[idx] - acc : acc;
    }
}

int set_smem2(const void* fn, size_t bytes) {
    ESR_CHECK_ARG(bytes <= 227 * 1024, "CEM 2-D stencil needs %zu bytes of shared memory", bytes);
    if (bytes > 48 * 1024) ESR_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    return ESR_OK;
}

int check_filters2d(const esr_cem_filters2d* f) {
    ESR_CHECK_ARG(f != nullptr && f->ds != nullptr && f->inv != nullptr, "null CEM 2-D filters");
    ESR_CHECK_ARG(f->sf >= 2 && f->sf <= 4 && f->pre >= 0 && f->pre < f->sf, "unsupported CEM scale factor %d", f->sf);
    ESR_CHECK_ARG(f->n_ds > 0 && f->n_ds <= ESR_CEM2D_MAX_SIDE && (f->n_ds & 1), "bad ds kernel side %d (odd, <= %d)", f->n_ds,
                  ESR_CEM2D_MAX_SIDE);
    ESR_CHECK_ARG(f->n_inv > 0 && f->n_inv <= ESR_CEM2D_MAX_SIDE && (f->n_inv & 1), "bad inv_hTh side %d (odd, <= %d)", f->n_inv,
                  ESR_CEM2D_MAX_SIDE);
    return ESR_OK;
}

int down2d(const esr_cem_filters2d& f, const float* y, const float* x, int planes, int H, int W, float* out, cudaStream_t s) {
    ESR_CHECK_ARG(H > 0 && W > 0 && H % f.sf == 0 && W % f.sf == 0, "HR size %dx%d not divisible by %d", H, W, f.sf);
    const int n = f.n_ds, rows_in = (D2_R - 1) * f.sf + n, cols_in = (D2_C - 1) * f.sf + n;
    const size_t sm = sizeof(float) * (static_cast<size_t>(n) * n + static_cast<size_t>(rows_in) * f.sf * ceil_div(cols_in, f.sf));
    int rc = set_smem2(reinterpret_cast<const void*>(cem2d_down_kernel), sm);
    if (rc) return rc;
    dim3 grid(ceil_div(W / f.sf, D2_C), ceil_div(H / f.sf, D2_R), planes);
    cem2d_down_kernel<<<grid, 256, sm, s>>>(f.ds, n, f.sf, f.pre, y, x, out, H, W);
    return check_launch("cem2d_down_kernel");
}

int inv2d(const esr_cem_filters2d& f, const float* x, int planes, int h, int w, float* out, cudaStream_t s) {
    const int n = f.n_inv;
    const size_t sm = sizeof(float) * (static_cast<size_t>(n) * n + static_cast<size_t>(I2_R + n - 1) * (I2_C + n - 1));
    int rc = set_smem2(reinterpret_cast<const void*>(cem2d_inv_kernel), sm);
    if (rc) return rc;
    dim3 grid(ceil_div(w, I2_C), ceil_div(h, I2_R), planes);
    cem2d_inv_kernel<<<grid, 256, sm, s>>>(f.inv, n, x, out, h, w);
    return check_launch("cem2d_inv_kernel");
}

int up2d(const esr_cem_filters2d& f, const float* x, const float* y, int planes, int h, int w, int crop, float sign, float* out,
         cudaStream_t s) {
    const int n = f.n_ds, ext = (n - 1) / f.sf + 2;
    const size_t sm = sizeof(float) * (static_cast<size_t>(n) * n + static_cast<size_t>(U2_R + 2 * ext) * (U2_C + 2 * ext));
    int rc = set_smem2(reinterpret_cast<const void*>(cem2d_up_kernel), sm);
    if (rc) return rc;
    dim3 grid(ceil_div(w, U2_C), ceil_div(h, U2_R), planes);
    cem2d_up_kernel<<<grid, 256, sm, s>>>(f.ds, n, f.sf, f.pre, x, y, out, h, w, crop, sign);
    return check_launch("cem2d_up_kernel");
}

int adj2d(const Adj2dArgs& a, cudaStream_t s) {
    const size_t total = static_cast<size_t>(a.planes) * a.nout_r * a.nout_c;
    const size_t want = (total + 255) / 256;
    const int grid = static_cast<int>(want < 148 * 16 ? (want ? want : 1) : 148 * 16);
    cem2d_adj_kernel<<<grid, 256, sizeof(float) * a.n * a.n, s>>>(a);
    return check_launch("cem2d_adj_kernel");
}

}  // namespace
}  // namespace esr

using namespace esr;

extern "C" int esr_cem2d_downscale(const esr_cem_filters2d* f, const float* y, int32_t B, int32_t C, int32_t H, int32_t W,
                                   float* out, void* stream) {
    int rc = check_filters2d(f);
    if (rc) return rc;
    ESR_CHECK_ARG(y && out && B > 0 && C > 0, "esr_cem2d_downscale: bad arguments");
    return down2d(*f, y, nullptr, B * C, H, W, out, static_cast<cudaStream_t>(stream));
}

extern "C" int esr_cem2d_inv_hth(const esr_cem_filters2d* f, const float* x, int32_t B, int32_t C, int32_t h, int32_t w,
                                 float* out, void* stream) {
    int rc = check_filters2d(f);
    if (rc) return rc;
    ESR_CHECK_ARG(x && out && B > 0 && C > 0 && h > 0 && w > 0, "esr_cem2d_inv_hth: bad arguments");
    return inv2d(*f, x, B * C, h, w, out, static_cast<cudaStream_t>(stream));
}

extern "C" int esr_cem2d_upscale(const esr_cem_filters2d* f, const float* x, int32_t B, int32_t C, int32_t h, int32_t w,
                                 float* out, void* stream) {
    int rc = check_filters2d(f);
    if (rc) return rc;
    ESR_CHECK_ARG(x && out && B > 0 && C > 0 && h > 0 && w > 0, "esr_cem2d_upscale: bad arguments");
    return up2d(*f, x, nullptr, B * C, h, w, 0, 1.f, out, static_cast<cudaStream_t>(stream));
}

extern "C" int esr_cem2d_project(const esr_cem_filters2d* f, const float* y, const float* x, int32_t B, int32_t C, int32_t H,
                                 int32_t W, int32_t crop, float* out, float* workspace, void* stream) {
    int rc = check_filters2d(f);
    if (rc) return rc;
    ESR_CHECK_ARG(y && x && out && workspace && B > 0 && C > 0, "esr_cem2d_project: bad arguments");
    ESR_CHECK_ARG(crop >= 0 && 2 * crop < H && 2 * crop < W, "esr_cem2d_project: crop %d too large", crop);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int h = H / f->sf, w = W / f->sf, planes = B * C;
    float* d = workspace;                                              // x - Down(y)
    float* e = workspace + static_cast<size_t>(planes) * h * w;        // K * d
    if ((rc = down2d(*f, y, x, planes, H, W, d, s))) return rc;
    if ((rc = inv2d(*f, d, planes, h, w, e, s))) return rc;
    return up2d(*f, e, y, planes, h, w, crop, 1.f, out, s);
}

extern "C" int esr_cem2d_project_bwd(const esr_cem_filters2d* f, const float* g_out, int32_t B, int32_t C, int32_t H,
                                     int32_t W, int32_t crop, float* g_y, float* workspace, void* stream) {
    int rc = check_filters2d(f);
    if (rc) return rc;
    ESR_CHECK_ARG(g_out && g_y && workspace && B > 0 && C > 0, "esr_cem2d_project_bwd: bad arguments");
    ESR_CHECK_ARG(H > 0 && W > 0 && H % f->sf == 0 && W % f->sf == 0 && crop >= 0 && 2 * crop < H && 2 * crop < W,
                  "esr_cem2d_project_bwd: bad geometry");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int sf = f->sf, h = H / sf, w = W / sf, planes = B * C;
    float* Gp = workspace;                                             // [planes,H,W] zero-padded g_out
    float* tA = Gp + static_cast<size_t>(planes) * H * W;              // [planes,h,w]
    float* tB = tA + static_cast<size_t>(planes) * h * w;              // [planes,h,w]
    if ((rc = cem_pad_zero(g_out, Gp, planes, H, W, crop, s))) return rc;
    Adj2dArgs a;
    a.planes = planes; a.base = nullptr;
    // Up^T, evaluated on the sample positions sf*j + pre of the zero-stuffed image
    a.taps = f->ds; a.n = f->n_ds; a.pad = f->n_ds / 2; a.flip = 0; a.gain = static_cast<float>(sf * sf);
    a.sa = 1; a.off = 0; a.so = sf; a.po = f->pre;
    a.na_r = H; a.na_c = W; a.Ls_r = H; a.Ls_c = W; a.nout_r = h; a.nout_c = w; a.g = Gp; a.out = tA;
    if ((rc = adj2d(a, s))) return rc;
    // K^T
    a.taps = f->inv; a.n = f->n_inv; a.pad = f->n_inv / 2; a.gain = 1.f; a.so = 1; a.po = 0;
    a.na_r = h; a.na_c = w; a.Ls_r = h; a.Ls_c = w; a.g = tA; a.out = tB;
    if ((rc = adj2d(a, s))) return rc;
    // g_y = Gp - Down^T(...)
    a.taps = f->ds; a.n = f->n_ds; a.pad = f->n_ds / 2; a.flip = 1; a.sa = sf; a.off = f->pre;
    a.Ls_r = H; a.Ls_c = W; a.nout_r = H; a.nout_c = W; a.g = tB; a.out = g_y; a.base = Gp;
    return adj2d(a, s);
}
