// Z optimisation plumbing around G+CEM as four HBM-bound kernels, so that one iteration of the editing loop
// (codes/Z_optimization.py:572-635) is a single CUDA graph with no ATen kernel in it:
//   esr_zopt_tanh_pack  Optimizable_Z.forward (:300-305): clamp to finite, Z_range * tanh(Z), written straight into the
//                       latent channels of the packed model input (SRRaGAN_model.py:249-255: the raw .view is the
//                       identity on memory, so this is one contiguous elementwise pass per image)
//   esr_zopt_loss       the scalar objectives built on fake_H: 'TV' (:474-475, :618-619, TV_Loss :322-324), 'max_STD' /
//                       'min_STD' / 'STD_increase' / 'STD_decrease' (:426-435, :603-607, Masked_STD :525-535, global form):
//                       per-image sums (deterministic two-stage reduction), the loss values, and the coefficients of
//   esr_zopt_loss_grad  dL/d fake_H (what autograd derives in the reference), written into the backward's input buffer
//   esr_zopt_adam       d tanh (Z_range * (1 - tanh^2)) chained with torch.optim.Adam's update (:512, defaults: betas
//                       (0.9, 0.999), eps 1e-8, no weight decay), in place on Z / exp_avg / exp_avg_sq
#include <cfloat>

#include "esr_common.cuh"

namespace esr {

constexpr int kZBlock = 256;

__global__ void zopt_tanh_pack_kernel(float* __restrict__ Z, float z_range, int n_lat, int n_img, float* __restrict__ mi) {
    const int b = blockIdx.y;
    float* zb = Z + static_cast<size_t>(b) * n_lat;
    float* ob = mi + static_cast<size_t>(b) * n_img;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n_lat; i += gridDim.x * blockDim.x * 4) {
        float4 z = *reinterpret_cast<const float4*>(zb + i);
        float4 c = make_float4(fminf(fmaxf(z.x, -FLT_MAX), FLT_MAX), fminf(fmaxf(z.y, -FLT_MAX), FLT_MAX),
                               fminf(fmaxf(z.z, -FLT_MAX), FLT_MAX), fminf(fmaxf(z.w, -FLT_MAX), FLT_MAX));
        if (c.x != z.x || c.y != z.y || c.z != z.z || c.w != z.w) *reinterpret_cast<float4*>(zb + i) = c;   // infinities only
        *reinterpret_cast<float4*>(ob + i) = make_float4(z_range * tanhf(c.x), z_range * tanhf(c.y), z_range * tanhf(c.z),
                                                          z_range * tanhf(c.w));
    }
}

// per (image, row chunk): sum x, sum x^2, sum |dx|, sum |dy| over the chunk's rows of every channel
__global__ void zopt_reduce_kernel(const float* __restrict__ x, int C, int H, int W, int rows_per_blk, float* __restrict__ partials) {
    const int b = blockIdx.y, nblk = gridDim.x;
    const int r0 = blockIdx.x * rows_per_blk, r1 = min(r0 + rows_per_blk, H);
    float s1 = 0.f, s2 = 0.f, tx = 0.f, ty = 0.f;
    for (int c = 0; c < C; ++c) {
        const float* p = x + (static_cast<size_t>(b) * C + c) * H * W;
        for (int idx = r0 * W + threadIdx.x; idx < r1 * W; idx += blockDim.x) {
            const int j = idx % W;
            const float v = p[idx];
            s1 += v;
            s2 += v * v;
            if (j + 1 < W) tx += fabsf(v - p[idx + 1]);
            if (idx + W < H * W) ty += fabsf(v - p[idx + W]);
        }
    }
    __shared__ float red[4][kZBlock / 32];
    float vals[4] = {s1, s2, tx, ty};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float v = vals[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float v = 0.f;
        for (int w = 0; w < kZBlock / 32; ++w) v += red[threadIdx.x][w];
        partials[(static_cast<size_t>(b) * nblk + blockIdx.x) * 4 + threadIdx.x] = v;
    }
}

// stats[b] = {mean, std, a, t, loss_b, 1/nx, 1/ny, 0}: dL/dx = a * (x - mean) + t * (sign differences weighted 1/nx, 1/ny)
// mode 0: TV (w_std * (std - target)^2 + TV), 1: sign * std, 2: (std - target)^2
__global__ void zopt_final_kernel(const float* __restrict__ partials, int nblk, int B, int C, int H, int W, int mode, float sign,
                                  float w_std, const float* __restrict__ target, float* __restrict__ stats,
                                  float* __restrict__ hist, int hist_len, int* __restrict__ step) {
    __shared__ double sh[4][kZBlock];
    __shared__ float loss_sum;
    if (threadIdx.x == 0) loss_sum = 0.f;
    for (int b = 0; b < B; ++b) {
        double acc[4] = {0, 0, 0, 0};
        for (int i = threadIdx.x; i < nblk; i += blockDim.x)
            for (int k = 0; k < 4; ++k) acc[k] += partials[(static_cast<size_t>(b) * nblk + i) * 4 + k];
        for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] = acc[k];
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s)
                for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double n = static_cast<double>(C) * H * W;
            const double mean = sh[0][0] / n;
            double var = (sh[1][0] - sh[0][0] * sh[0][0] / n) / (n - 1.0);     // torch.std: unbiased
            if (var < 0) var = 0;
            const double sd = sqrt(var);
            const double nx = static_cast<double>(C) * H * (W - 1), ny = static_cast<double>(C) * (H - 1) * W;
            const double tv = (nx > 0 ? sh[2][0] / nx : 0.0) + (ny > 0 ? sh[3][0] / ny : 0.0);
            const double tgt = target != nullptr ? target[b] : 0.0;
            double loss, dstd, t;
            if (mode == 0) { loss = w_std * (sd - tgt) * (sd - tgt) + tv; dstd = 2.0 * w_std * (sd - tgt); t = 1.0; }
            else if (mode == 1) { loss = sd; dstd = 1.0; t = 0.0; }
            else { loss = (sd - tgt) * (sd - tgt); dstd = 2.0 * (sd - tgt); t = 0.0; }
            loss *= sign; dstd *= sign; t *= sign;
            float* st = stats + b * 8;
            st[0] = static_cast<float>(mean);
            st[1] = static_cast<float>(sd);
            st[2] = sd > 0 ? static_cast<float>(dstd / ((n - 1.0) * sd) / B) : 0.f;      // d std / d x_i = (x_i - mean) / ((n-1) std)
            st[3] = static_cast<float>(t / B);
            st[4] = static_cast<float>(loss);
            st[5] = nx > 0 ? static_cast<float>(1.0 / nx) : 0.f;
            st[6] = ny > 0 ? static_cast<float>(1.0 / ny) : 0.f;
            st[7] = static_cast<float>(tv);
            loss_sum += static_cast<float>(loss);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int t = *step;                          // iterations completed so far
        if (hist != nullptr && t < hist_len) hist[t] = loss_sum / B;
        *step = t + 1;                                // Adam's step count of this iteration
    }
}

__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

__global__ void zopt_grad_kernel(const float* __restrict__ x, int C, int H, int W, const float* __restrict__ stats,
                                 float* __restrict__ g) {
    const int b = blockIdx.z;
    const float* st = stats + b * 8;
    const float mean = st[0], a = st[2], t = st[3], wx = st[5] * t, wy = st[6] * t;
    const size_t plane = static_cast<size_t>(H) * W, total = plane * C;
    const float* xb = x + static_cast<size_t>(b) * total;
    float* gb = g + static_cast<size_t>(b) * total;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = idx % plane;
        const int i = static_cast<int>(r / W), j = static_cast<int>(r % W);
        const float v = xb[idx];
        float gv = a * (v - mean);
        if (t != 0.f) {
            if (j + 1 < W) gv += wx * sgn(v - xb[idx + 1]);
            if (j > 0) gv -= wx * sgn(xb[idx - 1] - v);
            if (i + 1 < H) gv += wy * sgn(v - xb[idx + W]);
            if (i > 0) gv -= wy * sgn(xb[idx - W] - v);
        }
        gb[idx] = gv;
    }
}

// g_in: gradient w.r.t. the packed model input ([B, n_img], latent part first); Z, m, v: [B, n_lat]
__global__ void zopt_adam_kernel(float* __restrict__ Z, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g_in,
                                 float z_range, int n_lat, int n_img, float lr, float beta1, float beta2, float eps,
                                 const int* __restrict__ step) {
    const int b = blockIdx.y;
    const float tstep = static_cast<float>(*step);
    const float bc1 = 1.f - powf(beta1, tstep), bc2 = 1.f - powf(beta2, tstep);
    const float step_size = lr / bc1, inv_bc2_sqrt = 1.f / sqrtf(bc2);
    const size_t zo = static_cast<size_t>(b) * n_lat;
    const float* gb = g_in + static_cast<size_t>(b) * n_img;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n_lat; i += gridDim.x * blockDim.x * 4) {
        const float4 z4 = *reinterpret_cast<const float4*>(Z + zo + i);
        const float4 g4 = *reinterpret_cast<const float4*>(gb + i);
        float4 m4 = *reinterpret_cast<const float4*>(m + zo + i), v4 = *reinterpret_cast<const float4*>(v + zo + i);
        float zz[4] = {z4.x, z4.y, z4.z, z4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
        float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float th = tanhf(zz[k]);
            const float gz = gg[k] * z_range * (1.f - th * th);
            mm[k] = mm[k] + (gz - mm[k]) * (1.f - beta1);                 // exp_avg.lerp_(grad, 1 - beta1)
            vv[k] = vv[k] * beta2 + (1.f - beta2) * gz * gz;
            const float denom = sqrtf(vv[k]) * inv_bc2_sqrt + eps;
            zz[k] -= step_size * (mm[k] / denom);
        }
        *reinterpret_cast<float4*>(Z + zo + i) = make_float4(zz[0], zz[1], zz[2], zz[3]);
        *reinterpret_cast<float4*>(m + zo + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        *reinterpret_cast<float4*>(v + zo + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
}

static int grid_for4(int n) {
    const int want = (n / 4 + kZBlock - 1) / kZBlock;
    return want < 148 * 8 ? (want > 0 ? want : 1) : 148 * 8;
}

}  // namespace esr

using namespace esr;

extern "C" int esr_zopt_tanh_pack(float* Z, float z_range, int32_t B, int32_t n_lat, int32_t n_img, float* model_input, void* stream) {
    ESR_CHECK_ARG(Z && model_input && B > 0 && n_lat > 0 && n_lat % 4 == 0 && n_img >= n_lat && n_img % 4 == 0,
                  "esr_zopt_tanh_pack: bad arguments (sizes must be multiples of 4)");
    ESR_CHECK_ARG(((reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(model_input)) & 15) == 0, "esr_zopt_tanh_pack: misaligned");
    zopt_tanh_pack_kernel<<<dim3(grid_for4(n_lat), B), kZBlock, 0, static_cast<cudaStream_t>(stream)>>>(Z, z_range, n_lat, n_img, model_input);
    return check_launch("zopt_tanh_pack_kernel");
}

constexpr int kZRowsPerBlock = 2;      // 8 rows left 128 CTAs walking 96 dependent steps at config 3 (50 us for 12.6 MB)
extern "C" int32_t esr_zopt_loss_workspace_floats(int32_t B, int32_t H) {
    const int rows_per_blk = kZRowsPerBlock;
    return B * ((H + rows_per_blk - 1) / rows_per_blk) * 4;
}

extern "C" int esr_zopt_loss(const float* x, int32_t B, int32_t C, int32_t H, int32_t W, int32_t mode, float sign, float w_std,
                             const float* target, float* workspace, float* stats, float* hist, int32_t hist_len, int32_t* step,
                             void* stream) {
    ESR_CHECK_ARG(x && workspace && stats && step && B > 0 && C > 0 && H > 0 && W > 0, "esr_zopt_loss: bad arguments");
    ESR_CHECK_ARG(mode >= 0 && mode <= 2 && (mode == 1 || target != nullptr), "esr_zopt_loss: mode 0 / 2 need a target STD per image");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int rows_per_blk = kZRowsPerBlock, nblk = (H + rows_per_blk - 1) / rows_per_blk;
    zopt_reduce_kernel<<<dim3(nblk, B), kZBlock, 0, s>>>(x, C, H, W, rows_per_blk, workspace);
    int rc = check_launch("zopt_reduce_kernel");
    if (rc) return rc;
    zopt_final_kernel<<<1, kZBlock, 0, s>>>(workspace, nblk, B, C, H, W, mode, sign, w_std, target, stats, hist, hist_len, step);
    return check_launch("zopt_final_kernel");
}

extern "C" int esr_zopt_loss_grad(const float* x, int32_t B, int32_t C, int32_t H, int32_t W, const float* stats, float* g, void* stream) {
    ESR_CHECK_ARG(x && stats && g && B > 0 && C > 0 && H > 0 && W > 0, "esr_zopt_loss_grad: bad arguments");
    const size_t total = static_cast<size_t>(C) * H * W;
    const size_t want = (total + kZBlock - 1) / kZBlock;
    const int grid = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
    zopt_grad_kernel<<<dim3(grid, 1, B), kZBlock, 0, static_cast<cudaStream_t>(stream)>>>(x, C, H, W, stats, g);
    return check_launch("zopt_grad_kernel");
}

extern "C" int esr_zopt_adam(float* Z, float* exp_avg, float* exp_avg_sq, const float* g_in, float z_range, int32_t B, int32_t n_lat,
                             int32_t n_img, float lr, float beta1, float beta2, float eps, const int32_t* step, void* stream) {
    ESR_CHECK_ARG(Z && exp_avg && exp_avg_sq && g_in && step && B > 0 && n_lat > 0 && n_lat % 4 == 0 && n_img % 4 == 0,
                  "esr_zopt_adam: bad arguments");
    ESR_CHECK_ARG(((reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq) |
                    reinterpret_cast<uintptr_t>(g_in)) & 15) == 0, "esr_zopt_adam: misaligned");
    zopt_adam_kernel<<<dim3(grid_for4(n_lat), B), kZBlock, 0, static_cast<cudaStream_t>(stream)>>>(Z, exp_avg, exp_avg_sq, g_in, z_range, n_lat,
                                                                                                    n_img, lr, beta1, beta2, eps, step);
    return check_launch("zopt_adam_kernel");
}
