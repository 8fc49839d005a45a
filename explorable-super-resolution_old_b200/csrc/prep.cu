// Input preparation for G: the eval-mode replication padding of CEM_PyTorch.forward
// (CEM/CEMnet.py:170-181), the raw-view unpacking of Z and the bilinear 1/sf latent
// downscale of RRDBNet.forward (models/modules/architecture.py:152-157), and their adjoints.
#include <cuda_bf16.h>

#include "esr_common.cuh"

namespace esr {

__device__ __forceinline__ int clampi2(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

struct PrepArgs {
    const float* in;   // [B, nz*sf*sf + 3, h, w]
    int B, nz, h, w, m, sf;
    float* lr_pad;     // [B,3,h+2m,w+2m], batch stride lr_bs floats
    float* z_hr;       // [B,nz,sf(h+2m),sf(w+2m)], batch stride zhr_bs floats
    float* z_lr;       // [B,nz,h+2m,w+2m], batch stride zlr_bs floats
    long long lr_bs, zhr_bs, zlr_bs;
};

__global__ void prep_lr_kernel(const __grid_constant__ PrepArgs a) {
    const int hp = a.h + 2 * a.m, wp = a.w + 2 * a.m, cin = a.nz * a.sf * a.sf + 3;
    const size_t total = static_cast<size_t>(a.B) * 3 * hp * wp;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int X = static_cast<int>(idx % wp);
        const int Y = static_cast<int>((idx / wp) % hp);
        const int c = static_cast<int>((idx / (static_cast<size_t>(wp) * hp)) % 3);
        const int b = static_cast<int>(idx / (static_cast<size_t>(wp) * hp * 3));
        const int y = clampi2(Y - a.m, 0, a.h - 1), x = clampi2(X - a.m, 0, a.w - 1);
        a.lr_pad[static_cast<size_t>(b) * a.lr_bs + (static_cast<size_t>(c) * hp + Y) * wp + X] =
            __ldg(a.in + ((static_cast<size_t>(b) * cin + (cin - 3 + c)) * a.h + y) * a.w + x);
    }
}

// Z lives in the first nz*sf*sf channels, reinterpreted (raw .view) as [nz, sf*h, sf*w].
__device__ __forceinline__ float z_at(const PrepArgs& a, int b, int c, int Y, int X) {
    const int cin = a.nz * a.sf * a.sf + 3;
    const int Hh = a.sf * a.h, Wh = a.sf * a.w;
    const int y = clampi2(Y - a.sf * a.m, 0, Hh - 1), x = clampi2(X - a.sf * a.m, 0, Wh - 1);
    return __ldg(a.in + static_cast<size_t>(b) * cin * a.h * a.w + (static_cast<size_t>(c) * Hh + y) * Wh + x);
}

__global__ void prep_zhr_kernel(const __grid_constant__ PrepArgs a) {
    const int Hp = a.sf * (a.h + 2 * a.m), Wp = a.sf * (a.w + 2 * a.m);
    const size_t total = static_cast<size_t>(a.B) * a.nz * Hp * Wp;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int X = static_cast<int>(idx % Wp);
        const int Y = static_cast<int>((idx / Wp) % Hp);
        const int c = static_cast<int>((idx / (static_cast<size_t>(Wp) * Hp)) % a.nz);
        const int b = static_cast<int>(idx / (static_cast<size_t>(Wp) * Hp * a.nz));
        a.z_hr[static_cast<size_t>(b) * a.zhr_bs + (static_cast<size_t>(c) * Hp + Y) * Wp + X] = z_at(a, b, c, Y, X);
    }
}

// bilinear, scale 1/sf, align_corners=False: source coordinate sf*x + (sf-1)/2.
__global__ void prep_zlr_kernel(const __grid_constant__ PrepArgs a) {
    const int hp = a.h + 2 * a.m, wp = a.w + 2 * a.m;
    const size_t total = static_cast<size_t>(a.B) * a.nz * hp * wp;
    const int lo = (a.sf - 1) / 2, hi = a.sf / 2;     // equal when sf is odd
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(idx % wp);
        const int y = static_cast<int>((idx / wp) % hp);
        const int c = static_cast<int>((idx / (static_cast<size_t>(wp) * hp)) % a.nz);
        const int b = static_cast<int>(idx / (static_cast<size_t>(wp) * hp * a.nz));
        const int Y0 = a.sf * y + lo, Y1 = a.sf * y + hi, X0 = a.sf * x + lo, X1 = a.sf * x + hi;
        const float v = 0.25f * (z_at(a, b, c, Y0, X0) + z_at(a, b, c, Y0, X1) + z_at(a, b, c, Y1, X0) +
                                 z_at(a, b, c, Y1, X1));
        a.z_lr[static_cast<size_t>(b) * a.zlr_bs + (static_cast<size_t>(c) * hp + y) * wp + x] = v;
    }
}

struct PrepBwdArgs {
    const float* g_z_hr;  // [B,nz,sf*hp,sf*wp] or null
    const float* g_z_lr;  // [B,nz,hp,wp] or null
    int B, nz, h, w, m, sf;
    float* g_in;          // [B, nz*sf*sf+3, h, w]
};

__device__ __forceinline__ float gz_total(const PrepBwdArgs& a, int b, int c, int Y, int X) {
    const int hp = a.h + 2 * a.m, wp = a.w + 2 * a.m, Hp = a.sf * hp, Wp = a.sf * wp;
    float v = 0.f;
    if (a.g_z_hr) v = __ldg(a.g_z_hr + ((static_cast<size_t>(b) * a.nz + c) * Hp + Y) * Wp + X);
    if (a.g_z_lr) {
        const int lo = (a.sf - 1) / 2, hi = a.sf / 2;
        const int ry = Y % a.sf, rx = X % a.sf;
        const float wy = (ry == lo ? 0.5f : 0.f) + (ry == hi ? 0.5f : 0.f);
        const float wx = (rx == lo ? 0.5f : 0.f) + (rx == hi ? 0.5f : 0.f);
        if (wy != 0.f && wx != 0.f)
            v += wy * wx * __ldg(a.g_z_lr + ((static_cast<size_t>(b) * a.nz + c) * hp + Y / a.sf) * wp + X / a.sf);
    }
    return v;
}

// Gather form of the replicate-pad adjoint over the UNPADDED HR gradient: an interior pixel has one source (one thread
// each); a border pixel sums the margin samples that were copied from it - M+1 along an edge, (M+1)^2 in a corner - and
// is summed by a whole WARP (lanes stride over the samples in a fixed order, then a butterfly reduction), so the result is
// run-to-run reproducible (unlike the round-1 scatter with atomicAdd) without four corner threads walking 1681 samples each
// (that tail was 0.41 ms of a 10 ms Z-optimisation iteration).  Every element of the latent part of g_in is written.
__global__ void prep_bwd_kernel(const __grid_constant__ PrepBwdArgs a) {
    const int Hh = a.sf * a.h, Wh = a.sf * a.w, M = a.sf * a.m, cin = a.nz * a.sf * a.sf + 3;
    const size_t total = static_cast<size_t>(a.B) * a.nz * Hh * Wh;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(idx % Wh);
        const int y = static_cast<int>((idx / Wh) % Hh);
        if (x == 0 || y == 0 || x == Wh - 1 || y == Hh - 1) continue;          // border: the warp pass below
        const int c = static_cast<int>((idx / (static_cast<size_t>(Wh) * Hh)) % a.nz);
        const int b = static_cast<int>(idx / (static_cast<size_t>(Wh) * Hh * a.nz));
        a.g_in[static_cast<size_t>(b) * cin * a.h * a.w + (static_cast<size_t>(c) * Hh + y) * Wh + x] = gz_total(a, b, c, y + M, x + M);
    }
    const int per_plane = (Hh > 1 ? 2 * Wh : Wh) + (Wh > 1 ? 2 : 1) * (Hh > 2 ? Hh - 2 : 0);
    const size_t nborder = static_cast<size_t>(a.B) * a.nz * per_plane;
    const int lane = threadIdx.x & 31;
    const size_t warp0 = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5, nwarps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    for (size_t e = warp0; e < nborder; e += nwarps) {
        const int k = static_cast<int>(e % per_plane);
        const int c = static_cast<int>((e / per_plane) % a.nz);
        const int b = static_cast<int>(e / (static_cast<size_t>(per_plane) * a.nz));
        int y, x;
        if (k < Wh) { y = 0; x = k; }
        else if (Hh > 1 && k < 2 * Wh) { y = Hh - 1; x = k - Wh; }
        else {
            const int r = k - (Hh > 1 ? 2 * Wh : Wh);
            if (Wh > 1) { y = 1 + (r >> 1); x = (r & 1) ? Wh - 1 : 0; } else { y = 1 + r; x = 0; }
        }
        const int Y0 = y == 0 ? 0 : y + M, Y1 = y == Hh - 1 ? Hh - 1 + 2 * M : y + M;       // padded rows that clamp onto y
        const int X0 = x == 0 ? 0 : x + M, X1 = x == Wh - 1 ? Wh - 1 + 2 * M : x + M;
        const int nx = X1 - X0 + 1, n = (Y1 - Y0 + 1) * nx;
        float v = 0.f;
        for (int i = lane; i < n; i += 32) v += gz_total(a, b, c, Y0 + i / nx, X0 + i % nx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) a.g_in[static_cast<size_t>(b) * cin * a.h * a.w + (static_cast<size_t>(c) * Hh + y) * Wh + x] = v;
    }
}

__global__ void zero_kernel(float* p, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        p[i] = 0.f;
}

// Adjoint of esr_expand_rows: g_src[b,c,y,x] = sum over slots s with c_s == c of g_e[b, y-dy_s, x, s].
struct CollapseArgs {
    const float* g_e;  // f32 [B,H,W,stride] (NHWC) or [B,stride/8,H,W,8] (blocked), slots start at choff
    int B, C, H, W, nslots, stride, choff, blocked, n_acc, acc_stride;
    esr_xslot slots[64];
    float* g_src;      // NCHW f32 [B,C,H,W]
};

__global__ void collapse_rows_kernel(const __grid_constant__ CollapseArgs a) {
    const size_t total = static_cast<size_t>(a.B) * a.C * a.H * a.W;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(idx % a.W);
        const int y = static_cast<int>((idx / a.W) % a.H);
        const int c = static_cast<int>((idx / (static_cast<size_t>(a.W) * a.H)) % a.C);
        const int b = static_cast<int>(idx / (static_cast<size_t>(a.W) * a.H * a.C));
        float acc = 0.f;
        for (int s = 0; s < a.nslots; ++s) {
            if (a.slots[s].c != c) continue;
            const int yy = y - a.slots[s].dy;
            if (yy < 0 || yy >= a.H) continue;
            for (int j = 0; j < a.n_acc; ++j) {
            const int ch = a.choff + j * a.acc_stride + s;
            const size_t off = a.blocked
                ? ((static_cast<size_t>(b) * (a.stride >> 3) + (ch >> 3)) * a.H + yy) * (static_cast<size_t>(a.W) * 8) +
                      static_cast<size_t>(x) * 8 + (ch & 7)
                : ((static_cast<size_t>(b) * a.H + yy) * a.W + x) * a.stride + ch;
            acc += __ldg(a.g_e + off);
            }
        }
        a.g_src[idx] = acc;
    }
}

static int grid_for(size_t total) {
    const size_t want = (total + 255) / 256;
    return static_cast<int>(want < 148 * 16 ? (want ? want : 1) : 148 * 16);
}

}  // namespace esr

using namespace esr;

static int run_prep(PrepArgs a, cudaStream_t s) {
    const int hp = a.h + 2 * a.m, wp = a.w + 2 * a.m;
    if (a.lr_pad) {
        prep_lr_kernel<<<grid_for(static_cast<size_t>(a.B) * 3 * hp * wp), 256, 0, s>>>(a);
        if (int rc = check_launch("prep_lr_kernel")) return rc;
    }
    if (a.z_hr) {
        prep_zhr_kernel<<<grid_for(static_cast<size_t>(a.B) * a.nz * hp * wp * a.sf * a.sf), 256, 0, s>>>(a);
        if (int rc = check_launch("prep_zhr_kernel")) return rc;
    }
    if (a.z_lr) {
        prep_zlr_kernel<<<grid_for(static_cast<size_t>(a.B) * a.nz * hp * wp), 256, 0, s>>>(a);
        if (int rc = check_launch("prep_zlr_kernel")) return rc;
    }
    return ESR_OK;
}

extern "C" int esr_g_input_prep(const float* model_input, int32_t B, int32_t nz, int32_t h, int32_t w, int32_t m,
                                int32_t sf, float* lr_pad, float* fea_in, float* z_hr, float* z_lr, void* stream) {
    ESR_CHECK_ARG(model_input && B > 0 && nz >= 0 && h > 0 && w > 0 && m >= 0 && sf >= 1 && sf <= 4,
                  "esr_g_input_prep: bad arguments");
    ESR_CHECK_ARG(nz > 0 || (z_hr == nullptr && z_lr == nullptr), "esr_g_input_prep: latent outputs need nz > 0");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long hp = h + 2 * m, wp = w + 2 * m;
    PrepArgs a{model_input, B, nz, h, w, m, sf, lr_pad, z_hr, z_lr, 3 * hp * wp, nz * hp * wp * sf * sf, nz * hp * wp};
    if (int rc = run_prep(a, s)) return rc;
    if (fea_in) {  // cat([z_lr, lr_pad], 1): the 3+nz channel input of the first conv (architecture.py:160)
        PrepArgs c = a;
        c.z_hr = nullptr;
        c.lr_pad = fea_in + nz * hp * wp; c.lr_bs = (nz + 3) * hp * wp;
        c.z_lr = nz > 0 ? fea_in : nullptr; c.zlr_bs = (nz + 3) * hp * wp;
        if (int rc = run_prep(c, s)) return rc;
    }
    return ESR_OK;
}

// The padded input in the reference's own packed layout, for wrapped modules that are not ours.
extern "C" int esr_cem_pad_input(const float* model_input, int32_t B, int32_t nz, int32_t h, int32_t w, int32_t m,
                                 int32_t sf, float* packed_out, void* stream) {
    ESR_CHECK_ARG(model_input && packed_out && B > 0 && nz >= 0 && h > 0 && w > 0 && m >= 0 && sf >= 1 && sf <= 4,
                  "esr_cem_pad_input: bad arguments");
    const long long hp = h + 2 * m, wp = w + 2 * m, cin = nz * sf * sf + 3;
    PrepArgs a{model_input, B, nz, h, w, m, sf, packed_out + nz * sf * sf * hp * wp, nz > 0 ? packed_out : nullptr,
               nullptr, cin * hp * wp, cin * hp * wp, 0};
    return run_prep(a, static_cast<cudaStream_t>(stream));
}

extern "C" int esr_g_input_prep_bwd(const float* g_z_hr, const float* g_z_lr, int32_t B, int32_t nz, int32_t h,
                                    int32_t w, int32_t m, int32_t sf, float* g_model_input, void* stream) {
    ESR_CHECK_ARG(g_model_input && B > 0 && nz > 0 && h > 0 && w > 0 && m >= 0 && sf >= 1 && sf <= 4,
                  "esr_g_input_prep_bwd: bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t n_in = static_cast<size_t>(B) * (nz * sf * sf + 3) * h * w;
    zero_kernel<<<grid_for(n_in), 256, 0, s>>>(g_model_input, n_in);
    if (int rc = check_launch("zero_kernel")) return rc;
    PrepBwdArgs a{g_z_hr, g_z_lr, B, nz, h, w, m, sf, g_model_input};
    prep_bwd_kernel<<<grid_for(static_cast<size_t>(B) * nz * h * w * sf * sf), 256, 0, s>>>(a);
    return check_launch("prep_bwd_kernel");
}

extern "C" int esr_expand_rows_bwd(const float* g_e, int32_t stride, int32_t choff, int32_t blocked, int32_t n_acc,
                                   int32_t acc_stride, int32_t B, int32_t C, int32_t H, int32_t W,
                                   const esr_xslot* slots, int32_t nslots, float* g_src_nchw, void* stream) {
    ESR_CHECK_ARG(g_e && slots && g_src_nchw && nslots > 0 && nslots <= 64, "esr_expand_rows_bwd: bad arguments");
    ESR_CHECK_ARG(!blocked || stride % 8 == 0, "esr_expand_rows_bwd: blocked layout needs stride % 8 == 0");
    CollapseArgs a;
    a.g_e = g_e; a.B = B; a.C = C; a.H = H; a.W = W; a.nslots = nslots; a.stride = stride; a.choff = choff;
    a.blocked = blocked; a.n_acc = n_acc > 0 ? n_acc : 1; a.acc_stride = acc_stride;
    for (int i = 0; i < nslots; ++i) a.slots[i] = slots[i];
    a.g_src = g_src_nchw;
    collapse_rows_kernel<<<grid_for(static_cast<size_t>(B) * C * H * W), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("collapse_rows_kernel");
}

// ------------------------------------------------------------------ gradient plumbing
namespace esr {

struct CombineArgs {
    const float* src; int src_stride, src_choff, pool;
    const float* add; int add_stride, add_choff;
    int B, H, W;
    float* out_f32; int out_stride, out_choff;
    const __nv_bfloat16* mask; int mask_stride, mask_choff, mask_sub;
    float slope, scale;
    __nv_bfloat16* out_bf16; int bf16_stride, hi_choff, lo_choff;
};

// one thread per (pixel, 8-channel block) of a 64-channel tensor
__global__ void grad_combine_kernel(const __grid_constant__ CombineArgs a) {
    const size_t total = static_cast<size_t>(a.B) * a.H * a.W * 8;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        // channel block fastest: a warp covers 4 pixels x 64 channels = 512 contiguous bytes of the 16-bit NHWC mask / output
        // (with x fastest every lane touched 16 bytes of a different 128-byte pixel: half-used sectors both ways) and
        // 256-byte runs of each of the eight blocked fp32 planes
        const int blk = static_cast<int>(idx & 7);
        const int x = static_cast<int>((idx >> 3) % a.W);
        const int y = static_cast<int>((idx / (static_cast<size_t>(a.W) * 8)) % a.H);
        const int b = static_cast<int>(idx / (static_cast<size_t>(a.W) * a.H * 8));
        const int c0 = blk * 8;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        const int Hs = a.H * a.pool, Ws = a.W * a.pool;
        for (int dy = 0; dy < a.pool; ++dy)
            for (int dx = 0; dx < a.pool; ++dx) {
                const int ch = a.src_choff + c0;
                const float* p = a.src + ((static_cast<size_t>(b) * (a.src_stride >> 3) + (ch >> 3)) * Hs + (a.pool * y + dy)) *
                                             (static_cast<size_t>(Ws) * 8) + static_cast<size_t>(a.pool * x + dx) * 8;
                const float4 p0 = *reinterpret_cast<const float4*>(p), p1 = *reinterpret_cast<const float4*>(p + 4);
                v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w;
                v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
            }
        if (a.add != nullptr) {
            const int ch = a.add_choff + c0;
            const float* p = a.add + ((static_cast<size_t>(b) * (a.add_stride >> 3) + (ch >> 3)) * a.H + y) *
                                         (static_cast<size_t>(a.W) * 8) + static_cast<size_t>(x) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += p[i];
        }
        if (a.out_f32 != nullptr) {
            const int ch = a.out_choff + c0;
            float* p = a.out_f32 + ((static_cast<size_t>(b) * (a.out_stride >> 3) + (ch >> 3)) * a.H + y) *
                                       (static_cast<size_t>(a.W) * 8) + static_cast<size_t>(x) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = v[i];
        }
        if (a.out_bf16 != nullptr) {
            float w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = a.scale * v[i];
            if (a.mask != nullptr) {
                const __nv_bfloat16* m = a.mask + ((static_cast<size_t>(b) * a.H * a.mask_sub + static_cast<size_t>(y) * a.mask_sub) *
                                                       (static_cast<size_t>(a.W) * a.mask_sub) + static_cast<size_t>(x) * a.mask_sub) *
                                                      a.mask_stride + a.mask_choff + c0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {   // sign test on the raw bits: the stored activation may be bf16 or fp16
                    const uint16_t bits = reinterpret_cast<const uint16_t*>(m)[i];
                    w[i] *= ((bits & 0x8000u) == 0 && (bits & 0x7fffu) != 0) ? 1.f : a.slope;
                }
            }
            __nv_bfloat16* o = a.out_bf16 + ((static_cast<size_t>(b) * a.H + y) * a.W + x) * a.bf16_stride;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const __nv_bfloat16 h = __float2bfloat16_rn(w[i]);
                o[a.hi_choff + c0 + i] = h;
                if (a.lo_choff >= 0) o[a.lo_choff + c0 + i] = __float2bfloat16_rn(w[i] - __bfloat162float(h));
            }
        }
    }
}

}  // namespace esr

extern "C" int esr_grad_combine(const float* src, int32_t src_stride, int32_t src_choff, int32_t pool, const float* add,
                                int32_t add_stride, int32_t add_choff, int32_t B, int32_t H, int32_t W, float* out_f32,
                                int32_t out_stride, int32_t out_choff, const void* mask, int32_t mask_stride,
                                int32_t mask_choff, int32_t mask_sub, float slope, float scale, void* out_bf16,
                                int32_t bf16_stride, int32_t hi_choff, int32_t lo_choff, void* stream) {
    ESR_CHECK_ARG(src && B > 0 && H > 0 && W > 0 && (pool == 1 || pool == 2), "esr_grad_combine: bad arguments");
    ESR_CHECK_ARG(src_stride % 8 == 0 && src_choff % 8 == 0 && (!add || (add_stride % 8 == 0 && add_choff % 8 == 0)) &&
                  (!out_f32 || (out_stride % 8 == 0 && out_choff % 8 == 0)), "esr_grad_combine: f32 tensors are blocked by 8");
    CombineArgs a{src, src_stride, src_choff, pool, add, add_stride, add_choff, B, H, W, out_f32, out_stride, out_choff,
                  static_cast<const __nv_bfloat16*>(mask), mask_stride, mask_choff, mask_sub > 0 ? mask_sub : 1, slope, scale,
                  static_cast<__nv_bfloat16*>(out_bf16), bf16_stride, hi_choff, lo_choff};
    grad_combine_kernel<<<grid_for(static_cast<size_t>(B) * H * W * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return check_launch("grad_combine_kernel");
}
