// CEM x4 projection as ONE cooperative launch that reads y from HBM once (round 2, second design).
//
//   out = crop(y + Up(K * (x - Down y)))                                   (CEMnet.py:183-190)
//
// The two-launch forms in cem.cu read y twice (Down, then K + Up + add).  Here a CTA owns a SLAB of 16 full-width HR
// rows of one plane (= 4 LR rows; 128 KiB of shared memory at W = 2048) and keeps it on chip across a grid barrier:
//
//   phase 1   TMA tile loads bring the slab in (UTMALDG, 256 x 16 fp32 boxes).  Vertical Down taps from shared memory
//             into 8 register rows per thread (the slab's 16 HR rows touch LR rows I0-2 .. I0+5: its own four and two
//             partial rows on either side), horizontal Down taps through a shared tile, x subtracted on the owned rows,
//             then the HORIZONTAL half of K (27 taps) - K's two passes, the Down passes and the subtraction all commute
//             per axis, so u = K_h(x - D_h D_v y) can be accumulated per slab.  Every LR row receives exactly two
//             contributions (its owner's and one neighbour's): both are added to a zeroed buffer with red.global.add -
//             0 + a + b is independent of the order, so the result is deterministic.
//   barrier   grid-wide (cooperative groups); all slabs of a plane are resident in the same round, so the barrier is
//             the only dependency.
//   phase 2   every thread (one LR column) pulls the 34 rows of u around the slab from L2, finishes K vertically
//             (e rows I0-2 .. I0+5), applies the vertical Up taps in registers, exchanges the 16 result rows through
//             shared memory for the horizontal Up taps, adds y from the resident slab and stores out (float4, cropped).
//
// HBM traffic: y once + x + out (+ 3 % for u through L2) = the algorithmic 24.75 B per HR pixel.  Border semantics as
// in cem.cu: Down and K replicate-pad (rows / columns clamp), Up zero-pads.
#include <cuda.h>
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "cem_tabs.cuh"
#include "esr_common.cuh"
#include "ptx_sm100.cuh"

namespace cg = cooperative_groups;

namespace esr {

constexpr int kFT = 512;            // threads per CTA: one float4 column of the slab (= one LR column) each
constexpr int kFRows = 16;          // HR rows per slab
constexpr int kFBox = 256;          // TMA box width (fp32 elements; the hardware limit per dimension)
constexpr int kFVH = 8;             // halo of the V tile (HR columns either side: 2 cells)
constexpr int kFTH = 16;            // halo of the T tile (LR columns either side; 13 used, 16 keeps float4 alignment)
constexpr int kFFH = 4;             // halo of the F tile (2 used)

struct FusedArgs {
    const float* x;                 // [planes, h, w]
    float* u;                       // [planes, h, w] workspace: K_h (x - Down y), accumulated
    float* out;                     // [planes, H - 2 crop, W - 2 crop]
    int planes, H, W, h, w, crop;
    int slabs_per_plane, planes_per_round, rounds, ntiles;
    int prof;                       // ESR_CEM_PROF=1: phase time stamps into g_cemf_prof (tools/cem_fused_prof.py)
    float inv[27];
};

constexpr int kProfStamps = 8;
__device__ unsigned long long g_cemf_prof[160 * 4 * kProfStamps];
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define CEMF_STAMP(i)                                                                                         \
    do {                                                                                                      \
        if (A.prof && tid == 0 && round < 4 && blockIdx.x < 160)                                              \
            g_cemf_prof[(blockIdx.x * 4 + round) * kProfStamps + (i)] = gtime();                             \
    } while (0)

__device__ __forceinline__ void tma_load_3d_f(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__global__ void __launch_bounds__(kFT, 1)
cem_project4f_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CemTab T, const __grid_constant__ FusedArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x;
    const int W = A.W, w = A.w, H = A.H, h = A.h;
    const int Vs = W + 2 * kFVH, Ts = w + 2 * kFTH, Fs = w + 2 * kFFH;
    float* Y = reinterpret_cast<float*>(smem_raw);                         // [ntiles][16][256]
    float* S = Y + A.ntiles * kFRows * kFBox;                              // scratch: V [8][Vs] | F [16][Fs] (phase 2)
    float* Tt = S + 8 * Vs;                                                // T [8][Ts]
    uint64_t* bar = reinterpret_cast<uint64_t*>(Tt + 8 * Ts);
    const bool active = tid < w;                                           // w == W / 4: float4 column == LR column
    const int ytile = (tid >> 6) * (kFRows * kFBox) + 4 * (tid & 63);      // this thread's float4 column inside Y

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmY);
    }
    __syncthreads();

    auto slab_of = [&](int round, int& plane, int& sl) -> bool {
        const int b = blockIdx.x;
        if (b >= A.planes_per_round * A.slabs_per_plane) return false;
        plane = round * A.planes_per_round + b / A.slabs_per_plane;
        sl = b % A.slabs_per_plane;
        return plane < A.planes;
    };
    auto issue_load = [&](int round) {
        int plane, sl;
        if (!slab_of(round, plane, sl)) return;
        mbar_expect_tx(bar, static_cast<uint32_t>(A.ntiles) * kFRows * kFBox * 4);
        for (int k = 0; k < A.ntiles; ++k) tma_load_3d_f(Y + k * kFRows * kFBox, &tmY, bar, k * kFBox, sl * kFRows, plane);
    };
    if (tid == 0) issue_load(0);

    // zero the accumulation buffer while the first slab is in flight
    {
        const size_t n4 = static_cast<size_t>(A.planes) * h * w / 4;
        float4* u4 = reinterpret_cast<float4*>(A.u);
        for (size_t i = static_cast<size_t>(blockIdx.x) * kFT + tid; i < n4; i += static_cast<size_t>(gridDim.x) * kFT)
            u4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __threadfence();
    grid.sync();

    for (int round = 0; round < A.rounds; ++round) {
        int plane = 0, sl = 0;
        const bool have = slab_of(round, plane, sl);
        const int G0 = 4 * sl;                                             // first owned LR row
        CEMF_STAMP(0);
        if (have) {
            // ------------------------------------------------------------------ phase 1
            float xv[4];                                                   // x of the owned LR rows: in flight under the TMA wait
#pragma unroll
            for (int k = 0; k < 4; ++k) xv[k] = active ? __ldg(A.x + (static_cast<size_t>(plane) * h + G0 + k) * w + tid) : 0.f;
            mbar_wait(bar, round & 1);
            CEMF_STAMP(1);
            float4 acc[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) {
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 v = *reinterpret_cast<const float4*>(Y + ytile + (4 * g + q) * kFBox);
#pragma unroll
                        for (int a = g; a <= g + 4; ++a) {                 // LR row G0-2+a takes HR row 4(G0+g)+q with m = g+2-a
                            const float wv = T.down_v[q][g + 4 - a];
                            acc[a].x = fmaf(wv, v.x, acc[a].x); acc[a].y = fmaf(wv, v.y, acc[a].y);
                            acc[a].z = fmaf(wv, v.z, acc[a].z); acc[a].w = fmaf(wv, v.w, acc[a].w);
                        }
                    }
                // replicate padding: the HR rows above / below the plane equal its first / last row
                if (sl == 0) {
                    const float4 v = *reinterpret_cast<const float4*>(Y + ytile);
#pragma unroll
                    for (int g = -2; g < 0; ++g)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int a = 2; a <= g + 4; ++a) {             // only LR rows >= 0 exist (a >= 2)
                                const float wv = T.down_v[q][g + 4 - a];
                                acc[a].x = fmaf(wv, v.x, acc[a].x); acc[a].y = fmaf(wv, v.y, acc[a].y);
                                acc[a].z = fmaf(wv, v.z, acc[a].z); acc[a].w = fmaf(wv, v.w, acc[a].w);
                            }
                }
                if (sl == A.slabs_per_plane - 1) {
                    const float4 v = *reinterpret_cast<const float4*>(Y + ytile + 15 * kFBox);
#pragma unroll
                    for (int g = 4; g < 6; ++g)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int a = g; a <= 5; ++a) {                 // only LR rows <= h-1 exist (a <= 5)
                                const float wv = T.down_v[q][g + 4 - a];
                                acc[a].x = fmaf(wv, v.x, acc[a].x); acc[a].y = fmaf(wv, v.y, acc[a].y);
                                acc[a].z = fmaf(wv, v.z, acc[a].z); acc[a].w = fmaf(wv, v.w, acc[a].w);
                            }
                }
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    *reinterpret_cast<float4*>(S + a * Vs + kFVH + 4 * tid) = acc[a];
                    if (tid == 0) {
                        const float4 e = make_float4(acc[a].x, acc[a].x, acc[a].x, acc[a].x);
                        *reinterpret_cast<float4*>(S + a * Vs) = e;
                        *reinterpret_cast<float4*>(S + a * Vs + 4) = e;
                    }
                    if (tid == w - 1) {
                        const float4 e = make_float4(acc[a].w, acc[a].w, acc[a].w, acc[a].w);
                        *reinterpret_cast<float4*>(S + a * Vs + kFVH + W) = e;
                        *reinterpret_cast<float4*>(S + a * Vs + kFVH + W + 4) = e;
                    }
                }
            }
            __syncthreads();
            CEMF_STAMP(2);
            // horizontal Down taps; x subtracted on the owned rows (a = 2..5); rows outside the plane are skipped later
            if (active) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    float s = 0.f;
#pragma unroll
                    for (int k = -2; k <= 2; ++k) {
                        const float4 v = *reinterpret_cast<const float4*>(S + a * Vs + kFVH + 4 * (tid + k));
                        s = fmaf(T.down_h[0][k + 2], v.x, s); s = fmaf(T.down_h[1][k + 2], v.y, s);
                        s = fmaf(T.down_h[2][k + 2], v.z, s); s = fmaf(T.down_h[3][k + 2], v.w, s);
                    }
                    const float dv = (a >= 2 && a <= 5) ? xv[a - 2] - s : -s;
                    Tt[a * Ts + kFTH + tid] = dv;
                    if (tid == 0 || tid == w - 1) {                        // replicate halo (13 of the 16 slots are read)
                        float4* hp = reinterpret_cast<float4*>(Tt + a * Ts + (tid == 0 ? 0 : kFTH + w));
                        const float4 e = make_float4(dv, dv, dv, dv);
                        hp[0] = e; hp[1] = e; hp[2] = e; hp[3] = e;
                    }
                }
            }
            __syncthreads();
            CEMF_STAMP(3);
            // horizontal half of K: 8 consecutive outputs per task share 34 loaded values
            {
                const int groups = (w + 7) >> 3;
                for (int task = tid; task < 8 * groups; task += kFT) {
                    const int a = task / groups, j0 = (task - a * groups) * 8;
                    const int L = G0 - 2 + a;
                    if (L < 0 || L >= h) continue;
                    float tw[40];
                    const float* src = Tt + a * Ts + kFTH + j0;            // 16-byte aligned (kFTH, j0 multiples of 4)
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
                        const float4 v = *reinterpret_cast<const float4*>(src - kFTH + 4 * k);
                        tw[4 * k] = v.x; tw[4 * k + 1] = v.y; tw[4 * k + 2] = v.z; tw[4 * k + 3] = v.w;
                    }
                    // tw[i] = T[j0 - 16 + i]; output j0 + o uses T[j0 + o - 13 + c] = tw[o + 3 + c]
                    float o8[8];
#pragma unroll
                    for (int o = 0; o < 8; ++o) o8[o] = 0.f;
#pragma unroll
                    for (int c = 0; c < 27; ++c) {
                        const float kc = A.inv[c];
#pragma unroll
                        for (int o = 0; o < 8; ++o) o8[o] = fmaf(kc, tw[o + 3 + c], o8[o]);
                    }
                    float* dst = A.u + (static_cast<size_t>(plane) * h + L) * w + j0;
                    // w % 8 == 0 (host check): both float4 groups are inside the row; one L2 reduction per 16 bytes
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o8[0]), "f"(o8[1]), "f"(o8[2]), "f"(o8[3]) : "memory");
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(o8[4]), "f"(o8[5]), "f"(o8[6]), "f"(o8[7]) : "memory");
                }
            }
            __threadfence();
        }
        CEMF_STAMP(4);
        grid.sync();
        CEMF_STAMP(5);
        if (have) {
            // ------------------------------------------------------------------ phase 2
            float f16[16];
            if (active) {
                float e[8];
                {
                    float uc[34];
                    const float* ub = A.u + static_cast<size_t>(plane) * h * w + tid;
#pragma unroll
                    for (int k = 0; k < 34; ++k) {
                        int L = G0 - 15 + k;
                        L = L < 0 ? 0 : (L > h - 1 ? h - 1 : L);           // K replicate-pads
                        uc[k] = __ldcg(ub + static_cast<size_t>(L) * w);
                    }
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        float s = 0.f;
#pragma unroll
                        for (int c = 0; c < 27; ++c) s = fmaf(A.inv[c], uc[a + c], s);
                        const int L = G0 - 2 + a;
                        e[a] = (L < 0 || L >= h) ? 0.f : s;                // Up zero-pads
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float s = 0.f;
#pragma unroll
                        for (int m = -2; m <= 2; ++m) s = fmaf(T.up[q][m + 2], e[g + m + 2], s);
                        f16[4 * g + q] = s;
                    }
            }
            CEMF_STAMP(6);
            __syncthreads();                                               // the T / V tiles are dead: F aliases V
            if (active) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    S[r * Fs + kFFH + tid] = f16[r];
                    if (tid < 2) S[r * Fs + kFFH - 1 - tid] = 0.f;
                    if (tid >= w - 2) S[r * Fs + kFFH + tid + 2] = 0.f;
                }
            }
            __syncthreads();
            if (active) {
                const int col = 4 * tid - A.crop, Wo = W - 2 * A.crop, Ho = H - 2 * A.crop;
                const bool col_ok = col >= 0 && col < Wo;
                float* ob = A.out + static_cast<size_t>(plane) * Ho * Wo + col;
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int row = sl * kFRows + r - A.crop;
                    float4 o = *reinterpret_cast<const float4*>(Y + ytile + r * kFBox);
#pragma unroll
                    for (int k = -2; k <= 2; ++k) {
                        const float fv = S[r * Fs + kFFH + tid + k];
                        o.x = fmaf(T.up[0][k + 2], fv, o.x); o.y = fmaf(T.up[1][k + 2], fv, o.y);
                        o.z = fmaf(T.up[2][k + 2], fv, o.z); o.w = fmaf(T.up[3][k + 2], fv, o.w);
                    }
                    if (col_ok && row >= 0 && row < Ho) *reinterpret_cast<float4*>(ob + static_cast<size_t>(row) * Wo) = o;
                }
            }
        }
        __syncthreads();                                                   // everyone is done with Y and F
        CEMF_STAMP(7);
        if (tid == 0 && round + 1 < A.rounds) {
            fence_proxy_async();
            issue_load(round + 1);
        }
    }
}

// Measured at BASELINE config 4 (profiles/r02_notes.md): 62.6 us against 40.5 us for the two launches of cem.cu - three
// lock-step rounds of ~17 us (TMA wait 2.7, Down 3.3, K_h + red 4.0, grid.sync 1.8, K_v 2.6, Up + store 3.0) leave HBM idle
// while the SMs compute and the SMs idle while HBM moves.  It stays opt-in (ESR_CEM_FUSED=1, or esr_cem_project_fused).
bool cem_fused_enabled() {
    static const bool on = []() { const char* v = getenv("ESR_CEM_FUSED"); return v && atoi(v) == 1; }();
    return on;
}

// Returns ESR_OK when the fused launch ran, 1 when this shape is not its business (the caller falls back), < 0 on errors.
int cem_project4f(const esr_cem_filters& f, const float* y, const float* x, int planes, int H, int W, int crop, float* out,
                  float* workspace, cudaStream_t s) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    const int h = H / 4, w = W / 4, sms = num_sms_cached();
    if (f.sf != 4 || f.n_ds != 17 || f.n_inv != 27 || H % kFRows != 0 || W % 32 != 0 || w > kFT || W < 1024 ||
        crop % 4 != 0 || (static_cast<long>(planes) * h * w) % 4 != 0 || !al(y) || !al(x) || !al(out) || !al(workspace))
        return 1;
    FusedArgs A;
    A.x = x; A.u = workspace; A.out = out;
    A.planes = planes; A.H = H; A.W = W; A.h = h; A.w = w; A.crop = crop;
    A.slabs_per_plane = H / kFRows;
    if (A.slabs_per_plane > sms) return 1;                                 // a plane's slabs must be resident together
    A.planes_per_round = std::min(planes, sms / A.slabs_per_plane);
    A.rounds = ceil_div(planes, A.planes_per_round);
    A.ntiles = ceil_div(W, kFBox);
    static const int prof = []() { const char* v = getenv("ESR_CEM_PROF"); return v ? atoi(v) : 0; }();
    A.prof = prof;
    for (int i = 0; i < 27; ++i) A.inv[i] = f.inv[i];
    const int Vs = W + 2 * kFVH, Ts = w + 2 * kFTH, Fs = w + 2 * kFFH;
    const size_t scratch = std::max(static_cast<size_t>(8) * Vs, static_cast<size_t>(16) * Fs);
    if (scratch != static_cast<size_t>(8) * Vs) return 1;                  // T sits right after V
    const size_t smem = 4 * (static_cast<size_t>(A.ntiles) * kFRows * kFBox + 8 * Vs + 8 * Ts) + 64;
    if (smem > 227 * 1024) return 1;
    static const int coop = []() { int v = 0, dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, dev); return v; }();
    if (!coop) return 1;
    CUtensorMap tmY;
    int rc;
    if ((rc = make_plane_map(&tmY, y, planes, H, W, kFBox, kFRows))) return rc;
    ESR_ONCE_PER_DEVICE(ESR_CUDA(cudaFuncSetAttribute(cem_project4f_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)););
    CemTab T = make_tab(f);
    void* args[] = {&tmY, &T, &A};
    const int grid = A.planes_per_round * A.slabs_per_plane;
    const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(cem_project4f_kernel), dim3(grid), dim3(kFT), args, smem, s);
    if (e != cudaSuccess) { set_error("cem_project4f_kernel launch failed: %s", cudaGetErrorString(e)); return ESR_ERR_CUDA; }
    return check_launch("cem_project4f_kernel");
}

}  // namespace esr

extern "C" int esr_cem_project_fused(const esr_cem_filters* f, const float* y, const float* x, int32_t B, int32_t C, int32_t H,
                                     int32_t W, int32_t crop, float* out, float* workspace, void* stream) {
    ESR_CHECK_ARG(f && y && x && out && workspace && B > 0 && C > 0, "esr_cem_project_fused: bad arguments");
    ESR_CHECK_ARG(crop >= 0 && 2 * crop < H && 2 * crop < W, "esr_cem_project_fused: crop %d too large", crop);
    const int rc = esr::cem_project4f(*f, y, x, B * C, H, W, crop, out, workspace, static_cast<cudaStream_t>(stream));
    if (rc == 1) {
        esr::set_error("esr_cem_project_fused: shape %dx%d (crop %d) is outside the single-launch kernel's domain (x4 bicubic, "
                       "H %% 16 == 0, W %% 32 == 0, 1024 <= W <= 2048, H / 16 <= number of SMs, 16-byte aligned pointers)", H, W, crop);
        return ESR_ERR_UNSUPPORTED;
    }
    return rc;
}

// Phase time stamps of the last single-launch projection (ESR_CEM_PROF=1): [cta][round][8] globaltimer values.
extern "C" int esr_debug_cem_fused_prof(unsigned long long* out, int n) {
    const size_t total = sizeof(esr::g_cemf_prof) / sizeof(unsigned long long);
    ESR_CUDA(cudaMemcpyFromSymbol(out, esr::g_cemf_prof, sizeof(unsigned long long) * std::min(static_cast<size_t>(n), total)));
    return ESR_OK;
}
