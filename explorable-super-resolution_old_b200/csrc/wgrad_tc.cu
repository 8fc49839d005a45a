// Weight gradients of the trunk convolutions on the 5th-generation tensor cores (tcgen05 / TMEM), round 2.
//
//   dW[co, ci, ky, kx] = sum_p g[p, co] * X[p + (ky-1, kx-1), ci] = sum_q X[q, ci] * g[q - (ky-1, kx-1), co]
//
// is, per tap, a GEMM D[ci, co] = X^T g whose contraction index is the PIXEL.  Both tensors are NHWC in HBM (channels
// contiguous), i.e. the contracted dimension is the strided one: "MN-major" operands in tcgen05 terms.  A TMA box
// [pixels x 64 channels] written with SWIZZLE_128B (or [pixels x 32 channels] with SWIZZLE_64B) IS the canonical MN-major
// shared-memory layout (rows = K index = pixel, 8-row groups SBO apart, 64-channel column blocks LBO apart), so the NHWC
// tiles feed the MMA as they land - no transposition (csrc/wgrad.cu's ldmatrix.trans + mma.sync path is bound by the
// legacy HMMA rate of this part, ~90 TFLOP/s).
//
// One CTA per item = (conv, block of 128 input channels, block of 32 output channels): M = 128 (ci), N = 3 x 32 (kx, co), nine
// accumulators (one per tap) of 32 TMEM columns each live in tensor memory for the item's whole pixel loop and are read
// once at the end.  Per tile of 8 x 16 pixels the producer warp brings in the X tile (two 64-channel boxes, no halo) and
// the g tile with a halo of one pixel as THREE boxes shifted by one column each: the tap's row shift is then a whole-tile-
// row offset (1 KiB multiples) and its column shift selects the copy, so every descriptor start is aligned to its swizzle
// atom.  g is the small operand (64 B per pixel against 256 B of X): shifting it instead of X costs 30 KiB instead of 120.
// Pixels outside the image are zero-filled by TMA, which is the conv's zero padding in this form of the sum.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstring>

#include "esr_common.cuh"
#include "ptx_sm100.cuh"

namespace esr {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();

constexpr int kTcTH = 8, kTcTW = 16;                      // pixel tile
constexpr int kTcM = 128, kTcN = 32;                      // input / output channels per item
constexpr int kTcStages = 3;
constexpr uint32_t kTcXBox = kTcTH * kTcTW * 128;         // 16 KiB: 128 pixels x 64 channels
constexpr uint32_t kTcGBox = (kTcTH + 2) * kTcTW * 64;    // 10 KiB: 160 pixels x 32 channels
constexpr uint32_t kTcStageBytes = 2 * kTcXBox + 3 * kTcGBox;      // 62 KiB
constexpr uint32_t kTcTmemCols = 512;                     // 9 x 32 = 288 accumulator columns (power of two required)
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 1024 + 256;

// MN-major shared-memory operand descriptor (layout of cute::UMMA::SmemDescriptor): start [0,14), LBO [16,30) = byte
// distance between the swizzle-atom-wide column blocks along M / N, SBO [32,46) = byte distance between the 8-row groups
// along K, version [46,48) = 1, layout [61,64) (2 = SWIZZLE_128B, 4 = SWIZZLE_64B); all offsets in 16-byte units.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    d |= layout << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) wgrad_tc_kernel(const esr_wgrad_tc_item* __restrict__ items, const CUtensorMap* __restrict__ maps) {
    const esr_wgrad_tc_item it = items[blockIdx.x];
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcStages * kTcStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kTcStages;
    uint64_t* acc_full = bars + 2 * kTcStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const CUtensorMap* xmap = maps + it.x_map;
    const CUtensorMap* gmap = maps + it.g_map;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(xmap);
        tma_prefetch_desc(gmap);
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTcTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_x = (it.W + kTcTW - 1) / kTcTW, tiles_y = (it.H + kTcTH - 1) / kTcTH;
    const int tiles_img = tiles_x * tiles_y;
    const int t_begin = it.tile_begin, t_end = it.tile_end > 0 ? it.tile_end : it.B * tiles_img;

    if (warp == 0) {
        if (elect_one()) {                                 // ------------------------------------------------ TMA producer
            uint32_t stage = 0, phase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int n = t / tiles_img, r = t - n * tiles_img;
                const int y0 = (r / tiles_x) * kTcTH, x0 = (r % tiles_x) * kTcTW;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* st = smem + stage * kTcStageBytes;
                mbar_expect_tx(&full_bar[stage], kTcStageBytes);
                tma_load_4d(st, xmap, &full_bar[stage], it.x_c0, x0, y0, n);
                tma_load_4d(st + kTcXBox, xmap, &full_bar[stage], it.x_c0 + 64, x0, y0, n);
#pragma unroll
                for (int s = 0; s < 3; ++s)                // copy s holds g at columns x0 - 1 + s .., rows y0 - 1 ..
                    tma_load_4d(st + 2 * kTcXBox + s * kTcGBox, gmap, &full_bar[stage], it.g_c0, x0 - 1 + s, y0 - 1, n);
                if (++stage == kTcStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {                                 // ------------------------------------------------ MMA issuer
            // kind::f16, bf16 x bf16 -> fp32, A and B MN-major (bits 15, 16), M = 128, N = 96: the three column-shifted copies
            // of the gradient tile are three 32-channel swizzle atoms along N (LBO = one copy apart), so ONE instruction does
            // the three kx taps of a filter row.  (Nine N = 32 instructions per k-step re-read the 4 KiB X operand nine times:
            // 5 KiB of shared memory per 16-clock MMA, 1.94 us per tile; this form reads 7 KiB per 48-clock MMA.)
            // a_format [7,10): 1 = bf16, 0 = fp16; b_format [10,13) = bf16.  NOTE: the hardware rejects the mixed pair (fp16 x
            // bf16: illegal instruction), so the host only sends bf16 inputs (x_f16 = 0); the field is kept for an fp16 x fp16 use
            const uint32_t idesc = (1u << 4) | (it.x_f16 ? 0u : (1u << 7)) | (1u << 10) | (1u << 15) | (1u << 16) | (((3 * kTcN) >> 3) << 17) |
                                   ((kTcM >> 4) << 24);
            uint32_t stage = 0, phase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t xs = smem_u32(smem + stage * kTcStageBytes), gs = xs + 2 * kTcXBox;
#pragma unroll 1
                for (int y = 0; y < kTcTH; ++y) {          // one k-step = the 16 pixels of tile row y
                    const uint64_t adesc = umma_desc_mn(xs + y * (kTcTW * 128), kTcXBox, 1024, 2);
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        // X pixel (y, x) meets g pixel (y - ky + 1, x - kx + 1): halo-tile row y + 2 - ky of copy 2 - kx;
                        // column block n of the accumulator = copy n = tap kx = 2 - n
                        const uint64_t bdesc = umma_desc_mn(gs + (y + 2 - ky) * (kTcTW * 64), kTcGBox, 512, 4);
                        umma_bf16(tmem_base + ky * 3 * kTcN, adesc, bdesc, idesc, (t > t_begin || y > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[stage]);            // the stage is free once these MMAs retire
                if (++stage == kTcStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(acc_full);
        }
    }
    __syncwarp();
    // ---------------------------------------------------------------------------------------------------- epilogue
    if (t_begin < t_end) {
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int ci = warp * 32 + lane;                   // TMEM lane == row of D == input channel of the block
        const bool live = ci < it.n_ci;
        for (int blk = 0; blk < 9; ++blk) {                // accumulator column block (ky, copy n): tap (ky, kx = 2 - n)
            const int tap = (blk / 3) * 3 + 2 - blk % 3;
            float v[32];
            const uint32_t taddr = tmem_base + blk * kTcN + (static_cast<uint32_t>(warp * 32) << 16);
            tmem_ld_x16(taddr, *reinterpret_cast<float(*)[16]>(&v[0]));
            tmem_ld_x16(taddr + 16, *reinterpret_cast<float(*)[16]>(&v[16]));
            tmem_ld_wait();
            if (live) {
                float* dst = it.dw + (static_cast<size_t>(it.ci0) + ci) * 9 + tap;
#pragma unroll
                for (int co = 0; co < 32; ++co)
                    if (co < it.n_co) {
                        if (it.tile_end > 0) atomicAdd(dst + static_cast<size_t>(co) * it.cin_total * 9, v[co]);
                        else dst[static_cast<size_t>(co) * it.cin_total * 9] = v[co];
                    }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTcTmemCols);
    }
}

}  // namespace esr

using namespace esr;

extern "C" int32_t esr_wgrad_tc_map_bytes() { return static_cast<int32_t>(sizeof(CUtensorMap)); }

// Tensor map of an NHWC bf16 tensor [B, H, W, channels] for esr_wgrad_tc: kind 0 = conv input (box 64 channels x 16 x 8
// pixels, SWIZZLE_128B), kind 1 = gradient (box 32 channels x 16 x 10 pixels, SWIZZLE_64B).  Written to HOST memory; the
// caller uploads the table (64-byte aligned entries).
extern "C" int esr_wgrad_tc_make_map(void* map_host, const void* base, int32_t channels, int32_t B, int32_t H, int32_t W, int32_t kind) {
    ESR_CHECK_ARG(map_host && base && channels > 0 && channels % 8 == 0 && B > 0 && H > 0 && W > 0 && (kind == 0 || kind == 1),
                  "esr_wgrad_tc_make_map: bad arguments");
    ESR_CHECK_ARG((reinterpret_cast<uintptr_t>(base) & 15) == 0, "esr_wgrad_tc_make_map: base must be 16-byte aligned");
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return ESR_ERR_CUDA; }
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(channels), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(B)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(channels) * 2, static_cast<cuuint64_t>(W) * channels * 2,
                                   static_cast<cuuint64_t>(H) * W * channels * 2};
    const cuuint32_t box[4] = {kind == 0 ? 64u : 32u, static_cast<cuuint32_t>(kTcTW), static_cast<cuuint32_t>(kind == 0 ? kTcTH : kTcTH + 2), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     kind == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (wgrad map) failed with CUresult %d", static_cast<int>(r)); return ESR_ERR_CUDA; }
    memcpy(map_host, &tm, sizeof(tm));
    return ESR_OK;
}

extern "C" int esr_wgrad_tc(const esr_wgrad_tc_item* items_device, int32_t n_items, const void* maps_device, void* stream) {
    ESR_CHECK_ARG(items_device != nullptr && n_items > 0 && maps_device != nullptr, "esr_wgrad_tc: bad arguments");
    ESR_CHECK_ARG((reinterpret_cast<uintptr_t>(maps_device) & 63) == 0, "esr_wgrad_tc: the tensor-map table must be 64-byte aligned");
    // (items live in device memory: x_f16 != 0 cannot be refused here; the host mirror never sets it)
    ESR_ONCE_PER_DEVICE(ESR_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes)););
    wgrad_tc_kernel<<<n_items, 128, kTcSmemBytes, static_cast<cudaStream_t>(stream)>>>(items_device, static_cast<const CUtensorMap*>(maps_device));
    return check_launch("wgrad_tc_kernel");
}
