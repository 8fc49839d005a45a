// conv3x3 as a tcgen05 / TMEM implicit GEMM fed by TMA (sm_100a only).
//
// One persistent CTA per SM, 6 warps:
//   warp 0      TMA producer: per K block one 4-D tensor load of the 10 x 32 pixel halo
//               tile (32 channels, SWIZZLE_64B, zero fill outside the image == the conv's
//               zero padding) and one bulk copy of that block's pre-swizzled weights.
//   warp 1      MMA issuer (one elected lane): tcgen05.mma M=128, N=3*cout_tile, K=16;
//               filter rows are row-shifted views (start address + dy*2 KiB) of the halo tile.
//               Also owns the TMEM allocation (512 columns = 2 accumulator stages x 2 bands x 128).
//   warps 2..5  epilogue: tcgen05.ld the three dx slabs, combine them with warp shuffles,
//               apply bias / LeakyReLU / residuals, store bf16 / fp32.
// smem ring: kStages x (A 20 KiB + W up to 18 KiB); mbarriers full/empty per stage and
// acc_full/acc_empty per accumulator stage, so TMA, MMA and epilogue of consecutive tiles overlap.
#include <cuda.h>

#include "conv3x3.cuh"
#include "ptx_sm100.cuh"

namespace esr {

constexpr int kStages = 5;
constexpr int kAccStages = 2;
constexpr int kAccCols = 128;                      // TMEM columns reserved per band accumulator
constexpr int kTmemCols = kAccStages * kBands * kAccCols;  // 512
constexpr int kWBytesMax = 3 * 96 * kRowBytes;     // 18432
constexpr int kStageBytes = kABytes + kWBytesMax;  // 38912 (multiple of 1024)
constexpr int kNumThreads = 192;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;

template <int CT>
__global__ void __launch_bounds__(kNumThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                  const __grid_constant__ ConvLaunch L) {
    constexpr int N = 3 * CT;
    constexpr uint32_t kIdesc = umma_idesc_bf16_m128(N);
    const esr_conv_desc& d = L.d;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* acc_full = bars + 2 * kStages;
    uint64_t* acc_empty = bars + 2 * kStages + kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap0);
        tma_prefetch_desc(&tmap1);
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int nkb = d.num_kblocks;
    const int tiles_per_img = L.tiles_x * L.tiles_y;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            pdl_wait();
            uint32_t stage = 0, phase = 0;
            for (int t = blockIdx.x; t < L.total_tiles; t += gridDim.x) {
                const int ct = t % d.cout_tiles;
                const int sp = t / d.cout_tiles;
                const int n = sp / tiles_per_img;
                const int r = sp - n * tiles_per_img;
                const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
                const int x0 = tx * kTileWOut - 1, y0 = ty * kTileH - 1;
                const uint8_t* wt = reinterpret_cast<const uint8_t*>(d.wpack) + static_cast<size_t>(ct) * d.w_tile_bytes;
                for (int kb = 0; kb < nkb; ++kb) {
                    const esr_kblock& K = d.kblocks[kb];
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * kStageBytes;
                    const uint32_t wbytes = static_cast<uint32_t>(K.n_dy) * N * kRowBytes;
                    mbar_expect_tx(&full_bar[stage], kABytes + wbytes);
                    tma_load_4d(sa, K.src == 0 ? &tmap0 : &tmap1, &full_bar[stage], K.chan, x0, y0, n);
                    bulk_load_1d(sa + kABytes, wt + K.w_off, wbytes, &full_bar[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer
        if (elect_one()) {
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            for (int t = blockIdx.x; t < L.total_tiles; t += gridDim.x) {
                mbar_wait(&acc_empty[as], aphase ^ 1);
                tc_fence_after();
                uint32_t first[kBands];
#pragma unroll
                for (int b = 0; b < kBands; ++b) first[b] = 1;
                for (int kb = 0; kb < nkb; ++kb) {
                    const esr_kblock& K = d.kblocks[kb];
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint32_t sw = sa + kABytes;
#pragma unroll
                    for (int b = 0; b < kBands; ++b) {
                        const uint32_t acc = tmem_base + (as * kBands + b) * kAccCols;
                        int wi = 0;
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            if (!((K.dy_mask >> dy) & 1)) continue;
                            const uint32_t a_row = sa + (b * kBandRows + dy) * (kTileW * kRowBytes);
                            const uint32_t w_row = sw + wi * (N * kRowBytes);
                            ++wi;
#pragma unroll
                            for (int s = 0; s < 2; ++s) {
                                if (!((K.slice_mask >> s) & 1)) continue;
                                umma_bf16(acc, umma_smem_desc(a_row + s * 32, kRowBytes),
                                          umma_smem_desc(w_row + s * 32, kRowBytes), kIdesc, first[b] ? 0u : 1u);
                                first[b] = 0;
                            }
                        }
                    }
                    umma_commit(&empty_bar[stage]);   // frees the smem stage when these MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[as]);           // accumulators of this tile are complete
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue
        const int wq = warp & 3;                      // TMEM lane quadrant == row inside the band
        uint32_t as = 0, aphase = 0;
        for (int t = blockIdx.x; t < L.total_tiles; t += gridDim.x) {
            const int ct = t % d.cout_tiles;
            const int sp = t / d.cout_tiles;
            const int n = sp / tiles_per_img;
            const int r = sp - n * tiles_per_img;
            const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
            const int x = tx * kTileWOut - 1 + lane;
            const bool col_ok = lane >= 1 && lane <= kTileWOut && x < d.W;
            mbar_wait(&acc_full[as], aphase);
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < kBands; ++b) {
                const int y = ty * kTileH + b * kBandRows + wq;
                const uint32_t taddr = tmem_base + (as * kBands + b) * kAccCols + (static_cast<uint32_t>(wq * 32) << 16);
#pragma unroll
                for (int h = 0; h < CT / 16; ++h) {
                    float vl[16], vc[16], vr[16];
                    tmem_ld_x16(taddr + 0 * CT + h * 16, vl);
                    tmem_ld_x16(taddr + 1 * CT + h * 16, vc);
                    tmem_ld_x16(taddr + 2 * CT + h * 16, vr);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float fl = __shfl_up_sync(0xffffffffu, vl[i], 1);
                        const float fr = __shfl_down_sync(0xffffffffu, vr[i], 1);
                        vc[i] += fl + fr;
                    }
                    if (col_ok && y < d.H) conv_epilogue16(d, n, y, x, ct * CT + h * 16, vc);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
            if (++as == kAccStages) { as = 0; aphase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// NHWC bf16 [B,H,W,C] seen as a 4-D tensor (C, W, H, B); box = 32 ch x 32 px x 10 rows.
int make_act_tensor_map(CUtensorMap* tm, const esr_tensor_nhwc& t, int B, int H, int W) {
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return ESR_ERR_CUDA; }
    ESR_CHECK_ARG(t.ptr != nullptr && t.channels >= kKB && t.channels % 8 == 0,
                  "conv source must be NHWC bf16 with >=32 channels, multiple of 8 (got %d)", t.channels);
    ESR_CHECK_ARG((reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0, "conv source pointer must be 16-byte aligned");
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(t.channels), static_cast<cuuint64_t>(W),
                                static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(B)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(t.channels) * 2, static_cast<cuuint64_t>(W) * t.channels * 2,
                                   static_cast<cuuint64_t>(H) * W * t.channels * 2};
    const cuuint32_t box[4] = {kKB, kTileW, kHaloRows, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(t.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r)); return ESR_ERR_CUDA; }
    return ESR_OK;
}

int validate_conv_desc(const esr_conv_desc& d) {
    ESR_CHECK_ARG(d.B > 0 && d.H > 0 && d.W > 0, "bad conv geometry %dx%dx%d", d.B, d.H, d.W);
    ESR_CHECK_ARG(d.cout_tile == 32 || d.cout_tile == 16, "cout_tile must be 16 or 32");
    ESR_CHECK_ARG(d.cout_tiles > 0, "cout_tiles must be positive");
    ESR_CHECK_ARG(d.num_kblocks > 0 && d.num_kblocks <= ESR_MAX_KBLOCKS, "num_kblocks out of range");
    ESR_CHECK_ARG(d.wpack != nullptr && d.bias != nullptr, "wpack/bias missing");
    ESR_CHECK_ARG((reinterpret_cast<uintptr_t>(d.wpack) & 15) == 0 && d.w_tile_bytes % 16 == 0, "wpack misaligned");
    for (int i = 0; i < d.num_kblocks; ++i) {
        const esr_kblock& k = d.kblocks[i];
        ESR_CHECK_ARG(k.src == 0 || k.src == 1, "kblock %d: bad src", i);
        ESR_CHECK_ARG(d.src[k.src].ptr != nullptr, "kblock %d: src %d is null", i, k.src);
        ESR_CHECK_ARG(k.chan >= 0 && k.chan % 8 == 0 && k.chan + kKB <= d.src[k.src].channels,
                      "kblock %d: channel window [%d,%d) outside source (%d ch)", i, k.chan, k.chan + kKB,
                      d.src[k.src].channels);
        ESR_CHECK_ARG((k.dy_mask & 7) != 0 && (k.slice_mask & 3) != 0 && k.n_dy == __builtin_popcount(k.dy_mask & 7),
                      "kblock %d: bad masks", i);
        ESR_CHECK_ARG(k.w_off % 512 == 0 && k.w_off + k.n_dy * 3u * d.cout_tile * kRowBytes <= d.w_tile_bytes,
                      "kblock %d: weight slab outside tile image", i);
    }
    ESR_CHECK_ARG(d.up == 1 || d.up == 2, "up must be 1 or 2");
    if (d.out_bf16) ESR_CHECK_ARG(d.out_bf16_stride % 8 == 0 && d.out_bf16_choff % 8 == 0 &&
                                  (d.out_bf16_lo_choff < 0 || d.out_bf16_lo_choff % 8 == 0), "bf16 output misaligned");
    if (d.out_f32) ESR_CHECK_ARG(d.out_f32_stride % 4 == 0 && d.out_f32_choff % 4 == 0, "f32 output misaligned");
    if (d.flags & ESR_EPI_ACCUM) ESR_CHECK_ARG(d.out_f32 != nullptr, "ACCUM needs out_f32");
    if (d.flags & ESR_EPI_RES1) ESR_CHECK_ARG(d.res1 && d.res1_stride % 4 == 0 && d.res1_choff % 4 == 0, "res1 misaligned");
    if (d.flags & ESR_EPI_RES2) ESR_CHECK_ARG(d.res2 && d.res2_stride % 4 == 0 && d.res2_choff % 4 == 0, "res2 misaligned");
    if (d.flags & ESR_EPI_MASK) ESR_CHECK_ARG(d.mask && d.out_bf16 && d.up == 1 && d.mask_stride % 8 == 0 && d.mask_choff % 8 == 0, "mask misaligned");
    if (d.out_nchw) ESR_CHECK_ARG(d.cout_real > 0 && d.cout_real <= d.cout_tile * d.cout_tiles, "bad cout_real");
    return ESR_OK;
}

void fill_launch(ConvLaunch* L, const esr_conv_desc& d) {
    L->d = d;
    L->tiles_x = ceil_div(d.W, kTileWOut);
    L->tiles_y = ceil_div(d.H, kTileH);
    L->total_tiles = d.B * L->tiles_x * L->tiles_y * d.cout_tiles;
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

int launch_conv_tc(const CUtensorMap& tm0, const CUtensorMap& tm1, const ConvLaunch& L, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        ESR_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        ESR_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set = true;
    }
    const int grid = L.total_tiles < num_sms() ? L.total_tiles : num_sms();
    if (L.d.cout_tile == 32)
        conv3x3_tc_kernel<32><<<grid, kNumThreads, kSmemBytes, stream>>>(tm0, tm1, L);
    else
        conv3x3_tc_kernel<16><<<grid, kNumThreads, kSmemBytes, stream>>>(tm0, tm1, L);
    return check_launch("conv3x3_tc_kernel");
}

int build_conv_launch(const esr_conv_desc& d, CUtensorMap* tm0, CUtensorMap* tm1, ConvLaunch* L) {
    int rc = validate_conv_desc(d);
    if (rc != ESR_OK) return rc;
    rc = make_act_tensor_map(tm0, d.src[0], d.B, d.H, d.W);
    if (rc != ESR_OK) return rc;
    if (d.src[1].ptr != nullptr) {
        rc = make_act_tensor_map(tm1, d.src[1], d.B, d.H, d.W);
        if (rc != ESR_OK) return rc;
    } else {
        *tm1 = *tm0;
    }
    fill_launch(L, d);
    return ESR_OK;
}

}  // namespace esr

extern "C" int esr_conv3x3_tc(const esr_conv_desc* d, void* stream) {
    if (d == nullptr) { esr::set_error("null conv desc"); return ESR_ERR_INVALID; }
    CUtensorMap tm0, tm1;
    esr::ConvLaunch L;
    int rc = esr::build_conv_launch(*d, &tm0, &tm1, &L);
    if (rc != ESR_OK) return rc;
    return esr::launch_conv_tc(tm0, tm1, L, static_cast<cudaStream_t>(stream));
}
