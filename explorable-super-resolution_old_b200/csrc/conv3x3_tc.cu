// conv3x3 as a tcgen05 / TMEM implicit GEMM fed by TMA (sm_100a only).
//
// One persistent CTA per SM, 10 warps:
//   warp 0      TMA producer: per K block one 4-D tensor load of the 10 x 32 pixel halo
//               tile (32 channels, SWIZZLE_64B, zero fill outside the image == the conv's
//               zero padding) and one bulk copy of that block's pre-swizzled weights.
//   warp 1      MMA issuer (one elected lane): tcgen05.mma M=128, N=3*cout_tile, K=16;
//               filter rows are row-shifted views (start address + dy*2 KiB) of the halo tile.
//               Also owns the TMEM allocation (512 columns = 2 accumulator stages x 2 bands x 128).
//   warps 2..9  epilogue, one warp per (band, TMEM lane quadrant): tcgen05.ld the three dx slabs,
//               combine them with warp shuffles, apply bias / LeakyReLU / residuals, store bf16 / fp32.
// Every CTA owns one cout tile and keeps that tile's whole packed weight image (<= 124 KiB) resident
// in shared memory; only the 20 KiB activation halo tiles stream through a 4-6 deep TMA ring.
// mbarriers full/empty per ring stage and acc_full/acc_empty per accumulator stage let TMA, MMA and
// the epilogue of consecutive tiles overlap.
#include <cuda.h>

#include <cstdlib>

#include "conv3x3.cuh"
#include "ptx_sm100.cuh"

// Per-role cycle counters (tools/prof.py); compiled in only with -DESR_PROFILE_ROLES.
#ifdef ESR_PROFILE_ROLES
#define ESR_PROF(...) __VA_ARGS__
#else
#define ESR_PROF(...)
#endif

namespace esr {

constexpr int kMaxStages = 6;                      // A-tile ring depth (fewer when the weights are large)
constexpr int kAccStages = 2;
constexpr int kAccCols = 128;                      // TMEM columns reserved per band accumulator
constexpr int kTmemCols = kAccStages * kBands * kAccCols;  // 512
constexpr int kEpiWarps = 4 * kBands;              // one warp per (band, TMEM lane quadrant)
constexpr int kNumThreads = 64 + 32 * kEpiWarps;   // 320: leaves ~200 registers per epilogue thread
constexpr int kMaxBias = 256;                      // floats of bias staged in shared memory
constexpr int kCtrlBytes = 256 + kMaxBias * 4;     // barriers + tmem slot, bias
constexpr int kSmemMax = 227 * 1024;
// High word of every shared-memory matrix descriptor used here: SBO = 512 B (8 rows x 64 B),
// descriptor version 1, SWIZZLE_64B.  The low word is (smem address >> 4).
constexpr uint32_t kDescHi = ((8u * kRowBytes) >> 4) | (1u << 14) | (4u << 29);

__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Shared memory: [ weights of this CTA's cout tile (resident for the whole kernel) | A-tile ring |
// barriers | bias ].  Every CTA owns one cout tile (blockIdx.x % cout_tiles) and walks spatial tiles.
template <int CT, int MODE>
__global__ void __launch_bounds__(kNumThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                  const __grid_constant__ ConvLaunch L) {
    constexpr int N = 3 * CT;
    // operand format bits: a_format [7,10) and b_format [10,13) are 1 for bf16, 0 for fp16
    const uint32_t kIdesc = umma_idesc_bf16_m128(N) & ((L.d.flags & ESR_CONV_F16) ? ~((1u << 7) | (1u << 10)) : ~0u);
    constexpr uint32_t kARow16 = (kTileW * kRowBytes) >> 4;   // one halo-tile row, in 16-byte units
    constexpr uint32_t kWSlab16 = (N * kRowBytes) >> 4;       // one [N x 32ch] weight slab
    const esr_conv_desc& d = L.d;
    const int nstages = L.nstages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_w = smem;
    uint8_t* s_a = smem + L.w_smem_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_a + nstages * kABytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kMaxStages;
    uint64_t* acc_full = bars + 2 * kMaxStages;
    uint64_t* acc_empty = acc_full + kAccStages;
    uint64_t* w_full = acc_empty + kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    ESR_PROF(if (L.prof && threadIdx.x == 0) L.prof[blockIdx.x * 16 + 12] = gtime_ns();)
    const int ct = blockIdx.x % d.cout_tiles;
    const int sp0 = blockIdx.x / d.cout_tiles, sp_step = gridDim.x / d.cout_tiles;
    for (int i = threadIdx.x; i < CT; i += kNumThreads) s_bias[i] = d.bias[ct * CT + i];

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap0);
        tma_prefetch_desc(&tmap1);
        for (int s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], kEpiWarps); }
        mbar_init(w_full, 1);
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
    ESR_PROF(if (L.prof && threadIdx.x == 0) L.prof[blockIdx.x * 16 + 13] = gtime_ns();)
    pdl_launch_dependents();   // the next layer may start its prologue (weights, TMEM, barriers) on idle SMs

    const int nkb = d.num_kblocks;
    const int tiles_per_img = L.tiles_x * L.tiles_y;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            // weights first: they do not depend on the previous layer, so under programmatic dependent
            // launch this copy overlaps the tail of the preceding kernel
            const uint8_t* wt = reinterpret_cast<const uint8_t*>(d.wpack) + static_cast<size_t>(ct) * d.w_tile_bytes;
            mbar_expect_tx(w_full, d.w_tile_bytes);
            for (uint32_t off = 0; off < d.w_tile_bytes; off += 16384) {
                const uint32_t n = d.w_tile_bytes - off < 16384 ? d.w_tile_bytes - off : 16384;
                bulk_load_1d(s_w + off, wt + off, n, w_full);
            }
            pdl_wait();
            uint32_t stage = 0, phase = 0;
            ESR_PROF(long long p_wait = 0, p_t0 = clock64(), p_n = 0;)
            for (int spi = sp0; spi < L.spatial_tiles; spi += sp_step) {
                const int sp = L.reverse ? L.spatial_tiles - 1 - spi : spi;
                const int n = sp / tiles_per_img;
                const int r = sp - n * tiles_per_img;
                const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
                const int x0 = tx * kTileWOut - 1, y0 = ty * kTileH - 1;
                for (int kb = 0; kb < nkb; ++kb) {
                    const esr_kblock& K = d.kblocks[kb];
                    ESR_PROF(const long long w0c = clock64();)
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    ESR_PROF(p_wait += clock64() - w0c; ++p_n;)
                    if ((L.debug & 1) && (kb & 1)) {
                        mbar_arrive(&full_bar[stage]);            // timing experiment: no load
                    } else {
                        mbar_expect_tx(&full_bar[stage], kABytes);
                        tma_load_4d(s_a + stage * kABytes, K.src == 0 ? &tmap0 : &tmap1, &full_bar[stage], K.chan, x0, y0, n);
                    }
                    if (++stage == static_cast<uint32_t>(nstages)) { stage = 0; phase ^= 1; }
                }
            }
            ESR_PROF(if (L.prof) {
                unsigned long long* o = L.prof + blockIdx.x * 16;
                o[0] = clock64() - p_t0; o[1] = p_wait; o[2] = p_n;
            })
        }
    } else if (warp == 1) {
        // -------------------------------------------------------------- MMA issuer
        if (elect_one()) {
            const uint32_t w_lo = smem_u32(s_w) >> 4, a_lo = smem_u32(s_a) >> 4;
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            ESR_PROF(long long m_t0 = clock64(), m_wacc = 0, m_wfull = 0;)
            mbar_wait(w_full, 0);
            ESR_PROF(const long long m_tw = clock64() - m_t0;)
            for (int sp = sp0; sp < L.spatial_tiles; sp += sp_step) {
                ESR_PROF(long long c0 = clock64();)
                mbar_wait(&acc_empty[as], aphase ^ 1);
                ESR_PROF(m_wacc += clock64() - c0;)
                tc_fence_after();
                const uint32_t acc0 = tmem_base + as * (kBands * kAccCols);
                uint32_t nonfirst = 0;                 // 0 until the accumulators hold their first product
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint32_t masks = *reinterpret_cast<const uint32_t*>(&d.kblocks[kb].dy_mask);
                    const uint32_t dy_mask = masks & 0xff, slice_mask = (masks >> 8) & 0xff;
                    const uint32_t w0 = w_lo + (d.kblocks[kb].w_off >> 4);
                    ESR_PROF(c0 = clock64();)
                    mbar_wait(&full_bar[stage], phase);
                    ESR_PROF(m_wfull += clock64() - c0;)
                    tc_fence_after();
                    const uint32_t a0 = a_lo + stage * (kABytes >> 4);
                    if ((L.debug & 2) && (kb & 1)) {
                        // timing experiment: no MMAs for this block
                    } else if (dy_mask == 7u && slice_mask == 3u) {
                        // dense block: 2 bands x 3 filter rows x 2 K slices, all offsets immediate
                        // consecutive MMAs alternate between the two band accumulators so that the
                        // accumulate dependency of one never stalls the tensor pipe
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                            for (int s = 0; s < 2; ++s) {
#pragma unroll
                                for (int b = 0; b < kBands; ++b) {
                                    umma_issue(acc0 + b * kAccCols, a0 + (b * kBandRows + dy) * kARow16 + s * 2,
                                               w0 + dy * kWSlab16 + s * 2, kIdesc, (dy | s) ? 1u : nonfirst);
                                }
                            }
                        }
                    } else {
                        for (int b = 0; b < kBands; ++b) {
                            uint32_t acc_flag = nonfirst, wi = 0;
                            for (int dy = 0; dy < 3; ++dy) {
                                if (!((dy_mask >> dy) & 1u)) continue;
                                for (int s = 0; s < 2; ++s) {
                                    if (!((slice_mask >> s) & 1u)) continue;
                                    umma_issue(acc0 + b * kAccCols, a0 + (b * kBandRows + dy) * kARow16 + s * 2,
                                               w0 + wi * kWSlab16 + s * 2, kIdesc, acc_flag);
                                    acc_flag = 1;
                                }
                                ++wi;
                            }
                        }
                    }
                    nonfirst = 1;
                    umma_commit(&empty_bar[stage]);   // frees the A stage when these MMAs retire
                    if (++stage == static_cast<uint32_t>(nstages)) { stage = 0; phase ^= 1; }
                }
                umma_commit(&acc_full[as]);           // accumulators of this tile are complete
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
            }
            ESR_PROF(if (L.prof) {
                unsigned long long* o = L.prof + blockIdx.x * 16;
                o[3] = clock64() - m_t0; o[4] = m_wacc; o[5] = m_wfull; o[6] = m_tw; o[14] = gtime_ns();
            })
        }
    } else {
        // ---------------------------------------------------------------- epilogue
        const int wq = warp & 3;                      // TMEM lane quadrant == row inside the band
        const int b = (warp - 2) >> 2;                // band handled by this warp
        constexpr int NH = CT / 16;                   // 16-channel halves of the cout tile
        pdl_wait();                                   // residual / accumulate inputs come from earlier kernels
        uint32_t as = 0, aphase = 0;
        ESR_PROF(long long e_t0 = clock64(), e_wait = 0, e_ld = 0, e_shfl = 0, e_n = 0;)
        for (int spi = sp0; spi < L.spatial_tiles; spi += sp_step) {
            const int sp = L.reverse ? L.spatial_tiles - 1 - spi : spi;
            const int n = sp / tiles_per_img;
            const int r = sp - n * tiles_per_img;
            const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
            const int x = tx * kTileWOut - 1 + lane;
            const int y = ty * kTileH + b * kBandRows + wq;
            const bool ok = lane >= 1 && lane <= kTileWOut && x < d.W && y < d.H;
            EpiOperands ops[NH];
            if (ok) {                                 // issued before the wait: latency hides behind the MMAs
#pragma unroll
                for (int h = 0; h < NH; ++h) conv_epilogue_prefetch<MODE>(d, ct, n, y, x, tile_channel(d, ct, h * 16), ops[h]);
            }
            ESR_PROF(long long c0 = clock64();)
            mbar_wait(&acc_full[as], aphase);
            ESR_PROF(long long c1 = clock64(); e_wait += c1 - c0; ++e_n;)
            tc_fence_after();
            const uint32_t taddr = tmem_base + (as * kBands + b) * kAccCols + (static_cast<uint32_t>(wq * 32) << 16);
            float vc[CT];
#pragma unroll
            for (int h = 0; h < NH; ++h) {            // 16 channels at a time keeps the register footprint bounded
                float vl[16], vr[16];
                tmem_ld_x16(taddr + 0 * CT + h * 16, vl);
                tmem_ld_x16(taddr + 1 * CT + h * 16, *reinterpret_cast<float(*)[16]>(&vc[h * 16]));
                tmem_ld_x16(taddr + 2 * CT + h * 16, vr);
                tmem_ld_wait();
                if (h == NH - 1) {                    // all TMEM reads of this warp are done: release early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[as]);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float fl = __shfl_up_sync(0xffffffffu, vl[i], 1);
                    const float fr = __shfl_down_sync(0xffffffffu, vr[i], 1);
                    vc[h * 16 + i] += fl + fr;
                }
            }
            ESR_PROF(c0 = clock64(); e_ld += c0 - c1;)
            if (ok) {
#pragma unroll
                for (int h = 0; h < NH; ++h)
                    conv_epilogue16<MODE>(d, s_bias + h * 16, ct, n, y, x, tile_channel(d, ct, h * 16),
                                          *reinterpret_cast<float(*)[16]>(&vc[h * 16]), ops[h]);
            }
            if (++as == kAccStages) { as = 0; aphase ^= 1; }
        }
        ESR_PROF(if (L.prof && warp == 2 && lane == 1) {
            unsigned long long* o = L.prof + blockIdx.x * 16;
            o[7] = clock64() - e_t0; o[8] = e_wait; o[9] = e_ld; o[10] = e_shfl; o[11] = e_n; o[15] = gtime_ns();
        })
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

EncodeTiledFn get_encode_fn() {          // also used by csrc/cem.cu (fp32 plane maps of the CEM streaming kernels)
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
// lean != 0: 16-channel / 32-byte rows, SWIZZLE_32B (the lean latent blocks, esr_kblock.half)
int make_act_tensor_map(CUtensorMap* tm, const esr_tensor_nhwc& t, int B, int H, int W, int box_rows, int lean) {
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return ESR_ERR_CUDA; }
    ESR_CHECK_ARG(t.ptr != nullptr && t.channels >= kKB && t.channels % 8 == 0,
                  "conv source must be NHWC bf16 with >=32 channels, multiple of 8 (got %d)", t.channels);
    ESR_CHECK_ARG((reinterpret_cast<uintptr_t>(t.ptr) & 15) == 0, "conv source pointer must be 16-byte aligned");
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(t.channels), static_cast<cuuint64_t>(W),
                                static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(B)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(t.channels) * 2, static_cast<cuuint64_t>(W) * t.channels * 2,
                                   static_cast<cuuint64_t>(H) * W * t.channels * 2};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(lean ? 16 : kKB), kTileW, static_cast<cuuint32_t>(box_rows), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(t.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, lean ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r)); return ESR_ERR_CUDA; }
    return ESR_OK;
}

int validate_conv_desc(const esr_conv_desc& d) {
    ESR_CHECK_ARG(d.B > 0 && d.H > 0 && d.W > 0, "bad conv geometry %dx%dx%d", d.B, d.H, d.W);
    ESR_CHECK_ARG(d.pair ? (d.cout_tile == 32 || d.cout_tile == 64) : (d.cout_tile == 32 || d.cout_tile == 16),
                  "cout_tile must be 16 or 32 (32 or 64 in pair mode)");
    ESR_CHECK_ARG(!d.pair || d.w_tile_bytes % 32 == 0, "pair mode: w_tile_bytes must split into two 16-byte aligned halves");
    ESR_CHECK_ARG(d.cout_tiles > 0 && d.cout_tiles <= ESR_MAX_COUT_TILES, "cout_tiles out of range");
    for (int t = 0; t < d.cout_tiles; ++t)
        ESR_CHECK_ARG(d.tile_choff[t] < 0 || d.tile_choff[t] % 16 == 0, "tile_choff[%d] must be a multiple of 16", t);
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
        ESR_CHECK_ARG(k.w_off % 512 == 0 &&
                      k.w_off + k.n_dy * 3u * d.cout_tile * (k.half ? 32u : static_cast<uint32_t>(kRowBytes)) / (d.pair ? 2u : 1u) <=
                          d.w_tile_bytes / (d.pair ? 2u : 1u),
                      "kblock %d: weight slab outside tile image", i);
        ESR_CHECK_ARG(!k.half || (d.pair && k.src == 1 && (k.dy_mask & 7) == 2 && (k.slice_mask == 1 || k.slice_mask == 2)),
                      "kblock %d: a lean (half) block needs pair mode, source 1, the centre tap and one slice", i);
    }
    ESR_CHECK_ARG(d.up == 1 || d.up == 2, "up must be 1 or 2");
    if (d.out_bf16) ESR_CHECK_ARG(d.out_bf16_stride % 8 == 0 && d.out_bf16_choff % 8 == 0 &&
                                  (d.out_bf16_lo_choff < 0 || d.out_bf16_lo_choff % 8 == 0), "bf16 output misaligned");
    if (d.out_f32) ESR_CHECK_ARG(d.out_f32_stride % 4 == 0 && d.out_f32_choff % 4 == 0, "f32 output misaligned");
    if (d.flags & ESR_EPI_ACCUM) ESR_CHECK_ARG(d.out_f32 != nullptr, "ACCUM needs out_f32");
    if (d.flags & ESR_EPI_F32_BLOCKED) {
        const bool r1f32 = (d.flags & ESR_EPI_RES1) && !(d.flags & ESR_EPI_RES1_HILO);
        (void)r1f32;
        ESR_CHECK_ARG((d.out_f32 == nullptr || (d.out_f32_stride % 8 == 0 && d.out_f32_choff % 8 == 0 && (reinterpret_cast<uintptr_t>(d.out_f32) & 31) == 0)) &&
                      (d.res1 == nullptr || (d.res1_stride % 8 == 0 && d.res1_choff % 8 == 0 && (reinterpret_cast<uintptr_t>(d.res1) & 31) == 0)) &&
                      (d.res2 == nullptr || (d.res2_stride % 8 == 0 && d.res2_choff % 8 == 0 && (reinterpret_cast<uintptr_t>(d.res2) & 31) == 0)),
                      "blocked f32 tensors need 8-channel multiples and 32-byte alignment");
    }
    const auto pair16 = [](const void* p, int stride, int choff) {
        return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 31) == 0 && stride % 16 == 0 && choff % 16 == 0;
    };
    if ((d.flags & ESR_EPI_RES1) && (d.flags & ESR_EPI_RES1_HILO))
        ESR_CHECK_ARG(pair16(d.res1_hi, d.res1_hi_stride, d.res1_hi_choff) && pair16(d.res1_lo, d.res1_lo_stride, d.res1_lo_choff),
                      "res1_hi / res1_lo must be bf16 NHWC, 32-byte aligned, 16-channel multiples");
    else if (d.flags & ESR_EPI_RES1) ESR_CHECK_ARG(d.res1 && d.res1_stride % 4 == 0 && d.res1_choff % 4 == 0, "res1 misaligned");
    if (d.out_lo) ESR_CHECK_ARG(pair16(d.out_lo, d.out_lo_stride, d.out_lo_choff), "out_lo misaligned");
    if (d.flags & ESR_EPI_RES2) ESR_CHECK_ARG(d.res2 && d.res2_stride % 4 == 0 && d.res2_choff % 4 == 0, "res2 misaligned");
    if (d.flags & ESR_EPI_MASK) ESR_CHECK_ARG(d.mask && d.out_bf16 && d.up == 1 && d.mask_stride % 8 == 0 && d.mask_choff % 8 == 0, "mask misaligned");
    if (d.out_nchw) ESR_CHECK_ARG(d.cout_real > 0 && d.cout_real <= d.cout_tile * d.cout_tiles, "bad cout_real");
    return ESR_OK;
}

static unsigned long long* g_prof_buf = nullptr;
unsigned long long* rdb_prof_buffer() { return g_prof_buf; }
static int g_use_pdl = []() { const char* v = getenv("ESR_NO_PDL"); return (v && atoi(v)) ? 0 : 1; }();

static bool f32_wide_ok(const float* p, int stride, int choff) {
    return p == nullptr || ((reinterpret_cast<uintptr_t>(p) & 31) == 0 && stride % 8 == 0 && choff % 8 == 0);
}

void fill_launch(ConvLaunch* L, const esr_conv_desc& d) {
    L->d = d;
    L->d.flags &= ~static_cast<uint32_t>(ESR_EPI_WIDE_OK);
    if (f32_wide_ok(d.out_f32, d.out_f32_stride, d.out_f32_choff) && f32_wide_ok(d.res1, d.res1_stride, d.res1_choff) &&
        f32_wide_ok(d.res2, d.res2_stride, d.res2_choff))
        L->d.flags |= ESR_EPI_WIDE_OK;
    L->tiles_x = ceil_div(d.W, kTileWOut);
    L->tiles_y = ceil_div(d.H, kTileH);
    L->spatial_tiles = d.B * L->tiles_x * L->tiles_y;
    L->total_tiles = L->spatial_tiles * d.cout_tiles;
    L->w_smem_bytes = (d.w_tile_bytes + 1023u) & ~1023u;
    const int room = kSmemMax - 1024 - kCtrlBytes - static_cast<int>(L->w_smem_bytes);
    int st = room / kABytes;
    L->nstages = st > kMaxStages ? kMaxStages : st;
    const char* dbg = getenv("ESR_DEBUG_SKIP");
    L->debug = dbg ? atoi(dbg) : 0;
    L->prof = g_prof_buf;
    L->pair_nb = 0;
    L->reverse = 0;
    static const bool hints = []() { const char* v = getenv("ESR_NO_L2_HINTS"); return !(v && atoi(v)); }();
    L->a_stream = hints && classify_epilogue(L->d) == kEpiRes;
}

int num_sms_cached();
static int num_sms() { return num_sms_cached(); }
int num_sms_cached() {
    static std::atomic<int> per_dev[64];                  // zero-initialised; keyed by device ordinal
    const int dev = current_device_ordinal();
    int n = per_dev[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        per_dev[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

int launch_conv_tc2(const CUtensorMap& tm0, const CUtensorMap& tm1, const ConvLaunch& L, cudaStream_t stream, int use_pdl);

int launch_conv_tc(const CUtensorMap& tm0, const CUtensorMap& tm1, const ConvLaunch& L, cudaStream_t stream) {
    if (L.pair_nb) return launch_conv_tc2(tm0, tm1, L, stream, g_use_pdl);
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const ConvLaunch);
    static const KernelFn kernels[2][8] = {
        {conv3x3_tc_kernel<16, kEpiGeneric>, conv3x3_tc_kernel<16, kEpiGeneric>, conv3x3_tc_kernel<16, kEpiGeneric>,
         conv3x3_tc_kernel<16, kEpiGeneric>, conv3x3_tc_kernel<16, kEpiNchw>, conv3x3_tc_kernel<16, kEpiGeneric>,
         conv3x3_tc_kernel<16, kEpiGeneric>, conv3x3_tc_kernel<16, kEpiGeneric>},
        {conv3x3_tc_kernel<32, kEpiGeneric>, conv3x3_tc_kernel<32, kEpiTrunk>, conv3x3_tc_kernel<32, kEpiRes>,
         conv3x3_tc_kernel<32, kEpiAct>, conv3x3_tc_kernel<32, kEpiGeneric>, conv3x3_tc_kernel<32, kEpiMask>,
         conv3x3_tc_kernel<32, kEpiDx0>, conv3x3_tc_kernel<32, kEpiGeneric>}};
    ESR_ONCE_PER_DEVICE(
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 8; ++b)
                ESR_CUDA(cudaFuncSetAttribute(kernels[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    );
    ESR_CHECK_ARG(L.nstages >= 2, "conv weights (%u B per cout tile) leave no room for the A-tile ring", L.d.w_tile_bytes);
    // every CTA owns one cout tile: grid is a multiple of cout_tiles, at most one CTA per SM
    int per_ct = num_sms() / L.d.cout_tiles;
    if (per_ct > L.spatial_tiles) per_ct = L.spatial_tiles;
    ESR_CHECK_ARG(per_ct >= 1, "too many cout tiles (%d) for %d SMs", L.d.cout_tiles, num_sms());
    const int grid = per_ct * L.d.cout_tiles;
    const int smem = 1024 + static_cast<int>(L.w_smem_bytes) + L.nstages * kABytes + kCtrlBytes;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernels[L.d.cout_tile == 16 ? 0 : 1][classify_epilogue(L.d)], tm0, tm1, L);
    if (e != cudaSuccess) { set_error("conv3x3_tc_kernel launch failed: %s", cudaGetErrorString(e)); return ESR_ERR_CUDA; }
    return check_launch("conv3x3_tc_kernel");
}

bool fill_launch_pair(ConvLaunch* L);

int build_conv_launch(const esr_conv_desc& d, CUtensorMap* tm0, CUtensorMap* tm1, ConvLaunch* L) {
    int rc = validate_conv_desc(d);
    if (rc != ESR_OK) return rc;
    fill_launch(L, d);
    int box_rows = kHaloRows;
    if (d.pair) {
        ESR_CHECK_ARG(fill_launch_pair(L), "pair mode: weights (%u B per cout tile) leave no room for the A-tile ring", d.w_tile_bytes);
        box_rows = L->pair_nb * kBandRows + 2;
    }
    rc = make_act_tensor_map(tm0, d.src[0], d.B, d.H, d.W, box_rows, 0);
    if (rc != ESR_OK) return rc;
    if (d.src[1].ptr != nullptr) {
        // lean latent blocks: source 1 is loaded without halo rows, 16 channels wide (all or none of its blocks)
        int lean = -1;
        for (int i = 0; i < d.num_kblocks; ++i)
            if (d.kblocks[i].src == 1) {
                ESR_CHECK_ARG(lean < 0 || lean == (d.kblocks[i].half ? 1 : 0), "source 1: lean and full K blocks cannot be mixed");
                lean = d.kblocks[i].half ? 1 : 0;
            }
        rc = make_act_tensor_map(tm1, d.src[1], d.B, d.H, d.W, lean == 1 ? box_rows - 2 : box_rows, lean == 1);
        if (rc != ESR_OK) return rc;
    } else {
        *tm1 = *tm0;
    }
    return ESR_OK;
}

}  // namespace esr

// Debug aid: per-CTA, per-role cycle counters ([gridDim][16] uint64) written by subsequent tcgen05 conv
// launches that are *built* after this call; pass NULL to switch off.
extern "C" void esr_debug_set_profile_buffer(void* buf) { esr::g_prof_buf = static_cast<unsigned long long*>(buf); }

extern "C" int esr_conv3x3_tc(const esr_conv_desc* d, void* stream) {
    if (d == nullptr) { esr::set_error("null conv desc"); return ESR_ERR_INVALID; }
    CUtensorMap tm0, tm1;
    esr::ConvLaunch L;
    int rc = esr::build_conv_launch(*d, &tm0, &tm1, &L);
    if (rc != ESR_OK) return rc;
    return esr::launch_conv_tc(tm0, tm1, L, static_cast<cudaStream_t>(stream));
}
