// conv3x3 on CTA pairs: tcgen05.mma.cta_group::2 (M = 256 = two 4x32-pixel bands, one per CTA of a 2-CTA
// cluster), same formulation as conv3x3_tc.cu (filter columns stacked along N, filter rows as row-shifted views
// of one TMA halo tile, shuffle epilogue).
//
// Why pairs: a single-CTA M128 x N96 x K16 MMA reads 4 KiB of A and 3 KiB of B from shared memory and is bound
// by that read (~96 B/clk -> 74 clk instead of 48, measured).  In a pair each CTA still reads its own 4 KiB of A
// (its own spatial tile) but only HALF of B (the tensor cores exchange the halves), so
//   N = 96  (cout tile 32, 2 bands/tile):  5.5 KiB per CTA per MMA  -> ~57 clk vs 48 ideal,
//   N = 192 (cout tile 64, 1 band/tile):   7 KiB per CTA per 96-clk MMA -> tensor bound.
// Each CTA keeps its half of the cout tile's packed weights resident (<= 124 KiB) and streams its own halo
// tiles through a TMA ring; the leader CTA's elected thread issues every MMA; full / accumulator-empty barriers
// live in the leader and collect remote arrivals, empty / accumulator-full barriers are multicast commits.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <new>

#include "conv3x3.cuh"
#include "ptx_sm100.cuh"

// Per-role cycle counters (tools/prof.py); compiled in only with -DESR_PROFILE_ROLES.
#ifdef ESR_PROFILE_ROLES
#define ESR_PROF(...) __VA_ARGS__
#else
#define ESR_PROF(...)
#endif

namespace esr {

namespace pair {

__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int kMaxStages = 10;
constexpr int kAccStages = 2;
constexpr int kEpiWarps = 8;
constexpr int kNumThreads = 64 + 32 * kEpiWarps;   // 320
// Warp roles.  The schedulers of this architecture favour the HIGHEST warp id among eligible warps, so the two
// single-thread, latency-critical roles sit above the eight epilogue warps (which must be warps 0..7 anyway:
// warp w may only read TMEM lanes 32*(w % 4) ..): a TMA or MMA issue never queues behind epilogue arithmetic.
constexpr int kMmaWarp = kEpiWarps;                // 8  (also owns the TMEM allocation)
constexpr int kProducerWarp = kEpiWarps + 1;       // 9
constexpr int kTmemCols = 512;
constexpr int kCtrlBytes = 256 + 256;              // barriers + tmem slot, 64 floats of bias
constexpr int kSmemMax = 227 * 1024;
constexpr uint32_t kDescHi = ((8u * kRowBytes) >> 4) | (1u << 14) | (4u << 29);
// lean blocks: 32-byte rows, SBO = 256 B, SWIZZLE_32B (layout code 6)
constexpr uint32_t kDescHiLean = ((8u * 32u) >> 4) | (1u << 14) | (6u << 29);

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// Remote arrive WITHOUT cluster-scope release: that form compiles to MEMBAR.ALL.GPU (it would drain the epilogue's
// global stores on every tile).  Nothing written by the generic proxy is handed over through these barriers: the
// accumulator hand-back is ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_local(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA tile load whose completion bytes are credited to the barrier at `leader_bar_addr` (CTA 0 of the pair)
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const void* tmap, uint32_t leader_bar_addr, int c0, int c1,
                                                 int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar_addr), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair_hint(void* smem_dst, const void* tmap, uint32_t leader_bar_addr, int c0,
                                                      int c1, int c2, int c3, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar_addr), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_issue2_hi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}\n"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_issue2(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    umma_issue2_hi(tmem_d, a_lo, b_lo, kDescHi, idesc, accumulate);
}
// arrives on the barrier at this offset in BOTH CTAs once all prior MMAs of the pair have completed
__device__ __forceinline__ void umma_commit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
                 : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16_m256(uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

// CT = output channels per pair tile (32 -> N = 96, two bands per CTA; 64 -> N = 192, one band per CTA).
template <int CT, int MODE>
__global__ void __launch_bounds__(kNumThreads, 1)
conv3x3_tc2_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                   const __grid_constant__ ConvLaunch L) {
    constexpr int N = 3 * CT;                      // MMA N of the pair
    constexpr int NB = CT == 32 ? 2 : 1;           // bands per CTA tile
    constexpr int kRows = NB * kBandRows + 2;      // halo rows per A tile
    constexpr int kATile = kRows * kTileW * kRowBytes;
    constexpr int kLeanTile = NB * kBandRows * kTileW * 32;   // lean latent tile: no halo rows, 32-byte rows
    constexpr uint32_t kLeanRow16 = (kTileW * 32) >> 4;
    constexpr int kAccSlot = CT == 32 ? 128 : 256; // TMEM columns per band accumulator
    // operand format bits: a_format [7,10) and b_format [10,13) are 1 for bf16, 0 for fp16
    const uint32_t kIdesc = idesc_bf16_m256(N) & ((L.d.flags & ESR_CONV_F16) ? ~((1u << 7) | (1u << 10)) : ~0u);
    constexpr uint32_t kARow16 = (kTileW * kRowBytes) >> 4;
    constexpr uint32_t kWSlab16 = ((N / 2) * kRowBytes) >> 4;     // this CTA's half of one [N x 32ch] slab
    const esr_conv_desc& d = L.d;
    const int nstages = L.nstages;
    // The lean latent block (last in the list) rides in the stage of the last main block: its 2 MMAs per band would
    // otherwise pay a whole stage turn-over (barrier wait + commit, ~250 clk of MMA-thread time for 96 clk of tensor
    // work - removing the latent blocks altogether made the step 10 % faster, shrinking their tiles did nothing).
    const bool attach = d.num_kblocks >= 2 && d.kblocks[d.num_kblocks - 1].half != 0;
    const int nkb = attach ? d.num_kblocks - 1 : d.num_kblocks;          // stages per tile
    const int stage_bytes = kATile + (attach ? kLeanTile : 0);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_w = smem;
    uint8_t* s_a = smem + L.w_smem_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_a + nstages * stage_bytes);
    uint64_t* full_bar = bars;                      // waited in the leader only
    uint64_t* empty_bar = bars + kMaxStages;
    uint64_t* acc_full = bars + 2 * kMaxStages;
    uint64_t* acc_empty = acc_full + kAccStages;    // waited in the leader only
    uint64_t* w_full = acc_empty + kAccStages;      // local weight copy
    uint64_t* w_ready = w_full + 1;                 // leader: the peer's half is resident too
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_ready + 1);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    ESR_PROF(if (L.prof && threadIdx.x == 0) L.prof[blockIdx.x * 16 + 12] = gtime_ns();)
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int ct = cluster_id % d.cout_tiles;
    const int pair0 = cluster_id / d.cout_tiles, pair_step = num_clusters / d.cout_tiles;
    const int num_pairs = (L.spatial_tiles + 1) >> 1;
    for (int i = threadIdx.x; i < CT; i += kNumThreads) s_bias[i] = d.bias[ct * CT + i];

    if (warp == kProducerWarp && lane == 0) {
        tma_prefetch_desc(&tmap0);
        tma_prefetch_desc(&tmap1);
        for (int s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * kEpiWarps); }
        mbar_init(w_full, 1);
        mbar_init(w_ready, 1);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc2(tmem_slot, kTmemCols);
        tmem_relinquish2();
    }
    tc_fence_before();
    cluster_sync_all();                             // both CTAs' barriers and TMEM are ready
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ESR_PROF(if (L.prof && threadIdx.x == 0) L.prof[blockIdx.x * 16 + 13] = gtime_ns();)
    pdl_launch_dependents();

    const int tiles_per_img = L.tiles_x * L.tiles_y;

    if (warp == kProducerWarp) {
        // ------------------------------------------------------------ TMA producer (one per CTA)
        if (elect_one()) {
            const uint32_t half_bytes = d.w_tile_bytes >> 1;                         // this CTA's half image
            const uint8_t* wt = reinterpret_cast<const uint8_t*>(d.wpack) + static_cast<size_t>(ct) * d.w_tile_bytes +
                                static_cast<size_t>(rank) * half_bytes;
            mbar_expect_tx_local(w_full, half_bytes);
            for (uint32_t off = 0; off < half_bytes; off += 16384) {
                const uint32_t n = half_bytes - off < 16384 ? half_bytes - off : 16384;
                bulk_load_1d(s_w + off, wt + off, n, w_full);
            }
            pdl_wait();
            uint32_t stage = 0, phase = 0;
            const uint64_t stream_policy = l2_policy_evict_first();
            ESR_PROF(long long p_wait = 0, p_t0 = clock64(), p_n = 0; if (L.prof) L.prof[blockIdx.x * 16 + 6] = gtime_ns();)
            for (int p = pair0; p < num_pairs; p += pair_step) {
                int sp = 2 * (L.reverse ? num_pairs - 1 - p : p) + static_cast<int>(rank);
                if (sp >= L.spatial_tiles) sp = L.spatial_tiles - 1;       // odd tail: duplicate tile, stores masked
                const int n = sp / tiles_per_img;
                const int r = sp - n * tiles_per_img;
                const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
                const int x0 = tx * kTileWOut - 1, y0 = ty * (NB * kBandRows) - 1;
                for (int kb = 0; kb < nkb; ++kb) {
                    const esr_kblock& K = d.kblocks[kb];
                    ESR_PROF(const long long w0c = clock64();)
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    ESR_PROF(p_wait += clock64() - w0c; ++p_n;)
                    const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
                    // only the leader arrives, expecting both CTAs' tiles; the peer's bytes may be credited before
                    // that (a transiently negative tx-count is legal), never to an older phase: the peer issues
                    // only after its copy of the multicast "stage free" commit
                    uint8_t* dst = s_a + stage * stage_bytes;
                    if (K.half) {                   // lean latent block on its own (no main block before it)
                        if (rank == 0) mbar_expect_tx_local(&full_bar[stage], 2 * kLeanTile);
                        tma_load_4d_pair(dst, &tmap1, lead_full, K.chan + ((K.slice_mask & 1) ? 0 : 16), x0, y0 + 1, n);
                    } else {
                        const bool with_lean = attach && kb == nkb - 1;
                        if (rank == 0) mbar_expect_tx_local(&full_bar[stage], 2 * (kATile + (with_lean ? kLeanTile : 0)));
                        if (L.a_stream && K.src == 0)
                            tma_load_4d_pair_hint(dst, &tmap0, lead_full, K.chan, x0, y0, n, stream_policy);
                        else
                            tma_load_4d_pair(dst, K.src == 0 ? &tmap0 : &tmap1, lead_full, K.chan, x0, y0, n);
                        if (with_lean) {            // centre rows only, 16 channels (SWIZZLE_32B), same stage / barrier
                            const esr_kblock& Kl = d.kblocks[nkb];
                            tma_load_4d_pair(dst + kATile, &tmap1, lead_full, Kl.chan + ((Kl.slice_mask & 1) ? 0 : 16), x0, y0 + 1, n);
                        }
                    }
                    if (++stage == static_cast<uint32_t>(nstages)) { stage = 0; phase ^= 1; }
                }
            }
            ESR_PROF(if (L.prof) {
                unsigned long long* o = L.prof + blockIdx.x * 16;
                o[0] = clock64() - p_t0; o[1] = p_wait; o[2] = p_n;
            })
        }
    } else if (warp == kMmaWarp) {
        // -------------------------------------------------------------- MMA issuer (leader CTA only)
        if (rank != 0) {
            // the peer's weight half has landed: tell the leader (the producer thread must not wait for it, it
            // would stall the A ring; a tile may have more K blocks than the ring has stages)
            if (elect_one()) {
                mbar_wait(w_full, 0);
                mbar_arrive_cluster(mapa_u32(smem_u32(w_ready), 0));
            }
        } else if (elect_one()) {
            const uint32_t w_lo = smem_u32(s_w) >> 4, a_lo = smem_u32(s_a) >> 4;
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            ESR_PROF(long long m_t0 = clock64(), m_wacc = 0, m_wfull = 0; bool first_full = true;)
            mbar_wait(w_full, 0);
            mbar_wait(w_ready, 0);
            ESR_PROF(const long long m_tw = clock64() - m_t0;)
            for (int p = pair0; p < num_pairs; p += pair_step) {
                ESR_PROF(long long c0 = clock64();)
                mbar_wait(&acc_empty[as], aphase ^ 1);
                ESR_PROF(m_wacc += clock64() - c0;)
                tc_fence_after();
                const uint32_t acc0 = tmem_base + as * (NB * kAccSlot);
                uint32_t nonfirst = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint32_t masks = *reinterpret_cast<const uint32_t*>(&d.kblocks[kb].dy_mask);
                    const uint32_t dy_mask = masks & 0xff, slice_mask = (masks >> 8) & 0xff;
                    const uint32_t w0 = w_lo + (d.kblocks[kb].w_off >> 4);
                    ESR_PROF(c0 = clock64();)
                    mbar_wait(&full_bar[stage], phase);
                    ESR_PROF(m_wfull += clock64() - c0; if (first_full && L.prof) { L.prof[blockIdx.x * 16 + 10] = gtime_ns(); first_full = false; })
                    tc_fence_after();
                    const uint32_t a0 = a_lo + stage * (stage_bytes >> 4);
                    if (masks >> 24) {                // lean latent block: one K=16 MMA per band, SWIZZLE_32B operands
#pragma unroll
                        for (int b = 0; b < NB; ++b)
                            umma_issue2_hi(acc0 + b * kAccSlot, a0 + (b * kBandRows) * kLeanRow16, w0, kDescHiLean, kIdesc, nonfirst);
                    } else if (dy_mask == 7u && slice_mask == 3u) {
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                            for (int s = 0; s < 2; ++s) {
#pragma unroll
                                for (int b = 0; b < NB; ++b) {
                                    umma_issue2(acc0 + b * kAccSlot, a0 + (b * kBandRows + dy) * kARow16 + s * 2,
                                                w0 + dy * kWSlab16 + s * 2, kIdesc, (dy | s) ? 1u : nonfirst);
                                }
                            }
                        }
                    } else {
                        for (int b = 0; b < NB; ++b) {
                            uint32_t acc_flag = nonfirst, wi = 0;
                            for (int dy = 0; dy < 3; ++dy) {
                                if (!((dy_mask >> dy) & 1u)) continue;
                                for (int s = 0; s < 2; ++s) {
                                    if (!((slice_mask >> s) & 1u)) continue;
                                    umma_issue2(acc0 + b * kAccSlot, a0 + (b * kBandRows + dy) * kARow16 + s * 2,
                                                w0 + wi * kWSlab16 + s * 2, kIdesc, acc_flag);
                                    acc_flag = 1;
                                }
                                ++wi;
                            }
                        }
                    }
                    nonfirst = 1;
                    if (attach && kb == nkb - 1) {   // the latent rows, loaded behind this block's tile
                        const uint32_t wl = w_lo + (d.kblocks[nkb].w_off >> 4);
#pragma unroll
                        for (int b = 0; b < NB; ++b)
                            umma_issue2_hi(acc0 + b * kAccSlot, a0 + (kATile >> 4) + (b * kBandRows) * kLeanRow16, wl, kDescHiLean, kIdesc, 1u);
                    }
                    umma_commit2(&empty_bar[stage]);
                    if (++stage == static_cast<uint32_t>(nstages)) { stage = 0; phase ^= 1; }
                }
                umma_commit2(&acc_full[as]);
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
            }
            ESR_PROF(if (L.prof) {
                unsigned long long* o = L.prof + blockIdx.x * 16;
                o[3] = clock64() - m_t0; o[4] = m_wacc; o[5] = m_wfull; o[9] = m_tw; o[14] = gtime_ns();
            })
        }
    } else {
        // ---------------------------------------------------------------- epilogue (each CTA: its own tile)
        const int wq = warp & 3;
        const int e = warp >> 2;                         // 0/1
        const int b = NB == 2 ? e : 0;                 // band (N = 96) ...
        const int cbase = NB == 2 ? 0 : e * 32;        // ... or 32-channel half of the 64 (N = 192)
        const uint32_t lead_acc_empty0 = mapa_u32(smem_u32(&acc_empty[0]), 0);
        pdl_wait();
        uint32_t as = 0, aphase = 0;
        ESR_PROF(long long e_t0 = clock64(), e_wait = 0, e_n = 0;)
        for (int p = pair0; p < num_pairs; p += pair_step) {
            const int sp_raw = 2 * (L.reverse ? num_pairs - 1 - p : p) + static_cast<int>(rank);
            const bool tile_ok = sp_raw < L.spatial_tiles;
            const int sp = tile_ok ? sp_raw : L.spatial_tiles - 1;
            const int n = sp / tiles_per_img;
            const int r = sp - n * tiles_per_img;
            const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
            const int x = tx * kTileWOut - 1 + lane;
            const int y = ty * (NB * kBandRows) + b * kBandRows + wq;
            const bool ok = tile_ok && lane >= 1 && lane <= kTileWOut && x < d.W && y < d.H;
            EpiOperands ops[2];
            if (ok) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    conv_epilogue_prefetch<MODE>(d, ct, n, y, x, tile_channel(d, ct, cbase + h * 16), ops[h]);
            }
            ESR_PROF(long long c0 = clock64();)
            mbar_wait(&acc_full[as], aphase);
            ESR_PROF(e_wait += clock64() - c0; ++e_n;)
            tc_fence_after();
            const uint32_t taddr = tmem_base + as * (NB * kAccSlot) + b * kAccSlot + cbase + (static_cast<uint32_t>(wq * 32) << 16);
            float vc[32];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float vl[16], vr[16];
                tmem_ld_x16(taddr + 0 * CT + h * 16, vl);
                tmem_ld_x16(taddr + 1 * CT + h * 16, *reinterpret_cast<float(*)[16]>(&vc[h * 16]));
                tmem_ld_x16(taddr + 2 * CT + h * 16, vr);
                tmem_ld_wait();
                if (h == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(lead_acc_empty0 + as * 8);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float fl = __shfl_up_sync(0xffffffffu, vl[i], 1);
                    const float fr = __shfl_down_sync(0xffffffffu, vr[i], 1);
                    vc[h * 16 + i] += fl + fr;
                }
            }
            if (ok) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    conv_epilogue16<MODE>(d, s_bias + cbase + h * 16, ct, n, y, x, tile_channel(d, ct, cbase + h * 16),
                                          *reinterpret_cast<float(*)[16]>(&vc[h * 16]), ops[h]);
            }
            if (++as == kAccStages) { as = 0; aphase ^= 1; }
        }
        ESR_PROF(if (L.prof && warp == 0 && lane == 1) {
            unsigned long long* o = L.prof + blockIdx.x * 16;
            o[7] = clock64() - e_t0; o[8] = e_wait; o[11] = e_n; o[15] = gtime_ns();
        })
    }

    tc_fence_before();
    cluster_sync_all();                             // nobody frees TMEM / exits while the pair's MMAs may still read it
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, kTmemCols);
    }
}


// ============================================================ fused growth convs of a dense block
// One persistent launch = up to four convs (cout 32 each) over the same NHWC buffer; see esr_rdb_growth_desc.
// Work items are (chunk, layer, tile pair) in that order; cluster k takes items k, k+74, ...  An item of layer
// l > 0 may be loaded once layer l-1 of its tile's 3x3 neighbourhood has been stored: every epilogue warp bumps
// a per-(layer, tile) counter after its stores (generic->async proxy fence + __threadfence + atomicAdd), the TMA
// producer polls the up-to-nine counters (acquire) before its first load of the item.  Every dependency points to
// an item with a smaller index, all clusters are resident and walk their items in increasing order, and an item's
// MMAs / epilogue never wait on a flag, so the launch cannot deadlock.
struct RdbLayerDev {
    int nkb, out_choff;
    uint32_t w_smem_off, w_half_bytes;      // this CTA's half image: offset in the resident weight region, size
    const uint8_t* wpack;                   // cout tile image in global memory (two halves)
    const float* bias;
    esr_kblock kb[ESR_RDB_MAX_KBLOCKS];
};
struct RdbLaunch {
    int B, H, W, nlayers;
    RdbLayerDev layer[ESR_RDB_MAX_LAYERS];
    __nv_bfloat16* out;
    int out_stride, mode;
    float slope;
    const uint16_t* mask;
    int mask_stride;
    int tiles_x, tiles_y, tiles_per_img, tiles_c, ppc, chunks, total_items, spatial_tiles;
    uint32_t* flags;                        // [nlayers][spatial_tiles] counters of this launch
    uint32_t* flags_zero;                   // same size, cleared for a later launch
    int reverse;                            // walk chunks and tiles last-to-first (alternates launch by launch: L2 reuse)
    uint32_t m_per_chunk, m_ppc, m_tpi, m_tx;   // ceil(2^32 / d): n / d == __umulhi(n, m) for n * d < 2^32 (no XU divisions); 0 for d == 1
    int nstages;
    uint32_t w_smem_bytes;
    unsigned long long* prof;               // optional [gridDim][16] role counters (-DESR_PROFILE_ROLES)
};

constexpr int kRdbStages = 10;
constexpr int kRdbThreads = kNumThreads + 64;   // + publisher warp (GPU-scope fence + counter bump, off the epilogue's path)
                                                // + second TMA producer warp (a producer thread needs ~450-700 clk per
                                                //   K block: one alone cannot feed the 576-clk MMAs of a block)
constexpr int kRdbProducerB = kProducerWarp + 1; // warp index of the second producer
constexpr int kRdbPublisher = kProducerWarp + 2; // and of the publisher

// Dependency counters are read with a RELAXED load: an acquire load blocks the producer warp until it returns
// (~800 clk per item, on the path that feeds the MMAs).  Ordering is kept by construction instead: the publisher
// makes the tile's stores visible at GPU scope (fence + proxy fence) before it bumps the counter, and the TMA
// loads of the dependent item are only issued - control dependent - after the bumped value has been observed.
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct RdbItem { int layer, n, ty, tx, gt; bool ok; };
// n / d through the precomputed magic m = ceil(2^32 / d).  d == 1 has no 32-bit magic (2^32 wraps to 0, and
// __umulhi(n, 0) would be 0 instead of n - the round-1 fault on plans with one tile column / one pair per chunk):
// the host stores 0 for it and the division is the identity.
__device__ __forceinline__ int fast_div(int n, uint32_t m) {
    return m == 0u ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), m));
}
__device__ __forceinline__ RdbItem rdb_decode(const RdbLaunch& R, int item, uint32_t rank) {
    const int per_chunk = R.nlayers * R.ppc;
    int c = fast_div(item, R.m_per_chunk);
    const int rem = item - c * per_chunk;
    RdbItem it;
    it.layer = fast_div(rem, R.m_ppc);
    int p = rem - it.layer * R.ppc;
    if (R.reverse) { c = R.chunks - 1 - c; p = R.ppc - 1 - p; }
    int tl = 2 * p + static_cast<int>(rank);
    it.ok = tl < R.tiles_c;
    if (!it.ok) tl = R.tiles_c - 1;                                 // odd tail: duplicate tile, nothing stored
    it.gt = c * R.tiles_c + tl;
    it.n = fast_div(it.gt, R.m_tpi);
    const int r = it.gt - it.n * R.tiles_per_img;
    it.ty = fast_div(r, R.m_tx);
    it.tx = r - it.ty * R.tiles_x;
    return it;
}

template <int MODE>   // kEpiTrunk: bias + LeakyReLU;  kEpiMask: (acc + bias) * LeakyReLU'(mask)
__global__ void __launch_bounds__(kRdbThreads, 1)
conv3x3_rdb_growth_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
                          const __grid_constant__ RdbLaunch R) {
    constexpr int CT = 32, N = 96, NB = 2;
    constexpr int kATile = (NB * kBandRows + 2) * kTileW * kRowBytes;   // 20 KiB
    constexpr int kLeanTile = NB * kBandRows * kTileW * 32;             // 8 KiB
    constexpr uint32_t kLeanRow16 = (kTileW * 32) >> 4;
    constexpr int kAccSlot = 128;
    constexpr uint32_t kIdesc = idesc_bf16_m256(N);
    constexpr uint32_t kARow16 = (kTileW * kRowBytes) >> 4;
    constexpr uint32_t kWSlab16 = ((N / 2) * kRowBytes) >> 4;
    const int nstages = R.nstages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_w = smem;
    uint8_t* s_a = smem + R.w_smem_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_a + nstages * kATile);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kRdbStages;
    uint64_t* acc_full = bars + 2 * kRdbStages;
    uint64_t* acc_empty = acc_full + kAccStages;
    uint64_t* w_full = acc_empty + kAccStages;
    uint64_t* w_ready = w_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_ready + 1);
    uint32_t* stored_cnt = tmem_slot + 1;           // monotonic: +1 per epilogue warp per item whose stores are issued
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [nlayers][32]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    for (int i = threadIdx.x; i < R.nlayers * CT; i += kRdbThreads) s_bias[i] = R.layer[i / CT].bias[i % CT];

    if (warp == kProducerWarp && lane == 0) {
        tma_prefetch_desc(&tmap0);
        tma_prefetch_desc(&tmap1);
        for (int s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 2 * kEpiWarps); }
        *stored_cnt = 0u;
        mbar_init(w_full, 1);
        mbar_init(w_ready, 1);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc2(tmem_slot, kTmemCols);
        tmem_relinquish2();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();

    if (warp == kProducerWarp || warp == kRdbProducerB) {
        // ------------------------------------------------------------ TMA producers (two warps per CTA: K block q of the
        // launch-wide sequence goes to producer q & 1; both walk all items and poll the dependency counters)
        const uint32_t who = warp == kProducerWarp ? 0u : 1u;
        if (who == 0 && elect_one()) {
            uint32_t wbytes = 0;
            for (int l = 0; l < R.nlayers; ++l) wbytes += R.layer[l].w_half_bytes;
            mbar_expect_tx_local(w_full, wbytes);
            for (int l = 0; l < R.nlayers; ++l) {
                const RdbLayerDev& Ly = R.layer[l];
                const uint8_t* wt = Ly.wpack + static_cast<size_t>(rank) * Ly.w_half_bytes;
                for (uint32_t off = 0; off < Ly.w_half_bytes; off += 16384) {
                    const uint32_t n = Ly.w_half_bytes - off < 16384 ? Ly.w_half_bytes - off : 16384;
                    bulk_load_1d(s_w + Ly.w_smem_off + off, wt + off, n, w_full);
                }
            }
        }
        __syncwarp();
        pdl_wait();
        // Lane l < 9 watches neighbour (l/3-1, l%3-1) of the item's tile: one L2 round trip per item for all nine
        // counters, and the counters of the NEXT item are requested before this item's loads are issued, so the
        // answer is normally back (and positive: the neighbours ran ~2 rounds earlier) when it is needed.
        const int ndy = lane / 3 - 1, ndx = lane % 3 - 1;
        auto poll = [&](int item) -> uint32_t {
            if (lane >= 9 || item >= R.total_items) return kEpiWarps;
            const RdbItem it = rdb_decode(R, item, rank);
            const int yy = it.ty + ndy, xx = it.tx + ndx;
            if (it.layer == 0 || yy < 0 || yy >= R.tiles_y || xx < 0 || xx >= R.tiles_x) return kEpiWarps;
#ifdef ESR_RDB_NO_DEPS
            return kEpiWarps;      // timing experiment only (results invalid)
#endif
            return ld_acquire_u32(R.flags + static_cast<size_t>(it.layer - 1) * R.spatial_tiles +
                                  static_cast<size_t>(it.n) * R.tiles_per_img + yy * R.tiles_x + xx);
        };
        uint32_t q = 0;                                   // running K-block number (same in both producers)
        uint32_t seen = poll(cluster_id);
        ESR_PROF(long long p_t0 = clock64(), p_poll = 0, p_wait = 0, p_items = 0, p_respin = 0, p_issue = 0, p_sync = 0, p_poll2 = 0;)
        for (int item = cluster_id; item < R.total_items; item += num_clusters) {
            ESR_PROF(const long long pc0 = clock64(); ++p_items;)
            ESR_PROF(const int tslot = (item - cluster_id) / num_clusters;
                     if (R.prof && blockIdx.x == 0 && who == 0 && lane == 0 && tslot < 32) R.prof[148 * 16 + tslot * 8 + 0] = gtime_ns();)
#ifndef ESR_RDB_NO_POLL
            for (uint32_t spin = 0; !__all_sync(0xffffffffu, seen >= static_cast<uint32_t>(kEpiWarps)); ++spin) {
                if (spin > (1u << 24)) { __trap(); }
                seen = poll(item);
                ESR_PROF(++p_respin;)
            }
#endif
            ESR_PROF(p_poll += clock64() - pc0;)
            // (the generic->async proxy fence of this hand-over sits on the writer side, in publish(): a full proxy
            // fence here would also wait for this thread's TMA loads in flight and serialise the items)
            ESR_PROF(const long long pq0 = clock64();)
#ifndef ESR_RDB_NO_POLL
            seen = poll(item + num_clusters);
#endif
            ESR_PROF(p_poll2 += clock64() - pq0;)
            {
                const RdbItem it = rdb_decode(R, item, rank);
                const RdbLayerDev& Ly = R.layer[it.layer];
                if (lane == 0) {
                    const int x0 = it.tx * kTileWOut - 1, y0 = it.ty * (NB * kBandRows) - 1;
                    for (int kb = 0; kb < Ly.nkb; ++kb) {
                        const uint32_t qk = q + kb;
                        if ((qk & 1u) != who) continue;
                        const uint32_t stage = qk % static_cast<uint32_t>(nstages);
                        const uint32_t phase = (qk / static_cast<uint32_t>(nstages)) & 1u;
                        const esr_kblock& K = Ly.kb[kb];
                        ESR_PROF(const long long w0c = clock64();)
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        ESR_PROF(p_wait += clock64() - w0c; const long long i0c = clock64();)
                        const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        if (K.half) {                 // lean latent block: centre rows only, 16 channels (SWIZZLE_32B)
                            if (rank == 0) mbar_expect_tx_local(&full_bar[stage], 2 * kLeanTile);
                            tma_load_4d_pair(s_a + stage * kATile, &tmap1, lead_full, K.chan + ((K.slice_mask & 1) ? 0 : 16), x0,
                                             y0 + 1, it.n);
                        } else {
                            if (rank == 0) mbar_expect_tx_local(&full_bar[stage], 2 * kATile);
                            tma_load_4d_pair(s_a + stage * kATile, K.src == 0 ? &tmap0 : &tmap1, lead_full, K.chan, x0, y0, it.n);
                        }
                        ESR_PROF(p_issue += clock64() - i0c;)
                    }
                    ESR_PROF(if (R.prof && blockIdx.x == 0 && who == 0 && tslot < 32) R.prof[148 * 16 + tslot * 8 + 1] = gtime_ns();)
                }
                q += Ly.nkb;
            }
            ESR_PROF(const long long s0c = clock64();)
            __syncwarp();
            ESR_PROF(p_sync += clock64() - s0c;)
        }
        ESR_PROF(if (R.prof && lane == 0 && who == 0) {
            unsigned long long* o = R.prof + blockIdx.x * 16;
            o[0] = clock64() - p_t0; o[1] = p_wait; o[2] = p_items; o[6] = p_poll; o[10] = p_respin;
            o[12] = p_issue; o[13] = p_sync; o[14] = p_poll2;
        })
    } else if (warp == kMmaWarp) {
        // -------------------------------------------------------------- MMA issuer (leader CTA only)
        if (rank != 0) {
            if (elect_one()) {
                mbar_wait(w_full, 0);
                mbar_arrive_cluster(mapa_u32(smem_u32(w_ready), 0));
            }
        } else if (elect_one()) {
            const uint32_t w_lo = smem_u32(s_w) >> 4, a_lo = smem_u32(s_a) >> 4;
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            ESR_PROF(long long m_t0 = clock64(), m_wacc = 0, m_wfull = 0;)
            mbar_wait(w_full, 0);
            mbar_wait(w_ready, 0);
            ESR_PROF(const long long m_tw = clock64() - m_t0;)
            const int per_chunk = R.nlayers * R.ppc;
            for (int item = cluster_id; item < R.total_items; item += num_clusters) {
                const int c_ = fast_div(item, R.m_per_chunk);
                const RdbLayerDev& Ly = R.layer[fast_div(item - c_ * per_chunk, R.m_ppc)];
                ESR_PROF(long long c0 = clock64(); const int tslot = (item - cluster_id) / num_clusters;)
                mbar_wait(&acc_empty[as], aphase ^ 1);
                ESR_PROF(m_wacc += clock64() - c0; if (R.prof && blockIdx.x == 0 && tslot < 32) R.prof[148 * 16 + tslot * 8 + 2] = gtime_ns();)
                tc_fence_after();
                const uint32_t acc0 = tmem_base + as * (NB * kAccSlot);
                uint32_t nonfirst = 0;
                for (int kb = 0; kb < Ly.nkb; ++kb) {
                    const uint32_t masks = *reinterpret_cast<const uint32_t*>(&Ly.kb[kb].dy_mask);
                    const uint32_t dy_mask = masks & 0xff, slice_mask = (masks >> 8) & 0xff;
                    const uint32_t w0 = w_lo + ((Ly.w_smem_off + Ly.kb[kb].w_off) >> 4);
                    ESR_PROF(c0 = clock64();)
                    mbar_wait(&full_bar[stage], phase);
                    ESR_PROF(m_wfull += clock64() - c0;
                             if (R.prof && blockIdx.x == 0 && tslot < 32 && kb == 0) R.prof[148 * 16 + tslot * 8 + 3] = gtime_ns();
                             if (R.prof && blockIdx.x == 0 && tslot < 32 && kb == Ly.nkb - 1) R.prof[148 * 16 + tslot * 8 + 4] = gtime_ns();)
                    tc_fence_after();
                    const uint32_t a0 = a_lo + stage * (kATile >> 4);
                    if (masks >> 24) {                // lean latent block: one K=16 MMA per band, SWIZZLE_32B operands
#pragma unroll
                        for (int b = 0; b < NB; ++b)
                            umma_issue2_hi(acc0 + b * kAccSlot, a0 + (b * kBandRows) * kLeanRow16, w0, kDescHiLean, kIdesc, nonfirst);
                    } else if (dy_mask == 7u && slice_mask == 3u) {
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                            for (int s = 0; s < 2; ++s) {
#pragma unroll
                                for (int b = 0; b < NB; ++b) {
                                    umma_issue2(acc0 + b * kAccSlot, a0 + (b * kBandRows + dy) * kARow16 + s * 2,
                                                w0 + dy * kWSlab16 + s * 2, kIdesc, (dy | s) ? 1u : nonfirst);
                                }
                            }
                        }
                    } else {
                        for (int b = 0; b < NB; ++b) {
                            uint32_t acc_flag = nonfirst, wi = 0;
                            for (int dy = 0; dy < 3; ++dy) {
                                if (!((dy_mask >> dy) & 1u)) continue;
                                for (int s = 0; s < 2; ++s) {
                                    if (!((slice_mask >> s) & 1u)) continue;
                                    umma_issue2(acc0 + b * kAccSlot, a0 + (b * kBandRows + dy) * kARow16 + s * 2,
                                                w0 + wi * kWSlab16 + s * 2, kIdesc, acc_flag);
                                    acc_flag = 1;
                                }
                                ++wi;
                            }
                        }
                    }
                    nonfirst = 1;
                    umma_commit2(&empty_bar[stage]);
                    if (++stage == static_cast<uint32_t>(nstages)) { stage = 0; phase ^= 1; }
                }
                umma_commit2(&acc_full[as]);
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
            }
            ESR_PROF(if (R.prof) {
                unsigned long long* o = R.prof + blockIdx.x * 16;
                o[3] = clock64() - m_t0; o[4] = m_wacc; o[5] = m_wfull; o[9] = m_tw;
            })
        }
    } else if (warp == kRdbPublisher) {
        // ---------------------------------------------------------------- publisher (one thread per CTA)
        // Waits until the eight epilogue warps have issued the stores of an item (mbarrier, release/acquire at CTA
        // scope: cumulative), makes them visible GPU-wide and bumps the tile's counter.  The ~0.7 us of the GPU-scope
        // fence is paid here, not by the epilogue warps.
#ifdef ESR_RDB_NO_PUBLISHER
        if (false) {
#else
        if (lane == 0) {
#endif
            uint32_t want = 0;
            const uint32_t cnt_addr = smem_u32(stored_cnt);
            for (int item = cluster_id; item < R.total_items; item += num_clusters) {
                const RdbItem it = rdb_decode(R, item, rank);
                want += kEpiWarps;
                for (uint32_t spin = 0;; ++spin) {               // a counter, not an mbarrier: lagging behind cannot alias phases
                    uint32_t v;
                    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(cnt_addr) : "memory");
                    if (v >= want) break;
                    __nanosleep(200);                            // do not steal issue slots from the epilogue warps of this scheduler
                    if (spin > (1u << 24)) { __trap(); }
                }
                if (it.ok && it.layer + 1 < R.nlayers) {
                    __threadfence();
                    atomicAdd(R.flags + static_cast<size_t>(it.layer) * R.spatial_tiles + it.gt, static_cast<uint32_t>(kEpiWarps));
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue (each CTA: its own tile)
        const int wq = warp & 3;
        const int b = warp >> 2;
        const uint32_t lead_acc_empty0 = mapa_u32(smem_u32(&acc_empty[0]), 0);
        pdl_wait();
        {   // clear the counter third a later launch will use (nobody reads it during this launch)
            const int nwords = R.nlayers * R.spatial_tiles;
            const int t = blockIdx.x * (kEpiWarps * 32) + threadIdx.x;          // epilogue threads are 0 .. 255
            for (int i = t; i < nwords; i += gridDim.x * kEpiWarps * 32) R.flags_zero[i] = 0u;
        }
        uint32_t as = 0, aphase = 0;
        ESR_PROF(long long e_t0 = clock64(), e_wait = 0, e_n = 0;)
        for (int item = cluster_id; item < R.total_items; item += num_clusters) {
            const RdbItem it = rdb_decode(R, item, rank);
            const RdbLayerDev& Ly = R.layer[it.layer];
            const int x = it.tx * kTileWOut - 1 + lane;
            const int y = it.ty * (NB * kBandRows) + b * kBandRows + wq;
            const bool ok = it.ok && lane >= 1 && lane <= kTileWOut && x < R.W && y < R.H;
            const size_t pix = (static_cast<size_t>(it.n) * R.H + y) * R.W + x;
            uint4 mk[4];
            if (MODE == kEpiMask && ok) {
                const uint4* m = reinterpret_cast<const uint4*>(R.mask + pix * R.mask_stride + Ly.out_choff);
#pragma unroll
                for (int i = 0; i < 4; ++i) mk[i] = __ldg(m + i);
            }
            ESR_PROF(const long long ec0 = clock64(); const int tslot = (item - cluster_id) / num_clusters;)
            mbar_wait(&acc_full[as], aphase);
            ESR_PROF(e_wait += clock64() - ec0; ++e_n;
                     if (R.prof && blockIdx.x == 0 && warp == 0 && lane == 1 && tslot < 32) R.prof[148 * 16 + tslot * 8 + 5] = gtime_ns();)
            tc_fence_after();
            const uint32_t taddr = tmem_base + as * (NB * kAccSlot) + b * kAccSlot + (static_cast<uint32_t>(wq * 32) << 16);
            float vc[32];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float vl[16], vr[16];
                tmem_ld_x16(taddr + 0 * CT + h * 16, vl);
                tmem_ld_x16(taddr + 1 * CT + h * 16, *reinterpret_cast<float(*)[16]>(&vc[h * 16]));
                tmem_ld_x16(taddr + 2 * CT + h * 16, vr);
                tmem_ld_wait();
                if (h == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(lead_acc_empty0 + as * 8);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float fl = __shfl_up_sync(0xffffffffu, vl[i], 1);
                    const float fr = __shfl_down_sync(0xffffffffu, vr[i], 1);
                    vc[h * 16 + i] += fl + fr;
                }
            }
            if (ok) {
                const float* bias = s_bias + it.layer * CT;
                uint32_t pk[16];
                const uint16_t* mv = reinterpret_cast<const uint16_t*>(mk);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float v0 = vc[2 * i] + bias[2 * i], v1 = vc[2 * i + 1] + bias[2 * i + 1];
                    if (MODE == kEpiMask) {
                        v0 *= ((mv[2 * i] & 0x8000u) == 0 && (mv[2 * i] & 0x7fffu) != 0) ? 1.f : R.slope;
                        v1 *= ((mv[2 * i + 1] & 0x8000u) == 0 && (mv[2 * i + 1] & 0x7fffu) != 0) ? 1.f : R.slope;
                    } else {
                        v0 = fmaxf(v0, R.slope * v0);
                        v1 = fmaxf(v1, R.slope * v1);
                    }
                    pk[i] = pack_bf16x2(v0, v1);
                }
                __nv_bfloat16* o = R.out + pix * R.out_stride + Ly.out_choff;
                st_global_v8(o, *reinterpret_cast<const uint32_t(*)[8]>(&pk[0]));
                st_global_v8(o + 16, *reinterpret_cast<const uint32_t(*)[8]>(&pk[8]));
            }
            // generic-proxy stores -> later TMA (async proxy) reads by other SMs: every writer fences the proxies, the
            // publisher warp does the GPU-scope part
            ESR_PROF(if (R.prof && blockIdx.x == 0 && warp == 0 && lane == 1 && tslot < 32) R.prof[148 * 16 + tslot * 8 + 6] = gtime_ns();)
#ifndef ESR_RDB_NO_SIGNAL
#ifndef ESR_RDB_NO_PROXY_FENCE
            asm volatile("fence.proxy.async.global;" ::: "memory");
#endif
            __syncwarp();
#endif
            if (lane == 0) asm volatile(
#ifdef ESR_RDB_NO_SIGNAL
                "red.relaxed.cta.shared::cta.add.u32 [%0], 1;"
#else
                "red.release.cta.shared::cta.add.u32 [%0], 1;"
#endif
                ::"r"(smem_u32(stored_cnt)) : "memory");
            if (++as == kAccStages) { as = 0; aphase ^= 1; }
        }
        ESR_PROF(if (R.prof && warp == 0 && lane == 1) {
            unsigned long long* o = R.prof + blockIdx.x * 16;
            o[7] = clock64() - e_t0; o[8] = e_wait; o[11] = e_n;
        })
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, kTmemCols);
    }
}

}  // namespace pair

int num_sms_cached();

int make_act_tensor_map(CUtensorMap* tm, const esr_tensor_nhwc& t, int B, int H, int W, int box_rows, int lean);

unsigned long long* rdb_prof_buffer();   // conv3x3_tc.cu (esr_debug_set_profile_buffer)

struct RdbOp {
    alignas(64) CUtensorMap tm0;
    alignas(64) CUtensorMap tm1;
    pair::RdbLaunch R;
};

int build_rdb_growth(const esr_rdb_growth_desc& d, RdbOp* op) {
    using namespace pair;
    ESR_CHECK_ARG(d.B > 0 && d.H > 0 && d.W > 0, "rdb_growth: bad geometry");
    ESR_CHECK_ARG(d.num_layers >= 1 && d.num_layers <= ESR_RDB_MAX_LAYERS, "rdb_growth: 1..4 layers");
    ESR_CHECK_ARG(d.src[0].ptr && d.out && d.flags, "rdb_growth: null buffer");
    ESR_CHECK_ARG(d.out_stride % 16 == 0 && (reinterpret_cast<uintptr_t>(d.out) & 31) == 0, "rdb_growth: output misaligned");
    ESR_CHECK_ARG(d.mode == 0 || (d.mode == 1 && d.mask && d.mask_stride % 8 == 0), "rdb_growth: bad mode / mask");
    ESR_CHECK_ARG(d.flags_use >= 0 && d.flags_use < 3 && d.flags_zero >= 0 && d.flags_zero < 3 && d.flags_use != d.flags_zero,
                  "rdb_growth: flags_use / flags_zero must be different thirds");
    RdbLaunch& R = op->R;
    memset(&R, 0, sizeof(R));
    R.B = d.B; R.H = d.H; R.W = d.W; R.nlayers = d.num_layers;
    uint32_t woff = 0;
    for (int l = 0; l < d.num_layers; ++l) {
        const esr_rdb_layer& s = d.layers[l];
        ESR_CHECK_ARG(s.num_kblocks > 0 && s.num_kblocks <= ESR_RDB_MAX_KBLOCKS && s.wpack && s.bias, "rdb_growth: bad layer %d", l);
        ESR_CHECK_ARG(s.w_tile_bytes % 32 == 0 && (reinterpret_cast<uintptr_t>(s.wpack) & 15) == 0, "rdb_growth: wpack misaligned");
        ESR_CHECK_ARG(s.out_choff % 16 == 0 && s.out_choff + 32 <= d.out_stride, "rdb_growth: bad out_choff");
        RdbLayerDev& D = R.layer[l];
        D.nkb = s.num_kblocks; D.out_choff = s.out_choff; D.w_smem_off = woff; D.w_half_bytes = s.w_tile_bytes / 2;
        D.wpack = static_cast<const uint8_t*>(s.wpack); D.bias = s.bias;
        for (int k = 0; k < s.num_kblocks; ++k) {
            const esr_kblock& kb = s.kblocks[k];
            ESR_CHECK_ARG((kb.src == 0 || (kb.src == 1 && d.src[1].ptr)) && kb.chan >= 0 && kb.chan % 8 == 0 &&
                          kb.chan + kKB <= d.src[kb.src].channels && (kb.dy_mask & 7) && (kb.slice_mask & 3) && kb.w_off % 512 == 0 &&
                          kb.w_off + kb.n_dy * 48u * (kb.half ? 32u : static_cast<uint32_t>(kRowBytes)) <= D.w_half_bytes &&
                          (!kb.half || (kb.src == 1 && (kb.dy_mask & 7) == 2 && (kb.slice_mask == 1 || kb.slice_mask == 2))),
                          "rdb_growth: bad K block %d of layer %d", k, l);
            D.kb[k] = kb;
        }
        woff += (D.w_half_bytes + 1023u) & ~1023u;
    }
    R.w_smem_bytes = woff;
    R.out = static_cast<__nv_bfloat16*>(d.out); R.out_stride = d.out_stride; R.mode = d.mode; R.slope = d.slope;
    R.mask = static_cast<const uint16_t*>(d.mask); R.mask_stride = d.mask_stride;
    R.tiles_x = ceil_div(d.W, kTileWOut);
    R.tiles_y = ceil_div(d.H, 2 * kBandRows);
    R.tiles_per_img = R.tiles_x * R.tiles_y;
    R.spatial_tiles = d.B * R.tiles_per_img;
    // images per chunk: 4 large images keep a chunk's dense block in L2 between its layers; with SMALL images a chunk must
    // still hold at least two rounds of tile pairs per layer (2 x 74 clusters), or the clusters run into the next layer's
    // items and stall on their neighbours: 16 x 32x32 training batches took 16 ms per step with chunks of 4, 11.8 with 8,
    // 8.9 with all 16 in one chunk (separate launches: 9.05)
    int ic = d.imgs_per_chunk;
    if (ic <= 0) {
        ic = 4;
        const int pairs_per_img = (R.tiles_per_img + 1) / 2;
        while (ic < d.B && ic * pairs_per_img < 2 * 74) ++ic;
        while (ic < d.B && d.B % ic != 0) ++ic;
    }
    if (ic > d.B || d.B % ic != 0) ic = d.B;
    R.tiles_c = ic * R.tiles_per_img;
    R.ppc = (R.tiles_c + 1) / 2;
    R.chunks = d.B / ic;
    R.total_items = R.chunks * R.nlayers * R.ppc;
    // ceil(2^32 / d); d == 1 would need 2^32 itself: stored as 0 = "identity" (fast_div)
    const auto magic = [](int dd) { return dd <= 1 ? 0u : static_cast<uint32_t>(((1ull << 32) + dd - 1) / dd); };
    ESR_CHECK_ARG(static_cast<long long>(R.total_items) * R.nlayers * R.ppc < (1ll << 32) &&
                  static_cast<long long>(R.spatial_tiles) * R.tiles_per_img < (1ll << 32), "rdb_growth: problem too large");
    R.m_per_chunk = magic(R.nlayers * R.ppc); R.m_ppc = magic(R.ppc); R.m_tpi = magic(R.tiles_per_img); R.m_tx = magic(R.tiles_x);
    const size_t third = static_cast<size_t>(ESR_RDB_MAX_LAYERS) * R.spatial_tiles;
    R.reverse = 0;                                        // set by the sequence builder (alternates launch by launch)
    R.prof = rdb_prof_buffer();
    R.flags = d.flags + third * d.flags_use;
    R.flags_zero = d.flags + third * d.flags_zero;
    constexpr int a_tile = (2 * kBandRows + 2) * kTileW * kRowBytes;
    const int ctrl = 256 + ESR_RDB_MAX_LAYERS * 32 * 4;
    const int room = kSmemMax - 1024 - ctrl - static_cast<int>(R.w_smem_bytes);
    int st = room / a_tile;
    R.nstages = st > kRdbStages ? kRdbStages : st;
    ESR_CHECK_ARG(R.nstages >= 2, "rdb_growth: weights (%u B per CTA) leave no room for the A-tile ring", R.w_smem_bytes);
    int rc = make_act_tensor_map(&op->tm0, d.src[0], d.B, d.H, d.W, 2 * kBandRows + 2, 0);
    if (rc != ESR_OK) return rc;
    if (d.src[1].ptr != nullptr) {
        int lean = -1;
        for (int l = 0; l < d.num_layers; ++l)
            for (int k = 0; k < d.layers[l].num_kblocks; ++k)
                if (d.layers[l].kblocks[k].src == 1) {
                    const int h = d.layers[l].kblocks[k].half ? 1 : 0;
                    ESR_CHECK_ARG(lean < 0 || lean == h, "rdb_growth: lean and full K blocks of source 1 cannot be mixed");
                    lean = h;
                }
        return make_act_tensor_map(&op->tm1, d.src[1], d.B, d.H, d.W, lean == 1 ? 2 * kBandRows : 2 * kBandRows + 2, lean == 1);
    }
    op->tm1 = op->tm0;
    return ESR_OK;
}

RdbOp* new_rdb_op(const esr_rdb_growth_desc& d, int* rc) {
    RdbOp* op = new (std::nothrow) RdbOp();
    if (op == nullptr) { set_error("out of host memory"); *rc = ESR_ERR_INVALID; return nullptr; }
    *rc = build_rdb_growth(d, op);
    if (*rc != ESR_OK) { delete op; return nullptr; }
    return op;
}
void delete_rdb_op(RdbOp* op) { delete op; }
void rdb_op_set_reverse(RdbOp* op, int reverse) { op->R.reverse = reverse; }

int launch_rdb_growth(const RdbOp& op, cudaStream_t stream, int use_pdl) {
    using namespace pair;
    ESR_ONCE_PER_DEVICE(
        ESR_CUDA(cudaFuncSetAttribute(conv3x3_rdb_growth_kernel<kEpiTrunk>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
        ESR_CUDA(cudaFuncSetAttribute(conv3x3_rdb_growth_kernel<kEpiMask>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    );
    const pair::RdbLaunch& R = op.R;
    constexpr int a_tile = (2 * kBandRows + 2) * kTileW * kRowBytes;
    int clusters = num_sms_cached() / 2;
    if (clusters > R.ppc) clusters = R.ppc;               // never more clusters than items of one (chunk, layer)
    const int smem = 1024 + static_cast<int>(R.w_smem_bytes) + R.nstages * a_tile + 256 + ESR_RDB_MAX_LAYERS * 32 * 4;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kRdbThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 2 : 1;
    const cudaError_t e = R.mode == 0 ? cudaLaunchKernelEx(&cfg, conv3x3_rdb_growth_kernel<kEpiTrunk>, op.tm0, op.tm1, R)
                                      : cudaLaunchKernelEx(&cfg, conv3x3_rdb_growth_kernel<kEpiMask>, op.tm0, op.tm1, R);
    if (e != cudaSuccess) { set_error("conv3x3_rdb_growth_kernel launch failed: %s", cudaGetErrorString(e)); return ESR_ERR_CUDA; }
    return check_launch("conv3x3_rdb_growth_kernel");
}

// Fills the pair-mode launch geometry; returns false if the weights leave no room for the A ring.
bool fill_launch_pair(ConvLaunch* L) {
    const esr_conv_desc& d = L->d;
    const int nb = d.cout_tile == 32 ? 2 : 1;
    const bool attach = d.num_kblocks >= 2 && d.kblocks[d.num_kblocks - 1].half != 0;
    const int a_tile = (nb * kBandRows + 2) * kTileW * kRowBytes + (attach ? nb * kBandRows * kTileW * 32 : 0);
    L->pair_nb = nb;
    L->tiles_y = ceil_div(d.H, nb * kBandRows);
    L->spatial_tiles = d.B * L->tiles_x * L->tiles_y;
    L->total_tiles = L->spatial_tiles * d.cout_tiles;
    L->w_smem_bytes = (d.w_tile_bytes / 2 + 1023u) & ~1023u;
    const int room = pair::kSmemMax - 1024 - pair::kCtrlBytes - static_cast<int>(L->w_smem_bytes);
    int st = room / a_tile;
    static const int cap = []() { const char* v = getenv("ESR_MAX_STAGES"); return v ? atoi(v) : 0; }();   // tuning aid
    if (cap >= 2 && st > cap) st = cap;
    L->nstages = st > pair::kMaxStages ? pair::kMaxStages : st;
    return L->nstages >= 2;
}

int launch_conv_tc2(const CUtensorMap& tm0, const CUtensorMap& tm1, const ConvLaunch& L, cudaStream_t stream, int use_pdl) {
    using namespace pair;
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const ConvLaunch);
    // indexed by classify_epilogue(): Generic, Trunk, Res, Act, (Nchw -> Generic), Mask, Dx0, Dx0Mask, Outer (cout tile 64 only)
    static const KernelFn kernels[2][9] = {
        {conv3x3_tc2_kernel<32, kEpiGeneric>, conv3x3_tc2_kernel<32, kEpiTrunk>, conv3x3_tc2_kernel<32, kEpiRes>,
         conv3x3_tc2_kernel<32, kEpiAct>, conv3x3_tc2_kernel<32, kEpiGeneric>, conv3x3_tc2_kernel<32, kEpiMask>, conv3x3_tc2_kernel<32, kEpiDx0>,
         conv3x3_tc2_kernel<32, kEpiDx0Mask>, conv3x3_tc2_kernel<32, kEpiGeneric>},
        {conv3x3_tc2_kernel<64, kEpiGeneric>, conv3x3_tc2_kernel<64, kEpiGeneric>, conv3x3_tc2_kernel<64, kEpiRes>,
         conv3x3_tc2_kernel<64, kEpiAct>, conv3x3_tc2_kernel<64, kEpiGeneric>, conv3x3_tc2_kernel<64, kEpiGeneric>,
         conv3x3_tc2_kernel<64, kEpiGeneric>, conv3x3_tc2_kernel<64, kEpiGeneric>, conv3x3_tc2_kernel<64, kEpiOuter>}};
    ESR_ONCE_PER_DEVICE(
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 9; ++b)
                ESR_CUDA(cudaFuncSetAttribute(kernels[a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    );
    const int nb = L.d.cout_tile == 32 ? 2 : 1;
    const bool attach = L.d.num_kblocks >= 2 && L.d.kblocks[L.d.num_kblocks - 1].half != 0;
    const int a_tile = (nb * kBandRows + 2) * kTileW * kRowBytes + (attach ? nb * kBandRows * kTileW * 32 : 0);
    const int num_pairs = (L.spatial_tiles + 1) / 2;
    int per_ct = (num_sms_cached() / 2) / L.d.cout_tiles;       // clusters per cout tile
    if (per_ct > num_pairs) per_ct = num_pairs;
    ESR_CHECK_ARG(per_ct >= 1, "too many cout tiles (%d) for pair mode", L.d.cout_tiles);
    const int grid = 2 * per_ct * L.d.cout_tiles;
    const int smem = 1024 + static_cast<int>(L.w_smem_bytes) + L.nstages * a_tile + kCtrlBytes;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 2 : 1;
    const int mode = classify_epilogue(L.d);
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernels[L.d.cout_tile == 64 ? 1 : 0][mode], tm0, tm1, L);
    if (e != cudaSuccess) { set_error("conv3x3_tc2_kernel launch failed: %s", cudaGetErrorString(e)); return ESR_ERR_CUDA; }
    return check_launch("conv3x3_tc2_kernel");
}

}  // namespace esr
