// Auxiliary conv kernels: weight packing, the SIMT cross-check kernel that consumes the very
// same descriptor / packed weights as the tcgen05 kernel, and the small-channel row expansion.
#include <cuda_bf16.h>

#include <cstring>

#include "conv3x3.cuh"

namespace esr {

// ------------------------------------------------------------------ SIMT check
// One thread per (pixel, 16-channel chunk).  Walks the K-block list exactly like the MMA
// issuer does, so any disagreement with the tcgen05 kernel is a TMA/UMMA problem and any
// disagreement with the oracle is a packing/epilogue problem.
__global__ void conv3x3_simt_kernel(const __grid_constant__ ConvLaunch L) {
    const esr_conv_desc& d = L.d;
    const int CT = d.cout_tile, N = 3 * CT;
    const int chunks = d.cout_tiles * CT / 16;
    const size_t total = static_cast<size_t>(d.B) * d.H * d.W * chunks;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int chunk = static_cast<int>(idx % chunks);
        size_t pix = idx / chunks;
        const int x = static_cast<int>(pix % d.W); pix /= d.W;
        const int y = static_cast<int>(pix % d.H);
        const int n = static_cast<int>(pix / d.H);
        const int co0 = chunk * 16, ct = co0 / CT, col0 = co0 % CT;
        const uint8_t* wt = reinterpret_cast<const uint8_t*>(d.wpack) + static_cast<size_t>(ct) * d.w_tile_bytes;
        const int slab_rows = d.pair ? N / 2 : N;            // pair layout: each CTA's half image holds N/2 rows per slab
        const size_t half_bytes = d.w_tile_bytes / 2;
        const bool f16 = (d.flags & ESR_CONV_F16) != 0;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
        for (int kb = 0; kb < d.num_kblocks; ++kb) {
            const esr_kblock& K = d.kblocks[kb];
            const uint16_t* src = reinterpret_cast<const uint16_t*>(d.src[K.src].ptr);
            const int sc = d.src[K.src].channels;
            int wi = 0;
            const int half_k0 = K.half ? ((K.slice_mask & 1) ? 0 : 16) : 0;     // lean block: its 16-channel slice
            for (int dy = 0; dy < 3; ++dy) {
                if (!((K.dy_mask >> dy) & 1)) continue;
                const uint8_t* wslab = wt + K.w_off + static_cast<size_t>(wi) * slab_rows * (K.half ? 32 : kRowBytes);
                ++wi;
                const int yy = y + dy - 1;
                if (yy < 0 || yy >= d.H) continue;
                for (int dx = 0; dx < 3; ++dx) {
                    const int xx = x + dx - 1;
                    if (xx < 0 || xx >= d.W) continue;
                    const uint16_t* a = src + ((static_cast<size_t>(n) * d.H + yy) * d.W + xx) * sc + K.chan;
                    for (int k = 0; k < kKB; ++k) {
                        if (!((K.slice_mask >> (k >> 4)) & 1)) continue;
                        const float av = operand_value(a[k], f16);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int nrow = dx * CT + col0 + i;
                            const uint16_t w = *reinterpret_cast<const uint16_t*>(
                                wslab + (nrow / slab_rows) * half_bytes +
                                (K.half ? sw32_offset(nrow % slab_rows, k - half_k0) : sw64_offset(nrow % slab_rows, k)));
                            v[i] = fmaf(av, operand_value(w, f16), v[i]);
                        }
                    }
                }
            }
        }
        EpiOperands ops;
        const int cabs = tile_channel(d, ct, col0);
        conv_epilogue_prefetch<kEpiGeneric>(d, ct, n, y, x, cabs, ops);
        conv_epilogue16<kEpiGeneric>(d, d.bias + co0, ct, n, y, x, cabs, v, ops);
    }
}

// --------------------------------------------------------------- weight packing
struct PackArgs {
    const float* wsrc;
    long long off, s_row, s_slot, s_ky, s_kx;
    const float* bias_src;
    int cout_tile, cout_tiles, pair, num_kblocks;
    uint32_t w_tile_bytes;
    esr_kblock kblocks[ESR_MAX_KBLOCKS];
    const esr_wrow* rows;    // device [cout_tiles*cout_tile]
    const esr_wslot* slots;  // device [num_kblocks*32]
    uint8_t* out;
    float* bias_out;
};

__device__ __forceinline__ void pack_one(const PackArgs& a, const unsigned bx, const unsigned nbx) {
    const int CT = a.cout_tile, N = 3 * CT;
    // one thread per 16-byte chunk of the image = 8 consecutive k of one row (both swizzles move whole 16-byte chunks):
    // index space [ct][kb][wi(3)][n][k / 8].  (One thread per element: 8x the index arithmetic and 2-byte stores, 1.8 ms
    // for the ~770 images a training step re-packs.)
    constexpr int kChunks = kKB / 8;
    const size_t per_kb = static_cast<size_t>(3) * N * kChunks;
    const size_t total = static_cast<size_t>(a.cout_tiles) * a.num_kblocks * per_kb;
    for (size_t idx = bx * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(nbx) * blockDim.x) {
        size_t r = idx;
        const int k0 = static_cast<int>(r % kChunks) * 8; r /= kChunks;
        const int n = static_cast<int>(r % N); r /= N;
        const int wi = static_cast<int>(r % 3); r /= 3;
        const int kb = static_cast<int>(r % a.num_kblocks);
        const int ct = static_cast<int>(r / a.num_kblocks);
        const esr_kblock& K = a.kblocks[kb];
        if (wi >= K.n_dy) continue;
        const int hk0 = K.half ? ((K.slice_mask & 1) ? 0 : 16) : 0;      // lean block: only its 16-channel slice is stored
        if (K.half && (k0 < hk0 || k0 >= hk0 + 16)) continue;
        int dy = -1;  // filter row of the wi-th set bit
        for (int b = 0, c = 0; b < 3; ++b)
            if ((K.dy_mask >> b) & 1) { if (c == wi) dy = b; ++c; }
        const int dx = n / CT, col = n % CT;
        const esr_wrow row = a.rows[ct * CT + col];
        __align__(16) uint16_t vals[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const esr_wslot slot = a.slots[kb * kKB + k0 + i];
            float w = 0.f;
            if (row.idx >= 0 && slot.idx >= 0) {
                int ky = dy;
                bool live = true;
                if (row.ky >= 0 || slot.ky >= 0) {      // pre-expanded over dy on one side: centre tap only
                    live = (dy == 1) && !(row.ky >= 0 && slot.ky >= 0);
                    ky = row.ky >= 0 ? row.ky : slot.ky;
                }
                if (live) w = a.wsrc[a.off + row.idx * a.s_row + slot.idx * a.s_slot + ky * a.s_ky + dx * a.s_kx];
            }
            vals[i] = operand_bits(w, slot.term);
        }
        const int slab_rows = a.pair ? N / 2 : N;            // pair layout: rows [r*N/2, (r+1)*N/2) live in CTA r's half image
        uint8_t* dst = a.out + static_cast<size_t>(ct) * a.w_tile_bytes + (n / slab_rows) * (a.w_tile_bytes / 2) + K.w_off;
        if (K.half) dst += static_cast<size_t>(wi) * slab_rows * 32 + sw32_offset(n % slab_rows, k0 - hk0);   // [rows x 16 ch], SWIZZLE_32B
        else dst += static_cast<size_t>(wi) * slab_rows * kRowBytes + sw64_offset(n % slab_rows, k0);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(vals);
    }
    const int nb = a.cout_tiles * CT;
    for (int i = bx * blockDim.x + threadIdx.x; i < nb; i += nbx * blockDim.x) {
        const esr_wrow row = a.rows[i];
        a.bias_out[i] = (a.bias_src != nullptr && row.idx >= 0 && row.ky < 0) ? a.bias_src[row.idx] : 0.f;
    }
}

__global__ void pack_weights_kernel(const __grid_constant__ PackArgs a) { pack_one(a, blockIdx.x, gridDim.x); }

// Every conv of a network in ONE launch (grid.y = table entry): a training step re-packs 351 forward and ~420 dgrad weight
// images after each update; as separate launches that was 12.5 ms of 16 us launches, host bound.
__global__ void pack_table_kernel(const PackArgs* __restrict__ table) { pack_one(table[blockIdx.y], blockIdx.x, gridDim.x); }

// Strided row copies (gathering the five weight tensors of a dense block into the [co_slot, ci, 3, 3] array its dgrad
// images are packed from), one launch for all segments.
__global__ void copy_segments_kernel(const esr_copy_seg* __restrict__ segs) {
    const esr_copy_seg sg = segs[blockIdx.y];
    const int total = sg.rows * sg.row_elems;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / sg.row_elems, c = i - r * sg.row_elems;
        sg.dst[static_cast<size_t>(r) * sg.dst_pitch + c] = sg.src[static_cast<size_t>(r) * sg.src_pitch + c];
    }
}

// ---------------------------------------------------------- small-channel expand
struct ExpandArgs {
    const float* src;
    int B, C, H, W, nslots;
    esr_xslot slots[64];
    __nv_bfloat16* dst;
};

// One thread per pixel: reads its <= 3 rows x C source values (coalesced along x in NCHW) and writes all
// slots of the pixel as 16-byte vectors (consecutive pixels are contiguous in NHWC: fully coalesced).
// (One thread per (pixel, 8-slot chunk) was tried: whole-pixel store runs, but the loads of a warp then spread over four
// channel planes - 0.19 instead of 0.13 ms per Z-optimisation iteration.)
__global__ void expand_rows_kernel(const __grid_constant__ ExpandArgs a) {
    const size_t total = static_cast<size_t>(a.B) * a.H * a.W;
    for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total;
         pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(pix % a.W);
        const int y = static_cast<int>((pix / a.W) % a.H);
        const int n = static_cast<int>(pix / (static_cast<size_t>(a.W) * a.H));
        __nv_bfloat16* dst = a.dst + pix * a.nslots;
        for (int s0 = 0; s0 < a.nslots; s0 += 8) {
            __align__(16) uint16_t o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const esr_xslot sl = a.slots[s0 + i];
                float val = 0.f;
                const int yy = y + sl.dy;
                if (sl.c >= 0 && yy >= 0 && yy < a.H)    // <= 18 distinct addresses per pixel, coalesced along x, L1 hits
                    val = __ldg(a.src + ((static_cast<size_t>(n) * a.C + sl.c) * a.H + yy) * a.W + x);
                o[i] = operand_bits(val, sl.term);
            }
            *reinterpret_cast<uint4*>(dst + s0) = *reinterpret_cast<const uint4*>(o);
        }
    }
}

int validate_conv_desc(const esr_conv_desc& d);
void fill_launch(ConvLaunch* L, const esr_conv_desc& d);

int launch_conv_simt(const ConvLaunch& L, cudaStream_t stream) {
    const size_t total = static_cast<size_t>(L.d.B) * L.d.H * L.d.W * (L.d.cout_tiles * L.d.cout_tile / 16);
    const int block = 128;
    const int grid = static_cast<int>(total / block + 1 < 65535 * 16 ? total / block + 1 : 65535 * 16);
    conv3x3_simt_kernel<<<grid, block, 0, stream>>>(L);
    return check_launch("conv3x3_simt_kernel");
}

}  // namespace esr

extern "C" int esr_conv3x3_simt(const esr_conv_desc* d, void* stream) {
    if (d == nullptr) { esr::set_error("null conv desc"); return ESR_ERR_INVALID; }
    int rc = esr::validate_conv_desc(*d);
    if (rc != ESR_OK) return rc;
    esr::ConvLaunch L;
    esr::fill_launch(&L, *d);
    return esr::launch_conv_simt(L, static_cast<cudaStream_t>(stream));
}

static bool cout_tile_ok(int32_t cout_tile, int32_t pair) {
    return pair ? (cout_tile == 32 || cout_tile == 64) : (cout_tile == 16 || cout_tile == 32);
}

extern "C" int64_t esr_pack_layout(int32_t cout_tile, int32_t cout_tiles, int32_t pair, int32_t num_kblocks,
                                   esr_kblock* kblocks, uint32_t* w_tile_bytes) {
    if (!cout_tile_ok(cout_tile, pair) || cout_tiles <= 0 || num_kblocks <= 0 ||
        num_kblocks > ESR_MAX_KBLOCKS || kblocks == nullptr) {
        esr::set_error("esr_pack_layout: bad arguments");
        return ESR_ERR_INVALID;
    }
    uint32_t off = 0;
    for (int i = 0; i < num_kblocks; ++i) {
        kblocks[i].n_dy = static_cast<uint8_t>(__builtin_popcount(kblocks[i].dy_mask & 7));
        kblocks[i].w_off = off;
        if (kblocks[i].half && !(pair && (kblocks[i].dy_mask & 7) == 2 && (kblocks[i].slice_mask == 1 || kblocks[i].slice_mask == 2))) {
            esr::set_error("esr_pack_layout: K block %d: `half` needs pair mode, dy_mask 0b010 and a single slice", i);
            return ESR_ERR_INVALID;
        }
        off += kblocks[i].n_dy * 3u * cout_tile * (kblocks[i].half ? 32u : static_cast<uint32_t>(esr::kRowBytes)) / (pair ? 2u : 1u);
        off = (off + 511u) & ~511u;                              // slabs of lean blocks are 1.5 KiB: keep every block 512-byte aligned
    }
    if (pair) off *= 2;                                      // two half images per cout tile
    if (w_tile_bytes) *w_tile_bytes = off;
    return static_cast<int64_t>(off) * cout_tiles;
}

extern "C" int esr_pack_conv_weights(const float* wsrc, int64_t off, int64_t s_row, int64_t s_slot, int64_t s_ky,
                                     int64_t s_kx, const float* bias_src, int32_t cout_tile, int32_t cout_tiles,
                                     int32_t pair, int32_t num_kblocks, const esr_kblock* kblocks, uint32_t w_tile_bytes,
                                     const esr_wrow* rows_dev, const esr_wslot* slots_dev, void* wpack_out,
                                     float* bias_out, void* stream) {
    ESR_CHECK_ARG(wsrc && kblocks && rows_dev && slots_dev && wpack_out && bias_out, "esr_pack_conv_weights: null argument");
    ESR_CHECK_ARG(cout_tile_ok(cout_tile, pair) && cout_tiles > 0 && num_kblocks > 0 &&
                  num_kblocks <= ESR_MAX_KBLOCKS, "esr_pack_conv_weights: bad sizes");
    esr::PackArgs a;
    a.wsrc = wsrc; a.off = off; a.s_row = s_row; a.s_slot = s_slot; a.s_ky = s_ky; a.s_kx = s_kx;
    a.bias_src = bias_src; a.cout_tile = cout_tile; a.cout_tiles = cout_tiles; a.pair = pair ? 1 : 0; a.num_kblocks = num_kblocks;
    a.w_tile_bytes = w_tile_bytes;
    for (int i = 0; i < num_kblocks; ++i) a.kblocks[i] = kblocks[i];
    a.rows = rows_dev; a.slots = slots_dev; a.out = static_cast<uint8_t*>(wpack_out); a.bias_out = bias_out;
    const size_t total = static_cast<size_t>(cout_tiles) * num_kblocks * 3 * 3 * cout_tile * esr::kKB;
    const int block = 256;
    const int grid = static_cast<int>((total + block - 1) / block);
    esr::pack_weights_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return esr::check_launch("pack_weights_kernel");
}

static int fill_pack_args(esr::PackArgs& a, const float* wsrc, int64_t off, int64_t s_row, int64_t s_slot, int64_t s_ky, int64_t s_kx,
                          const float* bias_src, int32_t cout_tile, int32_t cout_tiles, int32_t pair, int32_t num_kblocks,
                          const esr_kblock* kblocks, uint32_t w_tile_bytes, const esr_wrow* rows_dev, const esr_wslot* slots_dev,
                          void* wpack_out, float* bias_out) {
    ESR_CHECK_ARG(wsrc && kblocks && rows_dev && slots_dev && wpack_out && bias_out, "pack: null argument");
    ESR_CHECK_ARG(cout_tile_ok(cout_tile, pair) && cout_tiles > 0 && num_kblocks > 0 && num_kblocks <= ESR_MAX_KBLOCKS, "pack: bad sizes");
    a.wsrc = wsrc; a.off = off; a.s_row = s_row; a.s_slot = s_slot; a.s_ky = s_ky; a.s_kx = s_kx;
    a.bias_src = bias_src; a.cout_tile = cout_tile; a.cout_tiles = cout_tiles; a.pair = pair ? 1 : 0; a.num_kblocks = num_kblocks;
    a.w_tile_bytes = w_tile_bytes;
    for (int i = 0; i < num_kblocks; ++i) a.kblocks[i] = kblocks[i];
    a.rows = rows_dev; a.slots = slots_dev; a.out = static_cast<uint8_t*>(wpack_out); a.bias_out = bias_out;
    return ESR_OK;
}

extern "C" int32_t esr_pack_entry_bytes() { return static_cast<int32_t>(sizeof(esr::PackArgs)); }

extern "C" int esr_pack_entry_fill(void* entry_host, const float* wsrc, int64_t off, int64_t s_row, int64_t s_slot, int64_t s_ky,
                                   int64_t s_kx, const float* bias_src, int32_t cout_tile, int32_t cout_tiles, int32_t pair,
                                   int32_t num_kblocks, const esr_kblock* kblocks, uint32_t w_tile_bytes, const esr_wrow* rows_dev,
                                   const esr_wslot* slots_dev, void* wpack_out, float* bias_out) {
    ESR_CHECK_ARG(entry_host != nullptr, "esr_pack_entry_fill: null entry");
    esr::PackArgs a;
    memset(&a, 0, sizeof(a));
    int rc = fill_pack_args(a, wsrc, off, s_row, s_slot, s_ky, s_kx, bias_src, cout_tile, cout_tiles, pair, num_kblocks, kblocks,
                            w_tile_bytes, rows_dev, slots_dev, wpack_out, bias_out);
    if (rc != ESR_OK) return rc;
    memcpy(entry_host, &a, sizeof(a));
    return ESR_OK;
}

extern "C" int esr_pack_table_run(const void* table_device, int32_t n, void* stream) {
    ESR_CHECK_ARG(table_device != nullptr && n > 0, "esr_pack_table_run: bad arguments");
    esr::pack_table_kernel<<<dim3(24, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const esr::PackArgs*>(table_device));
    return esr::check_launch("pack_table_kernel");
}

extern "C" int esr_copy_segments(const esr_copy_seg* segs_device, int32_t n, void* stream) {
    ESR_CHECK_ARG(segs_device != nullptr && n > 0, "esr_copy_segments: bad arguments");
    esr::copy_segments_kernel<<<dim3(8, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(segs_device);
    return esr::check_launch("copy_segments_kernel");
}

extern "C" int esr_expand_rows(const float* src_nchw, int32_t B, int32_t C, int32_t H, int32_t W,
                               const esr_xslot* slots, int32_t nslots, void* dst_nhwc, void* stream) {
    ESR_CHECK_ARG(src_nchw && slots && dst_nhwc, "esr_expand_rows: null argument");
    ESR_CHECK_ARG(nslots > 0 && nslots <= 64 && nslots % 8 == 0, "esr_expand_rows: nslots must be a multiple of 8, <= 64");
    esr::ExpandArgs a;
    a.src = src_nchw; a.B = B; a.C = C; a.H = H; a.W = W; a.nslots = nslots;
    for (int i = 0; i < nslots; ++i) {
        ESR_CHECK_ARG(slots[i].c < C && slots[i].dy >= -1 && slots[i].dy <= 1, "esr_expand_rows: bad slot %d", i);
        a.slots[i] = slots[i];
    }
    a.dst = static_cast<__nv_bfloat16*>(dst_nhwc);
    ESR_CHECK_ARG(C <= 8, "esr_expand_rows: at most 8 source channels");
    const size_t total = static_cast<size_t>(B) * H * W;
    const int block = 128;
    const size_t want = (total + block - 1) / block;
    const int grid = static_cast<int>(want < 148 * 64 ? want : 148 * 64);
    esr::expand_rows_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return esr::check_launch("expand_rows_kernel");
}
