/*
 * esr_b200.h - C ABI of libesr_b200.so: hand-written sm_100a kernels for the
 * Explorable-SR hot path (RRDBNet(+Z) -> CEM forward, and its data-gradient
 * backward used by Z_optimization).
 *
 * The reference (YuvalBahat/Explorable-Super-Resolution_old) is pure Python and
 * has no FFI; its boundary for this path is the nn.Module API, whose aten calls
 * each entry point below replaces (file:line under /root/reference/codes):
 *
 *   esr_conv3x3_*         nn.Conv2d(k=3,p=1)(+LeakyReLU 0.2, *0.2 + residual, nearest x2)
 *                         models/modules/block.py:129-155, :230-235, :262-270, :85-97, :294-301
 *                         models/modules/architecture.py:151-175
 *   esr_pack_conv_weights the OIHW fp32 state_dict tensors (models/base_model.py:100-144 contract)
 *   esr_g_input_prep      CEM_PyTorch.forward pre-pad branch + RRDBNet.forward head
 *                         CEM/CEMnet.py:170-181, models/modules/architecture.py:152-160
 *   esr_cem_downscale     CEM_PyTorch.DownscaleOP       CEM/CEMnet.py:152,157-162
 *   esr_cem_inv_hth       Conv_LR_with_Inv_hTh_OP       CEM/CEMnet.py:149-151
 *   esr_cem_upscale       Upscale_OP                    CEM/CEMnet.py:153-159
 *   esr_cem_project       CEM_PyTorch.forward :183-190  (out = y + Up(K*(x - Down y)), crop)
 *   esr_cem_project_bwd   its adjoint w.r.t. y (autograd in the reference, Z_optimization.py:633)
 *   esr_cem2d_*           the same five operators for non-default kernels (imresize_CEM.py:22-42),
 *                         whose filters are not rank-1
 *
 * Conventions: every function returns 0 on success or a negative esr_status and
 * records a message readable through esr_last_error() (thread-local).  No
 * exceptions cross the boundary.  No device memory is allocated inside: callers
 * pass device pointers, sizes and a cudaStream_t (as void*).  Kernels are
 * enqueued on that stream and the call returns without synchronising.
 * Layouts: "NCHW f32" = contiguous float [B,C,H,W]; "NHWC bf16" = contiguous
 * __nv_bfloat16 [B,H,W,C].
 */
#ifndef ESR_B200_H
#define ESR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum esr_status {
    ESR_OK = 0,
    ESR_ERR_INVALID = -1,   /* bad argument */
    ESR_ERR_CUDA = -2,      /* CUDA runtime / driver error */
    ESR_ERR_UNSUPPORTED = -3 /* device is not sm_100 or feature not built */
} esr_status;

const char* esr_last_error(void);
int esr_abi_version(void);
/* 0 if `device` can run the kernels (compute capability 10.x). */
int esr_device_check(int device);

/* ------------------------------------------------------------------ conv3x3 */

#define ESR_MAX_KBLOCKS 24
#define ESR_MAX_COUT_TILES 8
#define ESR_KBLOCK_CH 32 /* channels per K block (64-byte rows, SWIZZLE_64B) */

/* One K block = 32 consecutive channels of one NHWC bf16 source tensor. */
typedef struct esr_kblock {
    int32_t src;        /* index into esr_conv_desc.src[] */
    int32_t chan;       /* first channel (multiple of 8) */
    uint32_t w_off;     /* byte offset of the block's weights inside one cout tile image */
    uint8_t dy_mask;    /* bit d: filter row d (0..2, i.e. dy=d-1) is applied; 0b010 for
                           sources that were pre-expanded over dy */
    uint8_t slice_mask; /* bit s: channels [16s,16s+16) of the block are used */
    uint8_t n_dy;       /* popcount(dy_mask) */
    uint8_t half;       /* != 0 (pair launches only): lean block for the row-expanded latent - centre tap, one 16-channel
                           slice (dy_mask == 0b010, slice_mask == 0b01 or 0b10).  Its activation tile is loaded without
                           halo rows as 16-channel / 32-byte rows (SWIZZLE_32B): 8 KiB instead of 20 KiB per 8x32-pixel
                           tile, and its weight slabs are packed as [rows x 16 ch] in the same swizzle.  Every block of
                           source 1 of a launch must then be lean. */
} esr_kblock;

typedef struct esr_tensor_nhwc {
    const void* ptr;  /* bf16 [B,H,W,channels] */
    int32_t channels; /* multiple of 8 */
} esr_tensor_nhwc;

enum {
    ESR_EPI_LRELU = 1u << 0,   /* v = v > 0 ? v : slope*v              (block.py:10-23) */
    ESR_EPI_RES1 = 1u << 1,    /* v = alpha*v + res1                   (block.py:235)    */
    ESR_EPI_RES2 = 1u << 2,    /* v = beta*v + res2                    (block.py:270)    */
    ESR_EPI_ACCUM = 1u << 3,   /* v += out_f32 (read-modify-write; dgrad fan-in)         */
    ESR_EPI_MASK = 1u << 4,    /* bf16 output *= (mask > 0 ? 1 : slope) (LeakyReLU')     */
    ESR_EPI_F32_BLOCKED = 1u << 5, /* out_f32 / res1 / res2 use the blocked layout [B, C/8, H, W, 8] (coalesced
                                      32-byte accesses across the pixels of a warp) instead of NHWC; *_stride is C */
    ESR_EPI_RES1_HILO = 1u << 8, /* with ESR_EPI_RES1: the residual is the bf16 pair res1_hi + res1_lo, not `res1` */
    ESR_CONV_F16 = 1u << 6,    /* MMA operands of this launch (every K block and the packed weights, term 2) are IEEE
                                  fp16 instead of bf16: the convs outside the residual-scaled trunk, whose operand
                                  rounding dominates the output error, 11 significant bits instead of 8 */
    ESR_EPI_OUT_F16 = 1u << 7, /* out_bf16 receives fp16 (round to nearest, saturating) instead of bf16 */
    ESR_EPI_WIDE_OK = 1u << 16 /* internal: set by the library when 32-byte accesses are legal */
};

typedef struct esr_conv_desc {
    int32_t B, H, W;          /* conv resolution (input == output, stride 1, zero pad 1) */
    esr_tensor_nhwc src[2];
    int32_t cout_tile;        /* output channels per tile: 16 or 32; 64 only with pair != 0 */
    int32_t cout_tiles;       /* ceil(Cout / cout_tile) */
    int32_t num_kblocks;
    esr_kblock kblocks[ESR_MAX_KBLOCKS];
    const void* wpack;        /* esr_pack_conv_weights output */
    uint32_t w_tile_bytes;    /* bytes per cout tile in wpack */
    const float* bias;        /* [cout_tiles*cout_tile] f32 */
    uint32_t flags;
    float slope, alpha, beta;
    const float* res1;        /* NHWC f32 [B,H,W,res1_stride], channel offset res1_choff */
    int32_t res1_stride, res1_choff;
    const float* res2;
    int32_t res2_stride, res2_choff;
    void* out_bf16;           /* NHWC bf16 [B,up*H,up*W,out_bf16_stride] or NULL */
    int32_t out_bf16_stride, out_bf16_choff;
    int32_t out_bf16_lo_choff; /* >=0: also store the bf16 residue v-bf16(v) there */
    int32_t up;               /* 1, or 2 = write every output pixel as a 2x2 block (nn.Upsample nearest) */
    float out_bf16_scale;     /* bf16 output = bf16(out_bf16_scale * v) */
    float* out_f32;           /* NHWC f32 [B,H,W,out_f32_stride] or NULL */
    int32_t out_f32_stride, out_f32_choff;
    float* out_nchw;          /* NCHW f32 [B,cout_real,H,W] or NULL */
    int32_t cout_real;
    const void* mask;         /* NHWC bf16, same indexing as out_bf16 (ESR_EPI_MASK) */
    int32_t mask_stride, mask_choff;
    /* Per-cout-tile routing (used by the dgrad launches, where one launch produces the gradients of
     * several channel groups plus the latent rows).  Output channel of (tile t, column j) is
     * tile_choff[t] + j when tile_choff[t] >= 0, else t*cout_tile + j.  A set bit t in a no_* mask
     * switches that feature off for tile t. */
    int16_t tile_choff[ESR_MAX_COUT_TILES];
    uint16_t no_accum_tiles, no_bf16_tiles, no_res_tiles;
    uint16_t pair;            /* != 0: wpack is in the pair layout (esr_pack_layout pair=1) and the launch runs on
                                 2-CTA clusters with tcgen05.mma.cta_group::2 (M = 256: one 4x32-pixel band per CTA) */
    float gamma;              /* res1 multiplier: v = alpha*v + gamma*res1 */
    /* Residual trunk carried as a bf16 pair instead of fp32 (halves the trunk's HBM traffic inside a block chain):
     * with ESR_EPI_RES1_HILO, res1 = bf16 res1_hi[..] + bf16 res1_lo[..] (NHWC, e.g. hi = channels 0..63 of the
     * dense-block buffer the conv reads anyway, lo = the residue stored by the previous block); out_lo != NULL
     * stores bf16(v - bf16(v)) next to the bf16 output (x = hi + lo keeps ~17 significant bits). */
    const void* res1_hi;
    int32_t res1_hi_stride, res1_hi_choff;
    const void* res1_lo;
    int32_t res1_lo_stride, res1_lo_choff;
    void* out_lo;
    int32_t out_lo_stride, out_lo_choff;
} esr_conv_desc;

/* tcgen05/TMEM/TMA implicit-GEMM kernel (the product path). */
int esr_conv3x3_tc(const esr_conv_desc* d, void* stream);
/* Same contract on CUDA cores, one thread per output; used by tests to isolate
 * tensor-core/TMA faults from packing/epilogue faults.  Never used for timing. */
int esr_conv3x3_simt(const esr_conv_desc* d, void* stream);

/* Fused growth convs of a dense block (block.py:230-235: conv 0..3 of ResidualDenseBlock_5C; also the mirrored dgrad
 * launches 5..2 of the backward): ONE persistent launch runs up to four 3x3 convs that each read channel windows of
 * the same NHWC bf16 buffer (+ the latent rows) and write 32 new channels into it.  Layer l+1 of a tile starts as
 * soon as layer l of its 3x3 tile neighbourhood is stored (per-tile counters in `flags`), so there is no launch
 * boundary, pipeline drain or grid barrier between the layers; the batch is walked in chunks of `imgs_per_chunk`
 * images so that a chunk's dense block is still in L2 when the next layer reads it.  CTA pairs (cta_group::2),
 * cout tile 32, weights of all layers resident in shared memory. */
#define ESR_RDB_MAX_LAYERS 4
#define ESR_RDB_MAX_KBLOCKS 8
typedef struct esr_rdb_layer {
    int32_t num_kblocks;
    esr_kblock kblocks[ESR_RDB_MAX_KBLOCKS]; /* pair layout (esr_pack_layout pair=1), cout_tile 32, one cout tile */
    const void* wpack;
    uint32_t w_tile_bytes;
    const float* bias;                       /* 32 floats */
    int32_t out_choff;                       /* channel slice [out_choff, out_choff+32) of `out` written by this layer */
} esr_rdb_layer;
typedef struct esr_rdb_growth_desc {
    int32_t B, H, W;
    esr_tensor_nhwc src[2];  /* src[0]: the dense-block buffer the K blocks read; src[1]: latent rows or NULL */
    int32_t num_layers;      /* 1..4; layer l may read what layers < l wrote */
    esr_rdb_layer layers[ESR_RDB_MAX_LAYERS];
    void* out;               /* NHWC bf16 [B,H,W,out_stride]; normally == src[0].ptr */
    int32_t out_stride;
    int32_t mode;            /* 0: out = bf16(LeakyReLU(acc + bias))   1: out = bf16((acc + bias) * LeakyReLU'(mask)) */
    float slope;
    const void* mask;        /* mode 1: NHWC bf16/fp16 activations, sign read at the output's own channel offset */
    int32_t mask_stride;
    int32_t imgs_per_chunk;  /* <= 0: library default (4, or B if it does not divide B) */
    uint32_t* flags;         /* device scratch of esr_rdb_growth_flag_words() uint32, zero-filled ONCE by the caller */
    int32_t flags_use;       /* 0..2: third of `flags` this launch counts in (must be all zero when it starts) */
    int32_t flags_zero;      /* 0..2, != flags_use: third this launch clears for a later launch */
} esr_rdb_growth_desc;
int64_t esr_rdb_growth_flag_words(int32_t B, int32_t H, int32_t W);
int esr_rdb_growth_tc(const esr_rdb_growth_desc* d, void* stream);

/* Weight packing tables.  The kernels compute
 *   out[row] = sum over taps (ky,kx) and channel slots of  A[slot] * Wl[row, slot, ky, kx]
 * and the packer gathers the logical weights from an f32 device tensor:
 *   Wl[row, slot, ky, kx] = wsrc[off + row.idx*s_row + slot.idx*s_slot + ky*s_ky + kx*s_kx]
 * (element strides, possibly negative: dgrad is the transposed + flipped view of the
 * same OIHW tensor).  A row or a slot with ky >= 0 belongs to a tensor that was
 * pre-expanded over the filter rows: it contributes at the centre tap only, with
 * filter row ky.  idx < 0 = zero row / slot. */
typedef struct esr_wrow {
    int16_t idx;
    int8_t ky;   /* -1: follows the tap loop */
    int8_t reserved;
} esr_wrow;
typedef struct esr_wslot {
    int16_t idx;
    int8_t ky;   /* -1: follows the tap loop */
    int8_t term; /* 0: bf16(w)   1: bf16(w - bf16(w))   2: fp16(w) */
} esr_wslot;

/* Fills w_off / n_dy of every K block and returns the total packed size in bytes
 * (cout_tiles * *w_tile_bytes), or a negative esr_status.
 * pair = 0: one image per cout tile, slabs of N = 3*cout_tile rows (dx-major) x 32 channels.
 * pair = 1: the image of a cout tile is two half images (w_tile_bytes/2 each), one per CTA of a pair: CTA r holds
 *           rows [r*N/2, (r+1)*N/2) of every slab (the B operand split of cta_group::2); w_off is the offset
 *           inside a half image.  cout_tile may then also be 64. */
int64_t esr_pack_layout(int32_t cout_tile, int32_t cout_tiles, int32_t pair, int32_t num_kblocks, esr_kblock* kblocks,
                        uint32_t* w_tile_bytes);
/* rows_dev: DEVICE array [cout_tiles*cout_tile]; slots_dev: DEVICE array [num_kblocks*32].
 * bias_src (f32 device, indexed by row.idx) may be NULL.  bias_out: [cout_tiles*cout_tile]. */
int esr_pack_conv_weights(const float* wsrc, int64_t off, int64_t s_row, int64_t s_slot, int64_t s_ky, int64_t s_kx,
                          const float* bias_src, int32_t cout_tile, int32_t cout_tiles, int32_t pair, int32_t num_kblocks,
                          const esr_kblock* kblocks, uint32_t w_tile_bytes, const esr_wrow* rows_dev,
                          const esr_wslot* slots_dev, void* wpack_out, float* bias_out, void* stream);

/* The same for a whole network in one launch: fill one table entry per conv on the host (esr_pack_entry_bytes() bytes each,
 * same arguments as esr_pack_conv_weights), copy the table to the device, run it after every weight update. */
int32_t esr_pack_entry_bytes(void);
int esr_pack_entry_fill(void* entry_host, const float* wsrc, int64_t off, int64_t s_row, int64_t s_slot, int64_t s_ky,
                        int64_t s_kx, const float* bias_src, int32_t cout_tile, int32_t cout_tiles, int32_t pair,
                        int32_t num_kblocks, const esr_kblock* kblocks, uint32_t w_tile_bytes, const esr_wrow* rows_dev,
                        const esr_wslot* slots_dev, void* wpack_out, float* bias_out);
int esr_pack_table_run(const void* table_device, int32_t n, void* stream);
/* n strided row copies in one launch: dst[r*dst_pitch + c] = src[r*src_pitch + c], r < rows, c < row_elems (fp32) */
typedef struct esr_copy_seg {
    const float* src;
    float* dst;
    int32_t rows, row_elems, src_pitch, dst_pitch;
} esr_copy_seg;
int esr_copy_segments(const esr_copy_seg* segs_device, int32_t n, void* stream);

/* ------------------------------------------------------- small-channel sources */

typedef struct esr_xslot {
    int8_t c;    /* source channel, -1 = zero */
    int8_t dy;   /* row offset -1..1 */
    int8_t term; /* 0: bf16(v)  1: bf16(v - bf16(v))  2: fp16(v) */
    int8_t reserved;
} esr_xslot;

/* dst[b,y,x,s] (NHWC bf16, `nslots` channels) = term(src[b, c_s, y+dy_s, x]) with zeros
 * outside the image.  `slots` is a HOST array [nslots], nslots <= 64. */
int esr_expand_rows(const float* src_nchw, int32_t B, int32_t C, int32_t H, int32_t W, const esr_xslot* slots,
                    int32_t nslots, void* dst_nhwc, void* stream);

/* Adjoint of esr_expand_rows: g_src[b,c,y,x] = sum_{s: c_s == c} g_e[b, y-dy_s, x, choff+s]
 * (g_e: f32 with `stride` channels, NHWC or, if blocked != 0, [B, stride/8, H, W, 8]).  hi and lo
 * slots of the same value both contribute; n_acc accumulators at choff + j*acc_stride are summed. */
int esr_expand_rows_bwd(const float* g_e, int32_t stride, int32_t choff, int32_t blocked, int32_t n_acc,
                        int32_t acc_stride, int32_t B, int32_t C, int32_t H, int32_t W, const esr_xslot* slots,
                        int32_t nslots, float* g_src_nchw, void* stream);

/* Elementwise gradient plumbing between dgrad launches (nearest-upsample adjoint, residual sums,
 * LeakyReLU', bf16 hi/lo emission), 64-channel tensors, f32 blocked [B,8,H,W,8]:
 *   v[b,y,x,c] = sum_{a,b' < pool} src[b, pool*y+a, pool*x+b', c]  (+ add[b,y,x,c] if add != NULL)
 *   out_f32 = v (optional);   w = scale * v * (mask ? (mask > 0 ? 1 : slope) : 1)
 *   out_bf16[..., hi_choff + c] = bf16(w);  out_bf16[..., lo_choff + c] = bf16(w - bf16(w)) if lo_choff >= 0
 * src is [B, src_stride/8, pool*H, pool*W, 8] at channel offset src_choff; mask is NHWC bf16
 * [B, mask_sub*H, mask_sub*W, mask_stride] sampled at (mask_sub*y, mask_sub*x). */
int esr_grad_combine(const float* src, int32_t src_stride, int32_t src_choff, int32_t pool, const float* add,
                     int32_t add_stride, int32_t add_choff, int32_t B, int32_t H, int32_t W, float* out_f32,
                     int32_t out_stride, int32_t out_choff, const void* mask, int32_t mask_stride, int32_t mask_choff,
                     int32_t mask_sub, float slope, float scale, void* out_bf16, int32_t bf16_stride,
                     int32_t hi_choff, int32_t lo_choff, void* stream);

/* CEM_PyTorch pre-pad + RRDBNet head (CEMnet.py:170-181, architecture.py:152-160).
 * model_input: NCHW f32 [B, sf*sf*nz+3, h, w] = cat(Z.view(B,sf*sf*nz,h,w), LR)  (raw .view packing,
 * SRRaGAN_model.py:249-255).  Writes (any may be NULL), with hp=h+2m, wp=w+2m:
 *   lr_pad [B,3,hp,wp]          replication-padded LR image (the CEM's x)
 *   fea_in [B,nz+3,hp,wp]       cat(z_lr, lr_pad): input of the first conv
 *   z_hr   [B,nz,sf*hp,sf*wp]   Z replication-padded by sf*m
 *   z_lr   [B,nz,hp,wp]         bilinear(z_hr, 1/sf, align_corners=False) == mean of the centre
 *                               2x2 (even sf) / the centre pixel (odd sf) of each sf x sf block */
int esr_g_input_prep(const float* model_input, int32_t B, int32_t nz, int32_t h, int32_t w, int32_t m, int32_t sf,
                     float* lr_pad, float* fea_in, float* z_hr, float* z_lr, void* stream);
/* Same padding, emitted in the reference's packed layout [B, sf*sf*nz+3, hp, wp] (for wrapped
 * generators that are not this library's RRDBNet). */
int esr_cem_pad_input(const float* model_input, int32_t B, int32_t nz, int32_t h, int32_t w, int32_t m, int32_t sf,
                      float* packed_out, void* stream);
/* Adjoint of the above w.r.t. the Z channels: g_model_input[:, :16*nz] (+)= ... */
int esr_g_input_prep_bwd(const float* g_z_hr, const float* g_z_lr, int32_t B, int32_t nz, int32_t h, int32_t w,
                         int32_t m, int32_t sf, float* g_model_input, void* stream);

/* ---------------------------------------------------------------------- CEM */

#define ESR_CEM_MAX_TAPS 64
typedef struct esr_cem_filters {
    int32_t sf;       /* integer scale factor */
    int32_t pre;      /* sampling phase (imresize_CEM.py:73-86) */
    int32_t n_ds;     /* length of the 1-D factor of ds_kernel (17 for bicubic x4) */
    int32_t n_inv;    /* length of the 1-D factor of inv_hTh (27) */
    float ds[ESR_CEM_MAX_TAPS];  /* ds_kernel == outer(ds, ds) */
    float inv[ESR_CEM_MAX_TAPS]; /* inv_hTh   == outer(inv, inv) */
} esr_cem_filters;

int esr_cem_downscale(const esr_cem_filters* f, const float* y, int32_t B, int32_t C, int32_t H, int32_t W,
                      float* out, void* stream);
int esr_cem_inv_hth(const esr_cem_filters* f, const float* x, int32_t B, int32_t C, int32_t h, int32_t w, float* out,
                    void* stream);
int esr_cem_upscale(const esr_cem_filters* f, const float* x, int32_t B, int32_t C, int32_t h, int32_t w, float* out,
                    void* stream);
/* out[B,C,H-2*crop,W-2*crop] = crop(y + Up(K * (x - Down(y)))).  y: [B,C,H,W], x: [B,C,H/sf,W/sf].
 * workspace: 2*B*C*(H/sf)*(W/sf) floats. */
int esr_cem_project(const esr_cem_filters* f, const float* y, const float* x, int32_t B, int32_t C, int32_t H,
                    int32_t W, int32_t crop, float* out, float* workspace, void* stream);
/* The same projection as ONE cooperative launch that keeps y on chip between Down and Up (csrc/cem_fused.cu): y is read
 * from HBM once.  x4 bicubic only, H % 16 == 0, W % 32 == 0, 1024 <= W <= 2048, H / 16 <= SM count; other shapes return
 * ESR_ERR_UNSUPPORTED.  Slower than the two-launch path at BASELINE config 4 (lock-step rounds), hence not the default of
 * esr_cem_project (ESR_CEM_FUSED=1 routes it here).  Same workspace. */
int esr_cem_project_fused(const esr_cem_filters* f, const float* y, const float* x, int32_t B, int32_t C, int32_t H,
                          int32_t W, int32_t crop, float* out, float* workspace, void* stream);
/* g_y[B,C,H,W] = pad(g_out) - Down^T(K^T(Up^T(pad(g_out)))), exact adjoint including the
 * replicate-padding folds.  workspace: B*C*(H*W + H*(W/sf) + 2*(H/sf)*(W/sf)) floats. */
int esr_cem_project_bwd(const esr_cem_filters* f, const float* g_out, int32_t B, int32_t C, int32_t H, int32_t W,
                        int32_t crop, float* g_y, float* workspace, void* stream);

/* ------------------------------------------------- CEM, general (non-separable) filters
 * The reference accepts any square downscaling kernel (imresize_CEM.py:22-32: an estimated /
 * user-supplied ndarray, re-centred by Center_Mass) and `blurry_cubic_<sigma>` (:37-41); for those
 * ds_kernel and / or inv_hTh (whose Fourier-domain magnitude clamp, CEMnet.py:112, is not
 * separable) are not rank-1 and the operators run as direct 2-D stencils.  Same semantics,
 * workspace sizes and argument order as the esr_cem_* entry points above.
 *   ds  : DEVICE pointer, n_ds*n_ds floats row-major  == CEMnet.ds_kernel   (CEMnet.py:22)
 *   inv : DEVICE pointer, n_inv*n_inv floats row-major == CEMnet.inv_hTh    (CEMnet.py:105-126)
 * n_ds, n_inv odd, <= ESR_CEM2D_MAX_SIDE. */
#define ESR_CEM2D_MAX_SIDE 63
typedef struct esr_cem_filters2d {
    int32_t sf, pre, n_ds, n_inv;
    const float* ds;
    const float* inv;
} esr_cem_filters2d;

int esr_cem2d_downscale(const esr_cem_filters2d* f, const float* y, int32_t B, int32_t C, int32_t H, int32_t W,
                        float* out, void* stream);
int esr_cem2d_inv_hth(const esr_cem_filters2d* f, const float* x, int32_t B, int32_t C, int32_t h, int32_t w,
                      float* out, void* stream);
int esr_cem2d_upscale(const esr_cem_filters2d* f, const float* x, int32_t B, int32_t C, int32_t h, int32_t w,
                      float* out, void* stream);
int esr_cem2d_project(const esr_cem_filters2d* f, const float* y, const float* x, int32_t B, int32_t C, int32_t H,
                      int32_t W, int32_t crop, float* out, float* workspace, void* stream);
int esr_cem2d_project_bwd(const esr_cem_filters2d* f, const float* g_out, int32_t B, int32_t C, int32_t H, int32_t W,
                          int32_t crop, float* g_y, float* workspace, void* stream);

/* ------------------------------------------------- Z optimisation plumbing (csrc/zopt.cu)
 * One iteration of the editing loop (codes/Z_optimization.py:572-635) without ATen kernels: Optimizable_Z.forward
 * (:300-305), the 'TV' / global-STD objectives on fake_H (:322-324, :426-435, :474-475, :525-535, :603-607, :618-619)
 * with their gradient, and tanh' chained with torch.optim.Adam's update (:512).  All tensors fp32, contiguous.
 *   Z           [B, n_lat] pre-tanh control signal (n_lat = nz * sf^2 * h * w), clamped to finite in place
 *   model_input [B, n_img] packed [Z.view, LR] generator input (n_img = n_lat + 3*h*w): only the latent part is written
 *   mode        0: w_std * (std - target[b])^2 + TV,  1: sign * std,  2: (std - target[b])^2   (std = torch.std, unbiased)
 *   stats       [B, 8] out: mean, std, gradient coefficients, loss_b (index 4), 1/nx, 1/ny, tv_b
 *   hist        [hist_len] or NULL: mean loss of iteration *step is stored at hist[*step]; *step is then incremented
 *   step        device int: completed iterations before the call; esr_zopt_adam reads the incremented value */
int esr_zopt_tanh_pack(float* Z, float z_range, int32_t B, int32_t n_lat, int32_t n_img, float* model_input, void* stream);
int32_t esr_zopt_loss_workspace_floats(int32_t B, int32_t H);
int esr_zopt_loss(const float* fake_H, int32_t B, int32_t C, int32_t H, int32_t W, int32_t mode, float sign, float w_std,
                  const float* target, float* workspace, float* stats, float* hist, int32_t hist_len, int32_t* step,
                  void* stream);
int esr_zopt_loss_grad(const float* fake_H, int32_t B, int32_t C, int32_t H, int32_t W, const float* stats, float* g,
                       void* stream);
/* g_in: [B, n_img] gradient w.r.t. model_input; Z, exp_avg, exp_avg_sq: [B, n_lat], updated in place */
int esr_zopt_adam(float* Z, float* exp_avg, float* exp_avg_sq, const float* g_in, float z_range, int32_t B, int32_t n_lat,
                  int32_t n_img, float lr, float beta1, float beta2, float eps, const int32_t* step, void* stream);

/* ------------------------------------------------- rich Z objectives (csrc/zobj.cu)
 * The pairwise kernel-density sums of SoftHistogramLoss.ComputeSoftHistogram (codes/Z_optimization.py:168-200) without
 * the reference's [D, N, M] fp64 intermediates:
 *   E[i,j] = exp(-(1 / (temperature * D)) * sum_d (min(|x|, |x - period|, |x + period|) + eps)^2),  x = p[d,i] - b[d,j]
 * samples p: fp32 [D, N] (d-major);  bins b: fp64 [D, M].  All arithmetic fp64, fixed summation order.
 * esr_kde_sums: sums[i] = sum over the OTHER set of E for every OWN vector: own = samples (own_is_f64 0, other_is_f64 1)
 *   gives the per-sample sums of the dictionary objective (:193-194), own = bins (1, 0) the per-bin sums of the histogram
 *   (:195).  workspace: esr_kde_workspace_bytes(n_own, n_other) bytes (may be 0 -> NULL allowed).
 * esr_kde_grad: grad[d,i] = sum_j (w_sample[i] + w_bin[j]) * dE[i,j]/dp[d,i]  (either weight vector may be NULL): the
 *   chain rule of any scalar built on the two kinds of sums, i.e. what autograd derives in the reference. */
int64_t esr_kde_workspace_bytes(int64_t n_own, int64_t n_other);
int esr_kde_sums(const void* own, int32_t own_is_f64, int64_t n_own, const void* other, int32_t other_is_f64,
                 int64_t n_other, int32_t D, double period, double temperature, double eps, double* sums,
                 void* workspace, void* stream);
int64_t esr_kde_grad_workspace_bytes(int64_t n_samples, int64_t n_bins, int32_t D);
int esr_kde_grad(const float* samples, int64_t n_samples, const double* bins, int64_t n_bins, int32_t D, double period,
                 double temperature, double eps, const double* w_sample, const double* w_bin, float* grad, void* workspace,
                 void* stream);
/* Host code, no GPU: the greedy patch selection of ReturnPatchExtractionMat (codes/Z_optimization.py:236-254).
 * patches [n, D] int64 pixel indexes in visiting order; valid [n] and covered [span] (span = max - min index) are outputs. */
int esr_patch_select(const int64_t* patches, int64_t n, int32_t D, double overlap, int64_t min_index, int64_t span,
                     uint8_t* valid, uint8_t* covered);

/* ------------------------------------------------- generator weight gradients (csrc/wgrad.cu)
 * What autograd computes for the generator's conv parameters in the reference's training step
 * (codes/models/SRRaGAN_model.py:463-547 with the block definitions of modules/block.py:129-155):
 *   dW[co, ci, ky, kx] = sum g[n,y,x,co] * X[n, y+ky-1, x+kx-1, ci],   db[co] = sum g[n,y,x,co]
 * from the tensors the forward and the data-gradient backward leave in HBM.  Item tables live in DEVICE memory. */
typedef struct esr_wgrad_item {       /* one conv x one block of 16 input channels (x one spatial chunk) */
    const void* x;                    /* conv input, NHWC 16-bit (bf16, or fp16 when x_f16), 16-byte aligned at x_c0 */
    const void* g;                    /* gradient w.r.t. the conv's pre-activation output, NHWC bf16 */
    float* dw;                        /* [n_co, cin_total, 3, 3] fp32 */
    int32_t x_stride, x_c0, x_f16;    /* channels per pixel of x, first channel of this block */
    int32_t g_stride, g_c0, cout;     /* channels per pixel of g, first channel, channels staged (16, 32, 48 or 64) */
    int32_t n_co, n_ci;               /* valid output channels (<= cout) / end of the valid input channels of the block (<= 16) */
    int32_t ci_lo;                    /* first valid input channel of the block: channels [ci_lo, n_ci) map to dW[:, ci0 ...] */
    int32_t cin_total, ci0;           /* dW's input-channel extent and this block's first input channel in it */
    int32_t B, H, W;
    int32_t tile_begin, tile_end;     /* tile_end > 0: only tiles [tile_begin, tile_end) of the B*ceil(H/8)*ceil(W/16), dW accumulated
                                         with atomicAdd (must be zeroed by the caller); 0: all tiles, dW slice overwritten */
    float* db;                        /* optional [n_co]: this item also sums g over its tiles (the bias gradient; give it to ONE
                                         block of the conv).  Overwritten / accumulated like dW */
} esr_wgrad_item;
int esr_wgrad16(const esr_wgrad_item* items_device, int32_t n_items, void* stream);
/* Same items, cout <= 32 each (cut 64-channel convs in two): the tile's eight pixel rows are split over the warps, every
 * warp keeping all nine taps' accumulators (11 ldmatrix per 36 MMAs instead of 2 per 2). */
int esr_wgrad16r(const esr_wgrad_item* items_device, int32_t n_items, void* stream);

/* The same sums for one (conv, block of 128 input channels, block of 32 output channels) on the tcgen05 tensor cores
 * (csrc/wgrad_tc.cu): both operands are read as TMA boxes of the NHWC bf16 tensors (MN-major MMA operands, no transposition),
 * nine accumulators [128 x 32] live in tensor memory over the item's pixel loop.  x_map / g_map index a DEVICE table of
 * tensor maps built with esr_wgrad_tc_make_map (kind 0 for conv inputs, 1 for gradients). */
typedef struct esr_wgrad_tc_item {
    uint32_t x_map, g_map;
    int32_t x_c0, g_c0;               /* first input channel of the block (boxes at x_c0 and x_c0 + 64) / first output channel */
    float* dw;                        /* [n_co, cin_total, 3, 3] fp32, at the block's first output channel */
    int32_t n_ci, n_co;               /* valid input channels (<= 128) / output channels (<= 32) of the block */
    int32_t cin_total, ci0;           /* dW's input-channel extent, index of channel x_c0 in it */
    int32_t B, H, W;
    int32_t tile_begin, tile_end;     /* tile range as in the esr_wgrad_item struct; tiles of 8 x 16 pixels */
    int32_t x_f16;                    /* must be 0: tcgen05 kind::f16 rejects an fp16 input with the bf16 gradient (illegal instruction) */
} esr_wgrad_tc_item;
int32_t esr_wgrad_tc_map_bytes(void);
int esr_wgrad_tc_make_map(void* map_host, const void* base, int32_t channels, int32_t B, int32_t H, int32_t W, int32_t kind);
int esr_wgrad_tc(const esr_wgrad_tc_item* items_device, int32_t n_items, const void* maps_device, void* stream);

typedef struct esr_wgrad_small_item { /* one conv: its <= 8 fp32 NCHW input channels (latent, LR image) and its bias */
    const void* g;                    /* NHWC bf16 */
    const float* g32;                 /* optional [B, n_co, H, W] fp32 gradient: used for the bias sum instead of g */
    const float* s;                   /* [B, s_channels, H, W] fp32 */
    float* dw;                        /* [n_co, cin_total, 3, 3], ACCUMULATED (atomicAdd over row chunks): zero it first */
    float* db;                        /* [n_co] or NULL, accumulated likewise */
    int32_t g_stride, g_c0, cout, n_co;
    int32_t s_channels, s_c0, n_c;    /* planes per image, first plane, planes used (0: bias only) */
    int32_t cin_total, ci0;
    int32_t B, H, W;
} esr_wgrad_small_item;
/* max_rows: the largest B*H among the items (row chunks of the grid) */
int esr_wgrad_small(const esr_wgrad_small_item* items_device, int32_t n_items, int32_t max_rows, void* stream);

/* Debug aid: the x4 CEM streaming kernels record a ring wait that never completed instead of trapping;
 * out4 = {code (0 = none, 1 = Down, 2 = K+Up), block, thread, group}; reading clears it. */
int esr_debug_cem_timeout(uint32_t* out4);
/* ESR_CEM_PROF=1: phase time stamps of the last single-launch x4 projection, [cta < 160][round < 4][8] globaltimer ns. */
int esr_debug_cem_fused_prof(unsigned long long* out, int n);

/* Debug aid (tools/prof.py): per-CTA role cycle counters of later tcgen05 conv launches, when the
 * library was built with -DESR_PROFILE_ROLES.  buf: [148][16] uint64 device memory or NULL. */
void esr_debug_set_profile_buffer(void* buf);

/* --------------------------------------------------------- recorded sequences */
/* A sequence is a host-side list of fully resolved kernel launches (tensor maps
 * encoded once) that esr_seq_run replays on a stream without returning to the
 * caller between layers: the 351-conv forward is one call. */
typedef struct esr_seq esr_seq;
esr_seq* esr_seq_create(void);
void esr_seq_destroy(esr_seq* s);
int esr_seq_add_conv(esr_seq* s, const esr_conv_desc* d, int32_t use_simt);
int esr_seq_add_rdb_growth(esr_seq* s, const esr_rdb_growth_desc* d);
int esr_seq_run(const esr_seq* s, void* stream);
int32_t esr_seq_num_launches(const esr_seq* s);

#ifdef __cplusplus
}
#endif
#endif /* ESR_B200_H */
