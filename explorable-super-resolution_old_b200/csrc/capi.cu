// C-ABI glue: error reporting, device check and recorded launch sequences.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "conv3x3.cuh"

namespace esr {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int build_conv_launch(const esr_conv_desc& d, CUtensorMap* tm0, CUtensorMap* tm1, ConvLaunch* L);
int launch_conv_tc(const CUtensorMap& tm0, const CUtensorMap& tm1, const ConvLaunch& L, cudaStream_t stream);
int launch_conv_simt(const ConvLaunch& L, cudaStream_t stream);

struct SeqOp {
    alignas(64) CUtensorMap tm0;
    alignas(64) CUtensorMap tm1;
    ConvLaunch L;
    int use_simt;
};

struct RdbOp;                                   // conv3x3_tc2.cu
RdbOp* new_rdb_op(const esr_rdb_growth_desc& d, int* rc);
void rdb_op_set_reverse(RdbOp* op, int reverse);
void delete_rdb_op(RdbOp* op);
int launch_rdb_growth(const RdbOp& op, cudaStream_t stream, int use_pdl);

}  // namespace esr

struct esr_seq {
    std::vector<esr::SeqOp> ops;
    std::vector<esr::RdbOp*> rdb;               // ops[i].use_simt == -1 - k  <=>  fused growth launch rdb[k]
    ~esr_seq() { for (esr::RdbOp* r : rdb) esr::delete_rdb_op(r); }
};

extern "C" const char* esr_last_error(void) { return esr::g_error; }
extern "C" int esr_abi_version(void) { return 2; }   // 2: pair weight layout / esr_rdb_growth_* / hi-lo trunk fields

extern "C" int esr_device_check(int device) {
    cudaDeviceProp p;
    ESR_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major != 10) {
        esr::set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, p.major, p.minor);
        return ESR_ERR_UNSUPPORTED;
    }
    return ESR_OK;
}

extern "C" esr_seq* esr_seq_create(void) { return new (std::nothrow) esr_seq(); }
extern "C" void esr_seq_destroy(esr_seq* s) { delete s; }
extern "C" int32_t esr_seq_num_launches(const esr_seq* s) { return s ? static_cast<int32_t>(s->ops.size()) : 0; }

extern "C" int esr_seq_add_conv(esr_seq* s, const esr_conv_desc* d, int32_t use_simt) {
    if (s == nullptr || d == nullptr) { esr::set_error("esr_seq_add_conv: null argument"); return ESR_ERR_INVALID; }
    esr::SeqOp op;
    std::memset(&op, 0, sizeof(op));
    int rc = esr::build_conv_launch(*d, &op.tm0, &op.tm1, &op.L);
    if (rc != ESR_OK) return rc;
    op.use_simt = use_simt;
    static const bool snake = []() { const char* v = getenv("ESR_NO_SNAKE"); return !(v && atoi(v)); }();
    op.L.reverse = snake ? static_cast<int>(s->ops.size() & 1) : 0;     // alternate the tile walk layer by layer
    s->ops.push_back(op);
    return ESR_OK;
}

extern "C" int esr_seq_add_rdb_growth(esr_seq* s, const esr_rdb_growth_desc* d) {
    if (s == nullptr || d == nullptr) { esr::set_error("esr_seq_add_rdb_growth: null argument"); return ESR_ERR_INVALID; }
    int rc = ESR_OK;
    esr::RdbOp* r = esr::new_rdb_op(*d, &rc);
    if (r == nullptr) return rc;
    esr::SeqOp op;
    std::memset(&op, 0, sizeof(op));
    op.use_simt = -1 - static_cast<int>(s->rdb.size());
    static const bool snake = []() { const char* v = getenv("ESR_NO_SNAKE"); return !(v && atoi(v)); }();
    esr::rdb_op_set_reverse(r, snake ? static_cast<int>(s->ops.size() & 1) : 0);   // alternate with the neighbouring launches
    s->rdb.push_back(r);
    s->ops.push_back(op);
    return ESR_OK;
}

extern "C" int64_t esr_rdb_growth_flag_words(int32_t B, int32_t H, int32_t W) {
    if (B <= 0 || H <= 0 || W <= 0) return ESR_ERR_INVALID;
    const int64_t tiles = static_cast<int64_t>(B) * esr::ceil_div(W, esr::kTileWOut) * esr::ceil_div(H, 2 * esr::kBandRows);
    return 3 * ESR_RDB_MAX_LAYERS * tiles;
}

extern "C" int esr_rdb_growth_tc(const esr_rdb_growth_desc* d, void* stream) {
    if (d == nullptr) { esr::set_error("null rdb_growth desc"); return ESR_ERR_INVALID; }
    int rc = ESR_OK;
    esr::RdbOp* r = esr::new_rdb_op(*d, &rc);
    if (r == nullptr) return rc;
    rc = esr::launch_rdb_growth(*r, static_cast<cudaStream_t>(stream), 0);
    esr::delete_rdb_op(r);
    return rc;
}

extern "C" int esr_seq_run(const esr_seq* s, void* stream) {
    if (s == nullptr) { esr::set_error("esr_seq_run: null sequence"); return ESR_ERR_INVALID; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static const bool sync_each = []() { const char* v = getenv("ESR_SEQ_SYNC"); return v && atoi(v); }();   // debug aid
    int idx = 0;
    static const int use_pdl = []() { const char* v = getenv("ESR_NO_PDL"); return (v && atoi(v)) ? 0 : 1; }();
    for (const esr::SeqOp& op : s->ops) {
        int rc;
        if (op.use_simt < 0) rc = esr::launch_rdb_growth(*s->rdb[-1 - op.use_simt], st, use_pdl);
        else rc = op.use_simt ? esr::launch_conv_simt(op.L, st) : esr::launch_conv_tc(op.tm0, op.tm1, op.L, st);
        if (rc != ESR_OK) return rc;
        if (sync_each) {
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) {
                esr::set_error("sequence op %d (%s cout_tile %d x %d, pair %d, %d k-blocks, %dx%dx%d) failed: %s", idx,
                               op.use_simt < 0 ? "fused growth convs," : "conv,", op.L.d.cout_tile,
                               op.L.d.cout_tiles, op.L.d.pair, op.L.d.num_kblocks, op.L.d.B, op.L.d.H, op.L.d.W, cudaGetErrorString(e));
                return ESR_ERR_CUDA;
            }
        }
        ++idx;
    }
    return ESR_OK;
}
