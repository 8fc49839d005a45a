// Rich Z objectives: the pairwise kernel-density arithmetic behind the reference's histogram and
// dictionary losses (codes/Z_optimization.py:168-200, SoftHistogramLoss.ComputeSoftHistogram), which the reference
// evaluates by materialising [D, N, M] fp64 tensors (N image pixels / patches, M bins or dictionary atoms, D values per
// sample) - the second largest cost of the editing loop after G itself.  Here nothing of size N x M is stored:
//
//   E[i,j] = exp( -(1 / (T*D)) * sum_d ( min(|x|, |x - period|, |x + period|) + eps )^2 ),   x = p[d,i] - b[d,j]
//
//   esr_kde_sums   sum_j E[i,j] per sample ("dictionary": -log mean_j) or sum_i E[i,j] per bin ("histogram"), fp64
//   esr_kde_grad   d/dp[d,i] of any scalar built on those sums: sum_j (w_sample[i] + w_bin[j]) * E[i,j] * dlogE[i,j]/dp[d,i]
//
// One thread owns one "own" vector (kept in shared memory, column per thread), the "other" vectors stream through a
// shared chunk that every thread reads as a broadcast.  Sums and gradients split the other range over blockIdx.y into
// partial results (fp64) that a second kernel adds in a fixed order (deterministic, no atomics) whenever the own set alone
// would leave SMs idle (3600 patches = 29 CTAs: the unsplit gradient pass took 13.3 ms of a 13.9 ms forward + backward).
// All arithmetic is fp64 like the reference (`.type(torch.cuda.DoubleTensor)`, :172/:184); the samples arrive as fp32 (the
// generator's output) and the gradient leaves as fp32 (what autograd's cast gives).
//
// Also here: the greedy patch selection of ReturnPatchExtractionMat (:236-254) as a host loop in native code (the
// reference walks every candidate patch in Python).
#include <cmath>
#include <cstdint>

#include "esr_common.cuh"

namespace esr {

constexpr int kKdeBlock = 128;     // own vectors per CTA, one per thread
constexpr int kKdeChunk = 32;      // other vectors per shared-memory round
constexpr int kKdeMaxD = 81;       // 9x9 patches
constexpr int kKdeTargetCtas = 4 * 148;

__device__ __forceinline__ double kde_wrapped(double x, double period, double* signed_out) {
    // the reference's three-way min (:186-188): the distance on a circle of circumference `period`
    const double a0 = fabs(x), a1 = fabs(x - period), a2 = fabs(x + period);
    double t = x, m = a0;
    if (a1 < m) { m = a1; t = x - period; }
    if (a2 < m) { m = a2; t = x + period; }
    if (signed_out) *signed_out = t;
    return m;
}

template <typename TOwn, typename TOther, bool GRAD>
__global__ void __launch_bounds__(kKdeBlock) kde_pair_kernel(const TOwn* __restrict__ own, long long n_own,
                                                             const TOther* __restrict__ other, long long n_other, int D,
                                                             double period, double neg_inv_td, double eps,
                                                             const double* __restrict__ w_own, const double* __restrict__ w_other,
                                                             double* __restrict__ sums, float* __restrict__ grad,
                                                             double* __restrict__ grad_partial) {
    extern __shared__ double kde_smem[];
    double* so = kde_smem;                                   // [D][kKdeBlock]
    double* sc = so + static_cast<size_t>(D) * kKdeBlock;    // [D][kKdeChunk]
    double* sg = sc + static_cast<size_t>(D) * kKdeChunk;    // GRAD: [D][kKdeBlock]
    double* sw = sg + static_cast<size_t>(D) * kKdeBlock;    // GRAD: [kKdeChunk]
    const int tid = threadIdx.x;
    const long long i = static_cast<long long>(blockIdx.x) * kKdeBlock + tid;
    const bool live = i < n_own;
    for (int d = 0; d < D; ++d) {
        so[d * kKdeBlock + tid] = live ? static_cast<double>(own[static_cast<size_t>(d) * n_own + i]) : 0.0;
        if (GRAD) sg[d * kKdeBlock + tid] = 0.0;
    }
    const long long per = (n_other + gridDim.y - 1) / gridDim.y;
    const long long lo = per * blockIdx.y, hi = min(lo + per, n_other);
    const double wi = (GRAD && w_own != nullptr && live) ? w_own[i] : 0.0;
    double sum = 0.0;
    for (long long j0 = lo; j0 < hi; j0 += kKdeChunk) {
        const int nj = static_cast<int>(min(static_cast<long long>(kKdeChunk), hi - j0));
        __syncthreads();
        for (int e = tid; e < D * kKdeChunk; e += kKdeBlock) {
            const int d = e / kKdeChunk, jj = e % kKdeChunk;
            sc[e] = jj < nj ? static_cast<double>(other[static_cast<size_t>(d) * n_other + j0 + jj]) : 0.0;
        }
        if (GRAD && tid < kKdeChunk) sw[tid] = (w_other != nullptr && tid < nj) ? w_other[j0 + tid] : 0.0;
        __syncthreads();
        if (!live) continue;
        for (int jj = 0; jj < nj; ++jj) {
            double acc = 0.0;
            for (int d = 0; d < D; ++d) {
                const double m = kde_wrapped(so[d * kKdeBlock + tid] - sc[d * kKdeChunk + jj], period, nullptr) + eps;
                acc = fma(m, m, acc);
            }
            const double E = exp(acc * neg_inv_td);
            if (!GRAD) {
                sum += E;
            } else {
                const double w = (wi + sw[jj]) * E * 2.0 * neg_inv_td;
                if (w != 0.0) {
                    for (int d = 0; d < D; ++d) {
                        double t;
                        const double m = kde_wrapped(so[d * kKdeBlock + tid] - sc[d * kKdeChunk + jj], period, &t) + eps;
                        const double sgn = t > 0.0 ? 1.0 : (t < 0.0 ? -1.0 : 0.0);      // d|t|/dt, 0 at 0 like torch.abs
                        sg[d * kKdeBlock + tid] = fma(w * sgn, m, sg[d * kKdeBlock + tid]);
                    }
                }
            }
        }
    }
    if (!live) return;
    if (!GRAD) {
        sums[static_cast<size_t>(blockIdx.y) * n_own + i] = sum;
    } else if (gridDim.y == 1) {
        for (int d = 0; d < D; ++d) grad[static_cast<size_t>(d) * n_own + i] = static_cast<float>(sg[d * kKdeBlock + tid]);
    } else {                         // few samples: the bin range is split like the sums', partial gradients in fp64
        double* dst = grad_partial + static_cast<size_t>(blockIdx.y) * D * n_own;
        for (int d = 0; d < D; ++d) dst[static_cast<size_t>(d) * n_own + i] = sg[d * kKdeBlock + tid];
    }
}

__global__ void kde_reduce_kernel(const double* __restrict__ partial, int nsplit, long long n, double* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < nsplit; ++k) s += partial[static_cast<size_t>(k) * n + i];
    out[i] = s;
}

__global__ void kde_reduce_grad_kernel(const double* __restrict__ partial, int nsplit, long long n, float* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < nsplit; ++k) s += partial[static_cast<size_t>(k) * n + i];
    out[i] = static_cast<float>(s);
}

static int kde_nsplit(long long n_own, long long n_other) {
    const long long blocks = (n_own + kKdeBlock - 1) / kKdeBlock;
    long long s = (kKdeTargetCtas + blocks - 1) / blocks;
    const long long max_s = (n_other + kKdeChunk - 1) / kKdeChunk;
    if (s > max_s) s = max_s;
    if (s > 65535) s = 65535;
    return s < 1 ? 1 : static_cast<int>(s);
}

static size_t kde_smem_bytes(int D, bool grad) {
    size_t n = static_cast<size_t>(D) * (kKdeBlock + kKdeChunk);
    if (grad) n += static_cast<size_t>(D) * kKdeBlock + kKdeChunk;
    return n * sizeof(double);
}

template <typename K>
static int kde_allow_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) ESR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    return ESR_OK;
}

static int kde_check(const void* own, long long n_own, const void* other, long long n_other, int D, double period,
                     double temperature) {
    ESR_CHECK_ARG(own != nullptr && other != nullptr, "kde: null sample / bin pointer");
    ESR_CHECK_ARG(n_own > 0 && n_other > 0, "kde: empty sample or bin set (%lld, %lld)", n_own, n_other);
    ESR_CHECK_ARG(D >= 1 && D <= kKdeMaxD, "kde: %d values per sample (supported: 1..%d)", D, kKdeMaxD);
    ESR_CHECK_ARG(temperature > 0.0 && std::isfinite(temperature), "kde: temperature %g", temperature);
    ESR_CHECK_ARG(period > 0.0, "kde: period %g", period);
    ESR_CHECK_ARG((n_own + kKdeBlock - 1) / kKdeBlock <= 0x7fffffffLL, "kde: %lld own vectors exceed the grid", n_own);
    return ESR_OK;
}

}  // namespace esr

extern "C" int64_t esr_kde_workspace_bytes(int64_t n_own, int64_t n_other) {
    if (n_own <= 0 || n_other <= 0) return ESR_ERR_INVALID;
    const int s = esr::kde_nsplit(n_own, n_other);
    return s > 1 ? static_cast<int64_t>(s) * n_own * static_cast<int64_t>(sizeof(double)) : 0;
}

extern "C" int esr_kde_sums(const void* own, int32_t own_is_f64, int64_t n_own, const void* other, int32_t other_is_f64,
                            int64_t n_other, int32_t D, double period, double temperature, double eps, double* sums,
                            void* workspace, void* stream) {
    using namespace esr;
    int rc = kde_check(own, n_own, other, n_other, D, period, temperature);
    if (rc != ESR_OK) return rc;
    ESR_CHECK_ARG(sums != nullptr, "esr_kde_sums: null output");
    ESR_CHECK_ARG((own_is_f64 != 0) != (other_is_f64 != 0),
                  "esr_kde_sums: one side is the fp32 samples, the other the fp64 bins (got own_is_f64=%d, other_is_f64=%d)",
                  own_is_f64, other_is_f64);
    const int nsplit = kde_nsplit(n_own, n_other);
    ESR_CHECK_ARG(nsplit == 1 || workspace != nullptr, "esr_kde_sums: %d partial sums need the workspace of esr_kde_workspace_bytes", nsplit);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = kde_smem_bytes(D, false);
    const dim3 grid(static_cast<unsigned>((n_own + kKdeBlock - 1) / kKdeBlock), nsplit);
    double* dst = nsplit == 1 ? sums : static_cast<double*>(workspace);
    const double neg_inv_td = -1.0 / (temperature * D);
    if (own_is_f64) {
        auto k = kde_pair_kernel<double, float, false>;
        if ((rc = kde_allow_smem(k, smem)) != ESR_OK) return rc;
        k<<<grid, kKdeBlock, smem, st>>>(static_cast<const double*>(own), n_own, static_cast<const float*>(other), n_other, D,
                                        period, neg_inv_td, eps, nullptr, nullptr, dst, nullptr, nullptr);
    } else {
        auto k = kde_pair_kernel<float, double, false>;
        if ((rc = kde_allow_smem(k, smem)) != ESR_OK) return rc;
        k<<<grid, kKdeBlock, smem, st>>>(static_cast<const float*>(own), n_own, static_cast<const double*>(other), n_other, D,
                                        period, neg_inv_td, eps, nullptr, nullptr, dst, nullptr, nullptr);
    }
    if ((rc = check_launch("kde_pair_kernel")) != ESR_OK) return rc;
    if (nsplit > 1) {
        kde_reduce_kernel<<<static_cast<unsigned>((n_own + 255) / 256), 256, 0, st>>>(dst, nsplit, n_own, sums);
        rc = check_launch("kde_reduce_kernel");
    }
    return rc;
}

extern "C" int64_t esr_kde_grad_workspace_bytes(int64_t n_samples, int64_t n_bins, int32_t D) {
    if (n_samples <= 0 || n_bins <= 0 || D <= 0) return ESR_ERR_INVALID;
    const int s = esr::kde_nsplit(n_samples, n_bins);
    return s > 1 ? static_cast<int64_t>(s) * D * n_samples * static_cast<int64_t>(sizeof(double)) : 0;
}

extern "C" int esr_kde_grad(const float* samples, int64_t n_samples, const double* bins, int64_t n_bins, int32_t D,
                            double period, double temperature, double eps, const double* w_sample, const double* w_bin,
                            float* grad, void* workspace, void* stream) {
    using namespace esr;
    int rc = kde_check(samples, n_samples, bins, n_bins, D, period, temperature);
    if (rc != ESR_OK) return rc;
    ESR_CHECK_ARG(grad != nullptr, "esr_kde_grad: null output");
    ESR_CHECK_ARG(w_sample != nullptr || w_bin != nullptr, "esr_kde_grad: neither per-sample nor per-bin weights given");
    const size_t smem = kde_smem_bytes(D, true);
    auto k = kde_pair_kernel<float, double, true>;
    if ((rc = kde_allow_smem(k, smem)) != ESR_OK) return rc;
    const int nsplit = kde_nsplit(n_samples, n_bins);
    ESR_CHECK_ARG(nsplit == 1 || workspace != nullptr, "esr_kde_grad: %d partial gradients need the workspace of esr_kde_grad_workspace_bytes", nsplit);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid(static_cast<unsigned>((n_samples + kKdeBlock - 1) / kKdeBlock), nsplit);
    k<<<grid, kKdeBlock, smem, st>>>(samples, n_samples, bins, n_bins, D, period, -1.0 / (temperature * D), eps, w_sample, w_bin,
                                    nullptr, grad, static_cast<double*>(workspace));
    if ((rc = check_launch("kde_pair_kernel<grad>")) != ESR_OK) return rc;
    if (nsplit > 1) {
        const long long n = static_cast<long long>(D) * n_samples;
        kde_reduce_grad_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(static_cast<const double*>(workspace), nsplit, n, grad);
        rc = check_launch("kde_reduce_grad_kernel");
    }
    return rc;
}

// Greedy patch selection (codes/Z_optimization.py:236-254): candidates are visited in raster order; one is dropped when
// the share of its pixels already covered by kept patches exceeds `overlap` (any covered pixel when overlap == 0).
// The coverage table has span = max - min slots addressed with (pixel - min - 1) and Python's negative-index wrap, so
// the smallest and the largest pixel index share the last slot - kept, the selection must equal the reference's.
// patches: [n, D] pixel indexes (host);  valid: [n] out;  covered: [span] out (the table, for the caller's non-covered set)
extern "C" int esr_patch_select(const int64_t* patches, int64_t n, int32_t D, double overlap, int64_t min_index, int64_t span,
                                uint8_t* valid, uint8_t* covered) {
    ESR_CHECK_ARG(patches != nullptr && valid != nullptr && covered != nullptr, "esr_patch_select: null argument");
    ESR_CHECK_ARG(n >= 0 && D > 0 && span > 0, "esr_patch_select: n=%lld D=%d span=%lld", static_cast<long long>(n), D,
                  static_cast<long long>(span));
    for (int64_t s = 0; s < span; ++s) covered[s] = 0;
    for (int64_t p = 0; p < n; ++p) {
        const int64_t* px = patches + p * D;
        int taken = 0;
        for (int d = 0; d < D; ++d) {
            int64_t slot = px[d] - min_index - 1;
            if (slot < 0) slot += span;
            if (slot < 0 || slot >= span) {
                esr::set_error("esr_patch_select: pixel index %lld outside [%lld, %lld]", static_cast<long long>(px[d]),
                               static_cast<long long>(min_index), static_cast<long long>(min_index + span));
                return ESR_ERR_INVALID;
            }
            taken += covered[slot];
        }
        const bool drop = (overlap == 0.0 && taken > 0) || (static_cast<double>(taken) / D > overlap);
        valid[p] = drop ? 0 : 1;
        if (drop) continue;
        for (int d = 0; d < D; ++d) {
            int64_t slot = px[d] - min_index - 1;
            if (slot < 0) slot += span;
            covered[slot] = 1;
        }
    }
    return ESR_OK;
}
