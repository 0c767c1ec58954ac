// pnpb200_analysis.cu -- the analysis stage of face_variation_test.py (:631-759) on the device:
// the "top 10 % of |error|" selection the script does with four heaps of (-|err|, idx) tuples, and
// get_most_fragile_point_and_perturbation_direction over the selected problems (count of the
// landmark with the largest perturbation, Gram matrix of the perturbation vectors whose
// eigen-decomposition gives the script's SVD directions), for up to four error quantities at once.
//
// Selection = exact k-th largest of the composite key (|value| as the bit pattern of a non-negative
// double, then SMALLER global index first -- the order heapq pops (-|err|, idx) tuples in) by an
// MSB-first radix select: twelve 8-bit digits, one histogram kernel per digit.  The histograms
// are all a multi-GPU run has to exchange (SUM), so the global top-k needs no gather.
#include "pnpb200_common.cuh"
#include "pnpb200_math.cuh"

namespace pnpb200 {

constexpr int kMaxQ = 4;
constexpr int kDigits = 12;      // 8 digits of the value bits, 4 of the index

struct TopkIn {
    const double* v[kMaxQ];
    long long stride[kMaxQ];
    unsigned long long pre_hi[kMaxQ];   // digits decided so far (MSB-aligned), value part
    unsigned int pre_lo[kMaxQ];         // ... index part
    int nq;
};

PNP_DEV void make_key(double v, long long gidx, unsigned long long& hi, unsigned int& lo)
{
    const double a = fabs(v);
    hi = (a != a) ? 0x7ff0000000000000ull : (unsigned long long)__double_as_longlong(a);   // NaN counts as +inf
    lo = ~(unsigned int)gidx;                                                              // smaller index = larger key
}
// digit d of the key, d = 0 most significant
PNP_DEV unsigned key_digit(unsigned long long hi, unsigned int lo, int d)
{
    return d < 8 ? (unsigned)((hi >> (56 - 8 * d)) & 0xffu) : (unsigned)((lo >> (24 - 8 * (d - 8))) & 0xffu);
}
// do the first nd digits of the key equal the prefix?
PNP_DEV bool prefix_matches(unsigned long long hi, unsigned int lo, unsigned long long phi, unsigned int plo, int nd)
{
    if (nd <= 0) return true;
    if (nd <= 8) return (hi >> (64 - 8 * nd)) == (phi >> (64 - 8 * nd));
    if (hi != phi) return false;
    if (nd >= 12) return lo == plo;
    return (lo >> (32 - 8 * (nd - 8))) == (plo >> (32 - 8 * (nd - 8)));
}
PNP_DEV bool key_ge(unsigned long long hi, unsigned int lo, unsigned long long thi, unsigned int tlo)
{
    return hi > thi || (hi == thi && lo >= tlo);
}

// hist[q][256] += number of problems whose key matches the decided prefix, by their next digit
__global__ void __launch_bounds__(256) k_topk_hist(long long B, long long idx0, TopkIn in, int nd, unsigned long long* hist)
{
    __shared__ unsigned int sh[kMaxQ * 256];
    for (int e = threadIdx.x; e < kMaxQ * 256; e += blockDim.x) sh[e] = 0u;
    __syncthreads();
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) {
            if (q < in.nq) {
                unsigned long long hi; unsigned int lo;
                make_key(in.v[q][b * in.stride[q]], idx0 + b, hi, lo);
                if (prefix_matches(hi, lo, in.pre_hi[q], in.pre_lo[q], nd)) atomicAdd(&sh[q * 256 + key_digit(hi, lo, nd)], 1u);
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < in.nq * 256; e += blockDim.x)
        if (sh[e]) atomicAdd(&hist[e], (unsigned long long)sh[e]);
}

// Pass over the problems with the thresholds known: index list of the selected problems, their sum
// and maximum, and the count of the landmark with the largest perturbation norm (strict '>',
// first landmark wins, face_variation_test.py:668-677).
__global__ void __launch_bounds__(256) k_select(long long B, long long idx0, TopkIn thr, const double* __restrict__ perturb, int n,
                                                long long cap, long long* list, unsigned long long* n_sel,
                                                unsigned long long* count, double* vsum, double* vmax)
{
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        int worst = -2;                                   // computed once, for the first quantity that selects b
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) {
            if (q < thr.nq) {
                const double v = thr.v[q][b * thr.stride[q]];
                unsigned long long hi; unsigned int lo;
                make_key(v, idx0 + b, hi, lo);
                if (!key_ge(hi, lo, thr.pre_hi[q], thr.pre_lo[q])) continue;
                const unsigned long long pos = atomicAdd(&n_sel[q], 1ull);
                if ((long long)pos < cap) list[(size_t)q * cap + pos] = b;
                const double a = fabs(v);
                atomicAdd(&vsum[q], a);
                atomicMax(reinterpret_cast<unsigned long long*>(&vmax[q]), (unsigned long long)__double_as_longlong(a));
                if (worst == -2) {
                    double nm = -1.0;
                    worst = -1;
                    const double* p = perturb + (size_t)b * n * 3;
                    for (int i = 0; i < n; ++i) {
                        const double ni = sqrt(p[3 * i] * p[3 * i] + p[3 * i + 1] * p[3 * i + 1] + p[3 * i + 2] * p[3 * i + 2]);
                        if (ni > nm) { nm = ni; worst = i; }
                    }
                }
                if (worst >= 0) atomicAdd(&count[(size_t)q * n + worst], 1ull);
            }
        }
    }
}

// gram[q] (D x D, D = 3n, upper tiles only; the host mirrors it) += sum over the selected problems of v v^T.
// Block = 16 x 16 entries of one tile; blockIdx.z = quantity * slices + slice of the list.
constexpr int kGramTile = 16;
constexpr int kGramRows = 64;
__global__ void __launch_bounds__(kGramTile * kGramTile) k_gram(const double* __restrict__ perturb, int D, long long cap,
                                                                const long long* __restrict__ list,
                                                                const unsigned long long* __restrict__ n_sel, int slices, double* gram)
{
    const int ti = blockIdx.x, tj = blockIdx.y;
    if (tj < ti) return;
    const int q = blockIdx.z / slices, slice = blockIdx.z % slices;
    long long m = (long long)n_sel[q];
    if (m > cap) m = cap;
    const long long per = (m + slices - 1) / slices;
    const long long r0 = slice * per, r1 = (r0 + per < m) ? (r0 + per) : m;
    __shared__ double sa[kGramRows][kGramTile + 1], sb[kGramRows][kGramTile + 1];
    const int tx = threadIdx.x % kGramTile, ty = threadIdx.x / kGramTile;
    const int gi = ti * kGramTile + ty, gj = tj * kGramTile + tx;
    double acc = 0.0;
    for (long long r = r0; r < r1; r += kGramRows) {
        const int rows = (int)((r1 - r < kGramRows) ? (r1 - r) : kGramRows);
        for (int e = threadIdx.x; e < kGramRows * kGramTile; e += blockDim.x) {
            const int rr = e / kGramTile, c = e % kGramTile;
            double va = 0.0, vb = 0.0;
            if (rr < rows) {
                const double* row = perturb + (size_t)list[(size_t)q * cap + r + rr] * D;
                if (ti * kGramTile + c < D) va = row[ti * kGramTile + c];
                if (tj * kGramTile + c < D) vb = row[tj * kGramTile + c];
            }
            sa[rr][c] = va; sb[rr][c] = vb;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < kGramRows; ++rr) acc = fma(sa[rr][ty], sb[rr][tx], acc);
        __syncthreads();
    }
    if (gi < D && gj < D && acc != 0.0) atomicAdd(&gram[((size_t)q * D + gi) * D + gj], acc);
}

static inline unsigned grid_cap(long long n, int block, int cap)
{
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (unsigned)g;
}

}  // namespace pnpb200

using namespace pnpb200;

extern "C" {

int pnpb200_topk_histogram(int64_t B, int64_t idx0, int nq, const double* const* values, const int64_t* stride,
                           const uint64_t* prefix_hi, const uint32_t* prefix_lo, int n_digits_decided, uint64_t* hist,
                           void* stream)
{
    if (B < 0 || nq < 1 || nq > kMaxQ || !values || !stride || !hist || n_digits_decided < 0 || n_digits_decided >= kDigits)
        return PNPB200_EINVAL;
    if (n_digits_decided > 0 && (!prefix_hi || !prefix_lo)) return PNPB200_EINVAL;
    TopkIn in;
    in.nq = nq;
    for (int q = 0; q < kMaxQ; ++q) {
        in.v[q] = (q < nq) ? values[q] : nullptr;
        in.stride[q] = (q < nq) ? stride[q] : 0;
        in.pre_hi[q] = (q < nq && prefix_hi) ? prefix_hi[q] : 0ull;
        in.pre_lo[q] = (q < nq && prefix_lo) ? prefix_lo[q] : 0u;
        if (q < nq && !in.v[q]) return PNPB200_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PNP_CUDA_OK(cudaMemsetAsync(hist, 0, sizeof(uint64_t) * 256 * (size_t)nq, st));
    if (B == 0) return PNPB200_OK;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc != PNPB200_OK) return rc;
    k_topk_hist<<<grid_cap(B, 256, dp.sm_count * 8), 256, 0, st>>>(B, idx0, in, n_digits_decided, (unsigned long long*)hist); count_kernel_launches(1);
    PNP_CUDA_OK(cudaGetLastError());
    return PNPB200_OK;
}

int pnpb200_fragility_accumulate(int64_t B, int64_t idx0, int nq, const double* const* values, const int64_t* stride,
                                 const uint64_t* threshold_hi, const uint32_t* threshold_lo, const double* perturb, int n,
                                 int64_t list_capacity, int64_t* list, uint64_t* n_selected, uint64_t* count, double* value_sum,
                                 double* value_max, double* gram, void* stream)
{
    if (B < 0 || nq < 1 || nq > kMaxQ || !values || !stride || !threshold_hi || !threshold_lo || !perturb || n < 1 ||
        list_capacity < 1 || !list || !n_selected || !count || !value_sum || !value_max || !gram)
        return PNPB200_EINVAL;
    TopkIn in;
    in.nq = nq;
    for (int q = 0; q < kMaxQ; ++q) {
        in.v[q] = (q < nq) ? values[q] : nullptr;
        in.stride[q] = (q < nq) ? stride[q] : 0;
        in.pre_hi[q] = (q < nq) ? threshold_hi[q] : ~0ull;
        in.pre_lo[q] = (q < nq) ? threshold_lo[q] : ~0u;
        if (q < nq && !in.v[q]) return PNPB200_EINVAL;
    }
    const int D = 3 * n;
    cudaStream_t st = (cudaStream_t)stream;
    PNP_CUDA_OK(cudaMemsetAsync(n_selected, 0, sizeof(uint64_t) * (size_t)nq, st));
    PNP_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(uint64_t) * (size_t)nq * n, st));
    PNP_CUDA_OK(cudaMemsetAsync(value_sum, 0, sizeof(double) * (size_t)nq, st));
    PNP_CUDA_OK(cudaMemsetAsync(value_max, 0, sizeof(double) * (size_t)nq, st));
    PNP_CUDA_OK(cudaMemsetAsync(gram, 0, sizeof(double) * (size_t)nq * D * D, st));
    if (B == 0) return PNPB200_OK;
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc != PNPB200_OK) return rc;
    k_select<<<grid_cap(B, 256, dp.sm_count * 8), 256, 0, st>>>(B, idx0, in, perturb, n, list_capacity, (long long*)list,
                                                               (unsigned long long*)n_selected, (unsigned long long*)count,
                                                               value_sum, value_max); count_kernel_launches(1);
    const int tiles = (D + kGramTile - 1) / kGramTile;
    int slices = (dp.sm_count * 4) / (tiles * (tiles + 1) / 2 * nq);
    if (slices < 1) slices = 1;
    if (slices > 256) slices = 256;
    k_gram<<<dim3(tiles, tiles, nq * slices), kGramTile * kGramTile, 0, st>>>(perturb, D, list_capacity, (const long long*)list,
                                                                              (const unsigned long long*)n_selected, slices, gram); count_kernel_launches(1);
    PNP_CUDA_OK(cudaGetLastError());
    return PNPB200_OK;
}

}  // extern "C"
