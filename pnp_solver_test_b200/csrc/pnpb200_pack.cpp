// pnpb200_pack.cpp -- host side of the packed pixel transfer of the host-buffer pipeline.
//
// The pixels a landmark detector delivers are whole numbers (random_stress_test.py projects with
// is_quantized=True, PNP_SOLVER_LIB.py:4549-4552), yet the reference carries them as float64: 16 bytes
// per landmark over PCIe where 4 hold the same information.  pnpb200_pack_i16 turns a run of FP64 / FP32
// values into int16 and reports whether that was EXACT for every value (whole, inside [-32768, 32767],
// no NaN); the pipeline ships the packed chunk only then and widens it on the device, so the solver sees
// bit-identical inputs (a -0.0 arrives as +0.0) -- any other chunk travels as it is.
// AVX2 when the CPU has it (run-time check), several threads per call.
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <immintrin.h>
#include <stdint.h>
#include <thread>
#include <unistd.h>
#include <vector>

#include "../../include/pnpb200.h"

namespace {

template <typename T>
bool pack_scalar(const T* s, int64_t n, int16_t* d)
{
    bool ok = true;
    for (int64_t i = 0; i < n; ++i) {
        const T v = s[i];
        const bool in = (v >= (T)-32768) && (v <= (T)32767);      // false for NaN
        const int16_t q = in ? (int16_t)v : (int16_t)0;
        d[i] = q;
        ok = ok && in && ((T)q == v);
    }
    return ok;
}

// One core streams ~9 GB/s with demand loads alone (its fill buffers); prefetching 4 KB ahead lifts that to ~15 GB/s and
// non-temporal stores spare the read-for-ownership of the destination: 81 -> 111 GB/s of FP64 read on 15 threads
// (tools/pack_speed.py on the GPU box's 16-core Xeon).
constexpr int kPackPrefetchBytes = 4096;
__attribute__((target("avx2"))) bool pack_f64_avx2(const double* s, int64_t n, int16_t* d)
{
    __m256d bad = _mm256_setzero_pd();
    __m128i range = _mm_setzero_si128();
    const __m128i bias = _mm_set1_epi32(32768);
    int64_t i = 0;
    const bool nt = ((uintptr_t)d & 15u) == 0;
    for (; i + 16 <= n; i += 16) {
        _mm_prefetch((const char*)(s + i) + kPackPrefetchBytes, _MM_HINT_T0);        // (a prefetch never faults)
        _mm_prefetch((const char*)(s + i) + kPackPrefetchBytes + 64, _MM_HINT_T0);
        const __m256d a0 = _mm256_loadu_pd(s + i), a1 = _mm256_loadu_pd(s + i + 4);
        const __m256d a2 = _mm256_loadu_pd(s + i + 8), a3 = _mm256_loadu_pd(s + i + 12);
        const __m128i i0 = _mm256_cvtpd_epi32(a0), i1 = _mm256_cvtpd_epi32(a1);      // NaN / overflow -> 0x80000000
        const __m128i i2 = _mm256_cvtpd_epi32(a2), i3 = _mm256_cvtpd_epi32(a3);
        bad = _mm256_or_pd(bad, _mm256_cmp_pd(_mm256_cvtepi32_pd(i0), a0, _CMP_NEQ_UQ));
        bad = _mm256_or_pd(bad, _mm256_cmp_pd(_mm256_cvtepi32_pd(i1), a1, _CMP_NEQ_UQ));
        bad = _mm256_or_pd(bad, _mm256_cmp_pd(_mm256_cvtepi32_pd(i2), a2, _CMP_NEQ_UQ));
        bad = _mm256_or_pd(bad, _mm256_cmp_pd(_mm256_cvtepi32_pd(i3), a3, _CMP_NEQ_UQ));
        // inside int16 <=> (i + 32768) >> 16 == 0
        range = _mm_or_si128(range, _mm_srai_epi32(_mm_add_epi32(i0, bias), 16));
        range = _mm_or_si128(range, _mm_srai_epi32(_mm_add_epi32(i1, bias), 16));
        range = _mm_or_si128(range, _mm_srai_epi32(_mm_add_epi32(i2, bias), 16));
        range = _mm_or_si128(range, _mm_srai_epi32(_mm_add_epi32(i3, bias), 16));
        if (nt) {
            _mm_stream_si128((__m128i*)(d + i), _mm_packs_epi32(i0, i1));
            _mm_stream_si128((__m128i*)(d + i + 8), _mm_packs_epi32(i2, i3));
        } else {
            _mm_storeu_si128((__m128i*)(d + i), _mm_packs_epi32(i0, i1));
            _mm_storeu_si128((__m128i*)(d + i + 8), _mm_packs_epi32(i2, i3));
        }
    }
    if (nt) _mm_sfence();
    bool ok = _mm256_movemask_pd(bad) == 0 && _mm_testz_si128(range, range);
    if (i < n) ok = pack_scalar<double>(s + i, n - i, d + i) && ok;
    return ok;
}

__attribute__((target("avx2"))) bool pack_f32_avx2(const float* s, int64_t n, int16_t* d)
{
    __m256 bad = _mm256_setzero_ps();
    __m256i range = _mm256_setzero_si256();
    const __m256i bias = _mm256_set1_epi32(32768);
    int64_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m256 a0 = _mm256_loadu_ps(s + i), a1 = _mm256_loadu_ps(s + i + 8);
        const __m256i i0 = _mm256_cvtps_epi32(a0), i1 = _mm256_cvtps_epi32(a1);
        bad = _mm256_or_ps(bad, _mm256_cmp_ps(_mm256_cvtepi32_ps(i0), a0, _CMP_NEQ_UQ));
        bad = _mm256_or_ps(bad, _mm256_cmp_ps(_mm256_cvtepi32_ps(i1), a1, _CMP_NEQ_UQ));
        range = _mm256_or_si256(range, _mm256_srai_epi32(_mm256_add_epi32(i0, bias), 16));
        range = _mm256_or_si256(range, _mm256_srai_epi32(_mm256_add_epi32(i1, bias), 16));
        // packs works per 128-bit half: (i0.lo, i1.lo | i0.hi, i1.hi) -> reorder the 64-bit quarters to 0, 2, 1, 3
        const __m256i p = _mm256_permute4x64_epi64(_mm256_packs_epi32(i0, i1), 0xD8);
        _mm256_storeu_si256((__m256i*)(d + i), p);
    }
    bool ok = _mm256_movemask_ps(bad) == 0 && _mm256_testz_si256(range, range);
    if (i < n) ok = pack_scalar<float>(s + i, n - i, d + i) && ok;
    return ok;
}

bool pack_block(const void* src, int dtype, int64_t b, int64_t e, int16_t* dst, bool avx2)
{
    if (dtype == PNPB200_DTYPE_F64) {
        const double* s = (const double*)src + b;
        return avx2 ? pack_f64_avx2(s, e - b, dst + b) : pack_scalar<double>(s, e - b, dst + b);
    }
    const float* s = (const float*)src + b;
    return avx2 ? pack_f32_avx2(s, e - b, dst + b) : pack_scalar<float>(s, e - b, dst + b);
}

// Workers that outlive the call: a chunk is packed in ~1.3 ms, and starting / joining 15 threads for it cost 0.2 - 0.3 ms of
// that.  One job at a time (a process drives one pipeline per GPU); the pool is leaked on purpose -- its threads sleep on
// the condition variable until the process ends.
class PackPool {
public:
    void run(int n_threads, const std::function<void()>& job)
    {
        std::lock_guard<std::mutex> one_job(run_mu_);
        {
            std::unique_lock<std::mutex> lk(mu_);
            while ((int)workers_.size() < n_threads - 1) {
                const int id = (int)workers_.size();
                workers_.emplace_back([this, id]() { loop(id); });
                workers_.back().detach();
            }
            job_ = &job; participants_ = n_threads - 1; pending_ = participants_; ++generation_;
        }
        wake_.notify_all();
        job();
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this]() { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int id)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void()>* job = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                wake_.wait(lk, [&]() { return generation_ != seen; });
                seen = generation_;
                if (id < participants_) job = job_;
            }
            if (!job) continue;
            (*job)();
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::mutex run_mu_, mu_;
    std::condition_variable wake_, done_;
    std::vector<std::thread> workers_;
    const std::function<void()>* job_ = nullptr;
    int participants_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
};

PackPool& pack_pool()
{
    static std::mutex mu;
    static PackPool* pool = nullptr;        // never destroyed: see above
    static pid_t owner = 0;
    std::lock_guard<std::mutex> lk(mu);
    if (!pool || owner != getpid()) {       // a forked child has the parent's pool object and none of its threads
        pool = new PackPool;
        owner = getpid();
    }
    return *pool;
}

}  // namespace

extern "C" int pnpb200_pack_i16(int dtype, const void* src, int64_t n_values, int16_t* dst, int n_threads)
{
    if ((dtype != PNPB200_DTYPE_F64 && dtype != PNPB200_DTYPE_F32) || n_values < 0 || (n_values > 0 && (!src || !dst)))
        return PNPB200_EINVAL;
    if (n_values == 0) return 1;
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (n_threads < 1) n_threads = 1;
    // blocks of 32 Ki values handed out by an atomic counter (threads slowed down by the DMA traffic next to them
    // simply take fewer); everybody stops at the first block that does not pack exactly -- fractional pixels: the
    // chunk will travel as it is, no point in converting the rest
    const int64_t block = 1 << 15;
    const int64_t n_blocks = (n_values + block - 1) / block;
    if (n_blocks < n_threads) n_threads = (int)n_blocks;
    std::atomic<int64_t> next{0};
    std::atomic<int> exact{1};
    auto work = [&]() {
        for (;;) {
            const int64_t b = next.fetch_add(1, std::memory_order_relaxed);
            if (b >= n_blocks || !exact.load(std::memory_order_relaxed)) return;
            const int64_t lo = b * block, hi = (lo + block < n_values) ? lo + block : n_values;
            if (!pack_block(src, dtype, lo, hi, dst, avx2)) exact.store(0, std::memory_order_relaxed);
        }
    };
    pack_pool().run(n_threads, work);
    return exact.load();
}
