// pnpb200_api.cu -- the C ABI of the solve path (include/pnpb200.h): argument checks, dispatch to the
// per-(dtype, method group) objects of pnpb200_kernels.cu, per-kernel profiling, and the
// host-buffer pipeline.
#include <atomic>
#include <map>
#include <mutex>
#include <string.h>
#include <vector>

#include "pnpb200_common.cuh"

namespace pnpb200 {

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_last_error[512] = "";

void set_last_error(const char* where, cudaError_t e)
{
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

int get_device_props(DeviceProps* out)
{
    static std::mutex mu;
    static std::vector<DeviceProps> cache;
    int dev = 0;
    PNP_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    for (const DeviceProps& p : cache)
        if (p.device == dev) { *out = p; return PNPB200_OK; }
    DeviceProps p;
    p.device = dev;
    PNP_CUDA_OK(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev));
    PNP_CUDA_OK(cudaDeviceGetAttribute(&p.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    PNP_CUDA_OK(cudaDeviceGetAttribute(&p.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    PNP_CUDA_OK(cudaDeviceGetAttribute(&p.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    size_t free_b = 0;
    PNP_CUDA_OK(cudaMemGetInfo(&free_b, &p.total_mem));
    {   // keep stream-ordered scratch cached in the device's default pool instead of returning it
        // to the driver at every synchronisation (a moment-mapping call would re-map it each time)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        cudaGetLastError();
    }
    cache.push_back(p);
    *out = p;
    return PNPB200_OK;
}


thread_local ProfileRing g_prof;

static std::atomic<long long> g_launches{0};
void count_kernel_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

namespace {
struct KernelKey {
    int dev; const void* fn; int block; size_t smem;
    bool operator<(const KernelKey& o) const
    {
        if (dev != o.dev) return dev < o.dev;
        if (fn != o.fn) return fn < o.fn;
        if (block != o.block) return block < o.block;
        return smem < o.smem;
    }
};
}  // namespace

int make_row_tensor_map(CUtensorMap* out, const void* uv, int elem_bytes, long long B, int n_total, int chunk_points)
{
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static const EncodeFn encode = []() -> EncodeFn {
        if (getenv("PNPB200_NO_TENSOR_MAP")) return nullptr;                  // tuning / A-B comparisons only
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
        return (EncodeFn)fn;
    }();
    const unsigned long long row_bytes = (unsigned long long)n_total * 2ull * (unsigned long long)elem_bytes;
    if (!encode || B < 1 || B > 0x7fffffffLL || chunk_points < 1 || chunk_points * 2 > 256 || (row_bytes & 15ull) ||
        ((unsigned long long)chunk_points * 2ull * elem_bytes & 15ull) || ((uintptr_t)uv & 15u))
        return 0;
    const cuuint64_t dims[2] = { (cuuint64_t)n_total * 2u, (cuuint64_t)B };
    const cuuint64_t strides[1] = { (cuuint64_t)row_bytes };
    const cuuint32_t box[2] = { (cuuint32_t)chunk_points * 2u, 32u };
    const cuuint32_t estr[2] = { 1u, 1u };
    const CUresult r = encode(out, elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                              const_cast<void*>(uv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 1 : 0;
}

cudaError_t set_dynamic_smem(const void* kernel, size_t bytes)
{
    static std::mutex mu;
    static std::map<KernelKey, size_t> done;               // largest size set so far
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    size_t& cur = done[KernelKey{ dev, kernel, 0, 0 }];
    if (cur >= bytes && cur != 0) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

cudaError_t blocks_per_sm(int* out, const void* kernel, int block_threads, size_t smem_bytes)
{
    static std::mutex mu;
    static std::map<KernelKey, int> cache;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    const KernelKey key{ dev, kernel, block_threads, smem_bytes };
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return cudaSuccess; }
    int n = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, block_threads, smem_bytes);
    if (e == cudaSuccess) { cache[key] = n; *out = n; }
    return e;
}

void fill_default_params(pnpb200_params* p)
{
    p->max_it = 14; p->linear_it = 3;
    p->lm_lambda = 1e-5; p->exit_tol = 1e-2;
    p->f_weight = 225.68; p->meas_sigma_px = 3.0;
    p->proc_q = 1e-1; p->proc_d = 1e-2; p->omega0 = 1e-5; p->res_old0 = 1e-7;
    p->mapping = PNPB200_MAP_AUTO; p->flags = 0;
    p->workspace = nullptr; p->workspace_bytes = 0;
}


}  // namespace pnpb200

using namespace pnpb200;

extern "C" {

int pnpb200_version(void) { return PNPB200_VERSION; }
int64_t pnpb200_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
const char* pnpb200_last_error(void) { return g_last_error; }

int pnpb200_default_params(pnpb200_params* p)
{
    if (!p) return PNPB200_EINVAL;
    fill_default_params(p);
    return PNPB200_OK;
}

int pnpb200_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* hbm_bytes)
{
    DeviceProps dp;
    int rc = get_device_props(&dp);
    if (rc != PNPB200_OK) return rc;
    if (sm_count) *sm_count = dp.sm_count;
    if (cc_major) *cc_major = dp.cc_major;
    if (cc_minor) *cc_minor = dp.cc_minor;
    if (hbm_bytes) *hbm_bytes = (int64_t)dp.total_mem;
    return PNPB200_OK;
}

int pnpb200_profile_reset(void)
{
    g_prof.next = 0; g_prof.count = 0;
    if (g_prof.created)
        for (int s = 0; s < ProfileRing::kSlots; ++s) g_prof.used[s] = 0;
    return PNPB200_OK;
}

int pnpb200_profile_read(float* ms, int* n_calls)
{
    if (!ms) return PNPB200_EINVAL;
    double acc[3] = { 0, 0, 0 };
    int calls = 0;
    for (int s = 0; s < g_prof.count && g_prof.created; ++s) {
        const int u = g_prof.used[s];
        if (u < 2) continue;
        PNP_CUDA_OK(cudaEventSynchronize(g_prof.ev[s][u - 1]));
        for (int k = 0; k + 1 < u; ++k) {
            float t = 0.f;
            PNP_CUDA_OK(cudaEventElapsedTime(&t, g_prof.ev[s][k], g_prof.ev[s][k + 1]));
            acc[k] += t;
        }
        ++calls;
    }
    for (int k = 0; k < 3; ++k) ms[k] = calls ? (float)(acc[k] / calls) : 0.f;
    if (n_calls) *n_calls = calls;
    return PNPB200_OK;
}

int64_t pnpb200_workspace_bytes(int method, int dtype, int64_t B, int n_patterns, int mapping)
{
    const bool moment_form = (method == PNPB200_METHOD_LM || method == PNPB200_METHOD_LINEAR_F2 || method == PNPB200_METHOD_LM_PLUS) && n_patterns == 1;
    if (B <= 0 || !moment_form || (mapping != PNPB200_MAP_AUTO && mapping != PNPB200_MAP_MOMENT)) return 0;
    const int64_t esz = (dtype == PNPB200_DTYPE_F32) ? 4 : 8;
    return ((int64_t)(PNP_NMOM + PNP_NTAIL) * B + PNP_PATC) * esz;
}

int pnpb200_solve_batch(int method, int dtype, int64_t B, int n_total, int n, const void* uv, const void* pattern,
                        int n_patterns, const int32_t* point_index, const double* K, const pnpb200_params* params,
                        void* R, void* t, void* euler_deg, void* res_norm, int32_t* iters, int32_t* best_pattern,
                        void* stream)
{
    if (B < 0 || n_total < 1 || n < 1 || (!point_index && n != n_total) || !uv || !pattern || !K) return PNPB200_EINVAL;
    if (((uintptr_t)uv & 15u) != 0) return PNPB200_EINVAL;   // rows are staged with 16-byte bulk copies / vector loads
    if (n_patterns < 1 || n_patterns > PNPB200_MAX_PATTERNS) return PNPB200_EINVAL;
    if (method < 0 || method > 5 || (dtype != PNPB200_DTYPE_F64 && dtype != PNPB200_DTYPE_F32)) return PNPB200_EINVAL;
    if (B == 0) return PNPB200_OK;
    pnpb200_params prm;
    if (params) prm = *params; else fill_default_params(&prm);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* idx_dev = nullptr;
    if (point_index) {
        for (int i = 0; i < n; ++i)
            if (point_index[i] < 0 || point_index[i] >= n_total) return PNPB200_EINVAL;
        if (n > PNP_MAX_INLINE_IDX) {   // large selections go through device memory; small ones ride in the kernel arguments
            PNP_CUDA_OK(cudaMallocAsync((void**)&idx_dev, sizeof(int32_t) * (size_t)n, st));
            PNP_CUDA_OK(cudaMemcpyAsync(idx_dev, point_index, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
        }
    }
    const int group = (method == PNPB200_METHOD_QEIF || method == PNPB200_METHOD_LINEAR_F1) ? 0
                      : (method == PNPB200_METHOD_LINEAR_F2 ? 2 : (method == PNPB200_METHOD_EIF2 ? 3 : 1));
    typedef int (*part_fn)(PNP_SOLVE_PART_ARGS);
    static const part_fn parts[2][4] = { { solve_part_f64_g0, solve_part_f64_g1, solve_part_f64_g2, solve_part_f64_g3 },
                                         { solve_part_f32_g0, solve_part_f32_g1, solve_part_f32_g2, solve_part_f32_g3 } };
    const int rc = parts[dtype == PNPB200_DTYPE_F64 ? 0 : 1][group](method, B, n_total, n, uv, pattern, n_patterns, point_index,
                                                                   idx_dev, K, prm, R, t, euler_deg, res_norm, iters,
                                                                   best_pattern, st);
    if (idx_dev) cudaFreeAsync(idx_dev, st);
    return rc;
}

// ------------------------------------------------------------------------------------------
// host-buffer entry point: chunked, multi-stream H2D -> solve -> D2H pipeline
// ------------------------------------------------------------------------------------------
struct pnpb200_pipeline {
    int dtype, n_total, n_patterns, n_streams;
    int64_t chunk;
    size_t esz;
    std::vector<cudaStream_t> streams;
    std::vector<void*> d_uv, d_R, d_t, d_e, d_res;
    std::vector<int32_t*> d_it, d_best;
    std::vector<void*> d_ws;
    size_t ws_bytes;
    void* d_pattern;
};

int pnpb200_pipeline_create(pnpb200_pipeline** out, int dtype, int64_t chunk_problems, int n_total, int n_patterns,
                            int n_streams)
{
    if (!out || chunk_problems < 1 || n_total < 1 || n_patterns < 1 || n_patterns > PNPB200_MAX_PATTERNS) return PNPB200_EINVAL;
    if (dtype != PNPB200_DTYPE_F64 && dtype != PNPB200_DTYPE_F32) return PNPB200_EINVAL;
    if (n_streams < 1) n_streams = 3;
    pnpb200_pipeline* p = new pnpb200_pipeline();
    p->dtype = dtype; p->n_total = n_total; p->n_patterns = n_patterns; p->n_streams = n_streams;
    p->chunk = chunk_problems;
    p->esz = (dtype == PNPB200_DTYPE_F64) ? 8 : 4;
    p->d_pattern = nullptr;
    p->ws_bytes = ((size_t)(PNP_NMOM + PNP_NTAIL) * (size_t)chunk_problems + PNP_PATC) * p->esz;
    *out = p;
    PNP_CUDA_OK(cudaMalloc(&p->d_pattern, p->esz * (size_t)n_patterns * n_total * 3));
    for (int s = 0; s < n_streams; ++s) {
        cudaStream_t st;
        PNP_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        p->streams.push_back(st);
        void *a = nullptr, *b = nullptr, *c = nullptr, *d = nullptr, *e = nullptr, *f = nullptr, *g = nullptr;
        PNP_CUDA_OK(cudaMalloc(&a, p->esz * (size_t)chunk_problems * n_total * 2)); p->d_uv.push_back(a);
        PNP_CUDA_OK(cudaMalloc(&b, p->esz * (size_t)chunk_problems * 9)); p->d_R.push_back(b);
        PNP_CUDA_OK(cudaMalloc(&c, p->esz * (size_t)chunk_problems * 3)); p->d_t.push_back(c);
        PNP_CUDA_OK(cudaMalloc(&d, p->esz * (size_t)chunk_problems * 3)); p->d_e.push_back(d);
        PNP_CUDA_OK(cudaMalloc(&e, p->esz * (size_t)chunk_problems)); p->d_res.push_back(e);
        PNP_CUDA_OK(cudaMalloc(&f, sizeof(int32_t) * (size_t)chunk_problems)); p->d_it.push_back((int32_t*)f);
        PNP_CUDA_OK(cudaMalloc(&g, sizeof(int32_t) * (size_t)chunk_problems)); p->d_best.push_back((int32_t*)g);
        void* w = nullptr;
        PNP_CUDA_OK(cudaMalloc(&w, p->ws_bytes)); p->d_ws.push_back(w);
    }
    return PNPB200_OK;
}

int pnpb200_pipeline_destroy(pnpb200_pipeline* p)
{
    if (!p) return PNPB200_EINVAL;
    for (cudaStream_t st : p->streams) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (void* q : p->d_uv) cudaFree(q);
    for (void* q : p->d_R) cudaFree(q);
    for (void* q : p->d_t) cudaFree(q);
    for (void* q : p->d_e) cudaFree(q);
    for (void* q : p->d_res) cudaFree(q);
    for (int32_t* q : p->d_it) cudaFree(q);
    for (int32_t* q : p->d_best) cudaFree(q);
    for (void* q : p->d_ws) cudaFree(q);
    if (p->d_pattern) cudaFree(p->d_pattern);
    delete p;
    return PNPB200_OK;
}

int pnpb200_solve_batch_host(pnpb200_pipeline* p, int method, int64_t B, int n, const void* uv_host,
                             const void* pattern_host, const int32_t* point_index, const double* K,
                             const pnpb200_params* params, void* R, void* t, void* euler_deg, void* res_norm,
                             int32_t* iters, int32_t* best_pattern)
{
    if (!p || !uv_host || !pattern_host || !K || B < 0) return PNPB200_EINVAL;
    const size_t esz = p->esz;
    PNP_CUDA_OK(cudaMemcpyAsync(p->d_pattern, pattern_host, esz * (size_t)p->n_patterns * p->n_total * 3,
                                cudaMemcpyHostToDevice, p->streams[0]));
    PNP_CUDA_OK(cudaStreamSynchronize(p->streams[0]));
    int64_t done = 0;
    int s = 0;
    while (done < B) {
        const int64_t nb = (B - done < p->chunk) ? (B - done) : p->chunk;
        cudaStream_t st = p->streams[s];
        const char* src = (const char*)uv_host + esz * (size_t)done * p->n_total * 2;
        PNP_CUDA_OK(cudaMemcpyAsync(p->d_uv[s], src, esz * (size_t)nb * p->n_total * 2, cudaMemcpyHostToDevice, st));
        pnpb200_params prm;
        if (params) prm = *params; else fill_default_params(&prm);
        if (!prm.workspace) { prm.workspace = p->d_ws[s]; prm.workspace_bytes = (int64_t)p->ws_bytes; }
        int rc = pnpb200_solve_batch(method, p->dtype, nb, p->n_total, n, p->d_uv[s], p->d_pattern, p->n_patterns,
                                     point_index, K, &prm, R ? p->d_R[s] : nullptr, t ? p->d_t[s] : nullptr,
                                     euler_deg ? p->d_e[s] : nullptr, res_norm ? p->d_res[s] : nullptr,
                                     iters ? p->d_it[s] : nullptr, best_pattern ? p->d_best[s] : nullptr, st);
        if (rc != PNPB200_OK) return rc;
        if (R) PNP_CUDA_OK(cudaMemcpyAsync((char*)R + esz * (size_t)done * 9, p->d_R[s], esz * (size_t)nb * 9, cudaMemcpyDeviceToHost, st));
        if (t) PNP_CUDA_OK(cudaMemcpyAsync((char*)t + esz * (size_t)done * 3, p->d_t[s], esz * (size_t)nb * 3, cudaMemcpyDeviceToHost, st));
        if (euler_deg) PNP_CUDA_OK(cudaMemcpyAsync((char*)euler_deg + esz * (size_t)done * 3, p->d_e[s], esz * (size_t)nb * 3, cudaMemcpyDeviceToHost, st));
        if (res_norm) PNP_CUDA_OK(cudaMemcpyAsync((char*)res_norm + esz * (size_t)done, p->d_res[s], esz * (size_t)nb, cudaMemcpyDeviceToHost, st));
        if (iters) PNP_CUDA_OK(cudaMemcpyAsync(iters + done, p->d_it[s], sizeof(int32_t) * (size_t)nb, cudaMemcpyDeviceToHost, st));
        if (best_pattern) PNP_CUDA_OK(cudaMemcpyAsync(best_pattern + done, p->d_best[s], sizeof(int32_t) * (size_t)nb, cudaMemcpyDeviceToHost, st));
        done += nb;
        s = (s + 1) % p->n_streams;
        // a stream's buffers are reused n_streams chunks later; stream order protects them
    }
    for (cudaStream_t st : p->streams) PNP_CUDA_OK(cudaStreamSynchronize(st));
    return PNPB200_OK;
}

}  // extern "C"
