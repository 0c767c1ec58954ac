// pnpb200_writers.cpp -- host side of the result table: the per-problem CSV that
// TEST_TOOLBOX.write_result_to_csv (TEST_TOOLBOX.py:693-708) writes from the list of result dicts
// built by compare_result_and_generate_result_dict (:396-463), produced straight from the arrays the
// report kernel returns (no Python object per problem).  Formatting follows Python's csv module as
// the reference uses it: QUOTE_MINIMAL, "\r\n" line ends, floats as repr(float) (shortest digits
// that round-trip; fixed notation for 1e-4 <= |x| < 1e16, else d.ddde+XX), bools as True / False.
// Columns: every scalar, string, tuple and dict field of the result dict in the reference's order;
// the six ndarray fields (np_R_GT, np_R_est, np_R_err, np_t_GT_est, np_t_est, np_t_err -- NumPy's
// multi-line str() of a matrix in a CSV cell) are left out.
// Rows are formatted by several threads into per-chunk buffers and written in order.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pnpb200.h"

namespace {

// repr(float) of CPython (float_repr_style = 'short'), appended without temporaries
void append_repr(std::string& out, double v)
{
    if (std::isnan(v)) { out += "nan"; return; }
    if (std::isinf(v)) { out += (v < 0 ? "-inf" : "inf"); return; }
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest
    const char* p = buf;
    const char* end = r.ptr;
    if (*p == '-') { out += '-'; ++p; }
    const char* e = p;
    while (*e != 'e') ++e;
    char digits[24];
    int nd = 0;
    for (const char* c = p; c < e; ++c)
        if (*c != '.') digits[nd++] = *c;
    int exp10 = 0;
    for (const char* c = e + 2; c < end; ++c) exp10 = exp10 * 10 + (*c - '0');
    if (e[1] == '-') exp10 = -exp10;
    char o[48];
    int n = 0;
    if (exp10 >= -4 && exp10 < 16) {                      // fixed notation
        if (exp10 < 0) {
            o[n++] = '0'; o[n++] = '.';
            for (int k = 0; k < -exp10 - 1; ++k) o[n++] = '0';
            for (int k = 0; k < nd; ++k) o[n++] = digits[k];
        } else if (nd <= exp10 + 1) {
            for (int k = 0; k < nd; ++k) o[n++] = digits[k];
            for (int k = nd; k < exp10 + 1; ++k) o[n++] = '0';
            o[n++] = '.'; o[n++] = '0';
        } else {
            for (int k = 0; k <= exp10; ++k) o[n++] = digits[k];
            o[n++] = '.';
            for (int k = exp10 + 1; k < nd; ++k) o[n++] = digits[k];
        }
    } else {                                              // d.ddde+XX, at least two exponent digits
        o[n++] = digits[0];
        if (nd > 1) { o[n++] = '.'; for (int k = 1; k < nd; ++k) o[n++] = digits[k]; }
        o[n++] = 'e'; o[n++] = (exp10 < 0) ? '-' : '+';
        const int ae = exp10 < 0 ? -exp10 : exp10;
        if (ae >= 100) o[n++] = (char)('0' + ae / 100);
        o[n++] = (char)('0' + (ae / 10) % 10);
        o[n++] = (char)('0' + ae % 10);
    }
    out.append(o, (size_t)n);
}

// csv QUOTE_MINIMAL: quote when the field holds the delimiter, the quote char or a line break
void append_field(std::string& out, const std::string& f)
{
    if (f.find_first_of(",\"\r\n") == std::string::npos) { out += f; return; }
    out += '"';
    for (char c : f) { if (c == '"') out += '"'; out += c; }
    out += '"';
}

struct ClassSpec { const double* bins; int n_bins; const char* const* labels; };

// labels[np.digitize(value, bins)] (TEST_TOOLBOX.classify_drpy :239-247), bins ascending, right = False
const char* classify(const ClassSpec& c, double v)
{
    int k = 0;
    for (int i = 0; i < c.n_bins; ++i) k += (c.bins[i] <= v) ? 1 : 0;
    return c.labels[k];
}

const char* kHeader =
    "idx,file_name,drpy,class,fail_count,pass_count,is_depth_passed,is_roll_passed,is_pitch_passed,is_yaw_passed,"
    "distance_GT,t3_est,depth_err,abs_depth_err,roll_GT,roll_est,roll_err,abs_roll_err,pitch_GT,pitch_est,pitch_err,"
    "abs_pitch_err,yaw_GT,yaw_est,yaw_err,abs_yaw_err,res_norm,res_norm_1000x,res_norm_10000x_n_est,res_norm_10000x_n_GT,"
    "LM_GT_error_average_normalize,LM_GT_error_max_normalize,LM_GT_error_max_key,predict_LM_error_average_normalize,"
    "predict_LM_error_max_normalize,predict_LM_error_max_key,predict_GT_error_average_normalize,"
    "predict_GT_error_max_normalize,predict_GT_error_max_key\r\n";

}  // namespace

extern "C" {

int pnpb200_format_repr(double v, char* out, int out_size)
{
    if (!out || out_size < 32) return PNPB200_EINVAL;
    std::string s;
    append_repr(s, v);
    if ((int)s.size() + 1 > out_size) return PNPB200_EINVAL;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return PNPB200_OK;
}

int pnpb200_write_result_csv(const char* path, int append, int64_t B, int64_t idx0, const double* report,
                             const int32_t* flags, const int32_t* max_idx, const double* res_norm, const double* gt,
                             const char* const* key_names, int n_keys, const double* const* bins, const int32_t* n_bins,
                             const char* const* const* labels, int n_threads)
{
    if (!path || B < 0 || !report || !flags || !max_idx || !res_norm || !gt || !key_names || n_keys < 1 || !bins || !n_bins || !labels)
        return PNPB200_EINVAL;
    ClassSpec cls[4];
    for (int q = 0; q < 4; ++q) {
        if (!bins[q] || !labels[q] || n_bins[q] < 0) return PNPB200_EINVAL;
        cls[q].bins = bins[q]; cls[q].n_bins = n_bins[q]; cls[q].labels = labels[q];
    }
    FILE* f = std::fopen(path, append ? "ab" : "wb");
    if (!f) return PNPB200_EINVAL;
    if (!append) std::fputs(kHeader, f);
    if (n_threads < 1) {
        n_threads = (int)std::thread::hardware_concurrency();
        if (n_threads < 1) n_threads = 1;
        if (n_threads > 32) n_threads = 32;
    }
    const int64_t kChunk = 4096;                          // rows per formatting task
    const int64_t n_chunks = (B + kChunk - 1) / kChunk;
    int rc = PNPB200_OK;
    std::vector<std::string> text((size_t)n_threads);     // one buffer per slot, reused by every batch
    for (auto& t : text) t.reserve((size_t)kChunk * 800);
    for (int64_t c0 = 0; c0 < n_chunks && rc == PNPB200_OK; c0 += n_threads) {
        const int batch = (int)((n_chunks - c0 < n_threads) ? (n_chunks - c0) : n_threads);
        std::vector<std::thread> pool;
        for (int w = 0; w < batch; ++w) {
            pool.emplace_back([&, w]() {
                std::string& out = text[(size_t)w];
                out.clear();
                std::string tmp;
                tmp.reserve(160);
                const int64_t lo = (c0 + w) * kChunk, hi = (lo + kChunk < B) ? (lo + kChunk) : B;
                for (int64_t b = lo; b < hi; ++b) {
                    const double* rp = report + b * PNPB200_REPORT_WIDTH;
                    const double* g = gt + b * 4;
                    const double dist = rp[11], t3 = rp[10];
                    const char* cd = classify(cls[0], g[0] * 100.0);      // distance class is on centimetres
                    const char* cr = classify(cls[1], g[1]);
                    const char* cp = classify(cls[2], g[2]);
                    const char* cy = classify(cls[3], g[3]);
                    out += std::to_string(idx0 + b); out += ',';
                    out += "random_drpy_"; out += cd; out += '_'; out += cr; out += '_'; out += cp; out += '_'; out += cy; out += ',';
                    tmp = "(";                                        // str of the tuple (distance_GT, roll_GT, pitch_GT, yaw_GT)
                    append_repr(tmp, dist); tmp += ", "; append_repr(tmp, g[1]); tmp += ", ";
                    append_repr(tmp, g[2]); tmp += ", "; append_repr(tmp, g[3]); tmp += ")";
                    append_field(out, tmp); out += ',';
                    tmp = "{'distance': '"; tmp += cd; tmp += "', 'roll': '"; tmp += cr; tmp += "', 'pitch': '"; tmp += cp;
                    tmp += "', 'yaw': '"; tmp += cy; tmp += "'}";
                    append_field(out, tmp); out += ',';
                    const int32_t* fl = flags + b * 4;
                    const int pass = (fl[0] != 0) + (fl[1] != 0) + (fl[2] != 0) + (fl[3] != 0);
                    out += std::to_string(4 - pass); out += ','; out += std::to_string(pass); out += ',';
                    for (int q = 0; q < 4; ++q) { out += fl[q] ? "True" : "False"; out += ','; }
                    const double est[4] = { t3, rp[12], rp[13], rp[14] };
                    const double ref[4] = { dist, g[1], g[2], g[3] };
                    for (int q = 0; q < 4; ++q) {                     // GT, est, err, abs err
                        append_repr(out, ref[q]); out += ',';
                        append_repr(out, est[q]); out += ',';
                        append_repr(out, rp[q]); out += ',';
                        append_repr(out, std::fabs(rp[q])); out += ',';
                    }
                    const double rn = res_norm[b];
                    append_repr(out, rn); out += ',';
                    append_repr(out, rn * 1000.0); out += ',';
                    append_repr(out, rn * 1000.0 * t3); out += ',';
                    append_repr(out, rn * 1000.0 * dist); out += ',';
                    for (int k = 0; k < 3; ++k) {
                        append_repr(out, rp[4 + 2 * k]); out += ',';
                        append_repr(out, rp[5 + 2 * k]); out += ',';
                        const int mi = max_idx[b * 3 + k];
                        if (mi >= 0 && mi < n_keys) append_field(out, key_names[mi]);   // None -> empty field
                        out += (k < 2) ? "," : "\r\n";
                    }
                }
            });
        }
        for (auto& t : pool) t.join();
        for (int w = 0; w < batch; ++w)
            if (std::fwrite(text[(size_t)w].data(), 1, text[(size_t)w].size(), f) != text[(size_t)w].size()) rc = PNPB200_EINVAL;
    }
    if (std::fclose(f) != 0) rc = PNPB200_EINVAL;
    return rc;
}

}  // extern "C"
