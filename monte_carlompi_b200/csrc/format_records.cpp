// format_records.cpp -- native writer for the reference's output lines (host code, no CUDA).
//
// The reference writes one line per photon with '%d %r %r %r %d %r %r\n' % (condition, wvn, theta_n, phi_n, n_scat,
// path_length, snow_depth)  (monte_carloMPI/monte_carlo3D.py:1630-1636).  '%r' of a float is CPython's repr():
// the shortest decimal string that round-trips (David Gay's algorithm, mode 0), printed in fixed notation when
// -4 < decimal exponent <= 16 and in exponent notation otherwise, always with a '.0' if it would look like an
// integer (Python/pystrtod.c, format code 'r').  std::to_chars gives the same shortest digits; this file only
// re-lays them out the CPython way.  At 10^6+ photons the Python formatter costs more than the walk itself; this
// one runs at memory speed on all host cores and is byte-identical (tests/test_host_logic.py).
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mc3d.h"

namespace {

// Append repr(x) to p; returns the new end.  Needs up to 32 bytes.
char *py_repr(char *p, double x)
{
    if (std::isnan(x)) { memcpy(p, "nan", 3); return p + 3; }
    if (std::isinf(x)) { if (x < 0) *p++ = '-'; memcpy(p, "inf", 3); return p + 3; }
    if (std::signbit(x)) { *p++ = '-'; x = -x; }
    if (x == 0.0) { memcpy(p, "0.0", 3); return p + 3; }
    char sci[40];
    auto res = std::to_chars(sci, sci + sizeof sci, x, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest
    char *e = sci;
    while (*e != 'e') ++e;
    char digits[24];
    int nd = 0;
    for (char *c = sci; c < e; ++c)
        if (*c != '.') digits[nd++] = *c;
    int exp10 = 0;
    {
        const char *c = e + 1;
        const bool neg = (*c == '-');
        ++c;
        for (; c < res.ptr; ++c) exp10 = exp10 * 10 + (*c - '0');
        if (neg) exp10 = -exp10;
    }
    const int decpt = exp10 + 1;   // position of the decimal point relative to the digit string
    if (decpt <= -4 || decpt > 16) {   // exponent notation
        *p++ = digits[0];
        if (nd > 1) {
            *p++ = '.';
            memcpy(p, digits + 1, nd - 1);
            p += nd - 1;
        }
        *p++ = 'e';
        int ex = decpt - 1;
        *p++ = ex < 0 ? '-' : '+';
        if (ex < 0) ex = -ex;
        if (ex >= 100) { *p++ = char('0' + ex / 100); ex %= 100; *p++ = char('0' + ex / 10); *p++ = char('0' + ex % 10); }
        else { *p++ = char('0' + ex / 10); *p++ = char('0' + ex % 10); }
        return p;
    }
    if (decpt <= 0) {
        *p++ = '0';
        *p++ = '.';
        for (int k = 0; k < -decpt; ++k) *p++ = '0';
        memcpy(p, digits, nd);
        return p + nd;
    }
    if (decpt >= nd) {
        memcpy(p, digits, nd);
        p += nd;
        for (int k = 0; k < decpt - nd; ++k) *p++ = '0';
        *p++ = '.';
        *p++ = '0';
        return p;
    }
    memcpy(p, digits, decpt);
    p += decpt;
    *p++ = '.';
    memcpy(p, digits + decpt, nd - decpt);
    return p + (nd - decpt);
}

char *put_uint(char *p, uint64_t v)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = char('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

void format_range(uint64_t lo, uint64_t hi, const uint8_t *condition, const int16_t *wvl_row, const float *theta_n,
                  const float *phi_n, const uint32_t *n_scat, const float *path_length, const double *wvn_by_row,
                  const double *snow_depth_by_row, std::string &out)
{
    out.clear();
    out.reserve((hi - lo) * 112);
    char line[256];
    for (uint64_t k = lo; k < hi; ++k) {
        char *p = line;
        p = put_uint(p, condition[k]);
        *p++ = ' ';
        p = py_repr(p, wvn_by_row[wvl_row[k]]);
        *p++ = ' ';
        p = py_repr(p, (double)theta_n[k]);
        *p++ = ' ';
        p = py_repr(p, (double)phi_n[k]);
        *p++ = ' ';
        p = put_uint(p, n_scat[k]);
        *p++ = ' ';
        p = py_repr(p, (double)path_length[k]);
        *p++ = ' ';
        p = py_repr(p, snow_depth_by_row[wvl_row[k]]);
        *p++ = '\n';
        out.append(line, p - line);
    }
}

}  // namespace

extern "C" {

// Expand packed 16-byte records (mc3d_records.packed in include/mc3d.h) into columns.  Host code, threads over ranges.
int mc3d_unpack_records(const uint32_t *packed, uint64_t n, const mc3d_records *out, int n_threads)
{
    if (!out || (n && !packed)) return MC3D_EINVAL;
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<uint64_t>((uint64_t)n_threads, std::max<uint64_t>(1, n >> 16));
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t k = lo; k < hi; ++k) {
            const uint32_t m = packed[4 * k], t = packed[4 * k + 1], f = packed[4 * k + 2], l = packed[4 * k + 3];
            if (out->condition) out->condition[k] = (uint8_t)((t >> 31) | ((f >> 31) << 1) | ((l >> 31) << 2));
            if (out->wvl_row) out->wvl_row[k] = (int16_t)(m & 0x1ffu);
            if (out->n_scat) out->n_scat[k] = m >> 9;
            const uint32_t tb = t & 0x7fffffffu, fb = f & 0x7fffffffu, lb = l & 0x7fffffffu;
            if (out->theta_n) memcpy(&out->theta_n[k], &tb, 4);
            if (out->phi_n) memcpy(&out->phi_n[k], &fb, 4);
            if (out->path_length) memcpy(&out->path_length[k], &lb, 4);
        }
    };
    if (n_threads <= 1) { work(0, n); return MC3D_OK; }
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(work, n * t / n_threads, n * (t + 1) / n_threads);
    for (auto &t : th) t.join();
    return MC3D_OK;
}

// repr(x) of one double into buf (>= 32 bytes); returns the length.  For tests.
int mc3d_py_repr(double x, char *buf)
{
    char *end = py_repr(buf, x);
    *end = 0;
    return (int)(end - buf);
}

// Append the reference's text lines for photons [0, n) to `path` (created if missing; the caller writes the header).
// Columns as in mc3d_records; wvn and snow_depth are looked up by wvl_row.  n_threads <= 0: all host cores.
// Returns the number of bytes written, or a negative MC3D_E* code.
int64_t mc3d_write_records_text(const char *path, int append, uint64_t n, const uint8_t *condition, const int16_t *wvl_row,
                                const float *theta_n, const float *phi_n, const uint32_t *n_scat, const float *path_length,
                                const double *wvn_by_row, const double *snow_depth_by_row, int n_rows, int n_threads)
{
    if (!path || (n && (!condition || !wvl_row || !theta_n || !phi_n || !n_scat || !path_length || !wvn_by_row || !snow_depth_by_row)))
        return MC3D_EINVAL;
    for (uint64_t k = 0; k < n; ++k)
        if (wvl_row[k] < 0 || wvl_row[k] >= n_rows) return MC3D_EINVAL;
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return MC3D_EINVAL;
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const uint64_t block = 1u << 16;   // lines per work item; written in order
    int64_t total = 0;
    std::vector<std::string> bufs(n_threads);
    for (uint64_t base = 0; base < n; base += block * n_threads) {
        std::vector<std::thread> th;
        int used = 0;
        for (int t = 0; t < n_threads; ++t) {
            const uint64_t lo = base + (uint64_t)t * block;
            if (lo >= n) break;
            const uint64_t hi = std::min<uint64_t>(n, lo + block);
            ++used;
            th.emplace_back(format_range, lo, hi, condition, wvl_row, theta_n, phi_n, n_scat, path_length, wvn_by_row,
                            snow_depth_by_row, std::ref(bufs[t]));
        }
        for (auto &t : th) t.join();
        for (int t = 0; t < used; ++t) {
            if (fwrite(bufs[t].data(), 1, bufs[t].size(), f) != bufs[t].size()) { fclose(f); return MC3D_EINVAL; }
            total += (int64_t)bufs[t].size();
        }
    }
    if (fclose(f) != 0) return MC3D_EINVAL;
    return total;
}

}  // extern "C"
