// bed_scan.cpp -- host side of the BED -> TSV scan: PLINK readers, batch loop, TSV formatting.
//
// Replaces, for the B200 path:
//   run_unified_bed_scan_to_tsv_common (producer/consumer loop)   src/stats/lmm.rs:975-1477
//   read_fam / BimChunkReader / parse_bim_line                     src/io/gfcore.rs:112-324, 1426-1478
//   append_assoc_row_from_fields + header schemas                  src/io/assoc2tsv.rs:45-57, 430-517
//   AsyncTsvWriter (writer thread)                                 src/stats/common.rs:374-468
//   resolve_snp_name                                               src/stats/lmm.rs:1952-1958
//
// The device work of each batch is jxb_scan_staged (K1 -> K2 -> K3).  Three stages overlap, like the reference's
// double buffer (src/io/pipeline.rs:47-92): a producer thread parses the BIM rows of batch i+2 and copies its packed
// rows from the memory-mapped BED into a pinned ring slot; batch i+1 goes up to the second device buffer on a copy
// stream (jxb_stage_packed) while batch i computes; a writer thread formats and writes batch i-1.  The ring is the
// only host memory that scales with the batch, and `mmap_window_mb` (the CLI's -mem) bounds it
// (WindowedBedMatrix, src/io/gload.rs:523-800).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/jxb200.h"

namespace jxb {
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
}  // namespace jxb

namespace {

// ---- number formatting ------------------------------------------------------------------------------------------------
// The writer prints ten numbers per row; at more than a million rows per second and GPU the C library's exact-decimal
// printf (~0.6 us per number) is the bottleneck of the file-level scan.  fmt_fixed / fmt_exp therefore scale the value by
// an exactly representable power of ten, keep the rounding error of that one operation (fma), and round the exact
// scaled value to an integer -- the correctly rounded decimal, i.e. the digits printf prints -- whenever the exact
// value is clearly off a rounding boundary.  Ties, near-ties, exponents beyond 10^22 and non-finite values take the printf
// route (`*_slow`), so the output is printf's in every case (tests compare 10^7 values per format).

// Rust `{:.N}`: C "%.Nf" with Rust's spellings of the non-finite values
size_t fmt_fixed_slow(char* buf, size_t cap, double v, int prec) {
    if (std::isnan(v)) return (size_t)snprintf(buf, cap, "NaN");
    if (std::isinf(v)) return (size_t)snprintf(buf, cap, v > 0 ? "inf" : "-inf");
    return (size_t)snprintf(buf, cap, "%.*f", prec, v);
}

// Rust `{:.Ne}`: mantissa as C "%.Ne", exponent without sign padding or leading zeros
size_t fmt_exp_slow(char* buf, size_t cap, double v, int prec) {
    if (std::isnan(v)) return (size_t)snprintf(buf, cap, "NaN");
    if (std::isinf(v)) return (size_t)snprintf(buf, cap, v > 0 ? "inf" : "-inf");
    char tmp[64];
    int len = snprintf(tmp, sizeof tmp, "%.*e", prec, v);
    int epos = len - 1;
    while (epos > 0 && tmp[epos] != 'e') --epos;
    const int ex = atoi(tmp + epos + 1);
    tmp[epos] = '\0';
    return (size_t)snprintf(buf, cap, "%se%d", tmp, ex);
}

const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};   // exact doubles

// nearest integer of the exact real a * 10^k (k in [-22, 22], result < 2^52); false when the exact value lies within
// 2^-20 of a rounding boundary (ties included) -- the caller then lets printf decide
inline bool scaled_round(double a, int k, uint64_t* q) {
    if (k > 22 || k < -22) return false;
    double t, frac_err;
    if (k >= 0) {
        t = a * kPow10[k];
        frac_err = std::fma(a, kPow10[k], -t);                 // a * 10^k = t + frac_err exactly
    } else {
        t = a / kPow10[-k];
        frac_err = std::fma(-t, kPow10[-k], a) / kPow10[-k];   // a / 10^-k = t + (a - t * 10^-k) / 10^-k; remainder exact
    }
    if (!(t < 4503599627370496.0)) return false;
    const double fl = std::floor(t);
    const double r = (t - fl) + frac_err;                      // fractional part of the exact value (|frac_err| <= ulp(t)/2)
    if (std::fabs(r - 0.5) < 9.5367431640625e-07) return false;
    *q = (uint64_t)fl + (r > 0.5 ? 1u : 0u);
    return true;
}

inline char* put_uint(char* w, uint64_t v) {                   // decimal digits of v, no padding
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *w++ = tmp[--n];
    return w;
}

size_t fmt_fixed(char* buf, size_t cap, double v, int prec) {
    const double a = std::fabs(v);
    uint64_t q;
    if (!(a < 1e11) || prec < 1 || prec > 9 || cap < 40 || !scaled_round(a, prec, &q)) return fmt_fixed_slow(buf, cap, v, prec);
    char* w = buf;
    if (std::signbit(v)) *w++ = '-';
    const uint64_t scale = (uint64_t)kPow10[prec];
    w = put_uint(w, q / scale);
    *w++ = '.';
    uint64_t f = q % scale;
    for (int i = prec - 1; i >= 0; --i) { w[i] = (char)('0' + f % 10); f /= 10; }
    w += prec;
    *w = '\0';
    return (size_t)(w - buf);
}

size_t fmt_exp(char* buf, size_t cap, double v, int prec) {
    const double a = std::fabs(v);
    if (!(a >= 1e-17 && a < 1e21) || prec < 1 || prec > 9 || cap < 40) return fmt_exp_slow(buf, cap, v, prec);
    // decimal exponent E with 10^E <= a < 10^(E+1): estimate from the binary exponent, then settle it exactly
    int e2;
    (void)std::frexp(a, &e2);
    int E = (int)std::floor((e2 - 1) * 0.30102999566398120);
    auto at_least_pow10 = [&](int e) {                          // a >= 10^e, exactly
        if (e >= 0) return a >= kPow10[e];
        const double t = a * kPow10[-e];
        const double err = std::fma(a, kPow10[-e], -t);
        return t > 1.0 || (t == 1.0 && err >= 0.0);
    };
    if (E < -22 || E > 21) return fmt_exp_slow(buf, cap, v, prec);
    if (!at_least_pow10(E)) --E;
    else if (E + 1 <= 22 && E + 1 >= -22 && at_least_pow10(E + 1)) ++E;
    if (E < -22 || E > 21) return fmt_exp_slow(buf, cap, v, prec);
    uint64_t q;
    if (!scaled_round(a, prec - E, &q)) return fmt_exp_slow(buf, cap, v, prec);
    const uint64_t top = (uint64_t)kPow10[prec + 1];
    if (q >= top) { q /= 10; ++E; }                             // 9.99996 -> 1.0000e(E+1): q == 10^(prec+1) exactly
    if (q < top / 10) return fmt_exp_slow(buf, cap, v, prec);    // cannot happen; printf as the safety net
    char* w = buf;
    if (std::signbit(v)) *w++ = '-';
    const uint64_t scale = (uint64_t)kPow10[prec];
    *w++ = (char)('0' + q / scale);
    *w++ = '.';
    uint64_t f = q % scale;
    for (int i = prec - 1; i >= 0; --i) { w[i] = (char)('0' + f % 10); f /= 10; }
    w += prec;
    *w++ = 'e';
    if (E < 0) { *w++ = '-'; w = put_uint(w, (uint64_t)(-E)); }
    else w = put_uint(w, (uint64_t)E);
    *w = '\0';
    return (size_t)(w - buf);
}

// Pinned staging buffers are kept between calls (page-locking hundreds of MB costs more than a small scan):
// a process-wide pool of at most kPoolMax idle buffers.
struct PinnedPool {
    static constexpr size_t kPoolMax = 4;
    std::mutex mu;
    std::vector<std::pair<uint8_t*, size_t>> idle;
    uint8_t* acquire(size_t bytes, size_t* cap) {
        {
            std::lock_guard<std::mutex> lk(mu);
            size_t best = idle.size();
            for (size_t i = 0; i < idle.size(); ++i)
                if (idle[i].second >= bytes && (best == idle.size() || idle[i].second < idle[best].second)) best = i;
            if (best != idle.size()) {
                uint8_t* p = idle[best].first;
                *cap = idle[best].second;
                idle.erase(idle.begin() + (long)best);
                return p;
            }
        }
        uint8_t* p = nullptr;
        if (cudaHostAlloc((void**)&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
            (void)cudaGetLastError();
            return nullptr;
        }
        *cap = bytes;
        return p;
    }
    void release(uint8_t* p, size_t cap) {
        if (!p) return;
        std::lock_guard<std::mutex> lk(mu);
        if (idle.size() < kPoolMax) { idle.emplace_back(p, cap); return; }
        // keep the larger buffers
        size_t smallest = 0;
        for (size_t i = 1; i < idle.size(); ++i) if (idle[i].second < idle[smallest].second) smallest = i;
        if (idle[smallest].second < cap) { cudaFreeHost(idle[smallest].first); idle[smallest] = {p, cap}; }
        else cudaFreeHost(p);
    }
};
PinnedPool& pinned_pool() { static PinnedPool* g = new PinnedPool(); return *g; }

// One BIM row as views into the memory-mapped .bim (no per-token allocation: at > 1 M SNPs/s the producer thread must
// parse a line in well under a microsecond).  The mapping outlives every Site (it is released after the writer joins).
struct Str {
    const char* p = "";
    uint32_t n = 0;
};
struct Site {
    Str chrom, snp, a0, a1;
    int64_t pos = 0;
};

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// Rust str::parse::<i32>(): optional sign, decimal digits only, range-checked; failure -> 0
int64_t parse_i32_or_zero(const char* s, size_t n) {
    if (n == 0) return 0;
    size_t i = 0;
    bool neg = false;
    if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
    if (i >= n) return 0;
    int64_t v = 0;
    for (; i < n; ++i) {
        if (s[i] < '0' || s[i] > '9') return 0;
        v = v * 10 + (s[i] - '0');
        if (v > 2147483648LL) return 0;
    }
    v = neg ? -v : v;
    if (v < -2147483648LL || v > 2147483647LL) return 0;
    return v;
}

bool simple_snp_allele(const char* a, size_t n) {
    size_t b = 0, e = n;
    while (b < e && is_ws(a[b])) ++b;
    while (e > b && is_ws(a[e - 1])) --e;
    if (e - b != 1) return false;
    const char c = (char)toupper((unsigned char)a[b]);
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

std::vector<std::string> split_ws(const std::string& line) {
    std::vector<std::string> out;
    size_t i = 0, n = line.size();
    while (i < n) {
        while (i < n && isspace((unsigned char)line[i])) ++i;
        size_t j = i;
        while (j < n && !isspace((unsigned char)line[j])) ++j;
        if (j > i) out.emplace_back(line.substr(i, j - i));
        i = j;
    }
    return out;
}

// BimChunkReader / parse_bim_line (src/io/gfcore.rs:112-302, 1426-1478) over a read-only mapping of the file
struct BimReader {
    std::string path;
    const char* base = nullptr;
    size_t size = 0, off = 0;
    size_t next_row = 0;
    bool open(const std::string& prefix) {
        path = prefix + ".bim";
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); return false; }
        size = (size_t)st.st_size;
        if (size) {
            void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { ::close(fd); return false; }
            base = (const char*)m;
            madvise(m, size, MADV_SEQUENTIAL);
        }
        ::close(fd);
        return true;
    }
    ~BimReader() { if (base && size) munmap((void*)base, size); }
    // returns 0 ok, 1 EOF, -1 malformed
    int next(Site& s, std::string& err) {
        if (off >= size) return 1;
        const char* line = base + off;
        const char* eol = (const char*)memchr(line, '\n', size - off);
        const size_t len = eol ? (size_t)(eol - line) : size - off;
        off += len + (eol ? 1 : 0);
        ++next_row;
        Str tok[6];
        int nt = 0;
        size_t i = 0;
        while (i < len && nt < 6) {
            while (i < len && is_ws(line[i])) ++i;
            size_t j = i;
            while (j < len && !is_ws(line[j])) ++j;
            if (j > i) { tok[nt].p = line + i; tok[nt].n = (uint32_t)(j - i); ++nt; }
            i = j;
        }
        if (nt < 6) {
            size_t e = len;
            while (e > 0 && is_ws(line[e - 1])) --e;
            err = "Malformed BIM line at " + path + ":" + std::to_string(next_row) + ": " + std::string(line, e);
            return -1;
        }
        s.chrom = tok[0];
        s.snp = tok[1];
        s.pos = parse_i32_or_zero(tok[3].p, tok[3].n);
        s.a0 = tok[4];
        s.a1 = tok[5];
        return 0;
    }
};

size_t format_row_views(char* buf, size_t cap, const Site& s, float af, float miss_rate, const double* row, int out_cols);

struct Batch {
    std::vector<Site> sites;      // per source row
    std::vector<uint8_t> keep;
    std::vector<float> af;
    std::vector<int32_t> missing;
    std::vector<double> out;      // compacted
    size_t n_kept = 0;
    int out_cols = 3;
};

struct Writer {
    FILE* fp = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Batch*> q;
    bool done = false;
    bool io_error = false;
    size_t n_model = 0;
    size_t rows_written = 0;

    // Formatting is the serial part of the file-level scan at small n (1.3 M SNPs/s of device work at n = 5,000 vs
    // ~0.7 M rows/s for one formatter thread), so a batch is formatted by up to 8 threads, each into its own buffer
    // over a contiguous slice of kept rows, and the buffers are written in order.
    void run() {
        std::vector<std::string> parts;
        std::vector<uint32_t> kept_rows;
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const size_t max_threads = std::min<size_t>(8, std::max<size_t>(1, hw / 2));
        for (;;) {
            Batch* b = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return done || !q.empty(); });
                if (q.empty()) return;
                b = q.front();
                q.pop_front();
            }
            cv.notify_all();
            kept_rows.clear();
            for (size_t r = 0; r < b->keep.size(); ++r)
                if (b->keep[r]) kept_rows.push_back((uint32_t)r);
            const size_t k = kept_rows.size();
            const size_t nth = std::max<size_t>(1, std::min(max_threads, k / 2048));
            parts.assign(nth, std::string());
            auto work = [&](size_t t) {
                const size_t lo = k * t / nth, hi = k * (t + 1) / nth;
                std::string& text = parts[t];
                text.reserve((hi - lo) * 112);
                std::vector<char> buf(4096);   // grown on demand: BIM alleles (indels, SVs) have no length limit
                for (size_t i = lo; i < hi; ++i) {
                    const size_t r = kept_rows[i];
                    const Site& s = b->sites[r];
                    // lmm.rs:2667-2670: miss column = missing_count as f32 / n as f32
                    const float mr = n_model ? (float)b->missing[r] / (float)n_model : 0.0f;
                    size_t len = format_row_views(buf.data(), buf.size(), s, b->af[r], mr, b->out.data() + i * b->out_cols,
                                                  b->out_cols);
                    if (len > buf.size()) {
                        buf.resize(len);
                        len = format_row_views(buf.data(), buf.size(), s, b->af[r], mr, b->out.data() + i * b->out_cols,
                                               b->out_cols);
                    }
                    text.append(buf.data(), len);
                }
            };
            if (nth == 1) {
                work(0);
            } else {
                std::vector<std::thread> pool;
                for (size_t t = 1; t < nth; ++t) pool.emplace_back(work, t);
                work(0);
                for (auto& th2 : pool) th2.join();
            }
            for (const std::string& text : parts)
                if (!text.empty() && fwrite(text.data(), 1, text.size(), fp) != text.size()) io_error = true;
            rows_written += k;
            delete b;
        }
    }
    void push(Batch* b) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return q.size() < 4; });
            q.push_back(b);
        }
        cv.notify_all();
    }
    void finish() {
        {
            std::lock_guard<std::mutex> lk(mu);
            done = true;
        }
        cv.notify_all();
        if (th.joinable()) th.join();
    }
};

const char* header_for(int out_cols) {
    switch (out_cols) {
        case 4: return "chrom\tpos\tsnp\tallele0\tallele1\taf\tmiss\tbeta\tse\tchisq\tpwald\tplrt\n";
        case 6: return "chrom\tpos\tsnp\tallele0\tallele1\taf\tmiss\tbeta\tse\tchisq\tpwald\tlambda\tml\tplrt\n";
        default: return "chrom\tpos\tsnp\tallele0\tallele1\taf\tmiss\tbeta\tse\tchisq\tpwald\n";
    }
}

}  // namespace

// Bytes a row can need: the strings verbatim (chrom twice for the chrom_pos fallback name) plus nine numeric fields
// (a `{:.4}` rendering of a huge finite double is ~315 characters) plus separators.
static size_t format_row_need(size_t chrom, size_t snp, size_t a0, size_t a1) { return 2 * chrom + snp + a0 + a1 + 9 * 336 + 64; }

// Never writes past buf[cap-1].  Returns the row length when it fitted; otherwise a value > cap (the size that is
// guaranteed to fit) and the buffer content is unspecified -- callers retry with a larger buffer.
static size_t format_row_impl(char* buf, size_t cap, Str chrom, int64_t pos, Str snp, Str a0, Str a1, float af,
                              float miss_rate, const double* row, int out_cols, bool resolve_name) {
    const size_t need = format_row_need(chrom.n, snp.n, a0.n, a1.n);
    if (cap < need) return need > cap ? need : cap + 1;
    const double beta = row[0], se = row[1];
    const bool valid = std::isfinite(beta) && std::isfinite(se) && se > 0.0;
    // sanitize_assoc_pvalue, src/math/linalg.rs:99-108
    double pw = 1.0;
    if (valid) {
        const double p = row[2];
        if (std::isfinite(p)) pw = p < 2.2250738585072014e-308 ? 2.2250738585072014e-308 : (p > 1.0 ? 1.0 : p);
    }
    // chisq_from_beta_se_and_optional_plrt, src/math/linalg.rs:288-298
    double chisq = NAN;
    if (valid) { const double z = beta / se; chisq = z * z; }
    char* w = buf;
    char* end = buf + cap;
    // snprintf reports the untruncated length: clamp so `w` can never leave the buffer
    auto adv = [&](size_t k) { const size_t room = (size_t)(end - w) - 1; w += k < room ? k : room; };
    auto put_s = [&](Str t) { const size_t room = (size_t)(end - w) - 1; const size_t c = t.n < room ? t.n : room; memcpy(w, t.p, c); w += c; };
    auto tab = [&]() { if (w < end - 1) *w++ = '\t'; };
    put_s(chrom); tab();
    // (cap >= format_row_need(): 21 bytes of a decimal int64 always fit)
    auto put_pos = [&]() {
        if (pos < 0) { *w++ = '-'; w = put_uint(w, (uint64_t)0 - (uint64_t)pos); }
        else w = put_uint(w, (uint64_t)pos);
    };
    put_pos(); tab();
    if (resolve_name && (snp.n == 0 || (snp.n == 1 && snp.p[0] == '.'))) {
        put_s(chrom);
        if (w < end - 1) *w++ = '_';
        put_pos();
    } else {
        put_s(snp);
    }
    tab();
    put_s(a0); tab();
    put_s(a1); tab();
    adv(fmt_fixed(w, (size_t)(end - w), (double)af, 4)); tab();
    adv(fmt_fixed(w, (size_t)(end - w), (double)miss_rate, 4)); tab();
    adv(fmt_fixed(w, (size_t)(end - w), beta, 4)); tab();
    adv(fmt_fixed(w, (size_t)(end - w), se, 4)); tab();
    adv(fmt_exp(w, (size_t)(end - w), chisq, 4)); tab();
    adv(fmt_exp(w, (size_t)(end - w), pw, 4));
    if (out_cols == 4) {
        tab(); adv(fmt_exp(w, (size_t)(end - w), row[3], 4));
    } else if (out_cols == 6) {
        tab(); adv(fmt_exp(w, (size_t)(end - w), row[3], 6));
        tab(); adv(fmt_exp(w, (size_t)(end - w), row[4], 6));
        tab(); adv(fmt_exp(w, (size_t)(end - w), row[5], 4));
    }
    if (w < end - 1) *w++ = '\n';
    *w = '\0';
    return (size_t)(w - buf);
}

namespace {
size_t format_row_views(char* buf, size_t cap, const Site& s, float af, float miss_rate, const double* row, int out_cols) {
    return format_row_impl(buf, cap, s.chrom, s.pos, s.snp, s.a0, s.a1, af, miss_rate, row, out_cols, true);
}
}  // namespace

static Str cstr(const char* s) { Str t; t.p = s; t.n = (uint32_t)strlen(s); return t; }

extern "C" size_t jxb_format_row(char* buf, size_t cap, const char* chrom, int64_t pos, const char* snp,
                                 const char* a0, const char* a1, float af, float miss_rate, const double* row,
                                 int out_cols) {
    return format_row_impl(buf, cap, cstr(chrom), pos, cstr(snp), cstr(a0), cstr(a1), af, miss_rate, row, out_cols, true);
}

// transform_alleles_by_model, src/io/assoc2tsv.rs:117-137
static void alleles_by_model(const char* r, const char* a, int gm, std::string& a0, std::string& a1) {
    const std::string R(r), A(a);
    switch (gm) {
        case 1: a0 = R + R; a1 = R + A + "/" + A + A; break;           // dom
        case 2: a0 = R + A + "/" + R + R; a1 = A + A; break;           // rec
        case 3: a0 = R + R + "/" + A + A; a1 = R + A; break;           // het
        default: a0 = R; a1 = A; break;
    }
}

extern "C" size_t jxb_format_block(char* buf, size_t cap, size_t rows, const char* chrom, const int64_t* pos,
                                   const char* snp, const char* a0, const char* a1, const float* af,
                                   const float* miss_rate, const double* res, int out_cols, int genetic_model) {
    if (!(out_cols == 3 || out_cols == 4 || out_cols == 6) || genetic_model < 0 || genetic_model > 3) return 0;
    size_t used = 0;
    std::string t0, t1;
    std::vector<char> line;
    for (size_t r = 0; r < rows; ++r) {
        alleles_by_model(a0, a1, genetic_model, t0, t1);
        const size_t need = format_row_need(strlen(chrom), strlen(snp), t0.size(), t1.size());
        if (line.size() < need) line.resize(need);
        // write_chunk prints the caller's SNP names verbatim (no chrom_pos substitution)
        const size_t len = format_row_impl(line.data(), line.size(), cstr(chrom), pos[r], cstr(snp), cstr(t0.c_str()), cstr(t1.c_str()),
                                           af[r], miss_rate[r], res + r * (size_t)out_cols, out_cols, false);
        if (used + len <= cap) memcpy(buf + used, line.data(), len);
        used += len;
        chrom += strlen(chrom) + 1; snp += strlen(snp) + 1; a0 += strlen(a0) + 1; a1 += strlen(a1) + 1;
    }
    return used;
}

// Checksum of a host buffer at memory bandwidth (8 threads): the Python front end keys its resident-model cache on the
// content of U^T (1.6 GB at n = 20,000) on every call, so a matrix modified in place is never mistaken for the resident one.
extern "C" void jxb_host_checksum(const void* data, size_t bytes, uint64_t out2[2]) {
    const size_t words = bytes / 8;
    const uint64_t* w = (const uint64_t*)data;
    const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(8, words / (1u << 20)));
    std::vector<uint64_t> sum(nt, 0), mix(nt, 0);
    auto work = [&](unsigned t) {
        const size_t lo = words * t / nt, hi = words * (t + 1) / nt;
        uint64_t a = 0, x = 0;
        for (size_t i = lo; i < hi; ++i) { a += w[i] * (2 * (uint64_t)i + 1); x ^= w[i] + i; }   // position-dependent
        sum[t] = a; mix[t] = x;
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    uint64_t a = 0, x = 0;
    for (unsigned t = 0; t < nt; ++t) { a += sum[t]; x ^= mix[t]; }
    const uint8_t* tail = (const uint8_t*)data + words * 8;
    for (size_t i = 0; i < bytes % 8; ++i) a = a * 1099511628211ull + tail[i];
    out2[0] = a; out2[1] = x;
}

// fmt_fixed / fmt_exp against their printf routes on `count` doubles drawn to stress them: uniform bit patterns over the
// exponent range the columns see, short decimals (exact ties and near-ties at the printed precision), neighbours of powers
// of ten, f32 values, integers.  Returns the number of differing strings; the first one is described in `first_bad`.
extern "C" size_t jxb_selftest_format(size_t count, uint64_t seed, int prec, char* first_bad, size_t bad_cap) {
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    auto next = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    size_t bad = 0;
    char a[96], b[96];
    if (first_bad && bad_cap) first_bad[0] = '\0';
    for (size_t i = 0; i < count; ++i) {
        const uint64_t r = next();
        double v;
        switch (i % 8) {
            case 0: case 1: {   // random mantissa, binary exponent in [-80, 80]
                const int ex = (int)(next() % 161) - 80;
                v = std::ldexp(1.0 + (double)(r >> 12) / 4503599627370496.0, ex);
                break;
            }
            case 2: {           // short decimals: j / 10^d with d up to prec + 2 (ties at the printed precision are j = ...5)
                const int d = (int)(next() % (unsigned)(prec + 3));
                v = (double)(r % 20000000ull) / kPow10[d];
                break;
            }
            case 3: {           // dyadic fractions k / 2^s: exactly representable ties (0.03125 = 312.5e-4)
                v = (double)(r % 100000ull) / (double)(1ull << (next() % 12));
                break;
            }
            case 4: {           // neighbours of powers of ten
                const int e = (int)(next() % 41) - 20;
                const double p10 = e >= 0 ? kPow10[e] : 1.0 / kPow10[-e];
                v = p10;
                const int steps = (int)(r % 7) - 3;
                for (int q = 0; q < (steps < 0 ? -steps : steps); ++q) v = std::nextafter(v, steps < 0 ? 0.0 : 1e300);
                break;
            }
            case 5: v = (double)(float)std::ldexp(1.0 + (double)(r >> 41) / 8388608.0, (int)(next() % 40) - 30); break;   // f32
            case 6: v = (double)(r % 1000000000ull); break;
            default: {          // 9.9999x-type mantissas at every decade: carries into the next exponent
                const int e = (int)(next() % 31) - 15;
                const double m = 9.9 + (double)(r % 100000ull) / 1000000.0;
                v = e >= 0 ? m * kPow10[e] : m / kPow10[-e];
                break;
            }
        }
        if (next() & 1) v = -v;
        for (int kind = 0; kind < 2; ++kind) {
            if (kind == 0) { fmt_fixed(a, sizeof a, v, prec); fmt_fixed_slow(b, sizeof b, v, prec); }
            else { fmt_exp(a, sizeof a, v, prec); fmt_exp_slow(b, sizeof b, v, prec); }
            if (strcmp(a, b) != 0) {
                if (bad == 0 && first_bad && bad_cap) snprintf(first_bad, bad_cap, "%s of %.17g: fast '%s' printf '%s'", kind ? "exp" : "fixed", v, a, b);
                ++bad;
            }
        }
    }
    return bad;
}

extern "C" const char* jxb_tsv_header(int out_cols) {
    return (out_cols == 3 || out_cols == 4 || out_cols == 6) ? header_for(out_cols) : nullptr;
}

extern "C" int jxb_scan_bed_to_tsv(jxb_model* m, const jxb_bed_scan_cfg* cfg, size_t* rows_written,
                                   jxb_progress_cb cb, void* user) {
    using jxb::fail;
    if (!m || !cfg || !cfg->bed_prefix || !cfg->out_tsv) return fail(-2, "null argument");
    const std::string prefix = cfg->bed_prefix;

    // FAM (gfcore.rs:307-324)
    std::vector<std::string> fam;
    {
        std::ifstream f(prefix + ".fam");
        if (!f.good()) return fail(-20, "open " + prefix + ".fam: No such file or directory");
        std::string line;
        while (std::getline(f, line)) {
            auto tok = split_ws(line);
            if (tok.size() < 2) return fail(-21, "Malformed FAM line: " + line);
            fam.push_back(tok[1]);
        }
    }
    const size_t n_full = fam.size();
    if (n_full == 0) return fail(-22, "no samples in PLINK FAM");

    // sample mapping (lmm.rs:1010-1039)
    std::vector<int64_t> sidx;
    bool identity = true;
    size_t n = n_full;
    if (cfg->sample_ids) {
        std::unordered_map<std::string, size_t> pos;
        for (size_t i = 0; i < fam.size(); ++i) pos[fam[i]] = i;  // later duplicates win, like HashMap::collect
        sidx.resize(cfg->n_sample_ids);
        for (size_t k = 0; k < cfg->n_sample_ids; ++k) {
            auto it = pos.find(cfg->sample_ids[k]);
            if (it == pos.end()) return fail(-23, std::string("sample '") + cfg->sample_ids[k] + "' not found in PLINK FAM");
            sidx[k] = (int64_t)it->second;
        }
        n = sidx.size();
        identity = (n == n_full);
        for (size_t k = 0; identity && k < n; ++k) identity = (sidx[k] == (int64_t)k);
    }

    // BED (lmm.rs:1041-1076)
    const size_t bps = (n_full + 3) / 4;
    const std::string bed_path = prefix + ".bed";
    int fd = open(bed_path.c_str(), O_RDONLY);
    if (fd < 0) return fail(-24, "open " + bed_path + ": " + strerror(errno));
    struct stat st;
    fstat(fd, &st);
    const size_t fsize = (size_t)st.st_size;
    const uint8_t* map = fsize ? (const uint8_t*)mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    close(fd);
    if (fsize && map == MAP_FAILED) return fail(-25, "mmap " + bed_path + ": " + strerror(errno));
    auto unmap = [&]() { if (map && fsize) munmap((void*)map, fsize); };
    if (fsize < 3 || map[0] != 0x6C || map[1] != 0x1B || map[2] != 0x01) {
        unmap();
        return fail(-26, "only SNP-major BED supported");
    }
    const size_t data_len = fsize - 3;
    if (data_len % bps != 0) {
        unmap();
        return fail(-27, "BED payload length " + std::to_string(data_len) + " not a multiple of " + std::to_string(bps));
    }
    const size_t n_snps = data_len / bps;
    const size_t begin = cfg->snp_begin < n_snps ? cfg->snp_begin : n_snps;
    const size_t end = (cfg->snp_end == 0 || cfg->snp_end > n_snps) ? n_snps : cfg->snp_end;
    const uint8_t* payload = map + 3;

    BimReader bim;
    if (!bim.open(prefix)) { unmap(); return fail(-28, "open " + prefix + ".bim: No such file or directory"); }
    std::string err;
    Site skip;
    while (bim.next_row < begin) {
        int r = bim.next(skip, err);
        if (r != 0) {
            unmap();
            return fail(-29, r < 0 ? err : "BIM ended early: needed row " + std::to_string(begin) + " but only saw " +
                                               std::to_string(bim.next_row) + " rows from " + bim.path);
        }
    }

    jxb_solve_cfg solve = cfg->solve;
    const int mode = cfg->mode;
    // lmm.rs:2573-2575: a finite seed is clamped into [low, high]
    if (solve.has_init && mode != 2) {
        if (!std::isfinite(solve.init_log10_lbd)) solve.has_init = 0;
        else solve.init_log10_lbd = std::min(std::max(solve.init_log10_lbd, solve.low), solve.high);
    }
    if (mode == 1 && !solve.has_nullml) {
        // lmm.rs:2901-2924: fit the null ML with the same Brent settings
        double o2[2];
        int rc = jxb_ml_null(m, solve.low, solve.high, solve.max_iter, solve.tol, solve.has_init, solve.init_log10_lbd, o2);
        if (rc) { unmap(); return rc; }
        if (!std::isfinite(o2[1])) { unmap(); return fail(-30, "failed to optimize null ML for LMM2 unified scan"); }
        solve.has_nullml = 1;
        solve.nullml = o2[1];
    }
    const int out_cols = mode == 1 ? 6 : (solve.has_nullml ? 4 : 3);

    Writer wr;
    wr.fp = fopen(cfg->out_tsv, "wb");
    if (!wr.fp) { unmap(); return fail(-31, std::string("create ") + cfg->out_tsv + ": " + strerror(errno)); }
    static thread_local std::vector<char> iobuf;
    iobuf.resize(8u << 20);
    setvbuf(wr.fp, iobuf.data(), _IOFBF, iobuf.size());
    if (cfg->write_header) fputs(header_for(out_cols), wr.fp);
    wr.n_model = n;
    wr.th = std::thread([&wr] { wr.run(); });

    // prepared row metadata: scan only the listed rows (ascending), QC thresholds are not re-applied
    const bool prepared = cfg->row_indices != nullptr;
    jxb_qc_cfg qc = cfg->qc;
    if (prepared) {
        for (size_t i = 0; i < cfg->n_row_indices; ++i) {
            const int64_t v = cfg->row_indices[i];
            if (v < 0 || (size_t)v >= n_snps) { wr.finish(); fclose(wr.fp); unmap(); return fail(-34, "row_indices out of range"); }
            if (i && v < cfg->row_indices[i - 1]) {
                wr.finish(); fclose(wr.fp); unmap();
                return fail(-35, "prepared row_indices must be sorted in ascending BED order");
            }
        }
        qc.maf_thr = 0.0f; qc.miss_thr = 1.0f; qc.het_thr = 0.0f;
    }
    const size_t total = prepared ? cfg->n_row_indices : end - begin;
    size_t step = std::max<size_t>(1, std::min<size_t>(cfg->batch_rows ? cfg->batch_rows : 4096, std::max<size_t>(total, 1)));
    // pinned ring of kRing batch-sized slots: with a -mem window the batch shrinks so that the ring fits it
    constexpr int kRing = 3;
    if (cfg->mmap_window_mb) {
        const size_t budget = (cfg->mmap_window_mb << 20) / kRing / bps;
        step = std::max<size_t>(256, std::min(step, budget));
    }
    const size_t span_max = prepared ? std::min(step, end - begin) : std::min(step, std::max<size_t>(total, 1));
    uint8_t* ring[kRing] = {nullptr, nullptr, nullptr};
    size_t ring_cap[kRing] = {0, 0, 0};
    bool ring_free[kRing] = {true, true, true};
    auto free_ring = [&]() { for (int k = 0; k < kRing; ++k) if (ring[k]) { pinned_pool().release(ring[k], ring_cap[k]); ring[k] = nullptr; } };
    // a scan of fewer than kRing batches needs fewer slots.  Upper bound on the batch count: batches are BED-row spans of
    // at most `step` rows (prepared lists may leave most rows of a span unlisted, so the listed count says nothing)
    const size_t span_rows = prepared ? (end > begin ? end - begin : 0) : total;
    const size_t n_batches = (span_rows + step - 1) / std::max<size_t>(step, 1);
    for (int k = 0; k < kRing; ++k) {
        if ((size_t)k >= std::max<size_t>(n_batches, 1)) { ring_free[k] = false; continue; }
        ring[k] = pinned_pool().acquire(std::max<size_t>(span_max * bps, 16), &ring_cap[k]);
        if (!ring[k]) {
            free_ring();
            wr.finish(); fclose(wr.fp); unmap();
            return fail(-36, "pinned staging ring: cudaHostAlloc of " + std::to_string(span_max * bps) + " bytes failed");
        }
    }
    size_t next_emit = cfg->progress_every ? std::max<size_t>(1, std::min(cfg->progress_every, total)) : 0;
    // Producer thread (the reference's producer, src/io/pipeline.rs:49-92): parses the BIM rows of the next batches and
    // builds their masks while the device scans the current one; errors travel with the item and are raised on the
    // calling thread (the error message is thread-local).
    struct Prep {
        Batch* b = nullptr;
        size_t c0 = 0, rows = 0, listed = 0;
        std::vector<uint8_t> mask;
        std::vector<float> row_af;        // prepared metadata of the listed rows (by row of this batch)
        std::vector<uint8_t> row_flip;
        std::vector<int32_t> row_miss;    // missing COUNT recovered from the caller's rate; -1 = not listed
        bool has_mask = false, last = false;
        int slot = -1;                    // pinned ring slot holding this batch's packed rows
        int code = 0;
        std::string msg;
    };
    std::mutex pmu;
    std::condition_variable pcv;
    std::deque<Prep*> pq;
    bool stop = false;
    auto emit = [&](Prep* it) {
        std::unique_lock<std::mutex> lk(pmu);
        pcv.wait(lk, [&] { return stop || pq.size() < 2; });
        if (stop) { if (it->b) delete it->b; delete it; return false; }
        pq.push_back(it);
        lk.unlock();
        pcv.notify_all();
        return true;
    };
    std::thread producer([&] {
        size_t list_pos = 0;   // prepared mode: next entry of row_indices
        std::string perr;
        Site pskip;
        auto bail = [&](int code, const std::string& msg) {
            Prep* it = new Prep();
            it->last = true; it->code = code; it->msg = msg;
            emit(it);
        };
        for (size_t c0 = begin;;) {
            size_t rows;
            if (prepared) {
                // skip list entries before `begin`, stop at `end`
                while (list_pos < cfg->n_row_indices && (size_t)cfg->row_indices[list_pos] < begin) ++list_pos;
                if (list_pos >= cfg->n_row_indices || (size_t)cfg->row_indices[list_pos] >= end) break;
                c0 = (size_t)cfg->row_indices[list_pos];
                rows = std::min(step, end - c0);
            } else {
                if (c0 >= end) break;
                rows = std::min(step, end - c0);
            }
            Prep* it = new Prep();
            it->c0 = c0; it->rows = rows;
            Batch* b = it->b = new Batch();
            b->sites.resize(rows);
            b->keep.resize(rows);
            b->af.resize(rows);
            b->missing.resize(rows);
            b->out.resize(rows * out_cols);
            b->out_cols = out_cols;
            // packed rows of the batch: mmap (page cache / disk) -> a free pinned ring slot, copied by a helper thread
            // while this thread parses the batch's BIM rows
            {
                std::unique_lock<std::mutex> lk(pmu);
                pcv.wait(lk, [&] { return stop || ring_free[0] || ring_free[1] || ring_free[2]; });
                if (stop) { delete b; delete it; return; }
                for (int k = 0; k < kRing; ++k) if (ring_free[k]) { it->slot = k; ring_free[k] = false; break; }
            }
            std::thread copier([&, c0, rows] { memcpy(ring[it->slot], payload + c0 * bps, rows * bps); });
            // BIM is read sequentially: skip the lines between the previous batch and this one
            int bad = 0;
            size_t need = c0;
            while (!bad && bim.next_row < c0) bad = bim.next(pskip, perr);
            if (!bad) {
                need = c0 + rows;
                for (size_t r = 0; r < rows && !bad; ++r) bad = bim.next(b->sites[r], perr);
            }
            copier.join();
            if (bad) {
                { std::lock_guard<std::mutex> lk(pmu); ring_free[it->slot] = true; }
                delete b; delete it;
                bail(-29, bad < 0 ? perr : "BIM ended early: needed row " + std::to_string(need) + " but only saw " +
                                             std::to_string(bim.next_row) + " rows from " + bim.path);
                return;
            }
            if (cfg->snps_only || prepared) {
                it->has_mask = true;
                it->mask.assign(rows, prepared ? 0 : 1);
                if (prepared) {
                    if (cfg->row_maf) it->row_af.assign(rows, 0.0f);
                    if (cfg->row_flip) it->row_flip.assign(rows, 0);
                    if (cfg->row_missing) it->row_miss.assign(rows, -1);
                    while (list_pos < cfg->n_row_indices && (size_t)cfg->row_indices[list_pos] < c0 + rows) {
                        const size_t r = (size_t)cfg->row_indices[list_pos] - c0;
                        it->mask[r] = 1;
                        if (cfg->row_maf) it->row_af[r] = cfg->row_maf[list_pos];
                        if (cfg->row_flip) it->row_flip[r] = cfg->row_flip[list_pos] ? 1 : 0;
                        if (cfg->row_missing) {
                            // missing_count_from_rate, src/stats/lmm.rs:1934-1940
                            const float v = cfg->row_missing[list_pos];
                            it->row_miss[r] = (!std::isfinite(v) || v <= 0.0f) ? 0 : (int32_t)std::max(0.0, std::round((double)v * (double)n));
                        }
                        ++list_pos;
                        ++it->listed;
                    }
                }
                if (cfg->snps_only)
                    for (size_t r = 0; r < rows; ++r)
                        if (!simple_snp_allele(b->sites[r].a0.p, b->sites[r].a0.n) || !simple_snp_allele(b->sites[r].a1.p, b->sites[r].a1.n)) it->mask[r] = 0;
            }
            if (!emit(it)) return;
            c0 += rows;
        }
        if (end == n_snps && !prepared) {
            // gfcore.rs:265-280: the BIM must not hold more rows than the BED
            Site extra;
            if (bim.next(extra, perr) != 1) {
                bail(-32, "BIM site count exceeds BED SNP count: expected " + std::to_string(n_snps) +
                              ", saw extra row " + std::to_string(bim.next_row) + " in " + bim.path);
                return;
            }
        }
        bail(0, "");
    });

    int rc = 0;
    size_t scanned = 0;
    // JXB_BED_TIMING=1: where the calling thread's wall time goes (waiting for the producer, staging, device scan,
    // handing rows to the writer), printed to stderr at the end
    const bool timing = getenv("JXB_BED_TIMING") != nullptr;
    double t_pop = 0, t_stage = 0, t_scan = 0, t_push = 0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    auto pop = [&]() {
        const double t0 = now();
        Prep* it = nullptr;
        std::unique_lock<std::mutex> lk(pmu);
        pcv.wait(lk, [&] { return !pq.empty(); });
        it = pq.front();
        pq.pop_front();
        lk.unlock();
        pcv.notify_all();
        t_pop += now() - t0;
        return it;
    };
    auto release_slot = [&](Prep* it) {
        if (it && it->slot >= 0) {
            { std::lock_guard<std::mutex> lk(pmu); ring_free[it->slot] = true; }
            it->slot = -1;
            pcv.notify_all();
        }
    };
    auto drop = [&](Prep* it) { if (it) { release_slot(it); if (it->b) delete it->b; delete it; } };
    // cur = the batch on the device; nxt = the batch going up while cur computes
    Prep* cur = pop();
    if (!cur->last) rc = jxb_stage_packed(m, ring[cur->slot], bps, cur->rows);
    while (!rc && !cur->last) {
        Prep* nxt = pop();
        Batch* b = cur->b;
        // jxb_scan_staged waits for cur's copy and makes it the working buffer; nxt is staged right after the swap is
        // ordered, i.e. from inside the same call sequence: stage it first into the OTHER buffer -- the library keeps
        // two device buffers and swaps them at every scan
        double t0 = now();
        rc = jxb_scan_staged_begin(m);
        if (!rc && !nxt->last) rc = jxb_stage_packed(m, ring[nxt->slot], bps, nxt->rows);
        t_stage += now() - t0;
        t0 = now();
        if (!rc)
            rc = jxb_scan_staged(m, n_full, identity ? nullptr : sidx.data(), cur->has_mask ? cur->mask.data() : nullptr,
                                 cur->row_af.empty() ? nullptr : cur->row_af.data(),
                                 cur->row_flip.empty() ? nullptr : cur->row_flip.data(), &qc, &solve, mode, b->keep.data(),
                                 b->af.data(), b->missing.data(), b->out.data(), nullptr, &b->n_kept);
        if (rc) { drop(nxt); break; }
        if (!cur->row_miss.empty())
            for (size_t r = 0; r < cur->rows; ++r)
                if (cur->row_miss[r] >= 0) b->missing[r] = cur->row_miss[r];
        t_scan += now() - t0;
        t0 = now();
        wr.push(b);
        t_push += now() - t0;
        cur->b = nullptr;
        scanned += prepared ? cur->listed : cur->rows;
        drop(cur);
        cur = nxt;
        if (cb && next_emit && scanned >= next_emit) {
            if (cb(scanned < total ? scanned : total, total, user) != 0) { rc = fail(-40, "interrupted by progress callback"); break; }
            next_emit = std::min((scanned / cfg->progress_every + 1) * cfg->progress_every, total);
        }
    }
    if (!rc && cur->last && cur->code) rc = fail(cur->code, cur->msg);
    drop(cur);
    jxb_stage_cancel(m);
    {
        // stop the producer (no-op when it already delivered its last item) and drop what it had queued
        std::lock_guard<std::mutex> lk(pmu);
        stop = true;
    }
    pcv.notify_all();
    producer.join();
    for (Prep* it : pq) { if (it->b) delete it->b; delete it; }
    pq.clear();
    free_ring();
    if (rc == 0 && cb) cb(total, total, user);
    const double t_loop = now();
    wr.finish();
    if (timing)
        fprintf(stderr, "[jxb bed scan] batches of %zu rows: wait-producer %.3f s, stage %.3f s, device scan %.3f s, push %.3f s, "
                        "loop %.3f s, writer drain %.3f s\n", step, t_pop, t_stage, t_scan, t_push, t_loop - t_begin, now() - t_loop);
    const bool ioerr = wr.io_error || fclose(wr.fp) != 0;
    unmap();
    if (rc) return rc;
    if (ioerr) return fail(-33, std::string("write ") + cfg->out_tsv + " failed");
    if (rows_written) *rows_written = wr.rows_written;
    return 0;
}
