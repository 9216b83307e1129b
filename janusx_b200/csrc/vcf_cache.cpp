// vcf_cache.cpp -- SURVEY 8(f) N3: VCF(.gz) -> PLINK BED/BIM/FAM, the conversion `jx gwas -vcf` performs once before an
// exact-LMM scan (python/janusx/assoc/workflow.py:2431-2477 decides when; the scan itself only ever sees the BED).
//
// Restates, for the B200 path's host side:
//   VcfSnpIter::new / next_snp_raw (header, field split, GT codes, SNP naming)   src/io/gfcore.rs:2875-2980
//   plink2bits_from_g_f32 (dosage -> 2-bit code; missing = 01)                   src/io/gfreader.rs:2630-2641
//   is_simple_snp_allele (the snps_only filter)                                  src/io/gfreader.rs:7013-7019
// GT rules are the reference's exact string matches: 0/0 0|0 -> 0, 0/1 1/0 0|1 1|0 -> 1, 1/1 1|1 -> 2, anything else
// (./., multi-allelic indices, haploid calls) -> missing.  Dosage counts ALT; BIM column 5 = REF, column 6 = ALT.
// zlib is opened with dlopen (gzopen reads plain text transparently), so the library has no link-time dependency.
#include <dlfcn.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jxb200.h"

namespace jxb {
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
}  // namespace jxb

namespace {

struct Zlib {
    void* so = nullptr;
    void* (*open)(const char*, const char*) = nullptr;
    int (*read)(void*, void*, unsigned) = nullptr;
    int (*close)(void*) = nullptr;
    int (*buffer)(void*, unsigned) = nullptr;
    bool ok = false;
};

Zlib& zlib() {
    static Zlib z;
    static bool tried = false;
    if (tried) return z;
    tried = true;
    for (const char* name : {"libz.so.1", "libz.so"}) {
        z.so = dlopen(name, RTLD_NOW);
        if (z.so) break;
    }
    if (!z.so) return z;
    z.open = (void* (*)(const char*, const char*))dlsym(z.so, "gzopen");
    z.read = (int (*)(void*, void*, unsigned))dlsym(z.so, "gzread");
    z.close = (int (*)(void*))dlsym(z.so, "gzclose");
    z.buffer = (int (*)(void*, unsigned))dlsym(z.so, "gzbuffer");
    z.ok = z.open && z.read && z.close;
    return z;
}

// line reader over gz or plain input
struct LineReader {
    void* gz = nullptr;
    FILE* fp = nullptr;
    std::vector<char> buf;
    size_t lo = 0, hi = 0;
    bool eof = false;
    bool open(const char* path, std::string& err) {
        const size_t len = strlen(path);
        const bool is_gz = len > 3 && strcmp(path + len - 3, ".gz") == 0;
        buf.resize(1 << 22);
        if (zlib().ok) {
            gz = zlib().open(path, "rb");
            if (!gz) { err = std::string("open ") + path + ": No such file or directory"; return false; }
            if (zlib().buffer) zlib().buffer(gz, 1 << 20);
            return true;
        }
        if (is_gz) { err = "libz.so.1 could not be loaded: cannot read " + std::string(path); return false; }
        fp = fopen(path, "rb");
        if (!fp) { err = std::string("open ") + path + ": No such file or directory"; return false; }
        return true;
    }
    size_t fill(char* dst, size_t cap) {
        if (gz) { const int r = zlib().read(gz, dst, (unsigned)cap); return r > 0 ? (size_t)r : 0; }
        return fread(dst, 1, cap, fp);
    }
    // returns false at end of input; the line excludes the terminator
    bool next(std::string& line) {
        line.clear();
        for (;;) {
            if (lo == hi) {
                if (eof) return !line.empty();
                hi = fill(buf.data(), buf.size());
                lo = 0;
                if (hi == 0) { eof = true; return !line.empty(); }
            }
            const char* p = (const char*)memchr(buf.data() + lo, '\n', hi - lo);
            if (p) {
                line.append(buf.data() + lo, p - (buf.data() + lo));
                lo = (size_t)(p - buf.data()) + 1;
                return true;
            }
            line.append(buf.data() + lo, hi - lo);
            lo = hi;
        }
    }
    void close() {
        if (gz) zlib().close(gz);
        if (fp) fclose(fp);
        gz = nullptr; fp = nullptr;
    }
};

bool simple_allele(const char* a, size_t len) {
    while (len && isspace((unsigned char)*a)) { ++a; --len; }
    while (len && isspace((unsigned char)a[len - 1])) --len;
    if (len != 1) return false;
    const char c = (char)toupper((unsigned char)*a);
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

// 2-bit PLINK code of one sample field ("GT[:...]")
inline uint8_t gt_code(const char* f, size_t len) {
    size_t g = 0;
    while (g < len && f[g] != ':') ++g;
    if (g == 3 && (f[1] == '/' || f[1] == '|')) {
        const char a = f[0], b = f[2];
        if (a == '0' && b == '0') return 0b00;
        if ((a == '0' && b == '1') || (a == '1' && b == '0')) return 0b10;
        if (a == '1' && b == '1') return 0b11;
    }
    return 0b01;
}

}  // namespace

extern "C" int jxb_vcf_to_plink(const char* vcf_path, const char* out_prefix, int snps_only, size_t* n_samples_out,
                                size_t* n_sites_out) {
    using jxb::fail;
    if (!vcf_path || !out_prefix) return fail(-2, "null argument");
    LineReader in;
    std::string err;
    if (!in.open(vcf_path, err)) return fail(-50, err);
    std::string line;
    std::vector<std::string> samples;
    bool have_header = false;
    while (in.next(line)) {
        if (line.compare(0, 6, "#CHROM") == 0) {
            while (!line.empty() && (line.back() == '\r' || line.back() == '\n' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
            size_t col = 0, start = 0;
            for (size_t i = 0; i <= line.size(); ++i) {
                if (i == line.size() || line[i] == '\t') {
                    if (col >= 9) samples.emplace_back(line.substr(start, i - start));
                    ++col;
                    start = i + 1;
                }
            }
            if (col < 10) { in.close(); return fail(-51, "#CHROM header too short"); }
            have_header = true;
            break;
        }
    }
    if (!have_header) { in.close(); return fail(-52, "No #CHROM header found in VCF"); }
    const size_t n = samples.size();
    const size_t bps = (n + 3) / 4;
    const std::string prefix = out_prefix;
    FILE* fbed = fopen((prefix + ".bed").c_str(), "wb");
    FILE* fbim = fopen((prefix + ".bim").c_str(), "w");
    FILE* ffam = fopen((prefix + ".fam").c_str(), "w");
    auto close_all = [&]() {
        in.close();
        if (fbed) fclose(fbed);
        if (fbim) fclose(fbim);
        if (ffam) fclose(ffam);
    };
    if (!fbed || !fbim || !ffam) { close_all(); return fail(-53, "create " + prefix + ".bed/.bim/.fam failed"); }
    for (const std::string& s : samples) fprintf(ffam, "%s\t%s\t0\t0\t0\t-9\n", s.c_str(), s.c_str());
    const unsigned char magic[3] = {0x6C, 0x1B, 0x01};
    fwrite(magic, 1, 3, fbed);
    std::vector<uint8_t> row(bps);
    std::vector<const char*> fld;
    std::vector<size_t> flen;
    size_t n_sites = 0;
    while (in.next(line)) {
        if (line.empty() || line[0] == '#') continue;
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        if (line.empty()) continue;
        fld.clear(); flen.clear();
        size_t start = 0;
        for (size_t i = 0; i <= line.size(); ++i) {
            if (i == line.size() || line[i] == '\t') {
                fld.push_back(line.data() + start);
                flen.push_back(i - start);
                start = i + 1;
            }
        }
        if (fld.size() < 10) continue;                               // gfcore.rs:2940-2943
        bool has_gt = false;                                         // FORMAT must list GT (gfcore.rs:2945-2948)
        for (size_t i = 0, s0 = 0; i <= flen[8]; ++i)
            if (i == flen[8] || fld[8][i] == ':') {
                if (i - s0 == 2 && fld[8][s0] == 'G' && fld[8][s0 + 1] == 'T') has_gt = true;
                s0 = i + 1;
            }
        if (!has_gt) continue;
        if (snps_only && (!simple_allele(fld[3], flen[3]) || !simple_allele(fld[4], flen[4]))) continue;
        std::fill(row.begin(), row.end(), 0);
        const size_t ns = std::min(n, fld.size() - 9);
        for (size_t j = 0; j < ns; ++j) row[j >> 2] |= (uint8_t)(gt_code(fld[9 + j], flen[9 + j]) << (2 * (j & 3)));
        for (size_t j = ns; j < n; ++j) row[j >> 2] |= (uint8_t)(0b01 << (2 * (j & 3)));   // short line: missing
        fwrite(row.data(), 1, bps, fbed);
        // pos: parse().unwrap_or(0); snp: ID unless empty or "." -> chrom_pos (gfcore.rs:2950-2960)
        const std::string chrom(fld[0], flen[0]), pos_s(fld[1], flen[1]), id(fld[2], flen[2]);
        char* endp = nullptr;
        long long pos = strtoll(pos_s.c_str(), &endp, 10);
        if (pos_s.empty() || *endp != '\0' || pos < -2147483648LL || pos > 2147483647LL) pos = 0;
        std::string id_t = id;
        while (!id_t.empty() && isspace((unsigned char)id_t.back())) id_t.pop_back();
        const std::string name = (!id_t.empty() && id != ".") ? id : chrom + "_" + pos_s;
        fprintf(fbim, "%s\t%s\t0\t%lld\t%.*s\t%.*s\n", chrom.c_str(), name.c_str(), pos, (int)flen[3], fld[3],
                (int)flen[4], fld[4]);
        ++n_sites;
    }
    const bool bad = ferror(fbed) || ferror(fbim) || ferror(ffam);
    close_all();
    if (bad) return fail(-54, "write " + prefix + ".bed/.bim/.fam failed");
    if (n_samples_out) *n_samples_out = n;
    if (n_sites_out) *n_sites_out = n_sites;
    return 0;
}
