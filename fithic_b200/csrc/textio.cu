// Text boundary of the path (host code): gzipped contact-counts file -> SoA arrays, arrays -> `.significances.txt.gz`.
//
// The reference spends almost all of its wall time here: `for line in gzip.open(...)`, `split`, `int`, `float`
// (fithic/fithic.py:406-417, :1017-1023) and `outfile.write("%s\t%d\t%s\t%d\t%d\t%e\t%e\t%e\t%e\t%f\n" % ...)` through
// gzip level 9 (:1166-1212).  Once the kernels take milliseconds the CLI is I/O bound, so both directions are native:
//   reader  one thread inflates (zlib) into 4 MB blocks, the caller's thread parses them into growing arrays;
//   writer  rows are formatted and deflated in blocks of 64 k rows by a pool of threads, every block becomes one gzip
//           member and the members are written in order -- `gunzip` sees the same bytes the reference writes (the
//           compressed stream differs, as it does between any two gzip runs).
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <math.h>
#include <zlib.h>

#include "common.cuh"

namespace {

struct ContactsFile {
    std::vector<int32_t> mid1, mid2, cnt;
    std::vector<uint32_t> chrs;
    std::vector<std::string> chroms;
    std::unordered_map<std::string, uint32_t> ids;
};

inline const char *skip_ws(const char *p, const char *e) {
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\v' || *p == '\f')) ++p;
    return p;
}
inline const char *token_end(const char *p, const char *e) {
    while (p < e && !(*p == ' ' || *p == '\t' || *p == '\r' || *p == '\v' || *p == '\f' || *p == '\n')) ++p;
    return p;
}

// int(text) for the mid point columns: optional sign, digits.  Returns false on anything else.
inline bool parse_int(const char *b, const char *e, long long &out) {
    if (b == e) return false;
    bool neg = false;
    if (*b == '+' || *b == '-') {
        neg = *b == '-';
        ++b;
        if (b == e) return false;
    }
    long long v = 0;
    for (; b < e; ++b) {
        if (*b < '0' || *b > '9') return false;
        v = v * 10 + (*b - '0');
        if (v > (1ll << 40)) return false;
    }
    out = neg ? -v : v;
    return true;
}

uint32_t chrom_id(ContactsFile &f, const char *b, const char *e) {
    std::string name(b, e);
    auto it = f.ids.find(name);
    if (it != f.ids.end()) return it->second;
    const uint32_t id = (uint32_t)f.chroms.size();
    f.ids.emplace(name, id);
    f.chroms.push_back(std::move(name));
    return id;
}

// parses complete lines of [b, e); returns false (with the error set) on a malformed line
bool parse_block(ContactsFile &f, const char *b, const char *e, long long &lineno) {
    const char *p = b;
    while (p < e) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(e - p));
        if (!eol) eol = e;
        ++lineno;
        const char *t[5], *te[5];
        const char *q = p;
        int k = 0;
        for (; k < 5; ++k) {
            q = skip_ws(q, eol);
            if (q == eol) break;
            t[k] = q;
            q = token_end(q, eol);
            te[k] = q;
        }
        if (k == 0) {  // the reference would raise on an empty line; tolerate a trailing one
            p = eol + 1;
            continue;
        }
        q = skip_ws(q, eol);
        if (k != 5 || q != eol) {
            fhc::set_error("contact file: line %lld does not have 5 fields", lineno);
            return false;
        }
        long long m1, m2;
        if (!parse_int(t[1], te[1], m1) || !parse_int(t[3], te[3], m2) || m1 < 0 || m2 < 0 || m1 > 0x7fffffffll ||
            m2 > 0x7fffffffll) {
            fhc::set_error("contact file: line %lld has a mid point that is not an int32 >= 0", lineno);
            return false;
        }
        // contactCount = float(text); hitCount = int(contactCount): truncation toward zero (fithic/fithic.py:415)
        char buf[64];
        const size_t len = (size_t)(te[4] - t[4]);
        if (len >= sizeof(buf)) {
            fhc::set_error("contact file: line %lld has an over-long count", lineno);
            return false;
        }
        memcpy(buf, t[4], len);
        buf[len] = 0;
        char *endp = nullptr;
        double c;
        // fast path: plain digits
        long long ci;
        if (parse_int(t[4], te[4], ci)) {
            c = (double)ci;
        } else {
            c = strtod(buf, &endp);
            if (endp != buf + len || !(c == c)) {
                fhc::set_error("contact file: line %lld has a count that is not a number", lineno);
                return false;
            }
        }
        const double tr = c < 0 ? ceil(c) : floor(c);
        if (!(fabs(tr) <= 2147483647.0)) {
            fhc::set_error("contact file: line %lld has a count outside int32", lineno);
            return false;
        }
        const uint32_t c1 = chrom_id(f, t[0], te[0]), c2 = chrom_id(f, t[2], te[2]);
        if (f.chroms.size() >= 65536) {
            fhc::set_error("contact file: more than 65535 chromosome names");
            return false;
        }
        f.mid1.push_back((int32_t)m1);
        f.mid2.push_back((int32_t)m2);
        f.cnt.push_back((int32_t)tr);
        f.chrs.push_back(c1 | (c2 << 16));
        p = eol + 1;
    }
    return true;
}

struct Block {
    std::vector<char> data;
    bool last = false;
};

}  // namespace

extern "C" void *fhc_io_read_contacts(const char *path) {
    gzFile gz = gzopen(path, "rb");
    if (!gz) {
        fhc::set_error("cannot open %s", path);
        return nullptr;
    }
    gzbuffer(gz, 1 << 20);
    ContactsFile *f = new ContactsFile();
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Block> queue;
    bool failed = false;
    // producer: inflate into blocks that end on a line boundary
    std::thread producer([&]() {
        std::vector<char> carry;
        const size_t kBlock = 4u << 20;
        while (true) {
            Block b;
            b.data.resize(carry.size() + kBlock);
            if (!carry.empty()) memcpy(b.data.data(), carry.data(), carry.size());
            const int got = gzread(gz, b.data.data() + carry.size(), (unsigned)kBlock);
            if (got < 0) {
                std::lock_guard<std::mutex> lk(mu);
                failed = true;
                b.last = true;
                b.data.clear();
                queue.push_back(std::move(b));
                cv.notify_all();
                return;
            }
            const size_t total = carry.size() + (size_t)got;
            b.data.resize(total);
            carry.clear();
            if (got == 0) {
                b.last = true;
            } else {
                // keep the partial last line for the next block
                size_t cut = total;
                while (cut > 0 && b.data[cut - 1] != '\n') --cut;
                if (cut < total) {
                    carry.assign(b.data.begin() + cut, b.data.end());
                    b.data.resize(cut);
                }
            }
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&]() { return queue.size() < 8; });
                queue.push_back(std::move(b));
            }
            cv.notify_all();
            if (got == 0) return;
        }
    });
    long long lineno = 0;
    bool ok = true;
    while (true) {
        Block b;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&]() { return !queue.empty(); });
            b = std::move(queue.front());
            queue.pop_front();
        }
        cv.notify_all();
        if (ok && !b.data.empty()) ok = parse_block(*f, b.data.data(), b.data.data() + b.data.size(), lineno);
        if (b.last) break;
    }
    producer.join();
    gzclose(gz);
    if (failed) {
        fhc::set_error("read error in %s (not a gzip file?)", path);
        ok = false;
    }
    if (!ok) {
        delete f;
        return nullptr;
    }
    return f;
}

extern "C" int64_t fhc_io_contacts_n(void *h) { return h ? (int64_t) static_cast<ContactsFile *>(h)->mid1.size() : -1; }
extern "C" int32_t fhc_io_contacts_nchrom(void *h) { return h ? (int32_t) static_cast<ContactsFile *>(h)->chroms.size() : -1; }
extern "C" const char *fhc_io_contacts_chrom(void *h, int32_t i) {
    ContactsFile *f = static_cast<ContactsFile *>(h);
    if (!f || i < 0 || (size_t)i >= f->chroms.size()) return nullptr;
    return f->chroms[(size_t)i].c_str();
}
extern "C" int fhc_io_contacts_copy(void *h, int32_t *mid1, int32_t *mid2, int32_t *cnt, uint32_t *chrs) {
    ContactsFile *f = static_cast<ContactsFile *>(h);
    FHC_REQUIRE(f && mid1 && mid2 && cnt && chrs, FHC_E_INVALID, "fhc_io_contacts_copy: null argument");
    const size_t n = f->mid1.size();
    memcpy(mid1, f->mid1.data(), n * sizeof(int32_t));
    memcpy(mid2, f->mid2.data(), n * sizeof(int32_t));
    memcpy(cnt, f->cnt.data(), n * sizeof(int32_t));
    memcpy(chrs, f->chrs.data(), n * sizeof(uint32_t));
    return FHC_OK;
}
extern "C" void fhc_io_free(void *h) { delete static_cast<ContactsFile *>(h); }

// ---------------------------------------------------------------------------------------------------------------------
// writer
// ---------------------------------------------------------------------------------------------------------------------
namespace {

// ---- "%e" / "%f" with six decimals, correctly rounded like printf, without printf ------------------------------------
// The decimal digits come from one extended-precision (x87, 64-bit mantissa) multiplication by a power of ten; the
// result is within ~1e-12 of the exact scaled value, so rounding to an integer is certain unless the fraction lies
// within 1e-6 of a half -- then (and for subnormals / huge values) snprintf decides.  Exact ties (short dyadic values
// such as 2^-11) therefore always take the exact path.
static long double g_pow10[700];  // 10^(i - 350)
static bool g_pow10_ready = false;
static void init_pow10() {
    if (g_pow10_ready) return;
    // exact up to 10^27 in 64-bit mantissa; beyond that each entry is built from two exactly rounded factors
    for (int i = 0; i < 700; ++i) {
        char buf[16];
        snprintf(buf, sizeof(buf), "1e%d", i - 350);
        g_pow10[i] = strtold(buf, nullptr);  // correctly rounded by libc
    }
    g_pow10_ready = true;
}

inline char *put_digits(char *o, unsigned long long d, int ndigits) {  // exactly ndigits digits, zero padded
    for (int k = ndigits - 1; k >= 0; --k) {
        o[k] = (char)('0' + d % 10);
        d /= 10;
    }
    return o + ndigits;
}

inline char *put_e(char *o, double v) {  // Python's "%e"
    if (v != v) {
        memcpy(o, "nan", 3);
        return o + 3;
    }
    if (isinf(v)) {
        if (v < 0) *o++ = '-';
        memcpy(o, "inf", 3);
        return o + 3;
    }
    if (v == 0.0) {
        if (signbit(v)) *o++ = '-';
        memcpy(o, "0.000000e+00", 12);
        return o + 12;
    }
    const double a = fabs(v);
    if (a < 1e-290 || a > 1e290) return o + snprintf(o, 32, "%e", v);
    int e2;
    frexp(a, &e2);
    int E = (int)floor((e2 - 1) * 0.30102999566398120);  // floor(log10 a) or one less
    if ((long double)a >= g_pow10[E + 1 + 350]) ++E;
    long double scaled = (long double)a * g_pow10[6 - E + 350];  // in [1e6, 1e7)
    unsigned long long D = (unsigned long long)scaled;           // truncation
    const long double frac = scaled - (long double)D;
    if (frac > 0.499999L && frac < 0.500001L) return o + snprintf(o, 32, "%e", v);
    if (frac >= 0.5L) ++D;
    if (D >= 10000000ull) {
        D = 1000000ull;
        ++E;
    } else if (D < 1000000ull) {  // scaled landed just below 1e6 through rounding of the estimate: let printf decide
        return o + snprintf(o, 32, "%e", v);
    }
    if (v < 0) *o++ = '-';
    *o++ = (char)('0' + D / 1000000ull);
    *o++ = '.';
    o = put_digits(o, D % 1000000ull, 6);
    *o++ = 'e';
    *o++ = E < 0 ? '-' : '+';
    const int ae = E < 0 ? -E : E;
    if (ae >= 100) {
        o = put_digits(o, (unsigned long long)ae, 3);
    } else {
        o = put_digits(o, (unsigned long long)ae, 2);
    }
    return o;
}
inline char *put_f(char *o, double v) {  // Python's "%f"
    if (v != v) {
        memcpy(o, "nan", 3);
        return o + 3;
    }
    if (isinf(v)) {
        if (v < 0) *o++ = '-';
        memcpy(o, "inf", 3);
        return o + 3;
    }
    const double a = fabs(v);
    if (a >= 1e12) return o + snprintf(o, 400, "%f", v);
    const long double scaled = (long double)a * 1000000.0L;
    unsigned long long D = (unsigned long long)scaled;
    const long double frac = scaled - (long double)D;
    if (frac > 0.499999L && frac < 0.500001L) return o + snprintf(o, 400, "%f", v);
    if (frac >= 0.5L) ++D;
    if (signbit(v)) *o++ = '-';
    const unsigned long long ip = D / 1000000ull;
    char tmp[24];
    int k = 0;
    unsigned long long u = ip;
    do {
        tmp[k++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    while (k) *o++ = tmp[--k];
    *o++ = '.';
    return put_digits(o, D % 1000000ull, 6);
}
inline char *put_i(char *o, long long v) {
    char tmp[24];
    int k = 0;
    bool neg = v < 0;
    unsigned long long u = neg ? (unsigned long long)(-v) : (unsigned long long)v;
    do {
        tmp[k++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (neg) *o++ = '-';
    while (k) *o++ = tmp[--k];
    return o;
}

bool deflate_member(const std::vector<char> &in, std::vector<char> &out, int level) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(deflateBound(&zs, (uLong)in.size()) + 64);
    zs.next_in = (Bytef *)in.data();
    zs.avail_in = (uInt)in.size();
    zs.next_out = (Bytef *)out.data();
    zs.avail_out = (uInt)out.size();
    const int rc = deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    return rc == Z_STREAM_END;
}

}  // namespace

// Rows of `${lib}.spline_passN.resR.significances.txt.gz` (fithic/fithic.py:1178, :1197-1212).  A line is written when it
// is inter-chromosomal and mode is All / interOnly, or intra-chromosomal, in range (L <= d <= U, -1 = unbounded) and mode
// is All / intraOnly.  bias (nullable): the dense table of fhc_pvalues; without it both bias columns are 1.
extern "C" int64_t fhc_io_write_significances(const char *path, const char *const *chrom_names, int32_t nchrom,
                                              const int32_t *mid1, const int32_t *mid2, const int32_t *cnt,
                                              const uint32_t *chrs, const double *p, const double *q, const double *expcc,
                                              int64_t n, int32_t mode, int64_t L, int64_t U, const double *bias,
                                              const int32_t *bias_mid, const int64_t *chr_off, int32_t nbias_chr,
                                              int32_t res, int32_t nthreads, int32_t level, int32_t header) {
    FHC_REQUIRE(path && chrom_names && nchrom >= 0 && n >= 0, FHC_E_INVALID, "fhc_io_write_significances: bad arguments");
    FHC_REQUIRE(n == 0 || (mid1 && mid2 && cnt && chrs && p && q && expcc), FHC_E_INVALID,
                "fhc_io_write_significances: null array");
    FHC_REQUIRE(bias == nullptr || (bias_mid && chr_off && res >= 0), FHC_E_INVALID,
                "fhc_io_write_significances: bias needs bias_mid, chr_off, res >= 0");
    FILE *fp = fopen(path, "wb");
    FHC_REQUIRE(fp != nullptr, FHC_E_INVALID, "cannot create %s", path);
    init_pow10();
    if (nthreads < 1) nthreads = 1;
    if (level < 0 || level > 9) level = 6;
    const int64_t kRows = 1 << 16;
    const int64_t nblocks = (n + kRows - 1) / kRows;
    const long long Llo = L < 0 ? 0 : L;
    const long long Uhi = U < 0 ? INT64_MAX : U;
    auto bias_of = [&](uint32_t chr, int32_t mid) -> double {
        if (!bias) return 1.0;
        if ((int32_t)chr >= nbias_chr || mid < 0) return -1.0;
        if (res == 0) {  // restriction fragments: each chromosome's loci in ascending mid order, binary search
            const int32_t *lo = bias_mid + chr_off[chr], *hi = bias_mid + chr_off[chr + 1];
            const int32_t *it = std::lower_bound(lo, hi, mid);
            return (it < hi && *it == mid) ? bias[it - bias_mid] : -1.0;
        }
        const int64_t s = chr_off[chr] + mid / res;
        if (s >= chr_off[chr + 1] || bias_mid[s] != mid) return -1.0;
        return bias[s];
    };
    std::mutex mu;
    std::condition_variable cv;
    int64_t next_block = 0, next_write = 0, rows_written = 0;
    std::unordered_map<int64_t, std::vector<char>> done;
    bool failed = false;
    auto worker = [&]() {
        std::vector<char> text, comp;
        while (true) {
            int64_t b;
            {
                std::unique_lock<std::mutex> lk(mu);
                // do not run far ahead of the writer: bounded memory
                cv.wait(lk, [&]() { return failed || next_block >= nblocks + 1 || next_block - next_write < 4 * nthreads; });
                if (failed || next_block >= nblocks + 1) return;
                b = next_block++;
            }
            text.clear();
            int64_t rows = 0;
            if (b == 0 && header) {
                const char *hdr = "chr1\tfragmentMid1\tchr2\tfragmentMid2\tcontactCount\tp-value\tq-value\tbias1\tbias2\tExpCC\n";
                text.insert(text.end(), hdr, hdr + strlen(hdr));
            }
            if (b < nblocks) {
                const int64_t lo = b * kRows, hi = lo + kRows < n ? lo + kRows : n;
                text.reserve(text.size() + (size_t)(hi - lo) * 160);
                char row[1024];
                for (int64_t i = lo; i < hi; ++i) {
                    const uint32_t c1 = chrs[i] & 0xffffu, c2 = chrs[i] >> 16;
                    const bool inter = c1 != c2;
                    if (inter) {
                        if (mode == FHC_MODE_INTRA_ONLY) continue;
                    } else {
                        if (mode == FHC_MODE_INTER_ONLY) continue;
                        long long d = (long long)mid1[i] - (long long)mid2[i];
                        d = d < 0 ? -d : d;
                        if (d < Llo || d > Uhi) continue;
                    }
                    const char *n1 = c1 < (uint32_t)nchrom ? chrom_names[c1] : "?";
                    const char *n2 = c2 < (uint32_t)nchrom ? chrom_names[c2] : "?";
                    const size_t l1 = strlen(n1), l2 = strlen(n2);
                    if (l1 + l2 > 512) continue;
                    char *o = row;
                    memcpy(o, n1, l1); o += l1; *o++ = '\t';
                    o = put_i(o, mid1[i]); *o++ = '\t';
                    memcpy(o, n2, l2); o += l2; *o++ = '\t';
                    o = put_i(o, mid2[i]); *o++ = '\t';
                    o = put_i(o, cnt[i]); *o++ = '\t';
                    o = put_e(o, p[i]); *o++ = '\t';
                    o = put_e(o, q[i]); *o++ = '\t';
                    o = put_e(o, bias_of(c1, mid1[i])); *o++ = '\t';
                    o = put_e(o, bias_of(c2, mid2[i])); *o++ = '\t';
                    o = put_f(o, expcc[i]); *o++ = '\n';
                    text.insert(text.end(), row, o);
                    ++rows;
                }
            }
            const bool ok = text.empty() ? (comp.clear(), true) : deflate_member(text, comp, level);
            {
                std::unique_lock<std::mutex> lk(mu);
                if (!ok) failed = true;
                done[b] = comp;
                rows_written += rows;
                // whoever holds the next block in order writes it (and any successors already finished)
                while (!failed) {
                    auto it = done.find(next_write);
                    if (it == done.end()) break;
                    if (!it->second.empty() && fwrite(it->second.data(), 1, it->second.size(), fp) != it->second.size())
                        failed = true;
                    done.erase(it);
                    ++next_write;
                }
            }
            cv.notify_all();
        }
    };
    // block index nblocks is a sentinel so that an empty input still gets its header (block 0 carries it)
    std::vector<std::thread> pool;
    const int64_t total_blocks = nblocks > 0 ? nblocks : 1;
    (void)total_blocks;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker);
    for (auto &t : pool) t.join();
    const bool close_ok = fclose(fp) == 0;
    FHC_REQUIRE(!failed && close_ok, FHC_E_INVALID, "write error on %s", path);
    return rows_written;
}

// formatting of one value, for tests: kind 'e' or 'f'; returns the length written into out (>= 400 bytes)
extern "C" int fhc_io_format_double(double v, int kind, char *out) {
    init_pow10();
    char *e = kind == 'f' ? put_f(out, v) : put_e(out, v);
    *e = 0;
    return (int)(e - out);
}
