// Run-file writer: byte-identical to the reference's regroup + f-string loops
// (retriever/retrieve_top_passages.py:90-109, retriever/retrieve_top_queries.py:65-82):
//
//     f.write(f"{qid}\t{docid}\t{i+1}\t{s}\n")
//
// where `s` is a Python float produced by np.float32 ndarray.tolist(), i.e. the fp32 score
// widened to double and printed with float.__repr__ (shortest digits that round-trip the
// DOUBLE, CPython's format_float_short 'r' rules for where the point / exponent goes).
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "common_host.h"

namespace {

// Where CPython's format_float_short ('r') puts the point / the exponent: `digits` are the nd <= 24 shortest
// round-trip digits, decpt the position of the decimal point relative to them.  Exponent form iff decpt <= -4 or
// decpt > 16.  The digits are moved with fixed 24-byte copies (three register moves instead of a memcpy call per
// piece): `digits` must be readable for 48 bytes and `p` writable for 48 bytes; the bytes behind the returned
// pointer are scratch.
inline void copy24(char* d, const char* s) { memcpy(d, s, 24); }

inline char* layout_digits(char* p, const char* digits, int nd, int decpt) {
    if (decpt <= -4 || decpt > 16) {
        p[0] = digits[0];
        if (nd > 1) {
            p[1] = '.';
            copy24(p + 2, digits + 1);
            p += nd + 1;
        } else {
            ++p;
        }
        *p++ = 'e';
        int e = decpt - 1;
        if (e < 0) {
            *p++ = '-';
            e = -e;
        } else {
            *p++ = '+';
        }
        char eb[8];
        int ne = 0;
        do {
            eb[ne++] = char('0' + e % 10);
            e /= 10;
        } while (e);
        if (ne < 2) eb[ne++] = '0';  // at least two exponent digits
        while (ne) *p++ = eb[--ne];
    } else if (decpt <= 0) {
        *p++ = '0';
        *p++ = '.';
        for (int i = 0; i < -decpt; ++i) *p++ = '0';
        copy24(p, digits);
        p += nd;
    } else if (decpt >= nd) {
        copy24(p, digits);
        p += nd;
        for (int i = 0; i < decpt - nd; ++i) *p++ = '0';
        *p++ = '.';
        *p++ = '0';
    } else {
        copy24(p, digits);
        copy24(p + decpt + 1, digits + decpt);
        p[decpt] = '.';
        p += nd + 1;
    }
    return p;
}

// Python's repr(float(x)) into buf; returns length.  General path: any double, digits from std::to_chars
// (shortest round-trip, closest to the value among the shortest: the same contract as CPython's repr).
int py_repr_double(double v, char* buf) {
    if (std::isnan(v)) {
        memcpy(buf, "nan", 3);
        return 3;
    }
    if (std::isinf(v)) {
        if (v < 0) {
            memcpy(buf, "-inf", 4);
            return 4;
        }
        memcpy(buf, "inf", 3);
        return 3;
    }
    char* p = buf;
    if (std::signbit(v)) {
        *p++ = '-';
        v = -v;
    }
    if (v == 0.0) {
        memcpy(p, "0.0", 3);
        return int(p + 3 - buf);
    }
    // shortest round-trip digits in scientific form: d[.ddd]e[+-]XX
    char sci[48];
    auto res = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);
    int len = int(res.ptr - sci);
    int epos = 0;
    while (epos < len && sci[epos] != 'e') ++epos;
    char digits[48] = {0};
    int nd = 0;
    for (int i = 0; i < epos; ++i)
        if (sci[i] != '.') digits[nd++] = sci[i];
    int exp10 = 0;
    {
        int i = epos + 1;
        bool neg = false;
        if (sci[i] == '+' || sci[i] == '-') {
            neg = sci[i] == '-';
            ++i;
        }
        for (; i < len; ++i) exp10 = exp10 * 10 + (sci[i] - '0');
        if (neg) exp10 = -exp10;
    }
    return int(layout_digits(p, digits, nd, exp10 + 1) - buf);
}

// "00" "01" ... "99"
struct Pairs {
    char c[200];
    constexpr Pairs() : c() {
        for (int i = 0; i < 100; ++i) {
            c[2 * i] = char('0' + i / 10);
            c[2 * i + 1] = char('0' + i % 10);
        }
    }
};
constexpr Pairs kPairs;

// v < 10^8, no leading zeros
inline char* put_u32_short(char* p, uint32_t v) {
    const int n = v < 10000u ? (v < 100u ? (v < 10u ? 1 : 2) : (v < 1000u ? 3 : 4))
                             : (v < 1000000u ? (v < 100000u ? 5 : 6) : (v < 10000000u ? 7 : 8));
    char* q = p + n;
    while (v >= 100u) {
        const uint32_t t = v / 100u;
        memcpy(q -= 2, kPairs.c + 2 * (v - t * 100u), 2);
        v = t;
    }
    if (v >= 10u)
        memcpy(q - 2, kPairs.c + 2 * v, 2);
    else
        q[-1] = char('0' + v);
    return p + n;
}

// exactly eight digits of v < 10^8, leading zeros included: two independent chains of two pairs
inline char* put_8digits(char* p, uint32_t v) {
    const uint32_t hi = v / 10000u, lo = v - hi * 10000u;
    const uint32_t a = hi / 100u, b = hi - a * 100u, c = lo / 100u, d = lo - c * 100u;
    memcpy(p, kPairs.c + 2 * a, 2);
    memcpy(p + 2, kPairs.c + 2 * b, 2);
    memcpy(p + 4, kPairs.c + 2 * c, 2);
    memcpy(p + 6, kPairs.c + 2 * d, 2);
    return p + 8;
}

inline char* put_u64(char* p, uint64_t v) {
    if (v < 100000000ull) return put_u32_short(p, uint32_t(v));
    const uint64_t hi = v / 100000000ull;
    const uint32_t lo = uint32_t(v - hi * 100000000ull);
    if (hi < 100000000ull) return put_8digits(put_u32_short(p, uint32_t(hi)), lo);
    const uint64_t top = hi / 100000000ull;          // < 1845 for any 64-bit v
    const uint32_t mid = uint32_t(hi - top * 100000000ull);
    return put_8digits(put_8digits(put_u32_short(p, uint32_t(top)), mid), lo);
}

inline char* put_i64(char* p, int64_t v) {
    if (v < 0) {
        *p++ = '-';
        return put_u64(p, uint64_t(0) - uint64_t(v));
    }
    return put_u64(p, uint64_t(v));
}

typedef unsigned __int128 u128;

constexpr uint64_t kPow5[27] = {1ull,
                                5ull,
                                25ull,
                                125ull,
                                625ull,
                                3125ull,
                                15625ull,
                                78125ull,
                                390625ull,
                                1953125ull,
                                9765625ull,
                                48828125ull,
                                244140625ull,
                                1220703125ull,
                                6103515625ull,
                                30517578125ull,
                                152587890625ull,
                                762939453125ull,
                                3814697265625ull,
                                19073486328125ull,
                                95367431640625ull,
                                476837158203125ull,
                                2384185791015625ull,
                                11920928955078125ull,
                                59604644775390625ull,
                                298023223876953125ull,
                                1490116119384765625ull};

constexpr int kFastEmin = -33, kFastEmax = 52;   // binary exponents floor(log2 v) the fast path takes

// P(E) = 16 - floor(E * log10 2): v * 10^P lies in [1e16, 2e17) for every v in [2^E, 2^(E+1))
struct ScaleTable {
    int8_t p[kFastEmax - kFastEmin + 1];
    constexpr ScaleTable() : p() {
        // floor(E * log10(2)), log10(2) = 0.30102999566398...; 14 decimals decide the floor for every |E| < 64
        for (int E = kFastEmin; E <= kFastEmax; ++E) {
            const long long num = (long long)E * 30102999566398LL;   // E * log10(2) * 1e14
            long long fl = num / 100000000000000LL;
            if (num < 0 && fl * 100000000000000LL != num) --fl;
            p[E - kFastEmin] = int8_t(16 - fl);
        }
    }
};
constexpr ScaleTable kScale;

// Shortest round-trip digits of double(s) for an fp32 s: the only values the writer ever prints.
//
// s = m * 2^e with m < 2^24, so s * 10^P = m * 5^P * 2^(P+e) is ONE 64x64 -> 128-bit product and a shift, exactly:
// W = its integer part (17-18 digits), r / 2^sh its fraction.  The decimals that parse back to double(s) are those
// within h = 2^(E-53) of it (E = floor(log2 s); both ends included: the double's mantissa has 29 trailing zero
// bits, so it is even; below a power of two the lower half is h / 2).  In the unit of W, h = 5^P * 2^(-30-t) with
// t = -P-e, so the integers inside the interval are [a, b] = [W + ceil((r*2^31 - 2 hd) / 2^(31+sh)),
// W + floor((r*2^31 + 2 hu) / 2^(31+sh))], exact in 64-bit arithmetic (hu = 5^P * 2^max(0,-t) < 2^60.4, hd = hu, or
// hu / 2 below a power of two).  The shortest text is the largest j for which a multiple of 10^j lies in [a, b]
// (monotone in j), and of the multiples just below and just above s at that j the closer one that is inside, the even
// one on an exact tie: the contract of CPython's repr (David Gay's mode 0) and of std::to_chars, against which
// cldrd_format_score_selfcheck sweeps all 2^32 bit patterns (tools/check_score_text.py).
// Returns false outside the fast range (zero, denormals, |s| < 2^-33, |s| >= 2^53, inf, nan): those go through
// py_repr_double.
inline bool shortest_f32(uint32_t abs_bits, char* digits, int* nd, int* decpt) {
    const int E = int(abs_bits >> 23) - 127;
    if (E < kFastEmin || E > kFastEmax) return false;
    const uint32_t frac = abs_bits & 0x7fffffu;
    const uint64_t m = frac | 0x800000u;
    const int e = E - 23;
    const int P = kScale.p[E - kFastEmin];
    const int t = -P - e;
    const u128 prod = u128(m) * kPow5[P];
    uint64_t W, r, hnum;
    int sh;
    if (t >= 0) {
        W = uint64_t(prod >> t);
        r = uint64_t(prod) & ((uint64_t(1) << t) - 1);
        sh = t;
        hnum = kPow5[P];
    } else {
        W = uint64_t(prod) << (-t);
        r = 0;
        sh = 0;
        hnum = kPow5[P] << (-t);
    }
    const int s31 = 31 + sh;
    const int64_t r31 = int64_t(r << 31);
    const uint64_t b = W + (uint64_t(r31 + int64_t(2 * hnum)) >> s31);
    const int64_t num_lo = r31 - int64_t(frac ? 2 * hnum : hnum);
    const uint64_t a = W + uint64_t((num_lo + ((int64_t(1) << s31) - 1)) >> s31);   // arithmetic shift: floor
    // largest j with a multiple of 10^j in [a, b]; qb = b / 10^j, qw = W / 10^j
    uint64_t qb = b, qw = W, p10 = 1;
    int j = 0;
    for (;;) {
        const uint64_t qb1 = qb / 10, p10n = p10 * 10;
        if (qb1 * p10n < a) break;
        qb = qb1;
        qw /= 10;
        p10 = p10n;
        ++j;
    }
    const uint64_t cd = qw * p10;             // the multiple at or below s ...
    const bool in_d = cd >= a, in_u = cd + p10 <= b;   // ... and the one above it
    bool up;
    if (in_d && in_u) {
        const uint64_t rm = W - cd;            // s - cd = rm + r / 2^sh
        if (j == 0) {
            const uint64_t half = sh ? uint64_t(1) << (sh - 1) : 0;      // r > 0 only when sh >= 1
            up = r > half || (r == half && r && (qw & 1));
        } else {
            const uint64_t half = p10 / 2;
            up = rm > half || (rm == half && (r || (qw & 1)));
        }
    } else {
        up = in_u;
    }
    const uint64_t D = qw + (up ? 1 : 0);
    const int n = int(put_u64(digits, D) - digits);
    *nd = n;
    *decpt = n + j - P;
    return true;
}

inline char* put_score(char* p, float s) {
    uint32_t bits;
    memcpy(&bits, &s, 4);
    char digits[48] = {0};
    int nd, decpt;
    if (shortest_f32(bits & 0x7fffffffu, digits, &nd, &decpt)) {
        if (bits >> 31) *p++ = '-';
        return layout_digits(p, digits, nd, decpt);
    }
    return p + py_repr_double(double(s), p);
}

}  // namespace

extern "C" {

int cldrd_format_score(float s, char* buf) {
    if (!buf) return cldrd::fail(CLDRD_EINVAL, "format_score: NULL buffer");
    char tmp[96];                       // put_score uses the bytes behind its result as scratch
    const int n = int(put_score(tmp, s) - tmp);
    memcpy(buf, tmp, size_t(n));
    return n;
}

// The writer's fast path against the general one (std::to_chars digits) on `count` fp32 bit patterns
// first, first + stride, ... (wrapping mod 2^32).  Returns the number of patterns whose text differs and the first
// such pattern; how many of them the fast path took itself goes to *fast_taken.
int64_t cldrd_format_score_selfcheck(uint32_t first, uint32_t stride, int64_t count, uint32_t* first_bad,
                                     int64_t* fast_taken) {
    int64_t bad = 0, fast = 0;
    uint32_t bits = first;
    char a[96], b[96];
    for (int64_t i = 0; i < count; ++i, bits += stride) {
        float s;
        memcpy(&s, &bits, 4);
        const int la = int(put_score(a, s) - a);
        const int lb = py_repr_double(double(s), b);
        char dg[48];
        int nd, dp;
        fast += shortest_f32(bits & 0x7fffffffu, dg, &nd, &dp) ? 1 : 0;
        if (la != lb || memcmp(a, b, size_t(la)) != 0) {
            if (!bad && first_bad) *first_bad = bits;
            ++bad;
        }
    }
    if (fast_taken) *fast_taken = fast;
    return bad;
}

// One formatted line per hit.  `rank` is the 1-based rank of the first hit of this row.
static inline char* format_row(char* p, int64_t qid, const float* s, const int64_t* d, int32_t k, int64_t rank) {
    // qid + tab and the rank column are kept as text and moved with fixed 24-byte copies (the bytes behind the
    // field are overwritten by the next field); the rank counts up by one per line and is incremented in place
    char qbuf[48] = {0}, rbuf[48] = {0};
    const int qlen = int(put_i64(qbuf, qid) - qbuf) + 1;
    qbuf[qlen - 1] = '\t';
    int rlen = int(put_i64(rbuf, rank) - rbuf);
    const bool counting = rank >= 0 && rank <= INT64_MAX - k;
    for (int32_t j = 0; j < k; ++j) {
        copy24(p, qbuf);
        p = put_i64(p + qlen, d[j]);
        *p++ = '\t';
        if (counting) {
            copy24(p, rbuf);
            p += rlen;
            char* c = rbuf + rlen - 1;
            while (c >= rbuf && *c == '9') *c-- = '0';
            if (c >= rbuf) {
                ++*c;
            } else {                      // 99...9 -> 100...0
                rbuf[0] = '1';
                rbuf[rlen++] = '0';
            }
        } else {
            p = put_i64(p, rank + j);
        }
        *p++ = '\t';
        p = put_score(p, s[j]);
        *p++ = '\n';
    }
    return p;
}

// longest line: 20 (qid) + 20 (docid) + 20 (rank) + 24 (score) + 4 separators = 88; the fixed-size copies of
// format_row / layout_digits touch up to 48 bytes behind the field they write, which the piece buffers' tail covers
static constexpr size_t kMaxLine = 96;

static int pwrite_all(int fd, const char* q, size_t n, off_t off, const char* path) {
    while (n) {
        ssize_t w = pwrite(fd, q, n, off);
        if (w < 0) {
            if (errno == EINTR) continue;
            return cldrd::fail(CLDRD_EIO, "write to '%s' failed: %s", path, strerror(errno));
        }
        q += w;
        off += w;
        n -= size_t(w);
    }
    return CLDRD_OK;
}

int cldrd_write_run_mt(const char* path, const int64_t* qids, const float* scores, const int64_t* ids,
                       int64_t nq, int32_t k, int32_t append, int32_t threads, int64_t* lines_written) {
    if (!path || nq < 0 || k < 0 || (nq && k && (!qids || !scores || !ids)))
        return cldrd::fail(CLDRD_EINVAL, "write_run: bad argument");
    // explicit offsets instead of O_APPEND: the formatting threads pwrite their pieces in place
    int fd = open(path, O_WRONLY | O_CREAT | (append ? 0 : O_TRUNC), 0644);
    if (fd < 0) return cldrd::fail(CLDRD_EIO, "cannot open run file '%s': %s", path, strerror(errno));
    off_t base = 0;
    if (append) {
        base = lseek(fd, 0, SEEK_END);
        if (base < 0) {
            int rc = cldrd::fail(CLDRD_EIO, "cannot seek '%s': %s", path, strerror(errno));
            close(fd);
            return rc;
        }
    }
    // the reference's dict regroup: a qid seen again continues its rank sequence; callers pass rows
    // grouped so that equal qids are consecutive.  first_rank[i] = rank of row i's first hit.
    std::vector<int64_t> first_rank(size_t(nq), 1);
    for (int64_t i = 1; i < nq; ++i)
        if (qids[i] == qids[i - 1]) first_rank[size_t(i)] = first_rank[size_t(i - 1)] + k;
    int T = threads;
    if (T <= 0) {
        if (const char* e = getenv("CLDRD_WRITER_THREADS")) T = atoi(e);
        if (T <= 0) T = int(std::thread::hardware_concurrency());
    }
    // Pieces of ~2 MiB of text are handed out in file order.  A worker formats its piece into its own buffer,
    // learns the piece's file offset from its predecessor (start[i+1] = start[i] + length, published as soon as the
    // length is known, before the write), and writes it in place with pwrite: formatting and writing of different
    // pieces overlap, the bytes land exactly where the single-threaded loop would have put them.
    const int64_t piece = std::max<int64_t>(1, (int64_t(2) << 20) / std::max<int64_t>(1, int64_t(k) * 48));
    const int64_t npieces = (nq + piece - 1) / piece;
    T = int(std::max<int64_t>(1, std::min<int64_t>(std::min(T, 256), npieces)));
    std::vector<std::atomic<int64_t>> start(size_t(npieces) + 1);
    for (auto& x : start) x.store(-1, std::memory_order_relaxed);
    start[0].store(int64_t(base), std::memory_order_release);
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    std::vector<int> rcs(size_t(T), CLDRD_OK);
    std::vector<std::string> errs{size_t(T)};
    auto worker = [&](int t) {
        std::vector<char> buf(size_t(piece) * size_t(std::max(k, 1)) * kMaxLine + 128);
        for (;;) {
            const int64_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= npieces) break;
            const int64_t lo = i * piece, hi = std::min<int64_t>(nq, lo + piece);
            char* p = buf.data();
            if (!failed.load(std::memory_order_relaxed))
                for (int64_t r = lo; r < hi; ++r)
                    p = format_row(p, qids[r], scores + r * int64_t(k), ids + r * int64_t(k), k, first_rank[size_t(r)]);
            const int64_t len = int64_t(p - buf.data());
            int64_t at;
            while ((at = start[size_t(i)].load(std::memory_order_acquire)) < 0) std::this_thread::yield();
            start[size_t(i) + 1].store(at + len, std::memory_order_release);
            if (len && !failed.load(std::memory_order_relaxed)) {
                const int wrc = pwrite_all(fd, buf.data(), size_t(len), off_t(at), path);
                if (wrc) {
                    rcs[size_t(t)] = wrc;
                    errs[size_t(t)] = cldrd_last_error();   // the message is thread-local
                    failed.store(1, std::memory_order_relaxed);
                }
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(worker, t);
        worker(0);
        for (auto& x : th) x.join();
    }
    int rc = CLDRD_OK;
    for (int t = 0; t < T && !rc; ++t)
        if (rcs[size_t(t)]) rc = cldrd::fail(rcs[size_t(t)], "%s", errs[size_t(t)].c_str());
    if (close(fd) != 0 && !rc) rc = cldrd::fail(CLDRD_EIO, "close '%s' failed: %s", path, strerror(errno));
    if (lines_written) *lines_written = rc ? 0 : nq * int64_t(k);
    return rc;
}

int cldrd_write_run(const char* path, const int64_t* qids, const float* scores, const int64_t* ids,
                    int64_t nq, int32_t k, int32_t append, int64_t* lines_written) {
    return cldrd_write_run_mt(path, qids, scores, ids, nq, k, append, 0, lines_written);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Run-file reader: "qid\tpid[\trank[\tscore]]" lines back into arrays (what evaluation/retrieval_evaluator.py:46-63
// and the curriculum post-processing read; at config 5 the file has 100 M lines).  Same acceptance rule as the
// reference's reader: a line is `line.strip().split("\t")`, must have 2 to 4 fields, fields 0 and 1 are integers.
// The file is cut into byte ranges, one per thread; a line belongs to the range that holds its '\n'.
// ---------------------------------------------------------------------------------------------------------------
namespace {

inline bool is_py_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// int(field) for the canonical spellings: optional blanks, optional sign, decimal digits.  false otherwise
// (including values outside int64, which no id of the retriever reaches).
inline bool parse_i64_field(const char* b, const char* e, int64_t* out) {
    while (b < e && is_py_space(*b)) ++b;
    while (e > b && is_py_space(e[-1])) --e;
    bool neg = false;
    if (b < e && (*b == '+' || *b == '-')) neg = *b++ == '-';
    if (b == e) return false;
    uint64_t v = 0;
    for (; b < e; ++b) {
        const unsigned dgt = unsigned(*b) - unsigned('0');
        if (dgt > 9) return false;
        if (v > (UINT64_MAX - dgt) / 10) return false;
        v = v * 10 + dgt;
    }
    if (v > uint64_t(INT64_MAX) + (neg ? 1u : 0u)) return false;
    *out = neg ? int64_t(uint64_t(0) - v) : int64_t(v);
    return true;
}

// 0 ok, 1 = field count outside 2..4, 2 = field 0 or 1 is not an integer
inline int parse_run_line(const char* b, const char* e, int64_t* qid, int64_t* pid) {
    while (b < e && is_py_space(*b)) ++b;
    while (e > b && is_py_space(e[-1])) --e;
    const char* t1 = static_cast<const char*>(memchr(b, '\t', size_t(e - b)));
    if (!t1) return 1;
    const char* t2 = static_cast<const char*>(memchr(t1 + 1, '\t', size_t(e - t1 - 1)));
    const char* f1_end = t2 ? t2 : e;
    if (t2) {
        const char* t3 = static_cast<const char*>(memchr(t2 + 1, '\t', size_t(e - t2 - 1)));
        if (t3 && memchr(t3 + 1, '\t', size_t(e - t3 - 1))) return 1;      // five fields or more
    }
    if (!parse_i64_field(b, t1, qid) || !parse_i64_field(t1 + 1, f1_end, pid)) return 2;
    return 0;
}

struct RunSpan {
    int64_t r0 = 0, r1 = 0;       // byte range scanned for '\n'
    int64_t lines = 0;            // lines owned
    int64_t last_nl = -1;         // position of the last '\n' in the range
    int64_t begin = 0, end = 0;   // bytes of the owned lines
    int64_t first_line = 0;       // index of the first owned line
};

}  // namespace

extern "C" {

int cldrd_read_run(const char* path, int64_t* qids, int64_t* pids, int64_t capacity, int32_t threads,
                   int64_t* nlines, int64_t* bad_line) {
    if (!path || !nlines || (qids == nullptr) != (pids == nullptr) || capacity < 0)
        return cldrd::fail(CLDRD_EINVAL, "read_run: bad argument");
    if (bad_line) *bad_line = -1;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return cldrd::fail(CLDRD_EIO, "cannot open run file '%s': %s", path, strerror(errno));
    const int64_t size = int64_t(lseek(fd, 0, SEEK_END));
    if (size < 0) {
        close(fd);
        return cldrd::fail(CLDRD_EIO, "cannot seek '%s': %s", path, strerror(errno));
    }
    int T = threads;
    if (T <= 0) {
        if (const char* e = getenv("CLDRD_WRITER_THREADS")) T = atoi(e);
        if (T <= 0) T = int(std::thread::hardware_concurrency());
    }
    T = int(std::max<int64_t>(1, std::min<int64_t>(std::min(T, 256), (size + (int64_t(4) << 20) - 1) / (int64_t(4) << 20))));
    std::vector<RunSpan> sp{size_t(T)};
    std::atomic<int> io_failed{0};
    constexpr size_t kBlock = size_t(4) << 20;
    auto in_threads = [&](auto&& fn) {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(fn, t);
        fn(0);
        for (auto& x : th) x.join();
    };
    // pass 1: '\n' per byte range
    in_threads([&](int t) {
        RunSpan& s = sp[size_t(t)];
        s.r0 = size * t / T;
        s.r1 = size * (t + 1) / T;
        std::vector<char> buf(kBlock);
        for (int64_t pos = s.r0; pos < s.r1;) {
            const size_t want = size_t(std::min<int64_t>(int64_t(kBlock), s.r1 - pos));
            const ssize_t got = pread(fd, buf.data(), want, off_t(pos));
            if (got <= 0) {
                if (got < 0 && errno == EINTR) continue;
                io_failed.store(1);
                return;
            }
            const char* p = buf.data();
            const char* lim = p + got;
            while (const char* nl = static_cast<const char*>(memchr(p, '\n', size_t(lim - p)))) {
                ++s.lines;
                s.last_nl = pos + (nl - buf.data());
                p = nl + 1;
            }
            pos += got;
        }
    });
    if (io_failed.load()) {
        close(fd);
        return cldrd::fail(CLDRD_EIO, "read from '%s' failed", path);
    }
    int64_t total = 0, owned_to = 0;
    for (int t = 0; t < T; ++t) {
        RunSpan& s = sp[size_t(t)];
        s.begin = owned_to;
        s.end = s.last_nl >= 0 ? s.last_nl + 1 : owned_to;
        owned_to = s.end;
        s.first_line = total;
        total += s.lines;
    }
    if (owned_to < size) {            // the last line has no '\n': it belongs to the last range
        sp[size_t(T - 1)].end = size;
        sp[size_t(T - 1)].lines += 1;
        total += 1;
    }
    *nlines = total;
    if (!qids) {
        close(fd);
        return CLDRD_OK;
    }
    if (capacity < total) {
        close(fd);
        return cldrd::fail(CLDRD_EINVAL, "read_run: %lld lines, room for %lld", (long long)total, (long long)capacity);
    }
    // pass 2: parse the owned lines
    std::vector<int64_t> bad_at(size_t(T), -1);
    std::vector<int> bad_kind(size_t(T), 0);
    in_threads([&](int t) {
        const RunSpan& s = sp[size_t(t)];
        std::vector<char> buf(kBlock + (size_t(1) << 16));
        size_t have = 0;
        int64_t line = s.first_line;
        auto take = [&](const char* b, const char* e) {
            const int k = parse_run_line(b, e, qids + line, pids + line);
            if (k && bad_at[size_t(t)] < 0) {
                bad_at[size_t(t)] = line;
                bad_kind[size_t(t)] = k;
            }
            ++line;
        };
        for (int64_t pos = s.begin; pos < s.end || have;) {
            const size_t want = size_t(std::min<int64_t>(int64_t(buf.size() - have), s.end - pos));
            size_t got = 0;
            while (got < want) {
                const ssize_t r = pread(fd, buf.data() + have + got, want - got, off_t(pos + int64_t(got)));
                if (r <= 0) {
                    if (r < 0 && errno == EINTR) continue;
                    io_failed.store(1);
                    return;
                }
                got += size_t(r);
            }
            pos += int64_t(got);
            have += got;
            const char* p = buf.data();
            const char* lim = p + have;
            while (const char* nl = static_cast<const char*>(memchr(p, '\n', size_t(lim - p)))) {
                take(p, nl);
                p = nl + 1;
            }
            if (pos >= s.end) {
                if (p < lim) take(p, lim);          // the unterminated last line of the file
                break;
            }
            have = size_t(lim - p);
            if (have == buf.size()) {               // a "line" longer than the buffer: not a run file
                bad_at[size_t(t)] = line;
                bad_kind[size_t(t)] = 1;
                return;
            }
            memmove(buf.data(), p, have);
        }
    });
    close(fd);
    if (io_failed.load()) return cldrd::fail(CLDRD_EIO, "read from '%s' failed", path);
    for (int t = 0; t < T; ++t)
        if (bad_at[size_t(t)] >= 0) {
            if (bad_line) *bad_line = bad_at[size_t(t)];
            return cldrd::fail(CLDRD_EFORMAT, "%s: line %lld of '%s'",
                               bad_kind[size_t(t)] == 1 ? "array length is not legal" : "not an integer id",
                               (long long)(bad_at[size_t(t)] + 1), path);
        }
    return CLDRD_OK;
}

}  // extern "C"
