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

// Python's repr(float(x)) into buf; returns length.
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
    char digits[32];
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
    int decpt = exp10 + 1;  // position of the decimal point relative to the digit string
    // CPython format_float_short, 'r': exponent form iff decpt <= -4 or decpt > 16
    if (decpt <= -4 || decpt > 16) {
        *p++ = digits[0];
        if (nd > 1) {
            *p++ = '.';
            memcpy(p, digits + 1, nd - 1);
            p += nd - 1;
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
        memcpy(p, digits, nd);
        p += nd;
    } else if (decpt >= nd) {
        memcpy(p, digits, nd);
        p += nd;
        for (int i = 0; i < decpt - nd; ++i) *p++ = '0';
        *p++ = '.';
        *p++ = '0';
    } else {
        memcpy(p, digits, decpt);
        p += decpt;
        *p++ = '.';
        memcpy(p, digits + decpt, nd - decpt);
        p += nd - decpt;
    }
    return int(p - buf);
}

inline char* put_i64(char* p, int64_t v) {
    auto r = std::to_chars(p, p + 24, v);
    return r.ptr;
}

}  // namespace

extern "C" {

int cldrd_format_score(float s, char* buf) {
    if (!buf) return cldrd::fail(CLDRD_EINVAL, "format_score: NULL buffer");
    return py_repr_double(double(s), buf);
}

// One formatted line per hit.  `rank` is the 1-based rank of the first hit of this row.
static inline char* format_row(char* p, int64_t qid, const float* s, const int64_t* d, int32_t k, int64_t rank) {
    char qbuf[24];
    const int qlen = int(put_i64(qbuf, qid) - qbuf);
    for (int32_t j = 0; j < k; ++j) {
        memcpy(p, qbuf, qlen);
        p += qlen;
        *p++ = '\t';
        p = put_i64(p, d[j]);
        *p++ = '\t';
        p = put_i64(p, rank + j);
        *p++ = '\t';
        p += py_repr_double(double(s[j]), p);
        *p++ = '\n';
    }
    return p;
}

// longest line: 20 (qid) + 20 (docid) + 20 (rank) + 24 (score) + 4 separators
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
        std::vector<char> buf(size_t(piece) * size_t(std::max(k, 1)) * kMaxLine + 64);
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
