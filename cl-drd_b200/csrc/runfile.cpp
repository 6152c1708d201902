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

int cldrd_write_run(const char* path, const int64_t* qids, const float* scores, const int64_t* ids,
                    int64_t nq, int32_t k, int32_t append, int64_t* lines_written) {
    if (!path || nq < 0 || k < 0 || (nq && k && (!qids || !scores || !ids)))
        return cldrd::fail(CLDRD_EINVAL, "write_run: bad argument");
    int fd = open(path, O_WRONLY | O_CREAT | (append ? O_APPEND : O_TRUNC), 0644);
    if (fd < 0) return cldrd::fail(CLDRD_EIO, "cannot open run file '%s': %s", path, strerror(errno));
    const size_t kBuf = size_t(8) << 20;
    std::vector<char> buf(kBuf + 256);
    char* p = buf.data();
    int64_t lines = 0;
    int64_t rank = 0;
    int rc = CLDRD_OK;
    auto flush = [&]() {
        const char* q = buf.data();
        size_t n = size_t(p - buf.data());
        while (n) {
            ssize_t w = write(fd, q, n);
            if (w < 0) {
                if (errno == EINTR) continue;
                rc = cldrd::fail(CLDRD_EIO, "write to '%s' failed: %s", path, strerror(errno));
                return;
            }
            q += w;
            n -= size_t(w);
        }
        p = buf.data();
    };
    for (int64_t i = 0; i < nq && !rc; ++i) {
        // the reference's dict regroup: a qid seen again continues its rank sequence; callers
        // pass rows grouped so that equal qids are consecutive.
        if (i == 0 || qids[i] != qids[i - 1]) rank = 0;
        char qbuf[24];
        int qlen = int(put_i64(qbuf, qids[i]) - qbuf);
        const float* s = scores + i * int64_t(k);
        const int64_t* d = ids + i * int64_t(k);
        for (int32_t j = 0; j < k; ++j) {
            memcpy(p, qbuf, qlen);
            p += qlen;
            *p++ = '\t';
            p = put_i64(p, d[j]);
            *p++ = '\t';
            p = put_i64(p, ++rank);
            *p++ = '\t';
            p += py_repr_double(double(s[j]), p);
            *p++ = '\n';
            ++lines;
            if (size_t(p - buf.data()) >= kBuf) {
                flush();
                if (rc) break;
            }
        }
    }
    if (!rc) flush();
    if (close(fd) != 0 && !rc) rc = cldrd::fail(CLDRD_EIO, "close '%s' failed: %s", path, strerror(errno));
    if (lines_written) *lines_written = lines;
    return rc;
}

}  // extern "C"
