// Index-file reader/writer: faiss' on-disk layout for IndexIDMap{IndexFlatIP}
// (what retriever/index_text.py:91-105 writes and retrieve_top_passages.py:85 reads).
//
//   'IxMp'|'IxM2'  d:i32  ntotal:i64  dummy:i64  dummy:i64  is_trained:u8  metric:i32     (37 B)
//   'IxFI'         d:i32  ntotal:i64  dummy:i64  dummy:i64  is_trained:u8  metric:i32     (37 B)
//   count:u64 (= ntotal*d)   float32[ntotal*d]                 <- payload starts at byte 82
//   count:u64 (= ntotal)     int64[ntotal]                     <- id_map
//
// Little-endian, packed.  A bare 'IxFI' file (no id map) starts at the second header.
#include <cerrno>
#include <cstring>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <vector>

#include "common_host.h"

namespace cldrd {

std::string& last_error_ref() {
    static thread_local std::string msg;
    return msg;
}

namespace {

constexpr int64_t kDummy = 1 << 20;
constexpr int kHeaderBytes = 4 + 4 + 8 + 8 + 8 + 1 + 4;  // 37

struct Header {
    char fourcc[4];
    int32_t d;
    int64_t ntotal;
    int32_t metric;
};

void put_header(unsigned char* p, const char* fourcc, int32_t d, int64_t ntotal, int32_t metric) {
    memcpy(p, fourcc, 4);
    memcpy(p + 4, &d, 4);
    memcpy(p + 8, &ntotal, 8);
    memcpy(p + 16, &kDummy, 8);
    memcpy(p + 24, &kDummy, 8);
    p[32] = 1;
    memcpy(p + 33, &metric, 4);
}

bool pread_all(int fd, void* buf, size_t n, off_t off) {
    char* p = static_cast<char*>(buf);
    while (n > 0) {
        ssize_t r = pread(fd, p, n, off);
        if (r < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        if (r == 0) return false;
        p += r;
        off += r;
        n -= static_cast<size_t>(r);
    }
    return true;
}

bool pwrite_all(int fd, const void* buf, size_t n, off_t off) {
    const char* p = static_cast<const char*>(buf);
    while (n > 0) {
        ssize_t r = pwrite(fd, p, n, off);
        if (r < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += r;
        off += r;
        n -= static_cast<size_t>(r);
    }
    return true;
}

// returns bytes consumed, or <0
int get_header(int fd, off_t off, Header* h) {
    unsigned char b[kHeaderBytes + 4];
    if (!pread_all(fd, b, kHeaderBytes, off)) return -1;
    memcpy(h->fourcc, b, 4);
    memcpy(&h->d, b + 4, 4);
    memcpy(&h->ntotal, b + 8, 8);
    memcpy(&h->metric, b + 33, 4);
    int used = kHeaderBytes;
    if (h->metric > 1) used += 4;  // faiss stores metric_arg:f32 for the exotic metrics
    return used;
}

bool is_flat(const char* f) {
    return !memcmp(f, "IxFI", 4) || !memcmp(f, "IxF2", 4) || !memcmp(f, "IxFl", 4);
}

}  // namespace

int probe_index_file(const char* path, IndexFileInfo* info) {
    if (!path || !info) return fail(CLDRD_EINVAL, "probe: NULL argument");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(CLDRD_EIO, "cannot open index file '%s': %s", path, strerror(errno));
    struct stat st;
    fstat(fd, &st);
    Header h;
    off_t off = 0;
    int used = get_header(fd, off, &h);
    if (used < 0) {
        close(fd);
        return fail(CLDRD_EFORMAT, "'%s': truncated header", path);
    }
    off += used;
    *info = IndexFileInfo();
    bool wrapped = !memcmp(h.fourcc, "IxMp", 4) || !memcmp(h.fourcc, "IxM2", 4);
    if (wrapped) {
        info->has_ids = 1;
        info->idmap2 = !memcmp(h.fourcc, "IxM2", 4);
        Header h2;
        used = get_header(fd, off, &h2);
        if (used < 0 || !is_flat(h2.fourcc)) {
            close(fd);
            return fail(CLDRD_EFORMAT, "'%s': IndexIDMap does not wrap a flat index", path);
        }
        if (h2.d != h.d || h2.ntotal != h.ntotal) {
            close(fd);
            return fail(CLDRD_EFORMAT, "'%s': inner/outer header mismatch", path);
        }
        off += used;
        h = h2;
    } else if (!is_flat(h.fourcc)) {
        close(fd);
        return fail(CLDRD_EFORMAT, "'%s': fourcc '%.4s' is not IxMp/IxM2/IxFI", path, h.fourcc);
    }
    if (h.metric != 0) {
        close(fd);
        return fail(CLDRD_EFORMAT, "'%s': metric %d is not inner product", path, h.metric);
    }
    if (h.d <= 0 || h.ntotal < 0) {
        close(fd);
        return fail(CLDRD_EFORMAT, "'%s': bad d=%d ntotal=%lld", path, h.d, (long long)h.ntotal);
    }
    uint64_t cnt = 0;
    if (!pread_all(fd, &cnt, 8, off) || cnt != uint64_t(h.ntotal) * uint64_t(h.d)) {
        close(fd);
        return fail(CLDRD_EFORMAT, "'%s': payload count mismatch", path);
    }
    off += 8;
    info->ntotal = h.ntotal;
    info->d = h.d;
    info->metric = h.metric;
    info->data_off = off;
    off += off_t(cnt) * 4;
    if (wrapped) {
        uint64_t cnt2 = 0;
        if (!pread_all(fd, &cnt2, 8, off) || cnt2 != uint64_t(h.ntotal)) {
            close(fd);
            return fail(CLDRD_EFORMAT, "'%s': id_map count mismatch", path);
        }
        off += 8;
        info->ids_off = off;
        off += off_t(cnt2) * 8;
    }
    if (off > st.st_size) {
        close(fd);
        return fail(CLDRD_EFORMAT, "'%s': file shorter (%lld B) than its headers claim (%lld B)",
                    path, (long long)st.st_size, (long long)off);
    }
    close(fd);
    return CLDRD_OK;
}

}  // namespace cldrd

using namespace cldrd;

struct cldrd_index_writer {
    int fd = -1;
    int64_t n = 0;
    int32_t d = 0;
    int32_t with_ids = 0;
    int64_t rows_written = 0;
    int64_t data_off = 0;
    int64_t row0 = 0;        // this writer covers rows [row0, row0 + nrows) of the n the file declares
    int64_t nrows = 0;
    bool creator = true;     // wrote the headers; writes the id array in finish
    std::string path;
};

extern "C" {

const char* cldrd_last_error(void) { return last_error_ref().c_str(); }
int cldrd_abi_version(void) { return CLDRD_ABI_VERSION; }

int cldrd_index_probe(const char* path, int64_t* ntotal, int32_t* d, int32_t* metric,
                      int32_t* has_ids, int32_t* idmap2, int64_t* data_off, int64_t* ids_off) {
    IndexFileInfo info;
    int rc = probe_index_file(path, &info);
    if (rc) return rc;
    if (ntotal) *ntotal = info.ntotal;
    if (d) *d = info.d;
    if (metric) *metric = info.metric;
    if (has_ids) *has_ids = info.has_ids;
    if (idmap2) *idmap2 = info.idmap2;
    if (data_off) *data_off = info.data_off;
    if (ids_off) *ids_off = info.ids_off;
    return CLDRD_OK;
}

int cldrd_index_writer_open_range(cldrd_index_writer** out, const char* path, int64_t n, int32_t d, int32_t with_ids,
                                  int32_t idmap2, int64_t row0, int64_t nrows, int32_t create) {
    if (!out || !path || n < 0 || d <= 0 || row0 < 0 || nrows < 0 || row0 + nrows > n)
        return fail(CLDRD_EINVAL, "writer_open_range: bad argument (rows [%lld, +%lld) of %lld)", (long long)row0,
                    (long long)nrows, (long long)n);
    int fd = open(path, create ? (O_WRONLY | O_CREAT | O_TRUNC) : O_WRONLY, 0644);
    if (fd < 0) return fail(CLDRD_EIO, "cannot %s '%s': %s", create ? "create" : "open", path, strerror(errno));
    unsigned char hdr[2 * kHeaderBytes + 8];
    size_t len = 0;
    if (with_ids) {
        put_header(hdr, idmap2 ? "IxM2" : "IxMp", d, n, 0);
        len += kHeaderBytes;
    }
    put_header(hdr + len, "IxFI", d, n, 0);
    len += kHeaderBytes;
    uint64_t cnt = uint64_t(n) * uint64_t(d);
    memcpy(hdr + len, &cnt, 8);
    len += 8;
    if (create && !pwrite_all(fd, hdr, len, 0)) {
        close(fd);
        return fail(CLDRD_EIO, "write to '%s' failed: %s", path, strerror(errno));
    }
    if (!create) {      // joining a file somebody else created (or a build that is being continued): same layout?
        unsigned char have[sizeof(hdr)];
        int rfd = open(path, O_RDONLY);
        const bool got = rfd >= 0 && pread_all(rfd, have, len, 0);
        if (rfd >= 0) close(rfd);
        if (!got || memcmp(have, hdr, len) != 0) {
            close(fd);
            return fail(CLDRD_EFORMAT, "'%s' does not declare %lld rows of dimension %d in this layout", path,
                        (long long)n, (int)d);
        }
    }
    auto* w = new cldrd_index_writer();
    w->fd = fd;
    w->n = n;
    w->d = d;
    w->with_ids = with_ids;
    w->data_off = int64_t(len);
    w->row0 = row0;
    w->nrows = nrows;
    w->creator = create != 0;
    w->path = path;
    *out = w;
    return CLDRD_OK;
}

int cldrd_index_writer_begin(cldrd_index_writer** out, const char* path, int64_t n, int32_t d,
                             int32_t with_ids, int32_t idmap2) {
    if (n < 0) return fail(CLDRD_EINVAL, "writer_begin: bad argument");
    return cldrd_index_writer_open_range(out, path, n, d, with_ids, idmap2, 0, n, 1);
}

int cldrd_index_writer_append(cldrd_index_writer* w, const float* rows_host, int64_t nrows) {
    if (!w || (!rows_host && nrows) || nrows < 0) return fail(CLDRD_EINVAL, "writer_append: bad argument");
    if (w->rows_written + nrows > w->nrows)
        return fail(CLDRD_EINVAL, "writer_append: %lld rows exceed the %lld of this writer's range",
                    (long long)(w->rows_written + nrows), (long long)w->nrows);
    off_t off = w->data_off + off_t(w->row0 + w->rows_written) * w->d * 4;
    if (!pwrite_all(w->fd, rows_host, size_t(nrows) * w->d * 4, off))
        return fail(CLDRD_EIO, "write to '%s' failed: %s", w->path.c_str(), strerror(errno));
    w->rows_written += nrows;
    return CLDRD_OK;
}

int cldrd_index_writer_sync(cldrd_index_writer* w) {
    if (!w) return fail(CLDRD_EINVAL, "writer_sync: NULL");
    if (fdatasync(w->fd) != 0) return fail(CLDRD_EIO, "fdatasync '%s' failed: %s", w->path.c_str(), strerror(errno));
    return CLDRD_OK;
}

int cldrd_index_writer_finish(cldrd_index_writer* w, const int64_t* ids_host) {
    if (!w) return fail(CLDRD_EINVAL, "writer_finish: NULL");
    int rc = CLDRD_OK;
    if (w->rows_written != w->nrows)
        rc = fail(CLDRD_ESTATE, "writer_finish: %lld of %lld rows written", (long long)w->rows_written,
                  (long long)w->nrows);
    // the id array of ALL n rows: required from the writer that created the file, accepted from any other (the
    // process that continues an interrupted build did not create the file)
    if (!rc && w->with_ids && (w->creator || ids_host)) {
        if (!ids_host && w->n)
            rc = fail(CLDRD_EINVAL, "writer_finish: ids required");
        else {
            off_t off = w->data_off + off_t(w->n) * w->d * 4;
            uint64_t cnt = uint64_t(w->n);
            if (!pwrite_all(w->fd, &cnt, 8, off) || !pwrite_all(w->fd, ids_host, size_t(w->n) * 8, off + 8))
                rc = fail(CLDRD_EIO, "write to '%s' failed: %s", w->path.c_str(), strerror(errno));
        }
    }
    if (close(w->fd) != 0 && !rc) rc = fail(CLDRD_EIO, "close '%s' failed: %s", w->path.c_str(), strerror(errno));
    delete w;
    return rc;
}

int cldrd_index_write(const char* path, const float* xb_host, const int64_t* ids_host, int64_t n,
                      int32_t d, int32_t idmap2) {
    cldrd_index_writer* w = nullptr;
    int rc = cldrd_index_writer_begin(&w, path, n, d, ids_host != nullptr, idmap2);
    if (rc) return rc;
    rc = cldrd_index_writer_append(w, xb_host, n);
    int rc2 = cldrd_index_writer_finish(w, ids_host);
    return rc ? rc : rc2;
}

int cldrd_index_read_rows(const char* path, int64_t row0, int64_t nrows, float* out_host) {
    IndexFileInfo info;
    int rc = probe_index_file(path, &info);
    if (rc) return rc;
    if (row0 < 0 || nrows < 0 || row0 + nrows > info.ntotal || (!out_host && nrows))
        return fail(CLDRD_EINVAL, "read_rows: range [%lld,+%lld) outside ntotal=%lld", (long long)row0,
                    (long long)nrows, (long long)info.ntotal);
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(CLDRD_EIO, "cannot open '%s': %s", path, strerror(errno));
    bool ok = pread_all(fd, out_host, size_t(nrows) * info.d * 4, info.data_off + off_t(row0) * info.d * 4);
    close(fd);
    return ok ? CLDRD_OK : fail(CLDRD_EIO, "short read from '%s'", path);
}

int cldrd_index_read_ids(const char* path, int64_t row0, int64_t nrows, int64_t* out_host) {
    IndexFileInfo info;
    int rc = probe_index_file(path, &info);
    if (rc) return rc;
    if (!info.has_ids) return fail(CLDRD_EFORMAT, "'%s' has no id map", path);
    if (row0 < 0 || nrows < 0 || row0 + nrows > info.ntotal || (!out_host && nrows))
        return fail(CLDRD_EINVAL, "read_ids: range outside ntotal");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(CLDRD_EIO, "cannot open '%s': %s", path, strerror(errno));
    bool ok = pread_all(fd, out_host, size_t(nrows) * 8, info.ids_off + off_t(row0) * 8);
    close(fd);
    return ok ? CLDRD_OK : fail(CLDRD_EIO, "short read from '%s'", path);
}

}  // extern "C"
