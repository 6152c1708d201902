// Host-side helpers shared by the translation units of libcldrd.so.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "cldrd.h"

namespace cldrd {

// thread-local last-error message (cldrd_last_error)
std::string& last_error_ref();

inline int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

// Layout of an index file.
struct IndexFileInfo {
    int64_t ntotal = 0;
    int32_t d = 0;
    int32_t metric = 0;
    int32_t has_ids = 0;
    int32_t idmap2 = 0;
    int64_t data_off = 0;
    int64_t ids_off = 0;
};
int probe_index_file(const char* path, IndexFileInfo* info);

}  // namespace cldrd
