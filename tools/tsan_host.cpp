// Writer and reader threads of libcldrd under ThreadSanitizer (CPU only):
//   g++ -O1 -g -fsanitize=thread -std=c++17 -Iinclude -Icl-drd_b200/csrc tools/tsan_host.cpp cl-drd_b200/csrc/runfile.cpp cl-drd_b200/csrc/index_io.cpp -o /tmp/tsan_host -lpthread && /tmp/tsan_host
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "cldrd.h"
int main() {
    const int nq = 3000, k = 200;
    std::mt19937_64 g(1);
    std::vector<float> s(size_t(nq) * k);
    std::vector<int64_t> ids(size_t(nq) * k), q(nq);
    for (auto& x : s) x = float(g() % 100000) / 997.f;
    for (auto& x : ids) x = int64_t(g() % 8841823);
    for (int i = 0; i < nq; ++i) q[i] = 1000 + i;
    int64_t lines = 0;
    for (int rep = 0; rep < 3; ++rep) {
        if (cldrd_write_run_mt("/dev/shm/tsan.tsv", q.data(), s.data(), ids.data(), nq, k, rep > 0, 8, &lines)) { puts(cldrd_last_error()); return 1; }
    }
    int64_t n = 0, bad = -1;
    if (cldrd_read_run("/dev/shm/tsan.tsv", nullptr, nullptr, 0, 8, &n, nullptr)) return 2;
    std::vector<int64_t> qq(n), pp(n);
    if (cldrd_read_run("/dev/shm/tsan.tsv", qq.data(), pp.data(), n, 8, &n, &bad)) { puts(cldrd_last_error()); return 3; }
    for (int64_t i = 0; i < n; ++i) if (pp[i] != ids[i % (int64_t(nq) * k)] || qq[i] != q[(i / k) % nq]) return 4;
    remove("/dev/shm/tsan.tsv");
    printf("tsan run ok: %lld lines written per call, %lld read\n", (long long)lines, (long long)n);
}
