// ThreadSanitizer driver for the host-side wall-layer gather (tools only; no GPU is touched).
// Includes csrc/compact.cu as one translation unit so that the persistent thread pool of a handle (HostPool, anonymous
// namespace) is exercised the way vh_push_snapshots uses it: ONE pool, hundreds of run() calls of varying size, and the
// handle-free vh_host_gather (a transient pool per call) beside it.  Results are compared with a serial gather.
//   bash tools/tsan/run.sh      ->  profiles/r2_tsan_host_gather.txt
#include <cstdarg>
#include <cstdio>

#include "../../vasp_b200/csrc/compact.cu"

void vh_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}

int main() {
    const int64_t n_nodes = 200003, n_rows = 9;
    std::vector<int32_t> slots;
    for (int64_t i = 0; i < n_nodes; ++i)
        if ((i * 2654435761u) % 7 < 2) slots.push_back((int32_t)i);
    const int64_t nW = (int64_t)slots.size();
    std::vector<double> u((size_t)(n_rows * 3 * n_nodes));
    for (size_t i = 0; i < u.size(); ++i) u[i] = (double)(i % 1000003) * 1e-3;
    const int64_t off[3] = {0, n_nodes, 2 * n_nodes};
    std::vector<double> want((size_t)(n_rows * 3 * nW)), got(want.size());
    for (int64_t r = 0; r < n_rows; ++r)
        for (int c = 0; c < 3; ++c)
            for (int64_t i = 0; i < nW; ++i) want[(r * 3 + c) * nW + i] = u[r * 3 * n_nodes + off[c] + slots[i]];
    int bad = 0, calls = 0;
    for (int threads : {1, 2, 5, 16}) {
        HostPool pool(threads);  // persistent: reused by every call below
        for (int rep = 0; rep < 60; ++rep) {
            const int64_t rows = 1 + rep % n_rows;
            std::fill(got.begin(), got.end(), -1.0);
            gather_rows(&pool, nullptr, u.data(), 3 * n_nodes, rows, slots.data(), nW, off, got.data(), 3 * nW);
            for (int64_t k = 0; k < rows * 3 * nW; ++k) bad += got[k] != want[k];
            ++calls;
        }
    }
    for (int rep = 0; rep < 10; ++rep) {
        std::fill(got.begin(), got.end(), -1.0);
        bad += vh_host_gather(nullptr, u.data(), 3 * n_nodes * 8, n_rows, slots.data(), nW, off, got.data(), 3 * nW * 8, 1 + rep) != 0;
        bad += got != want;
        ++calls;
    }
    printf("host gather under ThreadSanitizer: %lld wall nodes of %lld, %d calls (4 persistent pools x 60 + 10 transient), "
           "mismatches/errors: %d\n", (long long)nW, (long long)n_nodes, calls, bad);
    return bad != 0;
}
