// Wall-layer compaction in front of the PCIe bus (host side).
//
// The reference only ever integrates over ds (compute_hemodynamics.py:113-115), so of a velocity snapshot
// (compute_hemodynamics.py:274) only the dofs of cells that own an exterior facet can reach the result: 72 % of the
// nodes on the tutorial-size mesh, 20 % at 2 M tets, 7 % at 10 M tets.  Instead of copying whole vectors to the device
// and gathering there, the host gathers
//
//     C[c * nWn_pad + i] = vec[comp_offset[c] + wall_slot[i]]          c < 3, i < nWn_pad   ("compact block")
//
// into a ring of pinned buffers with a small thread pool, piece by piece, while the previous piece crosses the bus;
// the device then receives 24 * nWn_pad bytes per snapshot and K1 is a pure transpose.  The source rows may live
// anywhere in host memory -- in particular in an mmap of u.h5, so the page cache is gathered directly and never
// copied whole (vh_compact_rows).
//
// This is data movement, not arithmetic: every value the kernels use still comes from the caller's vectors, bit for
// bit.  There is no CPU compute path here.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace {

// Persistent workers; run(n_items, fn) executes fn(item) for item in [0, n_items) on the workers AND the caller.
class HostPool {
  public:
    explicit HostPool(int n_threads) {
        for (int t = 0; t + 1 < n_threads; ++t) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    int size() const { return (int)workers_.size() + 1; }

    template <class F>
    void run(int64_t n_items, F&& fn) {
        if (n_items <= 0) return;
        std::function<void(int64_t)> f = fn;
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &f;
            n_items_ = n_items;
            next_.store(0, std::memory_order_relaxed);
            pending_ = (int)workers_.size();
            ++generation_;
        }
        cv_.notify_all();
        drain(f, n_items);
        std::unique_lock<std::mutex> lk(m_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

  private:
    void drain(const std::function<void(int64_t)>& f, int64_t n) {
        for (;;) {
            const int64_t i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= n) return;
            f(i);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int64_t)>* f;
            int64_t n;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                f = job_;
                n = n_items_;
            }
            drain(*f, n);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int64_t)>* job_ = nullptr;
    int64_t n_items_ = 0;
    std::atomic<int64_t> next_{0};
    int pending_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

int auto_threads() {
    if (const char* e = getenv("VASP_B200_HOST_THREADS")) {
        const int n = atoi(e);
        if (n > 0) return n;
    }
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    // one process per GPU: the ranks of a node share its cores
    int local = 1;
    for (const char* name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS", "SLURM_NTASKS_PER_NODE"})
        if (const char* e = getenv(name)) {
            local = atoi(e) > 0 ? atoi(e) : 1;
            break;
        }
    int n = hw / local;
    if (n > 32) n = 32;
    return n < 1 ? 1 : n;
}

HostPool* pool_of(vh_handle* h) {
    if (!h->host_pool) h->host_pool = new HostPool(h->host_threads > 0 ? h->host_threads : auto_threads());
    return static_cast<HostPool*>(h->host_pool);
}

constexpr int64_t GATHER_NODES = 16384;  // nodes per work item: 3 x 128 KiB written, a few MB of the vector swept
constexpr int PREFETCH_AHEAD = 24;

// out[c * ld + i] = src[off[c] + slot[i]], i in [i0, i1).  The slots ascend, so the three reads per node walk the
// vector forward; the prefetch covers the jumps between wall-layer nodes (one node in 5 .. 14 on the large meshes).
void gather_range(const double* __restrict__ src, const int32_t* __restrict__ slot, int64_t i0, int64_t i1,
                  const int64_t off[3], double* __restrict__ out, int64_t ld) {
    const double* s0 = src + off[0];
    const double* s1 = src + off[1];
    const double* s2 = src + off[2];
    double* o0 = out;
    double* o1 = out + ld;
    double* o2 = out + 2 * ld;
    const bool interleaved = off[1] == off[0] + 1 && off[2] == off[0] + 2;  // one line serves the three components
    const int64_t ipf = i1 - PREFETCH_AHEAD;
    int64_t i = i0;
    for (; i < ipf; ++i) {
        const int64_t sp = slot[i + PREFETCH_AHEAD];
        __builtin_prefetch(s0 + sp, 0, 0);
        if (!interleaved) {
            __builtin_prefetch(s1 + sp, 0, 0);
            __builtin_prefetch(s2 + sp, 0, 0);
        }
        const int64_t s = slot[i];
        o0[i] = s0[s];
        o1[i] = s1[s];
        o2[i] = s2[s];
    }
    for (; i < i1; ++i) {
        const int64_t s = slot[i];
        o0[i] = s0[s];
        o1[i] = s1[s];
        o2[i] = s2[s];
    }
}

}  // namespace

bool compact_wanted(const vh_handle* h) {
    if (h->compact_mode == 1) return false;
    if (h->compact_mode == 2) return true;
    // auto.  Measured on the GPU box (profiles/r2_bus_variants.md): the gather streams the vector through the host
    // cores at ~10 GB/s per thread up to ~160 GB/s (it touches nearly every cache line once the wall layer is more than
    // a few per cent of the nodes), the plain copy moves every byte at 54 GB/s over PCIe.  With >= 6 threads the
    // gather wins up to a wall-layer share of about a third (1.5x at 20 %, 2.7x at 7 %); at 72 % (the tutorial-size
    // mesh) it loses.  With 3-5 threads (eight ranks on a 32-core host) it breaks even below ~12 % (+4 % at 7 %, -9 % at
    // 9 %, -36 % at 20 %) -- and is still taken there, because on the device the staging kernel is then a transpose at
    // the HBM peak instead of a sector-bound gather (5x less K1 time on the 10 M-tet mesh).  With fewer threads only a
    // very thin wall layer pays.
    static const double limit_env = [] {
        const char* e = getenv("VASP_B200_COMPACT_RATIO");
        return e && *e ? atof(e) : -1.0;
    }();
    const int threads = h->host_threads > 0 ? h->host_threads : auto_threads();
    const double limit = limit_env >= 0.0 ? limit_env : (threads >= 6 ? 0.35 : threads >= 3 ? 0.12 : 0.05);
    const double slots = (double)((h->vec_len + 2) / 3);
    return slots > 0 && (double)h->nWn_pad <= limit * slots;
}

namespace {
void gather_rows(HostPool* pool, const double* const* rows, const double* base, int64_t stride_elems, int64_t n,
                 const int32_t* slot, int64_t nW, const int64_t off[3], double* out, int64_t out_stride_elems) {
    const int64_t chunks = (nW + GATHER_NODES - 1) / GATHER_NODES;
    pool->run(n * chunks, [&](int64_t item) {
        const int64_t r = item / chunks, ch = item % chunks;
        const double* src = rows ? rows[r] : base + r * stride_elems;
        const int64_t i0 = ch * GATHER_NODES, i1 = i0 + GATHER_NODES < nW ? i0 + GATHER_NODES : nW;
        gather_range(src, slot, i0, i1, off, out + r * out_stride_elems, nW);
    });
}
}  // namespace

int compact_gather(vh_handle* h, const double* const* rows, const double* base, int64_t stride_elems, int64_t n,
                   double* out, int64_t out_stride_elems) {
    VH_CHECK(h->nWn_pad > 0 && (int64_t)h->h_wall_slot.size() == h->nWn_pad, VH_ERR_ARG,
             "compaction: call vh_set_velocity_layout first");
    if (n <= 0) return VH_OK;
    const int64_t off[3] = {h->comp_offset[0], h->comp_offset[1], h->comp_offset[2]};
    const auto t0 = std::chrono::steady_clock::now();
    gather_rows(pool_of(h), rows, base, stride_elems, n, h->h_wall_slot.data(), h->nWn_pad, off, out, out_stride_elems);
    h->gather_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return VH_OK;
}

// Handle-free form of the same gather (needs no GPU): host tools and the CPU test-suite use it.
extern "C" int vh_host_gather(const double* const* rows, const double* base, int64_t stride_bytes, int64_t n,
                              const int32_t* slots, int64_t n_slots, const int64_t comp_offset[3], double* out,
                              int64_t out_stride_bytes, int threads) {
    VH_CHECK((rows || base) && slots && comp_offset && out && n >= 0 && n_slots > 0, VH_ERR_ARG,
             "vh_host_gather: null argument");
    VH_CHECK(stride_bytes % 8 == 0 && out_stride_bytes % 8 == 0 && out_stride_bytes >= 24 * n_slots, VH_ERR_ARG,
             "vh_host_gather: bad stride");
    HostPool pool(threads > 0 ? threads : auto_threads());
    gather_rows(&pool, rows, base, stride_bytes / 8, n, slots, n_slots, comp_offset, out, out_stride_bytes / 8);
    return VH_OK;
}

int compact_ring_ensure(vh_handle* h, int64_t slot_bytes) {
    if (h->cstage_bytes >= slot_bytes) return VH_OK;
    for (int i = 0; i < 3; ++i) {
        if (h->h_cstage[i]) cudaFreeHost(h->h_cstage[i]);
        h->h_cstage[i] = nullptr;
        h->cstage_busy[i] = false;
    }
    h->cstage_bytes = 0;
    for (int i = 0; i < 3; ++i) {
        VH_CUDA(cudaHostAlloc((void**)&h->h_cstage[i], (size_t)slot_bytes, cudaHostAllocDefault));
        if (!h->ev_cstage[i]) VH_CUDA(cudaEventCreateWithFlags(&h->ev_cstage[i], cudaEventDisableTiming));
    }
    h->cstage_bytes = slot_bytes;
    return VH_OK;
}

void compact_release(vh_handle* h) {
    if (h->host_pool) delete static_cast<HostPool*>(h->host_pool);
    h->host_pool = nullptr;
    for (int i = 0; i < 3; ++i) {
        if (h->h_cstage[i]) cudaFreeHost(h->h_cstage[i]);
        if (h->ev_cstage[i]) cudaEventDestroy(h->ev_cstage[i]);
        h->h_cstage[i] = nullptr;
        h->ev_cstage[i] = nullptr;
        h->cstage_busy[i] = false;
    }
    h->cstage_bytes = 0;
}
