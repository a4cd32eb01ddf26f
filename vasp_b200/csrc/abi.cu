// extern "C" surface of libvasp_hemo.so (include/vasp_hemo.h): handle lifetime, double-buffered snapshot staging,
// result export and the dlopen'ed NCCL reduction.  No torch, no CPU compute path: every numeric result comes from the
// kernels in k0_precompute.cu / k1_stage.cu / k2_wall.cu.
#include <dlfcn.h>
#include <nccl.h>  // types only; the library is resolved at run time
#include <nvtx3/nvToolsExt.h>  // header-only; ranges show up in nsys / ncu --nvtx
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void vh_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

namespace {

inline void nvtx_push(const char* name) { nvtxRangePushA(name); }
inline void nvtx_pop() { nvtxRangePop(); }
struct NvtxRange {  // a named range over a scope (shows up in nsys / ncu --nvtx)
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

int nccl_load() {
    if (g_nccl.lib) return VH_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    VH_CHECK(g_nccl.lib, VH_ERR_NCCL, "NCCL not found: %s", dlerror());
#define VH_SYM(field, name)                                                          \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.lib, name);                              \
    VH_CHECK(g_nccl.field, VH_ERR_NCCL, "NCCL symbol %s missing", name)
    VH_SYM(GetUniqueId, "ncclGetUniqueId");
    VH_SYM(CommInitRank, "ncclCommInitRank");
    VH_SYM(AllReduce, "ncclAllReduce");
    VH_SYM(AllGather, "ncclAllGather");
    VH_SYM(CommDestroy, "ncclCommDestroy");
    VH_SYM(GetErrorString, "ncclGetErrorString");
#undef VH_SYM
    return VH_OK;
}

#define VH_NCCL(call)                                                                                    \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) {                                                                        \
            vh_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__));      \
            return VH_ERR_NCCL;                                                                          \
        }                                                                                                \
    } while (0)

int ensure_run_buffers(vh_handle* h) {
    const int64_t nF = h->nF;
    if (!h->d_sums_block) {
        // two halves of (15 nF sums + the snapshot count), then the arrival counters of the peer reduction and one
        // word that a reduction sets when a peer never arrived (bounded wait, see k4_peer_indices)
        h->sum_stride = (VH_NSUM * nF + 1 + 15) / 16 * 16;
        const size_t bytes = sizeof(double) * (2 * h->sum_stride + VH_MAX_PEERS + 1);
        VH_CUDA(cudaMalloc(&h->d_sums_block, bytes));
        VH_CUDA(cudaMemset(h->d_sums_block, 0, bytes));
        VH_CUDA(cudaMalloc(&h->d_sums_red, sizeof(double) * (VH_NSUM * nF + 1)));
        h->loop_parity = 0;
        h->d_sums = h->d_sums_block;
        VH_CUDA(cudaMalloc(&h->d_tau_last[0], sizeof(double) * 9 * nF));
        VH_CUDA(cudaMalloc(&h->d_tau_last[1], sizeof(double) * 9 * nF));
        VH_CUDA(cudaMalloc(&h->d_out5, sizeof(double) * 15 * nF));
        VH_CUDA(cudaMalloc(&h->d_out5_peer, sizeof(double) * 15 * nF));
    }
    return VH_OK;
}

// the running sums are zeroed lazily (see vh_begin); anyone who reads them before the first push calls this
int settle_sums(vh_handle* h) {
    vh_join_peer(h);
    if (h->sums_pending_zero) {
        VH_CUDA(cudaMemsetAsync(h->d_sums, 0, sizeof(double) * (VH_NSUM * h->nF + 1), h->s_compute));
        h->sums_pending_zero = false;
    }
    return VH_OK;
}

int check_ready(vh_handle* h, const char* who) {
    VH_CHECK(h, VH_ERR_ARG, "%s: null handle", who);
    VH_CUDA(cudaSetDevice(h->device));
    VH_CHECK(h->nF > 0, VH_ERR_ARG, "%s: call vh_set_mesh first", who);
    VH_CHECK(h->order != 0, VH_ERR_ARG, "%s: call vh_set_velocity_layout first", who);
    VH_CHECK(h->begun, VH_ERR_ARG, "%s: call vh_begin first", who);
    return VH_OK;
}

__global__ void k_set_scalar(double* p, double v) { *p = v; }

__global__ void k_flush(double* __restrict__ p, int64_t n, double v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) p[i] = v;
}

// second half of the flush: stream a buffer larger than L2 through it so that the cache is left full of CLEAN
// foreign lines (a write-only flush leaves 126 MB of dirty lines whose write-back would be billed to the next kernel)
__global__ void k_flush_read(const double* __restrict__ p, int64_t n, double* __restrict__ sink) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t step = (int64_t)gridDim.x * blockDim.x;
    double t = 0.0;
    for (; i < n; i += step) t += p[i];
    if (t == 12345.678) *sink = t;  // never true; keeps the loads alive
}

}  // namespace

extern "C" {

const char* vh_last_error(void) { return g_err; }

int vh_device_count(int* n) {
    VH_CUDA(cudaGetDeviceCount(n));
    return VH_OK;
}

int vh_create(int device, vh_handle** out) {
    VH_CHECK(out, VH_ERR_ARG, "vh_create: null out");
    int ndev = 0;
    VH_CUDA(cudaGetDeviceCount(&ndev));
    VH_CHECK(ndev > 0, VH_ERR_CUDA, "vh_create: no CUDA device visible (this library has no CPU path)");
    VH_CHECK(device >= 0 && device < ndev, VH_ERR_ARG, "vh_create: device %d out of range [0,%d)", device, ndev);
    VH_CUDA(cudaSetDevice(device));
    vh_handle* h = new vh_handle();
    h->device = device;
    cudaDeviceProp prop;
    VH_CUDA(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("VASP_B200_PDL")) h->pdl = atoi(e);
    VH_CUDA(cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking));
    VH_CUDA(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    VH_CUDA(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    VH_CUDA(cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking));
    VH_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    VH_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        VH_CUDA(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
        VH_CUDA(cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming));
        VH_CUDA(cudaEventCreateWithFlags(&h->ev_wss[i], cudaEventDisableTiming));
    }
    VH_CUDA(cudaEventCreate(&h->ev_t0));
    VH_CUDA(cudaEventCreate(&h->ev_t1));
    VH_CUDA(cudaEventCreate(&h->ev_k0));
    VH_CUDA(cudaEventCreate(&h->ev_k1));
    *out = h;
    return VH_OK;
}

int vh_destroy(vh_handle* h) {
    if (!h) return VH_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    vh_nccl_destroy(h);
    k_free_run_buffers(h);
    compact_release(h);
    void* ptrs[] = {h->d_xyz, h->d_tets, h->d_facet_cell, h->d_facet_verts, h->d_bcell_parent, h->d_btopology,
                    h->d_bvert_parent, h->d_facet_local, h->d_bcell_local, h->d_blocal_soa, h->d_glam, h->d_normal,
                    h->d_area, h->d_work, h->d_m_lf, h->d_m_w, h->d_m_mat, h->d_facet_nodes, h->d_row, h->d_wall_slot, h->d_flush, h->d_scalar};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(h->ev_copied[i]);
        cudaEventDestroy(h->ev_consumed[i]);
        cudaEventDestroy(h->ev_wss[i]);
    }
    cudaEventDestroy(h->ev_t0);
    cudaEventDestroy(h->ev_t1);
    cudaEventDestroy(h->ev_k0);
    cudaEventDestroy(h->ev_k1);
    for (auto& e : h->prof_pool) cudaEventDestroy(e);
    for (auto& e : h->batch_events) cudaEventDestroy(e);
    if (h->h_out5) cudaFreeHost(h->h_out5);
    cudaEventDestroy(h->ev_fork);
    cudaEventDestroy(h->ev_join);
    cudaStreamDestroy(h->s_compute);
    cudaStreamDestroy(h->s_copy);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    cudaStreamDestroy(h->s_aux);
    delete h;
    return VH_OK;
}

int vh_set_mesh(vh_handle* h, const double* xyz, int64_t nv, const int64_t* tets, int64_t nc) {
    VH_CHECK(h, VH_ERR_ARG, "vh_set_mesh: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    nvtx_push("K0 mesh precompute");
    const int rc = k0_build_mesh(h, xyz, nv, tets, nc);
    nvtx_pop();
    return rc;
}

int vh_set_velocity_layout(vh_handle* h, int order, const double* refined_xyz, int64_t n_nodes, double tol,
                           const int64_t* node_perm, const int64_t comp_offset[3], int64_t node_stride) {
    VH_CHECK(h, VH_ERR_ARG, "vh_set_velocity_layout: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    VH_CHECK(comp_offset && node_stride >= 1, VH_ERR_ARG, "vh_set_velocity_layout: bad layout");
    // node_perm may be an injection into a longer vector (e.g. the fluid nodes inside a whole-domain array)
    int64_t n_slots = n_nodes;
    if (node_perm) {
        n_slots = 0;
        for (int64_t i = 0; i < n_nodes; ++i) {
            VH_CHECK(node_perm[i] >= 0, VH_ERR_ARG, "vh_set_velocity_layout: node_perm[%lld] is negative", (long long)i);
            if (node_perm[i] + 1 > n_slots) n_slots = node_perm[i] + 1;
        }
    }
    int64_t max_slot = (n_slots - 1) * node_stride;
    for (int c = 0; c < 3; ++c) VH_CHECK(comp_offset[c] >= 0, VH_ERR_ARG, "negative component offset");
    VH_CHECK(max_slot < (1LL << 31), VH_ERR_ARG, "velocity vector too long for int32 gather slots");
    h->node_stride = node_stride;
    for (int c = 0; c < 3; ++c) h->comp_offset[c] = comp_offset[c];
    int64_t top = comp_offset[0];
    for (int c = 1; c < 3; ++c) top = comp_offset[c] > top ? comp_offset[c] : top;
    h->vec_len = top + max_slot + 1;  // doubles per snapshot vector that the gather can touch
    k_free_run_buffers(h);
    nvtx_push("K0 velocity map");
    const int rc = k0_build_velocity_map(h, order, refined_xyz, n_nodes, tol, node_perm, n_slots);
    nvtx_pop();
    return rc;
}

int vh_get_sizes(vh_handle* h, int64_t n[7]) {
    VH_CHECK(h && n, VH_ERR_ARG, "vh_get_sizes: null argument");
    n[0] = h->nF; n[1] = h->nBV; n[2] = h->nW; n[3] = h->nMulti; n[4] = h->ndof; n[5] = h->n_nodes;
    n[6] = h->order ? h->nWn : 0;
    return VH_OK;
}

int vh_get_maps(vh_handle* h, int32_t* facet_cell, int8_t* facet_local, int32_t* facet_verts, int32_t* bcell_parent,
                int32_t* btopology, int32_t* bvert_parent, int8_t* bcell_local, int32_t* facet_nodes) {
    VH_CHECK(h && h->nF > 0, VH_ERR_ARG, "vh_get_maps: call vh_set_mesh first");
    VH_CUDA(cudaSetDevice(h->device));
    const int64_t nF = h->nF;
#define VH_D2H(dst, src, bytes) \
    if (dst) VH_CUDA(cudaMemcpy(dst, src, (size_t)(bytes), cudaMemcpyDeviceToHost))
    VH_D2H(facet_cell, h->d_facet_cell, 4 * nF);
    VH_D2H(facet_local, h->d_facet_local, nF);
    VH_D2H(facet_verts, h->d_facet_verts, 12 * nF);
    VH_D2H(bcell_parent, h->d_bcell_parent, 12 * nF);
    VH_D2H(btopology, h->d_btopology, 12 * nF);
    VH_D2H(bvert_parent, h->d_bvert_parent, 4 * h->nBV);
    VH_D2H(bcell_local, h->d_bcell_local, 3 * nF);
    if (facet_nodes) {
        VH_CHECK(h->order != 0, VH_ERR_ARG, "vh_get_maps: facet_nodes needs vh_set_velocity_layout");
        // device layout is [ndof][nF]; export as [nF][ndof]
        std::vector<int32_t> tmp((size_t)h->ndof * nF);
        VH_CUDA(cudaMemcpy(tmp.data(), h->d_facet_nodes, sizeof(int32_t) * tmp.size(), cudaMemcpyDeviceToHost));
        for (int64_t f = 0; f < nF; ++f)
            for (int k = 0; k < h->ndof; ++k) facet_nodes[f * h->ndof + k] = tmp[(size_t)k * nF + f];
    }
    return VH_OK;
}

int vh_get_geometry(vh_handle* h, double* normal, double* area, double* glam) {
    VH_CHECK(h && h->nF > 0, VH_ERR_ARG, "vh_get_geometry: call vh_set_mesh first");
    VH_CUDA(cudaSetDevice(h->device));
    const int64_t nF = h->nF;
    VH_D2H(area, h->d_area, 8 * nF);
    auto export_rows = [&](double* dst, const double* d_src, int rows) -> int {
        std::vector<double> tmp((size_t)rows * nF);
        VH_CUDA(cudaMemcpy(tmp.data(), d_src, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
        for (int64_t f = 0; f < nF; ++f)
            for (int r = 0; r < rows; ++r) dst[f * rows + r] = tmp[(size_t)r * nF + f];
        return VH_OK;
    };
    if (normal) VH_TRY(export_rows(normal, h->d_normal, 3));
    if (glam) VH_TRY(export_rows(glam, h->d_glam, 12));
    return VH_OK;
}
#undef VH_D2H

int vh_begin(vh_handle* h, double mu, double dt) {
    VH_CHECK(h, VH_ERR_ARG, "vh_begin: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    VH_CHECK(h->nF > 0 && h->order != 0, VH_ERR_ARG, "vh_begin: set mesh and velocity layout first");
    VH_CHECK(dt != 0.0, VH_ERR_ARG, "vh_begin: dt must be non-zero");
    VH_TRY(ensure_run_buffers(h));
    h->mu = mu;
    h->dt = dt;
    h->count = 0;
    h->count_on_device = false;
    h->have_tau_last = false;
    h->kernel_ms = h->h2d_ms = h->gather_ms = 0.0;
    h->h2d_bytes = 0;
    // no memsets on the hot path: the first K3 of the loop overwrites the sums instead of adding to them, and
    // tau_last is only read after a launch has written it
    h->sums_pending_zero = true;
    h->out5_count = -1;
    h->sums_reduced = false;
    h->loop_parity ^= 1;  // alternate halves of the peer-visible block (see common.cuh)
    h->d_sums = h->d_sums_block + (int64_t)h->loop_parity * h->sum_stride;
    h->begun = true;  // stream-ordered: no host sync needed before the first push
    return VH_OK;
}

int vh_set_tuning(vh_handle* h, int64_t batch_snapshots, int64_t chunk_snapshots) {
    VH_CHECK(h, VH_ERR_ARG, "vh_set_tuning: null handle");
    VH_CHECK(batch_snapshots >= 0 && chunk_snapshots >= 0, VH_ERR_ARG, "vh_set_tuning: negative value");
    if (batch_snapshots != h->batch_snapshots) {
        for (int i = 0; i < 2; ++i) {
            if (h->d_stage[i]) cudaFree(h->d_stage[i]);
            if (h->d_wss_stage[i]) cudaFree(h->d_wss_stage[i]);
            h->d_stage[i] = h->d_wss_stage[i] = nullptr;
        }
        h->stage_cap = h->wss_stage_cap = 0;
        h->stage_row_bytes = 0;
        if (h->d_W) cudaFree(h->d_W);  // the staged block is re-sized on the next launch
        h->d_W = nullptr;
        h->w_ld = 0;
    }
    h->batch_snapshots = batch_snapshots;
    h->chunk_snapshots = chunk_snapshots;
    return VH_OK;
}

int vh_set_wss_layout(vh_handle* h, int64_t ld, int64_t col0) {
    VH_CHECK(h, VH_ERR_ARG, "vh_set_wss_layout: null handle");
    VH_CHECK(ld >= 0 && col0 >= 0 && (ld == 0 ? col0 == 0 : col0 <= ld), VH_ERR_ARG,
             "vh_set_wss_layout: ld %lld, first column %lld", (long long)ld, (long long)col0);
    h->wss_ld = ld;
    h->wss_col = col0;
    return VH_OK;
}

// D2H of the five result fields: one copy into a pinned staging buffer (user arrays are usually pageable numpy
// memory, where every cudaMemcpyAsync degenerates into a staged synchronous copy), then host memcpy
static int export_out5(vh_handle* h, double* const outs[5], const double* d_src = nullptr, cudaStream_t st = nullptr) {
    const int64_t n3 = 3 * h->nF;
    bool any = false;
    for (int i = 0; i < 5; ++i) any = any || outs[i];
    if (!any) return VH_OK;  // results stay in HBM, the call stays asynchronous
    if (!st) st = h->s_compute;
    if (!h->h_out5) VH_CUDA(cudaHostAlloc((void**)&h->h_out5, sizeof(double) * 5 * n3, cudaHostAllocDefault));
    VH_CUDA(cudaMemcpyAsync(h->h_out5, d_src ? d_src : h->d_out5, sizeof(double) * 5 * n3, cudaMemcpyDeviceToHost, st));
    VH_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 5; ++i)
        if (outs[i]) memcpy(outs[i], h->h_out5 + i * n3, sizeof(double) * n3);
    return VH_OK;
}

static int prev_mode_for(vh_handle* h, int flags, const char* who, int* mode) {
    if (flags & VH_PUSH_HALO_FIRST) {
        *mode = 2;
    } else if (flags & VH_PUSH_GLOBAL_FIRST) {
        *mode = 0;
    } else {
        VH_CHECK(h->have_tau_last, VH_ERR_ARG,
                 "%s: first push of a time loop needs VH_PUSH_GLOBAL_FIRST or VH_PUSH_HALO_FIRST", who);
        *mode = 1;
    }
    return VH_OK;
}

// Device-resident snapshots: whole vectors (K1 gathers the wall layer) or compact blocks (K1 transposes).
static int push_device(vh_handle* h, const char* who, const double* d_u, int64_t n_snap, int64_t stride_bytes, int flags,
                       double* d_wss_out, bool dense) {
    NvtxRange range(dense ? "push resident compact blocks (K1 transpose, K2, K3)" : "push resident vectors (K1 gather, K2, K3)");
    VH_TRY(check_ready(h, who));
    VH_CHECK(d_u && n_snap > 0, VH_ERR_ARG, "%s: nothing to push", who);
    const int64_t row_elems = dense ? 3 * h->nWn_pad : h->vec_len;
    VH_CHECK(stride_bytes % 8 == 0 && stride_bytes >= 8 * row_elems, VH_ERR_ARG,
             "%s: stride %lld smaller than a %s (%lld doubles)", who, (long long)stride_bytes,
             dense ? "compact block" : "vector", (long long)row_elems);
    int mode = 0;
    VH_TRY(prev_mode_for(h, flags, who, &mode));
    const int64_t stride = stride_bytes / 8;
    VH_CHECK(mode != 2 || n_snap >= 2, VH_ERR_ARG, "halo push needs at least one real snapshot after the halo");
    const int64_t n_real = n_snap - (mode == 2 ? 1 : 0);
    if (d_wss_out && h->wss_ld > 0) {
        VH_CHECK(h->wss_col + n_real <= h->wss_ld, VH_ERR_ARG,
                 "%s: WSS matrix has %lld columns, %lld already written, %lld more pushed", who,
                 (long long)h->wss_ld, (long long)h->wss_col, (long long)n_real);
        d_wss_out += h->wss_col;
        h->wss_col += n_real;
    }
    return k2_launch(h, mode == 2 ? d_u + stride : d_u, n_real, stride, mode, d_wss_out, h->wss_ld, dense);
}

int vh_push_snapshots_device(vh_handle* h, const double* d_u, int64_t n_snap, int64_t stride_bytes, int flags,
                             double* d_wss_out) {
    return push_device(h, "vh_push_snapshots_device", d_u, n_snap, stride_bytes, flags, d_wss_out, false);
}

int vh_push_compact_device(vh_handle* h, const double* d_c, int64_t n_snap, int64_t stride_bytes, int flags,
                           double* d_wss_out) {
    return push_device(h, "vh_push_compact_device", d_c, n_snap, stride_bytes, flags, d_wss_out, true);
}

// Host snapshots -> device, double buffered against the kernels.  Three sources:
//   SRC_FULL     whole vectors cross the bus, K1 gathers the wall layer on the device
//   SRC_GATHER   whole vectors in host memory, the wall layer is gathered on the host (compact.cu) piece by piece into
//                a pinned ring while the previous piece crosses the bus; K1 transposes
//   SRC_COMPACT  the caller already holds compact blocks (vh_compact_snapshots / vh_compact_rows)
enum { SRC_FULL = 0, SRC_GATHER = 1, SRC_COMPACT = 2 };

static int push_host(vh_handle* h, const char* who, const double* u, int64_t n_snap, int64_t stride_bytes, int flags,
                     double* wss_out, int kind) {
    NvtxRange range(kind == SRC_GATHER ? "push host vectors (host gather, H2D, K1-K3)"
                    : kind == SRC_COMPACT ? "push host compact blocks (H2D, K1-K3)" : "push host vectors (H2D, K1-K3)");
    VH_TRY(check_ready(h, who));
    VH_CHECK(u && n_snap > 0, VH_ERR_ARG, "%s: nothing to push", who);
    const bool dense = kind != SRC_FULL;
    const int64_t row_elems = dense ? 3 * h->nWn_pad : h->vec_len;  // doubles per snapshot on the device
    const int64_t row_bytes = 8 * row_elems;
    const int64_t src_bytes = kind == SRC_COMPACT ? row_bytes : 8 * h->vec_len;
    VH_CHECK(stride_bytes % 8 == 0 && stride_bytes >= src_bytes, VH_ERR_ARG,
             "%s: stride %lld smaller than a %s (%lld bytes)", who, (long long)stride_bytes,
             kind == SRC_COMPACT ? "compact block" : "vector", (long long)src_bytes);
    int mode = 0;
    VH_TRY(prev_mode_for(h, flags, who, &mode));
    const bool halo = mode == 2;
    VH_CHECK(!halo || n_snap >= 2, VH_ERR_ARG, "halo push needs at least one real snapshot after the halo");
    const int64_t nF = h->nF;
    const bool wss_matrix = wss_out && h->wss_ld > 0;
    VH_CHECK(!wss_matrix || h->wss_col + n_snap - (halo ? 1 : 0) <= h->wss_ld, VH_ERR_ARG,
             "%s: WSS matrix has %lld columns, %lld already written, %lld more pushed", who,
             (long long)h->wss_ld, (long long)h->wss_col, (long long)(n_snap - (halo ? 1 : 0)));

    // stage capacity: user value, else half of the push -- the batches then shrink geometrically (1/2, 1/4, 1/8, 1/8):
    // large copies move faster over PCIe (55 GB/s for 72 MB in one piece against 52 for eight 9 MB pieces, measured),
    // a small last batch leaves little kernel time exposed behind the last copy; at least 32 snapshots per batch, at
    // most what fits in ~30 % of free memory / 2 buffers or 8 GiB
    if (h->stage_cap != 0 && h->stage_row_bytes != row_bytes) {  // the other kind of source was pushed before
        for (int i = 0; i < 2; ++i) {
            if (h->d_stage[i]) cudaFree(h->d_stage[i]);
            h->d_stage[i] = nullptr;
        }
        h->stage_cap = 0;
    }
    if (h->stage_cap == 0) {
        int64_t cap = h->batch_snapshots;
        if (cap <= 0) {
            size_t free_b = 0, total_b = 0;
            VH_CUDA(cudaMemGetInfo(&free_b, &total_b));
            int64_t per_snap = row_bytes + (wss_out ? 72 * nF : 0);
            int64_t budget = (int64_t)(0.3 * (double)free_b) / 2;
            if (budget > (8LL << 30)) budget = 8LL << 30;
            cap = (n_snap + 1) / 2;
            if (cap < 32) cap = 32;
            if (cap > budget / per_snap) cap = budget / per_snap;
            if (cap > n_snap) cap = n_snap;
        }
        if (cap < 2) cap = 2;
        for (int i = 0; i < 2; ++i) VH_CUDA(cudaMalloc(&h->d_stage[i], (size_t)(cap * row_bytes)));
        h->stage_cap = cap;
        h->stage_row_bytes = row_bytes;
    }
    if (wss_out && h->wss_stage_cap < h->stage_cap) {
        for (int i = 0; i < 2; ++i) {
            if (h->d_wss_stage[i]) cudaFree(h->d_wss_stage[i]);
            VH_CUDA(cudaMalloc(&h->d_wss_stage[i], (size_t)(h->stage_cap * 72 * nF)));
        }
        h->wss_stage_cap = h->stage_cap;
    }
    // gather pieces: at most ~32 MiB of compact blocks each (a few ms of bus time), and at least four per batch so
    // that gathering, copying and computing overlap even inside a single batch
    int64_t piece = 1;
    if (kind == SRC_GATHER) {
        piece = (32LL << 20) / row_bytes;
        const int64_t quarter = (h->stage_cap + 3) / 4;
        if (piece > quarter) piece = quarter;
        if (piece < 1) piece = 1;
        VH_TRY(compact_ring_ensure(h, piece * row_bytes));
    }

    // (h2d start, h2d stop, kernel start, kernel stop) per batch, from a pool that lives with the handle: creating
    // and destroying events inside the call cost more than a batch's kernels on small meshes
    size_t ev_used = 0;
    auto new_event = [&](cudaEvent_t* e) -> int {
        if (ev_used == h->batch_events.size()) {
            cudaEvent_t ne;
            VH_CUDA(cudaEventCreate(&ne));
            h->batch_events.push_back(ne);
        }
        *e = h->batch_events[ev_used++];
        return VH_OK;
    };
    int64_t pos = 0, real_done = 0;
    int b = 0, ring = 0;
    bool first = true;
    int rc = VH_OK;
    while (pos < n_snap && rc == VH_OK) {
        const int buf = b & 1;
        int64_t nb = n_snap - pos;
        if (h->batch_snapshots <= 0) {  // auto: halve the remainder down to an eighth of the push
            int64_t floor_nb = (n_snap + 7) / 8;
            if (floor_nb < 32) floor_nb = 32;
            int64_t half = (nb + 1) / 2;
            if (half < floor_nb) half = floor_nb;
            if (nb - half >= floor_nb) nb = half;  // else: take the rest in one piece
        }
        if (nb > h->stage_cap) nb = h->stage_cap;
        cudaEvent_t c0, c1, k0, k1;
        if ((rc = new_event(&c0)) || (rc = new_event(&c1)) || (rc = new_event(&k0)) || (rc = new_event(&k1))) break;
        // copy stream: wait until the kernel that last read this buffer is done, then H2D
        cudaStreamWaitEvent(h->s_copy, h->ev_consumed[buf], 0);
        cudaEventRecord(c0, h->s_copy);
        cudaError_t ce = cudaSuccess;
        if (kind == SRC_GATHER) {
            nvtx_push("gather+h2d batch");
            for (int64_t p0 = 0; p0 < nb && rc == VH_OK && ce == cudaSuccess; p0 += piece) {
                const int64_t np = nb - p0 < piece ? nb - p0 : piece;
                if (h->cstage_busy[ring]) cudaEventSynchronize(h->ev_cstage[ring]);  // its last copy has left the slot
                rc = compact_gather(h, nullptr, (const double*)((const char*)u + (pos + p0) * stride_bytes),
                                    stride_bytes / 8, np, h->h_cstage[ring], row_elems);
                if (rc != VH_OK) break;
                ce = cudaMemcpyAsync(h->d_stage[buf] + p0 * row_elems, h->h_cstage[ring], (size_t)(np * row_bytes),
                                     cudaMemcpyHostToDevice, h->s_copy);
                cudaEventRecord(h->ev_cstage[ring], h->s_copy);
                h->cstage_busy[ring] = true;
                ring = (ring + 1) % 3;
            }
            nvtx_pop();
            if (rc != VH_OK) break;
        } else {
            // contiguous rows (the usual case: one block of u.h5 vectors) go as one flat copy
            ce = stride_bytes == row_bytes
                     ? cudaMemcpyAsync(h->d_stage[buf], (const char*)u + pos * stride_bytes, (size_t)(nb * row_bytes),
                                       cudaMemcpyHostToDevice, h->s_copy)
                     : cudaMemcpy2DAsync(h->d_stage[buf], (size_t)row_bytes, (const char*)u + pos * stride_bytes,
                                         (size_t)stride_bytes, (size_t)row_bytes, (size_t)nb, cudaMemcpyHostToDevice,
                                         h->s_copy);
        }
        if (ce != cudaSuccess) {
            vh_set_error("%s: H2D copy failed: %s", who, cudaGetErrorString(ce));
            rc = VH_ERR_CUDA;
            break;
        }
        h->h2d_bytes += nb * row_bytes;
        cudaEventRecord(c1, h->s_copy);
        cudaEventRecord(h->ev_copied[buf], h->s_copy);
        // compute stream
        cudaStreamWaitEvent(h->s_compute, h->ev_copied[buf], 0);
        if (wss_out) cudaStreamWaitEvent(h->s_compute, h->ev_wss[buf], 0);  // previous D2H of this wss buffer done
        const double* d_u = h->d_stage[buf];
        int64_t n_real = nb;
        int pm = first ? mode : 1;
        if (first && halo) {
            d_u += row_elems;
            n_real = nb - 1;
        }
        cudaEventRecord(k0, h->s_compute);
        // per-snapshot vectors, or a (9 nF) x stage_cap time-major block whose first n_real columns are copied out
        rc = k2_launch(h, d_u, n_real, row_elems, pm, wss_out ? h->d_wss_stage[buf] : nullptr,
                       wss_matrix ? h->stage_cap : 0, dense);
        cudaEventRecord(k1, h->s_compute);
        cudaEventRecord(h->ev_consumed[buf], h->s_compute);
        if (rc == VH_OK && wss_out && n_real > 0) {
            // on its own stream: on the copy stream the next batch's H2D would queue behind this D2H, which waits for
            // this batch's kernels, and nothing would overlap (PCIe is full duplex, the GPU has several copy engines)
            cudaStreamWaitEvent(h->s_d2h, h->ev_consumed[buf], 0);
            ce = wss_matrix
                     ? cudaMemcpy2DAsync(wss_out + h->wss_col + real_done, (size_t)(8 * h->wss_ld), h->d_wss_stage[buf],
                                         (size_t)(8 * h->stage_cap), (size_t)(8 * n_real), (size_t)(9 * nF),
                                         cudaMemcpyDeviceToHost, h->s_d2h)
                     : cudaMemcpyAsync(wss_out + real_done * 9 * nF, h->d_wss_stage[buf], (size_t)(n_real * 72 * nF),
                                       cudaMemcpyDeviceToHost, h->s_d2h);
            if (ce != cudaSuccess) {
                vh_set_error("%s: D2H copy failed: %s", who, cudaGetErrorString(ce));
                rc = VH_ERR_CUDA;
            }
            cudaEventRecord(h->ev_wss[buf], h->s_d2h);
        }
        real_done += n_real;
        pos += nb;
        first = false;
        ++b;
    }
    if (wss_matrix) h->wss_col += real_done;
    cudaError_t e1 = cudaStreamSynchronize(h->s_copy), e2 = cudaStreamSynchronize(h->s_compute);
    const cudaError_t e3 = cudaStreamSynchronize(h->s_d2h);
    for (int i = 0; i < 3; ++i) h->cstage_busy[i] = false;  // the copy stream has drained
    if (e1 == cudaSuccess) e1 = e3;
    if (rc == VH_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        vh_set_error("%s: %s", who, cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        rc = VH_ERR_CUDA;
    }
    for (size_t i = 0; i + 3 < ev_used && rc == VH_OK; i += 4) {
        float ms = 0.f;
        const std::vector<cudaEvent_t>& evs = h->batch_events;
        if (cudaEventElapsedTime(&ms, evs[i], evs[i + 1]) == cudaSuccess) h->h2d_ms += ms;
        if (cudaEventElapsedTime(&ms, evs[i + 2], evs[i + 3]) == cudaSuccess) h->kernel_ms += ms;
    }
    return rc;
}

int vh_push_snapshots(vh_handle* h, const double* u, int64_t n_snap, int64_t stride_bytes, int flags, double* wss_out) {
    VH_CHECK(h, VH_ERR_ARG, "vh_push_snapshots: null handle");
    return push_host(h, "vh_push_snapshots", u, n_snap, stride_bytes, flags, wss_out,
                     h->order != 0 && compact_wanted(h) ? SRC_GATHER : SRC_FULL);
}

int vh_push_compact(vh_handle* h, const double* c, int64_t n_snap, int64_t stride_bytes, int flags, double* wss_out) {
    VH_CHECK(h, VH_ERR_ARG, "vh_push_compact: null handle");
    return push_host(h, "vh_push_compact", c, n_snap, stride_bytes, flags, wss_out, SRC_COMPACT);
}

// ---- wall-layer compaction (host side, compact.cu) ------------------------------------------------------------------
int vh_set_host_compaction(vh_handle* h, int mode, int threads) {
    VH_CHECK(h, VH_ERR_ARG, "vh_set_host_compaction: null handle");
    VH_CHECK(mode >= 0 && mode <= 2 && threads >= 0, VH_ERR_ARG, "vh_set_host_compaction: mode 0|1|2, threads >= 0");
    h->compact_mode = mode;
    if (threads != h->host_threads) {
        h->host_threads = threads;
        compact_release(h);  // the pool is re-made with the new size on the next gather
    }
    return VH_OK;
}

int vh_get_compact_info(vh_handle* h, int64_t* compact_len, int* active) {
    VH_CHECK(h && h->order != 0, VH_ERR_ARG, "vh_get_compact_info: call vh_set_velocity_layout first");
    if (compact_len) *compact_len = 3 * h->nWn_pad;
    if (active) *active = compact_wanted(h) ? 1 : 0;
    return VH_OK;
}

int vh_get_wall_slots(vh_handle* h, int64_t* slots) {
    VH_CHECK(h && h->order != 0 && slots, VH_ERR_ARG, "vh_get_wall_slots: call vh_set_velocity_layout first");
    for (int64_t i = 0; i < h->nWn; ++i) slots[i] = h->h_wall_slot[(size_t)i];
    return VH_OK;
}

int vh_compact_snapshots(vh_handle* h, const double* u, int64_t n_snap, int64_t stride_bytes, double* out) {
    VH_CHECK(h && h->order != 0, VH_ERR_ARG, "vh_compact_snapshots: call vh_set_velocity_layout first");
    VH_CHECK(u && out && n_snap > 0, VH_ERR_ARG, "vh_compact_snapshots: null argument");
    VH_CHECK(stride_bytes % 8 == 0 && stride_bytes >= 8 * h->vec_len, VH_ERR_ARG,
             "vh_compact_snapshots: stride %lld smaller than a vector (%lld bytes)", (long long)stride_bytes,
             (long long)(8 * h->vec_len));
    return compact_gather(h, nullptr, u, stride_bytes / 8, n_snap, out, 3 * h->nWn_pad);
}

int vh_compact_rows(vh_handle* h, const double* const* rows, int64_t n_snap, double* out) {
    VH_CHECK(h && h->order != 0, VH_ERR_ARG, "vh_compact_rows: call vh_set_velocity_layout first");
    VH_CHECK(rows && out && n_snap > 0, VH_ERR_ARG, "vh_compact_rows: null argument");
    return compact_gather(h, rows, nullptr, 0, n_snap, out, 3 * h->nWn_pad);
}

int vh_get_io_stats(vh_handle* h, double* gather_ms, int64_t* h2d_bytes) {
    VH_CHECK(h, VH_ERR_ARG, "vh_get_io_stats: null handle");
    if (gather_ms) *gather_ms = h->gather_ms;
    if (h2d_bytes) *h2d_bytes = h->h2d_bytes;
    return VH_OK;
}

int vh_get_sums(vh_handle* h, double* sums, int64_t* count) {
    VH_TRY(check_ready(h, "vh_get_sums"));
    VH_TRY(settle_sums(h));
    VH_CUDA(cudaStreamSynchronize(h->s_compute));
    VH_CUDA(cudaStreamSynchronize(h->s_aux));
    const double* src = h->sums_reduced ? h->d_sums_red : h->d_sums;  // after a fused peer reduction: the global sums
    if (sums) VH_CUDA(cudaMemcpy(sums, src, sizeof(double) * VH_NSUM * h->nF, cudaMemcpyDeviceToHost));
    if (h->count_on_device) {  // left there by the all-reduce / peer reduction, which do not synchronise
        double cnt = 0.0;
        VH_CUDA(cudaMemcpy(&cnt, src + VH_NSUM * h->nF, sizeof(double), cudaMemcpyDeviceToHost));
        h->count = (int64_t)(cnt + 0.5);
        h->count_on_device = false;
    }
    if (count) *count = h->count;
    return VH_OK;
}

int vh_set_sums(vh_handle* h, const double* sums, int64_t count) {
    VH_TRY(check_ready(h, "vh_set_sums"));
    VH_CHECK(sums, VH_ERR_ARG, "vh_set_sums: null sums");
    h->sums_pending_zero = false;
    h->sums_reduced = false;
    h->out5_count = -1;
    vh_join_peer(h);
    VH_CUDA(cudaStreamSynchronize(h->s_compute));
    VH_CUDA(cudaMemcpy(h->d_sums, sums, sizeof(double) * VH_NSUM * h->nF, cudaMemcpyHostToDevice));
    h->count = count;
    h->count_on_device = false;
    return VH_OK;
}

int vh_sums_device_ptr(vh_handle* h, double** d_sums) {
    VH_TRY(check_ready(h, "vh_sums_device_ptr"));
    VH_TRY(settle_sums(h));
    *d_sums = h->d_sums;
    return VH_OK;
}

int vh_get_tau_last(vh_handle* h, double* tau) {
    VH_TRY(check_ready(h, "vh_get_tau_last"));
    VH_CHECK(tau, VH_ERR_ARG, "vh_get_tau_last: null output");
    VH_CUDA(cudaStreamSynchronize(h->s_compute));
    const int64_t nF = h->nF;
    std::vector<double> tmp((size_t)9 * nF);
    VH_CUDA(cudaMemcpy(tmp.data(), h->d_tau_last[h->tau_cur], sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
    for (int64_t f = 0; f < nF; ++f)
        for (int r = 0; r < 9; ++r) tau[f * 9 + r] = tmp[(size_t)r * nF + f];
    return VH_OK;
}

int vh_finalize(vh_handle* h, int64_t n_total, double* tawss, double* osi, double* rrt, double* ecap, double* twssg) {
    NvtxRange range("finalize (indices, D2H)");
    VH_TRY(check_ready(h, "vh_finalize"));
    VH_CHECK(n_total > 0, VH_ERR_ARG, "vh_finalize: n_total must be positive");
    VH_TRY(settle_sums(h));
    // the fold of the last launch already evaluated the indices for the snapshots it had seen
    if (h->out5_count != n_total) VH_TRY(k4_finalize(h, n_total, h->d_out5));
    double* const outs[5] = {tawss, osi, rrt, ecap, twssg};
    // with no host outputs the call stays asynchronous (device-resident timing); results remain in HBM
    return export_out5(h, outs);
}

// A fused peer reduction whose wait ran out marks the word behind the arrival counters; every call that synchronises
// after one looks at it, so a lost rank ends in an error on the survivors instead of a hung GPU.
static int peer_check(vh_handle* h) {
    if (!h->peer_unchecked) return VH_OK;
    VH_CUDA(cudaStreamSynchronize(h->s_aux));
    uint64_t lost = 0;
    VH_CUDA(cudaMemcpyAsync(&lost, h->d_sums_block + 2 * h->sum_stride + VH_MAX_PEERS, sizeof(lost),
                            cudaMemcpyDeviceToHost, h->s_compute));
    VH_CUDA(cudaStreamSynchronize(h->s_compute));
    h->peer_unchecked = false;
    VH_CHECK(lost == 0, VH_ERR_NCCL,
             "vh_peer_reduce_finalize: rank %d never signalled epoch %llu within %.0f s (results are invalid)",
             (int)(lost - 1), (unsigned long long)h->peer_epoch, VH_PEER_WAIT_NS * 1e-9);
    return VH_OK;
}

int vh_sync(vh_handle* h) {
    VH_CHECK(h, VH_ERR_ARG, "vh_sync: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    VH_CUDA(cudaStreamSynchronize(h->s_copy));
    VH_CUDA(cudaStreamSynchronize(h->s_compute));
    VH_CUDA(cudaStreamSynchronize(h->s_aux));
    return peer_check(h);
}

int vh_get_timers(vh_handle* h, double* kernel_ms, double* h2d_ms, int64_t* launches) {
    VH_CHECK(h, VH_ERR_ARG, "vh_get_timers: null handle");
    if (kernel_ms) *kernel_ms = h->kernel_ms;
    if (h2d_ms) *h2d_ms = h->h2d_ms;
    if (launches) *launches = h->launches;
    return VH_OK;
}

int vh_set_profile(vh_handle* h, int on) {
    VH_CHECK(h, VH_ERR_ARG, "vh_set_profile: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    if (on && h->prof_pool.empty()) {
        h->prof_pool.resize(3 * 512);
        for (auto& e : h->prof_pool) VH_CUDA(cudaEventCreate(&e));
    }
    h->profile = on != 0;
    h->prof_used = 0;
    return VH_OK;
}

int vh_get_kernel_profile(vh_handle* h, double* k1_ms, double* k2_ms, int64_t* launches) {
    VH_CHECK(h, VH_ERR_ARG, "vh_get_kernel_profile: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    VH_CUDA(cudaStreamSynchronize(h->s_compute));
    double t1 = 0.0, t2 = 0.0;
    for (size_t i = 0; i + 2 < h->prof_used; i += 3) {
        float ms = 0.f;
        VH_CUDA(cudaEventElapsedTime(&ms, h->prof_pool[i], h->prof_pool[i + 1]));
        t1 += ms;
        VH_CUDA(cudaEventElapsedTime(&ms, h->prof_pool[i + 1], h->prof_pool[i + 2]));
        t2 += ms;
    }
    if (k1_ms) *k1_ms = t1;
    if (k2_ms) *k2_ms = t2;
    if (launches) *launches = (int64_t)(h->prof_used / 3);
    h->prof_used = 0;
    return VH_OK;
}

int vh_timer_start(vh_handle* h) {
    VH_CHECK(h, VH_ERR_ARG, "vh_timer_start: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    VH_CUDA(cudaEventRecord(h->ev_t0, h->s_compute));
    return VH_OK;
}

int vh_timer_stop(vh_handle* h, double* ms) {
    VH_CHECK(h && ms, VH_ERR_ARG, "vh_timer_stop: null argument");
    VH_CUDA(cudaSetDevice(h->device));
    if (h->peer_pending) VH_CUDA(cudaStreamWaitEvent(h->s_compute, h->ev_join, 0));  // the timed region ends with the
                                                                                       // last cross-GPU reduction
    VH_CUDA(cudaEventRecord(h->ev_t1, h->s_compute));
    VH_CUDA(cudaEventSynchronize(h->ev_t1));
    float f = 0.f;
    VH_CUDA(cudaEventElapsedTime(&f, h->ev_t0, h->ev_t1));
    *ms = f;
    return VH_OK;
}

int vh_alloc_pinned(void** p, int64_t nbytes) {
    VH_CHECK(p && nbytes > 0, VH_ERR_ARG, "vh_alloc_pinned: bad argument");
    VH_CUDA(cudaHostAlloc(p, (size_t)nbytes, cudaHostAllocDefault));
    return VH_OK;
}

int vh_free_pinned(void* p) {
    if (p) VH_CUDA(cudaFreeHost(p));
    return VH_OK;
}

int vh_alloc_device(vh_handle* h, void** d_p, int64_t nbytes) {
    VH_CHECK(h && d_p && nbytes > 0, VH_ERR_ARG, "vh_alloc_device: bad argument");
    VH_CUDA(cudaSetDevice(h->device));
    VH_CUDA(cudaMalloc(d_p, (size_t)nbytes));
    return VH_OK;
}

int vh_free_device(vh_handle* h, void* d_p) {
    VH_CHECK(h, VH_ERR_ARG, "vh_free_device: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    if (d_p) VH_CUDA(cudaFree(d_p));
    return VH_OK;
}

int vh_memcpy_h2d(vh_handle* h, void* d_dst, const void* src, int64_t nbytes) {
    VH_CHECK(h, VH_ERR_ARG, "vh_memcpy_h2d: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    VH_CUDA(cudaMemcpyAsync(d_dst, src, (size_t)nbytes, cudaMemcpyHostToDevice, h->s_compute));
    VH_CUDA(cudaStreamSynchronize(h->s_compute));
    return VH_OK;
}

int vh_memcpy_d2h(vh_handle* h, void* dst, const void* d_src, int64_t nbytes) {
    VH_CHECK(h, VH_ERR_ARG, "vh_memcpy_d2h: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    VH_CUDA(cudaMemcpyAsync(dst, d_src, (size_t)nbytes, cudaMemcpyDeviceToHost, h->s_compute));
    VH_CUDA(cudaStreamSynchronize(h->s_compute));
    return VH_OK;
}

int vh_flush_l2(vh_handle* h) {
    VH_CHECK(h, VH_ERR_ARG, "vh_flush_l2: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    if (!h->d_flush) {
        h->flush_bytes = 512LL << 20;  // two halves of 256 MiB, each 2x the 126 MB L2: one written, one read
        
        VH_CUDA(cudaMalloc(&h->d_flush, (size_t)h->flush_bytes));
    }
    static double tick = 0.0;
    tick += 1.0;
    const int64_t half = h->flush_bytes / 16;  // doubles in each half of the scratch buffer
    k_flush<<<h->sm_count * 8, 256, 0, h->s_compute>>>((double*)h->d_flush, half, tick);
    k_flush_read<<<h->sm_count * 8, 256, 0, h->s_compute>>>((const double*)h->d_flush + half, half, (double*)h->d_flush);
    VH_CUDA(cudaGetLastError());
    return VH_OK;
}

int vh_mem_info(vh_handle* h, int64_t* free_bytes, int64_t* total_bytes) {
    VH_CHECK(h, VH_ERR_ARG, "vh_mem_info: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    size_t f = 0, t = 0;
    VH_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return VH_OK;
}

// ---- NCCL ---------------------------------------------------------------------------------------------------------
int vh_nccl_unique_id(char id[128]) {
    VH_TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId uid;
    VH_NCCL(g_nccl.GetUniqueId(&uid));
    memcpy(id, &uid, 128);
    return VH_OK;
}

int vh_nccl_init(vh_handle* h, const char id[128], int rank, int world) {
    VH_CHECK(h && id, VH_ERR_ARG, "vh_nccl_init: null argument");
    VH_CHECK(world >= 1 && rank >= 0 && rank < world, VH_ERR_ARG, "vh_nccl_init: bad rank/world %d/%d", rank, world);
    VH_TRY(nccl_load());
    VH_CUDA(cudaSetDevice(h->device));
    vh_nccl_destroy(h);
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclComm_t comm;
    VH_NCCL(g_nccl.CommInitRank(&comm, world, uid, rank));
    h->nccl_comm = comm;
    h->rank = rank;
    h->world = world;
    if (!h->d_scalar) VH_CUDA(cudaMalloc(&h->d_scalar, sizeof(double) * 2));
    return VH_OK;
}

int vh_nccl_allreduce_sums(vh_handle* h) {
    VH_TRY(check_ready(h, "vh_nccl_allreduce_sums"));
    VH_CHECK(h->nccl_comm, VH_ERR_NCCL, "vh_nccl_allreduce_sums: call vh_nccl_init first");
    VH_TRY(settle_sums(h));  // also orders s_compute behind a fused reduction still in flight
    // one collective, stream-ordered, no host synchronisation: the snapshot count rides behind the 15 * nF sums
    const int64_t n = VH_NSUM * h->nF;
    if (!h->count_on_device) {
        k_set_scalar<<<1, 1, 0, h->s_compute>>>(h->d_sums + n, (double)h->count);
        VH_CUDA(cudaGetLastError());
    }
    VH_NCCL(g_nccl.AllReduce(h->d_sums, h->d_sums, (size_t)(n + 1), ncclDouble, ncclSum, (ncclComm_t)h->nccl_comm,
                             h->s_compute));
    h->count_on_device = true;
    h->out5_count = -1;  // the sums are global now
    h->launches += 1;
    return VH_OK;
}

int vh_nccl_allreduce_max(vh_handle* h, double* value) {
    VH_CHECK(h && value, VH_ERR_ARG, "vh_nccl_allreduce_max: null argument");
    VH_CHECK(h->nccl_comm, VH_ERR_NCCL, "vh_nccl_allreduce_max: call vh_nccl_init first");
    VH_CUDA(cudaSetDevice(h->device));
    VH_CUDA(cudaMemcpyAsync(h->d_scalar, value, sizeof(double), cudaMemcpyHostToDevice, h->s_compute));
    VH_NCCL(g_nccl.AllReduce(h->d_scalar, h->d_scalar, 1, ncclDouble, ncclMax, (ncclComm_t)h->nccl_comm, h->s_compute));
    VH_CUDA(cudaMemcpyAsync(value, h->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, h->s_compute));
    VH_CUDA(cudaStreamSynchronize(h->s_compute));
    return VH_OK;
}

int vh_nccl_barrier(vh_handle* h) {
    double v = 0.0;
    return vh_nccl_allreduce_max(h, &v);
}

static void peer_close(vh_handle* h) {
    for (int q = 0; q < VH_MAX_PEERS; ++q) {
        if (h->peer_block[q] && q != h->rank) cudaIpcCloseMemHandle(h->peer_block[q]);
        h->peer_block[q] = nullptr;
    }
    h->peer_ready = false;
}

int vh_peer_init(vh_handle* h) {
    VH_CHECK(h, VH_ERR_ARG, "vh_peer_init: null handle");
    VH_CUDA(cudaSetDevice(h->device));
    VH_CHECK(h->nccl_comm, VH_ERR_NCCL, "vh_peer_init: call vh_nccl_init first (it carries the handle exchange)");
    VH_CHECK(h->nF > 0 && h->order != 0, VH_ERR_ARG, "vh_peer_init: set mesh and velocity layout first");
    VH_CHECK(h->world <= VH_MAX_PEERS, VH_ERR_ARG, "vh_peer_init: at most %d ranks", VH_MAX_PEERS);
    peer_close(h);
    VH_TRY(ensure_run_buffers(h));
    VH_CUDA(cudaStreamSynchronize(h->s_compute));
    cudaIpcMemHandle_t mine;
    VH_CUDA(cudaIpcGetMemHandle(&mine, h->d_sums_block));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    // every rank also publishes sum_stride: all ranks must have built the same mesh
    struct Card {
        cudaIpcMemHandle_t handle;
        int64_t stride;
    } card{mine, h->sum_stride};
    char *d_send = nullptr, *d_recv = nullptr;
    VH_CUDA(cudaMalloc(&d_send, sizeof(Card)));
    VH_CUDA(cudaMalloc(&d_recv, sizeof(Card) * h->world));
    VH_CUDA(cudaMemcpyAsync(d_send, &card, sizeof(Card), cudaMemcpyHostToDevice, h->s_compute));
    VH_NCCL(g_nccl.AllGather(d_send, d_recv, sizeof(Card), ncclChar, (ncclComm_t)h->nccl_comm, h->s_compute));
    std::vector<Card> all(h->world);
    VH_CUDA(cudaMemcpyAsync(all.data(), d_recv, sizeof(Card) * h->world, cudaMemcpyDeviceToHost, h->s_compute));
    VH_CUDA(cudaStreamSynchronize(h->s_compute));
    cudaFree(d_send);
    cudaFree(d_recv);
    int rc = VH_OK;
    for (int q = 0; q < h->world && rc == VH_OK; ++q) {
        if (all[q].stride != h->sum_stride) {
            vh_set_error("vh_peer_init: rank %d has a different mesh (%lld vs %lld sums)", q, (long long)all[q].stride,
                         (long long)h->sum_stride);
            rc = VH_ERR_ARG;
        } else if (q == h->rank) {
            h->peer_block[q] = h->d_sums_block;
        } else {
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, all[q].handle, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                vh_set_error("vh_peer_init: cannot map rank %d's sums (%s); use vh_nccl_allreduce_sums", q,
                             cudaGetErrorString(e));
                cudaGetLastError();
                rc = VH_ERR_CUDA;
            } else {
                h->peer_block[q] = (double*)p;
            }
        }
    }
    // agree on the outcome (a rank that failed must not leave the others spinning later) and make sure every
    // rank's counters are zeroed before anyone signals
    double ok = rc == VH_OK ? 0.0 : 1.0;
    int rc2 = vh_nccl_allreduce_max(h, &ok);
    if (rc == VH_OK && rc2 != VH_OK) rc = rc2;
    if (rc == VH_OK && ok != 0.0) {
        vh_set_error("vh_peer_init: another rank could not map peer memory; use vh_nccl_allreduce_sums");
        rc = VH_ERR_CUDA;
    }
    if (rc != VH_OK) {
        peer_close(h);
        return rc;
    }
    // The arrival counters in d_sums_block are only zeroed at allocation, so a second vh_peer_init (new communicator,
    // same engine) must not restart the epochs below values the counters already hold: every rank continues from the
    // largest epoch any rank has used.
    double ep = (double)h->peer_epoch;
    VH_TRY(vh_nccl_allreduce_max(h, &ep));
    h->peer_epoch = (uint64_t)(ep + 0.5);
    h->peer_ready = true;
    return VH_OK;
}

int vh_peer_reduce_finalize(vh_handle* h, int64_t n_total, double* tawss, double* osi, double* rrt, double* ecap,
                            double* twssg) {
    NvtxRange range("fused cross-GPU reduction + indices");
    VH_TRY(check_ready(h, "vh_peer_reduce_finalize"));
    VH_CHECK(h->peer_ready, VH_ERR_ARG, "vh_peer_reduce_finalize: call vh_peer_init first");
    VH_CHECK(n_total > 0, VH_ERR_ARG, "vh_peer_reduce_finalize: n_total must be positive");
    VH_TRY(settle_sums(h));
    PeerBlocks pb;
    for (int q = 0; q < VH_MAX_PEERS; ++q) pb.block[q] = h->peer_block[q];
    const int64_t half_off = (int64_t)h->loop_parity * h->sum_stride, flags_off = 2 * h->sum_stride;
    h->peer_epoch += 1;
    VH_TRY(k4_peer_reduce_finalize(h, pb, half_off, flags_off, h->peer_epoch, n_total, h->d_sums_red, h->d_out5_peer));
    h->sums_reduced = true;
    h->count_on_device = true;
    h->peer_unchecked = true;
    double* const outs[5] = {tawss, osi, rrt, ecap, twssg};
    VH_TRY(export_out5(h, outs, h->d_out5_peer, h->s_aux));
    if (tawss || osi || rrt || ecap || twssg) return peer_check(h);  // export_out5 synchronised
    return VH_OK;
}

int vh_nccl_destroy(vh_handle* h) {
    if (h) peer_close(h);
    if (h && h->nccl_comm && g_nccl.CommDestroy) {
        g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
        h->nccl_comm = nullptr;
    }
    return VH_OK;
}

}  // extern "C"
