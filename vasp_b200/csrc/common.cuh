// Shared declarations of libvasp_hemo.so (sm_100a only).  See include/vasp_hemo.h for the ABI and DESIGN.md for
// the data layout.  Reference path: src/vasp/postprocessing/postprocessing_fenics/compute_hemodynamics.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/vasp_hemo.h"

// ---- error plumbing -------------------------------------------------------------------------------------------
void vh_set_error(const char* fmt, ...);

#define VH_CUDA(call)                                                                                     \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            vh_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));          \
            return VH_ERR_CUDA;                                                                           \
        }                                                                                                 \
    } while (0)

#define VH_CHECK(cond, code, ...)                                                                         \
    do {                                                                                                  \
        if (!(cond)) {                                                                                    \
            vh_set_error(__VA_ARGS__);                                                                    \
            return (code);                                                                                \
        }                                                                                                 \
    } while (0)

#define VH_TRY(expr)                                                                                      \
    do {                                                                                                  \
        int rc__ = (expr);                                                                                \
        if (rc__ != VH_OK) return rc__;                                                                   \
    } while (0)

// Number of running sums per facet: 9 (sum tau) + 3 (sum |tau|) + 3 (sum P(|dtau/dt|)); SoA rows of length nF.
constexpr int VH_NSUM = 15;
constexpr int VH_MAX_CONTRIB = 4;  // a tet has at most 4 exterior facets
constexpr int VH_MAX_PEERS = 8;     // GPUs of one NVSwitch node
// Programmatic dependent launch (sm_90+): a kernel launched with vh_launch_pdl may start while its predecessor in the
// stream drains; it must not touch anything the predecessor (or, transitively, an earlier kernel) writes or reads-
// then-overwrites before pdl_wait().  Every thread of every hot kernel executes pdl_wait(), so completion is transitive
// along the stream.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t vh_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                 Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

constexpr unsigned long long VH_PEER_WAIT_NS = 60ull * 1000000000ull;  // longest a fused reduction waits for a peer
constexpr int VH_MROW = 10;        // row length of the multi-facet operator: 9 outputs padded for 16-byte loads

// Per-facet constants in HBM, all SoA over facets (row r of an array with R rows: ptr[r * nF + f]).
struct FacetTables {
    int64_t nF;
    const int32_t* row;    // [ndof][nF]  wall-node number (row triple of the staged block W) of cell dof k
    const double* glam;    // [12][nF]    grad lambda_a (a-major, xyz minor), a in facet-canonical labels
    const double* normal;  // [3][nF]     outward unit normal
    // work list (one warp per entry): single-facet-cell facets first, then facets whose cell owns >= 2 exterior
    // facets, starting at multi_start; -1 entries are padding
    const int32_t* work;
    int64_t n_work, multi_start;
    // multi-facet cells (SurfaceProjector's 4x4 blocks): per multi facet m = work index - multi_start the dense
    // operator tau[3 j + ci] = -mu * sum_q m_mat[m][q][3 j + ci] * u[q],  q = 3 * (cell dof) + component
    const double* m_mat;     // [nMulti][3 * ndof][VH_MROW]
    int64_t nMulti;
};

struct vh_handle {
    int device = 0;
    int sm_count = 148;
    cudaStream_t s_compute = nullptr, s_copy = nullptr, s_aux = nullptr;  // s_aux: the fused cross-GPU reduction
    cudaStream_t s_d2h = nullptr;  // WSS blocks back to the host: its own stream, so that the next H2D does not queue behind it
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    // mesh
    int64_t nv = 0, nc = 0;
    double* d_xyz = nullptr;    // [nv][3]
    int32_t* d_tets = nullptr;  // [nc][4] rows ascending
    int64_t nF = 0, nBV = 0, nW = 0, nMulti = 0;
    int32_t *d_facet_cell = nullptr, *d_facet_verts = nullptr, *d_bcell_parent = nullptr, *d_btopology = nullptr,
            *d_bvert_parent = nullptr;
    int8_t *d_facet_local = nullptr, *d_bcell_local = nullptr;      // bcell_local AoS [nF][3] (map export)
    int8_t* d_blocal_soa = nullptr;                                   // [3][nF]
    double *d_glam = nullptr, *d_normal = nullptr, *d_area = nullptr;  // SoA
    int32_t* d_work = nullptr;
    int64_t n_work = 0, multi_start = 0;
    int8_t* d_m_lf = nullptr;     // [4][nMulti] canonical label a of contributor face (opposite vertex a), -1: none
    double* d_m_w = nullptr;      // [4*9][nMulti] W[a][j][kk] of SurfaceProjector's block solve
    double* d_m_mat = nullptr;    // [nMulti][3 * ndof][VH_MROW], built with the velocity layout (needs the order)

    // velocity layout
    int order = 0, ndof = 0;
    int64_t n_nodes = 0, vec_len = 0, node_stride = 1;
    int64_t comp_offset[3] = {0, 0, 0};
    int32_t* d_facet_nodes = nullptr;  // [ndof][nF] velocity node ids (map export)
    int32_t* d_row = nullptr;          // [ndof][nF] wall-node number of each cell dof, facet-canonical dof order
    // wall-layer nodes = velocity nodes referenced by any wall cell, ascending in vector position
    int64_t nWn = 0, nWn_pad = 0;      // padded to a multiple of 32 (padding repeats the last node)
    int32_t* d_wall_slot = nullptr;    // [nWn_pad] element offset of the node inside a snapshot vector
    std::vector<int32_t> h_wall_slot;  // host copy (nWn_pad entries) for the host-side wall-layer compaction
    // Wall-layer compaction in front of PCIe (compact.cu): a snapshot travels as the dense block C[c][i] =
    // vec[comp_offset[c] + wall_slot[i]] (3 * nWn_pad doubles) instead of the whole vector; K1 is then a pure transpose.
    int compact_mode = 0;              // 0 auto (by nWn_pad / vector slots), 1 never, 2 always
    int host_threads = 0;              // gather threads (0: auto)
    double* h_cstage[3] = {nullptr, nullptr, nullptr};  // pinned ring of gathered pieces
    int64_t cstage_bytes = 0;          // bytes per ring slot
    cudaEvent_t ev_cstage[3] = {nullptr, nullptr, nullptr};  // H2D out of the slot has completed
    bool cstage_busy[3] = {false, false, false};
    void* host_pool = nullptr;         // HostPool* (compact.cu)
    double gather_ms = 0.0;            // host time spent gathering since vh_begin
    int64_t h2d_bytes = 0;             // bytes copied host -> device since vh_begin
    // K1 output: W[((wall node * (w_ld / 32) + column / 32) * 3 + component) * 32 + column % 32], columns = snapshots
    // of the current launch (time tiles of 32, 768 contiguous bytes per node and tile)
    double* d_W = nullptr;
    int64_t w_ld = 0;                  // columns allocated per row (multiple of 32)

    // run state
    double mu = 0.0, dt = 0.0;
    bool begun = false;
    int64_t count = 0;         // snapshots accumulated
    bool sums_pending_zero = false;  // vh_begin happened, nothing accumulated yet: the next K3 overwrites the sums
    bool count_on_device = false;  // after an all-reduce the global count sits behind the sums until someone asks
    bool have_tau_last = false;
    double* d_sums = nullptr;      // [15][nF] + 1 (snapshot count, filled for the all-reduce); one half of d_sums_block
    // Peer-visible block (CUDA IPC over NVLink, vh_peer_*): two halves of sum_stride doubles used by alternate time
    // loops, then VH_MAX_PEERS arrival counters.  Double buffering makes one cross-GPU barrier per reduction enough:
    // a rank can only overwrite a half two loops later, after every peer has signalled the loop in between, which
    // each peer does (stream order) after it finished reading that half.
    double* d_sums_block = nullptr;
    int64_t sum_stride = 0;
    int loop_parity = 0;
    double* d_sums_red = nullptr;  // [15][nF] + 1 globally reduced sums (local), written by the fused peer reduction
    bool sums_reduced = false;     // vh_get_sums reads d_sums_red
    bool peer_ready = false;
    uint64_t peer_epoch = 0;
    // programmatic stream serialization per kernel: bit 0 K1, bit 1 K2, bit 2 K3 (VASP_B200_PDL).  Measured (profiles/
    // r1pdl): K1 + K2 is the best mask (55.8 us per headline step against 64 without); adding K3 costs 30 us on P2.
    int pdl = 3;                   // (the fused peer reduction used to be bit 3; it now runs on its own stream, s_aux)
    bool k2_configured[8] = {false, false, false, false, false, false, false, false};  // cudaFuncSetAttribute done on this handle's device, per (order, launch shape)
    bool peer_unchecked = false;   // a fused reduction was enqueued and its "peer lost" word not looked at yet
    // The fused reduction runs on s_aux behind the K3 of its time loop, so the next loop's K1/K2 overlap the cross-GPU
    // wait; whoever next WRITES the running sums on s_compute first waits for it (ev_join) -- see vh_join_peer.
    bool peer_pending = false;
    double* d_out5_peer = nullptr; // [5][3*nF] global indices written by the fused reduction (K3 keeps writing d_out5)
    double* peer_block[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double* d_tau_last[2] = {nullptr, nullptr};  // [9][nF], ping-pong (read by first chunk, written by last)
    int tau_cur = 0;
    double* d_out5 = nullptr;      // [5][3*nF] TAWSS, OSI, RRT, ECAP, TWSSG
    double* h_out5 = nullptr;      // pinned staging copy of d_out5 for the D2H export
    int64_t out5_count = -1;       // snapshot count d_out5 was evaluated for by the last fold (-1: stale)
    std::vector<cudaEvent_t> batch_events;  // timing events of vh_push_snapshots, reused across calls
    double* d_part = nullptr;      // [groups][15][nF] partial sums of one launch
    int64_t part_cap = 0;          // capacity in groups
    int64_t batch_snapshots = 0, chunk_snapshots = 0;
    int64_t wss_ld = 0, wss_col = 0;  // vh_set_wss_layout: leading dimension (0: one vector per snapshot), next column

    // staging (double buffered)
    double* d_stage[2] = {nullptr, nullptr};
    int64_t stage_cap = 0;  // snapshots per stage buffer
    int64_t stage_row_bytes = 0;  // bytes per snapshot the stage buffers were sized for (whole vector | compact block)
    double* d_wss_stage[2] = {nullptr, nullptr};
    int64_t wss_stage_cap = 0;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr}, ev_wss[2] = {nullptr, nullptr};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    void* d_flush = nullptr;
    int64_t flush_bytes = 0;

    // timers
    double kernel_ms = 0.0, h2d_ms = 0.0;
    int64_t launches = 0;
    // per-launch timing of K1 (k1_stage) and K2 (k2_wall) for the roofline: event triples from a pre-made pool
    // (before K1, between K1 and K2, after K2)
    bool profile = false;
    std::vector<cudaEvent_t> prof_pool;
    size_t prof_used = 0;

    // NCCL (dlopen'ed)
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    double* d_scalar = nullptr;
};

// ---- K0 (k0_precompute.cu) --------------------------------------------------------------------------------------
int k0_build_mesh(vh_handle* h, const double* xyz, int64_t nv, const int64_t* tets, int64_t nc);
int k0_build_velocity_map(vh_handle* h, int order, const double* refined_xyz, int64_t n_nodes, double tol,
                          const int64_t* node_perm, int64_t n_slots);

// ---- K1 (k1_stage.cu) ------------------------------------------------------------------------------------------------
// W[((i * (w_ld / 32) + col / 32) * 3 + c) * 32 + col % 32] = u[col * stride_elems + comp_offset[c] + wall_slot[i]]
// for col < ncol, i < nWn_pad (zero up to the next multiple of 32 columns)
// dense: d_u holds compact blocks (u[col * stride + c * nWn_pad + i]); K1 is then a pure transpose
int k1_launch(vh_handle* h, const double* d_u, int64_t ncol, int64_t stride_elems, bool dense);

// ---- K2/K3/K4 (k2_wall.cu) ---------------------------------------------------------------------------------------
// `n_snap` resident snapshots (d_u + s * stride_elems), staged (K1) and reduced (K2, K3) in column blocks.
// prev_mode: 0 tau_prev=0, 1 tau_prev from h->d_tau_last, 2 recompute from the snapshot just before d_u (halo).
// d_wss (may be null): [n_snap][nF][9] if wss_ld == 0, else the (9 nF) x wss_ld time-major matrix (columns 0..n_snap).
int k2_launch(vh_handle* h, const double* d_u, int64_t n_snap, int64_t stride_elems, int prev_mode, double* d_wss,
              int64_t wss_ld, bool dense = false);
int k4_finalize(vh_handle* h, int64_t n_total, double* d_out5);  // d_out5: 5 arrays [nF*3] TAWSS,OSI,RRT,ECAP,TWSSG
// Fused cross-GPU reduction + final formulas: waits until every rank has signalled `epoch`, adds the partial sums of
// all ranks in rank order straight from their memory (NVLink peer loads), writes the reduced sums and the indices.
struct PeerBlocks {
    const double* block[VH_MAX_PEERS];
};
int k4_peer_reduce_finalize(vh_handle* h, const PeerBlocks& pb, int64_t half_off, int64_t flags_off, uint64_t epoch,
                            int64_t n_total, double* d_red, double* d_out5);
int k_free_run_buffers(vh_handle* h);
// s_compute waits for a fused peer reduction still in flight on s_aux (before anything overwrites the running sums)
void vh_join_peer(vh_handle* h);

FacetTables vh_tables(const vh_handle* h);

// ---- host-side wall-layer compaction (compact.cu) ------------------------------------------------------------------
// out[r * out_stride + c * nWn_pad + i] = rows[r][comp_offset[c] + wall_slot[i]] for r < n, on the handle's thread pool
int compact_gather(vh_handle* h, const double* const* rows, const double* base, int64_t stride_elems, int64_t n,
                   double* out, int64_t out_stride_elems);
void compact_release(vh_handle* h);  // thread pool + pinned ring
int compact_ring_ensure(vh_handle* h, int64_t slot_bytes);
bool compact_wanted(const vh_handle* h);
