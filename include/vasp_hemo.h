/*
 * vasp_hemo.h -- C ABI of libvasp_hemo.so: B200-native wall-shear-stress post-processing.
 *
 * The reference (KVSlab/VaSP) exposes no FFI for this path: its boundary is the Python function
 * compute_hemodyanamics() (src/vasp/postprocessing/postprocessing_fenics/compute_hemodynamics.py:160-372)
 * whose numerics run inside dolfin.  Each entry point below replaces the dolfin object(s) named in its comment;
 * INTEGRATION.md shows the ctypes stub a VaSP maintainer would add in that file.
 *
 * Conventions: plain pointers and sizes only.  All pointers are HOST pointers unless the name starts with d_.
 * Every function returns 0 on success and a negative vh_status on failure; vh_last_error() gives the message
 * of the last failure on the calling thread.  A handle owns all device memory, is bound to one GPU and is not
 * thread-safe; use one handle per GPU (one process per GPU under a launcher).  Host buffers passed to
 * vh_push_snapshots must stay valid until the call returns (the call itself overlaps copies and kernels
 * internally and returns after the last copy has been issued and consumed).
 *
 * Floating point is IEEE fp64 throughout; indices are int64 at the boundary (dolfin's HDF5 layout) and int32 on
 * the device.
 */
#ifndef VASP_HEMO_H
#define VASP_HEMO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vh_handle vh_handle;

enum vh_status {
    VH_OK = 0,
    VH_ERR_CUDA = -1,       /* CUDA runtime error (message has the CUDA string) */
    VH_ERR_ARG = -2,        /* bad argument / call order */
    VH_ERR_MESH = -3,       /* degenerate mesh, unmatched P2 node, ... */
    VH_ERR_NCCL = -4,       /* NCCL missing or failed */
    VH_ERR_NOMEM = -5
};

/* push flags */
#define VH_PUSH_GLOBAL_FIRST 1 /* first snapshot of this push is the first of the whole series: tau_prev = 0
                                  (compute_hemodynamics.py:244,309) */
#define VH_PUSH_HALO_FIRST 2   /* first snapshot of this push only seeds tau_prev (time-shard halo) */

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int vh_create(int device, vh_handle** out);
int vh_destroy(vh_handle* h);
const char* vh_last_error(void);
int vh_device_count(int* n);

/* ---- mesh precompute (K0) -------------------------------------------------------------------------------
 * Replaces Mesh read + BoundaryMesh(mesh,"exterior") (:187-191), the facet->cell maps of InterpolateDG.__init__
 * (:59-61), SurfaceProjector.__init__ (:103-110) and FacetNormal/geometry inside the UFL form (:142-150).
 * xyz: nv*3 doubles, tets: nc*4 int64 (rows need not be sorted). Runs entirely on the device. */
int vh_set_mesh(vh_handle* h, const double* xyz, int64_t nv, const int64_t* tets, int64_t nc);

/* Velocity layout + P1(refined)->P2 map.  Replaces VectorFunctionSpace(refined_mesh,"CG",1) / (mesh,"CG",2) and
 * PETScDMCollection.create_transfer_matrix (:204-206,223,275), which for nested meshes is a permutation.
 *   order = 2: refined_xyz (n_nodes*3) are the refined-mesh vertices the velocity vectors are indexed by; every
 *              P2 node (vertex / edge midpoint) of a wall cell is matched to one of them within tol.
 *   order = 1: refined_xyz may be NULL (velocity lives on the mesh vertices, n_nodes = nv).
 * node_perm (n_nodes int64, may be NULL = identity) maps a velocity node to its slot in the vector; component c
 * of node v is read at  vec[comp_offset[c] + node_stride * node_perm[v]]  (u.h5 written by create_hdf5.py:158-174
 * is comp_offset = {0, n, 2n}, node_stride = 1; dolfin's reordered layout is {0,1,2}, 3).  node_perm may also be an
 * injection into a longer vector: a raw turtleFSI array VisualisationVector/<i> of shape (N_all, 3) is read in
 * place with node_perm = fluid node ids, comp_offset = {0,1,2}, node_stride = 3 (what create_hdf5.py:150-163 slices
 * on the host); a snapshot vector then has comp_offset_max + node_stride * max(node_perm) + 1 entries. */
int vh_set_velocity_layout(vh_handle* h, int order, const double* refined_xyz, int64_t n_nodes, double tol,
                           const int64_t* node_perm, const int64_t comp_offset[3], int64_t node_stride);

/* Sizes of what K0 produced: n[0]=nF exterior facets, n[1]=nBV boundary vertices, n[2]=wall cells,
 * n[3]=facets in cells with >=2 exterior facets, n[4]=dofs per cell (4|10), n[5]=velocity nodes,
 * n[6]=wall-layer velocity nodes (nodes of cells that own an exterior facet; what K1 stages per snapshot). */
int vh_get_sizes(vh_handle* h, int64_t n[7]);

/* Index maps for bit-exact checks against the oracle and for the output writer (any pointer may be NULL):
 *   facet_cell[nF], facet_local[nF] (face opposite local vertex k), facet_verts[nF*3] ascending parent vertex ids,
 *   bcell_parent[nF*3] parent vertex ids in boundary-cell order (BoundaryMesh(..., order=True): ascending in
 *   boundary vertex number),
 *   btopology[nF*3] boundary-vertex numbers, bvert_parent[nBV], bcell_local[nF*3] local cell vertex of each
 *   boundary dof (InterpolateDG's copy map :65-89), facet_nodes[nF*ndof] velocity node of each cell dof. */
int vh_get_maps(vh_handle* h, int32_t* facet_cell, int8_t* facet_local, int32_t* facet_verts,
                int32_t* bcell_parent, int32_t* btopology, int32_t* bvert_parent, int8_t* bcell_local,
                int32_t* facet_nodes);
/* normal[nF*3], area[nF], glam[nF*12] (grad lambda_a of the owning cell) */
int vh_get_geometry(vh_handle* h, double* normal, double* area, double* glam);

/* ---- per-run parameters ---------------------------------------------------------------------------------
 * mu: dynamic viscosity (:142); dt: timestamp gap of the first two selected snapshots (:269).
 * Also zeroes the running sums (a new time loop). */
int vh_begin(vh_handle* h, double mu, double dt);

/* Tuning: max snapshots per host->device batch (0 = auto from free memory), snapshots per time segment of the
 * K2 grid (0 = auto so that the grid fills 148 SMs). */
int vh_set_tuning(vh_handle* h, int64_t batch_snapshots, int64_t chunk_snapshots);

/* Layout of the per-snapshot WSS output of the push calls below.
 *   ld = 0 (default): one dolfin vector per snapshot, [snapshot][facet][boundary dof j][component c] -- what
 *            WSS.h5 stores (write_checkpoint, compute_hemodynamics.py:285-286).
 *   ld > 0: the (dof x time) matrix that VaSP's spectral tools build from WSS.h5 by re-reading every step
 *            (create_transformed_matrix, postprocessing_h5py_common.py:226-246,337-343: row = position in the
 *            WSS vector, column = snapshot): tau of the k-th non-halo snapshot pushed since this call goes to
 *            wss_out[(9 f + 3 j + c) * ld + col0 + k]; pass the SAME base pointer to every push.  K2's lanes run
 *            along time, so these rows are written as full 256-byte lines. */
int vh_set_wss_layout(vh_handle* h, int64_t ld, int64_t col0);

/* ---- snapshot loop (K2/K3) -------------------------------------------------------------------------------
 * Replaces one or more iterations of the loop at :272-318: P1->P2 transfer, Stress.__call__ (assemble + LU solve
 * + InterpolateDG), |tau| / sum(tau) / TWSSG accumulation.  u: n_snap vectors, consecutive ones stride_bytes
 * apart (pinned memory makes the copies asynchronous; see vh_alloc_pinned).  If wss_out != NULL it receives
 * tau of every non-halo snapshot as [snapshot][facet][boundary dof j][component c] doubles (or as columns of the
 * time-major matrix selected with vh_set_wss_layout). */
int vh_push_snapshots(vh_handle* h, const double* u, int64_t n_snap, int64_t stride_bytes, int flags,
                      double* wss_out);
/* Same, for vectors already resident in device memory (wss_out is a device pointer or NULL). */
int vh_push_snapshots_device(vh_handle* h, const double* d_u, int64_t n_snap, int64_t stride_bytes, int flags,
                             double* d_wss_out);

/* ---- wall-layer compaction in front of the bus ----------------------------------------------------------------
 * The reference integrates over ds only (compute_hemodynamics.py:113-115): of a snapshot vector (:274) only the dofs
 * of cells that own an exterior facet reach the result (n[6] of vh_get_sizes: 72 % of the nodes on the tutorial-size
 * mesh, 7 % at 10 M tets).  A COMPACT BLOCK of a snapshot is C[c * nWp + i] = vec[comp_offset[c] + slot[i]], c < 3,
 * i < nWp = n[6] rounded up to a multiple of 32 (the padding repeats the last node); compact_len = 3 * nWp doubles.
 * vh_push_snapshots gathers it on the host (a thread pool inside the library, pinned ring, overlapped with the copies)
 * when that pays: mode 0 = automatic (wall-layer share of the vector below ~1/3), 1 = never, 2 = always.
 * threads = gather threads (0 = hardware threads / ranks on the node, at most 32). */
int vh_set_host_compaction(vh_handle* h, int mode, int threads);
int vh_get_compact_info(vh_handle* h, int64_t* compact_len, int* active);
/* slot[i], i < n[6]: element offset of wall-layer node i inside a snapshot vector (ascending). */
int vh_get_wall_slots(vh_handle* h, int64_t* slots);
/* Host gather only (no device work): out[r * compact_len ...] = compact block of snapshot r.  The rows may live in
 * any host memory, e.g. an mmap of u.h5 -- the page cache is then gathered in place instead of being copied whole
 * (replaces HDF5File.read(u, name), compute_hemodynamics.py:274, for the dofs that matter). */
int vh_compact_snapshots(vh_handle* h, const double* u, int64_t n_snap, int64_t stride_bytes, double* out);
int vh_compact_rows(vh_handle* h, const double* const* rows, int64_t n_snap, double* out);
/* The same gather without a handle (no GPU needed): out[r][c * n_slots + i] = row_r[comp_offset[c] + slots[i]]; rows
 * are given either by address (rows != NULL) or as base + r * stride_bytes. */
int vh_host_gather(const double* const* rows, const double* base, int64_t stride_bytes, int64_t n,
                   const int32_t* slots, int64_t n_slots, const int64_t comp_offset[3], double* out,
                   int64_t out_stride_bytes, int threads);
/* Same contract as vh_push_snapshots / vh_push_snapshots_device for snapshots that already are compact blocks
 * (consecutive ones stride_bytes >= 8 * compact_len apart); K1 is then a pure transpose. */
int vh_push_compact(vh_handle* h, const double* c, int64_t n_snap, int64_t stride_bytes, int flags, double* wss_out);
int vh_push_compact_device(vh_handle* h, const double* d_c, int64_t n_snap, int64_t stride_bytes, int flags,
                           double* d_wss_out);
/* Host milliseconds spent gathering and bytes copied host -> device since vh_begin. */
int vh_get_io_stats(vh_handle* h, double* gather_ms, int64_t* h2d_bytes);

/* ---- reductions / results ------------------------------------------------------------------------------- */
/* Running sums as 15 rows of nF doubles (SoA): rows 0-8 sum tau (row 3*j+c: boundary dof j, component c),
 * rows 9-11 sum |tau| (dof j), rows 12-14 sum P(|dtau/dt|) (dof j); count = snapshots added. */
int vh_get_sums(vh_handle* h, double* sums, int64_t* count);
int vh_set_sums(vh_handle* h, const double* sums, int64_t count);
/* Device pointer to the same 15*nF block (for NCCL or peer access). */
int vh_sums_device_ptr(vh_handle* h, double** d_sums);
/* tau of the last snapshot pushed ([facet][j][c]); needed by nobody but tests and hand-offs between shards. */
int vh_get_tau_last(vh_handle* h, double* tau);

/* Final formulas (:326-346) on the device, each output nF*3 doubles (any may be NULL; with all NULL the call only
 * enqueues the kernel and the results stay in device memory). */
int vh_finalize(vh_handle* h, int64_t n_total, double* tawss, double* osi, double* rrt, double* ecap,
                double* twssg);

int vh_sync(vh_handle* h);
/* Milliseconds spent in kernels / in H2D copies since vh_begin (CUDA events), and number of kernel launches. */
int vh_get_timers(vh_handle* h, double* kernel_ms, double* h2d_ms, int64_t* launches);
/* Per-launch CUDA-event timing of the two hot kernels, for the roofline: switch on, run, then read the summed
 * durations of k1_stage (wall-layer staging) and k2_wall (traction + reductions) and the number of launches of each
 * since the last read (reading synchronises and resets). */
int vh_set_profile(vh_handle* h, int on);
int vh_get_kernel_profile(vh_handle* h, double* k1_ms, double* k2_ms, int64_t* launches);
/* Start/stop markers on the compute stream (CUDA events); stop synchronises and returns the elapsed ms. */
int vh_timer_start(vh_handle* h);
int vh_timer_stop(vh_handle* h, double* ms);

/* ---- memory helpers ------------------------------------------------------------------------------------- */
int vh_alloc_pinned(void** p, int64_t nbytes);
int vh_free_pinned(void* p);
int vh_alloc_device(vh_handle* h, void** d_p, int64_t nbytes);
int vh_free_device(vh_handle* h, void* d_p);
int vh_memcpy_h2d(vh_handle* h, void* d_dst, const void* src, int64_t nbytes);
int vh_memcpy_d2h(vh_handle* h, void* dst, const void* d_src, int64_t nbytes);
/* Writes `nbytes` of device scratch (evicts L2 between timed iterations). */
int vh_flush_l2(vh_handle* h);
int vh_mem_info(vh_handle* h, int64_t* free_bytes, int64_t* total_bytes);

/* ---- multi-GPU (one process per GPU; NCCL is dlopen'ed on first use) -------------------------------------
 * Replaces nothing in the reference (its snapshot loop is sequential, :272-318); SURVEY §8e. */
int vh_nccl_unique_id(char id[128]);
int vh_nccl_init(vh_handle* h, const char id[128], int rank, int world);
/* One ncclAllReduce(sum, fp64) over the 15*nF running sums, plus the snapshot count. */
int vh_nccl_allreduce_sums(vh_handle* h);
/* max-reduce a scalar over ranks (timing) and barrier */
int vh_nccl_allreduce_max(vh_handle* h, double* value);
int vh_nccl_barrier(vh_handle* h);
int vh_nccl_destroy(vh_handle* h);

/* Fused reduction + final formulas over NVLink peer memory (all ranks on one NVSwitch node, <= 8).
 * vh_peer_init (collective; after vh_nccl_init, vh_set_mesh and vh_set_velocity_layout on every rank) maps every
 * rank's running sums into this process with CUDA IPC.  vh_peer_reduce_finalize (collective) then replaces
 * vh_nccl_allreduce_sums + vh_finalize by ONE kernel per rank: it waits for every rank's arrival counter, adds the
 * 15*nF partial sums of all ranks in rank order straight from their memory (bitwise identical on every rank) and
 * evaluates :326-346.  Stream-ordered, no host synchronisation unless host outputs are requested.  If peer memory
 * cannot be mapped vh_peer_init fails on every rank and the NCCL path above remains. */
int vh_peer_init(vh_handle* h);
int vh_peer_reduce_finalize(vh_handle* h, int64_t n_total, double* tawss, double* osi, double* rrt, double* ecap,
                            double* twssg);

#ifdef __cplusplus
}
#endif
#endif /* VASP_HEMO_H */
