// K1: stage the wall-layer nodes of a block of snapshots in time-major order.
//
// The reference reads one velocity vector per loop iteration (compute_hemodynamics.py:274) and multiplies it by the
// P1(refined)->P2 transfer matrix (:275); only the dofs of cells that own an exterior facet ever reach the wall
// traction (:113-115 integrates over ds only).  K1 is that transfer restricted to the wall layer, applied to a
// whole block of snapshots at once and written transposed:
//
//     W[((i * ntile + col / 32) * 3 + c) * 32 + col % 32] = u[col * stride + comp_offset[c] + wall_slot[i]]
//                                                                                  i < nWn_pad, col < 32 * ceil(ncol / 32)
//
// (columns past ncol are zero) so that K2 can give one facet to a warp and the 32 snapshots of a time tile to its
// lanes: every K2 load is then 32 consecutive, 256-byte-aligned doubles (8 full sectors, 2 L1 wavefronts) instead of
// 32 scattered 8-byte gathers, and the three components of a node sit 256 bytes apart (immediate offsets).
//
// Tile = 32 wall nodes x 32 snapshots x 3 components through shared memory.  Reads run along the node list, which
// K0 sorted by position in the vector: as coalesced as the mesh numbering allows, and each wall node is read from
// HBM once per snapshot (not once per facet that touches it).  Writes: 768 contiguous bytes per (node, time tile).
#include "common.cuh"

namespace {

constexpr int TILE = 32;
constexpr int ROWS = 8;  // blockDim.y

__device__ __forceinline__ double ld_stream(const double* p) {
    // read-once data: do not let it displace the staged block in L1/L2
    double v;
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// DENSE: u holds compact blocks (the host gathered the wall layer in front of the bus, compact.cu): node i of
// component c sits at c * nWn_pad + i, so the "gather" is the identity and K1 a pure transpose of 256-byte runs.
template <bool DENSE>
__global__ void __launch_bounds__(TILE* ROWS)
    k1_stage(const double* __restrict__ u, int64_t stride, int ncol, const int32_t* __restrict__ wall_slot,
             int64_t off0, int64_t off1, int64_t off2, double* __restrict__ W, int64_t ntile) {
    __shared__ double tile[3][TILE][TILE + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int64_t node0 = (int64_t)blockIdx.x * TILE;
    const int col0 = blockIdx.y * TILE;
    const int64_t slot = DENSE ? node0 + tx : (int64_t)wall_slot[node0 + tx];
    double v[TILE / ROWS][3];
#pragma unroll
    for (int r = 0; r < TILE / ROWS; ++r) {
        const int col = col0 + ty + ROWS * r;
        if (col < ncol) {
            const double* p = u + (int64_t)col * stride + slot;
            v[r][0] = ld_stream(p + off0);
            v[r][1] = ld_stream(p + off1);
            v[r][2] = ld_stream(p + off2);
        } else {
            v[r][0] = v[r][1] = v[r][2] = 0.0;
        }
    }
    // the loads above touch only the input vectors; W is still being read by the previous launch's K2 until here
    pdl_wait();
    pdl_launch_dependents();
#pragma unroll
    for (int r = 0; r < TILE / ROWS; ++r) {
        tile[0][ty + ROWS * r][tx] = v[r][0];
        tile[1][ty + ROWS * r][tx] = v[r][1];
        tile[2][ty + ROWS * r][tx] = v[r][2];
    }
    __syncthreads();
    // the block is padded to a multiple of 32 columns, so whole 256-byte rows are written
#pragma unroll
    for (int r = 0; r < TILE / ROWS; ++r) {
        const int64_t node = node0 + ty + ROWS * r;
        double* q = W + ((node * ntile + blockIdx.y) * 3) * TILE + tx;
        q[0] = tile[0][tx][ty + ROWS * r];
        q[TILE] = tile[1][tx][ty + ROWS * r];
        q[2 * TILE] = tile[2][tx][ty + ROWS * r];
    }
}

}  // namespace

int k1_launch(vh_handle* h, const double* d_u, int64_t ncol, int64_t stride_elems, bool dense) {
    VH_CHECK(ncol > 0 && ncol <= h->w_ld, VH_ERR_ARG, "k1_launch: %lld columns do not fit the staged block (%lld)",
             (long long)ncol, (long long)h->w_ld);
    const int64_t gy = (ncol + 2 * TILE - 1) / (2 * TILE) * 2;  // zero-filled up to whole 64-column passes of K2
    VH_CHECK(gy <= 65535, VH_ERR_ARG, "k1_launch: too many snapshots in one block");
    dim3 grid((unsigned)(h->nWn_pad / TILE), (unsigned)gy), block(TILE, ROWS);
    const bool pdl = (h->pdl & 1) && !h->profile;
    if (dense)
        VH_CUDA(vh_launch_pdl(k1_stage<true>, grid, block, 0, h->s_compute, pdl, d_u, stride_elems, (int)ncol,
                              (const int32_t*)nullptr, (int64_t)0, h->nWn_pad, 2 * h->nWn_pad, h->d_W, h->w_ld / TILE));
    else
        VH_CUDA(vh_launch_pdl(k1_stage<false>, grid, block, 0, h->s_compute, pdl, d_u, stride_elems, (int)ncol,
                              (const int32_t*)h->d_wall_slot, h->comp_offset[0], h->comp_offset[1], h->comp_offset[2],
                              h->d_W, h->w_ld / TILE));
    h->launches += 1;
    return VH_OK;
}
