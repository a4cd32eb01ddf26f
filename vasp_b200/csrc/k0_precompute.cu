// K0: once-per-mesh precompute, entirely on the device (integer kernels + one geometry kernel).
//
// Replaces, in the reference's compute_hemodynamics.py: BoundaryMesh(mesh,"exterior") (:191), the facet->cell
// connectivity and entity_map of InterpolateDG.__init__ (:59-61), its per-snapshot coordinate matching (:68-83,
// constant, so done once here), SurfaceProjector.__init__'s mass matrix (:103-110, only the 4x4 blocks of cells
// owning >= 2 exterior facets survive), FacetNormal / grad of the UFL form (:142-150) and
// PETScDMCollection.create_transfer_matrix (:223), which for nested meshes is a node permutation.
//
// No sort is needed: faces are chained per smallest vertex with atomicExch, duplicates found by walking the short
// chains, and dolfin's facet order (lexicographic in the sorted vertex triple) is recovered as
// scan(exterior faces per smallest vertex) + rank inside the chain.
#include <math.h>
#include <stdio.h>

#include "common.cuh"

namespace {

constexpr int TPB = 256;
inline int nblk(int64_t n, int tpb = TPB) { return (int)((n + tpb - 1) / tpb); }

// ---------------------------------------------------------------------------------------------------------------
// exclusive scan (int32), recursive block scan; n up to 2^31
// ---------------------------------------------------------------------------------------------------------------
constexpr int SCAN_TPB = 256, SCAN_IPT = 4, SCAN_TILE = SCAN_TPB * SCAN_IPT;

__global__ void scan_tile(const int32_t* __restrict__ in, int32_t* __restrict__ out, int32_t* __restrict__ tile_sum,
                          int64_t n) {
    __shared__ int32_t warp_tot[SCAN_TPB / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_IPT;
    int32_t v[SCAN_IPT], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int32_t woff = 0;
    for (int w = 0; w < wid; ++w) woff += warp_tot[w];
    int32_t run = woff + incl - s;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == SCAN_TPB - 1) tile_sum[blockIdx.x] = woff + incl;
}

__global__ void scan_add(int32_t* __restrict__ out, const int32_t* __restrict__ tile_off, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x;
    int32_t off = tile_off[blockIdx.x];
    for (int k = 0; k < SCAN_IPT; ++k, i += SCAN_TPB)
        if (i < n) out[i] += off;
}

// out[i] = sum_{j<i} in[j]; *total (device) = sum of all.  in/out may alias.
int exclusive_scan(const int32_t* d_in, int32_t* d_out, int64_t n, int32_t* d_total, cudaStream_t st) {
    if (n <= 0) {
        VH_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int32_t), st));
        return VH_OK;
    }
    int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int32_t* d_tiles = nullptr;
    VH_CUDA(cudaMalloc(&d_tiles, sizeof(int32_t) * (tiles + 1)));
    scan_tile<<<(int)tiles, SCAN_TPB, 0, st>>>(d_in, d_out, d_tiles, n);
    if (tiles == 1) {
        VH_CUDA(cudaMemcpyAsync(d_total, d_tiles, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    } else {
        int rc = exclusive_scan(d_tiles, d_tiles, tiles, d_total, st);
        if (rc != VH_OK) {
            cudaFree(d_tiles);
            return rc;
        }
        scan_add<<<(int)tiles, SCAN_TPB, 0, st>>>(d_out, d_tiles, n);
    }
    VH_CUDA(cudaGetLastError());
    VH_CUDA(cudaStreamSynchronize(st));
    VH_CUDA(cudaFree(d_tiles));
    return VH_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// topology
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sort4(int32_t& a, int32_t& b, int32_t& c, int32_t& d) {
#define VH_CSWAP(x, y)       \
    if (x > y) {             \
        int32_t t__ = x;     \
        x = y;               \
        y = t__;             \
    }
    VH_CSWAP(a, b) VH_CSWAP(c, d) VH_CSWAP(a, c) VH_CSWAP(b, d) VH_CSWAP(b, c)
#undef VH_CSWAP
}

// dolfin orders cell vertices ascending on read (Mesh.order()); also narrows int64 -> int32 and validates.
__global__ void k0_order_cells(const int64_t* __restrict__ t64, int32_t* __restrict__ t32, int64_t nc, int64_t nv,
                               int32_t* __restrict__ bad) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    int64_t r0 = t64[4 * c], r1 = t64[4 * c + 1], r2 = t64[4 * c + 2], r3 = t64[4 * c + 3];
    if (r0 < 0 || r1 < 0 || r2 < 0 || r3 < 0 || r0 >= nv || r1 >= nv || r2 >= nv || r3 >= nv) {
        atomicAdd(bad, 1);
        r0 = r1 = r2 = r3 = 0;
    }
    int32_t a = (int32_t)r0, b = (int32_t)r1, cc = (int32_t)r2, d = (int32_t)r3;
    sort4(a, b, cc, d);
    if (a == b || b == cc || cc == d) atomicAdd(bad, 1);
    reinterpret_cast<int4*>(t32)[c] = make_int4(a, b, cc, d);
}

// face id = 4*cell + k, k = local vertex the face is opposite to; vertices ascending
__device__ __forceinline__ void face_verts(const int32_t* __restrict__ tets, int64_t fid, int32_t& a, int32_t& b,
                                           int32_t& c) {
    int4 t = reinterpret_cast<const int4*>(tets)[fid >> 2];
    int k = (int)(fid & 3);
    a = (k == 0) ? t.y : t.x;
    b = (k <= 1) ? t.z : t.y;
    c = (k <= 2) ? t.w : t.z;
}

__global__ void k0_link_faces(const int32_t* __restrict__ tets, int64_t nfaces, int32_t* __restrict__ head,
                              int32_t* __restrict__ next) {
    int64_t fid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (fid >= nfaces) return;
    int32_t a, b, c;
    face_verts(tets, fid, a, b, c);
    next[fid] = atomicExch(&head[a], (int32_t)fid);
}

// exterior <=> the (sorted) vertex triple occurs once.  Also counts exterior faces per smallest vertex.
__global__ void k0_mark_exterior(const int32_t* __restrict__ tets, int64_t nfaces, const int32_t* __restrict__ head,
                                 const int32_t* __restrict__ next, int8_t* __restrict__ ext,
                                 int32_t* __restrict__ vcount) {
    int64_t fid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (fid >= nfaces) return;
    int32_t a, b, c;
    face_verts(tets, fid, a, b, c);
    int hits = 0;
    for (int32_t g = head[a]; g >= 0; g = next[g]) {
        int32_t ga, gb, gc;
        face_verts(tets, g, ga, gb, gc);
        hits += (gb == b && gc == c);
    }
    int8_t e = (hits == 1);
    ext[fid] = e;
    if (e) atomicAdd(&vcount[a], 1);
}

// facet number = (#exterior faces with smaller first vertex) + rank of (b,c) among the chain's exterior faces
__global__ void k0_rank_exterior(const int32_t* __restrict__ tets, int64_t nfaces, const int32_t* __restrict__ head,
                                 const int32_t* __restrict__ next, const int8_t* __restrict__ ext,
                                 const int32_t* __restrict__ voff, int32_t* __restrict__ face_to_facet,
                                 int32_t* __restrict__ facet_cell, int8_t* __restrict__ facet_local,
                                 int32_t* __restrict__ facet_verts) {
    int64_t fid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (fid >= nfaces) return;
    if (!ext[fid]) {
        face_to_facet[fid] = -1;
        return;
    }
    int32_t a, b, c;
    face_verts(tets, fid, a, b, c);
    int rank = 0;
    for (int32_t g = head[a]; g >= 0; g = next[g]) {
        if (!ext[g]) continue;
        int32_t ga, gb, gc;
        face_verts(tets, g, ga, gb, gc);
        rank += (gb < b) || (gb == b && gc < c);
    }
    int32_t i = voff[a] + rank;
    face_to_facet[fid] = i;
    facet_cell[i] = (int32_t)(fid >> 2);
    facet_local[i] = (int8_t)(fid & 3);
    facet_verts[3 * i] = a;
    facet_verts[3 * i + 1] = b;
    facet_verts[3 * i + 2] = c;
}

// ---------------------------------------------------------------------------------------------------------------
// geometry + boundary mesh (dolfin BoundaryComputation's first-encounter vertex numbering, then Mesh.order())
// ---------------------------------------------------------------------------------------------------------------
__global__ void k0_geometry(const double* __restrict__ xyz, const int32_t* __restrict__ tets, int64_t nF,
                            const int32_t* __restrict__ facet_cell, const int8_t* __restrict__ facet_local,
                            const int32_t* __restrict__ facet_verts, int32_t* __restrict__ bcell_parent,
                            int8_t* __restrict__ bcell_local, int8_t* __restrict__ blocal_soa,
                            double* __restrict__ glam, double* __restrict__ normal, double* __restrict__ area,
                            int32_t* __restrict__ first_seen, int32_t* __restrict__ bad) {
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    int4 t = reinterpret_cast<const int4*>(tets)[facet_cell[f]];
    int32_t tv[4] = {t.x, t.y, t.z, t.w};
    double p[4][3];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int d = 0; d < 3; ++d) p[a][d] = xyz[3 * (int64_t)tv[a] + d];
    // grad lambda_a: rows of J^-1 for a = 1..3 (J columns e_a = p_a - p_0), grad lambda_0 = -sum
    double e[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int d = 0; d < 3; ++d) e[a][d] = p[a + 1][d] - p[0][d];
    double c12[3] = {e[1][1] * e[2][2] - e[1][2] * e[2][1], e[1][2] * e[2][0] - e[1][0] * e[2][2],
                     e[1][0] * e[2][1] - e[1][1] * e[2][0]};
    double c20[3] = {e[2][1] * e[0][2] - e[2][2] * e[0][1], e[2][2] * e[0][0] - e[2][0] * e[0][2],
                     e[2][0] * e[0][1] - e[2][1] * e[0][0]};
    double c01[3] = {e[0][1] * e[1][2] - e[0][2] * e[1][1], e[0][2] * e[1][0] - e[0][0] * e[1][2],
                     e[0][0] * e[1][1] - e[0][1] * e[1][0]};
    double det = e[0][0] * c12[0] + e[0][1] * c12[1] + e[0][2] * c12[2];
    if (!(fabs(det) > 0.0)) atomicAdd(bad, 1);
    double inv = 1.0 / det;
    double g[4][3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        g[1][d] = c12[d] * inv;
        g[2][d] = c20[d] * inv;
        g[3][d] = c01[d] * inv;
        g[0][d] = -(g[1][d] + g[2][d] + g[3][d]);
    }
    int k = facet_local[f];
    double gk[3] = {g[k][0], g[k][1], g[k][2]};
    double gn = sqrt(gk[0] * gk[0] + gk[1] * gk[1] + gk[2] * gk[2]);
#pragma unroll
    for (int d = 0; d < 3; ++d) normal[(int64_t)d * nF + f] = -gk[d] / gn;
    // facet vertices ascending = local vertices != k
    int lv[3] = {k == 0 ? 1 : 0, k <= 1 ? 2 : 1, k <= 2 ? 3 : 2};
    double a1[3], a2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        a1[d] = p[lv[1]][d] - p[lv[0]][d];
        a2[d] = p[lv[2]][d] - p[lv[0]][d];
    }
    double nx = a1[1] * a2[2] - a1[2] * a2[1], ny = a1[2] * a2[0] - a1[0] * a2[2], nz = a1[0] * a2[1] - a1[1] * a2[0];
    area[f] = 0.5 * sqrt(nx * nx + ny * ny + nz * nz);
    // the boundary-cell vertex order is settled later (k0_order_bcells): BoundaryMesh(..., order=True) sorts it
    // facet-canonical labels for the hot loop: 0,1,2 = boundary-cell vertices, 3 = opposite vertex
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        glam[(int64_t)(0 + d) * nF + f] = g[lv[0]][d];
        glam[(int64_t)(3 + d) * nF + f] = g[lv[1]][d];
        glam[(int64_t)(6 + d) * nF + f] = g[lv[2]][d];
        glam[(int64_t)(9 + d) * nF + f] = g[k][d];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        bcell_parent[3 * f + j] = tv[lv[j]];
        bcell_local[3 * f + j] = (int8_t)lv[j];
        blocal_soa[(int64_t)j * nF + f] = (int8_t)lv[j];
        // first encounter over facets in facet order, vertices in facet (ascending) order
        atomicMin(&first_seen[facet_verts[3 * f + j]], (int32_t)(3 * f + j));
    }
}

__global__ void k0_flag_first(const int32_t* __restrict__ facet_verts, const int32_t* __restrict__ first_seen,
                              int64_t n3, int32_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    flag[i] = (first_seen[facet_verts[i]] == (int32_t)i);
}

__global__ void k0_number_bverts(const int32_t* __restrict__ facet_verts, const int32_t* __restrict__ first_seen,
                                 const int32_t* __restrict__ pos, int64_t n3, int32_t* __restrict__ bvert_parent,
                                 int32_t* __restrict__ vnumber) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    int32_t v = facet_verts[i];
    if (first_seen[v] == (int32_t)i) {
        bvert_parent[pos[i]] = v;
        vnumber[v] = pos[i];
    }
}

// BoundaryMesh(mesh, "exterior") is built with dolfin's default order = True: after BoundaryComputation has numbered
// the boundary vertices (first encounter), Mesh.order() sorts every boundary cell's vertices ascending in BOUNDARY
// vertex number.  k0_geometry left the three facet vertices ascending in parent id; this kernel puts them (and
// everything indexed by boundary dof: bcell_parent, bcell_local, the first three grad-lambda rows) in that order.
__global__ void k0_order_bcells(const int32_t* __restrict__ vnumber, int64_t nF, int32_t* __restrict__ bcell_parent,
                                int8_t* __restrict__ bcell_local, int8_t* __restrict__ blocal_soa,
                                double* __restrict__ glam, int32_t* __restrict__ btopology) {
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    int32_t pv[3], bn[3];
    int8_t lv[3];
    double g[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        pv[j] = bcell_parent[3 * f + j];
        bn[j] = vnumber[pv[j]];
        lv[j] = bcell_local[3 * f + j];
#pragma unroll
        for (int d = 0; d < 3; ++d) g[j][d] = glam[(int64_t)(3 * j + d) * nF + f];
    }
    int o[3] = {0, 1, 2};
#define VH_OSWAP(x, y)                 \
    if (bn[o[x]] > bn[o[y]]) {         \
        int t__ = o[x];                \
        o[x] = o[y];                   \
        o[y] = t__;                    \
    }
    VH_OSWAP(0, 1) VH_OSWAP(1, 2) VH_OSWAP(0, 1)
#undef VH_OSWAP
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int s = o[j];
        bcell_parent[3 * f + j] = pv[s];
        btopology[3 * f + j] = bn[s];
        bcell_local[3 * f + j] = lv[s];
        blocal_soa[(int64_t)j * nF + f] = lv[s];
#pragma unroll
        for (int d = 0; d < 3; ++d) glam[(int64_t)(3 * j + d) * nF + f] = g[s][d];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// work list + multi-facet cells
// ---------------------------------------------------------------------------------------------------------------
__global__ void k0_flag_multi(const int32_t* __restrict__ facet_cell, const int8_t* __restrict__ ext, int64_t nF,
                              int32_t* __restrict__ is_multi, int32_t* __restrict__ is_single) {
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    int64_t c = facet_cell[f];
    int m = ext[4 * c] + ext[4 * c + 1] + ext[4 * c + 2] + ext[4 * c + 3];
    is_multi[f] = (m >= 2);
    is_single[f] = (m < 2);
}

__global__ void k0_fill_work(const int32_t* __restrict__ is_multi, const int32_t* __restrict__ pos_single,
                             const int32_t* __restrict__ pos_multi, int64_t nF, int64_t multi_start,
                             int32_t* __restrict__ work) {
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    if (is_multi[f])
        work[multi_start + pos_multi[f]] = (int32_t)f;
    else
        work[pos_single[f]] = (int32_t)f;
}

__global__ void k0_count_wall_cells(const int32_t* __restrict__ facet_cell, const int8_t* __restrict__ facet_local,
                                    const int8_t* __restrict__ ext, int64_t nF, int32_t* __restrict__ count) {
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nF) return;
    int64_t c = facet_cell[f];
    int k = facet_local[f];
    bool first = true;
    for (int j = 0; j < k; ++j) first = first && !ext[4 * c + j];
    if (first) atomicAdd(count, 1);
}

// 4x4 SPD inverse by Gauss-Jordan (A = sum of facet mass matrices of a cell with >= 2 exterior facets)
__device__ void invert4(double a[4][4], double inv[4][4]) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) inv[i][j] = (i == j) ? 1.0 : 0.0;
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        double best = fabs(a[col][col]);
        for (int r = col + 1; r < 4; ++r)
            if (fabs(a[r][col]) > best) {
                best = fabs(a[r][col]);
                piv = r;
            }
        if (piv != col)
            for (int j = 0; j < 4; ++j) {
                double t = a[col][j];
                a[col][j] = a[piv][j];
                a[piv][j] = t;
                t = inv[col][j];
                inv[col][j] = inv[piv][j];
                inv[piv][j] = t;
            }
        double d = 1.0 / a[col][col];
        for (int j = 0; j < 4; ++j) {
            a[col][j] *= d;
            inv[col][j] *= d;
        }
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            double fct = a[r][col];
            for (int j = 0; j < 4; ++j) {
                a[r][j] -= fct * a[col][j];
                inv[r][j] -= fct * inv[col][j];
            }
        }
    }
}

// Weights of SurfaceProjector's block solve for facet f = work[multi_start + m], in f's canonical labels
// (0,1,2 = boundary dofs, 3 = opposite vertex).  Contributor a = the cell's face opposite canonical vertex a
// (a = 3 is f itself), its vertices taken in ascending canonical label:
//   tau_f(j) = sum_a sum_kk W[a][j][kk] Ft_a(vertex kk of face a),
//   W[a][j][kk] = sum_a' Ainv[j][a'] M_a[a'][v_kk],  M_a = area_a/12 (1 + delta) on face a's vertices,
//   A = sum_a M_a  (compute_hemodynamics.py:103-117; zero rows cannot occur when >= 2 facets are exterior).
__global__ void k0_multi_weights(const int32_t* __restrict__ work, int64_t multi_start, int64_t nMulti,
                                 const int32_t* __restrict__ facet_cell, const int8_t* __restrict__ facet_local,
                                 const int32_t* __restrict__ face_to_facet, const int8_t* __restrict__ bcell_local,
                                 const double* __restrict__ area, int8_t* __restrict__ m_lf,
                                 double* __restrict__ m_w) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nMulti) return;
    int32_t f = work[multi_start + m];
    int64_t c = facet_cell[f];
    int nl[4] = {bcell_local[3 * f], bcell_local[3 * f + 1], bcell_local[3 * f + 2], facet_local[f]};
    double A[4][4] = {{0}}, Ainv[4][4], s[4];
    for (int a = 0; a < 4; ++a) {  // canonical contributor a
        int32_t gf = face_to_facet[4 * c + nl[a]];
        s[a] = gf >= 0 ? area[gf] / 12.0 : 0.0;
        for (int p = 0; p < 4; ++p)
            for (int q = 0; q < 4; ++q)
                if (p != a && q != a) A[p][q] += s[a] * (p == q ? 2.0 : 1.0);
    }
    invert4(A, Ainv);
    for (int a = 0; a < 4; ++a) {
        m_lf[(int64_t)a * nMulti + m] = (int8_t)(s[a] > 0.0 ? a : -1);
        int vc[3] = {a == 0 ? 1 : 0, a <= 1 ? 2 : 1, a <= 2 ? 3 : 2};
        for (int j = 0; j < 3; ++j)
            for (int kk = 0; kk < 3; ++kk) {
                double w = 0.0;
                for (int p = 0; p < 4; ++p)
                    if (p != a) w += Ainv[j][p] * s[a] * (p == vc[kk] ? 2.0 : 1.0);
                m_w[(int64_t)(9 * a + 3 * j + kk) * nMulti + m] = w;
            }
    }
}

// Dense operator of a multi-facet-cell facet (thread = (multi facet m, cell dof k), canonical labels):
//   tau_f(j) = sum_a sum_kk W[a][j][kk] Ft_a(v_kk),   Ft_a(V) = -mu (I - n_a n_a^T) (G(V) + G(V)^T) n_a,
//   G(V) = sum_k u_k (x) d_k(V),  d_k(V) = grad phi_k at vertex V:
//     P1: g_k;  P2 vertex dof b: 3 g_V if b == V else -g_b;  P2 edge dof (p,q): 4 g_q if p == V, 4 g_p if q == V, else 0
//   => (G + G^T) n = sum_k [(d_k . n) I + d_k n^T] u_k.
// m_mat[m][3 k + c][3 j + ci] collects the coefficient of u_k[c] in tau_f(j)[ci] without the factor -mu.
__global__ void k0_multi_matrix(int64_t nMulti, int ndof, const int32_t* __restrict__ work, int64_t multi_start,
                                int64_t nF, const double* __restrict__ glam, const int8_t* __restrict__ m_lf,
                                const double* __restrict__ m_w, double* __restrict__ m_mat) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nMulti * ndof) return;
    const int64_t m = idx / ndof;
    const int k = (int)(idx % ndof);
    const int32_t f = work[multi_start + m];
    double g[4][3];
    for (int b = 0; b < 4; ++b)
        for (int d = 0; d < 3; ++d) g[b][d] = glam[(int64_t)(3 * b + d) * nF + f];
    double M[3][9];  // [c][3 j + ci]
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 9; ++i) M[c][i] = 0.0;
    const int ea[6] = {0, 0, 1, 0, 1, 2}, eb[6] = {1, 2, 2, 3, 3, 3};  // canonical edge dofs 4..9
    for (int a = 0; a < 4; ++a) {
        if (m_lf[(int64_t)a * nMulti + m] < 0) continue;
        double inv = -1.0 / sqrt(g[a][0] * g[a][0] + g[a][1] * g[a][1] + g[a][2] * g[a][2]);
        double n[3] = {g[a][0] * inv, g[a][1] * inv, g[a][2] * inv};
        const int vc[3] = {a == 0 ? 1 : 0, a <= 1 ? 2 : 1, a <= 2 ? 3 : 2};
        for (int kk = 0; kk < 3; ++kk) {
            const int V = vc[kk];
            double d[3] = {0.0, 0.0, 0.0};
            if (ndof == 4) {
                for (int t = 0; t < 3; ++t) d[t] = g[k][t];
            } else if (k < 4) {
                for (int t = 0; t < 3; ++t) d[t] = (k == V) ? 3.0 * g[V][t] : -g[k][t];
            } else {
                const int p = ea[k - 4], q = eb[k - 4];
                if (p == V)
                    for (int t = 0; t < 3; ++t) d[t] = 4.0 * g[q][t];
                else if (q == V)
                    for (int t = 0; t < 3; ++t) d[t] = 4.0 * g[p][t];
                else
                    continue;
            }
            const double dn = d[0] * n[0] + d[1] * n[1] + d[2] * n[2];
            // B = (I - n n^T) [(d.n) I + d n^T]
            double B[3][3];
            for (int ci = 0; ci < 3; ++ci)
                for (int c = 0; c < 3; ++c) {
                    double acc = 0.0;
                    for (int t = 0; t < 3; ++t) {
                        const double P = (ci == t ? 1.0 : 0.0) - n[ci] * n[t];
                        const double A = (t == c ? dn : 0.0) + d[t] * n[c];
                        acc += P * A;
                    }
                    B[ci][c] = acc;
                }
            for (int j = 0; j < 3; ++j) {
                const double wj = m_w[(int64_t)(9 * a + 3 * j + kk) * nMulti + m];
                for (int ci = 0; ci < 3; ++ci)
                    for (int c = 0; c < 3; ++c) M[c][3 * j + ci] += wj * B[ci][c];
            }
        }
    }
    for (int c = 0; c < 3; ++c) {
        double* out = m_mat + ((m * ndof + k) * 3 + c) * VH_MROW;
        for (int i = 0; i < 9; ++i) out[i] = M[c][i];
        out[9] = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// P2 node map: uniform grid of the refined-mesh vertices, chained per grid cell
// ---------------------------------------------------------------------------------------------------------------
struct Grid {
    double lo[3], inv_h[3];
    int n[3];
};

__device__ __forceinline__ int grid_coord(const Grid& g, int d, double x) {
    int i = (int)floor((x - g.lo[d]) * g.inv_h[d]);
    return min(max(i, 0), g.n[d] - 1);
}

__global__ void k0_grid_insert(const double* __restrict__ pts, int64_t n, Grid g, int32_t* __restrict__ head,
                               int32_t* __restrict__ next) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = grid_coord(g, 0, pts[3 * i]), cy = grid_coord(g, 1, pts[3 * i + 1]), cz = grid_coord(g, 2, pts[3 * i + 2]);
    int64_t cell = ((int64_t)cz * g.n[1] + cy) * g.n[0] + cx;
    next[i] = atomicExch(&head[cell], (int32_t)i);
}

// UFC P2 local dof order: 4 vertices, then edges (2,3),(1,3),(1,2),(0,3),(0,2),(0,1)
__constant__ int c_edge_a[6] = {2, 1, 1, 0, 0, 0};
__constant__ int c_edge_b[6] = {3, 3, 2, 3, 2, 1};

__global__ void k0_match_p2(const double* __restrict__ xyz, const int32_t* __restrict__ tets, int64_t nF,
                            const int32_t* __restrict__ facet_cell, const double* __restrict__ pts, Grid g,
                            const int32_t* __restrict__ head, const int32_t* __restrict__ next, double tol,
                            int32_t* __restrict__ facet_nodes, int32_t* __restrict__ unmatched) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nF * 10) return;
    int64_t f = idx % nF;
    int k = (int)(idx / nF);
    int4 t = reinterpret_cast<const int4*>(tets)[facet_cell[f]];
    int32_t tv[4] = {t.x, t.y, t.z, t.w};
    double q[3];
    if (k < 4) {
        for (int d = 0; d < 3; ++d) q[d] = xyz[3 * (int64_t)tv[k] + d];
    } else {
        int a = tv[c_edge_a[k - 4]], b = tv[c_edge_b[k - 4]];
        for (int d = 0; d < 3; ++d) q[d] = 0.5 * (xyz[3 * (int64_t)a + d] + xyz[3 * (int64_t)b + d]);
    }
    int lo[3], hi[3];
    for (int d = 0; d < 3; ++d) {
        lo[d] = grid_coord(g, d, q[d] - tol);
        hi[d] = grid_coord(g, d, q[d] + tol);
    }
    double best = tol * tol;
    int32_t arg = -1;
    for (int cz = lo[2]; cz <= hi[2]; ++cz)
        for (int cy = lo[1]; cy <= hi[1]; ++cy)
            for (int cx = lo[0]; cx <= hi[0]; ++cx) {
                int64_t cell = ((int64_t)cz * g.n[1] + cy) * g.n[0] + cx;
                for (int32_t i = head[cell]; i >= 0; i = next[i]) {
                    double dx = pts[3 * (int64_t)i] - q[0], dy = pts[3 * (int64_t)i + 1] - q[1],
                           dz = pts[3 * (int64_t)i + 2] - q[2];
                    double d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 < best || (d2 == best && (arg < 0 || i < arg))) {
                        best = d2;
                        arg = i;
                    }
                }
            }
    if (arg < 0) atomicAdd(unmatched, 1);
    facet_nodes[idx] = arg;
}

__global__ void k0_p1_nodes(const int32_t* __restrict__ tets, int64_t nF, const int32_t* __restrict__ facet_cell,
                            int32_t* __restrict__ facet_nodes) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nF * 4) return;
    int64_t f = idx % nF;
    int k = (int)(idx / nF);
    facet_nodes[idx] = tets[4 * (int64_t)facet_cell[f] + k];
}

// Position (permuted velocity node) of facet-canonical cell dof kc: vertices (b0,b1,b2,opp), then edges
// e01,e02,e12,e03,e13,e23.  facet_nodes is in UFC cell order (4 vertices, edges (2,3),(1,3),(1,2),(0,3),(0,2),(0,1)).
__device__ __forceinline__ int64_t canonical_dof_node(const int32_t* __restrict__ facet_nodes,
                                                      const int8_t* __restrict__ bcell_local,
                                                      const int8_t* __restrict__ facet_local,
                                                      const int64_t* __restrict__ perm, int64_t nF, int64_t f, int kc) {
    int nl[4] = {bcell_local[3 * f], bcell_local[3 * f + 1], bcell_local[3 * f + 2], facet_local[f]};
    int old;
    if (kc < 4) {
        old = nl[kc];
    } else {
        const int ea[6] = {0, 0, 1, 0, 1, 2}, eb[6] = {1, 2, 2, 3, 3, 3};
        int oa = nl[ea[kc - 4]], ob = nl[eb[kc - 4]];
        int lo = oa < ob ? oa : ob, hi = oa < ob ? ob : oa;
        old = 9 - (lo == 0 ? hi - 1 : lo == 1 ? hi + 1 : 5);
    }
    int64_t v = facet_nodes[(int64_t)old * nF + f];
    return perm ? perm[v] : v;
}

// wall-layer nodes: every velocity node some wall cell touches
__global__ void k0_mark_wall_nodes(const int32_t* __restrict__ facet_nodes, const int8_t* __restrict__ bcell_local,
                                   const int8_t* __restrict__ facet_local, const int64_t* __restrict__ perm, int64_t nF,
                                   int ndof, int64_t n_nodes, int32_t* __restrict__ flag, int32_t* __restrict__ bad) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nF * ndof) return;
    int64_t p = canonical_dof_node(facet_nodes, bcell_local, facet_local, perm, nF, idx % nF, (int)(idx / nF));
    if (p < 0 || p >= n_nodes) {
        atomicAdd(bad, 1);
        return;
    }
    flag[p] = 1;
}

// wall_slot[i] = offset inside a snapshot vector of the i-th wall node (ascending => K1 reads are as coalesced as the
// numbering allows); entries [nWn, nWn_pad) repeat the last node so that K1 needs no bounds checks
__global__ void k0_wall_slots(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int64_t n_nodes,
                              int64_t node_stride, int64_t nWn, int64_t nWn_pad, int32_t* __restrict__ wall_slot) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_nodes || !flag[p]) return;
    int32_t s = (int32_t)(p * node_stride);
    wall_slot[pos[p]] = s;
    if (pos[p] == nWn - 1)
        for (int64_t i = nWn; i < nWn_pad; ++i) wall_slot[i] = s;
}

__global__ void k0_rows(const int32_t* __restrict__ facet_nodes, const int8_t* __restrict__ bcell_local,
                        const int8_t* __restrict__ facet_local, const int64_t* __restrict__ perm,
                        const int32_t* __restrict__ pos, int64_t nF, int ndof, int32_t* __restrict__ row) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nF * ndof) return;
    row[idx] = pos[canonical_dof_node(facet_nodes, bcell_local, facet_local, perm, nF, idx % nF, (int)(idx / nF))];
}

template <typename T>
int dev_alloc(T** p, int64_t n) {
    VH_CUDA(cudaMalloc((void**)p, sizeof(T) * (size_t)(n > 0 ? n : 1)));
    return VH_OK;
}

template <typename T>
void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

}  // namespace

// =================================================================================================================
int k0_build_mesh(vh_handle* h, const double* xyz, int64_t nv, const int64_t* tets, int64_t nc) {
    VH_CHECK(nv > 0 && nc > 0 && xyz && tets, VH_ERR_ARG, "vh_set_mesh: empty mesh");
    VH_CHECK(nv < (1LL << 31) && 4 * nc < (1LL << 31), VH_ERR_ARG, "vh_set_mesh: mesh too large for int32 indices");
    cudaStream_t st = h->s_compute;
    // drop a previous mesh
    dev_free(h->d_xyz); dev_free(h->d_tets); dev_free(h->d_facet_cell); dev_free(h->d_facet_verts);
    dev_free(h->d_bcell_parent); dev_free(h->d_btopology); dev_free(h->d_bvert_parent); dev_free(h->d_facet_local);
    dev_free(h->d_bcell_local); dev_free(h->d_blocal_soa); dev_free(h->d_glam); dev_free(h->d_normal);
    dev_free(h->d_area); dev_free(h->d_work); dev_free(h->d_m_lf); dev_free(h->d_m_w); dev_free(h->d_m_mat);
    dev_free(h->d_facet_nodes); dev_free(h->d_row); dev_free(h->d_wall_slot);
    k_free_run_buffers(h);
    h->order = 0;
    h->nv = nv;
    h->nc = nc;
    const int64_t nfaces = 4 * nc;

    int64_t* d_t64 = nullptr;
    int32_t *d_head = nullptr, *d_next = nullptr, *d_vcount = nullptr, *d_f2f = nullptr, *d_cnt = nullptr;
    int8_t* d_ext = nullptr;
    VH_TRY(dev_alloc(&h->d_xyz, 3 * nv));
    VH_TRY(dev_alloc(&h->d_tets, 4 * nc));
    VH_TRY(dev_alloc(&d_t64, 4 * nc));
    VH_TRY(dev_alloc(&d_head, nv));
    VH_TRY(dev_alloc(&d_next, nfaces));
    VH_TRY(dev_alloc(&d_vcount, nv));
    VH_TRY(dev_alloc(&d_f2f, nfaces));
    VH_TRY(dev_alloc(&d_ext, nfaces));
    VH_TRY(dev_alloc(&d_cnt, 8));
    VH_CUDA(cudaMemcpyAsync(h->d_xyz, xyz, sizeof(double) * 3 * nv, cudaMemcpyHostToDevice, st));
    VH_CUDA(cudaMemcpyAsync(d_t64, tets, sizeof(int64_t) * 4 * nc, cudaMemcpyHostToDevice, st));
    VH_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int32_t) * 8, st));
    VH_CUDA(cudaMemsetAsync(d_head, 0xFF, sizeof(int32_t) * nv, st));
    VH_CUDA(cudaMemsetAsync(d_vcount, 0, sizeof(int32_t) * nv, st));

    k0_order_cells<<<nblk(nc), TPB, 0, st>>>(d_t64, h->d_tets, nc, nv, d_cnt + 0);
    k0_link_faces<<<nblk(nfaces), TPB, 0, st>>>(h->d_tets, nfaces, d_head, d_next);
    k0_mark_exterior<<<nblk(nfaces), TPB, 0, st>>>(h->d_tets, nfaces, d_head, d_next, d_ext, d_vcount);
    VH_CUDA(cudaGetLastError());
    VH_TRY(exclusive_scan(d_vcount, d_vcount, nv, d_cnt + 1, st));
    int32_t cnt[8];
    VH_CUDA(cudaMemcpy(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
    VH_CHECK(cnt[0] == 0, VH_ERR_MESH, "vh_set_mesh: %d cells with out-of-range or repeated vertex ids", cnt[0]);
    const int64_t nF = cnt[1];
    VH_CHECK(nF > 0, VH_ERR_MESH, "vh_set_mesh: mesh has no exterior facets");
    h->nF = nF;

    VH_TRY(dev_alloc(&h->d_facet_cell, nF));
    VH_TRY(dev_alloc(&h->d_facet_local, nF));
    VH_TRY(dev_alloc(&h->d_facet_verts, 3 * nF));
    VH_TRY(dev_alloc(&h->d_bcell_parent, 3 * nF));
    VH_TRY(dev_alloc(&h->d_bcell_local, 3 * nF));
    VH_TRY(dev_alloc(&h->d_blocal_soa, 3 * nF));
    VH_TRY(dev_alloc(&h->d_btopology, 3 * nF));
    VH_TRY(dev_alloc(&h->d_glam, 12 * nF));
    VH_TRY(dev_alloc(&h->d_normal, 3 * nF));
    VH_TRY(dev_alloc(&h->d_area, nF));
    k0_rank_exterior<<<nblk(nfaces), TPB, 0, st>>>(h->d_tets, nfaces, d_head, d_next, d_ext, d_vcount, d_f2f,
                                                    h->d_facet_cell, h->d_facet_local, h->d_facet_verts);
    // boundary mesh + geometry; d_head is recycled as first_seen[v], d_next as scratch flags/positions
    VH_CUDA(cudaMemsetAsync(d_head, 0x7F, sizeof(int32_t) * nv, st));
    k0_geometry<<<nblk(nF), TPB, 0, st>>>(h->d_xyz, h->d_tets, nF, h->d_facet_cell, h->d_facet_local,
                                          h->d_facet_verts, h->d_bcell_parent, h->d_bcell_local, h->d_blocal_soa,
                                          h->d_glam, h->d_normal, h->d_area, d_head, d_cnt + 2);
    int32_t *d_flag = nullptr, *d_pos = nullptr;
    VH_TRY(dev_alloc(&d_flag, 3 * nF));
    VH_TRY(dev_alloc(&d_pos, 3 * nF));
    k0_flag_first<<<nblk(3 * nF), TPB, 0, st>>>(h->d_facet_verts, d_head, 3 * nF, d_flag);
    VH_CUDA(cudaGetLastError());
    VH_TRY(exclusive_scan(d_flag, d_pos, 3 * nF, d_cnt + 3, st));
    VH_CUDA(cudaMemcpy(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
    VH_CHECK(cnt[2] == 0, VH_ERR_MESH, "vh_set_mesh: %d wall cells are degenerate (zero volume)", cnt[2]);
    h->nBV = cnt[3];
    VH_TRY(dev_alloc(&h->d_bvert_parent, h->nBV));
    // vnumber[v] reuses d_vcount (voff no longer needed after k0_rank_exterior)
    k0_number_bverts<<<nblk(3 * nF), TPB, 0, st>>>(h->d_facet_verts, d_head, d_pos, 3 * nF, h->d_bvert_parent, d_vcount);
    k0_order_bcells<<<nblk(nF), TPB, 0, st>>>(d_vcount, nF, h->d_bcell_parent, h->d_bcell_local, h->d_blocal_soa, h->d_glam,
                                              h->d_btopology);

    // work list: singles (ascending facet id, padded to a warp multiple with -1), then multi-facet-cell facets
    int32_t *d_is_multi = d_flag, *d_is_single = d_pos, *d_pos_multi = nullptr, *d_pos_single = nullptr;
    VH_TRY(dev_alloc(&d_pos_multi, nF));
    VH_TRY(dev_alloc(&d_pos_single, nF));
    k0_flag_multi<<<nblk(nF), TPB, 0, st>>>(h->d_facet_cell, d_ext, nF, d_is_multi, d_is_single);
    k0_count_wall_cells<<<nblk(nF), TPB, 0, st>>>(h->d_facet_cell, h->d_facet_local, d_ext, nF, d_cnt + 6);
    VH_CUDA(cudaGetLastError());
    VH_TRY(exclusive_scan(d_is_multi, d_pos_multi, nF, d_cnt + 4, st));
    VH_TRY(exclusive_scan(d_is_single, d_pos_single, nF, d_cnt + 5, st));
    VH_CUDA(cudaMemcpy(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
    h->nMulti = cnt[4];
    h->nW = cnt[6];
    const int64_t nSingle = cnt[5];
    h->multi_start = (nSingle + 31) / 32 * 32;
    h->n_work = h->multi_start + h->nMulti;
    VH_TRY(dev_alloc(&h->d_work, h->n_work));
    VH_CUDA(cudaMemsetAsync(h->d_work, 0xFF, sizeof(int32_t) * h->n_work, st));
    k0_fill_work<<<nblk(nF), TPB, 0, st>>>(d_is_multi, d_pos_single, d_pos_multi, nF, h->multi_start, h->d_work);
    VH_TRY(dev_alloc(&h->d_m_lf, 4 * h->nMulti));
    VH_TRY(dev_alloc(&h->d_m_w, 36 * h->nMulti));
    if (h->nMulti > 0)
        k0_multi_weights<<<nblk(h->nMulti, 64), 64, 0, st>>>(h->d_work, h->multi_start, h->nMulti, h->d_facet_cell,
                                                              h->d_facet_local, d_f2f, h->d_bcell_local, h->d_area,
                                                              h->d_m_lf, h->d_m_w);
    VH_CUDA(cudaGetLastError());
    VH_CUDA(cudaStreamSynchronize(st));
    h->launches += 14;
    dev_free(d_t64); dev_free(d_head); dev_free(d_next); dev_free(d_vcount); dev_free(d_f2f); dev_free(d_ext);
    dev_free(d_cnt); dev_free(d_flag); dev_free(d_pos); dev_free(d_pos_multi); dev_free(d_pos_single);
    return VH_OK;
}

int k0_build_velocity_map(vh_handle* h, int order, const double* refined_xyz, int64_t n_nodes, double tol,
                          const int64_t* node_perm, int64_t n_slots) {
    VH_CHECK(h->nF > 0, VH_ERR_ARG, "vh_set_velocity_layout: call vh_set_mesh first");
    VH_CHECK(order == 1 || order == 2, VH_ERR_ARG, "vh_set_velocity_layout: order must be 1 or 2");
    cudaStream_t st = h->s_compute;
    const int64_t nF = h->nF;
    const int ndof = order == 2 ? 10 : 4;
    dev_free(h->d_facet_nodes);
    dev_free(h->d_row);
    dev_free(h->d_wall_slot);
    VH_TRY(dev_alloc(&h->d_facet_nodes, ndof * nF));
    VH_TRY(dev_alloc(&h->d_row, ndof * nF));
    if (order == 1) {
        VH_CHECK(n_nodes == h->nv, VH_ERR_ARG, "vh_set_velocity_layout: order 1 needs n_nodes == nv (%lld != %lld)",
                 (long long)n_nodes, (long long)h->nv);
        k0_p1_nodes<<<nblk(4 * nF), TPB, 0, st>>>(h->d_tets, nF, h->d_facet_cell, h->d_facet_nodes);
    } else {
        VH_CHECK(refined_xyz && n_nodes > 0, VH_ERR_ARG, "vh_set_velocity_layout: order 2 needs refined_xyz");
        VH_CHECK(n_nodes < (1LL << 31), VH_ERR_ARG, "too many velocity nodes for int32");
        VH_CHECK(tol > 0.0, VH_ERR_ARG, "vh_set_velocity_layout: tol must be > 0");
        // bounding box on the host (one pass over data that is being uploaded anyway)
        Grid g;
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int64_t i = 0; i < n_nodes; ++i)
            for (int d = 0; d < 3; ++d) {
                double x = refined_xyz[3 * i + d];
                lo[d] = x < lo[d] ? x : lo[d];
                hi[d] = x > hi[d] ? x : hi[d];
            }
        double ext[3], vol = 1.0, maxext = 0.0;
        for (int d = 0; d < 3; ++d) {
            ext[d] = hi[d] - lo[d];
            maxext = ext[d] > maxext ? ext[d] : maxext;
        }
        VH_CHECK(maxext > 0.0, VH_ERR_MESH, "refined mesh has zero extent");
        for (int d = 0; d < 3; ++d) vol *= (ext[d] > 1e-6 * maxext ? ext[d] : 1e-6 * maxext);
        double hcell = cbrt(vol / (double)n_nodes);
        if (hcell < 4.0 * tol) hcell = 4.0 * tol;
        int64_t ncell = 1;
        for (int d = 0; d < 3; ++d) {
            int64_t n = (int64_t)(ext[d] / hcell) + 1;
            if (n > 2048) n = 2048;
            g.n[d] = (int)n;
            g.lo[d] = lo[d];
            g.inv_h[d] = 1.0 / hcell;
            ncell *= n;
        }
        double* d_pts = nullptr;
        int32_t *d_head = nullptr, *d_next = nullptr, *d_un = nullptr;
        VH_TRY(dev_alloc(&d_pts, 3 * n_nodes));
        VH_TRY(dev_alloc(&d_head, ncell));
        VH_TRY(dev_alloc(&d_next, n_nodes));
        VH_TRY(dev_alloc(&d_un, 1));
        VH_CUDA(cudaMemcpyAsync(d_pts, refined_xyz, sizeof(double) * 3 * n_nodes, cudaMemcpyHostToDevice, st));
        VH_CUDA(cudaMemsetAsync(d_head, 0xFF, sizeof(int32_t) * ncell, st));
        VH_CUDA(cudaMemsetAsync(d_un, 0, sizeof(int32_t), st));
        k0_grid_insert<<<nblk(n_nodes), TPB, 0, st>>>(d_pts, n_nodes, g, d_head, d_next);
        k0_match_p2<<<nblk(10 * nF), TPB, 0, st>>>(h->d_xyz, h->d_tets, nF, h->d_facet_cell, d_pts, g, d_head, d_next,
                                                    tol, h->d_facet_nodes, d_un);
        VH_CUDA(cudaGetLastError());
        int32_t un = 0;
        VH_CUDA(cudaMemcpy(&un, d_un, sizeof(int32_t), cudaMemcpyDeviceToHost));
        dev_free(d_pts); dev_free(d_head); dev_free(d_next); dev_free(d_un);
        VH_CHECK(un == 0, VH_ERR_MESH,
                 "vh_set_velocity_layout: %d P2 nodes of wall cells have no refined-mesh vertex within tol=%g "
                 "(is <mesh>_refined_fluid.h5 the refinement of <mesh>_fluid.h5?)", un, tol);
        h->launches += 2;
    }
    int64_t* d_perm = nullptr;
    if (node_perm) {
        VH_TRY(dev_alloc(&d_perm, n_nodes));
        VH_CUDA(cudaMemcpyAsync(d_perm, node_perm, sizeof(int64_t) * n_nodes, cudaMemcpyHostToDevice, st));
    }
    VH_CHECK(n_slots < (1LL << 31), VH_ERR_ARG, "too many vector slots for int32");
    // wall-layer node list (ascending vector position) and the per-facet row table K2 reads the staged block with
    int32_t *d_flag = nullptr, *d_pos = nullptr, *d_cnt = nullptr;
    VH_TRY(dev_alloc(&d_flag, n_slots));
    VH_TRY(dev_alloc(&d_pos, n_slots));
    VH_TRY(dev_alloc(&d_cnt, 2));
    VH_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int32_t) * n_slots, st));
    VH_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int32_t) * 2, st));
    k0_mark_wall_nodes<<<nblk(ndof * nF), TPB, 0, st>>>(h->d_facet_nodes, h->d_bcell_local, h->d_facet_local, d_perm, nF,
                                                        ndof, n_slots, d_flag, d_cnt + 1);
    VH_CUDA(cudaGetLastError());
    VH_TRY(exclusive_scan(d_flag, d_pos, n_slots, d_cnt, st));
    int32_t cnt[2];
    VH_CUDA(cudaMemcpy(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
    if (cnt[1] != 0) {
        dev_free(d_flag); dev_free(d_pos); dev_free(d_cnt); dev_free(d_perm);
        VH_CHECK(false, VH_ERR_ARG, "vh_set_velocity_layout: node_perm has %d entries outside the vector", cnt[1]);
    }
    h->nWn = cnt[0];
    h->nWn_pad = (h->nWn + 31) / 32 * 32;
    VH_TRY(dev_alloc(&h->d_wall_slot, h->nWn_pad));
    k0_wall_slots<<<nblk(n_slots), TPB, 0, st>>>(d_flag, d_pos, n_slots, h->node_stride, h->nWn, h->nWn_pad,
                                                 h->d_wall_slot);
    k0_rows<<<nblk(ndof * nF), TPB, 0, st>>>(h->d_facet_nodes, h->d_bcell_local, h->d_facet_local, d_perm, d_pos, nF,
                                             ndof, h->d_row);
    dev_free(h->d_m_mat);
    VH_TRY(dev_alloc(&h->d_m_mat, h->nMulti * 3 * ndof * VH_MROW));
    if (h->nMulti > 0) {
        k0_multi_matrix<<<nblk(h->nMulti * ndof, 64), 64, 0, st>>>(h->nMulti, ndof, h->d_work, h->multi_start, nF, h->d_glam,
                                                                   h->d_m_lf, h->d_m_w, h->d_m_mat);
        h->launches += 1;
    }
    VH_CUDA(cudaGetLastError());
    VH_CUDA(cudaStreamSynchronize(st));
    // host copy of the wall-layer slots: the host-side compaction in front of the bus gathers with it (compact.cu)
    h->h_wall_slot.resize((size_t)h->nWn_pad);
    VH_CUDA(cudaMemcpy(h->h_wall_slot.data(), h->d_wall_slot, sizeof(int32_t) * (size_t)h->nWn_pad, cudaMemcpyDeviceToHost));
    dev_free(d_flag); dev_free(d_pos); dev_free(d_cnt);
    dev_free(d_perm);
    h->launches += 5;
    h->order = order;
    h->ndof = ndof;
    h->n_nodes = n_nodes;
    return VH_OK;
}
