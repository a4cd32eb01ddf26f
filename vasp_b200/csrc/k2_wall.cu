// K2/K3: per-snapshot wall traction + fused time reductions; K4: final index formulas.
//
// Replaces the body of the reference's snapshot loop (compute_hemodynamics.py:272-318):
//   u_p2 = T * u_p1                      (:275)  -> K1 staged the wall-layer dofs; the row table built by K0 is T
//   tau = stress()                       (:282)  -> closed-form P2/P1 gradient at the facet vertices, sigma n,
//                                                   tangential part; SurfaceProjector's solve is the identity for
//                                                   cells with one exterior facet and a precomputed 3x3-per-contributor
//                                                   weight for cells with several (K0)
//   TAWSS += |tau|, WSS_mean += tau      (:289-306)
//   TWSSG += project_dg(|dtau/dt|)       (:309-312) 7-point degree-5 rule, closed-form P1 mass inverse
// and the final formulas (:326-346).
//
// Work decomposition: one WARP per (facet, time segment); the 32 LANES are 32 consecutive snapshots.  With the
// time-major block W that K1 wrote, every load of a cell dof is 32 consecutive doubles (256 B): full sectors, two L1
// wavefronts.  Facet geometry is warp-uniform (broadcast loads), the multi-facet-cell branch is warp-uniform (no
// divergence), tau of the previous snapshot comes from the neighbouring lane (shuffle) and each lane keeps its
// share of the 15 running sums in registers until one fixed-order butterfly at the end.  A segment that has a
// predecessor column computes it in lane 0 of its first pass instead of communicating with the previous segment
// (TWSSG's one-step dependence).  Partial sums of the segments go to `part` and are folded into the running sums
// by k3_fold in fixed order: results are bitwise reproducible for a given launch shape.
//
// Local vertex labels are facet-canonical (K0): 0,1,2 = the facet's vertices in boundary-cell order, 3 = the
// opposite vertex; P2 edge dofs 4..9 = e01,e02,e12,e03,e13,e23.
#include <math.h>

#include "common.cuh"

namespace {

template <int ORDER>
struct Dofs {
    static constexpr int N = ORDER == 2 ? 10 : 4;
};

__host__ __device__ constexpr int edge_dof(int a, int b) {
    // canonical edge order e01,e02,e12,e03,e13,e23 -> 4..9
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    return hi == 1 ? 4 : hi == 2 ? 5 + lo : 7 + lo;
}
static_assert(edge_dof(0, 1) == 4 && edge_dof(0, 2) == 5 && edge_dof(1, 2) == 6 && edge_dof(0, 3) == 7 &&
                  edge_dof(1, 3) == 8 && edge_dof(2, 3) == 9 && edge_dof(3, 1) == 8,
              "edge table");

// FIAT default degree-5 triangle rule (Strang-Fix 7 points) in barycentric form; weights are area fractions.
constexpr double Q_A = 0.10128650732345633, Q_B = 0.79742698535308720;
constexpr double Q_C = 0.47014206410511505, Q_D = 0.05971587178976981;
constexpr double Q_W0 = 0.225, Q_W1 = 0.12593918054482717, Q_W2 = 0.13239415278850616;

struct Vel {
    // velocity of the cell dofs, one snapshot
    double x[10], y[10], z[10];
};

// velocity of the cell dofs at column `col` of the staged block; base[k] = 3 * row[k] * ld
template <int ORDER>
__device__ __forceinline__ void load_vel(const double* __restrict__ W, const int64_t (&base)[10], int64_t ld,
                                         int64_t col, Vel& v) {
#pragma unroll
    for (int k = 0; k < Dofs<ORDER>::N; ++k) {
        const double* p = W + base[k] + col;
        v.x[k] = __ldg(p);
        v.y[k] = __ldg(p + ld);
        v.z[k] = __ldg(p + 2 * ld);
    }
}

// Tangential traction Ft = F - (F.n) n, F = -mu (grad u + grad u^T) n, at local vertices listed in VS..., for the
// face with unit normal n.  g[a] = grad lambda_a.
//   P2: grad u (v_a) = H + 4 (u_a (x) g_a + sum_{b != a} u_ab (x) g_b),  H = -sum_b u_b (x) g_b
//   P1: grad u       = -H  (constant)
// Only G n and G^T n are formed:  G n = sum_nodes u_node (grad phi_node . n),  G^T n = sum grad phi_node (u_node . n).
template <int ORDER, int V>
__device__ __forceinline__ void ft_vertex(const double (&g)[4][3], const double (&n)[3], const double (&gam)[4],
                                          const Vel& v, const double (&un)[10], const double (&c)[3], double mu,
                                          double (&ft)[3]) {
    double s[3];
    if (ORDER == 2) {
        double ex = v.x[V] * gam[V], ey = v.y[V] * gam[V], ez = v.z[V] * gam[V];
        double tx = g[V][0] * un[V], ty = g[V][1] * un[V], tz = g[V][2] * un[V];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b == V) continue;
            const int e = edge_dof(V, b);
            ex = fma(v.x[e], gam[b], ex);
            ey = fma(v.y[e], gam[b], ey);
            ez = fma(v.z[e], gam[b], ez);
            tx = fma(g[b][0], un[e], tx);
            ty = fma(g[b][1], un[e], ty);
            tz = fma(g[b][2], un[e], tz);
        }
        s[0] = fma(4.0, ex + tx, c[0]);
        s[1] = fma(4.0, ey + ty, c[1]);
        s[2] = fma(4.0, ez + tz, c[2]);
    } else {
        s[0] = -c[0];
        s[1] = -c[1];
        s[2] = -c[2];
    }
    double fx = -mu * s[0], fy = -mu * s[1], fz = -mu * s[2];
    double fn = fx * n[0] + fy * n[1] + fz * n[2];
    ft[0] = fma(-fn, n[0], fx);
    ft[1] = fma(-fn, n[1], fy);
    ft[2] = fma(-fn, n[2], fz);
}

// common part c = H n + H^T n and un = u_node . n
template <int ORDER>
__device__ __forceinline__ void face_common(const double (&g)[4][3], const double (&n)[3], const double (&gam)[4],
                                            const Vel& v, double (&un)[10], double (&c)[3]) {
#pragma unroll
    for (int k = 0; k < Dofs<ORDER>::N; ++k) un[k] = fma(v.z[k], n[2], fma(v.y[k], n[1], v.x[k] * n[0]));
    c[0] = c[1] = c[2] = 0.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        c[0] = fma(-v.x[b], gam[b], fma(-g[b][0], un[b], c[0]));
        c[1] = fma(-v.y[b], gam[b], fma(-g[b][1], un[b], c[1]));
        c[2] = fma(-v.z[b], gam[b], fma(-g[b][2], un[b], c[2]));
    }
}

// tau[3*j + c] for a facet whose cell owns no other exterior facet
template <int ORDER>
__device__ __forceinline__ void tau_single(const double (&g)[4][3], const double (&n)[3], const double (&gam)[4],
                                           const Vel& v, double mu, double (&tau)[9]) {
    double un[10], c[3], ft[3];
    face_common<ORDER>(g, n, gam, v, un, c);
    ft_vertex<ORDER, 0>(g, n, gam, v, un, c, mu, ft);
    tau[0] = ft[0]; tau[1] = ft[1]; tau[2] = ft[2];
    if (ORDER == 2) {
        ft_vertex<ORDER, 1>(g, n, gam, v, un, c, mu, ft);
        tau[3] = ft[0]; tau[4] = ft[1]; tau[5] = ft[2];
        ft_vertex<ORDER, 2>(g, n, gam, v, un, c, mu, ft);
        tau[6] = ft[0]; tau[7] = ft[1]; tau[8] = ft[2];
    } else {
        tau[3] = tau[6] = ft[0]; tau[4] = tau[7] = ft[1]; tau[5] = tau[8] = ft[2];
    }
}

template <int ORDER, int A, int V, int KK>
__device__ __forceinline__ void multi_vertex(const double (&g)[4][3], const double (&n)[3], const double (&gam)[4],
                                             const Vel& v, const double (&un)[10], const double (&c)[3], double mu,
                                             const double* __restrict__ w, int64_t wstride, double (&tau)[9]) {
    double ft[3];
    ft_vertex<ORDER, V>(g, n, gam, v, un, c, mu, ft);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double wj = w[(int64_t)(9 * A + 3 * j + KK) * wstride];
        tau[3 * j + 0] = fma(wj, ft[0], tau[3 * j + 0]);
        tau[3 * j + 1] = fma(wj, ft[1], tau[3 * j + 1]);
        tau[3 * j + 2] = fma(wj, ft[2], tau[3 * j + 2]);
    }
}

// contributor = face opposite canonical local vertex A; its vertices are the other three in ascending label order
template <int ORDER, int A>
__device__ __forceinline__ void multi_face(const double (&g)[4][3], const Vel& v, double mu,
                                           const double* __restrict__ w, int64_t wstride, double (&tau)[9]) {
    double n[3], gam[4], un[10], c[3];
    double inv = -1.0 / sqrt(g[A][0] * g[A][0] + g[A][1] * g[A][1] + g[A][2] * g[A][2]);
    n[0] = g[A][0] * inv; n[1] = g[A][1] * inv; n[2] = g[A][2] * inv;
#pragma unroll
    for (int b = 0; b < 4; ++b) gam[b] = g[b][0] * n[0] + g[b][1] * n[1] + g[b][2] * n[2];
    face_common<ORDER>(g, n, gam, v, un, c);
    constexpr int V0 = A == 0 ? 1 : 0, V1 = A <= 1 ? 2 : 1, V2 = A <= 2 ? 3 : 2;
    multi_vertex<ORDER, A, V0, 0>(g, n, gam, v, un, c, mu, w, wstride, tau);
    multi_vertex<ORDER, A, V1, 1>(g, n, gam, v, un, c, mu, w, wstride, tau);
    multi_vertex<ORDER, A, V2, 2>(g, n, gam, v, un, c, mu, w, wstride, tau);
}

template <int ORDER>
__device__ __noinline__ void tau_multi(const double (&g)[4][3], const Vel& v, double mu,
                                       const int8_t* __restrict__ m_lf, const double* __restrict__ m_w, int64_t m,
                                       int64_t nMulti, double (&tau)[9]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) tau[i] = 0.0;
    const double* w = m_w + m;
    if (m_lf[0 * nMulti + m] >= 0) multi_face<ORDER, 0>(g, v, mu, w, nMulti, tau);
    if (m_lf[1 * nMulti + m] >= 0) multi_face<ORDER, 1>(g, v, mu, w, nMulti, tau);
    if (m_lf[2 * nMulti + m] >= 0) multi_face<ORDER, 2>(g, v, mu, w, nMulti, tau);
    multi_face<ORDER, 3>(g, v, mu, w, nMulti, tau);
}

__device__ __forceinline__ double norm3(double a, double b, double c) { return sqrt(fma(a, a, fma(b, b, c * c))); }

// P(|w|) on the boundary triangle: p_j = 12 s_j - 3 sum_i s_i,  s_i = sum_q wq phi_i(x_q) |w(x_q)|   (area cancels)
__device__ __forceinline__ void twssg_project(const double (&w)[9], double (&p)[3]) {
    double sx = w[0] + w[3] + w[6], sy = w[1] + w[4] + w[7], sz = w[2] + w[5] + w[8];
    double m0 = norm3(sx, sy, sz) * (1.0 / 3.0);
    double ax = Q_A * sx, ay = Q_A * sy, az = Q_A * sz;
    double cx = Q_C * sx, cy = Q_C * sy, cz = Q_C * sz;
    constexpr double BA = Q_B - Q_A, DC = Q_D - Q_C;
    double mb[3], md[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        mb[j] = norm3(fma(BA, w[3 * j], ax), fma(BA, w[3 * j + 1], ay), fma(BA, w[3 * j + 2], az));  // weight b on j
        md[j] = norm3(fma(DC, w[3 * j], cx), fma(DC, w[3 * j + 1], cy), fma(DC, w[3 * j + 2], cz));  // weight d on j
    }
    double sb = mb[0] + mb[1] + mb[2], sd = md[0] + md[1] + md[2];
    double tot = Q_W0 * m0 + Q_W1 * sb + Q_W2 * sd;  // sum_i s_i  (phi sums to one)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        double sj = Q_W0 * (1.0 / 3.0) * m0 + Q_W1 * fma(BA, mb[j], Q_A * sb) + Q_W2 * fma(DC, md[j], Q_C * sd);
        p[j] = 12.0 * sj - 3.0 * tot;
    }
}

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }

struct K2Args {
    FacetTables T;
    const double* W;        // staged block (K1)
    int64_t ld;
    int ncol;               // columns in the block
    int r0;                 // first real column (1 when column 0 is a halo snapshot that only seeds tau_prev)
    int seg_len;            // real snapshots per segment (blockIdx.y), a multiple of K2_COLS
    int prev_mode;          // tau_prev of the first real column when r0 == 0: 0 zero, 1 tau_last_in
    const double* tau_last_in;
    double* tau_last_out;   // [9][nF]
    double* part;           // [gridDim.y][15][nF]
    double* wss_out;        // [ncol - r0][nF][9] or null
    double mu, inv_dt;
};

constexpr int K2_WARPS = 4;
constexpr int K2_COLS = 31;  // real columns per lane pass; lane 0 recomputes the column before them

// MULTI = false: facets whose cell owns no other exterior facet (work[0, multi_start)); MULTI = true: the rest.
template <int ORDER, bool MULTI>
__global__ void __launch_bounds__(32 * K2_WARPS) k2_wall(const K2Args a) {
    const FacetTables& T = a.T;
    const int64_t nF = T.nF;
    const int lane = threadIdx.x & 31;
    const int64_t wi = (int64_t)blockIdx.x * K2_WARPS + (threadIdx.x >> 5);
    const int64_t w = MULTI ? wi + T.multi_start : wi;
    if (w >= (MULTI ? T.n_work : T.multi_start)) return;
    const int32_t f = T.work[w];
    if (f < 0) return;  // padding entry (warp-uniform)
    // P1 data in a cell with one exterior facet: tau is the same at the three facet vertices, so one magnitude and
    // P(|w|) = |w| exactly (the projection reproduces constants)
    constexpr bool FLAT = (ORDER == 1) && !MULTI;
    constexpr int NT = FLAT ? 3 : 9;

    int64_t base[10];
    double g[4][3], n[3], gam[4];
#pragma unroll
    for (int k = 0; k < Dofs<ORDER>::N; ++k) base[k] = 3 * (int64_t)T.row[(int64_t)k * nF + f] * a.ld;
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int d = 0; d < 3; ++d) g[b][d] = T.glam[(int64_t)(3 * b + d) * nF + f];
#pragma unroll
    for (int d = 0; d < 3; ++d) n[d] = T.normal[(int64_t)d * nF + f];
#pragma unroll
    for (int b = 0; b < 4; ++b) gam[b] = g[b][0] * n[0] + g[b][1] * n[1] + g[b][2] * n[2];

    const int seg0 = a.r0 + (int)blockIdx.y * a.seg_len;  // first real column of this segment
    const int seg1 = min(seg0 + a.seg_len, a.ncol);

    double acc[VH_NSUM];
#pragma unroll
    for (int i = 0; i < VH_NSUM; ++i) acc[i] = 0.0;

    for (int c0 = seg0; c0 < seg1; c0 += K2_COLS) {
        const int col = c0 + lane - 1;  // lane 0: the column before this pass
        const bool live = lane > 0 && col < seg1;
        Vel v;
        load_vel<ORDER>(a.W, base, a.ld, min(max(col, 0), seg1 - 1), v);
        double tau[9];
        if (MULTI)
            tau_multi<ORDER>(g, v, a.mu, T.m_lf, T.m_w, wi, T.nMulti, tau);
        else
            tau_single<ORDER>(g, n, gam, v, a.mu, tau);
        if (col < 0) {  // no column before the block's first: tau_prev is zero or carried over from the last launch
#pragma unroll
            for (int i = 0; i < NT; ++i) tau[i] = a.prev_mode == 1 ? a.tau_last_in[(int64_t)i * nF + f] : 0.0;
        }
        double dw[9];
#pragma unroll
        for (int i = 0; i < NT; ++i) dw[i] = (tau[i] - shfl_up1(tau[i])) * a.inv_dt;
        if (live) {
            if (FLAT) {
#pragma unroll
                for (int i = 0; i < 3; ++i) acc[i] += tau[i];
                acc[9] += norm3(tau[0], tau[1], tau[2]);
                acc[12] += norm3(dw[0], dw[1], dw[2]);
#pragma unroll
                for (int i = 3; i < 9; ++i) tau[i] = tau[i - 3];
            } else {
                double p[3];
#pragma unroll
                for (int i = 0; i < 9; ++i) acc[i] += tau[i];
#pragma unroll
                for (int j = 0; j < 3; ++j) acc[9 + j] += norm3(tau[3 * j], tau[3 * j + 1], tau[3 * j + 2]);
                twssg_project(dw, p);
#pragma unroll
                for (int j = 0; j < 3; ++j) acc[12 + j] += p[j];
            }
            if (a.wss_out) {
                double* o = a.wss_out + ((int64_t)(col - a.r0) * nF + f) * 9;
#pragma unroll
                for (int i = 0; i < 9; ++i) o[i] = tau[i];
            }
            if (col == a.ncol - 1) {
#pragma unroll
                for (int i = 0; i < 9; ++i) a.tau_last_out[(int64_t)i * nF + f] = tau[i];
            }
        }
    }

    // fixed-order butterfly over the 32 lanes, then lane 0 stores the segment's partial sums
#pragma unroll
    for (int i = 0; i < VH_NSUM; ++i) {
        if (FLAT && !(i < 3 || i == 9 || i == 12)) continue;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], d);
    }
    if (FLAT) {
#pragma unroll
        for (int i = 0; i < 3; ++i) acc[3 + i] = acc[6 + i] = acc[i];
        acc[10] = acc[11] = acc[9];
        acc[13] = acc[14] = acc[12];
    }
    if (lane == 0) {
        double* p = a.part + (int64_t)blockIdx.y * VH_NSUM * nF + f;
#pragma unroll
        for (int i = 0; i < VH_NSUM; ++i) p[(int64_t)i * nF] = acc[i];
    }
}

// sums[r][f] += part[0][r][f] + part[1][r][f] + ...   (fixed order)
__global__ void k3_fold(double* __restrict__ sums, const double* __restrict__ part, int64_t n, int64_t groups) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double t = sums[i];
    for (int64_t gq = 0; gq < groups; ++gq) t += part[gq * n + i];
    sums[i] = t;
}

// compute_hemodynamics.py:326-346
__global__ void k4_indices(const double* __restrict__ sums, int64_t nF, double count, double* __restrict__ tawss,
                           double* __restrict__ osi, double* __restrict__ rrt, double* __restrict__ ecap,
                           double* __restrict__ twssg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * nF) return;
    int64_t f = i % nF;
    int j = (int)(i / nF);
    double mx = sums[(int64_t)(3 * j + 0) * nF + f] / count;
    double my = sums[(int64_t)(3 * j + 1) * nF + f] / count;
    double mz = sums[(int64_t)(3 * j + 2) * nF + f] / count;
    double mean_mag = norm3(mx, my, mz);
    double ta = sums[(int64_t)(9 + j) * nF + f] / count;
    double o = 0.5 * (1.0 - mean_mag / ta);
    int64_t q = 3 * f + j;
    tawss[q] = ta;
    osi[q] = o;
    rrt[q] = 1.0 / mean_mag;
    ecap[q] = o / ta;
    twssg[q] = sums[(int64_t)(12 + j) * nF + f] / count;
}

}  // namespace

FacetTables vh_tables(const vh_handle* h) {
    FacetTables T;
    T.nF = h->nF;
    T.row = h->d_row;
    T.glam = h->d_glam;
    T.normal = h->d_normal;
    T.work = h->d_work;
    T.n_work = h->n_work;
    T.multi_start = h->multi_start;
    T.m_lf = h->d_m_lf;
    T.m_w = h->d_m_w;
    T.nMulti = h->nMulti;
    return T;
}

int k_free_run_buffers(vh_handle* h) {
    if (h->d_sums) cudaFree(h->d_sums);
    if (h->d_tau_last[0]) cudaFree(h->d_tau_last[0]);
    if (h->d_tau_last[1]) cudaFree(h->d_tau_last[1]);
    if (h->d_part) cudaFree(h->d_part);
    if (h->d_out5) cudaFree(h->d_out5);
    h->d_sums = h->d_tau_last[0] = h->d_tau_last[1] = h->d_part = h->d_out5 = nullptr;
    h->part_cap = 0;
    for (int i = 0; i < 2; ++i) {
        if (h->d_stage[i]) cudaFree(h->d_stage[i]);
        if (h->d_wss_stage[i]) cudaFree(h->d_wss_stage[i]);
        h->d_stage[i] = h->d_wss_stage[i] = nullptr;
    }
    h->stage_cap = h->wss_stage_cap = 0;
    if (h->d_W) cudaFree(h->d_W);
    h->d_W = nullptr;
    h->w_ld = 0;
    h->begun = false;
    return VH_OK;
}

// Columns per staged block: as many as fit ~1/8 of the free memory (at most 4 GiB), between 64 and 4096; a
// batch_snapshots tuning value caps it (one column more than the batch, for the halo)
static int ensure_stage_block(vh_handle* h, int64_t want_cols) {
    int64_t cap = h->w_ld;
    if (h->batch_snapshots > 0 && want_cols > h->batch_snapshots + 1) want_cols = h->batch_snapshots + 1;
    if (cap >= want_cols || cap >= 4096) return VH_OK;
    size_t free_b = 0, total_b = 0;
    VH_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (h->d_W) free_b += (size_t)(h->nWn_pad * 24 * h->w_ld);
    int64_t budget = (int64_t)(free_b / 8);
    if (budget > (4LL << 30)) budget = 4LL << 30;
    int64_t cols = budget / (h->nWn_pad * 24);
    if (cols > 4096) cols = 4096;
    if (cols > want_cols) cols = want_cols;
    if (cols < 64 && h->batch_snapshots <= 0) cols = 64;
    cols = (cols + 31) / 32 * 32;
    if (cols <= cap) return VH_OK;
    if (h->d_W) cudaFree(h->d_W);
    h->d_W = nullptr;
    h->w_ld = 0;
    VH_CUDA(cudaMalloc(&h->d_W, (size_t)(h->nWn_pad * 24 * cols)));
    h->w_ld = cols;
    return VH_OK;
}

int k2_launch(vh_handle* h, const double* d_u, int64_t n_snap, int64_t stride_elems, int prev_mode, double* d_wss) {
    if (n_snap <= 0) return VH_OK;
    const int64_t nF = h->nF;
    VH_TRY(ensure_stage_block(h, n_snap + 1));
    const int64_t n_single = h->multi_start, n_multi = h->n_work - h->multi_start;
    const unsigned gx_single = (unsigned)((n_single + K2_WARPS - 1) / K2_WARPS);
    const unsigned gx_multi = (unsigned)((n_multi + K2_WARPS - 1) / K2_WARPS);
    int64_t pos = 0;
    while (pos < n_snap) {
        const int halo = (pos == 0 && prev_mode == 2) ? 1 : 0;
        int64_t nb = n_snap - pos;
        if (nb + halo > h->w_ld) nb = h->w_ld - halo;
        const int64_t ncol = nb + halo;
        // segment length = p lane passes of 31 real snapshots: the fewest segments that still give every SM ~64
        // (facet, segment) warps to schedule
        int64_t p = (h->chunk_snapshots + K2_COLS - 1) / K2_COLS;
        if (p <= 0) {
            const int64_t target_warps = (int64_t)h->sm_count * 64;
            for (p = (nb + K2_COLS - 1) / K2_COLS; p > 1; --p)
                if (h->n_work * ((nb + K2_COLS * p - 1) / (K2_COLS * p)) >= target_warps) break;
        }
        const int64_t seg = K2_COLS * p;
        const int64_t gy = (nb + seg - 1) / seg;
        VH_CHECK(gy <= 65535, VH_ERR_ARG, "k2_launch: too many segments (%lld); raise chunk_snapshots", (long long)gy);
        if (gy > h->part_cap) {
            if (h->d_part) cudaFree(h->d_part);
            h->d_part = nullptr;
            h->part_cap = 0;
            VH_CUDA(cudaMalloc(&h->d_part, sizeof(double) * VH_NSUM * nF * gy));
            h->part_cap = gy;
        }
        const bool prof = h->profile && h->prof_used + 3 <= h->prof_pool.size();
        if (prof) cudaEventRecord(h->prof_pool[h->prof_used], h->s_compute);
        VH_TRY(k1_launch(h, d_u + (pos - halo) * stride_elems, ncol, stride_elems));
        if (prof) cudaEventRecord(h->prof_pool[h->prof_used + 1], h->s_compute);
        K2Args a;
        a.T = vh_tables(h);
        a.W = h->d_W;
        a.ld = h->w_ld;
        a.ncol = (int)ncol;
        a.r0 = halo;
        a.seg_len = (int)seg;
        a.prev_mode = pos == 0 ? prev_mode : 1;
        a.tau_last_in = h->d_tau_last[h->tau_cur];
        a.tau_last_out = h->d_tau_last[h->tau_cur ^ 1];
        h->tau_cur ^= 1;
        a.part = h->d_part;
        a.wss_out = d_wss ? d_wss + pos * nF * 9 : nullptr;
        a.mu = h->mu;
        a.inv_dt = 1.0 / h->dt;
        dim3 block(32 * K2_WARPS);
        if (gx_single) {
            dim3 grid(gx_single, (unsigned)gy);
            if (h->order == 2)
                k2_wall<2, false><<<grid, block, 0, h->s_compute>>>(a);
            else
                k2_wall<1, false><<<grid, block, 0, h->s_compute>>>(a);
            h->launches += 1;
        }
        if (gx_multi) {
            dim3 grid(gx_multi, (unsigned)gy);
            if (h->order == 2)
                k2_wall<2, true><<<grid, block, 0, h->s_compute>>>(a);
            else
                k2_wall<1, true><<<grid, block, 0, h->s_compute>>>(a);
            h->launches += 1;
        }
        if (prof) {
            cudaEventRecord(h->prof_pool[h->prof_used + 2], h->s_compute);
            h->prof_used += 3;
        }
        VH_CUDA(cudaGetLastError());
        const int64_t n = VH_NSUM * nF;
        k3_fold<<<(unsigned)((n + 255) / 256), 256, 0, h->s_compute>>>(h->d_sums, h->d_part, n, gy);
        VH_CUDA(cudaGetLastError());
        h->launches += 1;
        pos += nb;
    }
    h->count += n_snap;
    h->have_tau_last = true;
    return VH_OK;
}

int k4_finalize(vh_handle* h, int64_t n_total, double* d_out5) {
    const int64_t nF = h->nF, n3 = 3 * nF;
    k4_indices<<<(unsigned)((n3 + 255) / 256), 256, 0, h->s_compute>>>(h->d_sums, nF, (double)n_total, d_out5,
                                                                       d_out5 + n3, d_out5 + 2 * n3, d_out5 + 3 * n3,
                                                                       d_out5 + 4 * n3);
    VH_CUDA(cudaGetLastError());
    h->launches += 1;
    return VH_OK;
}
