// K2/K3: per-snapshot wall traction + fused time reductions; K4: final index formulas.
//
// Replaces the body of the reference's snapshot loop (compute_hemodynamics.py:272-318):
//   u_p2 = T * u_p1                      (:275)  -> K1 staged the wall-layer dofs; the row table built by K0 is T
//   tau = stress()                       (:282)  -> closed-form P2/P1 gradient at the facet vertices, sigma n,
//                                                   tangential part; SurfaceProjector's solve is the identity for
//                                                   cells with one exterior facet and a precomputed dense operator
//                                                   (K0: block solve folded with the contributing faces) otherwise
//   TAWSS += |tau|, WSS_mean += tau      (:289-306)
//   TWSSG += project_dg(|dtau/dt|)       (:309-312) 7-point degree-5 rule, closed-form P1 mass inverse
// and the final formulas (:326-346).
//
// Work decomposition: one WARP per (facet, time segment); the 32 LANES are 32 consecutive snapshots.  With the
// time-major block W that K1 wrote, every load of a cell dof is 32 consecutive doubles (256 B): full sectors, two L1
// wavefronts.  Facet geometry is warp-uniform (broadcast loads), facets of multi-facet cells get their own launch
// (no divergence), tau of the previous snapshot comes from the neighbouring lane (shuffle) and each lane keeps its
// share of the 15 running sums in registers until one fixed-order butterfly at the end.  Lane 0 of every lane pass
// recomputes the column before the pass instead of communicating with the previous pass or segment (TWSSG's
// one-step dependence), so a pass advances 31 snapshots.  The rows of the next pass are in flight while the
// current one is computed: cp.async into a per-warp shared-memory ring for P2 (30 rows; every lane copies and later
// reads only its own column, so no barrier), a register double buffer for P1 (12 rows).  Partial sums of the
// segments go to `part` and are folded into the running sums by k3_fold in fixed order: results are bitwise
// reproducible for a given launch shape.  (A TMA variant -- one cp.async.bulk of 288 B per row and pass, mbarrier
// completion -- was measured 10-20 % slower than cp.async.ca: the copies bypass L1, where neighbouring facets share
// rows, and are too small to amortise; see DESIGN.md.)
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
    int pass_base, pass_extra;  // segment y owns pass_base (+1 if y < pass_extra) lane passes of K2_COLS columns
    int prev_mode;          // tau_prev of the first real column when r0 == 0: 0 zero, 1 tau_last_in
    const double* tau_last_in;
    double* tau_last_out;   // [9][nF]
    double* part;           // [gridDim.y][15][n_work]
    double* wss_out;        // [ncol - r0][nF][9] or null
    double mu, inv_dt;
};

constexpr int K2_WARPS = 4;
constexpr int K2_COLS = 31;  // real columns per lane pass; lane 0 recomputes the column before them

__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Shared-memory ring of the cp.async path: [warp][stage][3 * dof + c][lane]
template <int ORDER>
constexpr int k2_smem_bytes() {
    return ORDER == 2 ? K2_WARPS * 2 * 3 * Dofs<ORDER>::N * 32 * (int)sizeof(double) : 0;
}

// tau of a facet whose cell owns several exterior facets: dense operator from K0, warp-uniform coefficient loads
template <int ORDER>
__device__ __forceinline__ void tau_dense(const double* __restrict__ M, const Vel& v, double mu, double (&tau)[9]) {
#pragma unroll
    for (int i = 0; i < 9; ++i) tau[i] = 0.0;
#pragma unroll
    for (int k = 0; k < Dofs<ORDER>::N; ++k) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double val = c == 0 ? v.x[k] : c == 1 ? v.y[k] : v.z[k];
            const double2* row = reinterpret_cast<const double2*>(M + (3 * k + c) * VH_MROW);
#pragma unroll
            for (int h2 = 0; h2 < 5; ++h2) {
                const double2 m2 = __ldg(row + h2);
                tau[2 * h2] = fma(m2.x, val, tau[2 * h2]);
                if (h2 < 4) tau[2 * h2 + 1] = fma(m2.y, val, tau[2 * h2 + 1]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) tau[i] *= -mu;
}

// One warp = one facet x one time segment; the 32 lanes are 32 consecutive columns of the staged block.
// MULTI = false: facets whose cell owns no other exterior facet (work[0, multi_start)); MULTI = true: the rest.
template <int ORDER, bool MULTI>
__global__ void __launch_bounds__(32 * K2_WARPS, (ORDER == 2 || MULTI) ? 3 : 4) k2_wall(const K2Args a) {
    extern __shared__ __align__(16) double k2_smem[];
    constexpr int N = Dofs<ORDER>::N;
    constexpr bool RING = ORDER == 2;  // cp.async shared-memory ring (else: register double buffer)
    const FacetTables& T = a.T;
    const int64_t nF = T.nF;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t wi = (int64_t)blockIdx.x * K2_WARPS + wib;
    const int64_t w = MULTI ? wi + T.multi_start : wi;
    if (w >= (MULTI ? T.n_work : T.multi_start)) return;
    const int32_t f = T.work[w];
    if (f < 0) return;  // padding entry (warp-uniform)
    // P1 data in a cell with one exterior facet: tau is the same at the three facet vertices, so one magnitude and
    // P(|w|) = |w| exactly (the projection reproduces constants)
    constexpr bool FLAT = (ORDER == 1) && !MULTI;
    constexpr int NT = FLAT ? 3 : 9;

    const int y = blockIdx.y;
    const int pass0 = y * a.pass_base + min(y, a.pass_extra);
    const int npass = a.pass_base + (y < a.pass_extra ? 1 : 0);
    const int seg0 = a.r0 + pass0 * K2_COLS;  // first real column of this segment
    const int seg1 = min(seg0 + npass * K2_COLS, a.ncol);

    // x-component row of every cell dof (y, z follow at +ld, +2 ld), already offset to this lane's first column
    const double* rp[N];
#pragma unroll
    for (int k = 0; k < N; ++k)
        rp[k] = a.W + 3 * (int64_t)T.row[(int64_t)k * nF + f] * a.ld + (seg0 + lane - 1);
    const int64_t ld = a.ld;
    // column offset of pass j relative to rp: the lane's column is clamped into [0, seg1)
    auto pass_off = [&](int j) {
        const int col = seg0 + j * K2_COLS + lane - 1;
        return min(max(col, 0), seg1 - 1) - (seg0 + lane - 1);
    };

    double* const ring = k2_smem + (size_t)wib * 2 * 3 * N * 32 + lane;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    Vel vn;  // register double buffer (P1)
    auto prefetch = [&](int j) {
        const int off = pass_off(j);
        if (RING) {
            const uint32_t dst = ring_s + (uint32_t)((j & 1) * 3 * N * 32 * sizeof(double));
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double* p = rp[k] + off;
                cp_async8(dst + (3 * k + 0) * 256, p);
                cp_async8(dst + (3 * k + 1) * 256, p + ld);
                cp_async8(dst + (3 * k + 2) * 256, p + 2 * ld);
            }
            cp_async_commit();
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double* p = rp[k] + off;
                vn.x[k] = __ldg(p);
                vn.y[k] = __ldg(p + ld);
                vn.z[k] = __ldg(p + 2 * ld);
            }
        }
    };
    prefetch(0);

    double g[4][3], n[3], gam[4];
    const double* M = nullptr;
    if (MULTI) {
        M = T.m_mat + (size_t)wi * 3 * N * VH_MROW;
    } else {
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int d = 0; d < 3; ++d) g[b][d] = T.glam[(int64_t)(3 * b + d) * nF + f];
#pragma unroll
        for (int d = 0; d < 3; ++d) n[d] = T.normal[(int64_t)d * nF + f];
#pragma unroll
        for (int b = 0; b < 4; ++b) gam[b] = g[b][0] * n[0] + g[b][1] * n[1] + g[b][2] * n[2];
    }

    // tau_prev of the block's first column is external (zero, or carried over from the last launch): lane i holds
    // component i.  It is fetched here and handed to lane 0 by shuffle in the first pass -- a (predicated) load inside
    // the pass loop shares a scoreboard with the prefetch loads and makes every pass wait for its own prefetch
    // (ncu r1i: 41 % of all stall samples on the first shuffle after the prefetch).
    const bool seeded = seg0 == 0;
    double prev_seed = 0.0;
    if (seeded && a.prev_mode == 1 && lane < NT) prev_seed = a.tau_last_in[(int64_t)lane * nF + f];

    double acc[VH_NSUM];
#pragma unroll
    for (int i = 0; i < VH_NSUM; ++i) acc[i] = 0.0;

    for (int j = 0; j < npass; ++j) {
        Vel v;
        if (RING) {
            if (j + 1 < npass) {
                prefetch(j + 1);
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            const double* sv = ring + (j & 1) * 3 * N * 32;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                v.x[k] = sv[(3 * k + 0) * 32];
                v.y[k] = sv[(3 * k + 1) * 32];
                v.z[k] = sv[(3 * k + 2) * 32];
            }
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) {
                v.x[k] = vn.x[k];
                v.y[k] = vn.y[k];
                v.z[k] = vn.z[k];
            }
            if (j + 1 < npass) prefetch(j + 1);
        }
        const int col = seg0 + j * K2_COLS + lane - 1;  // lane 0: the column before this pass
        const bool live = lane > 0 && col < seg1;
        double tau[9];
        if (MULTI)
            tau_dense<ORDER>(M, v, a.mu, tau);
        else
            tau_single<ORDER>(g, n, gam, v, a.mu, tau);
        if (j == 0 && seeded) {  // warp-uniform; lane 0 sits on the column before the block's first (col < 0)
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const double t = __shfl_sync(0xffffffffu, prev_seed, i);
                if (lane == 0) tau[i] = t;
            }
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
                for (int j2 = 0; j2 < 3; ++j2) acc[9 + j2] += norm3(tau[3 * j2], tau[3 * j2 + 1], tau[3 * j2 + 2]);
                twssg_project(dw, p);
#pragma unroll
                for (int j2 = 0; j2 < 3; ++j2) acc[12 + j2] += p[j2];
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
        double* p = a.part + (int64_t)y * VH_NSUM * T.n_work + w;
#pragma unroll
        for (int i = 0; i < VH_NSUM; ++i) p[(int64_t)i * T.n_work] = acc[i];
    }
}

// sums[i][f] += part[0][i][w] + part[1][i][w] + ...  for f = work[w], in fixed order; the single-facet and the
// multi-facet launches have their own segment counts and partial-sum blocks
__global__ void k3_fold(double* __restrict__ sums, const double* __restrict__ part_s, int gy_s,
                        const double* __restrict__ part_m, int gy_m, const int32_t* __restrict__ work, int64_t n_work,
                        int64_t multi_start, int64_t nF, int overwrite) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= VH_NSUM * n_work) return;
    const int64_t w = idx % n_work, i = idx / n_work;
    const int32_t f = work[w];
    if (f < 0) return;
    const bool multi = w >= multi_start;
    const double* p = (multi ? part_m : part_s) + i * n_work + w;
    const int gq = multi ? gy_m : gy_s;
    double t = overwrite ? 0.0 : sums[i * nF + f];  // first launch of a time loop: no memset needed
    for (int q = 0; q < gq; ++q) t += p[(int64_t)q * VH_NSUM * n_work];
    sums[i * nF + f] = t;
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

// ---- peer-memory reduction (one process per GPU, memory mapped with CUDA IPC) ------------------------------------------
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_peer(const double* p) {
    // system-scope relaxed load: goes to the owner's memory (peer lines are never in the local L2, and the local L1
    // holds nothing of them at this point of a fresh kernel), and -- unlike volatile -- loads may overlap
    double v;
    asm("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// thread q tells rank q "rank `rank` has finished epoch `epoch`": everything this GPU wrote before (K3's sums) is
// made visible system-wide first
__global__ void k_peer_signal(PeerBlocks pb, int world, int rank, int64_t flags_off, uint64_t epoch, double* count_slot,
                              double count) {
    if (threadIdx.x == 0) *count_slot = count;  // the snapshot count rides behind the sums
    __syncthreads();
    if ((int)threadIdx.x >= world) return;
    __threadfence_system();
    uint64_t* flags = reinterpret_cast<uint64_t*>(const_cast<double*>(pb.block[threadIdx.x]) + flags_off);
    st_release_sys(flags + rank, epoch);
}

// compute_hemodynamics.py:326-346 on sums that are still spread over the GPUs of the node
__global__ void k4_peer_indices(PeerBlocks pb, int world, int rank, int64_t half_off, int64_t flags_off, uint64_t epoch,
                                int64_t nF, double count, double* __restrict__ red, double* __restrict__ tawss,
                                double* __restrict__ osi, double* __restrict__ rrt, double* __restrict__ ecap,
                                double* __restrict__ twssg) {
    if ((int)threadIdx.x < world) {
        const uint64_t* flags = reinterpret_cast<const uint64_t*>(pb.block[rank] + flags_off);  // local memory
        while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
        }
    }
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {  // snapshot count
        double t = 0.0;
        for (int q = 0; q < world; ++q) t += ld_peer(pb.block[q] + half_off + VH_NSUM * nF);
        red[VH_NSUM * nF] = t;
    }
    if (i >= 3 * nF) return;
    const int64_t f = i % nF;
    const int j = (int)(i / nF);
    const int rows[5] = {3 * j, 3 * j + 1, 3 * j + 2, 9 + j, 12 + j};
    double pv[VH_MAX_PEERS][5];  // all peer loads in flight together (NVLink round trip ~ 1-2 us)
#pragma unroll
    for (int q = 0; q < VH_MAX_PEERS; ++q)
#pragma unroll
        for (int r = 0; r < 5; ++r)
            pv[q][r] = q < world ? ld_peer(pb.block[q] + half_off + (int64_t)rows[r] * nF + f) : 0.0;
    double v[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < VH_MAX_PEERS; ++q)
            if (q < world) t += pv[q][r];  // rank order: bitwise identical on every rank
        v[r] = t;
        red[(int64_t)rows[r] * nF + f] = t;
    }
    const double mean_mag = norm3(v[0] / count, v[1] / count, v[2] / count);
    const double ta = v[3] / count;
    const double o = 0.5 * (1.0 - mean_mag / ta);
    const int64_t q = 3 * f + j;
    tawss[q] = ta;
    osi[q] = o;
    rrt[q] = 1.0 / mean_mag;
    ecap[q] = o / ta;
    twssg[q] = v[4] / count;
}

}  // namespace

int k4_peer_signal(vh_handle* h, const PeerBlocks& pb, int64_t flags_off, uint64_t epoch) {
    k_peer_signal<<<1, 32, 0, h->s_compute>>>(pb, h->world, h->rank, flags_off, epoch, h->d_sums + VH_NSUM * h->nF,
                                              (double)h->count);
    VH_CUDA(cudaGetLastError());
    h->launches += 1;
    return VH_OK;
}

int k4_peer_reduce_finalize(vh_handle* h, const PeerBlocks& pb, int64_t half_off, int64_t flags_off, uint64_t epoch,
                            int64_t n_total, double* d_red, double* d_out5) {
    const int64_t nF = h->nF, n3 = 3 * nF;
    k4_peer_indices<<<(unsigned)((n3 + 255) / 256), 256, 0, h->s_compute>>>(
        pb, h->world, h->rank, half_off, flags_off, epoch, nF, (double)n_total, d_red, d_out5, d_out5 + n3,
        d_out5 + 2 * n3, d_out5 + 3 * n3, d_out5 + 4 * n3);
    VH_CUDA(cudaGetLastError());
    h->launches += 1;
    return VH_OK;
}

FacetTables vh_tables(const vh_handle* h) {
    FacetTables T;
    T.nF = h->nF;
    T.row = h->d_row;
    T.glam = h->d_glam;
    T.normal = h->d_normal;
    T.work = h->d_work;
    T.n_work = h->n_work;
    T.multi_start = h->multi_start;
    T.m_mat = h->d_m_mat;
    T.nMulti = h->nMulti;
    return T;
}

int k_free_run_buffers(vh_handle* h) {
    if (h->d_sums_block) cudaFree(h->d_sums_block);
    if (h->d_sums_red) cudaFree(h->d_sums_red);
    h->d_sums_block = h->d_sums_red = nullptr;
    h->peer_ready = false;  // peers must map the new block again (vh_peer_init)
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

namespace {

// Segments of one launch: `total` lane passes split as evenly as possible over gy segments.
struct SegPlan {
    int gy, base, extra;
};

SegPlan plan_segments(int64_t nb, int64_t n_items, int64_t target_warps, int64_t chunk_snapshots) {
    const int64_t total = (nb + K2_COLS - 1) / K2_COLS;
    int64_t gy;
    if (chunk_snapshots > 0) {
        const int64_t p = (chunk_snapshots + K2_COLS - 1) / K2_COLS;
        gy = (total + p - 1) / p;
    } else {
        gy = n_items > 0 ? (target_warps + n_items - 1) / n_items : 1;
    }
    if (gy < 1) gy = 1;
    if (gy > total) gy = total;
    if (gy > 65535) gy = 65535;
    return {(int)gy, (int)(total / gy), (int)(total % gy)};
}

template <int ORDER, bool MULTI>
int launch_k2(const K2Args& a, unsigned gx, unsigned gy, cudaStream_t st) {
    constexpr int smem = k2_smem_bytes<ORDER>();
    static bool configured = false;  // per instantiation
    if (!configured && smem > 48 * 1024) {
        VH_CUDA(cudaFuncSetAttribute(k2_wall<ORDER, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    k2_wall<ORDER, MULTI><<<dim3(gx, gy), 32 * K2_WARPS, smem, st>>>(a);
    VH_CUDA(cudaGetLastError());
    return VH_OK;
}

}  // namespace

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
        // the fewest segments that still give every SM ~64 (facet, segment) warps to schedule; the few multi-facet-cell
        // facets run beside them on a second stream, cut finer so that they never become the tail
        const SegPlan ps = plan_segments(nb, n_single, (int64_t)h->sm_count * 64, h->chunk_snapshots);
        const SegPlan pm = plan_segments(nb, n_multi, (int64_t)h->sm_count * 16, h->chunk_snapshots);
        const int64_t groups = ps.gy + (n_multi ? pm.gy : 0);
        if (groups > h->part_cap) {
            if (h->d_part) cudaFree(h->d_part);
            h->d_part = nullptr;
            h->part_cap = 0;
            VH_CUDA(cudaMalloc(&h->d_part, sizeof(double) * VH_NSUM * h->n_work * groups));
            h->part_cap = groups;
        }
        double* part_s = h->d_part;
        double* part_m = h->d_part + (int64_t)ps.gy * VH_NSUM * h->n_work;
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
        a.prev_mode = pos == 0 ? prev_mode : 1;
        a.tau_last_in = h->d_tau_last[h->tau_cur];
        a.tau_last_out = h->d_tau_last[h->tau_cur ^ 1];
        h->tau_cur ^= 1;
        a.wss_out = d_wss ? d_wss + pos * nF * 9 : nullptr;
        a.mu = h->mu;
        a.inv_dt = 1.0 / h->dt;
        if (gx_multi) {  // fork: multi-facet cells on the auxiliary stream, after K1
            VH_CUDA(cudaEventRecord(h->ev_fork, h->s_compute));
            VH_CUDA(cudaStreamWaitEvent(h->s_aux, h->ev_fork, 0));
            a.part = part_m;
            a.pass_base = pm.base;
            a.pass_extra = pm.extra;
            if (h->order == 2)
                VH_TRY((launch_k2<2, true>(a, gx_multi, pm.gy, h->s_aux)));
            else
                VH_TRY((launch_k2<1, true>(a, gx_multi, pm.gy, h->s_aux)));
            VH_CUDA(cudaEventRecord(h->ev_join, h->s_aux));
            h->launches += 1;
        }
        if (gx_single) {
            a.part = part_s;
            a.pass_base = ps.base;
            a.pass_extra = ps.extra;
            if (h->order == 2)
                VH_TRY((launch_k2<2, false>(a, gx_single, ps.gy, h->s_compute)));
            else
                VH_TRY((launch_k2<1, false>(a, gx_single, ps.gy, h->s_compute)));
            h->launches += 1;
        }
        if (gx_multi) VH_CUDA(cudaStreamWaitEvent(h->s_compute, h->ev_join, 0));
        if (prof) {
            cudaEventRecord(h->prof_pool[h->prof_used + 2], h->s_compute);
            h->prof_used += 3;
        }
        const int64_t n = VH_NSUM * h->n_work;
        k3_fold<<<(unsigned)((n + 255) / 256), 256, 0, h->s_compute>>>(h->d_sums, part_s, gx_single ? ps.gy : 0, part_m,
                                                                       pm.gy, h->d_work, h->n_work, h->multi_start, nF,
                                                                       h->sums_pending_zero ? 1 : 0);
        h->sums_pending_zero = false;
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
