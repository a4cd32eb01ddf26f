// K2/K3: per-snapshot wall traction + fused time reductions; K4: final index formulas.
//
// Replaces the body of the reference's snapshot loop (compute_hemodynamics.py:272-318):
//   u_p2 = T * u_p1                      (:275)  -> folded into the gather slots built by K0
//   tau = stress()                       (:282)  -> closed-form P2/P1 gradient at the facet vertices, sigma n,
//                                                   tangential part; SurfaceProjector's solve is the identity for
//                                                   cells with one exterior facet and a precomputed 3x3-per-contributor
//                                                   weight for cells with several (K0)
//   TAWSS += |tau|, WSS_mean += tau      (:289-306)
//   TWSSG += project_dg(|dtau/dt|)       (:309-312) 7-point degree-5 rule, closed-form P1 mass inverse
// and the final formulas (:326-346).
//
// Thread = one facet x one contiguous chunk of snapshots; tau_prev, the 15 running sums and the facet geometry stay
// in registers for the whole chunk.  A warp covers 32 consecutive work items (facets) so the table loads and the
// partial-sum stores are coalesced; the velocity gather goes through the read-only path.  Chunks other than the
// first recompute tau of the snapshot before them (TWSSG's one-step dependence) instead of communicating.  Partial
// sums of the blockDim.y chunks of a CTA are added in shared memory in fixed order, written to `part`, and folded
// into the running sums by k3_fold in fixed order: results are bitwise reproducible for a given launch shape.
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
    return (a > b) ? edge_dof(b, a)
                   : (b == 1) ? 4 : (b == 2) ? 5 + a : 7 + a;
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

template <int ORDER>
__device__ __forceinline__ void load_vel(const double* __restrict__ u, const int32_t (&slot)[10], int64_t o0,
                                         int64_t o1, int64_t o2, Vel& v) {
#pragma unroll
    for (int k = 0; k < Dofs<ORDER>::N; ++k) {
        v.x[k] = __ldg(u + o0 + slot[k]);
        v.y[k] = __ldg(u + o1 + slot[k]);
        v.z[k] = __ldg(u + o2 + slot[k]);
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

struct K2Args {
    FacetTables T;
    const double* u;        // first non-halo snapshot
    int64_t stride;         // doubles between snapshots
    int64_t n_snap;
    int64_t chunk;          // snapshots per thread
    int prev_mode;          // 0: zero, 1: tau_last, 2: recompute from snapshot -1
    const double* tau_last_in;
    double* tau_last_out;   // [9][nF]
    double* part;           // [gridDim.y][15][nF]
    double* wss_out;        // [n_snap][nF][9] or null
    double mu, inv_dt;
    int64_t off0, off1, off2;
};

template <int ORDER>
__global__ void __launch_bounds__(256) k2_traction(const K2Args a) {
    extern __shared__ double sm[];  // [blockDim.y][15][blockDim.x]
    const FacetTables& T = a.T;
    const int64_t nF = T.nF;
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t cid = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
    const int64_t s0 = cid * a.chunk;
    const int64_t s1 = min(s0 + a.chunk, a.n_snap);
    const int32_t f = (w < T.n_work) ? T.work[w] : -1;
    const bool is_multi = w >= T.multi_start;
    const int64_t m = w - T.multi_start;

    double acc[VH_NSUM];
#pragma unroll
    for (int i = 0; i < VH_NSUM; ++i) acc[i] = 0.0;

    if (f >= 0 && s0 < s1) {
        int32_t slot[10];
        double g[4][3], n[3], gam[4];
#pragma unroll
        for (int k = 0; k < Dofs<ORDER>::N; ++k) slot[k] = T.slot[(int64_t)k * nF + f];
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int d = 0; d < 3; ++d) g[b][d] = T.glam[(int64_t)(3 * b + d) * nF + f];
#pragma unroll
        for (int d = 0; d < 3; ++d) n[d] = T.normal[(int64_t)d * nF + f];
#pragma unroll
        for (int b = 0; b < 4; ++b) gam[b] = g[b][0] * n[0] + g[b][1] * n[1] + g[b][2] * n[2];

        Vel v;
        double prev[9], tau[9];
        if (s0 > 0 || a.prev_mode == 2) {
            load_vel<ORDER>(a.u + (s0 - 1) * a.stride, slot, a.off0, a.off1, a.off2, v);
            if (is_multi)
                tau_multi<ORDER>(g, v, a.mu, T.m_lf, T.m_w, m, T.nMulti, prev);
            else
                tau_single<ORDER>(g, n, gam, v, a.mu, prev);
        } else if (a.prev_mode == 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) prev[i] = a.tau_last_in[(int64_t)i * nF + f];
        } else {
#pragma unroll
            for (int i = 0; i < 9; ++i) prev[i] = 0.0;
        }

        for (int64_t s = s0; s < s1; ++s) {
            load_vel<ORDER>(a.u + s * a.stride, slot, a.off0, a.off1, a.off2, v);
            if (is_multi)
                tau_multi<ORDER>(g, v, a.mu, T.m_lf, T.m_w, m, T.nMulti, tau);
            else
                tau_single<ORDER>(g, n, gam, v, a.mu, tau);
            if (a.wss_out) {
                double* o = a.wss_out + (s * nF + f) * 9;
#pragma unroll
                for (int i = 0; i < 9; ++i) o[i] = tau[i];
            }
            double dw[9], p[3];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                acc[i] += tau[i];
                dw[i] = (tau[i] - prev[i]) * a.inv_dt;
                prev[i] = tau[i];
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[9 + j] += norm3(tau[3 * j], tau[3 * j + 1], tau[3 * j + 2]);
            twssg_project(dw, p);
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[12 + j] += p[j];
        }
        if (s1 == a.n_snap) {
#pragma unroll
            for (int i = 0; i < 9; ++i) a.tau_last_out[(int64_t)i * nF + f] = prev[i];
        }
    }

    // fixed-order reduction over the CTA's chunk rows, then one coalesced store per sum
    const int bx = blockDim.x, by = blockDim.y;
    if (by > 1) {
#pragma unroll
        for (int i = 0; i < VH_NSUM; ++i) sm[((int64_t)threadIdx.y * VH_NSUM + i) * bx + threadIdx.x] = acc[i];
        __syncthreads();
        if (threadIdx.y == 0) {
#pragma unroll
            for (int i = 0; i < VH_NSUM; ++i) {
                double t = acc[i];
                for (int y = 1; y < by; ++y) t += sm[((int64_t)y * VH_NSUM + i) * bx + threadIdx.x];
                acc[i] = t;
            }
        }
    }
    if (threadIdx.y == 0 && f >= 0) {
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
    T.slot = h->d_slot;
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
    h->begun = false;
    return VH_OK;
}

int k2_launch(vh_handle* h, const double* d_u, int64_t n_snap, int64_t stride_elems, int prev_mode, double* d_wss) {
    if (n_snap <= 0) return VH_OK;
    const int64_t nF = h->nF;
    // launch shape: x = 64 work items, y = up to 4 chunk rows; chunk length so that the grid is a few waves of
    // 148 SMs x resident CTAs, but never shorter than 4 snapshots (halo recompute <= 25 %)
    const int bx = 64;
    const int64_t gx = (h->n_work + bx - 1) / bx;
    int64_t chunk = h->chunk_snapshots;
    if (chunk <= 0) {
        const int64_t target_threads = (int64_t)h->sm_count * 512 * 2;
        int64_t want_chunks = (target_threads + gx * bx - 1) / (gx * bx);
        if (want_chunks < 1) want_chunks = 1;
        chunk = (n_snap + want_chunks - 1) / want_chunks;
        if (chunk < 4) chunk = 4;
    }
    if (chunk > n_snap) chunk = n_snap;
    const int64_t n_chunks = (n_snap + chunk - 1) / chunk;
    int by = n_chunks >= 4 ? 4 : (n_chunks >= 2 ? 2 : 1);
    const int64_t gy = (n_chunks + by - 1) / by;
    VH_CHECK(gy <= 65535, VH_ERR_ARG, "k2_launch: too many chunk groups (%lld); raise chunk_snapshots", (long long)gy);
    if (gy > h->part_cap) {
        if (h->d_part) cudaFree(h->d_part);
        h->d_part = nullptr;
        VH_CUDA(cudaMalloc(&h->d_part, sizeof(double) * VH_NSUM * nF * gy));
        h->part_cap = gy;
    }
    K2Args a;
    a.T = vh_tables(h);
    a.u = d_u;
    a.stride = stride_elems;
    a.n_snap = n_snap;
    a.chunk = chunk;
    a.prev_mode = prev_mode;
    a.tau_last_in = h->d_tau_last[h->tau_cur];
    a.tau_last_out = h->d_tau_last[h->tau_cur ^ 1];
    h->tau_cur ^= 1;
    a.part = h->d_part;
    a.wss_out = d_wss;
    a.mu = h->mu;
    a.inv_dt = 1.0 / h->dt;
    a.off0 = h->comp_offset[0];
    a.off1 = h->comp_offset[1];
    a.off2 = h->comp_offset[2];
    dim3 grid((unsigned)gx, (unsigned)gy), block(bx, by);
    size_t smem = by > 1 ? sizeof(double) * VH_NSUM * bx * by : 0;
    const bool prof = h->profile && h->prof_used + 2 <= h->prof_pool.size();
    if (prof) cudaEventRecord(h->prof_pool[h->prof_used], h->s_compute);
    if (h->order == 2)
        k2_traction<2><<<grid, block, smem, h->s_compute>>>(a);
    else
        k2_traction<1><<<grid, block, smem, h->s_compute>>>(a);
    if (prof) {
        cudaEventRecord(h->prof_pool[h->prof_used + 1], h->s_compute);
        h->prof_used += 2;
    }
    VH_CUDA(cudaGetLastError());
    const int64_t n = VH_NSUM * nF;
    k3_fold<<<(unsigned)((n + 255) / 256), 256, 0, h->s_compute>>>(h->d_sums, h->d_part, n, gy);
    VH_CUDA(cudaGetLastError());
    h->launches += 2;
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
