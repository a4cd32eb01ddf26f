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
// Work decomposition: one WARP per (facet, time segment); the LANES are the snapshots of a time tile of the staged
// block W that K1 wrote (W[node][tile][component][32], 768 contiguous bytes per node and tile).  Every load of a cell
// dof is a run of consecutive, 256-byte-aligned doubles; the three components sit at immediate offsets.  Three code
// paths share one launch (see the table before k2_body_p2); DESIGN.md section 3 has what bounds them (fp64 issue, not
// HBM: W is L2-resident after K1 and each row is shared by ~5 facets) and the variants that were measured and dropped.
//
// TWSSG's one-step dependence: inside a pass tau of the previous snapshot comes from the neighbouring lane (shuffle),
// across passes lane 0 keeps the last lane's tau, and across SEGMENTS nothing is communicated: each segment records tau
// of its first and last column and k3_fold adds the missing boundary terms P(|tau_first(s) - tau_last(s-1)| / dt) when
// it folds the segments' partial sums -- in fixed order, so results are bitwise reproducible for a given launch shape.
//
// Local vertex labels are facet-canonical (K0): 0,1,2 = the facet's vertices in boundary-cell order, 3 = the
// opposite vertex; P2 edge dofs 4..9 = e01,e02,e12,e03,e13,e23.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

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

constexpr int K2_WARPS = 4;
// Default launch shape (K2Launch).  K2 by itself is indifferent to the shape (it is bound by fp64
// operand bandwidth, not by occupancy), but the whole step is not: with 128 registers the blocks of the next kernel in
// the programmatic-dependent-launch chain find room earlier.  Measured (profiles/r2_k2_occupancy.md): step 2.95 -> 2.63 ms
// on the 5 M-tet P2 mesh, 0.724 -> 0.695 ms on the 2 M-tet P1 mesh, 249 -> 243 us on the small P2 mesh, but 55.2 -> 57.7 us
// on the small P1 mesh (its K2 loses the register double buffer).
constexpr int K2_OCC_P2 = 4, K2_OCC_P1_LARGE = 4, K2_OCC_P1_SMALL = 3;
constexpr int64_t K2_P1_LARGE_FACETS = 16384;

// Three code paths share one launch (k2_wall<ORDER>):
//   k2_body_flat2     P1 data, cell with one exterior facet: values by ld.global.nc straight into registers (register
//                     double buffer), two snapshots per lane, closed-form traction, constants in registers
//   k2_body_p2        P2 data, cell with one exterior facet: values by cp.async into a per-warp shared-memory landing
//                     buffer of two tiles (30 doubles per lane would not fit twice in registers; each lane copies and
//                     later reads only its own column: no barrier), closed-form traction, constants in registers
//   k2_body_multi2    cells with several exterior facets: dense operator (K0) staged once per warp in shared memory,
//                     two snapshots per lane so that every operator entry read feeds two evaluations
// Alternatives that were measured and dropped: per-facet constants or a 3 x 12 traction operator
// in shared memory to reach 32 warps per SM -- every warp-uniform LDS still costs its data-path cycles, the kernel
// became MIO-bound; a single-stage landing buffer with 20 warps per SM (spills); L1 prefetch two tiles ahead (no
// effect: W is L2-resident after K1); a TMA bulk copy per row (bypasses L1, where neighbouring facets share rows).
// Shared memory per warp, in doubles.
constexpr int K2_P2_STAGE = 30 * 32;                 // one tile of the 30 dof components
// Launch shapes (template parameter OCC = resident blocks per SM the kernel is compiled for):
//   OCC 3  168 registers, 12 warps per SM; P2 landing buffer of two tiles, P1 register double buffer      (round 1)
//   OCC 4  128 registers, 16 warps per SM; P2 landing buffer of ONE tile that is refilled as soon as the traction of
//          the current tile has been formed (the time reductions -- 45 % of a pass -- cover the copy); P1 loads its
//          pass at the top of the loop and leaves the latency to the other warps
template <int ORDER, int OCC>
struct K2Launch {
    static constexpr int NV = ORDER == 2 ? 30 : 12;
    static constexpr bool DIRECT = OCC == 6;      // P2 values by direct loads (k2_body_p2_direct), 3 blocks per SM
    static constexpr int BLOCKS = DIRECT ? 3 : OCC;
    static constexpr int STAGES = DIRECT ? 0 : OCC >= 4 ? 1 : 2;
    static constexpr int P2_WARP = STAGES * K2_P2_STAGE + 10;  // landing stages + tau of the last lane of the previous pass
    static constexpr int MULTI_WARP = NV * VH_MROW + 10;  // operator + carry
    static constexpr int WARP_DOUBLES = ORDER == 2 ? (P2_WARP > MULTI_WARP ? P2_WARP : MULTI_WARP) : MULTI_WARP;
    static constexpr int SMEM_BYTES = K2_WARPS * WARP_DOUBLES * (int)sizeof(double);
};

// sqrt(x) for x >= 0 without the library's special-case branch: rsqrt seed (MUFU.RSQ64H, ~2^-22), two coupled
// Newton steps on g ~ sqrt(x), h ~ 1 / (2 sqrt(x)) -> error of a few ulp.  Zero and anything below 2^-1020 give 0.
// Branch-free, so that the independent norms of a pass (two for P1, ten for P2) interleave instead of forming one
// serial chain of ten dependent fp64 operations each.
__device__ __forceinline__ double sqrt_nb(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    if (__double2hiint(x) < 0x00300000) y = 0.0;
    double g = x * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    return fma(g, r, g);
}

__device__ __forceinline__ double norm3(double a, double b, double c) { return sqrt_nb(fma(a, a, fma(b, b, c * c))); }

// Per-facet constants of the closed-form traction, as the kernels hold them in registers (19 doubles):
//   C(3 b + d) = hh_b[d] = g_b[d] - 2 gamma_b n[d],   C(12 + d) = n[d],   C(15 + b) = gamma_b = g_b . n
// (g_b = grad lambda_b, n = outward unit normal).  They fold the tangential projection P = I - n n^T into the operator:
// a dyad u (x) g_b inside grad u contributes  P [(u (x) g_b) + (u (x) g_b)^T] n = gamma_b u + (u . n) hh_b  to the
// tangential part of (grad u + grad u^T) n, so F . n is never formed and nothing is projected afterwards
// (9 fp64 instructions fewer per P1 unit, 33 per P2 unit, than forming F and projecting it).
__device__ __forceinline__ void traction_constants(const FacetTables& T, int32_t f, double (&gr)[19]) {
    const int64_t nF = T.nF;
#pragma unroll
    for (int i = 0; i < 12; ++i) gr[i] = T.glam[(int64_t)i * nF + f];
#pragma unroll
    for (int d = 0; d < 3; ++d) gr[12 + d] = T.normal[(int64_t)d * nF + f];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double gam = gr[3 * b] * gr[12] + gr[3 * b + 1] * gr[13] + gr[3 * b + 2] * gr[14];
        gr[15 + b] = gam;
#pragma unroll
        for (int d = 0; d < 3; ++d) gr[3 * b + d] = fma(-2.0 * gam, gr[12 + d], gr[3 * b + d]);
    }
}

// Tangential traction Ft = F - (F.n) n, F = -mu (grad u + grad u^T) n, of P1 data (grad u = sum_b u_b (x) g_b, constant
// in the cell):  Ft = -mu sum_b [gamma_b u_b + (u_b . n) hh_b].  C(i) as above, U(3 b + d) = component d of the velocity
// at vertex b.
template <class GF, class UF>
__device__ __forceinline__ void tau_p1(GF C, UF U, double mu, double (&ft)[3]) {
    const double n0 = C(12), n1 = C(13), n2 = C(14);
    double s[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const double un = fma(U(3 * b + 2), n2, fma(U(3 * b + 1), n1, U(3 * b) * n0));
#pragma unroll
        for (int d = 0; d < 3; ++d) s[d] = fma(U(3 * b + d), C(15 + b), fma(C(3 * b + d), un, s[d]));
    }
    ft[0] = -mu * s[0];
    ft[1] = -mu * s[1];
    ft[2] = -mu * s[2];
}

// P2 data, cell with one exterior facet: tau[3 V + d] at the facet vertices V = 0, 1, 2.
//   grad u (v_a) = H + 4 (u_a (x) g_a + sum_{b != a} u_ab (x) g_b),  H = -sum_b u_b (x) g_b
// With the projected operator above: tau(V) = mu B - 4 mu S_V,  B = sum_b [gamma_b u_b + (u_b . n) hh_b] over the four
// vertex dofs, S_V the same sum over (u_V, g_V) and the three edge dofs at V paired with the far vertex's constants.
// The constants C(i) and the dof values U(3 k + d) are fetched where they are needed (from registers or shared
// memory), so that little more than the ten u_k . n stays live.
template <class GF, class UF>
__device__ __forceinline__ void tau_p2(GF C, UF U, double mu, double (&tau)[9]) {
    const double n0 = C(12), n1 = C(13), n2 = C(14);
    double un[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) un[k] = fma(U(3 * k + 2), n2, fma(U(3 * k + 1), n1, U(3 * k) * n0));
    double mb[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
        for (int d = 0; d < 3; ++d) mb[d] = fma(U(3 * b + d), C(15 + b), fma(C(3 * b + d), un[b], mb[d]));
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) mb[d] *= mu;
    const double m4 = -4.0 * mu;
#pragma unroll
    for (int V = 0; V < 3; ++V) {
        double sv[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) sv[d] = fma(U(3 * V + d), C(15 + V), C(3 * V + d) * un[V]);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b == V) continue;
            const int k = edge_dof(V, b);
#pragma unroll
            for (int d = 0; d < 3; ++d) sv[d] = fma(U(3 * k + d), C(15 + b), fma(C(3 * b + d), un[k], sv[d]));
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) tau[3 * V + d] = fma(sv[d], m4, mb[d]);
    }
}

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

// One launch covers both kinds of work items: blocks [0, gx_multi * gy_multi) take the facets of multi-facet cells
// (the long-running ones, scheduled first), the rest the single-facet-cell facets.  Each kind has its own split of
// the block's tiles into segments and its own partial-sum / boundary buffers.
struct K2Seg {
    int gx, gy;             // blocks along the work list, segments
    int tile_base, tile_extra;  // segment y owns tile_base (+1 if y < tile_extra) consecutive tiles
    double* part;           // [gy][15][n_work] partial sums of the segments
    double* bnd;            // [gy][2][9][n_work] tau of every segment's first and last column
};
struct K2Args {
    FacetTables T;
    const double* W;        // staged block (K1): W[((node * ntile_ld + tile) * 3 + c) * 32 + column in tile]
    int ntile_ld;           // tiles allocated per wall node
    int ncol;               // columns in the block
    int r0;                 // first real column (1 when column 0 is a halo snapshot that only seeds tau_prev)
    int prev_mode;          // tau_prev of the first real column when r0 == 0: 0 zero, 1 tau_last_in
    K2Seg multi, single;
    const double* tau_last_in;
    double* tau_last_out;   // [9][nF]
    double* wss_out;        // tau of every real column, or null: entry (column s, facet f, dof-component i) at
    int64_t ws_col, ws_f, ws_i;  // wss_out[s * ws_col + f * ws_f + i * ws_i]: {9 nF, 9, 1} = one dolfin vector per
                                 // snapshot (WSS.h5), {1, 9 ld, ld} = (dof x time) matrix with lanes along a row
    double mu, inv_dt;
};

__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Time reductions of one column of a facet with three distinct vertex tractions: sum tau, sum |tau|, sum P(|dtau|).
__device__ __forceinline__ void reduce9(const double (&tau)[9], const double (&dw)[9], bool live, bool tw_live,
                                        double (&acc)[VH_NSUM]) {
    double m[3], p[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) m[j] = norm3(tau[3 * j], tau[3 * j + 1], tau[3 * j + 2]);
    twssg_project(dw, p);
    if (live) {
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] += tau[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[9 + j] += m[j];
    }
    if (tw_live) {
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[12 + j] += p[j];
    }
}

// fixed-order butterfly over the 32 lanes, then lane 0 stores the segment's partial sums (TWSSG rows scaled by 1 / dt)
__device__ __forceinline__ void store_partials(double (&acc)[VH_NSUM], double* part, int64_t n_work, double inv_dt, int lane) {
#pragma unroll
    for (int i = 0; i < VH_NSUM; ++i) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], d);
    }
    if (lane == 0) {
        const double s = fabs(inv_dt);
#pragma unroll
        for (int i = 0; i < VH_NSUM; ++i) part[(int64_t)i * n_work] = i < 12 ? acc[i] : acc[i] * s;
    }
}

// compute_hemodynamics.py:326-346 for one boundary dof: v = {sum tau_x, sum tau_y, sum tau_z, sum |tau|, sum P(|dtau/dt|)}
__device__ __forceinline__ void hemo_indices(const double (&v)[5], double count, double (&out)[5]) {
    const double mean_mag = norm3(v[0] / count, v[1] / count, v[2] / count);
    const double ta = v[3] / count;
    const double o = 0.5 * (1.0 - mean_mag / ta);
    out[0] = ta;
    out[1] = o;
    out[2] = 1.0 / mean_mag;
    out[3] = o / ta;
    out[4] = v[4] / count;
}

// One warp = one facet x one time segment; the 32 lanes are the 32 columns of a tile of the staged block.
// P2 data, facets whose cell owns no other exterior facet (work[0, multi_start)).
template <int OCC>
__device__ __forceinline__ void k2_body_p2(const K2Args& a, const K2Seg& sg, int bx, int y, double* k2_smem) {
    using L = K2Launch<2, OCC>;
    const FacetTables& T = a.T;
    const int64_t nF = T.nF;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)bx * K2_WARPS + wib;
    if (w >= T.multi_start) return;
    const int32_t f = T.work[w];
    if (f < 0) return;  // padding entry (warp-uniform)
    const int t0 = y * sg.tile_base + min(y, sg.tile_extra);
    const int nt = sg.tile_base + (y < sg.tile_extra ? 1 : 0);

    double* const ring = k2_smem + (size_t)wib * L::WARP_DOUBLES;
    double* const carry_s = ring + L::STAGES * K2_P2_STAGE;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring + lane);

    // element offset inside W of (cell dof k, tile t0, component 0, this lane's column); W holds < 2^31 doubles
    int32_t rb[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) rb[k] = (T.row[(int64_t)k * nF + f] * a.ntile_ld + t0) * 96 + lane;
    auto fetch = [&](int j) {  // tile t0 + j -> landing stage j % STAGES
        const uint32_t dst = ring_s + (uint32_t)((j % L::STAGES) * K2_P2_STAGE * sizeof(double));
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            const double* p = a.W + (rb[k] + j * 96);
#pragma unroll
            for (int d = 0; d < 3; ++d) cp_async8(dst + (3 * k + d) * 256, p + 32 * d);
        }
        cp_async_commit();
    };
    fetch(0);

    // per-facet constants in registers (traction_constants): {hh [4][3], n [3], gamma [4]}
    double gr[19];
    traction_constants(T, f, gr);
    // tau_prev of the segment's first column.  Segment 0: zero, or carried over from the last launch; later
    // segments: unknown here -- lane 0 skips that one TWSSG term and k3_fold adds it from the boundary records.
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) carry_s[i] = (y == 0 && a.prev_mode == 1) ? a.tau_last_in[(int64_t)i * nF + f] : 0.0;
    }
    __syncwarp();

    double acc[VH_NSUM];
#pragma unroll
    for (int i = 0; i < VH_NSUM; ++i) acc[i] = 0.0;

    for (int j = 0; j < nt; ++j) {
        if (L::STAGES == 2 && j + 1 < nt) {
            fetch(j + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        const double* sv = ring + (j % L::STAGES) * K2_P2_STAGE + lane;  // sv[(3 k + d) * 32]
        double tau[9], dw[9];
        tau_p2([&](int i) { return gr[i]; }, [&](int q) { return sv[q * 32]; }, a.mu, tau);
        // one landing stage: every value of this tile has been read (each lane reads and refills only its own column,
        // and the asm memory clobber keeps the reads above the copy), so the next tile can start to arrive now
        if (L::STAGES == 1 && j + 1 < nt) fetch(j + 1);
        const int col = (t0 + j) * 32 + lane;
        const bool live = col >= a.r0 && col < a.ncol;
        // w = tau - tau_prev (the 1 / dt is applied to the sums at the end: |.| and P are homogeneous)
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const double r = __shfl_sync(0xffffffffu, tau[i], (lane + 31) & 31);
            double prev = r;
            if (lane == 0) {
                prev = carry_s[i];
                carry_s[i] = r;  // tau of lane 31: the predecessor of the next pass's first column
            }
            dw[i] = tau[i] - prev;
        }
        reduce9(tau, dw, live, live && !(j == 0 && lane == 0 && y > 0), acc);
        if (live && a.wss_out) {
            double* o = a.wss_out + (int64_t)(col - a.r0) * a.ws_col + f * a.ws_f;
#pragma unroll
            for (int i = 0; i < 9; ++i) o[i * a.ws_i] = tau[i];
        }
        if (col == a.ncol - 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) a.tau_last_out[(int64_t)i * nF + f] = tau[i];
        }
        // boundary records for k3_fold: first column (segments after the first), last column
        if ((j == 0 && lane == 0 && y > 0) || (j == nt - 1 && lane == 31)) {
            double* b = sg.bnd + ((int64_t)(2 * y + (lane == 0 ? 0 : 1)) * 9) * T.n_work + w;
#pragma unroll
            for (int i = 0; i < 9; ++i) b[(int64_t)i * T.n_work] = tau[i];
        }
    }
    store_partials(acc, sg.part + (int64_t)y * VH_NSUM * T.n_work + w, T.n_work, a.inv_dt, lane);
}

// P2 data, cell with one exterior facet, values straight from global memory into registers (ld.global.nc.f64, no
// shared-memory landing buffer): 30 loads per pass instead of 30 cp.async + 30 LDS, i.e. a third of the wavefronts on
// the L1/shared-memory data path.  No prefetch across passes (60 registers per tile would not fit twice): the other
// resident warps cover the latency.
__device__ __forceinline__ void k2_body_p2_direct(const K2Args& a, const K2Seg& sg, int bx, int y, double* k2_smem,
                                                  int warp_doubles) {
    const FacetTables& T = a.T;
    const int64_t nF = T.nF;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)bx * K2_WARPS + wib;
    if (w >= T.multi_start) return;
    const int32_t f = T.work[w];
    if (f < 0) return;  // padding entry (warp-uniform)
    const int t0 = y * sg.tile_base + min(y, sg.tile_extra);
    const int nt = sg.tile_base + (y < sg.tile_extra ? 1 : 0);
    double* const carry_s = k2_smem + (size_t)wib * warp_doubles;  // tau of the last lane of the previous pass

    int32_t rb[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) rb[k] = (T.row[(int64_t)k * nF + f] * a.ntile_ld + t0) * 96 + lane;
    double gr[19];
    traction_constants(T, f, gr);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) carry_s[i] = (y == 0 && a.prev_mode == 1) ? a.tau_last_in[(int64_t)i * nF + f] : 0.0;
    }
    __syncwarp();

    double acc[VH_NSUM];
#pragma unroll
    for (int i = 0; i < VH_NSUM; ++i) acc[i] = 0.0;

    for (int j = 0; j < nt; ++j) {
        double u[30];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            const double* p = a.W + (rb[k] + j * 96);
#pragma unroll
            for (int d = 0; d < 3; ++d) u[3 * k + d] = __ldg(p + 32 * d);
        }
        double tau[9], dw[9];
        tau_p2([&](int i) { return gr[i]; }, [&](int q) { return u[q]; }, a.mu, tau);
        const int col = (t0 + j) * 32 + lane;
        const bool live = col >= a.r0 && col < a.ncol;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const double r = __shfl_sync(0xffffffffu, tau[i], (lane + 31) & 31);
            double prev = r;
            if (lane == 0) {
                prev = carry_s[i];
                carry_s[i] = r;
            }
            dw[i] = tau[i] - prev;
        }
        reduce9(tau, dw, live, live && !(j == 0 && lane == 0 && y > 0), acc);
        if (live && a.wss_out) {
            double* o = a.wss_out + (int64_t)(col - a.r0) * a.ws_col + f * a.ws_f;
#pragma unroll
            for (int i = 0; i < 9; ++i) o[i * a.ws_i] = tau[i];
        }
        if (col == a.ncol - 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) a.tau_last_out[(int64_t)i * nF + f] = tau[i];
        }
        if ((j == 0 && lane == 0 && y > 0) || (j == nt - 1 && lane == 31)) {
            double* b = sg.bnd + ((int64_t)(2 * y + (lane == 0 ? 0 : 1)) * 9) * T.n_work + w;
#pragma unroll
            for (int i = 0; i < 9; ++i) b[(int64_t)i * T.n_work] = tau[i];
        }
    }
    store_partials(acc, sg.part + (int64_t)y * VH_NSUM * T.n_work + w, T.n_work, a.inv_dt, lane);
}

// Facets whose cell owns several exterior facets (work[multi_start, n_work)): tau = -mu M^T u with the dense operator
// K0 built (SurfaceProjector's block solve folded with the contributing faces).  The operator is staged once per warp
// in shared memory, already scaled by -mu; a pass covers 64 columns, lane l owns columns 2 l and 2 l + 1, so every
// warp-uniform operator read feeds two evaluations.
template <int ORDER, int OCC>
__device__ __forceinline__ void k2_body_multi2(const K2Args& a, const K2Seg& sg, int bx, int y, double* k2_smem) {
    constexpr int N = ORDER == 2 ? 10 : 4, NV = 3 * N;
    const FacetTables& T = a.T;
    const int64_t nF = T.nF;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t wi = (int64_t)bx * K2_WARPS + wib;
    const int64_t w = wi + T.multi_start;
    if (w >= T.n_work) return;
    const int32_t f = T.work[w];
    if (f < 0) return;  // padding entry (warp-uniform)
    const int p0 = y * sg.tile_base + min(y, sg.tile_extra);  // passes of 64 columns
    const int np = sg.tile_base + (y < sg.tile_extra ? 1 : 0);

    double* const Ms = k2_smem + (size_t)wib * K2Launch<ORDER, OCC>::WARP_DOUBLES;  // [NV][VH_MROW]
    double* const carry_s = Ms + NV * VH_MROW;
    {
        const double2* Mg = reinterpret_cast<const double2*>(T.m_mat + (size_t)wi * NV * VH_MROW);
        const double s = -a.mu;
        for (int i = lane; i < NV * VH_MROW / 2; i += 32) {
            double2 m = __ldg(Mg + i);
            m.x *= s;
            m.y *= s;
            reinterpret_cast<double2*>(Ms)[i] = m;
        }
        if (lane < 9) carry_s[lane] = (y == 0 && a.prev_mode == 1) ? a.tau_last_in[(int64_t)lane * nF + f] : 0.0;
    }
    __syncwarp();
    // element offset inside W of (cell dof k, tile 2 p0 + lane / 16, component 0, column 2 (lane % 16))
    int32_t rb[N];
#pragma unroll
    for (int k = 0; k < N; ++k)
        rb[k] = (T.row[(int64_t)k * nF + f] * a.ntile_ld + 2 * p0 + (lane >> 4)) * 96 + 2 * (lane & 15);

    double acc[VH_NSUM];
#pragma unroll
    for (int i = 0; i < VH_NSUM; ++i) acc[i] = 0.0;

    for (int j = 0; j < np; ++j) {
        double t0[9], t1[9], d0[9], d1[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) t0[i] = t1[i] = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double2* p = reinterpret_cast<const double2*>(a.W + (rb[k] + j * 192));
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double2 u = __ldg(p + 16 * d);
                const double2* row = reinterpret_cast<const double2*>(Ms + (3 * k + d) * VH_MROW);
#pragma unroll
                for (int h2 = 0; h2 < 5; ++h2) {
                    const double2 m2 = row[h2];
                    t0[2 * h2] = fma(m2.x, u.x, t0[2 * h2]);
                    t1[2 * h2] = fma(m2.x, u.y, t1[2 * h2]);
                    if (h2 < 4) {
                        t0[2 * h2 + 1] = fma(m2.y, u.x, t0[2 * h2 + 1]);
                        t1[2 * h2 + 1] = fma(m2.y, u.y, t1[2 * h2 + 1]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const double r = __shfl_sync(0xffffffffu, t1[i], (lane + 31) & 31);
            double prev = r;
            if (lane == 0) {
                prev = carry_s[i];
                carry_s[i] = r;  // tau of the last column of this pass
            }
            d0[i] = t0[i] - prev;
            d1[i] = t1[i] - t0[i];
        }
        const int col = 64 * (p0 + j) + 2 * lane;
        const bool live0 = col >= a.r0 && col < a.ncol, live1 = col + 1 >= a.r0 && col + 1 < a.ncol;
        reduce9(t0, d0, live0, live0 && !(j == 0 && lane == 0 && y > 0), acc);
        reduce9(t1, d1, live1, live1, acc);
        if (a.wss_out) {
            if (live0) {
                double* o = a.wss_out + (int64_t)(col - a.r0) * a.ws_col + f * a.ws_f;
#pragma unroll
                for (int i = 0; i < 9; ++i) o[i * a.ws_i] = t0[i];
            }
            if (live1) {
                double* o = a.wss_out + (int64_t)(col + 1 - a.r0) * a.ws_col + f * a.ws_f;
#pragma unroll
                for (int i = 0; i < 9; ++i) o[i * a.ws_i] = t1[i];
            }
        }
        if (col == a.ncol - 1 || col + 1 == a.ncol - 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) a.tau_last_out[(int64_t)i * nF + f] = col == a.ncol - 1 ? t0[i] : t1[i];
        }
        // boundary records for k3_fold: first column (segments after the first), last column
        if (j == 0 && lane == 0 && y > 0) {
            double* b = sg.bnd + ((int64_t)(2 * y) * 9) * T.n_work + w;
#pragma unroll
            for (int i = 0; i < 9; ++i) b[(int64_t)i * T.n_work] = t0[i];
        }
        if (j == np - 1 && lane == 31) {
            double* b = sg.bnd + ((int64_t)(2 * y + 1) * 9) * T.n_work + w;
#pragma unroll
            for (int i = 0; i < 9; ++i) b[(int64_t)i * T.n_work] = t1[i];
        }
    }
    store_partials(acc, sg.part + (int64_t)y * VH_NSUM * T.n_work + w, T.n_work, a.inv_dt, lane);
}

// P1 data in a cell with one exterior facet, TWO consecutive snapshots per lane: a pass covers two tiles (64 columns),
// lane l owns columns 2 l and 2 l + 1 (one 16-byte load per dof component; the lower half-warp reads the first tile,
// the upper one the second).  The per-facet constants are shared by two independent evaluations -- twice the
// instruction-level parallelism per register -- tau_prev of the second column is the lane's own first column, and the
// shuffles and the integer work per unit halve.  (The fp64 pipe of an SM retires 2 warp instructions per clock at 8
// clocks latency, measured with tools/ubench/fp64_lat.cu; with 4 warps per scheduler and one column per lane the
// kernel issued 0.57 instructions per clock.)
template <bool PREFETCH>
__device__ __forceinline__ void k2_body_flat2(const K2Args& a, const K2Seg& sg, int bx, int y) {
    const FacetTables& T = a.T;
    const int64_t nF = T.nF;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)bx * K2_WARPS + wib;
    if (w >= T.multi_start) return;
    const int32_t f = T.work[w];
    if (f < 0) return;  // padding entry (warp-uniform)
    const int p0 = y * sg.tile_base + min(y, sg.tile_extra);  // passes of 64 columns
    const int np = sg.tile_base + (y < sg.tile_extra ? 1 : 0);

    // element offset inside W of (cell dof k, tile 2 p0 + lane / 16, component 0, column 2 (lane % 16))
    int32_t rb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        rb[k] = (T.row[(int64_t)k * nF + f] * a.ntile_ld + 2 * p0 + (lane >> 4)) * 96 + 2 * (lane & 15);
    double2 v[12], vn[12];  // register double buffer: the next pass's loads are issued before this pass's arithmetic
    auto fetch = [&](int j, double2* dst) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double2* p = reinterpret_cast<const double2*>(a.W + (rb[k] + j * 192));
#pragma unroll
            for (int d = 0; d < 3; ++d) dst[3 * k + d] = __ldg(p + 16 * d);
        }
    };
    if (PREFETCH) fetch(0, vn);

    double gr[19];
    traction_constants(T, f, gr);
    auto G = [&](int i) { return gr[i]; };

    // tau_prev of the segment's first column (see k2_body)
    double carry[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        carry[i] = (y == 0 && a.prev_mode == 1 && lane == 0) ? a.tau_last_in[(int64_t)i * nF + f] : 0.0;

    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int j = 0; j < np; ++j) {
        if (PREFETCH) {
#pragma unroll
            for (int q = 0; q < 12; ++q) v[q] = vn[q];
            if (j + 1 < np) fetch(j + 1, vn);
        } else {
            fetch(j, v);
        }
        double t0[3], t1[3], d0[3], d1[3];
        tau_p1(G, [&](int q) { return v[q].x; }, a.mu, t0);
        tau_p1(G, [&](int q) { return v[q].y; }, a.mu, t1);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double r = __shfl_sync(0xffffffffu, t1[i], (lane + 31) & 31);
            d0[i] = t0[i] - (lane == 0 ? carry[i] : r);
            carry[i] = r;  // lane 0: tau of the last column of this pass
            d1[i] = t1[i] - t0[i];
        }
        const double m0 = norm3(t0[0], t0[1], t0[2]), m1 = norm3(t1[0], t1[1], t1[2]);
        const double w0 = norm3(d0[0], d0[1], d0[2]), w1 = norm3(d1[0], d1[1], d1[2]);
        const int col = 64 * (p0 + j) + 2 * lane;
        const bool live0 = col >= a.r0 && col < a.ncol, live1 = col + 1 >= a.r0 && col + 1 < a.ncol;
        if (live0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) acc[i] += t0[i];
            acc[3] += m0;
        }
        if (live0 && !(j == 0 && lane == 0 && y > 0)) acc[4] += w0;
        if (live1) {
#pragma unroll
            for (int i = 0; i < 3; ++i) acc[i] += t1[i];
            acc[3] += m1;
            acc[4] += w1;
        }
        if (a.wss_out) {
            if (live0) {
                double* o = a.wss_out + (int64_t)(col - a.r0) * a.ws_col + f * a.ws_f;
#pragma unroll
                for (int i = 0; i < 9; ++i) o[i * a.ws_i] = t0[i % 3];
            }
            if (live1) {
                double* o = a.wss_out + (int64_t)(col + 1 - a.r0) * a.ws_col + f * a.ws_f;
#pragma unroll
                for (int i = 0; i < 9; ++i) o[i * a.ws_i] = t1[i % 3];
            }
        }
        if (col == a.ncol - 1 || col + 1 == a.ncol - 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) a.tau_last_out[(int64_t)i * nF + f] = (col == a.ncol - 1 ? t0 : t1)[i % 3];
        }
        // boundary records for k3_fold: first column (segments after the first), last column
        if (j == 0 && lane == 0 && y > 0) {
            double* b = sg.bnd + ((int64_t)(2 * y) * 9) * T.n_work + w;
#pragma unroll
            for (int i = 0; i < 9; ++i) b[(int64_t)i * T.n_work] = t0[i % 3];
        }
        if (j == np - 1 && lane == 31) {
            double* b = sg.bnd + ((int64_t)(2 * y + 1) * 9) * T.n_work + w;
#pragma unroll
            for (int i = 0; i < 9; ++i) b[(int64_t)i * T.n_work] = t1[i % 3];
        }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], d);
    }
    if (lane == 0) {
        const double s = fabs(a.inv_dt);
        double* p = sg.part + (int64_t)y * VH_NSUM * T.n_work + w;
#pragma unroll
        for (int i = 0; i < VH_NSUM; ++i) p[(int64_t)i * T.n_work] = i < 9 ? acc[i % 3] : i < 12 ? acc[3] : acc[4] * s;
    }}

template <int ORDER, int OCC>
__global__ void __launch_bounds__(32 * K2_WARPS, (K2Launch<ORDER, OCC>::BLOCKS)) k2_wall(const K2Args a) {
    extern __shared__ __align__(16) double k2_smem[];
    pdl_wait();  // W comes from the K1 just before; part/bnd were read by the K3 before that
    pdl_launch_dependents();
    const int nb_multi = a.multi.gx * a.multi.gy;
    if ((int)blockIdx.x < nb_multi) {
        k2_body_multi2<ORDER, OCC>(a, a.multi, blockIdx.x % a.multi.gx, blockIdx.x / a.multi.gx, k2_smem);
    } else {
        const int b = blockIdx.x - nb_multi;
        if constexpr (ORDER == 1)
            k2_body_flat2<(OCC < 4 || OCC == 6)>(a, a.single, b % a.single.gx, b / a.single.gx);
        else if constexpr (K2Launch<ORDER, OCC>::DIRECT)
            k2_body_p2_direct(a, a.single, b % a.single.gx, b / a.single.gx, k2_smem, K2Launch<ORDER, OCC>::WARP_DOUBLES);
        else
            k2_body_p2<OCC>(a, a.single, b % a.single.gx, b / a.single.gx, k2_smem);
    }
}

// sums[i][f] (+)= sum over the segments of part[q][i][w], f = work[w], in fixed order, plus the TWSSG terms of the
// segment boundaries, P(|tau_first(s) - tau_last(s - 1)| / dt), which no segment could form on its own; then the final
// formulas (:326-346) for the snapshots seen so far -- vh_finalize reuses them when the count matches, so a step on the
// hot path is three launches (K1, K2, K3).  A block takes 16 work items: thread (row i, item) folds one sum row
// (16 consecutive items per row: 128-byte runs), the rows meet in shared memory, 48 threads evaluate the indices.
// The single-facet and the multi-facet paths have their own segment counts and buffers.
constexpr int K3_ITEMS = 16;
__global__ void __launch_bounds__(16 * K3_ITEMS)
    k3_fold(double* __restrict__ sums, const double* __restrict__ part_s, const double* __restrict__ bnd_s, int gy_s,
            const double* __restrict__ part_m, const double* __restrict__ bnd_m, int gy_m,
            const int32_t* __restrict__ work, int64_t n_work, int64_t multi_start, int64_t nF, double inv_dt,
            int overwrite, double count, double* __restrict__ out5) {
    __shared__ double rows[VH_NSUM][K3_ITEMS];
    const int wl = threadIdx.x % K3_ITEMS, i = threadIdx.x / K3_ITEMS;  // i = 15: idle row
    const int64_t w = (int64_t)blockIdx.x * K3_ITEMS + wl;
    const int32_t f = w < n_work ? work[w] : -1;  // K0's table: older than any kernel in flight
    pdl_wait();
    pdl_launch_dependents();
    if (f >= 0 && i < VH_NSUM) {
        const bool multi = w >= multi_start;
        const double* part = (multi ? part_m : part_s) + (int64_t)i * n_work + w;
        const double* bnd = (multi ? bnd_m : bnd_s) + w;
        const int gq = multi ? gy_m : gy_s;
        double t = overwrite ? 0.0 : sums[(int64_t)i * nF + f];  // first launch of a time loop: no memset needed
        for (int q = 0; q < gq; ++q) {
            t += part[(int64_t)q * VH_NSUM * n_work];
            if (i >= 12 && q > 0) {  // the three TWSSG rows each evaluate the boundary projection and keep their entry
                double dw[9], p[3];
#pragma unroll
                for (int c = 0; c < 9; ++c)
                    dw[c] = bnd[((int64_t)(2 * q) * 9 + c) * n_work] - bnd[((int64_t)(2 * q - 1) * 9 + c) * n_work];
                twssg_project(dw, p);
                t += (i == 12 ? p[0] : i == 13 ? p[1] : p[2]) * fabs(inv_dt);
            }
        }
        sums[(int64_t)i * nF + f] = t;
        rows[i][wl] = t;
    }
    __syncthreads();
    if (f >= 0 && i < 3) {  // boundary dof j = i of facet f
        const double v[5] = {rows[3 * i][wl], rows[3 * i + 1][wl], rows[3 * i + 2][wl], rows[9 + i][wl], rows[12 + i][wl]};
        double o[5];
        hemo_indices(v, count, o);
#pragma unroll
        for (int k = 0; k < 5; ++k) out5[(int64_t)k * 3 * nF + 3 * f + i] = o[k];
    }
}

// compute_hemodynamics.py:326-346
__global__ void k4_indices(const double* __restrict__ sums, int64_t nF, double count, double* __restrict__ tawss,
                           double* __restrict__ osi, double* __restrict__ rrt, double* __restrict__ ecap,
                           double* __restrict__ twssg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * nF) return;
    int64_t f = i % nF;
    int j = (int)(i / nF);
    const double v[5] = {sums[(int64_t)(3 * j + 0) * nF + f], sums[(int64_t)(3 * j + 1) * nF + f],
                         sums[(int64_t)(3 * j + 2) * nF + f], sums[(int64_t)(9 + j) * nF + f],
                         sums[(int64_t)(12 + j) * nF + f]};
    double o[5];
    hemo_indices(v, count, o);
    const int64_t q = 3 * f + j;
    tawss[q] = o[0];
    osi[q] = o[1];
    rrt[q] = o[2];
    ecap[q] = o[3];
    twssg[q] = o[4];
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
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double ld_peer(const double* p) {
    // system-scope relaxed load: goes to the owner's memory (peer lines are never in the local L2, and the local L1
    // holds nothing of them at this point of a fresh kernel), and -- unlike volatile -- loads may overlap
    double v;
    asm("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// compute_hemodynamics.py:326-346 on sums that are still spread over the GPUs of the node
__global__ void k4_peer_indices(PeerBlocks pb, int world, int rank, int64_t half_off, int64_t flags_off, uint64_t epoch,
                                int64_t nF, double count, double* __restrict__ red, double* __restrict__ tawss,
                                double* __restrict__ osi, double* __restrict__ rrt, double* __restrict__ ecap,
                                double* __restrict__ twssg, double* count_slot, double my_count) {
    // (launched on its own stream behind an event recorded after K3: this rank's sums are folded)
    // Block 0 first tells every rank "rank `rank` has finished epoch `epoch`": thread q stores the epoch into rank q's
    // arrival counter after a system-scope fence, so everything this GPU wrote before (K3's sums, the snapshot count
    // that rides behind them) is visible to a peer that has seen the counter.
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) *count_slot = my_count;
        __syncthreads();
        if ((int)threadIdx.x < world) {
            __threadfence_system();
            uint64_t* peer_flags = reinterpret_cast<uint64_t*>(const_cast<double*>(pb.block[threadIdx.x]) + flags_off);
            st_release_sys(peer_flags + rank, epoch);
        }
    }
    if ((int)threadIdx.x < world) {
        uint64_t* flags = reinterpret_cast<uint64_t*>(const_cast<double*>(pb.block[rank]) + flags_off);  // local memory
        const uint64_t t0 = global_ns();
        while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
            if (global_ns() - t0 > VH_PEER_WAIT_NS) {  // a rank died: give up, tell the host which one
                flags[VH_MAX_PEERS] = threadIdx.x + 1;
                break;
            }
            __nanosleep(20);
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

int k4_peer_reduce_finalize(vh_handle* h, const PeerBlocks& pb, int64_t half_off, int64_t flags_off, uint64_t epoch,
                            int64_t n_total, double* d_red, double* d_out5) {
    const int64_t nF = h->nF, n3 = 3 * nF;
    // One launch: block 0 signals this rank's arrival, every block waits for all ranks, then reduces its entries.  It
    // runs on s_aux behind this loop's K3, so that the K1/K2 of the next time loop overlap the wait for the peers; the
    // halves of the sum block alternate between loops, and the next K3 that writes waits for this kernel (vh_join_peer).
    vh_join_peer(h);  // reductions stay in order, and s_aux never runs two of them at once
    VH_CUDA(cudaEventRecord(h->ev_fork, h->s_compute));
    VH_CUDA(cudaStreamWaitEvent(h->s_aux, h->ev_fork, 0));
    k4_peer_indices<<<dim3((unsigned)((n3 + 255) / 256)), dim3(256), 0, h->s_aux>>>(
        pb, h->world, h->rank, half_off, flags_off, epoch, nF, (double)n_total, d_red, d_out5, d_out5 + n3, d_out5 + 2 * n3,
        d_out5 + 3 * n3, d_out5 + 4 * n3, h->d_sums + VH_NSUM * h->nF, (double)h->count);
    VH_CUDA(cudaGetLastError());
    VH_CUDA(cudaEventRecord(h->ev_join, h->s_aux));
    h->peer_pending = true;
    h->launches += 1;
    return VH_OK;
}

void vh_join_peer(vh_handle* h) {
    if (h->peer_pending) {
        cudaStreamWaitEvent(h->s_compute, h->ev_join, 0);
        h->peer_pending = false;
    }
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
    h->out5_count = -1;
    if (h->d_out5) cudaFree(h->d_out5);
    if (h->d_out5_peer) cudaFree(h->d_out5_peer);
    h->d_out5_peer = nullptr;
    h->peer_pending = false;
    if (h->h_out5) cudaFreeHost(h->h_out5);
    h->h_out5 = nullptr;
    h->d_sums = h->d_tau_last[0] = h->d_tau_last[1] = h->d_part = h->d_out5 = nullptr;
    h->part_cap = 0;
    for (int i = 0; i < 2; ++i) {
        if (h->d_stage[i]) cudaFree(h->d_stage[i]);
        if (h->d_wss_stage[i]) cudaFree(h->d_wss_stage[i]);
        h->d_stage[i] = h->d_wss_stage[i] = nullptr;
    }
    h->stage_cap = h->wss_stage_cap = 0;
    h->stage_row_bytes = 0;
    if (h->d_W) cudaFree(h->d_W);
    h->d_W = nullptr;
    h->w_ld = 0;
    h->begun = false;
    return VH_OK;
}

// Columns per staged block: as many as fit ~1/8 of the free memory (at most 8 GiB), between 64 and 4096; a
// batch_snapshots tuning value caps it (one column more than the batch, for the halo)
static int ensure_stage_block(vh_handle* h, int64_t want_cols) {
    int64_t cap = h->w_ld;
    if (h->batch_snapshots > 0 && want_cols > h->batch_snapshots + 1) want_cols = h->batch_snapshots + 1;
    if (cap >= want_cols || cap >= 4096) return VH_OK;
    size_t free_b = 0, total_b = 0;
    VH_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (h->d_W) free_b += (size_t)(h->nWn_pad * 24 * h->w_ld);
    int64_t budget = (int64_t)(free_b / 8);
    if (budget > (8LL << 30)) budget = 8LL << 30;
    int64_t cols = budget / (h->nWn_pad * 24);
    if (cols > 4096) cols = 4096;
    if (cols > want_cols) cols = want_cols;
    if (cols < 64 && h->batch_snapshots <= 0) cols = 64;
    cols = (cols + 63) / 64 * 64;  // whole 64-column passes (two tiles)
    while (cols > 64 && h->nWn_pad * 3 * cols >= (1LL << 31)) cols -= 64;  // K2 addresses W with int32 element offsets
    if (cols <= cap) return VH_OK;
    if (h->d_W) cudaFree(h->d_W);
    h->d_W = nullptr;
    h->w_ld = 0;
    VH_CUDA(cudaMalloc(&h->d_W, (size_t)(h->nWn_pad * 24 * cols)));
    h->w_ld = cols;
    return VH_OK;
}

namespace {

// Segments of one launch: the tiles of the block split as evenly as possible over gy segments.
struct SegPlan {
    int gy, base, extra;
};

SegPlan plan_segments(int64_t ncol, int64_t pass_cols, int64_t n_items, int64_t target_warps, int64_t chunk_snapshots) {
    const int64_t total = (ncol + pass_cols - 1) / pass_cols;
    int64_t gy;
    if (chunk_snapshots > 0) {
        const int64_t p = (chunk_snapshots + pass_cols - 1) / pass_cols;
        gy = (total + p - 1) / p;
    } else {
        gy = n_items > 0 ? (target_warps + n_items - 1) / n_items : 1;
    }
    if (gy < 1) gy = 1;
    if (gy > total) gy = total;
    if (gy > 65535) gy = 65535;
    return {(int)gy, (int)(total / gy), (int)(total % gy)};
}

template <int ORDER, int OCC>
int launch_k2(vh_handle* h, const K2Args& a, cudaStream_t st, bool pdl) {
    using L = K2Launch<ORDER, OCC>;
    // function attributes belong to the (function, device) pair: remembered per handle, i.e. per GPU
    bool& configured = h->k2_configured[4 * (ORDER - 1) + (OCC - 3)];
    if (!configured && L::SMEM_BYTES > 0) {
        VH_CUDA(cudaFuncSetAttribute(k2_wall<ORDER, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM_BYTES));
        VH_CUDA(cudaFuncSetAttribute(k2_wall<ORDER, OCC>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    const unsigned blocks = (unsigned)(a.multi.gx * a.multi.gy + a.single.gx * a.single.gy);
    VH_CUDA(vh_launch_pdl(k2_wall<ORDER, OCC>, dim3(blocks), dim3(32 * K2_WARPS), L::SMEM_BYTES, st, pdl, a));
    return VH_OK;
}


int64_t env_int(const char* name, int64_t dflt) {
    const char* v = getenv(name);
    return v && *v ? atoll(v) : dflt;
}

}  // namespace

int k2_launch(vh_handle* h, const double* d_u, int64_t n_snap, int64_t stride_elems, int prev_mode, double* d_wss,
              int64_t wss_ld, bool dense) {
    if (n_snap <= 0) return VH_OK;
    const int64_t nF = h->nF;
    VH_TRY(ensure_stage_block(h, n_snap + 1));
    const int64_t n_single = h->multi_start, n_multi = h->n_work - h->multi_start;
    const int gx_single = (int)((n_single + K2_WARPS - 1) / K2_WARPS);
    const int gx_multi = (int)((n_multi + K2_WARPS - 1) / K2_WARPS);
    // (facet, segment) warps per launch: a few waves of the resident warps, so that the tail is short and the
    // prologue (three dependent loads) of one warp hides behind the passes of the others
    // measured (headline mesh, 1000 snapshots): P1 step 53.3 / 56.2 / 60.4 us at 1 / 2 / 3 waves (K2 itself is flat, every
    // extra segment costs K3), P2 255 / 248 / 251 us
    static const int64_t waves_env = env_int("VASP_B200_K2_WAVES", 0);
    const int64_t waves = waves_env > 0 ? waves_env : (h->order == 1 ? 1 : 2);
    // blocks per SM the kernel is compiled for (see K2Launch): measured per order, profiles/r2_k2_occupancy.md
    static const int64_t occ_env = env_int("VASP_B200_K2_OCC", 0);
    int occ = occ_env >= 3 && occ_env <= 6 ? (int)occ_env
              : h->order == 2 ? K2_OCC_P2 : (h->nF >= K2_P1_LARGE_FACETS ? K2_OCC_P1_LARGE : K2_OCC_P1_SMALL);
    if (occ == 6 && h->order == 1) occ = 3;  // the direct-load shape only exists for P2
    int64_t pos = 0;
    while (pos < n_snap) {
        const int halo = (pos == 0 && prev_mode == 2) ? 1 : 0;
        int64_t nb = n_snap - pos;
        if (nb + halo > h->w_ld) nb = h->w_ld - halo;
        const int64_t ncol = nb + halo;
        const int64_t target = (int64_t)h->sm_count * (occ == 6 ? 3 : occ) * K2_WARPS * waves;
        const SegPlan ps = plan_segments(ncol, h->order == 1 ? 64 : 32, n_single, target, h->chunk_snapshots);
        // the few multi-facet-cell facets are cut finer and scheduled first, so that they never are the tail
        const SegPlan pm = plan_segments(ncol, 64, n_multi, target / 4, h->chunk_snapshots);
        const int64_t groups = (n_single ? ps.gy : 0) + (n_multi ? pm.gy : 0);
        if (groups > h->part_cap) {
            if (h->d_part) cudaFree(h->d_part);
            h->d_part = nullptr;
            h->part_cap = 0;
            // per segment: 15 partial sums + tau of the first and of the last column (2 x 9)
            VH_CUDA(cudaMalloc(&h->d_part, sizeof(double) * (VH_NSUM + 18) * h->n_work * groups));
            h->part_cap = groups;
        }
        const bool prof = h->profile && h->prof_used + 3 <= h->prof_pool.size();
        const int pdl = h->profile ? 0 : h->pdl;  // events between the kernels would serialise them anyway
        if (prof) cudaEventRecord(h->prof_pool[h->prof_used], h->s_compute);
        VH_TRY(k1_launch(h, d_u + (pos - halo) * stride_elems, ncol, stride_elems, dense));
        if (prof) cudaEventRecord(h->prof_pool[h->prof_used + 1], h->s_compute);
        K2Args a;
        a.T = vh_tables(h);
        a.W = h->d_W;
        a.ntile_ld = (int)(h->w_ld / 32);
        a.ncol = (int)ncol;
        a.r0 = halo;
        a.prev_mode = pos == 0 ? prev_mode : 1;
        a.single = {n_single ? gx_single : 0, n_single ? ps.gy : 0, ps.base, ps.extra, h->d_part,
                    h->d_part + (int64_t)h->part_cap * VH_NSUM * h->n_work};
        a.multi = {n_multi ? gx_multi : 0, n_multi ? pm.gy : 0, pm.base, pm.extra,
                   a.single.part + (int64_t)a.single.gy * VH_NSUM * h->n_work,
                   a.single.bnd + (int64_t)a.single.gy * 18 * h->n_work};
        static const int skip = (int)env_int("VASP_B200_K2_SKIP", 0);  // timing experiments only (wrong results)
        if (skip == 1) a.multi.gx = a.multi.gy = 0;
        if (skip == 2) a.single.gx = a.single.gy = 0;
        a.tau_last_in = h->d_tau_last[h->tau_cur];
        a.tau_last_out = h->d_tau_last[h->tau_cur ^ 1];
        h->tau_cur ^= 1;
        a.ws_col = wss_ld > 0 ? 1 : 9 * nF;
        a.ws_f = wss_ld > 0 ? 9 * wss_ld : 9;
        a.ws_i = wss_ld > 0 ? wss_ld : 1;
        a.wss_out = d_wss ? d_wss + pos * a.ws_col : nullptr;
        a.mu = h->mu;
        a.inv_dt = 1.0 / h->dt;
        const bool k2_pdl = (pdl & 2) != 0;
        if (h->order == 2) {
            if (occ == 6) VH_TRY((launch_k2<2, 6>(h, a, h->s_compute, k2_pdl)));
            else if (occ == 5) VH_TRY((launch_k2<2, 5>(h, a, h->s_compute, k2_pdl)));
            else if (occ == 4) VH_TRY((launch_k2<2, 4>(h, a, h->s_compute, k2_pdl)));
            else VH_TRY((launch_k2<2, 3>(h, a, h->s_compute, k2_pdl)));
        } else {
            if (occ == 5) VH_TRY((launch_k2<1, 5>(h, a, h->s_compute, k2_pdl)));
            else if (occ == 4) VH_TRY((launch_k2<1, 4>(h, a, h->s_compute, k2_pdl)));
            else VH_TRY((launch_k2<1, 3>(h, a, h->s_compute, k2_pdl)));
        }
        h->launches += 1;
        if (prof) {
            cudaEventRecord(h->prof_pool[h->prof_used + 2], h->s_compute);
            h->prof_used += 3;
        }
        vh_join_peer(h);  // K3 writes the running sums: a fused reduction of the loop before last must have read them
        VH_CUDA(vh_launch_pdl(k3_fold, dim3((unsigned)((h->n_work + K3_ITEMS - 1) / K3_ITEMS)), dim3(16 * K3_ITEMS), 0,
                              h->s_compute, (pdl & 4) != 0, h->d_sums, (const double*)a.single.part, (const double*)a.single.bnd,
                              a.single.gy, (const double*)a.multi.part, (const double*)a.multi.bnd, a.multi.gy,
                              (const int32_t*)h->d_work, h->n_work, h->multi_start, nF, a.inv_dt,
                              h->sums_pending_zero ? 1 : 0, (double)(h->count + pos + nb), h->d_out5));
        h->launches += 1;
        h->sums_pending_zero = false;
        pos += nb;
    }
    h->count += n_snap;
    h->out5_count = h->count;  // the last fold left the indices of exactly these snapshots in d_out5
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
