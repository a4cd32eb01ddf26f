/*
 * hemo_oracle.c -- plain-C twin of oracle/hemo_oracle.py's time loop (fp64), threaded over time shards.
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ (cross-checked against the numpy restatement), by
 * __graft_entry__.smoke() and as the timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`).
 * Nothing under vasp_b200/ links or loads it.  PARITY UNPINNED at 1e-10 (see hemo_oracle.py's header): the
 * reference's numerics live in dolfin/FFC/PETSc, which cannot be installed here.
 *
 * Follows src/vasp/postprocessing/postprocessing_fenics/compute_hemodynamics.py of the reference:
 *   assemble(inner(Ft, v)*ds)  :112-115,142-150  -> 3-point degree-2 facet rule, P2 (or P1) basis gradients
 *   LUSolver(A).solve          :105-110,116      -> per-cell 4x4 blocks, factor reused (inverse passed in)
 *   InterpolateDG.__call__     :65-89            -> copy map bcell_local (built once by the numpy oracle)
 *   |tau|, sum tau             :289-306
 *   project_dg(|dtau/dt|)      :309-312 + postprocessing_fenics_common.py:31-54 -> 7-point degree-5 rule
 * The reference runs this loop sequentially; threads here take contiguous time ranges and recompute the tau of
 * the snapshot before their range, partial sums are added in thread order.
 *
 * Build: see oracle/Makefile (gcc -O2 -pthread -shared -fPIC; the image has no libgomp, so plain pthreads).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <pthread.h>
#include <string.h>
#include <unistd.h>

typedef struct {
    int32_t order;          /* 1 or 2 */
    int32_t ndof;           /* 4 or 10 */
    int64_t nF, nW;
    const int64_t* facet_wall;  /* [nF] row of the owning cell in the wall-cell arrays */
    const int64_t* wall_nodes;  /* [nW][ndof] velocity node per cell dof, UFC order */
    const double* glam;         /* [nW][4][3] grad lambda_a */
    const double* normal;       /* [nF][3] */
    const double* area;         /* [nF] */
    const int8_t* facet_lv;     /* [nF][3] local vertices of the facet, ascending */
    const int8_t* bcell_local;  /* [nF][3] local vertex copied to boundary dof j */
    const double* Ainv;         /* [nW][4][4] inverse of the ident_zeros'ed surface mass block */
} hemo_maps;

static const int EDGE_A[6] = {2, 1, 1, 0, 0, 0};
static const int EDGE_B[6] = {3, 3, 2, 3, 2, 1};
static const double Q2_PTS[3][3] = {{2. / 3, 1. / 6, 1. / 6}, {1. / 6, 1. / 6, 2. / 3}, {1. / 6, 2. / 3, 1. / 6}};
#define QA 0.10128650732345633
#define QB 0.79742698535308720
#define QC 0.47014206410511505
#define QD 0.05971587178976981
static const double Q5_PTS[7][3] = {{1. / 3, 1. / 3, 1. / 3}, {QA, QB, QA}, {QA, QA, QB}, {QB, QA, QA},
                                    {QC, QD, QC}, {QC, QC, QD}, {QD, QC, QC}};
static const double Q5_WTS[7] = {0.225, 0.12593918054482717, 0.12593918054482717, 0.12593918054482717,
                                 0.13239415278850616, 0.13239415278850616, 0.13239415278850616};

/* tau[nF][3][3] for one snapshot vector; b is scratch [nW][4][3] */
static void snapshot_tau(const hemo_maps* m, const double* u, const int64_t off[3], int64_t node_stride, double mu,
                         double* b, double* tau) {
    const int nd = m->ndof;
    memset(b, 0, sizeof(double) * (size_t)m->nW * 12);
    for (int64_t f = 0; f < m->nF; ++f) {
        const int64_t w = m->facet_wall[f];
        const double(*g)[3] = (const double(*)[3])(m->glam + 12 * w);
        const double* n = m->normal + 3 * f;
        double uc[10][3];
        for (int k = 0; k < nd; ++k) {
            const int64_t s = node_stride * m->wall_nodes[w * nd + k];
            for (int c = 0; c < 3; ++c) uc[k][c] = u[off[c] + s];
        }
        for (int q = 0; q < 3; ++q) {
            double lam[4] = {0, 0, 0, 0};
            for (int i = 0; i < 3; ++i) lam[m->facet_lv[3 * f + i]] = Q2_PTS[q][i];
            double gphi[10][3];
            if (m->order == 2) {
                for (int a = 0; a < 4; ++a)
                    for (int d = 0; d < 3; ++d) gphi[a][d] = (4.0 * lam[a] - 1.0) * g[a][d];
                for (int e = 0; e < 6; ++e)
                    for (int d = 0; d < 3; ++d)
                        gphi[4 + e][d] = 4.0 * (lam[EDGE_A[e]] * g[EDGE_B[e]][d] + lam[EDGE_B[e]] * g[EDGE_A[e]][d]);
            } else {
                for (int a = 0; a < 4; ++a)
                    for (int d = 0; d < 3; ++d) gphi[a][d] = g[a][d];
            }
            double G[3][3] = {{0}};
            for (int k = 0; k < nd; ++k)
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) G[i][j] += uc[k][i] * gphi[k][j];
            double F[3], Fn = 0.0;
            for (int i = 0; i < 3; ++i) {
                double s = 0.0;
                for (int j = 0; j < 3; ++j) s += mu * (G[i][j] + G[j][i]) * n[j];
                F[i] = -s;
                Fn += F[i] * n[i];
            }
            const double wq = m->area[f] / 3.0;
            for (int a = 0; a < 4; ++a)
                for (int c = 0; c < 3; ++c) b[(w * 4 + a) * 3 + c] += wq * lam[a] * (F[c] - Fn * n[c]);
        }
    }
    for (int64_t f = 0; f < m->nF; ++f) {
        const int64_t w = m->facet_wall[f];
        const double* Ai = m->Ainv + 16 * w;
        for (int j = 0; j < 3; ++j) {
            const int a = m->bcell_local[3 * f + j];
            for (int c = 0; c < 3; ++c) {
                double x = 0.0;
                for (int p = 0; p < 4; ++p) x += Ai[4 * a + p] * b[(w * 4 + p) * 3 + c];
                tau[(f * 3 + j) * 3 + c] = x;
            }
        }
    }
}

static void accumulate(const hemo_maps* m, const double* tau, double* prev, double inv_dt_is_div, double* wss_sum,
                       double* tawss_sum, double* twssg_sum) {
    const double dt = inv_dt_is_div;
    for (int64_t f = 0; f < m->nF; ++f) {
        const double* t = tau + 9 * f;
        double* p = prev + 9 * f;
        double w[3][3];
        for (int j = 0; j < 3; ++j) {
            tawss_sum[3 * f + j] += sqrt(t[3 * j] * t[3 * j] + t[3 * j + 1] * t[3 * j + 1] + t[3 * j + 2] * t[3 * j + 2]);
            for (int c = 0; c < 3; ++c) {
                wss_sum[9 * f + 3 * j + c] += t[3 * j + c];
                w[j][c] = (t[3 * j + c] - p[3 * j + c]) / dt;
                p[3 * j + c] = t[3 * j + c];
            }
        }
        /* project_dg: M p = rhs, M = area/12 (1 + delta), rhs_i = area sum_q wq phi_i |w(x_q)| */
        double rhs[3] = {0, 0, 0};
        for (int q = 0; q < 7; ++q) {
            double v[3] = {0, 0, 0};
            for (int j = 0; j < 3; ++j)
                for (int c = 0; c < 3; ++c) v[c] += Q5_PTS[q][j] * w[j][c];
            const double mag = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            for (int i = 0; i < 3; ++i) rhs[i] += m->area[f] * Q5_WTS[q] * Q5_PTS[q][i] * mag;
        }
        const double s = rhs[0] + rhs[1] + rhs[2];
        for (int i = 0; i < 3; ++i) twssg_sum[3 * f + i] += (12.0 / m->area[f]) * (rhs[i] - 0.25 * s);
    }
}

typedef struct {
    const hemo_maps* m;
    const double* u;
    int64_t n_snap, stride, node_stride;
    const int64_t* off;
    double mu, dt;
    const double* tau_prev0;
    double *part, *tau_last, *wss_series;
    int t, nthreads, fail;
} shard_args;

static void* shard_main(void* p) {
    shard_args* a = (shard_args*)p;
    const hemo_maps* m = a->m;
    const int64_t nF = m->nF;
    const int64_t k0 = a->n_snap * a->t / a->nthreads, k1 = a->n_snap * (a->t + 1) / a->nthreads;
    double* b = (double*)malloc(sizeof(double) * (size_t)m->nW * 12);
    double* tau = (double*)malloc(sizeof(double) * (size_t)nF * 9);
    double* prev = (double*)calloc((size_t)nF * 9, sizeof(double));
    if (!b || !tau || !prev) {
        a->fail = 1;
    } else {
        double* ws = a->part + (size_t)a->t * 15 * nF;
        if (k0 > 0)
            snapshot_tau(m, a->u + (k0 - 1) * a->stride, a->off, a->node_stride, a->mu, b, prev);
        else if (a->tau_prev0)
            memcpy(prev, a->tau_prev0, sizeof(double) * (size_t)nF * 9);
        for (int64_t k = k0; k < k1; ++k) {
            snapshot_tau(m, a->u + k * a->stride, a->off, a->node_stride, a->mu, b, tau);
            if (a->wss_series) memcpy(a->wss_series + (size_t)k * nF * 9, tau, sizeof(double) * (size_t)nF * 9);
            accumulate(m, tau, prev, a->dt, ws, ws + 9 * nF, ws + 12 * nF);
        }
        if (k1 == a->n_snap && k1 > k0 && a->tau_last) memcpy(a->tau_last, prev, sizeof(double) * (size_t)nF * 9);
    }
    free(b);
    free(tau);
    free(prev);
    return NULL;
}

/* Runs snapshots [0, n_snap) of u (stride in doubles).  tau_prev0: tau before snapshot 0 ([nF][3][3]) or NULL for
 * zero (compute_hemodynamics.py:244).  Outputs are SUMS (not divided by the count).  wss_series may be NULL.
 * Returns 0, or -1 on allocation failure. */
int hemo_oracle_run(const hemo_maps* m, const double* u, int64_t n_snap, int64_t stride, const int64_t off[3],
                    int64_t node_stride, double mu, double dt, const double* tau_prev0, double* wss_sum,
                    double* tawss_sum, double* twssg_sum, double* tau_last, double* wss_series, int nthreads) {
    const int64_t nF = m->nF;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n_snap) nthreads = (int)(n_snap > 0 ? n_snap : 1);
    double* part = (double*)calloc((size_t)nthreads * 15 * nF, sizeof(double));
    shard_args* args = (shard_args*)calloc((size_t)nthreads, sizeof(shard_args));
    pthread_t* tid = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    if (!part || !args || !tid) {
        free(part);
        free(args);
        free(tid);
        return -1;
    }
    int fail = 0;
    for (int t = 0; t < nthreads; ++t) {
        shard_args a = {m, u, n_snap, stride, node_stride, off, mu, dt, tau_prev0, part, tau_last, wss_series,
                        t, nthreads, 0};
        args[t] = a;
    }
    for (int t = 1; t < nthreads; ++t)
        if (pthread_create(&tid[t], NULL, shard_main, &args[t]) != 0) args[t].fail = 2;
    shard_main(&args[0]);
    for (int t = 1; t < nthreads; ++t) {
        if (args[t].fail == 2) {
            args[t].fail = 0;
            shard_main(&args[t]); /* could not spawn: run the shard here */
        } else {
            pthread_join(tid[t], NULL);
        }
    }
    for (int t = 0; t < nthreads; ++t) fail |= args[t].fail;
    if (!fail) {
        for (int t = 0; t < nthreads; ++t) {
            const double* ws = part + (size_t)t * 15 * nF;
            for (int64_t i = 0; i < 9 * nF; ++i) wss_sum[i] += ws[i];
            for (int64_t i = 0; i < 3 * nF; ++i) tawss_sum[i] += ws[9 * nF + i];
            for (int64_t i = 0; i < 3 * nF; ++i) twssg_sum[i] += ws[12 * nF + i];
        }
    }
    free(part);
    free(args);
    free(tid);
    return fail ? -1 : 0;
}

int hemo_oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
