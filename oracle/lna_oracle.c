/*
 * lna_oracle.c -- CPU ORACLE (test infrastructure; see lna_oracle.h for the rules of use).
 *
 * A literal scalar restatement of the reference's arithmetic for the hot path, INCLUDING its
 * quirks: the non-standard "RK4" whose k4 uses dt/2, explicit-Euler
 * fundamental matrix / covariance, the 1e-8 Cholesky jitter, the -1e7 sentinels, the
 * logOrigin that drops its last row in one sampler only, ...
 * Compile with `-O2 -ffp-contract=off` (no FMA contraction, like a default R package build).
 *
 * Every function cites the reference file:line (under /root/reference/) it restates.
 */
#include "lna_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323846264338327950288 /* src/LNA_functional.h:6 */
#define MAXP 3

/* ---------------------------------------------------------------- small dense helpers (col-major) */
static void mat_mul(const double *A, const double *B, int p, double *C) { /* C = A*B */
    for (int c = 0; c < p; c++)
        for (int r = 0; r < p; r++) {
            double s = A[r] * B[c * p];
            for (int k = 1; k < p; k++) s += A[r + k * p] * B[k + c * p];
            C[r + c * p] = s;
        }
}
static void mat_mul_bt(const double *A, const double *B, int p, double *C) { /* C = A*B^T */
    for (int c = 0; c < p; c++)
        for (int r = 0; r < p; r++) {
            double s = A[r] * B[c];
            for (int k = 1; k < p; k++) s += A[r + k * p] * B[c + k * p];
            C[r + c * p] = s;
        }
}
/* arma::accu / sum() of a vector: two interleaved accumulators (arrayops::accumulate). */
static double accu2(const double *x, int n) {
    double a1 = 0.0, a2 = 0.0;
    int i, j;
    for (i = 0, j = 1; j < n; i += 2, j += 2) {
        a1 += x[i];
        a2 += x[j];
    }
    if (i < n) a1 += x[i];
    return a1 + a2;
}

/* ---------------------------------------------------------------- change-point transmission rate */

/* betaTs: LNA_functional.cpp:31-58.  Single sweep over sorted times. */
void orc_betaTs(const double *param, const double *times, int m, const double *x_r, const int *x_i,
                double *betas) {
    double R0 = param[x_i[2]];
    int i = 0, nch = x_i[0];
    for (int j = 0; j < m; j++) {
        if (i < nch) {
            while (x_r[i + 1] <= times[j]) {
                R0 *= param[x_i[1] + i];
                i++;
                if (i == nch) break;
            }
        }
        betas[j] = R0 * param[x_i[3]] / x_r[0];
    }
}

/* param_transform: LNA_functional.cpp:63-86.  Rescans from scratch; only entry 0 changes. */
void orc_param_transform(double t, const double *param, int nparam_total, const double *x_r,
                         const int *x_i, double *param2) {
    memcpy(param2, param, sizeof(double) * (size_t)nparam_total);
    double R0 = param[x_i[2]];
    int i = 0, nch = x_i[0];
    if (nch > 0) {
        while (x_r[i + 1] <= t) {
            R0 *= param[x_i[1] + i];
            if (i == nch - 1) break;
            i++;
        }
    }
    param2[0] = R0 * param[x_i[3]] / x_r[0];
}

/* Only thetas[0..2] are ever read by rhs / F / h, so the transform is applied on a 3-vector view. */
static void thetas3(const orc_cfg *c, double t, const double *param, double *th) {
    double R0 = param[c->x_i[2]];
    int i = 0, nch = c->x_i[0];
    if (nch > 0) {
        while (c->x_r[i + 1] <= t) {
            R0 *= param[c->x_i[1] + i];
            if (i == nch - 1) break;
            i++;
        }
    }
    th[0] = R0 * param[c->x_i[3]] / c->x_r[0];
    th[1] = c->nparam_total > 1 ? param[1] : 0.0;
    th[2] = c->nparam_total > 2 ? param[2] : 0.0;
}

/* ---------------------------------------------------------------- model rhs / Jacobian / hazards */

/* ODE_SIR_one :107-127, ODE_SEIR2_one :130-153, ODE_SEIR_one :156-198. */
void orc_ode_one(const orc_cfg *c, const double *x, const double *param, double t, double *res) {
    double th[3];
    thetas3(c, t, param, th);
    if (c->model == ORC_SIR) {
        if (c->transX == ORC_STANDARD) {
            res[0] = -th[0] * x[0] * x[1];
            res[1] = th[0] * x[0] * x[1] - th[1] * x[1];
        } else {
            res[0] = -th[0] * exp(x[1]) - th[0] * exp(x[1] - x[0]) / 2;
            res[1] = th[0] * exp(x[0]) - th[1] - th[0] * exp(x[0] - x[1]) / 2 - th[1] * exp(-x[1]) / 2;
        }
    } else if (c->model == ORC_SEIR2) {
        double N = c->x_r[0];
        res[0] = th[0] * N * x[1] - th[1] * x[0];
        res[1] = th[1] * x[0] - th[2] * x[1];
    } else { /* SEIR */
        if (c->transX == ORC_STANDARD) {
            res[0] = -th[0] * x[0] * x[2];
            res[1] = th[0] * x[0] * x[2] - th[1] * x[1];
            res[2] = th[1] * x[1] - th[2] * x[2];
        } else {
            res[0] = -th[0] * exp(x[2]) - th[0] * exp(x[2] - x[0]) / 2;
            res[1] = th[0] * exp(x[0] + x[2] - x[1]) - th[1] - th[0] * exp(x[0] + x[2] - 2 * x[1]) / 2 -
                     th[1] * exp(-x[1]) / 2;
            res[2] = th[1] * exp(x[1] - x[2]) - th[2] - th[2] * exp(-x[2]) / 2 -
                     th[1] * exp(x[1] - 2 * x[2]) / 2;
        }
    }
}

/* SIR_F :248-261, SEIR2_F :265-271, SEIR_F :276-296.  F is p x p column-major. */
void orc_F(const orc_cfg *c, const double *x, const double *th, double *F) {
    if (c->model == ORC_SIR) {
        if (c->transX == ORC_STANDARD) {
            F[0] = -th[0] * x[1];            F[2] = -th[0] * x[0];
            F[1] = th[0] * x[1];             F[3] = th[0] * x[0] - th[1];
        } else {
            F[0] = th[0] * exp(x[1] - x[0]) / 2;
            F[2] = -th[0] * exp(x[1]) - th[0] / 2 * exp(x[1] - x[0]);
            F[1] = th[0] * exp(x[0]) - th[0] * exp(x[0] - x[1]) / 2;
            F[3] = th[0] * exp(x[0] - x[1]) / 2 + th[1] / 2 * exp(-x[1]);
        }
    } else if (c->model == ORC_SEIR2) {
        double N = 1000000; /* hard-coded in the reference (:267) */
        F[0] = -th[1];  F[2] = th[0] * N;
        F[1] = th[1];   F[3] = -th[2];
    } else {
        if (c->transX == ORC_STANDARD) {
            F[0] = -th[0] * x[2];  F[3] = 0;       F[6] = -th[0] * x[0];
            F[1] = th[0] * x[2];   F[4] = -th[1];  F[7] = th[0] * x[0];
            F[2] = 0;              F[5] = th[1];   F[8] = -th[2];
        } else {
            F[0] = th[0] * exp(x[2] - x[0]) / 2;
            F[3] = 0;
            F[6] = -th[0] * exp(x[2]) - th[0] / 2 * exp(x[2] - x[0]);
            F[1] = th[0] * exp(x[0] + x[2] - x[1]) - th[0] * exp(x[0] + x[2] - 2 * x[1]) / 2;
            F[4] = -th[0] * exp(x[0] + x[2] - x[1]) + th[1] * exp(-x[1]) / 2 +
                   th[0] * exp(x[0] + x[2] - 2 * x[1]);
            F[7] = th[0] * exp(x[0] + x[2] - x[1]) - th[0] * exp(-x[0] + x[2] - 2 * x[1]);
            F[2] = 0;
            F[5] = th[1] * exp(x[1] - x[2]) - th[1] * exp(x[1] - 2 * x[2]) / 2;
            F[8] = -th[1] * exp(x[1] - x[2]) + th[2] * exp(-x[2]) / 2 + th[1] * exp(x[1] - 2 * x[2]);
        }
    }
}

/* SIR_h :301-314, SEIR2_h :317-324, SEIR_h :328-344 (log branch keeps the states(1) quirk, A.16).
 * Returns the number of reactions m. */
int orc_h(const orc_cfg *c, const double *x, const double *th, double *h) {
    if (c->model == ORC_SIR) {
        if (c->transX == ORC_STANDARD) {
            h[0] = th[0] * x[0] * x[1];
            h[1] = th[1] * x[1];
        } else {
            h[0] = th[0] * exp(x[0]) * exp(x[1]);
            h[1] = th[1] * exp(x[1]);
        }
        return 2;
    }
    if (c->model == ORC_SEIR2) {
        double N = 1000000;
        h[0] = th[0] * N * x[1];
        h[1] = th[1] * x[0];
        h[2] = th[2] * x[1];
        return 3;
    }
    if (c->transX == ORC_STANDARD) {
        h[0] = th[0] * x[0] * x[2];
        h[1] = th[1] * x[1];
        h[2] = th[2] * x[2];
    } else {
        h[0] = th[0] * exp(x[0]) * exp(x[1]);
        h[1] = th[1] * exp(x[1]);
        h[2] = th[2] * exp(x[2]);
    }
    return 3;
}

/* Stoichiometry A (m x p, rows = reactions), SigmaF :558-564.  Column-major m x p. */
static int stoich(const orc_cfg *c, double *A) {
    if (c->model == ORC_SIR) {
        /* [-1 1; 0 -1] */
        A[0] = -1; A[2] = 1;
        A[1] = 0;  A[3] = -1;
        return 2;
    }
    if (c->model == ORC_SEIR) {
        /* [-1 1 0; 0 -1 1; 0 0 -1] */
        A[0] = -1; A[3] = 1;  A[6] = 0;
        A[1] = 0;  A[4] = -1; A[7] = 1;
        A[2] = 0;  A[5] = 0;  A[8] = -1;
        return 3;
    }
    /* SEIR2: 3 reactions x 2 states: [1 0; -1 1; 0 -1] */
    A[0] = 1;  A[3] = 0;
    A[1] = -1; A[4] = 1;
    A[2] = 0;  A[5] = -1;
    return 3;
}

/* ---------------------------------------------------------------- integrator */

/* ODE_rk45: LNA_functional.cpp:455-491.  NOT a textbook RK4: all four stages at t[i-1] and k4 is
 * evaluated at X0 + k3*dt/2.  `(k*dt)/2`, `(k1/6+k2/3+k3/3+k4/6)*dt` as written. */
void orc_ode_rk45(const orc_cfg *c, const double *initial, const double *t, int n, const double *param,
                  double *OdeTraj) {
    int p = c->p;
    double dt = t[1] - t[0];
    double X0[MAXP], X1[MAXP], k1[MAXP], k2[MAXP], k3[MAXP], k4[MAXP], xs[MAXP];
    for (int i = 0; i < n; i++) OdeTraj[i] = t[i];
    for (int j = 0; j < p; j++) {
        OdeTraj[(size_t)(j + 1) * n] = initial[j];
        X1[j] = initial[j];
    }
    for (int i = 1; i < n; i++) {
        for (int j = 0; j < p; j++) X0[j] = X1[j];
        orc_ode_one(c, X0, param, t[i - 1], k1);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k1[j] * dt / 2;
        orc_ode_one(c, xs, param, t[i - 1], k2);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k2[j] * dt / 2;
        orc_ode_one(c, xs, param, t[i - 1], k3);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k3[j] * dt / 2;
        orc_ode_one(c, xs, param, t[i - 1], k4);
        for (int j = 0; j < p; j++) {
            X1[j] = X0[j] + (k1[j] / 6 + k2[j] / 3 + k3[j] / 3 + k4[j] / 6) * dt;
            OdeTraj[i + (size_t)(j + 1) * n] = X1[j];
        }
    }
}

/* ODE_rk45_stop: LNA_functional.cpp:495-530. */
double orc_ode_rk45_stop(const orc_cfg *c, const double *initial, const double *t, int n,
                         const double *param, double tol) {
    int p = c->p;
    double dt = t[1] - t[0];
    double X0[MAXP], X1[MAXP], k1[MAXP], k2[MAXP], k3[MAXP], k4[MAXP], xs[MAXP];
    for (int j = 0; j < p; j++) X1[j] = initial[j];
    int i = 1;
    for (; i < n; i++) {
        for (int j = 0; j < p; j++) X0[j] = X1[j];
        orc_ode_one(c, X0, param, t[i - 1], k1);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k1[j] * dt / 2;
        orc_ode_one(c, xs, param, t[i - 1], k2);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k2[j] * dt / 2;
        orc_ode_one(c, xs, param, t[i - 1], k3);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k3[j] * dt / 2;
        orc_ode_one(c, xs, param, t[i - 1], k4);
        for (int j = 0; j < p; j++) X1[j] = X0[j] + (k1[j] / 6 + k2[j] / 3 + k3[j] / 3 + k4[j] / 6) * dt;
        if (X1[c->x_i[3]] < tol) return t[i];
    }
    return t[i - 1] + 0.1;
}

/* SigmaF: LNA_functional.cpp:536-599.  Traj_par = `nrow` rows of a (ld x (p+1)) col-major matrix.
 * Explicit Euler at the rows' left endpoints; dt = row1.time - row0.time of THIS block (A.3). */
void orc_sigmaF(const orc_cfg *c, const double *Traj_par, int nrow, int ld, const double *param,
                double *expF, double *Sigma) {
    int p = c->p;
    double Q[MAXP * MAXP] = {0}, F0[MAXP * MAXP] = {0}, Sg[MAXP * MAXP] = {0};
    double A[9], At[9], F[MAXP * MAXP], h[3], state[MAXP], th[3];
    double T1[9], T2[9], M2[9], M3[9], M4[9], FF0[9];
    for (int j = 0; j < p; j++) Q[j + j * p] = 1.0, F0[j + j * p] = 1.0;
    int m = stoich(c, A); /* m x p */
    double dt = Traj_par[1] - Traj_par[0];
    for (int i = 0; i < nrow; i++) {
        double t = Traj_par[i];
        thetas3(c, t, param, th);
        for (int j = 0; j < p; j++) {
            state[j] = Traj_par[i + (size_t)(j + 1) * ld];
            if (c->transX == ORC_LOG) Q[j + j * p] = exp(-state[j]);
        }
        orc_h(c, state, th, h);
        orc_F(c, state, th, F);
        /* F0 = F0 + (F*F0)*dt */
        mat_mul(F, F0, p, FF0);
        for (int e = 0; e < p * p; e++) F0[e] = F0[e] + FF0[e] * dt;
        /* Sigma = Sigma + (Sigma*F' + F*Sigma + Q*A'*H*A*Q)*dt ; the 5-factor product left to right */
        mat_mul_bt(Sg, F, p, T1);
        mat_mul(F, Sg, p, T2);
        /* M1 = Q*A' (p x m) */
        for (int r = 0; r < p; r++)
            for (int q = 0; q < m; q++) {
                double s = 0.0;
                for (int k = 0; k < p; k++) s += Q[r + k * p] * A[q + k * m];
                At[r + q * p] = s;
            }
        /* M2 = M1*H (p x m), H = diag(h) */
        for (int r = 0; r < p; r++)
            for (int q = 0; q < m; q++) M2[r + q * p] = At[r + q * p] * h[q];
        /* M3 = M2*A (p x p) */
        for (int r = 0; r < p; r++)
            for (int q = 0; q < p; q++) {
                double s = M2[r] * A[q * m];
                for (int k = 1; k < m; k++) s += M2[r + k * p] * A[k + q * m];
                M3[r + q * p] = s;
            }
        mat_mul(M3, Q, p, M4);
        for (int e = 0; e < p * p; e++) Sg[e] = Sg[e] + ((T1[e] + T2[e]) + M4[e]) * dt;
    }
    memcpy(expF, F0, sizeof(double) * (size_t)(p * p));
    memcpy(Sigma, Sg, sizeof(double) * (size_t)(p * p));
}

/* KF_param: LNA_functional.cpp:603-634 (centred: Sigma itself). */
void orc_kf_param(const orc_cfg *c, const double *OdeTraj, int nrow, const double *param, int gridsize,
                  double *Acube, double *Scube) {
    int p = c->p, k = (nrow - 1) / gridsize;
    for (int i = 0; i < k; i++)
        orc_sigmaF(c, OdeTraj + (size_t)i * gridsize, gridsize, nrow, param, Acube + (size_t)i * p * p,
                   Scube + (size_t)i * p * p);
}

/* chol(S + jitter*I)^T (lower), as LAPACK dpotrf('U') would compute it on the upper triangle:
 * LNA_functional.cpp:661.  Returns 0 on success, 1 when a pivot is <= 0 or NaN (dpotrf info>0). */
int orc_chol_lower(const double *S, int p, double jitter, double *L) {
    double U[MAXP * MAXP] = {0};
    for (int j = 0; j < p; j++) {
        double ajj = S[j + j * p] + jitter;
        for (int k = 0; k < j; k++) ajj -= U[k + j * p] * U[k + j * p];
        if (!(ajj > 0.0)) return 1;
        ajj = sqrt(ajj);
        U[j + j * p] = ajj;
        for (int q = j + 1; q < p; q++) {
            double s = S[j + q * p];
            for (int k = 0; k < j; k++) s -= U[k + j * p] * U[k + q * p];
            U[j + q * p] = s / ajj;
        }
    }
    for (int r = 0; r < p; r++)
        for (int q = 0; q < p; q++) L[r + q * p] = U[q + r * p];
    return 0;
}

/* KF_param_chol: LNA_functional.cpp:638-670.  The "Sigma" slot holds L (A.4). */
int orc_kf_param_chol(const orc_cfg *c, const double *OdeTraj, int nrow, const double *param,
                      int gridsize, double *Acube, double *Lcube) {
    int p = c->p, k = (nrow - 1) / gridsize;
    double Sg[MAXP * MAXP];
    for (int i = 0; i < k; i++) {
        orc_sigmaF(c, OdeTraj + (size_t)i * gridsize, gridsize, nrow, param, Acube + (size_t)i * p * p, Sg);
        if (orc_chol_lower(Sg, p, 0.00000001, Lcube + (size_t)i * p * p)) return ORC_CHOL_FAIL;
    }
    return ORC_OK;
}

/* Ode_Coarse_Slicer: LNA_functional.cpp:1179-1188. */
void orc_ode_coarse_slicer(const double *Ode_thin, int nrow, int ncol, int gridsize, double *Ode_coarse) {
    int d = nrow / gridsize + 1;
    for (int i = 0; i < d; i++)
        for (int j = 0; j < ncol; j++) Ode_coarse[i + (size_t)j * d] = Ode_thin[(size_t)i * gridsize + (size_t)j * nrow];
}

/* New_Param_List: LNA_functional.cpp:1191-1208. */
int orc_new_param_list(const orc_cfg *c, const double *param, const double *initial, int gridsize,
                       const double *t, int nt, double *Acube, double *Lcube, double *Ode_coarse,
                       double *betaN) {
    int p = c->p;
    double *thin = (double *)malloc(sizeof(double) * (size_t)nt * (size_t)(p + 1));
    orc_ode_rk45(c, initial, t, nt, param, thin);
    int st = orc_kf_param_chol(c, thin, nt, param, gridsize, Acube, Lcube);
    if (st == ORC_OK) {
        orc_ode_coarse_slicer(thin, nt, p + 1, gridsize, Ode_coarse);
        orc_betaTs(param, Ode_coarse, nt / gridsize + 1, c->x_r, c->x_i, betaN);
    }
    free(thin);
    return st;
}

/* ---------------------------------------------------------------- non-centred trajectory */

/* TransformTraj: LNA_functional.cpp:877-900.  X1 = A_i X0 + L_i eps_i ; traj = ODE + X. */
void orc_transform_traj(const double *OdeTraj, int nrow, int p, const double *OriginLatent, int k,
                        const double *Acube, const double *Lcube, double *LNA_traj) {
    double X0[MAXP] = {0}, X1[MAXP];
    memcpy(LNA_traj, OdeTraj, sizeof(double) * (size_t)nrow * (size_t)(p + 1));
    for (int i = 0; i < k; i++) {
        const double *A = Acube + (size_t)i * p * p, *L = Lcube + (size_t)i * p * p;
        for (int r = 0; r < p; r++) {
            double ax = A[r] * X0[0], le = L[r] * OriginLatent[i];
            for (int q = 1; q < p; q++) {
                ax += A[r + q * p] * X0[q];
                le += L[r + q * p] * OriginLatent[i + (size_t)q * k];
            }
            X1[r] = ax + le;
        }
        for (int r = 0; r < p; r++) {
            LNA_traj[(i + 1) + (size_t)(r + 1) * nrow] = X1[r] + LNA_traj[(i + 1) + (size_t)(r + 1) * nrow];
            X0[r] = X1[r];
        }
    }
}

static double det_small(const double *M, int p) {
    if (p == 2) return M[0] * M[3] - M[2] * M[1];
    return M[0] * (M[4] * M[8] - M[7] * M[5]) - M[3] * (M[1] * M[8] - M[7] * M[2]) +
           M[6] * (M[1] * M[5] - M[4] * M[2]);
}

/* Traj_sim_general_noncentral: LNA_functional.cpp:824-872, eps supplied (k*p normals, step-major:
 * step i consumes normals[i*p .. i*p+p-1]).  OriginLatent is k x p column-major. */
void orc_traj_sim_noncentral(const double *OdeTraj, int nrow, int p, const double *Acube,
                             const double *Lcube, double t_correct, const double *normals,
                             double *LNA_traj, double *OriginLatent, double *logMultiNorm,
                             double *logOrigin) {
    int k = nrow - 1;
    double X0[MAXP] = {0}, X1[MAXP], loglike = 0.0, logTraj = 0.0;
    memcpy(LNA_traj, OdeTraj, sizeof(double) * (size_t)nrow * (size_t)(p + 1));
    for (int i = 0; i < k; i++) {
        const double *eps = normals + (size_t)i * p;
        const double *A = Acube + (size_t)i * p * p, *L = Lcube + (size_t)i * p * p;
        for (int q = 0; q < p; q++) OriginLatent[i + (size_t)q * k] = eps[q];
        for (int r = 0; r < p; r++) {
            double ax = A[r] * X0[0], le = L[r] * eps[0];
            for (int q = 1; q < p; q++) {
                ax += A[r + q * p] * X0[q];
                le += L[r + q * p] * eps[q];
            }
            X1[r] = ax + le;
        }
        for (int r = 0; r < p; r++) {
            LNA_traj[(i + 1) + (size_t)(r + 1) * nrow] = X1[r] + LNA_traj[(i + 1) + (size_t)(r + 1) * nrow];
            X0[r] = X1[r];
        }
        if (OdeTraj[i + 1] <= t_correct) {
            double ee = eps[0] * eps[0];
            for (int q = 1; q < p; q++) ee += eps[q] * eps[q];
            loglike += -log(det_small(L, p));
            logTraj += (-0.5) * ee;
        }
    }
    *logMultiNorm = loglike;
    *logOrigin = logTraj;
}

/* log_like_traj_general2: LNA_functional.cpp:675-722 (centred Gaussian transition density; used by
 * the reference's "Test SIR model" identity).  solve(S, v) by explicit small inverse. */
double orc_log_like_traj_general2(const double *SdeTraj, const double *OdeTraj, int nrow, int p,
                                  const double *Acube, const double *Scube, double t_correct) {
    int k = nrow - 1;
    double loglik = 0.0, Xd0[MAXP], Xd1[MAXP], v[MAXP], w[MAXP];
    for (int j = 0; j < p; j++) Xd1[j] = SdeTraj[(size_t)(j + 1) * nrow] - OdeTraj[(size_t)(j + 1) * nrow];
    for (int i = 0; i < k; i++) {
        const double *A = Acube + (size_t)i * p * p, *S = Scube + (size_t)i * p * p;
        for (int j = 0; j < p; j++) Xd0[j] = Xd1[j];
        for (int j = 0; j < p; j++)
            Xd1[j] = SdeTraj[(i + 1) + (size_t)(j + 1) * nrow] - OdeTraj[(i + 1) + (size_t)(j + 1) * nrow];
        if (SdeTraj[i + 1] <= t_correct) {
            for (int r = 0; r < p; r++) {
                double ax = 0.0;
                for (int q = 0; q < p; q++) ax += A[r + q * p] * Xd0[q];
                v[r] = Xd1[r] - ax;
            }
            double d = det_small(S, p);
            if (p == 2) {
                w[0] = (S[3] * v[0] - S[2] * v[1]) / d;
                w[1] = (-S[1] * v[0] + S[0] * v[1]) / d;
            } else {
                /* Cramer via cofactors */
                double C[9];
                C[0] = S[4] * S[8] - S[7] * S[5]; C[3] = -(S[3] * S[8] - S[6] * S[5]); C[6] = S[3] * S[7] - S[6] * S[4];
                C[1] = -(S[1] * S[8] - S[7] * S[2]); C[4] = S[0] * S[8] - S[6] * S[2]; C[7] = -(S[0] * S[7] - S[6] * S[1]);
                C[2] = S[1] * S[5] - S[4] * S[2]; C[5] = -(S[0] * S[5] - S[3] * S[2]); C[8] = S[0] * S[4] - S[3] * S[1];
                for (int r = 0; r < 3; r++) w[r] = (C[r] * v[0] + C[r + 3] * v[1] + C[r + 6] * v[2]) / d;
            }
            double q = 0.0;
            for (int r = 0; r < p; r++) q += v[r] * w[r];
            loglik += -log(d) / 2.0 - 0.5 * q;
        }
    }
    return loglik;
}

/* ---------------------------------------------------------------- coalescent likelihoods */

static double volz_core(const orc_init *init, const double *f1, int nrow, int p, const double *betaN,
                        double t_correct, const int *index, int transX, int nan_guard, int *st_out) {
    /* volz_loglik_nh2: LNA_functional.cpp:1119-1176 (nh3 :1058-1113 = same without the NaN guard) */
    int n2 = nrow, n0 = init->ng - 1, L = 0;
    while (f1[L] < t_correct) {
        L++;
        if (L == n2 - 1) break;
    }
    for (int j = 1; j <= p; j++)
        for (int r = 0; r <= n0; r++)
            if (f1[r + (size_t)j * nrow] < 0) {
                if (st_out) *st_out |= ORC_NEG_STATE;
                return ORC_SENTINEL;
            }
    int E = init->E;
    double *ll = (double *)malloc(sizeof(double) * (size_t)(E > 0 ? E : 1));
    int start = 0, has_nan = 0;
    for (int i = 0; i < n0; i++) {
        int row = L - i; /* f2(n0-i-1,.) = f1(L-i,.)  (A.9) */
        int reps = (int)init->gridrep[i];
        for (int j = 0; j < reps; j++) {
            double f, s, b = betaN[row];
            if (transX == ORC_STANDARD) {
                f = log(f1[row + (size_t)(index[1] + 1) * nrow]);
                s = log(f1[row + (size_t)(index[0] + 1) * nrow]);
            } else {
                f = f1[row + (size_t)(index[1] + 1) * nrow];
                s = f1[row + (size_t)(index[0] + 1) * nrow];
            }
            double v = -2 * (init->C[start] * init->D[start] * b * exp(-f) * exp(s)) +
                       init->y[start] * (log(b) - f + s);
            if (v != v) has_nan = 1;
            ll[start++] = v;
        }
    }
    double res;
    if (nan_guard && has_nan) {
        res = ORC_SENTINEL;
        if (st_out) *st_out |= ORC_NAN;
    } else {
        res = accu2(ll, start);
    }
    free(ll);
    return res;
}

double orc_volz_loglik_nh2(const orc_init *init, const double *f1, int nrow, int p, const double *betaN,
                           double t_correct, const int *index, int transX) {
    return volz_core(init, f1, nrow, p, betaN, t_correct, index, transX, 1, 0);
}
double orc_volz_loglik_nh3(const orc_init *init, const double *f1, int nrow, int p, const double *betaN,
                           double t_correct, const int *index, int transX) {
    return volz_core(init, f1, nrow, p, betaN, t_correct, index, transX, 0, 0);
}

static double kingman_core(const orc_init *init, const double *f1, int nrow, int col, double t_correct,
                           double lambda, int take_log) {
    int n0 = 0;
    while (f1[n0] < t_correct) n0++;
    int E = init->E, start = 0;
    double *ll = (double *)malloc(sizeof(double) * (size_t)(E > 0 ? E : 1));
    /* f2(n0-i) = [log] f1(i,col), i = n0..1 ; cell c = n0-i uses row n0-c */
    for (int c = 0; c < n0 && c < init->ng; c++) {
        int row = n0 - c, reps = (int)init->gridrep[c];
        double f = f1[row + (size_t)col * nrow];
        if (take_log) f = log(f);
        for (int j = 0; j < reps; j++) {
            ll[start] = -lambda * (init->C[start] * init->D[start] * exp(-f)) + init->y[start] * (log(lambda) - f);
            start++;
        }
    }
    (void)nrow;
    double res = accu2(ll, start);
    free(ll);
    return res;
}

/* coal_loglik3: LNA_functional.cpp:1016-1054 (Kingman with scale lambda; Index = 0-based state col). */
double orc_coal_loglik3(const orc_init *init, const double *f1, int nrow, int p, double t_correct,
                        double lambda, int Index, int transX) {
    (void)p;
    return kingman_core(init, f1, nrow, Index + 1, t_correct, lambda, transX == ORC_STANDARD);
}

/* coal_loglik: SIR_phylodyn.cpp:436-468 (column 2 of f1 already holds log I). */
double orc_coal_loglik(const orc_init *init, const double *f1, int nrow, int p, double t_correct,
                       double lambda) {
    (void)p;
    return kingman_core(init, f1, nrow, 2, t_correct, lambda, 0);
}

/* ---------------------------------------------------------------- FULL evaluation + MH step */

int orc_full_loglik(const orc_cfg *c, const orc_init *init, const double *param, const double *initial,
                    const double *t, int nt, int gridsize, const double *OriginTraj, double t_correct,
                    double *Acube, double *Lcube, double *Ode_coarse, double *betaN, double *LatentTraj,
                    double *coalLog) {
    int p = c->p, nrow = nt / gridsize + 1, k = (nt - 1) / gridsize;
    int st = orc_new_param_list(c, param, initial, gridsize, t, nt, Acube, Lcube, Ode_coarse, betaN);
    if (st != ORC_OK) {
        *coalLog = ORC_SENTINEL;
        return st;
    }
    orc_transform_traj(Ode_coarse, nrow, p, OriginTraj, k, Acube, Lcube, LatentTraj);
    int vst = 0;
    *coalLog = volz_core(init, LatentTraj, nrow, p, betaN, t_correct, &c->x_i[2], c->transX, 1, &vst);
    return vst; /* ORC_NEG_STATE / ORC_NAN are informational: the value is the in-band -1e7 */
}

/* Update_Param: LNA_functional.cpp:1212-1243.  One uniform; accept iff log(u) < a. */
int orc_update_param(const orc_cfg *c, const orc_init *init, const double *param, const double *initial,
                     const double *t, int nt, const double *OriginTraj, int gridsize, double coal_log,
                     double prior_proposal_offset, double t_correct, double unif, int *accept,
                     double *Acube, double *Lcube, double *Ode_coarse, double *betaN, double *LatentTraj,
                     double *coalLog_new) {
    int st = orc_full_loglik(c, init, param, initial, t, nt, gridsize, OriginTraj, t_correct, Acube, Lcube,
                             Ode_coarse, betaN, LatentTraj, coalLog_new);
    if (st & ORC_CHOL_FAIL) { /* the reference throws here (-> R tryCatch keeps the old state) */
        *accept = 0;
        return st;
    }
    double a = *coalLog_new - coal_log + prior_proposal_offset;
    *accept = (log(unif) < a) ? 1 : 0;
    return st;
}

/* ---------------------------------------------------------------- ESS rotations */

/* Param_Slice_update: LNA_functional.cpp:1246-1253. */
void orc_param_slice_update(const double *param, int nparam_total, const int *x_i, double theta,
                            const double *newChs, double *param_new) {
    memcpy(param_new, param, sizeof(double) * (size_t)nparam_total);
    for (int j = 0; j < x_i[0]; j++) {
        double od = log(param[x_i[1] + j]);
        param_new[x_i[1] + j] = exp(od * cos(theta) + newChs[j] * sin(theta));
    }
}

/* Param_Slice_update_all2: LNA_functional.cpp:1268-1314 (hard-wired p=2). */
void orc_param_slice_update_all2(const double *par_old, int npar, const int *x_i, double theta,
                                 const double *newChs, const int *ESS_vec, const double *pr,
                                 double *par_new) {
    int nz = x_i[1] + x_i[0] + 2; /* length of par_old.subvec(1, x_i1+x_i0+2) */
    int ih = x_i[0] + x_i[1] + 1; /* hyper inside OdeChs */
    double *z = (double *)malloc(sizeof(double) * (size_t)nz);
    double *zp = (double *)malloc(sizeof(double) * (size_t)nz);
    memcpy(par_new, par_old, sizeof(double) * (size_t)npar);
    for (int i = 0; i < nz; i++) z[i] = log(par_old[1 + i]);
    for (int i = 0; i < 3; i++) z[i] = (z[i] - pr[2 * i]) / pr[2 * i + 1];
    z[ih] = (z[ih] - pr[6]) / pr[7];
    for (int i = 3; i <= 2 + x_i[0]; i++) z[i] = z[i] * par_new[x_i[0] + x_i[1] + 2];
    double ct = cos(theta), sn = sin(theta);
    for (int i = 0; i < nz; i++) zp[i] = z[i] * ct + newChs[i] * sn;
    for (int i = 0; i < 3; i++)
        if (ESS_vec[i] != 0) par_new[1 + i] = exp(zp[i] * pr[2 * i + 1] + pr[2 * i]);
    if (ESS_vec[4] != 0) par_new[x_i[0] + x_i[1] + 2] = exp(zp[ih] * pr[7] + pr[6]);
    if (ESS_vec[5] != 0)
        for (int i = 3; i <= x_i[0] + x_i[1]; i++) par_new[1 + i] = exp(zp[i] / par_new[x_i[0] + x_i[1] + 2]);
    free(z);
    free(zp);
}

/* R::runif(a,b) = a + (b-a)*unif_rand()  (R nmath/runif.c) */
static double runif_ab(double a, double b, double u) { return a + (b - a) * u; }

/* ESlice_par_General: LNA_functional.cpp:1532-1688.
 * normals: npar-1 (then k*p more, discarded, iff ESS_vec[6]==1); uniforms: u, theta_1, one per shrink. */
int orc_eslice_par_general(const orc_cfg *c, const orc_init *init, const double *par_old, int npar,
                           const double *t, int nt, const double *OriginTraj, const double *pr8,
                           int gridsize, const int *ESS_vec, double coal_log, double t_correct,
                           const double *normals, const double *uniforms, int *n_norm, int *n_unif,
                           int *n_evals, double *par_new, double *Acube, double *Lcube,
                           double *Ode_coarse, double *betaN, double *LatentTraj,
                           double *OriginTraj_new, double *CoalLog, double *logOrigin) {
    int p = 2, k = (nt - 1) / gridsize, nrow = nt / gridsize + 1;
    int u_param = npar - 1, in = 0, iu = 0, evals = 0, status = 0;
    const double *new_par_raw = normals;
    in += u_param;
    if (ESS_vec[6] == 1) in += k * p; /* drawn into a block-scoped variable and discarded (:1576-1578) */
    memcpy(OriginTraj_new, OriginTraj, sizeof(double) * (size_t)k * p);
    double u = uniforms[iu++];
    double logy = coal_log + log(u);
    double theta = 2 * ORC_PI, theta_min = 0, theta_max = 2 * ORC_PI, loglike = 0.0;
    memcpy(par_new, par_old, sizeof(double) * (size_t)npar);
    int i = 0;
    do {
        i += 1;
        if (i > 20) {
            theta = 0;
            memcpy(par_new, par_old, sizeof(double) * (size_t)npar);
            memcpy(OriginTraj_new, OriginTraj, sizeof(double) * (size_t)k * p);
            int st = orc_full_loglik(c, init, par_new + 2, par_new, t, nt, gridsize, OriginTraj_new, t_correct,
                                     Acube, Lcube, Ode_coarse, betaN, LatentTraj, &loglike);
            evals++;
            status |= ORC_CAP_HIT | st; /* a chol failure here propagates as an exception in the reference */
            break;
        }
        if (i == 1) {
            theta = runif_ab(0, 2 * ORC_PI, uniforms[iu++]);
            theta_min = theta - 2 * ORC_PI;
            theta_max = theta;
        } else {
            if (theta < 0) theta_min = theta;
            else theta_max = theta;
            theta = runif_ab(theta_min, theta_max, uniforms[iu++]);
        }
        orc_param_slice_update_all2(par_old, npar, c->x_i, theta, new_par_raw, ESS_vec, pr8, par_new);
        evals++;
        int st = orc_new_param_list(c, par_new + 2, par_new, gridsize, t, nt, Acube, Lcube, Ode_coarse, betaN);
        if (st != ORC_OK) { /* try/catch(...) -> loglike = -1e7 ; continue (:1641-1651) */
            loglike = ORC_SENTINEL;
            continue;
        }
        if (ESS_vec[6] == 1) { /* mixes OriginTraj with (the running) itself, :1660-1662 */
            double ct = cos(theta), sn = sin(theta);
            for (int e = 0; e < k * p; e++) OriginTraj_new[e] = OriginTraj[e] * ct + OriginTraj_new[e] * sn;
        }
        orc_transform_traj(Ode_coarse, nrow, p, OriginTraj_new, k, Acube, Lcube, LatentTraj);
        loglike = orc_volz_loglik_nh2(init, LatentTraj, nrow, p, betaN, t_correct, &c->x_i[2], c->transX);
    } while (loglike <= logy);
    double lo = 0.0;
    for (int j = 0; j < k; j++)
        for (int q = 0; q < p; q++) lo -= 0.5 * OriginTraj_new[j + (size_t)q * k] * OriginTraj_new[j + (size_t)q * k];
    *logOrigin = lo;
    *CoalLog = loglike;
    *n_norm = in;
    *n_unif = iu;
    *n_evals = evals;
    return status;
}

/* LogTraj: SIR_LNA.cpp:40-46 (time column copied, log of every state column). */
static void log_traj(const double *traj, int nrow, int p, double *out) {
    for (int r = 0; r < nrow; r++) out[r] = traj[r];
    for (int e = nrow; e < nrow * (p + 1); e++) out[e] = log(traj[e]);
}

/* ESlice_general_NC: LNA_functional.cpp:1690-1778.
 * normals: k*p (column-major k x p); uniforms: u, theta_1, one per shrink (<= 20). */
int orc_eslice_general_nc(const orc_init *init, const double *f_cur, int k, int p, const double *OdeTraj,
                          const double *Acube, const double *Lcube, const double *betaN, double t_correct,
                          double lambda, double coal_log, int volz, const int *Index, int transX,
                          const double *normals, const double *uniforms, int *n_norm, int *n_unif,
                          int *n_evals, double *LatentTraj, double *OriginTraj_new, double *logOrigin,
                          double *CoalLog) {
    int nrow = k + 1, iu = 0, evals = 0, status = 0;
    const double *v_traj = normals;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)nrow * (p + 1));
    double u = uniforms[iu++], logy;
    /* volz = FALSE (:1717-1721, 1733-1760): the Kingman likelihood of LogTraj(newTraj) (SIR_LNA.cpp:40-46: log of every state
     * column) through coal_loglik3, which with transX = "standard" takes the log AGAIN -- log(log I), NaN as soon as I < 1.
     * A NaN compares false in `loglike <= logy`, so it ENDS the loop and is returned.  The 20-shrink cap still recomputes with
     * volz_loglik_nh2 (:1743).  With coal_log == -99999999 the reference hands the k x p increment matrix f_cur to
     * coal_loglik3 as if it were a trajectory (:1720): not restated (rejected), no driver reaches it. */
#define NC_LOGLIKE(traj)                                                                              \
    (volz ? orc_volz_loglik_nh2(init, traj, nrow, p, betaN, t_correct, Index, transX)                  \
          : (log_traj(traj, nrow, p, tmp), orc_coal_loglik3(init, tmp, nrow, p, t_correct, lambda, Index[1], transX)))
    if (coal_log != -99999999) {
        logy = coal_log + log(u);
    } else {
        if (!volz) {
            free(tmp);
            return -1;
        }
        orc_transform_traj(OdeTraj, nrow, p, f_cur, k, Acube, Lcube, tmp);
        logy = orc_volz_loglik_nh2(init, tmp, nrow, p, betaN, t_correct, Index, transX) + log(u);
        evals++;
    }
    double theta = runif_ab(0, 2 * ORC_PI, uniforms[iu++]);
    double theta_min = theta - 2 * ORC_PI, theta_max = theta;
    double *f_prime = OriginTraj_new;
    double ct = cos(theta), sn = sin(theta);
    for (int e = 0; e < k * p; e++) f_prime[e] = f_cur[e] * ct + v_traj[e] * sn;
    orc_transform_traj(OdeTraj, nrow, p, f_prime, k, Acube, Lcube, LatentTraj);
    double loglike = NC_LOGLIKE(LatentTraj);
    evals++;
    int i = 0;
    for (;;) {
        double mn = LatentTraj[nrow];
        for (int e = nrow; e < nrow * (p + 1); e++)
            if (LatentTraj[e] < mn) mn = LatentTraj[e];
        if (!(mn < 0 || loglike <= logy)) break;
        i += 1;
        if (i > 20) {
            orc_transform_traj(OdeTraj, nrow, p, f_cur, k, Acube, Lcube, LatentTraj);
            loglike = orc_volz_loglik_nh2(init, LatentTraj, nrow, p, betaN, t_correct, Index, transX);
            evals++;
            memcpy(f_prime, f_cur, sizeof(double) * (size_t)k * p);
            status |= ORC_CAP_HIT;
            break;
        }
        if (theta < 0) theta_min = theta;
        else theta_max = theta;
        theta = runif_ab(theta_min, theta_max, uniforms[iu++]);
        ct = cos(theta);
        sn = sin(theta);
        for (int e = 0; e < k * p; e++) f_prime[e] = f_cur[e] * ct + v_traj[e] * sn;
        orc_transform_traj(OdeTraj, nrow, p, f_prime, k, Acube, Lcube, LatentTraj);
        loglike = NC_LOGLIKE(LatentTraj);
        evals++;
    }
#undef NC_LOGLIKE
    double lo = 0.0;
    for (int j = 0; j < k - 1; j++) /* omits the LAST row (:1768) */
        for (int q = 0; q < p; q++) lo -= 0.5 * f_prime[j + (size_t)q * k] * f_prime[j + (size_t)q * k];
    *logOrigin = lo;
    *CoalLog = loglike;
    *n_norm = k * p;
    *n_unif = iu;
    *n_evals = evals;
    free(tmp);
    return status;
}

/* ESlice_change_points: LNA_functional.cpp:1317-1409.
 * uniforms: u, theta_1 FIRST; then normals: nch (divided by hyper); then one uniform per shrink. */
int orc_eslice_change_points(const orc_cfg *c, const orc_init *init, const double *param,
                             const double *initial, const double *t, int nt, const double *OriginTraj,
                             int gridsize, double coal_log, double t_correct, const double *normals,
                             const double *uniforms, int *n_norm, int *n_unif, int *n_evals,
                             double *param_new, double *Acube, double *Lcube, double *Ode_coarse,
                             double *betaN, double *LatentTraj, double *CoalLog, double *LogChProb) {
    int nch = c->x_i[0], npt = c->nparam_total, iu = 0, evals = 0, status = 0;
    double u = uniforms[iu++];
    double logy = coal_log + log(u);
    double theta = runif_ab(0, 2 * ORC_PI, uniforms[iu++]);
    double theta_min = theta - 2 * ORC_PI, theta_max = theta;
    double *newChs = (double *)malloc(sizeof(double) * (size_t)(nch > 0 ? nch : 1));
    for (int j = 0; j < nch; j++) newChs[j] = normals[j] / param[c->x_i[0] + c->x_i[1]];
    orc_param_slice_update(param, npt, c->x_i, theta, newChs, param_new);
    double loglike;
    int st = orc_full_loglik(c, init, param_new, initial, t, nt, gridsize, OriginTraj, t_correct, Acube, Lcube,
                             Ode_coarse, betaN, LatentTraj, &loglike);
    evals++;
    int i = 0;
    /* a Cholesky failure is an uncaught exception in the reference: the update aborts */
    while (!(st & ORC_CHOL_FAIL) && loglike <= logy) {
        i += 1;
        if (i > 20) {
            theta = 0;
            memcpy(param_new, param, sizeof(double) * (size_t)npt);
            st = orc_full_loglik(c, init, param, initial, t, nt, gridsize, OriginTraj, t_correct, Acube, Lcube,
                                 Ode_coarse, betaN, LatentTraj, &loglike);
            evals++;
            status |= ORC_CAP_HIT;
            break;
        }
        if (theta < 0) theta_min = theta;
        else theta_max = theta;
        theta = runif_ab(theta_min, theta_max, uniforms[iu++]);
        orc_param_slice_update(param, npt, c->x_i, theta, newChs, param_new);
        st = orc_full_loglik(c, init, param_new, initial, t, nt, gridsize, OriginTraj, t_correct, Acube, Lcube,
                             Ode_coarse, betaN, LatentTraj, &loglike);
        evals++;
    }
    status |= st; /* status of the evaluation the call ended on */
    double lcp = 0.0;
    for (int j = 0; j < nch; j++) {
        double d = log(param_new[c->x_i[1] + j]) * param_new[c->x_i[0] + c->x_i[1]];
        lcp += -0.5 * d * d;
    }
    *LogChProb = lcp;
    *CoalLog = loglike;
    *n_norm = nch;
    *n_unif = iu;
    *n_evals = evals;
    free(newChs);
    return status;
}

/* ---------------------------------------------------------------- basic_util.cpp */

/* inv2: basic_util.cpp:5-14 */
void orc_inv2(const double *a, double *res) {
    double r00 = a[3], r11 = a[0], r10 = -a[1], r01 = -a[2];
    double D = r00 * r11 - r10 * r01;
    res[0] = r00 * (1 / D);
    res[1] = r10 * (1 / D);
    res[2] = r01 * (1 / D);
    res[3] = r11 * (1 / D);
}

/* chols: basic_util.cpp:17-25 (hand 2x2 Cholesky, returns lower) */
void orc_chols(const double *S, double *L) {
    double r00 = sqrt(S[0]);
    double r01 = S[1] / r00;
    double r11 = sqrt(S[3] - r01 * r01);
    L[0] = r00;
    L[1] = r01;
    L[2] = 0;
    L[3] = r11;
}

/* ---------------------------------------------------------------- centred parametrisation */

/* mvrnormArma: basic_util.cpp:44-58.  n == 2: chols(sigma) * z, else chol(sigma + 1e-13 I)^T * z. */
static int mvrnorm_small(const double *Sig, int p, const double *z, double *out) {
    double L[MAXP * MAXP];
    if (p == 2) {
        orc_chols(Sig, L);
    } else if (orc_chol_lower(Sig, p, 0.0000000000001, L)) {
        return 1;
    }
    for (int r = 0; r < p; r++) {
        double s = 0.0;
        for (int q = 0; q < p; q++) s += L[r + q * p] * z[q];
        out[r] = s;
    }
    return 0;
}

static void solve_small(const double *S, int p, const double *v, double *w) {
    double d = det_small(S, p);
    if (p == 2) {
        w[0] = (S[3] * v[0] - S[2] * v[1]) / d;
        w[1] = (-S[1] * v[0] + S[0] * v[1]) / d;
    } else {
        double C[9];
        C[0] = S[4] * S[8] - S[7] * S[5]; C[3] = -(S[3] * S[8] - S[6] * S[5]); C[6] = S[3] * S[7] - S[6] * S[4];
        C[1] = -(S[1] * S[8] - S[7] * S[2]); C[4] = S[0] * S[8] - S[6] * S[2]; C[7] = -(S[0] * S[7] - S[6] * S[1]);
        C[2] = S[1] * S[5] - S[4] * S[2]; C[5] = -(S[0] * S[5] - S[3] * S[2]); C[8] = S[0] * S[4] - S[3] * S[1];
        for (int r = 0; r < 3; r++) w[r] = (C[r] * v[0] + C[r + 3] * v[1] + C[r + 6] * v[2]) / d;
    }
}

/* Traj_sim_general3: LNA_functional.cpp:771-817 (centred LNA simulation; normals step-major k*p).
 * Quirk kept: LNA_traj(0,0) = 0 whatever OdeTraj(0,0) is. */
int orc_traj_sim_general3(const double *OdeTraj, int nrow, int p, const double *Acube, const double *Scube,
                          double t_correct, const double *normals, double *LNA_traj, double *loglike_out) {
    int k = nrow - 1;
    double X0[MAXP], X1[MAXP], eta0[MAXP], eta1[MAXP], noise[MAXP], w[MAXP], loglike = 0.0;
    for (int j = 0; j < p; j++) {
        X1[j] = OdeTraj[(size_t)(j + 1) * nrow];
        eta1[j] = X1[j];
        LNA_traj[(size_t)(j + 1) * nrow] = X1[j];
    }
    LNA_traj[0] = 0;
    for (int i = 0; i < k; i++) {
        const double *A = Acube + (size_t)i * p * p, *Sig = Scube + (size_t)i * p * p;
        for (int j = 0; j < p; j++) {
            X0[j] = X1[j];
            eta0[j] = eta1[j];
            eta1[j] = OdeTraj[(i + 1) + (size_t)(j + 1) * nrow];
        }
        if (mvrnorm_small(Sig, p, normals + (size_t)i * p, noise)) return ORC_CHOL_FAIL;
        for (int r = 0; r < p; r++) {
            double ax = 0.0;
            for (int q = 0; q < p; q++) ax += A[r + q * p] * (X0[q] - eta0[q]);
            X1[r] = ax + eta1[r] + noise[r];
        }
        if (OdeTraj[i + 1] <= t_correct) {
            solve_small(Sig, p, noise, w);
            double q = 0.0;
            for (int r = 0; r < p; r++) q += noise[r] * w[r];
            loglike += -log(det_small(Sig, p)) / 2.0 + (-0.5) * q;
        }
        LNA_traj[i + 1] = OdeTraj[i + 1];
        for (int j = 0; j < p; j++) LNA_traj[(i + 1) + (size_t)(j + 1) * nrow] = X1[j];
    }
    *loglike_out = loglike;
    return ORC_OK;
}

/* Traj_sim_ezG2: LNA_functional.cpp:907-975 = ODE_rk45 + coarse rows + KF_param + the loop of Traj_sim_general3. */
int orc_traj_sim_ezG2(const orc_cfg *c, const double *initial, const double *t, int nt, const double *param,
                      int gridsize, double t_correct, const double *normals, double *LNA_traj, double *loglike) {
    int p = c->p, k = nt / gridsize, nrow = k + 1;
    double *thin = (double *)malloc(sizeof(double) * (size_t)nt * (p + 1));
    double *coarse = (double *)malloc(sizeof(double) * (size_t)nrow * (p + 1));
    double *A = (double *)malloc(sizeof(double) * (size_t)p * p * k * 2), *S = A + (size_t)p * p * k;
    orc_ode_rk45(c, initial, t, nt, param, thin);
    for (int i = 0; i < nrow; i++)
        for (int j = 0; j <= p; j++) coarse[i + (size_t)j * nrow] = thin[(size_t)i * gridsize + (size_t)j * nt];
    orc_kf_param(c, thin, nt, param, gridsize, A, S);
    int st = orc_traj_sim_general3(coarse, nrow, p, A, S, t_correct, normals, LNA_traj, loglike);
    free(thin);
    free(coarse);
    free(A);
    return st;
}

/* ESlice_general2: LNA_functional.cpp:1783-1857 (centred elliptical slice, volz likelihood).
 * normals: k*p (consumed by Traj_sim_general3); uniforms: u, theta_1, one per shrink. */
int orc_eslice_general2(const orc_init *init, const double *f_cur, int nrow, int p, const double *OdeTraj,
                        const double *Acube, const double *Scube, const double *betaN, double t_correct,
                        const int *Index, int transX, const double *normals, const double *uniforms, int *n_unif,
                        int *n_evals, double *newTraj) {
    int iu = 0, evals = 0, n = nrow * p;
    double *v = (double *)malloc(sizeof(double) * (size_t)nrow * (p + 1));
    double ll_sim;
    int st = orc_traj_sim_general3(OdeTraj, nrow, p, Acube, Scube, t_correct, normals, v, &ll_sim);
    if (st) {
        free(v);
        return st;
    }
    double u = uniforms[iu++];
    double logy = orc_volz_loglik_nh2(init, f_cur, nrow, p, betaN, t_correct, Index, transX) + log(u);
    evals++;
    double theta = runif_ab(0, 2 * ORC_PI, uniforms[iu++]);
    double theta_min = theta - 2 * ORC_PI, theta_max = theta;
    for (int r = 0; r < nrow; r++) newTraj[r] = f_cur[r];
    double ct = cos(theta), sn = sin(theta);
    for (int e = 0; e < n; e++) {
        double fc = f_cur[nrow + e] - OdeTraj[nrow + e], vt = v[nrow + e] - OdeTraj[nrow + e];
        newTraj[nrow + e] = (fc * ct + vt * sn) + OdeTraj[nrow + e];
    }
    double loglike = orc_volz_loglik_nh2(init, newTraj, nrow, p, betaN, t_correct, Index, transX);
    evals++;
    int i = 0, status = 0;
    for (;;) {
        double mn = newTraj[nrow];
        for (int e = nrow; e < nrow * (p + 1); e++)
            if (newTraj[e] < mn) mn = newTraj[e];
        if (!(mn < 0 || loglike <= logy)) break;
        i += 1;
        if (i > 20) {
            memcpy(newTraj, f_cur, sizeof(double) * (size_t)nrow * (p + 1));
            status |= ORC_CAP_HIT;
            break;
        }
        if (theta < 0) theta_min = theta;
        else theta_max = theta;
        theta = runif_ab(theta_min, theta_max, uniforms[iu++]);
        ct = cos(theta);
        sn = sin(theta);
        for (int e = 0; e < n; e++) {
            double fc = f_cur[nrow + e] - OdeTraj[nrow + e], vt = v[nrow + e] - OdeTraj[nrow + e];
            newTraj[nrow + e] = (fc * ct + vt * sn) + OdeTraj[nrow + e];
        }
        loglike = orc_volz_loglik_nh2(init, newTraj, nrow, p, betaN, t_correct, Index, transX);
        evals++;
    }
    *n_unif = iu;
    *n_evals = evals;
    free(v);
    return status;
}

/* ---------------------------------------------------------------- seasonal SIRS (SIRS_period.cpp) */

/* SIRS_ODE: SIRS_period.cpp:122-144.  param = (beta, gamma, mu, alpha); states (S, I, R). */
static void sirs_rhs(const double *x, const double *param, double t, double period, double *res) {
    double th1 = param[0] * (1 + param[3] * sin(2 * ORC_PI * t / period));
    res[0] = -th1 * x[0] * x[1] + param[2] * x[2];
    res[1] = th1 * x[0] * x[1] - param[1] * x[1];
    res[2] = param[1] * x[1] - param[2] * x[2];
}

/* ODE2: SIRS_period.cpp:174-193 (same RK variant as ODE_rk45). */
void orc_sirs_ode2(const double *initial, const double *t, int n, const double *param, double period,
                   double *OdeTraj) {
    const int p = 3;
    double dt = t[1] - t[0], X0[3], X1[3], k1[3], k2[3], k3[3], k4[3], xs[3];
    for (int i = 0; i < n; i++) OdeTraj[i] = t[i];
    for (int j = 0; j < p; j++) {
        OdeTraj[(size_t)(j + 1) * n] = initial[j];
        X1[j] = initial[j];
    }
    for (int i = 1; i < n; i++) {
        for (int j = 0; j < p; j++) X0[j] = X1[j];
        sirs_rhs(X0, param, t[i - 1], period, k1);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k1[j] * dt / 2;
        sirs_rhs(xs, param, t[i - 1], period, k2);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k2[j] * dt / 2;
        sirs_rhs(xs, param, t[i - 1], period, k3);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k3[j] * dt / 2;
        sirs_rhs(xs, param, t[i - 1], period, k4);
        for (int j = 0; j < p; j++) {
            X1[j] = X0[j] + (k1[j] / 6 + k2[j] / 3 + k3[j] / 3 + k4[j] / 6) * dt;
            OdeTraj[i + (size_t)(j + 1) * n] = X1[j];
        }
    }
}

/* SIRS_IntSigma :50-79 + SIRS_KOM_Filter :147-168 (global dt; A = [-1 1 0; 0 -1 1; 1 0 -1]). */
void orc_sirs_kom_filter(const double *OdeTraj, int nrow, const double *param, int gridsize, double period,
                         double *Acube, double *Scube) {
    const int p = 3;
    int k = (nrow - 1) / gridsize;
    double dt = OdeTraj[1] - OdeTraj[0];
    double th[3] = {param[0], param[1], param[2]}, alpha = param[3];
    for (int seg = 0; seg < k; seg++) {
        double F0[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Sg[9] = {0}, F[9], Q[9], T1[9], T2[9], FF0[9];
        for (int i = 0; i < gridsize; i++) {
            int r = seg * gridsize + i;
            double t = OdeTraj[r], X = OdeTraj[r + (size_t)nrow], Y = OdeTraj[r + (size_t)2 * nrow],
                   Z = OdeTraj[r + (size_t)3 * nrow];
            double th1 = th[0] * (1 + alpha * sin(t / period * 2 * ORC_PI));
            double h0 = th1 * X * Y, h1 = th[1] * Y, h2 = th[2] * Z;
            /* F (col-major), SIRS_Fm_LNA_period :6-14 */
            F[0] = -th1 * Y; F[3] = -th1 * X;        F[6] = th[2];
            F[1] = th1 * Y;  F[4] = th1 * X - th[1]; F[7] = 0;
            F[2] = 0;        F[5] = th[1];           F[8] = -th[2];
            /* A' H A */
            Q[0] = h0 + h2; Q[3] = -h0;     Q[6] = -h2;
            Q[1] = -h0;     Q[4] = h0 + h1; Q[7] = -h1;
            Q[2] = -h2;     Q[5] = -h1;     Q[8] = h1 + h2;
            mat_mul(F, F0, p, FF0);
            for (int e = 0; e < 9; e++) F0[e] = F0[e] + FF0[e] * dt;
            mat_mul_bt(Sg, F, p, T1);
            mat_mul(F, Sg, p, T2);
            for (int e = 0; e < 9; e++) Sg[e] = Sg[e] + ((T1[e] + T2[e]) + Q[e]) * dt;
        }
        memcpy(Acube + (size_t)seg * 9, F0, sizeof F0);
        memcpy(Scube + (size_t)seg * 9, Sg, sizeof Sg);
    }
}

/* DegenerateDet3 :17-19 and PseudoInverse :22-48 of the rank-2 covariance. */
static double degenerate_det3(const double *M) {
    return (M[4] - M[3]) * (M[8] - M[6]) - (M[7] - M[3]) * (M[7] - M[6]);
}
static void pseudo_inverse3(const double *M, double *out) {
    /* M(r,c) = M[r + 3c] */
    double m01 = M[3], m02 = M[6], m11 = M[4], m12 = M[7], m22 = M[8];
    double tr = m11 + m22 - m01 - m02;
    double delta = tr * tr - 4 * ((m11 - m01) * (m22 - m02) - (m12 - m01) * (m12 - m02));
    double l1 = (tr + sqrt(delta)) / 2, l2 = (tr - sqrt(delta)) / 2;
    double e1[3], e2[3];
    e1[1] = (m12 - m01) / (m11 - m01 - l1); e1[0] = 1 - e1[1]; e1[2] = -1;
    e2[1] = (m12 - m01) / (m11 - m01 - l2); e2[0] = 1 - e2[1]; e2[2] = -1;
    double n1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
    double n2 = sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
    for (int i = 0; i < 3; i++) {
        e1[i] /= n1;
        e2[i] /= n2;
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) out[r + 3 * c] = e1[r] * (1 / l1) * e1[c] + e2[r] * (1 / l2) * e2[c];
}

/* log_like_trajSIRS :82-118 (mode 0) and Traj_sim_SIRS :196-240 (mode 1, normals step-major 3 per step). */
int orc_sirs_centred(int mode, const double *X_or_normals, const double *initial, const double *OdeTraj, int nrow,
                     const double *Acube, const double *Scube, double t_correct, double *traj_out, double *loglik) {
    const int p = 3;
    int k = nrow - 1;
    double ll = 0.0, PI3[9], v[3], w[3];
    if (mode == 0) {
        const double *X = X_or_normals;
        double d0[3], d1[3];
        for (int j = 0; j < p; j++) d1[j] = X[(size_t)(j + 1) * nrow] - OdeTraj[(size_t)(j + 1) * nrow];
        for (int i = 0; i < k; i++) {
            const double *A = Acube + (size_t)i * 9, *S = Scube + (size_t)i * 9;
            for (int j = 0; j < p; j++) {
                d0[j] = d1[j];
                d1[j] = X[(i + 1) + (size_t)(j + 1) * nrow] - OdeTraj[(i + 1) + (size_t)(j + 1) * nrow];
            }
            if (X[i + 1] <= t_correct) {
                pseudo_inverse3(S, PI3);
                for (int r = 0; r < p; r++) {
                    double ax = 0.0;
                    for (int q = 0; q < p; q++) ax += A[r + q * p] * d0[q];
                    v[r] = d1[r] - ax;
                }
                for (int r = 0; r < p; r++) w[r] = PI3[r] * v[0] + PI3[r + 3] * v[1] + PI3[r + 6] * v[2];
                ll += -log(degenerate_det3(S)) * 0.5 - 0.5 * (v[0] * w[0] + v[1] * w[1] + v[2] * w[2]);
            }
        }
    } else {
        double X0[3], X1[3], e0[3], e1[3], noise[3], L[9];
        for (int j = 0; j < p; j++) {
            X1[j] = initial[j];
            e1[j] = initial[j];
            traj_out[(size_t)(j + 1) * nrow] = initial[j];
        }
        traj_out[0] = 0;
        for (int i = 0; i < k; i++) {
            const double *A = Acube + (size_t)i * 9, *S = Scube + (size_t)i * 9;
            for (int j = 0; j < p; j++) {
                X0[j] = X1[j];
                e0[j] = e1[j];
                e1[j] = OdeTraj[(i + 1) + (size_t)(j + 1) * nrow];
            }
            if (orc_chol_lower(S, 3, 0.0000000000001, L)) return ORC_CHOL_FAIL; /* mvrnormArma2 */
            for (int r = 0; r < 3; r++) {
                double s2 = 0.0;
                for (int q = 0; q < 3; q++) s2 += L[r + q * 3] * X_or_normals[(size_t)i * 3 + q];
                noise[r] = s2;
            }
            for (int r = 0; r < p; r++) {
                double ax = 0.0;
                for (int q = 0; q < p; q++) ax += A[r + q * p] * (X0[q] - e0[q]);
                X1[r] = ax + e1[r] + noise[r];
            }
            if (OdeTraj[i + 1] <= t_correct) {
                pseudo_inverse3(S, PI3);
                for (int r = 0; r < p; r++) w[r] = PI3[r] * noise[0] + PI3[r + 3] * noise[1] + PI3[r + 6] * noise[2];
                ll += -log(degenerate_det3(S)) * 0.5 + (-0.5) * (noise[0] * w[0] + noise[1] * w[1] + noise[2] * w[2]);
            }
            traj_out[i + 1] = OdeTraj[i + 1];
            for (int j = 0; j < p; j++) traj_out[(i + 1) + (size_t)(j + 1) * nrow] = X1[j];
        }
    }
    *loglik = ll;
    return ORC_OK;
}

/* ---------------------------------------------------------------- legacy copies of the pipeline
 * src/LNA_NS_general.cpp (the SIR-only predecessor of LNA_functional.cpp), class Foo of src/generalLNA.cpp and
 * ESlice_SIRS of src/SIRS_period.cpp.  x_i has two entries here: (nch, nparam); param = (R0, gamma, .., ch_1..). */

static orc_cfg legacy_cfg(const double *x_r, const int *x_i2, int nparam_total) {
    orc_cfg c;
    c.model = ORC_SIR;
    c.transX = ORC_STANDARD;
    c.p = 2;
    c.nparam_total = nparam_total;
    c.x_r = x_r;
    c.x_i[0] = x_i2[0];
    c.x_i[1] = x_i2[1];
    c.x_i[2] = 0;
    c.x_i[3] = 1;
    return c;
}

/* betafs: LNA_NS_general.cpp:25-47 (betaf :6-21 is its single-time form) = betaTs with R0 = param[0], gamma = param[1] */
void orc_general_betafs(const double *ts, int m, const double *param, const double *x_r, const int *x_i2, double *betas) {
    int x_i[4] = {x_i2[0], x_i2[1], 0, 1};
    orc_betaTs(param, ts, m, x_r, x_i, betas);
}
/* ODE_general_one :52-62 (mode 0: 2 values) and general_F :66-72 (mode 1: 2 x 2 column-major) */
void orc_general_point(int mode, const double *states, const double *param, int nparam_total, double t, const double *x_r,
                       const int *x_i2, double *out) {
    orc_cfg c = legacy_cfg(x_r, x_i2, nparam_total);
    double th[3];
    thetas3(&c, t, param, th);
    if (mode == 0) {
        out[0] = -th[0] * states[0] * states[1];
        out[1] = th[0] * states[0] * states[1] - param[1] * states[1];
    } else {
        out[0] = -th[0] * states[1];
        out[1] = th[0] * states[1];
        out[2] = -th[0] * states[0];
        out[3] = th[0] * states[0] - param[1];
    }
}
/* General_ODE_rk45 :76-98: the same stages as ODE_rk45 (all at t[i-1], k4 at dt/2) */
void orc_general_ode_rk45(const double *initial, const double *t, int n, const double *param, int nparam_total,
                          const double *x_r, const int *x_i2, double *OdeTraj) {
    orc_cfg c = legacy_cfg(x_r, x_i2, nparam_total);
    orc_ode_rk45(&c, initial, t, n, param, OdeTraj);
}
/* General_IntSigmaF :103-135 (SigmaF without the Q factors, which are the identity on the standard scale) */
void orc_general_intsigmaf(const double *Traj_par, int nrow, const double *param, int nparam_total, const double *x_r,
                           const int *x_i2, double *expF, double *Sigma) {
    orc_cfg c = legacy_cfg(x_r, x_i2, nparam_total);
    orc_sigmaF(&c, Traj_par, nrow, nrow, param, expF, Sigma);
}
/* General_KOM_Filter :139-159 */
void orc_general_kom_filter(const double *OdeTraj, int nrow, const double *param, int nparam_total, int gridsize,
                            const double *x_r, const int *x_i2, double *Acube, double *Scube) {
    orc_cfg c = legacy_cfg(x_r, x_i2, nparam_total);
    orc_kf_param(&c, OdeTraj, nrow, param, gridsize, Acube, Scube);
}

/* log_like_traj_general :164-200 (mode 0) and Traj_sim_general :205-250 (mode 1; normals step-major, 2 per step):
 * p = 2, the density through inv2(Sigma) (basic_util.cpp:5-14), noise = mvrnormArma(2, Sigma) = chols(Sigma) z.
 * Foo::LNA_log_like_traj / Foo::LNA_Traj_sim (generalLNA.cpp:178-214, 217-262) are the same text. */
int orc_legacy_centred(int mode, const double *X_or_normals, const double *OdeTraj, int nrow, const double *Acube,
                       const double *Scube, double t_correct, double *traj_out, double *loglik) {
    const int p = 2;
    int k = nrow - 1;
    double ll = 0.0, SI[4];
    if (mode == 0) {
        const double *X = X_or_normals;
        double d0[2], d1[2], v[2];
        for (int j = 0; j < p; j++) d1[j] = X[(size_t)(j + 1) * nrow] - OdeTraj[(size_t)(j + 1) * nrow];
        for (int i = 0; i < k; i++) {
            const double *A = Acube + (size_t)i * 4, *S = Scube + (size_t)i * 4;
            orc_inv2(S, SI);
            for (int j = 0; j < p; j++) {
                d0[j] = d1[j];
                d1[j] = X[(i + 1) + (size_t)(j + 1) * nrow] - OdeTraj[(i + 1) + (size_t)(j + 1) * nrow];
            }
            if (X[i + 1] <= t_correct) {
                for (int r = 0; r < p; r++) v[r] = d1[r] - (A[r] * d0[0] + A[r + 2] * d0[1]);
                /* (v' SigInv) v */
                double w0 = v[0] * SI[0] + v[1] * SI[1], w1 = v[0] * SI[2] + v[1] * SI[3];
                ll += -log(S[0] * S[3] - S[2] * S[1]) / 2.0 - 0.5 * (w0 * v[0] + w1 * v[1]);
            }
        }
    } else {
        double X0[2], X1[2], e0[2], e1[2], noise[2], L[4];
        for (int j = 0; j < p; j++) {
            X1[j] = OdeTraj[(size_t)(j + 1) * nrow];
            e1[j] = X1[j];
            traj_out[(size_t)(j + 1) * nrow] = X1[j];
        }
        traj_out[0] = 0;
        for (int i = 0; i < k; i++) {
            const double *A = Acube + (size_t)i * 4, *S = Scube + (size_t)i * 4;
            for (int j = 0; j < p; j++) {
                X0[j] = X1[j];
                e0[j] = e1[j];
                e1[j] = OdeTraj[(i + 1) + (size_t)(j + 1) * nrow];
            }
            orc_chols(S, L);
            for (int r = 0; r < p; r++) noise[r] = L[r] * X_or_normals[(size_t)i * 2] + L[r + 2] * X_or_normals[(size_t)i * 2 + 1];
            for (int r = 0; r < p; r++) X1[r] = (A[r] * (X0[0] - e0[0]) + A[r + 2] * (X0[1] - e0[1])) + e1[r] + noise[r];
            if (OdeTraj[i + 1] <= t_correct) {
                orc_inv2(S, SI);
                /* ((-0.5) noise') inv2(Sig) noise */
                double a0 = (-0.5) * noise[0], a1 = (-0.5) * noise[1];
                double w0 = a0 * SI[0] + a1 * SI[1], w1 = a0 * SI[2] + a1 * SI[3];
                ll += -log(S[0] * S[3] - S[2] * S[1]) / 2 + (w0 * noise[0] + w1 * noise[1]);
            }
            traj_out[i + 1] = OdeTraj[i + 1];
            for (int j = 0; j < p; j++) traj_out[(i + 1) + (size_t)(j + 1) * nrow] = X1[j];
        }
    }
    *loglik = ll;
    return ORC_OK;
}

/* Traj_sim_ezG :257-314 = General_ODE_rk45 + coarse rows + General_KOM_Filter + the loop of Traj_sim_general started from
 * `initial` (k = length(times) / gridsize). */
int orc_traj_sim_ezG(const double *initial, const double *t, int nt, const double *param, int nparam_total, int gridsize,
                     const double *x_r, const int *x_i2, double t_correct, const double *normals, double *LNA_traj,
                     double *loglike) {
    const int p = 2;
    int k = nt / gridsize, nrow = k + 1;
    double *thin = (double *)malloc(sizeof(double) * (size_t)nt * (p + 1));
    double *coarse = (double *)malloc(sizeof(double) * (size_t)nrow * (p + 1));
    double *A = (double *)malloc(sizeof(double) * (size_t)p * p * k * 2), *S = A + (size_t)p * p * k;
    orc_general_ode_rk45(initial, t, nt, param, nparam_total, x_r, x_i2, thin);
    for (int i = 0; i < nrow; i++)
        for (int j = 0; j <= p; j++) coarse[i + (size_t)j * nrow] = thin[(size_t)i * gridsize + (size_t)j * nt];
    orc_general_kom_filter(thin, nt, param, nparam_total, gridsize, x_r, x_i2, A, S);
    /* (row 0 of the coarse ODE is `initial`, so the loop is Traj_sim_general's) */
    int st = orc_legacy_centred(1, normals, coarse, nrow, A, S, t_correct, LNA_traj, loglike);
    free(thin);
    free(coarse);
    free(A);
    return st;
}

/* volz_loglik_nh: SIR_phylodyn.cpp:517-560.  f1 is a LOG trajectory (LogTraj): columns 1, 2 = log S, log I; cell i of the
 * genealogy uses row L - i, L = first row with time >= t_correct; -1e7 when a LOG state of rows 0..n0 is negative. */
double orc_volz_loglik_nh(const orc_init *init, const double *f1, int nrow, const double *betaN, double t_correct) {
    int n0 = init->ng - 1, L = 0;
    while (f1[L] < t_correct) L++;
    for (int j = 1; j <= 2; j++)
        for (int r = 0; r <= n0; r++)
            if (f1[r + (size_t)j * nrow] < 0) return ORC_SENTINEL;
    int E = init->E, start = 0;
    double *ll = (double *)malloc(sizeof(double) * (size_t)(E > 0 ? E : 1));
    for (int i = 0; i < n0; i++) {
        int row = L - i, reps = (int)init->gridrep[i];
        double f = f1[row + (size_t)2 * nrow], s = f1[row + (size_t)nrow], b = betaN[row];
        for (int j = 0; j < reps; j++) {
            ll[start] = -2 * (init->C[start] * init->D[start] * b * exp(-f) * exp(s)) + init->y[start] * (log(b) - f + s);
            start++;
        }
    }
    double res = accu2(ll, start);
    free(ll);
    return res;
}

/* ESlice_general (LNA_NS_general.cpp:320-394; kind 0), Foo::LNA_ESlice (generalLNA.cpp:296-372, the same text; kind 2) and
 * ESlice_SIRS (SIRS_period.cpp:313-402; kind 1): the trajectory slice update in the CENTRED parametrisation, `reps` times.
 *   proposal v: Traj_sim_general(OdeTraj, FTs) [kinds 0, 2; p = 2]  /  Traj_sim_SIRS(state, OdeTraj, FTs) [kind 1; p = 3]
 *   likelihood: volz_loglik_nh(LogTraj(.), betaN) or coal_loglik(LogTraj(.), lambda)
 *   kind 1 computes betaN = betaDyn(params4[0], params4[1], times, params4[2]) * params4[3] (SIR_phylodyn.cpp:40-46)
 *   20-shrink cap: kinds 0, 2 return to f_cur; kind 1 just leaves the loop, the LAST candidate stays (:377-380)
 * normals: reps x k*p (step-major), uniforms: consumed contiguously (u, theta_1, one per shrink; per repetition). */
int orc_eslice_legacy(const orc_init *init, int kind, const double *f_cur0, int nrow, int p, const double *OdeTraj,
                      const double *Acube, const double *Scube, const double *state, const double *betaN_or_params4,
                      double t_correct, double lambda, int reps, int volz, const double *normals, const double *uniforms,
                      int *n_unif, int *n_evals, double *newTraj) {
    int iu = 0, evals = 0, n = nrow * p, k = nrow - 1, status = 0;
    size_t sz = (size_t)nrow * (p + 1);
    double *v = (double *)malloc(sizeof(double) * sz * 4), *f_cur = v + sz, *lg = f_cur + sz, *bet = lg + sz;
    memcpy(f_cur, f_cur0, sizeof(double) * sz);
    if (kind == 1) {
        const double *q = betaN_or_params4;
        for (int r = 0; r < nrow; r++) bet[r] = q[0] * (1 + q[1] * sin(2 * ORC_PI * f_cur[r] / q[2])) * q[3];
    } else {
        memcpy(bet, betaN_or_params4, sizeof(double) * (size_t)nrow);
    }
#define LEGACY_LL(T) (log_traj((T), nrow, p, lg), volz ? orc_volz_loglik_nh(init, lg, nrow, bet, t_correct) \
                                                       : orc_coal_loglik(init, lg, nrow, p, t_correct, lambda))
    for (int count = 0; count < reps; count++) {
        double ll_sim;
        int st = kind == 1 ? orc_sirs_centred(1, normals + (size_t)count * k * 3, state, OdeTraj, nrow, Acube, Scube, t_correct, v, &ll_sim)
                           : orc_legacy_centred(1, normals + (size_t)count * k * 2, OdeTraj, nrow, Acube, Scube, t_correct, v, &ll_sim);
        if (st) {
            free(v);
            return st;
        }
        double u = uniforms[iu++];
        double logy = LEGACY_LL(f_cur) + log(u);
        evals++;
        double theta = runif_ab(0, 2 * ORC_PI, uniforms[iu++]);
        double theta_min = theta - 2 * ORC_PI, theta_max = theta;
        for (int r = 0; r < nrow; r++) newTraj[r] = f_cur[r];
        double ct = cos(theta), sn = sin(theta);
        for (int e = 0; e < n; e++) {
            double fc = f_cur[nrow + e] - OdeTraj[nrow + e], vt = v[nrow + e] - OdeTraj[nrow + e];
            newTraj[nrow + e] = (fc * ct + vt * sn) + OdeTraj[nrow + e];
        }
        double loglike = LEGACY_LL(newTraj);
        evals++;
        int i = 0;
        for (;;) {
            double mn = newTraj[nrow];
            for (int e = nrow; e < nrow * (p + 1); e++)
                if (newTraj[e] < mn) mn = newTraj[e];
            if (!(mn < 0 || loglike <= logy)) break;
            i += 1;
            if (i > 20) {
                if (kind != 1) memcpy(newTraj, f_cur, sizeof(double) * sz);
                status |= ORC_CAP_HIT;
                break;
            }
            if (theta < 0) theta_min = theta;
            else theta_max = theta;
            theta = runif_ab(theta_min, theta_max, uniforms[iu++]);
            ct = cos(theta);
            sn = sin(theta);
            for (int e = 0; e < n; e++) {
                double fc = f_cur[nrow + e] - OdeTraj[nrow + e], vt = v[nrow + e] - OdeTraj[nrow + e];
                newTraj[nrow + e] = (fc * ct + vt * sn) + OdeTraj[nrow + e];
            }
            loglike = LEGACY_LL(newTraj);
            evals++;
        }
        memcpy(f_cur, newTraj, sizeof(double) * sz);
    }
#undef LEGACY_LL
    *n_unif = iu;
    *n_evals = evals;
    free(v);
    return status;
}

/* class Foo (generalLNA.cpp:9-372): a reaction network with hard-wired SIR-with-immigration hazards
 *   h = (param0 * param1 / N * S * I, param1 * I, param2 * N)  (rate_h :76-83, transform_param :50-55)
 * and effect matrix A = A_post - A_pre (m = 3 reactions x p = 2 species, column-major).
 *   mode 0: rate_h (3), 1: dh (3 x 2, :84-90), 2: ODE_ones = A' h (2, :96-98), 3: A' dh (2 x 2: what F_vec :92-94 computes
 *   before its arma::vec return type rejects it) */
void orc_foo_point(int mode, const double *A_pre, const double *A_post, const double *param, double N, const double *states,
                   double *out) {
    double pn0 = param[0] * param[1] / N, h[3], dh[6], A[6];
    for (int e = 0; e < 6; e++) A[e] = A_post[e] - A_pre[e];
    h[0] = pn0 * states[0] * states[1];
    h[1] = param[1] * states[1];
    h[2] = param[2] * N;
    dh[0] = pn0 * states[1]; dh[3] = pn0 * states[0];
    dh[1] = 0;               dh[4] = param[1];
    dh[2] = 0.0;             dh[5] = 0.0;
    if (mode == 0) {
        memcpy(out, h, sizeof h);
    } else if (mode == 1) {
        memcpy(out, dh, sizeof dh);
    } else if (mode == 2) {
        for (int j = 0; j < 2; j++) out[j] = A[3 * j] * h[0] + A[3 * j + 1] * h[1] + A[3 * j + 2] * h[2];
    } else {
        for (int c = 0; c < 2; c++)
            for (int j = 0; j < 2; j++) out[j + 2 * c] = A[3 * j] * dh[3 * c] + A[3 * j + 1] * dh[3 * c + 1] + A[3 * j + 2] * dh[3 * c + 2];
    }
}
/* Foo::ODE_rk45 :100-122 */
void orc_foo_ode_rk45(const double *A_pre, const double *A_post, const double *param, double N, const double *initial,
                      const double *t, int n, double *OdeTraj) {
    const int p = 2;
    double dt = t[1] - t[0];
    double X0[2], X1[2], k1[2], k2[2], k3[2], k4[2], xs[2];
    for (int i = 0; i < n; i++) OdeTraj[i] = t[i];
    for (int j = 0; j < p; j++) {
        OdeTraj[(size_t)(j + 1) * n] = initial[j];
        X1[j] = initial[j];
    }
    for (int i = 1; i < n; i++) {
        for (int j = 0; j < p; j++) X0[j] = X1[j];
        orc_foo_point(2, A_pre, A_post, param, N, X0, k1);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k1[j] * dt / 2;
        orc_foo_point(2, A_pre, A_post, param, N, xs, k2);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k2[j] * dt / 2;
        orc_foo_point(2, A_pre, A_post, param, N, xs, k3);
        for (int j = 0; j < p; j++) xs[j] = X0[j] + k3[j] * dt / 2;
        orc_foo_point(2, A_pre, A_post, param, N, xs, k4);
        for (int j = 0; j < p; j++) {
            X1[j] = X0[j] + (k1[j] / 6 + k2[j] / 3 + k3[j] / 3 + k4[j] / 6) * dt;
            OdeTraj[i + (size_t)(j + 1) * n] = X1[j];
        }
    }
}

/* ---------------------------------------------------------------- data simulators */

/* volz_thin_sir: R/coal_simulation.R:11-70.  Coalescent times by thinning against the Volz rate on an SIR trajectory
 * (t, S, I on a fine grid, betaN on the same grid).  R's rexp / runif are replaced by a pre-drawn stream indexed by
 * PROPOSAL number: proposal j waits -log(ua[j]) / rate and, if it is not overtaken by the next sampling time, is accepted
 * when ub[j] <= lower / thed (ub[j] is skipped otherwise, like the short-circuited runif(1) of the reference).
 * Returns 0, or 1 when the stream (n_avail proposals) or the output (max_coal) ran out. */
int orc_volz_thin_sir(const double *t, const double *S, const double *I, const double *betaN, int n, double t_correct,
                      const double *samp_times, const int *n_sampled, int ns, double lower, const double *ua,
                      const double *ub, long n_avail, double *coal_times, int *lineages, int max_coal, int *n_coal,
                      long *n_draws) {
    int first_false = n;
    for (int r = 0; r < n; r++)
        if (!(t_correct - t[r] >= 0)) {
            first_false = r;
            break;
        }
    int i = first_false >= n ? 0 : (first_false - 1 > 0 ? first_false - 1 : 0);
    double max_samp = samp_times[0];
    for (int s = 1; s < ns; s++)
        if (samp_times[s] > max_samp) max_samp = samp_times[s];
    int curr = 0, active = n_sampled[0], ncoal = 0, status = 0;
    double time = samp_times[0];
    long draw = 0;
    while (time <= max_samp || active > 1) {
        if (active == 1) {
            if (curr + 1 >= ns) {
                status = 1;
                break;
            }
            curr = curr + 1;
            active = active + n_sampled[curr];
            time = samp_times[curr];
        }
        if (draw >= n_avail || ncoal >= max_coal) {
            status = 1;
            break;
        }
        time = time + (-log(ua[draw])) / (0.5 * active * (active - 1) / lower);
        while (time > t_correct - t[i]) {
            i = i - 1;
            if (i < 0) {
                i = 0;
                break;
            }
        }
        double thed = 1.0 / (betaN[i] * S[i] / (I[i] / 2.0));
        if (curr < ns - 1 && time >= samp_times[curr + 1]) {
            curr = curr + 1;
            active = active + n_sampled[curr];
            time = samp_times[curr];
        } else if (ub[draw] <= lower / thed) {
            time = time < t_correct ? time : t_correct;
            coal_times[ncoal] = time;
            lineages[ncoal] = active;
            ncoal++;
            active = active - 1;
        }
        draw++;
    }
    *n_coal = ncoal;
    *n_draws = draw;
    return status;
}

/* ---------------------------------------------------------------- batch driver (bench baseline) */

/* B independent FULL evaluations, item-major inputs like the product's C ABI; one core.
 * Used by bench.py (cpu_baseline / --impl reference) and by the parity tests. */
int orc_full_loglik_batch(const orc_cfg *c, const orc_init *init, long B, const double *param,
                          const double *initial, const double *t, int nt, int gridsize,
                          const double *OriginTraj, double t_correct, double *coalLog, int *status) {
    int p = c->p, nrow = nt / gridsize + 1, k = (nt - 1) / gridsize;
    double *A = (double *)malloc(sizeof(double) * (size_t)(p * p * k) * 2);
    double *ode = (double *)malloc(sizeof(double) * (size_t)nrow * (p + 1) * 2);
    double *bn = (double *)malloc(sizeof(double) * (size_t)nrow);
    for (long b = 0; b < B; b++) {
        int st = orc_full_loglik(c, init, param + b * c->nparam_total, initial + b * p, t, nt, gridsize,
                                 OriginTraj + b * (long)k * p, t_correct, A, A + p * p * k, ode, bn,
                                 ode + (size_t)nrow * (p + 1), &coalLog[b]);
        if (status) status[b] = st;
    }
    free(A);
    free(ode);
    free(bn);
    return 0;
}
