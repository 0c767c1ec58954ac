/*
 * pnp_oracle.c -- CPU restatement of the reference PnP solve path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for pnp_solver_test_b200.  It restates, in plain scalar C
 * (FP64, one problem at a time), the algorithms of the reference's
 * scripts/PNP_SOLVER_LIB.py and scripts/TEST_TOOLBOX.py that lie on the hot path
 * (SURVEY.md section 8a).  Every function cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (pnp_solver_test_b200/) never does.
 *
 * Parity status: PINNED.  The reference is pure Python and imports in the build
 * container; oracle/make_golden.py runs the unmodified reference on seeded inputs and
 * commits its outputs under tests/golden/, and tests/test_oracle_golden.py checks this
 * file against them.  The reference itself ships no solver golden vectors; its only
 * fixture (scripts/ground_truth_R_t/GT_R_t_dict.pkl) pins Euler->R and is checked too.
 *
 * Numerics deliberately mirror the reference where it matters for the outcome:
 *  - np.linalg.pinv is restated as an SVD pseudo-inverse with the rcond = 1e-15 cut-off
 *    (one-sided Jacobi SVD instead of LAPACK gesdd; same mathematical object);
 *  - the halved constraint Jacobians of EKF2_get_hx_H are reproduced, not fixed;
 *  - res_norm is the residual of the state BEFORE the last update.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/* parameters (defaults = the reference's inline constants)                              */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    int32_t max_it;        /* PNP_SOLVER_LIB.py:2635, :2857  num_it = 14                 */
    int32_t linear_it;     /* :223, :762                    num_it = 3                  */
    double  lm_lambda;     /* :2631                         10**-5                      */
    double  exit_tol;      /* :2952                         1e-2                        */
    double  f_weight;      /* :2845                         225.68 (NOT K[0,0])         */
    double  meas_sigma_px; /* :2846                         3.0                         */
    double  proc_q;        /* :2792                         1e-1 (quaternion states)    */
    double  proc_d;        /* :2793                         1e-2 (delta states)         */
    double  omega0;        /* :2836                         1e-5                        */
    double  res_old0;      /* :2863                         1e-7                        */
} oracle_params_t;

ORACLE_API void pnp_oracle_default_params(oracle_params_t *p)
{
    p->max_it = 14; p->linear_it = 3;
    p->lm_lambda = 1e-5; p->exit_tol = 1e-2;
    p->f_weight = 225.68; p->meas_sigma_px = 3.0;
    p->proc_q = 1e-1; p->proc_d = 1e-2; p->omega0 = 1e-5; p->res_old0 = 1e-7;
}

/* ------------------------------------------------------------------------------------ */
/* small dense linear algebra                                                            */
/* ------------------------------------------------------------------------------------ */

/* One-sided (Hestenes) Jacobi SVD of a row-major m x n matrix, m >= n.
 * On return: U (m x n, orthonormal columns where s > 0), s (n, descending), V (n x n).
 * Stands in for the LAPACK SVD behind np.linalg.svd / np.linalg.pinv. */
static void svd_jacobi(int m, int n, const double *A, double *U, double *s, double *V)
{
    int i, j, k, sweep;
    memcpy(U, A, sizeof(double) * (size_t)m * n);
    for (i = 0; i < n; ++i)
        for (j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (sweep = 0; sweep < 60; ++sweep) {
        int rotated = 0;
        for (i = 0; i < n - 1; ++i) {
            for (j = i + 1; j < n; ++j) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
                for (k = 0; k < m; ++k) {
                    double ui = U[k * n + i], uj = U[k * n + j];
                    alpha += ui * ui; beta += uj * uj; gamma += ui * uj;
                }
                if (gamma == 0.0) continue;
                if (fabs(gamma) <= 2e-16 * sqrt(alpha * beta)) continue;
                rotated = 1;
                {
                    double zeta = (beta - alpha) / (2.0 * gamma);
                    double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                    for (k = 0; k < m; ++k) {
                        double ui = U[k * n + i], uj = U[k * n + j];
                        U[k * n + i] = c * ui - sn * uj;
                        U[k * n + j] = sn * ui + c * uj;
                    }
                    for (k = 0; k < n; ++k) {
                        double vi = V[k * n + i], vj = V[k * n + j];
                        V[k * n + i] = c * vi - sn * vj;
                        V[k * n + j] = sn * vi + c * vj;
                    }
                }
            }
        }
        if (!rotated) break;
    }
    for (j = 0; j < n; ++j) {
        double nrm = 0.0;
        for (k = 0; k < m; ++k) nrm += U[k * n + j] * U[k * n + j];
        s[j] = sqrt(nrm);
        if (s[j] > 0.0) for (k = 0; k < m; ++k) U[k * n + j] /= s[j];
    }
    /* sort descending (selection sort on columns) */
    for (i = 0; i < n - 1; ++i) {
        int p = i;
        for (j = i + 1; j < n; ++j) if (s[j] > s[p]) p = j;
        if (p != i) {
            double tmp = s[i]; s[i] = s[p]; s[p] = tmp;
            for (k = 0; k < m; ++k) { tmp = U[k * n + i]; U[k * n + i] = U[k * n + p]; U[k * n + p] = tmp; }
            for (k = 0; k < n; ++k) { tmp = V[k * n + i]; V[k * n + i] = V[k * n + p]; V[k * n + p] = tmp; }
        }
    }
}

/* np.linalg.pinv(A) with the default rcond = 1e-15: singular values <= rcond * s_max are
 * dropped.  A is m x n row-major (m >= n); out is n x m row-major.  work: m*n + n + n*n. */
static void pinv_svd(int m, int n, const double *A, double *out, double *work)
{
    double *U = work, *s = work + (size_t)m * n, *V = s + n;
    int i, j, k;
    double cutoff;
    svd_jacobi(m, n, A, U, s, V);
    cutoff = 1e-15 * s[0];
    for (i = 0; i < n; ++i)
        for (j = 0; j < m; ++j) {
            double acc = 0.0;
            for (k = 0; k < n; ++k)
                if (s[k] > cutoff) acc += V[i * n + k] * (1.0 / s[k]) * U[j * n + k];
            out[i * m + j] = acc;
        }
}

static double det3(const double *M)
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
           M[2] * (M[3] * M[7] - M[4] * M[6]);
}

/* np.linalg.inv for the 3x3 camera matrix (PNP_SOLVER_LIB.py:2596, :2811) */
static void inv3(const double *M, double *out)
{
    double d = det3(M), id = 1.0 / d;
    out[0] = (M[4] * M[8] - M[5] * M[7]) * id;
    out[1] = (M[2] * M[7] - M[1] * M[8]) * id;
    out[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    out[3] = (M[5] * M[6] - M[3] * M[8]) * id;
    out[4] = (M[0] * M[8] - M[2] * M[6]) * id;
    out[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    out[6] = (M[3] * M[7] - M[4] * M[6]) * id;
    out[7] = (M[1] * M[6] - M[0] * M[7]) * id;
    out[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}

#define RAD2DEG (180.0 / M_PI)
#define DEG2RAD (M_PI / 180.0)

/* ------------------------------------------------------------------------------------ */
/* Euler <-> R                                                                            */
/* ------------------------------------------------------------------------------------ */

/* get_rotation_matrix_from_Euler, PNP_SOLVER_LIB.py:4442-4472.  Angles in degrees when
 * is_degree != 0.  R = E_roll @ E_pitch @ E_yaw with the yaw sign flipped on input. */
ORACLE_API void pnp_oracle_R_from_euler(double roll, double yaw, double pitch, int is_degree, double *R)
{
    double c1, s1, c2, s2, c3, s3;
    yaw = -yaw;                                   /* :4451 */
    if (is_degree) { roll *= DEG2RAD; yaw *= DEG2RAD; pitch *= DEG2RAD; }
    c1 = cos(roll); s1 = sin(roll); c2 = cos(yaw); s2 = sin(yaw); c3 = cos(pitch); s3 = sin(pitch);
    /* E_roll = [[c1,s1,0],[-s1,c1,0],[0,0,1]]; E_pitch = [[1,0,0],[0,c3,s3],[0,-s3,c3]];
     * E_yaw = [[c2,0,-s2],[0,1,0],[s2,0,c2]]  (:4464-4466); product :4470 */
    {
        /* EP = E_pitch @ E_yaw */
        double ep[9] = { c2, 0.0, -s2,
                         s3 * s2, c3, s3 * c2,
                         c3 * s2, -s3, c3 * c2 };
        R[0] = c1 * ep[0] + s1 * ep[3]; R[1] = c1 * ep[1] + s1 * ep[4]; R[2] = c1 * ep[2] + s1 * ep[5];
        R[3] = -s1 * ep[0] + c1 * ep[3]; R[4] = -s1 * ep[1] + c1 * ep[4]; R[5] = -s1 * ep[2] + c1 * ep[5];
        R[6] = ep[6]; R[7] = ep[7]; R[8] = ep[8];
    }
}

/* get_Euler_from_rotation_matrix, PNP_SOLVER_LIB.py:4474-4517.  out = (roll, yaw, pitch). */
ORACLE_API void pnp_oracle_euler_from_R(const double *R, int is_degree, double *out)
{
    const double eps = 1e-7;                      /* :4480 */
    double th1, th2, th3;
    if (fabs(M_PI / 2.0 - asin(fabs(R[7]))) <= eps) {          /* :4482 gimbal lock */
        double sg = (-R[7] > 0.0) ? 1.0 : ((-R[7] < 0.0) ? -1.0 : 0.0);
        th2 = sg * (M_PI / 2.0);
        th3 = 0.0;
        th1 = atan2(-R[3], R[0]);
    } else {
        double c1, c3, c2;
        th1 = atan2(R[1], R[4]);                  /* :4490 */
        th3 = atan2(R[6], R[8]);                  /* :4491 */
        c1 = cos(th1); c3 = cos(th3);
        if (fabs(c1) > fabs(c3)) c2 = R[4] / c1;  /* :4495-4502 */
        else                     c2 = R[8] / c3;
        th2 = atan2(-R[7], c2);                   /* :4503 */
    }
    {
        double roll = th1, pitch = th2, yaw = th3;
        if (is_degree) { roll *= RAD2DEG; yaw *= RAD2DEG; pitch *= RAD2DEG; }
        yaw = -yaw;                               /* :4515 */
        out[0] = roll; out[1] = yaw; out[2] = pitch;   /* :4517 order (roll, yaw, pitch) */
    }
}

/* ------------------------------------------------------------------------------------ */
/* projection (PNP_SOLVER_LIB.py:4532-4557)                                               */
/* ------------------------------------------------------------------------------------ */

/* rint() rounds half to even in the default rounding mode, like np.around (:4551). */
ORACLE_API void pnp_oracle_project(int n, const double *P, const double *K, const double *R,
                                   const double *t, int is_quantized, double q, double *uvw)
{
    int i, r;
    for (i = 0; i < n; ++i) {
        double X[3], ray[3], az;
        for (r = 0; r < 3; ++r)
            X[r] = R[r * 3 + 0] * P[i * 3 + 0] + R[r * 3 + 1] * P[i * 3 + 1] + R[r * 3 + 2] * P[i * 3 + 2] + t[r];
        for (r = 0; r < 3; ++r)
            ray[r] = K[r * 3 + 0] * X[0] + K[r * 3 + 1] * X[1] + K[r * 3 + 2] * X[2];
        az = fabs(ray[2]);                        /* :4548 divides by |z| */
        for (r = 0; r < 3; ++r) {
            double v = ray[r] / az;
            if (is_quantized) v = rint(v / q) * q;
            uvw[i * 3 + r] = v;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* packing / normalisation: f2_get_B_xy, PNP_SOLVER_LIB.py:3291-3312                      */
/* ------------------------------------------------------------------------------------ */
static void normalise(int n, const double *uv, const double *Kinv, double *bx, double *by)
{
    int i;
    for (i = 0; i < n; ++i) {
        double u = uv[2 * i], v = uv[2 * i + 1];
        bx[i] = Kinv[0] * u + Kinv[1] * v + Kinv[2] * 1.0;
        by[i] = Kinv[3] * u + Kinv[4] * v + Kinv[5] * 1.0;
    }
}

/* ------------------------------------------------------------------------------------ */
/* QEIF: solve_pnp_QEIF_single_pattern :2771-3025, QEKF_get_hx_H :3902-3983,              */
/*       QEKF_reconstruct_R_t_m1 :3542-3609                                               */
/* ------------------------------------------------------------------------------------ */
static void qekf_phi(const double *x, double *phi1, double *phi2, double *phi3, double *gamma_out)
{
    double qr = x[0], qi = x[1], qj = x[2], qk = x[3];
    double nq = sqrt(qr * qr + qi * qi + qj * qj + qk * qk);
    double gamma = nq * nq;                       /* (np.linalg.norm(q))**2, :3918 */
    double qii = qi * qi, qjj = qj * qj, qkk = qk * qk;
    double qij = qi * qj, qjk = qj * qk, qik = qi * qk;
    double qri = qr * qi, qrj = qr * qj, qrk = qr * qk;
    phi1[0] = gamma - 2 * (qjj + qkk); phi1[1] = 2 * (qij - qrk); phi1[2] = 2 * (qik + qrj);   /* :3934 */
    phi2[0] = 2 * (qij + qrk); phi2[1] = gamma - 2 * (qii + qkk); phi2[2] = 2 * (qjk - qri);   /* :3935 */
    phi3[0] = 2 * (qik - qrj); phi3[1] = 2 * (qjk + qri); phi3[2] = gamma - 2 * (qii + qjj);   /* :3936 */
    *gamma_out = gamma;
}

/* hx (2n) and H (2n x 6), rows 0..n-1 = x equations, n..2n-1 = y equations. */
static void qekf_hx_H(int n, const double *x, const double *bx, const double *by, const double *P,
                      double *hx, double *H)
{
    double phi1[3], phi2[3], phi3[3], gamma;
    double r2 = 2.0 * x[0], i2 = 2.0 * x[1], j2 = 2.0 * x[2], k2 = 2.0 * x[3];
    double Q1[12] = { r2, i2, -j2, -k2,   -k2, j2, i2, -r2,   j2, k2, r2, i2 };      /* :3945 */
    double Q2[12] = { k2, j2, i2, r2,     r2, -i2, j2, -k2,   -i2, -r2, k2, j2 };    /* :3949 */
    double Q3[12] = { -j2, k2, -r2, i2,   i2, r2, k2, j2,     r2, -i2, -j2, k2 };    /* :3953 */
    int i, c;
    qekf_phi(x, phi1, phi2, phi3, &gamma);
    for (i = 0; i < n; ++i) {
        const double *th = P + 3 * i;
        double p1 = th[0] * phi1[0] + th[1] * phi1[1] + th[2] * phi1[2];
        double p2 = th[0] * phi2[0] + th[1] * phi2[1] + th[2] * phi2[2];
        double p3 = th[0] * phi3[0] + th[1] * phi3[1] + th[2] * phi3[2];
        hx[i]     = p1 - bx[i] * p3 + x[4];       /* :3969 */
        hx[n + i] = p2 - by[i] * p3 + x[5];       /* :3970 */
        for (c = 0; c < 4; ++c) {
            double pq1 = th[0] * Q1[c] + th[1] * Q1[4 + c] + th[2] * Q1[8 + c];
            double pq2 = th[0] * Q2[c] + th[1] * Q2[4 + c] + th[2] * Q2[8 + c];
            double pq3 = th[0] * Q3[c] + th[1] * Q3[4 + c] + th[2] * Q3[8 + c];
            H[i * 6 + c]       = pq1 - bx[i] * pq3;    /* :3979 */
            H[(n + i) * 6 + c] = pq2 - by[i] * pq3;    /* :3980 */
        }
        H[i * 6 + 4] = 1.0; H[i * 6 + 5] = 0.0;
        H[(n + i) * 6 + 4] = 0.0; H[(n + i) * 6 + 5] = 1.0;
    }
}

/* One QEIF solve.  P: n x 3, uv: n x 2 (pixels).  trace (optional): max_it x 6 states after
 * each update.  Returns the number of iterations run. */
ORACLE_API int pnp_oracle_qeif(int n, const double *P, const double *uv, const double *K,
                               const oracle_params_t *prm, double *R, double *t, double *euler,
                               double *res_norm_out, double *trace)
{
    double Kinv[9];
    double *bx = (double *)malloc(sizeof(double) * (size_t)n * (2 + 2 + 12));
    double *by = bx + n, *hx = by + n, *H = hx + 2 * n;
    double x[6] = { 1, 0, 0, 0, 0, 0 };            /* :2831-2833 */
    double Omega[36], Sigma[36], zeta[6], Rd[6], tmp[36], work[36 + 6 + 36];
    double w, res_old, res = 1e5;
    int i, j, k, r, it = 0;

    inv3(K, Kinv);
    normalise(n, uv, Kinv, bx, by);

    for (i = 0; i < 4; ++i) Rd[i] = prm->proc_q;   /* :2791-2794 */
    Rd[4] = Rd[5] = prm->proc_d;
    for (i = 0; i < 36; ++i) Omega[i] = 0.0;
    for (i = 0; i < 6; ++i) Omega[i * 6 + i] = prm->omega0;   /* :2836 */
    pinv_svd(6, 6, Omega, Sigma, work);            /* :2837 */
    /* eif_Q_diag = 9 / f^2, eif_Q_pinv = pinv(diag) = 1 / eif_Q_diag (:2844-2852) */
    w = 1.0 / ((prm->meas_sigma_px * prm->meas_sigma_px) / (prm->f_weight * prm->f_weight));
    res_old = prm->res_old0;                       /* :2863 */

    while (it < prm->max_it) {
        double ratio;
        ++it;
        /* predict: Omega = pinv(G Sigma G^T + R), G = I (:2887); zeta = Omega x (:2889) */
        memcpy(tmp, Sigma, sizeof(tmp));
        for (i = 0; i < 6; ++i) tmp[i * 6 + i] += Rd[i];
        pinv_svd(6, 6, tmp, Omega, work);
        for (i = 0; i < 6; ++i) {
            zeta[i] = 0.0;
            for (j = 0; j < 6; ++j) zeta[i] += Omega[i * 6 + j] * x[j];
        }
        /* update (:2895-2898) */
        qekf_hx_H(n, x, bx, by, P, hx, H);
        {
            double res2 = 0.0;
            for (r = 0; r < 2 * n; ++r) {
                double z = (r < n) ? bx[r] : by[r - n];
                double dz = z - hx[r];
                double Hx = 0.0, v;
                for (j = 0; j < 6; ++j) Hx += H[r * 6 + j] * x[j];
                v = dz + Hx;                        /* z - hx + H x */
                for (i = 0; i < 6; ++i) {
                    double hw = H[r * 6 + i] * w;   /* (H^T Q^-1)[i, r] */
                    zeta[i] += hw * v;
                    for (j = 0; j < 6; ++j) Omega[i * 6 + j] += hw * H[r * 6 + j];
                }
                res2 += dz * dz;
            }
            res = sqrt(res2);                       /* :2905-2907 */
        }
        /* x = pinv(Omega) zeta (:2924-2925) */
        pinv_svd(6, 6, Omega, Sigma, work);
        for (i = 0; i < 6; ++i) {
            double acc = 0.0;
            for (j = 0; j < 6; ++j) acc += Sigma[i * 6 + j] * zeta[j];
            x[i] = acc;
        }
        if (trace) for (k = 0; k < 6; ++k) trace[(it - 1) * 6 + k] = x[k];
        ratio = (res - res_old) / res_old;          /* :2945 */
        res_old = res;
        if (fabs(ratio) < prm->exit_tol) break;     /* :2952 */
    }

    /* QEKF_reconstruct_R_t_m1 :3542-3609 */
    {
        double phi1[3], phi2[3], phi3[3], gamma, t3;
        qekf_phi(x, phi1, phi2, phi3, &gamma);
        for (k = 0; k < 3; ++k) { R[k] = phi1[k] / gamma; R[3 + k] = phi2[k] / gamma; R[6 + k] = phi3[k] / gamma; }
        t3 = 1.0 / gamma;
        t[0] = x[4] * t3; t[1] = x[5] * t3; t[2] = 1.0 * t3;
    }
    pnp_oracle_euler_from_R(R, 1, euler);
    *res_norm_out = res;
    free(bx);
    return it;
}

/* ------------------------------------------------------------------------------------ */
/* LM: solve_pnp_LM_single_pattern :2567-2769, EKF2_get_hx_H :3718-3836 (defaults          */
/*     is_hc0_4_6_linear=False, is_hc1_linear=True), EKF2_reconstruct_R_t_m1 :3500-3540    */
/* ------------------------------------------------------------------------------------ */
/* true_jac = 0: the reference's Jacobians (rows 4-6 use u, rows 7-9 use u/(2|u|)).
 * true_jac = 1: the actual gradients (2u and u/|u|) -- only used by the non-parity LM+ mode. */
static void ekf2_hx_H_ex(int n, const double *x, const double *bx, const double *by, const double *P,
                         double *hx, double *J, int true_jac)
{
    const double *u1 = x, *u2 = x + 3, *u3 = x + 6;
    double d1 = x[9], d2 = x[10], g = x[11];
    int i, k, Z = 2 * n + 9;
    double n1, n2, n3;
    memset(J, 0, sizeof(double) * (size_t)Z * 12);
    for (i = 0; i < n; ++i) {
        const double *th = P + 3 * i;
        double hu1 = 0.0, hu2 = 0.0;
        /* hu1_bar = [P, 0, -B_x*P] @ u_all  (:3738-3741) */
        for (k = 0; k < 3; ++k) hu1 += th[k] * u1[k];
        for (k = 0; k < 3; ++k) hu1 += (-bx[i] * th[k]) * u3[k];
        for (k = 0; k < 3; ++k) hu2 += th[k] * u2[k];
        for (k = 0; k < 3; ++k) hu2 += (-by[i] * th[k]) * u3[k];
        hx[i]     = g * hu1 + d1;                 /* :3750 */
        hx[n + i] = g * hu2 + d2;                 /* :3751 */
        for (k = 0; k < 3; ++k) {                 /* H1, H2 :3743-3744 */
            J[i * 12 + k]           = g * th[k];
            J[i * 12 + 6 + k]       = g * (-bx[i] * th[k]);
            J[(n + i) * 12 + 3 + k] = g * th[k];
            J[(n + i) * 12 + 6 + k] = g * (-by[i] * th[k]);
        }
        J[i * 12 + 9] = 1.0;        J[i * 12 + 11] = hu1;
        J[(n + i) * 12 + 10] = 1.0; J[(n + i) * 12 + 11] = hu2;
    }
    {
        double u11 = 0, u22 = 0, u33 = 0, u13 = 0, u23 = 0, u12 = 0;
        double *r;
        for (k = 0; k < 3; ++k) {
            u11 += u1[k] * u1[k]; u22 += u2[k] * u2[k]; u33 += u3[k] * u3[k];
            u13 += u1[k] * u3[k]; u23 += u2[k] * u3[k]; u12 += u1[k] * u2[k];
        }
        hx[2 * n + 0] = u13; hx[2 * n + 1] = u23; hx[2 * n + 2] = u12;            /* :3753-3755 */
        hx[2 * n + 3] = u11 - u33; hx[2 * n + 4] = u22 - u33; hx[2 * n + 5] = u11 - u22;  /* :3764-3766 */
        hx[2 * n + 6] = sqrt(u11); hx[2 * n + 7] = sqrt(u22); hx[2 * n + 8] = sqrt(u33);  /* :3770-3772 */
        n1 = sqrt(u11); n2 = sqrt(u22); n3 = sqrt(u33);   /* np.linalg.norm */
        r = J + (size_t)(2 * n) * 12;
        for (k = 0; k < 3; ++k) {
            r[0 * 12 + k] = u3[k];  r[0 * 12 + 6 + k] = u1[k];                    /* :3787-3788 */
            r[1 * 12 + 3 + k] = u3[k]; r[1 * 12 + 6 + k] = u2[k];                 /* :3790-3791 */
            r[2 * 12 + k] = u2[k];  r[2 * 12 + 3 + k] = u1[k];                    /* :3793-3794 */
            {
                const double kq = true_jac ? 2.0 : 1.0, kn = true_jac ? 1.0 : 2.0;
                r[3 * 12 + k] = kq * u1[k];  r[3 * 12 + 6 + k] = -kq * u3[k];         /* :3808-3809 (u, not 2u) */
                r[4 * 12 + 3 + k] = kq * u2[k]; r[4 * 12 + 6 + k] = -kq * u3[k];      /* :3811-3812 */
                r[5 * 12 + k] = kq * u1[k];  r[5 * 12 + 3 + k] = -kq * u2[k];         /* :3814-3815 */
                r[6 * 12 + k] = u1[k] / (kn * n1);                                    /* :3819 halved gradient */
                r[7 * 12 + 3 + k] = u2[k] / (kn * n2);                                /* :3821 */
                r[8 * 12 + 6 + k] = u3[k] / (kn * n3);                                /* :3823 */
            }
        }
    }
}

static void ekf2_hx_H(int n, const double *x, const double *bx, const double *by, const double *P,
                      double *hx, double *J)
{
    ekf2_hx_H_ex(n, x, bx, by, P, hx, J, 0);
}

/* EKF2_reconstruct_R_t_m1 :3500-3540 */
static void ekf2_reconstruct(const double *x, double *R, double *t)
{
    double U[9], s[3], V[9], UVt[9], D, t3, valueG;
    int i, j, k;
    svd_jacobi(3, 3, x, U, s, V);                 /* Gamma rows = x[0:3], x[3:6], x[6:9] */
    for (i = 0; i < 3; ++i)
        for (j = 0; j < 3; ++j) {
            double a = 0.0;
            for (k = 0; k < 3; ++k) a += U[i * 3 + k] * V[j * 3 + k];
            UVt[i * 3 + j] = a;
        }
    D = det3(UVt);                                /* :3513 */
    for (i = 0; i < 3; ++i)
        for (j = 0; j < 3; ++j)
            R[i * 3 + j] = U[i * 3 + 0] * V[j * 3 + 0] + U[i * 3 + 1] * V[j * 3 + 1] + D * U[i * 3 + 2] * V[j * 3 + 2];
    valueG = s[0] * x[11];                        /* np.linalg.norm(Gamma, ord=2) * gamma :3530 */
    t3 = 1.0 / valueG;
    t[0] = x[9] * t3; t[1] = x[10] * t3; t[2] = 1.0 * t3;     /* :3536 */
}

ORACLE_API int pnp_oracle_lm(int n, const double *P, const double *uv, const double *K,
                             const oracle_params_t *prm, double *R, double *t, double *euler,
                             double *res_norm_out, double *trace)
{
    int Z = 2 * n + 9, i, j, r, it = 0;
    double Kinv[9];
    double *bx = (double *)malloc(sizeof(double) * ((size_t)n * 2 + (size_t)Z * 14));
    double *by = bx + n, *hx = by + n, *J = hx + Z, *dz = J + (size_t)Z * 12;
    double x[12] = { 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 1 };   /* :2619-2624 */
    double A[144], Ainv[144], g[12], work[144 + 12 + 144];
    double res = 1e5;

    inv3(K, Kinv);
    normalise(n, uv, Kinv, bx, by);

    while (it < prm->max_it) {                    /* :2642, no exit test */
        double res2 = 0.0;
        ++it;
        ekf2_hx_H(n, x, bx, by, P, hx, J);
        for (i = 0; i < 12; ++i)
            for (j = 0; j < 12; ++j) {
                double a = 0.0;
                for (r = 0; r < Z; ++r) a += J[r * 12 + i] * J[r * 12 + j];
                A[i * 12 + j] = a + ((i == j) ? prm->lm_lambda : 0.0);   /* :2666-2667 */
            }
        pinv_svd(12, 12, A, Ainv, work);          /* :2675 */
        for (r = 0; r < Z; ++r) {
            double z;
            if (r < n) z = bx[r];
            else if (r < 2 * n) z = by[r - n];
            else if (r < 2 * n + 6) z = 0.0;      /* :2603 */
            else z = 1.0;                         /* :2604 */
            dz[r] = z - hx[r];
            if (r < 2 * n) res2 += dz[r] * dz[r];
        }
        res = sqrt(res2);                         /* :2681 */
        for (i = 0; i < 12; ++i) {
            double a = 0.0;
            for (r = 0; r < Z; ++r) a += J[r * 12 + i] * dz[r];
            g[i] = a;
        }
        for (i = 0; i < 12; ++i) {                /* :2684, :2702 */
            double a = 0.0;
            for (j = 0; j < 12; ++j) a += Ainv[i * 12 + j] * g[j];
            x[i] += a;
        }
        if (trace) for (i = 0; i < 12; ++i) trace[(it - 1) * 12 + i] = x[i];
    }
    ekf2_reconstruct(x, R, t);
    pnp_oracle_euler_from_R(R, 1, euler);
    *res_norm_out = res;
    free(bx);
    return it;
}

ORACLE_API int pnp_oracle_linear_f2(int n, const double *P, const double *uv, const double *K,
                                    const oracle_params_t *prm, double *R, double *t, double *euler,
                                    double *res_norm_out, double *trace);

/* ------------------------------------------------------------------------------------ */
/* EIF2: solve_pnp_EIF2_single_pattern :2001-2276 -- the iterated information filter on the  */
/* 12-state model of EKF2_get_hx_H (:3718-3836, same defaults as LM) with the state-dependent */
/* process covariance of EKF2_get_process_covariance_R (:3668-3716) and QEIF's early exit.     */
/* ------------------------------------------------------------------------------------ */
/* EKF2_get_process_covariance_R :3668-3716.  Rk is 12 x 12 row-major. */
static void ekf2_process_cov(const double *x, double *Rk)
{
    const double sig_th = 60.0 * (3.14159265358979323846 / 180.0);   /* np.deg2rad(60.0) :3686 */
    const double sig_t12 = 0.05, sig_t3 = 2.0;                       /* :3687, :3689 */
    double C[27];                                                     /* vstack(-so3(u1), -so3(u2), -so3(u3)) :3694 */
    int b, i, j, k;
    memset(Rk, 0, sizeof(double) * 144);
    for (b = 0; b < 3; ++b) {
        const double *u = x + 3 * b;
        double *c = C + 9 * b;                                        /* get_so3_matrix_from_vec3 :3661-3666, negated */
        c[0] = -0.0;   c[1] = u[2];   c[2] = -u[1];
        c[3] = -u[2];  c[4] = -0.0;   c[5] = u[0];
        c[6] = u[1];   c[7] = -u[0];  c[8] = -0.0;
    }
    for (i = 0; i < 9; ++i)                                           /* Ck @ (sigma^2 I) @ Ck.T :3695-3696 */
        for (j = 0; j < 9; ++j) {
            double a = 0.0;
            for (k = 0; k < 3; ++k) a += (C[i * 3 + k] * (sig_th * sig_th)) * C[j * 3 + k];
            Rk[i * 12 + j] = a;
        }
    Rk[9 * 12 + 9] = Rk[10 * 12 + 10] = fabs(x[11]) * (sig_t12 * sig_t12);   /* :3700 */
    Rk[11 * 12 + 11] = (x[11] * x[11]) * (sig_t3 * sig_t3);                  /* :3701 */
}

ORACLE_API int pnp_oracle_eif2(int n, const double *P, const double *uv, const double *K,
                               const oracle_params_t *prm, double *R, double *t, double *euler,
                               double *res_norm_out, double *trace)
{
    int Z = 2 * n + 9, i, j, r, it = 0;
    double Kinv[9];
    double *bx = (double *)malloc(sizeof(double) * ((size_t)n * 2 + (size_t)Z * 15));
    double *by = bx + n, *hx = by + n, *J = hx + Z, *qinv = J + (size_t)Z * 12, *z = qinv + Z;
    double x[12] = { 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 1 };   /* :2058-2063 */
    double Omega[144], Sigma[144], Rk[144], tmp[144], zeta[12], work[144 + 12 + 144];
    double res = 1e5, res_old = prm->res_old0;                /* :2101-2102 (1e-7) */

    inv3(K, Kinv);
    normalise(n, uv, Kinv, bx, by);
    for (r = 0; r < Z; ++r) {
        /* eif_z :2049-2055 and eif_Q_diag :2080-2091, eif_Q_pinv = pinv(diag) = 1 / diag */
        double q;
        if (r < n) z[r] = bx[r];
        else if (r < 2 * n) z[r] = by[r - n];
        else if (r < 2 * n + 6) z[r] = 0.0;
        else z[r] = 1.0;
        if (r < 2 * n) q = (prm->meas_sigma_px * prm->meas_sigma_px) / (prm->f_weight * prm->f_weight);
        else if (r < 2 * n + 3) q = 1e-2 * 16.0;
        else if (r < 2 * n + 6) q = (2.0 * 2.0) * 1e-2 * 16.0;
        else q = 1.0;
        qinv[r] = 1.0 / q;
    }
    for (i = 0; i < 144; ++i) Omega[i] = 0.0;
    for (i = 0; i < 12; ++i) Omega[i * 12 + i] = prm->omega0;  /* :2067 (1e-5) */
    pinv_svd(12, 12, Omega, Sigma, work);                     /* eif_Omega_pinv :2068 */

    while (it < prm->max_it) {                                /* :2107 */
        double ratio, res2 = 0.0;
        ++it;
        /* predict :2146-2152: Omega = pinv(G Omega_pinv G^T + R_k), G = I; zeta = Omega x */
        ekf2_process_cov(x, Rk);
        for (i = 0; i < 144; ++i) tmp[i] = Sigma[i] + Rk[i];
        pinv_svd(12, 12, tmp, Omega, work);
        for (i = 0; i < 12; ++i) {
            zeta[i] = 0.0;
            for (j = 0; j < 12; ++j) zeta[i] += Omega[i * 12 + j] * x[j];
        }
        /* update :2157-2164 */
        ekf2_hx_H(n, x, bx, by, P, hx, J);
        for (r = 0; r < Z; ++r) {
            double dz = z[r] - hx[r], Hx = 0.0, v;
            for (j = 0; j < 12; ++j) Hx += J[r * 12 + j] * x[j];
            v = dz + Hx;                                      /* z - hx + H x */
            for (i = 0; i < 12; ++i) {
                double hw = J[r * 12 + i] * qinv[r];          /* (H^T Q^-1)[i, r] */
                if (hw == 0.0) continue;
                zeta[i] += hw * v;
                for (j = 0; j < 12; ++j) Omega[i * 12 + j] += hw * J[r * 12 + j];
            }
            if (r < 2 * n) res2 += dz * dz;
        }
        res = sqrt(res2);                                     /* res_norm over the 2n measurement rows :2178 */
        pinv_svd(12, 12, Omega, Sigma, work);                 /* :2184 */
        for (i = 0; i < 12; ++i) {                            /* :2185 */
            double a = 0.0;
            for (j = 0; j < 12; ++j) a += Sigma[i * 12 + j] * zeta[j];
            x[i] = a;
        }
        if (trace) for (i = 0; i < 12; ++i) trace[(it - 1) * 12 + i] = x[i];
        ratio = (res - res_old) / res_old;                    /* :2196 */
        res_old = res;
        if (fabs(ratio) < prm->exit_tol) break;               /* :2202 */
    }
    ekf2_reconstruct(x, R, t);
    pnp_oracle_euler_from_R(R, 1, euler);
    *res_norm_out = res;
    free(bx);
    return it;
}

/* ------------------------------------------------------------------------------------ */
/* LM+ -- NOT a reference method: the pipeline BASELINE.json's north star   */
/* describes.  Linear stage F2 (3 iterations) for the initial pose, then the same 12-state      */
/* damped Gauss-Newton as solve_pnp_LM_single_pattern but with the TRUE constraint gradients     */
/* and a convergence test: stop when max |dx| <= 1e-10, at most max_it iterations.               */
/* res_norm = ||z - hx|| over the 2n measurement rows at the RETURNED state.                      */
/* ------------------------------------------------------------------------------------ */
ORACLE_API int pnp_oracle_lm_plus(int n, const double *P, const double *uv, const double *K,
                                  const oracle_params_t *prm, double *R, double *t, double *euler,
                                  double *res_norm_out, double *trace)
{
    int Z = 2 * n + 9, i, j, r, it = 0;
    double Kinv[9];
    double *bx = (double *)malloc(sizeof(double) * ((size_t)n * 2 + (size_t)Z * 14));
    double *by = bx + n, *hx = by + n, *J = hx + Z, *dz = J + (size_t)Z * 12;
    double x[12], A[144], Ainv[144], g[12], work[144 + 12 + 144];
    double R0[9], t0[3], e0[3], r0, res2 = 0.0;
    (void)trace;
    pnp_oracle_linear_f2(n, P, uv, K, prm, R0, t0, e0, &r0, NULL);
    for (i = 0; i < 9; ++i) x[i] = R0[i];
    x[9] = t0[0] / t0[2]; x[10] = t0[1] / t0[2]; x[11] = 1.0 / t0[2];
    inv3(K, Kinv);
    normalise(n, uv, Kinv, bx, by);
    while (it < prm->max_it) {
        double step = 0.0;
        ++it;
        ekf2_hx_H_ex(n, x, bx, by, P, hx, J, 1);
        for (i = 0; i < 12; ++i)
            for (j = 0; j < 12; ++j) {
                double a = 0.0;
                for (r = 0; r < Z; ++r) a += J[r * 12 + i] * J[r * 12 + j];
                A[i * 12 + j] = a + ((i == j) ? prm->lm_lambda : 0.0);
            }
        pinv_svd(12, 12, A, Ainv, work);
        for (r = 0; r < Z; ++r) {
            double z = (r < n) ? bx[r] : (r < 2 * n) ? by[r - n] : (r < 2 * n + 6) ? 0.0 : 1.0;
            dz[r] = z - hx[r];
        }
        for (i = 0; i < 12; ++i) {
            double a = 0.0;
            for (r = 0; r < Z; ++r) a += J[r * 12 + i] * dz[r];
            g[i] = a;
        }
        for (i = 0; i < 12; ++i) {
            double a = 0.0;
            for (j = 0; j < 12; ++j) a += Ainv[i * 12 + j] * g[j];
            x[i] += a;
            if (fabs(a) > step) step = fabs(a);
        }
        if (step <= 1e-10) break;
    }
    ekf2_hx_H_ex(n, x, bx, by, P, hx, J, 1);
    for (r = 0; r < 2 * n; ++r) {
        double z = (r < n) ? bx[r] : by[r - n];
        res2 += (z - hx[r]) * (z - hx[r]);
    }
    ekf2_reconstruct(x, R, t);
    pnp_oracle_euler_from_R(R, 1, euler);
    *res_norm_out = sqrt(res2);
    free(bx);
    return it;
}

/* ------------------------------------------------------------------------------------ */
/* block reconstruction :4158-4364 and phi_3 update :4002-4017                             */
/* phi = [phi_1(3), phi_2(3), delta_1, delta_2]                                            */
/* ------------------------------------------------------------------------------------ */
static void block_reconstruct(const double *phi, double *R, double *t, double *t3_out)
{
    double K00 = phi[0], K01 = phi[1], K10 = phi[3], K11 = phi[4];
    double k1 = K00 * K00 + K10 * K10;            /* (K^T K)[0,0] :4205-4211 */
    double k2 = K00 * K01 + K10 * K11;
    double k3 = K01 * K01 + K11 * K11;
    double Dd = (k1 - k3) * (k1 - k3) + 4 * k2 * k2;      /* :4215 */
    double gamma2 = 0.5 * ((k1 + k3) + sqrt(Dd));
    double gamma = sqrt(gamma2);
    double e_se = sqrt(gamma2 - k1);              /* :4230-4231 */
    double sgn = (k2 > 0.0) ? 1.0 : ((k2 < 0.0) ? -1.0 : 0.0);
    double d_se = -sgn * sqrt(gamma2 - k3);       /* :4233-4234 */
    /* delta_se = inv(K^T) @ beta_se :4242 ; K^T = [[K00,K10],[K01,K11]] */
    double detK = K00 * K11 - K01 * K10;
    double ds0 = (K11 * e_se - K10 * d_se) / detK;
    double ds1 = (-K01 * e_se + K00 * d_se) / detK;
    double c = detK / gamma;                      /* :4247 */
    double a0 = -(c * ds0), a1 = -(c * ds1);      /* :4249 */
    double sim = a0 * phi[2] + a1 * phi[5];       /* alpha_se . Gamma[0:2,2] :4261-4262 */
    double se = (sim < 0.0) ? -1.0 : 1.0;         /* :4266-4269 */
    double G[9], t3;
    int i;
    G[0] = K00; G[1] = K01; G[2] = se * a0;
    G[3] = K10; G[4] = K11; G[5] = se * a1;
    G[6] = se * e_se; G[7] = se * d_se; G[8] = c;   /* :4326-4328 */
    for (i = 0; i < 9; ++i) R[i] = G[i] / gamma;  /* :4340 */
    t3 = 1.0 / gamma;                             /* :4357 */
    t[0] = phi[6] * t3; t[1] = phi[7] * t3; t[2] = 1.0 * t3;    /* :4360 */
    *t3_out = t3;
}

/* ------------------------------------------------------------------------------------ */
/* Linear stage, formulation 2: solve_pnp_formulation_2_single_pattern :693-953,           */
/* helpers :3260-3375                                                                      */
/* ------------------------------------------------------------------------------------ */
ORACLE_API int pnp_oracle_linear_f2(int n, const double *P, const double *uv, const double *K,
                                    const oracle_params_t *prm, double *R, double *t, double *euler,
                                    double *res_norm_out, double *trace)
{
    double Kinv[9];
    size_t nn = (size_t)n;
    double *bx = (double *)malloc(sizeof(double) * (nn * 2 + nn * 4 + 4 * nn + (nn * 4 + 4 + 16)));
    double *by = bx + n, *D = by + n, *Dp = D + 4 * nn, *work = Dp + 4 * nn;
    double v0x[4], v0y[4], Mx[12], My[12], phi3[3] = { 0.0, 0.0, 1.0 };   /* :758 */
    double res_norm = 30.0, t3 = 1.0;
    int i, k, c, it = 0;

    inv3(K, Kinv);
    normalise(n, uv, Kinv, bx, by);
    for (i = 0; i < n; ++i) {                     /* D = [P|1] :3274-3281 */
        D[i * 4 + 0] = P[i * 3 + 0]; D[i * 4 + 1] = P[i * 3 + 1]; D[i * 4 + 2] = P[i * 3 + 2]; D[i * 4 + 3] = 1.0;
    }
    pinv_svd(n, 4, D, Dp, work);                  /* :3283-3289, Dp is 4 x n */
    for (k = 0; k < 4; ++k) {                     /* :3314-3328 */
        double ax = 0.0, ay = 0.0;
        for (i = 0; i < n; ++i) { ax += Dp[k * nn + i] * bx[i]; ay += Dp[k * nn + i] * by[i]; }
        v0x[k] = ax; v0y[k] = ay;
        for (c = 0; c < 3; ++c) {
            double mx = 0.0, my = 0.0;
            for (i = 0; i < n; ++i) {
                mx += Dp[k * nn + i] * (bx[i] * P[i * 3 + c]);
                my += Dp[k * nn + i] * (by[i] * P[i * 3 + c]);
            }
            Mx[k * 3 + c] = mx; My[k * 3 + c] = my;
        }
    }
    while (it < prm->linear_it) {                 /* :779 */
        double phix[4], phiy[4], phi[8], pxn[4], pyn[4], phi3_new[3];
        double rx2 = 0.0, ry2 = 0.0;
        ++it;
        for (k = 0; k < 4; ++k) {                 /* :3337 */
            phix[k] = v0x[k] + (Mx[k * 3] * phi3[0] + Mx[k * 3 + 1] * phi3[1] + Mx[k * 3 + 2] * phi3[2]);
            phiy[k] = v0y[k] + (My[k * 3] * phi3[0] + My[k * 3 + 1] * phi3[1] + My[k * 3 + 2] * phi3[2]);
        }
        for (k = 0; k < 3; ++k) { phi[k] = phix[k]; phi[3 + k] = phiy[k]; }   /* :3158 */
        phi[6] = phix[3]; phi[7] = phiy[3];
        block_reconstruct(phi, R, t, &t3);        /* :862 */
        for (k = 0; k < 3; ++k) phi3_new[k] = (1.0 / t3) * R[6 + k];          /* :4002-4010 */
        for (k = 0; k < 3; ++k) { pxn[k] = R[k] / t3; pyn[k] = R[3 + k] / t3; }   /* :868-869 */
        pxn[3] = t[0] / t3; pyn[3] = t[1] / t3;
        for (i = 0; i < n; ++i) {                 /* f2_cal_res_all with the OLD phi_3 :877, :3368 */
            double db = 1.0 + (P[i * 3] * phi3[0] + P[i * 3 + 1] * phi3[1] + P[i * 3 + 2] * phi3[2]);
            double dx = D[i * 4] * pxn[0] + D[i * 4 + 1] * pxn[1] + D[i * 4 + 2] * pxn[2] + D[i * 4 + 3] * pxn[3];
            double dy = D[i * 4] * pyn[0] + D[i * 4 + 1] * pyn[1] + D[i * 4 + 2] * pyn[2] + D[i * 4 + 3] * pyn[3];
            double ex = bx[i] * db - dx, ey = by[i] * db - dy;
            rx2 += ex * ex; ry2 += ey * ey;
        }
        {   /* res_norm_all = sqrt(res_norm_x**2 + res_norm_y**2) :3374 */
            double nx = sqrt(rx2), ny = sqrt(ry2);
            res_norm = sqrt(nx * nx + ny * ny);
        }
        for (k = 0; k < 3; ++k) phi3[k] = phi3_new[k];
        if (trace) { for (k = 0; k < 8; ++k) trace[(it - 1) * 11 + k] = phi[k]; for (k = 0; k < 3; ++k) trace[(it - 1) * 11 + 8 + k] = phi3[k]; }
    }
    pnp_oracle_euler_from_R(R, 1, euler);
    *res_norm_out = res_norm;
    free(bx);
    return it;
}

/* ------------------------------------------------------------------------------------ */
/* Linear stage, formulation 1: solve_pnp_single_pattern :205-430, helpers :3031-3111      */
/* ------------------------------------------------------------------------------------ */
ORACLE_API int pnp_oracle_linear_f1(int n, const double *P, const double *uv, const double *K,
                                    const oracle_params_t *prm, double *R, double *t, double *euler,
                                    double *res_norm_out, double *trace)
{
    double Kinv[9];
    size_t m = 2 * (size_t)n;
    double *bx = (double *)malloc(sizeof(double) * ((size_t)n * 2 + m + m * 8 + 8 * m + (m * 8 + 8 + 64)));
    double *by = bx + n, *Ball = by + n, *A = Ball + m, *Ap = A + m * 8, *work = Ap + 8 * m;
    double phi3[3] = { 0.0, 0.0, 1.0 }, phi[8], t3 = 1.0, res_norm = 30.0;
    int i, k, it = 0;
    size_t r;

    inv3(K, Kinv);
    normalise(n, uv, Kinv, bx, by);
    for (i = 0; i < n; ++i) { Ball[2 * i] = bx[i]; Ball[2 * i + 1] = by[i]; }   /* get_B_all :3094-3111 */

    while (it < prm->linear_it) {
        double phin[8], phi3_new[3], res2 = 0.0;
        ++it;
        memset(A, 0, sizeof(double) * m * 8);
        for (i = 0; i < n; ++i) {                 /* get_Delta_i :3031-3046, get_A_i :3048-3062 */
            const double *th = P + 3 * i;
            double Delta = th[0] * phi3[0] + th[1] * phi3[1] + th[2] * phi3[2] + 1.0;
            if (fabs(Delta) <= 1e-7) Delta = (Delta < 0) ? -1e-7 : 1e-7;
            for (k = 0; k < 3; ++k) {
                A[(2 * (size_t)i) * 8 + k] = th[k] / Delta;
                A[(2 * (size_t)i + 1) * 8 + 3 + k] = th[k] / Delta;
            }
            A[(2 * (size_t)i) * 8 + 6] = 1.0 / Delta;
            A[(2 * (size_t)i + 1) * 8 + 7] = 1.0 / Delta;
        }
        pinv_svd((int)m, 8, A, Ap, work);         /* :262 phi = pinv(A) @ B */
        for (k = 0; k < 8; ++k) {
            double a = 0.0;
            for (r = 0; r < m; ++r) a += Ap[k * m + r] * Ball[r];
            phi[k] = a;
        }
        block_reconstruct(phi, R, t, &t3);        /* :365 */
        for (k = 0; k < 3; ++k) phi3_new[k] = (1.0 / t3) * R[6 + k];
        for (k = 0; k < 3; ++k) { phin[k] = R[k] / t3; phin[3 + k] = R[3 + k] / t3; }   /* :381 */
        phin[6] = t[0] / t3; phin[7] = t[1] / t3;
        for (r = 0; r < m; ++r) {                 /* :383-387 */
            double a = 0.0, e;
            for (k = 0; k < 8; ++k) a += A[r * 8 + k] * phin[k];
            e = Ball[r] - a;
            res2 += e * e;
        }
        res_norm = sqrt(res2);
        for (k = 0; k < 3; ++k) phi3[k] = phi3_new[k];
        if (trace) { for (k = 0; k < 8; ++k) trace[(it - 1) * 11 + k] = phi[k]; for (k = 0; k < 3; ++k) trace[(it - 1) * 11 + 8 + k] = phi3[k]; }
    }
    block_reconstruct(phi, R, t, &t3);            /* :413 (same phi, same result) */
    pnp_oracle_euler_from_R(R, 1, euler);
    *res_norm_out = res_norm;
    free(bx);
    return it;
}


/* ------------------------------------------------------------------------------------ */
/* pthread parallel-for (dynamic chunks) -- the batch drivers below use it so the CPU      */
/* baseline can use every host core without depending on an OpenMP runtime                 */
/* ------------------------------------------------------------------------------------ */
typedef void (*range_fn)(int64_t lo, int64_t hi, void *ctx);
typedef struct { range_fn fn; void *ctx; int64_t B, chunk; int64_t next; } pf_job_t;

static void *pf_worker(void *arg)
{
    pf_job_t *job = (pf_job_t *)arg;
    for (;;) {
        int64_t lo = __atomic_fetch_add(&job->next, job->chunk, __ATOMIC_RELAXED), hi;
        if (lo >= job->B) break;
        hi = lo + job->chunk; if (hi > job->B) hi = job->B;
        job->fn(lo, hi, job->ctx);
    }
    return NULL;
}

ORACLE_API int pnp_oracle_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return (n < 1) ? 1 : (int)n;
}

static void parallel_for(int64_t B, int n_threads, int64_t chunk, range_fn fn, void *ctx)
{
    pf_job_t job;
    pthread_t *th;
    int i, started = 0;
    if (n_threads <= 0) n_threads = pnp_oracle_num_threads();
    if ((int64_t)n_threads > (B + chunk - 1) / chunk) n_threads = (int)((B + chunk - 1) / chunk);
    job.fn = fn; job.ctx = ctx; job.B = B; job.chunk = chunk; job.next = 0;
    if (n_threads <= 1) { pf_worker(&job); return; }
    th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (i = 0; i < n_threads - 1; ++i)
        if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) ++started;
    pf_worker(&job);
    for (i = 0; i < started; ++i) pthread_join(th[i], NULL);
    free(th);
}

/* ------------------------------------------------------------------------------------ */
/* batch driver (the reference's per-problem script loop, OpenMP over problems)            */
/* method: 0 = QEIF, 1 = LM, 2 = linear F2, 3 = linear F1                                  */
/* solve_pnp's pattern arg-min (:166-199): strict '<', first pattern wins ties             */
/* ------------------------------------------------------------------------------------ */
typedef int (*solve_fn)(int, const double *, const double *, const double *, const oracle_params_t *,
                        double *, double *, double *, double *, double *);

typedef struct {
    solve_fn fn; int n, n_patterns; const double *uv, *patterns, *K; const oracle_params_t *prm;
    double *R, *t, *euler, *res_norm; int32_t *iters, *best_pattern;
} solve_ctx_t;

static void solve_range(int64_t lo, int64_t hi, void *vctx)
{
    solve_ctx_t *c = (solve_ctx_t *)vctx;
    int64_t b;
    int n = c->n;
    for (b = lo; b < hi; ++b) {
        int p, have = 0;
        double best = 0.0;
        for (p = 0; p < c->n_patterns; ++p) {
            double Rp[9], tp[3], ep[3], rn;
            int it = c->fn(n, c->patterns + (size_t)p * n * 3, c->uv + (size_t)b * n * 2, c->K, c->prm, Rp, tp, ep, &rn, NULL);
            if (!have || rn < best) {             /* :185 */
                have = 1; best = rn;
                memcpy(c->R + b * 9, Rp, sizeof(Rp)); memcpy(c->t + b * 3, tp, sizeof(tp));
                memcpy(c->euler + b * 3, ep, sizeof(ep));
                c->res_norm[b] = rn; c->iters[b] = it; c->best_pattern[b] = p;
            }
        }
    }
}

ORACLE_API int pnp_oracle_solve_batch(int method, int64_t B, int n, const double *uv,
                                      const double *patterns, int n_patterns, const double *K,
                                      const oracle_params_t *prm, double *R, double *t, double *euler,
                                      double *res_norm, int32_t *iters, int32_t *best_pattern,
                                      int n_threads)
{
    solve_ctx_t c;
    switch (method) {
    case 0: c.fn = pnp_oracle_qeif; break;
    case 1: c.fn = pnp_oracle_lm; break;
    case 2: c.fn = pnp_oracle_linear_f2; break;
    case 3: c.fn = pnp_oracle_linear_f1; break;
    case 4: c.fn = pnp_oracle_lm_plus; break;
    case 5: c.fn = pnp_oracle_eif2; break;
    default: return -1;
    }
    c.n = n; c.n_patterns = n_patterns; c.uv = uv; c.patterns = patterns; c.K = K; c.prm = prm;
    c.R = R; c.t = t; c.euler = euler; c.res_norm = res_norm; c.iters = iters; c.best_pattern = best_pattern;
    parallel_for(B, n_threads, 16, solve_range, &c);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* error reporting: TEST_TOOLBOX.py check_if_the_sample_passed :55-62,                     */
/* cal_LM_error_distances :252-286, compare_result_and_generate_result_dict :291-465,      */
/* and the glue of random_stress_test.py:353-377                                           */
/* report[16] = depth_err(m), roll_err, pitch_err, yaw_err (deg),                          */
/*              LM_GT avg/max, predict_LM avg/max, predict_GT avg/max (all x distance_GT), */
/*              t3_est, distance_GT, roll_GT? ... see pnpb200.h PNPB200_REPORT_*           */
/* ------------------------------------------------------------------------------------ */
static void err_dist(int n, const double *a, const double *b, double *avg, double *mx, int *imx)
{
    double tot = 0.0, m = 0.0;
    int i, im = -1;
    for (i = 0; i < n; ++i) {
        double d0 = a[3 * i] - b[3 * i], d1 = a[3 * i + 1] - b[3 * i + 1], d2 = a[3 * i + 2] - b[3 * i + 2];
        double e = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        tot += e;
        if (e > m) { m = e; im = i; }            /* strict '>' :274 */
    }
    *avg = tot / n; *mx = m; *imx = im;
}

/* gt = (distance_GT [m], roll_GT, pitch_GT, yaw_GT [deg]); est euler order (roll, yaw, pitch).
 * uv: the n x 2 measured pixels (homogeneous third component taken as 1).
 * flags[4] = depth, roll, pitch, yaw pass (bounds = 10 cm / 10 deg by default). */
ORACLE_API void pnp_oracle_report(int n, const double *P, const double *uv, const double *K,
                                  const double *R_est, const double *t_est, const double *euler_est,
                                  const double *gt, const double *bounds, double *report,
                                  int32_t *flags, int32_t *max_idx)
{
    double t3 = t_est[2], dist = gt[0];
    double roll_e = euler_est[0], yaw_e = euler_est[1], pitch_e = euler_est[2];
    double R_gt[9], t_gt[3];
    double *buf = (double *)malloc(sizeof(double) * (size_t)n * 9);
    double *rep = buf, *gtp = buf + 3 * (size_t)n, *meas = gtp + 3 * (size_t)n;
    int i;
    pnp_oracle_R_from_euler(gt[1], gt[3], gt[2], 1, R_gt);      /* random_stress_test.py:365 */
    for (i = 0; i < 3; ++i) t_gt[i] = (t_est[i] / t3) * dist;   /* :367-368 */
    flags[0] = fabs(t3 * 100.0 - dist * 100.0) < bounds[0];     /* :374-377 (depth in cm) */
    flags[1] = fabs(roll_e - gt[1]) < bounds[1];
    flags[2] = fabs(pitch_e - gt[2]) < bounds[2];
    flags[3] = fabs(yaw_e - gt[3]) < bounds[3];
    pnp_oracle_project(n, P, K, R_est, t_est, 0, 1.0, rep);     /* TEST_TOOLBOX.py:312 */
    pnp_oracle_project(n, P, K, R_gt, t_gt, 0, 1.0, gtp);       /* :314 */
    for (i = 0; i < n; ++i) { meas[3 * i] = uv[2 * i]; meas[3 * i + 1] = uv[2 * i + 1]; meas[3 * i + 2] = 1.0; }
    report[0] = t3 - dist;                                      /* depth_err :423 */
    report[1] = roll_e - gt[1]; report[2] = pitch_e - gt[2]; report[3] = yaw_e - gt[3];
    {
        double a, m;
        err_dist(n, meas, gtp, &a, &m, &max_idx[0]);            /* LM_GT :321 */
        report[4] = a * dist; report[5] = m * dist;             /* :452-453 */
        err_dist(n, rep, meas, &a, &m, &max_idx[1]);            /* predict_LM :326 */
        report[6] = a * dist; report[7] = m * dist;
        err_dist(n, rep, gtp, &a, &m, &max_idx[2]);             /* predict_GT :331 */
        report[8] = a * dist; report[9] = m * dist;
    }
    report[10] = t3; report[11] = dist;
    report[12] = roll_e; report[13] = pitch_e; report[14] = yaw_e;
    report[15] = 0.0;
    free(buf);
}

typedef struct {
    int n; const double *P, *uv, *K, *R_est, *t_est, *euler_est, *gt, *bounds;
    double *report; int32_t *flags, *max_idx;
} report_ctx_t;

static void report_range(int64_t lo, int64_t hi, void *vctx)
{
    report_ctx_t *c = (report_ctx_t *)vctx;
    int64_t b;
    for (b = lo; b < hi; ++b)
        pnp_oracle_report(c->n, c->P, c->uv + (size_t)b * c->n * 2, c->K, c->R_est + b * 9, c->t_est + b * 3,
                          c->euler_est + b * 3, c->gt + b * 4, c->bounds, c->report + b * 16,
                          c->flags + b * 4, c->max_idx + b * 3);
}

ORACLE_API void pnp_oracle_report_batch(int64_t B, int n, const double *P, const double *uv,
                                        const double *K, const double *R_est, const double *t_est,
                                        const double *euler_est, const double *gt, const double *bounds,
                                        double *report, int32_t *flags, int32_t *max_idx, int n_threads)
{
    report_ctx_t c;
    c.n = n; c.P = P; c.uv = uv; c.K = K; c.R_est = R_est; c.t_est = t_est; c.euler_est = euler_est;
    c.gt = gt; c.bounds = bounds; c.report = report; c.flags = flags; c.max_idx = max_idx;
    parallel_for(B, n_threads, 256, report_range, &c);
}

/* ------------------------------------------------------------------------------------ */
/* synthetic workload generator shared by CPU and GPU (counter-based, keyed by the GLOBAL  */
/* problem index so that sharding never changes the data).  Philox4x32-10.                  */
/* Draw order per problem follows random_stress_test.py:246-254:                           */
/*   roll, pitch, yaw ~ U(-a, a); depth ~ U(d0, d1) [m]; FOV_x, FOV_y ~ U(-f, f) [deg]       */
/* then one N(0,1) pair per point for optional pixel noise (LM_noise_test.py style).        */
/* ------------------------------------------------------------------------------------ */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t *out)
{
    int r;
    for (r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 53-bit uniform in [0,1) from two 32-bit words */
static double u01(uint32_t hi, uint32_t lo)
{
    uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}

typedef struct {
    uint64_t seed;
    double angle_range_deg;   /* 45   random_stress_test.py:246 */
    double depth_min_m;       /* 0.20 :251 */
    double depth_max_m;       /* 2.25 */
    double fov_max_deg;       /* 45   :252 */
    int32_t is_quantized;     /* :50 */
    double quantize_q;        /* 1.0 */
    double noise_sigma_px;    /* 0 = none */
    double roll_center_deg, pitch_center_deg, yaw_center_deg;   /* angle = centre + U(-range, range) */
} oracle_synth_t;

typedef struct {
    int64_t b0; int n; const double *P, *K; const oracle_synth_t *cfg; double *uv, *gt, *R_gt, *t_gt;
    double perturb_radius; int fixed_idx; double *perturb;   /* face_variation_test.py:317-345 */
} synth_ctx_t;

/* three N(0,1) draws for point i of problem gidx (the pattern perturbation) */
static void perturb_normals(uint64_t gidx, int i, uint32_t k0, uint32_t k1, double *z)
{
    uint32_t a[4], b[4];
    double r1, r2;
    philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), 16u + (uint32_t)i, 2u, k0, k1, a);
    philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), 16u + (uint32_t)i, 3u, k0, k1, b);
    r1 = sqrt(-2.0 * log(1.0 - u01(a[0], a[1]))); r2 = sqrt(-2.0 * log(1.0 - u01(b[0], b[1])));
    z[0] = r1 * cos(2.0 * M_PI * u01(a[2], a[3])); z[1] = r1 * sin(2.0 * M_PI * u01(a[2], a[3]));
    z[2] = r2 * cos(2.0 * M_PI * u01(b[2], b[3]));
}

static void synth_range(int64_t lo, int64_t hi, void *vctx)
{
    synth_ctx_t *c = (synth_ctx_t *)vctx;
    const oracle_synth_t *cfg = c->cfg;
    int n = c->n;
    int64_t bb;
    uint32_t k0 = (uint32_t)cfg->seed, k1 = (uint32_t)(cfg->seed >> 32);
    double *uvw = (double *)malloc(sizeof(double) * (size_t)n * 6);
    double *Pp = uvw + (size_t)n * 3;                 /* the (perturbed) pattern the pixels come from */
    for (bb = lo; bb < hi; ++bb) {
        uint64_t gidx = (uint64_t)(c->b0 + bb);
        uint32_t r[12];
        double roll, pitch, yaw, depth, fx, fy, R[9], t[3];
        int i;
        philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, 0u, k0, k1, r);
        philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), 1u, 0u, k0, k1, r + 4);
        philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), 2u, 0u, k0, k1, r + 8);
        roll  = cfg->roll_center_deg + (-cfg->angle_range_deg + 2.0 * cfg->angle_range_deg * u01(r[0], r[1]));
        pitch = cfg->pitch_center_deg + (-cfg->angle_range_deg + 2.0 * cfg->angle_range_deg * u01(r[2], r[3]));
        yaw   = cfg->yaw_center_deg + (-cfg->angle_range_deg + 2.0 * cfg->angle_range_deg * u01(r[4], r[5]));
        depth = cfg->depth_min_m + (cfg->depth_max_m - cfg->depth_min_m) * u01(r[6], r[7]);
        fx    = -cfg->fov_max_deg + 2.0 * cfg->fov_max_deg * u01(r[8], r[9]);
        fy    = -cfg->fov_max_deg + 2.0 * cfg->fov_max_deg * u01(r[10], r[11]);
        t[0] = depth * tan(fx * DEG2RAD); t[1] = depth * tan(fy * DEG2RAD); t[2] = depth;
        pnp_oracle_R_from_euler(roll, yaw, pitch, 1, R);
        memcpy(Pp, c->P, sizeof(double) * (size_t)n * 3);
        if (c->perturb_radius > 0.0) {
            /* unit_vec of a 3 (n - 1) standard normal vector times the radius, one landmark fixed
             * (face_variation_test.py:321-345; the i.i.d. draws make the reshape order immaterial) */
            double ss = 0.0, sc, z[3];
            for (i = 0; i < n; ++i) {
                if (i == c->fixed_idx) continue;
                perturb_normals(gidx, i, k0, k1, z);
                ss += z[0] * z[0] + z[1] * z[1] + z[2] * z[2];
            }
            sc = c->perturb_radius / sqrt(ss);
            for (i = 0; i < n; ++i) {
                int k;
                z[0] = z[1] = z[2] = 0.0;
                if (i != c->fixed_idx) perturb_normals(gidx, i, k0, k1, z);
                for (k = 0; k < 3; ++k) {
                    Pp[3 * i + k] += sc * z[k];
                    if (c->perturb) c->perturb[((size_t)bb * n + i) * 3 + k] = sc * z[k];
                }
            }
        }
        pnp_oracle_project(n, Pp, c->K, R, t, 0, 1.0, uvw);
        for (i = 0; i < n; ++i) {
            double u = uvw[3 * i], v = uvw[3 * i + 1];
            /* quantise first (random_stress_test.py:290), then add noise (LM_noise_test.py:272-286) */
            if (cfg->is_quantized) { u = rint(u / cfg->quantize_q) * cfg->quantize_q; v = rint(v / cfg->quantize_q) * cfg->quantize_q; }
            if (cfg->noise_sigma_px > 0.0) {
                uint32_t g[4];
                double a, bq, rad;
                philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), 16u + (uint32_t)i, 1u, k0, k1, g);
                a = u01(g[0], g[1]); bq = u01(g[2], g[3]);
                rad = sqrt(-2.0 * log(1.0 - a));                 /* Box-Muller, 1-a in (0,1] */
                u += cfg->noise_sigma_px * rad * cos(2.0 * M_PI * bq);
                v += cfg->noise_sigma_px * rad * sin(2.0 * M_PI * bq);
            }
            c->uv[(size_t)bb * n * 2 + 2 * i] = u; c->uv[(size_t)bb * n * 2 + 2 * i + 1] = v;
        }
        if (c->gt) { c->gt[bb * 4] = depth; c->gt[bb * 4 + 1] = roll; c->gt[bb * 4 + 2] = pitch; c->gt[bb * 4 + 3] = yaw; }
        if (c->R_gt) memcpy(c->R_gt + bb * 9, R, sizeof(R));
        if (c->t_gt) memcpy(c->t_gt + bb * 3, t, sizeof(t));
    }
    free(uvw);
}

/* gt[b] = (distance [m], roll, pitch, yaw [deg]); problems b0 .. b0+B-1 of the global stream */
ORACLE_API void pnp_oracle_synth(int64_t b0, int64_t B, int n, const double *P, const double *K,
                                 const oracle_synth_t *cfg, double *uv, double *gt, double *R_gt,
                                 double *t_gt, int n_threads)
{
    synth_ctx_t c;
    c.b0 = b0; c.n = n; c.P = P; c.K = K; c.cfg = cfg; c.uv = uv; c.gt = gt; c.R_gt = R_gt; c.t_gt = t_gt;
    c.perturb_radius = 0.0; c.fixed_idx = -1; c.perturb = NULL;
    parallel_for(B, n_threads, 256, synth_range, &c);
}

/* the face_variation_test.py workload: pixels of a perturbed pattern; perturb [B,n,3] */
ORACLE_API void pnp_oracle_synth_face_variation(int64_t b0, int64_t B, int n, const double *P, const double *K,
                                                const oracle_synth_t *cfg, double radius_m, int fixed_index,
                                                double *uv, double *gt, double *R_gt, double *t_gt,
                                                double *perturb, int n_threads)
{
    synth_ctx_t c;
    c.b0 = b0; c.n = n; c.P = P; c.K = K; c.cfg = cfg; c.uv = uv; c.gt = gt; c.R_gt = R_gt; c.t_gt = t_gt;
    c.perturb_radius = radius_m; c.fixed_idx = fixed_index; c.perturb = perturb;
    parallel_for(B, n_threads, 256, synth_range, &c);
}
