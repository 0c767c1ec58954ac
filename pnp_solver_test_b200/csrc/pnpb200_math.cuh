// pnpb200_math.cuh -- register-resident small linear algebra for the per-problem systems.
//
// Everything here is fully unrolled with compile-time indices so that the 6x6 / 12x12
// normal equations, their factorisations and the 3x3 SVD live in registers (no local
// memory).  Templated on the scalar type: double = parity mode, float = FP32 mode.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pnpb200 {

#define PNP_DEV __device__ __forceinline__

// packed upper-triangular storage of a symmetric N x N matrix, row-major: (i,j), i <= j
template <int N>
PNP_DEV constexpr int sidx(int i, int j) { return i * N - (i * (i - 1)) / 2 + (j - i); }
template <int N>
PNP_DEV constexpr int sym(int i, int j) { return i <= j ? sidx<N>(i, j) : sidx<N>(j, i); }

template <typename T> PNP_DEV T t_sqrt(T x);
template <> PNP_DEV double t_sqrt<double>(double x) { return sqrt(x); }
template <> PNP_DEV float t_sqrt<float>(float x) { return sqrtf(x); }
// Branch-free reciprocal and reciprocal square root for normal, finite, non-zero arguments (pivots
// of SPD systems, squared norms).  The CUDA library versions (1.0 / x, __drcp_rn, sqrt) carry a
// slow-path branch for denormals and the exponent extremes; inside the fully unrolled 10 x 10
// factorisation every such branch ends a basic block and stops ptxas from scheduling the seed and
// its Newton steps in the shadow of the independent trailing updates.  Seed: MUFU.RCP64H /
// MUFU.RSQ64H (about 2^-20 relative); one third-order step brings it to ~2^-60 before rounding
// (faithfully rounded, not correctly rounded; verified on the device by
// tests/test_gpu_parity.py::test_fast_reciprocals).
template <typename T> PNP_DEV T t_rcp(T x);
template <> PNP_DEV double t_rcp<double>(double a)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    e = fma(e, e, e);                                     // e + e^2
    return fma(x, e, x);                                  // x (1 + e + e^2): error e^3 ~ 2^-60, <= 0.5003 ulp measured
}
template <> PNP_DEV float t_rcp<float>(float a)
{
    float x;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(a));   // MUFU.RCP, ~1 ulp
    return fmaf(x, fmaf(-a, x, 1.0f), x);
}
// 1 / sqrt(a)
template <typename T> PNP_DEV T t_rsqrt(T x);
template <> PNP_DEV double t_rsqrt<double>(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double ha = 0.5 * a;
    double e = fma(-ha * y, y, 0.5);                      // (1 - a y^2) / 2
    return fma(y, fma(1.5 * e, e, e), y);                 // y (1 + e' / 2 + 3 e'^2 / 8), e' = 2 e: third order
}
template <> PNP_DEV float t_rsqrt<float>(float a)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(a));  // MUFU.RSQ, ~2 ulp
    return fmaf(y, fmaf(-0.5f * a * y, y, 0.5f), y);
}
// sqrt(a) = a / sqrt(a) with one correction step (a > 0, normal)
template <typename T> PNP_DEV T t_sqrt_fast(T a, T rs);   // rs = t_rsqrt(a)
template <> PNP_DEV double t_sqrt_fast<double>(double a, double rs)
{
    const double s = a * rs;
    return fma(fma(-s, s, a), 0.5 * rs, s);
}
template <> PNP_DEV float t_sqrt_fast<float>(float a, float rs)
{
    const float s = a * rs;
    return fmaf(fmaf(-s, s, a), 0.5f * rs, s);
}
// branch-free sqrt for a >= 0 including exact zeros (distances).  The seed of a = 0 is +inf; an integer min
// on its high word turns it into the largest finite power of two (every other seed is far below it), so
// that s0 = 0 * y is 0 instead of NaN -- one ALU instruction instead of an FP64 one.  s0 = a y0 carries
// the seed's 2^-20 error e' = 1 - a y0^2; one third-order step s0 (1 + e'/2 + 3 e'^2 / 8) leaves ~e'^3:
// five FP64 instructions and the MUFU, ~1 ulp.
PNP_DEV double sqrt_nonneg(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    y = __hiloint2double(min(__double2hiint(y), 0x7fe00000), __double2loint(y));
    const double s0 = a * y;
    const double e = fma(-s0, y, 1.0);
    const double p = fma(0.375, e, 0.5) * e;
    return fma(s0, p, s0);
}
// sqrt of a non-negative number, zero included, without the slow-path branch of sqrt()
template <typename T> PNP_DEV T t_sqrt_nn(T a);
template <> PNP_DEV double t_sqrt_nn<double>(double a) { return sqrt_nonneg(a); }
template <> PNP_DEV float t_sqrt_nn<float>(float a) { return sqrtf(a); }
template <typename T> PNP_DEV T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> PNP_DEV T t_fma(T a, T b, T c);
template <> PNP_DEV double t_fma<double>(double a, double b, double c) { return fma(a, b, c); }
template <> PNP_DEV float t_fma<float>(float a, float b, float c) { return fmaf(a, b, c); }

// nu = K^-1 [u, v, 1]^T, first two rows (f2_get_B_xy, PNP_SOLVER_LIB.py:3305): two FMAs per coordinate, the same in every kernel
template <typename T>
PNP_DEV void normalise_px(T u, T v, T k00, T k01, T k02, T k10, T k11, T k12, T& bx, T& by)
{
    bx = t_fma(k00, u, t_fma(k01, v, k02));
    by = t_fma(k10, u, t_fma(k11, v, k12));
}

// In-place LDL^T of a packed SPD matrix.  On return A(j,j) holds 1/d_j and A(j,i), i > j,
// holds L(i,j).  Replaces the SVD-based np.linalg.pinv of the reference for the SPD systems
// (PNP_SOLVER_LIB.py:2675, :2887, :2924); equivalent to ~1e-13 at the condition numbers seen
// (SURVEY.md 7.3).
template <typename T, int N>
PNP_DEV void ldlt_factor(T (&A)[N * (N + 1) / 2])
{
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const T inv = t_rcp<T>(A[sidx<N>(j, j)]);
        // The next pivot is the serial chain of the factorisation (rcp -> scale -> update -> rcp ...):
        // its update uses a^2 / d, whose square does not wait for the reciprocal.
        if (j + 1 < N) {
            const T a = A[sidx<N>(j, j + 1)];
            A[sidx<N>(j + 1, j + 1)] = t_fma(-(a * a), inv, A[sidx<N>(j + 1, j + 1)]);
        }
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            const T lij = A[sidx<N>(j, i)] * inv;
#pragma unroll
            for (int r = i; r < N; ++r)
                if (!(i == j + 1 && r == j + 1)) A[sidx<N>(i, r)] = t_fma(-lij, A[sidx<N>(j, r)], A[sidx<N>(i, r)]);
            A[sidx<N>(j, i)] = lij;
        }
        A[sidx<N>(j, j)] = inv;
    }
}

// Solve (L D L^T) x = b in place, A as left by ldlt_factor.
template <typename T, int N>
PNP_DEV void ldlt_solve(const T (&A)[N * (N + 1) / 2], T (&b)[N])
{
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int i = j + 1; i < N; ++i) b[i] = t_fma(-A[sidx<N>(j, i)], b[j], b[i]);
    }
#pragma unroll
    for (int j = 0; j < N; ++j) b[j] *= A[sidx<N>(j, j)];
#pragma unroll
    for (int j = N - 1; j >= 0; --j) {
        // oldest unknowns first: the one computed last (b[j+1]) enters last, so the dependent chain
        // through the back substitution is N FMAs long instead of N (N - 1) / 2
#pragma unroll
        for (int i = N - 1; i > j; --i) b[j] = t_fma(-A[sidx<N>(j, i)], b[i], b[j]);
    }
}

// Explicit inverse of a packed SPD matrix, in place and without a second array: A <- A^-1 = X^T D^-1 X
// with X = L^-1 (unit lower).  (The first version formed the product in a temporary of the same size;
// for the 12 x 12 systems of EIF2 that alone was 78 more doubles of local memory per thread.)
template <typename T, int N>
PNP_DEV void spd_inverse(T (&A)[N * (N + 1) / 2])
{
    ldlt_factor<T, N>(A);
    // X = L^-1.  Stored in the strictly-upper slots: X(i,j), i > j, at (j,i).
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            // X(i,j) = -( L(i,j) + sum_{k=j+1}^{i-1} L(i,k) X(k,j) ); L(i,k) is still needed for
            // later columns j' < k, so write X over column j only (slots (j, *)), which are the
            // L(*, j) entries -- those are consumed for this (i, j) before being overwritten
            // only if we go in increasing i and read L(i,j) first.
            T acc = A[sidx<N>(j, i)];
#pragma unroll
            for (int k = j + 1; k < i; ++k) acc = t_fma(A[sidx<N>(k, i)], A[sidx<N>(j, k)], acc);
            A[sidx<N>(j, i)] = -acc;
        }
    }
    // Ainv(i,j) = sum_{k >= j} X(k,i) X(k,j) / d_k for i <= j, with X(k,k) = 1; X(k,i) sits at (i,k),
    // 1/d_k at (k,k).  Rows top to bottom, columns left to right: entry (i,j) reads (i,k) and (j,k)
    // for k >= j only -- to its right in its own row, and rows below that are still untouched.
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = i; j < N; ++j) {
            // k = j term: X(j,i) / d_j   (X(j,i) = 1 if i == j)
            T acc = (i == j) ? A[sidx<N>(j, j)] : A[sidx<N>(i, j)] * A[sidx<N>(j, j)];
#pragma unroll
            for (int k = j + 1; k < N; ++k)
                acc = t_fma(A[sidx<N>(i, k)] * A[sidx<N>(j, k)], A[sidx<N>(k, k)], acc);
            A[sidx<N>(i, j)] = acc;
        }
    }
}

// y = S x for packed symmetric S
template <typename T, int N>
PNP_DEV void sym_matvec(const T (&S)[N * (N + 1) / 2], const T (&x)[N], T (&y)[N])
{
#pragma unroll
    for (int i = 0; i < N; ++i) {
        T acc = T(0);
#pragma unroll
        for (int j = 0; j < N; ++j) acc = t_fma(S[sym<N>(i, j)], x[j], acc);
        y[i] = acc;
    }
}

// 3x3 singular value decomposition by one-sided (Hestenes) Jacobi, enough of it to form
//   Rproj = U diag(1, 1, det(U V^T)) V^T   and   sigma_max
// as EKF2_reconstruct_R_t_m1 does with np.linalg.svd / det / norm(ord=2)
// (PNP_SOLVER_LIB.py:3509-3518, :3530).  G row-major.  det(U V^T) = sign(det G) when G is
// non-singular, and the -1 lands on the smallest singular value's pair.
template <typename T> PNP_DEV T t_sqrt_pos(T a) { return t_sqrt_fast<T>(a, t_rsqrt<T>(a)); }   // a > 0, normal

// Branch-free rotations: a rotation that is not needed (|gamma| <= tol sqrt(alpha beta)) becomes the
// identity by selects, divisions and square roots are the MUFU-seeded ones, and the sweep loop ends
// when no lane of the warp rotated -- the library division / sqrt slow paths and the per-lane
// divergence of the first version made this epilogue a third of k_iterate's time.
template <typename T>
PNP_DEV void svd3_project(const T (&G)[9], T (&R)[9], T& sigma_max)
{
    T W[9], V[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { W[i] = G[i]; V[i] = T(0); }
    V[0] = V[4] = V[8] = T(1);
    const T tol = (sizeof(T) == 8) ? T(2e-16) : T(1e-7);
    const T tol2 = tol * tol;
    for (int sweep = 0; sweep < 12; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
            const int i = (pr == 2) ? 1 : 0;
            const int j = (pr == 0) ? 1 : 2;
            T alpha = T(0), beta = T(0), gamma = T(0);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                alpha = t_fma(W[k * 3 + i], W[k * 3 + i], alpha);
                beta = t_fma(W[k * 3 + j], W[k * 3 + j], beta);
                gamma = t_fma(W[k * 3 + i], W[k * 3 + j], gamma);
            }
            const bool rot = gamma * gamma > tol2 * (alpha * beta);      // gamma != 0 and |gamma| > tol sqrt(alpha beta)
            rotated = rotated || rot;
            const T gs = rot ? gamma : T(1);
            const T zeta = (beta - alpha) * t_rcp<T>(gs + gs);
            const T tt0 = t_rcp<T>(t_abs(zeta) + t_sqrt_pos<T>(t_fma(zeta, zeta, T(1))));
            const T tt = (zeta >= T(0)) ? tt0 : -tt0;
            const T c0 = t_rsqrt<T>(t_fma(tt, tt, T(1)));
            const T c = rot ? c0 : T(1), s = rot ? c0 * tt : T(0);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const T wi = W[k * 3 + i], wj = W[k * 3 + j];
                W[k * 3 + i] = c * wi - s * wj;
                W[k * 3 + j] = s * wi + c * wj;
                const T vi = V[k * 3 + i], vj = V[k * 3 + j];
                V[k * 3 + i] = c * vi - s * vj;
                V[k * 3 + j] = s * vi + c * vj;
            }
        }
        if (!__any_sync(__activemask(), rotated)) break;
    }
    T sg[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        sg[j] = t_sqrt(W[j] * W[j] + W[3 + j] * W[3 + j] + W[6 + j] * W[6 + j]);
    const T det = G[0] * (G[4] * G[8] - G[5] * G[7]) - G[1] * (G[3] * G[8] - G[5] * G[6]) +
                  G[2] * (G[3] * G[7] - G[4] * G[6]);
    const T D = (det < T(0)) ? T(-1) : T(1);
    int jmin = 0;
    if (sg[1] < sg[jmin]) jmin = 1;
    if (sg[2] < sg[jmin]) jmin = 2;
    sigma_max = fmax(sg[0], fmax(sg[1], sg[2]));
#pragma unroll
    for (int e = 0; e < 9; ++e) R[e] = T(0);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const T f = ((j == jmin) ? D : T(1)) / sg[j];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) R[r * 3 + c] = t_fma(W[r * 3 + j] * f, V[c * 3 + j], R[r * 3 + c]);
    }
}

// get_Euler_from_rotation_matrix, PNP_SOLVER_LIB.py:4474-4517; out = (roll, yaw, pitch).
// Always evaluated in double (the trig is a negligible share; keeps FP32 mode's angles honest).
PNP_DEV void euler_from_R(const double (&R)[9], bool is_degree, double (&out)[3])
{
    const double kPi = 3.14159265358979323846;
    // |pi/2 - asin(|R7|)| <= 1e-7  <=>  cos(1e-7) <= |R7| <= 1  (asin is NaN above 1 and the test then fails, :4482)
    const double kGimbal = 0.999999999999995;              // cos(1e-7) = 1 - 5.0e-15
    const double a7 = fabs(R[7]);
    double th1, th2, th3;
    if (a7 >= kGimbal && a7 <= 1.0) {                       // gimbal lock (:4482)
        const double m = -R[7];
        const double sg = (m > 0.0) ? 1.0 : ((m < 0.0) ? -1.0 : 0.0);
        th2 = sg * (kPi / 2.0);
        th3 = 0.0;
        th1 = atan2(-R[3], R[0]);
    } else {
        th1 = atan2(R[1], R[4]);
        th3 = atan2(R[6], R[8]);
        // cos(atan2(y, x)) = x / hypot(x, y): no second trip through the trigonometric routines (:4490-4493)
        const double h1 = R[1] * R[1] + R[4] * R[4], h3 = R[6] * R[6] + R[8] * R[8];
        const double c1 = (h1 > 0.0) ? R[4] * t_rsqrt<double>(h1) : 1.0;
        const double c3 = (h3 > 0.0) ? R[8] * t_rsqrt<double>(h3) : 1.0;
        const double c2 = (fabs(c1) > fabs(c3)) ? (R[4] / c1) : (R[8] / c3);
        th2 = atan2(-R[7], c2);
    }
    double roll = th1, pitch = th2, yaw = th3;
    if (is_degree) {
        const double k = 180.0 / kPi;
        roll *= k; yaw *= k; pitch *= k;
    }
    out[0] = roll; out[1] = -yaw; out[2] = pitch;
}

// get_rotation_matrix_from_Euler, PNP_SOLVER_LIB.py:4442-4472 (R = E_roll E_pitch E_yaw, yaw negated)
PNP_DEV void R_from_euler(double roll, double yaw, double pitch, bool is_degree, double (&R)[9])
{
    const double kPi = 3.14159265358979323846;
    yaw = -yaw;
    if (is_degree) {
        const double k = kPi / 180.0;
        roll *= k; yaw *= k; pitch *= k;
    }
    double s1, c1, s2, c2, s3, c3;
    sincos(roll, &s1, &c1); sincos(yaw, &s2, &c2); sincos(pitch, &s3, &c3);
    const double e0 = c2, e1 = 0.0, e2 = -s2;
    const double e3 = s3 * s2, e4 = c3, e5 = s3 * c2;
    const double e6 = c3 * s2, e7 = -s3, e8 = c3 * c2;
    R[0] = c1 * e0 + s1 * e3; R[1] = c1 * e1 + s1 * e4; R[2] = c1 * e2 + s1 * e5;
    R[3] = -s1 * e0 + c1 * e3; R[4] = -s1 * e1 + c1 * e4; R[5] = -s1 * e2 + c1 * e5;
    R[6] = e6; R[7] = e7; R[8] = e8;
}

// sin and cos of a bounded angle (|x| < ~1e4 rad) without the library's large-argument slow path: the
// same scheme as sincos() -- k = rint(x 2/pi), two-term Cody-Waite reduction with FMAs, the fdlibm
// kernel polynomials on [-pi/4, pi/4], quadrant by selects -- but in ONE basic block, so that the three
// angles of a rotation interleave instead of running as three serial ~60-instruction chains.  ~1 ulp.
PNP_DEV void sincos_bounded(double x, double& s, double& c)
{
    const int k = __double2int_rn(x * 0.63661977236758134308);
    const double q = (double)k;
    double r = fma(-q, 1.57079632679489655800e+00, x);
    r = fma(-q, 6.12323399573676603587e-17, r);
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    ps = fma(ps, z, 2.75573137070700676789e-06);   pc = fma(pc, z, -2.75573143513906633035e-07);
    ps = fma(ps, z, -1.98412698298579493134e-04);  pc = fma(pc, z, 2.48015872894767294178e-05);
    ps = fma(ps, z, 8.33333333332248946124e-03);   pc = fma(pc, z, -1.38888888888741095749e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);  pc = fma(pc, z, 4.16666666666666019037e-02);
    const double sr = fma(r * z, ps, r);
    const double cr = fma(z * z, pc, fma(-0.5, z, 1.0));
    const double s0 = (k & 1) ? cr : sr, c0 = (k & 1) ? sr : cr;
    s = (k & 2) ? -s0 : s0;
    c = ((k + 1) & 2) ? -c0 : c0;
}

// R_from_euler for angles in degrees that are bounded (ground-truth poses): branch-free
PNP_DEV void R_from_euler_deg_bounded(double roll, double yaw, double pitch, double (&R)[9])
{
    const double k = 3.14159265358979323846 / 180.0;
    double s1, c1, s2, c2, s3, c3;
    sincos_bounded(roll * k, s1, c1); sincos_bounded(-yaw * k, s2, c2); sincos_bounded(pitch * k, s3, c3);
    const double e3 = s3 * s2, e5 = s3 * c2;
    R[0] = c1 * c2 + s1 * e3; R[1] = s1 * c3; R[2] = c1 * -s2 + s1 * e5;
    R[3] = -s1 * c2 + c1 * e3; R[4] = c1 * c3; R[5] = -s1 * -s2 + c1 * e5;
    R[6] = c3 * s2; R[7] = -s3; R[8] = c3 * c2;
}

}  // namespace pnpb200
