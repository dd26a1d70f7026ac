// fivept.cuh -- Nister's five-point minimal solver for the essential matrix (cv::findEssentialMat's runKernel,
// calib3d five-point.cpp, un-vendored OpenCV 4.12: conanfile.txt:2; the reference calls it at
// src/frontend/pose_estimator.cpp:42).  Same mathematics as OpenCV with its own null-space basis and a direct
// real-root finder: the candidate E's agree with OpenCV's to rounding, their order within one sample may differ.
// __host__ __device__ so tests/native/host_exact.cpp can check it on the CPU against the oracle without a GPU.
#pragma once
#include <math.h>

#include "exact.cuh"

namespace slamcu {

constexpr int kMaxModels = 10;  // essential matrices per sample

// linear polynomial (x, y, z, 1) products; monomial orders documented in DESIGN.md (x, y, z, 1 | xx, yy, zz, xy, xz, yz, x, y, z, 1 | Nister order)
SLAM_HDN void mul11(const double* a, const double* b, double* out /*10, accumulated*/, double sign) {
    constexpr int T11[4][4] = {{0, 3, 4, 6}, {3, 1, 5, 7}, {4, 5, 2, 8}, {6, 7, 8, 9}};
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) out[T11[i][j]] += sign * (a[i] * b[j]);
}
SLAM_HDN void mul21(const double* a, const double* b, double* out /*20, accumulated*/) {
    constexpr int T21[10][4] = {{0, 2, 4, 5},    {3, 1, 6, 7},    {10, 13, 16, 17}, {2, 3, 8, 9},     {4, 8, 10, 11},
                                {8, 6, 13, 14},  {5, 9, 11, 12},  {9, 7, 14, 15},   {11, 14, 17, 18}, {12, 15, 18, 19}};
    for (int i = 0; i < 10; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) out[T21[i][j]] += a[i] * b[j];
}

// c = a * b for dense univariate polynomials, highest power first; na, nb = number of coefficients
SLAM_HDN void polymul(const double* a, int na, const double* b, int nb, double* c, double sign, bool clear) {
    if (clear)
        for (int i = 0; i < na + nb - 1; i++) c[i] = 0.0;
    for (int i = 0; i < na; i++)
        for (int j = 0; j < nb; j++) c[i + j] += sign * (a[i] * b[j]);
}

SLAM_HDN double polyval(const double* a, int n, double x) {
    double v = a[0];
    for (int i = 1; i < n; i++) v = v * x + a[i];
    return v;
}

// ---- real roots of a real polynomial ------------------------------------------------------------------------------
// OpenCV runs 300 Durand-Kerner sweeps in complex arithmetic and keeps the roots with |Im| <= 1e-10
// (five-point.cpp); only the real roots matter, so they are isolated directly: a real root of f lies between two
// consecutive critical points of f (roots of f'), hence the chain f^(n-1), ..., f', f is solved bottom-up, each level
// bracketing its roots by sign changes between the previous level's roots and finishing them with a bracketed
// Newton iteration.  Roots with |x| <= 1 come from p itself, the others from the reversed polynomial y^n p(1/y) on
// (-1, 1), so every bracket is bounded.  ~1e4 flops per degree-10 polynomial instead of ~1e6.
//
// T holds the scaled derivatives f_k = f^(k) / k! (binomial weights, ascending powers): f_k has n-k+1 coefficients
// starting at T[toff(k)], and f_k' = (k+1) f_{k+1}.
SLAM_HDN int poly_toff(int k, int n) { return k * (n + 1) - (k * (k - 1)) / 2; }

SLAM_HDN double poly_eval_asc(const double* a, int deg, double x) {
    double v = a[deg];
    for (int i = deg - 1; i >= 0; i--) v = v * x + a[i];
    return v;
}

// root of f (ascending coefficients, degree deg; df = derivative coefficients / scale) inside the bracket (lo, hi)
// with f(lo) * f(hi) < 0: Newton steps guarded by bisection, then two free Newton steps for the last bits.
SLAM_HDN double poly_bracketed_root(const double* f, const double* df, int deg, double dscale, double lo, double hi, double flo) {
    if (flo > 0.0) { const double t = lo; lo = hi; hi = t; }  // orient: f(lo) < 0 < f(hi)
    double x = 0.5 * (lo + hi), dxold = fabs(hi - lo), dx = dxold;
    double fx = poly_eval_asc(f, deg, x), dfx = dscale * poly_eval_asc(df, deg - 1, x);
    for (int it = 0; it < 200; it++) {
        const bool out = ((x - hi) * dfx - fx) * ((x - lo) * dfx - fx) > 0.0;
        if (out || fabs(2.0 * fx) > fabs(dxold * dfx)) {
            dxold = dx;
            dx = 0.5 * (hi - lo);
            x = lo + dx;
            if (x == lo) break;
        } else {
            dxold = dx;
            dx = fx / dfx;
            const double t = x;
            x -= dx;
            if (t == x) break;
        }
        if (fabs(dx) < 1e-13 * fmax(fabs(x), 1e-3)) break;
        fx = poly_eval_asc(f, deg, x);
        dfx = dscale * poly_eval_asc(df, deg - 1, x);
        if (fx == 0.0) return x;
        if (fx < 0.0) lo = x; else hi = x;
    }
    for (int k = 0; k < 2; k++) {
        fx = poly_eval_asc(f, deg, x);
        dfx = dscale * poly_eval_asc(df, deg - 1, x);
        if (dfx != 0.0 && fx != 0.0) {
            const double xn = x - fx / dfx;
            if ((xn - lo) * (xn - hi) <= 0.0) x = xn;  // stay inside the final bracket
        }
    }
    return x;
}

// real roots of a (ascending, degree n <= 10, a[n] != 0) in [-1, 1] (closed) or (-1, 1) (open); ascending order
SLAM_HDN int poly_roots_unit(const double* a, int n, bool closed, double* out) {
    double T[66];
    for (int i = 0; i <= n; i++) T[i] = a[i];
    for (int k = 1; k < n; k++) {  // f_k[j] = a[j + k] * C(j + k, k) = f_{k-1}[j + 1] * (j + 1) / k
        const double* prev = T + poly_toff(k - 1, n);
        double* cur = T + poly_toff(k, n);
        for (int j = 0; j <= n - k; j++) cur[j] = prev[j + 1] * (double)(j + 1) / (double)k;
    }
    double cp[10], nx[10];
    int m = 0;
    for (int k = n - 1; k >= 0; k--) {
        const double* f = T + poly_toff(k, n);
        const int deg = n - k;
        const double* df = deg >= 2 ? T + poly_toff(k + 1, n) : f + 1;  // degree 1: f' = f[1]
        const double dscale = deg >= 2 ? (double)(k + 1) : 1.0;
        const bool top = k == 0;
        int cnt = 0;
        double xa = -1.0, fa = poly_eval_asc(f, deg, xa);
        if (top && closed && fa == 0.0) nx[cnt++] = xa;
        for (int i = 0; i <= m; i++) {
            const double xb = i < m ? cp[i] : 1.0;
            const double fb = poly_eval_asc(f, deg, xb);
            if (xb > xa) {
                if ((fa < 0.0 && fb > 0.0) || (fa > 0.0 && fb < 0.0)) {
                    nx[cnt++] = poly_bracketed_root(f, df, deg, dscale, xa, xb, fa);
                }
                if (fb == 0.0 && (i < m || (top && closed))) nx[cnt++] = xb;
            }
            xa = xb;
            fa = fb;
        }
        m = cnt;
        for (int i = 0; i < m; i++) cp[i] = nx[i];
    }
    for (int i = 0; i < m; i++) out[i] = cp[i];
    return m;
}

// Real roots of a real polynomial (highest power first, n_in coefficients, degree <= 10).
SLAM_HDN int real_roots(const double* c_in, int n_in, double* out) {
    int lead = 0;
    while (lead < n_in && c_in[lead] == 0.0) lead++;
    const int n = n_in - lead - 1;  // degree
    if (n < 1) return 0;
    double asc[11], rev[11];
    for (int i = 0; i <= n; i++) {
        rev[i] = c_in[lead + i] / c_in[lead];  // reversed polynomial, ascending: y^n p(1/y)
        asc[n - i] = rev[i];
    }
    int cnt = poly_roots_unit(asc, n, true, out);
    int nr = n;
    while (nr > 0 && rev[nr] == 0.0) nr--;  // p(0) = 0: the reversed polynomial loses a degree
    if (nr >= 1) {
        double ys[10];
        const int my = poly_roots_unit(rev, nr, false, ys);
        for (int i = 0; i < my && cnt < n; i++)
            if (ys[i] != 0.0) out[cnt++] = 1.0 / ys[i];
    }
    return cnt;
}

// Nister's 5-point solver.  x1, x2: 5 normalised correspondences.  models: up to kMaxModels row-major 3x3, |E|_F = 1.
SLAM_HDN int five_point(const double (*x1)[2], const double (*x2)[2], double* models) {
    // ---- null space of the 5x9 epipolar system (row-major E), full-pivot Gauss-Jordan + twice MGS
    double Q[5][9];
    int cols[9];
    for (int p = 0; p < 5; p++) {
        const double a[3] = {x1[p][0], x1[p][1], 1.0}, b[3] = {x2[p][0], x2[p][1], 1.0};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) Q[p][3 * i + j] = b[i] * a[j];
    }
    for (int k = 0; k < 9; k++) cols[k] = k;
    for (int r = 0; r < 5; r++) {
        int pr = r, pc = r;
        double best = -1.0;
        for (int i = r; i < 5; i++)
            for (int j = r; j < 9; j++)
                if (fabs(Q[i][j]) > best) { best = fabs(Q[i][j]); pr = i; pc = j; }
        if (best <= 0.0) return 0;
        for (int j = 0; j < 9; j++) { const double t = Q[r][j]; Q[r][j] = Q[pr][j]; Q[pr][j] = t; }
        for (int i = 0; i < 5; i++) { const double t = Q[i][r]; Q[i][r] = Q[i][pc]; Q[i][pc] = t; }
        { const int t = cols[r]; cols[r] = cols[pc]; cols[pc] = t; }
        const double piv = Q[r][r];
        for (int j = 0; j < 9; j++) Q[r][j] = Q[r][j] / piv;
        for (int k = 0; k < 5; k++) {
            if (k == r) continue;
            const double fct = Q[k][r];
            for (int j = 0; j < 9; j++) Q[k][j] = Q[k][j] - fct * Q[r][j];
        }
    }
    double B4[4][9];
    for (int k = 0; k < 4; k++) {
        for (int j = 0; j < 9; j++) B4[k][j] = 0.0;
        for (int r = 0; r < 5; r++) B4[k][cols[r]] = -Q[r][5 + k];
        B4[k][cols[5 + k]] = 1.0;
    }
    for (int rep = 0; rep < 2; rep++)
        for (int k = 0; k < 4; k++) {
            for (int j = 0; j < k; j++) {
                double d = 0.0;
                for (int t = 0; t < 9; t++) d += B4[k][t] * B4[j][t];
                for (int t = 0; t < 9; t++) B4[k][t] = B4[k][t] - d * B4[j][t];
            }
            double nn = 0.0;
            for (int t = 0; t < 9; t++) nn += B4[k][t] * B4[k][t];
            nn = sqrt(nn);
            for (int t = 0; t < 9; t++) B4[k][t] = B4[k][t] / nn;
        }
    // ---- the ten cubic constraints: det(E) and (E E' - tr(E E')/2 I) E, for E = x B0 + y B1 + z B2 + B3
    double A[10][20];
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 20; j++) A[i][j] = 0.0;
    double Ep[9][4];
    for (int e = 0; e < 9; e++)
        for (int k = 0; k < 4; k++) Ep[e][k] = B4[k][e];
    {
        double t2[10];
        for (int j = 0; j < 10; j++) t2[j] = 0.0;
        mul11(Ep[1], Ep[5], t2, 1.0); mul11(Ep[2], Ep[4], t2, -1.0); mul21(t2, Ep[6], A[0]);
        for (int j = 0; j < 10; j++) t2[j] = 0.0;
        mul11(Ep[2], Ep[3], t2, 1.0); mul11(Ep[0], Ep[5], t2, -1.0); mul21(t2, Ep[7], A[0]);
        for (int j = 0; j < 10; j++) t2[j] = 0.0;
        mul11(Ep[0], Ep[4], t2, 1.0); mul11(Ep[1], Ep[3], t2, -1.0); mul21(t2, Ep[8], A[0]);
    }
    {
        double L[3][3][10];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                for (int t = 0; t < 10; t++) L[i][j][t] = 0.0;
                for (int k = 0; k < 3; k++) mul11(Ep[3 * i + k], Ep[3 * j + k], L[i][j], 1.0);
            }
        double tr[10];
        for (int t = 0; t < 10; t++) tr[t] = (L[0][0][t] + L[1][1][t]) + L[2][2][t];
        for (int i = 0; i < 3; i++)
            for (int t = 0; t < 10; t++) L[i][i][t] = L[i][i][t] - 0.5 * tr[t];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 3; k++) mul21(L[i][k], Ep[3 * k + j], A[1 + 3 * i + j]);
    }
    // ---- Gauss-Jordan on the first ten columns, partial pivoting
    for (int col = 0; col < 10; col++) {
        int piv = col;
        double best = fabs(A[col][col]);
        for (int r = col + 1; r < 10; r++)
            if (fabs(A[r][col]) > best) { best = fabs(A[r][col]); piv = r; }
        if (best == 0.0) return 0;
        if (piv != col)
            for (int j = 0; j < 20; j++) { const double t = A[col][j]; A[col][j] = A[piv][j]; A[piv][j] = t; }
        const double d = A[col][col];
        for (int j = 0; j < 20; j++) A[col][j] = A[col][j] / d;
        for (int r = 0; r < 10; r++) {
            if (r == col) continue;
            const double fct = A[r][col];
            if (fct == 0.0) continue;
            for (int j = 0; j < 20; j++) A[r][j] = A[r][j] - fct * A[col][j];
        }
    }
    // ---- rows x^2z - z x^2, y^2z - z y^2, xyz - z xy  ->  B(z) (3 x 3 polynomial matrix), det B(z) of degree 10
    double bx[3][4], by[3][4], b1[3][5];
    for (int i = 0; i < 3; i++) {
        const double* a = &A[4 + 2 * i][10];
        const double* b = &A[5 + 2 * i][10];
        bx[i][0] = 0.0 - b[0]; bx[i][1] = a[0] - b[1]; bx[i][2] = a[1] - b[2]; bx[i][3] = a[2] - 0.0;
        by[i][0] = 0.0 - b[3]; by[i][1] = a[3] - b[4]; by[i][2] = a[4] - b[5]; by[i][3] = a[5] - 0.0;
        b1[i][0] = 0.0 - b[6]; b1[i][1] = a[6] - b[7]; b1[i][2] = a[7] - b[8]; b1[i][3] = a[8] - b[9]; b1[i][4] = a[9] - 0.0;
    }
    double det[11], m[7];
    polymul(bx[1], 4, by[2], 4, m, 1.0, true); polymul(by[1], 4, bx[2], 4, m, -1.0, false);
    polymul(b1[0], 5, m, 7, det, 1.0, true);
    polymul(bx[0], 4, by[2], 4, m, 1.0, true); polymul(by[0], 4, bx[2], 4, m, -1.0, false);
    polymul(b1[1], 5, m, 7, det, -1.0, false);
    polymul(bx[0], 4, by[1], 4, m, 1.0, true); polymul(by[0], 4, bx[1], 4, m, -1.0, false);
    polymul(b1[2], 5, m, 7, det, 1.0, false);
    double zs[10];
    const int nz = real_roots(det, 11, zs);
    int count = 0;
    for (int k = 0; k < nz && count < kMaxModels; k++) {
        const double z = zs[k];
        const double z2 = z * z, z3 = z2 * z, z4 = z3 * z;
        double Bz[3][3];
        for (int i = 0; i < 3; i++) {
            Bz[i][0] = ((bx[i][0] * z3 + bx[i][1] * z2) + bx[i][2] * z) + bx[i][3];
            Bz[i][1] = ((by[i][0] * z3 + by[i][1] * z2) + by[i][2] * z) + by[i][3];
            Bz[i][2] = (((b1[i][0] * z4 + b1[i][1] * z3) + b1[i][2] * z2) + b1[i][3] * z) + b1[i][4];
        }
        // null vector of the rank-2 matrix: the largest cross product of two rows
        double v[3] = {0, 0, 0}, vn = -1.0;
        const int pa[3] = {0, 0, 1}, pb[3] = {1, 2, 2};
        for (int c = 0; c < 3; c++) {
            const double* r0 = Bz[pa[c]];
            const double* r1 = Bz[pb[c]];
            const double cx = r0[1] * r1[2] - r0[2] * r1[1], cy = r0[2] * r1[0] - r0[0] * r1[2],
                         cz = r0[0] * r1[1] - r0[1] * r1[0];
            const double nn = (cx * cx + cy * cy) + cz * cz;
            if (nn > vn) { vn = nn; v[0] = cx; v[1] = cy; v[2] = cz; }
        }
        if (fabs(v[2]) < 1e-10 * sqrt(vn)) continue;
        const double x = v[0] / v[2], y = v[1] / v[2];
        double e[9], nn = 0.0;
        for (int t = 0; t < 9; t++) {
            e[t] = ((x * B4[0][t] + y * B4[1][t]) + z * B4[2][t]) + B4[3][t];
            nn += e[t] * e[t];
        }
        nn = sqrt(nn);
        for (int t = 0; t < 9; t++) models[count * 9 + t] = e[t] / nn;
        count++;
    }
    return count;
}

}  // namespace slamcu
