// fivept_warp.cuh -- the five-point minimal solver of fivept.cuh as a WARP-COOPERATIVE device routine.
//
// One thread per sample (fivept.cuh) keeps ~5 KB of matrices in local memory and runs ~3e4 dependent FP64 operations:
// ~0.8 ms of latency per RANSAC wave and a working set that falls out of L1.  Here one warp solves one sample with
// the matrices distributed over lanes / a 3.4 KB shared-memory scratch:
//   A  null space of the 5x9 epipolar system: lane = column, full-pivot Gauss-Jordan by shuffles, MGS twice with
//      butterfly dot products;
//   B  the ten cubic constraints: lane = one entry of E E' / one row of the 10x20 coefficient matrix;
//   C  Gauss-Jordan on the 10x20 matrix: lane = column (registers), pivot column broadcast by shuffles;
//   D  det B(z): lane = cofactor, then lane = coefficient;
//   E  real roots: half-warp 0 takes p on [-1, 1], half-warp 1 the reversed polynomial on (-1, 1); per level of the
//      derivative chain lane = bracket (bracketed Newton, fivept.cuh), ballot compaction of the roots;
//   F  back-substitution: lane = root, ballot compaction of the models.
// Same mathematics and tolerances as fivept.cuh (and the same null-space basis order: the conditioning of the 10x10
// elimination depends on it); a few summation orders differ and divisions / square roots are reciprocal seeds refined by
// Newton steps, so the two agree to rounding (tests/test_gpu_essential.py compares both with the oracle solver).
#pragma once
#include "fivept.cuh"

namespace slamcu {

constexpr int kFiveptScratchDoubles = 424;  // per warp

namespace fpw {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double sel5(const double (&q)[5], int i) {
    return i == 0 ? q[0] : i == 1 ? q[1] : i == 2 ? q[2] : i == 3 ? q[3] : q[4];
}

// butterfly sum over lanes 0..15 of each half-warp (identical result in every lane of the half)
__device__ __forceinline__ double half_sum(double v) {
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
    return v;
}

__device__ __forceinline__ void mul11_acc(const double* a, const double* b, double (&out)[10], double sign) {
    constexpr int T11[4][4] = {{0, 3, 4, 6}, {3, 1, 5, 7}, {4, 5, 2, 8}, {6, 7, 8, 9}};
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) out[T11[i][j]] += sign * (a[i] * b[j]);
}

__device__ __forceinline__ void mul21_acc(const double* a, const double* b, double (&out)[20]) {
    constexpr int T21[10][4] = {{0, 2, 4, 5},    {3, 1, 6, 7},    {10, 13, 16, 17}, {2, 3, 8, 9},     {4, 8, 10, 11},
                                {8, 6, 13, 14},  {5, 9, 11, 12},  {9, 7, 14, 15},   {11, 14, 17, 18}, {12, 15, 18, 19}};
#pragma unroll
    for (int i = 0; i < 10; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) out[T21[i][j]] += a[i] * b[j];
}

// a / b for the Newton iterations: reciprocal seed (MUFU.RCP64H, ~20 bits) refined by two Newton steps and one
// residual correction -- ~8 instructions and half the latency of the IEEE division; the last bit may differ from it,
// which only moves an iterate by an ulp.  b = 0 / inf / NaN propagate to a NaN quotient: the caller then bisects.
__device__ __forceinline__ double fast_div(double a, double b) {
    double r;
#if defined(SLAM_SIMT_EMULATION)
    r = 1.0 / b;  // CPU emulation in the tests: any seed converges to the same refined value to an ulp
#else
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
#endif
    double e = __fma_rn(-b, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-b, r, 1.0);
    r = __fma_rn(r, e, r);
    const double q = a * r;
    return __fma_rn(r, __fma_rn(-b, q, a), q);
}

// 1 / sqrt(s) for the normalisations: rsqrt seed refined by two Newton steps (relative error ~1e-16)
__device__ __forceinline__ double fast_rsqrt(double s) {
    double y;
#if defined(SLAM_SIMT_EMULATION)
    y = 1.0 / sqrt(s);
#else
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
#endif
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const double e = __fma_rn(-(s * y), y, 1.0);  // 1 - s y^2
        y = __fma_rn(0.5 * y, e, y);
    }
    return y;
}

// One level of the derivative chain for both half-warps in lockstep: f = ascending coefficients of a degree-DEG
// polynomial (per half), cp[0..m) = the roots of f' inside the interval (ascending), closed = include the end points
// -1 and 1 themselves.  Writes the roots of f inside the interval to nx (ascending) and returns their number (uniform
// within the half).  Lane hl of a half owns end point e_hl (e_0 = -1, e_i = cp[i-1], e_(m+1) = 1) and the bracket
// (e_hl, e_(hl+1)); all brackets run the same bracketed Newton iteration with selects instead of branches.
template <int DEG>
__device__ __forceinline__ void horner_pd(const double (&c)[DEG + 1], double x, double& p, double& dp) {
    p = c[DEG];
    dp = 0.0;
#pragma unroll
    for (int i = DEG - 1; i >= 0; i--) {
        dp = dp * x + p;
        p = p * x + c[i];
    }
}

template <int DEG>
__device__ __noinline__ int root_level(const double* f, const double* cp, double* nx, int m, bool top, bool closed) {
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
    double c[DEG + 1];
#pragma unroll
    for (int i = 0; i <= DEG; i++) c[i] = f[i];
    const bool has_pt = hl <= m + 1;
    const double xa = !has_pt ? 0.0 : hl == 0 ? -1.0 : hl <= m ? cp[hl - 1] : 1.0;
    double fa, dfa;
    horner_pd<DEG>(c, xa, fa, dfa);
    const double xb = __shfl_down_sync(FULL, xa, 1, 16), fb = __shfl_down_sync(FULL, fa, 1, 16);
    const bool brk = hl <= m;  // lane owns a bracket
    const bool z0 = brk && hl == 0 && top && closed && fa == 0.0;
    const bool wide = brk && xb > xa;
    const bool sc = wide && ((fa < 0.0 && fb > 0.0) || (fa > 0.0 && fb < 0.0));
    const bool zb = wide && fb == 0.0 && (hl < m || (top && closed));
    // lower levels only provide brackets for the next one: 1e-8 is plenty (OpenCV's own root finder cannot resolve
    // closer pairs); the top level is polished to the last bits afterwards
    const double tol = top ? 1e-10 : 1e-8;
    double lo = sc ? xa : 0.0, hi = sc ? xb : 0.0;
    const bool up = fa < 0.0;  // f(lo) < 0 < f(hi)
    double x = sc ? xa - fa * fast_div(xb - xa, fb - fa) : 0.0;  // regula falsi start
    if (!(x > lo && x < hi)) x = 0.5 * (lo + hi);
    double dxold = hi - lo;
    bool done = !sc;
    for (int it = 0; it < 128; it++) {
        double fx, dfx;
        horner_pd<DEG>(c, x, fx, dfx);
        if ((fx < 0.0) == up) lo = x; else hi = x;  // f(x) has the sign of f(lo): move lo
        const double xn = x - fast_div(fx, dfx);
        const bool newton = xn > lo && xn < hi && fabs(2.0 * fx) <= fabs(dxold * dfx);
        const double xnew = newton ? xn : 0.5 * (lo + hi);
        const double dx = fabs(xnew - x);
        const bool stop = fx == 0.0 || dx <= tol * fmax(fabs(xnew), 1e-3) || hi - lo <= 4e-16 * fmax(fabs(xnew), 1e-300);
        dxold = done ? dxold : dx;
        x = (done || fx == 0.0) ? x : xnew;
        done = done || stop;
        if (!__any_sync(FULL, !done)) break;
    }
    if (top) {  // two free Newton steps for the last bits, kept inside the final bracket
#pragma unroll
        for (int r = 0; r < 2; r++) {
            double fx, dfx;
            horner_pd<DEG>(c, x, fx, dfx);
            const double xn = x - fast_div(fx, dfx);
            if (sc && fx != 0.0 && dfx != 0.0 && xn >= lo && xn <= hi) x = xn;
        }
    }
    const unsigned hm = 0xffffu << (16 * half), lt = (1u << lane) - 1u;
    const unsigned b0 = __ballot_sync(FULL, z0) & hm, b1 = __ballot_sync(FULL, sc) & hm, b2 = __ballot_sync(FULL, zb) & hm;
    int pos = __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt);
    if (z0) nx[pos++] = -1.0;
    if (sc) nx[pos++] = x;
    if (zb) nx[pos++] = xb;
    __syncwarp();
    return __popc(b0) + __popc(b1) + __popc(b2);
}

__device__ __forceinline__ double binom(int n, int k) {  // C(n, k), n <= 10
    double v = 1.0;
    for (int i = 1; i <= k; i++) v = v * (double)(n - k + i) / (double)i;  // exact: every partial product is an integer
    return v;
}

}  // namespace fpw

// x1, x2: the problem's normalised correspondences, idx: the sample's 5 indices into them.  S: this warp's scratch (kFiveptScratchDoubles
// doubles of shared memory).  models: up to kMaxModels row-major 3x3 matrices, |E|_F = 1 (global or shared).
// Must be called by all 32 lanes of a warp; returns the number of models (uniform).
__device__ inline int five_point_warp(const double2* x1, const double2* x2, const int* idx, double* S, double* models) {
    using namespace fpw;
    const int lane = threadIdx.x & 31;
    double* B4 = S;            // [4][9]   null-space basis, kept to the end
    double* R = S + 36;        // 388 doubles reused by the stages
    // ================= A: null space of the 5x9 system, lane c < 9 holds column c =====================================
    double q[5];
    {
        const int c = lane < 9 ? lane : 0, i = c / 3, j = c - 3 * i;
#pragma unroll
        for (int p = 0; p < 5; p++) {
            const double2 a = x1[idx[p]], b = x2[idx[p]];
            const double av = j == 0 ? a.x : j == 1 ? a.y : 1.0, bv = i == 0 ? b.x : i == 1 ? b.y : 1.0;
            q[p] = lane < 9 ? bv * av : 0.0;
        }
    }
    unsigned free_cols = 0x1ffu, used_rows = 0;
    int prow[5], pcol[5];
    // the column permutation fivept.cuh's swaps would produce (nibble p = column at position p): the free columns are
    // taken in that order so that both solvers build the same basis (the conditioning of the elimination depends on it)
    unsigned long long perm = 0x876543210ULL;
#pragma unroll
    for (int r = 0; r < 5; r++) {
        double best = -1.0;
        int bi = 0, bl = lane;
        if (lane < 9 && ((free_cols >> lane) & 1u)) {
#pragma unroll
            for (int i = 0; i < 5; i++)
                if (!((used_rows >> i) & 1u) && fabs(q[i]) > best) { best = fabs(q[i]); bi = i; }
        }
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) {
            const double ob = __shfl_xor_sync(FULL, best, off);
            const int oi = __shfl_xor_sync(FULL, bi, off), ol = __shfl_xor_sync(FULL, bl, off);
            if (ob > best || (ob == best && ol < bl)) { best = ob; bi = oi; bl = ol; }
        }
        best = __shfl_sync(FULL, best, 0);
        bi = __shfl_sync(FULL, bi, 0);
        bl = __shfl_sync(FULL, bl, 0);
        if (!(best > 0.0)) return 0;
        const int pr = bi, pc = bl;
        const double piv = __shfl_sync(FULL, sel5(q, pr), pc);
        const double qn = fast_div(sel5(q, pr), piv);  // normalised pivot row, this lane's column
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const double f = __shfl_sync(FULL, q[k], pc);
            q[k] = k == pr ? qn : q[k] - f * qn;
        }
        used_rows |= 1u << pr;
        free_cols &= ~(1u << pc);
        prow[r] = pr;
        pcol[r] = pc;
        {
            int pos = r;
#pragma unroll
            for (int t = 0; t < 9; t++)
                if ((int)((perm >> (4 * t)) & 15ULL) == pc) pos = t;
            const unsigned long long at_r = (perm >> (4 * r)) & 15ULL;
            perm &= ~((15ULL << (4 * r)) | (15ULL << (4 * pos)));
            perm |= ((unsigned long long)pc << (4 * r)) | (at_r << (4 * pos));
        }
    }
    double bk[4];
    {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int f = (int)((perm >> (4 * (5 + k))) & 15ULL);  // k-th free column
            double v = lane == f ? 1.0 : 0.0;
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const double t = __shfl_sync(FULL, sel5(q, prow[r]), f);
                if (lane == pcol[r]) v = -t;
            }
            bk[k] = lane < 9 ? v : 0.0;
        }
    }
#pragma unroll
    for (int rep = 0; rep < 2; rep++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int j = 0; j < k; j++) {
                const double d = half_sum(bk[k] * bk[j]);
                bk[k] = bk[k] - d * bk[j];
            }
            bk[k] = bk[k] * fast_rsqrt(half_sum(bk[k] * bk[k]));
        }
    // ================= B: the ten cubic constraints ====================================================================
    double* Ep = R;              // [9][4]
    double* L = R + 36;          // [9][10]
    double* Af = R + 126;        // [10][20]
    double* Dp = R + 326;        // [3][20] the three terms of det(E)
    if (lane < 9) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            Ep[lane * 4 + k] = bk[k];
            B4[k * 9 + lane] = bk[k];
        }
    }
    __syncwarp();
    if (lane < 9) {  // L[i][j] = sum_k E[i][k] E[j][k]
        const int i = lane / 3, j = lane - 3 * i;
        double acc[10];
#pragma unroll
        for (int t = 0; t < 10; t++) acc[t] = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) mul11_acc(Ep + (3 * i + k) * 4, Ep + (3 * j + k) * 4, acc, 1.0);
#pragma unroll
        for (int t = 0; t < 10; t++) L[lane * 10 + t] = acc[t];
    } else if (lane < 12) {  // the three terms of det(E)
        const int c = lane - 9;
        const int a0 = c == 0 ? 1 : c == 1 ? 2 : 0, a1 = c == 0 ? 5 : c == 1 ? 3 : 4;
        const int b0 = c == 0 ? 2 : c == 1 ? 0 : 1, b1 = c == 0 ? 4 : c == 1 ? 5 : 3;
        double t2[10], out[20];
#pragma unroll
        for (int t = 0; t < 10; t++) t2[t] = 0.0;
#pragma unroll
        for (int t = 0; t < 20; t++) out[t] = 0.0;
        mul11_acc(Ep + a0 * 4, Ep + a1 * 4, t2, 1.0);
        mul11_acc(Ep + b0 * 4, Ep + b1 * 4, t2, -1.0);
        mul21_acc(t2, Ep + (6 + c) * 4, out);
#pragma unroll
        for (int t = 0; t < 20; t++) Dp[c * 20 + t] = out[t];
    }
    __syncwarp();
    if (lane < 10) {  // L <- E E' - tr(E E') / 2 I
        const double tr = (L[0 * 10 + lane] + L[4 * 10 + lane]) + L[8 * 10 + lane];
        L[0 * 10 + lane] = L[0 * 10 + lane] - 0.5 * tr;
        L[4 * 10 + lane] = L[4 * 10 + lane] - 0.5 * tr;
        L[8 * 10 + lane] = L[8 * 10 + lane] - 0.5 * tr;
    }
    __syncwarp();
    if (lane < 9) {  // row 1 + 3i + j = sum_k L[i][k] E[k][j]
        const int i = lane / 3, j = lane - 3 * i;
        double out[20];
#pragma unroll
        for (int t = 0; t < 20; t++) out[t] = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) mul21_acc(L + (3 * i + k) * 10, Ep + (3 * k + j) * 4, out);
#pragma unroll
        for (int t = 0; t < 20; t++) Af[(1 + lane) * 20 + t] = out[t];
    } else if (lane < 29) {  // row 0 = the sum of the three det terms
        const int t = lane - 9;
        Af[t] = (Dp[t] + Dp[20 + t]) + Dp[40 + t];
    }
    __syncwarp();
    // ================= C: Gauss-Jordan on the first ten columns, lane j < 20 holds column j ===========================
    double col[10];
    {
        const int j = lane < 20 ? lane : 0;
#pragma unroll
        for (int r = 0; r < 10; r++) col[r] = Af[r * 20 + j];
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 10; c++) {
        int piv = c;
        double best = fabs(col[c]);
#pragma unroll
        for (int r = c + 1; r < 10; r++)
            if (fabs(col[r]) > best) { best = fabs(col[r]); piv = r; }
        piv = __shfl_sync(FULL, piv, c);
        best = __shfl_sync(FULL, best, c);
        if (best == 0.0) return 0;
        {
            const double t = col[c];
#pragma unroll
            for (int r = c + 1; r < 10; r++)
                if (r == piv) { col[c] = col[r]; col[r] = t; }
        }
        const double d = __shfl_sync(FULL, col[c], c);
        col[c] = fast_div(col[c], d);
#pragma unroll
        for (int r = 0; r < 10; r++) {
            if (r == c) continue;
            const double f = __shfl_sync(FULL, col[r], c);
            col[r] = col[r] - f * col[c];
        }
    }
    // ================= D: B(z) and det B(z) ============================================================================
    double* Ar = R;              // [6][10] rows 4..9, columns 10..19 of the reduced matrix
    double* bxyz = R + 60;       // [3][13] bx(4) by(4) b1(5)
    double* mco = R + 100;       // [3][7]  cofactors
    double* det = R + 124;       // [11]
    double* pa = R + 136;        // [2][11] the two polynomials, ascending
    double* cpb = R + 160;       // [2][2][12] critical points, double buffered
    double* roots = R + 208;     // [10]
    double* T = R + 220;         // [2][66] derivative tables
    if (lane >= 10 && lane < 20) {
#pragma unroll
        for (int r = 4; r < 10; r++) Ar[(r - 4) * 10 + (lane - 10)] = col[r];
    }
    __syncwarp();
    if (lane < 3) {
        const double* a = Ar + (2 * lane) * 10;
        const double* b = Ar + (2 * lane + 1) * 10;
        double* o = bxyz + lane * 13;
        o[0] = 0.0 - b[0]; o[1] = a[0] - b[1]; o[2] = a[1] - b[2]; o[3] = a[2] - 0.0;
        o[4] = 0.0 - b[3]; o[5] = a[3] - b[4]; o[6] = a[4] - b[5]; o[7] = a[5] - 0.0;
        o[8] = 0.0 - b[6]; o[9] = a[6] - b[7]; o[10] = a[7] - b[8]; o[11] = a[8] - b[9]; o[12] = a[9] - 0.0;
    }
    __syncwarp();
    if (lane < 3) {  // cofactor of b1[lane]: rows (1,2), (0,2), (0,1)
        const int r0 = lane == 0 ? 1 : 0, r1 = lane == 2 ? 1 : 2;
        const double* bx0 = bxyz + r0 * 13;
        const double* by0 = bx0 + 4;
        const double* bx1 = bxyz + r1 * 13;
        const double* by1 = bx1 + 4;
        double m[7];
#pragma unroll
        for (int t = 0; t < 7; t++) m[t] = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) m[i + j] += bx0[i] * by1[j];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) m[i + j] += -1.0 * (by0[i] * bx1[j]);
#pragma unroll
        for (int t = 0; t < 7; t++) mco[lane * 7 + t] = m[t];
    }
    __syncwarp();
    if (lane < 11) {
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double sign = c == 1 ? -1.0 : 1.0;
            const double* b1 = bxyz + c * 13 + 8;
#pragma unroll
            for (int i = 0; i < 5; i++) {
                const int j = lane - i;
                if (j >= 0 && j < 7) v += sign * (b1[i] * mco[c * 7 + j]);
            }
        }
        det[lane] = v;
    }
    __syncwarp();
    // ================= E: real roots (half-warp 0: p on [-1, 1]; half-warp 1: reversed p on (-1, 1)) =================
    int lead = 0;
    while (lead < 11 && det[lead] == 0.0) lead++;
    const int n = 10 - lead;  // degree
    if (n < 1) return 0;
    const int half = lane >> 4, hl = lane & 15;
    double* a = pa + half * 11;
    if (hl <= n) {
        const double v = fast_div(det[lead + hl], det[lead]);  // coefficient of x^(n - hl)
        a[half == 0 ? n - hl : hl] = v;
    }
    __syncwarp();
    // (when p(0) = 0 the reversed polynomial simply keeps zero leading coefficients: its first derivative levels are
    // identically zero or constant and produce no brackets)
    double* Th = T + half * 66;
    for (int k = 0; k < n; k++) {
        if (hl <= n - k) Th[poly_toff(k, n) + hl] = a[hl + k] * binom(hl + k, k);
    }
    __syncwarp();
    int m = 0;  // roots of the previous level (uniform within the half)
    const bool closed = half == 0;
    for (int k = n - 1; k >= 0; k--) {  // both halves walk the levels together: warp-uniform control flow
        const double* f = Th + poly_toff(k, n);
        const double* cp = cpb + (half * 2 + (k & 1)) * 12;
        double* nx = cpb + (half * 2 + ((k & 1) ^ 1)) * 12;
        switch (n - k) {
            case 1: m = root_level<1>(f, cp, nx, m, k == 0, closed); break;
            case 2: m = root_level<2>(f, cp, nx, m, k == 0, closed); break;
            case 3: m = root_level<3>(f, cp, nx, m, k == 0, closed); break;
            case 4: m = root_level<4>(f, cp, nx, m, k == 0, closed); break;
            case 5: m = root_level<5>(f, cp, nx, m, k == 0, closed); break;
            case 6: m = root_level<6>(f, cp, nx, m, k == 0, closed); break;
            case 7: m = root_level<7>(f, cp, nx, m, k == 0, closed); break;
            case 8: m = root_level<8>(f, cp, nx, m, k == 0, closed); break;
            case 9: m = root_level<9>(f, cp, nx, m, k == 0, closed); break;
            default: m = root_level<10>(f, cp, nx, m, k == 0, closed); break;
        }
    }
    // final roots of each half are in its buffer 1
    const double* fin = cpb + (half * 2 + 1) * 12;
    const int m0 = __shfl_sync(FULL, m, 0), m1 = __shfl_sync(FULL, m, 16);
    {
        const bool ok = half == 1 && hl < m1 && fin[hl < m1 ? hl : 0] != 0.0;
        const unsigned bb = __ballot_sync(FULL, ok) >> 16;
        const int pos = m0 + __popc(bb & ((1u << hl) - 1u));
        if (half == 0 && hl < m0) roots[hl] = fin[hl];
        if (ok && pos < 10) roots[pos] = 1.0 / fin[hl];
        m = __popc(bb);
    }
    __syncwarp();
    const int nz = min(m0 + m, n);
    // ================= F: back-substitution, lane = root =================================================================
    bool good = false;
    double e[9];
    if (lane < nz) {
        const double z = roots[lane];
        const double z2 = z * z, z3 = z2 * z, z4 = z3 * z;
        double Bz[3][3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const double* o = bxyz + i * 13;
            Bz[i][0] = ((o[0] * z3 + o[1] * z2) + o[2] * z) + o[3];
            Bz[i][1] = ((o[4] * z3 + o[5] * z2) + o[6] * z) + o[7];
            Bz[i][2] = (((o[8] * z4 + o[9] * z3) + o[10] * z2) + o[11] * z) + o[12];
        }
        double v0 = 0, v1 = 0, v2 = 0, vn = -1.0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int ra = c == 2 ? 1 : 0, rb = c == 0 ? 1 : 2;
            const double cx = Bz[ra][1] * Bz[rb][2] - Bz[ra][2] * Bz[rb][1], cy = Bz[ra][2] * Bz[rb][0] - Bz[ra][0] * Bz[rb][2],
                         cz = Bz[ra][0] * Bz[rb][1] - Bz[ra][1] * Bz[rb][0];
            const double nn = (cx * cx + cy * cy) + cz * cz;
            if (nn > vn) { vn = nn; v0 = cx; v1 = cy; v2 = cz; }
        }
        if (!(fabs(v2) < 1e-10 * sqrt(vn))) {
            const double x = fast_div(v0, v2), y = fast_div(v1, v2);
            double nn = 0.0;
#pragma unroll
            for (int t = 0; t < 9; t++) {
                e[t] = ((x * B4[t] + y * B4[9 + t]) + z * B4[18 + t]) + B4[27 + t];
                nn += e[t] * e[t];
            }
            nn = fast_rsqrt(nn);
#pragma unroll
            for (int t = 0; t < 9; t++) e[t] = e[t] * nn;
            good = true;
        }
    }
    const unsigned gb = __ballot_sync(FULL, good);
    if (good) {
        const int pos = __popc(gb & ((1u << lane) - 1u));
#pragma unroll
        for (int t = 0; t < 9; t++) models[pos * 9 + t] = e[t];
    }
    __syncwarp();
    return __popc(gb);
}

}  // namespace slamcu
