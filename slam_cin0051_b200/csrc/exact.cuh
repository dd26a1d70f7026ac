// exact.cuh -- bit-exact re-statements of the host-library behaviour the reference's results depend on.
//
// The reference (CPU, g++ 13 / libstdc++ 13 / glibc 2.39, baseline x86-64 => no FMA) produces results
// that depend on three library behaviours the GPU does not have natively:
//   * glibc atan2f / sinf / cosf last-bit behaviour   (feature_detector.cpp:229, :247-249)
//   * the permutation std::sort produces on ties      (feature_detector.cpp:160-161, feature_matcher.cpp:201)
//   * the permutation std::partial_sort produces      (feature_matcher.cpp:197-199)
// Everything here is __host__ __device__ so tests/host_check.cpp can verify it on the CPU against the
// real libm / libstdc++ without a GPU.  Compile device code with -fmad=false (no FMA contraction).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SLAM_HD __host__ __device__ __forceinline__
#define SLAM_HDN __host__ __device__ inline
#else
#define SLAM_HD inline
#define SLAM_HDN inline
#endif

namespace slamcu {

SLAM_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } v; v.f = f; return v.u;
#endif
}
SLAM_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } v; v.u = u; return v.f;
#endif
}

// ---------------------------------------------------------------------------------------------
// glibc 2.39 atanf / atan2f (fdlibm single-precision code path; every op is an IEEE float op).
// Restated from the published algorithm; verified against libm on the host.
// ---------------------------------------------------------------------------------------------
SLAM_HDN float glibc_atanf(float x) {
    const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
                          9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f,
                          4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
    const uint32_t hx = f2u(x);
    const uint32_t ix = hx & 0x7fffffffu;
    int id;
    if (ix >= 0x4c000000u) {  // |x| >= 2^25
        if (ix > 0x7f800000u) return x + x;
        const float r = atanhi[3] + atanlo[3];
        return (hx >> 31) ? -r : r;
    }
    if (ix < 0x3ee00000u) {  // |x| < 0.4375
        if (ix < 0x31000000u) return x;
        id = -1;
    } else {
        x = u2f(ix);
        if (ix < 0x3f980000u) {
            if (ix < 0x3f300000u) { id = 0; x = (2.0f * x - 1.0f) / (2.0f + x); }
            else                  { id = 1; x = (x - 1.0f) / (x + 1.0f); }
        } else {
            if (ix < 0x401c0000u) { id = 2; x = (x - 1.5f) / (1.0f + 1.5f * x); }
            else                  { id = 3; x = -1.0f / x; }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    const float s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const float r = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
    return (hx >> 31) ? -r : r;
}

SLAM_HDN float glibc_atan2f(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    const int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return glibc_atanf(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) {
        switch (m) {
            case 0:
            case 1: return y;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000 || iy == 0x7f800000) {
        // moments are finite; kept for completeness
        if (ix == 0x7f800000) {
            if (iy == 0x7f800000) {
                switch (m) {
                    case 0: return 7.8539818525e-01f + tiny;
                    case 1: return -7.8539818525e-01f - tiny;
                    case 2: return 3.0f * 7.8539818525e-01f + tiny;
                    default: return -3.0f * 7.8539818525e-01f - tiny;
                }
            }
            switch (m) {
                case 0: return 0.0f;
                case 1: return -0.0f;
                case 2: return pi + tiny;
                default: return -pi - tiny;
            }
        }
        return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    }
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else {
        float q = y / x;
        z = glibc_atanf(u2f(f2u(q) & 0x7fffffffu));
    }
    switch (m) {
        case 0: return z;
        case 1: return u2f(f2u(z) ^ 0x80000000u);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

// ---------------------------------------------------------------------------------------------
// glibc 2.39 sinf / cosf (double-precision polynomial scheme; |y| < 120 path only -- the reference
// feeds |angle| <= pi).  SURVEY.md Appendix F.2.
// ---------------------------------------------------------------------------------------------
struct SinCosTab { double sign[4]; double hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3; };

SLAM_HD float sincos_poly(double x, double x2, double cs, int n) {
    // cs = +1 for table 0, -1 for table 1 (cosine coefficients negated)
    const double c0 = 1.0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
                 c4 = 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = x * x2;
        const double t = s2 + x2 * s3;
        const double x7 = x3 * x2;
        const double s = x + x3 * s1;
        return (float)(s + x7 * t);
    }
    const double x4 = x2 * x2;
    const double u = cs * c3 + x2 * (cs * c4);
    const double v = cs * c0 + x2 * (cs * c1);
    const double x6 = x4 * x2;
    const double c = v + x4 * (cs * c2);
    return (float)(c + x6 * u);
}

SLAM_HD uint32_t abstop12(float f) { return (f2u(f) >> 20) & 0x7ffu; }

// which: 0 = sinf, 1 = cosf
SLAM_HDN float glibc_sincosf_one(float y, int which) {
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    double x = (double)y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {  // |y| < pi/4
        const double x2 = x * x;
        if (abstop12(y) < abstop12(0x1p-12f)) return which ? 1.0f : y;
        return sincos_poly(x, x2, 1.0, which);
    }
    // |y| < 120 reduction
    const double r = x * hpi_inv;
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = x - (double)n * hpi;
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const double cs = (n & 2) ? -1.0 : 1.0;
    return sincos_poly(x * sgn, x * x, cs, which ? (n ^ 1) : n);
}
SLAM_HD float glibc_sinf(float y) { return glibc_sincosf_one(y, 0); }
SLAM_HD float glibc_cosf(float y) { return glibc_sincosf_one(y, 1); }

// ---------------------------------------------------------------------------------------------
// libstdc++ 13 std::sort / std::partial_sort, restated so the *permutation* on ties is identical.
// (bits/stl_algo.h: __move_median_to_first, __unguarded_partition, __introsort_loop (_S_threshold 16,
//  depth 2*lg n, heapsort fallback), __final_insertion_sort; bits/stl_heap.h: __adjust_heap, __push_heap,
//  __make_heap, __pop_heap, __sort_heap, and __heap_select.)   `less(a, b)` plays the role of comp.
// ---------------------------------------------------------------------------------------------
template <class T, class Less>
SLAM_HDN void std_adjust_heap(T* a, int hole, int len, T value, Less less) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(a[child], a[child - 1])) child--;
        a[hole] = a[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[hole] = a[child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && less(a[parent], value)) {
        a[hole] = a[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[hole] = value;
}

template <class T, class Less>
SLAM_HDN void std_make_heap(T* a, int len, Less less) {
    if (len < 2) return;
    int parent = (len - 2) / 2;
    while (true) {
        T v = a[parent];
        std_adjust_heap(a, parent, len, v, less);
        if (parent == 0) return;
        parent--;
    }
}

// __pop_heap(first, last, result): heap is a[0..len), result is element index `res`
template <class T, class Less>
SLAM_HD void std_pop_heap(T* a, int len, int res, Less less) {
    T v = a[res];
    a[res] = a[0];
    std_adjust_heap(a, 0, len, v, less);
}

// std::partial_sort(a, a+mid, a+n)
template <class T, class Less>
SLAM_HDN void std_partial_sort(T* a, int mid, int n, Less less) {
    std_make_heap(a, mid, less);
    for (int i = mid; i < n; i++)
        if (less(a[i], a[0])) std_pop_heap(a, mid, i, less);
    int last = mid;
    while (last > 1) {
        --last;
        std_pop_heap(a, last, last, less);
    }
}

template <class T, class Less>
SLAM_HD void std_unguarded_linear_insert(T* a, int last, Less less) {
    T v = a[last];
    int next = last - 1;
    while (less(v, a[next])) {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = v;
}

template <class T, class Less>
SLAM_HDN void std_insertion_sort(T* a, int first, int last, Less less) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (less(a[i], a[first])) {
            T v = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];
            a[first] = v;
        } else {
            std_unguarded_linear_insert(a, i, less);
        }
    }
}

// one partition step on [first, last): returns the cut
template <class T, class Less>
SLAM_HDN int std_partition_pivot(T* a, int first, int last, Less less) {
    const int mid = first + (last - first) / 2;
    const int ia = first + 1, ib = mid, ic = last - 1;
    int pick;
    if (less(a[ia], a[ib])) {
        if (less(a[ib], a[ic])) pick = ib;
        else if (less(a[ia], a[ic])) pick = ic;
        else pick = ia;
    } else if (less(a[ia], a[ic])) pick = ia;
    else if (less(a[ib], a[ic])) pick = ic;
    else pick = ib;
    { T t = a[first]; a[first] = a[pick]; a[pick] = t; }
    const T pivot = a[first];
    int lo = first + 1, hi = last;
    while (true) {
        while (less(a[lo], pivot)) ++lo;
        --hi;
        while (less(pivot, a[hi])) --hi;
        if (!(lo < hi)) return lo;
        T t = a[lo]; a[lo] = a[hi]; a[hi] = t;
        ++lo;
    }
}

SLAM_HD int std_lg(int n) {  // floor(log2 n), n > 0
    int l = 0;
    while (n > 1) { n >>= 1; l++; }
    return l;
}

// std::sort(a, a+n) -- serial emulation with an explicit stack replacing the recursion on the right part.
template <class T, class Less>
SLAM_HDN void std_sort(T* a, int n, Less less) {
    if (n <= 0) return;
    int stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * std_lg(n); sp = 1;
    while (sp > 0) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > 16) {
            if (depth == 0) {
                std_partial_sort(a + first, last - first, last - first, less);
                break;
            }
            --depth;
            const int cut = std_partition_pivot(a, first, last, less);
            // recursion on [cut, last) with the decremented depth; segments are independent, so deferring
            // it on a stack gives the same array as the depth-first order libstdc++ uses.
            stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; sp++;
            last = cut;
        }
    }
    if (n > 16) {
        std_insertion_sort(a, 0, 16, less);
        for (int i = 16; i < n; ++i) std_unguarded_linear_insert(a, i, less);
    } else {
        std_insertion_sort(a, 0, n, less);
    }
}

}  // namespace slamcu
