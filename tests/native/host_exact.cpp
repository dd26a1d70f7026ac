// Host build of slam_cin0051_b200/csrc/exact.cuh so the bit-exact emulations (glibc float libm,
// libstdc++ sort permutations) can be checked on the CPU against the real libraries without a GPU.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../slam_cin0051_b200/csrc/exact.cuh"

extern "C" {
// returns number of mismatches (bitwise) between the port and libm over [start, start+count) float bit patterns
long long hx_check_atanf_bits(uint32_t start, long long count) {
    long long bad = 0;
    for (long long i = 0; i < count; i++) {
        float x = slamcu::u2f((uint32_t)(start + i));
        if (x != x) continue;
        float a = slamcu::glibc_atanf(x), b = atanf(x);
        if (slamcu::f2u(a) != slamcu::f2u(b)) bad++;
    }
    return bad;
}
long long hx_check_atan2f(const float* y, const float* x, long long n) {
    long long bad = 0;
    for (long long i = 0; i < n; i++) {
        float a = slamcu::glibc_atan2f(y[i], x[i]), b = atan2f(y[i], x[i]);
        if (slamcu::f2u(a) != slamcu::f2u(b)) bad++;
    }
    return bad;
}
long long hx_check_sincosf_bits(uint32_t start, long long count) {
    long long bad = 0;
    for (long long i = 0; i < count; i++) {
        float x = slamcu::u2f((uint32_t)(start + i));
        if (x != x) continue;
        float s = slamcu::glibc_sinf(x), c = slamcu::glibc_cosf(x);
        if (slamcu::f2u(s) != slamcu::f2u(sinf(x))) bad++;
        if (slamcu::f2u(c) != slamcu::f2u(cosf(x))) bad++;
    }
    return bad;
}
struct KeyDesc { bool operator()(uint32_t a, uint32_t b) const { return (a >> 20) > (b >> 20); } };
// keys: (score << 20) | raster index.  In-place emulated std::sort (descending by score).
void hx_sort_keys_desc(uint32_t* keys, int n) { slamcu::std_sort(keys, n, KeyDesc()); }
void hx_ref_sort_keys_desc(uint32_t* keys, int n) { std::sort(keys, keys + n, KeyDesc()); }
struct KeyAsc { bool operator()(uint32_t a, uint32_t b) const { return (a >> 16) < (b >> 16); } };
void hx_partial_sort_asc(uint32_t* keys, int mid, int n) { slamcu::std_partial_sort(keys, mid, n, KeyAsc()); }
void hx_ref_partial_sort_asc(uint32_t* keys, int mid, int n) { std::partial_sort(keys, keys + mid, keys + n, KeyAsc()); }
void hx_sort_asc(uint32_t* keys, int n) { slamcu::std_sort(keys, n, KeyAsc()); }
void hx_ref_sort_asc(uint32_t* keys, int n) { std::sort(keys, keys + n, KeyAsc()); }
}

// ---- scalar model of the *parallel* formulation used by the CUDA sort kernel -------------------
// Hoare partition expressed through the ordered lists of left-stops / right-stops (see sortnms.cu);
// level-synchronous segment queue; final per-segment stable insertion.  Must equal std::sort.
namespace {
int model_partition(uint32_t* a, int first, int last, std::vector<int>& L, std::vector<int>& R) {
    KeyDesc less;
    const int mid = first + (last - first) / 2;
    const int ia = first + 1, ib = mid, ic = last - 1;
    int pick;
    if (less(a[ia], a[ib])) { if (less(a[ib], a[ic])) pick = ib; else if (less(a[ia], a[ic])) pick = ic; else pick = ia; }
    else if (less(a[ia], a[ic])) pick = ia; else if (less(a[ib], a[ic])) pick = ic; else pick = ib;
    std::swap(a[first], a[pick]);
    const uint32_t p = a[first];
    L.clear(); R.clear();
    for (int i = first + 1; i < last; i++) {
        if (!less(a[i], p)) L.push_back(i);
        if (!less(p, a[i])) R.push_back(i);   // ascending; R_k = R[nR-k]
    }
    const int nL = (int)L.size(), nR = (int)R.size();
    int K = 0;
    while (K < nL && K < nR && L[K] < R[nR - 1 - K]) K++;
    for (int k = 0; k < K; k++) std::swap(a[L[k]], a[R[nR - 1 - k]]);
    const int rk = (K > 0) ? R[nR - K] : last;
    return (K < nL && L[K] < rk) ? L[K] : rk;
}
}
extern "C" void hx_model_sort_keys_desc(uint32_t* a, int n) {
    if (n <= 0) return;
    KeyDesc less;
    std::vector<std::pair<int,int>> cur, nxt;
    std::vector<char> bnd(n + 1, 0);
    bnd[0] = 1;
    int depth = 2 * slamcu::std_lg(n);
    if (n > 16) cur.push_back({0, n});
    std::vector<int> L, R;
    while (!cur.empty()) {
        nxt.clear();
        for (auto [f, l] : cur) {
            if (depth == 0) { slamcu::std_partial_sort(a + f, l - f, l - f, less); continue; }
            int cut = model_partition(a, f, l, L, R);
            bnd[cut] = 1;
            if (cut - f > 16) nxt.push_back({f, cut});
            if (l - cut > 16) nxt.push_back({cut, l});
        }
        depth--;
        cur.swap(nxt);
    }
    int s = 0;
    for (int e = 1; e <= n; e++) {
        if (e == n || bnd[e]) {
            for (int i = s + 1; i < e; i++) {
                uint32_t v = a[i]; int j = i;
                while (j > s && less(v, a[j - 1])) { a[j] = a[j - 1]; j--; }
                a[j] = v;
            }
            s = e;
        }
    }
}

// ---- host build of the five-point solver (slam_cin0051_b200/csrc/fivept.cuh) ------------------------------------------
#include "../../slam_cin0051_b200/csrc/fivept.cuh"
extern "C" int hx_real_roots(const double* coeffs_high_first, int n_coeffs, double* out) {
    return slamcu::real_roots(coeffs_high_first, n_coeffs, out);
}
// x1, x2: [n_samples][5][2]; models: [n_samples][10][9]; counts: [n_samples]
extern "C" void hx_five_point(const double* x1, const double* x2, int n_samples, double* models, int* counts) {
    for (int s = 0; s < n_samples; s++) {
        double a[5][2], b[5][2];
        for (int i = 0; i < 5; i++) {
            a[i][0] = x1[(s * 5 + i) * 2]; a[i][1] = x1[(s * 5 + i) * 2 + 1];
            b[i][0] = x2[(s * 5 + i) * 2]; b[i][1] = x2[(s * 5 + i) * 2 + 1];
        }
        counts[s] = slamcu::five_point(a, b, models + (size_t)s * 90);
    }
}
