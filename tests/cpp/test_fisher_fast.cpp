// CPU test of csrc/bv_fisher_fast.h (the product's O(log range) Fisher) against the oracle's restatement of
// kt_fisher_exact (oracle/bv_oracle.c, itself pinned to the compiled reference), on random and adversarial tables.
#include <dlfcn.h>
#include <math.h>

#include <cstdio>
#include <random>
#include <string>

#include "../../basevar_b200/csrc/bv_fisher_fast.h"

typedef double (*fisher_fn)(int, int, int, int);

int main(int argc, char** argv) {
    const std::string root = argc > 1 ? argv[1] : ".";
    void* ho = dlopen((root + "/oracle/libbvoracle.so").c_str(), RTLD_NOW);
    if (!ho) { printf("FAIL cannot load oracle: %s\n", dlerror()); return 1; }
    fisher_fn oracle = (fisher_fn)dlsym(ho, "bvo_fisher_two_sided");
    long n_fast = 0;
    auto lf = [](int k) { return lgamma((double)k + 1.0); };
    auto ex = [](double x) { return exp(x); };
    auto fast = [&](int a, int b, int c, int d) {
        const int n1_ = a + b, n_1 = a + c, n = a + b + c + d;
        int hi = n_1 < n1_ ? n_1 : n1_, lo = n1_ + n_1 - n;
        if (lo < 0) lo = 0;
        if (lo == hi) return 1.0;
        const double q = bv::ff_pmf(lf, ex, a, n1_, n_1, n);
        if (q == 0.0) return 0.0;
        if (!bv::fisher_fast_applicable(lo, hi, q)) return oracle(a, b, c, d);   // the product runs the faithful loop there
        ++n_fast;
        return bv::fisher_two_sided_fast(lf, ex, a, n1_, n_1, n, lo, hi, q);
    };
    std::mt19937_64 rng(7);
    int fails = 0;
    long n_tab = 0;
    double worst = 0;
    int wa = 0, wb = 0, wc = 0, wd = 0;
    auto check = [&](int a, int b, int c, int d) {
        const double w = oracle(a, b, c, d), g = fast(a, b, c, d);
        ++n_tab;
        // p-values feed FS = -10 log10(p): compare in that scale too (absolute noise of the reference near p == 1)
        const double rel = fabs(g - w) / (w > 1e-300 ? w : 1e-300);
        const double fs_w = w > 0 ? -10 * log10(w) : 10000, fs_g = g > 0 ? -10 * log10(g) : 10000;
        const bool ok = (rel < 1e-9) || fabs(fs_g - fs_w) < 1e-9 + 1e-9 * fabs(fs_w);
        if (rel > worst) { worst = rel; wa = a; wb = b; wc = c; wd = d; }
        if (!ok && fails++ < 10) printf("FAIL (%d,%d,%d,%d): fast %.17g oracle %.17g\n", a, b, c, d, g, w);
    };
    // exhaustive small tables
    for (int a = 0; a < 14; ++a) for (int b = 0; b < 14; ++b) for (int c = 0; c < 14; ++c) for (int d = 0; d < 14; ++d) check(a, b, c, d);
    // symmetric tables: exact ties of the pmf
    for (int k = 1; k < 400; k += 3) { check(k, k, k, k); check(k, 2 * k, 2 * k, k); check(k, k + 1, k + 1, k); check(3 * k, k, k, 3 * k); }
    // random tables over the depths of the BASELINE shapes (dense 2,000-sample rows, 10^4..10^5 covered samples)
    for (int it = 0; it < 60000; ++it) {
        const int scale = (int[]){30, 300, 2000, 10000, 100000}[it % 5];
        const int a = (int)(rng() % scale), b = (int)(rng() % scale);
        const int alt = (int)(rng() % (1 + scale / (1 + (int)(rng() % 50))));
        const int c = alt ? (int)(rng() % (alt + 1)) : 0, d = alt - c;
        check(a, b, c, d);
        check(c, d, a, b);
        if (it % 7 == 0) check(a, c, b, d);
    }
    // strongly biased tables (tiny p, underflow to 0)
    check(600, 0, 0, 400); check(5000, 4900, 100, 20); check(50000, 49000, 300, 200); check(100000, 0, 0, 100000);
    check(1000, 3, 2, 900); check(500, 480, 20, 3);
    printf("%ld tables (%ld through the fast path), worst relative difference %.3g at (%d,%d,%d,%d)\n", n_tab, n_fast, worst, wa, wb, wc, wd);
    printf(fails ? "FAILED %d\n" : "ALL OK\n", fails);
    return fails ? 1 : 0;
}
