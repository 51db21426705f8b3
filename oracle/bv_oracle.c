/*
 * bv_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's per-site algorithm.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, load or call this file.  The product path (basevar_b200/csrc, include/basevar_b200.h)
 * never links or calls it.
 *
 * It follows the reference read by read (NOT the histogram form the CUDA kernels use), in plain
 * C doubles, in the reference's operation order, so that it is bit-identical with the compiled
 * reference (oracle/_ref/libbvref.so, built from /root/reference by oracle/build_ref.sh).  That
 * identity is checked by tests/test_oracle_vs_ref.py whenever oracle/_ref exists and is pinned by
 * the golden vectors in tests/golden/ (generated from the compiled reference).
 *
 * Reference lines restated (all paths relative to /root/reference):
 *   bvo_site()            src/basetype.cpp:22-72   BaseType::BaseType  (likelihood rows, depths)
 *   em_run()              src/algorithm.h:148-255  e_step, m_step, EM  (incl. the int abs() of :245)
 *   subsets               src/external/combinations.h:19-84 (lexicographic k-subsets)
 *   lrt                   src/basetype.cpp:93-199  _set_allele_initial_freq, _f, lrt
 *   gammaq_half()         htslib/kfunc.c:39-52,103-143  kf_lgamma, _kf_gammap, _kf_gammaq, kf_gammaq
 *   fisher_two_sided()    htslib/kfunc.c:197-313   lbinom, hypergeo, hypergeo_acc, kt_fisher_exact
 *   strand tables / FS    src/basetype.cpp:244-295 strand_bias
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/basevar_b200.h"
#include "bv_oracle.h"

static const double kMLN10TO10 = -0.23025850929940458; /* src/basetype.h:20 */

/* x86-64 cvttsd2si: out-of-range and NaN give the "integer indefinite" value. */
static int cvt_trunc_x86(double x) {
    if (!(x > -2147483649.0 && x < 2147483648.0)) return INT_MIN;
    return (int)x;
}
/* int abs(int) as compiled (neg wraps for INT_MIN). */
static int iabs_wrap(int v) { return v == INT_MIN ? INT_MIN : (v < 0 ? -v : v); }

/* ------------------------------------------------------------------------------------------------
 * EM over d reads with 4-wide likelihood rows (src/algorithm.h:148-255).
 * lik: d x 4, f: in/out 4, lml: out d (log marginal likelihood under the second-to-last f).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    double* post;  /* d x 4 */
    double* marg;  /* d */
} em_ws;

static void e_step(const double* f, const double* lik, size_t d, em_ws* w) {
    for (size_t i = 0; i < d; ++i) {
        double l[4], m = 0.0;
        for (int j = 0; j < 4; ++j) {
            l[j] = lik[4 * i + j] * f[j];
            m += l[j];
        }
        w->marg[i] = m;
        for (int j = 0; j < 4; ++j) w->post[4 * i + j] = l[j] / m;
    }
}

static void m_step(const em_ws* w, size_t d, double* f) {
    for (int j = 0; j < 4; ++j) {
        double s = 0.0;
        for (size_t i = 0; i < d; ++i) s += w->post[4 * i + j];
        f[j] = s / (double)d;
    }
}

static int em_run(const double* lik, size_t d, double* f, double* lml, em_ws* w, int max_iter, float eps,
                  int abs_mode) {
    int loops = 0;
    e_step(f, lik, d, w);
    for (size_t i = 0; i < d; ++i) lml[i] = log(w->marg[i]);
    m_step(w, d, f);
    int it = max_iter;
    while (it--) {
        ++loops;
        e_step(f, lik, d, w);
        m_step(w, d, f);
        double delta = 0.0;
        for (size_t i = 0; i < d; ++i) {
            double llh = log(w->marg[i]);
            double diff = llh - lml[i];
            if (abs_mode == BV_EM_ABS_INT_TRUNC)
                delta += (double)iabs_wrap(cvt_trunc_x86(diff)); /* algorithm.h:245 with int abs(int) */
            else
                delta += fabs(diff);
            lml[i] = llh;
        }
        if (delta < (double)eps) break;
    }
    m_step(w, d, f); /* algorithm.h:253: recomputes the same f */
    return loops;
}

/* ------------------------------------------------------------------------------------------------
 * htslib/kfunc.c restated
 * ------------------------------------------------------------------------------------------------ */
static double lgamma_as245(double z) { /* kfunc.c:39-52 */
    double x = 0;
    x += 0.1659470187408462e-06 / (z + 7);
    x += 0.9934937113930748e-05 / (z + 6);
    x -= 0.1385710331296526 / (z + 5);
    x += 12.50734324009056 / (z + 4);
    x -= 176.6150291498386 / (z + 3);
    x += 771.3234287757674 / (z + 2);
    x -= 1259.139216722289 / (z + 1);
    x += 676.5203681218835 / z;
    x += 0.9999999999995183;
    return log(x) - 5.58106146679532777 - z + (z - 0.5) * log(z + 6.5);
}

static double gammap_series(double s, double z) { /* kfunc.c:106-115 */
    double sum = 1.0, x = 1.0;
    for (int k = 1; k < 100; ++k) {
        x *= z / (s + k);
        sum += x;
        if (x / sum < 1e-14) break;
    }
    return exp(s * log(z) - z - lgamma_as245(s + 1.) + log(sum));
}

static double gammaq_cfrac(double s, double z) { /* kfunc.c:117-136, modified Lentz */
    const double tiny = 1e-290;
    double f = 1. + z - s, C = f, D = 0.;
    for (int j = 1; j < 100; ++j) {
        double a = j * (s - j), b = (j << 1) + 1 + z - s, d;
        D = b + a * D;
        if (D < tiny) D = tiny;
        C = b + a / C;
        if (C < tiny) C = tiny;
        D = 1. / D;
        d = C * D;
        f *= d;
        if (fabs(d - 1.) < 1e-14) break;
    }
    return exp(s * log(z) - z - lgamma_as245(s) - log(f));
}

double bvo_gammaq(double s, double z) { /* kfunc.c:140-143 */
    return (z <= 1. || z < s) ? 1. - gammap_series(s, z) : gammaq_cfrac(s, z);
}

double bvo_chi2_test(double chi, double dof) { /* src/algorithm.h:44-46 */
    return bvo_gammaq(dof / 2.0, chi / 2.0);
}

static double lbinom(int n, int k) { /* kfunc.c:197-201 */
    if (k == 0 || n == k) return 0;
    return lgamma(n + 1) - lgamma(k + 1) - lgamma(n - k + 1);
}

static double hypergeo(int n11, int n1_, int n_1, int n) { /* kfunc.c:209-212 */
    return exp(lbinom(n1_, n11) + lbinom(n - n1_, n_1 - n11) - lbinom(n, n_1));
}

typedef struct {
    int n11, n1_, n_1, n;
    double p;
} hg_state;

/* kfunc.c:220-243.  `full` = the call passed the margins; otherwise only n11 moves. */
static double hg_step(hg_state* st, int n11, int n1_, int n_1, int n, int full) {
    if (full) {
        st->n11 = n11; st->n1_ = n1_; st->n_1 = n_1; st->n = n;
    } else {
        int n22 = n11 + st->n - st->n1_ - st->n_1;
        if ((n11 % 11) && n22) {
            if (n11 == st->n11 + 1) {
                st->p *= (double)(st->n1_ - st->n11) / n11 * (st->n_1 - st->n11) / n22;
                st->n11 = n11;
                return st->p;
            }
            if (n11 == st->n11 - 1) {
                st->p *= (double)st->n11 / (st->n1_ - n11) * (st->n11 + st->n - st->n1_ - st->n_1) /
                         (st->n_1 - n11);
                st->n11 = n11;
                return st->p;
            }
        }
        st->n11 = n11;
    }
    st->p = hypergeo(st->n11, st->n1_, st->n_1, st->n);
    return st->p;
}

/* kfunc.c:245-313, two-sided p only (what fisher_exact_test returns, src/algorithm.h:62-74). */
double bvo_fisher_two_sided(int n11, int n12, int n21, int n22) {
    int n1_ = n11 + n12, n_1 = n11 + n21, n = n11 + n12 + n21 + n22;
    int hi = (n_1 < n1_) ? n_1 : n1_;
    int lo = n1_ + n_1 - n;
    if (lo < 0) lo = 0;
    if (lo == hi) return 1.;
    hg_state st;
    /* kfunc.c:257: hypergeo_acc(n11, n1_, n_1, n): an all-zero margin triple would take the
       "only n11 changed" branch on uninitialised state; lo != hi rules that out (n > 0). */
    double q = hg_step(&st, n11, n1_, n_1, n, 1);
    if (q == 0.0) return 0.0; /* kfunc.c:259-290: *two = 0 on both sides of the mode */
    int i, j;
    double p, left, right;
    p = hg_step(&st, lo, 0, 0, 0, 0);
    for (left = 0., i = lo + 1; p < 0.99999999 * q && i <= hi; ++i) {
        left += p;
        p = hg_step(&st, i, 0, 0, 0, 0);
    }
    --i;
    if (p < 1.00000001 * q) left += p;
    else --i;
    p = hg_step(&st, hi, 0, 0, 0, 0);
    for (right = 0., j = hi - 1; p < 0.99999999 * q && j >= 0; --j) {
        right += p;
        p = hg_step(&st, j, 0, 0, 0, 0);
    }
    ++j;
    if (p < 1.00000001 * q) right += p;
    else ++j;
    double two = left + right;
    if (two > 1.) two = 1.;
    (void)i; (void)j;
    return two;
}

/* src/basetype.cpp:277-283 */
double bvo_fs_from_table(int ref_fwd, int ref_rev, int alt_fwd, int alt_rev) {
    double fs = -10 * log10(bvo_fisher_two_sided(ref_fwd, ref_rev, alt_fwd, alt_rev));
    if (isinf(fs)) fs = 10000;
    else if (fs == 0) fs = 0.0;
    return fs;
}

/* src/basetype.cpp:286: int32 products, as in the reference */
double bvo_sor_from_table(int ref_fwd, int ref_rev, int alt_fwd, int alt_rev) {
    return (ref_rev * alt_fwd > 0) ? (double)(ref_fwd * alt_rev) / (double)(ref_rev * alt_fwd) : 10000;
}

/* ------------------------------------------------------------------------------------------------
 * k-subsets of an ordered set in lexicographic index order (combinations.h:19-84).
 * ------------------------------------------------------------------------------------------------ */
static int next_subset(int* idx, int k, int n) {
    int i = k - 1;
    while (i >= 0 && idx[i] == n - k + i) --i;
    if (i < 0) return 0;
    ++idx[i];
    for (int j = i + 1; j < k; ++j) idx[j] = idx[j - 1] + 1;
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * BaseType::lrt(specific_bases), src/basetype.cpp:130-199: active set over the ORDERED candidate list `cand`
 * (A,C,G,T for lrt(); [upper REF, ALT...] for the population-group calls, basetype_caller.cpp:750-753),
 * full-model EM, backward elimination.  Returns the number of active bases left (0: nothing to report).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
    int active[4];
    double f_act[4];
    double chi;
    unsigned em_calls;
} lrt_result;

static int lrt_core(const double* lik, size_t d, const double* depth, int total, const int* cand, int n_cand,
                    const bv_params* prm, unsigned* flags_io, lrt_result* res) {
    const double min_af = (double)prm->min_af; /* float -> double, basetype_caller.cpp:122,506 */
    unsigned flags = *flags_io;
    int active[4], n_active = 0;
    if (total > 0) {
        for (int k = 0; k < n_cand; ++k)
            if (depth[cand[k]] / total >= min_af) active[n_active++] = cand[k];
    }
    if (total == 0 || n_active == 0) return 0;

    em_ws w;
    w.post = (double*)malloc(sizeof(double) * 4 * d);
    w.marg = (double*)malloc(sizeof(double) * d);
    double* lml = (double*)malloc(sizeof(double) * d);
    unsigned em_calls = 0;

    double f_act[4] = {0, 0, 0, 0}, lr_alt, chi = 0.0;
    {   /* _f(active, |active|): one subset, src/basetype.cpp:105-128 */
        double f[4] = {0, 0, 0, 0};
        for (int k = 0; k < n_active; ++k) f[active[k]] = depth[active[k]] / (double)total;
        int loops = em_run(lik, d, f, lml, &w, prm->em_max_iter, prm->em_eps, prm->em_abs_mode);
        if (loops >= prm->em_max_iter) flags |= BV_FLAG_EM_MAXITER;
        ++em_calls;
        double s = 0.0;
        for (size_t i = 0; i < d; ++i) s += lml[i];
        lr_alt = s;
        memcpy(f_act, f, sizeof(f));
    }
    for (int n = n_active - 1; n > 0; --n) {
        int idx[4];
        for (int k = 0; k < n; ++k) idx[k] = k;
        double best_chi = 0, best_lr = 0, best_f[4] = {0, 0, 0, 0};
        int best_set[4] = {0, 0, 0, 0}, first = 1;
        do {
            double f[4] = {0, 0, 0, 0}, fsum = 0.0;
            for (int k = 0; k < n; ++k) f[active[idx[k]]] = depth[active[idx[k]]] / (double)total;
            for (int j = 0; j < 4; ++j) fsum += f[j];
            if (fsum == 0) flags |= BV_FLAG_ZERO_SUBSET; /* reference throws, basetype.cpp:113 */
            int loops = em_run(lik, d, f, lml, &w, prm->em_max_iter, prm->em_eps, prm->em_abs_mode);
            if (loops >= prm->em_max_iter) flags |= BV_FLAG_EM_MAXITER;
            if (em_calls < 255) ++em_calls;
            double s = 0.0;
            for (size_t i = 0; i < d; ++i) s += lml[i];
            double c = 2 * (lr_alt - s);
            /* Diagnostic only (the decision below is the reference's strict first minimum): two candidates whose statistics
             * agree to 1e-10 of the log-likelihoods they are differences of are a TIE -- alleles with identical read multisets;
             * which of them the reference keeps hangs on the rounding noise of its read-order sums.  The parity tests accept a
             * different pick of the CUDA path only at sites THIS flag (or NEAR_LRT below) marks. */
            if (!first && fabs(c - best_chi) <= 1e-10 * (fabs(lr_alt) + fabs(s))) flags |= BV_FLAG_LRT_TIE;
            if (first || c < best_chi) { /* std::min_element: first minimum, '<' only */
                first = 0;
                best_chi = c; best_lr = s;
                memcpy(best_f, f, sizeof(f));
                for (int k = 0; k < n; ++k) best_set[k] = active[idx[k]];
            }
        } while (next_subset(idx, n, n_active));
        lr_alt = best_lr;
        chi = best_chi;
        {
            double thr = (double)prm->lrt_threshold;
            if (fabs(chi - thr) < 1e-9 * thr) flags |= BV_FLAG_NEAR_LRT;
        }
        if (chi < prm->lrt_threshold) {
            n_active = n;
            for (int k = 0; k < n; ++k) active[k] = best_set[k];
            memcpy(f_act, best_f, sizeof(best_f));
        } else {
            break;
        }
    }

    free(lml); free(w.post); free(w.marg);
    for (int k = 0; k < n_active; ++k) res->active[k] = active[k];
    memcpy(res->f_act, f_act, sizeof(f_act));
    res->chi = chi;
    res->em_calls = em_calls;
    *flags_io = flags;
    return n_active;
}

/* ------------------------------------------------------------------------------------------------
 * One site.
 * ------------------------------------------------------------------------------------------------ */
int bvo_site(const uint8_t* base, const uint8_t* qual, const uint8_t* strand, uint32_t n_samples,
             uint8_t ref_char, const bv_params* prm, bv_site_out* out) {
    memset(out, 0, sizeof(*out));

    double* lik = (double*)malloc(sizeof(double) * 4 * (size_t)(n_samples ? n_samples : 1));
    size_t d = 0;
    double depth[4] = {0, 0, 0, 0};
    unsigned flags = 0;
    /* BaseType::BaseType, src/basetype.cpp:45-71 */
    for (uint32_t i = 0; i < n_samples; ++i) {
        uint8_t b = base[i];
        if (b >= BV_BASE_N) continue; /* 'N' and indels are skipped, basetype.cpp:51 */
        uint8_t q = qual[i];
        if (q > BV_QUAL_MAX) flags |= BV_FLAG_BAD_QUAL;
        double eps = exp((double)q * kMLN10TO10);
        if (b < 4) {
            depth[b] += 1;
            out->depth[b]++;
            if (strand[i] == BV_STRAND_FWD) out->fwd[b]++;
            else if (strand[i] == BV_STRAND_REV) out->rev[b]++;
        } else {
            out->depth_other++;
        }
        if (strand[i] != BV_STRAND_FWD && strand[i] != BV_STRAND_REV) flags |= BV_FLAG_BAD_STRAND;
        for (int j = 0; j < 4; ++j) lik[4 * d + j] = (b == j) ? 1.0 - eps : eps / 3;
        ++d;
    }
    const int total = (int)d;

    /* strand_bias for the CVG row: ref vs all non-ref ACGT (basetype_caller.cpp:1236-1245) */
    int up_ref = ref_char;
    if (up_ref >= 'a' && up_ref <= 'z') up_ref -= 32; /* toupper, basetype.cpp:171 */
    int ref_code = up_ref == 'A' ? 0 : up_ref == 'C' ? 1 : up_ref == 'G' ? 2 : up_ref == 'T' ? 3 : -1;
    {
        int rf = 0, rr = 0, af_ = 0, ar = 0;
        for (int b = 0; b < 4; ++b) {
            if (b == ref_code) { rf += out->fwd[b]; rr += out->rev[b]; }
            else { af_ += out->fwd[b]; ar += out->rev[b]; }
        }
        out->fs_cvg = bvo_fs_from_table(rf, rr, af_, ar);
    }

    /* BaseType::lrt, src/basetype.cpp:130-199 */
    static const int kACGT[4] = {0, 1, 2, 3};
    lrt_result res;
    int n_active = lrt_core(lik, d, depth, total, kACGT, 4, prm, &flags, &res);
    if (n_active == 0) {
        out->flags = (uint8_t)flags;
        free(lik);
        return 0;
    }
    const int* active = res.active;
    const double* f_act = res.f_act;
    const double chi = res.chi;
    const unsigned em_calls = res.em_calls;

    int n_alt = 0;
    for (int k = 0; k < n_active; ++k) {
        if (active[k] != ref_code) {
            out->alt[n_alt] = (uint8_t)active[k];
            out->af[n_alt] = f_act[active[k]];
            ++n_alt;
        }
    }
    out->n_alt = (uint8_t)n_alt;
    out->n_active = (uint8_t)n_active;
    out->chi2 = chi;
    out->em_calls = (uint8_t)em_calls;
    if (n_alt) {
        double r = depth[active[0]] / (double)total;
        if (n_active == 1 && total > 10 && r > 0.5) {
            out->qual = 5000.0;
            flags |= BV_FLAG_MONO_QUAL;
        } else {
            double p = bvo_chi2_test(chi, 1);
            if (isnan(p)) p = 1.0;
            double qv = (p != 0.0) ? -10 * log10(p) : 10000.0;
            if (qv == -0.0) qv = 0.0;
            out->qual = qv;
        }
        /* strand_bias for the VCF row: ref vs called ALT (basetype_caller.cpp:1164) */
        int rf = 0, rr = 0, af_ = 0, ar = 0;
        if (ref_code >= 0) { rf = out->fwd[ref_code]; rr = out->rev[ref_code]; }
        for (int k = 0; k < n_alt; ++k) { af_ += out->fwd[out->alt[k]]; ar += out->rev[out->alt[k]]; }
        out->fs_vcf = bvo_fs_from_table(rf, rr, af_, ar);
    }
    out->flags = (uint8_t)flags;
    free(lik);
    return 0;
}

int bvo_tile(const uint8_t* base, const uint8_t* qual, const uint8_t* strand, const uint8_t* ref_base,
             uint64_t pitch, uint32_t n_sites, uint32_t n_samples, const bv_params* prm, bv_site_out* out) {
    for (uint32_t s = 0; s < n_sites; ++s)
        bvo_site(base + (size_t)s * pitch, qual + (size_t)s * pitch, strand + (size_t)s * pitch, n_samples,
                 ref_base[s], prm, out + s);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Population group: __gb / __get_group_batchinfo (src/basetype_caller.cpp:767-797) = BaseType over the
 * group's samples, then lrt([upper REF, ALT...]) (src/basetype_caller.cpp:747-760).
 * order[]: the candidate base codes in list order (a REF that is not A/C/G/T is left out: it has no
 * depth entry and can never be active).
 * ------------------------------------------------------------------------------------------------ */
int bvo_group_site(const uint8_t* base, const uint8_t* qual, uint32_t n_samples, const uint8_t* sample_group,
                   uint32_t g, uint8_t ref_char, const uint8_t* order, int n_order, const bv_params* prm,
                   bv_group_out* out) {
    memset(out, 0, sizeof(*out));
    double* lik = (double*)malloc(sizeof(double) * 4 * (size_t)(n_samples ? n_samples : 1));
    size_t d = 0;
    double depth[4] = {0, 0, 0, 0};
    unsigned flags = 0;
    for (uint32_t i = 0; i < n_samples; ++i) {
        if (sample_group[i] != g) continue;
        uint8_t b = base[i];
        if (b >= BV_BASE_N) continue;
        uint8_t q = qual[i];
        if (q > BV_QUAL_MAX) flags |= BV_FLAG_BAD_QUAL;
        double eps = exp((double)q * kMLN10TO10);
        if (b < 4) depth[b] += 1;
        for (int j = 0; j < 4; ++j) lik[4 * d + j] = (b == j) ? 1.0 - eps : eps / 3;
        ++d;
    }
    int up_ref = ref_char;
    if (up_ref >= 'a' && up_ref <= 'z') up_ref -= 32;
    int ref_code = up_ref == 'A' ? 0 : up_ref == 'C' ? 1 : up_ref == 'G' ? 2 : up_ref == 'T' ? 3 : -1;
    int cand[4];
    for (int k = 0; k < n_order && k < 4; ++k) cand[k] = order[k];
    lrt_result res;
    int n_active = lrt_core(lik, d, depth, (int)d, cand, n_order < 4 ? n_order : 4, prm, &flags, &res);
    for (int k = 0; k < n_active; ++k) {
        if (res.active[k] != ref_code) {
            out->alt[out->n_alt] = (uint8_t)res.active[k];
            out->af[out->n_alt] = res.f_act[res.active[k]];
            out->n_alt++;
        }
    }
    out->flags = (uint8_t)flags;
    free(lik);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Rank sums: kf_erfc (htslib/kfunc.c:58-84), norm_dist / wilcoxon_ranksum_test (src/algorithm.h:48-50,
 * 76-136), ref_vs_alt_ranksumtest (src/basetype.cpp:201-242), (int) at src/basetype_caller.cpp:1151-1157.
 * ------------------------------------------------------------------------------------------------ */
double bvo_erfc(double x) {
    const double p0 = 220.2068679123761, p1 = 221.2135961699311, p2 = 112.0792914978709, p3 = 33.912866078383,
                 p4 = 6.37396220353165, p5 = .7003830644436881, p6 = .03526249659989109;
    const double q0 = 440.4137358247522, q1 = 793.8265125199484, q2 = 637.3336333788311, q3 = 296.5642487796737,
                 q4 = 86.78073220294608, q5 = 16.06417757920695, q6 = 1.755667163182642, q7 = .08838834764831844;
    double expntl, z, p;
    z = fabs(x) * M_SQRT2;
    if (z > 37.) return x > 0. ? 0. : 2.;
    expntl = exp(z * z * -.5);
    if (z < 10. / M_SQRT2)
        p = expntl * ((((((p6 * z + p5) * z + p4) * z + p3) * z + p2) * z + p1) * z + p0) /
            (((((((q7 * z + q6) * z + q5) * z + q4) * z + q3) * z + q2) * z + q1) * z + q0);
    else
        p = expntl / 2.506628274631001 / (z + 1. / (z + 2. / (z + 3. / (z + 4. / (z + .65)))));
    return x > 0. ? 2. * p : 2. * (1. - p);
}

typedef struct { double v; size_t i; } rk_item;
static int rk_desc(const void* a, const void* b) {
    const rk_item *x = (const rk_item*)a, *y = (const rk_item*)b;
    return x->v > y->v ? -1 : (x->v < y->v ? 1 : 0);
}

/* wilcoxon_ranksum_test: descending sort, average ranks for ties, normal approximation without tie correction */
double bvo_wilcoxon(const double* s1, size_t n1, const double* s2, size_t n2) {
    const size_t n = n1 + n2;
    rk_item* it = (rk_item*)malloc(sizeof(rk_item) * (n ? n : 1));
    double* rank = (double*)malloc(sizeof(double) * (n ? n : 1));
    for (size_t i = 0; i < n1; ++i) { it[i].v = s1[i]; it[i].i = i; }
    for (size_t i = 0; i < n2; ++i) { it[n1 + i].v = s2[i]; it[n1 + i].i = n1 + i; }
    qsort(it, n, sizeof(rk_item), rk_desc);
    for (size_t i = 0; i < n; ++i) rank[i] = (double)(i + 1);
    double ranksum = 0.0, same_n = 1;
    size_t i;
    for (i = 0; i < n; ++i) {
        if (i > 0 && it[i].v != it[i - 1].v) {
            if (same_n > 1) {
                double avg = ranksum / same_n;
                for (size_t j = i - (size_t)same_n; j < i; ++j) rank[j] = avg;
            }
            same_n = 1;
            ranksum = 0;
        } else if (i > 0) {
            same_n++;
        }
        ranksum += (double)(i + 1);
    }
    if (same_n > 1) {
        double avg = ranksum / same_n;
        for (size_t j = i - (size_t)same_n; j < i; ++j) rank[j] = avg;
    }
    double smp1 = 0.0;
    for (size_t k = 0; k < n; ++k)
        if (it[k].i < n1) smp1 += rank[k];
    free(it); free(rank);
    double e = (double)(n1 * (n1 + n2 + 1)) / 2.0;
    double z = (smp1 - e) / sqrt((double)(n1 * n2 * (n1 + n2 + 1)) / 12.0);
    return 2 * (bvo_erfc(fabs(z) / sqrt(2.0)) / 2.0);
}

/* one called site: the three INFO values from the packed planes; alt_mask bit b = base code b is a called ALT */
int bvo_ranksums(const uint8_t* base, const uint8_t* qual, const uint8_t* mapq, const uint16_t* rpr, uint32_t n_samples,
                 uint8_t ref_char, uint32_t alt_mask, int32_t out3[3]) {
    int up_ref = ref_char;
    if (up_ref >= 'a' && up_ref <= 'z') up_ref -= 32;
    int ref_code = up_ref == 'A' ? 0 : up_ref == 'C' ? 1 : up_ref == 'G' ? 2 : up_ref == 'T' ? 3 : -1;
    double* r = (double*)malloc(sizeof(double) * (n_samples ? n_samples : 1));
    double* a = (double*)malloc(sizeof(double) * (n_samples ? n_samples : 1));
    for (int t = 0; t < 3; ++t) {
        size_t n1 = 0, n2 = 0;
        for (uint32_t i = 0; i < n_samples; ++i) {
            uint8_t b = base[i];
            if (b > 3) continue; /* N, indels skipped; other characters match neither REF nor an ALT */
            double v = t == 0 ? (double)mapq[i] : t == 1 ? (double)rpr[i] : (double)(qual[i] + 33);
            if ((int)b == ref_code) r[n1++] = v;
            else if (alt_mask >> b & 1u) a[n2++] = v;
        }
        double ph;
        if (n1 > 0 && n2 > 0) {
            double p = bvo_wilcoxon(r, n1, a, n2);
            ph = -10 * log10(p);
            if (isinf(ph)) ph = 10000;
        } else {
            ph = 10000;
        }
        out3[t] = cvt_trunc_x86(ph);
    }
    free(r); free(a);
    return 0;
}
