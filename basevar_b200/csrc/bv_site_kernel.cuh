// bv_site_kernel.cuh -- the per-site statistical core of `basevar basetype` as one fused sm_100a kernel.
//
// One warp owns one genomic site at a time (sites are the embarrassingly parallel axis, samples the
// reduction axis).  For its site the warp
//   1. streams the three u8 planes of the site row (base, qual, strand) with 128-bit loads and counts
//      every read into a warp-private shared-memory histogram hist[base 0..4][phred 0..95] plus
//      per-lane packed strand counters                      (BaseType::BaseType, src/basetype.cpp:45-71;
//                                                            strand_bias counting, src/basetype.cpp:252-274)
//   2. compacts the non-empty (base, phred) bins in place, in (base, phred) order
//   3. runs EM + LRT backward elimination on the bins: all reads of one bin are exchangeable in
//      e_step/m_step (src/algorithm.h:148-198), so a bin of c reads contributes c * (per-read term);
//      lanes own bins, allele sums are warp-shuffle reductions   (EM, src/algorithm.h:210-255;
//                                                                 _f / lrt, src/basetype.cpp:105-199)
//   4. QUAL (chi2 survival via kf_gammaq) and the two Fisher strand-bias tests  (src/basetype.cpp:180-194,
//                                                                                 :244-295)
//   5. writes the fixed 128-byte bv_site_out record.
//
// The only FP64 work that scales with the number of samples is gone: the sample axis is byte loads and
// integer shared-memory atomics; FP64 work is O(bins) per site.  Per-read likelihood values (1-eps,
// eps/3) come from a host-computed table (glibc exp), so they are bit-identical with the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/basevar_b200.h"
#include "bv_math.cuh"

namespace bv {

constexpr int kQStride = 96;               // phred slots per base row of the histogram (0..93 used)
constexpr int kHistWords = 5 * kQStride;   // A,C,G,T,other
constexpr int kLutOneMinusEps = 0;         // lut[0][q] = 1 - eps(q)
constexpr int kLutEpsThird = 1;            // lut[1][q] = eps(q) / 3
constexpr int kLutLogMatch = 2;            // lut[2][q] = log(1 - eps(q))   (glibc)
constexpr int kLutLogMis = 3;              // lut[3][q] = log(eps(q) / 3)   (glibc)

struct SiteKernelArgs {
    const uint8_t* base;
    const uint8_t* qual;
    const uint8_t* strand;
    const uint8_t* ref_base;
    bv_site_out* out;
    const double* lut;       // [4][kQStride]
    const double* logfact;   // [max_samples + 2], lgamma(k+1) from glibc
    uint64_t pitch;
    uint32_t n_sites;
    uint32_t n_samples;
    double min_af;           // (double)(float)min_af
    double em_eps;           // (double)(float)0.001
    double lrt_threshold;
    int em_max_iter;
    int abs_mode;
};

// Per-warp shared-memory working set.
struct __align__(16) WarpScratch {
    uint32_t hist[kHistWords];  // dense histogram while streaming; packed compact bins afterwards
    double lml[kHistWords];     // log marginal likelihood of each compact bin (EM state)
    bv_site_out rec;            // record staging for one coalesced 128-byte store
};

__device__ __forceinline__ uint32_t pack_bin(uint32_t code, uint32_t count) { return (code << 22) | count; }
__device__ __forceinline__ uint32_t bin_code(uint32_t p) { return p >> 22; }
__device__ __forceinline__ uint32_t bin_count(uint32_t p) { return p & 0x3fffffu; }

// streaming loads: read once, do not pollute L1
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Per-lane counters of one site row.
struct LaneCounts {
    unsigned long long fwd, rev, nos;  // 4 x 16-bit fields (A,C,G,T) per strand class
    uint32_t other, qmin, qmax, flags;
};

// Count the (up to 4) cells of one 32-bit word of each plane.  MASKED: only the first `valid` cells exist
// (tail of the row; cells in [n_samples, pitch) are padding).
template <bool MASKED>
__device__ __forceinline__ void count_word(uint32_t wb, uint32_t wq, uint32_t ws, int valid, uint32_t* hist,
                                           LaneCounts& lc) {
    // bit 7 of each byte of `nc` is set iff the base code is >= 5 (N / indel / junk): not counted
    uint32_t nc = ((((wb | 0x80808080u) - 0x05050505u) | wb) & 0x80808080u);
    uint32_t m = ~nc & 0x80808080u;
    if (MASKED) m &= (valid <= 0) ? 0u : (valid >= 4 ? 0xffffffffu : (0xffffffffu >> (8 * (4 - valid))));
    while (m) {
        int sh = __ffs(m) - 8;  // bit index of the cell's LSB
        uint32_t b = (wb >> sh) & 0xffu;
        uint32_t q = (wq >> sh) & 0xffu;
        uint32_t s = (ws >> sh) & 0xffu;
        m &= m - 1;
        if (q > BV_QUAL_MAX) { lc.flags |= BV_FLAG_BAD_QUAL; q = BV_QUAL_MAX; }
        atomicAdd(&hist[b * kQStride + q], 1u);
        lc.qmin = min(lc.qmin, q);
        lc.qmax = max(lc.qmax, q);
        if (b < 4) {
            unsigned long long inc = 1ull << (16 * b);
            if (s == BV_STRAND_FWD) lc.fwd += inc;
            else if (s == BV_STRAND_REV) lc.rev += inc;
            else { lc.nos += inc; lc.flags |= BV_FLAG_BAD_STRAND; }
        } else {
            lc.other += 1;
            if (s > BV_STRAND_REV) lc.flags |= BV_FLAG_BAD_STRAND;
        }
    }
}

template <bool MASKED>
__device__ __forceinline__ void count_vec(const uint4& vb, const uint4& vq, const uint4& vs, int valid,
                                          uint32_t* hist, LaneCounts& lc) {
    count_word<MASKED>(vb.x, vq.x, vs.x, valid, hist, lc);
    count_word<MASKED>(vb.y, vq.y, vs.y, valid - 4, hist, lc);
    count_word<MASKED>(vb.z, vq.z, vs.z, valid - 8, hist, lc);
    count_word<MASKED>(vb.w, vq.w, vs.w, valid - 12, hist, lc);
}

// ---- EM on compact bins (src/algorithm.h:210-255) ----------------------------------------------------------
// subset: bit j set => allele j in the candidate combination.  f_io: initial frequencies in (NOT renormalised,
// src/basetype.cpp:93-103), estimated frequencies out.  Returns sum of log marginal likelihoods under the
// second-to-last frequency vector, exactly what _f() sums (src/basetype.cpp:119-120).
__device__ __noinline__ double em_bins(const uint32_t* bins, int nb, double* lml, const double* s_lut, int subset,
                                       double total, double* f_io, const SiteKernelArgs& a, uint32_t& flags) {
    const int lane = threadIdx.x & 31;
    double f0 = f_io[0], f1 = f_io[1], f2 = f_io[2], f3 = f_io[3];
    int it = a.em_max_iter;
    bool first = true;
    for (;;) {
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0, delta = 0;
        bool big = false;
        for (int i = lane; i < nb; i += 32) {
            uint32_t p = bins[i];
            uint32_t code = bin_code(p);
            uint32_t b = code / kQStride, q = code - b * kQStride;
            double cd = (double)bin_count(p);
            double ome = s_lut[kLutOneMinusEps * kQStride + q], e3 = s_lut[kLutEpsThird * kQStride + q];
            // e_step (algorithm.h:160-172): lik*freq summed in A,C,G,T order; alleles outside the subset have
            // freq 0 and add an exact +0.0, so they are skipped
            double l0 = 0, l1 = 0, l2 = 0, l3 = 0, m = 0;
            if (subset & 1) { l0 = (b == 0 ? ome : e3) * f0; m += l0; }
            if (subset & 2) { l1 = (b == 1 ? ome : e3) * f1; m += l1; }
            if (subset & 4) { l2 = (b == 2 ? ome : e3) * f2; m += l2; }
            if (subset & 8) { l3 = (b == 3 ? ome : e3) * f3; m += l3; }
            double llh = log(m);
            if (!first) {
                double diff = llh - lml[i];
                if (a.abs_mode == BV_EM_ABS_INT_TRUNC) {
                    // (double)abs((int)diff): non-zero iff |diff| >= 1; NaN/inf convert to INT_MIN whose
                    // "abs" stays negative and ends the loop (results are NaN by then)
                    if (fabs(diff) >= 1.0 && fabs(diff) < 2147483648.0) big = true;
                } else {
                    delta += cd * fabs(diff);
                }
            }
            lml[i] = llh;
            // m_step (algorithm.h:184-198): column sums of the posteriors; c equal reads add c * post
            if (subset & 1) s0 += cd * (l0 / m);
            if (subset & 2) s1 += cd * (l1 / m);
            if (subset & 4) s2 += cd * (l2 / m);
            if (subset & 8) s3 += cd * (l3 / m);
        }
        if (subset & 1) f0 = warp_sum(s0) / total;
        if (subset & 2) f1 = warp_sum(s1) / total;
        if (subset & 4) f2 = warp_sum(s2) / total;
        if (subset & 8) f3 = warp_sum(s3) / total;
        if (first) { first = false; continue; }
        bool more;
        if (a.abs_mode == BV_EM_ABS_INT_TRUNC) more = __any_sync(0xffffffffu, big);
        else more = !(warp_sum(delta) < a.em_eps);
        --it;
        if (it == 0) flags |= BV_FLAG_EM_MAXITER;
        if (!more || it == 0) break;
    }
    double ll = 0;
    for (int i = lane; i < nb; i += 32) ll += (double)bin_count(bins[i]) * lml[i];
    ll = warp_sum(ll);
    f_io[0] = f0; f_io[1] = f1; f_io[2] = f2; f_io[3] = f3;
    return ll;
}

// Log-likelihood of the single-allele model {b} (an EM whose answer is closed form):
// after the first m_step f_b == 1.0 exactly (every posterior is x/x), so every later marginal is L_b itself and
// the reported log marginal is log(1-eps) or log(eps/3) -- both tabulated on the host with glibc.  A bin of base b
// with phred 0 has L_b == 0: the reference then divides 0/0 and everything becomes NaN.
__device__ __forceinline__ double single_allele_ll(const uint32_t* bins, int nb, const double* s_lut, int b_allele,
                                                   bool& is_nan) {
    const int lane = threadIdx.x & 31;
    double ll = 0;
    bool bad = false;
    for (int i = lane; i < nb; i += 32) {
        uint32_t p = bins[i];
        uint32_t code = bin_code(p);
        uint32_t b = code / kQStride, q = code - b * kQStride;
        double cd = (double)bin_count(p);
        bool match = ((int)b == b_allele);
        if (match && q == 0) bad = true;
        ll += cd * s_lut[(match ? kLutLogMatch : kLutLogMis) * kQStride + q];
    }
    is_nan = __any_sync(0xffffffffu, bad);
    return warp_sum(ll);
}

// The LRT loop only ever asks for the (n-1)-subsets of the current n active bases.  In the lexicographic
// position order of src/external/combinations.h:19-84 the i-th of them drops position n-1-i.
__device__ __forceinline__ int subset_posmask(int n_active, int i) {
    return ((1 << n_active) - 1) ^ (1 << (n_active - 1 - i));
}

// ---- the warp-per-site core: everything after the row has been histogrammed -------------------------------------
__device__ __forceinline__ void site_finish(WarpScratch& ws, const double* s_lut, const SiteKernelArgs& a,
                                            uint32_t site, LaneCounts& lc) {
    const int lane = threadIdx.x & 31;
    // ---- reduce lane counters ----
    uint32_t fwd[4], rev[4], dep[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t f = (uint32_t)(lc.fwd >> (16 * b)) & 0xffffu;
        uint32_t r = (uint32_t)(lc.rev >> (16 * b)) & 0xffffu;
        uint32_t x = (uint32_t)(lc.nos >> (16 * b)) & 0xffffu;
        fwd[b] = __reduce_add_sync(0xffffffffu, f);
        rev[b] = __reduce_add_sync(0xffffffffu, r);
        dep[b] = fwd[b] + rev[b] + __reduce_add_sync(0xffffffffu, x);
    }
    const uint32_t other = __reduce_add_sync(0xffffffffu, lc.other);
    uint32_t flags = __reduce_or_sync(0xffffffffu, lc.flags);
    const uint32_t qmin = __reduce_min_sync(0xffffffffu, lc.qmin);
    const uint32_t qmax = __reduce_max_sync(0xffffffffu, lc.qmax);
    const uint32_t total = dep[0] + dep[1] + dep[2] + dep[3] + other;

    // ---- compact the non-empty bins in place, (base, phred) order ----
    __syncwarp();
    int nb = 0;
    if (total > 0) {
        for (int b = 0; b < 5; ++b) {
            for (uint32_t q0 = qmin; q0 <= qmax; q0 += 32) {
                uint32_t q = q0 + lane;
                uint32_t idx = b * kQStride + q;
                uint32_t v = 0;
                if (q <= qmax) { v = ws.hist[idx]; }
                __syncwarp();
                if (q <= qmax && v) ws.hist[idx] = 0;
                uint32_t bal = __ballot_sync(0xffffffffu, v != 0);
                __syncwarp();
                if (v) ws.hist[nb + __popc(bal & ((1u << lane) - 1u))] = pack_bin(idx, v);
                nb += __popc(bal);
                __syncwarp();
            }
        }
    }
    uint32_t* bins = ws.hist;

    // ---- reference base ----
    int ref_char = a.ref_base[site];
    if (ref_char >= 'a' && ref_char <= 'z') ref_char -= 32;  // toupper (src/basetype.cpp:171)
    const int ref_code = ref_char == 'A' ? 0 : ref_char == 'C' ? 1 : ref_char == 'G' ? 2 : ref_char == 'T' ? 3 : -1;

    // ---- lrt (src/basetype.cpp:130-199) ----
    int act[4];
    int n_act = 0;
    const double dtot = (double)total;
    if (total > 0) {
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if ((double)dep[b] / dtot >= a.min_af) act[n_act++] = b;   // exact-boundary compare (:137)
    }
    double f_act[4] = {0, 0, 0, 0};
    double chi = 0.0;
    uint32_t em_calls = 0;
    if (n_act == 1) {
        // One active allele: the reference still runs one EM; its answer is closed form (see single_allele_ll):
        // AF == 1.0 exactly, or NaN when a phred-0 read of that base exists.
        bool bad;
        (void)single_allele_ll(bins, nb, s_lut, act[0], bad);
        f_act[act[0]] = bad ? __longlong_as_double(0x7ff8000000000000ll) : 1.0;
        em_calls = 1;
    } else if (n_act > 1) {
        int mask = 0;
        for (int k = 0; k < n_act; ++k) { mask |= 1 << act[k]; f_act[act[k]] = (double)dep[act[k]] / dtot; }
        double lr_alt = em_bins(bins, nb, ws.lml, s_lut, mask, dtot, f_act, a, flags);
        em_calls = 1;
        for (int n = n_act - 1; n > 0; --n) {
            const int ns = n_act;   // C(n_act, n_act-1)
            double best_chi = 0, best_lr = 0, best_f[4] = {0, 0, 0, 0};
            int best_pm = 0;
            for (int i = 0; i < ns; ++i) {
                const int pm = subset_posmask(n_act, i);
                double f[4] = {0, 0, 0, 0};
                int sm = 0, single = -1;
                for (int k = 0; k < n_act; ++k)
                    if (pm & (1 << k)) { sm |= 1 << act[k]; f[act[k]] = (double)dep[act[k]] / dtot; single = act[k]; }
                double lr;
                if (n == 1) {
                    bool bad;
                    lr = single_allele_ll(bins, nb, s_lut, single, bad);
                    f[single] = 1.0;
                    if (bad) { lr = __longlong_as_double(0x7ff8000000000000ll); f[single] = lr; }
                } else {
                    lr = em_bins(bins, nb, ws.lml, s_lut, sm, dtot, f, a, flags);
                }
                if (em_calls < 255) ++em_calls;
                double c = 2 * (lr_alt - lr);
                // std::min_element keeps the FIRST minimum (algorithm.h:24-27).  Alleles with identical read
                // multisets give bit-identical chi in the reference (its per-read sums are symmetric under
                // relabelling); here the bin order is not symmetric, so values that agree to rounding noise
                // are treated as the tie they are and the earlier subset stays.
                const double tie_tol = 1e-11 * (fabs(lr_alt) + fabs(lr));
                if (i > 0 && fabs(c - best_chi) <= tie_tol) flags |= BV_FLAG_LRT_TIE;
                if (i == 0 || c < best_chi - tie_tol) {
                    best_chi = c; best_lr = lr; best_pm = pm;
#pragma unroll
                    for (int j = 0; j < 4; ++j) best_f[j] = f[j];
                }
            }
            lr_alt = best_lr;
            chi = best_chi;
            if (fabs(chi - a.lrt_threshold) < 1e-9 * a.lrt_threshold) flags |= BV_FLAG_NEAR_LRT;
            if (chi < a.lrt_threshold) {
                int k2 = 0;
                for (int k = 0; k < n_act; ++k)
                    if (best_pm & (1 << k)) act[k2++] = act[k];
                n_act = n;
#pragma unroll
                for (int j = 0; j < 4; ++j) f_act[j] = best_f[j];
            } else {
                break;
            }
        }
    }

    // ---- ALT / QUAL (src/basetype.cpp:170-196) ----
    int n_alt = 0;
    int alt[4] = {0, 0, 0, 0};
    double af[4] = {0, 0, 0, 0};
    for (int k = 0; k < n_act; ++k)
        if (act[k] != ref_code) { alt[n_alt] = act[k]; af[n_alt] = f_act[act[k]]; ++n_alt; }
    double qual = 0.0;
    if (n_alt) {
        double r = (double)dep[act[0]] / dtot;
        if (n_act == 1 && total > 10 && r > 0.5) { qual = 5000.0; flags |= BV_FLAG_MONO_QUAL; }
        else qual = qual_from_chi(chi);
    }

    // ---- strand bias (src/basetype.cpp:244-295): CVG row = ref vs all non-ref ACGT; VCF row = ref vs ALT ----
    double fs_cvg, fs_vcf = 0.0;
    {
        int rf = 0, rr = 0, af_ = 0, ar = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b == ref_code) { rf += fwd[b]; rr += rev[b]; } else { af_ += fwd[b]; ar += rev[b]; }
        }
        fs_cvg = fs_from_table(a.logfact, rf, rr, af_, ar);
        if (n_alt) {
            int vf = 0, vr = 0;
            for (int k = 0; k < n_alt; ++k) { vf += fwd[alt[k]]; vr += rev[alt[k]]; }
            if (vf == af_ && vr == ar) fs_vcf = fs_cvg;   // same 2x2 table
            else fs_vcf = fs_from_table(a.logfact, rf, rr, vf, vr);
        }
    }

    // ---- record ----
    __syncwarp();
    for (int i = lane; i < nb; i += 32) ws.hist[i] = 0;   // histogram back to all-zero for the next site
    if (lane == 0) {
        bv_site_out& r = ws.rec;
#pragma unroll
        for (int b = 0; b < 4; ++b) { r.depth[b] = dep[b]; r.fwd[b] = fwd[b]; r.rev[b] = rev[b]; }
        r.depth_other = other;
        r.reserved0 = 0;
        r.n_alt = (uint8_t)n_alt;
#pragma unroll
        for (int k = 0; k < 4; ++k) { r.alt[k] = (uint8_t)alt[k]; r.af[k] = af[k]; }
        r.n_active = (uint8_t)((total > 0) ? n_act : 0);
        r.flags = (uint8_t)flags;
        r.em_calls = (uint8_t)em_calls;
        r.qual = qual;
        r.chi2 = chi;
        r.fs_cvg = fs_cvg;
        r.fs_vcf = fs_vcf;
    }
    __syncwarp();
    if (lane < 8) {
        const uint4* src = reinterpret_cast<const uint4*>(&ws.rec);
        uint4* dst = reinterpret_cast<uint4*>(a.out + site);
        dst[lane] = src[lane];
    }
    __syncwarp();
}

}  // namespace bv
