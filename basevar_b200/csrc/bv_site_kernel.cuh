// bv_site_kernel.cuh -- the per-site statistical core of `basevar basetype` as one fused sm_100a kernel.
//
// One warp owns one genomic site at a time (sites are the embarrassingly parallel axis, samples the
// reduction axis).  For its site the warp
//   1. streams the three u8 planes of the site row (base, qual, strand) with 128-bit loads and counts every
//      read into a warp-private shared-memory histogram hist[strand 0..1][base 0..4][phred 0..95]
//                                                            (BaseType::BaseType, src/basetype.cpp:45-71;
//                                                             strand_bias counting, src/basetype.cpp:252-274)
//   2. sweeps the touched phred range once: per-base depths and the 2x4 strand table by warp reductions
//   3. decides the active alleles (depth/total >= min_af).  With ONE active allele the reference's EM has a
//      closed form (AF == 1.0 exactly) and nothing else is computed.  Otherwise the non-empty (base, phred)
//      bins are compacted and EM + LRT backward elimination run on the bins: all reads of one bin are
//      exchangeable in e_step/m_step (src/algorithm.h:148-198), so a bin of c reads contributes c * (per-read
//      term); lanes own bins, allele sums are warp-shuffle reductions
//                                                            (EM, src/algorithm.h:210-255;
//                                                             _f / lrt, src/basetype.cpp:105-199)
//   4. QUAL (chi2 survival via kf_gammaq) and the two Fisher strand-bias tests
//                                                            (src/basetype.cpp:180-194, :244-295)
//   5. writes the fixed 128-byte bv_site_out record with one coalesced store.
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

constexpr int kQStride = 128;                // phred slots per histogram row (0..93 used; 7-bit index never overflows)
constexpr int kHistRows = 10;                // (strand 0..1) x (A,C,G,T,other)
constexpr int kHistWords = kHistRows * kQStride;
constexpr int kSmemBins = 160;               // compact bins kept in shared memory; more spill to global scratch
constexpr int kMaxBins = 5 * kQStride;       // upper bound on distinct (base, phred) bins
constexpr int kLutOneMinusEps = 0;           // lut[0][q] = 1 - eps(q)
constexpr int kLutEpsThird = 1;              // lut[1][q] = eps(q) / 3
constexpr int kLutLogMatch = 2;              // lut[2][q] = log(1 - eps(q))   (glibc)
constexpr int kLutLogMis = 3;                // lut[3][q] = log(eps(q) / 3)   (glibc)

struct SiteKernelArgs {
    const uint8_t* base;
    const uint8_t* qual;
    const uint8_t* strand;
    const uint8_t* ref_base;
    bv_site_out* out;
    const double* lut;       // [4][kQStride]
    const double* logfact;   // [max_samples + 2], lgamma(k+1) from glibc
    uint32_t* bin_spill;     // [total warps][kMaxBins] global copy of the compact bins (used when > kSmemBins)
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
    uint32_t hist[kHistWords];   // dense histogram, all-zero between sites
    uint32_t bins[kSmemBins];    // compact non-empty (base, phred) bins: (base << 29) | (phred << 22) | count
    bv_site_out rec;             // record staging for one coalesced 128-byte store
};

__device__ __forceinline__ uint32_t pack_bin(uint32_t b, uint32_t q, uint32_t count) {
    return (b << 29) | (q << 22) | count;
}
__device__ __forceinline__ uint32_t bin_base(uint32_t p) { return p >> 29; }
__device__ __forceinline__ uint32_t bin_qual(uint32_t p) { return (p >> 22) & 0x7fu; }
__device__ __forceinline__ uint32_t bin_count(uint32_t p) { return p & 0x3fffffu; }

// streaming loads: read once, do not pollute L1
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Per-lane state of one site row.
struct LaneCounts {
    uint32_t qmin, qmax, flags;
};

// Count the 4 cells of one 32-bit word of each plane.  Histogram slot of a read: (base << 8) | (strand << 7) | phred.
__device__ __forceinline__ void count_word(uint32_t wb, uint32_t wq, uint32_t ws, uint32_t* hist, LaneCounts& lc) {
    // bit 7 of each byte of `m` is set iff the base code is < 5 (A,C,G,T,other): the cell is counted
    uint32_t m = ~((((wb | 0x80808080u) - 0x05050505u) | wb)) & 0x80808080u;
    if (m) {
        // rare input errors, checked per word: a counted cell with phred > 93 or a strand symbol other than +/-
        const uint32_t bytes = (m >> 7) * 0xffu;
        const uint32_t badq = ((wq + 0x22222222u) | wq) & m, bads = ws & 0xfefefefeu & bytes;
        if (badq | bads) lc.flags |= (badq ? BV_FLAG_BAD_QUAL : 0u) | (bads ? BV_FLAG_BAD_STRAND : 0u);
        const uint32_t sq = (wq & 0x7f7f7f7fu) | ((ws & 0x01010101u) << 7);   // per byte: strand << 7 | phred
        do {
            int top;                         // bit 7 of the highest counted cell
            asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(m));
            const int sh = top - 7;
            m ^= 1u << top;
            const uint32_t b = (wb >> sh) & 0xffu;
            const uint32_t x = (sq >> sh) & 0xffu;
            atomicAdd(&hist[(b << 8) | x], 1u);
            const uint32_t q = x & 0x7fu;
            lc.qmin = min(lc.qmin, q);
            lc.qmax = max(lc.qmax, q);
        } while (m);
    }
    // Lanes leave the cell loop after different trip counts; without an explicit barrier the warp stays split
    // into fragments for the rest of the row (measured: 6.5 active lanes per streaming load).
    __syncwarp();
}

// valid = number of real cells in this 16-cell vector (>= 16 for all but the row's last vector): padding cells
// are turned into 'N'
__device__ __forceinline__ void mask_tail(uint4& vb, int valid) {
    uint32_t w[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int left = valid - 4 * k;
        const uint32_t keep = left >= 4 ? 0xffffffffu : (left <= 0 ? 0u : (0xffffffffu >> (8 * (4 - left))));
        w[k] = (w[k] & keep) | (0x05050505u & ~keep);
    }
    vb = make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void count_vec(uint4 vb, const uint4& vq, const uint4& vs, int valid, uint32_t* hist,
                                          LaneCounts& lc) {
    if (valid < 16) mask_tail(vb, valid);
    count_word(vb.x, vq.x, vs.x, hist, lc);
    count_word(vb.y, vq.y, vs.y, hist, lc);
    count_word(vb.z, vq.z, vs.z, hist, lc);
    count_word(vb.w, vq.w, vs.w, hist, lc);
}

// =====================================================================================================================
// TMA-staged streaming (sm_100a): every warp owns a private ring of kStages stage buffers in shared memory.  One stage
// holds one chunk (<= kChunk cells) of the three planes of one site row.  Lane 0 issues the three bulk copies
// (cp.async.bulk, SASS UBLKCP) of the chunk kStages-1 units ahead and arms the stage's mbarrier with the byte count;
// all lanes wait on the mbarrier phase, then read the chunk from shared memory.  No register staging, no LDG in the
// hot loop, and the next rows are in flight while EM/Fisher of the current site run.
// =====================================================================================================================
constexpr int kChunk = 512;     // cells per stage and plane: one 16-cell vector per lane
constexpr int kStages = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One stage: three planes of one chunk.
struct __align__(128) Stage {
    uint8_t base[kChunk];
    uint8_t qual[kChunk];
    uint8_t strand[kChunk];
};

// Count one chunk that sits in shared memory.  `cells` = real cells of this chunk that belong to this lane's 16-cell
// vector and beyond (<= 0: the lane has nothing).  The cell loop is warp-uniform: its trip count is the warp maximum of
// the per-lane counted cells and lanes that run out are predicated off, so the warp never splits (a divergent loop
// left the warp in fragments for the rest of the row: 6.5 active lanes per load, measured with ncu).
__device__ __forceinline__ void count_chunk(const Stage& st, int lane_cells, uint32_t* hist, LaneCounts& lc) {
    const int lane = threadIdx.x & 31;
    uint4 vb = make_uint4(0x05050505u, 0x05050505u, 0x05050505u, 0x05050505u);
    if (lane_cells > 0) vb = *reinterpret_cast<const uint4*>(st.base + lane * 16);
    if (lane_cells < 16) mask_tail(vb, lane_cells);
    // t: bit (8*j + k) set <=> byte j of word k holds a counted base code (< 5)
    const uint32_t n0 = (((vb.x | 0x80808080u) - 0x05050505u) | vb.x) & 0x80808080u;
    const uint32_t n1 = (((vb.y | 0x80808080u) - 0x05050505u) | vb.y) & 0x80808080u;
    const uint32_t n2 = (((vb.z | 0x80808080u) - 0x05050505u) | vb.z) & 0x80808080u;
    const uint32_t n3 = (((vb.w | 0x80808080u) - 0x05050505u) | vb.w) & 0x80808080u;
    uint32_t t = ((n0 >> 7) | (n1 >> 6) | (n2 >> 5) | (n3 >> 4)) ^ 0x0f0f0f0fu;
    const int n = (int)__reduce_max_sync(0xffffffffu, (uint32_t)__popc(t));
    const uint8_t* cellp = st.base + lane * 16;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const bool on = t != 0u;
        int top;
        asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(t));
        t &= ~(1u << (top & 31));
        const int cell = ((top & 3) << 2) | ((top >> 3) & 3);   // word k = top & 7 (0..3), byte j = top >> 3
        if (on) {
            const uint32_t b = cellp[cell];
            const uint32_t q = cellp[cell + kChunk] & 0x7fu;     // rows have 128 slots: no overflow whatever the byte
            const uint32_t s = cellp[cell + 2 * kChunk];
            if (s > BV_STRAND_REV) lc.flags |= BV_FLAG_BAD_STRAND;
            atomicAdd(&hist[(b << 8) | ((s & 1u) << 7) | q], 1u);
            lc.qmin = min(lc.qmin, q);
            lc.qmax = max(lc.qmax, q);
        }
    }
}

// ---- EM on compact bins (src/algorithm.h:210-255) ----------------------------------------------------------------
// bins: nb packed (base, phred, count) entries; lml: nb doubles of scratch (log marginal likelihood per bin).
// subset: bit j set => allele j in the candidate combination.  f: initial frequencies in (NOT renormalised,
// src/basetype.cpp:93-103), estimated frequencies out.  Returns the sum of log marginal likelihoods under the
// second-to-last frequency vector, exactly what _f() sums (src/basetype.cpp:119-120).
struct Freq4 {
    double v0, v1, v2, v3;
};

__device__ __noinline__ double em_bins(const uint32_t* bins, double* lml, int nb, const double* s_lut, int subset,
                                       double total, Freq4& f, const SiteKernelArgs& a, uint32_t& flags) {
    const int lane = threadIdx.x & 31;
    double f0 = f.v0, f1 = f.v1, f2 = f.v2, f3 = f.v3;
    int it = a.em_max_iter;
    bool first = true;
    for (;;) {
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0, delta = 0;
        bool big = false;
#pragma unroll 1
        for (int i = lane; i < nb; i += 32) {
            const uint32_t p = bins[i];
            const uint32_t b = bin_base(p), q = bin_qual(p);
            const double cd = (double)bin_count(p);
            const double ome = s_lut[kLutOneMinusEps * kQStride + q], e3 = s_lut[kLutEpsThird * kQStride + q];
            // e_step (algorithm.h:160-172): lik*freq summed in A,C,G,T order; alleles outside the subset have
            // freq 0 and add an exact +0.0, so they are skipped
            double l0 = 0, l1 = 0, l2 = 0, l3 = 0, m = 0;
            if (subset & 1) { l0 = (b == 0 ? ome : e3) * f0; m += l0; }
            if (subset & 2) { l1 = (b == 1 ? ome : e3) * f1; m += l1; }
            if (subset & 4) { l2 = (b == 2 ? ome : e3) * f2; m += l2; }
            if (subset & 8) { l3 = (b == 3 ? ome : e3) * f3; m += l3; }
            const double llh = log(m);
            if (!first) {
                const double diff = llh - lml[i];
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
#pragma unroll 1
    for (int i = lane; i < nb; i += 32) ll += (double)bin_count(bins[i]) * lml[i];
    f.v0 = f0; f.v1 = f1; f.v2 = f2; f.v3 = f3;
    return warp_sum(ll);
}

// Log-likelihood of the single-allele model {b} (an EM whose answer is closed form):
// after the first m_step f_b == 1.0 exactly (every posterior is x/x), so every later marginal is L_b itself and
// the reported log marginal is log(1-eps) or log(eps/3) -- both tabulated on the host with glibc.  A bin of base b
// with phred 0 has L_b == 0: the reference then divides 0/0 and everything becomes NaN.
__device__ __noinline__ double single_allele_ll(const uint32_t* bins, int nb, const double* s_lut, int b_allele,
                                                bool& is_nan) {
    const int lane = threadIdx.x & 31;
    double ll = 0;
    bool bad = false;
#pragma unroll 1
    for (int i = lane; i < nb; i += 32) {
        const uint32_t p = bins[i];
        const uint32_t b = bin_base(p), q = bin_qual(p);
        const bool match = ((int)b == b_allele);
        if (match && q == 0) bad = true;
        ll += (double)bin_count(p) * s_lut[(match ? kLutLogMatch : kLutLogMis) * kQStride + q];
    }
    is_nan = __any_sync(0xffffffffu, bad);
    return warp_sum(ll);
}

__device__ __forceinline__ double sel4(int j, double v0, double v1, double v2, double v3) {
    return j == 0 ? v0 : j == 1 ? v1 : j == 2 ? v2 : v3;
}
__device__ __forceinline__ uint32_t sel4u(int j, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
    return j == 0 ? v0 : j == 1 ? v1 : j == 2 ? v2 : v3;
}
// position of the k-th (k >= 0) set bit of a 4-bit mask
__device__ __forceinline__ int nth_set_bit(uint32_t mask, int k) {
    int pos = -1;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        if (mask & (1u << b)) {
            if (k == 0 && pos < 0) pos = b;
            --k;
        }
    }
    return pos;
}

// exact `(double)dep / (double)total >= min_af` (src/basetype.cpp:137) with the trivial cases short-cut
__device__ __forceinline__ bool is_active(uint32_t dep, uint32_t total, double dtot, double min_af) {
    if (dep == 0) return 0.0 >= min_af;
    if (dep == total) return 1.0 >= min_af;
    return (double)dep / dtot >= min_af;
}

// State of the LRT of one site (warp-uniform).
struct LrtState {
    Freq4 fa;          // frequencies of the accepted model
    double chi;        // last chi_sqrt_value
    uint32_t act;      // bit b set => base b active
    int n_act;
    uint32_t em_calls;
    uint32_t flags;
};

// ---- sites with >= 2 active alleles: compact the bins, EM on the full set, backward elimination -----------------------
// (src/basetype.cpp:144-168).  Out of line: ~25 % of sites at N=1000/0.1x, ~2 % at N=10,000.
__device__ __noinline__ void lrt_multi(WarpScratch& ws, const double* s_lut, const SiteKernelArgs& a, uint32_t warp_global,
                                       uint32_t qmin, uint32_t qmax, uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3,
                                       double dtot, LrtState& st) {
    const int lane = threadIdx.x & 31;
    // sweep 2: histogram back to zero while the non-empty (base, phred) bins are compacted, in (base, phred) order
    int nb = 0;
    uint32_t* gbins = a.bin_spill + (size_t)warp_global * kMaxBins;
#pragma unroll 1
    for (int b = 0; b < 5; ++b) {
#pragma unroll 1
        for (uint32_t q0 = qmin; q0 <= qmax; q0 += 32) {
            const uint32_t q = q0 + lane;
            uint32_t v = 0;
            if (q <= qmax) {
                v = ws.hist[(2 * b) * kQStride + q] + ws.hist[(2 * b + 1) * kQStride + q];
                ws.hist[(2 * b) * kQStride + q] = 0;
                ws.hist[(2 * b + 1) * kQStride + q] = 0;
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, v != 0);
            if (v) {
                const int pos = nb + __popc(bal & ((1u << lane) - 1u));
                const uint32_t p = pack_bin(b, q, v);
                if (pos < kSmemBins) ws.bins[pos] = p;
                gbins[pos] = p;
            }
            nb += __popc(bal);
        }
    }
    __syncwarp();
    // bins live in shared memory unless there are more than kSmemBins of them (then the global copy is used);
    // the EM's per-bin state overlays the (now all-zero) histogram
    const uint32_t* bins = nb <= kSmemBins ? ws.bins : gbins;
    double* lml = reinterpret_cast<double*>(ws.hist);

    const double i0 = (double)d0 / dtot, i1 = (double)d1 / dtot, i2 = (double)d2 / dtot, i3 = (double)d3 / dtot;
    uint32_t act = st.act;
    int n_act = st.n_act;
    Freq4 fa = {(act & 1) ? i0 : 0.0, (act & 2) ? i1 : 0.0, (act & 4) ? i2 : 0.0, (act & 8) ? i3 : 0.0};
    uint32_t flags = st.flags;
    double chi = 0.0;
    double lr_alt = em_bins(bins, lml, nb, s_lut, (int)act, dtot, fa, a, flags);
    uint32_t em_calls = 1;
#pragma unroll 1
    for (int n = n_act - 1; n > 0; --n) {
        // the n-subsets of the n+1 active bases in the lexicographic order of
        // src/external/combinations.h:19-84: the i-th subset drops the (n-i)-th active base
        double best_chi = 0, best_lr = 0;
        Freq4 best_f = {0, 0, 0, 0};
        uint32_t best_set = 0;
#pragma unroll 1
        for (int i = 0; i <= n; ++i) {
            const uint32_t sub = act & ~(1u << nth_set_bit(act, n - i));
            Freq4 g = {(sub & 1) ? i0 : 0.0, (sub & 2) ? i1 : 0.0, (sub & 4) ? i2 : 0.0, (sub & 8) ? i3 : 0.0};
            if (g.v0 + g.v1 + g.v2 + g.v3 == 0) flags |= BV_FLAG_ZERO_SUBSET;   // the reference throws (basetype.cpp:113)
            double lr;
            if (n == 1) {
                const int single = __ffs(sub) - 1;
                bool bad;
                lr = single_allele_ll(bins, nb, s_lut, single, bad);
                double v = 1.0;
                if (bad) { lr = __longlong_as_double(0x7ff8000000000000ll); v = lr; }
                g.v0 = single == 0 ? v : 0.0; g.v1 = single == 1 ? v : 0.0; g.v2 = single == 2 ? v : 0.0; g.v3 = single == 3 ? v : 0.0;
            } else {
                lr = em_bins(bins, lml, nb, s_lut, (int)sub, dtot, g, a, flags);
            }
            if (em_calls < 255) ++em_calls;
            const double c = 2 * (lr_alt - lr);
            // std::min_element keeps the FIRST minimum (algorithm.h:24-27).  Alleles with identical read
            // multisets have equal likelihood; the reference's pick between them hangs on the rounding noise
            // of its read-order sums.  Values that agree to rounding noise are treated as the tie they are:
            // the earlier subset stays and the site is flagged.
            const double tie_tol = 1e-11 * (fabs(lr_alt) + fabs(lr));
            if (i > 0 && fabs(c - best_chi) <= tie_tol) flags |= BV_FLAG_LRT_TIE;
            if (i == 0 || c < best_chi - tie_tol) {
                best_chi = c; best_lr = lr; best_set = sub; best_f = g;
            }
        }
        lr_alt = best_lr;
        chi = best_chi;
        if (fabs(chi - a.lrt_threshold) < 1e-9 * a.lrt_threshold) flags |= BV_FLAG_NEAR_LRT;
        if (chi < a.lrt_threshold) {
            act = best_set; n_act = n; fa = best_f;
        } else {
            break;
        }
    }
    // the EM state overlaid the histogram: back to all-zero for the next site
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < 2 * nb; i += 32) ws.hist[i] = 0;
    st.fa = fa; st.chi = chi; st.act = act; st.n_act = n_act; st.em_calls = em_calls; st.flags = flags;
}

// ---- the warp-per-site core: everything after the row has been histogrammed ----------------------------------------
__device__ __forceinline__ void site_finish(WarpScratch& ws, const double* s_lut, const SiteKernelArgs& a,
                                            uint32_t site, uint32_t warp_global, LaneCounts& lc) {
    const int lane = threadIdx.x & 31;
    LrtState st;
    st.flags = __reduce_or_sync(0xffffffffu, lc.flags);
    const uint32_t qmin = __reduce_min_sync(0xffffffffu, lc.qmin);
    const uint32_t qmax = __reduce_max_sync(0xffffffffu, lc.qmax);
    if (qmin <= qmax && qmax > BV_QUAL_MAX) st.flags |= BV_FLAG_BAD_QUAL;   // phred 94..127 (>= 128 wraps modulo 128)
    __syncwarp();

    // ---- sweep 1: depths and strand table from the touched phred range ----
    // row 2*base + strand ('+' = 0, '-' = 1).  A counted cell whose strand is neither sets BV_FLAG_BAD_STRAND and is
    // counted by the low bit of its code: the reference throws on such a site (src/basetype.cpp:271-273), so only its
    // depths and flags are specified.
    uint32_t f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0, r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
    if (qmin <= qmax) {
#pragma unroll 1
        for (uint32_t q = qmin + lane; q <= qmax; q += 32) {
            f0 += ws.hist[0 * kQStride + q]; r0 += ws.hist[1 * kQStride + q];
            f1 += ws.hist[2 * kQStride + q]; r1 += ws.hist[3 * kQStride + q];
            f2 += ws.hist[4 * kQStride + q]; r2 += ws.hist[5 * kQStride + q];
            f3 += ws.hist[6 * kQStride + q]; r3 += ws.hist[7 * kQStride + q];
            f4 += ws.hist[8 * kQStride + q]; r4 += ws.hist[9 * kQStride + q];
        }
        f0 = __reduce_add_sync(0xffffffffu, f0); f1 = __reduce_add_sync(0xffffffffu, f1);
        f2 = __reduce_add_sync(0xffffffffu, f2); f3 = __reduce_add_sync(0xffffffffu, f3);
        r0 = __reduce_add_sync(0xffffffffu, r0); r1 = __reduce_add_sync(0xffffffffu, r1);
        r2 = __reduce_add_sync(0xffffffffu, r2); r3 = __reduce_add_sync(0xffffffffu, r3);
        f4 = __reduce_add_sync(0xffffffffu, f4 + r4);
    }
    const uint32_t d0 = f0 + r0, d1 = f1 + r1, d2 = f2 + r2, d3 = f3 + r3, other = f4;
    const uint32_t total = d0 + d1 + d2 + d3 + other;
    const double dtot = (double)total;

    // ---- reference base ----
    int ref_char = a.ref_base[site];
    if (ref_char >= 'a' && ref_char <= 'z') ref_char -= 32;  // toupper (src/basetype.cpp:171)
    const int ref_code = ref_char == 'A' ? 0 : ref_char == 'C' ? 1 : ref_char == 'G' ? 2 : ref_char == 'T' ? 3 : -1;

    // ---- lrt (src/basetype.cpp:130-199): active set ----
    st.act = 0;
    if (total > 0) {
        st.act |= is_active(d0, total, dtot, a.min_af) ? 1u : 0u;
        st.act |= is_active(d1, total, dtot, a.min_af) ? 2u : 0u;
        st.act |= is_active(d2, total, dtot, a.min_af) ? 4u : 0u;
        st.act |= is_active(d3, total, dtot, a.min_af) ? 8u : 0u;
    }
    st.n_act = __popc(st.act);
    st.fa.v0 = st.fa.v1 = st.fa.v2 = st.fa.v3 = 0.0;
    st.chi = 0.0;
    st.em_calls = 0;

    if (st.n_act >= 2) {
        lrt_multi(ws, s_lut, a, warp_global, qmin, qmax, d0, d1, d2, d3, dtot, st);
    } else {
        if (st.n_act == 1) {
            // One active allele: the reference still runs one EM; its answer is closed form (see single_allele_ll):
            // AF == 1.0 exactly, or NaN when a phred-0 read of that base exists.
            const int b = __ffs(st.act) - 1;
            bool bad = false;
            if (qmin == 0) bad = (ws.hist[(2 * b) * kQStride] + ws.hist[(2 * b + 1) * kQStride]) != 0;
            const double v = bad ? __longlong_as_double(0x7ff8000000000000ll) : 1.0;
            st.fa.v0 = b == 0 ? v : 0.0; st.fa.v1 = b == 1 ? v : 0.0; st.fa.v2 = b == 2 ? v : 0.0; st.fa.v3 = b == 3 ? v : 0.0;
            st.em_calls = 1;
        }
        __syncwarp();
        if (qmin <= qmax) {   // sweep 2: histogram back to zero
#pragma unroll 1
            for (uint32_t q = qmin + lane; q <= qmax; q += 32) {
#pragma unroll
                for (int r = 0; r < kHistRows; ++r) ws.hist[r * kQStride + q] = 0;
            }
        }
    }

    // ---- ALT / QUAL (src/basetype.cpp:170-196) ----
    const uint32_t alt_set = (ref_code >= 0) ? (st.act & ~(1u << ref_code)) : st.act;
    const int n_alt = __popc(alt_set);
    double qual = 0.0;
    if (n_alt) {
        const int first_act = __ffs(st.act) - 1;
        const double r = (double)sel4u(first_act, d0, d1, d2, d3) / dtot;
        if (st.n_act == 1 && total > 10 && r > 0.5) { qual = 5000.0; st.flags |= BV_FLAG_MONO_QUAL; }
        else qual = qual_from_chi(st.chi);
    }

    // ---- strand bias (src/basetype.cpp:244-295): CVG row = ref vs all non-ref ACGT; VCF row = ref vs ALT ----
    double fs_cvg = 0.0, fs_vcf = 0.0;
    {
        const int rf = ref_code < 0 ? 0 : (int)sel4u(ref_code, f0, f1, f2, f3);
        const int rr = ref_code < 0 ? 0 : (int)sel4u(ref_code, r0, r1, r2, r3);
        const int af_ = (int)(f0 + f1 + f2 + f3) - rf, ar = (int)(r0 + r1 + r2 + r3) - rr;
        // a table with an empty row or column has a single possible outcome: p == 1, FS == 0 (kfunc.c:256)
        if ((af_ | ar) != 0 && (rf | rr) != 0) fs_cvg = fs_from_table(a.logfact, rf, rr, af_, ar);
        if (n_alt) {
            const int vf = (int)(((alt_set & 1) ? f0 : 0u) + ((alt_set & 2) ? f1 : 0u) + ((alt_set & 4) ? f2 : 0u) + ((alt_set & 8) ? f3 : 0u));
            const int vr = (int)(((alt_set & 1) ? r0 : 0u) + ((alt_set & 2) ? r1 : 0u) + ((alt_set & 4) ? r2 : 0u) + ((alt_set & 8) ? r3 : 0u));
            if (vf == af_ && vr == ar) fs_vcf = fs_cvg;   // same 2x2 table
            else fs_vcf = fs_from_table(a.logfact, rf, rr, vf, vr);
        }
    }

    // ---- record ----
    __syncwarp();
    if (lane == 0) {
        bv_site_out& r = ws.rec;
        r.depth[0] = d0; r.depth[1] = d1; r.depth[2] = d2; r.depth[3] = d3;
        r.depth_other = other;
        r.reserved0 = 0;
        r.fwd[0] = f0; r.fwd[1] = f1; r.fwd[2] = f2; r.fwd[3] = f3;
        r.rev[0] = r0; r.rev[1] = r1; r.rev[2] = r2; r.rev[3] = r3;
        // ALT alleles in ACGT order of the active set (src/basetype.cpp:172-177)
        uint32_t alts = 0;
        double af[4] = {0.0, 0.0, 0.0, 0.0};
        int k = 0;
        if (alt_set & 1) { af[k] = st.fa.v0; alts |= 0u << (8 * k); ++k; }
        if (alt_set & 2) { af[k] = st.fa.v1; alts |= 1u << (8 * k); ++k; }
        if (alt_set & 4) { af[k] = st.fa.v2; alts |= 2u << (8 * k); ++k; }
        if (alt_set & 8) { af[k] = st.fa.v3; alts |= 3u << (8 * k); ++k; }
        r.n_alt = (uint8_t)n_alt;
        r.alt[0] = (uint8_t)alts; r.alt[1] = (uint8_t)(alts >> 8); r.alt[2] = (uint8_t)(alts >> 16); r.alt[3] = (uint8_t)(alts >> 24);
        r.af[0] = af[0]; r.af[1] = af[1]; r.af[2] = af[2]; r.af[3] = af[3];
        r.n_active = (uint8_t)st.n_act;
        r.flags = (uint8_t)st.flags;
        r.em_calls = (uint8_t)st.em_calls;
        r.qual = qual;
        r.chi2 = st.chi;
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
