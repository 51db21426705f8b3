// bv_site_kernel.cuh -- the per-site statistical core of `basevar basetype` as one fused sm_100a kernel.
//
// One warp owns one genomic site at a time (sites are the embarrassingly parallel axis, samples the reduction
// axis).  Persistent CTAs; every warp streams its own sequence of site rows through a private ring of
// TMA-filled shared-memory stages (cp.async.bulk + mbarrier, SASS UBLKCP / SYNCS).
//
//   pass 1 (every site, every cell)   base + strand planes only.  SIMD-in-register byte arithmetic: cells that
//        hold the reference base are counted with dp4a, 16 cells per lane per step, no per-cell work and no atomics;
//        the (rare) counted cells that are NOT the reference base are counted one by one into ten shared-memory
//        counters.  Result: per-base depths and the 2x4 strand table, bit-exact
//                                                            (BaseType::BaseType, src/basetype.cpp:45-71;
//                                                             strand_bias counting, src/basetype.cpp:252-274).
//   fast finish   a site whose counted cells all equal the reference base has a closed-form record (one active
//        allele == REF, no ALT, FS 0): 32 lanes compose the 128-byte record in registers and store it coalesced.
//   pass 2 (only sites whose result depends on base qualities: >= 2 active alleles, or a single active allele that
//        is not REF)   the row's base + qual chunks are fetched again (L2 / HBM) and the covered cells are
//        histogrammed by (base, phred); the non-empty bins are compacted and EM + LRT backward elimination run on
//        the bins: all reads of one bin are exchangeable in e_step/m_step (src/algorithm.h:148-198), so a bin of c
//        reads contributes c * (per-read term); lanes own bins, allele sums are warp-shuffle reductions
//                                                            (EM, src/algorithm.h:210-255;
//                                                             _f / lrt, src/basetype.cpp:105-199).
//   QUAL (chi2 survival via kf_gammaq) and the two Fisher strand-bias tests
//                                                            (src/basetype.cpp:180-194, :244-295).
//
// The qual plane is therefore read only where the reference's outputs depend on it; everywhere else the sample
// axis costs two byte loads and ~3 integer instructions per cell.  Per-read likelihood values (1-eps, eps/3) come
// from a host-computed table (glibc exp), so they are bit-identical with the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/basevar_b200.h"
#include "bv_math.cuh"

namespace bv {

#ifndef BV_WARPS
#define BV_WARPS 24
#endif
constexpr int kWarps = BV_WARPS;             // warps per CTA, one CTA per SM
constexpr int kChunk = 1024;                 // pass 1: cells per stage and plane, two 16-cell vectors per lane
constexpr int kStages = 2;                   // pass 1: ring depth per warp (one unit in flight while one is scanned)
constexpr int kP2Chunk = 512;                // pass 2: cells per buffer and plane, one 16-cell vector per lane
constexpr int kQStride = 128;                // phred slots per LUT row
constexpr int kQSlots = 96;                  // phred slots per histogram row (0..93 valid; larger values clamp to 95)
constexpr int kHistWords = 5 * kQSlots;      // (A,C,G,T,other) x phred
constexpr int kSmemBins = 160;               // compact bins kept in shared memory; more spill to global scratch
constexpr int kMaxBins = kHistWords;         // upper bound on distinct (base, phred) bins
constexpr int kLutOneMinusEps = 0;           // lut[0][q] = 1 - eps(q)
constexpr int kLutEpsThird = 1;              // lut[1][q] = eps(q) / 3
constexpr int kLutLogMatch = 2;              // lut[2][q] = log(1 - eps(q))   (glibc)
constexpr int kLutLogMis = 3;                // lut[3][q] = log(eps(q) / 3)   (glibc)
constexpr uint32_t kFull = 0xffffffffu;

struct SiteKernelArgs {
    const uint8_t* base;
    const uint8_t* qual;
    const uint8_t* strand;
    const uint8_t* ref_base;
    bv_site_out* out;
    const double* lut;       // [4][kQStride]
    const double* logfact;   // [max_samples + 2], lgamma(k+1) from glibc
    uint32_t* bin_spill;     // [total warps][kMaxBins] global copy of the compact bins (used when > kSmemBins)
    double* lml_spill;       // [total warps][kMaxBins] per-bin EM state for the same case
    uint64_t pitch;
    uint32_t n_sites;
    uint32_t n_samples;
    double min_af;           // (double)(float)min_af
    double em_eps;           // (double)(float)0.001
    double lrt_threshold;
    int em_max_iter;
    int abs_mode;
};

// ---- shared memory ---------------------------------------------------------------------------------------------------
struct __align__(128) Stage {      // pass 1: one chunk of the base and strand planes of one site row
    uint8_t base[kChunk];
    uint8_t strand[kChunk];
};
struct __align__(128) P2Buf {      // pass 2: one chunk of the base and qual planes
    uint8_t base[kP2Chunk];
    uint8_t qual[kP2Chunk];
};

struct __align__(128) WarpSmem {
    Stage stage[kStages];
    P2Buf p2[2];
    uint32_t hist[kHistWords];   // (base, phred) histogram of pass 2, all-zero between sites; the EM's per-bin state
                                 // (one double per compact bin) overlays it once the bins are compacted
    uint32_t bins[kSmemBins];    // compact non-empty bins: (base << 29) | (phred << 22) | count
    uint32_t nr_cnt[12];         // pass 1: counted cells that are not the reference base, [2*base + strand]
    double emf[4];               // EM: allele frequencies in / out
    double res_f[4];             // LRT: frequencies of the accepted model
    double best_f[4];            // LRT: frequencies of the best candidate subset of the current round
    double res_chi;              // LRT: last chi_sqrt_value
    uint32_t flag_word;          // BV_FLAG_* raised inside out-of-line code
    uint32_t p2_phase;           // mbarrier phase bits of p2bar[]
    // state of the streaming loop while the warp is away in the slow path (see stream_sites)
    uint32_t sv_p_site, sv_p_off, sv_p_slot, sv_c_slot, sv_c_par, sv_site, sv_ref_raw;
    // the slow site handed from stream_sites to site_slow
    uint32_t slow_site, slow_n_ref, slow_n_rev, slow_bad, slow_ref_code;
    uint64_t full[kStages];
    uint64_t p2bar[2];
    alignas(16) bv_site_out rec; // record staging for one coalesced 128-byte store (slow path)
};

struct __align__(128) CtaShared {
    double lut[4 * kQStride];
    uint32_t tail_keep[4];       // byte masks of the row's last, partial 16-cell vector (all ones when N % 16 == 0)
    uint32_t pad_[4];
    uint32_t gfix[kQSlots];      // ceil(2^20 * (log(1-eps(q)) - log(eps(q)/3))): log-likelihood gain of calling a read's
                                 // own base, fixed point, rounded up (see lrt_bound)
    SiteKernelArgs a;            // kernel parameters for out-of-line device functions (a reference to the
                                 // __global__ parameter itself would force a local-memory copy)
};

// The kernel's dynamic shared memory: CtaShared, then one WarpSmem per warp.  Device functions reach it through these
// accessors (not through pointer arguments) so that the compiler knows the address space and emits LDS/STS/ATOMS.
extern __shared__ __align__(128) unsigned char bv_smem_raw[];
__device__ __forceinline__ CtaShared& cta_shared() { return *reinterpret_cast<CtaShared*>(bv_smem_raw); }
__device__ __forceinline__ WarpSmem& warp_smem() {
    return reinterpret_cast<WarpSmem*>(bv_smem_raw + sizeof(CtaShared))[threadIdx.x >> 5];
}

__device__ __forceinline__ uint32_t pack_bin(uint32_t b, uint32_t q, uint32_t count) {
    return (b << 29) | (q << 22) | count;
}
__device__ __forceinline__ uint32_t bin_base(uint32_t p) { return p >> 29; }
__device__ __forceinline__ uint32_t bin_qual(uint32_t p) { return (p >> 22) & 0x7fu; }
__device__ __forceinline__ uint32_t bin_count(uint32_t p) { return p & 0x3fffffu; }

// ---- TMA bulk copies + mbarrier (sm_90+; SASS UBLKCP / SYNCS) -----------------------------------------------------------
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

// =====================================================================================================================
// Pass 1: SIMD-in-register scan of 4 cells (one 32-bit word of the base plane and of the strand plane).
// Every mask has its information in bit 7 of each byte.
// =====================================================================================================================
struct ScanAcc {
    uint32_t nref;     // 128 * (# cells holding the reference base)
    uint32_t nrev;     // 128 * (# of those on the '-' strand)
    uint32_t nonref;   // != 0: the row has counted cells that are not the reference base
    uint32_t bad;      // != 0: a counted cell has a strand code other than +/-
};

__device__ __forceinline__ uint32_t scan_word(uint32_t wb, uint32_t ws, uint32_t refw, ScanAcc& A) {
    // counted: base code < 5 (A,C,G,T,other); bytes >= 0x80 are never counted
    const uint32_t m = ~(((wb | 0x80808080u) - 0x05050505u) | wb) & 0x80808080u;
    // equal to the reference base (refw = code * 0x01010101, or 0x08080808 when REF is not A/C/G/T)
    const uint32_t x = (wb ^ refw) & 0x7f7f7f7fu;
    const uint32_t eq = ~((x + 0x7f7f7f7fu) | wb) & 0x80808080u;
    // strand code >= 2 in a counted cell: the reference throws (src/basetype.cpp:271-273)
    A.bad |= (((ws & 0x7f7f7f7fu) + 0x7e7e7e7eu) | ws) & m;
    A.nref = __dp4a(eq, 0x01010101u, A.nref);
    A.nrev = __dp4a(eq, ws, A.nrev);
    return m & ~eq;
}

// =====================================================================================================================
// Pass 2: the row's base + qual chunks, fetched again (L2 / HBM) into the warp's two pass-2 buffers.  Out of line code;
// runs on the sites whose result depends on base qualities.  f(cellp, vb, lane_cells): cellp points at this lane's 16
// base cells in shared memory (quals at cellp + kP2Chunk), vb holds them with padding cells masked to 'N'.
// =====================================================================================================================
template <class F>
__device__ __forceinline__ void for_each_p2_chunk(uint32_t site, F&& f) {
    WarpSmem& W = warp_smem();
    const CtaShared& cs = cta_shared();
    const int lane = threadIdx.x & 31;
    const uint32_t N = cs.a.n_samples;
    const uint32_t row_bytes = (N + 15u) & ~15u;
    const uint32_t nchunk = (row_bytes + kP2Chunk - 1) / kP2Chunk;
    const size_t row = (size_t)site * cs.a.pitch;
    const uint8_t* gb = cs.a.base + row;
    const uint8_t* gq = cs.a.qual + row;
    uint32_t phase = W.p2_phase;
    if (lane == 0) {
        const uint32_t bytes = min((uint32_t)kP2Chunk, row_bytes);
        mbar_expect_tx(&W.p2bar[0], 2 * bytes);
        bulk_g2s(W.p2[0].base, gb, bytes, &W.p2bar[0]);
        bulk_g2s(W.p2[0].qual, gq, bytes, &W.p2bar[0]);
    }
#pragma unroll 1
    for (uint32_t c = 0; c < nchunk; ++c) {
        const uint32_t buf = c & 1u;
        if (c + 1 < nchunk && lane == 0) {   // chunk c+1 goes where chunk c-1 was (all lanes are past it: __syncwarp below)
            const uint32_t off = (c + 1) * kP2Chunk;
            const uint32_t bytes = min((uint32_t)kP2Chunk, row_bytes - off);
            mbar_expect_tx(&W.p2bar[buf ^ 1u], 2 * bytes);
            bulk_g2s(W.p2[buf ^ 1u].base, gb + off, bytes, &W.p2bar[buf ^ 1u]);
            bulk_g2s(W.p2[buf ^ 1u].qual, gq + off, bytes, &W.p2bar[buf ^ 1u]);
        }
        mbar_wait(&W.p2bar[buf], (phase >> buf) & 1u);
        phase ^= 1u << buf;
        const int lane_cells = (int)N - (int)(c * kP2Chunk) - lane * 16;
        const uint8_t* cellp = W.p2[buf].base + lane * 16;
        uint4 vb = make_uint4(0x05050505u, 0x05050505u, 0x05050505u, 0x05050505u);
        if (lane_cells > 0) vb = *reinterpret_cast<const uint4*>(cellp);
        if (lane_cells < 16) mask_tail(vb, lane_cells);
        f(cellp, vb, lane_cells);
        __syncwarp();
    }
    if (lane == 0) W.p2_phase = phase;
}

// (base, phred) histogram of the covered cells of one row into W.hist.
// Returns qmin | qmax << 8 | flags << 16 (qmin > qmax: no counted cell).
__device__ __noinline__ uint32_t build_hist(uint32_t site) {
    WarpSmem& W = warp_smem();
    uint32_t qmin = 0xffu, qmax = 0, flags = 0;
    for_each_p2_chunk(site, [&](const uint8_t* cellp, const uint4& vb, int) {
        // t: bit (8*j + k) set <=> byte j of word k holds a counted base code (< 5)
        const uint32_t n0 = (((vb.x | 0x80808080u) - 0x05050505u) | vb.x) & 0x80808080u;
        const uint32_t n1 = (((vb.y | 0x80808080u) - 0x05050505u) | vb.y) & 0x80808080u;
        const uint32_t n2 = (((vb.z | 0x80808080u) - 0x05050505u) | vb.z) & 0x80808080u;
        const uint32_t n3 = (((vb.w | 0x80808080u) - 0x05050505u) | vb.w) & 0x80808080u;
        uint32_t t = ((n0 >> 7) | (n1 >> 6) | (n2 >> 5) | (n3 >> 4)) ^ 0x0f0f0f0fu;
        // warp-uniform trip count, lanes that run out are predicated off: the warp never splits
        const int n = (int)__reduce_max_sync(kFull, (uint32_t)__popc(t));
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const bool on = t != 0u;
            int top;
            asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(t));
            t &= ~(1u << (top & 31));
            const int cell = ((top & 3) << 2) | ((top >> 3) & 3);   // word k = top & 3, byte j = top >> 3
            if (on) {
                const uint32_t b = cellp[cell];
                uint32_t q = cellp[cell + kP2Chunk];
                if (q > BV_QUAL_MAX) { flags |= BV_FLAG_BAD_QUAL; q = min(q, (uint32_t)(kQSlots - 1)); }
                atomicAdd(&W.hist[b * kQSlots + q], 1u);
                qmin = min(qmin, q);
                qmax = max(qmax, q);
            }
        }
    });
    qmin = __reduce_min_sync(kFull, qmin);
    qmax = __reduce_max_sync(kFull, qmax);
    flags = __reduce_or_sync(kFull, flags);
    return qmin | (qmax << 8) | (flags << 16);
}

// ---- a bound that decides the LRT without running the EM ---------------------------------------------------------------
// Site with exactly two active alleles, REF (r) and one other base (o) carried by a few reads -- the signature of
// sequencing errors.  With L_ij the per-read likelihoods (src/basetype.cpp:61-64) and g_i = log(1-eps_i) - log(eps_i/3):
//   * every log-likelihood the EM can report for {r,o} is a sum of log(sum_j L_ij f_j) with sum_j f_j <= 1, hence
//       LL{r,o} <= sum_i log(max_j L_ij) = LL{r} + sum_{reads of o} g_i        (all phred >= 2, so 1-eps > eps/3)
//     where LL{r} = sum_{reads of r} log(1-eps_i) + sum_{other reads} log(eps_i/3) is the closed form of the
//     single-allele model (see single_allele_ll);
//   * LL{r} - LL{o} = sum_{reads of r} g_i - sum_{reads of o} g_i >= 0.56 * depth[r] - G,  G = sum_{reads of o} g_i.
// So when 2G < 23.9 and depth[r] >= 22, the first LRT round (src/basetype.cpp:151-168) picks subset {r}
// (chi_r < chi_o) with chi_r = 2 (LL{r,o} - LL{r}) <= 2G < 24 = LRT_THRESHOLD and drops o: the site ends with the
// single active allele REF, no ALT, whatever the EM would have returned.  Margins (23.9 vs 24, G rounded up in fixed
// point) dwarf the 1e-12 rounding noise of the reference's sums.
// Returns G in 2^-20 units, or 0xffffffff when a counted read has phred < 2 or > 93 (bound not applicable).
constexpr uint32_t kBoundLimit = 12530483u;   // floor(11.95 * 2^20)
__device__ __noinline__ uint32_t lrt_bound(uint32_t site, uint32_t o_code) {
    const CtaShared& cs = cta_shared();
    const uint32_t ow = o_code * 0x01010101u;
    uint32_t G = 0, bad = 0;
    for_each_p2_chunk(site, [&](const uint8_t* cellp, const uint4& vb, int lane_cells) {
        uint4 vq = make_uint4(0x02020202u, 0x02020202u, 0x02020202u, 0x02020202u);
        if (lane_cells > 0) vq = *reinterpret_cast<const uint4*>(cellp + kP2Chunk);
        const uint32_t wb[4] = {vb.x, vb.y, vb.z, vb.w}, wq[4] = {vq.x, vq.y, vq.z, vq.w};
        uint32_t eo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t m = ~(((wb[k] | 0x80808080u) - 0x05050505u) | wb[k]) & 0x80808080u;   // counted
            const uint32_t q7 = wq[k] & 0x7f7f7f7fu;
            bad |= (~(q7 + 0x7e7e7e7eu) | (q7 + 0x22222222u) | wq[k]) & m;                        // phred < 2 or > 93
            const uint32_t x = (wb[k] ^ ow) & 0x7f7f7f7fu;
            eo[k] = ~((x + 0x7f7f7f7fu) | wb[k]) & 0x80808080u;                                   // base == o
        }
        if (eo[0] | eo[1] | eo[2] | eo[3]) {
            uint32_t t = (eo[0] >> 7) | (eo[1] >> 6) | (eo[2] >> 5) | (eo[3] >> 4);   // bit (8*byte + word)
            do {
                int top;
                asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(t));
                t ^= 1u << top;
                const int cell = ((top & 3) << 2) | (top >> 3);
                const uint32_t q = cellp[cell + kP2Chunk];
                G += cs.gfix[min(q, (uint32_t)(kQSlots - 1))];
            } while (t);
        }
    });
    G = __reduce_add_sync(kFull, G);
    bad = __reduce_or_sync(kFull, bad);
    return bad ? 0xffffffffu : G;
}

// ---- EM on compact bins (src/algorithm.h:210-255) ----------------------------------------------------------------
// bins: nb packed (base, phred, count) entries; lml: nb doubles of scratch (log marginal likelihood per bin).
// subset: bit j set => allele j in the candidate combination.  W.emf: initial frequencies in (NOT renormalised,
// src/basetype.cpp:93-103), estimated frequencies out.  Returns the sum of log marginal likelihoods under the
// second-to-last frequency vector, exactly what _f() sums (src/basetype.cpp:119-120).
__device__ __noinline__ double em_bins(const uint32_t* bins, double* lml, int nb, int subset, double total) {
    WarpSmem& W = warp_smem();
    const CtaShared& cs = cta_shared();
    const double* s_lut = cs.lut;
    const int lane = threadIdx.x & 31;
    const int abs_mode = cs.a.abs_mode;
    double f0 = W.emf[0], f1 = W.emf[1], f2 = W.emf[2], f3 = W.emf[3];
    __syncwarp();
    int it = cs.a.em_max_iter;
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
                if (abs_mode == BV_EM_ABS_INT_TRUNC) {
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
        if (abs_mode == BV_EM_ABS_INT_TRUNC) more = __any_sync(kFull, big);
        else more = !(warp_sum(delta) < cs.a.em_eps);
        --it;
        if (it == 0 && lane == 0) W.flag_word |= BV_FLAG_EM_MAXITER;
        if (!more || it == 0) break;
    }
    double ll = 0;
#pragma unroll 1
    for (int i = lane; i < nb; i += 32) ll += (double)bin_count(bins[i]) * lml[i];
    if (lane == 0) { W.emf[0] = f0; W.emf[1] = f1; W.emf[2] = f2; W.emf[3] = f3; }
    __syncwarp();
    return warp_sum(ll);
}

// Log-likelihood of the single-allele model {b} (an EM whose answer is closed form):
// after the first m_step f_b == 1.0 exactly (every posterior is x/x), so every later marginal is L_b itself and
// the reported log marginal is log(1-eps) or log(eps/3) -- both tabulated on the host with glibc.  A bin of base b
// with phred 0 has L_b == 0: the reference then divides 0/0 and everything becomes NaN (returned as NaN).
__device__ __noinline__ double single_allele_ll(const uint32_t* bins, int nb, int b_allele) {
    const double* s_lut = cta_shared().lut;
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
    ll = warp_sum(ll);
    if (__any_sync(kFull, bad)) ll = __longlong_as_double(0x7ff8000000000000ll);
    return ll;
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
    // away from the boundary the product decides (one multiply instead of a division); within 1e-9 of it, the
    // reference's own expression
    const double thr = min_af * dtot, x = (double)dep;
    if (x > thr * 1.000000001) return true;
    if (x < thr * 0.999999999) return false;
    return x / dtot >= min_af;
}

// ---- sites with >= 2 active alleles: compact the bins, EM on the full set, backward elimination -----------------------
// (src/basetype.cpp:144-168).  In: the row's histogram in W.hist, phred range, depths in W.rec.depth[] (already final),
// active set.  Out: W.res_f / W.res_chi and the return value act | n_act << 4 | em_calls << 8.  The warp-uniform
// model state lives in shared memory (W.emf / W.best_f / W.res_f), not in registers that would have to survive the
// calls into em_bins.
__device__ __noinline__ uint32_t lrt_multi(uint32_t qmin, uint32_t qmax, uint32_t act) {
    WarpSmem& W = warp_smem();
    const CtaShared& cs = cta_shared();
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = blockIdx.x * kWarps + (threadIdx.x >> 5);
    // histogram back to zero while the non-empty (base, phred) bins are compacted, in (base, phred) order
    int nb = 0;
    uint32_t* gbins = cs.a.bin_spill + (size_t)warp_global * kMaxBins;
#pragma unroll 1
    for (int b = 0; b < 5; ++b) {
        if (b < 4 ? W.rec.depth[b] == 0 : W.rec.depth_other == 0) continue;
#pragma unroll 1
        for (uint32_t q0 = qmin; q0 <= qmax; q0 += 32) {
            const uint32_t q = q0 + lane;
            uint32_t v = 0;
            if (q <= qmax) {
                v = W.hist[b * kQSlots + q];
                W.hist[b * kQSlots + q] = 0;
            }
            const uint32_t bal = __ballot_sync(kFull, v != 0);
            if (v) {
                const int pos = nb + __popc(bal & ((1u << lane) - 1u));
                const uint32_t p = pack_bin(b, q, v);
                if (pos < kSmemBins) W.bins[pos] = p;
                gbins[pos] = p;
            }
            nb += __popc(bal);
        }
    }
    __syncwarp();
    // bins live in shared memory unless there are more than kSmemBins of them (then the global copy is used);
    // the EM's per-bin state overlays the (now all-zero) histogram
    const bool in_smem = nb <= kSmemBins;
    const uint32_t* bins = in_smem ? W.bins : gbins;
    double* lml = in_smem ? reinterpret_cast<double*>(W.hist) : cs.a.lml_spill + (size_t)warp_global * kMaxBins;

    const double dtot = (double)(W.rec.depth[0] + W.rec.depth[1] + W.rec.depth[2] + W.rec.depth[3] + W.rec.depth_other);
    int n_act = __popc(act);
    uint32_t flags = 0;
    double chi = 0.0;
    // initial frequencies of a subset: depth/total for its members, 0 elsewhere (src/basetype.cpp:93-103)
    if (lane < 4) W.emf[lane] = (act >> lane & 1u) ? (double)W.rec.depth[lane] / dtot : 0.0;
    __syncwarp();
    double lr_alt = em_bins(bins, lml, nb, (int)act, dtot);
    if (lane < 4) W.res_f[lane] = W.emf[lane];
    uint32_t em_calls = 1;
#pragma unroll 1
    for (int n = n_act - 1; n > 0; --n) {
        // the n-subsets of the n+1 active bases in the lexicographic order of
        // src/external/combinations.h:19-84: the i-th subset drops the (n-i)-th active base
        double best_chi = 0, best_lr = 0;
        uint32_t best_set = 0;
#pragma unroll 1
        for (int i = 0; i <= n; ++i) {
            const uint32_t sub = act & ~(1u << nth_set_bit(act, n - i));
            __syncwarp();
            if (lane < 4) W.emf[lane] = (sub >> lane & 1u) ? (double)W.rec.depth[lane] / dtot : 0.0;
            __syncwarp();
            if (W.emf[0] + W.emf[1] + W.emf[2] + W.emf[3] == 0) flags |= BV_FLAG_ZERO_SUBSET;   // the reference throws (basetype.cpp:113)
            double lr;
            if (n == 1) {
                const int single = __ffs(sub) - 1;
                lr = single_allele_ll(bins, nb, single);
                const double v = (lr != lr) ? lr : 1.0;
                if (lane < 4) W.emf[lane] = lane == single ? v : 0.0;
                __syncwarp();
            } else {
                lr = em_bins(bins, lml, nb, (int)sub, dtot);
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
                best_chi = c; best_lr = lr; best_set = sub;
                if (lane < 4) W.best_f[lane] = W.emf[lane];
            }
        }
        lr_alt = best_lr;
        chi = best_chi;
        const double lrt_threshold = cs.a.lrt_threshold;
        if (fabs(chi - lrt_threshold) < 1e-9 * lrt_threshold) flags |= BV_FLAG_NEAR_LRT;
        if (chi < lrt_threshold) {
            act = best_set; n_act = n;
            __syncwarp();
            if (lane < 4) W.res_f[lane] = W.best_f[lane];
        } else {
            break;
        }
    }
    // the EM state overlaid the histogram: back to all-zero for the next site
    __syncwarp();
    if (in_smem) {
#pragma unroll 1
        for (int i = lane; i < 2 * nb; i += 32) W.hist[i] = 0;
    }
    if (lane == 0) {
        W.res_chi = chi;
        W.flag_word |= flags;
    }
    __syncwarp();
    return act | ((uint32_t)n_act << 4) | (em_calls << 8);
}

// ---- slow finish: the row has non-reference reads (or REF is not A/C/G/T, or a bad strand code) -------------------------
// In (W.slow_*): the site, reads holding the reference base (all / '-' strand) from pass 1, REF code; the other counted
// cells are in W.nr_cnt.  Everything here is warp-uniform.
__device__ __noinline__ void site_slow() {
    WarpSmem& W = warp_smem();
    const CtaShared& cs = cta_shared();
    const int lane = threadIdx.x & 31;
    const uint32_t site = W.slow_site;
    const int ref_code = (int)W.slow_ref_code;
    const double min_af = cs.a.min_af;
    // ---- depths and strand table ----
    const uint4 c0 = *reinterpret_cast<const uint4*>(&W.nr_cnt[0]);
    const uint4 c1 = *reinterpret_cast<const uint4*>(&W.nr_cnt[4]);
    const uint2 c2 = *reinterpret_cast<const uint2*>(&W.nr_cnt[8]);
    uint32_t f0 = c0.x, r0 = c0.y, f1 = c0.z, r1 = c0.w, f2 = c1.x, r2 = c1.y, f3 = c1.z, r3 = c1.w;
    const uint32_t other = c2.x + c2.y;
    {
        const uint32_t n_rev = W.slow_n_rev, rf = W.slow_n_ref - n_rev;
        if (ref_code == 0) { f0 += rf; r0 += n_rev; }
        if (ref_code == 1) { f1 += rf; r1 += n_rev; }
        if (ref_code == 2) { f2 += rf; r2 += n_rev; }
        if (ref_code == 3) { f3 += rf; r3 += n_rev; }
    }
    const uint32_t bad_strand = W.slow_bad;
    __syncwarp();
    if (lane < 12) W.nr_cnt[lane] = 0;
    const uint32_t d0 = f0 + r0, d1 = f1 + r1, d2 = f2 + r2, d3 = f3 + r3;
    const uint32_t total = d0 + d1 + d2 + d3 + other;
    const double dtot = (double)total;
    if (lane == 0) {
        W.flag_word = bad_strand ? BV_FLAG_BAD_STRAND : 0u;
        bv_site_out& r = W.rec;
        r.depth[0] = d0; r.depth[1] = d1; r.depth[2] = d2; r.depth[3] = d3;
        r.depth_other = other;
        r.reserved0 = 0;
        r.fwd[0] = f0; r.fwd[1] = f1; r.fwd[2] = f2; r.fwd[3] = f3;
        r.rev[0] = r0; r.rev[1] = r1; r.rev[2] = r2; r.rev[3] = r3;
    }

    // ---- lrt (src/basetype.cpp:130-199): active set ----
    uint32_t act = 0;
    if (total > 0) {
        act |= is_active(d0, total, dtot, min_af) ? 1u : 0u;
        act |= is_active(d1, total, dtot, min_af) ? 2u : 0u;
        act |= is_active(d2, total, dtot, min_af) ? 4u : 0u;
        act |= is_active(d3, total, dtot, min_af) ? 8u : 0u;
    }
    int n_act = __popc(act);
    double chi = 0.0;
    uint32_t em_calls = 0;
    const uint32_t ref_bit = ref_code >= 0 ? (1u << ref_code) : 0u;
    __syncwarp();

    bool bounded = false;
    if (n_act == 2 && (act & ref_bit)) {
        // REF plus one minor allele: try to settle the LRT by the bound (see lrt_bound)
        const int o_code = __ffs(act & ~ref_bit) - 1;
        const uint32_t d_ref = sel4u(ref_code, d0, d1, d2, d3), d_o = sel4u(o_code, d0, d1, d2, d3);
        if (d_ref >= 22 && d_o <= 4 && lrt_bound(site, (uint32_t)o_code) < kBoundLimit) {
            bounded = true;
            act = ref_bit; n_act = 1;
            if (lane == 0) W.flag_word |= BV_FLAG_LRT_BOUND;
            if (lane < 4) W.res_f[lane] = 0.0;
        }
    }
    if (bounded) {
        // nothing else to compute: single active allele REF
    } else if (n_act >= 2 || (n_act == 1 && (act & ~ref_bit))) {
        // the result depends on base qualities: histogram the row by (base, phred)
        const uint32_t h = build_hist(site);
        const uint32_t qmin = h & 0xffu, qmax = (h >> 8) & 0xffu;
        if (lane == 0) W.flag_word |= h >> 16;
        __syncwarp();
        if (n_act >= 2) {
            const uint32_t r = lrt_multi(qmin, qmax, act);
            act = r & 0xfu; n_act = (int)((r >> 4) & 0xfu); em_calls = r >> 8;
            chi = W.res_chi;
        } else {
            // One active allele: the reference still runs one EM; its answer is closed form (see single_allele_ll):
            // AF == 1.0 exactly, or NaN when a phred-0 read of that base exists.
            const int b = __ffs(act) - 1;
            const bool bad = qmin == 0 && W.hist[b * kQSlots] != 0;
            const double v = bad ? __longlong_as_double(0x7ff8000000000000ll) : 1.0;
            __syncwarp();
            if (lane < 4) W.res_f[lane] = lane == b ? v : 0.0;
            em_calls = 1;
            if (qmin <= qmax) {   // histogram back to zero
#pragma unroll 1
                for (uint32_t q = qmin + lane; q <= qmax; q += 32) {
#pragma unroll
                    for (int r = 0; r < 5; ++r) W.hist[r * kQSlots + q] = 0;
                }
            }
        }
    } else {
        if (lane < 4) W.res_f[lane] = 0.0;
        if (n_act == 1) em_calls = 1;   // the single active allele is REF: AF is not reported
    }
    __syncwarp();
    uint32_t flags = W.flag_word;

    // ---- ALT / QUAL (src/basetype.cpp:170-196) ----
    const uint32_t alt_set = act & ~ref_bit;
    const int n_alt = __popc(alt_set);
    double qual = 0.0;
    if (n_alt) {
        const int first_act = __ffs(act) - 1;
        const double r = (double)sel4u(first_act, d0, d1, d2, d3) / dtot;
        if (n_act == 1 && total > 10 && r > 0.5) { qual = 5000.0; flags |= BV_FLAG_MONO_QUAL; }
        else qual = qual_from_chi(chi);
    }

    // ---- strand bias (src/basetype.cpp:244-295): CVG row = ref vs all non-ref ACGT; VCF row = ref vs ALT ----
    double fs_cvg = 0.0, fs_vcf = 0.0;
    {
        const int rf = ref_code < 0 ? 0 : (int)sel4u(ref_code, f0, f1, f2, f3);
        const int rr = ref_code < 0 ? 0 : (int)sel4u(ref_code, r0, r1, r2, r3);
        const int af_ = (int)(f0 + f1 + f2 + f3) - rf, ar = (int)(r0 + r1 + r2 + r3) - rr;
        // a table with an empty row or column has a single possible outcome: p == 1, FS == 0 (kfunc.c:256)
        if ((af_ | ar) != 0 && (rf | rr) != 0) fs_cvg = fs_from_table(cs.a.logfact, rf, rr, af_, ar);
        if (n_alt) {
            const int vf = (int)(((alt_set & 1) ? f0 : 0u) + ((alt_set & 2) ? f1 : 0u) + ((alt_set & 4) ? f2 : 0u) + ((alt_set & 8) ? f3 : 0u));
            const int vr = (int)(((alt_set & 1) ? r0 : 0u) + ((alt_set & 2) ? r1 : 0u) + ((alt_set & 4) ? r2 : 0u) + ((alt_set & 8) ? r3 : 0u));
            if (vf == af_ && vr == ar) fs_vcf = fs_cvg;   // same 2x2 table
            else if ((vf | vr) != 0 && (rf | rr) != 0) fs_vcf = fs_from_table(cs.a.logfact, rf, rr, vf, vr);
        }
    }

    // ---- record ----
    if (lane == 0) {
        bv_site_out& r = W.rec;
        // ALT alleles in ACGT order of the active set (src/basetype.cpp:172-177)
        uint32_t alts = 0;
        int k = 0;
        double af[4] = {W.res_f[0], W.res_f[1], W.res_f[2], W.res_f[3]};
        r.af[0] = 0.0; r.af[1] = 0.0; r.af[2] = 0.0; r.af[3] = 0.0;
        if (alt_set & 1) { r.af[k] = af[0]; alts |= 0u << (8 * k); ++k; }
        if (alt_set & 2) { r.af[k] = af[1]; alts |= 1u << (8 * k); ++k; }
        if (alt_set & 4) { r.af[k] = af[2]; alts |= 2u << (8 * k); ++k; }
        if (alt_set & 8) { r.af[k] = af[3]; alts |= 3u << (8 * k); ++k; }
        r.n_alt = (uint8_t)n_alt;
        r.alt[0] = (uint8_t)alts; r.alt[1] = (uint8_t)(alts >> 8); r.alt[2] = (uint8_t)(alts >> 16); r.alt[3] = (uint8_t)(alts >> 24);
        r.n_active = (uint8_t)n_act;
        r.flags = (uint8_t)flags;
        r.em_calls = (uint8_t)(em_calls > 255u ? 255u : em_calls);
        r.qual = qual;
        r.chi2 = chi;
        r.fs_cvg = fs_cvg;
        r.fs_vcf = fs_vcf;
    }
    __syncwarp();
    if (lane < 8) {
        const uint4* src = reinterpret_cast<const uint4*>(&W.rec);
        uint4* dst = reinterpret_cast<uint4*>(cs.a.out + site);
        dst[lane] = src[lane];
    }
    __syncwarp();
}

// =====================================================================================================================
// The streaming loop.  Call-free on purpose: with a call inside, the ABI makes the compiler keep the loop-carried state
// in local memory for the whole loop (measured: 4.9 GB of spill traffic per 1e9 cells).  It runs from the state saved
// in W.sv_* until it meets a site for the slow path, saves its state, describes the site in W.slow_* and returns 1;
// returns 0 when the warp's sites are exhausted.
// =====================================================================================================================
__device__ __noinline__ uint32_t stream_sites() {
    WarpSmem& W = warp_smem();
    const CtaShared& cs = cta_shared();
    const int lane = threadIdx.x & 31;
    const uint32_t N = cs.a.n_samples, n_sites = cs.a.n_sites;
    const uint64_t pitch = cs.a.pitch;
    const uint8_t* const g_base = cs.a.base;
    const uint8_t* const g_strand = cs.a.strand;
    const uint8_t* const g_ref = cs.a.ref_base;
    uint32_t* const g_out = reinterpret_cast<uint32_t*>(cs.a.out);
    const uint32_t row_bytes = (N + 15u) & ~15u;   // bytes of a row that hold cells
    const uint32_t total_warps = gridDim.x * kWarps;
    const uint32_t one_active = (1.0 >= cs.a.min_af) ? 1u : 0u;   // a site whose reads all agree has that allele active
    const uint32_t sW = smem_u32(&W);
    const uint32_t s_full0 = smem_u32(&W.full[0]);

    uint32_t p_site = W.sv_p_site, p_off = W.sv_p_off, p_slot = W.sv_p_slot;      // producer cursor (lane 0 issues)
    uint32_t c_slot = W.sv_c_slot, c_par = W.sv_c_par, site = W.sv_site, ref_raw = W.sv_ref_raw;

#pragma unroll 1
    for (; site < n_sites; site += total_warps) {
        // reference base of this site (fetched one site ahead), toupper (src/basetype.cpp:171)
        const uint32_t next_site = site + total_warps;
        const uint32_t ref_next = next_site < n_sites ? (uint32_t)__ldg(g_ref + next_site) : 0u;
        uint32_t rc = ref_raw;
        if (rc >= 'a' && rc <= 'z') rc -= 32;
        const int ref_code = rc == 'A' ? 0 : rc == 'C' ? 1 : rc == 'G' ? 2 : rc == 'T' ? 3 : -1;
        const uint32_t refw = ref_code >= 0 ? (uint32_t)ref_code * 0x01010101u : 0x08080808u;

        // ---- pass 1 ----
        ScanAcc A;
        A.nref = 0; A.nrev = 0; A.nonref = 0; A.bad = 0;
#pragma unroll 1
        for (uint32_t c_off = 0; c_off < row_bytes; c_off += kChunk) {
            // the unit kStages-1 ahead goes into the slot the previous unit used; every lane is past it (__syncwarp)
            if (p_site < n_sites) {
                if (lane == 0) {
                    const uint32_t bytes = min((uint32_t)kChunk, row_bytes - p_off);
                    const size_t g = (size_t)p_site * pitch + p_off;
                    const uint32_t bar = s_full0 + 8u * p_slot, dst = sW + (uint32_t)sizeof(Stage) * p_slot;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst), "l"(g_base + g), "r"(bytes), "r"(bar) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst + (uint32_t)kChunk), "l"(g_strand + g), "r"(bytes), "r"(bar) : "memory");
                }
                p_slot = (p_slot + 1 == kStages) ? 0 : p_slot + 1;
                p_off += kChunk;
                if (p_off >= row_bytes) { p_off = 0; p_site += total_warps; }
            }
            {   // wait for this unit's bytes
                const uint32_t bar = s_full0 + 8u * c_slot;
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    "WAIT_%=:\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                    "@p bra DONE_%=;\n\t"
                    "bra WAIT_%=;\n\t"
                    "DONE_%=:\n\t"
                    "}" ::"r"(bar), "r"(c_par) : "memory");
            }
#pragma unroll
            for (int v = 0; v < kChunk / 512; ++v) {
                const int lane_cells = (int)N - (int)c_off - v * 512 - lane * 16;
                if (lane_cells > 0) {
                    const uint8_t* cellp = W.stage[c_slot].base + v * 512 + lane * 16;
                    uint4 vb = *reinterpret_cast<const uint4*>(cellp);
                    const uint4 vs = *reinterpret_cast<const uint4*>(cellp + kChunk);
                    if (lane_cells < 16) {   // the row's last, partial vector: one lane, once per row
                        const uint4 k = *reinterpret_cast<const uint4*>(cs.tail_keep);
                        vb.x = (vb.x & k.x) | (0x05050505u & ~k.x); vb.y = (vb.y & k.y) | (0x05050505u & ~k.y);
                        vb.z = (vb.z & k.z) | (0x05050505u & ~k.z); vb.w = (vb.w & k.w) | (0x05050505u & ~k.w);
                    }
                    const uint32_t nr0 = scan_word(vb.x, vs.x, refw, A);
                    const uint32_t nr1 = scan_word(vb.y, vs.y, refw, A);
                    const uint32_t nr2 = scan_word(vb.z, vs.z, refw, A);
                    const uint32_t nr3 = scan_word(vb.w, vs.w, refw, A);
                    if (nr0 | nr1 | nr2 | nr3) {
                        // counted cells that are not the reference base (sequencing errors, ALT alleles): one by one
                        uint32_t t = (nr0 >> 7) | (nr1 >> 6) | (nr2 >> 5) | (nr3 >> 4);   // bit (8*byte + word)
                        A.nonref |= t;
                        do {
                            int top;
                            asm("bfind.u32 %0, %1;" : "=r"(top) : "r"(t));
                            t ^= 1u << top;
                            const int cell = ((top & 3) << 2) | (top >> 3);
                            const uint32_t b = cellp[cell];
                            const uint32_t s = cellp[cell + kChunk];
                            atomicAdd(&W.nr_cnt[2u * b + (s & 1u)], 1u);
                        } while (t);
                    }
                }
            }
            __syncwarp();
            if (++c_slot == kStages) { c_slot = 0; c_par ^= 1u; }
        }

        // ---- finish ----
        const uint32_t fl = __reduce_or_sync(kFull, A.nonref | (A.bad ? 0x80000000u : 0u));
        const uint32_t n_ref = __reduce_add_sync(kFull, A.nref) >> 7;
        const uint32_t n_rev = __reduce_add_sync(kFull, A.nrev) >> 7;
        if (fl != 0) {
            if (lane == 0) {
                W.sv_p_site = p_site; W.sv_p_off = p_off; W.sv_p_slot = p_slot; W.sv_c_slot = c_slot; W.sv_c_par = c_par;
                W.sv_site = next_site; W.sv_ref_raw = ref_next;
                W.slow_site = site; W.slow_n_ref = n_ref; W.slow_n_rev = n_rev; W.slow_bad = fl >> 31;
                W.slow_ref_code = (uint32_t)ref_code;
            }
            __syncwarp();
            return 1u;
        }
        // Every counted cell holds the reference base (or nothing is covered): depth[REF] = n_ref, one active
        // allele == REF, the EM's answer is f = 1 (src/algorithm.h:210-255 with a single column), no ALT,
        // QUAL / FS / chi2 = 0.  Lanes compose the 32 words of the record.
        const uint32_t n_active = (n_ref > 0) ? one_active : 0u;
        uint32_t w = 0;
        if (lane == ref_code) w = n_ref;                          // depth[REF]
        if (lane == ref_code + 6) w = n_ref - n_rev;              // fwd[REF]
        if (lane == ref_code + 10) w = n_rev;                     // rev[REF]
        if (lane == 15) w = (n_active << 8) | (n_active << 24);   // n_active | flags 0 | em_calls
        g_out[(size_t)site * 32 + lane] = w;
        ref_raw = ref_next;
    }
    return 0u;
}

}  // namespace bv
