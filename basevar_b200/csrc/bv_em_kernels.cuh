// bv_em_kernels.cuh -- K4: the sites whose result depends on base qualities (state kStateEM), as three kernels.
//
// The reference runs, per such site, one EM over the full active set and then backward elimination: for every
// (n-1)-subset of the n active alleles another EM, keep the best, stop when 2 dLL >= 24 (BaseType::lrt,
// src/basetype.cpp:130-199; _f, :105-128; EM, src/algorithm.h:210-255).  The EMs of one site are independent of each
// other -- only WHICH of them the elimination consults depends on earlier results -- and one EM is a short serial
// recurrence over the site's non-empty (base, phred) bins.  So the work is split by its shape:
//
//   K4a bv_hist_kernel     one WARP per site: streams the row's base + qual planes (TMA ring), (base, phred) histogram in
//                          shared memory, compaction to bins.  A site with one active allele is finished here (closed
//                          form).  A site with >= 2 gets a header, its bins in a pool, and one EM TASK per subset of
//                          >= 2 active alleles (1, 4 or 11 tasks for 2, 3 or 4 active alleles; single-allele models are
//                          closed form and travel in the header).
//   K4b bv_em_task_kernel  one THREAD per task: the whole EM of one candidate subset over the site's bins, staged in
//                          shared memory.  No shuffles, no warp-wide repetition of scalar work; every lane of the FP64
//                          pipe carries a different EM.  Tasks of all sites are one flat list, so a warp's 32 EMs have
//                          similar lengths whatever the sites' allele counts.
//   K4c bv_decide_kernel   one THREAD per site: replays the elimination loop on the table of task results (first-minimum
//                          argmin in the reference's subset order, threshold, flags), ALT / AF / QUAL
//                          (chi-square survival function) and the strand-bias Fisher test of the VCF row.
//
// At most 3 of the 11 tasks of a 4-allele site are never consulted (the elimination needs <= 8 EMs); evaluating them
// anyway removes every dependency between EMs.  A flag an unconsulted task raises (BV_FLAG_EM_MAXITER) is not reported.
// Scratch (headers, bin pool, task list) is sized per tile; a site that does not fit any more is finished inside K4a by
// the warp-per-site code K6 also uses (lrt_on_bins), so that the records never depend on the pool size.
#pragma once
#include "bv_finish_kernels.cuh"

namespace bv {

// bit m of the result: subset m (bit j = allele j) of `act` has >= 2 members, i.e. is an EM task of the site
__device__ __forceinline__ uint32_t em_task_mask(uint32_t act) {
    uint32_t v = 0;
#pragma unroll
    for (uint32_t m = 3; m < 16; ++m)
        if ((m & ~act) == 0 && (m & (m - 1)) != 0) v |= 1u << m;
    return v;
}

// ---- K4a: one site in state kStateEM -----------------------------------------------------------------------------------------
// Everything here is warp-uniform.  The record (counts, FS of the CVG row, flags) comes from K1 / K2.
__device__ __noinline__ void hist_site(uint32_t site) {
    QualWarp& W = warp_smem();
    const QualCta& cs = cta_shared();
    const int lane = threadIdx.x & 31;
    if (lane < 8) reinterpret_cast<uint4*>(&W.rec)[lane] = reinterpret_cast<const uint4*>(cs.a.out + site)[lane];
    const int ref_code = ref_code_of(cs.a.ref_base[site]);
    __syncwarp();
    const uint32_t d0 = W.rec.depth[0], d1 = W.rec.depth[1], d2 = W.rec.depth[2], d3 = W.rec.depth[3];
    const uint32_t total = d0 + d1 + d2 + d3 + W.rec.depth_other;
    const double dtot = (double)total;
    const double min_af = cs.a.min_af;
    if (lane == 0) W.flag_word = W.rec.flags;

    // ---- lrt (src/basetype.cpp:130-199): active set (total > 0 here) ----
    uint32_t act = 0;
    act |= is_active(d0, total, dtot, min_af) ? 1u : 0u;
    act |= is_active(d1, total, dtot, min_af) ? 2u : 0u;
    act |= is_active(d2, total, dtot, min_af) ? 4u : 0u;
    act |= is_active(d3, total, dtot, min_af) ? 8u : 0u;
    int n_act = __popc(act);
    double chi = 0.0;
    uint32_t em_calls = 0;
    const uint32_t ref_bit = ref_code >= 0 ? (1u << ref_code) : 0u;
    __syncwarp();

    // histogram the row by (base, phred)
    const uint32_t h = build_hist(site, nullptr, 0);
    const uint32_t qmin = h & 0xffu, qmax = (h >> 8) & 0xffu;
    if (lane == 0) W.flag_word |= h >> 16;
    __syncwarp();
    if (n_act >= 2) {
        const int nb = compact_bins(qmin, qmax);
        // ---- hand the site to K4b / K4c: bins into the pool, one task per subset of >= 2 active alleles ----
        const uint32_t tmask = em_task_mask(act);
        const uint32_t n_tasks = (uint32_t)__popc(tmask);
        uint32_t off = 0, t0 = 0, hdr = 0, ok = 0;
        if (lane == 0) {
            off = atomicAdd(cs.a.counters + kCntEmPool, (uint32_t)nb);
            if (off <= cs.a.em_pool_cap && (uint32_t)nb <= cs.a.em_pool_cap - off) {
                t0 = atomicAdd(cs.a.counters + kCntEmTask, n_tasks);
                if (t0 <= cs.a.em_task_cap && n_tasks <= cs.a.em_task_cap - t0) {
                    hdr = atomicAdd(cs.a.counters + kCntEmHdr, 1u);
                    ok = 1;
                } else {
                    ok = 2;   // task slots [t0, cap) stay unused: marked invalid below
                }
            }
        }
        ok = __shfl_sync(kFull, ok, 0); off = __shfl_sync(kFull, off, 0);
        t0 = __shfl_sync(kFull, t0, 0); hdr = __shfl_sync(kFull, hdr, 0);
        if (ok == 2) {
            for (uint32_t t = t0 + lane; t < cs.a.em_task_cap && t < t0 + n_tasks; t += 32) cs.a.em_tasks[t] = kEmTaskInvalid;
        }
        if (ok == 1) {
            const uint32_t* bins = nb <= kSmemBins ? W.bins : cs.a.bin_spill + (size_t)(blockIdx.x * kQualWarps + (threadIdx.x >> 5)) * kMaxBins;
            for (int i = lane; i < nb; i += 32) cs.a.em_pool[off + i] = bins[i];
            // single-allele models of the active alleles: closed form (see single_allele_ll)
            double sll = 0.0;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                if (!(act >> b & 1u)) continue;
                const double v = single_allele_ll(bins, nb, b);
                if (lane == b) sll = v;
            }
            // task m-th subset (ascending mask value) -> slot t0 + rank
            if (lane < 16 && (tmask >> lane & 1u))
                cs.a.em_tasks[t0 + __popc(tmask & ((1u << lane) - 1u))] = hdr | ((uint32_t)lane << 28);
            EmSiteHdr* H = cs.a.em_hdr + hdr;
            if (lane < 4) H->single_ll[lane] = sll;
            if (lane == 0) {
                uint4 w0, w1, w2;
                w0.x = site; w0.y = off; w0.z = (uint32_t)nb; w0.w = t0;
                w1.x = act; w1.y = W.flag_word; w1.z = d0; w1.w = d1;
                w2.x = d2; w2.y = d3; w2.z = total; w2.w = 0;
                reinterpret_cast<uint4*>(H)[0] = w0; reinterpret_cast<uint4*>(H)[1] = w1; reinterpret_cast<uint4*>(H)[2] = w2;
            }
            __syncwarp();
            return;   // the record is completed by K4c
        }
        // scratch pools full: finish the site here
        if (lane == 0) atomicAdd(cs.a.counters + kCntEmFallback, 1u);
        const uint32_t r = lrt_on_bins(nb, act, kOrderACGT);
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
    __syncwarp();
    uint32_t flags = W.flag_word;

    // ---- ALT / QUAL (src/basetype.cpp:170-196) ----
    // Here only the rule that needs no arithmetic (mono-allelic 5000); the chi-square survival function and the Fisher
    // test of the VCF row are scalar work: the site is queued and vcf_flush() does them one thread per site.
    const uint32_t alt_set = act & ~ref_bit;
    const int n_alt = __popc(alt_set);
    double qual = 0.0;
    if (n_alt) {
        const int first_act = __ffs(act) - 1;
        const double r = (double)sel4u(first_act, d0, d1, d2, d3) / dtot;
        if (n_act == 1 && total > 10 && r > 0.5) { qual = 5000.0; flags |= BV_FLAG_MONO_QUAL; }
    }

    // ---- record ----
    __syncwarp();
    if (lane == 0) {
        bv_site_out& r = W.rec;
        r.reserved0 = kStateDone;
        // ALT alleles in ACGT order of the active set (src/basetype.cpp:172-177)
        uint32_t alts = 0;
        int k = 0;
        const double af[4] = {W.res_f[0], W.res_f[1], W.res_f[2], W.res_f[3]};
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
        r.fs_vcf = 0.0;
        // called sites go on to the rank-sum / population-group kernels (bv_call_kernels.cuh)
        if (n_alt && cs.a.list_called) cs.a.list_called[atomicAdd(cs.a.counters + kCntCalled, 1u)] = site;
        if (n_alt) W.vcf_site[W.vcf_n++] = site;
    }
    __syncwarp();
    if (lane < 8) reinterpret_cast<uint4*>(cs.a.out + site)[lane] = reinterpret_cast<const uint4*>(&W.rec)[lane];
    __syncwarp();   // also orders the record's stores before vcf_flush() reads them from other lanes
    if (W.vcf_n == 32) vcf_flush();
}

// Persistent warps with dynamic work distribution over the EM list.
__global__ void __launch_bounds__(kQualWarps * 32, 1) bv_hist_kernel(const __grid_constant__ SiteKernelArgs a) {
    QualCta& cs = cta_shared();
    QualWarp& W = warp_smem();
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4 * kQStride; i += blockDim.x) cs.lut[i] = a.lut[i];
    if (threadIdx.x == 0) cs.a = a;
    for (int i = lane; i < kHistWords; i += 32) W.hist[i] = 0;
    if (lane == 0) {
        W.flag_word = 0;
        W.p2_phase = 0;
        W.vcf_n = 0;
        for (int b = 0; b < kP2Bufs; ++b) mbar_init(&W.p2bar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // the EM list is final (K3 is done); warps take one site at a time: the cost per site varies by an order of magnitude
    const uint32_t n_em = a.counters[kCntEm];
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(a.counters + kCntEmNext, 1u);
        i = __shfl_sync(kFull, i, 0);
        if (i >= n_em) break;
        hist_site(a.list_em[i]);
    }
    if (W.vcf_n) vcf_flush();
}

// =====================================================================================================================
// K4b: one thread per EM task.
// =====================================================================================================================
#ifndef BV_TASK_THREADS
#define BV_TASK_THREADS 128
#endif
#ifndef BV_TASK_STAGE_BINS
#define BV_TASK_STAGE_BINS 192
#endif
constexpr int kTaskThreads = BV_TASK_THREADS;
constexpr int kStageBins = BV_TASK_STAGE_BINS;     // bins per site staged in shared memory; longer lists are read from the pool
constexpr int kStageStride = kStageBins + 1;       // odd: the rows of 32 different sites start in 32 different banks

struct __align__(16) TaskCta {
    double ome[kQSlots];         // 1 - eps(q)
    double e3[kQSlots];          // eps(q) / 3
    uint32_t hdr_of[kTaskThreads];    // header index of thread t's task
    uint32_t slot_hdr[kTaskThreads];  // header index of staging row s
    uint32_t warp_leaders[kTaskThreads / 32];
    uint32_t bins[kTaskThreads * kStageStride];
};
constexpr size_t kTaskSmemBytes = sizeof(TaskCta);
static_assert(kTaskSmemBytes <= 232448, "shared memory of the EM task kernel exceeds 227 KB");

// One E-step + M-step over the bins (src/algorithm.h:148-198) under frequencies f; WITH_PREV: also the marginals under
// the previous frequencies fp, for the convergence test of EM() (src/algorithm.h:238-250) -- recomputed rather than
// kept per bin, so that a task needs no per-bin state.  Same operation order as the reference: lik * freq summed in
// A, C, G, T order (alleles outside the subset have freq 0 and add an exact +0.0), column sums of the posteriors with
// c equal reads adding c * post; posteriors are l_j * (1 / m) (one division per bin, inside the stated tolerance).
template <bool WITH_PREV>
__device__ __forceinline__ void em_pass(const uint32_t* bins, int nb, const double* s_ome, const double* s_e3,
                                        const double (&f)[4], const double (&fp)[4], bool int_mode,
                                        double (&s)[4], bool& big, double& delta, double* prev_log) {
    s[0] = 0; s[1] = 0; s[2] = 0; s[3] = 0;
#pragma unroll 2
    for (int i = 0; i < nb; ++i) {
        const uint32_t p = bins[i];
        const uint32_t b = bin_base(p), q = bin_qual(p);
        const double cd = (double)bin_count(p);
        const double ome = s_ome[q], e3 = s_e3[q];
        const double L0 = b == 0 ? ome : e3, L1 = b == 1 ? ome : e3, L2 = b == 2 ? ome : e3, L3 = b == 3 ? ome : e3;
        const double l0 = L0 * f[0], l1 = L1 * f[1], l2 = L2 * f[2], l3 = L3 * f[3];
        double m = 0.0;
        m += l0; m += l1; m += l2; m += l3;
        const double inv = 1.0 / m;
        s[0] += cd * (l0 * inv); s[1] += cd * (l1 * inv); s[2] += cd * (l2 * inv); s[3] += cd * (l3 * inv);
        if (WITH_PREV) {
            if (int_mode) {
                double mp = 0.0;
                mp += L0 * fp[0]; mp += L1 * fp[1]; mp += L2 * fp[2]; mp += L3 * fp[3];
                // (double)abs((int)diff) is non-zero iff |log m - log mp| >= 1, i.e. the marginal moved by a factor e:
                // inside (1/2.5, 2.5) the ratio decides without a logarithm, otherwise the logarithms themselves do.
                // NaN / inf convert to INT_MIN whose "abs" stays negative and ends the loop (results are NaN by then).
                if (!(m < 2.5 * mp && mp < 2.5 * m)) {
                    const double diff = nlog(m) - nlog(mp);
                    if (fabs(diff) >= 1.0 && fabs(diff) < 2147483648.0) big = true;
                }
            } else {
                const double llh = nlog(m);
                double lp;
                if (prev_log) { lp = prev_log[i]; prev_log[i] = llh; }
                else {
                    double mp = 0.0;
                    mp += L0 * fp[0]; mp += L1 * fp[1]; mp += L2 * fp[2]; mp += L3 * fp[3];
                    lp = nlog(mp);
                }
                delta += cd * fabs(llh - lp);
            }
        } else if (prev_log) {
            prev_log[i] = nlog(m);
        }
    }
}

__device__ __noinline__ void em_task(const SiteKernelArgs& a, const TaskCta& cs, const EmSiteHdr& H, const uint32_t* bins,
                                     uint32_t subset, double* res) {
    const int nb = (int)H.nb;
    const double total = (double)H.total;
    const bool int_mode = a.abs_mode == BV_EM_ABS_INT_TRUNC;
    const bool in[4] = {(subset & 1u) != 0, (subset & 2u) != 0, (subset & 4u) != 0, (subset & 8u) != 0};
    // initial frequencies: depth/total for the subset's members, 0 elsewhere, NOT renormalised (src/basetype.cpp:93-103)
    double f[4], fp[4], s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[j] = in[j] ? (double)H.depth[j] / total : 0.0; fp[j] = 0.0; }
    double prev_buf[kStageBins];   // double mode: log marginals of the previous E-step (local memory, lanes interleaved)
    double* prev_log = (!int_mode && nb <= kStageBins) ? prev_buf : nullptr;
    bool big = false;
    double delta = 0.0;
    uint64_t flags = 0;
    em_pass<false>(bins, nb, cs.ome, cs.e3, f, fp, int_mode, s, big, delta, prev_log);
#pragma unroll
    for (int j = 0; j < 4; ++j) { fp[j] = f[j]; f[j] = in[j] ? s[j] / total : 0.0; }
    int it = a.em_max_iter;
    for (;;) {
        big = false; delta = 0.0;
        em_pass<true>(bins, nb, cs.ome, cs.e3, f, fp, int_mode, s, big, delta, prev_log);
#pragma unroll
        for (int j = 0; j < 4; ++j) { fp[j] = f[j]; f[j] = in[j] ? s[j] / total : 0.0; }
        const bool more = int_mode ? big : !(delta < a.em_eps);
        --it;
        if (it == 0) flags |= BV_FLAG_EM_MAXITER;
        if (!more || it == 0) break;
    }
    // log marginal likelihoods under the second-to-last frequencies (those of the last E-step), summed: what _f() adds up
    // (src/basetype.cpp:119-120)
    double ll = 0.0;
#pragma unroll 2
    for (int i = 0; i < nb; ++i) {
        const uint32_t p = bins[i];
        const uint32_t b = bin_base(p), q = bin_qual(p);
        const double ome = cs.ome[q], e3 = cs.e3[q];
        double lml;
        if (prev_log) lml = prev_log[i];
        else {
            double mp = 0.0;
            mp += (b == 0 ? ome : e3) * fp[0]; mp += (b == 1 ? ome : e3) * fp[1];
            mp += (b == 2 ? ome : e3) * fp[2]; mp += (b == 3 ? ome : e3) * fp[3];
            lml = nlog(mp);
        }
        ll += (double)bin_count(p) * lml;
    }
    res[0] = ll; res[1] = f[0]; res[2] = f[1]; res[3] = f[2]; res[4] = f[3];
    res[5] = __longlong_as_double((long long)flags);
}

__global__ void __launch_bounds__(kTaskThreads) bv_em_task_kernel(const __grid_constant__ SiteKernelArgs a) {
    TaskCta& cs = *reinterpret_cast<TaskCta*>(bv_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int q = tid; q < kQSlots; q += kTaskThreads) {
        cs.ome[q] = a.lut[kLutOneMinusEps * kQStride + q];
        cs.e3[q] = a.lut[kLutEpsThird * kQStride + q];
    }
    const uint32_t n_tasks = min(a.counters[kCntEmTask], a.em_task_cap);
#pragma unroll 1
    for (uint32_t t0 = blockIdx.x * kTaskThreads; t0 < n_tasks; t0 += gridDim.x * kTaskThreads) {
        const uint32_t t = t0 + tid;
        const uint32_t word = t < n_tasks ? a.em_tasks[t] : kEmTaskInvalid;
        const bool valid = word != kEmTaskInvalid;
        const uint32_t hdr = word & 0x0fffffffu, subset = word >> 28;
        __syncthreads();   // the previous round's readers of cs.bins / cs.hdr_of are done
        cs.hdr_of[tid] = valid ? hdr : kEmTaskInvalid;
        __syncthreads();
        // staging rows: one per run of equal headers (the tasks of a site are consecutive)
        const bool leader = valid && (tid == 0 || cs.hdr_of[tid - 1] != hdr);
        const uint32_t bal = __ballot_sync(kFull, leader);
        if (lane == 0) cs.warp_leaders[warp] = (uint32_t)__popc(bal);
        __syncthreads();
        uint32_t row = (uint32_t)__popc(bal & ((2u << lane) - 1u)) - 1u;   // leaders up to and including this thread, minus one
        for (int w = 0; w < warp; ++w) row += cs.warp_leaders[w];
        // (a non-leader's row is the row of the last leader before it; a CTA's first valid thread is always a leader)
        if (leader) cs.slot_hdr[row] = hdr;
        uint32_t n_rows = 0;
        for (int w = 0; w < kTaskThreads / 32; ++w) n_rows += cs.warp_leaders[w];
        __syncthreads();
        for (uint32_t r = warp; r < n_rows; r += kTaskThreads / 32) {
            const EmSiteHdr& H = a.em_hdr[cs.slot_hdr[r]];
            const uint32_t nb = H.nb, off = H.bins_off;
            if (nb <= (uint32_t)kStageBins)
                for (uint32_t i = lane; i < nb; i += 32) cs.bins[r * kStageStride + i] = a.em_pool[off + i];
        }
        __syncthreads();
        if (valid) {
            const EmSiteHdr& H = a.em_hdr[hdr];
            const uint32_t* bins = H.nb <= (uint32_t)kStageBins ? cs.bins + row * kStageStride : a.em_pool + H.bins_off;
            em_task(a, cs, H, bins, subset, a.em_res + (size_t)t * kEmResDoubles);
        }
    }
}

// =====================================================================================================================
// K4c: one thread per EM site: backward elimination on the task results, ALT / AF / QUAL, FS of the VCF row.
// =====================================================================================================================
__device__ __noinline__ void decide_site(const SiteKernelArgs& a, const EmSiteHdr& H) {
    const uint32_t site = H.site;
    bv_site_out* rec = a.out + site;
    const int ref_code = ref_code_of(a.ref_base[site]);
    const uint32_t dep[4] = {H.depth[0], H.depth[1], H.depth[2], H.depth[3]};
    const uint32_t total = H.total;
    const double dtot = (double)total;
    uint32_t act = H.act & 0xfu;
    int n_act = __popc(act);
    uint32_t flags = H.flags;
    const uint32_t tmask = em_task_mask(act);
    auto result = [&](uint32_t sub) { return a.em_res + (size_t)(H.task0 + (uint32_t)__popc(tmask & ((1u << sub) - 1u))) * kEmResDoubles; };

    // (src/basetype.cpp:144-168) full model, then backward elimination
    const double* r0 = result(act);
    double lr_alt = r0[0];
    double res_f[4] = {r0[1], r0[2], r0[3], r0[4]};
    flags |= (uint32_t)__double_as_longlong(r0[5]);
    double chi = 0.0;
    uint32_t em_calls = 1;
#pragma unroll 1
    for (int n = n_act - 1; n > 0; --n) {
        // the n-subsets of the n+1 active bases in the lexicographic order of
        // src/external/combinations.h:19-84: the i-th subset drops the (n-i)-th active base
        double best_chi = 0, best_lr = 0, best_f[4] = {0, 0, 0, 0};
        uint32_t best_set = 0;
#pragma unroll 1
        for (int i = 0; i <= n; ++i) {
            const uint32_t sub = act & ~(1u << nth_active(kOrderACGT, act, n - i));
            double f0sum = 0.0;   // the subset's initial frequencies, summed in A,C,G,T order
#pragma unroll
            for (int j = 0; j < 4; ++j) f0sum += (sub >> j & 1u) ? (double)dep[j] / dtot : 0.0;
            if (f0sum == 0) flags |= BV_FLAG_ZERO_SUBSET;   // the reference throws (basetype.cpp:113)
            double lr, f[4] = {0, 0, 0, 0};
            if (n == 1) {
                const int single = __ffs(sub) - 1;
                lr = H.single_ll[single];
                f[single] = (lr != lr) ? lr : 1.0;
            } else {
                const double* r = result(sub);
                lr = r[0]; f[0] = r[1]; f[1] = r[2]; f[2] = r[3]; f[3] = r[4];
                flags |= (uint32_t)__double_as_longlong(r[5]);
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
                best_f[0] = f[0]; best_f[1] = f[1]; best_f[2] = f[2]; best_f[3] = f[3];
            }
        }
        lr_alt = best_lr;
        chi = best_chi;
        const double lrt_threshold = a.lrt_threshold;
        if (fabs(chi - lrt_threshold) < 1e-9 * lrt_threshold) flags |= BV_FLAG_NEAR_LRT;
        if (chi < lrt_threshold) {
            act = best_set; n_act = n;
            res_f[0] = best_f[0]; res_f[1] = best_f[1]; res_f[2] = best_f[2]; res_f[3] = best_f[3];
        } else {
            break;
        }
    }

    // ---- ALT / QUAL (src/basetype.cpp:170-196) ----
    const uint32_t ref_bit = ref_code >= 0 ? (1u << ref_code) : 0u;
    const uint32_t alt_set = act & ~ref_bit;
    const int n_alt = __popc(alt_set);
    double qual = 0.0;
    if (n_alt) {
        const int first_act = __ffs(act) - 1;
        const double r = (double)dep[first_act] / dtot;
        if (n_act == 1 && total > 10 && r > 0.5) { qual = 5000.0; flags |= BV_FLAG_MONO_QUAL; }
        else qual = qual_from_chi(chi);
    }
    // ---- strand bias of the VCF row, ref vs the called ALT alleles (src/basetype.cpp:244-295, basetype_caller.cpp:1164) ----
    double fs_vcf = 0.0;
    if (n_alt) {
        const uint32_t f[4] = {rec->fwd[0], rec->fwd[1], rec->fwd[2], rec->fwd[3]};
        const uint32_t rv[4] = {rec->rev[0], rec->rev[1], rec->rev[2], rec->rev[3]};
        int rf = 0, rr = 0, vf = 0, vr = 0, af_ = 0, ar = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b == ref_code) { rf = (int)f[b]; rr = (int)rv[b]; }
            else { af_ += (int)f[b]; ar += (int)rv[b]; }
            if (alt_set >> b & 1u) { vf += (int)f[b]; vr += (int)rv[b]; }
        }
        if (vf == af_ && vr == ar) fs_vcf = rec->fs_cvg;   // same 2x2 table as the CVG row
        else if ((vf | vr) != 0 && (rf | rr) != 0) fs_vcf = fs_from_table(a.logfact, rf, rr, vf, vr);
    }
    // ---- record: ALT alleles in ACGT order of the active set (src/basetype.cpp:172-177) ----
    uint32_t alts = 0;
    int k = 0;
    double af[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (alt_set >> b & 1u) { af[k] = res_f[b]; alts |= (uint32_t)b << (8 * k); ++k; }
    uint32_t* w = reinterpret_cast<uint32_t*>(rec);
    w[kWState] = kStateDone;
    w[kWAlt] = (uint32_t)n_alt | (alts << 8);                       // n_alt, alt[0..2]
    w[kWInfo] = (alts >> 24) | ((uint32_t)n_act << 8) | ((flags & 0xffu) << 16) | (em_calls << 24);   // alt[3], n_active, flags, em_calls
    rec->af[0] = af[0]; rec->af[1] = af[1]; rec->af[2] = af[2]; rec->af[3] = af[3];
    rec->qual = qual;
    rec->chi2 = chi;
    rec->fs_vcf = fs_vcf;
    // called sites go on to the rank-sum / population-group kernels (bv_call_kernels.cuh)
    if (n_alt && a.list_called) a.list_called[atomicAdd(a.counters + kCntCalled, 1u)] = site;
}

__global__ void __launch_bounds__(128) bv_decide_kernel(const __grid_constant__ SiteKernelArgs a) {
    const uint32_t n_hdr = a.counters[kCntEmHdr];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_hdr; i += gridDim.x * blockDim.x) decide_site(a, a.em_hdr[i]);
}

}  // namespace bv
