// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A C-ABI shim around the UNMODIFIED reference (`BaseType`, `strand_bias` of
// /root/reference/src/basetype.{h,cpp}, linked with its utils.cpp and htslib/kfunc.c).  It is
// compiled by oracle/build_ref.sh into oracle/_ref/libbvref.so (git-ignored; travels to the GPU
// box) and is used (a) to pin the C restatement oracle/bv_oracle.c, (b) to generate tests/golden,
// (c) as the `--impl reference` / cpu_baseline arm of bench.py.  No reference source is copied:
// this file only includes the reference's public header at build time.
//
// It drives the reference exactly the way `_basevar_caller` does (src/basetype_caller.cpp:738-761):
//   _out_cvg_line: strand_bias(upper_ref, <non-ref ACGT>, first chars, strands)      (:1236-1245)
//   BaseType bt(&samples_bi, min_af); bt.lrt();                                      (:742-743)
//   if ALT: strand_bias(upper_ref, ALT string, first chars, strands)                 (:1164)
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "basetype.h"  // the reference's header (-I /root/reference/src)

#include "../include/basevar_b200.h"

namespace {

const char kBaseChar[8] = {'A', 'C', 'G', 'T', 'R', 'N', '+', '-'};

struct SiteScratch {
    BatchInfo bi;
    std::vector<char> first;
    void resize(size_t n) {
        bi.n = n;
        bi.align_bases.assign(n, "N");
        bi.align_base_quals.assign(n, '!');
        bi.mapqs.assign(n, 0);
        bi.map_strands.assign(n, '.');
        bi.base_pos_ranks.assign(n, 0);
        first.assign(n, 'N');
    }
};

int base_code(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

// returns seconds spent inside the reference's own code (ctor + lrt + strand_bias)
double run_site(SiteScratch& sc, const uint8_t* base, const uint8_t* qual, const uint8_t* strand, uint32_t n,
                uint8_t ref_char, double min_af, bv_site_out* out) {
    std::memset(out, 0, sizeof(*out));
    BatchInfo& bi = sc.bi;
    bi.ref_id = "chrS";
    bi.ref_pos = 1;
    bi.ref_base.assign(1, (char)ref_char);
    bi.depth = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint8_t b = base[i] & 7;
        if (b == BV_BASE_INS) bi.align_bases[i] = "+A";
        else if (b == BV_BASE_DEL) bi.align_bases[i] = "-A";
        else bi.align_bases[i].assign(1, kBaseChar[b]);
        sc.first[i] = bi.align_bases[i][0];
        bi.align_base_quals[i] = (char)(qual[i] + 33);
        bi.map_strands[i] = strand[i] == BV_STRAND_FWD ? '+' : strand[i] == BV_STRAND_REV ? '-' : '.';
        if (b != BV_BASE_N) bi.depth++;
        if (b == BV_BASE_OTHER) out->depth_other++;
        if (b < 4) {
            if (strand[i] == BV_STRAND_FWD) out->fwd[b]++;
            else if (strand[i] == BV_STRAND_REV) out->rev[b]++;
        }
    }
    auto t0 = std::chrono::steady_clock::now();
    unsigned flags = 0;
    char upper_ref = (char)toupper(bi.ref_base[0]);
    try {
        std::string cvg_alt;
        for (char b : BASES) if (b != upper_ref) cvg_alt.push_back(b);
        StrandBiasInfo s = strand_bias(upper_ref, cvg_alt, sc.first, bi.map_strands);
        out->fs_cvg = s.fs;
    } catch (const std::runtime_error&) {
        flags |= BV_FLAG_BAD_STRAND;
    }
    try {
        BaseType bt(&bi, min_af);
        bt.lrt();
        for (int b = 0; b < 4; ++b) out->depth[b] = (uint32_t)bt.get_base_depth(BASES[b]);
        const std::vector<char>& alts = bt.get_alt_bases();
        out->n_alt = (uint8_t)alts.size();
        std::string alt_str;
        for (size_t k = 0; k < alts.size() && k < 4; ++k) {
            out->alt[k] = (uint8_t)base_code(alts[k]);
            out->af[k] = bt.get_lrt_af(alts[k]);
            alt_str.push_back(alts[k]);
        }
        if (!alts.empty()) {
            out->qual = bt.get_var_qual();
            if (out->qual == 5000.0) flags |= BV_FLAG_MONO_QUAL;
            try {
                StrandBiasInfo s = strand_bias(upper_ref, alt_str, sc.first, bi.map_strands);
                out->fs_vcf = s.fs;
            } catch (const std::runtime_error&) {
                flags |= BV_FLAG_BAD_STRAND;
            }
        }
    } catch (const std::runtime_error&) {
        flags |= BV_FLAG_ZERO_SUBSET;
    }
    auto t1 = std::chrono::steady_clock::now();
    out->flags = (uint8_t)flags;
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // namespace

extern "C" {

// Which abs() did the reference's EM get?  0 = int abs(int) (as built), 1 = double (-include stdlib.h).
int bvref_abs_mode(void) {
#ifdef BVREF_DBLABS
    return 1;
#else
    return 0;
#endif
}

int bvref_site(const uint8_t* base, const uint8_t* qual, const uint8_t* strand, uint32_t n_samples,
               uint8_t ref_char, float min_af, bv_site_out* out) {
    SiteScratch sc;
    sc.resize(n_samples);
    run_site(sc, base, qual, strand, n_samples, ref_char, (double)min_af, out);
    return 0;
}

// Sites are split into `n_threads` contiguous ranges (one std::thread each, like the reference's
// ThreadPool tasks).  core_seconds (may be NULL) receives max over threads of the time spent
// inside the reference's BaseType/lrt/strand_bias calls, i.e. without this shim's BatchInfo fill.
int bvref_tile(const uint8_t* base, const uint8_t* qual, const uint8_t* strand, const uint8_t* ref_base,
               uint64_t pitch, uint32_t n_sites, uint32_t n_samples, float min_af, int n_threads,
               bv_site_out* out, double* core_seconds) {
    if (n_threads < 1) n_threads = 1;
    std::vector<double> secs(n_threads, 0.0);
    auto work = [&](int t) {
        SiteScratch sc;
        sc.resize(n_samples);
        uint64_t s0 = (uint64_t)n_sites * t / n_threads, s1 = (uint64_t)n_sites * (t + 1) / n_threads;
        double acc = 0.0;
        for (uint64_t s = s0; s < s1; ++s)
            acc += run_site(sc, base + s * pitch, qual + s * pitch, strand + s * pitch, n_samples, ref_base[s],
                            (double)min_af, out + s);
        secs[t] = acc;
    };
    if (n_threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    if (core_seconds) {
        double m = 0.0;
        for (double v : secs) m = v > m ? v : m;
        *core_seconds = m;
    }
    return 0;
}

// Direct known-answer access to the reference's numerics (through its public functions only).
double bvref_fisher_fs(int ref_fwd, int ref_rev, int alt_fwd, int alt_rev) {
    // strand_bias() over a synthetic read list reproduces fisher_exact_test + the FS rule.
    std::vector<char> bases, strands;
    for (int i = 0; i < ref_fwd; ++i) { bases.push_back('A'); strands.push_back('+'); }
    for (int i = 0; i < ref_rev; ++i) { bases.push_back('A'); strands.push_back('-'); }
    for (int i = 0; i < alt_fwd; ++i) { bases.push_back('C'); strands.push_back('+'); }
    for (int i = 0; i < alt_rev; ++i) { bases.push_back('C'); strands.push_back('-'); }
    return strand_bias('A', "C", bases, strands).fs;
}

// The three rank-sum INFO values of one called site, exactly as _out_vcf_line computes them
// (src/basetype_caller.cpp:1151-1157): int = ref_vs_alt_ranksumtest(upper REF, ALT string, first bases, values).
int bvref_ranksums(const uint8_t* base, const uint8_t* qual, const uint8_t* mapq, const uint16_t* rpr, uint32_t n_samples,
                   uint8_t ref_char, uint32_t alt_mask, int32_t out3[3]) {
    std::vector<char> first(n_samples), quals(n_samples);
    std::vector<int> mapqs(n_samples), ranks(n_samples);
    for (uint32_t i = 0; i < n_samples; ++i) {
        uint8_t b = base[i] & 7;
        first[i] = b == BV_BASE_INS ? '+' : b == BV_BASE_DEL ? '-' : kBaseChar[b];
        quals[i] = (char)(qual[i] + 33);
        mapqs[i] = mapq[i];
        ranks[i] = rpr[i];
    }
    std::string alt_str;
    for (int b = 0; b < 4; ++b) if (alt_mask >> b & 1u) alt_str.push_back(BASES[b]);
    char upper_ref = (char)toupper((char)ref_char);
    int mq = ref_vs_alt_ranksumtest(upper_ref, alt_str, first, mapqs);
    int rp = ref_vs_alt_ranksumtest(upper_ref, alt_str, first, ranks);
    int bq = ref_vs_alt_ranksumtest(upper_ref, alt_str, first, quals);
    out3[0] = mq; out3[1] = rp; out3[2] = bq;
    return 0;
}

// One population group of one called site: __gb (src/basetype_caller.cpp:767-797) restated with the reference's own
// class -- BaseType over the group's BatchInfo, lrt([upper REF, ALT...]) (:747-760) -- and the ALT / AF read-out of
// _out_vcf_line (:1184-1194).  alts[]: the site's ALT base codes in get_alt_bases() order.
int bvref_group_site(const uint8_t* base, const uint8_t* qual, uint32_t n_samples, const uint8_t* sample_group,
                     uint32_t g, uint8_t ref_char, const uint8_t* alts, int n_alts, float min_af, bv_group_out* out) {
    std::memset(out, 0, sizeof(*out));
    BatchInfo bi;
    bi.ref_id = "chrS"; bi.ref_pos = 1; bi.ref_base.assign(1, (char)ref_char); bi.depth = 0;
    for (uint32_t i = 0; i < n_samples; ++i) {
        if (sample_group[i] != g) continue;
        uint8_t b = base[i] & 7;
        if (b == BV_BASE_INS) bi.align_bases.push_back("+A");
        else if (b == BV_BASE_DEL) bi.align_bases.push_back("-A");
        else bi.align_bases.push_back(std::string(1, kBaseChar[b]));
        bi.align_base_quals.push_back((char)(qual[i] + 33));
        bi.mapqs.push_back(0);
        bi.map_strands.push_back('+');
        bi.base_pos_ranks.push_back(0);
    }
    bi.n = bi.align_bases.size();
    std::vector<char> basecombination;
    basecombination.push_back((char)toupper((char)ref_char));
    for (int k = 0; k < n_alts; ++k) basecombination.push_back(BASES[alts[k] & 3]);
    try {
        BaseType bt(&bi, (double)min_af);
        bt.lrt(basecombination);
        for (char b : bt.get_alt_bases()) {
            if (out->n_alt >= 4) break;
            out->alt[out->n_alt] = (uint8_t)base_code(b);
            out->af[out->n_alt] = bt.get_lrt_af(b);
            out->n_alt++;
        }
    } catch (const std::runtime_error&) {
        out->flags |= BV_FLAG_ZERO_SUBSET;
    }
    return 0;
}

}  // extern "C"
