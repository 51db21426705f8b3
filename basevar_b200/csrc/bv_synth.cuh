// bv_synth.cuh -- counter-based synthetic pileup generator, shared by the device
// kernel and its bit-identical host twin.  Integer arithmetic only: every probability is a pre-computed
// integer threshold in bv_synth_model, so host and device produce the same bytes.
//
// A cell (site, sample) is a pure function of (seed, site, sample): no state, any tile of the synthetic
// genome can be generated on any GPU without communication.
#pragma once
#include <stdint.h>

#include "../../include/basevar_b200.h"

#ifdef __CUDACC__
#define BV_HD __host__ __device__ __forceinline__
#else
#define BV_HD inline
#endif

namespace bv {

BV_HD uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct SynthSite {
    uint64_t h;        // site hash, keys every cell of the row
    uint32_t thr[3];   // cumulative u32 thresholds: draw < thr[k] (first k that matches) => true base is alt[k]
    uint8_t ref;       // 0..3
    uint8_t alt[3];    // the three non-reference bases in a site-specific order
};

BV_HD SynthSite synth_site(const bv_synth_model* m, uint64_t site) {
    SynthSite s;
    const uint64_t h = mix64(m->seed ^ mix64(site + 0x9E3779B97F4A7C15ull));
    s.h = h;
    s.ref = (uint8_t)(h & 3);
    const uint32_t rot = (uint32_t)((h >> 2) & 0xffff) % 3u;
    for (uint32_t k = 0; k < 3; ++k) s.alt[k] = (uint8_t)((s.ref + 1 + (rot + k) % 3u) & 3);
    const uint64_t hv = mix64(h + 1);
    const bool variant = (uint32_t)hv < m->var_thr;
    uint64_t t1 = variant ? m->af_thr[(hv >> 32) & 1023] : 0;
    const uint64_t hm = mix64(h + 2);
    const bool multi = variant && ((uint32_t)hm < m->multi_thr);
    uint64_t t2 = multi ? m->af_extra_thr[(hm >> 32) & 255] : 0;
    uint64_t t3 = (multi && ((hm >> 40) & 1)) ? m->af_extra_thr[(hm >> 48) & 255] : 0;
    uint64_t c1 = t1, c2 = c1 + t2, c3 = c2 + t3;
    const uint64_t cap = 0xffffffffull;
    s.thr[0] = (uint32_t)(c1 > cap ? cap : c1);
    s.thr[1] = (uint32_t)(c2 > cap ? cap : c2);
    s.thr[2] = (uint32_t)(c3 > cap ? cap : c3);
    return s;
}

struct SynthCell {
    uint8_t base, qual, strand, mapq;
};

BV_HD SynthCell synth_cell(const bv_synth_model* m, const SynthSite& s, uint64_t sample) {
    SynthCell c;
    const uint64_t k1 = mix64(s.h + (sample + 1) * 0xD1B54A32D192ED03ull);
    if ((uint32_t)k1 >= m->cov_thr) {
        c.base = BV_BASE_N; c.qual = 0; c.strand = BV_STRAND_NONE; c.mapq = 0;
        return c;
    }
    const uint32_t q = m->q_lo + (uint32_t)(((k1 >> 32) * (uint64_t)m->q_span) >> 32);
    const uint64_t k2 = mix64(k1 ^ 0xA0761D6478BD642Full);
    const uint32_t a = (uint32_t)k2;
    uint32_t b = a < s.thr[0] ? s.alt[0] : a < s.thr[1] ? s.alt[1] : a < s.thr[2] ? s.alt[2] : s.ref;
    const uint32_t e = (uint32_t)(k2 >> 32) & 0xffffffu;
    if (e < m->err_thr[q]) b = (b + 1 + (uint32_t)((k2 >> 58) % 3u)) & 3;  // sequencing error: another base
    c.base = (uint8_t)b;
    c.qual = (uint8_t)q;
    c.strand = (uint8_t)((k2 >> 56) & 1);
    const uint64_t k3 = mix64(k2 + 0x632BE59BD9B4E019ull);
    c.mapq = ((k3 & 0xff) < 230) ? 60 : (uint8_t)(10 + ((k3 >> 8) & 0xffff) % 50u);
    return c;
}

// read position rank of a cell (BatchInfo::base_pos_ranks): 1..35 for covered cells, 0 otherwise
BV_HD uint16_t synth_rpr(const bv_synth_model* m, const SynthSite& s, uint64_t sample) {
    const uint64_t k1 = mix64(s.h + (sample + 1) * 0xD1B54A32D192ED03ull);
    if ((uint32_t)k1 >= m->cov_thr) return 0;
    const uint64_t k2 = mix64(k1 ^ 0xA0761D6478BD642Full);
    const uint64_t k3 = mix64(k2 + 0x632BE59BD9B4E019ull);
    return (uint16_t)(1u + (uint32_t)((k3 >> 24) & 0xffff) % 35u);
}

}  // namespace bv
