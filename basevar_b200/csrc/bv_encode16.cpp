// bv_encode16.cpp -- host side of the BV_CELLS_U16 transport: the inner loop of bv_sparse_encode16 (bv_api.cu), eight cells per
// step with AVX2 where the CPU has it.  Plain C++ (g++, no CUDA): the packer of a < 1x pileup produces one u32 cell per covered
// read (sample | base << 20 | strand << 23 | phred << 25, include/basevar_b200.h) and re-codes the cells of a tile into two-byte
// words right before the upload, so this loop runs once per read on the host path of every tile (`encode_submit` in the
// caller's stage timers; `e2e_from_cells.host_encode16` in bench.py).
//
// One word = gap | base << 5 | strand << 8 | phred << 9 with gap = sample - (previous sample of the site + 1); a gap of 31 or
// more is written as "skip 31 samples" words in front of the cell.  The vector step computes all eight gaps at once (the
// previous sample of lane k is lane k - 1), and writes the eight words when no lane needs a skip word or has a strand without a
// 16-bit form; otherwise the lanes in front of the first such cell are written, that cell goes the scalar way, and the loop
// resumes behind it.  At 0.1x one cell in 26 needs a skip word, so a step covers 6 to 7 cells on average.
#include <stdint.h>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define BV_HAVE_X86 1
#else
#define BV_HAVE_X86 0
#endif

#include "../../include/basevar_b200.h"

namespace bv {

bool encode16_have_avx2() {
#if BV_HAVE_X86
    static const bool ok = __builtin_cpu_supports("avx2");
    return ok;
#else
    return false;
#endif
}

#if BV_HAVE_X86
// Same contract as encode16_run (bv_api.cu): encodes cells [p, end) of one site while it can, returns the first cell it could
// not encode (a strand that is neither + nor -, or a sample below its predecessor), or end.  aux32 is pre-offset so that
// aux32[p - end] is the aux word of *p.  The caller guarantees room for the worst case of the site plus 8 words.
template <bool AUX>
__attribute__((target("avx2"))) const uint32_t* encode16_run_avx2(const uint32_t* __restrict__ p, const uint32_t* __restrict__ end,
                                                                   const uint32_t* __restrict__ aux32, uint16_t* __restrict__& o,
                                                                   uint32_t* __restrict__& oa, uint32_t& next) {
    uint32_t nx = next;
    uint16_t* out = o;
    uint32_t* outa = oa;
    const __m256i sample_mask = _mm256_set1_epi32((int)(BV_CELL_MAX_SAMPLES - 1u));
    const __m256i shift_lanes = _mm256_setr_epi32(0, 0, 1, 2, 3, 4, 5, 6);   // lane k takes lane k - 1 (lane 0 is replaced)
    const __m256i thirty = _mm256_set1_epi32((int)BV_CELL16_GAP_SKIP - 1);
    const __m256i one = _mm256_set1_epi32(1);
    while (p < end) {
        if (end - p >= 8) {
            const __m256i w = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
            const __m256i idx = _mm256_and_si256(w, sample_mask);
            __m256i prev = _mm256_permutevar8x32_epi32(idx, shift_lanes);
            prev = _mm256_insert_epi32(prev, (int)(nx - 1u), 0);
            const __m256i gap = _mm256_sub_epi32(_mm256_sub_epi32(idx, prev), one);   // wraps to a huge value when samples descend
            const __m256i f12 = _mm256_srli_epi32(w, 20);
            // ok lanes: gap <= 30 (unsigned) and the strand's high bit clear
            const __m256i gap_ok = _mm256_cmpeq_epi32(_mm256_min_epu32(gap, thirty), gap);
            const __m256i strand_ok = _mm256_cmpeq_epi32(_mm256_and_si256(f12, _mm256_set1_epi32(0x10)), _mm256_setzero_si256());
            const unsigned bad = ~(unsigned)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_and_si256(gap_ok, strand_ok))) & 0xffu;
            // word = gap | (base | strand << 3 | phred << 4) << 5
            const __m256i low = _mm256_and_si256(f12, _mm256_set1_epi32(0xf));
            const __m256i hi = _mm256_slli_epi32(_mm256_srli_epi32(f12, 5), 4);
            const __m256i word = _mm256_or_si256(gap, _mm256_slli_epi32(_mm256_or_si256(low, hi), 5));
            // 8 x u32 -> 8 x u16 (values < 2^16 in the ok lanes; the others are overwritten below)
            const __m256i packed = _mm256_permute4x64_epi64(_mm256_packus_epi32(_mm256_and_si256(word, _mm256_set1_epi32(0xffff)), _mm256_setzero_si256()), 0x08);
            _mm_storeu_si128(reinterpret_cast<__m128i*>(out), _mm256_castsi256_si128(packed));
            if (AUX) {
                const __m256i a = aux32 ? _mm256_loadu_si256(reinterpret_cast<const __m256i*>(aux32 + (p - end))) : _mm256_setzero_si256();
                _mm256_storeu_si256(reinterpret_cast<__m256i*>(outa), a);
            }
            if (bad == 0) {
                out += 8;
                if (AUX) outa += 8;
                nx = (p[7] & (BV_CELL_MAX_SAMPLES - 1u)) + 1u;
                p += 8;
                continue;
            }
            const int k = __builtin_ctz(bad);   // the lanes in front of the first cell that needs the scalar path are done
            out += k;
            if (AUX) outa += k;
            if (k) nx = (p[k - 1] & (BV_CELL_MAX_SAMPLES - 1u)) + 1u;
            p += k;
        }
        // one cell the scalar way (also the last < 8 cells of a site)
        const uint32_t w = *p, i = w & (BV_CELL_MAX_SAMPLES - 1u), f12 = w >> 20;
        uint32_t gap = i - nx;
        if ((f12 & 0x10u) || gap >= BV_CELL_MAX_SAMPLES) break;   // the caller reports it
        while (gap >= BV_CELL16_GAP_SKIP) {
            *out++ = (uint16_t)BV_CELL16_GAP_SKIP;
            if (AUX) *outa++ = 0;
            gap -= BV_CELL16_GAP_SKIP;
        }
        *out++ = (uint16_t)(gap | (((f12 & 0xfu) | ((f12 >> 5) << 4)) << 5));
        if (AUX) *outa++ = aux32 ? aux32[p - end] : 0;
        nx = i + 1;
        ++p;
    }
    next = nx; o = out; oa = outa;
    return p;
}

template const uint32_t* encode16_run_avx2<true>(const uint32_t* __restrict__, const uint32_t* __restrict__, const uint32_t* __restrict__,
                                                 uint16_t* __restrict__&, uint32_t* __restrict__&, uint32_t&);
template const uint32_t* encode16_run_avx2<false>(const uint32_t* __restrict__, const uint32_t* __restrict__, const uint32_t* __restrict__,
                                                  uint16_t* __restrict__&, uint32_t* __restrict__&, uint32_t&);
#else
template <bool AUX>
const uint32_t* encode16_run_avx2(const uint32_t* __restrict__ p, const uint32_t* __restrict__, const uint32_t* __restrict__, uint16_t* __restrict__&,
                                  uint32_t* __restrict__&, uint32_t&) { return p; }
template const uint32_t* encode16_run_avx2<true>(const uint32_t* __restrict__, const uint32_t* __restrict__, const uint32_t* __restrict__,
                                                 uint16_t* __restrict__&, uint32_t* __restrict__&, uint32_t&);
template const uint32_t* encode16_run_avx2<false>(const uint32_t* __restrict__, const uint32_t* __restrict__, const uint32_t* __restrict__,
                                                  uint16_t* __restrict__&, uint32_t* __restrict__&, uint32_t&);
#endif

}  // namespace bv
