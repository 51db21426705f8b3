// bv_expand_kernel.cuh -- K0: sparse host tile -> dense site-major planes in HBM (sm_100a).
//
// A sparse tile (bv_sparse_tile, include/basevar_b200.h) carries only the covered cells of a pileup tile, one packed
// u32 each, grouped by site.  This is what crosses PCIe: at 0.1x depth 0.4 bytes per sample-site instead of the 2-3
// bytes of the dense planes.  K0 turns it back into the planes every other kernel reads (the layout BatchInfo is
// replaced by, src/basetype.h:25-43).  One warp per site:
//
//   staged path: the row is built chunk by chunk (1 KB per plane) in shared memory -- the "uncovered" filler (`N`, phred 0,
//   strand none: the reference's `N ! 0 0 .`, src/basetype_caller.cpp:1063-1075) with 16-byte stores, then the site's cells
//   scattered into it with byte stores -- and leaves as three TMA bulk stores (cp.async.bulk.global.shared::cta; SASS UBLKCP)
//   from one of two buffers per warp, so that the next chunk is built while the previous one drains.  Every plane byte goes to
//   HBM exactly once, in full lines, and no thread waits for a store; offsets (two sites ahead) and cells (one site ahead) are
//   prefetched into registers.  BV_CELLS_U16 words ascend by sample, so they are consumed front to back across the chunks
//   (stream_cells16: about one extra batch of 128 words per chunk) and rows of ANY length are staged: a 9,472-site tile of
//   100,000-sample rows (2.84 GB of planes) takes 0.73 ms = 3.9 TB/s where the direct path below took 2.9 ms
//   (gpurun r2t, profiles/r02_k0_long_rows.txt).  BV_CELLS_U32 cells may come in any order within a site, so a row of k
//   chunks scans the site's cells k times (they stay in L1): staged up to 16 chunks;
//   direct path (longer BV_CELLS_U32 rows): filler and cells are written to global memory directly; the scatter hits lines the
//   same warp has just written and merges in L2.
//
// Two cell formats (include/basevar_b200.h): BV_CELLS_U32, one self-contained word per cell, any order within a site;
// BV_CELLS_U16, two bytes per cell with the sample index delta-coded against the previous cell of the site (ascending
// samples, "skip 31" words for longer gaps): the warp turns a batch of 32 words into sample indices with one inclusive
// scan of the deltas.
// The mapq / rpr planes of the called-site kernels (only with cells_aux) always take the direct path.
// Bound: HBM writes, 3 bytes per sample-site (+3 with the called-site planes), reads 4 bytes per covered cell.
#pragma once
#include "bv_common.cuh"

namespace bv {

constexpr int kExpandWarps = 16;
constexpr int kExChunk = 1024;                       // bytes (= samples) per staged chunk and plane
constexpr uint32_t kExStagedMaxPitch = 16 * kExChunk;

struct __align__(128) ExpandBuf {
    uint8_t base[kExChunk];
    uint8_t qual[kExChunk];
    uint8_t strand[kExChunk];
};
struct __align__(128) ExpandWarp {
    ExpandBuf buf[2];
};
constexpr size_t kExpandSmemBytes = (size_t)kExpandWarps * sizeof(ExpandWarp);   // 96 KB: two CTAs per SM

struct ExpandArgs {
    const void* cells;           // u32 (BV_CELLS_U32) or u16 (BV_CELLS_U16) words
    const uint32_t* cells_aux;   // null: no mapq / rpr planes
    const uint32_t* site_start;  // [n_sites + 1]
    uint8_t* base;
    uint8_t* qual;
    uint8_t* strand;
    uint8_t* mapq;               // [n_sites][pitch] or null
    uint16_t* rpr;               // [n_sites][rpr_pitch] or null
    uint32_t* counters;
    uint64_t pitch;              // bytes per row of the u8 planes, multiple of 16
    uint64_t rpr_pitch;          // elements per row of the rpr plane, multiple of 8
    uint64_t n_cells;
    uint32_t n_sites;
    uint32_t n_samples;
};

__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

constexpr int kExPre = 4;   // words per lane prefetched one site ahead (the first 128 words of a site)

template <int FMT>
__device__ __forceinline__ uint32_t load_word(const void* cells, uint64_t c) {
    if (FMT == BV_CELLS_U16) return __ldg(reinterpret_cast<const uint16_t*>(cells) + c);
    return __ldg(reinterpret_cast<const uint32_t*>(cells) + c);
}

// Word k of a site's prefetch registers: BV_CELLS_U32 keeps the strided order of coalesced 4-byte loads (pre[k] of lane l
// is word 32 k + l); BV_CELLS_U16 gives every lane kExPre CONSECUTIVE words (pre[k] of lane l is word 4 l + k), so that a
// batch of 128 delta-coded words needs one warp scan instead of four.
template <int FMT>
__device__ __forceinline__ uint64_t pre_index(uint32_t lane, int k) {
    return FMT == BV_CELLS_U16 ? (uint64_t)(kExPre * lane + (uint32_t)k) : (uint64_t)(32u * (uint32_t)k + lane);
}

// Visits the cells of one site, 128 words per step: f(sample, base, strand, phred, word index) for every cell.
// `upto` (BV_CELLS_U16 only; samples ascend): stops once every further cell has sample >= upto.
// Returns false when a cell's sample index is >= n_samples (malformed input).
template <int FMT, class F>
__device__ __forceinline__ bool for_each_cell(const void* cells, uint64_t beg, uint64_t end, const uint32_t (&pre)[kExPre],
                                              uint32_t lane, uint32_t n_samples, uint32_t upto, F&& f) {
    bool good = true;
    uint32_t next = 0;   // BV_CELLS_U16: the sample index a gap of 0 would mean (warp-uniform)
    bool first = true;
#pragma unroll 1
    for (uint64_t c0 = beg; c0 < end; c0 += 32 * kExPre) {
        uint32_t w[kExPre];
#pragma unroll
        for (int k = 0; k < kExPre; ++k) {
            const uint64_t c = c0 + pre_index<FMT>(lane, k);
            w[k] = first ? pre[k] : (c < end ? load_word<FMT>(cells, c) : 0u);   // (`first` is warp-uniform)
        }
        first = false;
        if (FMT == BV_CELLS_U16) {
            // lane l holds words 4 l .. 4 l + 3 of the batch: local prefix of the sample increments, one scan of the lane totals
            uint32_t inc[kExPre], tot = 0;
#pragma unroll
            for (int k = 0; k < kExPre; ++k) {
                const bool live = c0 + pre_index<FMT>(lane, k) < end;
                const uint32_t gap = w[k] & 31u;
                inc[k] = live ? (gap == BV_CELL16_GAP_SKIP ? BV_CELL16_GAP_SKIP : gap + 1u) : 0u;
                tot += inc[k];
            }
            uint32_t x = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(kFull, x, o);
                if ((int)lane >= o) x += y;
            }
            uint32_t at = next + (x - tot);   // the sample index a gap of 0 would mean at this lane's first word
            next += __shfl_sync(kFull, x, 31);
#pragma unroll
            for (int k = 0; k < kExPre; ++k) {
                const uint64_t c = c0 + pre_index<FMT>(lane, k);
                const uint32_t gap = w[k] & 31u;
                if (c < end && gap != BV_CELL16_GAP_SKIP) {
                    const uint32_t sample = at + gap;
                    if (sample >= n_samples) good = false;
                    else f(sample, (w[k] >> 5) & 7u, (w[k] >> 8) & 1u, w[k] >> 9, c);
                }
                at += inc[k];
            }
            if (next >= upto) break;   // warp-uniform
        } else {
#pragma unroll
            for (int k = 0; k < kExPre; ++k) {
                const uint64_t c = c0 + pre_index<FMT>(lane, k);
                if (c < end) {
                    const uint32_t sample = w[k] & (BV_CELL_MAX_SAMPLES - 1u);
                    if (sample >= n_samples) good = false;
                    else f(sample, (w[k] >> 20) & 7u, (w[k] >> 23) & 3u, w[k] >> 25, c);
                }
            }
        }
    }
    return good;
}

// BV_CELLS_U16 (samples ascend within a site): the words of a site consumed front to back, chunk by chunk of the row.
// `cur` (warp-uniform) is the first word not consumed yet and the sample index a gap of 0 would mean there.  Visits batches of
// 128 words from the cursor on -- f(sample, base, strand, phred, word index) for every cell -- and moves the cursor past every
// batch whose cells all lie below `limit`; a batch that reaches beyond it is visited again by the next call (for the next chunk
// of the row, whose f ignores the cells below its own range): about one batch per 1,024-sample chunk at 0.1x, whatever the
// length of the row.  Returns false when a cell's sample index is >= n_samples.
struct CellCursor {
    uint64_t c;
    uint32_t next;
};
template <class F>
__device__ __forceinline__ bool stream_cells16(const void* cells, uint64_t beg, uint64_t end, const uint32_t (&pre)[kExPre], uint32_t lane,
                                               uint32_t n_samples, uint32_t limit, CellCursor& cur, F&& f) {
    bool good = true;
#pragma unroll 1
    while (cur.c < end) {
        const bool first = cur.c == beg;   // (warp-uniform) the batch the caller prefetched
        uint32_t w[kExPre], inc[kExPre], tot = 0;
#pragma unroll
        for (int k = 0; k < kExPre; ++k) {
            const uint64_t c = cur.c + pre_index<BV_CELLS_U16>(lane, k);
            w[k] = first ? pre[k] : (c < end ? load_word<BV_CELLS_U16>(cells, c) : 0u);
            const uint32_t gap = w[k] & 31u;
            inc[k] = c < end ? (gap == BV_CELL16_GAP_SKIP ? BV_CELL16_GAP_SKIP : gap + 1u) : 0u;
            tot += inc[k];
        }
        uint32_t x = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, x, o);
            if ((int)lane >= o) x += y;
        }
        uint32_t at = cur.next + (x - tot);
        const uint32_t after = cur.next + __shfl_sync(kFull, x, 31);
#pragma unroll
        for (int k = 0; k < kExPre; ++k) {
            const uint64_t c = cur.c + pre_index<BV_CELLS_U16>(lane, k);
            const uint32_t gap = w[k] & 31u;
            if (c < end && gap != BV_CELL16_GAP_SKIP) {
                const uint32_t sample = at + gap;
                if (sample >= n_samples) good = false;
                else f(sample, (w[k] >> 5) & 7u, (w[k] >> 8) & 1u, w[k] >> 9, c);
            }
            at += inc[k];
        }
        if (after > limit) break;   // (warp-uniform) the batch reaches into the next chunk
        cur.c += 32 * kExPre;
        cur.next = after;
    }
    return good;
}

template <int FMT>
__global__ void __launch_bounds__(kExpandWarps * 32, 2) bv_expand_kernel(const ExpandArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * kExpandWarps + (threadIdx.x >> 5);
    const uint32_t n_warps = gridDim.x * kExpandWarps;
    ExpandWarp& W = reinterpret_cast<ExpandWarp*>(bv_smem_raw)[threadIdx.x >> 5];
    const uint32_t pitch = (uint32_t)a.pitch;
    // BV_CELLS_U16 rows of any length are staged (their cells are consumed front to back, stream_cells16); BV_CELLS_U32 cells
    // come in any order, so every chunk of the row scans all of them: staged up to kExStagedMaxPitch, written directly beyond
    const bool staged = FMT == BV_CELLS_U16 || a.pitch <= kExStagedMaxPitch;
    const uint4 fill_base = make_uint4(0x05050505u, 0x05050505u, 0x05050505u, 0x05050505u);     // BV_BASE_N
    const uint4 fill_strand = make_uint4(0x02020202u, 0x02020202u, 0x02020202u, 0x02020202u);   // BV_STRAND_NONE
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    bool bad = false;
    uint32_t cur = 0;   // staging buffer of the next chunk

    // Software pipeline over the warp's sites: the offsets are loaded two sites ahead and the first kExPre * 32 words
    // one site ahead, so that their DRAM latency (the cells have just arrived over PCIe) hides behind the site in hand.
    auto load_offsets = [&](uint32_t site, uint64_t& b, uint64_t& e) {
        b = 0; e = 0;
        if (site < a.n_sites) { b = a.site_start[site]; e = a.site_start[site + 1]; }
    };
    auto load_cells = [&](uint64_t b, uint64_t e, uint32_t (&pre)[kExPre]) {
        const bool ok = e >= b && e <= a.n_cells;
#pragma unroll
        for (int k = 0; k < kExPre; ++k) {
            const uint64_t c = b + pre_index<FMT>(lane, k);
            pre[k] = (ok && c < e) ? load_word<FMT>(a.cells, c) : 0u;
        }
    };
    uint32_t s = warp;
    uint64_t beg, end, nbeg, nend;
    uint32_t pre[kExPre], npre[kExPre];
    load_offsets(s, beg, end);
    load_offsets(s + n_warps, nbeg, nend);
    load_cells(beg, end, pre);
    while (s < a.n_sites) {
        const uint32_t s_next = s + n_warps;
        uint64_t n2beg, n2end;
        load_offsets(s_next + n_warps, n2beg, n2end);
        load_cells(nbeg, nend, npre);
        const bool ok = end >= beg && end <= a.n_cells;   // warp-uniform
        if (!ok) { bad = true; end = beg; }               // the row still gets its filler
        const size_t row = (size_t)s * a.pitch;

        if (a.mapq) {   // called-site planes: direct path
            uint4* rm = reinterpret_cast<uint4*>(a.mapq + row);
            for (uint32_t v = lane; v < (pitch >> 4); v += 32) rm[v] = zero;
            uint4* rr = reinterpret_cast<uint4*>(a.rpr + (size_t)s * a.rpr_pitch);
            const uint32_t rvecs = (uint32_t)(a.rpr_pitch >> 3);
            for (uint32_t v = lane; v < rvecs; v += 32) rr[v] = zero;
            __syncwarp();   // orders the filler before the cell stores of other lanes
            for_each_cell<FMT>(a.cells, beg, end, pre, lane, a.n_samples, 0xffffffffu,
                               [&](uint32_t i, uint32_t, uint32_t, uint32_t, uint64_t c) {
                                   const uint32_t x = __ldg(a.cells_aux + c);
                                   a.mapq[row + i] = (uint8_t)x;
                                   a.rpr[(size_t)s * a.rpr_pitch + i] = (uint16_t)(x >> 8);
                               });
        }

        if (staged) {
            CellCursor cursor;
            cursor.c = beg; cursor.next = 0;
#pragma unroll 1
            for (uint32_t off = 0; off < pitch; off += kExChunk) {
                const uint32_t bytes = min((uint32_t)kExChunk, pitch - off);
                ExpandBuf& B = W.buf[cur];
                // the bulk stores issued from this buffer two chunks ago must have read it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                for (uint32_t v = lane; v < (bytes >> 4); v += 32) {
                    reinterpret_cast<uint4*>(B.base)[v] = fill_base;
                    reinterpret_cast<uint4*>(B.qual)[v] = zero;
                    reinterpret_cast<uint4*>(B.strand)[v] = fill_strand;
                }
                __syncwarp();   // filler before the cells of other lanes
                // ascending samples (BV_CELLS_U16): a pass stops at the end of its chunk; the last one runs to the end of
                // the words so that every malformed cell (sample >= n_samples) is seen
                const uint32_t upto = off + kExChunk >= pitch ? 0xffffffffu : off + bytes;
                auto put = [&](uint32_t i, uint32_t b, uint32_t st, uint32_t q, uint64_t) {
                    const uint32_t k = i - off;
                    if (k < bytes) {   // (unsigned: also false for i < off)
                        B.base[k] = (uint8_t)b;
                        B.strand[k] = (uint8_t)st;
                        B.qual[k] = (uint8_t)q;
                    }
                };
                const bool good = FMT == BV_CELLS_U16 ? stream_cells16(a.cells, beg, end, pre, lane, a.n_samples, upto, cursor, put)
                                                      : for_each_cell<FMT>(a.cells, beg, end, pre, lane, a.n_samples, upto, put);
                if (!good) bad = true;
                // generic-proxy writes -> visible to the async proxy, then one lane hands the chunk to the TMA unit
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    bulk_s2g(a.base + row + off, smem_u32(B.base), bytes);
                    bulk_s2g(a.qual + row + off, smem_u32(B.qual), bytes);
                    bulk_s2g(a.strand + row + off, smem_u32(B.strand), bytes);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                cur ^= 1u;
            }
        } else {
            uint4* rb = reinterpret_cast<uint4*>(a.base + row);
            uint4* rq = reinterpret_cast<uint4*>(a.qual + row);
            uint4* rs = reinterpret_cast<uint4*>(a.strand + row);
            for (uint32_t v = lane; v < (pitch >> 4); v += 32) {
                rb[v] = fill_base;
                rq[v] = zero;
                rs[v] = fill_strand;
            }
            __syncwarp();   // orders the filler before the cell stores of other lanes
            const bool good = for_each_cell<FMT>(a.cells, beg, end, pre, lane, a.n_samples, 0xffffffffu,
                                                 [&](uint32_t i, uint32_t b, uint32_t st, uint32_t q, uint64_t) {
                                                     a.base[row + i] = (uint8_t)b;
                                                     a.strand[row + i] = (uint8_t)st;
                                                     a.qual[row + i] = (uint8_t)q;
                                                 });
            if (!good) bad = true;
        }
        s = s_next; beg = nbeg; end = nend; nbeg = n2beg; nend = n2end;
#pragma unroll
        for (int k = 0; k < kExPre; ++k) pre[k] = npre[k];
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the stores are complete before the warp leaves
    if (__any_sync(kFull, bad) && lane == 0) atomicAdd(a.counters + kCntBadCell, 1u);
}

}  // namespace bv
