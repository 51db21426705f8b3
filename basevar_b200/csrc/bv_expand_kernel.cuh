// bv_expand_kernel.cuh -- K0: sparse host tile -> dense site-major planes in HBM (sm_100a).
//
// A sparse tile (bv_sparse_tile, include/basevar_b200.h) carries only the covered cells of a pileup tile, one packed
// u32 each, grouped by site.  This is what crosses PCIe: at 0.1x depth 0.4 bytes per sample-site instead of the 2-3
// bytes of the dense planes.  K0 turns it back into the planes every other kernel reads (the layout BatchInfo is
// replaced by, src/basetype.h:25-43): one warp per site writes the row's "uncovered" filler
// (`N`, phred 0, strand none -- the reference's `N ! 0 0 .`, src/basetype_caller.cpp:1063-1075) with 16-byte stores and
// then scatters the site's cells into it.  The scatter hits lines the same warp has just written, so it merges in L2
// and every plane byte goes to DRAM once.
//
// Bound: HBM writes, 3 bytes per sample-site (+3 with the called-site planes), reads 4 bytes per covered cell.
#pragma once
#include "bv_common.cuh"

namespace bv {

constexpr int kExpandWarps = 16;

struct ExpandArgs {
    const uint32_t* cells;
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

__global__ void __launch_bounds__(kExpandWarps * 32) bv_expand_kernel(const ExpandArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * kExpandWarps + (threadIdx.x >> 5);
    const uint32_t n_warps = gridDim.x * kExpandWarps;
    const uint32_t vecs = (uint32_t)(a.pitch >> 4);
    const uint4 fill_base = make_uint4(0x05050505u, 0x05050505u, 0x05050505u, 0x05050505u);     // BV_BASE_N
    const uint4 fill_strand = make_uint4(0x02020202u, 0x02020202u, 0x02020202u, 0x02020202u);   // BV_STRAND_NONE
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    bool bad = false;
    for (uint32_t s = warp; s < a.n_sites; s += n_warps) {
        // the offsets first, so that the loads are in flight while the filler is written
        const uint64_t beg = a.site_start[s], end = a.site_start[s + 1];
        const size_t row = (size_t)s * a.pitch;
        uint4* rb = reinterpret_cast<uint4*>(a.base + row);
        uint4* rq = reinterpret_cast<uint4*>(a.qual + row);
        uint4* rs = reinterpret_cast<uint4*>(a.strand + row);
        for (uint32_t v = lane; v < vecs; v += 32) {
            rb[v] = fill_base;
            rq[v] = zero;
            rs[v] = fill_strand;
        }
        if (a.mapq) {
            uint4* rm = reinterpret_cast<uint4*>(a.mapq + row);
            for (uint32_t v = lane; v < vecs; v += 32) rm[v] = zero;
            uint4* rr = reinterpret_cast<uint4*>(a.rpr + (size_t)s * a.rpr_pitch);
            const uint32_t rvecs = (uint32_t)(a.rpr_pitch >> 3);
            for (uint32_t v = lane; v < rvecs; v += 32) rr[v] = zero;
        }
        if (end < beg || end > a.n_cells) { bad = true; continue; }   // warp-uniform
        __syncwarp();   // orders the filler before the cell stores of other lanes
        for (uint64_t c = beg + lane; c < end; c += 32) {
            const uint32_t w = __ldg(a.cells + c);
            const uint32_t i = w & (BV_CELL_MAX_SAMPLES - 1u);
            if (i >= a.n_samples) { bad = true; continue; }
            a.base[row + i] = (uint8_t)((w >> 20) & 7u);
            a.strand[row + i] = (uint8_t)((w >> 23) & 3u);
            a.qual[row + i] = (uint8_t)(w >> 25);
            if (a.mapq) {
                const uint32_t x = __ldg(a.cells_aux + c);
                a.mapq[row + i] = (uint8_t)x;
                a.rpr[(size_t)s * a.rpr_pitch + i] = (uint16_t)(x >> 8);
            }
        }
    }
    if (__any_sync(kFull, bad) && lane == 0) atomicAdd(a.counters + kCntBadCell, 1u);
}

}  // namespace bv
