// bv_caller.hpp -- the per-site driver of `basevar basetype` over the C ABI: batchfile rows in, VCF / CVG text out.
//
// Mirrors the reference's driver for the hot path (same argument meaning, same error texts, same output bytes):
//
//   reference (src/basetype_caller.{h,cpp})                      here (namespace bvhost)
//   ----------------------------------------------------------   -------------------------------------------------------
//   _basevar_caller(lines, group_smp_idx, min_af, n_sample,       BasevarCaller::call(lines)  (queues the position into the
//                   vcf_hd, cvg_hd)               cpp:667-765      current tile) + finish()
//   _out_cvg_line / __base_depth_and_indel        cpp:1211-1289   out_cvg_line(): text from the device record
//   _out_vcf_line                                 cpp:1103-1209   out_vcf_line(): text from the record, the rank sums
//                                                                 (K5) and the population-group calls (K6)
//   __gb / __get_group_batchinfo                  cpp:767-797     bv_set_groups + K6 (sample -> group plane)
//   _variant_calling_unit                         cpp:529-635     variant_calling_unit(): batchfiles (BGZF / gzip text)
//                                                                 read with zlib, lock-step, one row per file
//   vcf_header_define / cvg_header_define   basetype_utils.cpp:32-88   vcf_header_define() / cvg_header_define()
//
// What stays on the host is text: parsing the rows into the packed planes of a tile and formatting the records.
// Likelihoods, EM, LRT, QUAL, Fisher, rank sums and group frequencies all come from the GPU through
// include/basevar_b200.h; without a CUDA device the constructor throws.
//
// Positions are processed in tiles (CallerOptions::tile_sites) over n_slots in-flight slots: while the GPU works on a
// tile the host parses the next one and formats the previous one.  Output order is input order.
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "bv_host.hpp"

namespace bvhost {

using TextSink = std::function<void(const char* data, size_t len)>;

// Wall seconds per stage of the host pipeline, accumulated by whoever runs it (one instance per host worker).
struct StageTimes {
    double decode = 0;         // BAM records -> per-sample cell lists (BamPileup::load_span)
    double scatter = 0;        // cell lists -> site-major rows and the tile's cell list (BamPileup::scatter)
    double tile_reset = 0;     // begin_tile: clearing the tile's rows
    double encode_submit = 0;  // u32 cells -> u16 words, bv_tile_submit*
    double gpu_wait = 0;       // bv_tile_wait*: time the host stood waiting for the device
    double text = 0;           // records -> VCF / CVG rows (and handing them to the sinks)
    void add(const StageTimes& o) {
        decode += o.decode; scatter += o.scatter; tile_reset += o.tile_reset; encode_submit += o.encode_submit;
        gpu_wait += o.gpu_wait; text += o.text;
    }
};

struct CallerOptions {
    int device = 0;
    uint32_t tile_sites = 8192;
    uint32_t n_slots = 2;
    int em_abs_mode = BV_EM_ABS_INT_TRUNC;
    bool sparse_upload = true;   // tiles whose packer lists its covered cells cross PCIe as bv_sparse_tile (see TileRows)
    bool dense_rows = false;     // hand the packer dense planes to fill even when it lists its cells (tests: both forms of a tile)
    // Receives one line "CHROM\tPOS\tFLAG[,FLAG]\n" per covered position whose record carries BV_FLAG_NEAR_LRT (an LRT
    // statistic within 1e-9 of the threshold) or BV_FLAG_LRT_TIE (two candidate subsets tie to rounding): the positions
    // where the call can legitimately differ from the reference's, whose own choice hangs on rounding noise there.
    std::function<void(const char*, size_t)> flip_log;
    StageTimes* times = nullptr;   // when set, BasevarCaller adds its stages' wall time here
};

// ---- number formatting of the reference's text outputs ---------------------------------------------------------------
std::string to_string_f(double v);   // std::to_string(double): "%f"
std::string tostring_g(double v);    // ngslib::tostring / join (src/utils.h:37-85): ostream << double, 6 significant digits;
                                     // NaN prints "-nan" (x86 invalid-operation NaNs carry the sign bit)

// ---- one position as the host keeps it next to the packed planes -------------------------------------------------------
struct SiteMeta {
    std::string ref_id;
    std::string ref_base;     // column 3 verbatim (the reference prints the whole string)
    uint32_t ref_pos = 0;
    uint32_t depth = 0;       // sum of column 4 over the batchfiles
    // cells whose align_bases string is not one of "A","C","G","T","N": (sample index, string) -- indels ("+ACG",
    // "-AC") and other characters; needed for the Indels column of the CVG row and the AB field of the VCF row
    std::vector<std::pair<uint32_t, std::string>> specials;
    std::vector<std::pair<uint32_t, char>> odd_strands;   // strand characters other than '+', '-', '.'
};

// Row views of one site inside a tile (n_samples cells each).
struct SiteCells {
    const uint8_t* base;
    const uint8_t* qual;
    const uint8_t* strand;
    uint32_t n_samples;
};

// _out_cvg_line (src/basetype_caller.cpp:1211-1260): "" when the site has no A/C/G/T read.
std::string out_cvg_line(const SiteMeta& m, const SiteCells& c, const bv_site_out& rec);
// _out_vcf_line (src/basetype_caller.cpp:1103-1209).  group_names in ascending order, groups[g] belongs to group_names[g].
std::string out_vcf_line(const SiteMeta& m, const SiteCells& c, const bv_site_out& rec, const bv_call_out& call,
                         const std::vector<std::string>& group_names, const bv_group_out* groups);

// Direct row access for packers that fill a whole tile themselves (the BAM-driven packer, bv_pileup.hpp): the planes of the
// current tile, pre-filled with uncovered cells (N ! 0 0 .), and one SiteMeta per row for the caller to fill.  When the sparse
// transport is offered (site_start != nullptr) the planes may be NULL: the packer then lists its cells and nothing else.
struct TileRows {
    uint8_t *base, *qual, *strand, *mapq;
    uint16_t* rpr;
    uint64_t pitch, rpr_pitch;   // elements per row
    uint32_t n_rows;
    SiteMeta* meta;
    // Sparse transport (bv_sparse_tile), offered when CallerOptions::sparse_upload is set and the sample count allows it
    // (site_start != nullptr): a packer that knows its covered cells writes site_start[0 .. n_rows], asks reserve_cells(n)
    // for the two arrays of n words, fills them with BV_CELL_PACK / BV_CELL_AUX_PACK words grouped by row, and sets
    // *sparse_ready.  The tile then crosses PCIe as 8 bytes per covered cell instead of 5 bytes per sample-site.  The planes
    // above, when they are not NULL, are filled all the same.
    uint32_t* site_start;
    std::function<void(size_t n_cells, uint32_t** cells, uint32_t** aux)> reserve_cells;
    bool* sparse_ready;
};

std::string vcf_header_define(const std::vector<std::string>& contig_lines, const std::string& reference_line,
                              const std::vector<std::string>& addition_info, const std::vector<std::string>& samples);
std::string cvg_header_define();

class BasevarCaller {
public:
    // n_sample, group_smp_idx, min_af: as _basevar_caller's arguments (min_af is the CLI's float widened to double,
    // src/basetype_caller.cpp:122,506).  vcf / cvg receive the text the reference would bgzf_write to vcf_hd / cvg_hd.
    BasevarCaller(size_t n_sample, const std::map<std::string, std::vector<size_t>>& group_smp_idx, double min_af,
                  TextSink vcf, TextSink cvg, const CallerOptions& opt = CallerOptions());
    ~BasevarCaller();
    BasevarCaller(const BasevarCaller&) = delete;
    BasevarCaller& operator=(const BasevarCaller&) = delete;

    // One position: one row from each batchfile, in batchfile order (the smp_bf_line_vector of _basevar_caller).
    // Throws std::runtime_error with the reference's messages on malformed rows.  Text may be emitted for earlier tiles.
    void call(const std::vector<std::string>& smp_bf_line_vector);
    // Tile path: begin_tile(n) hands out n (<= tile_sites) rows of a free tile; the packer fills cells and meta (ref_id,
    // ref_pos, ref_base, depth, specials); commit_tile() sends them.  Rows whose depth stays 0 produce no output, as in
    // _basevar_caller (cpp:717-718).  Text of earlier tiles may be emitted by begin_tile().
    TileRows begin_tile(uint32_t n_rows);
    void commit_tile();
    // Drains the pipeline.  Returns true if any SNP row was written since construction (`has_data`, cpp:610).
    bool finish();

    uint64_t n_positions() const { return n_positions_; }   // positions with depth > 0 sent to the GPU
    uint64_t n_snps() const { return n_snps_; }
    uint64_t launch_count() const;

private:
    struct Tile;
    void call_row(const std::vector<std::string>& smp_bf_line_vector);
    void submit_current();
    void drain(uint32_t slot);

    size_t n_sample_;
    std::vector<std::string> group_names_;
    TextSink vcf_, cvg_;
    CallerOptions opt_;
    bv_ctx* ctx_ = nullptr;
    std::vector<std::unique_ptr<Tile>> tiles_;
    uint32_t cur_ = 0;
    uint64_t n_positions_ = 0, n_snps_ = 0;
};

// _variant_calling_unit (src/basetype_caller.cpp:529-635): the batchfiles are read in lock step, one row from each per
// position.  `region` ("chr", "chr:beg-end", or "" for everything) filters rows by coordinate (the reference asks the
// tabix index for the same rows).  Returns has_data.
bool variant_calling_unit(const std::vector<std::string>& batchfiles, const std::vector<std::string>& sample_ids,
                          const std::map<std::string, std::vector<size_t>>& group_smp_idx, double min_af,
                          const std::string& region, TextSink vcf, TextSink cvg, const CallerOptions& opt = CallerOptions());
// "##SampleIDs=" headers of the batchfiles, concatenated (_get_sampleid_from_batchfiles, cpp:637-665)
std::vector<std::string> get_sampleid_from_batchfiles(const std::vector<std::string>& batchfiles);

}  // namespace bvhost
