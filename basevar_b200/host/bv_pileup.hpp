// bv_pileup.hpp -- BAM-driven tile packer and the `basetype` runner above it.
//
// The reference assembles the pileup in two text hops: BAM -> per-sample position maps -> bgzipped text batchfiles
// (src/basetype_caller.cpp:800-1101), which `_variant_calling_unit` then re-parses (cpp:529-635, 686-736).  Here the
// alignments are decoded straight into the site-major planes of a pinned tile (base / qual / strand / mapq / rank per
// sample per site), the tile goes to the GPU, and the records come back as VCF / CVG rows: no intermediate files.
//
//   reference                                                here (namespace bvhost)
//   -------------------------------------------------------  -------------------------------------------------------
//   __fetch_base_in_region   cpp:876-938   (read filters)     pileup_sample(): one BAM over a span of positions
//   __seek_position          cpp:940-1021  (first read wins)
//   BamRecord::get_aligned_pairs  bam_record.cpp:217-283
//   __write_record_to_batchfile   cpp:1024-1101               BamPileup::scatter(): cells -> rows of a tile
//                                                             batchfile_row(): the same row as the reference's text (tests)
//   __create_a_batchfile     cpp:800-874   (500-kb steps)     BamPileup::load_span(); the 500-kb grid only survives as the
//                                                             rule for indels anchored on a step's last position
//   BaseTypeRunner           cpp:19-466, basetype_caller.h    BaseTypeRunner (same options, same output files)
//   merge_file_by_line       basetype_utils.cpp:90-123        rows are produced in coordinate order; one writer
//
// Behaviour kept on purpose: the first read (file order) that touches a (sample, position) wins, indel or not; reads
// with mapq < -q, duplicates and QC failures are skipped; an insertion / deletion is rolled back onto the base to its
// left and therefore nearly always loses to that base (cpp:977-1001); its quality is the read's mean quality; soft
// clips and padding advance the read coordinate; uncovered cells are `N ! 0 0 .`; REF is the FASTA character as it is.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "bv_bam.hpp"
#include "bv_caller.hpp"

namespace bvhost {

struct PileupCell {
    uint32_t off;        // position - span begin
    uint8_t base;        // BV_BASE_*
    uint8_t qual;        // phred
    uint8_t strand;      // BV_STRAND_*
    uint8_t mapq;
    uint16_t rpr;        // 1-based rank of the base in the read
    int32_t special;     // index into SamplePileup::specials, -1 for a plain A/C/G/T/N base
};

struct SamplePileup {
    std::vector<PileupCell> cells;        // sorted by off, at most one per position
    std::vector<std::string> specials;    // "+ACG" / "-AC" strings
};

// The reference's creation step length (cpp:810): indels whose anchor is the last position of a step are dropped.
static const uint32_t PILEUP_STEP_REGION_LEN = 500000;

// One sample over positions [span_beg, span_end] (1-based, inclusive) of the calling interval [reg_beg, reg_end] on
// contig `tid` of `bam`; fa_seq is the whole contig.  `occ` is scratch.
void pileup_sample(BamReader& bam, int tid, const std::string& fa_seq, uint32_t reg_beg, uint32_t reg_end,
                   uint32_t span_beg, uint32_t span_end, int mapq_thd, SamplePileup& out, std::vector<uint8_t>& occ);

class BamPileup {
public:
    // n_instances: how many BamPileup objects over the same files live at once (one per GPU shard): the descriptor budget
    // of the process is shared between them, so readers stay open only when n_files x n_instances descriptors fit.
    BamPileup(const std::vector<std::string>& bam_files, int mapq_thd, int n_threads, int n_instances = 1);
    ~BamPileup();
    size_t n_samples() const { return files_.size(); }
    // SM of the first @RG (bam_header.cpp:62-83) or the file name up to its first '.' (cpp:279-283)
    std::vector<std::string> sample_ids(bool filename_has_samplename);

    // Decode every sample over [span_beg, span_end]; returns false when no sample has a cell there.
    bool load_span(const std::string& ref_id, const std::string& fa_seq, uint32_t reg_beg, uint32_t reg_end,
                   uint32_t span_beg, uint32_t span_end);
    // Rows for positions [pos, pos + n) of the loaded span (rows must be pre-filled with uncovered cells).
    void scatter(uint32_t pos, uint32_t n, const std::string& ref_id, const std::string& fa_seq, TileRows& rows);

private:
    struct Impl;
    std::vector<std::string> files_;
    int mapq_thd_, n_threads_;
    std::unique_ptr<Impl> impl_;
};

// The row __write_record_to_batchfile would write for one site of a tile (cpp:1080-1086), without the newline.
std::string batchfile_row(const SiteMeta& m, const SiteCells& c, const uint8_t* mapq, const uint16_t* rpr);

// ---- the runner ----------------------------------------------------------------------------------------------------------
struct BaseTypeARGS {   // src/basetype_utils.h:74-96
    std::vector<std::string> input_bf;
    std::string in_bamfilelist, reference;
    float min_af = 0.01f;
    int mapq = 10;
    int batchcount = 200;   // accepted for compatibility: there are no batchfiles
    int thread_num = 4;
    std::string regions, pop_group_file, output_vcf, output_cvg;
    bool filename_has_samplename = false;
    bool smart_rerun = false;   // accepted for compatibility: there is nothing to resume from
    // additions
    std::vector<int> devices;   // GPUs to shard the calling intervals over (empty: device 0)
    uint32_t tile_sites = 8192;
    bool tile_sites_given = false;   // --tile-sites on the command line: taken as is; else lowered for very wide cohorts (run())
    int em_abs_mode = BV_EM_ABS_INT_TRUNC;
    bool dense_upload = false;   // upload the packed planes instead of the covered cells (bv_tile instead of bv_sparse_tile)
    std::string flip_log;        // file for the positions flagged NEAR_LRT / LRT_TIE (CHROM, POS, FLAGS); empty: count only
    bool timing = false;         // print the wall time per stage of the host pipeline (JSON, one line on stderr)
    int workers_per_gpu = 0;     // host workers (each with its own context and region shard) per GPU; 0: max(1, thread_num / 2)
};

class BaseTypeRunner {
public:
    BaseTypeRunner() {}
    BaseTypeRunner(int argc, char* argv[]) { set_arguments(argc, argv); }
    explicit BaseTypeRunner(const BaseTypeARGS& args) { set_arguments(args); }
    static std::string usage();
    void set_arguments(int argc, char* argv[]);   // same options as `basevar basetype` (cpp:19-142)
    void set_arguments(const BaseTypeARGS& args);
    void run();                                   // writes --output-vcf and --output-cvg

    const std::vector<std::string>& samples_id() const { return samples_id_; }
    const std::vector<std::tuple<std::string, uint32_t, uint32_t>>& calling_intervals() const { return intervals_; }
    uint64_t launch_count() const { return launches_; }

private:
    void finish_arguments();
    BaseTypeARGS args_;
    std::unique_ptr<Fasta> reference_;
    std::vector<std::string> samples_id_;
    std::map<std::string, std::vector<size_t>> groups_idx_;
    std::vector<std::tuple<std::string, uint32_t, uint32_t>> intervals_;
    uint64_t launches_ = 0;
    StageTimes times_;   // summed over the host workers
};

// BGZF output (".gz" outputs, basetype_utils.cpp:97) or plain text.
class TextWriter {
public:
    explicit TextWriter(const std::string& path);
    ~TextWriter();
    void write(const char* data, size_t n);
    void close();

private:
    void flush_block(const uint8_t* p, size_t n);
    std::string path_;
    FILE* f_ = nullptr;
    bool gz_ = false;
    std::vector<uint8_t> pend_;
};

}  // namespace bvhost
