// bv_bam.hpp -- alignment and reference input for the BAM-driven tile packer.
//
// The reference reads BAM and FASTA through htslib (src/bam.cpp, src/bam_record.cpp, src/bam_header.cpp, src/fasta.cpp).
// htslib is not part of this repository; these are self-contained readers of the published formats (SAM/BAM
// specification v1: BGZF section 4.1, BAM section 4.2, BAI section 5.2; samtools faidx .fai), over zlib only,
// providing exactly what the pileup needs:
//
//   reference                                                here (namespace bvhost)
//   -------------------------------------------------------  ----------------------------------------------------
//   ngslib::Bam(fn, "r") + fetch(region) + next(al)          BamReader::open / query(tid, beg0, end0) / next(rec)
//     src/bam.cpp:86-138  (sam_itr_querys / sam_itr_next)      same records, in file order: tid matches, pos < end0,
//                                                              bam_endpos > beg0
//   ngslib::BamHeader::get_sample_name  bam_header.cpp:62-83  BamReader::sample_name(): SM of the first @RG line
//   BamRecord::mapq / is_duplicate / is_qc_fail /             BamRec fields + helpers below
//     map_strand / map_ref_start_pos / map_ref_end_pos
//     src/bam_record.h:132-247
//   ngslib::Fasta(fn), operator[](ref_id), nseq, iseq_name,   Fasta: plain-text or BGZF-compressed FASTA (bgzip, the only
//     seq_length   src/fasta.cpp:17-95                         compressed form faidx accepts) + .fai; the index is computed
//                                                              in memory when the .fai file is missing, and the block table
//                                                              of a compressed file (what .gzi stores) is always read off the
//                                                              BGZF block headers
//
// Not supported (the constructor / open() throws): CRAM and SAM input, plain-gzip FASTA.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <zlib.h>

namespace bvhost {

// ---- BGZF ---------------------------------------------------------------------------------------------------------------
class BgzfReader {
public:
    explicit BgzfReader(const std::string& path);
    ~BgzfReader();
    BgzfReader(const BgzfReader&) = delete;
    BgzfReader& operator=(const BgzfReader&) = delete;

    void seek(uint64_t voffset);           // virtual offset: compressed block address << 16 | offset inside the block
    uint64_t tell() const;                 // virtual offset of the next byte (start of the next block when one is used up)
    size_t read(void* dst, size_t n);      // bytes read; fewer than n only at end of file
    const std::string& path() const { return path_; }
    // (compressed address, uncompressed start offset) of every non-empty block, from the block headers alone (no inflate):
    // what a .gzi index stores.  throws when the file is not BGZF.
    std::vector<std::pair<uint64_t, uint64_t>> block_table() const;

private:
    bool load_block();                     // the block at next_addr_; false at end of file
    std::string path_;
    int fd_ = -1;
    uint64_t block_addr_ = 0, next_addr_ = 0;
    std::vector<uint8_t> cbuf_, ubuf_;
    size_t ulen_ = 0, upos_ = 0;
    z_stream zs_;
    bool zs_ready_ = false;
};

// ---- BAM ----------------------------------------------------------------------------------------------------------------
enum : uint16_t { BAM_FLAG_REVERSE = 0x10, BAM_FLAG_UNMAP = 0x4, BAM_FLAG_QCFAIL = 0x200, BAM_FLAG_DUP = 0x400 };
enum : uint8_t { CIG_M = 0, CIG_I = 1, CIG_D = 2, CIG_N = 3, CIG_S = 4, CIG_H = 5, CIG_P = 6, CIG_EQ = 7, CIG_X = 8 };

struct BamRec {
    int32_t tid = -1;
    int32_t pos = -1;        // 0-based leftmost coordinate
    int32_t end = 0;         // bam_endpos: pos + reference length of the CIGAR; pos + 1 when that is 0 or the read is unmapped
    uint8_t mapq = 0;
    uint16_t flag = 0;
    int32_t l_seq = 0;
    std::vector<uint32_t> cigar;   // len << 4 | op
    std::vector<uint8_t> seq;      // 4-bit codes, two per byte, high nibble first
    std::vector<uint8_t> qual;     // phred, 0xff when absent
    bool is_mapped() const { return (flag & BAM_FLAG_UNMAP) == 0; }
    int seqi(int i) const { return (seq[(size_t)i >> 1] >> ((~i & 1) << 2)) & 0xf; }
};

class BamReader {
public:
    explicit BamReader(const std::string& path);   // reads the header; the index is loaded by the first query()
    const std::string& path() const { return bgzf_.path(); }
    const std::string& header_text() const { return text_; }
    int n_ref() const { return (int)ref_names_.size(); }
    const std::string& ref_name(int tid) const { return ref_names_[(size_t)tid]; }
    int64_t ref_length(int tid) const { return ref_lens_[(size_t)tid]; }
    int name2id(const std::string& name) const;    // -1 when the header does not list it
    bool sample_name(std::string& out) const;      // SM of the first @RG line; false when that line has none (or there is no @RG)

    // Records overlapping [beg0, end0) on tid, through the index: fn + ".bai", fn with .bam replaced by .bai, or the same
    // two names with .csi.
    void query(int tid, int64_t beg0, int64_t end0);
    // Without query(): every record of the file in order.  Returns false when the iteration is over.
    bool next(BamRec& rec);

private:
    struct Chunk { uint64_t beg, end; };
    struct RefIndex {
        std::vector<std::pair<uint32_t, std::vector<Chunk>>> bins;   // sorted by bin number
        std::vector<uint64_t> linear;
    };
    void load_index();
    bool read_record(BamRec& rec);

    BgzfReader bgzf_;
    std::string text_;
    std::vector<std::string> ref_names_;
    std::vector<int64_t> ref_lens_;
    std::vector<RefIndex> index_;
    bool index_loaded_ = false;
    int idx_min_shift_ = 14, idx_depth_ = 5;   // BAI geometry; a CSI index brings its own
    uint64_t first_record_voffset_ = 0;
    // iterator state
    bool querying_ = false, finished_ = false;
    int q_tid_ = -1;
    int64_t q_beg_ = 0, q_end_ = 0;
    std::vector<Chunk> chunks_;
    size_t chunk_i_ = 0;
    bool in_chunk_ = false;
    std::vector<uint8_t> buf_;
};

// ---- FASTA --------------------------------------------------------------------------------------------------------------
class Fasta {
public:
    explicit Fasta(const std::string& path);
    size_t nseq() const { return names_.size(); }
    const std::string& iseq_name(size_t i) const { return names_[i]; }
    bool has_seq(const std::string& name) const;
    uint32_t seq_length(const std::string& name) const;    // throws std::invalid_argument for an unknown name
    std::string fetch(const std::string& name) const;      // the whole sequence, characters as they are in the file
    const std::string& path() const { return path_; }

private:
    struct Entry { uint64_t length, offset, line_bases, line_width; };
    const Entry& entry(const std::string& name) const;
    void build_index();
    size_t read_at(void* dst, size_t n, uint64_t off) const;   // uncompressed bytes [off, off + n) of the file
    std::string path_;
    bool bgzf_ = false;
    std::vector<std::pair<uint64_t, uint64_t>> blocks_;   // BGZF: (compressed address, uncompressed offset) per block
    std::vector<std::string> names_;
    std::vector<Entry> entries_;
};

}  // namespace bvhost
