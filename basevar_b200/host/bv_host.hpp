// bv_host.hpp -- C++ host layer over the C ABI (include/basevar_b200.h): the reference-facing side of the drop-in.
//
// It mirrors the interface the reference's per-site driver uses (same names, argument meaning and error behaviour):
//
//   reference (src/basetype.h)                              here (namespace bvhost)
//   -----------------------------------------------------   ---------------------------------------------------------
//   struct BatchInfo                      :25-43            BatchInfo (same members)
//   struct StrandBiasInfo                 :56-61            StrandBiasInfo (same members)
//   class BaseType, ctor + lrt() + getters :64-153          BaseType: filled from a device record (batch path), or
//                                                           BaseType(const BatchInfo*, double) + lrt() for one site
//   strand_bias(ref, alts, bases, strands) :178-181         strand_bias(ref, alts, record)
//   BatchInfo rows -> BaseType per site,                    TilePacker (BatchInfo -> packed SoA planes in pinned memory)
//     src/basetype_caller.cpp:686-743                       Context (slots: H2D copy, kernels, D2H copy per tile)
//   100-kb tasks over a ThreadPool,                         shard_region / run_region: contiguous region shards, one host
//     src/basetype_caller.cpp:469-525                       worker + context per GPU, records delivered in coordinate order
//
// All the arithmetic happens on the GPU behind the C ABI; nothing here computes likelihoods, EM, LRT or Fisher.
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/basevar_b200.h"

namespace bvhost {

static const std::vector<char> BASES = {'A', 'C', 'G', 'T'};   // src/basetype.h:19
static const int LRT_THRESHOLD = 24;                           // src/basetype.h:21
static const int QUAL_THRESHOLD = 20;                          // src/basetype.h:22

struct BatchInfo {   // src/basetype.h:25-43
    size_t n;
    std::string ref_id;
    std::string ref_base;
    uint32_t ref_pos;
    uint32_t depth;
    std::vector<std::string> align_bases;
    std::vector<char> align_base_quals;
    std::vector<int> mapqs;
    std::vector<char> map_strands;
    std::vector<int> base_pos_ranks;
    BatchInfo() : n(0), ref_pos(0), depth(0) {}
};

struct StrandBiasInfo {   // src/basetype.h:56-61
    int ref_fwd, ref_rev;
    int alt_fwd, alt_rev;
    double fs;
    double sor;
};

// ---- cell encoding ---------------------------------------------------------------------------------------------------
// First character of align_bases[i] -> BV_BASE_* (src/basetype.cpp:50-56).  A multi-character string that is not an
// indel makes the reference throw; so does this.
uint8_t encode_base(const std::string& align_base);
inline uint8_t encode_strand(char c) { return c == '+' ? BV_STRAND_FWD : c == '-' ? BV_STRAND_REV : BV_STRAND_NONE; }
// min(100.0f / n_bam, min_af) in float (src/basetype_caller.cpp:122)
float cli_min_af(float min_af, size_t n_bam);

// ---- pinned planes + packer ------------------------------------------------------------------------------------------
class TilePacker {
public:
    TilePacker(uint32_t n_samples, uint32_t max_sites, bool pinned = true);
    ~TilePacker();
    TilePacker(const TilePacker&) = delete;
    TilePacker& operator=(const TilePacker&) = delete;

    void clear() { n_sites_ = 0; }
    uint32_t n_sites() const { return n_sites_; }
    uint32_t n_samples() const { return n_samples_; }
    uint32_t capacity() const { return max_sites_; }
    uint64_t pitch() const { return pitch_; }

    // One site from the reference's per-site struct (any type with BatchInfo's members).
    template <class BI>
    void add_site(const BI& bi) {
        if (bi.n != n_samples_) throw std::runtime_error("[ERROR] BatchInfo::n does not match the tile's sample count");
        uint8_t *b, *q, *s, *m;
        next_row(bi.ref_base.empty() ? 'N' : bi.ref_base[0], &b, &q, &s, &m);
        for (size_t i = 0; i < bi.n; ++i) {
            b[i] = encode_base(bi.align_bases[i]);
            q[i] = (uint8_t)(bi.align_base_quals[i] - 33);
            s[i] = encode_strand(bi.map_strands[i]);
            m[i] = (uint8_t)(bi.mapqs[i] > 255 ? 255 : bi.mapqs[i] < 0 ? 0 : bi.mapqs[i]);
        }
    }
    // One site of already encoded cells (row pointers of n_samples bytes each; mapq may be null).
    void add_site_cells(char ref_base, const uint8_t* base, const uint8_t* qual, const uint8_t* strand, const uint8_t* mapq);
    // Raw row access for packers that scatter per-sample runs (rows are pre-filled with uncovered cells).
    void next_row(char ref_base, uint8_t** base, uint8_t** qual, uint8_t** strand, uint8_t** mapq);

    bv_tile tile() const;   // BV_LOC_HOST tile over the packed sites
    const uint8_t* mapq_plane() const { return mapq_; }

private:
    uint32_t n_samples_, max_sites_, n_sites_ = 0;
    uint64_t pitch_;
    bool pinned_;
    uint8_t *base_ = nullptr, *qual_ = nullptr, *strand_ = nullptr, *mapq_ = nullptr, *ref_ = nullptr;
};

// ---- sparse packer: only the covered cells of a site travel (bv_sparse_tile) ----------------------------------------------
// What a pileup produces at < 1x depth: one packed word per (sample, position) a read covers.  A sample whose entry is
// the reference's filler for "no read here" (`N`, `!`, strand `.`; src/basetype_caller.cpp:1063-1075) adds nothing.
class SparsePacker {
public:
    SparsePacker(uint32_t n_samples, uint32_t max_sites, size_t reserve_cells = 0);
    ~SparsePacker();
    SparsePacker(const SparsePacker&) = delete;
    SparsePacker& operator=(const SparsePacker&) = delete;

    void clear() { n_sites_ = 0; n_cells_ = 0; }
    uint32_t n_sites() const { return n_sites_; }
    size_t n_cells() const { return n_cells_; }

    void begin_site(char ref_base);                                      // then add_cell() for its covered samples
    void add_cell(uint32_t sample, uint8_t base, uint8_t strand, uint8_t phred, uint8_t mapq = 0, uint16_t rpr = 0);
    template <class BI>
    void add_site(const BI& bi) {   // the reference's per-site struct (any type with BatchInfo's members)
        if (bi.n != n_samples_) throw std::runtime_error("[ERROR] BatchInfo::n does not match the tile's sample count");
        begin_site(bi.ref_base.empty() ? 'N' : bi.ref_base[0]);
        for (size_t i = 0; i < bi.n; ++i) {
            const uint8_t b = encode_base(bi.align_bases[i]), st = encode_strand(bi.map_strands[i]);
            const uint8_t q = (uint8_t)(bi.align_base_quals[i] - 33);
            const int mq = bi.mapqs[i] > 255 ? 255 : bi.mapqs[i] < 0 ? 0 : bi.mapqs[i];
            const int rp = bi.base_pos_ranks.empty() ? 0 : bi.base_pos_ranks[i];
            if (b == BV_BASE_N && q == 0 && st == BV_STRAND_NONE && mq == 0 && rp == 0) continue;
            add_cell((uint32_t)i, b, st, q, (uint8_t)mq, (uint16_t)(rp < 0 ? 0 : rp > 65535 ? 65535 : rp));
        }
    }
    bv_sparse_tile tile() const;   // valid until the next clear() / add

private:
    void grow(size_t want);
    uint32_t n_samples_, max_sites_, n_sites_ = 0;
    size_t n_cells_ = 0, cap_cells_ = 0;
    uint32_t *cells_ = nullptr, *aux_ = nullptr, *site_start_ = nullptr;   // pinned
    uint8_t* ref_ = nullptr;
};

// ---- a context bound to one GPU ---------------------------------------------------------------------------------------
class Context {
public:
    Context(int device, float min_af, uint32_t max_samples, uint32_t max_sites, uint32_t n_slots = 2,
            int em_abs_mode = BV_EM_ABS_INT_TRUNC);
    ~Context();
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;

    void submit(int slot, const bv_tile& tile);          // asynchronous; throws std::runtime_error on failure
    void wait(int slot, bv_site_out* out);               // blocks; `out` receives the tile's records
    std::vector<bv_site_out> run(const bv_tile& tile);   // submit + wait on slot 0
    void submit(int slot, const bv_sparse_tile& tile);   // sparse transport of the same tile (kernel K0 expands it)
    std::vector<bv_site_out> run(const bv_sparse_tile& tile);
    uint32_t n_slots() const { return n_slots_; }
    uint64_t launch_count() const;
    bv_ctx* raw() { return ctx_; }
    uint32_t max_samples() const { return max_samples_; }
    // FS of free-standing 2x2 strand tables {ref_fwd, ref_rev, alt_fwd, alt_rev} on the device (bv_fisher_fs)
    double fisher_fs(int ref_fwd, int ref_rev, int alt_fwd, int alt_rev);

private:
    bv_ctx* ctx_ = nullptr;
    uint32_t n_slots_;
    uint32_t max_samples_ = 0;
};

// ---- per-site mirror of the reference's class --------------------------------------------------------------------------
class BaseType {
public:
    BaseType() {}
    // Batch path: the record the device produced for this site.
    BaseType(const BatchInfo* smp_bi, const bv_site_out& rec);
    // Drop-in path (src/basetype.h:105): one site through a one-row tile on the calling thread's default context
    // (device BASEVAR_B200_DEVICE or 0).  `af` is the CLI's float min_af widened to double (basetype_caller.cpp:506).
    BaseType(const BatchInfo* smp_bi, double af);

    void lrt() { lrt(BASES); }                              // src/basetype.h:117
    void lrt(const std::vector<char>& specific_bases);      // only the full set {A,C,G,T} is computed on the device

    bool is_only_snp() const { return true; }
    const std::string& get_ref_id() const { return ref_id_; }
    const uint32_t& get_ref_pos() const { return ref_pos_; }
    const std::string& get_ref_base() const { return ref_base_; }
    const std::vector<char>& get_alt_bases() const { return alt_bases_; }
    double get_var_qual() const { return var_qual_; }
    int get_total_depth() const { return total_depth_; }
    double get_base_depth(char b) const;   // throws std::runtime_error for a key that is not A/C/G/T (basetype.h:127-139)
    double get_lrt_af(char b) const;       // throws std::runtime_error for a base that is not an ALT (basetype.h:141-151)
    const bv_site_out& record() const { return rec_; }

private:
    void fill_from_record();
    std::string ref_id_, ref_base_;
    uint32_t ref_pos_ = 0;
    std::vector<char> alt_bases_;
    std::map<char, double> af_by_lrt_;
    double var_qual_ = 0.0;
    int total_depth_ = 0;
    bool lrt_done_ = false;
    bv_site_out rec_{};
};

// strand_bias(ref, ALT string) from a device record (src/basetype.cpp:244-295).  The counts come from the record's 2x4
// strand table for any ALT set; FS is read from the record for the two sets the reference asks for -- all non-REF bases (CVG
// row, basetype_caller.cpp:1236-1245) and the called ALT alleles (VCF row, :1164) -- and computed on the device for any other
// set (bv_fisher_fs; `ctx`, or the calling thread's default context, created on first use on device BASEVAR_B200_DEVICE or 0).
// A record flagged BV_FLAG_BAD_STRAND throws the reference's "[ERROR] Get strange strand symbol" error.
StrandBiasInfo strand_bias(const char ref_base, const std::string alt_bases_string, const bv_site_out& rec, Context* ctx = nullptr);

// ---- region sharding over the GPUs of one box ---------------------------------------------------------------------------
struct Shard {
    int gpu;
    uint64_t beg, end;   // half-open site range
};
// Contiguous shards, cut at multiples of `step` (the reference's 100-kb task length, basetype_caller.cpp:474) from
// reg_beg, sizes differing by at most one step; fewer shards than GPUs when the region is short.
std::vector<Shard> shard_region(uint64_t reg_beg, uint64_t reg_end, int n_gpus, uint64_t step = 100000);

// fill(site0, n, packer): add sites [site0, site0+n) to the (cleared) packer, in order.
using TileSource = std::function<void(uint64_t site0, uint32_t n_sites, TilePacker& into)>;
// sink(site0, records, n): called on the calling thread, in coordinate order over the whole region.
using RecordSink = std::function<void(uint64_t site0, const bv_site_out* recs, uint32_t n_sites)>;

struct RunOptions {
    float min_af = 0.01f;
    uint32_t n_samples = 0;
    uint32_t tile_sites = 16384;
    uint32_t n_slots = 2;
    int em_abs_mode = BV_EM_ABS_INT_TRUNC;
    std::vector<int> devices;   // device of shard i is devices[i % devices.size()]; empty: all visible devices
    int n_shards = 0;           // 0: one shard per device
};
// One host worker thread and one context per shard; tiles are packed while the previous one is on the device.
// No inter-GPU communication: shards are independent, the merge is the concatenation in coordinate order.
void run_region(uint64_t reg_beg, uint64_t reg_end, const RunOptions& opt, const TileSource& fill, const RecordSink& sink);

}  // namespace bvhost
