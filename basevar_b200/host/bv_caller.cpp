// bv_caller.cpp -- see bv_caller.hpp.  Text in (batchfile rows), text out (VCF / CVG rows); the numbers come from the GPU.
#include "bv_caller.hpp"

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <stdexcept>

namespace bvhost {

// ---- number formatting ---------------------------------------------------------------------------------------------------
std::string to_string_f(double v) {
    char buf[400];
    const int n = snprintf(buf, sizeof(buf), "%f", v);
    return std::string(buf, (size_t)n);
}

std::string tostring_g(double v) {
    if (std::isnan(v)) return "-nan";   // 0/0 on x86 gives the NaN with the sign bit set; ostream prints its sign
    char buf[64];
    const int n = snprintf(buf, sizeof(buf), "%g", v);
    return std::string(buf, (size_t)n);
}

static const char kBaseChars[4] = {'A', 'C', 'G', 'T'};
static int code_of_char(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

// std::to_string(1.0 - exp((q - 33) * MLN10TO10)) for every quality character (src/basetype.cpp:47-48, the BP field
// of src/basetype_caller.cpp:1141): a function of the character only, formatted once
static const std::string* bp_text_table() {
    static std::string table[256];
    static bool done = false;
    if (!done) {
        const double MLN10TO10 = -0.23025850929940458;   // src/basetype.h:20
        for (int p = 0; p < 256; ++p) table[p] = to_string_f(1.0 - exp((double)p * MLN10TO10));
        done = true;
    }
    return table;
}

static const std::string* special_of(const SiteMeta& m, uint32_t sample) {
    for (const auto& s : m.specials)
        if (s.first == sample) return &s.second;
    return nullptr;
}

static void throw_bad_strand(const SiteMeta& m, const SiteCells& c) {
    char sym = '.';
    for (uint32_t i = 0; i < c.n_samples; ++i) {
        if (c.base[i] < BV_BASE_N && c.strand[i] == BV_STRAND_NONE) {
            for (const auto& o : m.odd_strands)
                if (o.first == i) sym = o.second;
            break;
        }
    }
    throw std::runtime_error(std::string("[ERROR] Get strange strand symbol: ") + sym);   // src/basetype.cpp:271-273
}

struct StrandTable {
    int ref_fwd, ref_rev, alt_fwd, alt_rev;
    double sor;
};
// the counts of strand_bias() for an ALT set (bit mask over A,C,G,T) and SOR (src/basetype.cpp:251-286)
static StrandTable strand_table(const bv_site_out& rec, int ref_code, unsigned alt_mask) {
    StrandTable t;
    t.ref_fwd = ref_code >= 0 ? (int)rec.fwd[ref_code] : 0;
    t.ref_rev = ref_code >= 0 ? (int)rec.rev[ref_code] : 0;
    t.alt_fwd = t.alt_rev = 0;
    for (int b = 0; b < 4; ++b)
        if (b != ref_code && (alt_mask >> b & 1u)) { t.alt_fwd += (int)rec.fwd[b]; t.alt_rev += (int)rec.rev[b]; }
    t.sor = (t.ref_rev * t.alt_fwd > 0) ? (double)(t.ref_fwd * t.alt_rev) / (double)(t.ref_rev * t.alt_fwd) : 10000;
    return t;
}

// ---- _out_cvg_line ---------------------------------------------------------------------------------------------------------
std::string out_cvg_line(const SiteMeta& m, const SiteCells& c, const bv_site_out& rec) {
    if (rec.flags & BV_FLAG_BAD_STRAND) throw_bad_strand(m, c);
    // __base_depth_and_indel (cpp:1262-1289): strings that do not start with A/C/G/T/N are listed as "indels"
    std::string indel_string = ".";
    if (!m.specials.empty()) {
        std::map<std::string, int> indel_depth;
        for (const auto& s : m.specials) indel_depth[s.second]++;
        indel_string.clear();
        for (const auto& kv : indel_depth) {
            if (!indel_string.empty()) indel_string += ",";
            indel_string += kv.first + "|" + std::to_string(kv.second);
        }
    }
    const int total_depth = (int)(rec.depth[0] + rec.depth[1] + rec.depth[2] + rec.depth[3]);
    if (total_depth <= 0) return std::string();
    const int ref_code = code_of_char((char)toupper((unsigned char)m.ref_base[0]));
    const StrandTable t = strand_table(rec, ref_code, 0xfu);
    std::string out;
    out.reserve(96 + m.ref_id.size() + indel_string.size());
    out += m.ref_id; out += '\t'; out += std::to_string(m.ref_pos); out += '\t';
    out += m.ref_base; out += '\t'; out += std::to_string(total_depth); out += '\t';
    for (int b = 0; b < 4; ++b) { out += std::to_string((int)rec.depth[b]); out += '\t'; }
    out += indel_string; out += '\t';
    out += to_string_f(rec.fs_cvg); out += '\t'; out += to_string_f(t.sor); out += '\t';
    out += std::to_string(t.ref_fwd); out += ','; out += std::to_string(t.ref_rev); out += ',';
    out += std::to_string(t.alt_fwd); out += ','; out += std::to_string(t.alt_rev); out += '\n';
    return out;
}

// ---- _out_vcf_line ---------------------------------------------------------------------------------------------------------
std::string out_vcf_line(const SiteMeta& m, const SiteCells& c, const bv_site_out& rec, const bv_call_out& call,
                         const std::vector<std::string>& group_names, const bv_group_out* groups) {
    const int n_alt = rec.n_alt < 4 ? rec.n_alt : 4;
    const int total_depth = (int)(rec.depth[0] + rec.depth[1] + rec.depth[2] + rec.depth[3] + rec.depth_other);
    std::string alt_gt[4];     // genotype text of A,C,G,T reads: "./k" for the k-th ALT, "./." for anything else
    for (int b = 0; b < 4; ++b) alt_gt[b] = "./.";
    std::string cm_ac, cm_af, cm_caf, alt_list;
    double ad_sum = 0;
    unsigned alt_mask = 0;
    for (int i = 0; i < n_alt; ++i) {
        const int b = rec.alt[i] & 3;
        alt_gt[b] = "./" + std::to_string(i + 1);
        alt_mask |= 1u << b;
        ad_sum = ad_sum + (double)rec.depth[b];
        if (i) { cm_ac += ','; cm_af += ','; cm_caf += ','; alt_list += ','; }
        cm_ac += std::to_string((int)rec.depth[b]);
        cm_af += tostring_g(rec.af[i]);
        cm_caf += tostring_g((double)rec.depth[b] / total_depth);
        alt_list += kBaseChars[b];
    }
    const char upper_ref_base = (char)toupper((unsigned char)m.ref_base[0]);
    const int ref_code = code_of_char(upper_ref_base);

    double qd = rec.qual / ad_sum;
    if (qd == 0) qd = 0.0;   // -0.0 => 0.0
    if (rec.flags & BV_FLAG_BAD_STRAND) throw_bad_strand(m, c);
    const StrandTable t = strand_table(rec, ref_code, alt_mask);

    std::string out;
    out.reserve(512 + (size_t)c.n_samples * 8);
    out += m.ref_id; out += '\t'; out += std::to_string(m.ref_pos); out += "\t.\t"; out += m.ref_base; out += '\t';
    out += alt_list; out += '\t'; out += to_string_f(rec.qual); out += '\t';
    out += (rec.qual > QUAL_THRESHOLD) ? "." : "LowQual";
    out += '\t';
    out += "CM_DP=" + std::to_string(total_depth);
    out += ";CM_AC=" + cm_ac;
    out += ";CM_AF=" + cm_af;
    out += ";CM_CAF=" + cm_caf;
    out += ";MQRankSum=" + std::to_string(call.mq_rank_sum);
    out += ";ReadPosRankSum=" + std::to_string(call.read_pos_rank_sum);
    out += ";BaseQRankSum=" + std::to_string(call.base_q_rank_sum);
    out += ";QD=" + to_string_f(qd);
    out += ";SOR=" + to_string_f(t.sor);
    out += ";FS=" + to_string_f(rec.fs_vcf);
    out += ";SB_REF=" + std::to_string(t.ref_fwd) + "," + std::to_string(t.ref_rev);
    out += ";SB_ALT=" + std::to_string(t.alt_fwd) + "," + std::to_string(t.alt_rev);
    for (size_t g = 0; g < group_names.size(); ++g) {   // groupID_AF=xxx,xxx, only for groups that report an ALT
        const bv_group_out& go = groups[g];
        if (go.n_alt == 0) continue;
        out += ";" + group_names[g] + "_AF=";
        for (int k = 0; k < go.n_alt && k < 4; ++k) {
            if (k) out += ',';
            out += tostring_g(go.af[k]);
        }
    }
    out += "\tGT:AB:SO:BP";
    // per-sample GT:AB:SO:BP (cpp:1125-1146)
    const std::string* bp = bp_text_table();
    for (uint32_t i = 0; i < c.n_samples; ++i) {
        const uint8_t b = c.base[i];
        if (b >= BV_BASE_N) { out += "\t./."; continue; }   // 'N' or indel
        char fb;
        const std::string* gt;
        static const std::string kRefGt = "0/.", kNoGt = "./.";
        if (b < 4) {
            fb = kBaseChars[b];
            gt = (fb == upper_ref_base) ? &kRefGt : &alt_gt[b];
        } else {   // another character: never an ALT
            const std::string* s = special_of(m, i);
            fb = s && !s->empty() ? (*s)[0] : '?';
            gt = (fb == upper_ref_base) ? &kRefGt : &kNoGt;
        }
        out += '\t'; out += *gt; out += ':'; out += fb; out += ':';
        char st = c.strand[i] == BV_STRAND_FWD ? '+' : c.strand[i] == BV_STRAND_REV ? '-' : '.';
        out += st; out += ':'; out += bp[c.qual[i]];
    }
    out += '\n';
    return out;
}

// ---- headers (src/basetype_utils.cpp:32-88) -----------------------------------------------------------------------------------
std::string vcf_header_define(const std::vector<std::string>& contig_lines, const std::string& reference_line,
                              const std::vector<std::string>& addition_info, const std::vector<std::string>& samples) {
    std::vector<std::string> header = {
        "##fileformat=VCFv4.2",
        "##FILTER=<ID=LowQual,Description=\"Low quality (QUAL < 60)\">",
        "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">",
        "##FORMAT=<ID=AB,Number=1,Type=String,Description=\"Allele Base\">",
        "##FORMAT=<ID=SO,Number=1,Type=String,Description=\"Strand orientation of the mapping base. Marked as + or -\">",
        "##FORMAT=<ID=BP,Number=1,Type=String,Description=\"Base Probability which calculate by base quality\">",
        "##INFO=<ID=CM_AF,Number=A,Type=Float,Description=\"An ordered, comma delimited list of allele frequencies base on LRT algorithm\">",
        "##INFO=<ID=CM_CAF,Number=A,Type=Float,Description=\"An ordered, comma delimited list of allele frequencies just base on read count\">",
        "##INFO=<ID=CM_AC,Number=A,Type=Integer,Description=\"An ordered, comma delimited allele depth in CMDB\">",
        "##INFO=<ID=CM_DP,Number=A,Type=Integer,Description=\"Total Depth in CMDB\">",
        "##INFO=<ID=SB_REF,Number=A,Type=Integer,Description=\"Read number support REF: Forward,Reverse\">",
        "##INFO=<ID=SB_ALT,Number=A,Type=Integer,Description=\"Read number support ALT: Forward,Reverse\">",
        "##INFO=<ID=FS,Number=1,Type=Float,Description=\"Phred-scaled p-value using Fisher's exact test to detect strand bias\">",
        "##INFO=<ID=BaseQRankSum,Number=1,Type=Float,Description=\"Phred-score from Wilcoxon rank sum test of Alt Vs. Ref base qualities\">",
        "##INFO=<ID=SOR,Number=1,Type=Float,Description=\"Symmetric Odds Ratio of 2x2 contingency table to detect strand bias\">",
        "##INFO=<ID=MQRankSum,Number=1,Type=Float,Description=\"Phred-score From Wilcoxon rank sum test of Alt vs. Ref read mapping qualities\">",
        "##INFO=<ID=ReadPosRankSum,Number=1,Type=Float,Description=\"Phred-score from Wilcoxon rank sum test of Alt vs. Ref read position bias\">",
        "##INFO=<ID=QD,Number=1,Type=Float,Description=\"Variant Confidence Quality by Depth\">"};
    header.insert(header.end(), addition_info.begin(), addition_info.end());
    header.insert(header.end(), contig_lines.begin(), contig_lines.end());
    if (!reference_line.empty()) header.push_back(reference_line);
    std::string cols = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
    for (const auto& s : samples) cols += "\t" + s;
    header.push_back(cols);
    std::string out;
    for (size_t i = 0; i < header.size(); ++i) {
        if (i) out += '\n';
        out += header[i];
    }
    return out;
}

std::string cvg_header_define() {
    return "##fileformat=CVGv1.0\n##Group information is the depth of A:C:G:T:Indel\n"
           "#CHROM\tPOS\tREF\tDepth\tA\tC\tG\tT\tIndels\tFS\tSOR\tStrand_Coverage(REF_FWD,REF_REV,ALT_FWD,ALT_REV)";
}

// ---- tiles ------------------------------------------------------------------------------------------------------------------
static void* pinned(size_t bytes) {
    void* p = nullptr;
    if (bv_host_alloc(&p, bytes ? bytes : 16) != BV_OK)
        throw std::runtime_error(std::string("[ERROR] bv_host_alloc: ") + bv_last_error(nullptr));
    return p;
}

struct BasevarCaller::Tile {
    uint8_t *base = nullptr, *qual = nullptr, *strand = nullptr, *mapq = nullptr, *ref = nullptr;
    uint16_t* rpr = nullptr;
    uint64_t pitch = 0, rpr_pitch = 0;
    uint32_t n_sites = 0;
    bool pending = false;
    std::vector<SiteMeta> meta;
    std::vector<bv_site_out> recs;
    std::vector<bv_call_out> calls;
    std::vector<bv_group_out> groups;
    std::vector<int32_t> call_of_site;
    // sparse transport of the tile in hand (pinned; filled by the packer through TileRows)
    uint32_t *sp_cells = nullptr, *sp_aux = nullptr, *sp_start = nullptr;
    size_t sp_cap = 0;
    bool sparse_ready = false;
    // the same cells in the compact form (BV_CELLS_U16), encoded at submit time
    uint16_t* sp_words16 = nullptr;
    uint32_t *sp_aux16 = nullptr, *sp_start16 = nullptr;
    size_t sp_cap16 = 0;

    void reserve_words16(size_t n) {
        if (n <= sp_cap16) return;
        bv_host_free(sp_words16); bv_host_free(sp_aux16);
        sp_words16 = nullptr; sp_aux16 = nullptr; sp_cap16 = 0;
        const size_t cap = n + n / 4 + 4096;
        sp_words16 = (uint16_t*)pinned(cap * sizeof(uint16_t)); sp_aux16 = (uint32_t*)pinned(cap * sizeof(uint32_t));
        sp_cap16 = cap;
    }

    void reserve_cells(size_t n) {   // contents are not kept: called once per tile, before the cells are written
        if (n <= sp_cap) return;
        bv_host_free(sp_cells); bv_host_free(sp_aux);
        sp_cells = sp_aux = nullptr; sp_cap = 0;
        const size_t cap = n + n / 4 + 4096;
        sp_cells = (uint32_t*)pinned(cap * sizeof(uint32_t)); sp_aux = (uint32_t*)pinned(cap * sizeof(uint32_t));
        sp_cap = cap;
    }

    // The dense rows of the tile on the host (pinned: they are uploaded as they are when the tile is not sparse).  A packer that
    // lists its covered cells does not need them: they are allocated the first time somebody asks for dense rows.
    uint32_t max_sites = 0;
    bool dense = false;        // the current tile's rows are in the planes (else only its sparse cells exist)
    void ensure_dense() {
        if (base) return;
        const size_t plane = (size_t)max_sites * pitch;
        base = (uint8_t*)pinned(plane); qual = (uint8_t*)pinned(plane); strand = (uint8_t*)pinned(plane);
        mapq = (uint8_t*)pinned(plane); rpr = (uint16_t*)pinned(plane * 2);
    }
    // one site's row rebuilt from the sparse cells (text of a called site): N ! . everywhere but at the site's cells
    std::vector<uint8_t> row_base, row_qual, row_strand;
    SiteCells site_cells(uint32_t i, uint32_t n_samples) {
        if (dense) return SiteCells{base + (size_t)i * pitch, qual + (size_t)i * pitch, strand + (size_t)i * pitch, n_samples};
        if (row_base.empty()) { row_base.assign(pitch, BV_BASE_N); row_qual.assign(pitch, 0); row_strand.assign(pitch, BV_STRAND_NONE); }
        for (uint32_t c = sp_start[i]; c < sp_start[i + 1]; ++c) {
            const uint32_t w = sp_cells[c], smp = w & (BV_CELL_MAX_SAMPLES - 1u);
            row_base[smp] = (uint8_t)((w >> 20) & 7u); row_strand[smp] = (uint8_t)((w >> 23) & 3u); row_qual[smp] = (uint8_t)(w >> 25);
        }
        return SiteCells{row_base.data(), row_qual.data(), row_strand.data(), n_samples};
    }
    void release_site_cells(uint32_t i) {   // back to the uncovered row, touching the site's cells only
        if (dense) return;
        for (uint32_t c = sp_start[i]; c < sp_start[i + 1]; ++c) {
            const uint32_t smp = sp_cells[c] & (BV_CELL_MAX_SAMPLES - 1u);
            row_base[smp] = BV_BASE_N; row_strand[smp] = BV_STRAND_NONE; row_qual[smp] = 0;
        }
    }

    Tile(uint32_t n_samples, uint32_t max_sites_, size_t n_groups) : max_sites(max_sites_) {
        pitch = ((uint64_t)n_samples + 15) / 16 * 16;
        rpr_pitch = pitch;
        ref = (uint8_t*)pinned(max_sites);
        sp_start = (uint32_t*)pinned(((size_t)max_sites + 1) * sizeof(uint32_t));
        sp_start16 = (uint32_t*)pinned(((size_t)max_sites + 1) * sizeof(uint32_t));
        meta.resize(max_sites);
        recs.resize(max_sites);
        calls.resize(max_sites);
        groups.resize((size_t)max_sites * n_groups);
        call_of_site.resize(max_sites);
    }
    ~Tile() {
        bv_host_free(base); bv_host_free(qual); bv_host_free(strand); bv_host_free(mapq); bv_host_free(rpr); bv_host_free(ref);
        bv_host_free(sp_start); bv_host_free(sp_cells); bv_host_free(sp_aux);
        bv_host_free(sp_start16); bv_host_free(sp_words16); bv_host_free(sp_aux16);
    }
};

static void check(int rc, bv_ctx* ctx, const char* what) {
    if (rc != BV_OK) throw std::runtime_error(std::string("[ERROR] ") + what + ": " + bv_last_error(ctx));
}

BasevarCaller::BasevarCaller(size_t n_sample, const std::map<std::string, std::vector<size_t>>& group_smp_idx, double min_af,
                             TextSink vcf, TextSink cvg, const CallerOptions& opt)
    : n_sample_(n_sample), vcf_(std::move(vcf)), cvg_(std::move(cvg)), opt_(opt) {
    if (n_sample == 0) throw std::invalid_argument("[ERROR] n_sample must be > 0");
    if (opt_.tile_sites == 0) opt_.tile_sites = 1;
    if (opt_.n_slots == 0) opt_.n_slots = 1;
    if (group_smp_idx.size() > BV_MAX_GROUPS) throw std::invalid_argument("[ERROR] more population groups than the device path supports");
    std::vector<uint8_t> sample_group(n_sample, (uint8_t)BV_GROUP_NONE);
    for (const auto& kv : group_smp_idx) {   // std::map order = the order the reference iterates the groups in
        for (size_t i : kv.second) {
            if (i >= n_sample) throw std::invalid_argument("[ERROR] population group holds a sample index out of range");
            if (sample_group[i] != BV_GROUP_NONE) throw std::invalid_argument("[ERROR] a sample belongs to two population groups");
            sample_group[i] = (uint8_t)group_names_.size();
        }
        group_names_.push_back(kv.first);
    }
    bv_params p;
    memset(&p, 0, sizeof(p));
    p.min_af = (float)min_af;   // the double is a widened float (src/basetype_caller.cpp:122,506)
    p.lrt_threshold = LRT_THRESHOLD;
    p.em_max_iter = 100;
    p.em_eps = 0.001f;
    p.em_abs_mode = opt_.em_abs_mode;
    p.max_samples = (uint32_t)n_sample;
    p.max_sites = opt_.tile_sites;
    p.n_slots = opt_.n_slots;
    check(bv_create(opt_.device, &p, &ctx_), nullptr, "bv_create");
    try {
        if (!group_names_.empty())
            check(bv_set_groups(ctx_, sample_group.data(), (uint32_t)n_sample, (uint32_t)group_names_.size()), ctx_, "bv_set_groups");
        for (uint32_t s = 0; s < opt_.n_slots; ++s)
            tiles_.emplace_back(new Tile((uint32_t)n_sample, opt_.tile_sites, group_names_.size()));
    } catch (...) {
        tiles_.clear();
        bv_destroy(ctx_);
        throw;
    }
}

BasevarCaller::~BasevarCaller() {
    // in-flight tiles read the pinned planes: wait before freeing them
    for (uint32_t s = 0; s < tiles_.size(); ++s)
        if (tiles_[s]->pending) bv_tile_wait(ctx_, (int)s, nullptr);
    bv_destroy(ctx_);
    tiles_.clear();
}

uint64_t BasevarCaller::launch_count() const { return bv_launch_count(ctx_); }

// ---- row parsing (src/basetype_caller.cpp:686-736) -------------------------------------------------------------------------------
namespace {

// ngslib::split semantics (src/utils.h:90-122, src/utils.cpp:81-99): every delimiter ends a token, tokens may be empty
template <class F>
inline size_t for_each_token(const char* p, const char* end, char delim, F&& f) {
    size_t n = 0;
    for (;;) {
        const char* e = (const char*)memchr(p, delim, (size_t)(end - p));
        const char* tok_end = e ? e : end;
        f(n, p, tok_end);
        ++n;
        if (!e) break;
        p = e + 1;
    }
    return n;
}

// istringstream >> int on one token: leading whitespace skipped, junk gives 0
inline long parse_int(const char* p, const char* e) {
    while (p < e && isspace((unsigned char)*p)) ++p;
    bool neg = false;
    if (p < e && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    long v = 0;
    bool any = false;
    while (p < e && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); ++p; any = true; if (v > 100000000L) break; }
    if (!any) return 0;
    return neg ? -v : v;
}

inline char parse_char(const char* p, const char* e) {   // istringstream >> char: first non-space character, 0 if none
    while (p < e && isspace((unsigned char)*p)) ++p;
    return p < e ? *p : '\0';
}

}  // namespace

void BasevarCaller::call(const std::vector<std::string>& lines) {
    try {
        call_row(lines);
    } catch (...) {
        // The reference writes position by position, so everything before a malformed row is on disk when it throws
        // (src/basetype_caller.cpp:586-611).  Here up to n_slots tiles of earlier positions are still queued or in flight: they
        // are emitted first (the row that failed was not counted), then the error goes on.
        try { finish(); } catch (...) {}
        throw;
    }
}

void BasevarCaller::call_row(const std::vector<std::string>& lines) {
    Tile& T = *tiles_[cur_];
    if (T.pending) drain(cur_);   // the slot comes round again: its previous tile is emitted first
    T.ensure_dense();
    T.dense = true;
    const uint32_t row = T.n_sites;
    const size_t n = n_sample_;
    uint8_t* b = T.base + (size_t)row * T.pitch;
    uint8_t* q = T.qual + (size_t)row * T.pitch;
    uint8_t* s = T.strand + (size_t)row * T.pitch;
    uint8_t* mq = T.mapq + (size_t)row * T.pitch;
    uint16_t* rp = T.rpr + (size_t)row * T.rpr_pitch;
    // uncovered = N ! 0 0 . (src/basetype_caller.cpp:1071-1077); also what the padding cells hold
    memset(b, BV_BASE_N, T.pitch); memset(q, 0, T.pitch); memset(s, BV_STRAND_NONE, T.pitch); memset(mq, 0, T.pitch);
    memset(rp, 0, T.rpr_pitch * sizeof(uint16_t));
    SiteMeta& m = T.meta[row];
    m.specials.clear();
    m.odd_strands.clear();
    m.depth = 0;
    size_t n_mq = 0, n_b = 0, n_q = 0, n_rp = 0, n_s = 0;

    for (size_t li = 0; li < lines.size(); ++li) {
        const std::string& line = lines[li];
        const char* col[10];
        const char* col_end[10];
        size_t ncol = 0;
        for_each_token(line.data(), line.data() + line.size(), '\t', [&](size_t k, const char* p, const char* e) {
            if (k < 10) { col[k] = p; col_end[k] = e; }
            ncol = k + 1;
        });
        if (ncol != 9) throw std::runtime_error("[ERROR] batchfile has invalid data:\n" + line);
        const std::string ref_id(col[0], col_end[0]), ref_base(col[2], col_end[2]);
        const uint32_t ref_pos = (uint32_t)std::stoi(std::string(col[1], col_end[1]));
        if (li == 0) {
            m.ref_id = ref_id; m.ref_pos = ref_pos; m.ref_base = ref_base;
        } else if (m.ref_id != ref_id || m.ref_pos != ref_pos || m.ref_base != ref_base) {
            throw std::runtime_error("[ERROR] Batchfiles must have the same genome coordinate in each line.");
        }
        m.depth += (uint32_t)std::stoi(std::string(col[3], col_end[3]));
        for_each_token(col[4], col_end[4], ' ', [&](size_t, const char* p, const char* e) {
            const long v = parse_int(p, e);
            if (n_mq < n) mq[n_mq] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            ++n_mq;
        });
        for_each_token(col[5], col_end[5], ' ', [&](size_t, const char* p, const char* e) {
            if (n_b < n) {
                const size_t len = (size_t)(e - p);
                const char fb = len ? p[0] : 'N';
                uint8_t code;
                if (len == 1 && (fb == 'A' || fb == 'C' || fb == 'G' || fb == 'T')) code = (uint8_t)code_of_char(fb);
                else if (fb == 'N') code = BV_BASE_N;
                else {
                    const std::string str(p, e);
                    code = encode_base(str);   // indel, other character, or the reference's "size is not 1" error
                    m.specials.emplace_back((uint32_t)n_b, str);
                }
                b[n_b] = code;
            }
            ++n_b;
        });
        for_each_token(col[6], col_end[6], ' ', [&](size_t, const char* p, const char* e) {
            const int c = (unsigned char)parse_char(p, e);
            if (n_q < n) q[n_q] = (uint8_t)(c >= 33 ? c - 33 : 0);
            ++n_q;
        });
        for_each_token(col[7], col_end[7], ' ', [&](size_t, const char* p, const char* e) {
            const long v = parse_int(p, e);
            if (v > 65535) throw std::runtime_error("[ERROR] read position rank above 65535 is not supported: " + line.substr(0, 64));
            if (n_rp < n) rp[n_rp] = (uint16_t)(v < 0 ? 0 : v);
            ++n_rp;
        });
        for_each_token(col[8], col_end[8], ' ', [&](size_t, const char* p, const char* e) {
            const char c = parse_char(p, e);
            if (n_s < n) {
                s[n_s] = encode_strand(c);
                if (c != '+' && c != '-' && c != '.') m.odd_strands.emplace_back((uint32_t)n_s, c);
            }
            ++n_s;
        });
    }
    if (m.depth == 0) return;   // coverage is 0 for all samples on this position: skipped (cpp:717-718)
    if (n_mq != n || n_b != n || n_q != n || n_s != n || n_rp != n)
        throw std::runtime_error("[ERROR] Something is wrong in batchfiles.");
    T.ref[row] = m.ref_base.empty() ? (uint8_t)'N' : (uint8_t)m.ref_base[0];
    ++T.n_sites;
    ++n_positions_;
    if (T.n_sites == opt_.tile_sites) submit_current();
}

namespace {
struct StageClock {   // adds the scope's wall time to *acc (no-op when acc is null)
    double* acc;
    std::chrono::steady_clock::time_point t0;
    explicit StageClock(double* a) : acc(a) { if (acc) t0 = std::chrono::steady_clock::now(); }
    ~StageClock() { if (acc) *acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
}  // namespace

TileRows BasevarCaller::begin_tile(uint32_t n_rows) {
    if (n_rows == 0 || n_rows > opt_.tile_sites) throw std::invalid_argument("[ERROR] begin_tile: row count outside the tile");
    if (!tiles_[cur_]->pending && tiles_[cur_]->n_sites) submit_current();   // rows queued through call() go first
    Tile& T = *tiles_[cur_];
    if (T.pending) drain(cur_);
    StageClock clk(opt_.times ? &opt_.times->tile_reset : nullptr);
    // A packer that lists the covered cells of its tile (sparse upload) gets no planes to fill: a tile of an < 1x cohort is
    // nine tenths filler, and clearing and scattering into 6 bytes per sample-site was most of the host's time.  The text of a
    // called site takes its row from the cell list (Tile::site_cells).
    const bool offer_sparse = opt_.sparse_upload && n_sample_ <= BV_CELL_MAX_SAMPLES;
    T.dense = !offer_sparse || opt_.dense_rows;
    if (T.dense) {
        T.ensure_dense();
        const size_t cells = (size_t)n_rows * T.pitch;
        memset(T.base, BV_BASE_N, cells); memset(T.qual, 0, cells); memset(T.strand, BV_STRAND_NONE, cells); memset(T.mapq, 0, cells);
        memset(T.rpr, 0, (size_t)n_rows * T.rpr_pitch * sizeof(uint16_t));
    }
    for (uint32_t i = 0; i < n_rows; ++i) {
        SiteMeta& m = T.meta[i];
        m.specials.clear(); m.odd_strands.clear(); m.depth = 0;
    }
    T.n_sites = n_rows;
    T.sparse_ready = false;
    TileRows rows{T.dense ? T.base : nullptr, T.dense ? T.qual : nullptr, T.dense ? T.strand : nullptr, T.dense ? T.mapq : nullptr,
                  T.dense ? T.rpr : nullptr, T.pitch, T.rpr_pitch, n_rows, T.meta.data(), nullptr, nullptr, nullptr};
    if (offer_sparse) {
        Tile* tp = &T;
        rows.site_start = T.sp_start;
        rows.reserve_cells = [tp](size_t n, uint32_t** cells, uint32_t** aux) {
            tp->reserve_cells(n);
            *cells = tp->sp_cells; *aux = tp->sp_aux;
        };
        rows.sparse_ready = &T.sparse_ready;
    }
    return rows;
}

void BasevarCaller::commit_tile() {
    Tile& T = *tiles_[cur_];
    for (uint32_t i = 0; i < T.n_sites; ++i) {
        T.ref[i] = T.meta[i].ref_base.empty() ? (uint8_t)'N' : (uint8_t)T.meta[i].ref_base[0];
        if (T.meta[i].depth) ++n_positions_;
    }
    submit_current();
}

void BasevarCaller::submit_current() {
    Tile& T = *tiles_[cur_];
    if (T.pending || T.n_sites == 0) return;   // a pending tile still holds the rows it was submitted with
    StageClock clk(opt_.times ? &opt_.times->encode_submit : nullptr);
    if (T.sparse_ready) {   // the packer listed the covered cells: 8 bytes per read instead of 5 per sample-site
        T.sparse_ready = false;
        bv_sparse_tile st;
        st.cells = T.sp_cells; st.cells_aux = T.sp_aux; st.site_start = T.sp_start; st.ref_base = T.ref; st.out = nullptr;
        st.n_sites = T.n_sites; st.n_samples = (uint32_t)n_sample_; st.format = BV_CELLS_U32; st.out_mode = BV_OUT_RECORDS;
        // the packer's cells ascend by sample within a row: two bytes per cell (plus the aux word) instead of four, unless a
        // cell has no compact form (a strand that is neither + nor -)
        // (one pass: the buffer is sized for the worst case, so that the encoder never has to count first)
        uint64_t n_words = 0;
        T.reserve_words16((size_t)bv_sparse_encode16_bound(T.sp_start[T.n_sites], T.n_sites, (uint32_t)n_sample_) + BV_CELL_MAX_SAMPLES / BV_CELL16_GAP_SKIP + 2);
        if (bv_sparse_encode16(T.sp_cells, T.sp_aux, T.sp_start, T.n_sites, T.sp_words16, T.sp_aux16, T.sp_cap16, T.sp_start16, &n_words) == BV_OK) {
            st.cells = T.sp_words16; st.cells_aux = T.sp_aux16; st.site_start = T.sp_start16; st.format = BV_CELLS_U16;
        }
        check(bv_tile_submit_sparse_calls(ctx_, (int)cur_, &st), ctx_, "bv_tile_submit_sparse_calls");
        T.pending = true;
        cur_ = (cur_ + 1) % (uint32_t)tiles_.size();
        return;
    }
    if (!T.dense) throw std::runtime_error("[BUG] tile without dense rows and without a cell list");
    bv_tile t;
    t.base = T.base; t.qual = T.qual; t.strand = T.strand; t.ref_base = T.ref;
    t.pitch = T.pitch; t.n_sites = T.n_sites; t.n_samples = (uint32_t)n_sample_;
    t.location = BV_LOC_HOST; t.out_mode = BV_OUT_RECORDS;
    bv_tile_aux a;
    a.mapq = T.mapq; a.rpr = T.rpr; a.rpr_pitch = T.rpr_pitch;
    check(bv_tile_submit_calls(ctx_, (int)cur_, &t, &a), ctx_, "bv_tile_submit_calls");
    T.pending = true;
    cur_ = (cur_ + 1) % (uint32_t)tiles_.size();
}

void BasevarCaller::drain(uint32_t slot) {
    Tile& T = *tiles_[slot];
    if (!T.pending) return;
    uint32_t n_calls = 0;
    const size_t G = group_names_.size();
    {
        StageClock clk(opt_.times ? &opt_.times->gpu_wait : nullptr);
        check(bv_tile_wait_calls(ctx_, (int)slot, T.recs.data(), T.calls.data(), (uint32_t)T.calls.size(), &n_calls,
                                 G ? T.groups.data() : nullptr), ctx_, "bv_tile_wait_calls");
    }
    StageClock clk_text(opt_.times ? &opt_.times->text : nullptr);
    T.pending = false;
    std::fill(T.call_of_site.begin(), T.call_of_site.begin() + T.n_sites, -1);
    for (uint32_t k = 0; k < n_calls; ++k) T.call_of_site[T.calls[k].site] = (int32_t)k;
    std::string cvg_text, vcf_text, flip_text;
    for (uint32_t i = 0; i < T.n_sites; ++i) {
        const bv_site_out& rec = T.recs[i];
        if (T.meta[i].depth == 0) continue;   // no sample covers the position: no row (cpp:717-718)
        if (rec.flags & BV_FLAG_ZERO_SUBSET)   // src/basetype.cpp:113-115
            throw std::runtime_error("[ERROR] The sum of frequence of active bases must always > 0. Check: " + T.meta[i].ref_id + ":" +
                                     std::to_string(T.meta[i].ref_pos));
        const bool need_cells = rec.n_alt || (rec.flags & BV_FLAG_BAD_STRAND);   // the per-sample columns of a VCF row; an error message
        const SiteCells c = need_cells ? T.site_cells(i, (uint32_t)n_sample_) : SiteCells{nullptr, nullptr, nullptr, (uint32_t)n_sample_};
        struct Release { Tile& t; uint32_t i; bool on; ~Release() { if (on) t.release_site_cells(i); } } release{T, i, need_cells};
        cvg_text += out_cvg_line(T.meta[i], c, rec);
        if ((rec.flags & (BV_FLAG_NEAR_LRT | BV_FLAG_LRT_TIE)) && opt_.flip_log) {
            flip_text += T.meta[i].ref_id + "\t" + std::to_string(T.meta[i].ref_pos) + "\t";
            if (rec.flags & BV_FLAG_NEAR_LRT) flip_text += "NEAR_LRT";
            if ((rec.flags & BV_FLAG_NEAR_LRT) && (rec.flags & BV_FLAG_LRT_TIE)) flip_text += ",";
            if (rec.flags & BV_FLAG_LRT_TIE) flip_text += "LRT_TIE";
            flip_text += "\n";
        }
        if (rec.n_alt) {   // only SNPs reach the VCF (cpp:745)
            const int32_t k = T.call_of_site[i];
            if (k < 0) throw std::runtime_error("[BUG] called site without its rank-sum record");
            vcf_text += out_vcf_line(T.meta[i], c, rec, T.calls[k], group_names_, G ? &T.groups[(size_t)k * G] : nullptr);
            ++n_snps_;
        }
    }
    if (!cvg_text.empty() && cvg_) cvg_(cvg_text.data(), cvg_text.size());
    if (!vcf_text.empty() && vcf_) vcf_(vcf_text.data(), vcf_text.size());
    if (!flip_text.empty()) opt_.flip_log(flip_text.data(), flip_text.size());
    T.n_sites = 0;
}

bool BasevarCaller::finish() {
    submit_current();
    // oldest first: after submit_current() the next slot to be reused is the oldest one in flight
    for (uint32_t k = 0; k < tiles_.size(); ++k) drain((cur_ + k) % (uint32_t)tiles_.size());
    return n_snps_ > 0;
}

// ---- batchfiles -------------------------------------------------------------------------------------------------------------
namespace {
struct GzReader {
    gzFile f = nullptr;
    std::string path;
    explicit GzReader(const std::string& p) : path(p) {
        f = gzopen(p.c_str(), "rb");
        if (!f) throw std::runtime_error("[ERROR] " + p + " open failure.");
        gzbuffer(f, 1 << 20);
    }
    ~GzReader() { if (f) gzclose(f); }
    GzReader(const GzReader&) = delete;
    GzReader& operator=(const GzReader&) = delete;
    bool getline(std::string& out) {   // BGZF is a series of gzip members: zlib reads through them
        out.clear();
        char buf[65536];
        while (gzgets(f, buf, (int)sizeof(buf))) {
            const size_t len = strlen(buf);
            out.append(buf, len);
            if (len && buf[len - 1] == '\n') { out.pop_back(); return true; }
        }
        return !out.empty();
    }
};

struct Region {
    bool any = true;
    std::string chrom;
    long beg = 1, end = 0x7fffffffL;
    bool holds(const std::string& line) const {
        if (any) return true;
        const size_t t1 = line.find('\t');
        if (t1 == std::string::npos || line.compare(0, t1, chrom) != 0) return false;
        const long pos = atol(line.c_str() + t1 + 1);
        return pos >= beg && pos <= end;
    }
};
Region parse_region(const std::string& r) {
    Region g;
    if (r.empty()) return g;
    g.any = false;
    const size_t c = r.rfind(':');
    if (c == std::string::npos) { g.chrom = r; return g; }
    g.chrom = r.substr(0, c);
    std::string rest;
    for (char ch : r.substr(c + 1)) if (ch != ',') rest += ch;
    const size_t d = rest.find('-');
    g.beg = atol(rest.substr(0, d).c_str());
    if (d != std::string::npos && d + 1 < rest.size()) g.end = atol(rest.substr(d + 1).c_str());
    return g;
}
}  // namespace

std::vector<std::string> get_sampleid_from_batchfiles(const std::vector<std::string>& batchfiles) {
    std::vector<std::string> ids;
    for (const auto& fn : batchfiles) {
        GzReader r(fn);
        std::string line;
        while (r.getline(line)) {
            if (line.empty() || line[0] != '#') break;
            if (line.compare(0, 12, "##SampleIDs=") == 0) {
                const std::string list = line.substr(12);
                for_each_token(list.data(), list.data() + list.size(), ',',
                               [&](size_t, const char* p, const char* e) { ids.emplace_back(p, e); });
                break;
            }
        }
    }
    return ids;
}

bool variant_calling_unit(const std::vector<std::string>& batchfiles, const std::vector<std::string>& sample_ids,
                          const std::map<std::string, std::vector<size_t>>& group_smp_idx, double min_af,
                          const std::string& region, TextSink vcf, TextSink cvg, const CallerOptions& opt) {
    const std::vector<std::string> bf_ids = get_sampleid_from_batchfiles(batchfiles);
    if (bf_ids != sample_ids) {
        auto join = [](const std::vector<std::string>& v) {
            std::string s;
            for (size_t i = 0; i < v.size(); ++i) { if (i) s += ","; s += v[i]; }
            return s;
        };
        throw std::runtime_error("[BUG] The order of sample ids in batchfiles must be the same as input bamfiles.\n"
                                 "Sample ids in batchfiles: " + join(bf_ids) + "\nSample ids in bamfiles  : " + join(sample_ids) + "\n");
    }
    std::vector<std::unique_ptr<GzReader>> readers;
    for (const auto& fn : batchfiles) readers.emplace_back(new GzReader(fn));
    const Region reg = parse_region(region);
    BasevarCaller caller(sample_ids.size(), group_smp_idx, min_af, std::move(vcf), std::move(cvg), opt);
    std::vector<std::string> lines(batchfiles.size());
    bool eof = false;
    while (!eof) {
        for (size_t i = 0; i < readers.size(); ++i) {
            // next data row of this file inside the region
            bool got = false;
            while (readers[i]->getline(lines[i])) {
                if (lines[i].empty() || lines[i][0] == '#') continue;
                if (!reg.holds(lines[i])) continue;
                got = true;
                break;
            }
            if (!got) { eof = true; break; }
        }
        if (eof) break;
        caller.call(lines);
    }
    return caller.finish();
}

}  // namespace bvhost
