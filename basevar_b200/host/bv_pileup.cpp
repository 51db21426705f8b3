// bv_pileup.cpp -- see bv_pileup.hpp.
#include "bv_pileup.hpp"

#include <getopt.h>
#include <sys/resource.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <tuple>

namespace bvhost {

namespace {

// bam_get_seq codes -> characters the reference keeps (src/bam_record.h:28-31): ambiguity codes and '=' become ' '
const char kSeqChar[16] = {' ', 'A', 'C', ' ', 'G', ' ', ' ', ' ', 'T', ' ', ' ', ' ', ' ', ' ', ' ', 'N'};

inline uint8_t base_code(char c) {
    return c == 'A' ? BV_BASE_A : c == 'C' ? BV_BASE_C : c == 'G' ? BV_BASE_G : c == 'T' ? BV_BASE_T : BV_BASE_N;
}

// A read character that the reference's space-joined batchfile row cannot carry (' ': ambiguity code, '=' or an absent
// quality): its phase 2 stops with this message when it re-parses the row (src/basetype_caller.cpp:720-736).
[[noreturn]] void throw_unrepresentable(const std::string& bam, int64_t pos1) {
    throw std::runtime_error("[ERROR] Something is wrong in batchfiles. (a read in " + bam + " covering position " + std::to_string(pos1) +
                             " has a base or quality the batchfile row cannot represent)");
}

template <class F>
void parallel_for(size_t n, int n_threads, F&& fn) {
    if (n == 0) return;
    const int T = (int)std::min<size_t>((size_t)std::max(1, n_threads), n);
    if (T == 1) {
        for (size_t i = 0; i < n; ++i) fn(i, 0);
        return;
    }
    std::atomic<size_t> next{0};
    std::exception_ptr err;
    std::mutex mu;
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            try {
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= n) break;
                    fn(i, t);
                }
            } catch (...) {
                std::lock_guard<std::mutex> g(mu);
                if (!err) err = std::current_exception();
                next.store(n);
            }
        });
    for (auto& x : th) x.join();
    if (err) std::rethrow_exception(err);
}

}  // namespace

// ---- one sample ---------------------------------------------------------------------------------------------------------
void pileup_sample(BamReader& bam, int tid, const std::string& fa, uint32_t reg_beg, uint32_t reg_end, uint32_t sb,
                   uint32_t se, int mapq_thd, SamplePileup& out, std::vector<uint8_t>& occ) {
    out.cells.clear();
    out.specials.clear();
    if (se < sb) return;
    occ.assign((size_t)(se - sb) + 1, 0);
    // last position of the reference's creation step that holds p (cpp:838-846): an indel rolled back onto p comes from
    // the aligned pair at p + 1, which that step only visits while p + 1 <= its end (cpp:969-970 run before the roll-back)
    auto step_end = [&](uint32_t p) -> uint64_t {
        const uint64_t e = (uint64_t)reg_beg + ((uint64_t)(p - reg_beg) / PILEUP_STEP_REGION_LEN + 1) * PILEUP_STEP_REGION_LEN - 1;
        return std::min<uint64_t>(e, reg_end);
    };
    auto anchor = [&](int64_t p1) -> char { return (p1 >= 1 && (uint64_t)p1 <= fa.size()) ? fa[(size_t)p1 - 1] : 'N'; };

    // reads overlapping positions sb .. se + 1 (a read that starts with an insertion at se + 1 is anchored on se)
    bam.query(tid, (int64_t)sb - 1, (int64_t)se + 1);
    BamRec r;
    while (bam.next(r)) {
        // mapq() is 0 for an unmapped read (src/bam_record.h:247) and -q is >= 1: unmapped reads never pass (cpp:906)
        if (!r.is_mapped() || (int)r.mapq < mapq_thd || (r.flag & BAM_FLAG_DUP) || (r.flag & BAM_FLAG_QCFAIL)) continue;
        const int64_t start1 = (int64_t)r.pos + 1, end1 = r.end;
        if ((int64_t)sb > end1 + 1) continue;
        if ((int64_t)se + 1 < start1) break;
        const uint8_t strand = (r.flag & BAM_FLAG_REVERSE) ? BV_STRAND_REV : BV_STRAND_FWD;
        int mean_q = -1;
        auto mean_qual = [&]() -> uint8_t {   // int(mean_qqual()) + 33 as a char (cpp:960, bam_record.cpp:332-343)
            if (mean_q < 0) {
                if (r.l_seq <= 0) throw_unrepresentable(bam.path(), start1);   // -1 + 33 is ' '
                double tot = 0;
                for (int32_t i = 0; i < r.l_seq; ++i) tot += r.qual[(size_t)i];
                mean_q = (int)(tot / r.l_seq);
            }
            if (mean_q > 93) throw_unrepresentable(bam.path(), start1);
            return (uint8_t)mean_q;
        };
        int64_t rpos = r.pos;   // 0-based reference coordinate of the next aligned column
        int64_t qpos = 0;       // 0-based read coordinate
        for (size_t ci = 0; ci < r.cigar.size(); ++ci) {
            const unsigned op = r.cigar[ci] & 0xf;
            const int64_t len = r.cigar[ci] >> 4;
            if (rpos + 1 > (int64_t)se + 1) break;   // every later pair lies beyond the span (cpp:969)
            if (op == CIG_M || op == CIG_EQ || op == CIG_X) {
                const int64_t lo = std::max<int64_t>(rpos + 1, sb), hi = std::min<int64_t>(rpos + len, se);
                for (int64_t m = lo; m <= hi; ++m) {
                    uint8_t& o = occ[(size_t)(m - sb)];
                    if (o) continue;
                    o = 1;
                    const int64_t q = qpos + (m - (rpos + 1));
                    if (q >= r.l_seq) throw_unrepresentable(bam.path(), m);
                    const char c = kSeqChar[r.seqi((int)q)];
                    const uint8_t ql = r.qual[(size_t)q];
                    if (c == ' ' || ql > 93) throw_unrepresentable(bam.path(), m);
                    if (q + 1 > 65535) throw std::runtime_error("[ERROR] read position rank above 65535 is not supported: " + bam.path());
                    out.cells.push_back(PileupCell{(uint32_t)(m - sb), base_code(c), ql, strand, r.mapq, (uint16_t)(q + 1), -1});
                }
                rpos += len;
                qpos += len;
            } else if (op == CIG_I || op == CIG_D) {
                const int64_t p = rpos;   // the pair sits at map_ref_pos = rpos + 1 and is rolled back one position (cpp:985,997)
                if (p >= (int64_t)sb && p <= (int64_t)se && (uint64_t)(rpos + 1) <= step_end((uint32_t)p) && !occ[(size_t)(p - sb)]) {
                    occ[(size_t)(p - sb)] = 1;
                    std::string s;
                    bool plain = false;   // equal lengths of ref and read strings are written as a plain base (cpp:1059-1061)
                    if (op == CIG_I) {
                        std::string ins;
                        for (int64_t k = 0; k < len && qpos + k < r.l_seq; ++k) {
                            const char c = kSeqChar[r.seqi((int)(qpos + k))];
                            if (c == ' ') throw_unrepresentable(bam.path(), p);
                            ins += c;
                        }
                        if (ins.empty()) { plain = true; s = std::string(1, anchor(p)); }
                        else s = std::string("+") + anchor(p) + ins;
                    } else {
                        const std::string del = (uint64_t)rpos < fa.size() ? fa.substr((size_t)rpos, (size_t)len) : std::string();
                        if (del.empty()) { plain = true; s = std::string(1, anchor(p)); }
                        else s = std::string("-") + anchor(p) + del;
                    }
                    if (qpos + 1 > 65535) throw std::runtime_error("[ERROR] read position rank above 65535 is not supported: " + bam.path());
                    PileupCell c{(uint32_t)(p - sb), BV_BASE_N, mean_qual(), strand, r.mapq, (uint16_t)(qpos + 1), -1};
                    if (plain && s.size() == 1 && (s[0] == 'A' || s[0] == 'C' || s[0] == 'G' || s[0] == 'T' || s[0] == 'N')) {
                        c.base = base_code(s[0]);
                    } else {
                        c.base = encode_base(s);
                        c.special = (int32_t)out.specials.size();
                        out.specials.push_back(s);
                    }
                    out.cells.push_back(c);
                }
                if (op == CIG_I) qpos += len; else rpos += len;
            } else if (op == CIG_S || op == CIG_P) {
                qpos += len;   // padding advances the read coordinate in get_aligned_pairs (bam_record.cpp:251-261)
            } else if (op == CIG_N) {
                rpos += len;
            }   // hard clips: nothing
        }
    }
    if (!std::is_sorted(out.cells.begin(), out.cells.end(), [](const PileupCell& a, const PileupCell& b) { return a.off < b.off; }))
        std::sort(out.cells.begin(), out.cells.end(), [](const PileupCell& a, const PileupCell& b) { return a.off < b.off; });
}

// ---- all samples --------------------------------------------------------------------------------------------------------
struct BamPileup::Impl {
    std::vector<std::unique_ptr<BamReader>> readers;   // kept open when the process may hold that many descriptors
    bool keep_open = false;
    std::vector<SamplePileup> piles;
    uint32_t span_beg = 0, span_end = 0;
};

BamPileup::BamPileup(const std::vector<std::string>& bam_files, int mapq_thd, int n_threads, int n_instances)
    : files_(bam_files), mapq_thd_(mapq_thd), n_threads_(std::max(1, n_threads)), impl_(new Impl) {
    impl_->readers.resize(files_.size());
    impl_->piles.resize(files_.size());
    struct rlimit rl;
    if (getrlimit(RLIMIT_NOFILE, &rl) == 0) {
        if (rl.rlim_cur < rl.rlim_max) { rl.rlim_cur = rl.rlim_max; setrlimit(RLIMIT_NOFILE, &rl); getrlimit(RLIMIT_NOFILE, &rl); }
        impl_->keep_open = (uint64_t)files_.size() * (uint64_t)std::max(1, n_instances) + 256 < (uint64_t)rl.rlim_cur;
    }
}

BamPileup::~BamPileup() {}

std::vector<std::string> BamPileup::sample_ids(bool filename_has_samplename) {
    std::vector<std::string> ids(files_.size());
    parallel_for(files_.size(), filename_has_samplename ? 1 : n_threads_, [&](size_t i, int) {
        std::string name;
        if (filename_has_samplename) {
            std::string fn = std::filesystem::path(files_[i]).filename().string();
            const size_t dot = fn.find_last_of('.');
            if (dot > 0 && dot != std::string::npos) fn.resize(dot);
            const size_t si = fn.find('.');
            name = (si > 0 && si != std::string::npos) ? fn.substr(0, si) : fn;
        } else {
            BamReader b(files_[i]);
            if (!b.sample_name(name))
                throw std::runtime_error("[bam_header.cpp::BamHeader:get_sample_name] Bam file format error: "
                                         "missing `SM` tag in `@RG` field in BAM/CRAM/SAM header.");
        }
        if (name.empty())
            throw std::invalid_argument("[BaseTypeRunner::_load_sample_id_from_bam] " + files_[i] + " sample ID not found.\n");
        ids[i] = name;
    });
    return ids;
}

bool BamPileup::load_span(const std::string& ref_id, const std::string& fa_seq, uint32_t reg_beg, uint32_t reg_end,
                          uint32_t span_beg, uint32_t span_end) {
    impl_->span_beg = span_beg;
    impl_->span_end = span_end;
    std::vector<std::vector<uint8_t>> occ((size_t)n_threads_);
    std::atomic<bool> any{false};
    parallel_for(files_.size(), n_threads_, [&](size_t i, int t) {
        std::unique_ptr<BamReader> local;
        BamReader* b;
        if (impl_->keep_open) {
            if (!impl_->readers[i]) impl_->readers[i].reset(new BamReader(files_[i]));
            b = impl_->readers[i].get();
        } else {
            local.reset(new BamReader(files_[i]));
            b = local.get();
        }
        const int tid = b->name2id(ref_id);
        if (tid < 0)   // sam_itr_querys fails on an unknown contig (src/bam.cpp:93-98)
            throw std::runtime_error("[bam.cpp::Bam:fetch] Fail to fetch the alignment data in : " + ref_id + ":" +
                                     std::to_string(span_beg) + "-" + std::to_string(span_end));
        pileup_sample(*b, tid, fa_seq, reg_beg, reg_end, span_beg, span_end, mapq_thd_, impl_->piles[i], occ[(size_t)t]);
        if (!impl_->piles[i].cells.empty()) any.store(true);
    });
    return any.load();
}

void BamPileup::scatter(uint32_t pos, uint32_t n, const std::string& ref_id, const std::string& fa_seq, TileRows& rows) {
    if (n > rows.n_rows) throw std::invalid_argument("[ERROR] scatter: more positions than rows");
    if (pos < impl_->span_beg || (uint64_t)pos + n - 1 > impl_->span_end) throw std::invalid_argument("[ERROR] scatter: positions outside the loaded span");
    const uint32_t off0 = pos - impl_->span_beg;
    const size_t N = files_.size();
    struct Spec { uint32_t row, sample; int32_t idx; };
    const int T = n_threads_;
    std::vector<std::vector<uint32_t>> depth((size_t)T);
    std::vector<std::vector<Spec>> specs((size_t)T);
    const size_t block = 256;   // samples per work item: neighbouring columns stay with one thread
    const size_t n_blocks = (N + block - 1) / block;
    // sparse transport: cells per (sample block, row), so that the second pass below knows where each block writes
    const bool sparse = rows.sparse_ready != nullptr && rows.site_start != nullptr && rows.reserve_cells;
    const bool dense = rows.base != nullptr;
    if (!dense && !sparse) throw std::invalid_argument("[ERROR] scatter: neither planes nor a cell list to fill");
    std::vector<uint32_t> block_cnt(sparse ? n_blocks * (size_t)n : 0, 0);
    parallel_for(n_blocks, T, [&](size_t bi, int t) {
        std::vector<uint32_t>& d = depth[(size_t)t];
        if (d.empty()) d.assign(n, 0);
        uint32_t* bc = sparse ? block_cnt.data() + bi * (size_t)n : nullptr;
        for (size_t s = bi * block; s < std::min(N, (bi + 1) * block); ++s) {
            const std::vector<PileupCell>& cells = impl_->piles[s].cells;
            auto it = std::lower_bound(cells.begin(), cells.end(), off0, [](const PileupCell& c, uint32_t v) { return c.off < v; });
            for (; it != cells.end() && it->off < off0 + n; ++it) {
                const uint32_t row = it->off - off0;
                if (dense) {
                    const size_t at = (size_t)row * rows.pitch + s;
                    rows.base[at] = it->base;
                    rows.qual[at] = it->qual;
                    rows.strand[at] = it->strand;
                    rows.mapq[at] = it->mapq;
                    rows.rpr[(size_t)row * rows.rpr_pitch + s] = it->rpr;
                }
                ++d[row];
                if (bc) ++bc[row];
                if (it->special >= 0) specs[(size_t)t].push_back(Spec{row, (uint32_t)s, it->special});
            }
        }
    });
    if (sparse) {
        // row offsets, then every block's write position inside its rows (exclusive prefix over the blocks)
        uint64_t total = 0;
        for (uint32_t i = 0; i < n; ++i) {
            rows.site_start[i] = (uint32_t)total;
            for (size_t bi = 0; bi < n_blocks; ++bi) {
                const uint32_t c = block_cnt[bi * (size_t)n + i];
                block_cnt[bi * (size_t)n + i] = (uint32_t)total;
                total += c;
            }
        }
        rows.site_start[n] = (uint32_t)total;
        if (total <= 0xffffffffull) {
            uint32_t *sp_cells = nullptr, *sp_aux = nullptr;
            rows.reserve_cells((size_t)total, &sp_cells, &sp_aux);
            parallel_for(n_blocks, T, [&](size_t bi, int) {
                uint32_t* at = block_cnt.data() + bi * (size_t)n;
                for (size_t s = bi * block; s < std::min(N, (bi + 1) * block); ++s) {
                    const std::vector<PileupCell>& cells = impl_->piles[s].cells;
                    auto it = std::lower_bound(cells.begin(), cells.end(), off0, [](const PileupCell& c, uint32_t v) { return c.off < v; });
                    for (; it != cells.end() && it->off < off0 + n; ++it) {
                        const uint32_t k = at[it->off - off0]++;
                        sp_cells[k] = BV_CELL_PACK((uint32_t)s, it->base, it->strand, it->qual);
                        sp_aux[k] = BV_CELL_AUX_PACK(it->mapq, it->rpr);
                    }
                }
            });
            *rows.sparse_ready = true;
        } else if (!dense) {
            throw std::runtime_error("[ERROR] more than 2^32 covered cells in one tile: lower --tile-sites");
        }
    }
    for (uint32_t i = 0; i < n; ++i) {
        SiteMeta& m = rows.meta[i];
        m.ref_id = ref_id;
        m.ref_pos = pos + i;
        const uint64_t p = (uint64_t)pos + i;
        m.ref_base.assign(1, (p >= 1 && p <= fa_seq.size()) ? fa_seq[(size_t)p - 1] : 'N');   // cpp:1080
        uint32_t d = 0;
        for (int t = 0; t < T; ++t)
            if (!depth[(size_t)t].empty()) d += depth[(size_t)t][i];
        m.depth = d;
    }
    std::vector<Spec> all;
    for (auto& v : specs) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end(), [](const Spec& a, const Spec& b) { return a.row != b.row ? a.row < b.row : a.sample < b.sample; });
    for (const Spec& s : all) rows.meta[s.row].specials.emplace_back(s.sample, impl_->piles[s.sample].specials[(size_t)s.idx]);
}

// ---- the reference's batchfile row, for tests -------------------------------------------------------------------------------
std::string batchfile_row(const SiteMeta& m, const SiteCells& c, const uint8_t* mapq, const uint16_t* rpr) {
    std::string mq, bs, qs, rp, st;
    size_t sp = 0;
    for (uint32_t i = 0; i < c.n_samples; ++i) {
        if (i) { mq += ' '; bs += ' '; qs += ' '; rp += ' '; st += ' '; }
        mq += std::to_string((int)mapq[i]);
        const uint8_t b = c.base[i];
        while (sp < m.specials.size() && m.specials[sp].first < i) ++sp;
        if (sp < m.specials.size() && m.specials[sp].first == i) bs += m.specials[sp].second;
        else bs += (b < 4 ? "ACGT"[b] : 'N');
        qs += (char)(c.qual[i] + 33);
        rp += std::to_string((int)rpr[i]);
        char sc = c.strand[i] == BV_STRAND_FWD ? '+' : c.strand[i] == BV_STRAND_REV ? '-' : '.';
        for (const auto& o : m.odd_strands)
            if (o.first == i) sc = o.second;
        st += sc;
    }
    return m.ref_id + "\t" + std::to_string(m.ref_pos) + "\t" + m.ref_base + "\t" + std::to_string(m.depth) + "\t" + mq + "\t" + bs +
           "\t" + qs + "\t" + rp + "\t" + st;
}

// ---- output files ---------------------------------------------------------------------------------------------------------------
TextWriter::TextWriter(const std::string& path) : path_(path) {
    gz_ = path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0;
    f_ = fopen(path.c_str(), "wb");
    if (!f_) throw std::runtime_error("[ERROR] " + path + " open failure.");
}

TextWriter::~TextWriter() {
    try { close(); } catch (...) {}
}

void TextWriter::flush_block(const uint8_t* p, size_t n) {
    // one BGZF block: a gzip member with the 'BC' extra field holding its total size - 1 (SAM specification 4.1)
    uint8_t out[65536 + 64];
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("[ERROR] zlib deflateInit2 failed");
    zs.next_in = const_cast<uint8_t*>(p);
    zs.avail_in = (uInt)n;
    zs.next_out = out + 18;
    zs.avail_out = (uInt)(sizeof(out) - 18 - 8);
    const int rc = deflate(&zs, Z_FINISH);
    const size_t clen = zs.total_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) throw std::runtime_error("[ERROR] fail to write data");
    const size_t total = 18 + clen + 8;
    const uint8_t head[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, (uint8_t)((total - 1) & 0xff), (uint8_t)((total - 1) >> 8)};
    memcpy(out, head, 18);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), p, (uInt)n);
    uint8_t* t = out + 18 + clen;
    for (int k = 0; k < 4; ++k) { t[k] = (uint8_t)(crc >> (8 * k)); t[4 + k] = (uint8_t)((uint32_t)n >> (8 * k)); }
    if (fwrite(out, 1, total, f_) != total) throw std::runtime_error("[ERROR] fail to write data");
}

void TextWriter::write(const char* data, size_t n) {
    if (!f_) throw std::runtime_error("[ERROR] fail to write data");
    if (!gz_) {
        if (fwrite(data, 1, n, f_) != n) throw std::runtime_error("[ERROR] fail to write data");
        return;
    }
    const size_t kBlock = 0xff00;   // uncompressed bytes per block, as bgzip cuts them
    pend_.insert(pend_.end(), (const uint8_t*)data, (const uint8_t*)data + n);
    size_t o = 0;
    while (pend_.size() - o >= kBlock) { flush_block(pend_.data() + o, kBlock); o += kBlock; }
    if (o) pend_.erase(pend_.begin(), pend_.begin() + (long)o);
}

void TextWriter::close() {
    if (!f_) return;
    if (gz_) {
        if (!pend_.empty()) flush_block(pend_.data(), pend_.size());
        pend_.clear();
        flush_block(nullptr, 0);   // the empty end-of-file block
    }
    const int rc = fclose(f_);
    f_ = nullptr;
    if (rc != 0) throw std::runtime_error("[ERROR] " + path_ + " fail close.");
}

// ---- the runner ---------------------------------------------------------------------------------------------------------------
std::string BaseTypeRunner::usage() {
    return "About: Call variants and estimate allele frequency by BaseVar (B200 build: pileup on the host, statistics on the GPU).\n"
           "Usage: basevar basetype [options] <-R Fasta> <--output-vcf> <--output-cvg> [-I input] ...\n\n"
           "optional arguments:\n"
           "  -I, --input=FILE             BAM file containing reads.\n"
           "  -L, --align-file-list=FILE   BAM files list, one file per row.\n"
           "  -R, --reference FILE         Input reference fasta file.\n\n"
           "  -m, --min-af=float           Prior precision of MAF; min(-m, 100/x) is used, x = number of input files. [0.01]\n"
           "  -q, --mapq=INT               Only include reads with mapping quality >= INT. [10]\n"
           "  -B, --batch-count=INT        Accepted and ignored (no batchfiles are written). [200]\n"
           "  -t, --thread=INT             Number of host threads decoding alignments. [4]\n\n"
           "  -G, --pop-group=FILE         Calculating the allele frequency for specific population.\n"
           "  -r, --regions=chr:start-end  Comma delimited list of regions.\n"
           "  --output-vcf FILE            Output VCF file.\n"
           "  --output-cvg FILE            Output position coverage file.\n\n"
           "  --filename-has-samplename    Take the sample id from a file name like 'SampleID.xxxx.bam'.\n"
           "  --smart-rerun                Accepted and ignored.\n"
           "  --gpus=LIST                  Comma delimited CUDA devices to shard the regions over. [0]\n"
           "  --tile-sites=INT             Positions per GPU tile. [8192; fewer for cohorts of more than ~40,000 samples:\n"
           "                               2 GiB of device planes per tile]\n"
           "  --dense-upload               Upload the packed planes of a tile instead of its covered cells.\n"
           "  --workers-per-gpu=INT        Host workers per GPU, each with its own region shard (pileup, text). [thread / 2]\n"
           "  --timing                     Print the wall seconds per stage of the host pipeline (JSON, stderr).\n"
           "  --flip-log=FILE              List the positions whose LRT sat on its threshold or tied (CHROM POS FLAGS):\n"
           "                               there a call may differ from the CPU caller's, whose choice is rounding noise.\n"
           "  -h, --help                   Show this help message and exit.";
}

void BaseTypeRunner::set_arguments(int argc, char* argv[]) {
    if (argc < 2) {
        std::cout << usage() << "\n" << std::endl;
        exit(1);
    }
    static const struct option lopts[] = {
        {"input", optional_argument, NULL, 'I'},       {"align-file-list", optional_argument, NULL, 'L'},
        {"reference", required_argument, NULL, 'R'},   {"min-af", optional_argument, NULL, 'm'},
        {"mapq", optional_argument, NULL, 'q'},        {"batch-count", optional_argument, NULL, 'B'},
        {"thread", optional_argument, NULL, 't'},      {"regions", optional_argument, NULL, 'r'},
        {"positions", optional_argument, NULL, 'p'},   {"pop-group", optional_argument, NULL, 'G'},
        {"output-vcf", required_argument, NULL, '1'},  {"output-cvg", required_argument, NULL, '2'},
        {"filename-has-samplename", no_argument, NULL, '3'}, {"smart-rerun", no_argument, NULL, '4'},
        {"gpus", required_argument, NULL, '5'},        {"tile-sites", required_argument, NULL, '6'},
        {"dense-upload", no_argument, NULL, '7'},      {"flip-log", required_argument, NULL, '8'},
        {"timing", no_argument, NULL, '9'},            {"workers-per-gpu", required_argument, NULL, 'w'},
        {"help", no_argument, NULL, 'h'},              {0, 0, 0, 0}};
    BaseTypeARGS a;
    optind = 1;
    int c;
    while ((c = getopt_long(argc, argv, "I:L:R:m:q:B:t:r:G:h", lopts, NULL)) >= 0) {
        std::stringstream ss(optarg ? optarg : "");
        switch (c) {
            case 'I': a.input_bf.push_back(optarg); break;
            case 'L': a.in_bamfilelist = optarg; break;
            case 'R': a.reference = optarg; break;
            case 'm': ss >> a.min_af; break;
            case 'q': ss >> a.mapq; break;
            case 'B': ss >> a.batchcount; break;
            case 't': ss >> a.thread_num; break;
            case 'r': a.regions = optarg; break;
            case 'G': a.pop_group_file = optarg; break;
            case '1': a.output_vcf = optarg; break;
            case '2': a.output_cvg = optarg; break;
            case '3': a.filename_has_samplename = true; break;
            case '4': a.smart_rerun = true; break;
            case '5': {
                std::string tok;
                while (std::getline(ss, tok, ','))
                    if (!tok.empty()) a.devices.push_back(std::stoi(tok));
                break;
            }
            case '6': ss >> a.tile_sites; a.tile_sites_given = true; break;
            case '7': a.dense_upload = true; break;
            case '8': a.flip_log = optarg; break;
            case '9': a.timing = true; break;
            case 'w': ss >> a.workers_per_gpu; break;
            case 'h': std::cout << usage() << std::endl; exit(1);
            default: std::cerr << "Unknown argument: " << (char)c << std::endl; exit(1);
        }
    }
    set_arguments(a);
}

void BaseTypeRunner::set_arguments(const BaseTypeARGS& args) {
    args_ = args;
    if (args_.input_bf.empty() && args_.in_bamfilelist.empty())
        throw std::invalid_argument("[ERROR] Missing argument '-I/--input' or '-L/--align-file-list'");
    if (args_.reference.empty()) throw std::invalid_argument("[ERROR] Missing argument '-R/--reference'");
    if (args_.output_vcf.empty()) throw std::invalid_argument("[ERROR] Missing argument '--output-vcf'");
    if (args_.output_cvg.empty()) throw std::invalid_argument("[ERROR] Missing argument '--output-cvg'");
    if (args_.min_af <= 0) throw std::invalid_argument("[ERROR] '-m/--min-af' argument must be > 0");
    if (args_.mapq <= 0) throw std::invalid_argument("[ERROR] '-q/--mapq' argument must be > 0");
    if (args_.batchcount <= 0) throw std::invalid_argument("[ERROR] '-B/--batch-count' argument must be > 0");
    if (args_.thread_num <= 0) throw std::invalid_argument("[ERROR] '-t/--thread' argument must be > 0");
    if (args_.tile_sites == 0) throw std::invalid_argument("[ERROR] '--tile-sites' argument must be > 0");
    args_.output_vcf = std::filesystem::absolute(args_.output_vcf).lexically_normal().string();
    args_.output_cvg = std::filesystem::absolute(args_.output_cvg).lexically_normal().string();
    finish_arguments();
}

void BaseTypeRunner::finish_arguments() {
    if (!args_.in_bamfilelist.empty()) {   // first column of every row (basetype_utils.cpp:10-30)
        std::ifstream in(args_.in_bamfilelist.c_str());
        if (!in) throw std::invalid_argument("[ERROR] Cannot open file: " + args_.in_bamfilelist);
        std::string first, skip;
        for (;;) {   // as the reference reads it: a last row without a newline is not taken
            in >> first;
            if (in.eof()) break;
            std::getline(in, skip, '\n');
            args_.input_bf.push_back(first);
        }
    }
    std::cout << "[INFO] Finish loading arguments and we have " << args_.input_bf.size() << " BAM files for variants calling.\n" << std::endl;
    // the resolution of AF (cpp:122): stays a float
    args_.min_af = std::min(float(100) / args_.input_bf.size(), args_.min_af);

    reference_.reset(new Fasta(args_.reference));

    // calling intervals (cpp:314-366)
    intervals_.clear();
    if (!args_.regions.empty()) {
        std::stringstream ss(args_.regions);
        std::string rg;
        while (std::getline(ss, rg, ',')) {
            if (rg.empty()) continue;
            const size_t colon = rg.find(':');
            const std::string ref_id = rg.substr(0, colon);
            uint32_t beg = 1, end = 0;
            if (colon != std::string::npos) {
                const std::string range = rg.substr(colon + 1);
                const size_t dash = range.find('-');
                beg = (uint32_t)std::strtoul(range.substr(0, dash).c_str(), nullptr, 10);
                end = dash != std::string::npos ? (uint32_t)std::strtoul(range.substr(dash + 1).c_str(), nullptr, 10) : reference_->seq_length(ref_id);
            } else {
                end = reference_->seq_length(ref_id);
            }
            if (beg > end) throw std::invalid_argument("[ERROR] start postion is larger than end position in -r/--regions " + rg);
            intervals_.emplace_back(ref_id, beg, end);
        }
    } else {
        for (size_t i = 0; i < reference_->nseq(); ++i)
            intervals_.emplace_back(reference_->iseq_name(i), 1u, reference_->seq_length(reference_->iseq_name(i)));
    }
    std::cout << "---- Calling Intervals ----\n";
    for (size_t i = 0; i < intervals_.size(); ++i)
        std::cout << i + 1 << " - " << std::get<0>(intervals_[i]) << ":" << std::get<1>(intervals_[i]) << "-" << std::get<2>(intervals_[i]) << "\n";
    std::cout << "\n";

    {
        BamPileup ids(args_.input_bf, args_.mapq, args_.thread_num);
        samples_id_ = ids.sample_ids(args_.filename_has_samplename);
    }
    {   // duplicated sample ids only draw a warning (cpp:131-137)
        std::vector<std::string> sorted = samples_id_;
        std::sort(sorted.begin(), sorted.end());
        std::vector<std::string> dup;
        for (size_t i = 1; i < sorted.size(); ++i)
            if (sorted[i] == sorted[i - 1] && (dup.empty() || dup.back() != sorted[i])) dup.push_back(sorted[i]);
        if (!dup.empty()) std::cout << "[WARNING] Find " << dup.size() << " duplicated samples within the input bamfiles\n" << std::endl;
    }

    groups_idx_.clear();
    if (!args_.pop_group_file.empty()) {   // sample -> group, two columns (cpp:383-422)
        std::ifstream in(args_.pop_group_file.c_str());
        if (!in) throw std::invalid_argument("[ERROR] Cannot open file: " + args_.pop_group_file);
        std::map<std::string, std::string> sample2group;
        std::string sn, gn, skip;
        for (;;) {
            in >> sn >> gn;
            if (in.eof()) break;   // as the reference: a last row without a newline is not read
            sample2group[sn] = gn;
            std::getline(in, skip, '\n');
        }
        for (size_t i = 0; i < samples_id_.size(); ++i) {
            auto it = sample2group.find(samples_id_[i]);
            if (it != sample2group.end()) groups_idx_[it->second].push_back(i);
        }
    }
}

// Text a shard produces before its turn to write: kept in memory up to a bound, then in a temporary file next to the output
// (the reference spills per-task temporary files and merges them, src/basetype_utils.cpp:90-123).
namespace {
class SpillBuffer {
public:
    SpillBuffer(std::string path, size_t mem_cap) : path_(std::move(path)), cap_(mem_cap) {}
    ~SpillBuffer() { discard(); }
    SpillBuffer(const SpillBuffer&) = delete;
    SpillBuffer& operator=(const SpillBuffer&) = delete;
    bool empty() const { return mem_.empty() && !f_; }
    void append(const char* d, size_t n) {
        if (!f_ && mem_.size() + n <= cap_) { mem_.append(d, n); return; }
        if (!f_) {
            f_ = fopen(path_.c_str(), "wb+");
            if (!f_) throw std::runtime_error("[ERROR] " + path_ + " open failure.");
        }
        if (!mem_.empty()) { put(mem_.data(), mem_.size()); mem_.clear(); mem_.shrink_to_fit(); }
        put(d, n);
    }
    template <class W>
    void move_to(W& w) {   // everything, in order, into the writer; the buffer is empty afterwards
        if (f_) {
            fflush(f_);
            rewind(f_);
            std::vector<char> buf(1 << 20);
            size_t n;
            while ((n = fread(buf.data(), 1, buf.size(), f_)) > 0) w.write(buf.data(), n);
        }
        if (!mem_.empty()) w.write(mem_.data(), mem_.size());
        discard();
    }
    void discard() {
        if (f_) { fclose(f_); f_ = nullptr; std::remove(path_.c_str()); }
        mem_.clear();
    }
private:
    void put(const char* d, size_t n) {
        if (fwrite(d, 1, n, f_) != n) throw std::runtime_error("[ERROR] " + path_ + ": write failure (disk full?)");
    }
    std::string path_;
    size_t cap_;
    std::string mem_;
    FILE* f_ = nullptr;
};
}  // namespace

void BaseTypeRunner::run() {
    std::vector<std::string> add_group_info;
    for (const auto& kv : groups_idx_)
        add_group_info.push_back("##INFO=<ID=" + kv.first + "_AF,Number=A,Type=Float,Description=\"Allele frequency in the " + kv.first +
                                 " populations calculated base on LRT, in the range (0,1)\">");
    std::vector<std::string> contigs;
    for (size_t i = 0; i < reference_->nseq(); ++i) {
        const std::string& n = reference_->iseq_name(i);
        contigs.push_back("##contig=<ID=" + n + ",length=" + std::to_string(reference_->seq_length(n)) + ",assembly=" + args_.reference + ">");
    }
    const std::string ref_line = "##reference=file://" + std::filesystem::absolute(args_.reference).lexically_normal().string();

    TextWriter vcf_out(args_.output_vcf), cvg_out(args_.output_cvg);
    const std::string vh = vcf_header_define(contigs, ref_line, add_group_info, samples_id_) + "\n";
    const std::string ch = cvg_header_define() + "\n";
    vcf_out.write(vh.data(), vh.size());
    cvg_out.write(ch.data(), ch.size());

    // One host worker per region shard.  The device side of a tile takes a fraction of the time its pileup and its text take on
    // one host thread, so every GPU gets several workers (each with its own context, streams and pinned tiles), and the
    // interval is cut into that many more shards.
    std::vector<int> devices;
    {
        const std::vector<int> gpus = args_.devices.empty() ? std::vector<int>{0} : args_.devices;
        const int w = args_.workers_per_gpu > 0 ? args_.workers_per_gpu : std::max(1, args_.thread_num / 2 / (int)gpus.size());
        for (int g : gpus)
            for (int k = 0; k < w; ++k) devices.push_back(g);
    }
    const size_t N = args_.input_bf.size();
    if (!args_.tile_sites_given) {
        // A tile's device planes are 6 bytes per sample and position (base, qual, strand + the called-site planes), per slot and
        // worker: 8,192 positions of 100,000 samples are 4.9 GB, of a million samples 49 GB.  Keep a tile at 2 GiB unless told otherwise.
        const uint64_t fit = (2ull << 30) / (6ull * ((N + 15) / 16 * 16 + 16));
        args_.tile_sites = (uint32_t)std::min<uint64_t>(args_.tile_sites, std::max<uint64_t>(fit, 64));
    }
    // positions decoded per pass: bounded by the reference's own step and by the memory of the per-sample cell lists
    const uint64_t span_len = std::max<uint64_t>(args_.tile_sites, std::min<uint64_t>(PILEUP_STEP_REGION_LEN, (uint64_t)4e8 / std::max<size_t>(N, 1)));
    std::string flips;   // positions flagged NEAR_LRT / LRT_TIE, in coordinate order

    for (const auto& iv : intervals_) {
        const std::string& ref_id = std::get<0>(iv);
        const uint32_t reg_beg = std::get<1>(iv), reg_end = std::get<2>(iv);
        const std::string fa_seq = reference_->fetch(ref_id);
        // contiguous shards cut at the reference's 100-kb task boundaries (cpp:474-510), one per GPU
        const std::vector<Shard> shards = shard_region(reg_beg, (uint64_t)reg_end + 1, (int)devices.size());
        // a shard that is not next in line keeps at most 64 MiB of each text in memory; the rest waits in a temporary file
        struct ShardOut {
            std::unique_ptr<SpillBuffer> vcf, cvg;
            std::string flips;
            StageTimes times;
            std::exception_ptr err;
            uint64_t launches = 0;
        };
        std::vector<ShardOut> outs(shards.size());
        for (size_t si = 0; si < shards.size(); ++si) {
            outs[si].vcf.reset(new SpillBuffer(args_.output_vcf + ".shard" + std::to_string(si) + ".tmp", (size_t)64 << 20));
            outs[si].cvg.reset(new SpillBuffer(args_.output_cvg + ".shard" + std::to_string(si) + ".tmp", (size_t)64 << 20));
        }
        std::mutex out_mu;
        size_t next_to_write = 0;   // shard whose text streams straight to the files; later shards buffer until their turn
        std::vector<bool> done(shards.size(), false);
        auto work = [&](size_t si) {
            ShardOut& O = outs[si];
            try {
                const Shard& sh = shards[si];
                CallerOptions opt;
                opt.device = devices[(size_t)sh.gpu % devices.size()];
                opt.tile_sites = args_.tile_sites;
                opt.n_slots = 2;
                opt.em_abs_mode = args_.em_abs_mode;
                opt.sparse_upload = !args_.dense_upload;
                opt.flip_log = [&O](const char* d, size_t n) { O.flips.append(d, n); };
                opt.times = &O.times;
                auto seconds_since = [](std::chrono::steady_clock::time_point t0) {
                    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                };
                auto sink = [&out_mu, &next_to_write, si](SpillBuffer* buf, TextWriter* w) {
                    return [=, &out_mu, &next_to_write](const char* d, size_t n) {
                        std::lock_guard<std::mutex> g(out_mu);
                        if (si == next_to_write) {   // this shard's turn: earlier text first, then straight to the file
                            if (!buf->empty()) buf->move_to(*w);
                            w->write(d, n);
                        } else buf->append(d, n);
                    };
                };
                BasevarCaller caller(N, groups_idx_, (double)args_.min_af, sink(O.vcf.get(), &vcf_out), sink(O.cvg.get(), &cvg_out), opt);
                BamPileup pile(args_.input_bf, args_.mapq, std::max(1, args_.thread_num / (int)shards.size()), (int)shards.size());
                for (uint64_t sb = sh.beg; sb < sh.end; sb += span_len) {
                    const uint64_t se = std::min<uint64_t>(sb + span_len, sh.end) - 1;
                    auto t0 = std::chrono::steady_clock::now();
                    const bool any = pile.load_span(ref_id, fa_seq, reg_beg, reg_end, (uint32_t)sb, (uint32_t)se);
                    O.times.decode += seconds_since(t0);
                    if (!any) continue;
                    for (uint64_t p = sb; p <= se; p += args_.tile_sites) {
                        const uint32_t n = (uint32_t)std::min<uint64_t>(args_.tile_sites, se - p + 1);
                        TileRows rows = caller.begin_tile(n);
                        t0 = std::chrono::steady_clock::now();
                        pile.scatter((uint32_t)p, n, ref_id, fa_seq, rows);
                        O.times.scatter += seconds_since(t0);
                        caller.commit_tile();
                    }
                }
                caller.finish();
                O.launches = caller.launch_count();
            } catch (...) {
                O.err = std::current_exception();
            }
            // the files pass to the next shard that is still running; finished shards in between are written out whole
            std::lock_guard<std::mutex> g(out_mu);
            done[si] = true;
            while (next_to_write < shards.size() && done[next_to_write]) {
                ShardOut& W = outs[next_to_write];
                if (!W.err) {
                    W.vcf->move_to(vcf_out);
                    W.cvg->move_to(cvg_out);
                }
                W.vcf->discard(); W.cvg->discard();
                ++next_to_write;
            }
        };
        if (shards.size() == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (size_t si = 0; si < shards.size(); ++si) th.emplace_back(work, si);
            for (auto& t : th) t.join();
        }
        for (auto& O : outs) {
            launches_ += O.launches;
            if (O.err) std::rethrow_exception(O.err);
            flips += O.flips;   // shards are in coordinate order
            times_.add(O.times);
        }
    }
    vcf_out.close();
    cvg_out.close();
    if (!flips.empty()) {
        const size_t n_flagged = (size_t)std::count(flips.begin(), flips.end(), '\n');
        std::cerr << "[INFO] " << n_flagged << " position(s) with an LRT statistic on its threshold or a tie between candidate allele sets"
                  << (args_.flip_log.empty() ? " (--flip-log=FILE lists them)" : ": listed in " + args_.flip_log) << std::endl;
    }
    if (args_.timing) {   // summed over the host workers (one per GPU shard); the shards run side by side
        char buf[512];
        snprintf(buf, sizeof(buf), "{\"stage_seconds\": {\"bam_decode\": %.3f, \"scatter\": %.3f, \"tile_reset\": %.3f, \"encode_submit\": %.3f, "
                 "\"gpu_wait\": %.3f, \"text\": %.3f}, \"workers\": %zu, \"kernel_launches\": %llu}",
                 times_.decode, times_.scatter, times_.tile_reset, times_.encode_submit, times_.gpu_wait, times_.text, devices.size(),
                 (unsigned long long)launches_);
        std::cerr << buf << std::endl;
    }
    if (!args_.flip_log.empty()) {
        std::ofstream fl(args_.flip_log.c_str());
        if (!fl) throw std::invalid_argument("[ERROR] Cannot open file: " + args_.flip_log);
        fl << "#CHROM\tPOS\tFLAGS\n" << flips;
    }
}

}  // namespace bvhost
