// bv_bam.cpp -- see bv_bam.hpp.  BGZF / BAM / BAI / FASTA readers written against the published format specifications.
#include "bv_bam.hpp"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <stdexcept>

namespace bvhost {

namespace {
inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | p[1] << 8); }
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | (uint64_t)le32(p + 4) << 32; }

size_t pread_full(int fd, void* dst, size_t n, uint64_t off) {
    size_t got = 0;
    while (got < n) {
        const ssize_t r = ::pread(fd, (char*)dst + got, n - got, (off_t)(off + got));
        if (r < 0) throw std::runtime_error("[ERROR] read failure");
        if (r == 0) break;
        got += (size_t)r;
    }
    return got;
}
}  // namespace

// ---- BGZF ---------------------------------------------------------------------------------------------------------------
BgzfReader::BgzfReader(const std::string& path) : path_(path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) throw std::runtime_error("[ERROR] " + path + " open failure.");
    cbuf_.resize(65536);
    ubuf_.resize(65536);
    memset(&zs_, 0, sizeof(zs_));
    if (inflateInit2(&zs_, -15) != Z_OK) {
        ::close(fd_);
        throw std::runtime_error("[ERROR] zlib inflateInit2 failed");
    }
    zs_ready_ = true;
}

BgzfReader::~BgzfReader() {
    if (zs_ready_) inflateEnd(&zs_);
    if (fd_ >= 0) ::close(fd_);
}

bool BgzfReader::load_block() {
    for (;;) {   // empty blocks (the end-of-file marker, or any in the middle) are skipped
        uint8_t h[18];
        const size_t got = pread_full(fd_, h, 12, next_addr_);
        if (got == 0) { block_addr_ = next_addr_; ulen_ = upos_ = 0; return false; }
        if (got < 12 || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4))
            throw std::runtime_error("[ERROR] " + path_ + " is not a BGZF file (is it BAM?)");
        const unsigned xlen = le16(h + 10);
        std::vector<uint8_t> extra(xlen);
        if (pread_full(fd_, extra.data(), xlen, next_addr_ + 12) != xlen) throw std::runtime_error("[ERROR] " + path_ + ": truncated BGZF block");
        int bsize = -1;
        for (unsigned o = 0; o + 4 <= xlen;) {
            const unsigned slen = le16(extra.data() + o + 2);
            if (extra[o] == 'B' && extra[o + 1] == 'C' && slen == 2 && o + 6 <= xlen) bsize = le16(extra.data() + o + 4);
            o += 4 + slen;
        }
        if (bsize < 0) throw std::runtime_error("[ERROR] " + path_ + ": gzip member without the BGZF block size field");
        const size_t total = (size_t)bsize + 1;
        const size_t head = 12 + xlen;
        if (total < head + 8) throw std::runtime_error("[ERROR] " + path_ + ": corrupt BGZF block");
        const size_t clen = total - head - 8;
        if (cbuf_.size() < clen + 8) cbuf_.resize(clen + 8);
        if (pread_full(fd_, cbuf_.data(), clen + 8, next_addr_ + head) != clen + 8) throw std::runtime_error("[ERROR] " + path_ + ": truncated BGZF block");
        const uint32_t isize = le32(cbuf_.data() + clen + 4);
        if (isize > 65536) throw std::runtime_error("[ERROR] " + path_ + ": corrupt BGZF block");
        block_addr_ = next_addr_;
        next_addr_ += total;
        upos_ = 0;
        ulen_ = isize;
        if (isize == 0) continue;
        inflateReset(&zs_);
        zs_.next_in = cbuf_.data();
        zs_.avail_in = (uInt)clen;
        zs_.next_out = ubuf_.data();
        zs_.avail_out = (uInt)ubuf_.size();
        const int rc = inflate(&zs_, Z_FINISH);
        if (rc != Z_STREAM_END || zs_.total_out != isize) throw std::runtime_error("[ERROR] " + path_ + ": BGZF inflate failed");
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), ubuf_.data(), isize) != le32(cbuf_.data() + clen))
            throw std::runtime_error("[ERROR] " + path_ + ": BGZF block checksum mismatch");
        return true;
    }
}

void BgzfReader::seek(uint64_t voffset) {
    const uint64_t addr = voffset >> 16;
    const size_t off = (size_t)(voffset & 0xffff);
    if (ulen_ == 0 || addr != block_addr_) {
        next_addr_ = addr;
        if (!load_block()) { upos_ = ulen_ = 0; return; }
        // load_block() may have skipped empty blocks: the offset only applies to the block that was asked for
        if (block_addr_ != addr) { upos_ = 0; return; }
    }
    upos_ = off > ulen_ ? ulen_ : off;
}

uint64_t BgzfReader::tell() const {
    if (upos_ >= ulen_) return next_addr_ << 16;
    return block_addr_ << 16 | (uint64_t)upos_;
}

std::vector<std::pair<uint64_t, uint64_t>> BgzfReader::block_table() const {
    std::vector<std::pair<uint64_t, uint64_t>> t;
    uint64_t addr = 0, uoff = 0;
    for (;;) {
        uint8_t h[12];
        const size_t got = pread_full(fd_, h, 12, addr);
        if (got == 0) break;
        if (got < 12 || h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4))
            throw std::runtime_error("[ERROR] " + path_ + " is not a BGZF file (compress FASTA with bgzip, not gzip)");
        const unsigned xlen = le16(h + 10);
        std::vector<uint8_t> extra(xlen);
        if (pread_full(fd_, extra.data(), xlen, addr + 12) != xlen) throw std::runtime_error("[ERROR] " + path_ + ": truncated BGZF block");
        int bsize = -1;
        for (unsigned o = 0; o + 4 <= xlen;) {
            const unsigned slen = le16(extra.data() + o + 2);
            if (extra[o] == 'B' && extra[o + 1] == 'C' && slen == 2 && o + 6 <= xlen) bsize = le16(extra.data() + o + 4);
            o += 4 + slen;
        }
        if (bsize < 0) throw std::runtime_error("[ERROR] " + path_ + " is not a BGZF file (compress FASTA with bgzip, not gzip)");
        uint8_t tail[4];
        if (pread_full(fd_, tail, 4, addr + (uint64_t)bsize + 1 - 4) != 4) throw std::runtime_error("[ERROR] " + path_ + ": truncated BGZF block");
        const uint32_t isize = le32(tail);
        if (isize) t.emplace_back(addr, uoff);
        uoff += isize;
        addr += (uint64_t)bsize + 1;
    }
    return t;
}

size_t BgzfReader::read(void* dst, size_t n) {
    size_t got = 0;
    while (got < n) {
        if (upos_ >= ulen_ && !load_block()) break;
        const size_t k = std::min(n - got, ulen_ - upos_);
        memcpy((char*)dst + got, ubuf_.data() + upos_, k);
        upos_ += k;
        got += k;
    }
    return got;
}

// ---- BAM ----------------------------------------------------------------------------------------------------------------
BamReader::BamReader(const std::string& path) : bgzf_(path) {
    uint8_t b[8];
    if (bgzf_.read(b, 8) != 8 || memcmp(b, "BAM\1", 4) != 0)
        throw std::runtime_error("[ERROR] " + path + " is not a BAM file (CRAM and SAM input are not supported)");
    const uint32_t l_text = le32(b + 4);
    text_.resize(l_text);
    if (bgzf_.read(&text_[0], l_text) != l_text) throw std::runtime_error("[ERROR] " + path + ": truncated BAM header");
    while (!text_.empty() && text_.back() == '\0') text_.pop_back();
    if (bgzf_.read(b, 4) != 4) throw std::runtime_error("[ERROR] " + path + ": truncated BAM header");
    const uint32_t n_ref = le32(b);
    for (uint32_t i = 0; i < n_ref; ++i) {
        if (bgzf_.read(b, 4) != 4) throw std::runtime_error("[ERROR] " + path + ": truncated BAM header");
        const uint32_t l_name = le32(b);
        std::string name(l_name, '\0');
        if (bgzf_.read(&name[0], l_name) != l_name || bgzf_.read(b, 4) != 4) throw std::runtime_error("[ERROR] " + path + ": truncated BAM header");
        while (!name.empty() && name.back() == '\0') name.pop_back();
        ref_names_.push_back(name);
        ref_lens_.push_back((int64_t)le32(b));
    }
    first_record_voffset_ = bgzf_.tell();
}

int BamReader::name2id(const std::string& name) const {
    for (size_t i = 0; i < ref_names_.size(); ++i)
        if (ref_names_[i] == name) return (int)i;
    return -1;
}

bool BamReader::sample_name(std::string& out) const {
    // The reference's loop over the @RG lines cannot advance (its status variable is unsigned, src/bam_header.cpp:63-72):
    // only the first @RG line is ever looked at.
    size_t p = 0;
    while (p < text_.size()) {
        size_t e = text_.find('\n', p);
        if (e == std::string::npos) e = text_.size();
        if (e - p >= 3 && text_.compare(p, 3, "@RG") == 0 && (e - p == 3 || text_[p + 3] == '\t')) {
            size_t f = p + 3;
            while (f < e) {   // tab-separated TAG:VALUE fields
                const size_t fe = std::min(text_.find('\t', f + 1), e);
                if (fe - f >= 4 && text_.compare(f + 1, 3, "SM:") == 0) { out = text_.substr(f + 4, fe - f - 4); return true; }
                f = fe;
            }
            return false;
        }
        p = e + 1;
    }
    return false;
}

void BamReader::load_index() {
    if (index_loaded_) return;
    const std::string& fn = bgzf_.path();
    // fn.bai, fn with its extension replaced by .bai, then the same two with .csi (hts_idx_load's order for BAM)
    const size_t dot = fn.rfind('.');
    const std::string stem = dot == std::string::npos ? fn : fn.substr(0, dot);
    const std::string cand[4] = {fn + ".bai", stem + ".bai", fn + ".csi", stem + ".csi"};
    std::string idx;
    struct stat st;
    for (const std::string& c : cand)
        if (::stat(c.c_str(), &st) == 0) { idx = c; break; }
    if (idx.empty()) throw std::runtime_error("[ERROR] could not load the index of " + fn + " (.bai or .csi)");
    std::vector<uint8_t> d;
    if (idx.size() > 4 && idx.compare(idx.size() - 4, 4, ".csi") == 0) {
        // a CSI file is BGZF compressed (SAM specification, CSIv1)
        BgzfReader z(idx);
        std::vector<uint8_t> buf(1 << 16);
        size_t n;
        while ((n = z.read(buf.data(), buf.size())) > 0) d.insert(d.end(), buf.begin(), buf.begin() + n);
    } else {
        d.resize((size_t)st.st_size);
        const int fd = ::open(idx.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("[ERROR] " + idx + " open failure.");
        const size_t got = pread_full(fd, d.data(), d.size(), 0);
        ::close(fd);
        if (got != d.size()) throw std::runtime_error("[ERROR] " + idx + ": short read");
    }
    auto need = [&](size_t o, size_t n) { if (o + n > d.size()) throw std::runtime_error("[ERROR] " + idx + ": truncated index"); };
    need(0, 8);
    size_t o;
    bool csi = false;
    if (memcmp(d.data(), "BAI\1", 4) == 0) {
        idx_min_shift_ = 14; idx_depth_ = 5;
        o = 4;
    } else if (memcmp(d.data(), "CSI\1", 4) == 0) {
        need(4, 12);
        idx_min_shift_ = (int)le32(d.data() + 4); idx_depth_ = (int)le32(d.data() + 8);
        const uint32_t l_aux = le32(d.data() + 12);
        if (idx_min_shift_ < 0 || idx_min_shift_ > 30 || idx_depth_ < 0 || idx_depth_ > 9) throw std::runtime_error("[ERROR] " + idx + ": unsupported CSI geometry");
        o = 16 + (size_t)l_aux;
        csi = true;
    } else {
        throw std::runtime_error("[ERROR] " + idx + " is neither a BAI nor a CSI index");
    }
    need(o, 4);
    const uint32_t n_ref = le32(d.data() + o); o += 4;
    const uint32_t meta_bin = (uint32_t)((((uint64_t)1 << (3 * (idx_depth_ + 1))) - 1) / 7 + 1);   // 37450 for BAI: metadata, not chunks
    index_.assign(n_ref, RefIndex());
    for (uint32_t r = 0; r < n_ref; ++r) {
        need(o, 4);
        const uint32_t n_bin = le32(d.data() + o); o += 4;
        RefIndex& R = index_[r];
        R.bins.reserve(n_bin);
        for (uint32_t k = 0; k < n_bin; ++k) {
            need(o, csi ? 16 : 8);
            const uint32_t bin = le32(d.data() + o);
            if (csi) o += 8;   // loffset: the per-bin form of the linear index; not needed for correctness (records are filtered)
            const uint32_t n_chunk = le32(d.data() + o + 4); o += 8;
            need(o, (size_t)n_chunk * 16);
            std::vector<Chunk> cs(n_chunk);
            for (uint32_t c = 0; c < n_chunk; ++c) { cs[c].beg = le64(d.data() + o); cs[c].end = le64(d.data() + o + 8); o += 16; }
            if (bin != meta_bin) R.bins.emplace_back(bin, std::move(cs));
        }
        std::sort(R.bins.begin(), R.bins.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        if (!csi) {
            need(o, 4);
            const uint32_t n_intv = le32(d.data() + o); o += 4;
            need(o, (size_t)n_intv * 8);
            R.linear.resize(n_intv);
            for (uint32_t i = 0; i < n_intv; ++i) { R.linear[i] = le64(d.data() + o); o += 8; }
        }
    }
    index_loaded_ = true;
}

void BamReader::query(int tid, int64_t beg0, int64_t end0) {
    load_index();
    querying_ = true;
    finished_ = false;
    in_chunk_ = false;
    chunk_i_ = 0;
    chunks_.clear();
    q_tid_ = tid;
    q_beg_ = beg0 < 0 ? 0 : beg0;
    q_end_ = end0;
    const int64_t max_end = (int64_t)1 << (idx_min_shift_ + 3 * idx_depth_);   // 2^29 for BAI
    if (q_end_ > max_end) q_end_ = max_end;
    if (tid < 0 || (size_t)tid >= index_.size() || q_beg_ >= q_end_) { finished_ = true; return; }
    const RefIndex& R = index_[(size_t)tid];
    // bins that may hold records overlapping [beg, end): SAM specification section 5.3 (reg2bins), for any (min_shift, depth)
    std::vector<uint32_t> bins;
    const int64_t b = q_beg_, e = q_end_ - 1;
    for (int l = 0; l <= idx_depth_; ++l) {
        const int sh = idx_min_shift_ + 3 * (idx_depth_ - l);
        const int64_t t = (((int64_t)1 << (3 * l)) - 1) / 7;
        for (int64_t k = t + (b >> sh); k <= t + (e >> sh); ++k) bins.push_back((uint32_t)k);
    }
    // records starting before this offset end before the 16-kb window of beg: they cannot overlap
    uint64_t min_off = 0;
    if (!R.linear.empty()) {
        const size_t w = (size_t)(b >> 14);
        min_off = w < R.linear.size() ? R.linear[w] : R.linear.back();
    }
    for (uint32_t bin : bins) {
        auto it = std::lower_bound(R.bins.begin(), R.bins.end(), bin, [](const auto& a, uint32_t v) { return a.first < v; });
        if (it == R.bins.end() || it->first != bin) continue;
        for (const Chunk& c : it->second)
            if (c.end > min_off) chunks_.push_back(c);
    }
    if (chunks_.empty()) { finished_ = true; return; }
    std::sort(chunks_.begin(), chunks_.end(), [](const Chunk& x, const Chunk& y) { return x.beg < y.beg; });
    size_t m = 0;
    for (size_t i = 1; i < chunks_.size(); ++i) {
        if (chunks_[i].beg <= chunks_[m].end) chunks_[m].end = std::max(chunks_[m].end, chunks_[i].end);
        else chunks_[++m] = chunks_[i];
    }
    chunks_.resize(m + 1);
}

bool BamReader::read_record(BamRec& rec) {
    uint8_t b4[4];
    const size_t got = bgzf_.read(b4, 4);
    if (got == 0) return false;
    if (got != 4) throw std::runtime_error("[ERROR] " + bgzf_.path() + ": truncated BAM record");
    const uint32_t block_size = le32(b4);
    if (block_size < 32) throw std::runtime_error("[ERROR] " + bgzf_.path() + ": corrupt BAM record");
    buf_.resize(block_size);
    if (bgzf_.read(buf_.data(), block_size) != block_size) throw std::runtime_error("[ERROR] " + bgzf_.path() + ": truncated BAM record");
    const uint8_t* p = buf_.data();
    rec.tid = (int32_t)le32(p);
    rec.pos = (int32_t)le32(p + 4);
    const unsigned l_read_name = p[8];
    rec.mapq = p[9];
    const unsigned n_cigar = le16(p + 12);
    rec.flag = le16(p + 14);
    rec.l_seq = (int32_t)le32(p + 16);
    if (rec.l_seq < 0) throw std::runtime_error("[ERROR] " + bgzf_.path() + ": corrupt BAM record");
    size_t o = 32 + l_read_name;
    const size_t seq_bytes = ((size_t)rec.l_seq + 1) / 2;
    if (o + 4 * (size_t)n_cigar + seq_bytes + (size_t)rec.l_seq > block_size) throw std::runtime_error("[ERROR] " + bgzf_.path() + ": corrupt BAM record");
    rec.cigar.resize(n_cigar);
    int64_t rlen = 0;
    for (unsigned i = 0; i < n_cigar; ++i) {
        const uint32_t c = le32(p + o + 4 * i);
        rec.cigar[i] = c;
        const unsigned op = c & 0xf;
        if (op == CIG_M || op == CIG_D || op == CIG_N || op == CIG_EQ || op == CIG_X) rlen += c >> 4;
    }
    o += 4 * (size_t)n_cigar;
    rec.seq.assign(p + o, p + o + seq_bytes);
    o += seq_bytes;
    rec.qual.assign(p + o, p + o + rec.l_seq);
    rec.end = (!rec.is_mapped() || n_cigar == 0 || rlen == 0) ? rec.pos + 1 : (int32_t)(rec.pos + rlen);
    return true;
}

bool BamReader::next(BamRec& rec) {
    if (!querying_) return read_record(rec);
    while (!finished_) {
        if (!in_chunk_) {
            if (chunk_i_ >= chunks_.size()) { finished_ = true; break; }
            bgzf_.seek(chunks_[chunk_i_].beg);
            in_chunk_ = true;
        }
        if (bgzf_.tell() >= chunks_[chunk_i_].end) { in_chunk_ = false; ++chunk_i_; continue; }
        if (!read_record(rec)) { finished_ = true; break; }
        if (rec.tid != q_tid_ || rec.pos >= q_end_) { finished_ = true; break; }   // sorted file: nothing further can overlap
        if (rec.end > q_beg_) return true;
    }
    return false;
}

// ---- FASTA --------------------------------------------------------------------------------------------------------------
// A ".gz" reference must be BGZF (bgzip), as for htslib's faidx (src/fasta.cpp:9-48 -> fai_load): offsets in the .fai index
// are offsets into the UNCOMPRESSED text; the block table maps them to BGZF virtual offsets.
Fasta::Fasta(const std::string& path) : path_(path) {
    {
        const int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) throw std::invalid_argument("[ERROR] " + path + " open failure.");
        uint8_t h[2] = {0, 0};
        const size_t got = pread_full(fd, h, 2, 0);
        ::close(fd);
        bgzf_ = got == 2 && h[0] == 31 && h[1] == 139;
    }
    if (bgzf_) blocks_ = BgzfReader(path).block_table();
    std::ifstream fai(path + ".fai");
    if (!fai) { build_index(); return; }
    std::string line;
    while (std::getline(fai, line)) {
        if (line.empty()) continue;
        std::vector<std::string> col;
        size_t p = 0;
        for (;;) {
            const size_t e = line.find('\t', p);
            col.push_back(line.substr(p, e == std::string::npos ? std::string::npos : e - p));
            if (e == std::string::npos) break;
            p = e + 1;
        }
        if (col.size() < 5) throw std::invalid_argument("[ERROR] malformed FASTA index: " + path + ".fai");
        names_.push_back(col[0]);
        entries_.push_back(Entry{std::stoull(col[1]), std::stoull(col[2]), std::stoull(col[3]), std::stoull(col[4])});
    }
}

void Fasta::build_index() {
    FILE* f = nullptr;
    std::unique_ptr<BgzfReader> z;
    if (bgzf_) z.reset(new BgzfReader(path_));
    else {
        f = fopen(path_.c_str(), "rb");
        if (!f) throw std::invalid_argument("[ERROR] " + path_ + " open failure.");
    }
    std::vector<char> buf(1 << 20);
    uint64_t off = 0;
    bool in_name = false, at_line_start = true;
    std::string name;
    Entry cur{0, 0, 0, 0};
    bool have = false;
    uint64_t line_bases = 0, line_len = 0;
    bool first_line_done = false;
    auto end_line = [&]() {
        if (have && !first_line_done && line_len > 0) { cur.line_bases = line_bases; cur.line_width = line_len; first_line_done = true; }
        line_bases = line_len = 0;
    };
    auto flush = [&]() {
        if (!have) return;
        if (!first_line_done) { cur.line_bases = line_bases; cur.line_width = line_len ? line_len : line_bases; }
        names_.push_back(name);
        entries_.push_back(cur);
    };
    size_t n;
    while ((n = z ? z->read(buf.data(), buf.size()) : fread(buf.data(), 1, buf.size(), f)) > 0) {
        for (size_t i = 0; i < n; ++i, ++off) {
            const char c = buf[i];
            if (in_name) {
                if (c == '\n') {
                    in_name = false; at_line_start = true;
                    const size_t sp = name.find_first_of(" \t\r");
                    if (sp != std::string::npos) name.resize(sp);
                    cur = Entry{0, off + 1, 0, 0};
                    have = true; first_line_done = false; line_bases = line_len = 0;
                } else name += c;
                continue;
            }
            if (at_line_start && c == '>') { end_line(); flush(); have = false; name.clear(); in_name = true; continue; }
            at_line_start = false;
            ++line_len;
            if (c == '\n') { end_line(); at_line_start = true; }
            else if (c != '\r') { ++line_bases; if (have) ++cur.length; }
        }
    }
    end_line();
    flush();
    if (f) fclose(f);
}

const Fasta::Entry& Fasta::entry(const std::string& name) const {
    for (size_t i = 0; i < names_.size(); ++i)
        if (names_[i] == name) return entries_[i];
    throw std::invalid_argument("Fasta::fetch - Fail to fetch sequence.");   // src/fasta.cpp:58
}

bool Fasta::has_seq(const std::string& name) const { return std::find(names_.begin(), names_.end(), name) != names_.end(); }

uint32_t Fasta::seq_length(const std::string& name) const { return (uint32_t)entry(name).length; }

std::string Fasta::fetch(const std::string& name) const {
    const Entry& e = entry(name);
    std::string out;
    out.reserve(e.length);
    if (e.length == 0) throw std::invalid_argument("Fasta::fetch - Fetch empty sequence on " + name);   // src/fasta.cpp:65
    const uint64_t lines = e.line_bases ? (e.length + e.line_bases - 1) / e.line_bases : 1;
    const uint64_t span = e.length + lines * (e.line_width - e.line_bases);
    std::vector<char> buf(1 << 22);
    uint64_t done = 0;
    std::unique_ptr<BgzfReader> z;
    int fd = -1;
    if (bgzf_) {
        z.reset(new BgzfReader(path_));
        auto it = std::upper_bound(blocks_.begin(), blocks_.end(), e.offset, [](uint64_t v, const std::pair<uint64_t, uint64_t>& b) { return v < b.second; });
        if (it == blocks_.begin()) throw std::invalid_argument("Fasta::fetch - Fail to fetch sequence.");
        --it;
        z->seek(it->first << 16 | (e.offset - it->second));
    } else {
        fd = ::open(path_.c_str(), O_RDONLY);
        if (fd < 0) throw std::invalid_argument("[ERROR] " + path_ + " open failure.");
    }
    // the index gives the line geometry: whole runs of bases are appended at once, line terminators are stepped over
    const uint64_t lb = e.line_bases ? e.line_bases : e.length, lw = e.line_width > lb ? e.line_width : lb + 1;
    while (done < span && out.size() < e.length) {
        const size_t want = (size_t)std::min<uint64_t>(buf.size(), span - done);
        const size_t got = z ? z->read(buf.data(), want) : pread_full(fd, buf.data(), want, e.offset + done);
        if (got == 0) break;
        size_t i = 0;
        while (i < got && out.size() < e.length) {
            const uint64_t in_line = (done + i) % lw;   // position inside the current line (bases first, then the terminator)
            if (in_line < lb) {
                const size_t k = (size_t)std::min<uint64_t>(std::min<uint64_t>(lb - in_line, got - i), e.length - out.size());
                out.append(buf.data() + i, k);
                i += k;
            } else {
                i += (size_t)std::min<uint64_t>(lw - in_line, got - i);
            }
        }
        done += got;
    }
    if (fd >= 0) ::close(fd);
    return out;
}

}  // namespace bvhost
