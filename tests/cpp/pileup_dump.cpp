// pileup_dump -- CPU-only test program: BAM files -> packed tile rows (bv_pileup.hpp) -> the text rows the reference's
// __write_record_to_batchfile would have written for the same calling interval (src/basetype_caller.cpp:1024-1101).
//   pileup_dump <fasta> <bam list file> <chr:beg-end> <mapq> <threads> [span_len] [tile_sites]
// Prints the batchfile header and one row per position.  No GPU is involved: this checks the packer alone.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "../../basevar_b200/host/bv_pileup.hpp"

using namespace bvhost;

// --query-check <bam> <seed> <n>: n random region queries through the index against a linear scan of the file
static int query_check(const char* bam, unsigned seed, int n) {
    BamReader scan(bam);
    std::vector<BamRec> all;
    BamRec r;
    while (scan.next(r)) all.push_back(r);
    BamReader idx(bam);
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    long total = 0;
    for (int k = 0; k < n; ++k) {
        const int tid = (int)(rnd() % (uint64_t)idx.n_ref());
        const int64_t len = idx.ref_length(tid);
        const int64_t beg = (int64_t)(rnd() % (uint64_t)len);
        const int64_t span = (k % 3 == 0) ? 1 + (int64_t)(rnd() % 50) : (k % 3 == 1) ? 1 + (int64_t)(rnd() % 5000) : 1 + (int64_t)(rnd() % (uint64_t)len);
        const int64_t end = beg + span;
        std::vector<const BamRec*> want;
        for (const BamRec& a : all)
            if (a.tid == tid && a.pos < end && a.end > beg) want.push_back(&a);
        idx.query(tid, beg, end);
        size_t i = 0;
        while (idx.next(r)) {
            if (i >= want.size() || want[i]->pos != r.pos || want[i]->end != r.end || want[i]->flag != r.flag || want[i]->mapq != r.mapq ||
                want[i]->cigar != r.cigar || want[i]->seq != r.seq || want[i]->qual != r.qual) {
                fprintf(stderr, "query %d (%d:%ld-%ld): record %zu differs from the linear scan\n", k, tid, (long)beg, (long)end, i);
                return 1;
            }
            ++i;
        }
        if (i != want.size()) { fprintf(stderr, "query %d (%d:%ld-%ld): %zu records, linear scan has %zu\n", k, tid, (long)beg, (long)end, i, want.size()); return 1; }
        total += (long)i;
    }
    printf("query-check ok: %zu records in the file, %d queries, %ld records returned\n", all.size(), n, total);
    return 0;
}

// --bgzf-roundtrip <out.gz> <bytes> <seed>: TextWriter -> file -> BgzfReader, byte for byte (also readable by gzip)
static int bgzf_roundtrip(const char* path, size_t bytes, unsigned seed) {
    std::string data(bytes, '\0');
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    for (size_t i = 0; i < bytes; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; data[i] = "ACGT\t\n0123456789"[(x >> 20) % 16]; }
    {
        TextWriter w(path);
        size_t o = 0;
        while (o < bytes) { const size_t k = std::min<size_t>(bytes - o, 1 + (x % 100000)); x ^= x << 13; x ^= x >> 7; x ^= x << 17; w.write(data.data() + o, k); o += k; }
        w.close();
    }
    BgzfReader rd(path);
    std::string back(bytes + 10, '\0');
    const size_t got = rd.read(&back[0], back.size());
    if (got != bytes || memcmp(back.data(), data.data(), bytes) != 0) { fprintf(stderr, "bgzf round trip differs (%zu of %zu bytes)\n", got, bytes); return 1; }
    printf("bgzf-roundtrip ok: %zu bytes\n", bytes);
    return 0;
}

int main(int argc, char** argv) {
    try {
        if (argc == 5 && strcmp(argv[1], "--query-check") == 0) return query_check(argv[2], (unsigned)atoi(argv[3]), atoi(argv[4]));
        if (argc == 5 && strcmp(argv[1], "--bgzf-roundtrip") == 0) return bgzf_roundtrip(argv[2], (size_t)atol(argv[3]), (unsigned)atoi(argv[4]));
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    if (argc < 6) { fprintf(stderr, "usage: pileup_dump fasta bamlist chr:beg-end mapq threads [span_len] [tile_sites]\n"); return 2; }
    try {
        Fasta fa(argv[1]);
        std::vector<std::string> bams;
        { std::ifstream in(argv[2]); std::string l; while (std::getline(in, l)) if (!l.empty()) bams.push_back(l); }
        const std::string reg = argv[3];
        const size_t colon = reg.find(':'), dash = reg.find('-', colon);
        const std::string ref_id = reg.substr(0, colon);
        const uint32_t reg_beg = (uint32_t)atol(reg.substr(colon + 1, dash - colon - 1).c_str()), reg_end = (uint32_t)atol(reg.substr(dash + 1).c_str());
        const int mapq = atoi(argv[4]), threads = atoi(argv[5]);
        const uint32_t span_len = argc > 6 ? (uint32_t)atol(argv[6]) : 500000, tile = argc > 7 ? (uint32_t)atol(argv[7]) : 1024;
        const std::string seq = fa.fetch(ref_id);
        BamPileup pile(bams, mapq, threads);
        const std::vector<std::string> ids = pile.sample_ids(false);
        std::string hdr = "##fileformat=BaseVarBatchFile_v1.0\n##SampleIDs=";
        for (size_t i = 0; i < ids.size(); ++i) hdr += (i ? "," : "") + ids[i];
        hdr += "\n#CHROM\tPOS\tREF\tDepth(CoveredSample)\tMappingQuality\tReadbases\tReadbasesQuality\tReadPositionRank\tStrand\n";
        fputs(hdr.c_str(), stdout);
        const size_t N = bams.size(), pitch = (N + 15) / 16 * 16;
        std::vector<uint8_t> base(tile * pitch), qual(tile * pitch), strand(tile * pitch), mq(tile * pitch);
        std::vector<uint16_t> rpr(tile * pitch);
        std::vector<SiteMeta> meta(tile);
        for (uint64_t sb = reg_beg; sb <= reg_end; sb += span_len) {
            const uint64_t se = std::min<uint64_t>(sb + span_len - 1, reg_end);
            pile.load_span(ref_id, seq, reg_beg, reg_end, (uint32_t)sb, (uint32_t)se);
            for (uint64_t p = sb; p <= se; p += tile) {
                const uint32_t n = (uint32_t)std::min<uint64_t>(tile, se - p + 1);
                memset(base.data(), BV_BASE_N, base.size()); memset(qual.data(), 0, qual.size());
                memset(strand.data(), BV_STRAND_NONE, strand.size()); memset(mq.data(), 0, mq.size());
                std::fill(rpr.begin(), rpr.end(), 0);
                for (auto& m : meta) { m.specials.clear(); m.odd_strands.clear(); m.depth = 0; }
                TileRows rows{base.data(), qual.data(), strand.data(), mq.data(), rpr.data(), pitch, pitch, n, meta.data(), nullptr, nullptr, nullptr};
                // the sparse transport of the same tile: its cells, expanded, must give exactly the planes
                std::vector<uint32_t> sp_start(n + 1), sp_cells, sp_aux;
                bool sp_ready = false;
                rows.site_start = sp_start.data();
                rows.reserve_cells = [&](size_t k, uint32_t** c, uint32_t** a) { sp_cells.assign(k, 0); sp_aux.assign(k, 0); *c = sp_cells.data(); *a = sp_aux.data(); };
                rows.sparse_ready = &sp_ready;
                pile.scatter((uint32_t)p, n, ref_id, seq, rows);
                {
                    if (!sp_ready || sp_start[0] != 0 || sp_start[n] != sp_cells.size()) { fprintf(stderr, "sparse tile: not filled / bad offsets\n"); return 1; }
                    std::vector<uint8_t> xb(n * pitch, BV_BASE_N), xq(n * pitch, 0), xs(n * pitch, BV_STRAND_NONE), xm(n * pitch, 0);
                    std::vector<uint16_t> xr(n * pitch, 0);
                    size_t covered = 0;
                    for (uint32_t i = 0; i < n; ++i) {
                        if (sp_start[i + 1] < sp_start[i]) { fprintf(stderr, "sparse tile: offsets descend at row %u\n", i); return 1; }
                        for (uint32_t k = sp_start[i]; k < sp_start[i + 1]; ++k) {
                            const uint32_t w = sp_cells[k], smp = w & (BV_CELL_MAX_SAMPLES - 1u);
                            const size_t at = (size_t)i * pitch + smp;
                            if (smp >= N || xs[at] != BV_STRAND_NONE || xb[at] != BV_BASE_N) { fprintf(stderr, "sparse tile: bad or repeated sample %u at row %u\n", smp, i); return 1; }
                            xb[at] = (uint8_t)((w >> 20) & 7u); xs[at] = (uint8_t)((w >> 23) & 3u); xq[at] = (uint8_t)(w >> 25);
                            xm[at] = (uint8_t)sp_aux[k]; xr[at] = (uint16_t)(sp_aux[k] >> 8);
                            ++covered;
                        }
                    }
                    if (memcmp(xb.data(), base.data(), n * pitch) || memcmp(xq.data(), qual.data(), n * pitch) || memcmp(xs.data(), strand.data(), n * pitch) ||
                        memcmp(xm.data(), mq.data(), n * pitch) || memcmp(xr.data(), rpr.data(), n * pitch * sizeof(uint16_t))) {
                        fprintf(stderr, "sparse tile at %lu: expanded cells differ from the planes (%zu cells)\n", (unsigned long)p, covered);
                        return 1;
                    }
                }
                std::string out;
                for (uint32_t i = 0; i < n; ++i) {
                    const SiteCells c{base.data() + i * pitch, qual.data() + i * pitch, strand.data() + i * pitch, (uint32_t)N};
                    out += batchfile_row(meta[i], c, mq.data() + i * pitch, rpr.data() + i * pitch);
                    out += '\n';
                }
                fwrite(out.data(), 1, out.size(), stdout);
            }
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
