// CPU-only tests of the C++ host layer (no CUDA call except the one that must fail): packer, encodings, error
// messages, BaseType getters from a record, strand_bias, region sharding.  Prints one line per check; exit code 0 = ok.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../basevar_b200/host/bv_host.hpp"

using namespace bvhost;

static int g_fail = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
    } while (0)

template <class F>
static std::string thrown(F&& f) {
    try { f(); } catch (const std::exception& e) { return e.what(); }
    return "";
}

static BatchInfo make_bi(const char* ref, const std::vector<std::string>& bases, const std::string& quals, const std::string& strands) {
    BatchInfo bi;
    bi.n = bases.size(); bi.ref_id = "chr1"; bi.ref_base = ref; bi.ref_pos = 1234; bi.depth = 0;
    bi.align_bases = bases;
    bi.align_base_quals.assign(quals.begin(), quals.end());
    bi.map_strands.assign(strands.begin(), strands.end());
    bi.mapqs.assign(bases.size(), 60);
    bi.base_pos_ranks.assign(bases.size(), 1);
    return bi;
}

int main() {
    // ---- encodings (src/basetype.cpp:50-64) ----
    CHECK(encode_base("A") == BV_BASE_A && encode_base("C") == BV_BASE_C && encode_base("G") == BV_BASE_G && encode_base("T") == BV_BASE_T);
    CHECK(encode_base("N") == BV_BASE_N && encode_base("+ACG") == BV_BASE_INS && encode_base("-AC") == BV_BASE_DEL);
    CHECK(encode_base("R") == BV_BASE_OTHER && encode_base("a") == BV_BASE_OTHER);
    CHECK(thrown([] { encode_base("AC"); }) == "[ERROR] Why dose the size of aligned base is not 1? Check: AC");
    CHECK(encode_strand('+') == BV_STRAND_FWD && encode_strand('-') == BV_STRAND_REV && encode_strand('.') == BV_STRAND_NONE && encode_strand('*') == BV_STRAND_NONE);
    CHECK(cli_min_af(0.01f, 1000) == 0.01f && cli_min_af(0.01f, 100000) == 100.0f / 100000 && (double)cli_min_af(0.01f, 1000) == 0.0099999997764825821);

    // ---- packer ----
    {
        TilePacker pk(5, 4, /*pinned=*/false);
        CHECK(pk.pitch() == 16 && pk.capacity() == 4 && pk.n_sites() == 0);
        BatchInfo bi = make_bi("g", {"A", "N", "+AT", "T", "R"}, "?!?I5", "+.+-+");
        pk.add_site(bi);
        bv_tile t = pk.tile();
        CHECK(t.n_sites == 1 && t.n_samples == 5 && t.pitch == 16 && t.location == BV_LOC_HOST);
        const uint8_t eb[5] = {0, 5, 6, 3, 4}, eq[5] = {30, 0, 30, 40, 20}, es[5] = {0, 2, 0, 1, 0};
        CHECK(memcmp(t.base, eb, 5) == 0 && memcmp(t.qual, eq, 5) == 0 && memcmp(t.strand, es, 5) == 0 && t.ref_base[0] == 'g');
        for (int i = 5; i < 16; ++i) CHECK(t.base[i] == BV_BASE_N && t.strand[i] == BV_STRAND_NONE && t.qual[i] == 0);
        BatchInfo bad = make_bi("A", {"A", "AC", "N", "N", "N"}, "?????", "++...");
        CHECK(thrown([&] { pk.add_site(bad); }).find("Why dose the size of aligned base is not 1? Check: AC") != std::string::npos);
        BatchInfo wrong_n = make_bi("A", {"A"}, "?", "+");
        CHECK(!thrown([&] { pk.add_site(wrong_n); }).empty());
        pk.clear();
        for (int i = 0; i < 4; ++i) pk.add_site(bi);
        CHECK(!thrown([&] { pk.add_site(bi); }).empty());   // full
    }

    // ---- BaseType from a record, getters and their errors (src/basetype.h:120-151) ----
    {
        bv_site_out r;
        memset(&r, 0, sizeof(r));
        r.depth[0] = 7; r.depth[2] = 3; r.depth_other = 1;
        r.fwd[0] = 3; r.rev[0] = 4; r.fwd[2] = 1; r.rev[2] = 2;
        r.n_alt = 1; r.alt[0] = 2; r.af[0] = 0.2998664889880594; r.qual = 86.650670492227547; r.n_active = 2;
        r.fs_cvg = 0.0; r.fs_vcf = 0.0;
        BatchInfo bi = make_bi("A", {"A"}, "?", "+");
        BaseType bt(&bi, r);
        bt.lrt();
        CHECK(bt.get_ref_id() == "chr1" && bt.get_ref_pos() == 1234 && bt.get_ref_base() == "A" && bt.is_only_snp());
        CHECK(bt.get_alt_bases().size() == 1 && bt.get_alt_bases()[0] == 'G');
        CHECK(bt.get_lrt_af('G') == 0.2998664889880594 && bt.get_var_qual() == 86.650670492227547);
        CHECK(bt.get_total_depth() == 11 && bt.get_base_depth('A') == 7.0 && bt.get_base_depth('G') == 3.0 && bt.get_base_depth('T') == 0.0);
        CHECK(thrown([&] { bt.get_lrt_af('C'); }).find("[ERROR] out_of_range::") == 0);
        CHECK(thrown([&] { bt.get_base_depth('N'); }).find("'N' not found.") != std::string::npos);
        CHECK(!thrown([&] { bt.lrt({'A', 'G'}); }).empty());
        // strand_bias: SB = 3,4,1,2; SOR = (3*2)/(4*1) = 1.5
        StrandBiasInfo s = strand_bias('A', "G", r);
        CHECK(s.ref_fwd == 3 && s.ref_rev == 4 && s.alt_fwd == 1 && s.alt_rev == 2 && s.sor == 1.5 && s.fs == 0.0);
        StrandBiasInfo s2 = strand_bias('A', "CGT", r);   // CVG row: all non-ref bases
        CHECK(s2.alt_fwd == 1 && s2.alt_rev == 2 && s2.sor == 1.5);
        StrandBiasInfo s3 = strand_bias('A', "C", r);     // no reads of C: single possible table
        CHECK(s3.alt_fwd == 0 && s3.alt_rev == 0 && s3.fs == 0.0 && s3.sor == 10000);
        r.flags = BV_FLAG_BAD_STRAND;
        CHECK(thrown([&] { strand_bias('A', "G", r); }).find("[ERROR] Get strange strand symbol") == 0);
        r.flags = BV_FLAG_ZERO_SUBSET;
        BaseType bz(&bi, r);
        CHECK(thrown([&] { bz.lrt(); }).find("The sum of frequence of active bases must always > 0") != std::string::npos);
    }

    // ---- region sharding (src/basetype_caller.cpp:469-525: 100-kb tasks; shards are whole tasks, in order) ----
    {
        for (int g : {1, 2, 3, 4, 8})
            for (uint64_t len : {1ull, 99999ull, 100000ull, 100001ull, 64000000ull, 6400001ull, 250000ull}) {
                const uint64_t beg = 5246595;
                std::vector<Shard> sh = shard_region(beg, beg + len, g);
                CHECK(!sh.empty() && (int)sh.size() <= g && sh.front().beg == beg && sh.back().end == beg + len);
                uint64_t mn = ~0ull, mx = 0;
                for (size_t i = 0; i < sh.size(); ++i) {
                    CHECK(sh[i].gpu == (int)i && sh[i].end > sh[i].beg);
                    if (i) CHECK(sh[i].beg == sh[i - 1].end);
                    CHECK((sh[i].beg - beg) % 100000 == 0);
                    const uint64_t tasks = (sh[i].end - sh[i].beg + 99999) / 100000;
                    mn = std::min(mn, tasks); mx = std::max(mx, tasks);
                }
                CHECK(mx - mn <= 1);
            }
        CHECK(shard_region(10, 10, 4).empty());
        std::vector<Shard> one = shard_region(0, 50000, 8);
        CHECK(one.size() == 1);   // a region shorter than one task is one shard (the reference: one thread)
    }

    // ---- no CPU fallback: a context needs a CUDA device ----
    {
        const char* expect_gpu = getenv("BV_EXPECT_GPU");
        if (!expect_gpu) {
            std::string e = thrown([] { Context c(0, 0.01f, 16, 16, 1); });
            CHECK(e.find("bv_create") != std::string::npos && e.find("no CPU fallback") != std::string::npos);
        }
    }
    printf(g_fail ? "FAILED %d checks\n" : "ALL OK\n", g_fail);
    return g_fail ? 1 : 0;
}
