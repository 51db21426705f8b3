// CPU-only tests of the per-site driver's text side (basevar_b200/host/bv_caller): number formatting, CVG / VCF rows
// from hand-made device records, headers, and the loud failure without a CUDA device.  Exit code 0 = ok.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>

#include "../../basevar_b200/host/bv_caller.hpp"

using namespace bvhost;

static int g_fail = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
    } while (0)
#define CHECK_EQ(a, b)                                                     \
    do {                                                                   \
        const std::string a_ = (a), b_ = (b);                              \
        if (a_ != b_) { printf("FAIL %s:%d\n  got  [%s]\n  want [%s]\n", __FILE__, __LINE__, a_.c_str(), b_.c_str()); ++g_fail; } \
    } while (0)

template <class F>
static std::string thrown(F&& f) {
    try { f(); } catch (const std::exception& e) { return e.what(); }
    return "";
}

int main() {
    // ---- std::to_string(double) / ostream << double --------------------------------------------------------------------
    CHECK_EQ(to_string_f(106.66461659), "106.664617");
    CHECK_EQ(to_string_f(0.0), "0.000000");
    CHECK_EQ(to_string_f(10000), "10000.000000");
    CHECK_EQ(tostring_g(0.62484312), "0.624843");
    CHECK_EQ(tostring_g(1.0), "1");
    CHECK_EQ(tostring_g(0.375), "0.375");
    CHECK_EQ(tostring_g(0.00012345678), "0.000123457");
    CHECK_EQ(tostring_g(1.5e-7), "1.5e-07");
    CHECK_EQ(tostring_g(std::numeric_limits<double>::quiet_NaN()), "-nan");

    // ---- CVG row (src/basetype_caller.cpp:1211-1260) ----------------------------------------------------------------------
    SiteMeta m;
    m.ref_id = "chr11"; m.ref_pos = 5247141; m.ref_base = "G"; m.depth = 9;
    // samples: A+ A- A+ A- A+ G+ G- G+ N  +AT(ins)
    const uint8_t base[10] = {0, 0, 0, 0, 0, 2, 2, 2, 5, 6};
    const uint8_t qual[10] = {21, 36, 37, 36, 38, 36, 34, 30, 0, 30};
    const uint8_t strand[10] = {0, 1, 0, 1, 0, 0, 1, 0, 2, 0};
    m.specials.emplace_back(9u, "+AT");
    SiteCells c{base, qual, strand, 10};
    bv_site_out rec;
    memset(&rec, 0, sizeof(rec));
    rec.depth[0] = 5; rec.depth[2] = 3;
    rec.fwd[0] = 3; rec.rev[0] = 2; rec.fwd[2] = 2; rec.rev[2] = 1;
    rec.fs_cvg = 0.0;
    CHECK_EQ(out_cvg_line(m, c, rec), "chr11\t5247141\tG\t8\t5\t0\t3\t0\t+AT|1\t0.000000\t1.333333\t2,1,3,2\n");
    {
        bv_site_out none = rec;
        none.depth[0] = none.depth[2] = 0;
        CHECK_EQ(out_cvg_line(m, c, none), "");   // no A/C/G/T read: no row
        bv_site_out bad = rec;
        bad.flags = BV_FLAG_BAD_STRAND;
        uint8_t st2[10];
        memcpy(st2, strand, 10);
        st2[1] = BV_STRAND_NONE;
        SiteMeta m2 = m;
        m2.odd_strands.emplace_back(1u, '*');
        SiteCells c2{base, qual, st2, 10};
        CHECK_EQ(thrown([&] { out_cvg_line(m2, c2, bad); }), "[ERROR] Get strange strand symbol: *");
    }

    // ---- VCF row (src/basetype_caller.cpp:1103-1209); numbers of the first record of the C1 fixture ---------------------------
    rec.n_alt = 1; rec.alt[0] = 0; rec.n_active = 2;
    rec.af[0] = 0.62484312; rec.qual = 106.66461659; rec.fs_vcf = 0.0;
    bv_call_out call{0, 0, 3, 5};
    bv_group_out groups[2];
    memset(groups, 0, sizeof(groups));
    groups[0].n_alt = 1; groups[0].af[0] = 1.0;
    const std::vector<std::string> gnames = {"BJ", "GD"};
    const std::string want =
        "chr11\t5247141\t.\tG\tA\t106.664617\t.\tCM_DP=8;CM_AC=5;CM_AF=0.624843;CM_CAF=0.625;MQRankSum=0;ReadPosRankSum=3;"
        "BaseQRankSum=5;QD=21.332923;SOR=1.333333;FS=0.000000;SB_REF=2,1;SB_ALT=3,2;BJ_AF=1\tGT:AB:SO:BP"
        "\t./1:A:+:0.992057\t./1:A:-:0.999749\t./1:A:+:0.999800\t./1:A:-:0.999749\t./1:A:+:0.999842"
        "\t0/.:G:+:0.999749\t0/.:G:-:0.999602\t0/.:G:+:0.999000\t./.\t./.\n";
    CHECK_EQ(out_vcf_line(m, c, rec, call, gnames, groups), want);
    {
        bv_site_out low = rec;
        low.qual = 20.0;   // FILTER is "." only above QUAL_THRESHOLD = 20 (cpp:1199)
        CHECK(out_vcf_line(m, c, low, call, gnames, groups).find("\t20.000000\tLowQual\t") != std::string::npos);
        bv_site_out nan = rec;
        nan.af[0] = std::numeric_limits<double>::quiet_NaN();
        CHECK(out_vcf_line(m, c, nan, call, {}, nullptr).find(";CM_AF=-nan;") != std::string::npos);
    }

    // ---- headers ----------------------------------------------------------------------------------------------------------------
    CHECK(cvg_header_define().find("#CHROM\tPOS\tREF\tDepth\tA\tC\tG\tT\tIndels\tFS\tSOR\tStrand_Coverage(REF_FWD,REF_REV,ALT_FWD,ALT_REV)") !=
          std::string::npos);
    const std::string vh = vcf_header_define({"##contig=<ID=chr1,length=10,assembly=x.fa>"}, "##reference=file:///x.fa",
                                             {"##INFO=<ID=BJ_AF,Number=A,Type=Float,Description=\"x\">"}, {"s1", "s2"});
    CHECK(vh.compare(0, 20, "##fileformat=VCFv4.2") == 0);
    CHECK(vh.find("\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\ts2") != std::string::npos);
    CHECK(vh.find("BJ_AF") < vh.find("##contig") && vh.find("##contig") < vh.find("##reference"));

    // ---- no CPU fallback: the driver needs a CUDA device ------------------------------------------------------------------------
    if (!getenv("BV_EXPECT_GPU")) {
        const std::string e = thrown([] {
            BasevarCaller bc(4, {}, 0.01f, nullptr, nullptr);
        });
        CHECK(e.find("bv_create") != std::string::npos && e.find("no CPU fallback") != std::string::npos);
    }
    if (g_fail) { printf("%d FAILED\n", g_fail); return 1; }
    printf("ALL OK\n");
    return 0;
}
