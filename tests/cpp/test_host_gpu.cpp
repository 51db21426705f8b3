// GPU tests of the C++ host layer: the BaseType / strand_bias mirror against the oracle on the same BatchInfo, the
// one-site drop-in constructor, and region sharding (several shards on the visible devices) against one big tile.
// TEST INFRASTRUCTURE: links oracle/libbvoracle.so (CPU restatement) and, when present, oracle/_ref/libbvref.so (the
// compiled reference) through dlopen; the host library itself never touches either.
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>

#include "../../basevar_b200/host/bv_host.hpp"

using namespace bvhost;

static int g_fail = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
    } while (0)

typedef int (*site_fn)(const uint8_t*, const uint8_t*, const uint8_t*, uint32_t, uint8_t, const bv_params*, bv_site_out*);
typedef int (*refsite_fn)(const uint8_t*, const uint8_t*, const uint8_t*, uint32_t, uint8_t, float, bv_site_out*);

static bool close_rel(double a, double b, double rtol = 1e-9, double atol = 1e-9) {
    if (std::isnan(a) && std::isnan(b)) return true;
    return a == b || std::fabs(a - b) <= atol + rtol * std::fabs(b);
}

static BatchInfo random_bi(std::mt19937_64& rng, size_t n, double cov, int n_alt_alleles) {
    std::uniform_real_distribution<double> U(0, 1);
    BatchInfo bi;
    bi.n = n; bi.ref_id = "chrT"; bi.ref_pos = (uint32_t)(rng() % 100000000);
    const char acgt[4] = {'A', 'C', 'G', 'T'};
    const int ref = (int)(rng() % 4);
    bi.ref_base.assign(1, acgt[ref]);
    double p[4] = {0.002, 0.002, 0.002, 0.002};
    p[ref] = 1.0;
    for (int k = 0; k < n_alt_alleles; ++k) p[(ref + 1 + k) % 4] = std::pow(10.0, -3 * U(rng));
    const double ps = p[0] + p[1] + p[2] + p[3];
    bi.depth = 0;
    for (size_t i = 0; i < n; ++i) {
        if (U(rng) >= cov) {
            bi.align_bases.push_back("N"); bi.align_base_quals.push_back('!'); bi.map_strands.push_back('.');
            bi.mapqs.push_back(0); bi.base_pos_ranks.push_back(0);
            continue;
        }
        double u = U(rng) * ps;
        int b = 0;
        while (b < 3 && u >= p[b]) { u -= p[b]; ++b; }
        const double v = U(rng);
        bi.align_bases.push_back(v < 0.01 ? "+AC" : v < 0.02 ? "-T" : std::string(1, acgt[b]));
        bi.align_base_quals.push_back((char)(33 + 2 + rng() % 40));
        bi.map_strands.push_back((rng() & 1) ? '+' : '-');
        bi.mapqs.push_back(60); bi.base_pos_ranks.push_back(1 + (int)(rng() % 35));
        ++bi.depth;
    }
    return bi;
}

int main(int argc, char** argv) {
    const std::string root = argc > 1 ? argv[1] : ".";
    void* ho = dlopen((root + "/oracle/libbvoracle.so").c_str(), RTLD_NOW);
    if (!ho) { printf("FAIL cannot load oracle: %s\n", dlerror()); return 1; }
    site_fn bvo_site = (site_fn)dlsym(ho, "bvo_site");
    void* hr = dlopen((root + "/oracle/_ref/libbvref.so").c_str(), RTLD_NOW);
    refsite_fn bvref_site = hr ? (refsite_fn)dlsym(hr, "bvref_site") : nullptr;
    printf("compiled reference %s\n", bvref_site ? "present" : "absent (oracle only)");

    std::mt19937_64 rng(20241017);
    const size_t N = 600;
    const float maf = cli_min_af(0.01f, N);
    bv_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.min_af = maf; prm.lrt_threshold = 24; prm.em_max_iter = 100; prm.em_eps = 0.001f;

    // ---- batch path: BatchInfo -> TilePacker -> Context -> BaseType(bi, record) vs oracle / reference ----
    const int S = 400;
    std::vector<BatchInfo> bis;
    TilePacker pk((uint32_t)N, S);
    for (int s = 0; s < S; ++s) {
        bis.push_back(random_bi(rng, N, s % 3 == 0 ? 0.9 : 0.12, s % 4));
        pk.add_site(bis.back());
    }
    Context ctx(0, maf, (uint32_t)N, S, 2);
    std::vector<bv_site_out> recs = ctx.run(pk.tile());
    CHECK(ctx.launch_count() == 7);   // K1, K2, K3, K4a, K4b (two builds of it are launched, one of them works), Fisher tests
    int n_var = 0;
    bv_tile t = pk.tile();
    for (int s = 0; s < S; ++s) {
        bv_site_out want;
        bvo_site(t.base + s * t.pitch, t.qual + s * t.pitch, t.strand + s * t.pitch, (uint32_t)N, t.ref_base[s], &prm, &want);
        BaseType bt(&bis[s], recs[s]);
        bt.lrt();
        const bool soft = ((recs[s].flags | want.flags) & (BV_FLAG_NEAR_LRT | BV_FLAG_LRT_TIE)) != 0;
        CHECK(bt.get_total_depth() == (int)(want.depth[0] + want.depth[1] + want.depth[2] + want.depth[3] + want.depth_other));
        for (int b = 0; b < 4; ++b) CHECK(bt.get_base_depth(BASES[b]) == (double)want.depth[b]);
        if (!soft) {
            CHECK(bt.get_alt_bases().size() == want.n_alt);
            for (size_t k = 0; k < bt.get_alt_bases().size() && k < want.n_alt; ++k) {
                CHECK(bt.get_alt_bases()[k] == BASES[want.alt[k]]);
                CHECK(close_rel(bt.get_lrt_af(bt.get_alt_bases()[k]), want.af[k]));
            }
            if (want.n_alt) { CHECK(close_rel(bt.get_var_qual(), want.qual)); ++n_var; }
        }
        // strand_bias, CVG row (all non-ref bases) and VCF row (called ALTs)
        const char up = (char)toupper(bis[s].ref_base[0]);
        std::string others;
        for (char b : BASES) if (b != up) others.push_back(b);
        StrandBiasInfo sb = strand_bias(up, others, recs[s]);
        CHECK(close_rel(sb.fs, want.fs_cvg));
        if (!soft && want.n_alt) {
            std::string alts(bt.get_alt_bases().begin(), bt.get_alt_bases().end());
            StrandBiasInfo sv = strand_bias(up, alts, recs[s]);
            CHECK(close_rel(sv.fs, want.fs_vcf));
            int af = 0, ar = 0;
            for (int k = 0; k < want.n_alt; ++k) { af += want.fwd[want.alt[k]]; ar += want.rev[want.alt[k]]; }
            CHECK(sv.alt_fwd == af && sv.alt_rev == ar);
        }
        if (bvref_site && !soft) {   // the compiled, unmodified reference on the same cells
            bv_site_out ref;
            bvref_site(t.base + s * t.pitch, t.qual + s * t.pitch, t.strand + s * t.pitch, (uint32_t)N, t.ref_base[s], maf, &ref);
            CHECK(bt.get_alt_bases().size() == ref.n_alt);
            for (size_t k = 0; k < bt.get_alt_bases().size() && k < ref.n_alt; ++k) CHECK(close_rel(bt.get_lrt_af(bt.get_alt_bases()[k]), ref.af[k]));
            if (ref.n_alt) CHECK(close_rel(bt.get_var_qual(), ref.qual));
            CHECK(close_rel(sb.fs, ref.fs_cvg));
        }
    }
    CHECK(n_var > 50);
    // strand_bias over an ALT set the record carries no FS for: the 2x2 table goes to the device (bv_fisher_fs).  Known answers of
    // kt_fisher_exact: (4,3,3,0) -> p 0.4749999999999977, (500,480,20,3) -> p 0.00050937499055459272.
    {
        bv_site_out r;
        memset(&r, 0, sizeof(r));
        r.depth[0] = 7; r.depth[1] = 3; r.depth[2] = 4; r.fwd[0] = 4; r.rev[0] = 3; r.fwd[1] = 3; r.rev[1] = 0; r.fwd[2] = 2; r.rev[2] = 2;
        r.fs_cvg = -1.0; r.fs_vcf = -1.0;   // (not the answer for {C}: must not be picked)
        StrandBiasInfo sc = strand_bias('A', "C", r, &ctx);
        CHECK(sc.ref_fwd == 4 && sc.ref_rev == 3 && sc.alt_fwd == 3 && sc.alt_rev == 0);
        CHECK(close_rel(sc.fs, 3.233063903751355));
        StrandBiasInfo sd = strand_bias('A', "C", r);   // the calling thread's default context
        CHECK(sd.fs == sc.fs);
        r.fwd[0] = 500; r.rev[0] = 480; r.fwd[3] = 20; r.rev[3] = 3;
        StrandBiasInfo st = strand_bias('A', "T", r);   // 1,003 reads: the default context grows
        CHECK(close_rel(st.fs, 32.92962381969127));
        printf("strand_bias over a free ALT set: FS %.6f, %.6f (device)\n", sc.fs, st.fs);
    }
    {   // the same sites through the sparse transport: byte-identical records
        SparsePacker sp((uint32_t)N, S);
        for (int s = 0; s < S; ++s) sp.add_site(bis[s]);
        CHECK(sp.n_sites() == (uint32_t)S && sp.n_cells() > 0 && sp.n_cells() < (size_t)S * N);
        const uint64_t l0 = ctx.launch_count();
        std::vector<bv_site_out> srecs = ctx.run(sp.tile());
        CHECK(ctx.launch_count() - l0 == 8);   // K0 expand + K1..K3, K4a, K4b x 2, Fisher tests
        CHECK(memcmp(srecs.data(), recs.data(), sizeof(bv_site_out) * S) == 0);
        printf("sparse transport: %zu cells for %d x %zu sample-sites, records identical\n", sp.n_cells(), S, N);
    }
    printf("batch path: %d sites, %d variant\n", S, n_var);

    // ---- drop-in constructor: BaseType bt(&bi, min_af); bt.lrt();  (src/basetype_caller.cpp:742-743) ----
    for (int s = 0; s < 25; ++s) {
        BaseType bt(&bis[s], (double)maf);
        bt.lrt();
        CHECK(memcmp(&bt.record(), &recs[s], sizeof(bv_site_out)) == 0);
        BaseType from_rec(&bis[s], recs[s]);
        CHECK(bt.get_alt_bases() == from_rec.get_alt_bases() && bt.get_total_depth() == from_rec.get_total_depth());
    }
    {
        BatchInfo bad = bis[0];
        bad.align_bases[3] = "AC";
        bool threw = false;
        try { BaseType bt(&bad, (double)maf); } catch (const std::runtime_error& e) { threw = std::string(e.what()).find("Check: AC") != std::string::npos; }
        CHECK(threw);
    }

    // ---- region sharding: 5 shards over the visible devices == one pass, in coordinate order ----
    {
        const uint64_t beg = 1000, len = 430000;   // five 100-kb tasks, the last one short
        const uint32_t Ns = 64;
        auto cell_site = [&](uint64_t site, uint8_t* b, uint8_t* q, uint8_t* st) {
            std::mt19937_64 r(site * 7919 + 13);
            for (uint32_t i = 0; i < Ns; ++i) {
                const bool cov = (r() % 10) < 3;
                b[i] = cov ? (uint8_t)((r() % 50) ? site % 4 : r() % 4) : BV_BASE_N;
                q[i] = cov ? (uint8_t)(10 + r() % 30) : 0;
                st[i] = cov ? (uint8_t)(r() & 1) : BV_STRAND_NONE;
            }
        };
        TileSource fill = [&](uint64_t s0, uint32_t n, TilePacker& into) {
            std::vector<uint8_t> b(Ns), q(Ns), st(Ns);
            for (uint32_t k = 0; k < n; ++k) {
                cell_site(s0 + k, b.data(), q.data(), st.data());
                into.add_site_cells("ACGT"[(s0 + k) % 4], b.data(), q.data(), st.data(), nullptr);
            }
        };
        auto run = [&](int n_shards, std::vector<bv_site_out>& out) {
            uint64_t next = beg;
            bool ordered = true;
            RunOptions o;
            o.min_af = 0.01f; o.n_samples = Ns; o.tile_sites = 30000; o.n_slots = 2; o.n_shards = n_shards;
            out.clear();
            run_region(beg, beg + len, o, fill, [&](uint64_t s0, const bv_site_out* r, uint32_t n) {
                ordered = ordered && s0 == next;
                next = s0 + n;
                out.insert(out.end(), r, r + n);
            });
            CHECK(ordered && next == beg + len && out.size() == len);
        };
        std::vector<bv_site_out> one, five;
        run(1, one);
        run(5, five);
        CHECK(one.size() == five.size() && memcmp(one.data(), five.data(), one.size() * sizeof(bv_site_out)) == 0);
        printf("region sharding: %zu sites, 5 shards == 1 shard\n", one.size());
    }
    printf(g_fail ? "FAILED %d checks\n" : "ALL OK\n", g_fail);
    return g_fail ? 1 : 0;
}
