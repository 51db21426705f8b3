// bv_host.cpp -- see bv_host.hpp.  Host-side plumbing only: packing, contexts, record -> BaseType, region sharding.
#include "bv_host.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <thread>

namespace bvhost {

static void check(int rc, bv_ctx* ctx, const char* what) {
    if (rc != BV_OK) throw std::runtime_error(std::string("[ERROR] ") + what + ": " + bv_last_error(ctx));
}

uint8_t encode_base(const std::string& s) {
    const char fb = s.empty() ? 'N' : s[0];
    if (fb == 'N') return BV_BASE_N;
    if (fb == '+') return BV_BASE_INS;
    if (fb == '-') return BV_BASE_DEL;
    if (s.size() != 1)   // src/basetype.cpp:54-56
        throw std::runtime_error("[ERROR] Why dose the size of aligned base is not 1? Check: " + s);
    switch (fb) {
        case 'A': return BV_BASE_A;
        case 'C': return BV_BASE_C;
        case 'G': return BV_BASE_G;
        case 'T': return BV_BASE_T;
        default: return BV_BASE_OTHER;   // counted in the total depth, never an allele (src/basetype.cpp:58-64)
    }
}

float cli_min_af(float min_af, size_t n_bam) { return std::min(float(100) / n_bam, min_af); }

// ---- TilePacker ------------------------------------------------------------------------------------------------------
static uint8_t* plane_alloc(size_t bytes, bool pinned) {
    void* p = nullptr;
    if (pinned) {
        if (bv_host_alloc(&p, bytes) != BV_OK) throw std::runtime_error(std::string("[ERROR] bv_host_alloc: ") + bv_last_error(nullptr));
    } else {
        if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) throw std::bad_alloc();
    }
    return static_cast<uint8_t*>(p);
}
static void plane_free(uint8_t* p, bool pinned) {
    if (!p) return;
    if (pinned) bv_host_free(p);
    else free(p);
}

TilePacker::TilePacker(uint32_t n_samples, uint32_t max_sites, bool pinned)
    : n_samples_(n_samples), max_sites_(max_sites), pitch_(((uint64_t)n_samples + 15) / 16 * 16), pinned_(pinned) {
    const size_t plane = (size_t)max_sites_ * pitch_;
    base_ = plane_alloc(plane, pinned_);
    qual_ = plane_alloc(plane, pinned_);
    strand_ = plane_alloc(plane, pinned_);
    mapq_ = plane_alloc(plane, pinned_);
    ref_ = plane_alloc(max_sites_, pinned_);
}

TilePacker::~TilePacker() {
    plane_free(base_, pinned_); plane_free(qual_, pinned_); plane_free(strand_, pinned_); plane_free(mapq_, pinned_);
    plane_free(ref_, pinned_);
}

void TilePacker::next_row(char ref_base, uint8_t** b, uint8_t** q, uint8_t** s, uint8_t** m) {
    if (n_sites_ >= max_sites_) throw std::runtime_error("[ERROR] TilePacker is full");
    const size_t row = (size_t)n_sites_ * pitch_;
    // uncovered = N ! 0 0 . (src/basetype_caller.cpp:1051-1077)
    memset(base_ + row, BV_BASE_N, pitch_);
    memset(qual_ + row, 0, pitch_);
    memset(strand_ + row, BV_STRAND_NONE, pitch_);
    memset(mapq_ + row, 0, pitch_);
    ref_[n_sites_] = (uint8_t)ref_base;
    *b = base_ + row; *q = qual_ + row; *s = strand_ + row; *m = mapq_ + row;
    ++n_sites_;
}

void TilePacker::add_site_cells(char ref_base, const uint8_t* base, const uint8_t* qual, const uint8_t* strand, const uint8_t* mapq) {
    uint8_t *b, *q, *s, *m;
    next_row(ref_base, &b, &q, &s, &m);
    memcpy(b, base, n_samples_); memcpy(q, qual, n_samples_); memcpy(s, strand, n_samples_);
    if (mapq) memcpy(m, mapq, n_samples_);
}

bv_tile TilePacker::tile() const {
    bv_tile t;
    t.base = base_; t.qual = qual_; t.strand = strand_; t.ref_base = ref_;
    t.pitch = pitch_; t.n_sites = n_sites_; t.n_samples = n_samples_;
    t.location = BV_LOC_HOST; t.out_mode = BV_OUT_RECORDS;
    return t;
}

// ---- SparsePacker -----------------------------------------------------------------------------------------------------
SparsePacker::SparsePacker(uint32_t n_samples, uint32_t max_sites, size_t reserve_cells) : n_samples_(n_samples), max_sites_(max_sites) {
    if (n_samples > BV_CELL_MAX_SAMPLES) throw std::runtime_error("[ERROR] sparse tiles hold at most 2^20 samples");
    site_start_ = reinterpret_cast<uint32_t*>(plane_alloc(((size_t)max_sites + 1) * sizeof(uint32_t), true));
    ref_ = plane_alloc(max_sites ? max_sites : 1, true);
    site_start_[0] = 0;
    grow(reserve_cells ? reserve_cells : (size_t)max_sites * 16 + 1024);
}

SparsePacker::~SparsePacker() {
    plane_free(reinterpret_cast<uint8_t*>(cells_), true); plane_free(reinterpret_cast<uint8_t*>(aux_), true);
    plane_free(reinterpret_cast<uint8_t*>(site_start_), true); plane_free(ref_, true);
}

void SparsePacker::grow(size_t want) {
    if (want <= cap_cells_) return;
    size_t cap = cap_cells_ ? cap_cells_ : 1024;
    while (cap < want) cap *= 2;
    uint32_t* c = reinterpret_cast<uint32_t*>(plane_alloc(cap * sizeof(uint32_t), true));
    uint32_t* a = reinterpret_cast<uint32_t*>(plane_alloc(cap * sizeof(uint32_t), true));
    if (n_cells_) { memcpy(c, cells_, n_cells_ * sizeof(uint32_t)); memcpy(a, aux_, n_cells_ * sizeof(uint32_t)); }
    plane_free(reinterpret_cast<uint8_t*>(cells_), true); plane_free(reinterpret_cast<uint8_t*>(aux_), true);
    cells_ = c; aux_ = a; cap_cells_ = cap;
}

void SparsePacker::begin_site(char ref_base) {
    if (n_sites_ >= max_sites_) throw std::runtime_error("[ERROR] SparsePacker is full");
    ref_[n_sites_] = (uint8_t)ref_base;
    ++n_sites_;
    site_start_[n_sites_] = (uint32_t)n_cells_;
}

void SparsePacker::add_cell(uint32_t sample, uint8_t base, uint8_t strand, uint8_t phred, uint8_t mapq, uint16_t rpr) {
    if (n_sites_ == 0) throw std::runtime_error("[ERROR] SparsePacker::add_cell before begin_site");
    if (sample >= n_samples_) throw std::runtime_error("[ERROR] SparsePacker: sample index out of range");
    if (n_cells_ >= 0xffffffffull) throw std::runtime_error("[ERROR] SparsePacker: more than 2^32 cells in one tile");
    grow(n_cells_ + 1);
    cells_[n_cells_] = BV_CELL_PACK(sample, base & 7u, strand & 3u, phred & 127u);
    aux_[n_cells_] = BV_CELL_AUX_PACK(mapq, rpr);
    ++n_cells_;
    site_start_[n_sites_] = (uint32_t)n_cells_;
}

bv_sparse_tile SparsePacker::tile() const {
    bv_sparse_tile t;
    t.cells = cells_; t.cells_aux = aux_; t.site_start = site_start_; t.ref_base = ref_; t.out = nullptr;
    t.n_sites = n_sites_; t.n_samples = n_samples_;
    t.format = BV_CELLS_U32; t.out_mode = BV_OUT_RECORDS;
    return t;
}

// ---- Context ----------------------------------------------------------------------------------------------------------
Context::Context(int device, float min_af, uint32_t max_samples, uint32_t max_sites, uint32_t n_slots, int em_abs_mode)
    : n_slots_(n_slots) {
    bv_params p;
    memset(&p, 0, sizeof(p));
    p.min_af = min_af;
    p.lrt_threshold = LRT_THRESHOLD;
    p.em_max_iter = 100;     // src/algorithm.h:213
    p.em_eps = 0.001f;       // src/algorithm.h:213
    p.em_abs_mode = em_abs_mode;
    p.max_samples = max_samples;
    p.max_sites = max_sites;
    p.n_slots = n_slots;
    check(bv_create(device, &p, &ctx_), nullptr, "bv_create");
    max_samples_ = max_samples;
}
Context::~Context() { bv_destroy(ctx_); }
void Context::submit(int slot, const bv_tile& tile) { check(bv_tile_submit(ctx_, slot, &tile), ctx_, "bv_tile_submit"); }
void Context::wait(int slot, bv_site_out* out) { check(bv_tile_wait(ctx_, slot, out), ctx_, "bv_tile_wait"); }
std::vector<bv_site_out> Context::run(const bv_tile& tile) {
    std::vector<bv_site_out> out(tile.n_sites);
    submit(0, tile);
    wait(0, out.data());
    return out;
}
void Context::submit(int slot, const bv_sparse_tile& tile) { check(bv_tile_submit_sparse(ctx_, slot, &tile), ctx_, "bv_tile_submit_sparse"); }
std::vector<bv_site_out> Context::run(const bv_sparse_tile& tile) {
    std::vector<bv_site_out> out(tile.n_sites);
    submit(0, tile);
    wait(0, out.data());
    return out;
}
uint64_t Context::launch_count() const { return bv_launch_count(ctx_); }
double Context::fisher_fs(int ref_fwd, int ref_rev, int alt_fwd, int alt_rev) {
    const int32_t t[4] = {ref_fwd, ref_rev, alt_fwd, alt_rev};
    double fs = 0.0;
    check(bv_fisher_fs(ctx_, t, 1, &fs), ctx_, "bv_fisher_fs");
    return fs;
}

// ---- BaseType ---------------------------------------------------------------------------------------------------------
BaseType::BaseType(const BatchInfo* bi, const bv_site_out& rec)
    : ref_id_(bi->ref_id), ref_base_(bi->ref_base), ref_pos_(bi->ref_pos), rec_(rec) {
    fill_from_record();
    lrt_done_ = true;
}

namespace {
struct OneSite {   // the calling thread's default context for the one-site drop-in constructor
    std::unique_ptr<Context> ctx;
    std::unique_ptr<TilePacker> packer;
    float min_af = -1.f;
    uint32_t n = 0;
};
thread_local OneSite t_one;
}  // namespace

BaseType::BaseType(const BatchInfo* bi, double af)
    : ref_id_(bi->ref_id), ref_base_(bi->ref_base), ref_pos_(bi->ref_pos) {
    const float maf = (float)af;
    const uint32_t n = (uint32_t)bi->n;
    if (!t_one.ctx || t_one.min_af != maf || t_one.n < n) {
        const char* d = getenv("BASEVAR_B200_DEVICE");
        t_one.ctx.reset();
        t_one.packer.reset();
        const uint32_t cap = std::max<uint32_t>(n, 1);
        t_one.ctx.reset(new Context(d ? atoi(d) : 0, maf, cap, 1, 1));
        t_one.min_af = maf;
        t_one.n = cap;
    }
    if (!t_one.packer || t_one.packer->n_samples() != n) t_one.packer.reset(new TilePacker(n, 1));
    t_one.packer->clear();
    t_one.packer->add_site(*bi);   // throws on a malformed base string, like the reference's constructor
    rec_ = t_one.ctx->run(t_one.packer->tile())[0];
    // the constructor gives depths; ALT / AF / QUAL appear with lrt() (src/basetype.cpp:130)
    total_depth_ = (int)(rec_.depth[0] + rec_.depth[1] + rec_.depth[2] + rec_.depth[3] + rec_.depth_other);
}

void BaseType::lrt(const std::vector<char>& specific_bases) {
    if (specific_bases != BASES)
        throw std::invalid_argument("[ERROR] lrt() over a subset of bases is not computed on the device (pop-group path)");
    if (rec_.flags & BV_FLAG_ZERO_SUBSET)   // src/basetype.cpp:113-115
        throw std::runtime_error("[ERROR] The sum of frequence of active bases must always > 0. Check: ");
    if (lrt_done_) return;   // (the reference appends the ALT alleles again on a second call; nobody relies on that)
    fill_from_record();
    lrt_done_ = true;
}

void BaseType::fill_from_record() {
    total_depth_ = (int)(rec_.depth[0] + rec_.depth[1] + rec_.depth[2] + rec_.depth[3] + rec_.depth_other);
    alt_bases_.clear();
    af_by_lrt_.clear();
    for (int k = 0; k < rec_.n_alt && k < 4; ++k) {
        const char b = BASES[rec_.alt[k] & 3];
        alt_bases_.push_back(b);
        af_by_lrt_[b] = rec_.af[k];
    }
    var_qual_ = rec_.qual;
}

double BaseType::get_base_depth(char b) const {
    for (size_t i = 0; i < BASES.size(); ++i)
        if (BASES[i] == b) return (double)rec_.depth[i];
    throw std::runtime_error(std::string("[ERROR] out_of_range:: map::at '") + b + "' not found.");
}

double BaseType::get_lrt_af(char b) const {
    std::map<char, double>::const_iterator it = af_by_lrt_.find(b);
    if (it == af_by_lrt_.end()) throw std::runtime_error(std::string("[ERROR] out_of_range:: map::at '") + b + "' not found.");
    return it->second;
}

// ---- strand_bias -------------------------------------------------------------------------------------------------------
static int code_of(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }

StrandBiasInfo strand_bias(const char ref_base, const std::string alt_bases_string, const bv_site_out& rec, Context* ctx) {
    if (rec.flags & BV_FLAG_BAD_STRAND)   // src/basetype.cpp:271-273
        throw std::runtime_error("[ERROR] Get strange strand symbol: ");
    StrandBiasInfo s;
    const int rc = code_of(ref_base);
    s.ref_fwd = rc >= 0 ? (int)rec.fwd[rc] : 0;
    s.ref_rev = rc >= 0 ? (int)rec.rev[rc] : 0;
    s.alt_fwd = s.alt_rev = 0;
    unsigned alt_set = 0;
    for (char c : alt_bases_string) {
        const int k = code_of(c);
        if (k < 0 || k == rc || (alt_set >> k & 1)) continue;   // a base equal to REF counts as REF (basetype.cpp:256-261)
        alt_set |= 1u << k;
        s.alt_fwd += (int)rec.fwd[k];
        s.alt_rev += (int)rec.rev[k];
    }
    unsigned cvg_set = 0xfu & ~(rc >= 0 ? 1u << rc : 0u), vcf_set = 0;
    for (int k = 0; k < rec.n_alt && k < 4; ++k) vcf_set |= 1u << (rec.alt[k] & 3);
    // bases without reads do not change the table
    unsigned covered = 0;
    for (int k = 0; k < 4; ++k) if (rec.fwd[k] + rec.rev[k]) covered |= 1u << k;
    if (((alt_set ^ cvg_set) & covered) == 0) s.fs = rec.fs_cvg;
    else if (rec.n_alt && ((alt_set ^ vcf_set) & covered) == 0) s.fs = rec.fs_vcf;
    else if ((s.alt_fwd | s.alt_rev) == 0 || (s.ref_fwd | s.ref_rev) == 0) s.fs = 0.0;   // one possible table: p == 1
    else {
        // neither of the two sets the record carries FS for: the 2x2 table goes to the device (bv_fisher_fs, the code of bv_fisher_kernel)
        const uint32_t reads = (uint32_t)(s.ref_fwd + s.ref_rev + s.alt_fwd + s.alt_rev);
        if (!ctx) {
            if (!t_one.ctx || t_one.ctx->max_samples() + 1 < reads) {
                const char* d = getenv("BASEVAR_B200_DEVICE");
                t_one.ctx.reset();
                t_one.packer.reset();
                t_one.n = std::max<uint32_t>(reads, 1);
                t_one.min_af = 0.01f;
                t_one.ctx.reset(new Context(d ? atoi(d) : 0, t_one.min_af, t_one.n, 1, 1));
            }
            ctx = t_one.ctx.get();
        }
        s.fs = ctx->fisher_fs(s.ref_fwd, s.ref_rev, s.alt_fwd, s.alt_rev);
    }
    // src/basetype.cpp:286, int32 products as in the reference
    s.sor = (s.ref_rev * s.alt_fwd > 0) ? (double)(s.ref_fwd * s.alt_rev) / (double)(s.ref_rev * s.alt_fwd) : 10000;
    return s;
}

// ---- region sharding ----------------------------------------------------------------------------------------------------
std::vector<Shard> shard_region(uint64_t reg_beg, uint64_t reg_end, int n_gpus, uint64_t step) {
    std::vector<Shard> out;
    if (reg_end <= reg_beg || n_gpus <= 0) return out;
    if (step == 0) step = 1;
    const uint64_t n_steps = (reg_end - reg_beg + step - 1) / step;   // 100-kb tasks, the last one may be short
    const uint64_t g = std::min<uint64_t>((uint64_t)n_gpus, n_steps);
    uint64_t s0 = 0;
    for (uint64_t i = 0; i < g; ++i) {
        const uint64_t cnt = n_steps / g + (i < n_steps % g ? 1 : 0);
        Shard sh;
        sh.gpu = (int)i;
        sh.beg = reg_beg + s0 * step;
        sh.end = std::min(reg_end, reg_beg + (s0 + cnt) * step);
        out.push_back(sh);
        s0 += cnt;
    }
    return out;
}

void run_region(uint64_t reg_beg, uint64_t reg_end, const RunOptions& opt, const TileSource& fill, const RecordSink& sink) {
    std::vector<int> devices = opt.devices;
    if (devices.empty()) {
        // probe: contexts on devices 0,1,... until creation fails
        for (int d = 0; d < 64; ++d) {
            bv_params p;
            memset(&p, 0, sizeof(p));
            p.min_af = opt.min_af; p.lrt_threshold = LRT_THRESHOLD; p.em_max_iter = 100; p.em_eps = 0.001f; p.max_samples = 1;
            bv_ctx* c = nullptr;
            if (bv_create(d, &p, &c) != BV_OK) break;
            bv_destroy(c);
            devices.push_back(d);
        }
        if (devices.empty()) throw std::runtime_error(std::string("[ERROR] no CUDA device: ") + bv_last_error(nullptr));
    }
    const int n_shards = opt.n_shards > 0 ? opt.n_shards : (int)devices.size();
    const std::vector<Shard> shards = shard_region(reg_beg, reg_end, n_shards);
    std::vector<std::vector<bv_site_out>> results(shards.size());
    std::vector<std::exception_ptr> errors(shards.size());
    std::vector<std::thread> workers;
    for (size_t si = 0; si < shards.size(); ++si) {
        workers.emplace_back([&, si]() {
            try {
                const Shard& sh = shards[si];
                const uint32_t n_slots = std::max<uint32_t>(opt.n_slots, 1);
                Context ctx(devices[si % devices.size()], opt.min_af, opt.n_samples, opt.tile_sites, n_slots, opt.em_abs_mode);
                std::vector<std::unique_ptr<TilePacker>> packers;
                for (uint32_t k = 0; k < n_slots; ++k) packers.emplace_back(new TilePacker(opt.n_samples, opt.tile_sites));
                std::vector<bv_site_out>& res = results[si];
                res.resize(sh.end - sh.beg);
                struct Pending { uint64_t site0; uint32_t n; };
                std::vector<Pending> pending(n_slots, Pending{0, 0});
                uint32_t slot = 0;
                for (uint64_t s0 = sh.beg; s0 < sh.end; s0 += opt.tile_sites) {
                    const uint32_t n = (uint32_t)std::min<uint64_t>(opt.tile_sites, sh.end - s0);
                    if (pending[slot].n) {   // the slot's previous tile: collect before its packer is reused
                        ctx.wait((int)slot, res.data() + (pending[slot].site0 - sh.beg));
                        pending[slot].n = 0;
                    }
                    TilePacker& pk = *packers[slot];
                    pk.clear();
                    fill(s0, n, pk);
                    if (pk.n_sites() != n) throw std::runtime_error("[ERROR] TileSource packed a wrong number of sites");
                    ctx.submit((int)slot, pk.tile());
                    pending[slot] = Pending{s0, n};
                    slot = (slot + 1) % n_slots;
                }
                for (uint32_t k = 0; k < n_slots; ++k) {
                    const uint32_t sl = (slot + k) % n_slots;
                    if (pending[sl].n) ctx.wait((int)sl, res.data() + (pending[sl].site0 - sh.beg));
                }
            } catch (...) {
                errors[si] = std::current_exception();
            }
        });
    }
    for (auto& w : workers) w.join();
    for (auto& e : errors)
        if (e) std::rethrow_exception(e);   // like future.get() in the reference (basetype_caller.cpp:513-517)
    // merge = concatenation in coordinate order (the reference's merge_file_by_line order, basetype_utils.cpp:90-123)
    for (size_t si = 0; si < shards.size(); ++si)
        for (uint64_t s0 = shards[si].beg; s0 < shards[si].end; s0 += opt.tile_sites) {
            const uint32_t n = (uint32_t)std::min<uint64_t>(opt.tile_sites, shards[si].end - s0);
            sink(s0, results[si].data() + (s0 - shards[si].beg), n);
        }
}

}  // namespace bvhost
