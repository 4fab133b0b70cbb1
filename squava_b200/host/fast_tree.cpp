// fast_tree.cpp -- the host side of leaf-parallel MCTS at scale (BASELINE configs[2]:
// "10^9 iterations, host tree + GPU leaf batches of 64K").
//
// Same algorithm as UCT() in squava_host.cpp (src/mcts/mcts.go:123-199: select by UCB1 while a node
// is fully expanded, expand one untried move picked uniformly, roll out, back-propagate), same
// batched-mode conventions (visits counted at selection time = virtual loss, wins when the batch
// returns, a node created inside a batch learns whether it is terminal from its own rollout).
// What differs is the data structure, built for 10^9 nodes and many selecting threads:
//   * nodes are 16-byte records in one lazily committed arena, addressed by 32-bit index (a node's move is its slot in
//     the parent's child block, its side the parity of its depth: neither is stored);
//   * the children of a node live in ONE contiguous block (slot = rank of the move among the
//     node's empty cells), allocated when the first child is made: a UCB1 scan is a linear read;
//   * counts are 32-bit atomics; a thread claims an untried move with one fetch_and, so any
//     number of threads select / back-propagate concurrently without locks;
//   * every batch is three phases -- parallel select, ONE sq_playouts call, parallel update.
// With one thread, batch 1 and the test hooks it grows exactly the reference's tree (children in
// creation order, first maximum wins ties): tests/test_host_uct_vs_reference.py.
#include <sys/mman.h>
#include <x86intrin.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>

#include "squava_host.hpp"

namespace squava {
namespace mcts {

namespace {

constexpr uint32_t kFull = 0x1FFFFFFu;
constexpr uint32_t kPending = 1u << 31;      // untried: the node's first rollout has not come back yet
constexpr int kRankShift = 26;               // untried bits 26..30: creation rank among the parent's children
constexpr uint32_t kRankMask = 31u << kRankShift;
const double kEps = std::numeric_limits<double>::denorm_min();

struct FNode {
    std::atomic<uint32_t> visits;        // includes selections still in flight
    std::atomic<uint32_t> wins;          // of the player who moved INTO this node (Update, mcts.go:268-271)
    std::atomic<uint32_t> first_child;   // arena index of the child block, 0 = none yet
    std::atomic<uint32_t> untried;       // moves nobody has claimed (25 bits) | creation rank (bits 26..30) | kPending
};
static_assert(sizeof(FNode) == 16, "node record");

struct Rng {                              // xoshiro256**, one per selecting thread
    uint64_t s[4];
    explicit Rng(uint64_t seed) {
        for (auto& w : s) { seed += 0x9e3779b97f4a7c15ull; uint64_t z = seed; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; w = z ^ (z >> 31); }
    }
    uint64_t next() {
        auto rotl = [](uint64_t x, int k) { return (x << k) | (x >> (64 - k)); };
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    int intn(int n) {                     // unbiased, like rand.Intn (Lemire's multiply-shift with rejection)
        uint64_t m = (uint64_t)(uint32_t)(next() >> 32) * (uint64_t)n;
        if ((uint32_t)m < (uint32_t)n) {
            const uint32_t t = (0u - (uint32_t)n) % (uint32_t)n;
            while ((uint32_t)m < t) m = (uint64_t)(uint32_t)(next() >> 32) * (uint64_t)n;
        }
        return (int)(m >> 32);
    }
};

inline int nth_set_bit(uint32_t m, int n) {
    for (int i = 0; i < n; i++) m &= m - 1;
    return __builtin_ctz(m);
}

// ln(n) for n >= 1 in single precision without libm: exponent + a cubic on the mantissa (|error| < 7e-4), for the
// throughput mode's exploration term (the exact mode keeps std::log in double)
inline float fast_logf(uint32_t n) {
    float x = (float)n;
    uint32_t b;
    std::memcpy(&b, &x, 4);
    const int e = (int)(b >> 23) - 127;
    b = (b & 0x7fffffu) | 0x3f800000u;
    float m;
    std::memcpy(&m, &b, 4);                                  // mantissa in [1, 2)
    const float l2 = ((0.44717955f * m - 1.4699568f) * m + 2.8212026f) * m - 1.7417939f;
    return ((float)e + l2) * 0.69314718f;
}

// run fn(t) for t in [0, T) on T threads; the pool lives as long as the search
class Pool {
public:
    explicit Pool(int n) : n_(n) {
        for (int t = 1; t < n_; t++) th_.emplace_back([this, t] { loop(t); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void run(const std::function<void(int)>& fn) {
        if (n_ == 1) { fn(0); return; }
        { std::lock_guard<std::mutex> lk(mu_); fn_ = &fn; left_ = n_ - 1; gen_++; }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return left_ == 0; });
    }
private:
    void loop(int t) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int)>* fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(t);
            { std::lock_guard<std::mutex> lk(mu_); if (--left_ == 0) done_.notify_one(); }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    uint64_t gen_ = 0;
    int left_ = 0;
    bool stop_ = false;
};

// UCB1 scan of one child block in single precision (throughput mode): w/v + K*sqrt(c*ln N / v) as
// w*r*r + K*sqrt(c*ln N)*r with r = 1/sqrt(v) (rsqrt + one Newton step, 8 children per instruction).  The exact double form
// (and the creation-order tie rule) is kept for the mode the tests pin against the reference.
int best_child_scalar(const FNode* blk, int slots, float k_sqrt_lg, const uint32_t* dv) {
    int best = -1; float bs = 0.f;
    for (int k = 0; k < slots; k++) {
        const uint32_t gv = blk[k].visits.load(std::memory_order_relaxed);
        const float v = gv ? (float)(gv + (dv ? dv[k] : 0u)) : 0.f;
        if (!(v > 0.f)) continue;                        // claimed, still being set up
        const float inv = 1.0f / v;
        const float sc = (float)blk[k].wins.load(std::memory_order_relaxed) * inv + k_sqrt_lg * std::sqrt(inv);
        if (sc > bs) { bs = sc; best = k; }
    }
    return best;
}

__attribute__((target("avx2,fma")))
int best_child_avx2(const FNode* blk, int slots, float k_sqrt_lg, const uint32_t* dv) {
    // Records are 4 dwords {visits, wins, first_child, untried}: two per 256-bit load.  Four loads and four in-lane
    // shuffles give the visits and the wins of 8 children, in slot order (0,2,4,6,1,3,5,7) -- no gathers.
    alignas(32) float sc[32];
    alignas(32) static const uint32_t zero[32] = {0};
    alignas(32) static const int kSlotOfLane[8] = {0, 2, 4, 6, 1, 3, 5, 7};
    if (!dv) dv = zero;
    const __m256 c = _mm256_set1_ps(k_sqrt_lg), half = _mm256_set1_ps(0.5f), three = _mm256_set1_ps(3.0f);
    const __m256i lane_slot = _mm256_load_si256((const __m256i*)kSlotOfLane);
    __m256 top = _mm256_set1_ps(-1.f);
    for (int k = 0; k < slots; k += 8) {
        // lanes past the block read the next records of the arena (mapped, possibly foreign): masked out below
        const float* p = reinterpret_cast<const float*>(blk + k);
        const __m256 l0 = _mm256_loadu_ps(p), l1 = _mm256_loadu_ps(p + 8), l2 = _mm256_loadu_ps(p + 16), l3 = _mm256_loadu_ps(p + 24);
        const __m256 a = _mm256_shuffle_ps(l0, l1, 0x44), b = _mm256_shuffle_ps(l2, l3, 0x44);     // v w v w | v w v w
        const __m256i gv = _mm256_castps_si256(_mm256_shuffle_ps(a, b, 0x88));                       // visits
        const __m256i gw = _mm256_castps_si256(_mm256_shuffle_ps(a, b, 0xDD));                       // wins
        const __m256i live = _mm256_cmpgt_epi32(_mm256_set1_epi32(slots - k), lane_slot);
        const __m256i some = _mm256_andnot_si256(_mm256_cmpeq_epi32(gv, _mm256_setzero_si256()), live);   // visits != 0: the child is set up
        const __m256i dvv = _mm256_permutevar8x32_epi32(_mm256_loadu_si256((const __m256i*)(dv + k)), lane_slot);
        const __m256 vv = _mm256_cvtepi32_ps(_mm256_add_epi32(gv, dvv));
        const __m256 ww = _mm256_cvtepi32_ps(gw);
        __m256 rs = _mm256_rsqrt_ps(vv);                                   // 1/sqrt(v), 12 bits ...
        rs = _mm256_mul_ps(_mm256_mul_ps(half, rs), _mm256_fnmadd_ps(_mm256_mul_ps(vv, rs), rs, three));   // ... one Newton step
        const __m256 s = _mm256_fmadd_ps(_mm256_mul_ps(ww, rs), rs, _mm256_mul_ps(c, rs));                  // w/v + K*sqrt(lg/v)
        const __m256 m = _mm256_blendv_ps(_mm256_set1_ps(-1.f), s, _mm256_castsi256_ps(some));
        _mm256_store_ps(sc + k, m);
        top = _mm256_max_ps(top, m);
    }
    // first SLOT holding the maximum (ties: the lowest slot, as the scalar scan)
    __m256 t = _mm256_max_ps(top, _mm256_permute2f128_ps(top, top, 1));
    t = _mm256_max_ps(t, _mm256_shuffle_ps(t, t, 0x4E));
    t = _mm256_max_ps(t, _mm256_shuffle_ps(t, t, 0xB1));
    if (!(_mm_cvtss_f32(_mm256_castps256_ps128(t)) > 0.f)) return -1;
    for (int k = 0; k < slots; k += 8) {
        const int hit = _mm256_movemask_ps(_mm256_cmp_ps(_mm256_load_ps(sc + k), t, _CMP_EQ_OQ));
        if (hit) {
            int best = 99;
            for (int j = 0; j < 8; j++) if ((hit >> j & 1) && kSlotOfLane[j] < best) best = kSlotOfLane[j];
            return k + best;
        }
    }
    return -1;
}

// the same scan with two gathers per 8 children (stride 4 dwords) instead of loads + shuffles: the default, the fastest of
// the three on the GPU box's host (profiles/r02_host_tree_layout.txt); SQUAVA_FAST_SCAN=shuffle / scalar select the others
__attribute__((target("avx2,fma")))
int best_child_avx2_gather(const FNode* blk, int slots, float k_sqrt_lg, const uint32_t* dv) {
    alignas(32) float sc[32];
    alignas(32) static const uint32_t zero[32] = {0};
    if (!dv) dv = zero;
    const __m256 c = _mm256_set1_ps(k_sqrt_lg), half = _mm256_set1_ps(0.5f), three = _mm256_set1_ps(3.0f);
    const __m256i stride = _mm256_setr_epi32(0, 4, 8, 12, 16, 20, 24, 28);
    __m256 top = _mm256_set1_ps(-1.f);
    const int* base = reinterpret_cast<const int*>(blk);
    for (int k = 0; k < slots; k += 8) {
        const __m256i live = _mm256_cmpgt_epi32(_mm256_set1_epi32(slots - k), _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7));
        const __m256i gv = _mm256_mask_i32gather_epi32(_mm256_setzero_si256(), base + 4 * k, stride, live, 4);
        const __m256i gw = _mm256_mask_i32gather_epi32(_mm256_setzero_si256(), base + 4 * k + 1, stride, live, 4);
        const __m256i some = _mm256_andnot_si256(_mm256_cmpeq_epi32(gv, _mm256_setzero_si256()), live);
        const __m256 vv = _mm256_cvtepi32_ps(_mm256_add_epi32(gv, _mm256_loadu_si256((const __m256i*)(dv + k))));
        const __m256 ww = _mm256_cvtepi32_ps(gw);
        __m256 rs = _mm256_rsqrt_ps(vv);
        rs = _mm256_mul_ps(_mm256_mul_ps(half, rs), _mm256_fnmadd_ps(_mm256_mul_ps(vv, rs), rs, three));
        const __m256 s = _mm256_fmadd_ps(_mm256_mul_ps(ww, rs), rs, _mm256_mul_ps(c, rs));
        const __m256 m = _mm256_blendv_ps(_mm256_set1_ps(-1.f), s, _mm256_castsi256_ps(some));
        _mm256_store_ps(sc + k, m);
        top = _mm256_max_ps(top, m);
    }
    __m256 t = _mm256_max_ps(top, _mm256_permute2f128_ps(top, top, 1));
    t = _mm256_max_ps(t, _mm256_shuffle_ps(t, t, 0x4E));
    t = _mm256_max_ps(t, _mm256_shuffle_ps(t, t, 0xB1));
    if (!(_mm_cvtss_f32(_mm256_castps256_ps128(t)) > 0.f)) return -1;
    for (int k = 0; k < slots; k += 8) {
        const int hit = _mm256_movemask_ps(_mm256_cmp_ps(_mm256_load_ps(sc + k), t, _CMP_EQ_OQ));
        if (hit) return k + __builtin_ctz((unsigned)hit);
    }
    return -1;
}

int best_child_float(const FNode* blk, int slots, float k_sqrt_lg, const uint32_t* dv) {
    static const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
    static const int mode = [] { const char* e = std::getenv("SQUAVA_FAST_SCAN"); return !e ? 1 : e[0] == 'g' ? 1 : e[0] == 's' ? (e[1] == 'h' ? 0 : 2) : 1; }();
    if (!avx2 || mode == 2) return best_child_scalar(blk, slots, k_sqrt_lg, dv);
    return mode == 1 ? best_child_avx2_gather(blk, slots, k_sqrt_lg, dv) : best_child_avx2(blk, slots, k_sqrt_lg, dv);
}

struct Chunk { uint64_t at = 0, end = 0; };            // a thread's private piece of the arena
constexpr uint64_t kChunk = 1 << 16;                   // slots per piece

// Every walk crosses the first two levels, so their counters are the lines all cores fight over.  In
// throughput mode each thread keeps its additions to those (at most 25 + 25*24) nodes privately and
// hands them over once per batch: a thread sees the tree as of the start of the batch plus its own
// selections in flight (virtual loss per thread).
struct Shard {
    uint32_t v1[32], v2[25][32];       // visits added to the root's children / grandchildren, by slot
    uint64_t w1[32], w2[25][32];       // wins, update phase
    Shard() { std::memset(this, 0, sizeof *this); }
};

struct Leaf { uint32_t node; uint32_t x, o; int8_t to_move; bool fresh; uint8_t depth; };

}  // namespace

FastResult fast_uct(const GameState& rootstate, long long itermax, double UCTK, const FastOptions& opt) {
    static std::atomic<uint64_t> stream_counter{1ull << 40};
    auto next_stream = [&] { return opt.stream_base + stream_counter.fetch_add(1); };
    const Hooks* hk = opt.hooks;
    int T = opt.threads > 0 ? opt.threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (hk && hk->intn) T = 1;                                   // a caller-supplied rand.Intn is one stream
    const int P = opt.playouts_per_leaf < 1 ? 1 : opt.playouts_per_leaf;
    if (itermax > 4000000000ll) { std::fprintf(stderr, "fast_uct: counts are 32 bits wide, at most 4e9 playouts per decision\n"); std::exit(2); }
    const long long B0 = opt.batch < 1 ? 1 : opt.batch;
    // worst case: every selection makes one node whose block has up to 24 slots... bounded in practice by
    // (selections + 1) blocks of <= 25 slots; the arena is virtual memory, pages are touched on use
    const uint64_t selections = (uint64_t)((itermax + P - 1) / P);
    // every thread claims arena space in private kChunk-slot pieces: each may leave one piece mostly unused
    uint64_t cap = opt.max_nodes ? opt.max_nodes : selections * 26 + 64 + (uint64_t)T * kChunk;
    if (cap > 0xFFFFFFF0ull) cap = 0xFFFFFFF0ull;
    FNode* nodes = (FNode*)mmap(nullptr, cap * sizeof(FNode), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (nodes == MAP_FAILED) { std::fprintf(stderr, "fast_uct: cannot reserve %llu node slots\n", (unsigned long long)cap); std::exit(2); }
    madvise(nodes, cap * sizeof(FNode), MADV_HUGEPAGE);          // random access over GBs: 2 MB pages if the kernel gives them
    std::atomic<uint64_t> next_free{2};                         // 0 = null, 1 = root
    std::atomic<bool> overflow{false};

    const uint32_t root = 1;
    {
        bool terminal;
        if (hk && hk->terminal) terminal = hk->terminal(rootstate.board);
        else {
            uint8_t flags = 0; sq_board rb = rootstate.board.raw();
            must(sq_classify(&rb, 1, &flags, nullptr, nullptr), "sq_classify");
            terminal = flags != 0;
        }
        nodes[root].untried.store(terminal ? 0u : rootstate.board.empty());
    }
    const double factor = opt.ucb_factor2 ? 2. : 1.;
    const bool exact = opt.exact_ucb || (hk && hk->intn);

    std::vector<Chunk> chunks((size_t)T);
    std::vector<Shard> shards(exact ? 0 : (size_t)T);
    std::vector<Rng> rngs;
    if (opt.rng_seed) { std::mt19937_64 g(opt.rng_seed); for (int t = 0; t < T; t++) rngs.emplace_back(g()); }
    else for (int t = 0; t < T; t++) rngs.emplace_back(host_rng()());

    // One walk from the root to a leaf is a chain of dependent cache misses (each level's child block is
    // cold once the tree is large), so every thread advances a GROUP of walks in turn, one level per
    // turn, prefetching the block the next turn will scan.  Visits are added on the way down (virtual
    // loss, visible to every later step of every walk); wins follow when the batch returns.
    struct Walk { uint32_t idx, x, o; int pjm; int depth; int slot1, slot2; bool fresh, done; };
    auto prefetch_block = [&](uint32_t fc, int slots) {
        const char* p = (const char*)&nodes[fc];
        const int bytes = slots * (int)sizeof(FNode);
        for (int off = 0; off < bytes; off += 64) __builtin_prefetch(p + off, 0, 1);
    };
    auto start_walk = [&](Walk& w, uint32_t* path) {
        w.idx = root; w.x = rootstate.board.x; w.o = rootstate.board.o; w.pjm = rootstate.playerJustMoved;
        w.depth = 0; w.slot1 = 0; w.slot2 = 0; w.fresh = false; w.done = false;
        // every walk passes the root: in throughput mode its visit is added once per thread and batch instead
        if (exact) nodes[root].visits.fetch_add((uint32_t)P, std::memory_order_release);
        path[w.depth++] = root;
    };
    // one level of one walk; returns true when the walk has reached its leaf
    auto step = [&](Rng& rng, Chunk& ck, Shard* sh, Walk& w, uint32_t* path) -> bool {
        FNode& nd = nodes[w.idx];
        const uint32_t u = nd.untried.load(std::memory_order_acquire);
        if (u & kPending) return true;                       // still waiting for its first rollout: a leaf
        const uint32_t um = u & kFull;
        const uint32_t emp = kFull & ~(w.x | w.o);
        if (um) {                                            // expand (mcts.go:166-171)
            const int cnt = __builtin_popcount(um);
            const int r = hk && hk->intn ? hk->intn(cnt) : rng.intn(cnt);
            const int m = nth_set_bit(um, r);
            const uint32_t bit = 1u << m;
            if (!(nd.untried.fetch_and(~bit, std::memory_order_acq_rel) & bit)) return false;   // another thread took it: look again
            const int slots = __builtin_popcount(emp);
            uint32_t fc = nd.first_child.load(std::memory_order_acquire);
            if (!fc) {
                if (ck.at + (uint64_t)slots > ck.end) { ck.at = next_free.fetch_add(kChunk); ck.end = ck.at + kChunk; }
                const uint64_t base = ck.at;
                ck.at += (uint64_t)slots;
                if (base + (uint64_t)slots >= cap) { overflow.store(true); nd.untried.fetch_or(bit); return true; }
                uint32_t expect = 0;
                if (nd.first_child.compare_exchange_strong(expect, (uint32_t)base, std::memory_order_acq_rel)) fc = (uint32_t)base;
                else fc = expect;                            // lost the race: that block is simply not used
            }
            const uint32_t child = fc + (uint32_t)__builtin_popcount(emp & (bit - 1u));
            w.pjm = -w.pjm;
            if (w.pjm == MAXIMIZER) w.x |= bit; else w.o |= bit;
            FNode& c = nodes[child];
            c.untried.store(kPending | ((uint32_t)(slots - cnt) << kRankShift), std::memory_order_relaxed);
            c.visits.store((uint32_t)P, std::memory_order_release);   // nobody touches a child before its visits are non-zero: a plain store, no locked RMW on a cold line
            if (w.depth == 1) w.slot1 = (int)(child - fc); else if (w.depth == 2) w.slot2 = (int)(child - fc);
            w.idx = child;
            path[w.depth++] = child;
            w.fresh = true;
            return true;
        }
        const uint32_t fc = nd.first_child.load(std::memory_order_acquire);
        if (!fc) return true;                                // terminal node: no moves at all
        // select (mcts.go:158-161, bestMove :218-238): first maximum in creation order
        const int slots = __builtin_popcount(emp);
        // this walk's own visit is already on the node (added on the way down): the reference sees it only after the rollout
        // w.depth == 1: at the root, 2: at one of its children -- the levels whose additions stay in the shard
        uint32_t* dv = !sh ? nullptr : w.depth == 1 ? sh->v1 : w.depth == 2 ? sh->v2[w.slot1] : nullptr;
        // visits this thread still holds in its shard for THIS node (its own walk's included)
        const uint32_t mine = !sh ? 0u : w.depth == 2 ? sh->v1[w.slot1] : w.depth == 3 ? sh->v2[w.slot1][w.slot2] : 0u;
        uint32_t seen = nd.visits.load(std::memory_order_relaxed) + mine - ((exact || w.idx != root) ? (uint32_t)P : 0u);
        // Other threads' visits to a first- or second-level node sit in THEIR shards until the batch ends, so a node
        // can be fully expanded while this thread has seen next to nothing of it.  Every child was made by a walk
        // through the node: that many visits it has had for certain (keeps ln N positive, the scores above 0).
        if (!exact && seen < (uint32_t)slots * (uint32_t)P + 1u) seen = (uint32_t)slots * (uint32_t)P + 1u;
        int best = -1;
        if (exact) {
            const double lg = factor * std::log((double)seen);
            double bestscore = kEps; uint32_t best_rank = 0;
            for (int k = 0; k < slots; k++) {
                const FNode& c = nodes[fc + (uint32_t)k];
                const uint32_t v = c.visits.load(std::memory_order_acquire);
                if (!v) continue;                                // claimed, still being set up
                const double sc = (double)c.wins.load(std::memory_order_relaxed) / ((double)v + kEps) + UCTK * std::sqrt(lg / ((double)v + kEps));
                const uint32_t rank = (c.untried.load(std::memory_order_relaxed) & kRankMask) >> kRankShift;
                if (sc > bestscore || (best >= 0 && sc == bestscore && rank < best_rank)) { bestscore = sc; best = k; best_rank = rank; }
            }
        } else {
            best = best_child_float(&nodes[fc], slots, (float)UCTK * __builtin_sqrtf((float)factor * fast_logf(seen)), dv);
        }
        if (best < 0 && !exact) {
            // no child scored above zero: take any child that is set up rather than wait (a walk that waits keeps
            // its thread's shard from being handed over, which others may be waiting for)
            for (int k = 0; k < slots && best < 0; k++) if (nodes[fc + (uint32_t)k].visits.load(std::memory_order_acquire)) best = k;
        }
        if (best < 0) {
            if (T == 1) {                                    // the reference dereferences nil here (mcts.go:161)
                std::fprintf(stderr, "Node has %d children\nplayer just moved %d\n", slots, w.pjm);
                std::exit(2);
            }
            return false;                                    // every child is still being set up by another thread
        }
        w.idx = fc + (uint32_t)best;
        FNode& c = nodes[w.idx];
        // the scan saw this child's visits non-zero through plain (vector) loads; this load pairs with the
        // creator's release store, so the child's other fields are visible before they are read
        (void)c.visits.load(std::memory_order_acquire);
        if (dv) { dv[best] += (uint32_t)P; if (w.depth == 1) w.slot1 = best; else w.slot2 = best; }
        else c.visits.fetch_add((uint32_t)P, std::memory_order_release);
        const uint32_t bit = 1u << nth_set_bit(emp, best);          // the child's move is its slot among the empty cells
        w.pjm = -w.pjm;
        if (w.pjm == MAXIMIZER) w.x |= bit; else w.o |= bit;
        path[w.depth++] = w.idx;
        const uint32_t next = c.first_child.load(std::memory_order_relaxed);
        if (next) prefetch_block(next, slots - 1);
        return false;
    };
    constexpr int kPathLen = 27;
    // walks a thread keeps in flight / paths it prefetches ahead while updating: memory-level parallelism knobs
    auto env_int = [](const char* name, int dflt, int lo, int hi) { const char* e = std::getenv(name); int v = e && *e ? std::atoi(e) : dflt; return v < lo ? lo : v > hi ? hi : v; };
    const int group = (hk && hk->intn) ? 1 : env_int("SQUAVA_FAST_GROUP", 8, 1, 64);
    const long long kAhead = env_int("SQUAVA_FAST_AHEAD", 4, 1, 64);

    FastResult res;
    std::memset(&res, 0, sizeof res);
    Pool pool(T);
    // One batch: its leaves, the paths that led to them, and the playout results.  Two of them alternate so that the GPU
    // call of batch k overlaps the selection of batch k+1 (double-buffered leaf batches).
    struct Batch {
        long long B = 0; int p = 1;
        std::vector<Leaf> leaves; std::vector<uint32_t> paths; std::vector<sq_board> boards; std::vector<int8_t> to_move;
        std::vector<uint64_t> out; std::vector<uint8_t> packed;
    };
    Batch bufs[2];
    long long done = 0, planned = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    // sizes of the next batch; false when the budget is used up
    auto plan = [&](Batch& bt) -> bool {
        const long long remaining = itermax - planned;
        if (remaining <= 0) return false;
        bt.p = P;
        if (bt.p > remaining) bt.p = (int)remaining;
        bt.B = B0;
        if (bt.B * bt.p > remaining) bt.B = remaining / bt.p;
        if (bt.B < 1) bt.B = 1;
        planned += bt.B * bt.p;
        bt.leaves.resize((size_t)bt.B); bt.boards.resize((size_t)bt.B); bt.to_move.resize((size_t)bt.B); bt.out.assign((size_t)bt.B * 4, 0);
        bt.paths.resize((size_t)bt.B * kPathLen);
        return true;
    };
    auto select_phase = [&](Batch& bt) {
        pool.run([&](int t) {
            const long long lo = bt.B * t / T, hi = bt.B * (t + 1) / T;
            Walk walks[64];
            long long slot_of[64];
            Chunk& ck = chunks[(size_t)t];
            Shard* sh = exact ? nullptr : &shards[(size_t)t];
            long long next = lo;
            int active = 0;
            for (; active < group && next < hi; active++, next++) { slot_of[active] = next; start_walk(walks[active], &bt.paths[(size_t)next * kPathLen]); }
            if (!exact && hi > lo) nodes[root].visits.fetch_add((uint32_t)((hi - lo) * P), std::memory_order_release);
            while (active > 0) {
                for (int g = 0; g < active; g++) {
                    Walk& w = walks[g];
                    const long long k = slot_of[g];
                    if (!step(rngs[(size_t)t], ck, sh, w, &bt.paths[(size_t)k * kPathLen])) continue;
                    Leaf& lf = bt.leaves[(size_t)k];
                    lf.node = w.idx; lf.x = w.x; lf.o = w.o; lf.to_move = (int8_t)-w.pjm; lf.fresh = w.fresh; lf.depth = (uint8_t)w.depth;
                    bt.boards[(size_t)k] = sq_board{w.x, w.o};
                    bt.to_move[(size_t)k] = lf.to_move;
                    if (next < hi) { slot_of[g] = next; start_walk(w, &bt.paths[(size_t)next * kPathLen]); next++; }
                    else { walks[g] = walks[active - 1]; slot_of[g] = slot_of[active - 1]; active--; g--; }
                }
            }
            if (sh) {                                        // hand this thread's first- and second-level visits over
                const uint32_t fc1 = nodes[root].first_child.load(std::memory_order_acquire);
                for (int a = 0; a < 25 && fc1; a++) {
                    if (sh->v1[a]) { nodes[fc1 + (uint32_t)a].visits.fetch_add(sh->v1[a], std::memory_order_release); sh->v1[a] = 0; }
                    const uint32_t fc2 = nodes[fc1 + (uint32_t)a].first_child.load(std::memory_order_acquire);
                    for (int b = 0; b < 25; b++)
                        if (sh->v2[a][b]) { nodes[fc2 + (uint32_t)b].visits.fetch_add(sh->v2[a][b], std::memory_order_release); sh->v2[a][b] = 0; }
                }
            }
        });
    };
    auto gpu_phase = [&](Batch& bt) {
        const uint64_t sid = next_stream();
        if (hk && hk->playouts) hk->playouts(bt.boards.data(), bt.to_move.data(), bt.B, (uint64_t)bt.p, sid, (uint64_t(*)[4])bt.out.data());
        else if (bt.p == 1) {
            // one playout per leaf: one outcome byte per leaf comes back (1 byte instead of 32), unpacked below
            bt.packed.resize((size_t)bt.B);
            if (opt.device_slot >= 0) must(sq_playouts_packed_on(opt.device_slot, bt.boards.data(), bt.to_move.data(), bt.B, sid, bt.packed.data()), "sq_playouts_packed_on");
            else must(sq_playouts_packed(bt.boards.data(), bt.to_move.data(), bt.B, sid, bt.packed.data()), "sq_playouts_packed");
            for (long long k = 0; k < bt.B; k++) {
                const uint8_t c = bt.packed[(size_t)k];
                uint64_t* r = &bt.out[(size_t)k * 4];
                r[0] = (c & 3) == 1; r[1] = (c & 3) == 2; r[2] = (c & 3) == 3; r[3] = c >> 2;
            }
        }
        else if (opt.device_slot >= 0) must(sq_playouts_on(opt.device_slot, bt.boards.data(), bt.to_move.data(), bt.B, (uint64_t)bt.p, sid, (uint64_t(*)[4])bt.out.data()), "sq_playouts_on");
        else must(sq_playouts(bt.boards.data(), bt.to_move.data(), bt.B, (uint64_t)bt.p, sid, (uint64_t(*)[4])bt.out.data()), "sq_playouts");
    };
    // the player who moved INTO the node at depth i of a path (the root is depth 0): sides alternate from the root's
    const bool root_max = rootstate.playerJustMoved == MAXIMIZER;
    auto is_max = [root_max](int i) { return (i & 1) ? !root_max : root_max; };
    auto update_phase = [&](Batch& bt) {
        pool.run([&](int t) {
            const long long lo = bt.B * t / T, hi = bt.B * (t + 1) / T;
            // the paths are known: their nodes are pulled in kAhead paths early
            uint64_t root_wins = 0;
            Shard* sh = exact ? nullptr : &shards[(size_t)t];
            const uint32_t fc1 = nodes[root].first_child.load(std::memory_order_acquire);
            for (long long k = lo; k < hi; k++) {
                if (k + kAhead < hi) {
                    const uint32_t* pp = &bt.paths[(size_t)(k + kAhead) * kPathLen];
                    const int d = bt.leaves[(size_t)(k + kAhead)].depth;
                    for (int i = 0; i < d; i++) __builtin_prefetch(&nodes[pp[i]], 1, 1);
                }
                const Leaf& lf = bt.leaves[(size_t)k];
                const uint64_t* r = &bt.out[(size_t)k * 4];
                if (lf.fresh) {
                    const uint32_t rank = nodes[lf.node].untried.load(std::memory_order_relaxed) & kRankMask;
                    nodes[lf.node].untried.store(rank | (r[3] != 0 ? (kFull & ~(lf.x | lf.o)) : 0u), std::memory_order_release);
                }
                const uint32_t* pp = &bt.paths[(size_t)k * kPathLen];
                root_wins += is_max(0) ? r[0] : r[1];                             // pp[0] is the root: one add per thread
                int from = 1;
                if (sh && lf.depth > 1) {                                          // first and second level: into the shard
                    const int a = (int)(pp[1] - fc1);
                    sh->w1[a] += is_max(1) ? r[0] : r[1];
                    from = 2;
                    if (lf.depth > 2) {
                        const int b = (int)(pp[2] - nodes[pp[1]].first_child.load(std::memory_order_relaxed));
                        sh->w2[a][b] += is_max(2) ? r[0] : r[1];
                        from = 3;
                    }
                }
                for (int i = from; i < lf.depth; i++) {                            // Update, mcts.go:190-192,268-271
                    FNode& nd = nodes[pp[i]];
                    const uint32_t add = (uint32_t)(is_max(i) ? r[0] : r[1]);
                    if (add) nd.wins.fetch_add(add, std::memory_order_relaxed);
                }
            }
            if (root_wins) nodes[root].wins.fetch_add((uint32_t)root_wins, std::memory_order_relaxed);
            if (sh && fc1)
                for (int a = 0; a < 25; a++) {
                    if (sh->w1[a]) { nodes[fc1 + (uint32_t)a].wins.fetch_add((uint32_t)sh->w1[a], std::memory_order_relaxed); sh->w1[a] = 0; }
                    const uint32_t fc2 = nodes[fc1 + (uint32_t)a].first_child.load(std::memory_order_relaxed);
                    for (int b = 0; b < 25; b++)
                        if (sh->w2[a][b]) { nodes[fc2 + (uint32_t)b].wins.fetch_add((uint32_t)sh->w2[a][b], std::memory_order_relaxed); sh->w2[a][b] = 0; }
                }
        });
    };
    // Pipelined (throughput mode): while the GPU plays batch k the host threads already select batch k+1 -- they see the
    // tree with batch k's visits on it (virtual loss) but not yet its wins.  The exact mode (reference tie rule, caller-
    // supplied rand.Intn) keeps the strict select / play / update sequence the golden trees were grown with.
    const bool pipelined = !exact && (!(hk && hk->playouts) || std::getenv("SQUAVA_FAST_PIPELINE") != nullptr) &&
                           std::getenv("SQUAVA_FAST_NO_PIPELINE") == nullptr;
    int cur = 0;
    bool have = plan(bufs[cur]);
    if (have) { auto t0 = now(); select_phase(bufs[cur]); res.select_seconds += secs(t0, now()); }
    while (have && !overflow.load()) {
        Batch& bt = bufs[cur];
        bool have_next = false;
        auto t1 = now();
        if (pipelined) {
            std::thread gpu_thread([&] { gpu_phase(bt); });
            have_next = plan(bufs[cur ^ 1]);
            auto s0 = now();
            if (have_next) select_phase(bufs[cur ^ 1]);
            auto s1 = now();
            gpu_thread.join();
            res.select_seconds += secs(s0, s1);
            res.gpu_seconds += secs(s1, now());          // what the GPU call still took after the selection was done
        } else {
            gpu_phase(bt);
            res.gpu_seconds += secs(t1, now());
        }
        auto t2 = now();
        update_phase(bt);
        res.update_seconds += secs(t2, now());
        res.batches++;
        done += bt.B * bt.p;
        if (!pipelined) {
            have_next = plan(bufs[cur ^ 1]);
            if (have_next) { auto t0 = now(); select_phase(bufs[cur ^ 1]); res.select_seconds += secs(t0, now()); }
        }
        cur ^= 1;
        have = have_next;
    }
    if (overflow.load()) { std::fprintf(stderr, "fast_uct: node arena exhausted (%llu slots)\n", (unsigned long long)cap); std::exit(2); }
    res.iterations = done;
    res.nodes = next_free.load();
    res.threads = T;
    res.huge_page_kb = -1;
    if (FILE* f = std::fopen("/proc/self/smaps_rollup", "r")) {
        char line[256];
        while (std::fgets(line, sizeof line, f)) if (std::sscanf(line, "AnonHugePages: %lld kB", &res.huge_page_kb) == 1) break;
        std::fclose(f);
    }
    // children of the root in creation order; the move returned is argmax UCB1 (mcts.go:196-198)
    const uint32_t fc = nodes[root].first_child.load();
    const int slots = __builtin_popcount(rootstate.board.empty());
    int order[25], n = 0;
    for (int k = 0; k < slots && fc; k++) if (nodes[fc + (uint32_t)k].visits.load()) order[n++] = k;
    auto rank_of = [&](int k) { return (nodes[fc + (uint32_t)k].untried.load() & kRankMask) >> kRankShift; };
    std::sort(order, order + n, [&](int a, int b) { return rank_of(a) < rank_of(b); });
    const double lg = factor * std::log((double)nodes[root].visits.load());
    double bestscore = kEps;
    res.best_move = -1;
    for (int i = 0; i < n; i++) {
        const FNode& c = nodes[fc + (uint32_t)order[i]];
        const double v = (double)c.visits.load(), w = (double)c.wins.load();
        res.child_move[i] = nth_set_bit(rootstate.board.empty(), order[i]); res.child_visits[i] = v; res.child_wins[i] = w;
        const double s = w / (v + kEps) + UCTK * std::sqrt(lg / (v + kEps));
        if (s > bestscore) { bestscore = s; res.best_move = res.child_move[i]; res.value = (int)(1000. * s); }
    }
    res.n_children = n;
    res.root_visits = (double)nodes[root].visits.load();
    munmap(nodes, cap * sizeof(FNode));
    if (res.best_move < 0) { std::fprintf(stderr, "Node has %d children\n", n); std::exit(2); }
    return res;
}

RootParallelResult root_parallel_uct(const GameState& rootstate, long long itermax, double UCTK, const FastOptions& opt,
                                     int n_searchers) {
    if (n_searchers < 1) n_searchers = 1;
    const int devices = std::max(1, sq_device_count());
    int T = opt.threads > 0 ? opt.threads : (int)std::thread::hardware_concurrency();
    if (T < n_searchers) T = n_searchers;
    RootParallelResult R;
    R.searcher.resize((size_t)n_searchers);
    const uint64_t seed0 = opt.rng_seed ? opt.rng_seed : host_rng()();
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int i = 0; i < n_searchers; i++)
        th.emplace_back([&, i] {
            FastOptions o = opt;
            o.threads = T * (i + 1) / n_searchers - T * i / n_searchers;
            o.device_slot = opt.hooks ? -1 : i % devices;
            o.stream_base = opt.stream_base + ((uint64_t)(i + 1) << 48);
            o.rng_seed = seed0 + 0x9e3779b97f4a7c15ull * (uint64_t)(i + 1);
            const long long share = itermax * (i + 1) / n_searchers - itermax * i / n_searchers;
            R.searcher[(size_t)i] = fast_uct(rootstate, share, UCTK, o);
        });
    for (auto& t : th) t.join();
    R.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // merge: sum the root children's statistics by move, then bestMove over the sums
    FastResult& M = R.merged;
    std::memset(&M, 0, sizeof M);
    double v[25] = {0}, w[25] = {0};
    for (const FastResult& r : R.searcher) {
        for (int k = 0; k < r.n_children; k++) { v[r.child_move[k]] += r.child_visits[k]; w[r.child_move[k]] += r.child_wins[k]; }
        M.iterations += r.iterations; M.batches += r.batches; M.root_visits += r.root_visits; M.nodes += r.nodes;
        M.select_seconds = std::max(M.select_seconds, r.select_seconds); M.gpu_seconds = std::max(M.gpu_seconds, r.gpu_seconds);
        M.update_seconds = std::max(M.update_seconds, r.update_seconds); M.threads += r.threads;
    }
    const double factor = opt.ucb_factor2 ? 2. : 1.;
    const double lg = factor * std::log(M.root_visits);
    double bestscore = kEps;
    M.best_move = -1;
    for (int m = 0; m < 25; m++) {
        if (!(v[m] > 0)) continue;
        const int i = M.n_children++;
        M.child_move[i] = m; M.child_visits[i] = v[m]; M.child_wins[i] = w[m];
        const double s = w[m] / (v[m] + kEps) + UCTK * std::sqrt(lg / (v[m] + kEps));
        if (s > bestscore) { bestscore = s; M.best_move = m; M.value = (int)(1000. * s); }
    }
    if (M.best_move < 0) { std::fprintf(stderr, "Node has %d children\n", M.n_children); std::exit(2); }
    return R;
}

}  // namespace mcts
}  // namespace squava
