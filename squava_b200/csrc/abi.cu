// abi.cu -- the extern "C" surface of libsquava_b200.so (see include/squava_b200.h).
//
// Host side only does plumbing: device contexts, staging buffers for the
// host-pointer entry points, sharding over the devices named in sq_init(), and
// launch accounting.  All arithmetic of the hot path lives in kernels.cuh /
// alphabeta.cuh.  There is no CPU fallback anywhere in this file.
#include "../../include/squava_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "kernels.cuh"
#include "alphabeta.cuh"
#include "negascout.cuh"
#include "sq_tables.h"

namespace {

// A call context: everything one call in flight on a device needs -- its own stream, staging
// buffers, pinned block, timing events and scheduling counters.  A device carries kCtx of them, so
// calls arriving on different host threads overlap on the SAME device (and calls on different
// devices never meet at all); a context is held for the duration of one call.
// NegaScout call items in flight (NsSplit): a stream of its own and pinned, mapped memory the kernel reads the items from
// and writes the results to in place -- a flight costs one launch and one event, no copies
struct NsFlight {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    sq::NsItem* items = nullptr;
    sq::NsItemOut* out = nullptr;
    int32_t* landed = nullptr;                           // item indices in the order they finish (-1: not yet), mapped too
    unsigned int* n_landed = nullptr;                    // device memory: next slot of `landed`
    size_t cap = 0, n = 0, seen = 0;
    bool busy = false;
    std::vector<std::pair<int32_t, uint32_t>> sent;      // (call, generation) of every item
};
constexpr int kNsFlights = 32;

struct Ctx {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // around the most recent playout kernel
    bool timed = false;
    unsigned long long* counters = nullptr;     // ring of playout round counters, one per launch in flight
    unsigned launch_seq = 0;
    // staging for host-pointer calls (grown on demand)
    void* buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[6] = {0, 0, 0, 0, 0, 0};
    // small-call path: pinned host staging [boards | to_move | out] and one device block
    // [counter | out | boards | to_move], so a call is memcpy-in, H2D, memset, kernel, D2H
    uint8_t* pin = nullptr;
    uint8_t* small = nullptr;
    NsFlight flight[kNsFlights];                // made on first use
};

constexpr int kCtx = 4;

struct Device {
    int id = -1;
    int sm_count = 0;
    int playout_blocks_per_sm = 0;
    int ab_blocks_per_sm = 0;
    Ctx ctx[kCtx];
    std::atomic<unsigned> rr{0};
    std::atomic<int> last_timed{-1};            // context that ran the most recent playout kernel
};

constexpr int64_t kSmallLeaves = 1 << 16;      // per call and device

constexpr unsigned kCounterRing = 256;

static sq::NsTables g_ns_tables;   // host copy: which positions the fast NegaScout kernel may take

// g_state: exclusive for sq_init / sq_shutdown, shared for the duration of every other call, so the
// device table cannot change under a call in flight.  Below it every context has its own mutex.
std::shared_mutex g_state;
std::vector<std::unique_ptr<Device>> g_dev;
bool g_init = false;
uint64_t g_seed = 0;
std::atomic<uint64_t> g_launches{0};
thread_local char t_err[512] = "";

// Takes a free context of the device; when all are busy, queues on one of them.
struct CtxHold {
    Ctx* c = nullptr;
    std::unique_lock<std::mutex> lk;
    Ctx* operator->() const { return c; }
    Ctx& operator*() const { return *c; }
};
CtxHold acquire(Device& d) {
    CtxHold h;
    for (int i = 0; i < kCtx; i++) {
        std::unique_lock<std::mutex> lk(d.ctx[i].mu, std::try_to_lock);
        if (lk.owns_lock()) { h.c = &d.ctx[i]; h.lk = std::move(lk); return h; }
    }
    const unsigned k = d.rr.fetch_add(1) % kCtx;
    h.lk = std::unique_lock<std::mutex>(d.ctx[k].mu);
    h.c = &d.ctx[k];
    return h;
}

int fail_cuda(cudaError_t e, const char* what) {
    snprintf(t_err, sizeof t_err, "%s: %s", what, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? SQ_ERR_NOMEM : SQ_ERR_CUDA;
}
#define CK(call)                                             \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call);  \
    } while (0)

int reserve(Ctx& c, int which, size_t bytes) {
    if (c.cap[which] >= bytes) return SQ_OK;
    if (c.buf[which]) { CK(cudaStreamSynchronize(c.stream)); CK(cudaFree(c.buf[which])); }
    c.buf[which] = nullptr; c.cap[which] = 0;
    size_t want = bytes + bytes / 4 + 256;
    CK(cudaMalloc(&c.buf[which], want));
    c.cap[which] = want;
    return SQ_OK;
}

// contiguous share of n units for slot k of m
void share(int64_t n, int m, int k, int64_t& lo, int64_t& hi) {
    lo = n * k / m; hi = n * (k + 1) / m;
}

int launch_classify(Device& d, const sq_board* boards, int64_t n, uint8_t* flags, int8_t* wm,
                    int8_t* wa, cudaStream_t st) {
    if (n <= 0) return SQ_OK;
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)d.sm_count * 8;
    if (blocks > cap) blocks = cap;
    sq::classify_kernel<<<(unsigned)blocks, 256, 0, st>>>(boards, n, flags, wm, wa);
    g_launches++;
    CK(cudaGetLastError());
    return SQ_OK;
}

// counter == nullptr: zero `out`, take a counter from the context's ring and zero it.  Otherwise the
// caller has already zeroed both (they sit in one block: one memset).  packed != nullptr: one playout
// per leaf, the outcome goes to packed[i] as one byte instead of out[i][0..3] (sq_playouts_packed).
int launch_playouts(Device& d, Ctx& c, int ctx_index, const sq_board* leaves, const int8_t* to_move, int64_t n_leaves,
                    uint64_t leaf_index_base, uint64_t per_leaf, uint64_t stream_id,
                    uint64_t* out, cudaStream_t st, unsigned long long* counter = nullptr, uint8_t* packed = nullptr) {
    if (n_leaves <= 0) return SQ_OK;
    if (!counter && out) CK(cudaMemsetAsync(out, 0, (size_t)n_leaves * 4 * sizeof(uint64_t), st));
    if (per_leaf == 0) return SQ_OK;
    sq::PlayoutParams P;
    P.leaves = leaves; P.to_move = to_move; P.n_leaves = n_leaves;
    P.leaf_index_base = leaf_index_base; P.per_leaf = per_leaf;
    uint64_t S = (per_leaf + sq::kItemQuota - 1) / sq::kItemQuota;
    if (S > 0xffffffffull) { snprintf(t_err, sizeof t_err, "playouts_per_leaf too large"); return SQ_ERR_INVALID; }
    P.items_per_leaf = (uint32_t)S;
    P.quota_lo = (uint32_t)(per_leaf / S);
    P.quota_extra = (uint32_t)(per_leaf % S);
    P.key0 = (uint32_t)g_seed ^ (uint32_t)(stream_id >> 32);
    P.key1 = (uint32_t)(g_seed >> 32);
    P.ctr3 = (uint32_t)stream_id;
    P.one = 1u;
    sq::playout_round_keys(P);
    unsigned long long items = (unsigned long long)n_leaves * S;
    P.n_rounds = (int64_t)((items + 31) / 32);
    P.out = reinterpret_cast<unsigned long long*>(out);
    P.packed = packed;
    if (counter) P.round_counter = counter;
    else {
        // each launch gets its own counter: launches on different streams may overlap on the device
        P.round_counter = c.counters + (c.launch_seq++ % kCounterRing);
        CK(cudaMemsetAsync(P.round_counter, 0, sizeof(unsigned long long), st));
    }
    const int warps_per_block = sq::kPlayoutThreads / 32;
    int64_t blocks = (int64_t)d.sm_count * d.playout_blocks_per_sm;
    int64_t need = (P.n_rounds + warps_per_block - 1) / warps_per_block;
    if (blocks > need) blocks = need;
    size_t smem = (size_t)sq::kMaxEmpty * sq::kPlayoutThreads * sizeof(uint32_t);
    CK(cudaEventRecord(c.ev0, st));
    sq::playout_kernel<<<(unsigned)blocks, sq::kPlayoutThreads, smem, st>>>(P);
    g_launches++;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c.ev1, st));
    c.timed = true;
    d.last_timed.store(ctx_index);
    return SQ_OK;
}

// Host-buffer calls over several devices: each share runs on its own host thread (the alpha-beta search
// is host-driven, one synchronisation per pass).  fn(k) returns a status; its message is kept.
template <typename F>
int for_each_device(int m, F fn) {
    if (m == 1) return fn(0);
    std::vector<int> rcs(m, SQ_OK);
    std::vector<std::string> msgs(m);
    std::vector<std::thread> th;
    for (int k = 0; k < m; k++)
        th.emplace_back([&, k] { rcs[k] = fn(k); if (rcs[k]) msgs[k] = t_err; });
    for (auto& t : th) t.join();
    for (int k = 0; k < m; k++)
        if (rcs[k]) { snprintf(t_err, sizeof t_err, "%s", msgs[k].c_str()); return rcs[k]; }
    return SQ_OK;
}

// callers hold g_state (shared)
int check_ready() {
    if (!g_init) { snprintf(t_err, sizeof t_err, "sq_init() has not succeeded"); return SQ_ERR_NOT_INIT; }
    return SQ_OK;
}
int check_slot(int slot) {
    int rc = check_ready();
    if (rc) return rc;
    if (slot < 0 || slot >= (int)g_dev.size()) { snprintf(t_err, sizeof t_err, "bad device slot %d", slot); return SQ_ERR_INVALID; }
    return SQ_OK;
}

void destroy_device(Device& d) {
    if (d.id < 0) return;
    cudaSetDevice(d.id);
    for (auto& c : d.ctx) {
        if (c.stream) cudaStreamSynchronize(c.stream);
        for (int i = 0; i < 6; i++) if (c.buf[i]) cudaFree(c.buf[i]);
        if (c.counters) cudaFree(c.counters);
        if (c.small) cudaFree(c.small);
        if (c.pin) cudaFreeHost(c.pin);
        for (auto& f : c.flight) {
            if (f.stream) { cudaStreamSynchronize(f.stream); cudaStreamDestroy(f.stream); }
            if (f.done) cudaEventDestroy(f.done);
            if (f.items) cudaFreeHost(f.items);
            if (f.out) cudaFreeHost(f.out);
            if (f.landed) cudaFreeHost(f.landed);
            if (f.n_landed) cudaFree(f.n_landed);
            f = NsFlight();
        }
        if (c.ev0) cudaEventDestroy(c.ev0);
        if (c.ev1) cudaEventDestroy(c.ev1);
        if (c.stream) cudaStreamDestroy(c.stream);
        c.stream = nullptr; c.ev0 = c.ev1 = nullptr; c.timed = false; c.counters = nullptr; c.launch_seq = 0;
        for (int i = 0; i < 6; i++) { c.buf[i] = nullptr; c.cap[i] = 0; }
        c.pin = c.small = nullptr;
    }
    d.id = -1;
}

void destroy_all() {
    for (auto& d : g_dev) destroy_device(*d);
    g_dev.clear();
    g_init = false;
}

int init_device(Device& d, const sq::ScanTables& tables, const sq::AbTables& ab_tables, const sq::NsTables& ns_tables,
                const sq::NsCellTables& ns_cell_tables) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, d.id));
    if (prop.major != 10) {
        snprintf(t_err, sizeof t_err, "device %d is sm_%d%d; this library carries sm_100a code only", d.id, prop.major, prop.minor);
        return SQ_ERR_NO_DEVICE;
    }
    d.sm_count = prop.multiProcessorCount;
    CK(cudaSetDevice(d.id));
    {   // keep the alpha-beta workspace in the stream-ordered pool between calls
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, d.id));
        uint64_t keep = ~0ull;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (auto& c : d.ctx) {
        CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&c.ev0));
        CK(cudaEventCreate(&c.ev1));
        CK(cudaMalloc(&c.counters, kCounterRing * sizeof(unsigned long long)));
        CK(cudaMemset(c.counters, 0, kCounterRing * sizeof(unsigned long long)));
        CK(cudaHostAlloc((void**)&c.pin, (size_t)kSmallLeaves * (8 + 1 + 32), cudaHostAllocDefault));
        CK(cudaMalloc((void**)&c.small, 16 + (size_t)kSmallLeaves * (32 + 8 + 1)));
    }
    CK(cudaMemcpyToSymbol(sq::c_scan, &tables, sizeof tables));
    CK(cudaMemcpyToSymbol(sq::c_ab, &ab_tables, sizeof ab_tables));
    CK(cudaMemcpyToSymbol(sq::c_ns, &ns_tables, sizeof ns_tables));
    CK(cudaMemcpyToSymbol(sq::c_ns_cell, &ns_cell_tables, sizeof ns_cell_tables));
    size_t smem = (size_t)sq::kMaxEmpty * sq::kPlayoutThreads * sizeof(uint32_t);
    CK(cudaFuncSetAttribute(sq::playout_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.playout_blocks_per_sm, sq::playout_kernel,
                                                     sq::kPlayoutThreads, smem));
    if (d.playout_blocks_per_sm < 1) d.playout_blocks_per_sm = 1;
    CK(sq::ab_configure(&d.ab_blocks_per_sm));
    return SQ_OK;
}

bool bad_board(const sq_board& b) { return (b.x & b.o) || ((b.x | b.o) >> 25); }

// NegaScout, one position over many warps (negascout.cuh: NsItem).  negaScout's recursion (negascout.go:232-266) and
// ChooseMove's loop over it (:92-113) are replayed HERE down to a split ply; every call below that is an item searched on the
// device.  A loop makes its first call alone (its window is the loop's own); after that ALL remaining children are
// searched at once with the null window (-alpha-1, -alpha) they get if alpha does not move.  Results are consumed in the
// reference's order, and only if alpha is still the one they were searched with: a null-window result above the score
// needs the re-search (-beta, -cur); if that raises alpha, the unconsumed calls are dropped and the rest is searched again
// under the new alpha; a cut-off drops the rest.  Every value, the visiting order, movekeeper's set and the leaf counter
// are exactly the sequential recursion's; what is speculative is only WHICH calls are in flight.
struct NsCall {
    int32_t parent = -1;            // the loop waiting for this call (-1: ChooseMove's, the loop IS the position)
    uint32_t gen = 0;               // bumped when the slot is released: results of dropped calls are recognised by it
    int32_t pos = 0;                // index in the list of positions
    uint8_t state = 0;              // 0 free, 1 item waiting for a flight, 4 item in flight, 2 loop, 3 done
    uint8_t ply = 0;                // of the node
    uint8_t path[7];
    int32_t a = 0, b = 0;           // the window the node is called with
    int32_t ret = 0;
    uint64_t nodes = 0;             // staticValue calls below, its own included
    // the loop (state 2), negascout.go:240-262
    uint32_t x = 0, o = 0;
    int32_t alpha = 0, beta = 0, score = 0, n = 0, i = 0, n_kids = 0, spec_alpha = 0;
    int32_t again = -1;             // the re-search of child i
    int32_t first[25];              // the calls (-n, -alpha) of the children, -1: none
    uint8_t cell[25];
};

struct NsRootOut {                  // what ChooseMove leaves behind, movekeeper.go:26-44
    int keeper_max = -30000, first_best = -1, n_visited = 0;
    uint32_t keeper_mask = 0;
};

class NsSplit {
public:
    NsSplit(Device& d, Ctx& c, const sq_board* h_pos, const sq_board* d_pos, const int* list, int64_t n_list, int max_depth,
            const sq::NsScores& sc, int32_t (*values)[25], int8_t (*visit_order)[25])
        : d_(d), c_(c), h_pos_(h_pos), d_pos_(d_pos), list_(list), n_list_(n_list), max_depth_(max_depth), sc_(sc),
          values_(values), visit_order_(visit_order) {
        const char* e = getenv("SQ_NS_SPLIT_PLY");
        // loops at ply < split_ply run here: deep enough for the items to be small next to the whole search, and never a
        // loop at the horizon or just above it (those are scans inside the kernel)
        int wide = e ? atoi(e) : 3;
        const char* e2 = getenv("SQ_NS_SPLIT_PLY_NULL");
        int null_w = e2 ? atoi(e2) : 2;
        const int most = max_depth - 3 < 6 ? max_depth - 3 : 6;
        split_wide_ = wide < 1 ? 1 : wide; if (split_wide_ > most) split_wide_ = most; if (split_wide_ < 1) split_wide_ = 1;
        split_null_ = null_w < 1 ? 1 : null_w; if (split_null_ > split_wide_) split_null_ = split_wide_;
        // an item that turns out long gives up and becomes a loop here, its children items
        const char* e3 = getenv("SQ_NS_ITEM_BUDGET");
        item_budget_ = e3 ? (uint32_t)strtoul(e3, nullptr, 10) : 8192u;
        deepest_loop_ = most < 1 ? 0 : most;
        const char* e6 = getenv("SQ_NS_ITEM_BUDGET_WIDE");
        item_budget_wide_ = e6 ? (uint32_t)strtoul(e6, nullptr, 10) : item_budget_ / 4;
        const char* e4 = getenv("SQ_NS_FLIGHTS");
        n_flights_ = e4 ? atoi(e4) : 16;
        if (n_flights_ < 1) n_flights_ = 1;
        if (n_flights_ > kNsFlights) n_flights_ = kNsFlights;
    }

    int run(int32_t* best_value, uint32_t* best_mask, int8_t* first_best, uint64_t* nodes) {
        t_run_ = std::chrono::steady_clock::now();
        for (auto& f : c_.flight) if (f.busy) { CK(cudaStreamSynchronize(f.stream)); f.busy = false; }      // left by a call that failed
        order_max_.resize((size_t)n_list_ * 25); order_min_.resize((size_t)n_list_ * 25);
        roots_.resize((size_t)n_list_);
        root_call_.resize((size_t)n_list_);
        for (int64_t q = 0; q < n_list_; q++) {
            const sq_board& bd = h_pos_[list_[q]];
            const uint32_t x = bd.x & sq::kFull25, o = bd.o & sq::kFull25;
            sq::ns_reorder_host(g_ns_tables, x, o, &order_max_[(size_t)q * 25], &order_min_[(size_t)q * 25]);
            for (int k = 0; k < 25; k++) { values_[list_[q]][k] = SQ_NO_MOVE; if (visit_order_) visit_order_[list_[q]][k] = -1; }
            const int ci = alloc();
            NsCall& c = calls_[(size_t)ci];
            c.parent = -1; c.pos = (int32_t)q; c.ply = 0; c.a = -20000; c.b = 20000;
            start_loop(c, x, o);
            root_call_[(size_t)q] = ci;
            ready_.push_back(ci);
        }
        pump();
        // Items go out in FLIGHTS -- concurrent launches on their own streams -- and every item lands by itself: the kernel
        // appends its index to the flight's landing list in mapped memory when its result is written, and the host reads the
        // lists while the kernels run.  A long search holds up the loops that wait for it and nothing else.  Wide windows
        // (the long searches) and null windows fly apart.  New items are sent when nothing more is landing (or a few hundred
        // have piled up): results that arrive together make ONE next flight.
        int busy = 0;
        std::vector<std::pair<int32_t, uint32_t>> later;
        const sq::NsOut none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        const int64_t cap = (int64_t)d_.sm_count * 6;
        auto last_progress = std::chrono::steady_clock::now();
        bool progressed = true;
        for (unsigned long long spin = 0;; spin++) {
            if (!outbox_.empty() && (!progressed || busy == 0 || outbox_.size() >= 256)) {
                for (int kind = 0; kind < 2; kind++) {
                    size_t n = 0;
                    for (const auto& w : outbox_) n += wanted(w) && wide(w) == (kind == 0);
                    if (!n) continue;
                    NsFlight* f = nullptr;
                    for (int k = 0; k < n_flights_ && !f; k++) if (!c_.flight[k].busy) f = &c_.flight[k];
                    if (!f) break;                                 // all in the air: these wait for a flight to empty
                    int r;
                    if ((r = ready_flight(*f, n))) return r;
                    f->n = 0; f->seen = 0; f->sent.clear();
                    for (const auto& w : outbox_) {
                        if (!wanted(w) || wide(w) != (kind == 0)) continue;
                        const NsCall& k = calls_[(size_t)w.first];
                        sq::NsItem& it = f->items[f->n];
                        it.pos = list_[k.pos]; it.a = k.a; it.b = k.b; it.n_path = k.ply;
                        it.budget = k.ply > deepest_loop_ ? 0u : kind == 0 ? item_budget_wide_ : item_budget_;
                        it.deepest = (uint8_t)deepest_loop_;
                        memcpy(it.path, k.path, 7);
                        f->landed[f->n] = -1;
                        f->n++;
                        f->sent.push_back(w);
                        calls_[(size_t)w.first].state = 4;         // in the air
                    }
                    const int64_t want = ((int64_t)f->n + sq::kNsWarps - 1) / sq::kNsWarps;
                    CK(cudaMemsetAsync(f->n_landed, 0, sizeof(unsigned int), f->stream));
                    sq::negascout_fast_kernel<true><<<(int)(want < cap ? want : cap), sq::kNsWarps * 32, 0, f->stream>>>(
                        d_pos_, nullptr, (int)f->n, max_depth_, sc_, none, nullptr, f->items, f->out, 0ull, f->landed, f->n_landed);
                    g_launches += 1;
                    CK(cudaGetLastError());
                    CK(cudaEventRecord(f->done, f->stream));
                    f->busy = true; busy++;
                    rounds_++; items_ += f->n;
                }
                // what did not get a flight stays in the outbox
                later.clear();
                for (const auto& w : outbox_) if (wanted(w)) later.push_back(w);
                outbox_.swap(later);
            }
            if (busy == 0) {
                if (outbox_.empty()) break;
                snprintf(t_err, sizeof t_err, "NegaScout split driver: items left and nothing in flight"); return SQ_ERR_CUDA;
            }
            // take what has landed
            progressed = false;
            for (int k = 0; k < n_flights_; k++) {
                NsFlight& f = c_.flight[k];
                if (!f.busy) continue;
                bool finished = false;
                if ((spin & 1023u) == 1023u) {                      // a kernel that died would never fill its list
                    const cudaError_t e = cudaEventQuery(f.done);
                    if (e == cudaSuccess) finished = true;
                    else if (e != cudaErrorNotReady) return fail_cuda(e, "cudaEventQuery(flight)");
                }
                for (int pass = 0; pass < (finished ? 2 : 1); pass++)
                    while (f.seen < f.n) {
                        const int t = *((volatile int32_t*)f.landed + f.seen);
                        if (t < 0) break;
                        std::atomic_thread_fence(std::memory_order_acquire);
                        f.seen++;
                        progressed = true;
                        NsCall& k2 = calls_[(size_t)f.sent[(size_t)t].first];
                        if (k2.gen != f.sent[(size_t)t].second || k2.state != 4) { dropped_++; continue; }   // dropped while in flight
                        if (f.out[t].nodes == sq::kNsAborted) {                              // too long for one warp: its loop runs here
                            resume(f.sent[(size_t)t].first, f.out[t]);
                            expanded_++;
                            continue;
                        }
                        k2.state = 3; k2.ret = f.out[t].ret; k2.nodes = f.out[t].nodes;
                        ready_.push_back(k2.parent);
                    }
                if (f.seen == f.n) { f.busy = false; busy--; }
                else if (finished) { snprintf(t_err, sizeof t_err, "NegaScout split driver: a flight ended with %zu of %zu items landed", f.seen, f.n); return SQ_ERR_CUDA; }
            }
            if (progressed) { pump(); last_progress = std::chrono::steady_clock::now(); }
            else if ((spin & 0xfffffu) == 0xfffffu && std::chrono::duration<double>(std::chrono::steady_clock::now() - last_progress).count() > 60.) {
                snprintf(t_err, sizeof t_err, "NegaScout split driver: nothing landed for 60 s"); return SQ_ERR_CUDA;
            }
        }
        for (int64_t q = 0; q < n_list_; q++) {
            const NsCall& c = calls_[(size_t)root_call_[(size_t)q]];
            if (c.state != 3) { snprintf(t_err, sizeof t_err, "NegaScout split driver: a position did not finish"); return SQ_ERR_CUDA; }
            const NsRootOut& ro = roots_[(size_t)q];
            const int p = list_[q];
            if (best_value) best_value[p] = ro.keeper_mask ? ro.keeper_max : 0;
            if (best_mask) best_mask[p] = ro.keeper_mask;
            if (first_best) first_best[p] = (int8_t)ro.first_best;
            if (nodes) nodes[p] = c.nodes;
        }
        if (getenv("SQ_NS_TRACE"))
            fprintf(stderr, "ns split: %.4f s, %lld positions, split ply %d/%d, item budget %u, %llu flights, %llu items (%llu dropped in flight, %llu given up), %zu call slots\n",
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t_run_).count(), (long long)n_list_, split_wide_, split_null_, item_budget_, (unsigned long long)rounds_, (unsigned long long)items_,
                    (unsigned long long)dropped_, (unsigned long long)expanded_, calls_.size());
        return SQ_OK;
    }

private:
    bool wanted(const std::pair<int32_t, uint32_t>& w) const { const NsCall& k = calls_[(size_t)w.first]; return k.gen == w.second && k.state == 1; }
    bool wide(const std::pair<int32_t, uint32_t>& w) const { const NsCall& k = calls_[(size_t)w.first]; return k.b - k.a > 1; }
    int ready_flight(NsFlight& f, size_t n) {
        CK(cudaSetDevice(d_.id));
        if (!f.stream) CK(cudaStreamCreateWithFlags(&f.stream, cudaStreamNonBlocking));
        if (!f.done) CK(cudaEventCreateWithFlags(&f.done, cudaEventDisableTiming));
        if (!f.n_landed) CK(cudaMalloc((void**)&f.n_landed, sizeof(unsigned int)));
        if (f.cap < n) {
            CK(cudaStreamSynchronize(f.stream));                 // the last kernel that used the old buffers has left
            if (f.items) CK(cudaFreeHost(f.items));
            if (f.out) CK(cudaFreeHost(f.out));
            if (f.landed) CK(cudaFreeHost(f.landed));
            f.items = nullptr; f.out = nullptr; f.landed = nullptr; f.cap = 0;
            size_t cap = 1024;
            while (cap < n) cap *= 2;
            CK(cudaHostAlloc((void**)&f.items, cap * sizeof(sq::NsItem), cudaHostAllocMapped));
            CK(cudaHostAlloc((void**)&f.out, cap * sizeof(sq::NsItemOut), cudaHostAllocMapped));
            CK(cudaHostAlloc((void**)&f.landed, cap * sizeof(int32_t), cudaHostAllocMapped));
            f.cap = cap;
        }
        return SQ_OK;
    }
    int alloc() {
        if (!free_.empty()) { const int ci = free_.back(); free_.pop_back(); return ci; }
        calls_.emplace_back();
        return (int)calls_.size() - 1;
    }
    void release(int ci) {
        NsCall& c = calls_[(size_t)ci];
        c.state = 0; c.gen++;
        free_.push_back(ci);
    }
    // a call that is no longer wanted, with everything below it
    void drop(int ci) {
        stack_.clear(); stack_.push_back(ci);
        while (!stack_.empty()) {
            const int k = stack_.back(); stack_.pop_back();
            NsCall& c = calls_[(size_t)k];
            if (c.state == 2) {
                if (c.again >= 0) stack_.push_back(c.again);
                for (int j = 0; j < c.n_kids; j++) if (c.first[j] >= 0) stack_.push_back(c.first[j]);
            }
            release(k);
        }
    }
    void start_loop(NsCall& c, uint32_t x, uint32_t o) {
        c.state = 2; c.x = x; c.o = o; c.nodes = 0;
        c.alpha = c.a; c.beta = c.b; c.score = -30000; c.n = c.b; c.i = 0; c.spec_alpha = c.a; c.again = -1;
        const uint8_t* ord = (c.ply & 1) ? &order_min_[(size_t)c.pos * 25] : &order_max_[(size_t)c.pos * 25];
        const int n_root = 25 - __builtin_popcount(h_pos_[list_[c.pos]].x | h_pos_[list_[c.pos]].o);
        c.n_kids = 0;
        for (int k = 0; k < n_root; k++) if (!(((x | o) >> ord[k]) & 1u)) { c.first[c.n_kids] = -1; c.cell[c.n_kids++] = ord[k]; }
    }
    // An item gave up (negascout.cuh: NsItemOut).  Its node's loop runs here now, and so do the loops of the first children it
    // had followed; the deepest of them continues where the item stood, minus the child it was in.
    void resume(int ci, const sq::NsItemOut& got) {
        {
            NsCall& k = calls_[(size_t)ci];
            uint32_t x = h_pos_[list_[k.pos]].x & sq::kFull25, o = h_pos_[list_[k.pos]].o & sq::kFull25;
            for (int j = 0; j < k.ply; j++) { if (j & 1) o |= 1u << k.path[j]; else x |= 1u << k.path[j]; }
            start_loop(k, x, o);
        }
        bool exact = true;                                           // the whole chain could be rebuilt
        for (int j = 0; j < got.n_chain; j++) {
            NsCall& k = calls_[(size_t)ci];
            if (k.n_kids == 0 || k.cell[0] != got.cells[j] || k.ply >= deepest_loop_) { exact = false; break; }
            const int child = spawn(ci, 0, -k.beta, -k.alpha, true);
            calls_[(size_t)ci].first[0] = child;
            ci = child;
        }
        NsCall& k = calls_[(size_t)ci];
        if (exact && got.i_done > 0 && got.i_done < k.n_kids) {
            k.i = got.i_done; k.alpha = got.alpha; k.score = got.score; k.n = got.alpha + 1; k.spec_alpha = got.alpha;
            k.nodes = got.nodes_done;
        }
        ready_.push_back(ci);
    }
    // the call negaScout(child j of loop `pi`, window (a, b)): a loop of its own above the split ply, an item below
    int spawn(int pi, int j, int a, int b, bool as_loop = false) {
        const int ci = alloc();
        NsCall& p = calls_[(size_t)pi];
        NsCall& c = calls_[(size_t)ci];
        const int cell = p.cell[j];
        const bool x_moves = !(p.ply & 1);
        c.parent = pi; c.pos = p.pos; c.ply = (uint8_t)(p.ply + 1); c.a = a; c.b = b;
        memcpy(c.path, p.path, 7);
        c.path[p.ply] = (uint8_t)cell;
        const bool ends = sq::ns_move_ends_game(g_ns_tables, x_moves ? p.x : p.o, cell);
        if ((as_loop && !ends && c.ply <= deepest_loop_) || (!ends && c.ply < (b - a > 1 ? split_wide_ : split_null_))) {
            start_loop(c, p.x | (x_moves ? 1u << cell : 0u), p.o | (x_moves ? 0u : 1u << cell));
            ready_.push_back(ci);
        } else {
            c.state = 1;
            outbox_.emplace_back(ci, c.gen);
        }
        return ci;
    }
    void pump() {
        while (!ready_.empty()) { const int ci = ready_.back(); ready_.pop_back(); if (ci >= 0) advance(ci); }
    }
    // run the loop of call `ci` as far as the results at hand reach
    void advance(int ci) {
        if (calls_[(size_t)ci].state != 2) return;
        for (;;) {
            NsCall& c = calls_[(size_t)ci];                          // (a deque: spawn() does not move it, but keep it local)
            if (c.again >= 0) {
                NsCall& r = calls_[(size_t)c.again];
                if (r.state != 3) return;
                c.score = -r.ret; c.nodes += r.nodes;                // score = -negaScout(.., -beta, -cur)
                release(c.again); c.again = -1;
            } else {
                if (c.i == c.n_kids) { finish(ci); return; }         // (a node without children)
                if (c.i == 0) {
                    if (c.first[0] < 0) { const int k = spawn(ci, 0, -c.n, -c.alpha); calls_[(size_t)ci].first[0] = k; }
                } else {
                    if (c.spec_alpha != c.alpha) {
                        for (int j = c.i; j < c.n_kids; j++) if (c.first[j] >= 0) { drop(c.first[j]); c.first[j] = -1; }
                        c.spec_alpha = c.alpha;
                    }
                    for (int j = c.i; j < c.n_kids; j++)
                        if (calls_[(size_t)ci].first[j] < 0) { const int k = spawn(ci, j, -c.alpha - 1, -c.alpha); calls_[(size_t)ci].first[j] = k; }
                }
                NsCall& r = calls_[(size_t)c.first[c.i]];
                if (r.state != 3) return;
                const int cur = -r.ret;
                c.nodes += r.nodes;
                release(c.first[c.i]); c.first[c.i] = -1;
                if (cur > c.score) {
                    if (c.n == c.beta || (c.ply > 0 && c.ply == max_depth_ - 2)) c.score = cur;
                    else { const int k = spawn(ci, c.i, -c.beta, -cur); calls_[(size_t)ci].again = k; continue; }
                }
            }
            if (c.score > c.alpha) c.alpha = c.score;
            if (c.ply == 0) {                                        // moves.SetMove(i, j, alpha), movekeeper.go:26-36
                NsRootOut& ro = roots_[(size_t)c.pos];
                const int p = list_[c.pos], cell = c.cell[c.i];
                values_[p][cell] = c.alpha;
                if (visit_order_) visit_order_[p][ro.n_visited] = (int8_t)cell;
                ro.n_visited++;
                if (c.alpha >= ro.keeper_max) {
                    if (c.alpha > ro.keeper_max) { ro.keeper_max = c.alpha; ro.keeper_mask = 0; ro.first_best = cell; }
                    ro.keeper_mask |= 1u << cell;
                }
            }
            c.i++;
            if (c.alpha >= c.beta || c.i == c.n_kids) { finish(ci); return; }
            c.n = c.alpha + 1;
        }
    }
    void finish(int ci) {
        NsCall& c = calls_[(size_t)ci];
        for (int j = 0; j < c.n_kids; j++) if (c.first[j] >= 0) { drop(c.first[j]); c.first[j] = -1; }
        c.state = 3; c.ret = c.score;
        if (c.ply > 0) c.nodes += 1;                                 // the call itself (its parent's loop counts it)
        if (c.parent >= 0) ready_.push_back(c.parent);
    }

    Device& d_; Ctx& c_;
    const sq_board* h_pos_; const sq_board* d_pos_; const int* list_; int64_t n_list_; int max_depth_;
    sq::NsScores sc_;
    int32_t (*values_)[25]; int8_t (*visit_order_)[25];
    int split_wide_ = 1, split_null_ = 1, deepest_loop_ = 0, n_flights_ = 8;
    uint32_t item_budget_ = 0, item_budget_wide_ = 0;
    std::chrono::steady_clock::time_point t_run_;
    std::deque<NsCall> calls_;
    std::vector<int> free_, ready_, stack_, root_call_;
    std::vector<std::pair<int32_t, uint32_t>> outbox_;
    std::vector<uint8_t> order_max_, order_min_;
    std::vector<NsRootOut> roots_;
    uint64_t rounds_ = 0, items_ = 0, dropped_ = 0, expanded_ = 0;
};

}  // namespace

extern "C" {

int sq_abi_version(void) { return SQ_ABI_VERSION; }

const char* sq_strerror(int status) {
    switch (status) {
    case SQ_OK: return "ok";
    case SQ_ERR_INVALID: return "invalid argument";
    case SQ_ERR_NOT_INIT: return "library not initialised (sq_init)";
    case SQ_ERR_CUDA: return "CUDA runtime error";
    case SQ_ERR_NO_DEVICE: return "no usable sm_100 CUDA device";
    case SQ_ERR_UNSUPPORTED: return "unsupported combination";
    case SQ_ERR_NOMEM: return "out of memory";
    default: return "unknown status";
    }
}
const char* sq_last_error(void) { return t_err; }
int sq_last_error_copy(char* buf, int len) {
    if (!buf || len <= 0) return SQ_ERR_INVALID;
    snprintf(buf, (size_t)len, "%s", t_err);
    return SQ_OK;
}

int sq_init(const int* device_ids, int n_devices, uint64_t seed) {
    std::unique_lock<std::shared_mutex> lk(g_state);
    destroy_all();
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        snprintf(t_err, sizeof t_err, "cudaGetDeviceCount: %s (count %d)", cudaGetErrorString(e), count);
        cudaGetLastError();
        return SQ_ERR_NO_DEVICE;
    }
    int one = 0;
    if (!device_ids || n_devices <= 0) { device_ids = &one; n_devices = 1; }
    sq::ScanTables tables;
    sq::build_scan_tables(tables);
    sq::AbTables ab_tables;
    sq::build_ab_tables(ab_tables);
    sq::NsTables ns_tables;
    sq::build_ns_tables(ns_tables);
    sq::NsCellTables ns_cell_tables;
    sq::build_ns_cell_tables(ns_tables, ns_cell_tables);
    g_ns_tables = ns_tables;
    for (int k = 0; k < n_devices; k++) {
        if (device_ids[k] < 0 || device_ids[k] >= count) {
            destroy_all();
            snprintf(t_err, sizeof t_err, "device id %d out of range", device_ids[k]);
            return SQ_ERR_INVALID;
        }
        // the device joins the table BEFORE anything fallible is created on it: whatever a failed
        // start-up leaves behind is released by destroy_all()
        g_dev.emplace_back(new Device());
        g_dev.back()->id = device_ids[k];
        int rc = init_device(*g_dev.back(), tables, ab_tables, ns_tables, ns_cell_tables);
        if (rc) { destroy_all(); return rc; }
    }
    g_seed = seed;
    g_init = true;
    return SQ_OK;
}

void sq_shutdown(void) {
    std::unique_lock<std::shared_mutex> lk(g_state);
    destroy_all();
}

int sq_device_count(void) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    return g_init ? (int)g_dev.size() : 0;
}
int sq_sm_count(int slot) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    if (check_slot(slot)) return 0;
    return g_dev[slot]->sm_count;
}
int sq_device_id(int slot) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    const int rc = check_slot(slot);
    if (rc) return rc;
    return g_dev[slot]->id;
}
uint64_t sq_launch_count(void) { return g_launches.load(); }

// ---- device-resident forms -------------------------------------------------
int sq_classify_dev(int slot, const sq_board* boards, int64_t n, uint8_t* flags, int8_t* wm,
                    int8_t* wa, void* stream) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    int rc = check_slot(slot);
    if (rc) return rc;
    if (n < 0 || (n > 0 && !boards)) { snprintf(t_err, sizeof t_err, "sq_classify_dev: null boards"); return SQ_ERR_INVALID; }
    Device& d = *g_dev[slot];
    CK(cudaSetDevice(d.id));
    return launch_classify(d, boards, n, flags, wm, wa, (cudaStream_t)stream);
}

int sq_playouts_dev(int slot, const sq_board* leaves, const int8_t* to_move, int64_t n_leaves,
                    uint64_t leaf_index_base, uint64_t per_leaf, uint64_t stream_id, uint64_t* out,
                    void* stream) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    int rc = check_slot(slot);
    if (rc) return rc;
    if (n_leaves < 0 || (n_leaves > 0 && (!leaves || !to_move || !out))) {
        snprintf(t_err, sizeof t_err, "sq_playouts_dev: null pointer"); return SQ_ERR_INVALID;
    }
    Device& d = *g_dev[slot];
    CtxHold c = acquire(d);                    // for its counter ring and timing events only
    CK(cudaSetDevice(d.id));
    return launch_playouts(d, *c, (int)(c.c - d.ctx), leaves, to_move, n_leaves, leaf_index_base, per_leaf, stream_id, out,
                           (cudaStream_t)stream);
}

// ---- host-buffer forms -------------------------------------------------------
int sq_classify(const sq_board* boards, int64_t n, uint8_t* flags, int8_t* wm, int8_t* wa) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    int rc = check_ready();
    if (rc) return rc;
    if (n < 0 || (n > 0 && !boards)) { snprintf(t_err, sizeof t_err, "sq_classify: null boards"); return SQ_ERR_INVALID; }
    int m = (int)g_dev.size();
    std::vector<CtxHold> held((size_t)m);
    for (int k = 0; k < m; k++) {
        Device& d = *g_dev[k];
        int64_t lo, hi; share(n, m, k, lo, hi);
        int64_t cnt = hi - lo;
        if (cnt <= 0) continue;
        held[k] = acquire(d);
        Ctx& c = *held[k];
        CK(cudaSetDevice(d.id));
        if (cnt <= kSmallLeaves) {          // small call: pinned staging, one copy each way
            memcpy(c.pin, boards + lo, cnt * sizeof(sq_board));
            uint8_t* din = c.small + 16 + (size_t)cnt * 32;
            uint8_t* dres = c.small + 16;
            CK(cudaMemcpyAsync(din, c.pin, cnt * sizeof(sq_board), cudaMemcpyHostToDevice, c.stream));
            rc = launch_classify(d, (const sq_board*)din, cnt, dres, (int8_t*)dres + cnt, (int8_t*)dres + 2 * cnt, c.stream);
            if (rc) return rc;
            CK(cudaMemcpyAsync(c.pin + (size_t)kSmallLeaves * 9, dres, cnt * 3, cudaMemcpyDeviceToHost, c.stream));
            continue;
        }
        if ((rc = reserve(c, 0, cnt * sizeof(sq_board)))) return rc;
        if ((rc = reserve(c, 1, cnt * 3))) return rc;
        uint8_t* dout = (uint8_t*)c.buf[1];
        CK(cudaMemcpyAsync(c.buf[0], boards + lo, cnt * sizeof(sq_board), cudaMemcpyHostToDevice, c.stream));
        rc = launch_classify(d, (const sq_board*)c.buf[0], cnt, dout, (int8_t*)dout + cnt, (int8_t*)dout + 2 * cnt, c.stream);
        if (rc) return rc;
    }
    // second sweep: a device-to-host copy into pageable memory waits for its stream, so all
    // devices are launched before the first result is collected
    for (int k = 0; k < m; k++) {
        Device& d = *g_dev[k];
        int64_t lo, hi; share(n, m, k, lo, hi);
        int64_t cnt = hi - lo;
        if (cnt <= 0) continue;
        Ctx& c = *held[k];
        CK(cudaSetDevice(d.id));
        if (cnt <= kSmallLeaves) {
            CK(cudaStreamSynchronize(c.stream));
            const uint8_t* h = c.pin + (size_t)kSmallLeaves * 9;
            if (flags) memcpy(flags + lo, h, cnt);
            if (wm) memcpy(wm + lo, h + cnt, cnt);
            if (wa) memcpy(wa + lo, h + 2 * cnt, cnt);
            continue;
        }
        uint8_t* dout = (uint8_t*)c.buf[1];
        if (flags) CK(cudaMemcpyAsync(flags + lo, dout, cnt, cudaMemcpyDeviceToHost, c.stream));
        if (wm) CK(cudaMemcpyAsync(wm + lo, dout + cnt, cnt, cudaMemcpyDeviceToHost, c.stream));
        if (wa) CK(cudaMemcpyAsync(wa + lo, dout + 2 * cnt, cnt, cudaMemcpyDeviceToHost, c.stream));
    }
    for (int k = 0; k < m; k++)
        if (held[k].c) { CK(cudaSetDevice(g_dev[k]->id)); CK(cudaStreamSynchronize(held[k]->stream)); }
    return SQ_OK;
}

// shared body of sq_playouts (out: u64[n][4]) and sq_playouts_packed (packed: u8[n], one playout per leaf)
static int playouts_host(const char* who, const sq_board* leaves, const int8_t* to_move, int64_t n_leaves,
                         uint64_t per_leaf, uint64_t stream_id, uint64_t (*out)[4], uint8_t* packed, int only_slot = -1) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    int rc = only_slot < 0 ? check_ready() : check_slot(only_slot);
    if (rc) return rc;
    if (n_leaves < 0 || (n_leaves > 0 && (!leaves || !to_move || (!out && !packed)))) {
        snprintf(t_err, sizeof t_err, "%s: null pointer", who); return SQ_ERR_INVALID;
    }
    for (int64_t i = 0; i < n_leaves; i++) {
        if (bad_board(leaves[i]) || (to_move[i] != 1 && to_move[i] != -1)) {
            snprintf(t_err, sizeof t_err, "%s: leaf %lld is not a board / to_move not +-1", who, (long long)i);
            return SQ_ERR_INVALID;
        }
    }
    const size_t res_bytes = packed ? 1 : 32;           // per leaf, device -> host
    int m = (int)g_dev.size();
    // only_slot >= 0: the whole call on that one device (root-parallel hosts keep one tree per device)
    auto part = [&](int k, int64_t& lo, int64_t& hi) {
        if (only_slot < 0) share(n_leaves, m, k, lo, hi);
        else { lo = 0; hi = k == only_slot ? n_leaves : 0; }
    };
    std::vector<CtxHold> held((size_t)m);
    for (int k = 0; k < m; k++) {
        Device& d = *g_dev[k];
        int64_t lo, hi; part(k, lo, hi);
        int64_t cnt = hi - lo;
        if (cnt <= 0) continue;
        held[k] = acquire(d);
        Ctx& c = *held[k];
        const int ci = (int)(held[k].c - d.ctx);
        CK(cudaSetDevice(d.id));
        if (cnt <= kSmallLeaves) {
            // small call: pinned staging, one copy in, one memset, one copy out
            uint8_t* hin = c.pin;                              // [boards | to_move]
            memcpy(hin, leaves + lo, cnt * sizeof(sq_board));
            memcpy(hin + cnt * sizeof(sq_board), to_move + lo, cnt);
            unsigned long long* dcounter = (unsigned long long*)c.small;
            uint8_t* dres = c.small + 16;
            uint8_t* din = c.small + 16 + (size_t)cnt * 32;
            CK(cudaMemcpyAsync(din, hin, cnt * 9, cudaMemcpyHostToDevice, c.stream));
            CK(cudaMemsetAsync(c.small, 0, 16 + (size_t)cnt * res_bytes, c.stream));
            rc = launch_playouts(d, c, ci, (const sq_board*)din, (const int8_t*)(din + cnt * sizeof(sq_board)), cnt, (uint64_t)lo,
                                 per_leaf, stream_id, packed ? nullptr : (uint64_t*)dres, c.stream, dcounter, packed ? dres : nullptr);
            if (rc) return rc;
            CK(cudaMemcpyAsync(c.pin + (size_t)kSmallLeaves * 9, dres, cnt * res_bytes, cudaMemcpyDeviceToHost, c.stream));
            continue;
        }
        if ((rc = reserve(c, 0, cnt * sizeof(sq_board)))) return rc;
        if ((rc = reserve(c, 1, cnt))) return rc;
        if ((rc = reserve(c, 2, cnt * res_bytes))) return rc;
        CK(cudaMemcpyAsync(c.buf[0], leaves + lo, cnt * sizeof(sq_board), cudaMemcpyHostToDevice, c.stream));
        CK(cudaMemcpyAsync(c.buf[1], to_move + lo, cnt, cudaMemcpyHostToDevice, c.stream));
        if (packed) CK(cudaMemsetAsync(c.buf[2], 0, cnt, c.stream));
        rc = launch_playouts(d, c, ci, (const sq_board*)c.buf[0], (const int8_t*)c.buf[1], cnt, (uint64_t)lo,
                             per_leaf, stream_id, packed ? nullptr : (uint64_t*)c.buf[2], c.stream, nullptr,
                             packed ? (uint8_t*)c.buf[2] : nullptr);
        if (rc) return rc;
    }
    for (int k = 0; k < m; k++) {      // collect after every device has been launched
        Device& d = *g_dev[k];
        int64_t lo, hi; part(k, lo, hi);
        int64_t cnt = hi - lo;
        if (cnt <= 0) continue;
        Ctx& c = *held[k];
        uint8_t* dst = packed ? packed + lo : (uint8_t*)(out + lo);
        CK(cudaSetDevice(d.id));
        if (cnt <= kSmallLeaves) {
            CK(cudaStreamSynchronize(c.stream));
            memcpy(dst, c.pin + (size_t)kSmallLeaves * 9, cnt * res_bytes);
            continue;
        }
        CK(cudaMemcpyAsync(dst, c.buf[2], cnt * res_bytes, cudaMemcpyDeviceToHost, c.stream));
    }
    for (int k = 0; k < m; k++)
        if (held[k].c) { CK(cudaSetDevice(g_dev[k]->id)); CK(cudaStreamSynchronize(held[k]->stream)); }
    return SQ_OK;
}

int sq_playouts(const sq_board* leaves, const int8_t* to_move, int64_t n_leaves,
                uint64_t per_leaf, uint64_t stream_id, uint64_t (*out)[4]) {
    if (n_leaves > 0 && !out) { snprintf(t_err, sizeof t_err, "sq_playouts: null pointer"); return SQ_ERR_INVALID; }
    return playouts_host("sq_playouts", leaves, to_move, n_leaves, per_leaf, stream_id, out, nullptr);
}

int sq_playouts_packed(const sq_board* leaves, const int8_t* to_move, int64_t n_leaves,
                       uint64_t stream_id, uint8_t* outcome) {
    if (n_leaves > 0 && !outcome) { snprintf(t_err, sizeof t_err, "sq_playouts_packed: null pointer"); return SQ_ERR_INVALID; }
    return playouts_host("sq_playouts_packed", leaves, to_move, n_leaves, 1, stream_id, nullptr, outcome);
}

int sq_playouts_on(int slot, const sq_board* leaves, const int8_t* to_move, int64_t n_leaves,
                   uint64_t per_leaf, uint64_t stream_id, uint64_t (*out)[4]) {
    if (slot < 0 || (n_leaves > 0 && !out)) { snprintf(t_err, sizeof t_err, "sq_playouts_on: bad slot / null pointer"); return SQ_ERR_INVALID; }
    return playouts_host("sq_playouts_on", leaves, to_move, n_leaves, per_leaf, stream_id, out, nullptr, slot);
}

int sq_playouts_packed_on(int slot, const sq_board* leaves, const int8_t* to_move, int64_t n_leaves,
                          uint64_t stream_id, uint8_t* outcome) {
    if (slot < 0 || (n_leaves > 0 && !outcome)) { snprintf(t_err, sizeof t_err, "sq_playouts_packed_on: bad slot / null pointer"); return SQ_ERR_INVALID; }
    return playouts_host("sq_playouts_packed_on", leaves, to_move, n_leaves, 1, stream_id, nullptr, outcome, slot);
}

int sq_alphabeta_root(const sq_board* positions, int64_t n, int dialect, int max_depth,
                      const int8_t scores[25], int32_t (*values)[25], int32_t* best_value,
                      uint32_t* best_mask, uint64_t* nodes) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    int rc = check_ready();
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!positions || !values)) || !scores) {
        snprintf(t_err, sizeof t_err, "sq_alphabeta_root: null pointer"); return SQ_ERR_INVALID;
    }
    if (dialect != SQ_AB_ALPHABETA && dialect != SQ_AB_SQUAVATHR && dialect != SQ_AB_ALPHABETA_AVOID &&
        dialect != SQ_AB_ABBOOK) {
        snprintf(t_err, sizeof t_err, "sq_alphabeta_root: dialect %d has no root-move search in the reference", dialect);
        return SQ_ERR_UNSUPPORTED;
    }
    if (max_depth < 0 || max_depth > sq::kAbMaxDepth) { snprintf(t_err, sizeof t_err, "max_depth out of range [0,%d]", sq::kAbMaxDepth); return SQ_ERR_INVALID; }
    for (int64_t i = 0; i < n; i++)
        if (bad_board(positions[i])) {
            snprintf(t_err, sizeof t_err, "sq_alphabeta_root: position %lld is not a board", (long long)i);
            return SQ_ERR_INVALID;
        }
    const int m = (int)g_dev.size();
    rc = for_each_device(m, [&](int k) -> int {
        Device& d = *g_dev[k];
        int64_t lo, hi; share(n, m, k, lo, hi);
        int64_t cnt = hi - lo;
        if (cnt <= 0) return SQ_OK;
        int r;
        CtxHold c = acquire(d);
        CK(cudaSetDevice(d.id));
        if ((r = reserve(*c, 0, cnt * sizeof(sq_board)))) return r;
        if ((r = reserve(*c, 2, cnt * 25 * sizeof(int32_t)))) return r;
        if ((r = reserve(*c, 3, cnt * sizeof(unsigned long long)))) return r;
        CK(cudaMemcpyAsync(c->buf[0], positions + lo, cnt * sizeof(sq_board), cudaMemcpyHostToDevice, c->stream));
        uint64_t launches = 0;
        cudaError_t e = sq::ab_root_search(positions + lo, (const sq_board*)c->buf[0], cnt, dialect, max_depth, scores,
                                           (int32_t*)c->buf[2], (unsigned long long*)c->buf[3],
                                           d.sm_count * d.ab_blocks_per_sm, c->stream, &launches);
        g_launches += launches;
        if (e != cudaSuccess) return fail_cuda(e, "ab_root_search");
        CK(cudaMemcpyAsync(values + lo, c->buf[2], cnt * 25 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        if (nodes) CK(cudaMemcpyAsync(nodes + lo, c->buf[3], cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return SQ_OK;
    });
    if (rc) return rc;
    // movekeeper.go:26-52 over each finished row (a 25-entry host reduction)
    for (int64_t i = 0; i < n; i++) {
        int mx = -20000; uint32_t mask = 0;
        for (int c = 0; c < 25; c++) {
            int32_t v = values[i][c];
            if (v == SQ_NO_MOVE) continue;
            if (v >= mx) { if (v > mx) { mx = v; mask = 0; } mask |= 1u << c; }
        }
        if (best_mask) best_mask[i] = mask;
        if (best_value) best_value[i] = mask ? mx : 0;
    }
    return SQ_OK;
}

int sq_negascout_root(const sq_board* positions, int64_t n, int max_depth, const int8_t scores[25],
                      int32_t (*values)[25], int8_t (*visit_order)[25], int32_t* best_value,
                      uint32_t* best_mask, int8_t* first_best, uint64_t* nodes) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    int rc = check_ready();
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!positions || !values)) || !scores) {
        snprintf(t_err, sizeof t_err, "sq_negascout_root: null pointer"); return SQ_ERR_INVALID;
    }
    if (max_depth < 0 || max_depth > sq::kNsMaxDepth) { snprintf(t_err, sizeof t_err, "max_depth out of range [0,%d]", sq::kNsMaxDepth); return SQ_ERR_INVALID; }
    for (int64_t i = 0; i < n; i++)
        if (bad_board(positions[i])) {
            snprintf(t_err, sizeof t_err, "sq_negascout_root: position %lld is not a board", (long long)i);
            return SQ_ERR_INVALID;
        }
    sq::NsScores sc;
    memcpy(sc.s, scores, 25);
    const int m = (int)g_dev.size();
    return for_each_device(m, [&](int k) -> int {
        Device& d = *g_dev[k];
        int64_t lo, hi; share(n, m, k, lo, hi);
        const int64_t cnt = hi - lo;
        if (cnt <= 0) return SQ_OK;
        int r;
        CtxHold c = acquire(d);
        CK(cudaSetDevice(d.id));
        // one device block: [queue | nodes | values | best_value | best_mask | visit_order | first_best]
        const size_t o_nodes = 16, o_values = o_nodes + 8 * cnt, o_bv = o_values + 100 * cnt, o_bm = o_bv + 4 * cnt,
                     o_visit = o_bm + 4 * cnt, o_first = o_visit + 25 * cnt, total = o_first + cnt;
        if ((r = reserve(*c, 0, cnt * sizeof(sq_board)))) return r;
        if ((r = reserve(*c, 4, total))) return r;
        uint8_t* blk = (uint8_t*)c->buf[4];
        CK(cudaMemcpyAsync(c->buf[0], positions + lo, cnt * sizeof(sq_board), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemsetAsync(blk, 0, 16, c->stream));
        sq::NsOut out;
        out.nodes = (unsigned long long*)(blk + o_nodes);
        out.values = (int32_t(*)[25])(blk + o_values);
        out.best_value = (int32_t*)(blk + o_bv);
        out.best_mask = (uint32_t*)(blk + o_bm);
        out.visit_order = (int8_t(*)[25])(blk + o_visit);
        out.first_best = (int8_t*)(blk + o_first);
        // positions the fast kernel cannot take (a complete line already on the board, or maxDepth 0) replay the
        // recursion with the full evaluator; the rest take the incremental path.  Both read an index list.
        std::vector<int> order((size_t)cnt);
        int64_t n_fast = 0, n_slow = 0;
        for (int64_t i = 0; i < cnt; i++) {
            const sq_board& b = positions[lo + i];
            bool line = false;
            for (int q = 0; q < 28; q++) line |= !(g_ns_tables.quad[q] & ~b.x) || !(g_ns_tables.quad[q] & ~b.o);
            for (int t = 0; t < 48; t++) line |= !(g_ns_tables.trip[t] & ~b.x) || !(g_ns_tables.trip[t] & ~b.o);
            if (line || max_depth < 1) order[(size_t)(cnt - 1 - n_slow++)] = (int)i; else order[(size_t)n_fast++] = (int)i;
        }
        // warps pull positions in list order: start the (probably) biggest searches first -- more empty cells,
        // bigger tree -- so that the batch does not end on one late, large search
        std::stable_sort(order.begin(), order.begin() + n_fast, [&](int a, int b) {
            return __builtin_popcount(positions[lo + a].x | positions[lo + a].o) < __builtin_popcount(positions[lo + b].x | positions[lo + b].o);
        });
        if ((r = reserve(*c, 5, cnt * sizeof(int)))) return r;
        CK(cudaMemcpyAsync(c->buf[5], order.data(), cnt * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        const int64_t cap = (int64_t)d.sm_count * 6;
        auto grid = [&](int64_t units) { int64_t want = (units + sq::kNsWarps - 1) / sq::kNsWarps; return (int)(want < cap ? want : cap); };
        // One warp replays the whole ChooseMove of a position -- up to `budget` staticValue calls; a position that needs more
        // is given up by its warp and finished by NsSplit over many warps, so that a batch does not end on a few long
        // searches.  Short lists go to NsSplit at once (the latency case: one position per call).  SQ_NS_SPLIT=0: no budget.
        const char* e_split = getenv("SQ_NS_SPLIT");
        const bool split = n_fast > 0 && max_depth >= 4 && !(e_split && e_split[0] == '0');
        const char* e_budget = getenv("SQ_NS_BUDGET");
        const unsigned long long budget = !split ? ~0ull : e_budget ? strtoull(e_budget, nullptr, 10) : (1ull << 17);
        const char* e_direct = getenv("SQ_NS_DIRECT");
        const bool direct = split && n_fast <= (e_direct ? atoi(e_direct) : 16);
        if (n_fast && !direct) {
            sq::negascout_fast_kernel<false><<<grid(n_fast), sq::kNsWarps * 32, 0, c->stream>>>(
                (const sq_board*)c->buf[0], (const int*)c->buf[5], (int)n_fast, max_depth, sc, out, (unsigned int*)blk, nullptr, nullptr,
                budget, nullptr, nullptr);
            g_launches += 1;
            CK(cudaGetLastError());
        }
        if (n_slow) {
            sq::negascout_kernel<<<grid(n_slow), sq::kNsWarps * 32, 0, c->stream>>>(
                (const sq_board*)c->buf[0], (const int*)c->buf[5] + n_fast, (int)n_slow, max_depth, sc, out, (unsigned int*)blk + 1);
            g_launches += 1;
            CK(cudaGetLastError());
        }
        std::vector<uint8_t> h(total);
        const auto t_pass = std::chrono::steady_clock::now();
        CK(cudaMemcpyAsync(h.data(), blk, total, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (getenv("SQ_NS_TRACE") && n_fast && !direct)
            fprintf(stderr, "ns one warp per position: %lld positions, budget %llu, %.4f s\n", (long long)n_fast, budget,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t_pass).count());
        memcpy(values + lo, h.data() + o_values, 100 * cnt);
        if (nodes) memcpy(nodes + lo, h.data() + o_nodes, 8 * cnt);
        if (best_value) memcpy(best_value + lo, h.data() + o_bv, 4 * cnt);
        if (best_mask) memcpy(best_mask + lo, h.data() + o_bm, 4 * cnt);
        if (visit_order) memcpy(visit_order + lo, h.data() + o_visit, 25 * cnt);
        if (first_best) memcpy(first_best + lo, h.data() + o_first, cnt);
        if (split) {
            std::vector<int> hard;
            const unsigned long long* got = (const unsigned long long*)(h.data() + o_nodes);
            for (int64_t q = 0; q < n_fast; q++) if (direct || got[order[(size_t)q]] == sq::kNsAborted) hard.push_back(order[(size_t)q]);
            if (!hard.empty()) {
                NsSplit driver(d, *c, positions + lo, (const sq_board*)c->buf[0], hard.data(), (int64_t)hard.size(), max_depth, sc,
                               values + lo, visit_order ? visit_order + lo : nullptr);
                return driver.run(best_value ? best_value + lo : nullptr, best_mask ? best_mask + lo : nullptr,
                                  first_best ? first_best + lo : nullptr, nodes ? nodes + lo : nullptr);
            }
        }
        return SQ_OK;
    });
}

int sq_alphabeta_subtrees(const sq_subtree* subtrees, int64_t n, int dialect, int max_depth,
                          const int8_t scores[25], int32_t* value, uint64_t* nodes) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    int rc = check_ready();
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!subtrees || !value)) || !scores) {
        snprintf(t_err, sizeof t_err, "sq_alphabeta_subtrees: null pointer"); return SQ_ERR_INVALID;
    }
    if (dialect < 1 || dialect > 6) { snprintf(t_err, sizeof t_err, "unknown dialect %d", dialect); return SQ_ERR_INVALID; }
    if (max_depth < 0 || max_depth > sq::kAbMaxDepth) { snprintf(t_err, sizeof t_err, "max_depth out of range [0,%d]", sq::kAbMaxDepth); return SQ_ERR_INVALID; }
    for (int64_t i = 0; i < n; i++) {
        const sq_subtree& s = subtrees[i];
        // the search keeps running sums in 16 bits: 26 evaluations of at most 240 + 127 + 300 leave this much
        if (s.carried > 15000 || s.carried < -15000) {
            snprintf(t_err, sizeof t_err, "sq_alphabeta_subtrees: subtree %lld: |carried| > 15000", (long long)i);
            return SQ_ERR_UNSUPPORTED;
        }
        bool ok = !bad_board(s.b) && s.last_cell >= 0 && s.last_cell < 25 &&
                  (((s.b.x | s.b.o) >> s.last_cell) & 1u) && (s.side_to_move == 1 || s.side_to_move == -1) &&
                  s.ply >= 0 && s.ply <= max_depth;
        if (!ok) { snprintf(t_err, sizeof t_err, "sq_alphabeta_subtrees: subtree %lld malformed", (long long)i); return SQ_ERR_INVALID; }
    }
    // one device's share: the subtrees ix[] searched to `depth`, results scattered back by index
    auto run_on = [&](int k, const std::vector<int64_t>& ix, int depth, int32_t* val_out, uint64_t* nodes_out) -> int {
        Device& d = *g_dev[k];
        const int64_t cnt = (int64_t)ix.size();
        if (cnt <= 0) return SQ_OK;
        int r;
        std::vector<sq_subtree> st((size_t)cnt);
        for (int64_t q = 0; q < cnt; q++) st[(size_t)q] = subtrees[ix[(size_t)q]];
        CtxHold c = acquire(d);
        CK(cudaSetDevice(d.id));
        if ((r = reserve(*c, 0, cnt * sizeof(sq_subtree)))) return r;
        if ((r = reserve(*c, 2, cnt * sizeof(int32_t)))) return r;
        if ((r = reserve(*c, 3, cnt * sizeof(unsigned long long)))) return r;
        CK(cudaMemcpyAsync(c->buf[0], st.data(), cnt * sizeof(sq_subtree), cudaMemcpyHostToDevice, c->stream));
        uint64_t launches = 0;
        cudaError_t e = sq::ab_subtree_search(st.data(), (const sq_subtree*)c->buf[0], cnt, dialect, depth, scores,
                                              (int32_t*)c->buf[2], (unsigned long long*)c->buf[3],
                                              d.sm_count * d.ab_blocks_per_sm, c->stream, &launches);
        g_launches += launches;
        if (e != cudaSuccess) return fail_cuda(e, "ab_subtree_search");
        std::vector<int32_t> hv((size_t)cnt); std::vector<uint64_t> hn((size_t)cnt);
        CK(cudaMemcpyAsync(hv.data(), c->buf[2], cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(hn.data(), c->buf[3], cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int64_t q = 0; q < cnt; q++) {
            if (val_out) val_out[ix[(size_t)q]] = hv[(size_t)q];
            if (nodes_out) nodes_out[ix[(size_t)q]] = hn[(size_t)q];
        }
        return SQ_OK;
    };
    const int m = (int)g_dev.size();
    std::vector<int64_t> all((size_t)n);
    for (int64_t i = 0; i < n; i++) all[(size_t)i] = i;
    if (m == 1 || n <= 1) return run_on(0, all, max_depth, value, nodes);
    // Several devices: whole subtrees are dealt out by SIZE, not by index -- opening3's 85 subtrees differ by
    // orders of magnitude although they all have 23 empty cells.  The size estimate is a PILOT search of
    // every subtree three plies shallower on the first device (a few hundred times cheaper than the search
    // itself): its evaluator counts are the weights; largest first, each to the least loaded device.
    std::vector<double> est((size_t)n, 1.0);
    int min_rem = 1 << 30;
    for (int64_t i = 0; i < n; i++) min_rem = std::min(min_rem, max_depth - (int)subtrees[i].ply);
    if (min_rem > 4) {
        std::vector<uint64_t> pilot((size_t)n, 0);
        rc = run_on(0, all, max_depth - 3, nullptr, pilot.data());
        if (rc) return rc;
        for (int64_t i = 0; i < n; i++) est[(size_t)i] = 1.0 + (double)pilot[(size_t)i];
    } else {
        for (int64_t i = 0; i < n; i++) {
            const int empt = 25 - __builtin_popcount(subtrees[i].b.x | subtrees[i].b.o);
            const int rem = std::min(max_depth - (int)subtrees[i].ply, empt);
            double e = 1; for (int k = 0; k < rem; k++) e *= 0.45 * (empt - k) + 1;
            est[(size_t)i] = e;
        }
    }
    std::vector<std::vector<int64_t>> mine((size_t)m);
    std::vector<double> load((size_t)m, 0.0);
    std::stable_sort(all.begin(), all.end(), [&](int64_t a, int64_t b) { return est[(size_t)a] > est[(size_t)b]; });
    for (int64_t i : all) {
        int best = 0;
        for (int k = 1; k < m; k++) if (load[(size_t)k] < load[(size_t)best]) best = k;
        mine[(size_t)best].push_back(i); load[(size_t)best] += est[(size_t)i];
    }
    return for_each_device(m, [&](int k) -> int { return run_on(k, mine[(size_t)k], max_depth, value, nodes); });
}

// ---- measurement helpers -------------------------------------------------------
int sq_int32_peak(int slot, int kind, int reps, double* ops_per_s) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    int rc = check_slot(slot);
    if (rc) return rc;
    if (!ops_per_s || kind < 0 || kind > 2) return SQ_ERR_INVALID;
    Device& d = *g_dev[slot];
    CtxHold c = acquire(d);
    CK(cudaSetDevice(d.id));
    if ((rc = reserve(*c, 4, 256))) return rc;
    const int iters = 4096, threads = 1024;
    const int blocks = d.sm_count * 2;
    double best = 0;
    if (reps < 1) reps = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int r = 0; r < reps + 1; r++) {
        CK(cudaEventRecord(e0, c->stream));
        uint32_t* o = (uint32_t*)c->buf[4];
        if (kind == 0) sq::int32_peak_kernel<0><<<blocks, threads, 0, c->stream>>>(o, 12345u + r, iters);
        else if (kind == 1) sq::int32_peak_kernel<1><<<blocks, threads, 0, c->stream>>>(o, 12345u + r, iters);
        else sq::int32_peak_kernel<2><<<blocks, threads, 0, c->stream>>>(o, 12345u + r, iters);
        g_launches++;
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, c->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double ops = (double)blocks * threads * (double)iters * 16.0 * 8.0;
        if (r > 0 && ms > 0 && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ops_per_s = best;
    return SQ_OK;
}

int sq_last_playout_kernel_ms(int slot, float* ms) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    int rc = check_slot(slot);
    if (rc) return rc;
    if (!ms) return SQ_ERR_INVALID;
    Device& d = *g_dev[slot];
    const int ci = d.last_timed.load();
    if (ci < 0) { snprintf(t_err, sizeof t_err, "no playout kernel has run on slot %d", slot); return SQ_ERR_INVALID; }
    Ctx& c = d.ctx[ci];
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.timed) { snprintf(t_err, sizeof t_err, "no playout kernel has run on slot %d", slot); return SQ_ERR_INVALID; }
    CK(cudaSetDevice(d.id));
    CK(cudaEventSynchronize(c.ev1));
    CK(cudaEventElapsedTime(ms, c.ev0, c.ev1));
    return SQ_OK;
}

int sq_abi_checked(void) { return SQ_CHECKED; }

// Checked build only: {failures, code of the first, its two arguments} summed / taken over the devices, then reset.
int sq_debug_check_failures(uint32_t out[4]) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    int rc = check_ready();
    if (rc) return rc;
    if (!out) return SQ_ERR_INVALID;
    out[0] = out[1] = out[2] = out[3] = 0;
#if SQ_CHECKED
    for (auto& d : g_dev) {
        CK(cudaSetDevice(d->id));
        CK(cudaDeviceSynchronize());
        unsigned int h[4] = {0, 0, 0, 0}, zero[4] = {0, 0, 0, 0};
        CK(cudaMemcpyFromSymbol(h, sq::g_sq_check_fail, sizeof h));
        CK(cudaMemcpyToSymbol(sq::g_sq_check_fail, zero, sizeof zero));
        if (h[0] && !out[0]) { out[1] = h[1]; out[2] = h[2]; out[3] = h[3]; }
        out[0] += h[0];
    }
#endif
    return SQ_OK;
}

int sq_debug_philox(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
    std::shared_lock<std::shared_mutex> sl(g_state);
    int rc = check_ready();
    if (rc) return rc;
    if (!counter || !key || !out) return SQ_ERR_INVALID;
    Device& d = *g_dev[0];
    CtxHold c = acquire(d);
    CK(cudaSetDevice(d.id));
    if ((rc = reserve(*c, 4, 256))) return rc;
    uint32_t* b = (uint32_t*)c->buf[4];
    CK(cudaMemcpyAsync(b, counter, 16, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(b + 4, key, 8, cudaMemcpyHostToDevice, c->stream));
    sq::philox_debug_kernel<<<1, 1, 0, c->stream>>>(b, b + 4, b + 8);
    g_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, b + 8, 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SQ_OK;
}

}  // extern "C"
