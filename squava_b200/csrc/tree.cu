// tree.cu -- sq_mcts_decide: one UCT decision (src/mcts/mcts.go:140-200) with the whole tree in HBM (tree.cuh).
// A second translation unit of libsquava_b200.so: it uses the library's own device-resident entry point
// (sq_playouts_dev) for the rollouts, so the playout kernel and its random streams are exactly the ones every other
// caller gets.
#include "../../include/squava_b200.h"

#include <cstdio>
#include <cstring>

#include <cuda_runtime.h>

#include "tree.cuh"

namespace {

thread_local char t_tree_err[256] = "";

// the first `levels` levels of a batch's descent, synchronised (tree.cuh): arrive / decide once per level over all walks
__global__ void __launch_bounds__(256)
tree_arrive_kernel(sqt::Tree T, int level, uint32_t root, uint32_t rx, uint32_t ro, int pjm, uint32_t P, long long B, sqt::Walks W) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (level == 0 && k == 0) atomicAdd(&T.visits[root], (uint32_t)(B * (long long)P));     // every walk passes the root: one add for the batch
    if (k < B) sqt::tree_arrive(T, level, root, rx, ro, pjm, P, k, W, true);
}

__global__ void __launch_bounds__(256)
tree_decide_kernel(sqt::Tree T, uint32_t P, float c2, uint64_t key, long long B, sqt::Walks W) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < B) sqt::tree_decide(T, P, c2, key, k, W);
}

// the rest of every walk, on its own.  levels == 0: the whole walk, and the root's visits
__global__ void __launch_bounds__(128)
tree_finish_kernel(sqt::Tree T, int levels, uint32_t root, uint32_t rx, uint32_t ro, int pjm, uint32_t P, float c2, uint64_t key, long long B,
                   sqt::Walks W) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, n = (long long)gridDim.x * blockDim.x;
    if (levels == 0 && tid == 0) atomicAdd(&T.visits[root], (uint32_t)(B * (long long)P));
    // a thread takes several walks of the batch, one after the other: the later ones see the visits of the earlier ones
    for (long long k = tid; k < B; k += n) sqt::tree_finish(T, levels, root, rx, ro, pjm, P, c2, key, k, W);
}

__global__ void __launch_bounds__(256)
tree_update_kernel(sqt::Tree T, uint32_t root, int pjm, const unsigned long long* counts, long long B, sqt::Walks W) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t r = 0;
    if (k < B) r = sqt::tree_update(T, pjm, counts + 4 * k, k, W);
    r = __reduce_add_sync(0xffffffffu, r);                      // the root's wins: one atomic per warp
    if ((threadIdx.x & 31) == 0 && r) atomicAdd(&T.wins[root], r);
}

__global__ void tree_init_kernel(sqt::Tree T, uint32_t root, uint32_t untried) {
    T.untried[root] = untried;
    *T.next_free = 2;                                           // 0 = none, 1 = the root
    *T.overflow = 0;
}

struct Buffers {                                                // everything the call allocates, freed on every way out
    static constexpr int kMost = 32;
    void* p[kMost];
    int n = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Buffers() {
        for (int i = 0; i < n; i++) cudaFree(p[i]);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (stream) cudaStreamDestroy(stream);
    }
    template <class T> cudaError_t get(T** out, size_t count) {
        void* q = nullptr;
        if (n >= kMost) { *out = nullptr; return cudaErrorMemoryAllocation; }
        const cudaError_t e = cudaMalloc(&q, count * sizeof(T));
        if (e == cudaSuccess) p[n++] = q;
        *out = (T*)q;
        return e;
    }
};

#define TCK(call)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) {                                                                               \
            snprintf(t_tree_err, sizeof t_tree_err, "sq_mcts_decide: %s: %s", #call, cudaGetErrorString(e_));  \
            return e_ == cudaErrorMemoryAllocation ? SQ_ERR_NOMEM : SQ_ERR_CUDA;                               \
        }                                                                                                      \
    } while (0)

}  // namespace

extern "C" {

const char* sq_mcts_last_error(void) { return t_tree_err; }

int sq_mcts_decide(int slot, const sq_board* root_board, int player_just_moved, int64_t iterations, double uctk, int ucb_factor2,
                   int64_t batch, int playouts_per_leaf, int walkers, int sync_levels, uint64_t max_nodes, uint64_t stream_base,
                   uint64_t seed, sq_mcts_result* out) {
    t_tree_err[0] = 0;
    (void)cudaGetLastError();                                   // whatever an earlier call left behind is not this call's
    if (!root_board || !out || iterations < 1 || batch < 1 || playouts_per_leaf < 1 || (player_just_moved != 1 && player_just_moved != -1) ||
        (root_board->x & root_board->o) || ((root_board->x | root_board->o) >> 25) || iterations > 4000000000ll) {        // 32-bit counts
        snprintf(t_tree_err, sizeof t_tree_err, "sq_mcts_decide: bad argument");
        return SQ_ERR_INVALID;
    }
    const int dev = sq_device_id(slot);
    if (dev < 0) { snprintf(t_tree_err, sizeof t_tree_err, "sq_mcts_decide: %s", sq_last_error()); return dev; }
    TCK(cudaSetDevice(dev));
    memset(out, 0, sizeof *out);
    const uint32_t P = (uint32_t)playouts_per_leaf;
    const uint32_t rx = root_board->x, ro = root_board->o, emp = sqt::kFull & ~(rx | ro);
    if (walkers < 1) walkers = 8192;
    if (sync_levels < 0) sync_levels = 4;
    if (sync_levels > 24) sync_levels = 24;
    if (batch > (1 << 22)) batch = 1 << 22;

    // the arena: every expanded node costs one slot per empty cell -- 5 slots per iteration in a small search, 2.2 after 10^9
    // iterations of the host tree (profiles/r02_config3_fast_tree_1e9_16byte.txt).  32-bit indices; when the arena is full
    // the search goes on without growing (arena_full).
    unsigned long long cap = max_nodes ? max_nodes : (unsigned long long)(6.0 * (double)(iterations / P + 1)) + (1ull << 20);
    if (cap > 0xfffffff0ull) cap = 0xfffffff0ull;
    size_t free_b = 0, total_b = 0;
    TCK(cudaMemGetInfo(&free_b, &total_b));
    const unsigned long long fits = (unsigned long long)((double)free_b * 0.8 / (sync_levels ? 20.0 : 16.0));
    if (cap > fits) cap = fits;
    if (cap < 64) { snprintf(t_tree_err, sizeof t_tree_err, "sq_mcts_decide: no device memory for the tree"); return SQ_ERR_NOMEM; }

    Buffers buf;
    sqt::Tree T;
    T.cap = cap;
    TCK(buf.get(&T.visits, cap)); TCK(buf.get(&T.wins, cap)); TCK(buf.get(&T.first_child, cap)); TCK(buf.get(&T.untried, cap));
    TCK(buf.get(&T.next_free, 1)); TCK(buf.get(&T.overflow, 1));
    T.arrivals = nullptr;
    if (sync_levels) TCK(buf.get(&T.arrivals, cap));
    sqt::Walks W;
    unsigned long long* counts = nullptr;
    TCK(buf.get(&W.boards, (size_t)batch)); TCK(buf.get(&W.to_move, (size_t)batch)); TCK(buf.get(&W.leaf, (size_t)batch));
    TCK(buf.get(&W.fresh, (size_t)batch)); TCK(buf.get(&W.depth, (size_t)batch)); TCK(buf.get(&W.paths, (size_t)batch * sqt::kPathLen));
    TCK(buf.get(&counts, (size_t)batch * 4));
    TCK(buf.get(&W.rank, (size_t)batch)); TCK(buf.get(&W.act, (size_t)batch)); TCK(buf.get(&W.arg, (size_t)batch)); TCK(buf.get(&W.pjm, (size_t)batch));
    TCK(cudaStreamCreateWithFlags(&buf.stream, cudaStreamNonBlocking));
    TCK(cudaEventCreate(&buf.e0)); TCK(cudaEventCreate(&buf.e1));
    cudaStream_t st = buf.stream;
    TCK(cudaMemsetAsync(T.visits, 0, cap * 4, st)); TCK(cudaMemsetAsync(T.wins, 0, cap * 4, st));
    TCK(cudaMemsetAsync(T.first_child, 0, cap * 4, st)); TCK(cudaMemsetAsync(T.untried, 0, cap * 4, st));
    if (T.arrivals) TCK(cudaMemsetAsync(T.arrivals, 0, cap * 4, st));
    const uint32_t root = 1;
    tree_init_kernel<<<1, 1, 0, st>>>(T, root, emp);
    TCK(cudaGetLastError());

    const float c2 = (float)(uctk * uctk * (ucb_factor2 ? 2.0 : 1.0));
    TCK(cudaEventRecord(buf.e0, st));
    long long done = 0, batches = 0;
    while (done < iterations) {
        const long long remaining = iterations - done;
        uint32_t p = P;
        if ((long long)p > remaining) p = (uint32_t)remaining;
        long long B = batch;
        if (batches < 16 && (256ll << batches) < B) B = 256ll << batches;       // the first batches grow with the tree: 256, 512, ...
        if (B * (long long)p > remaining) B = remaining / p;
        if (B < 1) B = 1;
        const uint64_t key = sqt::t_mix(seed ^ (uint64_t)batches * 0x9e3779b97f4a7c15ull);
        const unsigned per_walk = (unsigned)((B + 255) / 256);
        for (int level = 0; level < sync_levels; level++) {
            tree_arrive_kernel<<<per_walk, 256, 0, st>>>(T, level, root, rx, ro, player_just_moved, p, B, W);
            TCK(cudaGetLastError());
            tree_decide_kernel<<<per_walk, 256, 0, st>>>(T, p, c2, key, B, W);
            TCK(cudaGetLastError());
        }
        const long long w = B < walkers ? B : walkers;
        tree_finish_kernel<<<(unsigned)((w + 127) / 128), 128, 0, st>>>(T, sync_levels, root, rx, ro, player_just_moved, p, c2, key, B, W);
        TCK(cudaGetLastError());
        const int rc = sq_playouts_dev(slot, W.boards, W.to_move, B, 0, p, stream_base + (uint64_t)batches, (uint64_t*)counts, st);
        if (rc != SQ_OK) { snprintf(t_tree_err, sizeof t_tree_err, "sq_mcts_decide: %s", sq_last_error()); return rc; }
        tree_update_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(T, root, player_just_moved, counts, B, W);
        TCK(cudaGetLastError());
        done += B * (long long)p;
        batches++;
    }
    TCK(cudaEventRecord(buf.e1, st));
    // the decision: the root's children
    uint32_t fc = 0, root_visits = 0, overflow = 0;
    unsigned long long used = 0;
    TCK(cudaMemcpyAsync(&fc, T.first_child + root, 4, cudaMemcpyDeviceToHost, st));
    TCK(cudaMemcpyAsync(&root_visits, T.visits + root, 4, cudaMemcpyDeviceToHost, st));
    TCK(cudaMemcpyAsync(&overflow, T.overflow, 4, cudaMemcpyDeviceToHost, st));
    TCK(cudaMemcpyAsync(&used, T.next_free, 8, cudaMemcpyDeviceToHost, st));
    TCK(cudaStreamSynchronize(st));
    float ms = 0.f;
    TCK(cudaEventElapsedTime(&ms, buf.e0, buf.e1));
    const int slots = __builtin_popcount(emp);
    uint32_t v[25] = {0}, wn[25] = {0};
    if (fc && slots) {
        TCK(cudaMemcpy(v, T.visits + fc, (size_t)slots * 4, cudaMemcpyDeviceToHost));
        TCK(cudaMemcpy(wn, T.wins + fc, (size_t)slots * 4, cudaMemcpyDeviceToHost));
    }
    int n = 0;
    uint32_t e = emp;
    for (int k = 0; k < slots; k++, e &= e - 1u) {
        if (!v[k]) continue;
        out->child_move[n] = __builtin_ctz(e);
        out->child_visits[n] = (double)v[k];
        out->child_wins[n] = (double)wn[k];
        n++;
    }
    out->n_children = n;
    out->root_visits = (double)root_visits;
    out->iterations = done;
    out->batches = batches;
    out->nodes = used;
    out->arena_slots = cap;
    out->arena_full = (int32_t)overflow;
    out->device_ms = (double)ms;
    return SQ_OK;
}

}  // extern "C"
