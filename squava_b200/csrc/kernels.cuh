// kernels.cuh -- classification and random-playout kernels (sm_100a).
//
// Roofline: INT32 issue throughput (LOP3/IADD3 on the ALU pipe + IMAD on the FMA
// pipe); HBM traffic is 9 B in + 32 B out per LEAF, i.e. nothing.  No tensor
// cores: nothing here is a contraction.
#pragma once
#include "../../include/squava_b200.h"
#include "sq_device.cuh"

namespace sq {

__constant__ ScanTables c_scan;

// ---------------------------------------------------------------------------
// Terminal-position test
// ---------------------------------------------------------------------------

// First owned line in a reference scan order (slow path: both colours own one).
__device__ __noinline__ int ordered_winner(uint32_t x, uint32_t o, const uint32_t* quads,
                                           const uint32_t* trips) {
    for (int i = 0; i < 28; i++) {
        uint32_t q = quads[i];
        if ((x & q) == q) return 1;
        if ((o & q) == q) return -1;
    }
    for (int i = 0; i < 48; i++) {
        uint32_t t = trips[i];
        if ((x & t) == t) return -1;  // three in a row loses
        if ((o & t) == t) return 1;
    }
    return 0;
}

struct Classified { uint32_t flags; int winner_mcts, winner_ab; };

// x, o in API layout.  Restates FindWinner (src/mcts/mcts.go:92-121 and
// src/alphabeta/alphabeta.go:355-378): quads first, so four beats three.
__device__ __forceinline__ Classified classify_board(uint32_t x, uint32_t o) {
    Pairs px = pairs(pad(x)), po = pairs(pad(o));
    uint32_t x3 = any3(px), o3 = any3(po);
    uint32_t x4 = any4(px), o4 = any4(po);
    Classified c;
    c.flags = (x4 ? 1u : 0u) | (o4 ? 2u : 0u) | (x3 ? 4u : 0u) | (o3 ? 8u : 0u) |
              (((x | o) & kFull25) == kFull25 ? 16u : 0u);
    bool need_scan = (x4 && o4) || (!x4 && !o4 && x3 && o3);
    if (need_scan) {
        c.winner_mcts = ordered_winner(x, o, c_scan.mcts_quads, c_scan.mcts_triplets);
        c.winner_ab = ordered_winner(x, o, c_scan.ab_quads, c_scan.ab_triplets);
    } else {
        int w = x4 ? 1 : o4 ? -1 : x3 ? -1 : o3 ? 1 : 0;
        c.winner_mcts = c.winner_ab = w;
    }
    return c;
}

__global__ void __launch_bounds__(256)
classify_kernel(const sq_board* __restrict__ boards, int64_t n, uint8_t* __restrict__ flags,
                int8_t* __restrict__ wm, int8_t* __restrict__ wa) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint2 b = *reinterpret_cast<const uint2*>(boards + i);
        Classified c = classify_board(b.x & kFull25, b.y & kFull25);
        if (flags) flags[i] = (uint8_t)c.flags;
        if (wm) wm[i] = (int8_t)c.winner_mcts;
        if (wa) wa[i] = (int8_t)c.winner_ab;
    }
}

// ---------------------------------------------------------------------------
// Random playouts
// ---------------------------------------------------------------------------
//
// Work decomposition.  The playouts_per_leaf playouts of a leaf are cut into
// S = ceil(N / kItemQuota) ITEMS of (almost) equal quota; item t = leaf*S + j.
// A warp ROUND is 32 consecutive items, one per lane; warps pull rounds from a
// global counter.  Within a round every lane runs the same branch-free ply
// loop: a lane whose playout ends re-arms itself from its leaf in the same
// iteration (lane refill), so the warp stays converged at ply granularity.
//
// Per ply (mover's mask a, other's mask b, n empty cells):
//   draw   idx = hi32(r*n), r = lo32(r*n): mixed-radix extraction of 3 draws per
//          32-bit Philox word (bias <= n(n-1)(n-2)/2^32 < 3.3e-6 relative)
//   pick   the idx-th empty cell: per-lane Fisher-Yates array of cell BITS in
//          shared memory (word k of lane t at [k*blockDim + t]: bank = lane,
//          conflict-free).  Swap-with-last keeps it a permutation of the leaf's
//          empty cells, so a new playout needs no re-initialisation.
//   test   the mover's new mask only (only the mover can have made a line):
//          any3 != 0 or board full -> terminal; any4 decides win vs loss.
//
// Random stream of item (leaf, j): Philox4x32-10, key (seed_lo ^ stream_hi,
// seed_hi), counter (block, j, leaf_index, stream_lo); block b feeds plies
// 12b .. 12b+11 of the item, counted over ALL its playouts back to back.

#ifndef SQ_ITEM_QUOTA
#define SQ_ITEM_QUOTA 1024
#endif
constexpr int kItemQuota = SQ_ITEM_QUOTA;       // playouts per item (max); < 65536 (packed counters)
constexpr int kPlayoutThreads = 256;   // CTA size
constexpr int kMaxEmpty = 25;

struct PlayoutParams {
    const sq_board* leaves;
    const int8_t* to_move;
    int64_t n_leaves;
    uint64_t leaf_index_base;
    uint64_t per_leaf;         // N
    uint32_t items_per_leaf;   // S
    uint32_t quota_lo;         // N / S
    uint32_t quota_extra;      // N % S: items j < quota_extra run quota_lo + 1
    uint32_t key0, key1, ctr3;
    uint32_t rk[20];           // Philox round keys (key0 + r*W0, key1 + r*W1), r = 0..9: read straight from the
                               // constant bank by the round's LOP3, so the key schedule costs no instruction
    uint32_t one;              // == 1, but opaque to ptxas: x*one+y is issued as IMAD on the FMA pipe,
                               // which balances it against the ALU pipe (LOP3/SEL/ISETP) -- see DESIGN.md
    int64_t n_rounds;
    unsigned long long* round_counter;
    unsigned long long* out;   // [n_leaves][4]
    uint8_t* packed;           // non-null (per_leaf == 1 only): one outcome byte per leaf instead of `out`:
                               // bits 0-1 = 1 MAXIMIZER won, 2 MINIMIZER won, 3 draw; bits 2-7 = plies played
};

inline void playout_round_keys(PlayoutParams& P) {
    for (int r = 0; r < 10; r++) { P.rk[2 * r] = P.key0 + (uint32_t)r * kPhiloxW0; P.rk[2 * r + 1] = P.key1 + (uint32_t)r * kPhiloxW1; }
}

// Philox4x32-10 with the round keys precomputed (same function as philox4x32_10)
__device__ __forceinline__ void philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t (&rk)[20], uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
        uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
        c0 = hi1 ^ c1 ^ rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c3 = lo0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#ifndef SQ_PLAYOUT_CTAS_PER_SM
#define SQ_PLAYOUT_CTAS_PER_SM 5      // 40 warps per SM (48 registers per thread): measured best of 3..6
#endif
// tuning knobs of the ply loop (pipe balance; the random stream and the results do not depend on them)
#ifndef SQ_PL_ADDR_ALU
#define SQ_PL_ADDR_ALU 0      // element addresses by shift+add on the ALU pipe instead of IMAD (measured: IMAD is better)
#endif
#ifndef SQ_PL_LINE3
#define SQ_PL_LINE3 1         // line test from three-input ANDs of the shifted mask (22 instructions) instead of pair masks (24)
#endif
#ifndef SQ_PL_ACC_ALU
#define SQ_PL_ACC_ALU 0       // win counter by predicated add (ALU) instead of predicated IMAD
#endif
#ifndef SQ_PL_SHF
#define SQ_PL_SHF 0           // how many of the eight line-test shifts of a ply are forced onto the ALU pipe (SHF)
#endif

__device__ __forceinline__ uint32_t shl_alu(uint32_t x, int k) {      // x << k as SHF (ALU pipe); ptxas keeps it there
    uint32_t r;
    asm("shf.l.clamp.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(0u), "r"(x), "r"((uint32_t)k));
    return r;
}

struct PlyState {
    uint32_t a, b, n, inc_w, acc, acc_fin, plies;
    int remi;
};

// Twelve plies from one Philox block.  CAREFUL = false: no lane can finish its quota inside this block (every
// lane has more than 12 playouts to go), so the end-of-quota snapshot is not compiled in at all.
template <bool CAREFUL>
__device__ __forceinline__ void ply_block(PlyState& S, const uint32_t (&w)[4], uint32_t leaf_a, uint32_t leaf_b, uint32_t n0,
                                          uint32_t sbase, uint32_t one, uint32_t blk_after) {
    constexpr uint32_t kStride = kPlayoutThreads * 4;
    uint32_t a = S.a, b = S.b, n = S.n, inc_w = S.inc_w, acc = S.acc, acc_fin = S.acc_fin, plies = S.plies;
    int remi = S.remi;
#pragma unroll
    for (int wi = 0; wi < 4; wi++) {
        uint32_t r = w[wi];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            // draw: idx = floor(r*n / 2^32), r <- r*n mod 2^32
            const uint32_t idx = __umulhi(r, n);
            r = r * n;
#if SQ_PL_ADDR_ALU
            const uint32_t ai = shl_alu(idx, 10) + sbase;
            const uint32_t al = shl_alu(n, 10) + (sbase - kStride);
#else
            const uint32_t ai = idx * kStride + sbase;
            const uint32_t al = n * kStride + (sbase - kStride);
#endif
            uint32_t bit, last;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(bit) : "r"(ai));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(last) : "r"(al));
            asm volatile("st.shared.u32 [%0], %1;" :: "r"(ai), "r"(last));
            asm volatile("st.shared.u32 [%0], %1;" :: "r"(al), "r"(bit));
            SQ_CHECK(idx < n && n >= 1u && n <= 25u, 1, idx, n);                          // the draw stays inside the lane's array
            SQ_CHECK((bit & (bit - 1u)) == 0u && ((a | b) & bit) == 0u, 2, bit, a | b);    // one cell, and an empty one
            const uint32_t m = a + bit;          // bit is an empty cell: + is |
#if SQ_PL_LINE3
            // r_d = m & m<<d & m<<2d: bit at the END of every run of >= 3 in direction d; a run of >= 4 also has m<<3d
#define SQ_SHL(x, k, i) (SQ_PL_SHF > (i) ? shl_alu((x), (k)) : ((x) << (k)))      /* the first SQ_PL_SHF of these go to the ALU pipe */
            const uint32_t h1 = m << 1, h2 = m << 2, h3 = SQ_SHL(m, 3, 3), v1 = SQ_SHL(m, 6, 0), v2 = m << 12, v3 = SQ_SHL(m, 18, 4);
            const uint32_t d1 = SQ_SHL(m, 7, 1), d2 = m << 14, d3 = SQ_SHL(m, 21, 5), a1 = SQ_SHL(m, 5, 2), a2 = m << 10, a3 = m << 15;
#undef SQ_SHL
            const uint32_t rh = m & h1 & h2, rv = m & v1 & v2, rd = m & d1 & d2, ra = m & a1 & a2;
            const uint32_t t3 = rh | rv | rd | ra;
            const uint32_t t4 = (rh & h3) | (rv & v3) | (rd & d3) | (ra & a3);
#elif SQ_PL_SHF > 0
            Pairs p;
            p.h = m & (m << 1);
            p.v = m & (SQ_PL_SHF >= 1 ? shl_alu(m, 6) : (m << 6));
            p.d = m & (SQ_PL_SHF >= 2 ? shl_alu(m, 7) : (m << 7));
            p.a = m & (SQ_PL_SHF >= 3 ? shl_alu(m, 5) : (m << 5));
            const uint32_t t3 = any3(p), t4 = any4(p);
#else
            const Pairs p = pairs(m);
            const uint32_t t3 = any3(p), t4 = any4(p);
#endif
            const uint32_t n1 = n - 1u;
            const bool line = t3 != 0;
            const bool term = line | (n1 == 0);
            const uint32_t inc_l = inc_w ^ 0x10001u;
            const uint32_t inc = t4 ? inc_w : inc_l;
#if SQ_PL_ACC_ALU
            asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q add.u32 %0, %0, %2;\n\t}" : "+r"(acc) : "r"(t3), "r"(inc));
#else
            asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q mad.lo.u32 %0, %2, %3, %0;\n\t}"
                : "+r"(acc) : "r"(t3), "r"(inc), "r"(one));
#endif
            if (CAREFUL) {
                const bool fin = term & (remi == 1);
                const uint32_t plies_now = blk_after * 12u - 11u + (uint32_t)(wi * 3 + d);
                asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %4, 0;\n\t@q mov.u32 %0, %2;\n\t@q mov.u32 %1, %3;\n\t}"
                    : "+r"(acc_fin), "+r"(plies) : "r"(acc), "r"(plies_now), "r"((uint32_t)fin));
            }
            asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q add.s32 %0, %0, -1;\n\t}"
                : "+r"(remi) : "r"((uint32_t)term));
            a = term ? leaf_a : b;
            b = term ? leaf_b : m;
            n = term ? n0 : n1;
            inc_w = term ? 1u : inc_l;
        }
    }
    S.a = a; S.b = b; S.n = n; S.inc_w = inc_w; S.acc = acc; S.acc_fin = acc_fin; S.plies = plies; S.remi = remi;
}

__global__ void __launch_bounds__(kPlayoutThreads, SQ_PLAYOUT_CTAS_PER_SM)
playout_kernel(const __grid_constant__ PlayoutParams P) {
    extern __shared__ uint32_t s_arr[];  // [kMaxEmpty][blockDim.x]
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    uint32_t* arr = s_arr + tid;         // element k at arr[k * kPlayoutThreads]

    for (;;) {
        unsigned long long round = 0;
        if (lane == 0) round = atomicAdd(P.round_counter, 1ull);
        round = __shfl_sync(0xffffffffu, round, 0);
        if ((int64_t)round >= P.n_rounds) break;

        // ---- per-lane item setup -------------------------------------------
        const uint64_t item = round * 32ull + lane;
        const uint64_t leaf64 = item / P.items_per_leaf;
        const uint32_t j = (uint32_t)(item - leaf64 * P.items_per_leaf);
        const bool have = leaf64 < (uint64_t)P.n_leaves;
        const uint32_t quota = have ? P.quota_lo + (j < P.quota_extra ? 1u : 0u) : 0u;

        uint32_t leaf_a = 0, leaf_b = 0, n0 = 1, rem = 0, played = 0;
        uint32_t first_is_x = 1;
        uint32_t direct_x = 0, direct_o = 0, direct_d = 0;  // outcomes decided without playing
        arr[0] = 0;
        if (have) {
            uint2 bd = *reinterpret_cast<const uint2*>(P.leaves + leaf64);
            int tm = P.to_move[leaf64];
            uint32_t x = bd.x & kFull25, o = bd.y & kFull25;
            first_is_x = tm > 0;
            Classified c = classify_board(x, o);
            if (c.flags) {
                // GetMoves says terminal (mcts.go:316-330,336-345): zero plies, result
                // straight from GetResult's scan order (mcts.go:348-388).
                int w = c.winner_mcts;
                direct_x = w > 0 ? quota : 0u;
                direct_o = w < 0 ? quota : 0u;
                direct_d = w == 0 ? quota : 0u;
            } else {
                uint32_t xp = pad(x), op = pad(o);
                leaf_a = first_is_x ? xp : op;
                leaf_b = first_is_x ? op : xp;
                uint32_t e = kFullPadded & ~(xp | op);
                n0 = __popc(e);
                for (uint32_t k = 0; k < n0; k++) {
                    uint32_t bit = e & (0u - e);
                    e ^= bit;
                    arr[k * kPlayoutThreads] = bit;
                }
                rem = quota;
                played = quota;
            }
        }

        // ---- ply loop --------------------------------------------------------
        // No per-ply gating of finished lanes: a lane keeps playing after its quota is done, and
        // its counters are SNAPSHOT (two predicated moves) at the ply its last playout ends.
        // Plies need no counter at all: every iteration of an unfinished lane is one ply, so
        // the snapshot takes the (warp-uniform) iteration number.
        // Two phases: while EVERY playing lane still has more than 12 playouts to go no lane can finish inside
        // the next block of 12 plies, and the block runs without the snapshot logic; the last few blocks of a
        // round run the careful version.  Lanes with nothing to play (no leaf, terminal leaf) idle along.
        PlyState S;
        S.a = leaf_a; S.b = leaf_b; S.n = n0;
        S.inc_w = 1u;                 // what the mover's win adds: 1 = first player, 0x10000 = second
        S.acc = 0; S.acc_fin = 0; S.plies = 0;
        S.remi = (int)rem;            // playouts still to finish; <= 0: done
        uint32_t blk = 0;
        const uint32_t c1 = j, c2 = (uint32_t)(P.leaf_index_base + leaf64);
        const uint32_t one = P.one;
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(arr);   // element k at sbase + k*kStride
        const bool idle = rem == 0;
        while (__all_sync(0xffffffffu, idle | (S.remi > 12))) {
            if (__all_sync(0xffffffffu, idle)) break;
            uint32_t w[4];
            philox4x32_10_rk(blk, c1, c2, P.ctr3, P.rk, w);
            blk++;
            ply_block<false>(S, w, leaf_a, leaf_b, n0, sbase, one, blk);
        }
        if (idle) S.remi = 0;
        while (__any_sync(0xffffffffu, S.remi > 0)) {
            uint32_t w[4];
            philox4x32_10_rk(blk, c1, c2, P.ctr3, P.rk, w);
            blk++;
            ply_block<true>(S, w, leaf_a, leaf_b, n0, sbase, one, blk);
        }
        const uint32_t acc = S.acc_fin, plies = S.plies;

        // ---- fold into per-leaf counts ------------------------------------------
        uint32_t first_wins = acc & 0xffffu, second_wins = acc >> 16;
        uint32_t xw = first_is_x ? first_wins : second_wins;
        uint32_t ow = first_is_x ? second_wins : first_wins;
        uint32_t dr = played - first_wins - second_wins;   // playouts that filled the board lineless
        xw += direct_x; ow += direct_o; dr += direct_d;
        if (P.packed) {                                   // one playout per leaf: the lane owns its leaf's byte
            if (have) P.packed[leaf64] = (uint8_t)((xw ? 1u : ow ? 2u : 3u) | (plies << 2));
            __syncwarp();
            continue;
        }
        uint32_t key = have ? (uint32_t)leaf64 : 0xffffffffu;
        uint32_t peers = __match_any_sync(0xffffffffu, key);
        uint32_t sx = __reduce_add_sync(peers, xw);
        uint32_t so = __reduce_add_sync(peers, ow);
        uint32_t sd = __reduce_add_sync(peers, dr);
        uint32_t sp = __reduce_add_sync(peers, plies);
        if (have && lane == (__ffs(peers) - 1)) {
            unsigned long long* o4 = P.out + leaf64 * 4;
            if (sx) atomicAdd(o4 + 0, (unsigned long long)sx);
            if (so) atomicAdd(o4 + 1, (unsigned long long)so);
            if (sd) atomicAdd(o4 + 2, (unsigned long long)sd);
            if (sp) atomicAdd(o4 + 3, (unsigned long long)sp);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// INT32 issue-rate probe: the measured roofline denominator
// ---------------------------------------------------------------------------
// kind 0: 8 independent LOP3 chains (ALU pipe); kind 1: 8 IMAD chains (FMA pipe);
// kind 2: 4 + 4 interleaved.  Every op is one warp instruction = 32 thread-ops.
template <int KIND>
__global__ void __launch_bounds__(1024, 2)
int32_peak_kernel(uint32_t* out, uint32_t seed, int iters) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = seed + threadIdx.x * 0x9E3779B9u + k * 0x85EBCA6Bu;
    uint32_t c0 = seed | 1u, c1 = seed * 3u + 7u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                bool alu = KIND == 0 || (KIND == 2 && (k & 1) == 0);
                if (alu) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[k]) : "r"(c0), "r"(c1));
                } else {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[k]) : "r"(c0), "r"(c1));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= v[k];
    if (s == 0x12345678u) out[0] = s;  // keeps the chains alive
}

__global__ void philox_debug_kernel(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    uint32_t w[4];
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], w);
    for (int i = 0; i < 4; i++) out[i] = w[i];
}

}  // namespace sq
