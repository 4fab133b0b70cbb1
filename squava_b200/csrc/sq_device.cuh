// sq_device.cuh -- board primitives shared by the squava_b200 kernels (sm_100a).
//
// A board is two 25-bit masks held in registers.  Two layouts are used:
//   * API layout: bit c = 5*x + y            (what crosses the C ABI)
//   * padded layout: bit p = 6*x + y          (one always-zero pad column)
// In the padded layout a line is detected with shifts alone: the pad column
// breaks every wrap-around of the four line directions (steps 1, 6, 7, 5), so
// no start-position masks are needed.  Left shifts are used throughout: the
// compiler may issue them as IMAD.SHL on the FMA pipe, next to LOP3 on the ALU
// pipe, which is what balances the two INT32 issue ports.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sq {

// Checked build (-DSQ_CHECKED=1, `make checked`): compute-sanitizer is closed on the pool this was developed on
// (profiles/r02_sanitizer.md), so the kernels carry their own bounds checks -- node indices, queue slots, search-stack
// frames, Fisher-Yates slots -- compiled in only here and reported through sq_debug_check_failures().
#ifndef SQ_CHECKED
#define SQ_CHECKED 0
#endif
#if SQ_CHECKED
__device__ unsigned int g_sq_check_fail[4];   // [0] failures, [1] code of the first, [2] [3] its two arguments
__device__ __forceinline__ void sq_check_fail(unsigned code, unsigned a, unsigned b) {
    if (atomicAdd(&g_sq_check_fail[0], 1u) == 0) { g_sq_check_fail[1] = code; g_sq_check_fail[2] = a; g_sq_check_fail[3] = b; }
}
#define SQ_CHECK(cond, code, a, b) do { if (!(cond)) ::sq::sq_check_fail((code), (unsigned)(a), (unsigned)(b)); } while (0)
#else
#define SQ_CHECK(cond, code, a, b) do { } while (0)
#endif

constexpr uint32_t kFull25 = 0x1FFFFFFu;
constexpr uint32_t kFullPadded = 0x1F7DF7DFu;  // 5 cells in each of 5 six-bit rows

__host__ __device__ __forceinline__ uint32_t pad(uint32_t m) {
    return (m & 0x1Fu) | ((m & 0x3E0u) << 1) | ((m & 0x7C00u) << 2) | ((m & 0xF8000u) << 3) |
           ((m & 0x1F00000u) << 4);
}
__host__ __device__ __forceinline__ uint32_t unpad(uint32_t p) {
    return (p & 0x1Fu) | ((p >> 1) & 0x3E0u) | ((p >> 2) & 0x7C00u) | ((p >> 3) & 0xF8000u) |
           ((p >> 4) & 0x1F00000u);
}

// Padded mask -> bit set at the END cell of every run of >= 2 / 3 / 4 in each direction.
struct Pairs { uint32_t h, v, d, a; };
__device__ __forceinline__ Pairs pairs(uint32_t m) {
    Pairs p;
    p.h = m & (m << 1);
    p.v = m & (m << 6);
    p.d = m & (m << 7);
    p.a = m & (m << 5);
    return p;
}
__device__ __forceinline__ uint32_t any3(const Pairs& p) {
    return (p.h & (p.h << 1)) | (p.v & (p.v << 6)) | (p.d & (p.d << 7)) | (p.a & (p.a << 5));
}
__device__ __forceinline__ uint32_t any4(const Pairs& p) {
    return (p.h & (p.h << 2)) | (p.v & (p.v << 12)) | (p.d & (p.d << 14)) | (p.a & (p.a << 10));
}

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so any (leaf, sub-stream,
// block) triple is addressable without state.
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
        uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += kPhiloxW0;
        k1 += kPhiloxW1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Scan-order tables (API layout masks) for the slow path of the winner scans:
// only consulted when BOTH colours own a line, which no reachable position does.
struct ScanTables {
    uint32_t mcts_quads[28];     // src/mcts/mcts.go:409-425, important-cell order
    uint32_t mcts_triplets[48];  // src/mcts/mcts.go:430-446, first occurrences in scan order
    uint32_t ab_quads[28];       // src/alphabeta/alphabeta.go:299-328
    uint32_t ab_triplets[48];    // src/alphabeta/alphabeta.go:249-298
};

}  // namespace sq
