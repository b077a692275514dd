// wh_kernels.cuh — device code of the B200-native batched warehouse hot path (sm_100a).
//
// Mapping: one GROUP of G consecutive lanes (G = max(R, P/4): 4 for Small, 9 for Medium, 16 for
// Large) owns one environment; floor(32/G) environments share a warp and run in lock-step (the
// 32 mod G left-over lanes idle as "ghosts"). Within a group
//   * lane a  (a < R) holds agent a: its cell (x | y<<8), its delivery target, its action;
//   * lane l  holds pickup points 4l..4l+3: four int8 delivery targets packed in one 32-bit
//     register and four timers;
//   * lane r  (r < R) holds request r / "other agent" slot r while observations are written.
// Everything that is per-env scalar (time, num_agents, the 64-bit active-request mask) is kept
// redundantly in every lane of the group; all exchange is by warp shuffles and ballots. Shared
// memory is used only to stage the small-row observation keys for full-width stores.
//
// Reference semantics (file:line into ffahleraz/rllib-warehouse) are cited at each phase.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wh_b200.h"

// Observation stores are written once and never re-read by these kernels. Measured on B200
// (profiles/README.md): for the large-row variants plain write-back stores are ~4 % faster than the
// streaming hint (st.global.cs), for Small (tiny rows, many envs per warp) .cs is ~4 % faster.
#define WH_ST(p, v) ::wh::st_obs<(RC == 4)>((p), (v))

namespace wh {

// Environment index inside one launch. 32-bit on purpose: every per-key element index (at most
// e*R*R) then stays a single IMAD / IMAD.WIDE.U32; the launchers reject N*R*R >= 2^32.
typedef uint32_t env_t;

// L2 eviction priority for the STATE (createpolicy + .L2::cache_hint; the policy rides in the memory
// descriptor, no extra instruction per access). The state — a few tens of MB — is written by step t and read
// back by step t+1, but the 0.2 - 2.2 GB observation stream of a step pushes it out of the 126 MB L2 in
// between, so every step starts with DRAM-latency loads queued behind the write stream (the largest single
// stall of the step kernels, profiles/r02_hotspots_*.txt). Loading and storing the state evict_last keeps it
// resident: Medium 65 536 envs 0.855 -> 0.924 of the HBM peak, 131 072 envs 0.903 -> 0.953, Large 65 536 envs
// 0.966 -> 1.004. It costs L2 capacity the write-back path also wants, so it only pays while the state is
// small against L2: at 262 144 envs (37 / 64 MB) it is -0.4 % / -1.3 %, and Small (whose 143 MB observation
// stream leaves the state in L2 anyway) loses 3.6 %. A FRACTIONAL policy (half of the accesses evict_last)
// still pays for Medium at 37 MB (0.923 -> 0.939) but not for Large at 64 MB (-0.6 .. -1.5 %). Choosing the
// part BY ARRAY instead of at random is better again (Medium 262 144 envs 0.928 -> 0.950 / 0.960 on two boxes,
// with the in-kernel solver 0.921 -> 0.945): with a random half nearly every warp still has some load that
// misses, with "everything except the timers" (55 % of the bytes) the loads a warp needs FIRST always hit
// and the timers — first used after the move loop — arrive behind it. Dropping pickup_tgt as well: 0.936;
// the same for Large (all but the timers, or also without pickup_tgt): 0.965 / 0.977 vs 0.977 plain. Hence
// compile-time variants of the throughput (PLAIN) kernels — KEEP = 1 (all) / 2 (part) — that the launcher
// selects from the state size (WH_KEEP_MAX_MB, and twice that for the partial policy of Medium). Observation
// stores with evict_first were measured too: -4 % everywhere, not used.
#ifndef WH_KEEP_MAX_MB
#define WH_KEEP_MAX_MB 32
#endif
#ifndef WH_KEEP_PART_FRAC
#define WH_KEEP_PART_FRAC 0.5   // share of the state accesses that get evict_last under the partial policy (KEEP = 2)
#endif
#define WH_STR2(x) #x
#define WH_STR(x) WH_STR2(x)
// How KEEP = 2 picks its part of the state. 0 = a random share (WH_KEEP_PART_FRAC) of all accesses; 1 (shipped) =
// every array except the timers (45 % of the bytes, first used after the move loop); 2 = also without pickup_tgt
#ifndef WH_KEEP_PART_MODE
#define WH_KEEP_PART_MODE 1
#endif
// KEEP levels: 0 = plain accesses, 1 = every state access evict_last, 2 = part of them (see WH_KEEP_PART_MODE),
// 3 = every array except the timers and pickup_tgt (tuning builds)
template <int KEEP>
__device__ __forceinline__ uint64_t l2_evict_last() {
    uint64_t p;
    if constexpr (KEEP == 2 && WH_KEEP_PART_MODE == 0) asm("createpolicy.fractional.L2::evict_last.b64 %0, " WH_STR(WH_KEEP_PART_FRAC) ";" : "=l"(p));
    else asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// BIG: 1 = the timers, 2 = pickup_tgt, 0 = everything else
template <int KEEP, int BIG>
struct KeepThis { static constexpr bool value = (KEEP == 1 || KEEP == 2) ? !(KEEP == 2 && WH_KEEP_PART_MODE != 0 && BIG != 0 && BIG <= WH_KEEP_PART_MODE) : (KEEP == 3 && BIG == 0); };

template <typename T>
__device__ __forceinline__ void st_hint(T *p, const T &v, uint64_t pol) {
    static_assert(sizeof(T) == 16 || sizeof(T) == 8 || sizeof(T) == 4 || sizeof(T) == 2 || sizeof(T) == 1, "store width");
    if constexpr (sizeof(T) == 16) {
        const uint4 u = *reinterpret_cast<const uint4 *>(&v);
        asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w), "l"(pol) : "memory");
    } else if constexpr (sizeof(T) == 8) {
        const uint2 u = *reinterpret_cast<const uint2 *>(&v);
        asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(u.x), "r"(u.y), "l"(pol) : "memory");
    } else if constexpr (sizeof(T) == 4) {
        const uint32_t u = *reinterpret_cast<const uint32_t *>(&v);
        asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(u), "l"(pol) : "memory");
    } else if constexpr (sizeof(T) == 2) {
        const uint16_t u = *reinterpret_cast<const uint16_t *>(&v);
        asm volatile("st.global.L2::cache_hint.b16 [%0], %1, %2;" ::"l"(p), "h"(u), "l"(pol) : "memory");
    } else {
        const uint32_t u = *reinterpret_cast<const uint8_t *>(&v);
        asm volatile("st.global.L2::cache_hint.b8 [%0], %1, %2;" ::"l"(p), "r"(u), "l"(pol) : "memory");
    }
}

template <typename T>
__device__ __forceinline__ T ld_hint(const T *p, uint64_t pol) {
    static_assert(sizeof(T) == 16 || sizeof(T) == 8 || sizeof(T) == 4 || sizeof(T) == 2 || sizeof(T) == 1, "load width");
    T out;
    if constexpr (sizeof(T) == 16) {
        uint4 u;
        asm("ld.global.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p), "l"(pol));
        out = *reinterpret_cast<T *>(&u);
    } else if constexpr (sizeof(T) == 8) {
        uint2 u;
        asm("ld.global.L2::cache_hint.v2.b32 {%0,%1}, [%2], %3;" : "=r"(u.x), "=r"(u.y) : "l"(p), "l"(pol));
        out = *reinterpret_cast<T *>(&u);
    } else if constexpr (sizeof(T) == 4) {
        uint32_t u;
        asm("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(u) : "l"(p), "l"(pol));
        out = *reinterpret_cast<T *>(&u);
    } else if constexpr (sizeof(T) == 2) {
        uint16_t u;
        asm("ld.global.L2::cache_hint.b16 %0, [%1], %2;" : "=h"(u) : "l"(p), "l"(pol));
        out = *reinterpret_cast<T *>(&u);
    } else {
        uint32_t u;
        asm("ld.global.L2::cache_hint.b8 %0, [%1], %2;" : "=r"(u) : "l"(p), "l"(pol));
        const uint8_t b = (uint8_t)u;
        out = *reinterpret_cast<const T *>(&b);
    }
    return out;
}

// state accessors
template <int KEEP, int BIG = 0, typename T>
__device__ __forceinline__ T ld_state(const T *p) {
    if constexpr (KeepThis<KEEP, BIG>::value) return ld_hint(p, l2_evict_last<KEEP>());
    else return *p;
}
template <int KEEP, int BIG = 0, typename T>
__device__ __forceinline__ void st_state(T *p, const T &v) {
    if constexpr (KeepThis<KEEP, BIG>::value) st_hint(p, v, l2_evict_last<KEEP>());
    else *p = v;
}

template <bool STREAM, typename T>
__device__ __forceinline__ void st_obs(T *p, const T &v) {
    if (STREAM) __stcs(p, v);
    else *p = v;
}

constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t ABSENT_MOVE = 0xffffffffu;
constexpr uint32_t NO_CELL = 0xffff0000u;  // never equals a packed 16-bit cell

// Philox counter words reserved for reset draws (step draws use c2 = episode_time >= 1)
constexpr uint32_t CTR_NUM_AGENTS = 0xFFFFFFFFu;
constexpr uint32_t CTR_SPAWN_AGENT = 0xF0000000u;
constexpr uint32_t CTR_INIT_REQUESTS = 0xE0000000u;
// The solver's eps-random draws use the same (env, episode, time) counter words as the env's respawn
// draws; bit 31 of the lane word separates the two streams even when solver_seed == seed.
constexpr uint32_t CTR_SOLVER_TAG = 0x80000000u;

struct KParams {
    // geometry (core.py:92-108, variants.py)
    int R, dim, L, P, D, episode, wait, null_pos, max_agents, random_agents, regular_racks;
    int G;      // lanes per environment
    int invL;   // ceil(256 / L): q / L == (q * invL) >> 8 for q < 64, L <= 8
    int racks[WH_MAX_RACKS];
    // state (wh_state)
    int8_t *agent_pos, *agent_tgt, *pickup_tgt;
    int16_t *pickup_timer;
    int32_t *time;
    int8_t *num_agents;
    int32_t *episode_ctr, *acc;
    wh_obs obs;  // all NULL => no observation build
    // step inputs / outputs
    const int32_t *actions, *order;
    const int8_t *spawn_p, *spawn_t;
    float *rewards;
    uint8_t *dones;
    unsigned long long *stats;
    int32_t *actions_out;
    float *flat_out;    // RLlib-flattened observations [N,R,9R+1] (k_obs_flat)
    // reset replay
    const int8_t *r_agent_pos, *r_init_p, *r_init_t, *r_num_agents;
    const uint8_t *env_mask;
    // solver-from-obs inputs
    const uint8_t *is_random;
    const int32_t *random_actions;
    const int8_t *g_num_agents;
    const int32_t *g_episode, *g_time;
    long long N, env_id0;
    unsigned long long seed, solver_seed, rand_thr;
    int flags, flavour;
    int n_steps;        // k_rollout / k_multi: steps per launch
};

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              unsigned long long seed, uint32_t &o0, uint32_t &o1) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1;
}

__device__ __forceinline__ uint32_t bounded(uint32_t u, uint32_t n) { return __umulhi(u, n); }

// index of the n-th (0-based, ascending) set bit (n < popc(mask)); BITS = how many low bits of the
// mask can be set (16 / 32 / 64) — the binary search skips the levels that cannot matter
template <int BITS>
__device__ __forceinline__ int nth_set64(unsigned long long m, int n) {
    uint32_t w = (uint32_t)m;
    int pos = 0, c;
    if (BITS > 32) {
        c = __popc(w);
        if (n >= c) { n -= c; w = (uint32_t)(m >> 32); pos = 32; }
    }
    if (BITS > 16) {
        c = __popc(w & 0xFFFFu);
        if (n >= c) { n -= c; w >>= 16; pos += 16; }
    }
    c = __popc(w & 0xFFu);
    if (n >= c) { n -= c; w >>= 8; pos += 8; }
    c = __popc(w & 0xFu);
    if (n >= c) { n -= c; w >>= 4; pos += 4; }
    c = __popc(w & 0x3u);
    if (n >= c) { n -= c; w >>= 2; pos += 2; }
    if (n >= (int)(w & 1u)) pos += 1;
    return pos;
}

// r-th set bit when r < RMAX is tiny (Small): strip the lowest set bit r times, then find-first-set
template <int RMAX>
__device__ __forceinline__ int nth_set_small(uint32_t m, int r) {
#pragma unroll
    for (int i = 0; i < RMAX - 1; ++i)
        if (r > i) m &= m - 1u;
    return __ffs((int)m) - 1;
}

// Geometry as the kernels see it. The variant kernels (GC = 4 / 9 / 16, launched only for exactly
// WarehouseSmall / Medium / Large: variants.py:25-32,40-47,55-62) get every value as a compile-time
// constant (no parameter loads, folded arithmetic, no irregular-rack code); GC = 0 reads KParams.
#ifndef WH_GEO_CONST_LARGE
#define WH_GEO_CONST_LARGE 1
#endif
template <int GC>
struct Geo {
    int dim, L, PP, D, null_pos, invL;
    bool regular;
    __device__ __forceinline__ explicit Geo(const KParams &P) {
        constexpr bool FIXED = GC == 4 || GC == 9 || (GC == 16 && WH_GEO_CONST_LARGE);
        if (GC == 4) { dim = 12; L = 2; }
        else if (GC == 9) { dim = 16; L = 3; }
        else if (FIXED) { dim = 20; L = 4; }
        else { dim = P.dim; L = P.L; }
        PP = FIXED ? 4 * L * L : P.P;
        D = FIXED ? 4 * (dim - 4) : P.D;
        null_pos = FIXED ? dim / 2 : P.null_pos;                               // core.py:107
        invL = FIXED ? (256 + L - 1) / L : P.invL;
        regular = FIXED ? true : (P.regular_racks != 0);
    }
    __device__ __forceinline__ uint32_t null16() const { return (uint32_t)null_pos | ((uint32_t)null_pos << 8); }
};

// core.py:178-188  delivery point d -> cell, v = 2 + d/4, side = d%4: (v,0) (0,v) (v,dim-1) (dim-1,v)
__device__ __forceinline__ uint32_t delivery_cell16(int d, int dim) {
    // sides 0,2 (even): x = v, y = 0 | dim-1      sides 1,3 (odd): x = 0 | dim-1, y = v
    const int v = 2 + (d >> 2), s = d & 3;
    const int edge = (s >> 1) * (dim - 1);
    const bool odd = (s & 1) != 0;
    const int x = odd ? edge : v, y = odd ? v : edge;
    return (uint32_t)x | ((uint32_t)y << 8);
}

// core.py:171-175  pickup point p -> cell: racks[p/(4L)], racks[(p/4)%L], corner p%4 in
// (-1,-1) (0,-1) (-1,0) (0,0)
template <int GC>
__device__ __forceinline__ uint32_t pickup_cell16(const KParams &P, const Geo<GC> &geo, int p) {
    const int q = p >> 2, c = p & 3;
    const int rxi = (q * geo.invL) >> 8, ryi = q - rxi * geo.L;
    const int rx = geo.regular ? 4 * (rxi + 1) : P.racks[rxi];
    const int ry = geo.regular ? 4 * (ryi + 1) : P.racks[ryi];
    return (uint32_t)(rx - 1 + (c & 1)) | ((uint32_t)(ry - 1 + (c >> 1)) << 8);
}

// inverse of pickup_cell16: FIRST pickup index at cell (x,y) or -1 (core.py:319 argmax = first match)
template <int GC>
__device__ __forceinline__ int pickup_index(const KParams &P, const Geo<GC> &geo, int x, int y) {
    if (geo.regular) {  // racks = 4,8,12,... (all reference variants)
        const int qx = (x + 1) >> 2, ox = (x + 1) & 3, qy = (y + 1) >> 2, oy = (y + 1) & 3;
        const bool ok = ox < 2 && oy < 2 && qx >= 1 && qx <= geo.L && qy >= 1 && qy <= geo.L;
        return ok ? 4 * ((qx - 1) * geo.L + (qy - 1)) + ox + 2 * oy : -1;
    }
    int ix = -1, iy = -1, ox = 0, oy = 0;
    for (int i = geo.L - 1; i >= 0; --i) {  // descending so that the smallest matching index wins
        const int r = P.racks[i];
        if (x == r - 1 || x == r) { ix = i; ox = (x == r); }
        if (y == r - 1 || y == r) { iy = i; oy = (y == r); }
    }
    return (ix >= 0 && iy >= 0) ? 4 * (ix * geo.L + iy) + ox + 2 * oy : -1;
}

// A group of G consecutive lanes. GC = compile-time G (0 => runtime G from the launch parameters).
template <int GC>
struct Group {
    int lane, gl;     // lane in warp, lane in group
    int gshift;       // first lane of the group
    int gi;           // group index inside the warp
    int G, epw;       // lanes per group, groups (environments) per warp
    uint32_t gmask;   // this group's lanes
    bool ghost;       // one of the 32 mod G left-over lanes: computes along, owns nothing
    static constexpr bool WIDE = (GC == 0) || (4 * GC > 32);   // pickup masks may need > 32 bits
    static constexpr int PBITS = (GC == 0) ? 64 : (4 * GC <= 16 ? 16 : (4 * GC <= 32 ? 32 : 64));   // pickup-mask width
    static constexpr bool POW2 = GC != 0 && (GC & (GC - 1)) == 0;
    __device__ __forceinline__ explicit Group(int g_runtime) {
        G = GC ? GC : g_runtime;
        epw = 32 / G;
        lane = threadIdx.x & 31;
        gi = lane / G;
        ghost = gi >= epw;
        gshift = ghost ? 0 : gi * G;
        gl = ghost ? lane - epw * G : lane - gshift;
        gmask = (G == 32 ? FULL : ((1u << G) - 1u)) << gshift;
    }
    __device__ __forceinline__ uint32_t ballot(bool p) const {
        return (__ballot_sync(FULL, p) & gmask) >> gshift;
    }
    // does any lane of my group have `hit` set? hit = (mark == c) | (armed & (mm in {rev, ca, cb})):
    // one predicate chain (4 ISETP + 1 PLOP3) feeding the vote instead of 0/1 integers and SELs
    __device__ __forceinline__ bool any_hit(uint32_t mm, uint32_t mark, uint32_t c, uint32_t rev, uint32_t ca,
                                            uint32_t cb, uint32_t armed) const {
        uint32_t b;
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "setp.eq.u32 p, %1, %2;\n\t"
            "setp.eq.or.u32 p, %3, %2, p;\n\t"
            "setp.eq.or.u32 p, %4, %2, p;\n\t"
            "setp.ne.u32 q, %5, 0;\n\t"
            "and.pred p, p, q;\n\t"
            "setp.eq.or.u32 p, %6, %7, p;\n\t"
            "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t}"
            : "=r"(b) : "r"(rev), "r"(mm), "r"(ca), "r"(cb), "r"(armed), "r"(mark), "r"(c));
        return (b & gmask) != 0u;
    }
    __device__ __forceinline__ uint32_t shfl(uint32_t v, int src) const {
        return __shfl_sync(FULL, v, gshift + src);
    }
    __device__ __forceinline__ uint32_t shfl_down1(uint32_t v) const { return __shfl_down_sync(FULL, v, 1); }
    // OR over the group of a per-lane nibble placed at bit 4*gl (natural pickup-index order)
    __device__ __forceinline__ unsigned long long or_nibbles(uint32_t nib) const {
        if (POW2) {
            unsigned long long v = (gl < 16) ? ((unsigned long long)nib << (4 * gl)) : 0ull;
            uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
#pragma unroll
            for (int m = 1; m < (GC ? GC : 1); m <<= 1) {
                lo |= __shfl_xor_sync(FULL, lo, m);
                if (WIDE) hi |= __shfl_xor_sync(FULL, hi, m);
            }
            return (unsigned long long)lo | ((unsigned long long)hi << 32);
        }
        unsigned long long r = 0ull;
        const int n = G < 16 ? G : 16;
#pragma unroll
        for (int s = 0; s < (GC ? (GC < 16 ? GC : 16) : 16); ++s)
            if (s < n) r |= (unsigned long long)shfl(nib, s) << (4 * s);
        return r;
    }
    __device__ __forceinline__ unsigned long long or64(unsigned long long v) const {
        if (POW2) {
            uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
#pragma unroll
            for (int m = 1; m < (GC ? GC : 1); m <<= 1) {
                lo |= __shfl_xor_sync(FULL, lo, m);
                if (WIDE) hi |= __shfl_xor_sync(FULL, hi, m);
            }
            return (unsigned long long)lo | ((unsigned long long)hi << 32);
        }
        unsigned long long r = 0ull;
        for (int s = 0; s < G; ++s) {
            r |= (unsigned long long)shfl((uint32_t)v, s);
            if (WIDE) r |= (unsigned long long)shfl((uint32_t)(v >> 32), s) << 32;
        }
        return r;
    }
    __device__ __forceinline__ uint32_t min_u32(uint32_t v) const {
        if (POW2) {
#pragma unroll
            for (int m = 1; m < (GC ? GC : 1); m <<= 1) v = min(v, __shfl_xor_sync(FULL, v, m));
            return v;
        }
        uint32_t r = 0xffffffffu;
#pragma unroll
        for (int s = 0; s < (GC ? GC : 32); ++s)
            if (s < G) r = min(r, shfl(v, s));
        return r;
    }
    __device__ __forceinline__ int add(int v) const {
        if (POW2) {
#pragma unroll
            for (int m = 1; m < (GC ? GC : 1); m <<= 1) v += (int)__shfl_xor_sync(FULL, (uint32_t)v, m);
            return v;
        }
        int r = 0;
        for (int s = 0; s < G; ++s) r += (int)shfl((uint32_t)v, s);
        return r;
    }
};

// Registers of one lane of one environment.
struct EnvRegs {
    uint32_t pos16;  // my agent's cell (x | y<<8), 0xFFFF for rows >= A
    int atgt;        // my agent's delivery target or -1
    uint32_t pt4;    // delivery targets of my 4 pickup points (0xFF = inactive)
    uint2 tmr;       // their timers as loaded: four int16, t0 | t1 << 16, t2 | t3 << 16 (0xFFFF = -1 = none)
    int time, A, ep;
};

// timer j of a lane := v (16 bits); j is a runtime index into the two packed registers
__device__ __forceinline__ void set_timer(EnvRegs &s, int j, uint32_t v) {
    const uint32_t sh = (uint32_t)(j & 1) * 16u, m = 0xFFFFu << sh, val = (v & 0xFFFFu) << sh;
    if (j < 2) s.tmr.x = (s.tmr.x & ~m) | val;
    else s.tmr.y = (s.tmr.y & ~m) | val;
}

// PP = number of pickup points: a compile-time 4*G in the variant kernels, P.P otherwise
template <int GC, int KEEP = 0>
__device__ __forceinline__ void load_env(const KParams &P, const Group<GC> &g, env_t e, int R, int PP, EnvRegs &s) {
    s.time = ld_state<KEEP>(P.time + e);
    s.A = ld_state<KEEP>(P.num_agents + e);
    s.ep = ld_state<KEEP>(P.episode_ctr + e);
    s.pos16 = 0xFFFFu;
    s.atgt = -1;
    if (g.gl < R) {
        s.pos16 = ld_state<KEEP>(reinterpret_cast<const uint16_t *>(P.agent_pos) + e * R + g.gl);
        s.atgt = ld_state<KEEP>(P.agent_tgt + e * R + g.gl);
    }
    s.pt4 = 0xFFFFFFFFu;
    s.tmr = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    if (4 * g.gl < PP) {
        s.pt4 = ld_state<KEEP, 2>(reinterpret_cast<const uint32_t *>(P.pickup_tgt + e * PP) + g.gl);
        s.tmr = ld_state<KEEP, 1>(reinterpret_cast<const uint2 *>(P.pickup_timer + e * PP) + g.gl);
    }
}

template <int GC, int KEEP = 0>
__device__ __forceinline__ void store_env(const KParams &P, const Group<GC> &g, env_t e, int R, int PP,
                                          const EnvRegs &s, bool store_meta) {
    if (g.gl < R) {
        st_state<KEEP>(reinterpret_cast<uint16_t *>(P.agent_pos) + e * R + g.gl, (uint16_t)s.pos16);
        st_state<KEEP>(P.agent_tgt + e * R + g.gl, (int8_t)s.atgt);
    }
    if (4 * g.gl < PP) {
        st_state<KEEP, 2>(reinterpret_cast<uint32_t *>(P.pickup_tgt + e * PP) + g.gl, s.pt4);
        st_state<KEEP, 1>(reinterpret_cast<uint2 *>(P.pickup_timer + e * PP) + g.gl, s.tmr);
    }
    if (g.gl == 0) {
        st_state<KEEP>(P.time + e, s.time);
        if (store_meta) { st_state<KEEP>(P.num_agents + e, (int8_t)s.A); st_state<KEEP>(P.episode_ctr + e, s.ep); }
    }
}

// 64-bit mask of active requests in natural pickup-index order, replicated in every lane
template <int GC>
__device__ __forceinline__ unsigned long long active_mask(const Group<GC> &g, uint32_t pt4) {
    // byte != 0xFF  <=>  delivery index in 0..63  <=>  bit 7 clear
    const uint32_t inv = ~pt4;
    const uint32_t nib = ((inv >> 7) & 1u) | ((inv >> 14) & 2u) | ((inv >> 21) & 4u) | ((inv >> 28) & 8u);
    return g.or_nibbles(nib);
}

// ---------------------------------------------------------------------------------------------
// S2-S3  sequential, order-dependent, collision-resolved moves — core.py:275-300
// ---------------------------------------------------------------------------------------------
// The reference keeps an occupancy grid and a set of forbidden (from,to) moves. Restated per lane:
//   mark   : the cell whose occupancy bit this agent currently accounts for (NO_CELL if none).
//            occ[c] is True  <=>  some lane has mark == c.  A successful move from p clears the
//            mark of EVERY agent standing on p (core.py:290 clears the bit even if a co-located
//            agent remains) and sets the mover's mark to its new cell (core.py:291).
//   rev/ca/cb : the up-to-three moves this agent's successful move forbids (core.py:294-297): the
//            reverse move and, for a diagonal, the two crossing moves; armed (`moved`) on success.
// One ballot per processed agent answers "is the target marked, or is this move forbidden?".
template <int GC, int RC>
__device__ __forceinline__ void do_moves(const KParams &P, const Group<GC> &g, int R, int A,
                                         int act, int ord, bool have_order, uint32_t &pos16) {
    const Geo<GC> geo(P);
    const int px = pos16 & 0xFF, py = pos16 >> 8;
    uint32_t m = ABSENT_MOVE, rev = ABSENT_MOVE, ca = ABSENT_MOVE, cb = ABSENT_MOVE;
    if (g.gl < A && act >= 0 && act <= 8) {
        const int ax = (act * 11) >> 5;            // act / 3 for 0..8   (MOVES, core.py:38)
        int x = px + ax - 1, y = py + (act - 3 * ax) - 1;
        if ((unsigned)x >= (unsigned)geo.dim) x = px;  // per-axis clamp => wall sliding (core.py:284-287)
        if ((unsigned)y >= (unsigned)geo.dim) y = py;
        const uint32_t to = (uint32_t)x | ((uint32_t)y << 8);
        m = pos16 | (to << 16);                                // bytes [px, py, x, y]
        // the forbidden moves are byte permutations of m: the reverse move [x, y, px, py] (core.py:294)
        // and, for a diagonal, the two crossing moves (x,py)->(px,y) = [x, py, px, y] and
        // (px,y)->(x,py) = [px, y, x, py] (core.py:295-297); otherwise all three are the reverse move
        const bool diag = x != px && y != py;
        rev = __byte_perm(m, 0u, 0x1032u);
        ca = __byte_perm(m, 0u, diag ? 0x3012u : 0x1032u);
        cb = __byte_perm(m, 0u, diag ? 0x1230u : 0x1032u);
    }
    uint32_t mark = (g.gl < A) ? pos16 : NO_CELL;   // core.py:276: every agent marks its cell
    uint32_t moved = 0u;  // arms rev / ca / cb (an ABSENT move is rejected below whatever `hit` says)
    int n_order = R;
    if (have_order) {  // entries after the first -1 are ignored
        const uint32_t neg = g.ballot(g.gl < R && ord < 0);
        n_order = neg ? (__ffs(neg) - 1) : R;
    }
    const int RR = RC ? RC : R;
    if (have_order) {
#pragma unroll
        for (int t = 0; t < RR; ++t) {
            int cur = (int)g.shfl((uint32_t)ord, t);
            if (t >= n_order || cur >= R) cur = -1;
            uint32_t mm = g.shfl(m, cur < 0 ? 0 : cur);
            if (cur < 0) mm = ABSENT_MOVE;
            const uint32_t c = mm >> 16, from = mm & 0xFFFFu;
            const bool ok = !g.any_hit(mm, mark, c, rev, ca, cb, moved) && (mm != ABSENT_MOVE);   // core.py:289
            if (ok && mark == from) mark = NO_CELL;                            // core.py:290
            if (ok && g.gl == cur) { mark = c; moved = 1u; }                   // core.py:291-297
        }
    } else {
#pragma unroll
        for (int t = 0; t < RR; ++t) {                                         // ascending agent ids
            const uint32_t mm = g.shfl(m, t);
            const uint32_t c = mm >> 16, from = mm & 0xFFFFu;
            const bool ok = !g.any_hit(mm, mark, c, rev, ca, cb, moved) && (mm != ABSENT_MOVE);   // core.py:289
            if (ok && mark == from) mark = NO_CELL;                            // core.py:290
            if (ok && g.gl == t) { mark = c; moved = 1u; }                     // core.py:291-297
        }
    }
    if (moved) pos16 = m >> 16;                                                // core.py:299-300
}

// ---------------------------------------------------------------------------------------------
// S4-S8  expiry, pickups, respawn, deliveries — core.py:303-368
// ---------------------------------------------------------------------------------------------
struct StepOut {
    float reward;
    unsigned long long active;  // active-request mask after respawn
    uint32_t tpos16;            // delivery-target cell for the observation (null cell if none)
    int npick, ndeliv, nexp;    // per-env event counts of this step
};

template <int GC>
__device__ __forceinline__ StepOut do_world(const KParams &P, const Group<GC> &g, env_t e, int R,
                                            uint32_t env_id, EnvRegs &s, bool replay,
                                            unsigned long long active_in = 0ull, bool have_active = false) {
    // active_in / have_active: the active-request mask as it was BEFORE this step, when the caller already has
    // it (the in-kernel solver computed it): the mask after expiry and pickups is then derived from it instead
    // of being gathered from the lanes a second time.
    StepOut o;
    const Geo<GC> geo(P);
    // ---- core.py:303-306 expiry (before pickup detection) ----
    // The four timers stay packed as loaded. Point j is active iff bit 7 of byte j of pt4 is clear;
    // -1 is added to the active halves with the packed 16-bit add (VIADD.16x2: no borrow between the
    // halves), and a half that reached 0 — found with the packed unsigned minimum — has expired.
    int nexp = 0;
    bool expired_any;
    {
        const uint32_t inv = ~s.pt4;
        const uint32_t d0 = ((inv >> 7) & 1u) * 0xFFFFu + ((inv >> 15) & 1u) * 0xFFFF0000u;
        const uint32_t d1 = ((inv >> 23) & 1u) * 0xFFFFu + ((inv >> 31) & 1u) * 0xFFFF0000u;
        s.tmr.x = __vadd2(s.tmr.x, d0);
        s.tmr.y = __vadd2(s.tmr.y, d1);
        const bool any0 = (__vminu2(s.tmr.x, 0x00010001u) & __vminu2(s.tmr.y, 0x00010001u)) != 0x00010001u;
        expired_any = __any_sync(FULL, any0);
        if (expired_any) {                          // rare: a request expires at most once per episode
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t t = ((j < 2 ? s.tmr.x : s.tmr.y) >> (16 * (j & 1))) & 0xFFFFu;
                if (t == 0u) { s.pt4 |= 0xFFu << (8 * j); set_timer(s, j, 0xFFFFu); ++nexp; }
            }
            o.nexp = g.add(nexp);
        } else {
            o.nexp = 0;
        }
    }
    // ---- core.py:309-335 pickups: agent on a pickup cell, free, request waiting there ----
    const int x = s.pos16 & 0xFF, y = s.pos16 >> 8;
    const int cand = (g.gl < s.A) ? pickup_index(P, geo, x, y) : -1;
    const uint32_t w = g.shfl(s.pt4, (cand < 0 ? 0 : cand) >> 2);
    const int tg = (int)(int8_t)((w >> (8 * (cand & 3))) & 0xFFu);
    const bool picks = cand >= 0 && s.atgt == -1 && tg > -1;
    float reward = 0.0f;
    o.npick = 0;
    unsigned long long served = 0ull;
    if (__any_sync(FULL, picks)) {                  // rare with random actions, every other step with the greedy solver
        // every group walks its own pickers (usually one); a picker's point belongs to lane cand / 4
        const uint32_t pm = g.ballot(picks);
        o.npick = __popc(pm);
        for (uint32_t rem = pm; __any_sync(FULL, rem != 0u); rem &= rem - 1u) {
            const int c = (int)g.shfl((uint32_t)cand, rem ? __ffs((int)rem) - 1 : 0);
            if (rem != 0u) {
                served |= 1ull << c;
                if ((c >> 2) == g.gl) {                                         // core.py:330-331
                    const int j = c & 3;
                    s.pt4 |= 0xFFu << (8 * j);
                    set_timer(s, j, 0xFFFFu);
                }
            }
        }
        if (picks) { s.atgt = tg; reward = 1.0f; }                             // core.py:327-329,335
    }
    // ---- core.py:338-351 respawn until exactly R requests are active ----
    unsigned long long active = (have_active && !expired_any) ? (active_in & ~served) : active_mask(g, s.pt4);
    int sp = -1, st_ = -1, k;
    uint32_t up = 0, ut = 0;
    if (replay) {
        if (g.gl < R) { sp = P.spawn_p[e * R + g.gl]; st_ = P.spawn_t[e * R + g.gl]; }
        const uint32_t neg = g.ballot(sp < 0);       // lanes >= R hold -1
        k = neg ? (__ffs(neg) - 1) : g.G;
    } else {
        k = R - __popcll(active);
        if (k < 0) k = 0;
    }
    if (__any_sync(FULL, k > 0)) {
        if (!replay) philox4x32_10(env_id, (uint32_t)s.ep, (uint32_t)s.time, (uint32_t)g.gl, P.seed, up, ut);
        const unsigned long long pmask = (geo.PP >= 64) ? ~0ull : ((1ull << geo.PP) - 1ull);
        unsigned long long inactive = ~active & pmask;
        unsigned long long avail_d = (geo.D >= 64) ? ~0ull : ((1ull << geo.D) - 1ull);
        const int n_inact = __popcll(inactive);
        for (int i = 0; __any_sync(FULL, i < k); ++i) {
            int p, d;
            if (replay) {
                p = (int)g.shfl((uint32_t)sp, i);
                d = (int)g.shfl((uint32_t)st_, i);
            } else {
                const uint32_t a = g.shfl(up, i), b = g.shfl(ut, i);
                const int ni = n_inact - i, di = geo.D - i;
                p = nth_set64<Group<GC>::PBITS>(inactive, (int)bounded(a, (uint32_t)(ni > 0 ? ni : 1)));
                d = nth_set64<64>(avail_d, (int)bounded(b, (uint32_t)(di > 0 ? di : 1)));
            }
            if (i < k && p >= 0) {
                inactive &= ~(1ull << p);
                avail_d &= ~(1ull << d);
                active |= 1ull << p;
                if ((p >> 2) == g.gl) {                                       // core.py:344,351
                    const int j = p & 3;
                    s.pt4 = (s.pt4 & ~(0xFFu << (8 * j))) | (((uint32_t)d & 0xFFu) << (8 * j));
                    set_timer(s, j, (uint32_t)P.wait);
                }
            }
        }
    }
    // ---- core.py:354-368 deliveries (an agent that picked up THIS step is already delivering) ----
    o.tpos16 = geo.null16();
    bool delivered = false;
    if (g.gl < s.A && s.atgt > -1) {
        const uint32_t dcell = delivery_cell16(s.atgt, geo.dim);
        delivered = dcell == s.pos16;
        if (delivered) { s.atgt = -1; reward += 1.0f; }
        else o.tpos16 = dcell;
    }
    o.ndeliv = __any_sync(FULL, delivered) ? __popc(g.ballot(delivered)) : 0;
    o.reward = reward;
    o.active = active;
    return o;
}

// ---------------------------------------------------------------------------------------------
// S9  observation build — core.py:371-432 (step flavour), core.py:224-260 (reset flavour)
// ---------------------------------------------------------------------------------------------
// Lane r holds padded row r of the three per-agent tables (position, availability, delivery-target
// position) plus row r+1 (one shuffle), so "table with row a deleted" is a per-lane select; lane r
// also materialises request r (r-th set bit of the active mask -> pickup cell, delivery cell).
// requests: every 128-bit store instruction writes 16R contiguous bytes per environment straight from
// registers. other_*: one environment's block of a key is assembled in shared memory and streamed
// out with 128-bit stores (st.global.cs: written once, never re-read here).
// Shared-memory staging area of one environment while its observation is written (compile-time R).
#ifndef WH_FLAT_UNITS
#define WH_FLAT_UNITS 1     // Small flat copy-out: per-lane fixed sources over 2-environment units (tuning builds: 0)
#endif
#ifndef WH_COOP_FAST
#define WH_COOP_FAST 1      // Medium copy-out: the all-environments-live fast path (tuning builds: 0)
#endif
#ifndef WH_COOP_MEDIUM
#define WH_COOP_MEDIUM 1
#endif
template <int RC>
struct ObsStage {
    static constexpr int ROWS = RC * (RC - 1);          // "other" rows of one env
    static constexpr int POS_BYTES = 8 * ROWS;          // other_positions / other_delivery_targets
    static constexpr int AV_BYTES = ROWS;               // other_availabilities
    static constexpr int RAW = RC ? ((POS_BYTES + AV_BYTES + 15) / 16) * 16 : 16;
    // Warp-cooperative copy-out (Medium, see build_obs): the warp's EPW environments are staged back to
    // back per key — [EPW][R][R-1] int2 | [EPW][R][R-1] int8 | [EPW][R] int4 request rows — so one slot
    // is 1/EPW of that area.
    static constexpr bool COOP = RC == 9 && WH_COOP_MEDIUM;
    static constexpr int REQ_BYTES = 16 * RC;
    static constexpr int BYTES = COOP ? ((POS_BYTES + AV_BYTES + REQ_BYTES + 15) / 16) * 16 : RAW;
};

// PART: which keys this caller writes — bit 0: num_agents, self_*, requests; bit 1: other_* (k_multi_ws splits
// one environment's observation between two warps; everybody else writes all of it)
// FAST (Medium copy-out only): take the all-environments-live fast path. It executes ~100 fewer instructions per
// warp, which pays where the kernel is issue-bound — with the in-kernel solver: k_step Medium 65 536 greedy
// 0.866 -> 0.892, k_multi 31.2 -> 29.4 us per step — and costs where it is bound by the write stream (the stores
// of a warp then issue in one burst: k_step random actions -0.2 .. -0.8 %, k_multi open-loop per-step slices
// -14 %), so the callers enable it for the GREEDY instantiations only (same-box A/B, profiles/README.md).
template <int GC, int RC, int PART = 3, bool FAST = false>
__device__ __forceinline__ void build_obs(const KParams &P, const wh_obs &o, const Group<GC> &g, env_t e, int R,
                                          const EnvRegs &s, unsigned long long active, uint32_t tpos16,
                                          int flavour, bool live, unsigned char *stage,
                                          unsigned char *wstage = nullptr, env_t env0 = 0) {
    // o: where the observation goes (P.obs, or a per-step slice of [T,N,...] tensors in k_multi);
    // stage: this environment's staging slot; wstage / env0: the warp's whole staging area and its
    // first environment (used by the warp-cooperative copy-out only)
    const Geo<GC> geo(P);
    const int null_pos = geo.null_pos;
    const uint32_t null16 = geo.null16();
    const bool real = g.gl < s.A;
    const bool delivering = real && s.atgt > -1;
    // core.py:372-407 padded tables (reset flavour: availability 0 and null targets, core.py:233-236)
    const uint32_t ppos = real ? s.pos16 : null16;
    const uint32_t avail = (flavour == WH_OBS_STEP && real && !delivering) ? 1u : 0u;
    const uint32_t tpos = (flavour == WH_OBS_STEP && delivering) ? tpos16 : null16;
    const uint32_t mine = (ppos & 0x7Fu) | (((ppos >> 8) & 0x7Fu) << 7) | ((tpos & 0x7Fu) << 14) |
                          (((tpos >> 8) & 0x7Fu) << 21) | (avail << 28);
    const uint32_t next = g.shfl_down1(mine);
    const int2 my_p = make_int2(mine & 0x7F, (mine >> 7) & 0x7F), nx_p = make_int2(next & 0x7F, (next >> 7) & 0x7F);
    const int2 my_t = make_int2((mine >> 14) & 0x7F, (mine >> 21) & 0x7F),
               nx_t = make_int2((next >> 14) & 0x7F, (next >> 21) & 0x7F);
    const int my_a = (mine >> 28) & 1, nx_a = (next >> 28) & 1;

    // core.py:409-418 request list: active pickup points in ascending index, [px,py,dx,dy]
    int4 rq = make_int4(null_pos, null_pos, null_pos, null_pos);  // only if < R active (unreachable)
    if constexpr ((PART & 1) != 0) {
        const bool have = g.gl < R && g.gl < __popcll(active);
        const int p = ((RC != 0 && RC <= 4) ? nth_set_small<(RC ? RC : 1)>((uint32_t)active, have ? g.gl : 0)
                                            : nth_set64<Group<GC>::PBITS>(active, have ? g.gl : 0)) & 63;
        const uint32_t w4 = g.shfl(s.pt4, p >> 2);
        if (have) {
            const uint32_t pc = pickup_cell16(P, geo, p);
            const uint32_t dc = delivery_cell16((int)((w4 >> (8 * (p & 3))) & 0x3Fu), geo.dim);
            rq = make_int4(pc & 0xFF, pc >> 8, dc & 0xFF, dc >> 8);
        }
    }

    const uint32_t row0 = e * (uint32_t)R;
    if ((PART & 1) != 0 && live && g.gl < R) {
        o.num_agents[row0 + g.gl] = s.A;
        reinterpret_cast<int2 *>(o.self_position)[row0 + g.gl] = my_p;
        o.self_availability[row0 + g.gl] = (int8_t)my_a;
        reinterpret_cast<int2 *>(o.self_delivery_target)[row0 + g.gl] = my_t;
    }
    int4 *orq = reinterpret_cast<int4 *>(o.requests) + row0 * R + g.gl;
    // core.py:428 quirk: in step() other_delivery_targets always drops row 1 (reset drops row i)
    const int2 t_fixed = (g.gl >= 1) ? nx_t : my_t;
    const int RR = RC ? RC : R;

    if constexpr (RC != 0 && ObsStage<RC>::COOP) {
        // Warp-cooperative copy-out (Medium: 3 environments per warp, 144-byte rows). Written group by
        // group, every store instruction would emit three separate 144-byte runs = 4.5 sectors each,
        // half of them starting mid-sector: 98 sectors per environment for 84.7 sectors of data, and the
        // SM->L2 write path (not HBM) bounds the observation build. The three environments of a warp are
        // consecutive, so each key's block for the warp is ONE contiguous range: stage it in that
        // order and let all 32 lanes stream it out, 512 contiguous bytes per store instruction.
        using St = ObsStage<RC>;
        constexpr int EPW = 32 / GC;
        constexpr int ROWS4 = St::ROWS / 2;                        // int4 per env of an [R][R-1] int2 block
        int2 *w_pos = reinterpret_cast<int2 *>(wstage);                                   // [EPW][R][R-1]
        int8_t *w_av = reinterpret_cast<int8_t *>(wstage + EPW * St::POS_BYTES);          // [EPW][R][R-1]
        int4 *w_req = reinterpret_cast<int4 *>(wstage + EPW * (St::POS_BYTES + St::AV_BYTES + 8) / 16 * 16);  // [EPW][R]
        const int gi = g.ghost ? 0 : g.gi;
        const bool writer = !g.ghost && g.gl < RC - 1;
        // which of the warp's environments are written at all (envs past N, and envs a masked reset
        // leaves alone, keep their observations): bit k = environment env0 + k
        const uint32_t lm = __ballot_sync(FULL, live);
        auto env_live = [&](int env) { return ((lm >> (env * GC)) & 1u) != 0u; };
        if ((PART & 1) != 0 && !g.ghost && g.gl < RC) w_req[gi * RC + g.gl] = rq;           // core.py:409-418
#pragma unroll
        for (int a = 0; a < RC; ++a) {
            if ((PART & 2) != 0 && writer) {
                const bool sh = g.gl >= a;                                      // core.py:426-427
                w_pos[gi * St::ROWS + a * (RC - 1) + g.gl] = sh ? nx_p : my_p;
                w_av[gi * St::ROWS + a * (RC - 1) + g.gl] = (int8_t)(sh ? nx_a : my_a);
            }
        }
        __syncwarp();
        const int lane = g.lane;
        // Fast path (every step of a full batch): all EPW environments of the warp are written. No liveness
        // tests, and the request row / environment of every element follow from compile-time constants:
        // element i = u + 32k (u = lane - mis) has row (u + 5k) mod 9 and environment i / 81.
        constexpr uint32_t ALL_LIVE = 1u | (1u << GC) | (1u << (2 * GC));
        if (WH_COOP_FAST && FAST && (lm & ALL_LIVE) == ALL_LIVE && !__any_sync(FULL, flavour != WH_OBS_STEP)) {
            static_assert(EPW == 3 && RC == 9, "index arithmetic below is written for 3 x 9");
            if constexpr ((PART & 1) != 0) {
                constexpr int PER_ENV = RC * RC;
                const uint32_t base4 = env0 * (uint32_t)PER_ENV;
                const int u = lane - (int)(base4 & 1u);                         // -1 .. 31
                const int u9 = u + 9, r0 = u9 - RC * ((u9 * 57) >> 9);          // (u + 9) mod 9
                int4 *d = reinterpret_cast<int4 *>(o.requests) + base4 + u;
                const int4 *src = w_req + r0;
#pragma unroll
                for (int k = 0; k < (EPW * PER_ENV + 1 + 31) / 32; ++k) {
                    constexpr int K32 = 32;
                    const int ck = (5 * k) % 9;                                 // 32k mod 9
                    const int i = u + K32 * k;
                    // row r0 + ck, minus 9 on wrap-around: the unsigned minimum of the two candidates
                    const int off = (int)min((uint32_t)(r0 + ck), (uint32_t)(r0 + ck - RC)) - r0;
                    const int env = (i >= PER_ENV) + (i >= 2 * PER_ENV);
                    if ((k > 0 || i >= 0) && (K32 * k + 31 < EPW * PER_ENV || i < EPW * PER_ENV))
                        WH_ST(d + K32 * k, src[env * RC + off]);
                }
            }
            if constexpr ((PART & 2) == 0) return;
            {   // other_positions [N,R,R-1,2]: staged in output order
                int4 *d = reinterpret_cast<int4 *>(o.other_positions) + env0 * (uint32_t)ROWS4 + lane;
                const int4 *src = reinterpret_cast<const int4 *>(w_pos) + lane;
#pragma unroll
                for (int k = 0; k < (EPW * ROWS4 + 31) / 32; ++k)
                    if (32 * k + 31 < EPW * ROWS4 || lane + 32 * k < EPW * ROWS4) WH_ST(d + 32 * k, src[32 * k]);
            }
            if (lane < EPW * (St::ROWS / 8))   // other_availabilities [N,R,R-1]: 72 bytes per env = 9 int2
                WH_ST(reinterpret_cast<int2 *>(o.other_availabilities + (size_t)env0 * St::ROWS) + lane,
                      reinterpret_cast<const int2 *>(w_av)[lane]);
            __syncwarp();
            // core.py:428: every agent's block is the same (R-1)-row table; 8 rows = 4 int4 per env
            if (writer) w_pos[gi * (RC - 1) + g.gl] = t_fixed;
            __syncwarp();
            int4 *d_t = reinterpret_cast<int4 *>(o.other_delivery_targets) + env0 * (uint32_t)ROWS4 + lane;
            static_assert(ROWS4 % 4 == 0 && 32 % 4 == 0, "table period");
            const int4 *tsrc = reinterpret_cast<const int4 *>(w_pos) + (lane & 3);
#pragma unroll
            for (int k = 0; k < (EPW * ROWS4 + 31) / 32; ++k) {
                const int i = lane + 32 * k;
                const int env = (i >= ROWS4) + (i >= 2 * ROWS4);
                if (32 * k + 31 < EPW * ROWS4 || i < EPW * ROWS4) WH_ST(d_t + 32 * k, tsrc[env * ((RC - 1) / 2)]);
            }
            return;
        }
        if constexpr ((PART & 1) != 0) {   // requests [N,R,R,4]: R copies of the env's R request rows (core.py:429)
            constexpr int PER_ENV = RC * RC;                                    // int4 per env
            const uint32_t base4 = env0 * (uint32_t)PER_ENV;
            const int mis = (int)(base4 & 1u);                                  // tile starts mid-sector
            int4 *d = reinterpret_cast<int4 *>(o.requests) + base4;
#pragma unroll
            for (int k = 0; k < (EPW * PER_ENV + 1 + 31) / 32; ++k) {
                const int i = lane + 32 * k - mis;
                const int env = (i >= PER_ENV) + (i >= 2 * PER_ENV);
                if (i >= 0 && i < EPW * PER_ENV && env_live(env)) {
                    const int r = i - RC * ((i * 57) >> 9);                     // i % 9 for i < 256
                    static_assert(EPW == 3 && RC == 9, "index arithmetic below is written for 3 x 9");
                    WH_ST(d + i, w_req[env * RC + r]);
                }
            }
        }
        if constexpr ((PART & 2) == 0) return;
        {   // other_positions [N,R,R-1,2]
            int4 *d = reinterpret_cast<int4 *>(o.other_positions) + env0 * (uint32_t)ROWS4;
#pragma unroll
            for (int k = 0; k < (EPW * ROWS4 + 31) / 32; ++k) {
                const int i = lane + 32 * k;
                const int env = (i >= ROWS4) + (i >= 2 * ROWS4);
                if (i < EPW * ROWS4 && env_live(env)) WH_ST(d + i, reinterpret_cast<const int4 *>(w_pos)[i]);
            }
        }
        {   // other_availabilities [N,R,R-1]: 72 bytes per env = 9 int2
            int2 *d = reinterpret_cast<int2 *>(o.other_availabilities + (size_t)env0 * St::ROWS);
            const int env = (lane >= St::ROWS / 8) + (lane >= 2 * (St::ROWS / 8));
            if (lane < EPW * (St::ROWS / 8) && env_live(env)) WH_ST(d + lane, reinterpret_cast<const int2 *>(w_av)[lane]);
        }
        __syncwarp();
        int4 *d_t = reinterpret_cast<int4 *>(o.other_delivery_targets) + env0 * (uint32_t)ROWS4;
        // the flavour is per environment (an env that was just auto-reset shows its reset observation),
        // the copy-out is per warp: take the compact path only if the whole warp is in step flavour
        if (!__any_sync(FULL, flavour != WH_OBS_STEP)) {
            // core.py:428: every agent's block is the same (R-1)-row table; 8 rows = 4 int4 per env
            if (writer) w_pos[gi * (RC - 1) + g.gl] = t_fixed;
            __syncwarp();
#pragma unroll
            for (int k = 0; k < (EPW * ROWS4 + 31) / 32; ++k) {
                const int i = lane + 32 * k;
                const int env = (i >= ROWS4) + (i >= 2 * ROWS4);
                static_assert(ROWS4 % 4 == 0, "table period");
                if (i < EPW * ROWS4 && env_live(env)) WH_ST(d_t + i, reinterpret_cast<const int4 *>(w_pos)[env * ((RC - 1) / 2) + (i & 3)]);
            }
        } else {
#pragma unroll
            for (int a = 0; a < RC; ++a)
                if (writer)                                                     // core.py:428 (step) / :256 (reset)
                    w_pos[gi * St::ROWS + a * (RC - 1) + g.gl] = (flavour == WH_OBS_STEP) ? t_fixed : ((g.gl >= a) ? nx_t : my_t);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < (EPW * ROWS4 + 31) / 32; ++k) {
                const int i = lane + 32 * k;
                const int env = (i >= ROWS4) + (i >= 2 * ROWS4);
                if (i < EPW * ROWS4 && env_live(env)) WH_ST(d_t + i, reinterpret_cast<const int4 *>(w_pos)[i]);
            }
        }
        return;
    }
    if constexpr (RC != 0) {
        // The "other_*" keys have 8-byte / 1-byte rows: written straight from registers they would
        // leave partially filled 32-byte sectors on the SM->L2 path. Stage one environment's
        // [R, R-1, ...] block in shared memory and stream it out with 128-bit / 32-bit stores.
        using St = ObsStage<RC>;
        int2 *s_pos = reinterpret_cast<int2 *>(stage);
        int8_t *s_av = reinterpret_cast<int8_t *>(stage + St::POS_BYTES);
        int4 *d_pos = reinterpret_cast<int4 *>(o.other_positions) + row0 * (RC - 1) / 2;
        int4 *d_tgt = reinterpret_cast<int4 *>(o.other_delivery_targets) + row0 * (RC - 1) / 2;
        int32_t *d_av = reinterpret_cast<int32_t *>(o.other_availabilities + row0 * (RC - 1));
        const bool writer = !g.ghost && g.gl < RC - 1;   // ghost lanes shadow group 0: keep them off its stage
#pragma unroll
        for (int a = 0; a < RC; ++a) {
            if ((PART & 1) != 0 && live && g.gl < RC) WH_ST(orq + a * RC, rq);  // core.py:429
            if ((PART & 2) != 0 && writer) {
                const bool sh = g.gl >= a;                                      // core.py:426-427
                s_pos[a * (RC - 1) + g.gl] = sh ? nx_p : my_p;
                s_av[a * (RC - 1) + g.gl] = (int8_t)(sh ? nx_a : my_a);
            }
        }
        if constexpr ((PART & 2) == 0) return;
        __syncwarp();
        if (live) {
#pragma unroll
            for (int k = 0; k < (St::ROWS / 2 + GC - 1) / GC; ++k) {
                const int i = g.gl + k * GC;
                if (i < St::ROWS / 2) WH_ST(d_pos + i, reinterpret_cast<const int4 *>(stage)[i]);
            }
#pragma unroll
            for (int k = 0; k < (St::ROWS / 4 + GC - 1) / GC; ++k) {
                const int i = g.gl + k * GC;
                if (i < St::ROWS / 4) WH_ST(d_av + i, reinterpret_cast<const int32_t *>(s_av)[i]);
            }
        }
        __syncwarp();
        if (flavour == WH_OBS_STEP) {
            // every agent's other_delivery_targets block is the same (R-1)-row table (core.py:428):
            // stage two copies (= R-1 whole int4) and stream them out R/2 times over
            if (writer) { s_pos[g.gl] = t_fixed; s_pos[RC - 1 + g.gl] = t_fixed; }
            __syncwarp();
            if (live) {
#pragma unroll
                for (int k = 0; k < (St::ROWS / 2 + GC - 1) / GC; ++k) {
                    const int i = g.gl + k * GC;
                    // i mod (RC-1): GC == RC here, so i = gl + k (mod RC-1) and gl + k < 2(RC-1)
                    int src = g.gl + k * (GC % (RC - 1));
                    if (src >= RC - 1) src -= RC - 1;
                    if (src >= RC - 1) src -= RC - 1;
                    if (i < St::ROWS / 2) WH_ST(d_tgt + i, reinterpret_cast<const int4 *>(stage)[src]);
                }
            }
        } else {
#pragma unroll
            for (int a = 0; a < RC; ++a)
                if (writer) s_pos[a * (RC - 1) + g.gl] = (g.gl >= a) ? nx_t : my_t;   // core.py:256
            __syncwarp();
            if (live) {
#pragma unroll
                for (int k = 0; k < (St::ROWS / 2 + GC - 1) / GC; ++k) {
                    const int i = g.gl + k * GC;
                    if (i < St::ROWS / 2) WH_ST(d_tgt + i, reinterpret_cast<const int4 *>(stage)[i]);
                }
            }
        }
        return;
    }
    if (!live) return;
    int2 *op = reinterpret_cast<int2 *>(o.other_positions) + row0 * (R - 1) + g.gl;
    int2 *ot = reinterpret_cast<int2 *>(o.other_delivery_targets) + row0 * (R - 1) + g.gl;
    int8_t *oa = o.other_availabilities + row0 * (R - 1) + g.gl;
    for (int a = 0; a < RR; ++a) {
        if ((PART & 2) != 0 && g.gl < R - 1) {
            const bool sh = g.gl >= a;                                          // core.py:426-427
            WH_ST(op + a * (R - 1), sh ? nx_p : my_p);
            WH_ST(oa + a * (R - 1), (int8_t)(sh ? nx_a : my_a));
            WH_ST(ot + a * (R - 1), flavour == WH_OBS_STEP ? t_fixed : (sh ? nx_t : my_t));
        }
        if ((PART & 1) != 0 && g.gl < R) WH_ST(orq + a * R, rq);               // core.py:429
    }
}

// ---------------------------------------------------------------------------------------------
// RLlib-flattened observations: float32 [N, R, 9R+1], keys in RLlib's Dict
// flattening order (alphabetical): num_agents(1) other_availabilities(R-1)
// other_delivery_targets(2(R-1)) other_positions(2(R-1)) requests(4R) self_availability(1)
// self_delivery_target(2) self_position(2). Same values as build_obs (core.py:371-432 / 224-260).
// ---------------------------------------------------------------------------------------------
// Every float of an environment's [R, 9R+1] block is a copy of one entry of a small per-env VALUE TABLE
//   V = [ num_agents | availability[R] | delivery-target cell[R][2] | position[R][2] | requests[R][4] ]
// (9R+1 entries, the padded tables of core.py:372-418), and WHICH entry is a static function of the
// position in the block (row a, column f): the "table without row a" of core.py:424-427 is a +1 index
// shift behind the deleted row, core.py:428 always deletes row 1. In the reset flavour every delivery
// target is the null cell and every availability 0 (core.py:233-236), so the same map serves both.
// FlatMap<RC> is that map, built at compile time; entries are SWIZZLED table addresses (see swz).
template <int RC>
struct FlatMap {
    static constexpr int R = RC ? RC : 1;
    static constexpr int F = 9 * R + 1;                 // floats per agent row
    static constexpr int RF = R * F;                    // floats per environment (always even)
    static constexpr int S4 = (F + 3) / 4;              // table words per swizzle plane
    static constexpr int VS = 4 * S4;                   // table words per environment
    // Table entry s lives at word (s >> 2) + (s & 3) * S4: a lane that copies 4 consecutive floats reads
    // entries c + 4*lane + j (j = 0..3), i.e. for fixed j CONSECUTIVE words across the warp — no bank
    // conflicts inside a segment.
    __host__ __device__ static constexpr int swz(int s) { return (s >> 2) + (s & 3) * S4; }
    __host__ __device__ static constexpr int src(int a, int f) {
        if (f == 0) return 0;                                                              // num_agents
        if (f < R) { const int j = f - 1; return 1 + j + (j >= a ? 1 : 0); }               // other_availabilities
        if (f < 3 * R - 2) { const int j2 = f - R, j = j2 >> 1; return 1 + R + 2 * (j + (j >= 1 ? 1 : 0)) + (j2 & 1); }  // core.py:428
        if (f < 5 * R - 4) { const int j2 = f - (3 * R - 2), j = j2 >> 1; return 1 + 3 * R + 2 * (j + (j >= a ? 1 : 0)) + (j2 & 1); }
        if (f < 9 * R - 4) return 1 + 5 * R + (f - (5 * R - 4));                            // requests
        if (f == 9 * R - 4) return 1 + a;                                                  // self_availability
        if (f < 9 * R - 1) return 1 + R + 2 * a + (f - (9 * R - 3));                        // self_delivery_target
        return 1 + 3 * R + 2 * a + (f - (9 * R - 1));                                      // self_position
    }
    alignas(16) uint8_t m[RF + 8];
    constexpr FlatMap() : m{} {
        for (int a = 0; a < R; ++a)
            for (int f = 0; f < F; ++f) m[a * F + f] = (uint8_t)swz(src(a, f));
    }
};
__device__ const FlatMap<4> d_flat_map4{};
__device__ const FlatMap<9> d_flat_map9{};
__device__ const FlatMap<16> d_flat_map16{};
template <int RC>
__device__ __forceinline__ const uint8_t *flat_map() {
    if constexpr (RC == 4) return d_flat_map4.m;
    else if constexpr (RC == 9) return d_flat_map9.m;
    else return d_flat_map16.m;
}

// The same map composed with the warp layout: entry k (k-th float of the warp's contiguous output range of
// 32/G environments) = word offset of its source inside the warp's staging area (env * VS + swizzled table
// address). One 64-bit load yields the four sources of a 128-bit store; no per-element index arithmetic.
template <int RC>
struct FlatWarpMap {
    using M = FlatMap<RC>;
    static constexpr int EPW = 32 / (RC ? RC : 1);
    static constexpr int TOT = EPW * M::RF;
    alignas(16) uint16_t w[TOT + 8];
    constexpr FlatWarpMap() : w{} {
        for (int env = 0; env < EPW; ++env)
            for (int a = 0; a < M::R; ++a)
                for (int f = 0; f < M::F; ++f)
                    w[env * M::RF + a * M::F + f] = (uint16_t)(env * M::VS + M::swz(M::src(a, f)));
    }
};
__device__ const FlatWarpMap<4> d_flat_wmap4{};
__device__ const FlatWarpMap<9> d_flat_wmap9{};
__device__ const FlatWarpMap<16> d_flat_wmap16{};
template <int RC>
__device__ __forceinline__ const uint16_t *flat_warp_map() {
    if constexpr (RC == 4) return d_flat_wmap4.w;
    else if constexpr (RC == 9) return d_flat_wmap9.w;
    else return d_flat_wmap16.w;
}

template <int RC>
struct FlatStage {
    static constexpr int BYTES = RC ? FlatMap<RC>::VS * 4 : 16;            // one environment's value table
};

template <int GC, int RC>
__device__ __forceinline__ void build_obs_flat(const KParams &P, const Group<GC> &g, env_t e, int R,
                                               const EnvRegs &s, unsigned long long active, uint32_t tpos16,
                                               int flavour, bool live, float *out, float *stage,
                                               float *wstage = nullptr, env_t env0 = 0) {
    // stage: this environment's value table; wstage / env0: the warp's tables and its first environment
    const Geo<GC> geo(P);
    const int null_pos = geo.null_pos;
    const uint32_t null16 = geo.null16();
    const bool real = g.gl < s.A;
    const bool delivering = real && s.atgt > -1;
    const uint32_t ppos = real ? s.pos16 : null16;
    const uint32_t avail = (flavour == WH_OBS_STEP && real && !delivering) ? 1u : 0u;
    const uint32_t tpos = (flavour == WH_OBS_STEP && delivering) ? tpos16 : null16;
    const bool have = g.gl < R && g.gl < __popcll(active);
    const int p = ((RC != 0 && RC <= 4) ? nth_set_small<(RC ? RC : 1)>((uint32_t)active, have ? g.gl : 0)
                                        : nth_set64<Group<GC>::PBITS>(active, have ? g.gl : 0)) & 63;
    const uint32_t w4 = g.shfl(s.pt4, p >> 2);
    float4 rq = make_float4((float)null_pos, (float)null_pos, (float)null_pos, (float)null_pos);
    if (have) {
        const uint32_t pc = pickup_cell16(P, geo, p);
        const uint32_t dc = delivery_cell16((int)((w4 >> (8 * (p & 3))) & 0x3Fu), geo.dim);
        rq = make_float4((float)(pc & 0xFF), (float)(pc >> 8), (float)(dc & 0xFF), (float)(dc >> 8));
    }
    const float n_agents = (float)s.A;
    if constexpr (RC != 0) {
        using M = FlatMap<RC>;
        constexpr int EPW = 32 / GC;
        // ---- the environment's value table (each lane contributes the entries of its own row) ----
        if (!g.ghost && g.gl < RC) {
            if (g.gl == 0) stage[M::swz(0)] = n_agents;
            const int sa = 1 + g.gl, st_ = 1 + RC + 2 * g.gl, sp = 1 + 3 * RC + 2 * g.gl, sr = 1 + 5 * RC + 4 * g.gl;
            auto at = [&](int i) -> float & { return stage[(i >> 2) + (i & 3) * M::S4]; };
            at(sa) = (float)avail;
            at(st_) = (float)(tpos & 0xFF); at(st_ + 1) = (float)(tpos >> 8);
            at(sp) = (float)(ppos & 0xFF); at(sp + 1) = (float)(ppos >> 8);
            at(sr) = rq.x; at(sr + 1) = rq.y; at(sr + 2) = rq.z; at(sr + 3) = rq.w;
        }
        __syncwarp();
        // ---- copy-out: the warp's EPW environments are ONE contiguous range of the output; all 32 lanes
        // stream it with 128-bit stores, every float fetched through the static map. An environment that is
        // not written (past N, or masked out of this step) keeps its observation: bit env*GC of `lm`.
        const uint8_t *map = flat_map<RC>();
        const uint32_t lm = __ballot_sync(FULL, live);
        auto env_live = [&](int env) { return ((lm >> (env * GC)) & 1u) != 0u; };
        constexpr int TOT = EPW * M::RF;                                     // floats of this warp
        float *wout = out + (size_t)env0 * M::RF;
        const int head = (M::RF % 4 == 0) ? 0 : (int)((0u - env0 * (uint32_t)M::RF) & 3u);   // floats before the first 16-byte boundary (0 or 2)
        auto locate = [&](int k, int &env, int &kk) {    // float k of the warp's range -> (environment, float in it)
            env = 0; kk = k;
            if constexpr (EPW <= 3) {
#pragma unroll
                for (int t = 1; t < EPW; ++t) if (k >= t * M::RF) { env = t; kk = k - t * M::RF; }
            } else {
                env = (int)(((uint32_t)k * (uint32_t)((65536 + M::RF - 1) / M::RF)) >> 16);   // k / RF (range checked below)
                kk = k - env * M::RF;
            }
        };
        auto pair = [&](int env, int kk) -> float2 {     // floats kk, kk+1 of an environment (kk even)
            const uint32_t mm = *reinterpret_cast<const uint16_t *>(map + kk);
            const float *v = wstage + env * M::VS;
            return make_float2(v[mm & 0xFFu], v[mm >> 8]);
        };
        static_assert(M::RF % 2 == 0, "pairs never straddle environments");
        static_assert(EPW <= 3 || ((EPW * M::RF - 1) * ((65536 + M::RF - 1) / M::RF)) >> 16 == EPW - 1, "reciprocal division range");
        constexpr int ITERS = (TOT / 4 + 31) / 32;
        // Fast path (always, unless an env mask punches holes): the written environments are a prefix of the
        // warp's range, so liveness is one compare and the composed warp map gives the sources directly.
        int live_envs = 0;
        bool prefix = true;
#pragma unroll
        for (int t = 0; t < EPW; ++t) {
            const bool l = env_live(t);
            prefix = prefix && !(l && live_envs != t);               // a live env after a dead one: not a prefix
            live_envs += l ? 1 : 0;
        }
        // Small, every environment of the warp written (every step of a full batch): the warp's 8 environments are
        // 4 UNITS of 2 (296 floats = 74 float4 = 37 whole sectors each), and float4 c of a unit has the same four
        // sources in every unit — up to the unit's table offset, a compile-time constant once the loop is unrolled.
        // Lane L owns float4 L, L + 32 and (L < 10) L + 64 of every unit: it fetches their 12 source addresses
        // ONCE and then issues 4 shared-memory loads + one 128-bit store per float4 with immediate offsets
        // (~80 instructions per lane instead of ~170 for the map-indexed loop below; the kernel is issue-bound).
        if constexpr (RC == 4 && WH_FLAT_UNITS) {
            if (live_envs == EPW) {
                constexpr int UF4 = 2 * M::RF / 4;                             // float4 per unit (74)
                constexpr int CH = (UF4 + 31) / 32;                            // float4 per lane and unit (3)
                static_assert(EPW % 2 == 0 && (2 * M::RF) % 8 == 0, "units are whole sectors");
                const uint16_t *wm = flat_warp_map<RC>();
                const float *a[CH][4];
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const int c = g.lane + 32 * j;
                    const uint2 v = (c < UF4) ? __ldg(reinterpret_cast<const uint2 *>(wm + 4 * c)) : make_uint2(0u, 0u);
                    a[j][0] = wstage + (v.x & 0xFFFFu); a[j][1] = wstage + (v.x >> 16);
                    a[j][2] = wstage + (v.y & 0xFFFFu); a[j][3] = wstage + (v.y >> 16);
                }
                float4 *o4 = reinterpret_cast<float4 *>(wout) + g.lane;
#pragma unroll
                for (int u = 0; u < EPW / 2; ++u) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        if (32 * j + 31 < UF4 || g.lane + 32 * j < UF4)
                            WH_ST(o4 + u * UF4 + 32 * j, make_float4(a[j][0][u * 2 * M::VS], a[j][1][u * 2 * M::VS],
                                                                     a[j][2][u * 2 * M::VS], a[j][3][u * 2 * M::VS]));
                    }
                }
                return;
            }
        }
        // (Large keeps the per-element map: its 9.3 KB composed table measured 3 % slower, HBM-bound either way)
        if (prefix && RC != 16) {
            const uint16_t *wm = flat_warp_map<RC>();
            const int limit = live_envs * M::RF;
            const int nbl = limit > head ? (limit - head) >> 2 : 0;
#pragma unroll 4
            for (int it = 0; it < ITERS; ++it) {
                const int i = g.lane + 32 * it;
                if (i < nbl) {
                    const int k = head + 4 * i;
                    uint32_t lo, hi;
                    if constexpr (M::RF % 4 == 0) {
                        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(wm + k));
                        lo = v.x; hi = v.y;
                    } else {
                        lo = __ldg(reinterpret_cast<const uint32_t *>(wm + k));
                        hi = __ldg(reinterpret_cast<const uint32_t *>(wm + k + 2));
                    }
                    WH_ST(reinterpret_cast<float4 *>(wout + k),
                          make_float4(wstage[lo & 0xFFFFu], wstage[lo >> 16], wstage[hi & 0xFFFFu], wstage[hi >> 16]));
                }
            }
            if constexpr (M::RF % 4 != 0) {                                  // 8-byte head / tail of a misaligned range
                if (g.lane == 0 && head == 2 && limit >= 2) {
                    const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(wm));
                    WH_ST(reinterpret_cast<float2 *>(wout), make_float2(wstage[v & 0xFFFFu], wstage[v >> 16]));
                }
                const int tail = head + 4 * nbl;
                if (g.lane == 1 && limit > 0 && tail < limit) {
                    const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(wm + tail));
                    WH_ST(reinterpret_cast<float2 *>(wout + tail), make_float2(wstage[v & 0xFFFFu], wstage[v >> 16]));
                }
            }
            return;
        }
        const int nb = (TOT - head) >> 2;                                    // whole float4 in the range
#pragma unroll 4
        for (int it = 0; it < ITERS; ++it) {
            const int i = g.lane + 32 * it;
            if (i < nb) {
                const int k = head + 4 * i;
                int e0, k0, e1, k1;
                locate(k, e0, k0);
                if constexpr (M::RF % 4 == 0) { e1 = e0; k1 = k0 + 2; }
                else locate(k + 2, e1, k1);
                const bool l0 = env_live(e0), l1 = env_live(e1);
                if (l0 && l1) {
                    const float2 a = pair(e0, k0), b = pair(e1, k1);
                    WH_ST(reinterpret_cast<float4 *>(wout + k), make_float4(a.x, a.y, b.x, b.y));
                } else if (l0) {
                    WH_ST(reinterpret_cast<float2 *>(wout + k), pair(e0, k0));
                } else if (l1) {
                    WH_ST(reinterpret_cast<float2 *>(wout + k + 2), pair(e1, k1));
                }
            }
        }
        if constexpr (M::RF % 4 != 0) {                                      // 8-byte head / tail of a misaligned range
            if (g.lane == 0 && head == 2 && env_live(0)) WH_ST(reinterpret_cast<float2 *>(wout), pair(0, 0));
            const int tail = head + 4 * nb;
            if (g.lane == 1 && tail < TOT && env_live(EPW - 1))
                WH_ST(reinterpret_cast<float2 *>(wout + tail), pair(EPW - 1, tail - (EPW - 1) * M::RF));
        }
    } else {
        // runtime-R geometries: every lane writes its pieces of each agent row straight to global memory
        const uint32_t mine = (ppos & 0x7Fu) | (((ppos >> 8) & 0x7Fu) << 7) | ((tpos & 0x7Fu) << 14) |
                              (((tpos >> 8) & 0x7Fu) << 21) | (avail << 28);
        const uint32_t next = g.shfl_down1(mine);
        const int F = 9 * R + 1;
        auto px = [](uint32_t v) { return (float)(v & 0x7F); };
        auto py = [](uint32_t v) { return (float)((v >> 7) & 0x7F); };
        auto tx = [](uint32_t v) { return (float)((v >> 14) & 0x7F); };
        auto ty = [](uint32_t v) { return (float)((v >> 21) & 0x7F); };
        auto av = [](uint32_t v) { return (float)((v >> 28) & 1); };
        const uint32_t t_row1 = (g.gl >= 1) ? next : mine;       // core.py:428 (step) drops row 1
        float *env_out = out + (long long)e * R * F;
        if (live)
            for (int a = 0; a < R; ++a) {
                float *dst = env_out + a * F;
                if (g.gl == 0) dst[0] = n_agents;
                if (g.gl < R - 1) {
                    const uint32_t o = (g.gl >= a) ? next : mine;                   // core.py:426-427
                    const uint32_t t = (flavour == WH_OBS_STEP) ? t_row1 : o;
                    dst[1 + g.gl] = av(o);
                    dst[R + 2 * g.gl] = tx(t); dst[R + 2 * g.gl + 1] = ty(t);
                    dst[3 * R - 2 + 2 * g.gl] = px(o); dst[3 * R - 2 + 2 * g.gl + 1] = py(o);
                }
                if (g.gl < R) {
                    float *q = dst + 5 * R - 4 + 4 * g.gl;
                    q[0] = rq.x; q[1] = rq.y; q[2] = rq.z; q[3] = rq.w;
                }
                if (g.gl == a) {
                    dst[9 * R - 4] = av(mine);
                    dst[9 * R - 3] = tx(mine); dst[9 * R - 2] = ty(mine);
                    dst[9 * R - 1] = px(mine); dst[9 * R] = py(mine);
                }
            }
    }
}

// delivery-target cell of my agent for the observation tables (null cell when not delivering)
template <int GC>
__device__ __forceinline__ uint32_t target_cell16(const KParams &P, int atgt) {
    const Geo<GC> geo(P);
    return atgt > -1 ? delivery_cell16(atgt, geo.dim) : geo.null16();
}

// ---------------------------------------------------------------------------------------------
// reset — core.py:167-221, variants.py:69-74
// ---------------------------------------------------------------------------------------------
template <int GC>
__device__ __forceinline__ unsigned long long do_reset(const KParams &P, const Group<GC> &g, env_t e,
                                                       int R, uint32_t env_id, EnvRegs &s, bool replay,
                                                       bool doit) {
    // `doit` is uniform within the group; groups that skip still take part in warp-wide votes
    const Geo<GC> geo(P);
    EnvRegs n = s;
    n.ep = s.ep + 1;
    n.time = 0;                                                                // core.py:168
    uint32_t u0 = 0, u1 = 0;
    if (replay) {
        if (P.r_num_agents) n.A = P.r_num_agents[e];
    } else if (P.random_agents) {                                             // variants.py:70,74
        philox4x32_10(env_id, (uint32_t)n.ep, CTR_NUM_AGENTS, 0u, P.seed, u0, u1);
        n.A = 1 + (int)bounded(u0, (uint32_t)P.max_agents);
    }
    n.atgt = -1;                                                               // core.py:204
    n.pos16 = 0xFFFFu;
    if (replay) {
        if (g.gl < n.A) n.pos16 = reinterpret_cast<const uint16_t *>(P.r_agent_pos)[e * R + g.gl];
    } else {
        // core.py:192-201 rejection sampling over [1,dim-2]^2 minus pickup cells; other agents are
        // NOT checked, so agents may be co-located
        bool need = doit && g.gl < n.A;
        for (uint32_t j = 0; __any_sync(FULL, need); ++j) {
            if (need) {
                philox4x32_10(env_id, (uint32_t)n.ep, CTR_SPAWN_AGENT + j, (uint32_t)g.gl, P.seed, u0, u1);
                const int x = 1 + (int)bounded(u0, (uint32_t)(geo.dim - 2));
                const int y = 1 + (int)bounded(u1, (uint32_t)(geo.dim - 2));
                if (pickup_index(P, geo, x, y) < 0) { n.pos16 = (uint32_t)x | ((uint32_t)y << 8); need = false; }
            }
        }
    }
    n.pt4 = 0xFFFFFFFFu;                                                       // core.py:210-211
    n.tmr = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    // core.py:215-221 R distinct pickup points x R distinct delivery points, paired in draw order
    int sp = -1, st_ = -1;
    if (replay) {
        if (g.gl < R) { sp = P.r_init_p[e * R + g.gl]; st_ = P.r_init_t[e * R + g.gl]; }
    } else {
        philox4x32_10(env_id, (uint32_t)n.ep, CTR_INIT_REQUESTS, (uint32_t)g.gl, P.seed, u0, u1);
    }
    unsigned long long inactive = (geo.PP >= 64) ? ~0ull : ((1ull << geo.PP) - 1ull);
    unsigned long long avail_d = (geo.D >= 64) ? ~0ull : ((1ull << geo.D) - 1ull);
    unsigned long long active = 0ull;
    for (int i = 0; i < R; ++i) {
        int p, d;
        if (replay) {
            p = (int)g.shfl((uint32_t)sp, i);
            d = (int)g.shfl((uint32_t)st_, i);
        } else {
            const uint32_t a = g.shfl(u0, i), b = g.shfl(u1, i);
            p = nth_set64<Group<GC>::PBITS>(inactive, (int)bounded(a, (uint32_t)(geo.PP - i)));
            d = nth_set64<64>(avail_d, (int)bounded(b, (uint32_t)(geo.D - i)));
        }
        if (p >= 0) {
            inactive &= ~(1ull << p);
            avail_d &= ~(1ull << d);
            active |= 1ull << p;
            if ((p >> 2) == g.gl) {
                const int j = p & 3;
                n.pt4 = (n.pt4 & ~(0xFFu << (8 * j))) | (((uint32_t)d & 0xFFu) << (8 * j));
                set_timer(n, j, (uint32_t)P.wait);
            }
        }
    }
    if (doit) s = n;
    return active;
}

// ---------------------------------------------------------------------------------------------
// greedy solver evaluated from the state in registers — baseline/solvers.py:27-58
// ---------------------------------------------------------------------------------------------
// Equivalent to running the solver on the observation the previous step()/reset() returned:
// after reset (time == 0) every availability is 0 and every delivery target is the null position
// (core.py:233-236), so agents head for the map centre; otherwise a free agent goes to the
// L1-nearest request (first minimum in request order == ascending pickup index), a delivering
// agent to its delivery cell.
// active_io: out = the active-request mask of this state (do_world derives the post-step mask from it); with
// have_active it is also an INPUT — the multi-step loops pass the mask the previous step left (StepOut::active,
// or do_reset's) instead of gathering it from the lanes again.
template <int GC, int RC>
__device__ __forceinline__ int greedy_from_state(const KParams &P, const Group<GC> &g, int R,
                                                 uint32_t env_id, const EnvRegs &s, unsigned long long &active_io,
                                                 bool have_active = false) {
    const Geo<GC> geo(P);
    const unsigned long long active = have_active ? active_io : active_mask(g, s.pt4);
    active_io = active;
    const int px = s.pos16 & 0xFF, py = s.pos16 >> 8;
    // lane r takes the r-th active pickup point's cell
    const int nact = __popcll(active);
    uint32_t cell = geo.null16();
    if (g.gl < nact && g.gl < R)
        cell = pickup_cell16(P, geo, (RC != 0 && RC <= 4) ? nth_set_small<(RC ? RC : 1)>((uint32_t)active, g.gl)
                                                          : nth_set64<Group<GC>::PBITS>(active, g.gl));
    // L1 argmin with first-minimum tie-break (solvers.py:53-58) as a running minimum of the key
    // distance << 21 | request index << 16 | cell. The distance of two packed cells (x | y << 8, upper
    // bytes zero) is ONE instruction: the byte-wise sum of absolute differences (VABSDIFF4.U8.ACC).
    uint32_t best = 0xFFFFFFFFu;
    const int RR = RC ? RC : R;
#pragma unroll
    for (int r = 0; r < RR; ++r) {
        const uint32_t c = g.shfl(cell, r);
        const uint32_t d = __vsadu4(s.pos16 & 0xFFFFu, c);
        best = min(best, (d << 21) + (((uint32_t)r << 16) | c));
    }
    const uint32_t bcell = best & 0xFFFFu;
    uint32_t target;
    const bool free_agent = s.time > 0 && s.atgt == -1;                        // availability 1
    if (free_agent) target = bcell;
    else if (s.time > 0) target = delivery_cell16(s.atgt, geo.dim);              // solvers.py:33-34
    else target = geo.null16();
    const int sx = max(-1, min(1, (int)(target & 0xFF) - px));                 // solvers.py:41
    const int sy = max(-1, min(1, (int)(target >> 8) - py));
    int action = (sx + 1) * 3 + (sy + 1);                                      // solvers.py:47-49
    if (P.rand_thr) {                                                          // solvers.py:44-45
        uint32_t u0, u1;
        philox4x32_10(env_id, (uint32_t)s.ep, (uint32_t)s.time, (uint32_t)g.gl | CTR_SOLVER_TAG, P.solver_seed, u0, u1);
        if ((unsigned long long)u0 < P.rand_thr) action = (int)bounded(u1, 9u);
    }
    return (g.gl < s.A) ? action : -1;
}

}  // namespace wh
