// wh_tpe.cuh — EXPERIMENTAL thread-per-environment step kernel for the small-footprint variants.
//
// STATUS (round 1): bit-exact (the whole -m gpu suite passes with it dispatched for Small / Medium)
// but SLOWER than the lane-group kernels on B200. First version (everything inlined, int4 / int2
// staging): Small 0.57 vs 0.68, Medium 0.71 vs 0.86 of the HBM peak — 21 % / 37 % fewer warp
// instructions, but 10.9 k SASS instructions of mostly straight-line code (instruction-cache misses)
// and 18 KB of staging per warp (12 resident warps per SM for Medium). This version (cold helpers out
// of line, compact staging tables, 5.1 k SASS): Small 0.63, Medium 0.82; an occupancy sweep
// (32..114 registers) does not close the gap, the per-thread dependency chains do not hide behind
// 16-32 warps. Not compiled by default: build with -DWH_WITH_TPE and run with WH_ENABLE_TPE=1 to A/B
// it. See profiles/README.md.
//
// The lane-group kernels (wh_kernels.cuh) give one environment G = R lanes; for Small (R = 4) and
// Medium (R = 9) they are bound by instruction issue, not by HBM: every warp instruction serves only
// 8 / 3 environments. Here ONE THREAD owns one environment: the whole game logic runs on registers
// with compile-time-unrolled loops (no shuffles, no ballots), a warp instruction serves 32
// environments, and everything that goes to HBM — state, rewards, and the observation tensors —
// is first laid out per environment in shared memory (bank-conflict-free padded strides) and then
// streamed out by the whole warp over the CONTIGUOUS [32 envs, ...] block of each tensor.
//
// Semantics are identical to k_step (same reference lines, same Philox counters); the parity tests
// drive both. Supported here: ascending action order, native RNG, dict-layout observations,
// auto-reset, compact I/O, in-kernel greedy solver. Everything else (custom order, replayed draws,
// flattened observations, runtime geometry) stays on the lane-group kernels.
#pragma once
#include "../../rllib_warehouse_b200/csrc/wh_kernels.cuh"

namespace wh {

__host__ __device__ constexpr int tpe_pow2_div(int v) { return (v % 16 == 0) ? 16 : (v % 8 == 0) ? 8 : (v % 4 == 0) ? 4 : (v % 2 == 0) ? 2 : 1; }
__host__ __device__ constexpr int tpe_odd_words(int bytes) {   // stride in bytes: whole words, odd word count (conflict-free for 32-bit accesses)
    const int w = (bytes + 3) / 4;
    return 4 * ((w % 2) ? w : w + 1);
}

template <int RC>
struct Tpe {
    static constexpr int R = RC;
    static constexpr int P = 4 * RC;                 // variant kernels: P == 4R (Small 16, Medium 36)
    static constexpr int PW = P / 4;                 // pickup-target words (4 int8 each)
    // per-env rows (bytes in the global tensor) and padded strides (bytes in shared memory)
    static constexpr int ROW_PT = P, STR_PT = tpe_odd_words(ROW_PT);
    static constexpr int ROW_TM = 2 * P, STR_TM = tpe_odd_words(ROW_TM);
    static constexpr int ROW_POS = 2 * R, STR_POS = tpe_odd_words(ROW_POS);
    static constexpr int ROW_TGT = R, STR_TGT = tpe_odd_words(ROW_TGT);
    static constexpr int ROW_REW = 4 * R, STR_REW = tpe_odd_words(ROW_REW);
    static constexpr int STR_REQ = tpe_odd_words(4 * R);   // compact request list: 4 bytes (px,py,dx,dy) per request
    static constexpr int STR_PP = tpe_odd_words(2 * R);    // padded position / target-cell tables: one u16 cell per row
    static constexpr int ROW_SA = R, STR_SA = tpe_odd_words(ROW_SA);
    static constexpr int ROW_OA = R * (R - 1), STR_OA = tpe_odd_words(ROW_OA);
    __host__ __device__ static constexpr int al16(int v) { return (v + 15) / 16 * 16; }
    static constexpr int O_PT = 0;
    static constexpr int O_TM = al16(O_PT + 32 * STR_PT);
    static constexpr int O_POS = al16(O_TM + 32 * STR_TM);
    static constexpr int O_TGT = al16(O_POS + 32 * STR_POS);
    static constexpr int O_REW = al16(O_TGT + 32 * STR_TGT);
    static constexpr int O_TIME = al16(O_REW + 32 * STR_REW);     // int32 [32]
    static constexpr int O_EP = O_TIME + 128;                     // int32 [32]
    static constexpr int O_A = O_EP + 128;                        // int8  [32]
    static constexpr int O_DONE = O_A + 32;                       // u8    [32]
    static constexpr int O_REQ = al16(O_DONE + 32);
    static constexpr int O_PP = al16(O_REQ + 32 * STR_REQ);
    static constexpr int O_TP = al16(O_PP + 32 * STR_PP);
    static constexpr int O_SA = al16(O_TP + 32 * STR_PP);
    static constexpr int O_OA = al16(O_SA + 32 * STR_SA);
    static constexpr int BYTES = al16(O_OA + 32 * STR_OA);
};

template <int GB> struct TpeVec;
template <> struct TpeVec<16> { typedef int4 type; };
template <> struct TpeVec<8> { typedef int2 type; };
template <> struct TpeVec<4> { typedef int32_t type; };
template <> struct TpeVec<2> { typedef int16_t type; };
template <> struct TpeVec<1> { typedef int8_t type; };

// cold helpers are kept out of line so that the hot body stays small (instruction cache)
__device__ __noinline__ uint2 tpe_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, unsigned long long seed) {
    uint32_t a, b;
    philox4x32_10(c0, c1, c2, c3, seed, a, b);
    return make_uint2(a, b);
}

// regular racks only (4, 8, 12, ...): the dispatcher guarantees it
__device__ __forceinline__ int tpe_pickup_index(int L, int x, int y) {
    const int qx = (x + 1) >> 2, ox = (x + 1) & 3, qy = (y + 1) >> 2, oy = (y + 1) & 3;
    const bool ok = ox < 2 && oy < 2 && qx >= 1 && qx <= L && qy >= 1 && qy <= L;
    return ok ? 4 * ((qx - 1) * L + (qy - 1)) + ox + 2 * oy : -1;
}

// Streams the rows of `nl` consecutive environments (ROW bytes each, padded to STRIDE in shared
// memory) into the contiguous global block that starts at gdst.
template <int ROW, int STRIDE>
__device__ __forceinline__ void tpe_flush(void *gdst, const unsigned char *sbase, int nl, int lane) {
    // element width: the largest power of two dividing both the row size (global alignment for any
    // first env) and the padded stride (shared-memory alignment)
    constexpr int GB = tpe_pow2_div(ROW) < tpe_pow2_div(STRIDE) ? tpe_pow2_div(ROW) : tpe_pow2_div(STRIDE);
    constexpr unsigned PER = ROW / GB;
    typedef typename TpeVec<GB>::type V;
    V *d = reinterpret_cast<V *>(gdst);
    const unsigned total = (unsigned)nl * PER;
    for (unsigned i = lane; i < total; i += 32) {
        const unsigned el = i / PER, j = i - el * PER;
        d[i] = *reinterpret_cast<const V *>(sbase + el * STRIDE + j * GB);
    }
}

#ifndef WH_TPE_MB_SMALL
#define WH_TPE_MB_SMALL 8
#endif
#ifndef WH_TPE_MB_MEDIUM
#define WH_TPE_MB_MEDIUM 4
#endif
template <int RC, bool GREEDY>
__global__ void __launch_bounds__(128, (RC == 4 ? WH_TPE_MB_SMALL : WH_TPE_MB_MEDIUM)) k_step_tpe(const __grid_constant__ KParams P) {
    using T = Tpe<RC>;
    constexpr int R = RC, NP = T::P, PW = T::PW;
    constexpr int PBITS = NP <= 16 ? 16 : (NP <= 32 ? 32 : 64);
    extern __shared__ __align__(16) unsigned char tpe_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    unsigned char *W = tpe_smem + wib * T::BYTES;               // this warp's staging
    const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + wib;
    const uint32_t e0 = warp * 32u, n = (uint32_t)P.N;
    if (e0 >= n) return;                                        // whole warp out of range (warp-uniform)
    const int nl = (n - e0 < 32u) ? (int)(n - e0) : 32;         // live envs of this warp (tail warp < 32)
    const bool live = lane < nl;
    const uint32_t e = live ? e0 + lane : e0 + nl - 1;          // dead lanes shadow the last live env
    const uint32_t env_id = (uint32_t)P.env_id0 + e;
    const int dim = P.dim;
    const uint32_t null16 = (uint32_t)P.null_pos | ((uint32_t)P.null_pos << 8);

    // my rows in the per-warp regions
    uint8_t *s_pt = W + T::O_PT + lane * T::STR_PT;             // int8  [P]  pickup targets (dynamic indexing)
    int16_t *s_tm = reinterpret_cast<int16_t *>(W + T::O_TM + lane * T::STR_TM);   // int16 [P] timers
    uint32_t *s_req = reinterpret_cast<uint32_t *>(W + T::O_REQ + lane * T::STR_REQ);   // compact request list (px,py,dx,dy bytes)

    // ------------------------------------------------------------------------------- load state
    int time = P.time[e], A = P.num_agents[e], ep = P.episode_ctr[e];
    uint32_t pos16[R];
    int tgt[R];
    uint32_t ptw[PW];
    {
        const uint16_t *gpos = reinterpret_cast<const uint16_t *>(P.agent_pos) + e * R;
        const int8_t *gtgt = P.agent_tgt + e * R;
        const uint32_t *gpt = reinterpret_cast<const uint32_t *>(P.pickup_tgt + e * NP);
        const uint32_t *gtm = reinterpret_cast<const uint32_t *>(P.pickup_timer + e * NP);
#pragma unroll
        for (int a = 0; a < R; ++a) { pos16[a] = gpos[a]; tgt[a] = gtgt[a]; }
#pragma unroll
        for (int w = 0; w < PW; ++w) ptw[w] = gpt[w];
#pragma unroll
        for (int w = 0; w < NP / 2; ++w) reinterpret_cast<uint32_t *>(s_tm)[w] = gtm[w];
    }

    // request list of the state mirrored in s_pt (ascending pickup index) -> compact table in smem
    // (core.py:409-418). Scratch for the indices of the active points: my other_availabilities row,
    // which is only written after the last scan.
    uint8_t *s_idx = W + T::O_OA + lane * T::STR_OA;
    auto scan_requests = [&](const uint32_t (&pw)[PW]) {
        int rank = 0;
#pragma unroll
        for (int p = 0; p < NP; ++p)
            if ((((pw[p >> 2] >> (8 * (p & 3) + 7)) & 1u) == 0u) && rank < R) s_idx[rank++] = (uint8_t)p;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            uint32_t q = null16 | (null16 << 16);                              // only if < R active (unreachable)
            if (r < rank) {
                const int p = s_idx[r];
                q = pickup_cell16(P, Geo<0>(P), p) | (delivery_cell16((int)(s_pt[p] & 0x3Fu), dim) << 16);
            }
            s_req[r] = q;
        }
    };

    // ------------------------------------------------------------------------------- actions
    int act[R];
    if (GREEDY) {                                               // solvers.py:27-58 on the previous observation
#pragma unroll
        for (int w = 0; w < PW; ++w) reinterpret_cast<uint32_t *>(s_pt)[w] = ptw[w];
        scan_requests(ptw);
        uint32_t cells[R];
#pragma unroll
        for (int r = 0; r < R; ++r) cells[r] = s_req[r] & 0xFFFFu;
#pragma unroll
        for (int a = 0; a < R; ++a) {
            const int px = pos16[a] & 0xFF, py = pos16[a] >> 8;
            int best = 1 << 30;
            uint32_t bcell = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {                                       // solvers.py:53-58, first minimum
                const int d = abs(px - (int)(cells[r] & 0xFF)) + abs(py - (int)(cells[r] >> 8));
                if (d < best) { best = d; bcell = cells[r]; }
            }
            uint32_t target = null16;                                          // reset-flavour obs (core.py:233-236)
            if (time > 0) target = (tgt[a] == -1) ? bcell : delivery_cell16(tgt[a], dim);   // solvers.py:33-39
            const int sx = max(-1, min(1, (int)(target & 0xFF) - px)), sy = max(-1, min(1, (int)(target >> 8) - py));
            int action = (sx + 1) * 3 + (sy + 1);                              // solvers.py:41,47-49
            if (P.rand_thr) {                                                  // solvers.py:44-45
                uint32_t u0, u1;
                { const uint2 ph_ = tpe_philox(env_id, (uint32_t)ep, (uint32_t)time, (uint32_t)a, P.solver_seed); u0 = ph_.x; u1 = ph_.y; }
                if ((unsigned long long)u0 < P.rand_thr) action = (int)bounded(u1, 9u);
            }
            act[a] = (a < A) ? action : -1;
            if (P.actions_out && live) P.actions_out[e * R + a] = act[a];
        }
    } else {
#pragma unroll
        for (int a = 0; a < R; ++a)
            act[a] = (P.flags & WH_FLAG_COMPACT_IO) ? (int)reinterpret_cast<const int8_t *>(P.actions)[e * R + a]
                                                    : P.actions[e * R + a];
    }
    time += 1;                                                                 // core.py:267

    // ------------------------------------------------------------------------------- moves, core.py:275-300
    {
        uint32_t m[R], rev[R], ca[R], cb[R], mark[R];
        bool moved[R];
#pragma unroll
        for (int a = 0; a < R; ++a) {
            const int px = pos16[a] & 0xFF, py = pos16[a] >> 8;
            m[a] = rev[a] = ca[a] = cb[a] = ABSENT_MOVE;
            moved[a] = false;
            mark[a] = (a < A) ? pos16[a] : NO_CELL;                            // core.py:276
            if (a < A && act[a] >= 0 && act[a] <= 8) {
                const int ax = (act[a] * 11) >> 5;                             // MOVES, core.py:38
                int x = px + ax - 1, y = py + (act[a] - 3 * ax) - 1;
                if ((unsigned)x >= (unsigned)dim) x = px;                      // core.py:284-287
                if ((unsigned)y >= (unsigned)dim) y = py;
                const uint32_t to = (uint32_t)x | ((uint32_t)y << 8);
                m[a] = pos16[a] | (to << 16);
                rev[a] = to | (pos16[a] << 16);                                // core.py:294
                ca[a] = cb[a] = rev[a];
                if (x != px && y != py) {                                      // core.py:295-297
                    const uint32_t c1 = (uint32_t)x | ((uint32_t)py << 8), c2 = (uint32_t)px | ((uint32_t)y << 8);
                    ca[a] = c1 | (c2 << 16);
                    cb[a] = c2 | (c1 << 16);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < R; ++t) {                                          // ascending agent ids (core.py:279)
            const uint32_t mm = m[t], c = mm >> 16, from = mm & 0xFFFFu;
            bool hit = false;
#pragma unroll
            for (int j = 0; j < R; ++j) {
                hit |= (mark[j] == c);
                if (j < t) hit |= moved[j] & ((rev[j] == mm) | (ca[j] == mm) | (cb[j] == mm));
            }
            const bool ok = !hit && mm != ABSENT_MOVE;                         // core.py:289
#pragma unroll
            for (int j = 0; j < R; ++j)
                if (ok && mark[j] == from) mark[j] = NO_CELL;                  // core.py:290
            if (ok) { mark[t] = c; moved[t] = true; pos16[t] = c; }            // core.py:291-300
        }
    }

    // ------------------------------------------------------------------------------- expiry, core.py:303-306
    int nexp = 0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const uint32_t sh = 8 * (p & 3);
        const bool active = ((ptw[p >> 2] >> (sh + 7)) & 1u) == 0u;
        int tmr = s_tm[p] - (active ? 1 : 0);
        if (tmr == 0) { ptw[p >> 2] |= 0xFFu << sh; tmr = -1; ++nexp; }
        s_tm[p] = (int16_t)tmr;
    }
#pragma unroll
    for (int w = 0; w < PW; ++w) reinterpret_cast<uint32_t *>(s_pt)[w] = ptw[w];

    // ------------------------------------------------------------------------------- pickups, core.py:309-335
    float reward[R];
    int npick = 0;
    {
        int cand[R], tgv[R];
#pragma unroll
        for (int a = 0; a < R; ++a) {
            reward[a] = 0.0f;
            cand[a] = (a < A) ? tpe_pickup_index(P.L, pos16[a] & 0xFF, pos16[a] >> 8) : -1;
            tgv[a] = (cand[a] >= 0) ? (int)(int8_t)s_pt[cand[a]] : -1;        // pre-assignment targets (core.py:320-324)
        }
#pragma unroll
        for (int a = 0; a < R; ++a) {
            if (cand[a] >= 0 && tgt[a] == -1 && tgv[a] > -1) {
                tgt[a] = tgv[a]; reward[a] = 1.0f; ++npick;                     // core.py:327-329,335
                s_pt[cand[a]] = 0xFF; s_tm[cand[a]] = -1;                      // core.py:330-331
            }
        }
    }
    // ------------------------------------------------------------------------------- respawn, core.py:338-351
    unsigned long long active = 0ull;
#pragma unroll
    for (int w = 0; w < PW; ++w) {
        ptw[w] = reinterpret_cast<uint32_t *>(s_pt)[w];
        const uint32_t inv = ~ptw[w];
        const uint32_t nib = ((inv >> 7) & 1u) | ((inv >> 14) & 2u) | ((inv >> 21) & 4u) | ((inv >> 28) & 8u);
        active |= (unsigned long long)nib << (4 * w);
    }
    {
        const int k = R - __popcll(active);
        if (k > 0) {
            const unsigned long long pmask = (NP >= 64) ? ~0ull : ((1ull << NP) - 1ull);
            unsigned long long inactive = ~active & pmask;
            unsigned long long avail_d = (P.D >= 64) ? ~0ull : ((1ull << P.D) - 1ull);
            const int n_inact = __popcll(inactive);
            for (int i = 0; i < k; ++i) {
                uint32_t up, ut;
                { const uint2 ph_ = tpe_philox(env_id, (uint32_t)ep, (uint32_t)time, (uint32_t)i, P.seed); up = ph_.x; ut = ph_.y; }
                const int p = nth_set64<PBITS>(inactive, (int)bounded(up, (uint32_t)(n_inact - i)));
                const int d = nth_set64<64>(avail_d, (int)bounded(ut, (uint32_t)(P.D - i)));
                inactive &= ~(1ull << p);
                avail_d &= ~(1ull << d);
                s_pt[p] = (uint8_t)d;                                          // core.py:351
                s_tm[p] = (int16_t)P.wait;                                     // core.py:344
            }
#pragma unroll
            for (int w = 0; w < PW; ++w) ptw[w] = reinterpret_cast<uint32_t *>(s_pt)[w];
        }
    }
    // ------------------------------------------------------------------------------- deliveries, core.py:354-368
    int ndeliv = 0;
    uint32_t tcell[R];
#pragma unroll
    for (int a = 0; a < R; ++a) {
        tcell[a] = null16;
        if (a < A && tgt[a] > -1) {
            const uint32_t dc = delivery_cell16(tgt[a], dim);
            if (dc == pos16[a]) { tgt[a] = -1; reward[a] += 1.0f; ++ndeliv; }
            else tcell[a] = dc;
        }
    }

    // ------------------------------------------------------------------------------- dones, stats
    const bool done = time >= P.episode;                                       // core.py:438
    const bool auto_reset = (P.flags & WH_FLAG_AUTO_RESET) != 0;
    if (live && ((npick | ndeliv | nexp) != 0 || time == P.episode || (auto_reset && done))) {
        int4 a4 = reinterpret_cast<int4 *>(P.acc)[e];
        a4.x += npick; a4.y += ndeliv; a4.z += nexp;
        if (P.stats && time == P.episode) {                                    // train.py:18-23
            const unsigned long long ret = (unsigned long long)(a4.x + a4.y);
            atomicAdd(P.stats + 0, 1ull);
            atomicAdd(P.stats + 1, ret);
            atomicAdd(P.stats + 2, (unsigned long long)a4.x);
            atomicAdd(P.stats + 3, (unsigned long long)a4.y);
            atomicAdd(P.stats + 4, (unsigned long long)a4.z);
            atomicAdd(P.stats + 8 + 2 * (A - 1), 1ull);
            atomicAdd(P.stats + 9 + 2 * (A - 1), ret);
        }
        if (auto_reset && done) a4 = make_int4(0, 0, 0, 0);
        reinterpret_cast<int4 *>(P.acc)[e] = a4;
    }
    // rewards / dones of THIS step are staged before a possible reset overwrites the state
    const bool compact = (P.flags & WH_FLAG_COMPACT_IO) != 0;
    {
        unsigned char *row = W + T::O_REW + lane * T::STR_REW;
#pragma unroll
        for (int a = 0; a < R; ++a) {
            if (compact) row[a] = (uint8_t)reward[a];
            else reinterpret_cast<float *>(row)[a] = reward[a];                // core.py:435
        }
        (W + T::O_DONE)[lane] = done ? 1 : 0;
    }

    // ------------------------------------------------------------------------------- auto-reset, core.py:167-221
    int flavour = WH_OBS_STEP;
    if (auto_reset && done) {
        flavour = WH_OBS_RESET;
        ep += 1;
        time = 0;
        uint32_t u0, u1;
        if (P.random_agents) {                                                 // variants.py:70,74
            { const uint2 ph_ = tpe_philox(env_id, (uint32_t)ep, CTR_NUM_AGENTS, 0u, P.seed); u0 = ph_.x; u1 = ph_.y; }
            A = 1 + (int)bounded(u0, (uint32_t)P.max_agents);
        }
#pragma unroll
        for (int a = 0; a < R; ++a) {
            tgt[a] = -1;                                                       // core.py:204
            tcell[a] = null16;
            pos16[a] = 0xFFFFu;
            if (a < A) {
                for (uint32_t j = 0;; ++j) {                                   // core.py:192-201
                    { const uint2 ph_ = tpe_philox(env_id, (uint32_t)ep, CTR_SPAWN_AGENT + j, (uint32_t)a, P.seed); u0 = ph_.x; u1 = ph_.y; }
                    const int x = 1 + (int)bounded(u0, (uint32_t)(dim - 2)), y = 1 + (int)bounded(u1, (uint32_t)(dim - 2));
                    if (tpe_pickup_index(P.L, x, y) < 0) { pos16[a] = (uint32_t)x | ((uint32_t)y << 8); break; }
                }
            }
        }
#pragma unroll
        for (int w = 0; w < PW; ++w) reinterpret_cast<uint32_t *>(s_pt)[w] = 0xFFFFFFFFu;     // core.py:210-211
#pragma unroll
        for (int w = 0; w < NP / 2; ++w) reinterpret_cast<uint32_t *>(s_tm)[w] = 0xFFFFFFFFu;
        unsigned long long inactive = (NP >= 64) ? ~0ull : ((1ull << NP) - 1ull);
        unsigned long long avail_d = (P.D >= 64) ? ~0ull : ((1ull << P.D) - 1ull);
        for (int i = 0; i < R; ++i) {                                          // core.py:215-221
            { const uint2 ph_ = tpe_philox(env_id, (uint32_t)ep, CTR_INIT_REQUESTS, (uint32_t)i, P.seed); u0 = ph_.x; u1 = ph_.y; }
            const int p = nth_set64<PBITS>(inactive, (int)bounded(u0, (uint32_t)(NP - i)));
            const int d = nth_set64<64>(avail_d, (int)bounded(u1, (uint32_t)(P.D - i)));
            inactive &= ~(1ull << p);
            avail_d &= ~(1ull << d);
            s_pt[p] = (uint8_t)d;
            s_tm[p] = (int16_t)P.wait;
        }
#pragma unroll
        for (int w = 0; w < PW; ++w) ptw[w] = reinterpret_cast<uint32_t *>(s_pt)[w];
    }

    // ------------------------------------------------------------------------------- observation tables
    const bool want_obs = P.obs.requests != nullptr;
    if (want_obs) {
        scan_requests(ptw);                                     // s_pt mirrors ptw at this point
        uint16_t *s_pp = reinterpret_cast<uint16_t *>(W + T::O_PP + lane * T::STR_PP);
        uint16_t *s_tp = reinterpret_cast<uint16_t *>(W + T::O_TP + lane * T::STR_PP);
        int8_t *s_sa = reinterpret_cast<int8_t *>(W + T::O_SA + lane * T::STR_SA);
        uint32_t avail_bits = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {                                          // core.py:372-407 padded tables
            const bool real = r < A;
            const bool delivering = real && tgt[r] > -1;
            const uint32_t pp = real ? pos16[r] : null16;
            const uint32_t tp = (flavour == WH_OBS_STEP && delivering) ? tcell[r] : null16;
            const int av = (flavour == WH_OBS_STEP && real && !delivering) ? 1 : 0;
            avail_bits |= (uint32_t)av << r;
            s_pp[r] = (uint16_t)pp;
            s_tp[r] = (uint16_t)tp;
            s_sa[r] = (int8_t)av;
        }
        int8_t *s_oa = reinterpret_cast<int8_t *>(W + T::O_OA + lane * T::STR_OA);
#pragma unroll
        for (int a = 0; a < R; ++a)
#pragma unroll
            for (int o = 0; o < R - 1; ++o)
                s_oa[a * (R - 1) + o] = (int8_t)((avail_bits >> (o + (o >= a ? 1 : 0))) & 1u);   // core.py:427
    }
    // ------------------------------------------------------------------------------- stage the rest of the state
    {
        uint16_t *s_pos = reinterpret_cast<uint16_t *>(W + T::O_POS + lane * T::STR_POS);
        int8_t *s_tgt = reinterpret_cast<int8_t *>(W + T::O_TGT + lane * T::STR_TGT);
#pragma unroll
        for (int a = 0; a < R; ++a) { s_pos[a] = (uint16_t)pos16[a]; s_tgt[a] = (int8_t)tgt[a]; }
        reinterpret_cast<int32_t *>(W + T::O_TIME)[lane] = time;
        reinterpret_cast<int32_t *>(W + T::O_EP)[lane] = ep;
        (W + T::O_A)[lane] = (uint8_t)A;
    }
    __syncwarp();

    // ------------------------------------------------------------------------------- flush: state, rewards, dones
    tpe_flush<T::ROW_PT, T::STR_PT>(P.pickup_tgt + (size_t)e0 * NP, W + T::O_PT, nl, lane);
    tpe_flush<T::ROW_TM, T::STR_TM>(P.pickup_timer + (size_t)e0 * NP, W + T::O_TM, nl, lane);
    tpe_flush<T::ROW_POS, T::STR_POS>(P.agent_pos + (size_t)e0 * R * 2, W + T::O_POS, nl, lane);
    tpe_flush<T::ROW_TGT, T::STR_TGT>(P.agent_tgt + (size_t)e0 * R, W + T::O_TGT, nl, lane);
    tpe_flush<4, 4>(P.time + e0, W + T::O_TIME, nl, lane);
    if (compact) tpe_flush<R, T::STR_REW>(reinterpret_cast<uint8_t *>(P.rewards) + (size_t)e0 * R, W + T::O_REW, nl, lane);
    else tpe_flush<T::ROW_REW, T::STR_REW>(P.rewards + (size_t)e0 * R, W + T::O_REW, nl, lane);
    tpe_flush<1, 1>(P.dones + e0, W + T::O_DONE, nl, lane);
    if (auto_reset && __any_sync(FULL, done)) {
        tpe_flush<4, 4>(P.episode_ctr + e0, W + T::O_EP, nl, lane);
        tpe_flush<1, 1>(P.num_agents + e0, W + T::O_A, nl, lane);
    }
    if (!want_obs) return;

    // ------------------------------------------------------------------------------- flush: observations
    const wh_obs &o = P.obs;
    const size_t row0 = (size_t)e0 * R;
    {   // self_position / self_delivery_target [env][r] = padded table rows, expanded to int32 pairs
        int2 *dsp = reinterpret_cast<int2 *>(o.self_position) + row0, *dst = reinterpret_cast<int2 *>(o.self_delivery_target) + row0;
        for (unsigned i = lane; i < (unsigned)nl * R; i += 32) {
            const unsigned el = i / (unsigned)R, r = i - el * R;
            const uint32_t pc = reinterpret_cast<const uint16_t *>(W + T::O_PP + el * T::STR_PP)[r];
            const uint32_t tc = reinterpret_cast<const uint16_t *>(W + T::O_TP + el * T::STR_PP)[r];
            dsp[i] = make_int2(pc & 0xFF, pc >> 8);
            dst[i] = make_int2(tc & 0xFF, tc >> 8);
        }
    }
    tpe_flush<T::ROW_SA, T::STR_SA>(o.self_availability + row0, W + T::O_SA, nl, lane);
    tpe_flush<T::ROW_OA, T::STR_OA>(o.other_availabilities + row0 * (R - 1), W + T::O_OA, nl, lane);
    {   // num_agents key: A replicated over the R rows of an env
        int32_t *dst = o.num_agents + row0;
        const uint8_t *sa = W + T::O_A;
        for (unsigned i = lane; i < (unsigned)nl * R; i += 32) dst[i] = sa[i / (unsigned)R];
    }
    {   // requests [env][a][r] = request r of env: every list goes out R times (core.py:429)
        int4 *dst = reinterpret_cast<int4 *>(o.requests) + row0 * R;
        const unsigned total = (unsigned)nl * R * R;
        for (unsigned i = lane; i < total; i += 32) {
            const unsigned el = i / (unsigned)(R * R), rem = i - el * (R * R), r = rem % (unsigned)R;
            const uint32_t q = reinterpret_cast<const uint32_t *>(W + T::O_REQ + el * T::STR_REQ)[r];
            dst[i] = make_int4(q & 0xFF, (q >> 8) & 0xFF, (q >> 16) & 0xFF, q >> 24);
        }
    }
    {   // other_positions / other_delivery_targets [env][a][o] = padded row o + (o >= dropped row); 2 rows per int4
        constexpr unsigned H = R * (R - 1) / 2;
        int4 *dpos = reinterpret_cast<int4 *>(o.other_positions) + row0 * (R - 1) / 2;
        int4 *dtgt = reinterpret_cast<int4 *>(o.other_delivery_targets) + row0 * (R - 1) / 2;
        const uint8_t *sdone = W + T::O_DONE;
        const unsigned total = (unsigned)nl * H;
        for (unsigned i = lane; i < total; i += 32) {
            const unsigned el = i / H, k = i - el * H;
            const unsigned f0 = 2 * k, f1 = 2 * k + 1;
            const unsigned a0 = f0 / (unsigned)(R - 1), o0 = f0 - a0 * (R - 1);
            const unsigned a1 = f1 / (unsigned)(R - 1), o1 = f1 - a1 * (R - 1);
            const uint16_t *pp = reinterpret_cast<const uint16_t *>(W + T::O_PP + el * T::STR_PP);
            const uint16_t *tp = reinterpret_cast<const uint16_t *>(W + T::O_TP + el * T::STR_PP);
            const uint32_t p0 = pp[o0 + (o0 >= a0 ? 1 : 0)], p1 = pp[o1 + (o1 >= a1 ? 1 : 0)];
            dpos[i] = make_int4(p0 & 0xFF, p0 >> 8, p1 & 0xFF, p1 >> 8);       // core.py:426
            // core.py:428: step() always drops row 1; a reset observation drops row a (core.py:256)
            const bool rst = auto_reset && sdone[el] != 0;
            const unsigned d0 = rst ? a0 : 1u, d1 = rst ? a1 : 1u;
            const uint32_t t0 = tp[o0 + (o0 >= d0 ? 1 : 0)], t1 = tp[o1 + (o1 >= d1 ? 1 : 0)];
            dtgt[i] = make_int4(t0 & 0xFF, t0 >> 8, t1 & 0xFF, t1 >> 8);
        }
    }
}

}  // namespace wh
