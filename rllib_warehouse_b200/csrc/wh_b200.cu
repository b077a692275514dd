// wh_b200.cu — kernels' __global__ entry points, launch dispatch and the C ABI (include/wh_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <dlfcn.h>

#include "wh_kernels.cuh"
#ifdef WH_WITH_TPE
#include "../../tools/experimental/wh_tpe.cuh"   // rejected thread-per-env experiment (kept with the tuning tools, not in the product tree)
#endif

namespace wh {

#ifndef WH_BLOCK
#define WH_BLOCK 256
#endif
// Resident blocks per SM the register allocator must allow (measured optimum per variant,
// profiles/README.md): Large is fastest with FEWER, fatter warps (72 registers, 3 blocks = 24 warps:
// fewer concurrent write streams), Medium with 5, Small (issue-bound) with 6.
#ifndef WH_MIN_BLOCKS
#define WH_MIN_BLOCKS 5
#endif
#ifndef WH_MIN_BLOCKS_LARGE
#define WH_MIN_BLOCKS_LARGE 3
#endif
#ifndef WH_MIN_BLOCKS_MEDIUM
#define WH_MIN_BLOCKS_MEDIUM 5
#endif
#ifndef WH_MIN_BLOCKS_SMALL
#define WH_MIN_BLOCKS_SMALL 6
#endif
#ifndef WH_LARGE_DYN_SMEM
#define WH_LARGE_DYN_SMEM 24576   // 34.5 KB static + 24 KB: 3 blocks fit in 227 KB, 4 do not
#endif
#ifndef WH_MULTI_WS_DEFAULT
#define WH_MULTI_WS_DEFAULT 2           // observation warps per env tile of k_multi_ws (0 = never use it)
#endif
#ifndef WH_MULTI_WS_MAX_TILES_PER_SM
#define WH_MULTI_WS_MAX_TILES_PER_SM 8  // k_multi_ws while the launch has fewer env tiles per SM than this
#endif
#ifndef WH_WS_DIAG
#define WH_WS_DIAG 0
#endif
#ifndef WH_ROLLOUT_MIN_BLOCKS
#define WH_ROLLOUT_MIN_BLOCKS 4         // resident blocks per SM of k_rollout (Small / Medium / runtime geometries)
#endif
#ifndef WH_MULTI_MIN_BLOCKS
#define WH_MULTI_MIN_BLOCKS 3           // resident 256-thread blocks per SM of the throughput k_multi (80 registers: at 5 blocks = 48 registers it spilled 200 B and ran 14 % slower)
#endif
// Open-loop actions arrive from HBM (nobody has just written them) as small reads mixed into the step's write
// stream. WH_ACT_BURST_*: the blocks of the first wave prefetch the step's whole action tensor into L2 in one
// burst while they are parked in front of griddepcontrol.wait — the previous step is draining, the DRAM is
// nearly idle, and a prefetch is only a hint (L2 is the coherence point: it cannot expose stale data). Same-box
// A/B at 262 144 envs (profiles/README.md): Small 0.872 -> 0.895 of the HBM peak (its 4 MB of actions survive
// the 143 MB of observations a step writes), Medium 0.960 -> 0.952, Large 0.989 -> 0.983 — on for Small only.
// Rejected variants of the same idea: every warp prefetching the actions of the warp one (0.5, 2) wave(s) ahead
// (Small 0.869), the burst with the evict_last priority (Small 0.834), evict_first on the action loads, and a
// burst that also covers the state of the first 20 / 50 % of the envs (0.884 / 0.879 vs 0.894: it is still in L2).
#ifndef WH_ACT_BURST_SMALL
#define WH_ACT_BURST_SMALL 1
#endif
#ifndef WH_ACT_BURST_OTHER
#define WH_ACT_BURST_OTHER 0
#endif
#ifndef WH_ACT_BURST_MAX_MB
#define WH_ACT_BURST_MAX_MB 8           // lines beyond this would be evicted again before their warp arrives
#endif
#ifndef WH_KEEP_SMALL
#define WH_KEEP_SMALL 0                 // tuning builds: KEEP level (2 / 3) of the Small throughput kernels (0.882 / 0.886 vs 0.894 plain)
#endif
#ifndef WH_KEEP_PART_LARGE
#define WH_KEEP_PART_LARGE 0            // tuning builds: the partial evict_last policy (KEEP = 2) for Large as well
#endif
#ifndef WH_KEEP_PART_MAX_MB
#define WH_KEEP_PART_MAX_MB (2 * WH_KEEP_MAX_MB)
#endif
constexpr int BLOCK = WH_BLOCK;   // threads per block (tuning: -DWH_BLOCK / -DWH_MIN_BLOCKS)

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// Which environment a lane works for: warp w of the grid owns envs [w*epw, (w+1)*epw).
template <int GC>
struct Tile {
    env_t e, env0;   // this lane's environment; the first environment of the warp
    bool live;
    __device__ __forceinline__ Tile(const KParams &P, const Group<GC> &g)
        : Tile(P, g, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) {}                 // blockDim.x <= BLOCK
    // warp = index of the env tile (k_multi_ws: all warps of a block work on the block's one tile)
    __device__ __forceinline__ Tile(const KParams &P, const Group<GC> &g, uint32_t warp) {
        env0 = warp * (uint32_t)g.epw;
        const uint32_t env = env0 + (uint32_t)g.gi, n = (uint32_t)P.N;
        live = !g.ghost && env < n;
        e = (env < n) ? env : n - 1;   // dead lanes shadow a valid env so that loads stay in bounds
    }
};

// per-block staging for the observation build: one ObsStage per environment slot of the block
template <int GC, int RC, bool FLAT = false>
struct StageMem {
    static constexpr int EPW = GC ? 32 / GC : 1;
    // a slot serves one observation layout: dict keys (ObsStage) or RLlib-flattened (FlatStage)
    static constexpr int SLOT = FLAT ? FlatStage<RC>::BYTES : ObsStage<RC>::BYTES;
    static constexpr int BYTES = RC ? (BLOCK / 32) * EPW * SLOT : 16;
    __device__ static __forceinline__ unsigned char *mine(unsigned char *base, const Group<GC> &g) {
        if (RC == 0) return nullptr;
        const int slot = (threadIdx.x >> 5) * EPW + (g.ghost ? 0 : g.gi);
        return base + slot * SLOT;
    }
    // the whole staging area of this warp (EPW consecutive slots)
    __device__ static __forceinline__ unsigned char *warp_area(unsigned char *base) {
        if (RC == 0) return nullptr;
        return base + (threadIdx.x >> 5) * EPW * SLOT;
    }
};

// Per-episode event counters and, when an episode ends, the statistics vector that mirrors
// scripts/train.py:18-23. `owner` = the one lane of a live env that does it. The counters (`a`, 16 bytes per
// env) are loaded by the caller together with the rest of the state at kernel entry — loading them here, on
// demand, put a full DRAM round trip on the critical path of nearly every warp (some env of a warp has an
// event in most steps): 6-8 % of all warp stall samples (profiles/r02_hotspots_*.txt) — and are stored back
// only when this step changed them (`dirty`). Measured: Medium +1 %, Medium-65536 greedy +3.5 %, Large +0.3 %;
// Small -1.8 % (its 16 extra bytes per env are 2 % more traffic), so Small (EARLY = false) still loads on demand.
template <bool EARLY>
__device__ __forceinline__ void account_episode(const KParams &P, bool owner, env_t e, const StepOut &so,
                                                const EnvRegs &s, bool done, bool auto_reset, int4 &a, bool &dirty) {
    const bool flush = s.time == P.episode;
    if (owner && ((so.npick | so.ndeliv | so.nexp) != 0 || flush || (auto_reset && done))) {
        if (!EARLY) a = reinterpret_cast<const int4 *>(P.acc)[e];
        a.x += so.npick; a.y += so.ndeliv; a.z += so.nexp;
        if (P.stats && flush) {                                                // train.py:18-23
            const unsigned long long ret = (unsigned long long)(a.x + a.y);
            atomicAdd(P.stats + 0, 1ull);
            atomicAdd(P.stats + 1, ret);
            atomicAdd(P.stats + 2, (unsigned long long)a.x);
            atomicAdd(P.stats + 3, (unsigned long long)a.y);
            atomicAdd(P.stats + 4, (unsigned long long)a.z);
            atomicAdd(P.stats + 8 + 2 * (s.A - 1), 1ull);
            atomicAdd(P.stats + 9 + 2 * (s.A - 1), ret);
        }
        if (auto_reset && done) a = make_int4(0, 0, 0, 0);
        dirty = true;
    }
}

// Warehouse.step (+ optional in-kernel greedy solver, + optional observation build, + optional
// auto-reset) — core.py:262-442, solvers.py:27-58
// PLAIN: the caller passes int32 actions / float32 rewards, no dict order and no replayed draws (the
// throughput path); the instantiation then carries none of the code or tests for those options.
// KEEP (PLAIN only): the state is loaded / stored with the L2 evict_last priority (1 = all accesses, 2 = all arrays but the timers),
// see wh_kernels.cuh.
template <int GC, int RC, bool GREEDY, bool FLAT, bool PLAIN = false, int KEEP = 0>
__global__ void __launch_bounds__(BLOCK, (RC == 16 ? WH_MIN_BLOCKS_LARGE : RC == 9 ? WH_MIN_BLOCKS_MEDIUM : RC == 4 ? WH_MIN_BLOCKS_SMALL : WH_MIN_BLOCKS)) k_step(const __grid_constant__ KParams P) {
    __shared__ __align__(16) unsigned char smem[StageMem<GC, RC, FLAT>::BYTES];
    // Programmatic dependent launch (launch_step): the next step's blocks may be scheduled while
    // this grid drains, but touch no global memory before this grid has completed and flushed.
    // Both instructions are no-ops for a launch without the attribute.
    asm volatile("griddepcontrol.launch_dependents;");
    const Group<GC> g(P.G);
    const Tile<GC> t(P, g);
    const int R = RC ? RC : P.R;
    const env_t e = t.e;
    const uint32_t env_id = (uint32_t)P.env_id0 + e;
    EnvRegs s;
    if constexpr ((RC == 4 ? WH_ACT_BURST_SMALL : WH_ACT_BURST_OTHER) != 0 && PLAIN && !GREEDY && RC != 0) {
        // burst prefetch of this step's actions by the first wave (see WH_ACT_BURST_*); not for host-resident
        // (zero-copy) actions, whose launches carry WH_FLAG_NO_PDL
        constexpr uint32_t MIN_BLOCKS = RC == 16 ? WH_MIN_BLOCKS_LARGE : RC == 9 ? WH_MIN_BLOCKS_MEDIUM : WH_MIN_BLOCKS_SMALL;
        constexpr uint32_t WAVE_BLOCKS = 148 * MIN_BLOCKS;
        if (blockIdx.x < WAVE_BLOCKS && !(P.flags & WH_FLAG_NO_PDL)) {
            const uintptr_t a = (uintptr_t)P.actions, a_al = a & ~(uintptr_t)127;
            size_t lines = (a + (size_t)P.N * (RC * 4) - a_al + 127) >> 7;
            if (lines > ((size_t)WH_ACT_BURST_MAX_MB << 13)) lines = (size_t)WH_ACT_BURST_MAX_MB << 13;
            const uint32_t nthreads = (gridDim.x < WAVE_BLOCKS ? gridDim.x : WAVE_BLOCKS) * BLOCK;
            for (size_t line = (size_t)blockIdx.x * BLOCK + threadIdx.x; line < lines; line += nthreads)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a_al + line * 128));
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    load_env<GC, KEEP>(P, g, e, R, (RC ? 4 * GC : P.P), s);
    int4 acc4 = make_int4(0, 0, 0, 0);
    bool acc_dirty = false;
    constexpr bool EARLY_ACC = RC != 4;
    if (EARLY_ACC && g.gl == 0) acc4 = ld_state<KEEP>(reinterpret_cast<const int4 *>(P.acc) + e);   // with the rest of the state, see account_episode

    // envs masked out of this step (BaseEnv.send_actions for a subset of the envs) compute along but write nothing
    const bool live = t.live && (PLAIN || !P.env_mask || P.env_mask[e] != 0);
    int act = -1, ord = -1;
    unsigned long long active0 = 0ull;
    if (GREEDY) {
        act = greedy_from_state<GC, RC>(P, g, R, env_id, s, active0);
        if (P.actions_out && live && g.gl < R) P.actions_out[e * R + g.gl] = act;
    } else if (g.gl < R) {
        act = (!PLAIN && (P.flags & WH_FLAG_COMPACT_IO)) ? (int)reinterpret_cast<const int8_t *>(P.actions)[e * R + g.gl]
                                                          : P.actions[e * R + g.gl];
        if (!PLAIN && P.order) ord = P.order[e * R + g.gl];
    }
    s.time += 1;                                                               // core.py:267
    do_moves<GC, RC>(P, g, R, s.A, act, ord, !PLAIN && !GREEDY && P.order != nullptr, s.pos16);
    const StepOut so = do_world(P, g, e, R, env_id, s, !PLAIN && !GREEDY && P.spawn_p != nullptr, active0, GREEDY);
    unsigned long long active = so.active;
    uint32_t tpos16 = so.tpos16;

    const bool done = s.time >= P.episode;                                     // core.py:438
    if (live) {
        if (g.gl < R) {                                                         // core.py:435
            if (!PLAIN && (P.flags & WH_FLAG_COMPACT_IO)) reinterpret_cast<uint8_t *>(P.rewards)[e * R + g.gl] = (uint8_t)so.reward;
            else P.rewards[e * R + g.gl] = so.reward;
        }
        if (g.gl == 0) P.dones[e] = done ? 1 : 0;
    }
    const bool auto_reset = (P.flags & WH_FLAG_AUTO_RESET) != 0;
    account_episode<EARLY_ACC>(P, g.gl == 0 && live, e, so, s, done, auto_reset, acc4, acc_dirty);
    if (acc_dirty) st_state<KEEP>(reinterpret_cast<int4 *>(P.acc) + e, acc4);
    int flavour = WH_OBS_STEP;
    bool meta = false;
    if (auto_reset && __any_sync(FULL, done)) {
        const unsigned long long a2 = do_reset(P, g, e, R, env_id, s, false, done && live);
        if (done) { active = a2; flavour = WH_OBS_RESET; }
        meta = true;
    }
    if (live) store_env<GC, KEEP>(P, g, e, R, (RC ? 4 * GC : P.P), s, meta);
    if constexpr (FLAT)   // RLlib-flattened float32 layout instead of the dict keys (separate instantiation)
        build_obs_flat<GC, RC>(P, g, e, R, s, active, tpos16, flavour, live, P.flat_out,
                               reinterpret_cast<float *>(StageMem<GC, RC, true>::mine(smem, g)),
                               reinterpret_cast<float *>(StageMem<GC, RC, true>::warp_area(smem)), t.env0);
    else if (P.obs.requests)
        build_obs<GC, RC, 3, GREEDY>(P, P.obs, g, e, R, s, active, tpos16, flavour, live, StageMem<GC, RC>::mine(smem, g),
                          StageMem<GC, RC>::warp_area(smem), t.env0);
}

// baseline/run.py:42-62 for `n_steps` iterations in ONE launch: greedy solver -> step, the state
// stays in registers between the steps and no observation is materialised (the solver reads the
// state it would have been shown). Leaves the state, the per-agent reward sums of these steps, the
// last step's done flags and the episode statistics exactly as n_steps wh_greedy_step launches do.
// LOWOCC: launch-sized batches run 64-thread blocks spread over the SMs with no register cap (as k_multi does).
template <int GC, int RC, bool LOWOCC = false>
__global__ void __launch_bounds__(LOWOCC ? 64 : BLOCK, LOWOCC ? 1 : (RC == 16 ? 4 : WH_ROLLOUT_MIN_BLOCKS)) k_rollout(const __grid_constant__ KParams P) {
    const Group<GC> g(P.G);
    const Tile<GC> t(P, g);
    const int R = RC ? RC : P.R;
    const env_t e = t.e;
    const uint32_t env_id = (uint32_t)P.env_id0 + e;
    EnvRegs s;
    load_env(P, g, e, R, (RC ? 4 * GC : P.P), s);
    const bool auto_reset = (P.flags & WH_FLAG_AUTO_RESET) != 0;
    float ret = 0.0f;
    bool done = false;
    int4 acc4 = make_int4(0, 0, 0, 0);
    bool acc_dirty = false;
    if (g.gl == 0) acc4 = reinterpret_cast<const int4 *>(P.acc)[e];
    unsigned long long active0 = 0ull;   // the active-request mask of the current state, carried from step to step
    for (int it = 0; it < P.n_steps; ++it) {
        const int act = greedy_from_state<GC, RC>(P, g, R, env_id, s, active0, it > 0);
        s.time += 1;                                                           // core.py:267
        do_moves<GC, RC>(P, g, R, s.A, act, -1, false, s.pos16);
        const StepOut so = do_world(P, g, e, R, env_id, s, false, active0, true);
        ret += so.reward;
        done = s.time >= P.episode;                                            // core.py:438
        account_episode<true>(P, g.gl == 0 && t.live, e, so, s, done, auto_reset, acc4, acc_dirty);
        active0 = so.active;
        if (auto_reset && __any_sync(FULL, done)) {
            const unsigned long long a2 = do_reset(P, g, e, R, env_id, s, false, done && t.live);
            if (done && t.live) active0 = a2;
        }
    }
    if (t.live) {
        store_env(P, g, e, R, (RC ? 4 * GC : P.P), s, true);
        if (acc_dirty) reinterpret_cast<int4 *>(P.acc)[e] = acc4;
        if (g.gl < R) P.rewards[e * R + g.gl] = ret;
        if (g.gl == 0) P.dones[e] = done ? 1 : 0;
    }
}

// `n_steps` consecutive env.step calls in ONE launch (wh_multi_step): the state stays in registers between
// the steps, every step's rewards / dones / observations are written exactly as n_steps single launches
// would write them — either into per-step slices of [T,N,...] tensors (WH_FLAG_PER_STEP_OUT) or over the
// resident [N,...] tensors. Actions: open-loop int32 [T,N,R], or the in-kernel greedy solver. Envs are
// warp-private, so the steps of different warps drift apart freely: no per-step launch ramp / tail, which
// is what bounds launch-sized batches (BASELINE configs[1]: 4 096 Small envs, configs[2]: 65 536 Medium).
// LOWOCC: the instantiation for launch-sized batches (64-thread blocks spread over the SMs): no register cap —
// occupancy is irrelevant there and the capped kernel spills (Small: 232 B at 40 registers).
template <int GC, int RC, bool GREEDY, bool LOWOCC = false>
__global__ void __launch_bounds__(LOWOCC ? 64 : BLOCK, LOWOCC ? 1 : (RC == 16 ? WH_MIN_BLOCKS_LARGE : WH_MULTI_MIN_BLOCKS)) k_multi(const __grid_constant__ KParams P) {
    __shared__ __align__(16) unsigned char smem[StageMem<GC, RC>::BYTES];
    const Group<GC> g(P.G);
    const Tile<GC> t(P, g);
    const int R = RC ? RC : P.R;
    const env_t e = t.e;
    const uint32_t env_id = (uint32_t)P.env_id0 + e;
    EnvRegs s;
    load_env(P, g, e, R, (RC ? 4 * GC : P.P), s);
    const bool auto_reset = (P.flags & WH_FLAG_AUTO_RESET) != 0;
    const bool per_step = (P.flags & WH_FLAG_PER_STEP_OUT) != 0;
    wh_obs o = P.obs;
    const bool with_obs = o.requests != nullptr;
    const int32_t *acts = P.actions;
    float *rew = P.rewards;
    uint8_t *dn = P.dones;
    const size_t NR = (size_t)P.N * R, N = (size_t)P.N;
    float ret = 0.0f;
    bool done = false;
    int4 acc4 = make_int4(0, 0, 0, 0);
    bool acc_dirty = false;
    if (g.gl == 0) acc4 = reinterpret_cast<const int4 *>(P.acc)[e];
    // open-loop actions: step t+1's action is loaded while step t runs (the load is the first thing on a
    // step's dependent chain otherwise: an L2 / HBM round trip per step)
    int act_next = -1;
    if (!GREEDY && g.gl < R) act_next = acts[e * R + g.gl];
    unsigned long long active_cur = 0ull;   // GREEDY: the active-request mask of the current state, carried from step to step
    constexpr bool CARRY = LOWOCC || RC != 16;   // (the 80-register Large throughput kernel would spill the two registers)
    for (int it = 0; it < P.n_steps; ++it) {
        int act = -1;
        unsigned long long active0 = active_cur;
        if (GREEDY) act = greedy_from_state<GC, RC>(P, g, R, env_id, s, active0, CARRY && it > 0);
        else {
            act = act_next;
            acts += NR;
            if (it + 1 < P.n_steps && g.gl < R) act_next = acts[e * R + g.gl];
        }
        s.time += 1;                                                           // core.py:267
        do_moves<GC, RC>(P, g, R, s.A, act, -1, false, s.pos16);
        const StepOut so = do_world(P, g, e, R, env_id, s, false, active0, GREEDY);
        done = s.time >= P.episode;                                            // core.py:438
        if (per_step) {
            if (t.live && g.gl < R) rew[e * R + g.gl] = so.reward;              // core.py:435
            if (t.live && g.gl == 0) dn[e] = done ? 1 : 0;
        } else {
            ret += so.reward;
        }
        account_episode<true>(P, g.gl == 0 && t.live, e, so, s, done, auto_reset, acc4, acc_dirty);
        unsigned long long active = so.active;
        int flavour = WH_OBS_STEP;
        if (auto_reset && __any_sync(FULL, done)) {
            const unsigned long long a2 = do_reset(P, g, e, R, env_id, s, false, done && t.live);
            if (done) { active = a2; flavour = WH_OBS_RESET; }
        }
        active_cur = (done && !t.live) ? so.active : active;   // (do_reset leaves the state of a dead lane's env alone)
        if (with_obs) {
            build_obs<GC, RC, 3, GREEDY>(P, o, g, e, R, s, active, so.tpos16, flavour, t.live, StageMem<GC, RC>::mine(smem, g),
                              StageMem<GC, RC>::warp_area(smem), t.env0);
            __syncwarp();                                                      // staging is reused by the next step
        }
        if (per_step) {
            rew += NR; dn += N;
            if (with_obs) {
                o.num_agents += NR; o.self_position += 2 * NR; o.self_availability += NR; o.self_delivery_target += 2 * NR;
                o.other_positions += 2 * NR * (R - 1); o.other_availabilities += NR * (R - 1);
                o.other_delivery_targets += 2 * NR * (R - 1); o.requests += 4 * NR * R;
            }
        }
    }
    if (t.live) {
        store_env(P, g, e, R, (RC ? 4 * GC : P.P), s, true);
        if (acc_dirty) reinterpret_cast<int4 *>(P.acc)[e] = acc4;
        if (!per_step) {
            if (g.gl < R) P.rewards[e * R + g.gl] = ret;
            if (g.gl == 0) P.dones[e] = done ? 1 : 0;
        }
    }
}

// k_multi for LAUNCH-SIZED batches (fewer env tiles than the GPU has warp schedulers; BASELINE configs[1]:
// 4 096 Small envs = 512 tiles on 592 schedulers). There a step is one ~1 000-instruction dependent chain per
// warp at a single-warp IPC of ~0.2 (profiles/r02_ncu_multi_small4096.txt), and most schedulers idle. The
// loop-carried part of that chain is only solver -> moves -> world; the observation of step t hangs off it.
// So a block = one env tile worked on by 1 + NOBS warps: warp 0 runs the step logic and hands the post-step
// state of every lane (one 16-byte word: cell, delivery target, agent count, observation flavour, the
// lane's four pickup targets, the active-request mask) to the observation warp(s) through a two-deep
// shared-memory ring guarded by named barriers (bar.arrive / bar.sync: full[2], empty[2]); the observation
// warps build and store step t's observation while warp 0 is already in step t+1. With NOBS = 2 the keys are
// split (build_obs PART 1: num_agents / self_* / requests, PART 2: other_*). No register cap: occupancy is
// irrelevant in this regime, and the capped k_multi spills (Small: 232 B at 40 registers).
// Outputs are bit-identical to k_multi (same device functions, same order of stores per address).
// barrier ids are immediates (two ring slots -> ids 1,2 = full, 3,4 = empty), so ptxas reserves 5 barriers, not 16
template <int ID>
__device__ __forceinline__ void named_bar_sync_i(int n) { asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(n) : "memory"); }
template <int ID>
__device__ __forceinline__ void named_bar_arrive_i(int n) { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) {
    if (id == 1) named_bar_sync_i<1>(n); else if (id == 2) named_bar_sync_i<2>(n); else if (id == 3) named_bar_sync_i<3>(n); else named_bar_sync_i<4>(n);
}
__device__ __forceinline__ void named_bar_arrive(int id, int n) {
    if (id == 1) named_bar_arrive_i<1>(n); else if (id == 2) named_bar_arrive_i<2>(n); else if (id == 3) named_bar_arrive_i<3>(n); else named_bar_arrive_i<4>(n);
}
// The arrive that frees a ring slot must not be issued before the slot has been READ. `w0` is the first word loaded
// from the slot; its bit 31 is never set (cell 16 | target 8 | agent count 6 | flavour 1 bits), so the thread count
// below always equals n — but only the loaded value says so, which makes the arrive data-dependent on the load.
__device__ __forceinline__ void named_bar_arrive_after(int id, int n, uint32_t w0) {
    named_bar_arrive(id, n + (int)(w0 >> 31));
}

template <int GC, int RC, int PART>
__device__ __forceinline__ void multi_ws_obs_loop(const KParams &P, const Group<GC> &g, const Tile<GC> &t, const uint4 (*ring)[32],
                                                  unsigned char *smem, int nthreads) {
    constexpr int R = RC;
    const env_t e = t.e;
    const bool per_step = (P.flags & WH_FLAG_PER_STEP_OUT) != 0;
    const size_t NR = (size_t)P.N * R;
    wh_obs o = P.obs;
    const int T = P.n_steps;
    for (int it = 0; it < T; ++it) {
        const int b = it & 1;
        named_bar_sync(1 + b, nthreads);                                        // full[b]
        const uint4 h = ring[b][g.lane];
        if (it + 2 < T) named_bar_arrive_after(3 + b, nthreads, h.x);                      // empty[b]
        EnvRegs s;
        s.pos16 = h.x & 0xFFFFu;
        s.atgt = (int)(int8_t)((h.x >> 16) & 0xFFu);
        s.A = (int)((h.x >> 24) & 0x3Fu);
        const int flavour = (h.x >> 30) & 1u ? WH_OBS_RESET : WH_OBS_STEP;
        s.pt4 = h.y;
        s.tmr = make_uint2(0u, 0u); s.time = 0; s.ep = 0;                       // not read by build_obs
        const unsigned long long active = (unsigned long long)h.z | ((unsigned long long)h.w << 32);
#if WH_WS_DIAG == 1   // tuning only: how fast is the logic warp alone?
        if (s.pos16 == 0x12345u) o.num_agents[0] = (int)active + flavour;
#else
        build_obs<GC, RC, PART>(P, o, g, e, R, s, active, target_cell16<GC>(P, s.atgt), flavour, t.live,
                                StageMem<GC, RC>::mine(smem, g), StageMem<GC, RC>::warp_area(smem), t.env0);
        __syncwarp();                                                           // staging is reused by the next step
#endif
        if (per_step) {
            o.num_agents += NR; o.self_position += 2 * NR; o.self_availability += NR; o.self_delivery_target += 2 * NR;
            o.other_positions += 2 * NR * (R - 1); o.other_availabilities += NR * (R - 1);
            o.other_delivery_targets += 2 * NR * (R - 1); o.requests += 4 * NR * R;
        }
    }
}

template <int GC, int RC, bool GREEDY, int NOBS>
__global__ void __launch_bounds__(32 * (1 + NOBS)) k_multi_ws(const __grid_constant__ KParams P) {
    static_assert(RC != 0 && (NOBS == 1 || NOBS == 2), "variant kernels only");
    __shared__ __align__(16) unsigned char smem[StageMem<GC, RC>::BYTES / (BLOCK / 32) * (1 + NOBS)];
    __shared__ __align__(16) uint4 ring[2][32];
    constexpr int NT = 32 * (1 + NOBS);
    constexpr int R = RC;
    const int role = threadIdx.x >> 5;
    const Group<GC> g(P.G);
    const Tile<GC> t(P, g, blockIdx.x);
    if (role != 0) {
        if (NOBS == 1) multi_ws_obs_loop<GC, RC, 3>(P, g, t, ring, smem, NT);
        else if (role == 1) multi_ws_obs_loop<GC, RC, 1>(P, g, t, ring, smem, NT);
        else multi_ws_obs_loop<GC, RC, 2>(P, g, t, ring, smem, NT);
        return;
    }
    const env_t e = t.e;
    const uint32_t env_id = (uint32_t)P.env_id0 + e;
    EnvRegs s;
    load_env(P, g, e, R, 4 * GC, s);
    const bool auto_reset = (P.flags & WH_FLAG_AUTO_RESET) != 0;
    const bool per_step = (P.flags & WH_FLAG_PER_STEP_OUT) != 0;
    const int32_t *acts = P.actions;
    float *rew = P.rewards;
    uint8_t *dn = P.dones;
    const size_t NR = (size_t)P.N * R, N = (size_t)P.N;
    float ret = 0.0f;
    bool done = false;
    int4 acc4 = make_int4(0, 0, 0, 0);
    bool acc_dirty = false;
    if (g.gl == 0) acc4 = reinterpret_cast<const int4 *>(P.acc)[e];
    // open-loop actions: step t+1's action is loaded while step t runs (the load is the first thing on a
    // step's dependent chain otherwise: an L2 / HBM round trip per step)
    int act_next = -1;
    if (!GREEDY && g.gl < R) act_next = acts[e * R + g.gl];
    unsigned long long active_cur = 0ull;   // GREEDY: the active-request mask of the current state, carried from step to step
    for (int it = 0; it < P.n_steps; ++it) {
        int act = -1;
        unsigned long long active0 = active_cur;
        if (GREEDY) act = greedy_from_state<GC, RC>(P, g, R, env_id, s, active0, it > 0);
        else {
            act = act_next;
            acts += NR;
            if (it + 1 < P.n_steps && g.gl < R) act_next = acts[e * R + g.gl];
        }
        s.time += 1;                                                           // core.py:267
#if WH_WS_DIAG == 2   // tuning only: how fast are the observation warps alone?
        StepOut so; so.reward = 0.f; so.active = active_mask(g, s.pt4); so.tpos16 = 0; so.npick = so.ndeliv = so.nexp = 0;
        if (act == 12345) s.pos16 = 0;
#else
        do_moves<GC, RC>(P, g, R, s.A, act, -1, false, s.pos16);
        const StepOut so = do_world(P, g, e, R, env_id, s, false, active0, GREEDY);
#endif
        done = s.time >= P.episode;                                            // core.py:438
        unsigned long long active = so.active;
        uint32_t flav = 0u;
        // the hand-over comes first: everything after it is off the observation warps' critical path
        const bool resets = auto_reset && __any_sync(FULL, done);
        if (!resets) {
            active_cur = active;
            const int b = it & 1;
            if (it >= 2) named_bar_sync(3 + b, NT);                             // empty[b]
            ring[b][g.lane] = make_uint4((s.pos16 & 0xFFFFu) | (((uint32_t)s.atgt & 0xFFu) << 16) | ((uint32_t)s.A << 24),
                                         s.pt4, (uint32_t)active, (uint32_t)(active >> 32));
            named_bar_arrive(1 + b, NT);                                        // full[b]
        }
        if (per_step) {
            if (t.live && g.gl < R) rew[e * R + g.gl] = so.reward;              // core.py:435
            if (t.live && g.gl == 0) dn[e] = done ? 1 : 0;
        } else {
            ret += so.reward;
        }
        account_episode<true>(P, g.gl == 0 && t.live, e, so, s, done, auto_reset, acc4, acc_dirty);
        if (resets) {
            const unsigned long long a2 = do_reset(P, g, e, R, env_id, s, false, done && t.live);
            if (done) { active = a2; flav = 1u; }
            active_cur = (done && !t.live) ? so.active : active;   // (do_reset leaves the state of a dead lane's env alone)
            const int b = it & 1;
            if (it >= 2) named_bar_sync(3 + b, NT);                             // empty[b]
            ring[b][g.lane] = make_uint4((s.pos16 & 0xFFFFu) | (((uint32_t)s.atgt & 0xFFu) << 16) | ((uint32_t)s.A << 24) | (flav << 30),
                                         s.pt4, (uint32_t)active, (uint32_t)(active >> 32));
            named_bar_arrive(1 + b, NT);                                        // full[b]
        }
        if (per_step) { rew += NR; dn += N; }
    }
    if (t.live) {
        store_env(P, g, e, R, 4 * GC, s, true);
        if (acc_dirty) reinterpret_cast<int4 *>(P.acc)[e] = acc4;
        if (!per_step) {
            if (g.gl < R) P.rewards[e * R + g.gl] = ret;
            if (g.gl == 0) P.dones[e] = done ? 1 : 0;
        }
    }
}

// Warehouse.reset — core.py:167-260
template <int GC, int RC>
__global__ void __launch_bounds__(BLOCK) k_reset(const __grid_constant__ KParams P) {
    __shared__ __align__(16) unsigned char smem[StageMem<GC, RC>::BYTES];
    const Group<GC> g(P.G);
    const Tile<GC> t(P, g);
    const int R = RC ? RC : P.R;
    const env_t e = t.e;
    EnvRegs s;
    load_env(P, g, e, R, (RC ? 4 * GC : P.P), s);
    const bool doit = t.live && (!P.env_mask || P.env_mask[e]);
    const unsigned long long active =
        do_reset(P, g, e, R, (uint32_t)P.env_id0 + e, s, P.r_agent_pos != nullptr, doit);
    if (doit) {
        store_env(P, g, e, R, (RC ? 4 * GC : P.P), s, true);
        if (g.gl == 0) reinterpret_cast<int4 *>(P.acc)[e] = make_int4(0, 0, 0, 0);
    }
    if (P.obs.requests)
        build_obs<GC, RC>(P, P.obs, g, e, R, s, active, 0u, WH_OBS_RESET, doit, StageMem<GC, RC>::mine(smem, g),
                          StageMem<GC, RC>::warp_area(smem), t.env0);
}

// observation build alone — core.py:224-260 / 371-432
template <int GC, int RC>
__global__ void __launch_bounds__(BLOCK) k_obs(const __grid_constant__ KParams P) {
    __shared__ __align__(16) unsigned char smem[StageMem<GC, RC>::BYTES];
    const Group<GC> g(P.G);
    const Tile<GC> t(P, g);
    const int R = RC ? RC : P.R;
    EnvRegs s;
    load_env(P, g, t.e, R, (RC ? 4 * GC : P.P), s);
    const unsigned long long active = active_mask(g, s.pt4);
    build_obs<GC, RC>(P, P.obs, g, t.e, R, s, active, target_cell16<GC>(P, s.atgt), P.flavour, t.live, StageMem<GC, RC>::mine(smem, g),
                      StageMem<GC, RC>::warp_area(smem), t.env0);
}

// RLlib-flattened float32 observations from the resident state
template <int GC, int RC>
__global__ void __launch_bounds__(BLOCK) k_obs_flat(const __grid_constant__ KParams P) {
    __shared__ __align__(16) unsigned char smem[StageMem<GC, RC, true>::BYTES];
    const Group<GC> g(P.G);
    const Tile<GC> t(P, g);
    const int R = RC ? RC : P.R;
    EnvRegs s;
    load_env(P, g, t.e, R, (RC ? 4 * GC : P.P), s);
    const unsigned long long active = active_mask(g, s.pt4);
    build_obs_flat<GC, RC>(P, g, t.e, R, s, active, target_cell16<GC>(P, s.atgt), P.flavour, t.live,
                           P.flat_out, reinterpret_cast<float *>(StageMem<GC, RC, true>::mine(smem, g)),
                           reinterpret_cast<float *>(StageMem<GC, RC, true>::warp_area(smem)), t.env0);
}

// WarehouseRandomGreedySolver.compute_action on observation tensors — solvers.py:27-58.
// LR lanes cooperate on one AGENT ROW (4 for Large, 3 for Medium, 1 for Small): lane j loads
// requests j, j+LR, ... with 128-bit loads (each load instruction covers 16*LR contiguous bytes of
// every row in the warp, all loads of a lane are issued back to back), keeps its running
// (distance<<8 | r) minimum, and the LR partial minima are combined by shuffles: the L1 argmin with
// first-minimum tie-break of solvers.py:53-58.
template <int RC, int LR>
__global__ void __launch_bounds__(BLOCK) k_greedy(const __grid_constant__ KParams P) {
    const int R = RC ? RC : P.R;
    constexpr int RPW = 32 / LR;                       // agent rows per warp
    const int lane = threadIdx.x & 31, gi = lane / LR, j = lane - gi * LR;
    const bool ghost = gi >= RPW;
    const long long rows = P.N * R;
    const long long warp = ((long long)blockIdx.x * BLOCK + threadIdx.x) >> 5;
    const long long row_raw = warp * RPW + gi;
    const bool live = !ghost && row_raw < rows;
    const long long row = row_raw < rows ? row_raw : rows - 1;
    const long long e = row / R;
    const int a = (int)(row - e * R);
    const wh_obs &o = P.obs;
    const int4 *req = reinterpret_cast<const int4 *>(o.requests) + row * R;
    const int2 sp = reinterpret_cast<const int2 *>(o.self_position)[row];
    const int2 stg = reinterpret_cast<const int2 *>(o.self_delivery_target)[row];
    const int avail = o.self_availability[row];
    const int A = P.g_num_agents[e];
    uint32_t best = 0xffffffffu, bcell = 0;
    constexpr int RPL = RC ? (RC + LR - 1) / LR : 1;   // requests per lane (compile-time R)
    if (RC) {
        int4 rq[RPL];
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
            const int r = j + k * LR;
            rq[k] = (r < RC) ? __ldcs(req + r) : make_int4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
            const int r = j + k * LR;
            const uint32_t key = (r < RC) ? (((uint32_t)(abs(sp.x - rq[k].x) + abs(sp.y - rq[k].y)) << 8) | (uint32_t)r)
                                          : 0xffffffffu;                       // solvers.py:54-57
            if (key < best) { best = key; bcell = (uint32_t)(rq[k].x & 0xFFFF) | ((uint32_t)rq[k].y << 16); }
        }
    } else {
        for (int r = j; r < R; r += LR) {
            const int4 q = __ldcs(req + r);
            const uint32_t key = ((uint32_t)(abs(sp.x - q.x) + abs(sp.y - q.y)) << 8) | (uint32_t)r;
            if (key < best) { best = key; bcell = (uint32_t)(q.x & 0xFFFF) | ((uint32_t)q.y << 16); }
        }
    }
    if (LR == 2 || LR == 4) {                                                  // solvers.py:58 argmin
#pragma unroll
        for (int m = 1; m < LR; m <<= 1) {
            const uint32_t ob = __shfl_xor_sync(FULL, best, m), oc = __shfl_xor_sync(FULL, bcell, m);
            if (ob < best) { best = ob; bcell = oc; }
        }
    } else if (LR > 1) {
        const int base = ghost ? 0 : gi * LR;
        uint32_t b0 = best, c0 = bcell;
#pragma unroll
        for (int k = 0; k < LR; ++k) {
            const uint32_t ob = __shfl_sync(FULL, best, base + k), oc = __shfl_sync(FULL, bcell, base + k);
            if (ob < b0) { b0 = ob; c0 = oc; }
        }
        best = b0; bcell = c0;
    }
    int tx = (int)(bcell & 0xFFFF), ty = (int)(bcell >> 16);
    if (avail == 0) { tx = stg.x; ty = stg.y; }                                 // solvers.py:33-34
    const int sx = max(-1, min(1, tx - sp.x)), sy = max(-1, min(1, ty - sp.y)); // solvers.py:41
    int action = (sx + 1) * 3 + (sy + 1);                                      // solvers.py:47-49
    if (P.is_random) {                                                         // solvers.py:44-45 replay
        if (P.is_random[row]) action = P.random_actions[row];
    } else if (P.rand_thr) {
        uint32_t u0, u1;
        philox4x32_10((uint32_t)(P.env_id0 + e), (uint32_t)P.g_episode[e], (uint32_t)P.g_time[e],
                      (uint32_t)a | CTR_SOLVER_TAG, P.solver_seed, u0, u1);
        if ((unsigned long long)u0 < P.rand_thr) action = (int)bounded(u1, 9u);
    }
    if (live && j == 0) P.actions_out[row] = (a < A) ? action : -1;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Shape { int G, RC; };

static int fill_params(const wh_config *cfg, KParams &K, Shape &sh) {
    if (!cfg) return WH_E_ARG;
    memset(&K, 0, sizeof(K));
    K.R = cfg->num_requests; K.dim = cfg->area_dimension; K.L = cfg->num_racks;
    if (K.L < 1 || K.L > WH_MAX_RACKS) return WH_E_CONFIG;
    K.P = 4 * K.L * K.L; K.D = 4 * (K.dim - 4);
    K.episode = cfg->episode_duration; K.wait = cfg->pickup_wait_duration;
    K.null_pos = K.dim / 2;                                                   // core.py:107
    K.max_agents = cfg->max_num_agents > 0 ? cfg->max_num_agents : K.R;
    K.random_agents = cfg->random_num_agents;
    K.regular_racks = 1;
    for (int i = 0; i < K.L; ++i) { K.racks[i] = cfg->racks[i]; if (cfg->racks[i] != 4 * (i + 1)) K.regular_racks = 0; }
    if (K.R < 2 || K.R > 32 || K.P > 64 || K.D > 64 || K.D < 1 || K.dim > 127 || K.R > K.P || K.R > K.D ||
        K.max_agents > K.R || K.wait > 32767 || K.wait < 1)
        return WH_E_CONFIG;
    for (int i = 0; i < K.L; ++i) if (K.racks[i] < 1 || K.racks[i] >= K.dim) return WH_E_CONFIG;
    int G = K.R;                         // one lane per agent / request ...
    if ((K.P + 3) / 4 > G) G = (K.P + 3) / 4;   // ... and per 4 pickup points
    if (G < 2) G = 2;
    K.G = G;
    K.invL = (256 + K.L - 1) / K.L;
    sh.G = G; sh.RC = 0;
    // compile-time-shaped kernels: exactly the reference variants' geometry (R == G, P == 4G, their
    // area dimension and regular racks 4,8,12,..; see Geo<GC>); anything else runs the runtime kernels
    if (K.regular_racks) {
        if (K.R == 4 && K.L == 2 && K.dim == 12) sh.RC = 4;        // WarehouseSmall  (variants.py:25-32)
        if (K.R == 9 && K.L == 3 && K.dim == 16) sh.RC = 9;        // WarehouseMedium (variants.py:40-47): 3 envs per warp
        if (K.R == 16 && K.L == 4 && K.dim == 20) sh.RC = 16;      // WarehouseLarge  (variants.py:55-62)
    }
    return 0;
}

static void set_state(KParams &K, const wh_state *st) {
    K.agent_pos = st->agent_pos; K.agent_tgt = st->agent_tgt; K.pickup_tgt = st->pickup_tgt;
    K.pickup_timer = st->pickup_timer; K.time = st->time; K.num_agents = st->num_agents;
    K.episode_ctr = st->episode; K.acc = st->acc;
}

static bool state_ok(const wh_state *st) {
    return st && st->agent_pos && st->agent_tgt && st->pickup_tgt && st->pickup_timer && st->time &&
           st->num_agents && st->episode && st->acc;
}

static bool obs_ok(const wh_obs *o) {
    return o && o->num_agents && o->self_position && o->self_availability && o->self_delivery_target &&
           o->other_positions && o->other_availabilities && o->other_delivery_targets && o->requests;
}

enum Kind { K_STEP, K_GSTEP, K_STEP_FLAT, K_RESET, K_OBS, K_OBS_FLAT, K_GREEDY, K_ROLLOUT, K_MULTI, K_GMULTI };

// Step kernels are launched back to back, one per env.step, each depending on the one before
// through the state tensors. Programmatic stream serialization lets the driver schedule step
// k+1's blocks (which park at griddepcontrol.wait) while step k's last blocks finish, hiding the
// launch gap (a few us: ~7 % of a 40 us Small / Medium-65536 step). WH_B200_PDL=0 turns it off.
static bool pdl_enabled() {
    static const bool on = [] { const char *e = getenv("WH_B200_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

static void launch_step(void (*kern)(const KParams), unsigned grid, size_t dyn, cudaStream_t s, const KParams &K) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(BLOCK); cfg.dynamicSmemBytes = dyn; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_enabled() && !(K.flags & WH_FLAG_NO_PDL)) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, K);
}

template <int GC, int RC>
static void launch_kind(Kind kind, const KParams &K, cudaStream_t s) {
    const int G = GC ? GC : K.G;
    const long long epw = 32 / G, warps = (K.N + epw - 1) / epw;
    const unsigned grid = (unsigned)((warps * 32 + BLOCK - 1) / BLOCK);
    // unused dynamic shared memory caps the resident blocks per SM where fewer, fatter warps measured
    // faster than what the register count alone would allow (Large: 3 blocks, profiles/README.md)
    const size_t dyn = RC == 16 ? WH_LARGE_DYN_SMEM : 0;
    if (dyn) {   // static + dynamic > 48 KB needs the opt-in, once per device (function attributes are per device)
        static std::atomic<bool> done[64];   // setting the attribute twice is harmless: a flag per device is enough
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !done[dev].load(std::memory_order_acquire)) {
            cudaFuncSetAttribute(k_step<GC, RC, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            cudaFuncSetAttribute(k_step<GC, RC, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            cudaFuncSetAttribute(k_step<GC, RC, false, false, RC != 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            cudaFuncSetAttribute(k_step<GC, RC, true, false, RC != 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            cudaFuncSetAttribute(k_step<GC, RC, false, false, RC != 0, (RC != 0 ? 1 : 0)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            cudaFuncSetAttribute(k_step<GC, RC, true, false, RC != 0, (RC != 0 ? 1 : 0)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            constexpr int KP = (RC == 16 && WH_KEEP_PART_LARGE) ? 2 : (RC != 0 ? 1 : 0);   // the partial-policy instantiation, where it exists
            cudaFuncSetAttribute(k_step<GC, RC, false, false, RC != 0, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            cudaFuncSetAttribute(k_step<GC, RC, true, false, RC != 0, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            done[dev].store(true, std::memory_order_release);
        }
    }
    const bool plain = RC != 0 && !(K.flags & WH_FLAG_COMPACT_IO) && !K.order && !K.spawn_p && !K.env_mask;
    // KEEP variant: the whole state of this launch (narrow state + episode counters) fits the keep budget
    const double state_mb = (double)K.N * (3.0 * K.R + 3.0 * K.P + 9.0 + 16.0) / 1048576.0;
    static const double keep_mb = [] { const char *v = getenv("WH_B200_KEEP_MB"); return v ? atof(v) : (double)WH_KEEP_MAX_MB; }();
    // level 1 while the state fits the budget; Medium also pays at level 2 (all but the timers) up to twice the budget
    constexpr bool PART = RC == 9 || (RC == 16 && WH_KEEP_PART_LARGE);
    const int keep = !(plain && (RC == 9 || RC == 16)) ? 0 : state_mb <= keep_mb ? 1 : (PART && state_mb <= keep_mb * ((double)WH_KEEP_PART_MAX_MB / WH_KEEP_MAX_MB)) ? 2 : 0;
    constexpr bool KV = RC == 9 || RC == 16;      // the KEEP instantiations exist for Medium and Large only
    constexpr int K1 = KV ? 1 : 0, K2 = PART ? 2 : 0;
    constexpr int KSM = RC == 4 ? WH_KEEP_SMALL : 0;
    switch (kind) {
    case K_STEP:
        if (KSM && plain) launch_step(k_step<GC, RC, false, false, RC != 0, KSM>, grid, dyn, s, K);
        else if (keep == 2) launch_step(k_step<GC, RC, false, false, KV, K2>, grid, dyn, s, K);
        else if (keep) launch_step(k_step<GC, RC, false, false, KV, K1>, grid, dyn, s, K);
        else if (plain) launch_step(k_step<GC, RC, false, false, RC != 0>, grid, dyn, s, K);
        else launch_step(k_step<GC, RC, false, false>, grid, dyn, s, K);
        break;
    case K_GSTEP:
        if (keep == 2) launch_step(k_step<GC, RC, true, false, KV, K2>, grid, dyn, s, K);
        else if (keep) launch_step(k_step<GC, RC, true, false, KV, K1>, grid, dyn, s, K);
        else if (plain) launch_step(k_step<GC, RC, true, false, RC != 0>, grid, dyn, s, K);
        else launch_step(k_step<GC, RC, true, false>, grid, dyn, s, K);
        break;
    case K_STEP_FLAT:
        if (keep == 2) launch_step(k_step<GC, RC, false, true, KV, K2>, grid, 0, s, K);
        else if (keep) launch_step(k_step<GC, RC, false, true, KV, K1>, grid, 0, s, K);
        else if (plain) launch_step(k_step<GC, RC, false, true, RC != 0>, grid, 0, s, K);
        else launch_step(k_step<GC, RC, false, true>, grid, 0, s, K);
        break;
    case K_RESET: k_reset<GC, RC><<<grid, BLOCK, 0, s>>>(K); break;
    case K_OBS: k_obs<GC, RC><<<grid, BLOCK, 0, s>>>(K); break;
    case K_OBS_FLAT: k_obs_flat<GC, RC><<<grid, BLOCK, 0, s>>>(K); break;
    case K_ROLLOUT: {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        if (warps < (long long)sms * 16) k_rollout<GC, RC, true><<<(unsigned)((warps + 1) / 2), 64, 0, s>>>(K);
        else k_rollout<GC, RC><<<grid, BLOCK, 0, s>>>(K);
        break;
    }
    case K_MULTI:
    case K_GMULTI: {
        // launch-sized batches: with fewer warps than the GPU has schedulers, spread them over the SMs
        // (2-warp blocks) instead of packing 8 per block. (Giving each warp fewer environments to put more
        // warps in flight was measured and is worse: 4 096 Small envs 2.7 -> 5.5 us per step — a step is one
        // long dependent chain per warp, so extra warps only add instructions.)
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        const int force = (K.flags >> 4) & 7;                // WH_FLAG_MULTI_KERNEL
        if constexpr (RC != 0) {
            // launch-sized batches that write observations: the warp-specialised kernel (see k_multi_ws);
            // WH_B200_MULTI_WS = 0 (off) / 1 / 2 observation warps per env tile
            static const int ws_default = [] { const char *v = getenv("WH_B200_MULTI_WS"); return v ? atoi(v) : WH_MULTI_WS_DEFAULT; }();
            const int ws = force >= 3 ? force - 2 : force ? 0 : (K.obs.requests && warps < (long long)sms * WH_MULTI_WS_MAX_TILES_PER_SM) ? ws_default : 0;
            if (ws > 0) {
                const bool gr = kind == K_GMULTI;
                if (ws == 1) { if (gr) k_multi_ws<GC, RC, true, 1><<<(unsigned)warps, 64, 0, s>>>(K); else k_multi_ws<GC, RC, false, 1><<<(unsigned)warps, 64, 0, s>>>(K); }
                else { if (gr) k_multi_ws<GC, RC, true, 2><<<(unsigned)warps, 96, 0, s>>>(K); else k_multi_ws<GC, RC, false, 2><<<(unsigned)warps, 96, 0, s>>>(K); }
                break;
            }
        }
        const unsigned threads = force == 1 ? BLOCK : force == 2 ? 64 : warps >= (long long)sms * 16 ? BLOCK : 64;
        const unsigned mgrid = (unsigned)((warps * 32 + threads - 1) / threads);
        if (dyn) {
            static std::atomic<bool> once[2];
            if (!once[kind == K_GMULTI].exchange(true)) {
                cudaFuncSetAttribute(k_multi<GC, RC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
                cudaFuncSetAttribute(k_multi<GC, RC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
            }
        }
        if (threads == 64) {
            if (kind == K_GMULTI) k_multi<GC, RC, true, true><<<mgrid, threads, 0, s>>>(K);
            else k_multi<GC, RC, false, true><<<mgrid, threads, 0, s>>>(K);
        } else if (kind == K_GMULTI) k_multi<GC, RC, true><<<mgrid, threads, dyn, s>>>(K);
        else k_multi<GC, RC, false><<<mgrid, threads, dyn, s>>>(K);
        break;
    }
    default: break;
    }
}

template <int RC, int LR>
static void launch_greedy(const KParams &K, cudaStream_t s) {
    const long long rpw = 32 / LR, warps = (K.N * K.R + rpw - 1) / rpw;   // agent rows per warp
    const unsigned grid = (unsigned)((warps * 32 + BLOCK - 1) / BLOCK);
    k_greedy<RC, LR><<<grid, BLOCK, 0, s>>>(K);
}

#ifdef WH_WITH_TPE
// Thread-per-environment kernels (tools/experimental/wh_tpe.cuh) for the default path of Small / Medium.
template <int RC>
static void launch_tpe(Kind kind, const KParams &K, cudaStream_t s) {
    constexpr int WARPS = 4;
    const size_t smem = (size_t)WARPS * Tpe<RC>::BYTES;
    const unsigned grid = (unsigned)((K.N + 32 * WARPS - 1) / (32 * WARPS));
    if (kind == K_GSTEP) {
        static const cudaError_t once = cudaFuncSetAttribute(k_step_tpe<RC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        (void)once;
        k_step_tpe<RC, true><<<grid, 32 * WARPS, smem, s>>>(K);
    } else {
        static const cudaError_t once = cudaFuncSetAttribute(k_step_tpe<RC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        (void)once;
        k_step_tpe<RC, false><<<grid, 32 * WARPS, smem, s>>>(K);
    }
}

static bool tpe_enabled() {
    static const bool on = getenv("WH_ENABLE_TPE") != nullptr;
    return on;
}
#endif

static int launch(Kind kind, const KParams &K, const Shape &sh, void *stream) {
    if (K.N <= 0) return 0;
    if ((unsigned long long)K.N * (K.R * K.R > 2 * K.P ? K.R * K.R : 2 * K.P) >= (1ull << 32)) return WH_E_CONFIG;   // 32-bit in-launch indexing: shard the batch
    cudaStream_t s = (cudaStream_t)stream;
    if (kind == K_GREEDY) {
        if (K.R == 4) launch_greedy<4, 1>(K, s);
        else if (K.R == 9) launch_greedy<9, 3>(K, s);
        else if (K.R == 16) launch_greedy<16, 4>(K, s);
        else launch_greedy<0, 4>(K, s);
    }
#ifdef WH_WITH_TPE
    else if ((kind == K_STEP || kind == K_GSTEP) && (sh.RC == 4 || sh.RC == 9) && !K.order && !K.spawn_p &&
             K.regular_racks && tpe_enabled()) {
        if (sh.RC == 4) launch_tpe<4>(kind, K, s);
        else launch_tpe<9>(kind, K, s);
    }
#endif
    else if (sh.RC == 4) launch_kind<4, 4>(kind, K, s);
    else if (sh.RC == 9) launch_kind<9, 9>(kind, K, s);
    else if (sh.RC == 16) launch_kind<16, 16>(kind, K, s);
    else launch_kind<0, 0>(kind, K, s);
    return (int)cudaGetLastError();
}

}  // namespace wh

using namespace wh;

#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return (int)_e; } while (0)

// ---------------------------------------------------------------------------------------------
// C ABI — layer 1
// ---------------------------------------------------------------------------------------------
extern "C" {

int wh_version(void) { return 101; }

const char *wh_error_string(int code) {
    if (code == 0) return "ok";
    if (code == WH_E_CONFIG) return "unsupported warehouse configuration (limits: 2<=R<=32, P<=64, D<=64, dim<=127)";
    if (code == WH_E_ARG) return "NULL or inconsistent argument";
    if (code == WH_E_NCCL) return "ncclAllReduce could not be resolved in this process (load NCCL first)";
    if (code > 11000 && code < 11100) return "NCCL error (code - 11000 = ncclResult_t)";
    return cudaGetErrorString((cudaError_t)code);
}

int wh_num_pickup_points(const wh_config *cfg) { return 4 * cfg->num_racks * cfg->num_racks; }
int wh_num_delivery_points(const wh_config *cfg) { return 4 * (cfg->area_dimension - 4); }

int wh_reset(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
             const int8_t *agent_pos, const int8_t *init_pickups, const int8_t *init_targets,
             const int8_t *num_agents, const uint8_t *env_mask, const wh_obs *obs, void *stream) {
    KParams K; Shape sh;
    if (int rc = fill_params(cfg, K, sh)) return rc;
    if (n_envs == 0) return 0;   // empty batch: nothing to launch
    if (!state_ok(st) || (obs && !obs_ok(obs))) return WH_E_ARG;
    if (agent_pos && (!init_pickups || !init_targets)) return WH_E_ARG;
    set_state(K, st);
    if (obs) K.obs = *obs;
    K.N = n_envs; K.env_id0 = env_id0; K.seed = seed;
    K.r_agent_pos = agent_pos; K.r_init_p = init_pickups; K.r_init_t = init_targets;
    K.r_num_agents = num_agents; K.env_mask = env_mask;
    return launch(K_RESET, K, sh, stream);
}

int wh_step(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
            const int32_t *actions, const int32_t *order,
            const int8_t *spawn_pickups, const int8_t *spawn_targets,
            float *rewards, uint8_t *dones, unsigned long long *stats,
            const wh_obs *obs, int flags, const uint8_t *env_mask, void *stream) {
    KParams K; Shape sh;
    if (int rc = fill_params(cfg, K, sh)) return rc;
    if (n_envs == 0) return 0;   // empty batch: nothing to launch
    if (!state_ok(st) || !actions || !rewards || !dones || (obs && !obs_ok(obs))) return WH_E_ARG;
    K.env_mask = env_mask;
    if ((spawn_pickups == nullptr) != (spawn_targets == nullptr)) return WH_E_ARG;
    if ((flags & WH_FLAG_AUTO_RESET) && spawn_pickups) return WH_E_ARG;  // auto-reset needs the native RNG
    set_state(K, st);
    if (obs) K.obs = *obs;
    K.N = n_envs; K.env_id0 = env_id0; K.seed = seed;
    K.actions = actions; K.order = order; K.spawn_p = spawn_pickups; K.spawn_t = spawn_targets;
    K.rewards = rewards; K.dones = dones; K.stats = stats; K.flags = flags;
    return launch(K_STEP, K, sh, stream);
}

int wh_step_flat(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
                 const int32_t *actions, const int32_t *order, float *rewards, uint8_t *dones,
                 unsigned long long *stats, float *flat_obs, int flags, const uint8_t *env_mask, void *stream) {
    KParams K; Shape sh;
    if (int rc = fill_params(cfg, K, sh)) return rc;
    if (n_envs == 0) return 0;   // empty batch: nothing to launch
    if (!state_ok(st) || !actions || !rewards || !dones || !flat_obs) return WH_E_ARG;
    K.env_mask = env_mask;
    set_state(K, st);
    K.flat_out = flat_obs;
    K.N = n_envs; K.env_id0 = env_id0; K.seed = seed;
    K.actions = actions; K.order = order;
    K.rewards = rewards; K.dones = dones; K.stats = stats; K.flags = flags;
    return launch(K_STEP_FLAT, K, sh, stream);
}

int wh_greedy_step(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0,
                   uint64_t seed, uint64_t solver_seed, uint64_t rand_threshold,
                   int32_t *actions_out, float *rewards, uint8_t *dones,
                   unsigned long long *stats, const wh_obs *obs, int flags, void *stream) {
    KParams K; Shape sh;
    if (int rc = fill_params(cfg, K, sh)) return rc;
    if (n_envs == 0) return 0;   // empty batch: nothing to launch
    if (!state_ok(st) || !rewards || !dones || (obs && !obs_ok(obs))) return WH_E_ARG;
    set_state(K, st);
    if (obs) K.obs = *obs;
    K.N = n_envs; K.env_id0 = env_id0; K.seed = seed; K.solver_seed = solver_seed;
    K.rand_thr = rand_threshold; K.actions_out = actions_out;
    K.rewards = rewards; K.dones = dones; K.stats = stats; K.flags = flags;
    return launch(K_GSTEP, K, sh, stream);
}

int wh_greedy_rollout(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0,
                      uint64_t seed, uint64_t solver_seed, uint64_t rand_threshold, int n_steps,
                      float *reward_sums, uint8_t *dones, unsigned long long *stats, int flags, void *stream) {
    KParams K; Shape sh;
    if (int rc = fill_params(cfg, K, sh)) return rc;
    if (n_envs == 0 || n_steps == 0) return 0;
    if (!state_ok(st) || !reward_sums || !dones || n_steps < 0) return WH_E_ARG;
    set_state(K, st);
    K.N = n_envs; K.env_id0 = env_id0; K.seed = seed; K.solver_seed = solver_seed;
    K.rand_thr = rand_threshold; K.n_steps = n_steps;
    K.rewards = reward_sums; K.dones = dones; K.stats = stats; K.flags = flags;
    return launch(K_ROLLOUT, K, sh, stream);
}

int wh_multi_step(const wh_config *cfg, const wh_state *st, int64_t n_envs, int64_t env_id0, uint64_t seed,
                  int n_steps, const int32_t *actions, uint64_t solver_seed, uint64_t rand_threshold,
                  float *rewards, uint8_t *dones, unsigned long long *stats, const wh_obs *obs, int flags,
                  void *stream) {
    KParams K; Shape sh;
    if (int rc = fill_params(cfg, K, sh)) return rc;
    if (n_envs == 0 || n_steps == 0) return 0;
    if (!state_ok(st) || !rewards || !dones || n_steps < 0 || (obs && !obs_ok(obs))) return WH_E_ARG;
    if (flags & ~(WH_FLAG_AUTO_RESET | WH_FLAG_PER_STEP_OUT | WH_FLAG_MULTI_KERNEL(7))) return WH_E_ARG;
    const int force = (flags >> 4) & 7;
    if (force > 4 || (force >= 3 && (!obs || sh.RC == 0))) return WH_E_ARG;
    set_state(K, st);
    if (obs) K.obs = *obs;
    K.N = n_envs; K.env_id0 = env_id0; K.seed = seed; K.solver_seed = solver_seed;
    K.rand_thr = rand_threshold; K.n_steps = n_steps; K.actions = actions;
    K.rewards = rewards; K.dones = dones; K.stats = stats; K.flags = flags;
    return launch(actions ? K_MULTI : K_GMULTI, K, sh, stream);
}

int wh_build_obs(const wh_config *cfg, const wh_state *st, int64_t n_envs, int flavour,
                 const wh_obs *obs, void *stream) {
    KParams K; Shape sh;
    if (int rc = fill_params(cfg, K, sh)) return rc;
    if (n_envs == 0) return 0;   // empty batch: nothing to launch
    if (!state_ok(st) || !obs_ok(obs) || (flavour != WH_OBS_STEP && flavour != WH_OBS_RESET)) return WH_E_ARG;
    set_state(K, st);
    K.obs = *obs; K.N = n_envs; K.flavour = flavour;
    return launch(K_OBS, K, sh, stream);
}

int wh_build_obs_flat(const wh_config *cfg, const wh_state *st, int64_t n_envs, int flavour,
                      float *out, void *stream) {
    KParams K; Shape sh;
    if (int rc = fill_params(cfg, K, sh)) return rc;
    if (n_envs == 0) return 0;   // empty batch: nothing to launch
    if (!state_ok(st) || !out || (flavour != WH_OBS_STEP && flavour != WH_OBS_RESET)) return WH_E_ARG;
    set_state(K, st);
    K.flat_out = out;
    K.N = n_envs; K.flavour = flavour;
    return launch(K_OBS_FLAT, K, sh, stream);
}

int wh_greedy(const wh_config *cfg, const wh_obs *obs, const int8_t *num_agents,
              const int32_t *episode, const int32_t *time, int64_t n_envs, int64_t env_id0,
              uint64_t seed, uint64_t rand_threshold, const uint8_t *is_random,
              const int32_t *random_actions, int32_t *actions, void *stream) {
    KParams K; Shape sh;
    if (int rc = fill_params(cfg, K, sh)) return rc;
    if (n_envs == 0) return 0;   // empty batch: nothing to launch
    if (!obs || !obs->requests || !obs->self_position || !obs->self_availability ||
        !obs->self_delivery_target || !num_agents || !actions)
        return WH_E_ARG;
    if (is_random && !random_actions) return WH_E_ARG;
    if (!is_random && rand_threshold && (!episode || !time)) return WH_E_ARG;
    K.obs = *obs; K.N = n_envs; K.env_id0 = env_id0; K.solver_seed = seed; K.rand_thr = rand_threshold;
    K.is_random = is_random; K.random_actions = random_actions; K.g_num_agents = num_agents;
    K.g_episode = episode; K.g_time = time; K.actions_out = actions;
    return launch(K_GREEDY, K, sh, stream);
}

int wh_save_prev(const wh_config *cfg, const wh_state *st, const wh_prev *prev, int64_t n_envs, void *stream) {
    KParams K; Shape sh;
    if (int rc = fill_params(cfg, K, sh)) return rc;
    if (n_envs == 0) return 0;
    if (!state_ok(st) || !prev || !prev->agent_pos || !prev->agent_tgt || !prev->pickup_tgt) return WH_E_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    // core.py:270-272: copies taken at the top of step(), used only by render(animate=True)
    CK(cudaMemcpyAsync(prev->agent_pos, st->agent_pos, (size_t)n_envs * K.R * 2, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(prev->agent_tgt, st->agent_tgt, (size_t)n_envs * K.R, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(prev->pickup_tgt, st->pickup_tgt, (size_t)n_envs * K.P, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int wh_stats_allreduce(unsigned long long *stats, void *nccl_comm, void *stream) {
    if (!stats || !nccl_comm) return WH_E_ARG;
    // ncclResult_t ncclAllReduce(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)
    typedef int (*allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    static allreduce_fn fn = nullptr;
    if (!fn) {
        fn = (allreduce_fn)dlsym(RTLD_DEFAULT, "ncclAllReduce");
        if (!fn) {
            void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (h) fn = (allreduce_fn)dlsym(h, "ncclAllReduce");
        }
        if (!fn) return WH_E_NCCL;
    }
    const int ncclUint64 = 5, ncclSum = 0;   // nccl.h enums (stable since NCCL 2.0)
    const int rc = fn(stats, stats, WH_NUM_STATS, ncclUint64, ncclSum, nccl_comm, (cudaStream_t)stream);
    return rc == 0 ? 0 : 11000 + rc;
}

// ---------------------------------------------------------------------------------------------
// C ABI — layer 2: host-buffer environment handle
// ---------------------------------------------------------------------------------------------
void wh_env_destroy(wh_env *E);

struct wh_env {
    wh_config cfg;
    int64_t N, env_id0;
    uint64_t seed;
    int device, n_chunks, n_streams, R, P;
    int direct;      // 0 = copy pipeline; 1 = the kernel reads the actions from / writes the rewards to the host buffers itself
    wh_state st;
    wh_obs obs;
    int32_t *d_actions;
    float *d_rewards;
    uint8_t *d_dones;
    unsigned long long *d_stats;
    cudaStream_t *streams;
    int64_t launches;
};

static wh_state offset_state(const wh_env *E, int64_t e0) {
    wh_state s = E->st;
    s.agent_pos += e0 * E->R * 2; s.agent_tgt += e0 * E->R; s.pickup_tgt += e0 * E->P;
    s.pickup_timer += e0 * E->P; s.time += e0; s.num_agents += e0; s.episode += e0; s.acc += e0 * 4;
    return s;
}

static wh_obs offset_obs(const wh_obs &b, int64_t e0, int64_t R) {
    wh_obs o = b;
    o.num_agents += e0 * R; o.self_position += e0 * R * 2; o.self_availability += e0 * R;
    o.self_delivery_target += e0 * R * 2; o.other_positions += e0 * R * (R - 1) * 2;
    o.other_availabilities += e0 * R * (R - 1); o.other_delivery_targets += e0 * R * (R - 1) * 2;
    o.requests += e0 * R * R * 4;
    return o;
}

static int env_alloc(wh_env *E, const KParams &K, int n_chunks) {
    const int64_t N = E->N, R = K.R, P = K.P;
    CK(cudaMalloc(&E->st.agent_pos, N * R * 2)); CK(cudaMalloc(&E->st.agent_tgt, N * R));
    CK(cudaMalloc(&E->st.pickup_tgt, N * P)); CK(cudaMalloc(&E->st.pickup_timer, N * P * 2));
    CK(cudaMalloc(&E->st.time, N * 4)); CK(cudaMalloc(&E->st.num_agents, N));
    CK(cudaMalloc(&E->st.episode, N * 4)); CK(cudaMalloc(&E->st.acc, N * 16));
    CK(cudaMemset(E->st.agent_pos, 0xFF, N * R * 2)); CK(cudaMemset(E->st.agent_tgt, 0xFF, N * R));
    CK(cudaMemset(E->st.pickup_tgt, 0xFF, N * P)); CK(cudaMemset(E->st.pickup_timer, 0xFF, N * P * 2));
    CK(cudaMemset(E->st.time, 0, N * 4)); CK(cudaMemset(E->st.num_agents, K.max_agents, N));
    CK(cudaMemset(E->st.episode, 0xFF, N * 4)); CK(cudaMemset(E->st.acc, 0, N * 16));
    CK(cudaMalloc(&E->obs.num_agents, N * R * 4)); CK(cudaMalloc(&E->obs.self_position, N * R * 8));
    CK(cudaMalloc(&E->obs.self_availability, N * R)); CK(cudaMalloc(&E->obs.self_delivery_target, N * R * 8));
    CK(cudaMalloc(&E->obs.other_positions, N * R * (R - 1) * 8));
    CK(cudaMalloc(&E->obs.other_availabilities, N * R * (R - 1)));
    CK(cudaMalloc(&E->obs.other_delivery_targets, N * R * (R - 1) * 8));
    CK(cudaMalloc(&E->obs.requests, N * R * R * 16));
    CK(cudaMalloc(&E->d_actions, N * R * 4)); CK(cudaMalloc(&E->d_rewards, N * R * 4));
    CK(cudaMalloc(&E->d_dones, N)); CK(cudaMalloc(&E->d_stats, WH_NUM_STATS * 8));
    CK(cudaMemset(E->d_stats, 0, WH_NUM_STATS * 8));
    E->n_streams = n_chunks < 1 ? 1 : n_chunks;
    E->streams = (cudaStream_t *)calloc((size_t)E->n_streams, sizeof(cudaStream_t));
    if (!E->streams) return WH_E_ARG;
    for (int i = 0; i < E->n_streams; ++i) CK(cudaStreamCreateWithFlags(&E->streams[i], cudaStreamNonBlocking));
    CK(cudaDeviceSynchronize());
    return 0;
}

int wh_env_create(const wh_config *cfg, int64_t n_envs, int device, int64_t env_id0, uint64_t seed,
                  int n_chunks, wh_env **out) {
    if (!cfg || !out || n_envs <= 0) return WH_E_ARG;
    *out = nullptr;
    KParams K; Shape sh;
    if (int rc = fill_params(cfg, K, sh)) return rc;
    CK(cudaSetDevice(device));
    wh_env *E = new (std::nothrow) wh_env();
    if (!E) return WH_E_ARG;
    memset(E, 0, sizeof(*E));
    E->cfg = *cfg; E->N = n_envs; E->env_id0 = env_id0; E->seed = seed; E->device = device;
    E->R = K.R; E->P = K.P;
    // n_chunks <= 0 selects the direct (zero-copy) mode, see wh_env_step_host
    E->direct = n_chunks <= 0 ? 1 : 0;
    if (n_chunks <= 0) n_chunks = 1;
    if (n_chunks > 64) n_chunks = 64;
    E->n_chunks = n_chunks;
    if (int rc = env_alloc(E, K, n_chunks)) {   // a failed allocation leaves nothing behind
        wh_env_destroy(E);
        return rc;
    }
    *out = E;
    return 0;
}

void wh_env_destroy(wh_env *E) {
    if (!E) return;
    cudaSetDevice(E->device);
    cudaDeviceSynchronize();
    void *ptrs[] = {E->st.agent_pos, E->st.agent_tgt, E->st.pickup_tgt, E->st.pickup_timer, E->st.time,
                    E->st.num_agents, E->st.episode, E->st.acc, E->obs.num_agents, E->obs.self_position,
                    E->obs.self_availability, E->obs.self_delivery_target, E->obs.other_positions,
                    E->obs.other_availabilities, E->obs.other_delivery_targets, E->obs.requests,
                    E->d_actions, E->d_rewards, E->d_dones, E->d_stats};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (E->streams) {
        for (int i = 0; i < E->n_streams; ++i) if (E->streams[i]) cudaStreamDestroy(E->streams[i]);
        free(E->streams);
    }
    delete E;
}

int wh_env_reset(wh_env *E) {
    if (!E) return WH_E_ARG;
    CK(cudaSetDevice(E->device));
    int rc = wh_reset(&E->cfg, &E->st, E->N, E->env_id0, E->seed, nullptr, nullptr, nullptr, nullptr,
                      nullptr, &E->obs, E->streams[0]);
    E->launches += 1;
    if (rc) return rc;
    CK(cudaStreamSynchronize(E->streams[0]));
    return 0;
}

// Is `p` page-locked host memory this device can address (cudaHostAlloc / cudaHostRegister / torch pin_memory)?
static bool device_can_address(const void *p, void **dev_ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
    *dev_ptr = a.devicePointer;
    return true;
}

// all eight observation tensors of n envs, device -> host, on stream s
static int obs_to_host(const wh_obs &ob, const wh_obs &oh, int64_t n, int64_t R, cudaStream_t s) {
    CK(cudaMemcpyAsync(oh.num_agents, ob.num_agents, n * R * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(oh.self_position, ob.self_position, n * R * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(oh.self_availability, ob.self_availability, n * R, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(oh.self_delivery_target, ob.self_delivery_target, n * R * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(oh.other_positions, ob.other_positions, n * R * (R - 1) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(oh.other_availabilities, ob.other_availabilities, n * R * (R - 1), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(oh.other_delivery_targets, ob.other_delivery_targets, n * R * (R - 1) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(oh.requests, ob.requests, n * R * R * 16, cudaMemcpyDeviceToHost, s));
    return 0;
}

static int env_step_issue(wh_env *E, const int32_t *actions, float *rewards, uint8_t *dones,
                         const wh_obs *obs_host, bool greedy, bool compact) {
    CK(cudaSetDevice(E->device));
    const int64_t R = E->R;
    // Direct mode: the step kernel itself moves the step's I/O over PCIe — it reads the actions from and writes
    // the rewards into the caller's page-locked buffers, so there is no separate copy to wait for before / after
    // the kernel. Falls back to the copy pipeline when a buffer is not page-locked.
    void *d_act = nullptr, *d_rew = nullptr;
    if (E->direct && device_can_address(rewards, &d_rew) && (greedy || device_can_address(actions, &d_act))) {
        cudaStream_t s = E->streams[0];
        int rc;
        const int fl = WH_FLAG_AUTO_RESET | (compact ? WH_FLAG_COMPACT_IO : 0);
        // dones (one byte per env) stay a device-side tensor + one small copy: single-byte stores over PCIe
        // are the one thing the direct mode must not do (262 144 one-byte write transactions per step)
        if (greedy)
            rc = wh_greedy_step(&E->cfg, &E->st, E->N, E->env_id0, E->seed, E->seed ^ 0x5EEDull, 0, nullptr,
                                (float *)d_rew, E->d_dones, E->d_stats, &E->obs, fl, s);
        else
            rc = wh_step(&E->cfg, &E->st, E->N, E->env_id0, E->seed, (const int32_t *)d_act, nullptr, nullptr, nullptr,
                         (float *)d_rew, E->d_dones, E->d_stats, &E->obs, fl, nullptr, s);
        E->launches += 1;
        if (rc) return rc;
        CK(cudaMemcpyAsync(dones, E->d_dones, (size_t)E->N, cudaMemcpyDeviceToHost, s));
        if (obs_host) return obs_to_host(E->obs, *obs_host, E->N, R, s);
        return 0;
    }
    for (int c = 0; c < E->n_chunks; ++c) {
        const int64_t e0 = E->N * c / E->n_chunks, e1 = E->N * (c + 1) / E->n_chunks, n = e1 - e0;
        if (n <= 0) continue;
        cudaStream_t s = E->streams[c];
        const wh_state st = offset_state(E, e0);
        const wh_obs ob = offset_obs(E->obs, e0, R);
        int rc;
        if (greedy) {
            rc = wh_greedy_step(&E->cfg, &st, n, E->env_id0 + e0, E->seed, E->seed ^ 0x5EEDull, 0, nullptr,
                                E->d_rewards + e0 * R, E->d_dones + e0, E->d_stats, &ob,
                                WH_FLAG_AUTO_RESET | WH_FLAG_NO_PDL, s);
        } else {
            if (compact) {
                int8_t *da = reinterpret_cast<int8_t *>(E->d_actions) + e0 * R;
                CK(cudaMemcpyAsync(da, reinterpret_cast<const int8_t *>(actions) + e0 * R, n * R, cudaMemcpyHostToDevice, s));
                rc = wh_step(&E->cfg, &st, n, E->env_id0 + e0, E->seed, reinterpret_cast<const int32_t *>(da), nullptr,
                             nullptr, nullptr, reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(E->d_rewards) + e0 * R),
                             E->d_dones + e0, E->d_stats, &ob, WH_FLAG_AUTO_RESET | WH_FLAG_COMPACT_IO | WH_FLAG_NO_PDL, nullptr, s);
            } else {
                CK(cudaMemcpyAsync(E->d_actions + e0 * R, actions + e0 * R, n * R * 4, cudaMemcpyHostToDevice, s));
                rc = wh_step(&E->cfg, &st, n, E->env_id0 + e0, E->seed, E->d_actions + e0 * R, nullptr, nullptr,
                             nullptr, E->d_rewards + e0 * R, E->d_dones + e0, E->d_stats, &ob,
                             WH_FLAG_AUTO_RESET | WH_FLAG_NO_PDL, nullptr, s);
            }
        }
        E->launches += 1;
        if (rc) return rc;
        if (compact)
            CK(cudaMemcpyAsync(reinterpret_cast<uint8_t *>(rewards) + e0 * R, reinterpret_cast<uint8_t *>(E->d_rewards) + e0 * R,
                               n * R, cudaMemcpyDeviceToHost, s));
        else
            CK(cudaMemcpyAsync(rewards + e0 * R, E->d_rewards + e0 * R, n * R * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(dones + e0, E->d_dones + e0, n, cudaMemcpyDeviceToHost, s));
        if (obs_host)
            if (int rc2 = obs_to_host(ob, offset_obs(*obs_host, e0, R), n, R, s)) return rc2;
    }
    return 0;
}

// Issues the per-chunk copy -> kernel -> copy pipelines, then ALWAYS drains every chunk stream — also when
// issuing failed half-way — so that no copy into the caller's host buffers is still in flight on return.
static int env_step_impl(wh_env *E, const int32_t *actions, float *rewards, uint8_t *dones,
                         const wh_obs *obs_host, bool greedy, bool compact = false) {
    if (!E || !rewards || !dones || (!greedy && !actions)) return WH_E_ARG;
    int rc = env_step_issue(E, actions, rewards, dones, obs_host, greedy, compact);
    for (int c = 0; c < E->n_streams; ++c) {
        const cudaError_t e = cudaStreamSynchronize(E->streams[c]);
        if (!rc && e != cudaSuccess) rc = (int)e;
    }
    return rc;
}

int wh_env_step_host(wh_env *E, const int32_t *actions, float *rewards, uint8_t *dones,
                     const wh_obs *obs_host) {
    return env_step_impl(E, actions, rewards, dones, obs_host, false);
}

int wh_env_step_host_compact(wh_env *E, const int8_t *actions, uint8_t *rewards, uint8_t *dones) {
    return env_step_impl(E, reinterpret_cast<const int32_t *>(actions), reinterpret_cast<float *>(rewards), dones,
                         nullptr, false, true);
}

int wh_env_greedy_step_host(wh_env *E, float *rewards, uint8_t *dones) {
    return env_step_impl(E, nullptr, rewards, dones, nullptr, true);
}

int wh_env_obs_ptrs(wh_env *E, wh_obs *out) { if (!E || !out) return WH_E_ARG; *out = E->obs; return 0; }
int wh_env_state_ptrs(wh_env *E, wh_state *out) { if (!E || !out) return WH_E_ARG; *out = E->st; return 0; }

int wh_env_stats_host(wh_env *E, unsigned long long *stats_out) {
    if (!E || !stats_out) return WH_E_ARG;
    CK(cudaSetDevice(E->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(stats_out, E->d_stats, WH_NUM_STATS * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int64_t wh_env_launch_count(wh_env *E) { return E ? E->launches : -1; }

}  // extern "C"
